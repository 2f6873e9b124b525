"""Row-partitioned leg of bench.py (N > 1): one process per GPU, launched by torch.distributed.run.

STRONG scaling: the same S3-hex-n system is split into N contiguous block-row ranges balanced by
stored blocks; each rank generates its own rows directly in HBM.  The p-halo is pushed straight into
the neighbours' memory over NVLink (cudaIpc-mapped peer stores + device-side flags), overlapped with
the interior SpMV; the two dot products per iteration are combined through peer mailboxes, every rank
adding the partial sums in rank order (xfem-amie_b200/csrc/dist.cu; AMIE_B200_TRANSPORT=nccl selects
the NCCL send/recv + allreduce path).  Time = max over ranks of the CUDA-event solve time.
The line carries `x_checksum` (sum of |x| over all ranks) and the iteration count, to be compared with
the N = 1 line of bench.py: the partitioned solve must reproduce the single-device answer.
"""
import ctypes
import json
import time

import numpy as np


# sum |x| of the single-device solve (bench.py, N = 1, profiles/): the partitioned solve must land on the same field
X_CHECKSUM_1GPU = {("S3-hex", 256): 6789467.724700802}


def run_distributed(args, pkg, dist, rank, world, local_rank):
    import torch
    from bench import METRIC, UNIT, ClockSampler, measured_peak, stdout_to_stderr

    dev = torch.device("cuda", local_rank)
    syn = pkg.Synth(args.preset, args.n)
    rs, _ = syn.row_sizes()
    bounds = pkg.partition_rows(rs, world)
    del rs

    with stdout_to_stderr():
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        id128 = bytes(idt.cpu().numpy().tobytes())
        asm = pkg.Assembly(device=local_rank)
        asm.dist_init(rank, world, id128, bounds)
    t0 = time.time()
    asm.dist_synth_to_device(syn)
    gen_s = time.time() - t0
    asm.set_option("time_spmv", 1)
    st = asm.stats()
    info = asm.dist_info()
    N_loc, s = st.ndof, st.stride
    N_glob = syn.nb * s

    def maxf(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumf(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def step():
        asm.upload_x0(None)
        ok, nit, err, rho = asm.pcg_resident(nssor=32)
        return ok, nit, asm.stats()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall = time.time()
    dev_ms = spmv_ms = 0.0
    its = spmv_n = launches = 0
    conv = True
    for _ in range(args.steps):
        ok, nit, s_ = step()
        conv &= ok
        its += nit
        dev_ms += s_.solve_ms
        spmv_ms += s_.spmv_ms_total
        spmv_n += s_.spmv_timed
        launches += s_.kernel_launches
    torch.cuda.synchronize()
    dist.barrier()
    wall_ms = 1e3 * (time.time() - t_wall)
    clocks = sampler.stop()

    x_checksum = sumf(float(np.abs(asm.download_x()).sum()))
    dev_ms_max = maxf(dev_ms)
    spmv_avg_ms = maxf(spmv_ms / max(1, spmv_n))
    algo_bytes_total = sumf(st.spmv_algorithmic_bytes)
    launches_total = int(sumf(launches))
    value = its / (dev_ms_max * 1e-3)
    peak, peak_src = measured_peak()
    achieved_per_gpu = st.spmv_algorithmic_bytes / (spmv_avg_ms * 1e-3) / 1e9

    # e2e: host (pinned) local slices through the C-ABI
    e2e = None
    if not args.no_e2e:
        b_host = torch.empty(N_loc, dtype=torch.float64, pin_memory=True).numpy()
        x_host = torch.empty(N_loc, dtype=torch.float64, pin_memory=True).numpy()
        x0_host = torch.zeros(N_loc, dtype=torch.float64, pin_memory=True).numpy()
        b_host[:] = asm.download_rhs()
        L = pkg.lib()
        nit_c, err_c, rho_c = ctypes.c_uint64(), ctypes.c_double(), ctypes.c_double()

        def e2e_step():
            rc = L.amie_b200_pcg(asm.ctx, b_host.ctypes.data, x0_host.ctypes.data, N_loc, 0, 1e-10, -1, 32, 0, 0,
                                 x_host.ctypes.data, ctypes.byref(nit_c), ctypes.byref(err_c), ctypes.byref(rho_c))
            asm.check(rc)
            return nit_c.value
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.time()
        e_its = 0
        for _ in range(args.steps):
            e_its += e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        e_wall = maxf(time.time() - t0)
        e2e = {"value": e_its / e_wall, "unit": UNIT, "h2d_bytes_per_step": int(2 * N_glob * 8), "d2h_bytes_per_step": int(N_glob * 8)}

    halo_max = maxf(info["halo"])
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.preset}-{args.n}", "ndof": int(N_glob), "stride": int(s), "eps": 1e-10, "nssor": 32,
                           "maxit": -1, "precond": "InverseDiagonal", "parallelism": f"row-partition x{world}; halo + 2-double reductions over " + ("NVLink peer memory (device-side flags, no NCCL call per iteration)" if info["transport"] == "peer" else "NCCL send/recv + allreduce"),
                           "step": "one full PCG solve (reference control flow)", "iterations_per_step": its / max(1, args.steps),
                           "halo_block_columns_max": int(halo_max), "l2": "per-rank matrix is far larger than L2; no flush needed",
                           "generate_s": gen_s},
                "converged": bool(conv), "wall_ms_per_step": wall_ms / max(1, args.steps), "dof_iter_per_s": value * N_glob,
                "x_checksum": x_checksum, "nit": int(its // max(1, args.steps)),
                "x_checksum_rel_to_1gpu": (abs(x_checksum / X_CHECKSUM_1GPU[(args.preset, args.n)] - 1.0)
                                           if (args.preset, args.n) in X_CHECKSUM_1GPU else None),
                "clocks": clocks, "gpu_launches": launches_total,
                "roofline": {"bound": "hbm", "achieved": achieved_per_gpu, "peak": peak, "unit": "GB/s", "frac": achieved_per_gpu / peak,
                             "traffic": None, "kernel": "k_spmv_s3_rt (per GPU, interior + boundary launches incl. halo wait)",
                             "algorithmic_bytes_per_launch": int(st.spmv_algorithmic_bytes), "launch_ms": spmv_avg_ms,
                             "aggregate_GBs": algo_bytes_total / (spmv_avg_ms * 1e-3) / 1e9, "peak_source": peak_src},
                "e2e": e2e, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    asm.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0
