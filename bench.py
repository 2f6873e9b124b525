#!/usr/bin/env python3
"""bench.py -- PCG iterations/s and SpMV HBM GB/s on the synthetic 3D elastic mesh (BASELINE.json).

A "step" is ONE full reference-flow PCG solve (Jacobi, eps 1e-10, nssor 32, maxit -1, x0 = 0) of the
S3-hex-n system (n = 256 -> 50.3 M DOF).  Per step the solver runs its ~2.5 k
iterations, so `value` = iterations of all timed steps / device time of those steps.

  value    : device-resident solves (matrix, b, x0 already in HBM; CUDA-event time inside the library)
  e2e      : the same solve through the reference-facing C-ABI call amie_b200_pcg with HOST (pinned)
             b / x0 / x buffers: H2D and D2H inside the timed region
  roofline : block-row SpMV, algorithmic bytes nnzb*(8 s^2+4) + 4 (nb+1) + 16 N per launch over the
             mean CUDA-event duration of the SpMV launches of the timed steps
  cpu_baseline : the reference's own CPU solver (oracle/_ref, OpenMP, all host cores) on a bounded sample

`--impl reference` times only that CPU arm and prints its own line.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# experiment switch read by the library at context creation (include/amie_b200.h, option "split_dot")
SPLIT_DOT = os.environ.get("AMIE_B200_SPLIT_DOT", "0") not in ("", "0")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PCG iters/s & SpMV HBM GB/s (% peak), 3D elastic mesh, 1/2/4/8 B200 vs CPU"
UNIT = "PCG iterations/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class stdout_to_stderr:
    """NCCL prints its version banner on stdout at communicator creation: keep stdout for the ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nme in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm

def cpu_reference_rate(pkg, n_cpu, threads=0):
    """The reference's own ConjugateGradient::solve (oracle/_ref, OpenMP) on S3-hex-n_cpu.
    Returns dict(dof_iter_per_s, it_per_s, nit, wall, cores, kind, n)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    syn = pkg.Synth("S3-hex", n_cpu)
    rs, ci, arr, b = syn.rows()
    S = ol.Sys(syn.stride, syn.nb, rs, ci, arr, b)
    if ol.ref() is not None:
        # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on
        cores = threads or len(os.sched_getaffinity(0)) or ol.ref_max_threads()
        ok, x, nit, wall, _ = ol.ref_cg(S, nssor=32, nthreads=cores)
        kind = "reference"
        # the reference's own SpMV, assign(y, A*x), 20 launches: GB/s by the algorithmic bytes
        # every roofline here uses, and by the bytes of the reference's padded layout
        _, spmv_s = ol.ref_spmv(S, b, mode=0, nthreads=cores, reps=20)
        st, cl = S.stride, S.stride + S.stride % 2
        spmv = {"ms": 1e3 * spmv_s,
                "gbs_algorithmic": (S.nnzb * (8 * st * st + 4) + 4 * (S.nb + 1) + 16 * S.n) / spmv_s / 1e9,
                "gbs_padded_layout": (S.nnzb * (8 * st * cl + 4) + 8 * S.nb + 16 * S.n) / spmv_s / 1e9}
    else:
        cores = 1
        t0 = time.time()
        ok, x, info = ol.oracle_cg(S, nssor=32)
        wall, nit = time.time() - t0, info.nit
        kind = "port"
        spmv = None
    return dict(dof_iter_per_s=S.n * nit / wall, it_per_s=nit / wall, nit=int(nit), wall=wall, cores=int(cores),
                kind=kind, n=n_cpu, ndof=S.n, converged=bool(ok), spmv=spmv)


def run_reference_arm(args, pkg, rank):
    if rank != 0:
        return
    N_work = 3 * args.n ** 3
    rates = []
    for _ in range(args.warmup):
        cpu_reference_rate(pkg, max(8, args.cpu_n // 2))
    t_all = time.time()
    for _ in range(args.steps):
        rates.append(cpu_reference_rate(pkg, args.cpu_n))
    wall = time.time() - t_all
    r = rates[-1]
    dof_it = float(np.mean([q["dof_iter_per_s"] for q in rates]))
    value = dof_it / N_work
    sample = (f"each step = one full ConjugateGradient::solve of S3-hex-{r['n']} ({r['ndof']} DOF, {r['nit']} it) by the "
              f"{'compiled reference (oracle/_ref, OpenMP)' if r['kind'] == 'reference' else 'C oracle port'}; DOF*iter/s scaled by the DOF ratio "
              f"to the S3-hex-{args.n} workload ({N_work} DOF)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"S3-hex-{args.n}", "ndof": N_work, "eps": 1e-10, "nssor": 32, "maxit": -1},
            "dof_iter_per_s": dof_it,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample, "spmv": r["spmv"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh-n", dest="n", type=int, default=int(os.environ.get("AMIE_BENCH_N", 256)), help="nodes per side of S3-hex-n (256 -> 50.3 M DOF)")
    ap.add_argument("--preset", default="S3-hex")
    ap.add_argument("--cpu-n", type=int, default=int(os.environ.get("AMIE_BENCH_CPU_N", 64)), help="size of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--spmv-variant", type=int, default=0, help="kernel selection for A/B runs (option spmv_variant; 0 = the shipped default)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    import __graft_entry__ as g
    pkg = g.load_package()

    if args.impl == "reference":
        run_reference_arm(args, pkg, rank)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        from bench_dist import run_distributed          # row-partitioned path
        return run_distributed(args, pkg, dist, rank, world, local_rank)

    # ---- build the system directly in HBM
    syn = pkg.Synth(args.preset, args.n)
    asm = pkg.Assembly(device=local_rank)
    t0 = time.time()
    syn.to_device(asm)
    gen_s = time.time() - t0
    asm.set_option("time_spmv", 1)
    if args.spmv_variant:
        asm.set_option("spmv_variant", args.spmv_variant)
    st = asm.stats()
    N, nb, nnzb, s = st.ndof, st.nb, st.nnzb, st.stride
    algo_bytes = st.spmv_algorithmic_bytes

    def resident_step():
        asm.upload_x0(None)
        ok, nit, err, rho = asm.pcg_resident(nssor=32)
        return ok, nit, asm.stats()

    for _ in range(args.warmup):
        resident_step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    t_wall = time.time()
    dev_ms = spmv_ms = 0.0
    its = spmv_n = launches = smoothing = 0
    conv = True
    for _ in range(args.steps):
        ok, nit, s_ = resident_step()
        conv &= ok
        its += nit
        dev_ms += s_.solve_ms
        spmv_ms += s_.spmv_ms_total
        spmv_n += s_.spmv_timed
        launches += s_.kernel_launches
        smoothing += s_.smoothing_spmv
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.time() - t_wall)
    clocks = sampler.stop()

    value = its / (dev_ms * 1e-3)
    spmv_avg_ms = spmv_ms / max(1, spmv_n)
    peak, peak_src = measured_peak()
    achieved = algo_bytes / (spmv_avg_ms * 1e-3) / 1e9
    iter_bytes = algo_bytes + 120 * N           # + K-Update (7R+4W) + K-Dir (3R+1W), Kahan compensators kept

    # ---- e2e: host (pinned) buffers through the C-ABI call
    e2e = None
    if not args.no_e2e:
        b_host = torch.empty(N, dtype=torch.float64, pin_memory=True).numpy()
        x_host = torch.empty(N, dtype=torch.float64, pin_memory=True).numpy()
        x0_host = torch.zeros(N, dtype=torch.float64, pin_memory=True).numpy()
        b_host[:] = asm.download_rhs()
        L = pkg.lib()
        nit_c, err_c, rho_c = ctypes.c_uint64(), ctypes.c_double(), ctypes.c_double()

        def e2e_step():
            rc = L.amie_b200_pcg(asm.ctx, b_host.ctypes.data, x0_host.ctypes.data, N, 0, 1e-10, -1, 32, 0, 0,
                                 x_host.ctypes.data, ctypes.byref(nit_c), ctypes.byref(err_c), ctypes.byref(rho_c))
            asm.check(rc)
            return nit_c.value
        torch.cuda.synchronize()         # (already warm: W + K resident solves ran on this context)
        t0 = time.time()
        e_its = 0
        for _ in range(args.steps):
            e_its += e2e_step()
        torch.cuda.synchronize()
        e_wall = time.time() - t0
        e2e = {"value": e_its / e_wall, "unit": UNIT, "h2d_bytes_per_step": int(2 * N * 8), "d2h_bytes_per_step": int(N * 8),
               "x_checksum": float(np.abs(x_host).sum())}

    # ---- CPU baseline (rank 0, bounded sample)
    cpu = None
    if not args.no_cpu:
        r = cpu_reference_rate(pkg, args.cpu_n)
        cpu = {"value": r["dof_iter_per_s"] / N, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
               "dof_iter_per_s": r["dof_iter_per_s"], "spmv": r["spmv"],
               "sample": f"one full ConjugateGradient::solve of S3-hex-{r['n']} ({r['ndof']} DOF, {r['nit']} it, {r['wall']:.1f} s); "
                         f"DOF*iter/s scaled by the DOF ratio to this workload ({N} DOF)"}

    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{args.preset}-{args.n}/1", {}).get("bytes")
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.preset}-{args.n}", "ndof": int(N), "block_rows": int(nb), "blocks": int(nnzb), "stride": int(s),
                       "eps": 1e-10, "nssor": 32, "maxit": -1, "precond": "InverseDiagonal",
                       "step": "one full PCG solve (reference control flow)", "iterations_per_step": its / max(1, args.steps),
                       "l2": "matrix (%.1f GB) is far larger than L2; no flush needed" % (nnzb * (8 * s * s + 4) / 1e9),
                       "generate_s": gen_s},
            "converged": bool(conv), "wall_ms_per_step": wall_ms / max(1, args.steps),
            "dof_iter_per_s": value * N, "smoothing_spmv_per_step": smoothing / max(1, args.steps),
            "pcg_iteration_gbs": iter_bytes * value / 1e9,
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": (("k_spmv_s3_rt<DOT_NONE> (q = A p; p.q as a separate pass: AMIE_B200_SPLIT_DOT)" if SPLIT_DOT else
                                                          "k_spmv_s3_rt<DOT_YX> (q = A p fused with p.q; TMA bulk-copy pipeline)") if s == 3 else "k_spmv_s2_rt<DOT_YX> (row-thread TMA pipeline, 2x2 blocks)"),
                         "algorithmic_bytes_per_launch": int(algo_bytes), "launch_ms": spmv_avg_ms, "launches_timed": int(spmv_n),
                         "peak_source": peak_src},
            "e2e": e2e, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    asm.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
