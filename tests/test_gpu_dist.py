"""Row-partitioned CUDA path (dist.cu) against the oracle.  world 1 runs anywhere there is a GPU
(it still goes through the deferred all-reduce + scalar-step kernels); world 2 needs two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _solve_rank(pkg, syn, rank, world, id128, bounds, device, bicg=False, x0=None, id128_b=None):
    s = syn.stride
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    rs, ci, arr, b = syn.rows(r0, r1)
    asm = pkg.Assembly(device=device)
    asm.dist_init(rank, world, id128, bounds)
    asm.dist_set_structure(s, syn.nb, rs, ci)
    asm.set_values(arr)
    asm.upload_rhs(b)
    asm.upload_x0(None if x0 is None else x0[r0 * s:r1 * s])
    if bicg:
        ok, nit, err = asm.bicgstab_resident()
    else:
        ok, nit, err, rho = asm.pcg_resident(nssor=32)
    x = asm.download_x()
    info = asm.dist_info()
    asm.close()
    if id128_b is None:
        return ok, nit, x, info, None
    # the device generator must give the same partitioned system (a ncclUniqueId serves ONE communicator)
    asm2 = pkg.Assembly(device=device)
    asm2.dist_init(rank, world, id128_b, bounds)
    asm2.dist_synth_to_device(syn)
    asm2.upload_x0(None)
    ok2, nit2, _, _ = asm2.pcg_resident(nssor=32)
    x2 = asm2.download_x()
    asm2.close()
    return ok, nit, x, info, (ok2, nit2, x2)


@pytest.mark.parametrize("preset,n", [("S3-hex", 12), ("S2-tri", 30)])
def test_dist_world1(pkg, ol, systems, preset, n):
    syn = pkg.Synth(preset, n)
    S = systems(preset, n)
    bounds = np.array([0, syn.nb], np.uint64)
    ok, nit, x, info, (ok2, nit2, x2) = _solve_rank(pkg, syn, 0, 1, pkg.nccl_unique_id(), bounds, 0, id128_b=pkg.nccl_unique_id())
    ret, x_ref, oinfo = ol.oracle_cg(S, nssor=32)
    assert ok == bool(ret) and abs(int(nit) - int(oinfo.nit)) <= 2
    assert rel_l2(x, x_ref) <= 1e-8
    assert info["halo"] == 0 and info["peers"] == 0 and info["interior_rows"] == syn.nb
    assert ok2 and nit2 == nit and np.array_equal(x, x2)
    ok, nit, x, _, _ = _solve_rank(pkg, syn, 0, 1, pkg.nccl_unique_id(), bounds, 0, bicg=True)
    ret, x_ref, binfo = ol.oracle_bicgstab(S)
    assert ok == bool(ret) and rel_l2(x, x_ref) <= 1e-8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker2(rank, world, port, preset, n, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    syn = pkg.Synth(preset, n)
    rs, _ = syn.row_sizes()
    bounds = pkg.partition_rows(rs, world)
    def fresh_id():
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return idt.numpy().tobytes()
    ok, nit, x, info, (ok2, nit2, x2) = _solve_rank(pkg, syn, rank, world, fresh_id(), bounds, rank, id128_b=fresh_id())
    okb, nitb, xb, _, _ = _solve_rank(pkg, syn, rank, world, fresh_id(), bounds, rank, bicg=True)
    out[rank] = (ok, nit, x, info, ok2, nit2, x2, okb, nitb, xb, int(bounds[rank]), int(bounds[rank + 1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("preset,n", [("S3-hex", 14), ("S3-tet", 12), ("S2-tri", 40)])
def test_dist_world2(pkg, ol, systems, preset, n):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    S = systems(preset, n)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker2, args=(2, _free_port(), preset, n, out), nprocs=2, join=True)
    ret, x_ref, oinfo = ol.oracle_cg(S, nssor=32)
    x = np.zeros(S.n)
    xb = np.zeros(S.n)
    for rank in range(2):
        ok, nit, xl, info, ok2, nit2, x2, okb, nitb, xbl, r0, r1 = out[rank]
        assert ok == bool(ret) and abs(int(nit) - int(oinfo.nit)) <= 2
        assert ok2 and abs(int(nit2) - int(nit)) <= 0 and np.array_equal(xl, x2)
        assert info["peers"] == 1 and info["halo"] > 0 and info["send"] > 0
        assert info["transport"] == ("nccl" if os.environ.get("AMIE_B200_TRANSPORT") == "nccl" else "peer")
        x[r0 * S.stride:r1 * S.stride] = xl
        xb[r0 * S.stride:r1 * S.stride] = xbl
        assert okb
    assert rel_l2(x, x_ref) <= 1e-8
    _, xb_ref, _ = ol.oracle_bicgstab(S)
    assert rel_l2(xb, xb_ref) <= 1e-8
