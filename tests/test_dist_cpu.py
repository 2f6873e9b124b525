"""Host-side logic of the row-partitioned path, exercised with world_size 2 over gloo on the CPU:
partition bounds, halo lists, the owned|halo renumbering of dist.cu restated in numpy, the halo
exchange and the all-reduced dot product -- against the oracle on the undivided system."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, preset, n, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    import oracle_lib as ol
    pkg = g.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    syn = pkg.Synth(preset, n)
    s = syn.stride
    rs_all, _ = syn.row_sizes()
    bounds = pkg.partition_rows(rs_all, world).astype(np.int64)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    rs, ci, arr, b = syn.rows(r0, r1)
    halo = pkg.partition_halo(r0, r1, rs, ci).astype(np.int64)
    nbl = r1 - r0
    # renumber like k_remap_cols: owned -> c - r0, halo -> nbl + position in the sorted halo list
    ci64 = ci.astype(np.int64)
    own = (ci64 >= r0) & (ci64 < r1)
    loc = np.where(own, ci64 - r0, nbl + np.searchsorted(halo, ci64))
    # what each rank needs from each owner, exchanged with all_to_all (the NCCL send/recv of dist.cu)
    owner = np.searchsorted(bounds, halo, side="right") - 1
    need = [torch.from_numpy(halo[owner == q].copy()) for q in range(world)]
    counts = torch.tensor([t.numel() for t in need])
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, counts)
    asked = []
    for q in range(world):          # pairwise, ordered to avoid deadlock
        if q == rank:
            asked.append(torch.zeros(0, dtype=torch.int64))
            continue
        recv = torch.zeros(int(allc[q][rank]), dtype=torch.int64)
        if rank < q:
            dist.send(need[q], q)
            dist.recv(recv, q)
        else:
            dist.recv(recv, q)
            dist.send(need[q], q)
        asked.append(recv)
    # a global vector known to every rank (same seed); the local copy holds the owned part only
    xg = np.random.default_rng(5).standard_normal(syn.nb * s)
    xl = np.zeros((nbl + halo.size) * s)
    xl[: nbl * s] = xg[r0 * s: r1 * s]
    for q in range(world):          # halo exchange
        if q == rank:
            continue
        send = torch.from_numpy(xg.reshape(-1, s)[asked[q].numpy()].copy()) if asked[q].numel() else torch.zeros(0, s, dtype=torch.float64)
        assert np.all((asked[q].numpy() >= r0) & (asked[q].numpy() < r1))
        recv = torch.zeros(int(counts[q]), s, dtype=torch.float64)
        if rank < q:
            dist.send(send, q)
            dist.recv(recv, q)
        else:
            dist.recv(recv, q)
            dist.send(send, q)
        off = int(np.searchsorted(halo, bounds[q]))
        xl[(nbl + off) * s:(nbl + off + recv.shape[0]) * s] = recv.numpy().ravel()
    # local SpMV in local numbering == the rank's slice of the global SpMV (oracle on both sides)
    rs_g, ci_g, arr_g, b_g = syn.rows()
    Sg = ol.Sys(s, syn.nb, rs_g, ci_g, arr_g, b_g)
    yg = ol.oracle_assign(Sg, xg)
    cl = s + s % 2
    blocks = arr.reshape(-1, s, cl)[:, :, :s]
    xs = xl.reshape(-1, s)[loc]
    contrib = np.einsum("kcr,kc->kr", blocks, xs)
    rowid = np.repeat(np.arange(nbl), rs)
    yl = np.zeros((nbl, s))
    np.add.at(yl, rowid, contrib)
    err = np.abs(yl.ravel() - yg[r0 * s:r1 * s]).max() / np.abs(yg).max()
    # all-reduced dot product p.q
    part = torch.tensor([float(yl.ravel() @ xl[: nbl * s])], dtype=torch.float64)
    dist.all_reduce(part)
    derr = abs(part.item() - float(yg @ xg)) / abs(float(yg @ xg))
    # interior rows (no halo column) form one long run for slab partitions
    touches = np.zeros(nbl, bool)
    np.logical_or.at(touches, rowid, ~own)
    out[rank] = (err, derr, int(halo.size), int((~touches).sum()), nbl)
    dist.destroy_process_group()


@pytest.mark.parametrize("preset,n", [("S3-hex", 8), ("S3-tet", 7), ("S2-tri", 14)])
def test_partitioned_spmv_world2_gloo(preset, n):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, preset, n, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        err, derr, nhalo, ninterior, nbl = out[rank]
        assert err < 1e-13
        assert derr < 1e-12
        plane = n * n if preset != "S2-tri" else n
        assert 0 < nhalo <= plane + n + 2          # one node plane (line in 2D) per neighbour
        assert ninterior >= nbl - 2 * plane - 2 * n - 2


def _assembly_worker(rank, world, port, name, out):
    """One rank of a row-partitioned value assembly + elimination: the whole element list and the global id lists on
    every rank, the rank's own rows out (kernel sources through the host emulation), gathered over gloo."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    import emu_lib as em
    import oracle_lib as ol
    pkg = g.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = np.load(os.path.join(ROOT, "tests", "golden", name))
    s, nb = int(G["stride"]), int(G["nb"])
    el = ol.Elements(s, G["elem_ids"], G["elem_ke"], G["scales"])
    bounds = pkg.partition_rows(G["row_size"], world).astype(np.int64)
    part = em.split_rows(G["row_size"], G["column_index"], [int(b) for b in bounds])[rank]
    r0, r1, k0, k1 = part[:4]
    # the halo list of this numbering is what csrc/partition.cpp computes for the rank
    assert np.array_equal(part[6], pkg.partition_halo(r0, r1, part[4], G["column_index"][k0:k1]))
    rc, vals = em.assemble_part(s, part, nb, el.ids, el.ke, el.scales)
    assert rc == 0
    vals, forces, _, _ = em.dirichlet_part(s, part, nb, vals, np.zeros((r1 - r0) * s), G["fix_ids"], G["fix_values"])
    # every rank learns every part (sizes differ: pad to the largest)
    sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([vals.size, forces.size]))
    nv, nf = max(int(t[0]) for t in sizes), max(int(t[1]) for t in sizes)
    pv, pf = torch.zeros(nv, dtype=torch.float64), torch.zeros(nf, dtype=torch.float64)
    pv[:vals.size] = torch.from_numpy(vals)
    pf[:forces.size] = torch.from_numpy(forces)
    allv = [torch.zeros(nv, dtype=torch.float64) for _ in range(world)]
    allf = [torch.zeros(nf, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allv, pv)
    dist.all_gather(allf, pf)
    V = np.concatenate([allv[q][:int(sizes[q][0])].numpy() for q in range(world)])
    F = np.concatenate([allf[q][:int(sizes[q][1])].numpy() for q in range(world)])
    same = lambda a, b: np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))
    out[rank] = (bool(same(em.padded(V, s), G["array_post"])),
                 bool(same(F, G["forces_post"])) if bool(G["forces_comparable"]) else True, int(part[6].size))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["AMIE-2d-s20-assembly.npz", "AMIE-3d-s400-assembly.npz"])
def test_partitioned_assembly_world2_gloo(name):
    """N > 1 form of SURVEY section 8 row f1: two processes, each assembling and eliminating the block rows it owns from
    the whole element list; the parts gathered over gloo are the matrix (and forces) the unmodified FeatureTree solved."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_assembly_worker, args=(world, port, name, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        matrix_ok, forces_ok, nhalo = out[rank]
        assert matrix_ok and forces_ok and nhalo > 0
