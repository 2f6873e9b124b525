#!/usr/bin/env python3
"""First device timings of the SURVEY section 8(f) rows next to the solve: value assembly + Dirichlet elimination (f1)
and field recovery (f2), at sizes well beyond L2.  Prints one JSON line per row; run under ncu to capture the kernels
(k_assemble_gather, k_dirichlet, k_element_fields)."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
PEAK = 6548.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def hex_grid(n):
    node = np.arange(n ** 3).reshape(n, n, n)
    corners = [node[tuple(slice(o, n - 1 + o) for o in off)].reshape(-1) for off in np.ndindex(2, 2, 2)]
    return np.stack(corners, 1).astype(np.uint32)


def pattern(ids, nb):
    i64 = ids.astype(np.int64)
    keys = [np.arange(nb, dtype=np.int64) * nb + np.arange(nb)]
    for j in range(ids.shape[1]):
        for k in range(ids.shape[1]):
            keys.append(i64[:, j] * nb + i64[:, k])
    key = np.unique(np.concatenate(keys))
    return np.bincount(key // nb, minlength=nb).astype(np.uint32), (key % nb).astype(np.uint32)


def assembly_row(n=64, reps=5):
    s = 3
    ids = hex_grid(n)
    nb = n ** 3
    rs, ci = pattern(ids, nb)
    rng = np.random.default_rng(0)
    ke = rng.standard_normal((ids.shape[0], 8, 8, 9))
    asm = pkg.Assembly(device=0)
    asm.set_structure_only(s, rs, ci)
    asm.set_elements(ids)
    asm.update_elements(0, ke)
    asm.assemble()
    full = []
    for _ in range(reps):
        asm.update_elements(0, ke)          # marks everything
        asm.assemble()
        full.append(asm.stats().assemble_ms)
    part = []
    cnt = ids.shape[0] // 100               # a damage step touching 1 % of the elements
    for r in range(reps):
        asm.update_elements(r * cnt, ke[r * cnt:(r + 1) * cnt])
        asm.assemble()
        part.append(asm.stats().assemble_ms)
    nnzb = int(ci.size)
    alg = ke.size * 8 + nnzb * 4 + nnzb * (8 * s * s + 4)
    asm.upload_rhs(np.zeros(nb * s))
    fix = np.arange(0, nb * s, 97, dtype=np.uint32)
    asm.set_boundary_conditions(fix, np.ones(fix.size))
    rec = dict(row="f1 assembly", nodes=nb, elements=int(ids.shape[0]), nnzb=nnzb, algorithmic_bytes=alg,
               assemble_full_ms=min(full), assemble_full_gbs=alg / (min(full) * 1e-3) / 1e9,
               frac_of_peak=alg / (min(full) * 1e-3) / 1e9 / PEAK, assemble_1pct_ms=min(part),
               dirichlet_ms=asm.stats().bc_ms, set_elements_ms=asm.stats().elements_ms)
    rec["algorithmic_bytes_note"] = "element blocks once + 4 B of run offsets per stored block + stored blocks written once"
    # what a damage step costs on the two routes (BASELINE.json config 4: repeated re-solves on one topology):
    #   route A (drop-in shim): the host re-assembles and set_values uploads the whole padded array;
    #   route B (rows f1): update_elements of the elements that changed (1 %) + assemble of the touched blocks
    import time
    cl = s + s % 2
    arr = np.zeros(nnzb * s * cl)
    asm.set_values(arr)
    asm.set_values(arr)
    rec["route_A_set_values_ms"] = asm.stats().values_ms
    rec["route_A_host_bytes"] = int(arr.nbytes)
    t = []
    for r in range(reps):
        t0 = time.time()
        asm.update_elements(r * cnt, ke[r * cnt:(r + 1) * cnt])
        asm.assemble()
        t.append(1e3 * (time.time() - t0))
    rec["route_B_update_1pct_plus_assemble_ms"] = min(t)
    rec["route_B_host_bytes"] = int(ke[:cnt].nbytes)
    t = []
    for r in range(2):
        t0 = time.time()
        asm.update_elements(0, ke)
        asm.assemble()
        t.append(1e3 * (time.time() - t0))
    rec["route_B_update_all_plus_assemble_ms"] = min(t)
    rec["route_B_all_host_bytes"] = int(ke.nbytes)
    asm.close()
    return rec


def fields_row(n=100, reps=5):
    """n^3 nodes, 6 linear tetrahedra per cube (the S3-tet connectivity), one behaviour per element (damage-like)
    and a 2-entry table."""
    dim, npe, nc = 3, 4, 6
    cube = hex_grid(n)
    # Kuhn split of the cube (corner numbering of np.ndindex(2,2,2): bit 2 = x, 1 = y, 0 = z)
    tets = [(0, 4, 6, 7), (0, 4, 5, 7), (0, 2, 6, 7), (0, 2, 3, 7), (0, 1, 5, 7), (0, 1, 3, 7)]
    ids = np.concatenate([cube[:, t] for t in tets]).astype(np.uint32)
    ne, nb = ids.shape[0], n ** 3
    rng = np.random.default_rng(1)
    ds = rng.standard_normal((ne, npe, dim))
    ji = rng.standard_normal((ne, dim, dim))
    u = rng.standard_normal(nb * dim)
    asm = pkg.Assembly(device=0)
    rs = np.ones(nb, np.uint32)
    asm.set_structure_only(dim, rs, np.arange(nb, dtype=np.uint32))     # the kernel needs N only
    asm.set_element_kinematics(dim, ids, ds, ji)
    asm.upload_x0(u)
    out = {}
    for label, ntab in (("table2", 2), ("per_element", ne)):
        C = rng.standard_normal((ntab, nc, nc))
        toe = None if ntab == ne else (np.arange(ne) % ntab).astype(np.uint32)
        asm.set_element_behaviour(C, None, None, toe)
        lib, ctx = pkg.lib(), asm.ctx
        ms = []
        for _ in range(reps):
            asm.check(lib.amie_b200_element_fields(ctx, None, 0, None, None, None))   # results stay on the device
            ms.append(asm.stats().fields_ms)
        # the solution is counted once per node (the per-slot gathers hit L2), csrc/fields.cu header
        alg = ne * (4 * npe + 8 * npe * dim + 8 * dim * dim + 4 + 3 * 8 * nc) + 8 * dim * nb
        if ntab == ne:
            alg += ne * 8 * nc * (nc + 2)
        out[label] = dict(ms=min(ms), algorithmic_bytes=alg, gbs=alg / (min(ms) * 1e-3) / 1e9,
                          frac_of_peak=alg / (min(ms) * 1e-3) / 1e9 / PEAK, melem_per_s=ne / (min(ms) * 1e-3) / 1e6)
        # the unrolled, phase-split kernel (option "fields_variant" = 1; same bits)
        ref = [np.zeros((ne, nc)) for _ in range(3)]
        asm.check(lib.amie_b200_element_fields(ctx, None, 0, *[r.ctypes.data_as(ctypes.c_void_p) for r in ref]))
        asm.set_option("fields_variant", 1)
        got = [np.zeros((ne, nc)) for _ in range(3)]
        ms1 = []
        for _ in range(reps):
            asm.check(lib.amie_b200_element_fields(ctx, None, 0, None, None, None))
            ms1.append(asm.stats().fields_ms)
        asm.check(lib.amie_b200_element_fields(ctx, None, 0, *[r.ctypes.data_as(ctypes.c_void_p) for r in got]))
        asm.set_option("fields_variant", 0)
        out[label].update(variant1_ms=min(ms1), variant1_frac_of_peak=alg / (min(ms1) * 1e-3) / 1e9 / PEAK,
                          variant1_same_bits=bool(all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in zip(ref, got))))
    asm.close()
    return dict(row="f2 element fields", elements=ne, nodes=nb, **out)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "fields"):
        print(json.dumps(fields_row()), flush=True)
    if which in ("all", "assembly"):
        print(json.dumps(assembly_row()), flush=True)
