#!/usr/bin/env python3
"""Generates tests/golden/*.npz with the REAL reference (oracle/_ref/libamie_ref_oracle.so, compiled
from /root/reference by oracle/build_ref.py).  Run in the build container, where the reference is
present; the .npz fixtures are committed so the oracle stays pinned where it is not.

Each fixture holds a small system in the reference layout plus what the reference computed on it,
single-threaded (OMP result is thread-count dependent in the last bits; 1 thread is the
--no-openmp behaviour): assign(y, A*x), assign(y, A*x-b) with rowstart/colstart, the serial
operator path, inverseDiagonal, ConjugateGradient::solve and BiConjugateGradientStabilized::solve.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g   # noqa: E402
import oracle_lib as ol       # noqa: E402

CASES = [("S3-hex", 6), ("S3-tet", 6), ("S2-tri", 12), ("ASR-hex", 7)]
# systems assembled by AMIE ITSELF (unstructured Delaunay meshes, AMIE's own numbering and Dirichlet
# elimination), dumped by oracle/_ref/amie_e2e_ref after one FeatureTree::step():
#   2d: plain-elastic twin of examples/main_tension_benchmark.cpp ; 3d: S1 sphere-in-cube of examples/main_3d_benchmark.cpp
AMIE_CASES = [("2d", 20), ("3d", 400)]


def read_dump(path):
    raw = open(path, "rb").read()
    stride, nb, nnzb = np.frombuffer(raw, np.uint64, 3)
    stride, nb, nnzb = int(stride), int(nb), int(nnzb)
    off = 24
    rs = np.frombuffer(raw, np.uint32, nb, off); off += 4 * nb
    ci = np.frombuffer(raw, np.uint32, nnzb, off); off += 4 * nnzb
    cl = stride + stride % 2
    arr = np.frombuffer(raw, np.float64, nnzb * stride * cl, off); off += 8 * nnzb * stride * cl
    b = np.frombuffer(raw, np.float64, nb * stride, off)
    return stride, nb, rs.copy(), ci.copy(), arr.copy(), b.copy()


def reference_outputs(S, rng_seed=7):
    rng = np.random.default_rng(rng_seed)
    v = rng.standard_normal(S.n)
    out = dict(stride=S.stride, nb=S.nb, row_size=S.row_size, column_index=S.column_index, array=S.array, b=S.b, v=v)
    out["assign"] = ol.ref_spmv(S, v, None, mode=0)[0]
    out["assign_minus_b"] = ol.ref_spmv(S, v, S.b, mode=1)[0]
    rsn = S.stride * 2
    out["rowstart"] = rsn
    out["assign_minus_b_rowstart"] = ol.ref_spmv(S, v, S.b, mode=1, rowstart=rsn, colstart=rsn)[0]
    out["serial"] = ol.ref_spmv(S, v, None, mode=2)[0]
    out["serial_minus_b"] = ol.ref_spmv(S, v, S.b, mode=3)[0]
    out["inverse_diagonal"] = ol.ref_inverse_diagonal(S)
    ok, x, nit, _, _ = ol.ref_cg(S, nssor=32, nthreads=1)
    out["cg_ok"], out["cg_x"], out["cg_nit"] = ok, x, nit
    ok, x, nit, _, _ = ol.ref_cg(S, nssor=32, nthreads=1, rowstart=rsn, colstart=rsn)
    out["cg_rs_ok"], out["cg_rs_x"], out["cg_rs_nit"] = ok, x, nit
    ok, x, nit, _, _ = ol.ref_cg(S, nssor=0, nthreads=1, x0=0.5 * out["cg_x"])
    out["cg_warm_ok"], out["cg_warm_x"], out["cg_warm_nit"] = ok, x, nit
    ok, x, nit, _, _ = ol.ref_bicgstab(S, nthreads=1)
    out["bicg_ok"], out["bicg_x"], out["bicg_nit"] = ok, x, nit
    return out


def main():
    pkg = g.load_package()
    assert ol.ref() is not None, "oracle/_ref is not built: python oracle/build_ref.py"
    for preset, n in CASES:
        syn = pkg.Synth(preset, n)
        rs, ci, arr, b = syn.rows()
        S = ol.Sys(syn.stride, syn.nb, rs, ci, arr, b)
        rng = np.random.default_rng(7)
        v = rng.standard_normal(S.n)
        out = dict(stride=S.stride, nb=S.nb, row_size=rs, column_index=ci, array=arr, b=b, v=v)
        out["assign"] = ol.ref_spmv(S, v, None, mode=0)[0]
        out["assign_minus_b"] = ol.ref_spmv(S, v, b, mode=1)[0]
        rsn = S.stride * 2
        out["rowstart"] = rsn
        out["assign_minus_b_rowstart"] = ol.ref_spmv(S, v, b, mode=1, rowstart=rsn, colstart=rsn)[0]
        out["serial"] = ol.ref_spmv(S, v, None, mode=2)[0]
        out["serial_minus_b"] = ol.ref_spmv(S, v, b, mode=3)[0]
        out["inverse_diagonal"] = ol.ref_inverse_diagonal(S)
        ok, x, nit, _, _ = ol.ref_cg(S, nssor=32, nthreads=1)
        out["cg_ok"], out["cg_x"], out["cg_nit"] = ok, x, nit
        ok, x, nit, _, _ = ol.ref_cg(S, nssor=32, nthreads=1, rowstart=rsn, colstart=rsn)
        out["cg_rs_ok"], out["cg_rs_x"], out["cg_rs_nit"] = ok, x, nit
        ok, x, nit, _, _ = ol.ref_cg(S, nssor=0, nthreads=1, x0=0.5 * out["cg_x"])
        out["cg_warm_ok"], out["cg_warm_x"], out["cg_warm_nit"] = ok, x, nit
        ok, x, nit, _, _ = ol.ref_bicgstab(S, nthreads=1)
        out["bicg_ok"], out["bicg_x"], out["bicg_nit"] = ok, x, nit
        path = os.path.join(HERE, f"{preset}-{n}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes; CG nit", out["cg_nit"], "BiCGStab nit", out["bicg_nit"])
    from conftest import random_spd_blocks
    for stride in (1, 4, 6):                      # inner_product's other stride cases
        rs, ci, arr, b = random_spd_blocks(stride, 60, 30 + stride)
        S = ol.Sys(stride, 60, rs, ci, arr, b)
        out = reference_outputs(S)
        path = os.path.join(HERE, f"rand-s{stride}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes; CG nit", out["cg_nit"], "BiCGStab nit", out["bicg_nit"])
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "amie_e2e_ref")
    for mode, sampling in AMIE_CASES:
        with tempfile.TemporaryDirectory() as tmp:
            subprocess.run([exe, mode, str(sampling), os.path.join(tmp, "u.bin"), os.path.join(tmp, "dump.bin")],
                           check=True, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            stride, nb, rs, ci, arr, b = read_dump(os.path.join(tmp, "dump.bin"))
            u = np.fromfile(os.path.join(tmp, "u.bin"), np.float64, offset=8)
        S = ol.Sys(stride, nb, rs, ci, arr, b)
        out = reference_outputs(S)
        out["featuretree_displacements"] = u          # what F.getDisplacements() returned after the step
        path = os.path.join(HERE, f"AMIE-{mode}-s{sampling}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes;", S.n, "DOF,", S.nnzb / S.nb, "blocks/row; CG nit", out["cg_nit"],
              "BiCGStab nit", out["bicg_nit"], "| max row", rs.max(), "min row", rs.min())


if __name__ == "__main__":
    main()
