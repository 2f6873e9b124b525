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


if __name__ == "__main__":
    main()
