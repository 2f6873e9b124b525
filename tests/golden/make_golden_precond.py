#!/usr/bin/env python3
"""Generates the fixtures for the reference's diagonal preconditioners (solvers/inversediagonal.cpp) with the REAL
reference (oracle/_ref; run in the build container, the .npz files are committed).

  precond-{S3-hex-6,S2-tri-12,rand-s3}.npz : for the preconditioner classes InverseDiagonal (kind 0),
      InverseDiagonalSquared (2), InverseLumpedDiagonal (3) the `diagonal` the class built from the matrix; the result
      of ConjugateGradient::solve / BiConjugateGradientStabilized::solve (x, nit, return value) with an object of
      kind 2 / 3 passed as `precond`, and with a user-written Preconditionner (kind 4: precondition(v, t) is
      t = v * user_diagonal, oracle/ref_harness.cpp:UserDiagonal).
  blockprecond.npz : Inverse2x2Diagonal (solvers/inversediagonal.cpp:84-133) on two stride-2 systems (its `blocks`
      and ConjugateGradient::solve with it), and Amie::det / Amie::invert3x3Matrix on the 3x3 node blocks of a
      stride-3 system and on random 3x3 matrices (utilities/matrixops.cpp:826-847, :681-702).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol                     # noqa: E402
from conftest import make_sys, random_spd_blocks      # noqa: E402
import __graft_entry__ as g                 # noqa: E402


def main():
    assert ol.ref() is not None, "oracle/_ref is not built: python oracle/build_ref.py"
    pkg = g.load_package()
    cases = {"S3-hex-6": make_sys(pkg, ol, "S3-hex", 6), "S2-tri-12": make_sys(pkg, ol, "S2-tri", 12)}
    rs, ci, arr, b = random_spd_blocks(3, 90, 11)
    cases["rand-s3"] = ol.Sys(3, 90, rs, ci, arr, b)
    for name, S in cases.items():
        out = dict(stride=S.stride, nb=S.nb, row_size=S.row_size, column_index=S.column_index, array=S.array, b=S.b)
        for kind in (0, 2, 3):
            out[f"diag{kind}"] = ol.ref_precond_diagonal(S, kind)
        rng = np.random.default_rng(3)
        out["user_diagonal"] = out["diag0"] * rng.uniform(0.5, 1.5, S.n)
        for kind in (2, 3, 4):
            ret, x, nit, _, _ = ol.ref_cg(S, precond=kind, nssor=32, diag=out["user_diagonal"])
            out[f"cg{kind}_ok"], out[f"cg{kind}_x"], out[f"cg{kind}_nit"] = ret, x, nit
            ret, x, nit, _, _ = ol.ref_bicgstab(S, precond=kind, diag=out["user_diagonal"])
            out[f"bicg{kind}_ok"], out[f"bicg{kind}_x"], out[f"bicg{kind}_nit"] = ret, x, nit
        path = os.path.join(HERE, f"precond-{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes;", {k: int(out[k]) for k in out if k.endswith(("_ok", "_nit"))})


def blocks():
    pkg = g.load_package()
    out = {}
    rs, ci, arr, b = random_spd_blocks(2, 60, 5)
    for name, S in (("tri", make_sys(pkg, ol, "S2-tri", 10)), ("rand", ol.Sys(2, 60, rs, ci, arr, b))):
        for k, v in dict(nb=S.nb, row_size=S.row_size, column_index=S.column_index, array=S.array, b=S.b).items():
            out[f"{name}_{k}"] = v
        out[f"{name}_blocks"] = ol.ref_precond_blocks2(S)
        ret, x, nit, _, _ = ol.ref_cg(S, precond=5, nssor=32)
        out[f"{name}_cg_ok"], out[f"{name}_cg_x"], out[f"{name}_cg_nit"] = ret, x, nit
    S3 = make_sys(pkg, ol, "ASR-hex", 5)
    A = S3.to_scipy()
    m = np.stack([A[3 * k:3 * k + 3, 3 * k:3 * k + 3].toarray().ravel() for k in range(S3.nb)])
    m = np.concatenate([m, np.random.default_rng(9).standard_normal((50, 9))])
    det, inv = ol.ref_det_invert3x3(m)
    for k, v in dict(nb=S3.nb, row_size=S3.row_size, column_index=S3.column_index, array=S3.array).items():
        out[f"hex_{k}"] = v
    out["m3"], out["det3"], out["inv3"] = m, det, inv
    path = os.path.join(HERE, "blockprecond.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    blocks()
