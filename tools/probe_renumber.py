#!/usr/bin/env python3
"""What a numbering without locality costs the SpMV and what the renumbering gives back: an S3-tet-n system whose nodes
are renumbered at random on the host (a stand-in for the mesher's numbering of a large unstructured mesh, which the
reference mesher makes too expensive to produce), uploaded as is and with Assembly(renumber=True).  One JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
syn = pkg.Synth("S3-tet", n)
rs, ci, arr, b = syn.rows()
s = syn.stride
nb = syn.nb
cl = s + s % 2
out = {"case": f"S3-tet-{n}", "ndof": int(nb * s), "blocks": int(ci.size)}


def spmv_rate(matrix_arrays, renumber, label):
    rs_, ci_, arr_, b_ = matrix_arrays
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(rs_, ci_, s, arr_), b_, device=0, renumber=renumber)
    asm.sync_matrix()
    asm.upload_rhs(asm.to_device_order(b_))
    asm.upload_x0(asm.to_device_order(b_))
    ms = asm.spmv_resident(reps=20)
    st = asm.stats()
    out[label] = {"spmv_ms": ms, "gbs": st.spmv_algorithmic_bytes / (ms * 1e-3) / 1e9,
                  "structure_ms": st.structure_ms, "values_ms": st.values_ms}
    cg = pkg.ConjugateGradient(asm)
    cg.nssor = 32
    ok = cg.solve(None, None, 1e-10, 200)          # 200 iterations are enough to compare rates
    out[label]["pcg_it_per_s"] = cg.nit / (asm.stats().solve_ms * 1e-3)
    x = cg.x
    asm.close()
    return x


x_ref = spmv_rate((rs, ci, arr, b), False, "lexicographic numbering")
# scramble: node i of the generator becomes node scr[i]
rng = np.random.default_rng(0)
scr = rng.permutation(nb).astype(np.uint32)
rs2, ci2, frm = pkg.permute_structure(rs, ci, scr)
arr2 = arr.reshape(-1, s * cl)[frm].reshape(-1)
b2 = np.empty_like(b)
b2.reshape(-1, s)[scr] = b.reshape(-1, s)
x_scr = spmv_rate((rs2, ci2, arr2, b2), False, "random numbering")
x_rcm = spmv_rate((rs2, ci2, arr2, b2), True, "random numbering, renumber=True")
back = lambda x: x.reshape(-1, s)[scr].reshape(-1)
out["same_iterate_after_200"] = {"scrambled": float(np.linalg.norm(back(x_scr) - x_ref) / np.linalg.norm(x_ref)),
                                 "renumbered": float(np.linalg.norm(back(x_rcm) - x_ref) / np.linalg.norm(x_ref))}
print(json.dumps(out), flush=True)
