#!/usr/bin/env python3
"""Generates the value-assembly / Dirichlet-elimination fixtures with the REAL
reference.  Run in the build container (needs oracle/_ref, built by oracle/build_ref.py); the .npz files are committed.

  AMIE-{2d-s20,3d-s400}-assembly.npz : the elements the unmodified FeatureTree::assemble handed to Assembly
      (ids, cached elementary matrices, scales; dumped by oracle/_ref/amie_e2e_ref) together with the matrix
      Assembly::make_final + setBoundaryConditions produced from them (`array_post`), the eliminated dofs read back
      from that matrix (unit rows) and their imposed values (forces_post[id]).
  bc-rand-s{2,3}.npz : Assembly::setBoundaryConditions itself (oracle/ref_harness.cpp:amie_ref_set_boundary_conditions)
      on a random block system with displacement and nodal-force multipliers, natural-BC and add-to-forces vectors.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol                     # noqa: E402
from conftest import random_spd_blocks      # noqa: E402
from make_golden import read_dump           # noqa: E402


def read_elements(path):
    raw = open(path, "rb").read()
    ne, npe, s = [int(v) for v in np.frombuffer(raw, np.uint64, 3)]
    off = 24
    ids = np.frombuffer(raw, np.uint32, ne * npe, off).reshape(ne, npe).copy(); off += 4 * ne * npe
    ke = np.frombuffer(raw, np.float64, ne * npe * npe * s * s, off).reshape(ne, npe, npe, s * s).copy()
    off += 8 * ne * npe * npe * s * s
    scales = np.frombuffer(raw, np.float64, ne, off).copy()
    return s, ids, ke, scales


def unit_rows(stride, nb, row_size, column_index, array):
    """dofs whose matrix row is e_i^T: the rows setBoundaryConditions rewrote (solvers/assembly.cpp:193-204)."""
    s, cl = stride, stride + stride % 2
    acc = np.concatenate([[0], np.cumsum(row_size)]).astype(np.int64)
    blocks = array.reshape(-1, s, cl)
    out = []
    for k in range(nb):
        cols = column_index[acc[k]:acc[k + 1]]
        B = blocks[acc[k]:acc[k + 1]]
        for m in range(s):
            nz = np.argwhere(B[:, :, m] != 0)
            if len(nz) == 1 and cols[nz[0][0]] == k and nz[0][1] == m and B[nz[0][0], m, m] == 1.0:
                out.append(k * s + m)
    return np.array(out, np.uint32)


def main():
    assert ol.ref() is not None, "oracle/_ref is not built: python oracle/build_ref.py"
    exe = os.path.join(ROOT, "oracle", "_ref", "amie_e2e_ref")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for mode, sampling in (("2d", 20), ("3d", 400)):
        with tempfile.TemporaryDirectory() as tmp:
            subprocess.run([exe, mode, str(sampling), "u.bin", "dump.bin", "el.bin"], check=True, cwd=tmp, env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            stride, nb, rs, ci, arr, f = read_dump(os.path.join(tmp, "dump.bin"))
            s, ids, ke, scales = read_elements(os.path.join(tmp, "el.bin"))
        assert s == stride
        fix = unit_rows(stride, nb, rs, ci, arr)
        out = dict(stride=stride, nb=nb, row_size=rs, column_index=ci, elem_ids=ids, elem_ke=ke, scales=scales,
                   fix_ids=fix, fix_values=f[fix], array_post=arr, forces_post=f,
                   # 2d: every load is an imposed displacement, so forces_post is entirely the elimination's work;
                   # 3d: SET_STRESS_XI adds surface loads outside this path, only the matrix is comparable
                   forces_comparable=(mode == "2d"))
        path = os.path.join(HERE, f"AMIE-{mode}-s{sampling}-assembly.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes;", ids.shape[0], "elements of", ids.shape[1], "nodes;", fix.size, "fixed dofs")
    for stride in (2, 3):
        nb = 60
        rs, ci, arr, b = random_spd_blocks(stride, nb, 100 + stride)
        n = nb * stride
        rng = np.random.default_rng(40 + stride)
        fix = np.sort(rng.choice(n, 23, replace=False)).astype(np.uint32)
        fv = rng.standard_normal(fix.size)
        frc = np.sort(rng.choice(np.setdiff1d(np.arange(n), fix), 11, replace=False)).astype(np.uint32)
        frv = rng.standard_normal(frc.size)
        nat, add = rng.standard_normal(n), rng.standard_normal(n)
        a1, f1, n1, d1 = ol.ref_set_bcs(stride, nb, rs, ci, arr, b, fix, fv, frc, frv, nat, add)
        path = os.path.join(HERE, f"bc-rand-s{stride}.npz")
        np.savez_compressed(path, stride=stride, nb=nb, row_size=rs, column_index=ci, array=arr, forces=b, natural=nat,
                            add_to_forces=add, fix_ids=fix, fix_values=fv, force_ids=frc, force_values=frv,
                            array_post=a1, forces_post=f1, natural_post=n1, add_to_forces_post=d1)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
