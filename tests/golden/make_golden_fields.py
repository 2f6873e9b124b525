#!/usr/bin/env python3
"""Generates the field-recovery fixtures with the REAL reference.  Run in the build
container (needs oracle/_ref, built by oracle/build_ref.py); the .npz files are committed.

  AMIE-{2d-s20,3di-s400}-fields.npz : an unmodified FeatureTree run (oracle/e2e_harness.cpp; 3di = the S1
      sphere-in-cube with a non-zero imposed strain in the inclusion).  For every element, at its centre: the
      answers ElementState::getField gave for TOTAL_STRAIN_FIELD, MECHANICAL_STRAIN_FIELD, REAL_STRESS_FIELD and their
      PRINCIPAL_* fields,
      and the operands it used (dof ids, shape-function derivatives, the cached inverse Jacobian, the behaviour's
      tensor and imposed strain / stress), plus the solution vector `u` the element states were stepped with.
      Tensors are stored as a table of the distinct ones + one index per element.
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def read_fields(path):
    raw = open(path, "rb").read()
    ne, npe, dim, nc = [int(v) for v in np.frombuffer(raw, np.uint64, 4)]
    off = [32]

    def take(shape, t=np.float64):
        n = int(np.prod(shape))
        a = np.frombuffer(raw, t, n, off[0]).reshape(shape).copy()
        off[0] += a.nbytes
        return a
    d = dict(dim=dim, ids=take((ne, npe), np.uint32), dshape=take((ne, npe, dim)), jinv=take((ne, dim, dim)))
    tensor, istrain, istress = take((ne, nc, nc)), take((ne, nc)), take((ne, nc))
    d.update(total_strain=take((ne, nc)), mechanical_strain=take((ne, nc)), real_stress=take((ne, nc)),
             state_displacements=take((ne, npe, dim)), principal_total_strain=take((ne, dim)),
             principal_mechanical_strain=take((ne, dim)), principal_real_stress=take((ne, dim)))
    assert off[0] == len(raw)
    # behaviours -> table of distinct (tensor, imposed strain, imposed stress) + index per element
    key = np.concatenate([tensor.reshape(ne, -1), istrain, istress], axis=1)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)
    d.update(tensors=uniq[:, :nc * nc].reshape(-1, nc, nc).copy(), imposed_strain=uniq[:, nc * nc:nc * nc + nc].copy(),
             imposed_stress=uniq[:, nc * nc + nc:].copy(), tensor_of_elem=inv.reshape(-1).astype(np.uint32))
    return d


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "amie_e2e_ref")
    assert os.path.exists(exe), "oracle/_ref is not built: python oracle/build_ref.py"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for mode, sampling in (("2d", 20), ("3di", 400)):
        with tempfile.TemporaryDirectory() as tmp:
            subprocess.run([exe, mode, str(sampling), "u.bin", "dump.bin", "el.bin", "fields.bin"], check=True, cwd=tmp,
                           env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            d = read_fields(os.path.join(tmp, "fields.bin"))
            raw = open(os.path.join(tmp, "u.bin"), "rb").read()
            n = int(np.frombuffer(raw, np.uint64, 1)[0])
            d["u"] = np.frombuffer(raw, np.float64, n, 8).copy()
        # what ElementState::step gathered is the solution at the element's dof ids
        assert np.array_equal(d["u"].reshape(-1, d["dim"])[d["ids"]], d.pop("state_displacements"))
        path = os.path.join(HERE, f"AMIE-{mode}-s{sampling}-fields.npz")
        np.savez_compressed(path, **d)
        print("wrote", path, os.path.getsize(path), "bytes;", d["ids"].shape[0], "elements of", d["ids"].shape[1],
              "nodes;", d["tensors"].shape[0], "distinct behaviours;", np.count_nonzero(d["imposed_strain"]), "imposed-strain entries")


if __name__ == "__main__":
    main()
