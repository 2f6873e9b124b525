#!/usr/bin/env python3
"""tests/golden/e2e-3d-<sampling>-rounding-sensitive-dofs.npy: the DOFs of the 3D S1 end-to-end case
(oracle/e2e_harness.cpp, mode 3d; sampling 500: 4 641 DOF, sampling 1000: 26 088 DOF = the runnable form of
examples/main_3d_benchmark.cpp, BASELINE.json config 2) on which the UNMODIFIED reference differs from its own 1-thread
run when only OMP_NUM_THREADS changes (2, 3, 4, 8 threads; more than 1e-7 of max|u|).  Run in the build container (needs
oracle/_ref/amie_e2e_ref): `python make_golden_e2e_dofs.py [sampling ...]`; tests/test_gpu_e2e.py sets exactly these
DOFs aside."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "amie_e2e_ref")


def run(threads, tmp, sampling):
    out = os.path.join(tmp, f"u{threads}.bin")
    subprocess.run([REF, "3d", str(sampling), out], cwd=tmp, env=dict(os.environ, OMP_NUM_THREADS=str(threads)),
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True, timeout=600)
    return np.fromfile(out, np.float64, offset=8)


def main(sampling):
    with tempfile.TemporaryDirectory() as tmp:
        ref = run(1, tmp, sampling)
        loose = np.zeros(ref.size, bool)
        for t in (2, 3, 4, 8):
            u = run(t, tmp, sampling)
            l = np.abs(u - ref) > 1e-7 * np.abs(ref).max()
            print(f"{t} threads: {int(l.sum())} DOF differ from the 1-thread run, rel-L2 {np.linalg.norm(u - ref) / np.linalg.norm(ref):.3e}")
            loose |= l
    idx = np.flatnonzero(loose).astype(np.int64)
    np.save(os.path.join(HERE, f"e2e-3d-{sampling}-rounding-sensitive-dofs.npy"), idx)
    print(len(idx), "DOF pinned:", idx.tolist())


if __name__ == "__main__":
    # (sampling 1000 has no closed set: the union keeps growing with every thread count -- tests/test_gpu_e2e.py)
    for smp in ([int(a) for a in sys.argv[1:]] or [500]):
        main(smp)
