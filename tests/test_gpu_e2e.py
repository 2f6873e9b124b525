"""End to end through an UNMODIFIED FeatureTree (SURVEY.md §8c "secondary oracle"): the same driver
(oracle/e2e_harness.cpp) linked with the reference solvers and with the drop-in translation units
(xfem-amie_b200/host/shim/*.cpp -> C-ABI -> CUDA).  Both binaries are prebuilt by
oracle/build_ref.py where /root/reference exists and travel in oracle/_ref/."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "amie_e2e_ref")
B200 = os.path.join(ROOT, "oracle", "_ref", "amie_e2e_b200")


def run(exe, mode, sampling, tmp):
    out = os.path.join(tmp, os.path.basename(exe) + ".bin")
    # one OpenMP thread: the reference's own pipeline (meshing / assembly, outside the solver) is not
    # run-to-run reproducible with several threads -- on the 3D case a dozen DOFs of F.getDisplacements()
    # flip between two values from run to run of the UNMODIFIED reference (rel-L2 7e-3 between 1 and 8 threads)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([exe, mode, str(sampling), out], cwd=tmp, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    u = np.fromfile(out, np.float64, offset=8)
    cg = [int(m) for m in re.findall(r"CG \d+ converged after (\d+) iterations", p.stderr)]
    bi = [int(m) for m in re.findall(r"BiCGStab \d+ converged after (\d+) iterations", p.stderr)]
    return u, cg, bi, p.stderr


@pytest.mark.parametrize("mode,sampling", [("2d", 48), ("3d", 500)])
def test_featuretree_step_with_dropin_solvers(tmp_path, mode, sampling):
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("oracle/_ref e2e binaries not prebuilt (no /root/reference at build time)")
    u_ref, cg_ref, bi_ref, _ = run(REF, mode, sampling, str(tmp_path))
    u_gpu, cg_gpu, bi_gpu, log = run(B200, mode, sampling, str(tmp_path))
    assert "amie_b200:" not in log, log[-1500:]
    # FeatureTree::step issues CG, CG, BiCGStab
    assert len(cg_ref) == len(cg_gpu) == 2 and len(bi_ref) == len(bi_gpu) == 1
    for a, b in zip(cg_ref, cg_gpu):
        assert abs(a - b) <= 2, (cg_ref, cg_gpu)
    assert u_ref.size == u_gpu.size and u_ref.size > 1000
    # The 3D S1 system has a handful of DOFs (6 nodes at sampling 500) that the Krylov iteration does not pin
    # down: the UNMODIFIED reference returns different values for exactly these DOFs when only its OpenMP thread
    # count changes (rel-L2 7e-3 between 1 and 8 threads; tools/ notes in profiles/r01_notes.md), i.e. they
    # react to last-bit rounding of the dot products.  They are excluded (and counted) here; every other DOF
    # must agree to 1e-8.
    d = np.abs(u_gpu - u_ref)
    loose = d > 1e-7 * np.abs(u_ref).max()
    err_all = rel_l2(u_gpu, u_ref)
    err = rel_l2(u_gpu[~loose], u_ref[~loose])
    print(f"e2e {mode}-{sampling}: {u_ref.size} DOF, CG {cg_ref} vs {cg_gpu}, BiCGStab {bi_ref} vs {bi_gpu}, "
          f"rel-L2 {err:.3e} on {int((~loose).sum())} DOF ({int(loose.sum())} rounding-sensitive DOF excluded, rel-L2 with them {err_all:.3e})")
    assert loose.sum() <= (0 if mode == "2d" else 24), np.flatnonzero(loose)
    assert err <= 1e-8, err
