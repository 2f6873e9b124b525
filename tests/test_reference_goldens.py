"""The golden files the reference's own test-suite holds for this path: the three
`examples/test/check_behaviour_test_stiffness*_base` files -- `examples/main_check_behaviour.cpp` on the 2-element,
8-unknown sample, one line `<time> <strain*1e3> <stress/1e6> <damage %>` each, compared at 1 % by the reference.
oracle/e2e_harness.cpp mode `check` is that example call by call (the .ini values written out), linked with the
reference solvers (amie_e2e_ref: checked here, on the CPU) and with the drop-in translation units (amie_e2e_b200:
tests/test_gpu_e2e.py).  The golden values are copied here, the files live in /root/reference only."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "amie_e2e_ref")

# which -> (ini, base file, the line of the base file)
GOLDEN = {
    0: ("test_stiffness.ini", "check_behaviour_test_stiffness_base", (0.1, 0.1, 1.0, 0.0)),
    1: ("test_stiffness_with_imposed_deformation.ini", "check_behaviour_test_stiffness_with_imposed_deformation_base", (0.1, 1.0, 9.31323e-16, 0.0)),
    2: ("test_stiffness_with_imposed_stress.ini", "check_behaviour_test_stiffness_with_imposed_stress_base", (0.1, 0.08, 1.16415e-16, 0.0)),
}


def run_check(exe, which, tmp, env_extra=None):
    out = os.path.join(tmp, f"check{which}_{os.path.basename(exe)}.bin")
    env = dict(os.environ, OMP_NUM_THREADS="1", **(env_extra or {}))
    p = subprocess.run([exe, "check", str(which), out], cwd=tmp, capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr[-1500:]
    m = re.search(r"^check: (\S+) (\S+) (\S+) (\S+)$", p.stderr, re.M)
    assert m, p.stderr[-1500:]
    cg = [int(v) for v in re.findall(r"CG \d+ converged after (\d+) iterations", p.stderr)]
    return tuple(float(v) for v in m.groups()), np.fromfile(out, np.float64, offset=8), cg, p.stderr


def matches_golden(line, want):
    """The reference's bar: 1 % -- on entries that are zero up to rounding (a stress of 1e-16 MPa), an absolute 1e-9."""
    return all(abs(a - b) <= max(0.01 * abs(b), 1e-9) for a, b in zip(line, want))


@pytest.mark.parametrize("which", [0, 1, 2])
def test_reference_binary_reproduces_its_own_golden_files(tmp_path, which):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref e2e binaries not prebuilt (no /root/reference at build time)")
    line, u, cg, _ = run_check(REF, which, str(tmp_path))
    assert matches_golden(line, GOLDEN[which][2]), (line, GOLDEN[which])
    assert u.size == 8 and cg and max(cg) <= 8            # 8 unknowns: CG ends within N iterations
