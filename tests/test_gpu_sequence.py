"""Assembly::cgsolve's solver part with the displacement history in HBM (SURVEY.md section 8(f) row 3,
csrc/cgsolve.cu): extrapolation bit for bit against the oracle (pinned against the real Assembly::extrapolate,
tests/test_oracle_sequence.py), and a sequence of load steps against the same sequence driven through the oracle."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def device_assembly(pkg, S):
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(S.row_size, S.column_index, S.stride, S.array), S.b, device=0)
    asm.sync_matrix()
    asm.upload_rhs(S.b)
    asm.upload_x0(None)
    return asm


def test_extrapolate_and_history_bits(pkg, ol, systems):
    S = systems("S2-tri", 24)
    asm = device_assembly(pkg, S)
    rng = np.random.default_rng(2)
    a, b, c = rng.standard_normal(S.n), rng.standard_normal(S.n), rng.standard_normal(S.n)
    b[3] = np.nan
    # no history: x0 is the resident x ("displacements"), untouched
    asm.upload_x0(a)
    x0, case = asm.extrapolate()
    assert case == 0 and same_bits(x0, a)
    # first push: history = [x*0, x]  ->  x0 = x + (x - x*0)*factor
    asm.push_history()
    for factor in (1.0, 0.37):
        x0, case = asm.extrapolate(factor)
        want, _ = ol.oracle_extrapolate(a * 0., a, factor)
        assert case == 1 and same_bits(x0, want), factor
        asm.upload_x0(a)
    # second push shifts: history = [a, b]; the NaN in b is scrubbed (in the history too)
    asm.upload_x0(b)
    asm.push_history()
    x0, case = asm.extrapolate(1.0)
    want, b_scrubbed = ol.oracle_extrapolate(a, b, 1.0)
    assert case == 1 and same_bits(x0, want) and not np.isnan(x0).any()
    asm.upload_x0(c)
    asm.push_history()                     # history = [b scrubbed, c]
    x0, case = asm.extrapolate(-1.5)
    want, _ = ol.oracle_extrapolate(b_scrubbed, c, -1.5)
    assert case == 1 and same_bits(x0, want)
    asm.reset_history()
    asm.upload_x0(c)
    x0, case = asm.extrapolate()
    assert case == 0 and same_bits(x0, c)
    asm.close()


def test_history_cleared_when_the_system_size_changes(pkg, ol, systems):
    S, T = systems("S2-tri", 24), systems("S2-tri", 20)
    asm = device_assembly(pkg, S)
    asm.upload_x0(np.ones(S.n))
    asm.push_history()
    asm.push_history()
    # a new topology with another number of dofs on the same context (Assembly re-meshed)
    asm.coordinateIndexedMatrix = pkg.CoordinateIndexedSparseMatrix(T.row_size, T.column_index, T.stride, T.array)
    asm.externalForces = T.b
    asm.sync_matrix()
    asm.upload_x0(np.full(T.n, 7.0))
    x0, case = asm.extrapolate()
    assert case == 2 and not x0.any()      # Vector(0): the solver starts from zeros (solvers/assembly.cpp:1781-1785)
    x0, case = asm.extrapolate()
    assert case == 0                       # and the history is gone
    asm.close()


@pytest.mark.parametrize("preset,n", [("S3-hex", 12), ("S2-tri", 40)])
def test_load_steps_on_the_device_follow_the_reference_sequence(pkg, ol, systems, preset, n):
    """Five load steps (forces scaled and perturbed, as in a loading ramp): device-resident cgsolve vs the same
    sequence through the oracle -- x0 = extrapolate(history), CG, history update."""
    S = systems(preset, n)
    asm = device_assembly(pkg, S)
    asm.nssor = 32
    rng = np.random.default_rng(1)
    hist = []
    x_prev = np.zeros(0)
    nits = []
    for step in range(5):
        b = S.b * (1.0 + 0.25 * step) + 1e-3 * np.abs(S.b).max() * rng.standard_normal(S.n) * (S.b != 0)
        # reference sequence (solvers/assembly.cpp:1850-1868)
        if len(hist) == 2:
            x0, hist[1] = ol.oracle_extrapolate(hist[0], hist[1], 1.0)
        else:
            x0 = x_prev
        ret, x_ref, info = ol.oracle_cg(S, x0=x0 if x0.size else None, nssor=32, b=b)
        hist = [hist[1], x_ref] if len(hist) == 2 else [x_ref * 0., x_ref]
        x_prev = x_ref
        # device
        asm.upload_rhs(b)
        ok, nit, err, rho = asm.cgsolve_resident()
        assert ok == bool(ret)
        assert abs(int(nit) - int(info.nit)) <= 2, (step, nit, info.nit)
        assert rel_l2(asm.download_x(), x_ref) <= 1e-8, step
        nits.append(int(nit))
    # the warm start is worth something: later steps take fewer iterations than the cold first one
    assert min(nits[2:]) < nits[0], nits
    asm.close()


def test_space_time_tension_benchmark_through_featuretree(tmp_path):
    """BASELINE.json configs[0] in the form the reference can run (SURVEY.md section 8 config map, row 1):
    examples/main_tension_benchmark.cpp --space-time, sampling 16 -- a notched space-time damage sample whose
    solves carry rowstart = colstart > 0 -- through an UNMODIFIED FeatureTree, first step + three load steps, once
    with the reference solvers and once with the drop-in translation units (oracle/e2e_harness.cpp, mode 2dst)."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exes = [os.path.join(root, "oracle", "_ref", n) for n in ("amie_e2e_ref", "amie_e2e_b200")]
    if not all(os.path.exists(e) for e in exes):
        pytest.skip("oracle/_ref e2e binaries not prebuilt (no /root/reference at build time)")
    res = []
    for exe in exes:
        out = os.path.join(str(tmp_path), os.path.basename(exe) + ".bin")
        p = subprocess.run([exe, "2dst", "16", out], cwd=str(tmp_path), capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert p.returncode == 0, p.stderr[-2000:]
        cg = [int(m) for m in re.findall(r"CG \d+ converged after (\d+) iterations", p.stderr)]
        bi = [int(m) for m in re.findall(r"BiCGStab \d+ converged after (\d+) iterations", p.stderr)]
        rs = re.search(r"2dst: rowstart (\d+) colstart (\d+)", p.stderr)
        res.append((np.fromfile(out, np.float64, offset=8), cg, bi, rs, p.stderr))
    (u_ref, cg_ref, bi_ref, rs_ref, _), (u_gpu, cg_gpu, bi_gpu, rs_gpu, log) = res
    assert "amie_b200:" not in log, log[-1500:]
    assert rs_ref and int(rs_ref.group(1)) > 0 and rs_ref.group(0) == rs_gpu.group(0)
    assert len(cg_ref) == len(cg_gpu) >= 6 and len(bi_ref) == len(bi_gpu) >= 3
    for a, b in zip(cg_ref, cg_gpu):
        assert abs(a - b) <= 2, (cg_ref, cg_gpu)
    err = rel_l2(u_gpu, u_ref)
    print(f"e2e 2dst-16: {u_ref.size} DOF, rowstart {rs_ref.group(1)}, CG {cg_ref} vs {cg_gpu}, BiCGStab {bi_ref} vs {bi_gpu}, rel-L2 {err:.3e}")
    assert u_ref.size == u_gpu.size and err <= 1e-8, err


def test_early_returns_are_flagged_for_the_shim(pkg, systems):
    """Where the reference returns before its loop it prints "homogeneous" (CG) or nothing (BiCGStab) instead of
    "converged after": the stats carry that so the drop-in TUs print the same lines."""
    S = systems("S3-hex", 12)
    asm = device_assembly(pkg, S)
    ok, nit, err, rho = asm.pcg_resident(nssor=32)
    assert ok and nit > 0 and asm.stats().early_return == 0
    asm.upload_rhs(np.zeros(S.n))
    asm.upload_x0(None)
    ok, nit, err, rho = asm.pcg_resident(nssor=32)
    assert ok and nit == 0 and asm.stats().early_return == 1
    ok, nit, err = asm.bicgstab_resident()
    assert ok and nit == 0 and asm.stats().early_return == 1
    asm.upload_rhs(S.b)
    ok, nit, err = asm.bicgstab_resident()
    assert ok and nit > 0 and asm.stats().early_return == 0
    asm.close()


@pytest.mark.parametrize("mode,sampling", [("2d", 48), ("3d", 400)])
def test_featuretree_step_with_renumbered_device_matrix(tmp_path, monkeypatch, mode, sampling):
    """AMIE_B200_RENUMBER=1: the drop-in TUs hand the device a reverse-Cuthill-McKee renumbering of the assembled
    system and permute b, x0, x at the boundary; the FeatureTree sees its own numbering and the same answers."""
    import os
    import test_gpu_e2e as e2e
    if not (os.path.exists(e2e.REF) and os.path.exists(e2e.B200)):
        pytest.skip("oracle/_ref e2e binaries not prebuilt (no /root/reference at build time)")
    u_ref, cg_ref, bi_ref, _ = e2e.run(e2e.REF, mode, sampling, str(tmp_path))
    monkeypatch.setenv("AMIE_B200_RENUMBER", "1")
    u_gpu, cg_gpu, bi_gpu, log = e2e.run(e2e.B200, mode, sampling, str(tmp_path))
    assert "amie_b200:" not in log, log[-1500:]
    assert len(cg_ref) == len(cg_gpu) == 2 and len(bi_ref) == len(bi_gpu) == 1
    for a, b in zip(cg_ref, cg_gpu):
        assert abs(a - b) <= 2, (cg_ref, cg_gpu)
    d = np.abs(u_gpu - u_ref)
    loose = d > 1e-7 * np.abs(u_ref).max()          # the rounding-sensitive DOFs of the 3D case (tests/test_gpu_e2e.py)
    assert loose.sum() <= (0 if mode == "2d" else 24)
    assert rel_l2(u_gpu[~loose], u_ref[~loose]) <= 1e-8


def test_a_whole_step_without_host_vectors(pkg, ol):
    """INTEGRATION.md section 6 on the 2D FeatureTree run of the fixtures: elements -> assemble -> eliminate -> cgsolve
    with the history in HBM -> strains / stresses from the resident solution.  Only the elementary matrices and the
    multipliers go in, only the fields come out; neither the matrix nor x0 nor x crosses PCIe.  The device-assembled
    matrix and forces equal AMIE's bit for bit, the fields equal ElementState::getField applied to the downloaded x
    bit for bit, and they match the fields AMIE itself computed from its own solve to the solvers' tolerance."""
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    A = np.load(os.path.join(golden, "AMIE-2d-s20-assembly.npz"))
    F = np.load(os.path.join(golden, "AMIE-2d-s20-fields.npz"))
    s, nb = int(A["stride"]), int(A["nb"])
    assert np.array_equal(A["elem_ids"], F["ids"])                      # the two fixtures are one FeatureTree run
    asm = pkg.Assembly(None, None, device=0)
    asm.nssor = 32
    asm.set_structure_only(s, A["row_size"], A["column_index"])
    asm.set_elements(A["elem_ids"])
    asm.set_element_kinematics(s, F["ids"], F["dshape"], F["jinv"])
    asm.set_element_behaviour(F["tensors"], F["imposed_strain"], F["imposed_stress"], F["tensor_of_elem"])
    for step in range(3):                                              # history: none -> [0, x] (x0 = 2x) -> [x, x] (x0 = x)
        asm.update_elements(0, A["elem_ke"], A["scales"])
        asm.assemble()
        asm.upload_rhs(np.zeros(nb * s))
        asm.set_boundary_conditions(A["fix_ids"], A["fix_values"])
        ok, nit, err, rho = asm.cgsolve_resident()
        assert ok
        tot, mech, sig = asm.element_fields()                          # resident x
        if step == 0:
            nit0 = nit
    assert nit < nit0 // 2                                             # third step: x0 = x + (x - x), the answer is already there
    rs, ci, arr, f = asm.download_matrix()
    assert np.array_equal(arr, A["array_post"]) and np.array_equal(f, A["forces_post"])
    x = asm.download_x()
    want = ol.oracle_element_fields(s, F["ids"], F["dshape"], F["jinv"], x, F["tensors"], F["imposed_strain"],
                                    F["imposed_stress"], F["tensor_of_elem"])
    for a, b in zip((tot, mech, sig), want):
        assert same_bits(a, b)
    assert rel_l2(x, F["u"]) <= 1e-8
    assert rel_l2(tot.reshape(-1), F["total_strain"].reshape(-1)) <= 1e-7
    assert rel_l2(sig.reshape(-1), F["real_stress"].reshape(-1)) <= 1e-7
    asm.close()
