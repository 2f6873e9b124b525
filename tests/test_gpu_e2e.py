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


@pytest.mark.parametrize("mode,sampling", [("2d", 48), ("3d", 500), ("3d", 1000), ("asr2d", 64)])
def test_featuretree_step_with_dropin_solvers(tmp_path, mode, sampling):
    """2d / 3d: BASELINE.json configs 1 and 2 (3d-1000 = 26 088 unknowns, the runnable form of main_3d_benchmark).
    asr2d: an aggregate with six ExpansiveZone gel pockets -- real XFEM enrichment through the drop-in: 455 enrichment
    block rows behind the 15 124 regular ones, enriched nodes with rows of 10-15 blocks among rows of 7 (31 158 unknowns).
    The 3D member of that family (mode asr, examples/main_3d_asr.cpp, BASELINE.json config 5) cannot be a test: the
    UNMODIFIED reference drops ExpansiveZone3D features when it samples the tree (features/features.cpp:1811, :2561-2580)
    and corrupts its heap in the elementary matrices of the enriched tetrahedra when they are added afterwards
    (`malloc_consolidate(): invalid chunk size` before any solver call)."""
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
    # The 3D S1 system has a handful of DOFs that the Krylov iteration does not pin down: the UNMODIFIED reference
    # returns different values for exactly these DOFs when only its OpenMP thread count changes (rel-L2 7e-3 to 9e-3
    # between 1 thread and 2, 3 or 8 threads; none with 4), i.e. they react to last-bit rounding of the dot products.
    # The set is PINNED: tests/golden/e2e-3d-500-rounding-sensitive-dofs.npy = the 27 DOFs (9 nodes) on which the
    # reference differs from its own 1-thread run at 2 / 3 / 4 / 8 threads (generated in the build container with
    # oracle/_ref/amie_e2e_ref, tests/golden/make_golden_e2e_dofs.py).  Every DOF outside it must agree to 1e-8; nothing
    # outside it may be loose.  (Both binaries run on ONE thread here, and every GPU run so far left none of the pinned
    # DOFs loose either.)
    d = np.abs(u_gpu - u_ref)
    loose = d > 1e-7 * np.abs(u_ref).max()
    if (mode, sampling) == ("3d", 1000):
        # 26 088 unknowns: FeatureTree::step ends with a BiCGStab pass whose stopping test is loose (error ~2e-6) and whose
        # path is chaotic (55 iterations in the reference, 89 here); on this mesh it leaves a few
        # dozen DOFs undetermined to ~1e-3 -- the UNMODIFIED reference moves 31 to 61 DOFs by a relative L2 of 3.2e-3 to
        # 4.3e-3 when only its OpenMP thread count changes, and the union of those sets keeps growing with every thread
        # count tried (101 DOFs after nine), so there is no closed set to pin as for sampling 500.  What is checked: the two
        # CG solves take the reference's iteration counts (above), BiCGStab converges, more than 99 % of the DOFs agree
        # with the reference to 1e-7 of max|u|, and the field as a whole is as close to the reference as the reference
        # is to itself across thread counts.
        err_all = rel_l2(u_gpu, u_ref)
        print(f"e2e {mode}-{sampling}: {u_ref.size} DOF, CG {cg_ref} vs {cg_gpu}, BiCGStab {bi_ref} vs {bi_gpu}, "
              f"{int(loose.sum())} DOF differ by more than 1e-7 max|u|, rel-L2 {err_all:.3e} "
              f"(on the others {rel_l2(u_gpu[~loose], u_ref[~loose]):.3e})")
        # observed on B200: 19 and 51 such DOFs, rel-L2 1.5e-3 and 2.6e-3 (the reference binary itself lands on 673 or 675
        # iterations for the first solve from one 1-thread run to the next); the bounds leave twice the reference's spread
        assert loose.sum() <= 0.01 * u_ref.size, int(loose.sum())
        assert err_all <= 1e-2, err_all
        return
    pinned = np.zeros(u_ref.size, bool)
    if mode == "3d":
        pinned[np.load(os.path.join(ROOT, "tests", "golden", f"e2e-3d-{sampling}-rounding-sensitive-dofs.npy"))] = True
    err_all = rel_l2(u_gpu, u_ref)
    err = rel_l2(u_gpu[~pinned], u_ref[~pinned])
    print(f"e2e {mode}-{sampling}: {u_ref.size} DOF, CG {cg_ref} vs {cg_gpu}, BiCGStab {bi_ref} vs {bi_gpu}, "
          f"rel-L2 {err:.3e} on {int((~pinned).sum())} DOF ({int(pinned.sum())} pinned rounding-sensitive DOF set aside, "
          f"{int(loose.sum())} of them loose in this run, rel-L2 with them {err_all:.3e})")
    assert not (loose & ~pinned).any(), np.flatnonzero(loose & ~pinned)
    assert err <= 1e-8, err


def read_records(path):
    """out.bin of the tripoint mode: one (uint64 n, n doubles) record per load step."""
    raw = np.fromfile(path, np.uint8)
    recs, off = [], 0
    while off + 8 <= raw.size:
        n = int(raw[off:off + 8].view(np.uint64)[0])
        if off + 8 + 8 * n > raw.size:
            break
        recs.append(raw[off + 8:off + 8 + 8 * n].view(np.float64).copy())
        off += 8 + 8 * n
    return recs


def run_until(exe, args, tmp, max_solves, timeout_s, env_extra=None):
    """Run a harness mode and stop it after `max_solves` solver lines: FeatureTree::step has no usable iteration bound
    once damage starts (features/features.cpp:6244 needs a checkpoint), and the run is deterministic, so the first
    K solves of two binaries are comparable."""
    import time
    out = os.path.join(tmp, os.path.basename(exe) + ".bin")
    env = dict(os.environ, OMP_NUM_THREADS="1", **(env_extra or {}))
    p = subprocess.Popen([exe] + [a if a != "@OUT" else out for a in args], cwd=tmp, stdout=subprocess.DEVNULL,
                         stderr=subprocess.PIPE, text=True, env=env)
    cg, bi, log, notes, t0 = [], [], [], set(), time.time()
    pat_cg = re.compile(r"CG \d+ converged after (\d+) iterations")
    pat_bi = re.compile(r"BiCGStab \d+ converged after (\d+) iterations")
    for line in p.stderr:
        log.append(line[-300:])
        for part in line.split("amie_b200: ")[1:]:
            notes.add(part.split(":")[0].strip()[:60])           # the shim's own messages, whatever progress output surrounds them
        m = pat_cg.search(line)
        if m:
            cg.append(int(m.group(1)))
        m = pat_bi.search(line)
        if m:
            bi.append(int(m.group(1)))
        if len(cg) + len(bi) >= max_solves or time.time() - t0 > timeout_s:
            break
    p.kill()
    p.wait()
    return read_records(out), cg, bi, "".join(log[-40:]) + "\nshim notes: " + " | ".join(sorted(notes))


def test_tripoint_damage_resolves_through_the_dropin(tmp_path):
    """BASELINE.json config 4: examples/main_tripoint.cpp (`24 0 3.9 1.2`, 8 994 unknowns): load steps on ONE topology,
    the damage iterations re-assembling the values and re-solving again and again.  The same unmodified FeatureTree
    driver with the reference solvers and with the drop-in translation units: the elastic load steps must give the
    same displacement field (1e-8), and the CG iteration counts of the first 48 solver calls -- damage iterations included, where
    every solution feeds the next matrix -- must agree within +-2."""
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("oracle/_ref e2e binaries not prebuilt (no /root/reference at build time)")
    args = ["tripoint", "24", "@OUT", "6", "900", "4e-5"]
    K = 48
    u_ref, cg_ref, bi_ref, _ = run_until(REF, args, str(tmp_path), K, 600)
    u_gpu, cg_gpu, bi_gpu, log = run_until(B200, args, str(tmp_path), K, 600, {"AMIE_B200_SHIM_TRACE": "1"})
    assert "amie_b200: set_" not in log and "no CPU fallback" not in log, log[-1500:]
    # the shim uploads the matrix when it changed and only then: the second CG and the BiCGStab of a step see the same array
    assert "matrix unchanged since the last solve" in log and "matrix upload" in log, log[-600:]
    n = min(len(cg_ref), len(cg_gpu))
    print(f"tripoint: {len(u_ref)} / {len(u_gpu)} load steps written, {len(cg_ref)} / {len(cg_gpu)} CG solves, "
          f"first counts {cg_ref[:8]} vs {cg_gpu[:8]}, last {cg_ref[n - 4:n]} vs {cg_gpu[n - 4:n]}")
    assert n >= 24, (len(cg_ref), len(cg_gpu))           # 2 CG + 1 BiCGStab line per solve triple
    assert len(u_ref) >= 3 and len(u_gpu) >= 3
    worst = max(abs(a - b) for a, b in zip(cg_ref[:n], cg_gpu[:n]))
    assert worst <= 2, [(i, a, b) for i, (a, b) in enumerate(zip(cg_ref[:n], cg_gpu[:n])) if abs(a - b) > 2][:5]
    for step, (a, b) in enumerate(zip(u_ref, u_gpu)):
        assert a.size == b.size == 8994
        if np.abs(a).max() == 0.0:
            assert np.abs(b).max() == 0.0          # the first step carries no load
        else:
            assert rel_l2(b, a) <= 1e-8, (step, rel_l2(b, a))


@pytest.mark.parametrize("which", [0, 1, 2])
def test_reference_golden_files_through_the_dropin(tmp_path, which):
    """The three golden files of the reference's own test-suite on this path (examples/test/check_behaviour_test_stiffness*
    _base, 1 % bar; tests/test_reference_goldens.py) with the drop-in solvers under the unmodified FeatureTree: the line
    the example writes, the 8 displacements against the reference binary, the CG counts."""
    from test_reference_goldens import GOLDEN, run_check, matches_golden
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("oracle/_ref e2e binaries not prebuilt (no /root/reference at build time)")
    line_ref, u_ref, cg_ref, _ = run_check(REF, which, str(tmp_path))
    line_gpu, u_gpu, cg_gpu, log = run_check(B200, which, str(tmp_path))
    assert "amie_b200:" not in log, log[-1500:]
    assert matches_golden(line_gpu, GOLDEN[which][2]), (line_gpu, GOLDEN[which])
    assert len(cg_ref) == len(cg_gpu) and all(abs(a - b) <= 2 for a, b in zip(cg_ref, cg_gpu)), (cg_ref, cg_gpu)
    assert np.abs(u_gpu - u_ref).max() <= 1e-8 * np.abs(u_ref).max()


def test_reference_xfem_test_program_through_the_dropin(tmp_path):
    """examples/test/main_test_xfem.cpp (the reference's own XFEM test; tests/test_reference_goldens.py) with the drop-in
    solvers: the enrichment changes between the stages, so the drop-in sees three different topologies (222, 246 and 186
    unknowns) behind one Assembly and must notice each change.  Displacements of every stage against the reference
    binary, CG counts, the averaged fields the example prints."""
    from test_reference_goldens import run_xfem
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("oracle/_ref e2e binaries not prebuilt (no /root/reference at build time)")
    lines_ref, recs_ref, cg_ref, _ = run_xfem(REF, str(tmp_path))
    lines_gpu, recs_gpu, cg_gpu, log = run_xfem(B200, str(tmp_path), {"AMIE_B200_SHIM_TRACE": "1"})
    assert "no CPU fallback" not in log, log[-1500:]
    assert len(recs_ref) == len(recs_gpu) == 3 and [r.size for r in recs_ref] == [r.size for r in recs_gpu]
    assert len(cg_ref) == len(cg_gpu) and all(abs(a - b) <= 2 for a, b in zip(cg_ref, cg_gpu)), (cg_ref, cg_gpu)
    for a, b in zip(recs_ref, recs_gpu):
        assert rel_l2(b, a) <= 1e-8
    for a, b in zip(lines_ref, lines_gpu):
        assert all(abs(x - y) <= max(1e-6 * abs(x), 1e-9) for x, y in zip(a, b)), (a, b)
    print(f"xfem test program: unknowns {[r.size for r in recs_ref]}, CG {cg_ref} vs {cg_gpu}")
