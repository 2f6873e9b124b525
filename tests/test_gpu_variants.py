"""Kernel selections that must not change results: the generic slot loop of the field-recovery kernel against the
unrolled default, and the renumbered device matrix against the caller's numbering.  (The assembly / elimination variants
of round 1 lost their A/B on the GPU and are gone: profiles/r02_notes.md.)"""
import numpy as np
import pytest


pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 1])
def test_fields_variant_matches_reference_bits(pkg, ol, variant):
    """Option "fields_variant": 1 (default) the unrolled, phase-split field kernel for linear triangles / tetrahedra,
    0 the generic slot loop."""
    import glob
    import os
    from test_gpu_recovery import context_for, same_bits
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*-fields.npz"))):
        g = np.load(path)
        asm, dim = context_for(pkg, g)
        asm.set_option("fields_variant", variant)
        asm.set_element_kinematics(dim, g["ids"], g["dshape"], g["jinv"])
        asm.set_element_behaviour(g["tensors"], g["imposed_strain"], g["imposed_stress"], g["tensor_of_elem"])
        tot, mech, sig = asm.element_fields(g["u"])
        assert same_bits(tot, g["total_strain"]) and same_bits(mech, g["mechanical_strain"]) and same_bits(sig, g["real_stress"])
        # unused slots inside the element with garbage derivatives, per-element behaviours
        rng = np.random.default_rng(5)
        ne = g["ids"].shape[0]
        nc = tot.shape[1]
        idsr, dsr = g["ids"].copy(), g["dshape"].copy()
        hole = rng.integers(0, idsr.shape[1], ne)
        pick = rng.random(ne) < 0.3
        idsr[pick, hole[pick]] = 0xFFFFFFFF
        dsr[pick, hole[pick], :] = np.nan
        C = g["tensors"][g["tensor_of_elem"]] * rng.uniform(0.2, 1.0, (ne, 1, 1))
        es, ss = 1e-4 * rng.standard_normal((ne, nc)), rng.standard_normal((ne, nc))
        asm.set_element_kinematics(dim, idsr, dsr, g["jinv"])
        asm.set_element_behaviour(C, es, ss, None)
        got = asm.element_fields(g["u"])
        want = ol.oracle_element_fields(dim, idsr, dsr, g["jinv"], g["u"], C, es, ss, None)
        for a, b in zip(got, want):
            assert same_bits(a, b)
        asm.close()


@pytest.mark.parametrize("name", ["AMIE-3d-s400.npz", "AMIE-2d-s20.npz", "rand-s4.npz"])
def test_renumbered_device_matrix_gives_the_same_solve(pkg, ol, name):
    """Assembly(renumber=True): the device works on the reverse-Cuthill-McKee numbering (structure permuted on the host,
    values scattered through the block map by set_values), the caller keeps its own.  Same SpMV, inverse diagonal,
    PCG and BiCGStab answers as the reference on the original numbering."""
    import os
    from conftest import rel_l2
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    s, nb = int(G["stride"]), int(G["nb"])
    S = ol.Sys(s, nb, G["row_size"], G["column_index"], G["array"], G["b"])
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(G["row_size"], G["column_index"], s, G["array"]), G["b"], device=0,
                       renumber=True)
    v = G["v"]
    y = asm.spmv(v)
    assert asm.perm is not None and not np.array_equal(asm.perm, np.arange(nb))
    assert np.abs(y - G["assign"]).max() <= 1e-12 * (np.abs(G["assign"]).max() + 1e-300)
    assert np.array_equal(asm.inverse_diagonal(), G["inverse_diagonal"])
    cg = pkg.ConjugateGradient(asm)
    cg.nssor = 32
    assert cg.solve() == bool(G["cg_ok"])
    assert abs(int(cg.nit) - int(G["cg_nit"])) <= 2 and rel_l2(cg.x, G["cg_x"]) <= 1e-8
    cg.nssor = 0
    assert cg.solve(0.5 * G["cg_x"]) == bool(G["cg_warm_ok"])
    assert abs(int(cg.nit) - int(G["cg_warm_nit"])) <= 2 and rel_l2(cg.x, G["cg_warm_x"]) <= 1e-8
    bi = pkg.BiConjugateGradientStabilized(asm)
    assert bi.solve() == bool(G["bicg_ok"]) and rel_l2(bi.x, G["bicg_x"]) <= 1e-8
    r, nrm = asm.residual(cg.x)
    assert nrm <= 1e-6 * np.linalg.norm(G["b"])
    # values change, structure and map stay
    asm.getMatrix().array[:] = G["array"] * 2.0
    asm.values_changed()
    cg.nssor = 32
    assert cg.solve() and rel_l2(cg.x, 0.5 * G["cg_x"]) <= 1e-8
    asm.close()


def test_spmv_2x2_pipeline_on_ragged_rows(pkg, ol, systems):
    """The 2x2 pipeline (lane <-> (block row, COLUMN): one LDS.128 + one LDS.64 per block, partial sums exchanged once
    per tile) forced on any row lengths ("spmv_variant" 3) against the plain kernel ("spmv_variant" 8) and the oracle:
    uniform T3 rows, the rowstart / colstart forms, ragged rows (short tiles, oversize tiles that read from global
    memory).  On B200 the column mapping gave the row mapping's bits on these very cases before the row mapping was
    deleted (profiles/r02l_pytest_rows.log); what stays checkable is the 1e-13 bar against an independent kernel."""
    from test_gpu_parity import assembly_of
    rng = np.random.default_rng(11)
    S = systems("S2-tri", 40)
    cases = [S]
    nb = 500
    rows = []
    for r in range(nb):
        k = int(rng.choice([1, 2, 5, 7, 7, 7, 9, 12, 30]))
        rows.append(sorted(set(rng.integers(0, nb, k).tolist()) | {r}))
    rs = np.array([len(c) for c in rows], np.uint32)
    ci = np.array([c for row in rows for c in row], np.uint32)
    cases.append(ol.Sys(2, nb, rs, ci, rng.standard_normal(ci.size * 4), rng.standard_normal(nb * 2)))
    for T in cases:
        asm = assembly_of(pkg, T)
        v = rng.standard_normal(T.n)
        out = {}
        cs = 2 * (T.nb // 3)
        for variant in (8, 3):
            asm.set_option("spmv_variant", variant)
            out[variant] = (asm.spmv(v), asm.spmv(v, minus_b=T.b), asm.spmv(v, rowstart=cs, colstart=cs))
        want = (ol.oracle_assign(T, v, None, 0, 0), ol.oracle_assign(T, v, T.b, 0, 0), ol.oracle_assign(T, v, None, cs, cs))
        for a, b, w in zip(out[8], out[3], want):
            scale = np.abs(w).max()
            assert np.abs(a - b).max() <= 1e-13 * scale and np.abs(b - w).max() <= 1e-12 * scale
        # the same launch twice: the same bits
        asm.set_option("spmv_variant", 3)
        assert np.array_equal(asm.spmv(v), out[3][0])
        if T is S:
            cg = pkg.ConjugateGradient(asm)
            cg.nssor = 32
            ok = cg.solve(None, None, 1e-10, -1)
            ret, x_ref, info = ol.oracle_cg(T, nssor=32)
            assert ok == bool(ret) and abs(int(cg.nit) - int(info.nit)) <= 2
        asm.close()
