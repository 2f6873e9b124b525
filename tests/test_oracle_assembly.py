"""Pins the CPU restatement of value assembly + Dirichlet elimination (oracle/amie_oracle_assembly.c; SURVEY.md
section 8 row f1) against the real reference: golden fixtures made by tests/golden/make_golden_assembly.py everywhere,
the live oracle/_ref library where it was prebuilt."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import random_spd_blocks

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.mark.parametrize("name", ["AMIE-2d-s20-assembly.npz", "AMIE-3d-s400-assembly.npz"])
def test_assemble_and_eliminate_match_featuretree_bit_for_bit(name):
    G = load(name)
    s, nb = int(G["stride"]), int(G["nb"])
    el = ol.Elements(s, G["elem_ids"], G["elem_ke"], G["scales"])
    pre = ol.oracle_assemble(s, nb, G["row_size"], G["column_index"], el)
    arr, f, _, _ = ol.oracle_set_bcs(s, nb, G["row_size"], G["column_index"], pre, np.zeros(nb * s), G["fix_ids"],
                                     G["fix_values"])
    assert np.array_equal(arr, G["array_post"])
    if bool(G["forces_comparable"]):
        assert np.array_equal(f, G["forces_post"])
    # the sparsity pattern AMIE built is the union of the element node pairs
    rs, ci = el.pattern(nb)
    assert np.array_equal(rs, G["row_size"]) and np.array_equal(ci, G["column_index"])


@pytest.mark.parametrize("stride", [2, 3])
def test_set_boundary_conditions_golden(stride):
    G = load(f"bc-rand-s{stride}.npz")
    a, f, n, d = ol.oracle_set_bcs(stride, int(G["nb"]), G["row_size"], G["column_index"], G["array"], G["forces"],
                                   G["fix_ids"], G["fix_values"], G["force_ids"], G["force_values"], G["natural"],
                                   G["add_to_forces"])
    assert np.array_equal(a, G["array_post"]) and np.array_equal(f, G["forces_post"])
    assert np.array_equal(n, G["natural_post"]) and np.array_equal(d, G["add_to_forces_post"])


@pytest.mark.parametrize("stride", [1, 2, 3, 4, 6])
def test_set_boundary_conditions_live_reference(stride):
    if ol.ref() is None:
        pytest.skip("oracle/_ref not prebuilt")
    nb = 48
    rs, ci, arr, b = random_spd_blocks(stride, nb, 300 + stride)
    n = nb * stride
    rng = np.random.default_rng(stride)
    for nfix in (0, 1, n // 5, n):
        fix = np.sort(rng.choice(n, nfix, replace=False)).astype(np.uint32)
        fv = rng.standard_normal(nfix)
        rest = np.setdiff1d(np.arange(n), fix)
        frc = np.sort(rng.choice(rest, min(7, rest.size), replace=False)).astype(np.uint32)
        frv = rng.standard_normal(frc.size)
        nat, add = rng.standard_normal(n), rng.standard_normal(n)
        got = ol.oracle_set_bcs(stride, nb, rs, ci, arr, b, fix, fv, frc, frv, nat, add)
        want = ol.ref_set_bcs(stride, nb, rs, ci, arr, b, fix, fv, frc, frv, nat, add)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)


def test_elimination_keeps_the_solution_of_the_constrained_problem():
    """Property: after elimination, solving K'u = f' gives u[fixed] = imposed values and satisfies the free rows of K."""
    stride, nb = 3, 40
    rs, ci, arr, b = random_spd_blocks(stride, nb, 5)
    n = nb * stride
    rng = np.random.default_rng(0)
    fix = np.sort(rng.choice(n, 9, replace=False)).astype(np.uint32)
    fv = rng.standard_normal(9)
    a1, f1, _, _ = ol.oracle_set_bcs(stride, nb, rs, ci, arr, b, fix, fv)
    K = ol.Sys(stride, nb, rs, ci, arr, b).to_scipy().toarray()
    K1 = ol.Sys(stride, nb, rs, ci, a1, f1).to_scipy().toarray()
    u = np.linalg.solve(K1, f1)
    assert np.allclose(u[fix], fv, rtol=0, atol=1e-12)
    free = np.setdiff1d(np.arange(n), fix)
    assert np.allclose((K @ u)[free], b[free], rtol=1e-9, atol=1e-9)
    assert np.allclose(K1, K1.T)
