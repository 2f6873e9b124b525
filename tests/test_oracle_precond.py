"""Pins the oracle's restatement of the reference's diagonal preconditioners (solvers/inversediagonal.cpp:
InverseDiagonalSquared, InverseLumpedDiagonal, and a user-written t = v*d Preconditionner) bit for bit against the
golden vectors the real classes produced (tests/golden/precond-*.npz, make_golden_precond.py) and, where oracle/_ref
is present, against the live reference."""
import glob
import os

import numpy as np
import pytest

from conftest import random_spd_blocks

PRECOND = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "precond-*.npz")))


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def test_fixtures_exist():
    assert len(PRECOND) >= 3


@pytest.mark.parametrize("path", PRECOND, ids=[os.path.basename(p)[:-4] for p in PRECOND])
def test_oracle_matches_precond_golden(ol, path):
    g = np.load(path)
    S = ol.Sys(int(g["stride"]), int(g["nb"]), g["row_size"], g["column_index"], g["array"], g["b"])
    for kind in (0, 2, 3):
        assert same_bits(ol.oracle_precond_diagonal(S, kind), g[f"diag{kind}"]), kind
    for kind in (2, 3, 4):
        ret, x, info = ol.oracle_cg(S, precond=kind, nssor=32, diag=g["user_diagonal"])
        assert ret == int(g[f"cg{kind}_ok"]) and info.nit == int(g[f"cg{kind}_nit"]), kind
        assert same_bits(x, g[f"cg{kind}_x"]), kind
        ret, x, info = ol.oracle_bicgstab(S, precond=kind, diag=g["user_diagonal"])
        assert ret == int(g[f"bicg{kind}_ok"]) and info.nit == int(g[f"bicg{kind}_nit"]), kind
        assert same_bits(x, g[f"bicg{kind}_x"]), kind


@pytest.mark.parametrize("stride", [1, 2, 3, 4, 6])
def test_oracle_precond_matches_live_reference_all_strides(ol, stride):
    if ol.ref() is None:
        pytest.skip("oracle/_ref not built here: tests/golden/precond-*.npz pin the oracle instead")
    rs, ci, arr, b = random_spd_blocks(stride, 40, 30 + stride)
    S = ol.Sys(stride, 40, rs, ci, arr, b)
    for kind in (0, 2, 3):
        assert same_bits(ol.oracle_precond_diagonal(S, kind), ol.ref_precond_diagonal(S, kind)), kind
    ud = ol.oracle_precond_diagonal(S, 0) * np.random.default_rng(stride).uniform(0.5, 1.5, S.n)
    for kind in (2, 3, 4):
        r1 = ol.oracle_cg(S, precond=kind, diag=ud)
        r2 = ol.ref_cg(S, precond=kind, diag=ud)
        assert r1[0] == r2[0] and r1[2].nit == r2[2] and same_bits(r1[1], r2[1]), kind
        b1 = ol.oracle_bicgstab(S, precond=kind, diag=ud)
        b2 = ol.ref_bicgstab(S, precond=kind, diag=ud)
        assert b1[0] == b2[0] and b1[2].nit == b2[2] and same_bits(b1[1], b2[1]), kind
