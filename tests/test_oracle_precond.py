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
BLOCKS = os.path.join(os.path.dirname(__file__), "golden", "blockprecond.npz")


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


def test_oracle_block_preconditioners_match_golden(ol):
    """Inverse2x2Diagonal's blocks and a PCG solve with it; det / invert3x3Matrix of the 3x3 node blocks (precond kinds
    5 and 6 of include/amie_b200.h) against what the reference itself produced (make_golden_precond.py)."""
    g = np.load(BLOCKS)
    for name in ("tri", "rand"):
        S = ol.Sys(2, int(g[f"{name}_nb"]), g[f"{name}_row_size"], g[f"{name}_column_index"], g[f"{name}_array"], g[f"{name}_b"])
        assert same_bits(ol.oracle_precond_blocks(S), g[f"{name}_blocks"]), name
        ret, x, info = ol.oracle_cg(S, precond=5, nssor=32)
        assert ret == int(g[f"{name}_cg_ok"]) and info.nit == int(g[f"{name}_cg_nit"]), name
        assert same_bits(x, g[f"{name}_cg_x"]), name
    nb = int(g["hex_nb"])
    S3 = ol.Sys(3, nb, g["hex_row_size"], g["hex_column_index"], g["hex_array"], np.zeros(3 * nb))
    B = ol.oracle_precond_blocks(S3).reshape(-1, 9)
    ok = np.abs(g["det3"][:nb]) > 1e-8
    assert ok.any() and same_bits(B[ok], g["inv3"][:nb][ok])
    # where the determinant is tiny the block falls back to the inverse diagonal (1 where that is tiny too)
    for k in np.flatnonzero(~ok):
        d = g["m3"][k].reshape(3, 3).diagonal()
        want = np.diag(np.where(np.abs(d) > 1e-8, 1. / np.where(d == 0, 1, d), 1.))
        assert np.array_equal(B[k].reshape(3, 3), want)


def test_oracle_block_preconditioners_match_live_reference(ol):
    if ol.ref() is None:
        pytest.skip("oracle/_ref not built here: tests/golden/blockprecond.npz pins the oracle instead")
    for seed in (1, 2):
        rs, ci, arr, b = random_spd_blocks(2, 50, 70 + seed)
        S = ol.Sys(2, 50, rs, ci, arr, b)
        assert same_bits(ol.oracle_precond_blocks(S), ol.ref_precond_blocks2(S))
        r1 = ol.oracle_cg(S, precond=5)
        r2 = ol.ref_cg(S, precond=5)
        assert r1[0] == r2[0] and r1[2].nit == r2[2] and same_bits(r1[1], r2[1])
    rs, ci, arr, b = random_spd_blocks(3, 40, 77)
    S3 = ol.Sys(3, 40, rs, ci, arr, b)
    A = S3.to_scipy()
    m = np.stack([A[3 * k:3 * k + 3, 3 * k:3 * k + 3].toarray().ravel() for k in range(40)])
    det, inv = ol.ref_det_invert3x3(m)
    assert (np.abs(det) > 1e-8).all() and same_bits(ol.oracle_precond_blocks(S3).reshape(-1, 9), inv)
