"""Host half of the node renumbering (csrc/reorder.cpp): reverse Cuthill-McKee on the block graph and the structure in
the new numbering.  Checked on FeatureTree-assembled systems (the mesher's numbering has no locality) and on the
synthetic ones: valid permutation, structure preserved, bandwidth reduced, and -- through the oracle -- the solve of
the renumbered system is the renumbered solve."""
import os

import numpy as np
import pytest

from conftest import rel_l2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def renumbered(pkg, G):
    s, nb = int(G["stride"]), int(G["nb"])
    rs, ci = G["row_size"], G["column_index"]
    perm = pkg.rcm_order(rs, ci)
    rs2, ci2, frm = pkg.permute_structure(rs, ci, perm)
    return s, nb, rs, ci, perm, rs2, ci2, frm


@pytest.mark.parametrize("name", ["AMIE-3d-s400.npz", "AMIE-2d-s20.npz", "S3-tet-6.npz", "rand-s4.npz"])
def test_rcm_and_permuted_structure(pkg, ol, name):
    G = np.load(os.path.join(GOLDEN, name))
    s, nb, rs, ci, perm, rs2, ci2, frm = renumbered(pkg, G)
    assert np.array_equal(np.sort(perm), np.arange(nb))
    rows, rows2 = np.repeat(np.arange(nb), rs), np.repeat(np.arange(nb), rs2)
    # same blocks, renumbered: block k of the new structure is block frm[k] of the old one
    assert np.array_equal(np.sort(frm), np.arange(ci.size))
    assert np.array_equal(rows2, perm[rows[frm]]) and np.array_equal(ci2, perm[ci[frm]])
    # columns ascending inside every row (the reference binary-searches them)
    starts = np.concatenate([[0], np.cumsum(rs2, dtype=np.int64)])[:-1].astype(np.int64)
    inner = np.ones(ci2.size, bool)
    inner[starts[rs2 > 0]] = False
    assert (np.diff(ci2.astype(np.int64))[inner[1:]] > 0).all()
    if name.startswith("AMIE"):
        bw, bw2 = np.abs(ci.astype(np.int64) - rows).mean(), np.abs(ci2.astype(np.int64) - rows2).mean()
        assert bw2 < 0.5 * bw, (bw, bw2)
    # the solve of the renumbered system is the renumbered solve
    cl = s + s % 2
    arr2 = G["array"].reshape(-1, s * cl)[frm].reshape(-1)
    b2 = np.zeros_like(G["b"])
    b2.reshape(-1, s)[perm] = G["b"].reshape(-1, s)
    r1 = ol.oracle_cg(ol.Sys(s, nb, rs, ci, G["array"], G["b"]), nssor=32)
    r2 = ol.oracle_cg(ol.Sys(s, nb, rs2, ci2, arr2, b2), nssor=32)
    assert r1[0] == r2[0] and abs(int(r1[2].nit) - int(r2[2].nit)) <= 2
    assert rel_l2(r2[1].reshape(-1, s)[perm].reshape(-1), r1[1]) <= 1e-8


def test_permute_structure_rejects_a_non_permutation(pkg):
    G = np.load(os.path.join(GOLDEN, "S3-hex-6.npz"))
    perm = np.zeros(int(G["nb"]), np.uint32)
    with pytest.raises(pkg.AmieB200Error):
        pkg.permute_structure(G["row_size"], G["column_index"], perm)


def test_rcm_handles_disconnected_graphs_and_is_deterministic(pkg):
    # two disjoint chains and an isolated node
    rs = np.array([2, 3, 2, 1, 2, 3, 2], np.uint32)
    ci = np.array([0, 1, 0, 1, 2, 1, 2, 3, 4, 5, 4, 5, 6, 5, 6], np.uint32)
    p1, p2 = pkg.rcm_order(rs, ci), pkg.rcm_order(rs, ci)
    assert np.array_equal(p1, p2) and np.array_equal(np.sort(p1), np.arange(7))


def test_gather_of_a_device_slice_under_a_block_map(pkg):
    """amie_b200_set_block_map on a multi-device context (csrc/group.cu): device r holds the stored blocks
    [k0, k1) of the renumbered structure, which are blocks block_from[k0:k1] of the caller's array -- gathered on the
    host by amie_b200_gather_blocks before the upload.  Against numpy indexing, padded 3x3 and 2x2 blocks."""
    import ctypes
    from conftest import random_spd_blocks
    L = pkg.lib()
    L.amie_b200_gather_blocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    L.amie_b200_gather_blocks.restype = None
    for stride in (2, 3):
        rs, ci, arr, b = random_spd_blocks(stride, 60, 70 + stride)
        per_block = stride * (stride + stride % 2)
        perm = pkg.rcm_order(rs, ci)
        rs2, ci2, frm = pkg.permute_structure(rs, ci, perm)
        want = arr.reshape(-1, per_block)[frm]
        k0, k1 = frm.size // 3, frm.size - 5                       # some device's range of stored blocks
        src = np.ascontiguousarray(frm[k0:k1], np.uint32)
        out = np.zeros((k1 - k0) * per_block)
        L.amie_b200_gather_blocks(arr.ctypes.data, src.ctypes.data, k1 - k0, per_block, out.ctypes.data)
        assert np.array_equal(out.reshape(-1, per_block), want[k0:k1])


def tile_imbalance(row_size, rows_per_tile=10):
    """(rows per tile x longest row, summed over the tiles) / stored blocks: what a lane-per-row tile pays for unequal rows."""
    rs = np.asarray(row_size, np.int64)
    pad = (-rs.size) % rows_per_tile
    t = np.concatenate([rs, np.zeros(pad, np.int64)]).reshape(-1, rows_per_tile)
    return float((t.max(1) * rows_per_tile).sum() / rs.sum())


@pytest.mark.parametrize("name", ["AMIE-3d-s400.npz", "AMIE-2d-s20.npz"])
def test_rows_grouped_by_length_inside_windows(pkg, ol, name):
    """amie_b200_group_rows_by_length on top of the Cuthill-McKee numbering of the FeatureTree fixtures: still a permutation,
    nodes stay inside their window, the per-tile imbalance drops, and the solve of the renumbered system is the
    renumbered solve (oracle)."""
    G = np.load(os.path.join(GOLDEN, name))
    s, nb = int(G["stride"]), int(G["nb"])
    rs, ci = G["row_size"], G["column_index"]
    perm = pkg.rcm_order(rs, ci)
    W = 80
    perm2 = pkg.group_rows_by_length(rs, perm, W)
    assert np.array_equal(np.sort(perm2), np.arange(nb))
    assert np.array_equal(perm // W, perm2 // W)                       # nobody left its window
    rs1, ci1, _ = pkg.permute_structure(rs, ci, perm)
    rs2, ci2, frm = pkg.permute_structure(rs, ci, perm2)
    for w0 in range(0, nb, W):                                         # longest first inside each window
        assert np.all(np.diff(rs2[w0:w0 + W].astype(np.int64)) <= 0)
    rows = 10 if s == 3 else 16
    before, after = tile_imbalance(rs1, rows), tile_imbalance(rs2, rows)
    print(f"{name}: tile imbalance {tile_imbalance(rs, rows):.2f} (mesher) -> {before:.2f} (RCM) -> {after:.2f} (RCM + window {W})")
    assert after < before
    # the renumbered system solves to the renumbered solution
    cl = s + s % 2
    arr2 = G["array"].reshape(-1, s * cl)[frm].reshape(-1)
    b2 = np.empty_like(G["b"])
    b2.reshape(-1, s)[perm2] = G["b"].reshape(-1, s)
    S2 = ol.Sys(s, nb, rs2, ci2, arr2, b2)
    ret, x2, info = ol.oracle_cg(S2, nssor=32)
    assert ret == bool(G["cg_ok"]) and abs(int(info.nit) - int(G["cg_nit"])) <= 2
    x_back = x2.reshape(-1, s)[perm2].reshape(-1)
    assert np.linalg.norm(x_back - G["cg_x"]) <= 1e-8 * np.linalg.norm(G["cg_x"])
    # error paths
    with pytest.raises(pkg.AmieB200Error):
        pkg.group_rows_by_length(rs, np.zeros(nb, np.uint32), W)
    with pytest.raises(pkg.AmieB200Error):
        pkg.group_rows_by_length(rs, perm, 0)
