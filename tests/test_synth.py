"""Synthetic structured elastic systems (host generator) and the row partitioner (host only)."""
import numpy as np
import pytest


@pytest.mark.parametrize("preset,n,bpr", [("S3-hex", 8, 27), ("S3-tet", 8, 15), ("S2-tri", 16, 7), ("ASR-hex", 8, 27)])
def test_structure_and_symmetry(pkg, ol, systems, preset, n, bpr):
    S = systems(preset, n)
    assert S.row_size.max() == bpr
    assert S.row_size.sum() == S.nnzb
    ptr = np.concatenate([[0], np.cumsum(S.row_size, dtype=np.int64)]).astype(np.int64)
    for r in range(S.nb):                       # sorted ascending, diagonal present (assembly.cpp:503-521)
        c = S.column_index[ptr[r]:ptr[r + 1]]
        assert np.all(np.diff(c.astype(np.int64)) > 0)
        assert r in c
    A = S.to_scipy()
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
    # pad slots of stride-3 blocks are zero (sparse/sparse_matrix.h layout)
    if S.stride == 3:
        assert not S.array.reshape(-1, 3, 4)[:, :, 3].any()
    # SPD after Dirichlet elimination
    import scipy.sparse.linalg as sla
    lam = sla.eigsh(A, k=1, which="SA", return_eigenvectors=False, tol=1e-6)
    assert lam[0] > 0


def test_row_ranges_concatenate(pkg):
    syn = pkg.Synth("S3-tet", 6)
    rs, ci, arr, b = syn.rows()
    cut = syn.nb // 3
    rs0, ci0, arr0, b0 = syn.rows(0, cut)
    rs1, ci1, arr1, b1 = syn.rows(cut, syn.nb)
    assert np.array_equal(np.concatenate([rs0, rs1]), rs)
    assert np.array_equal(np.concatenate([ci0, ci1]), ci)
    assert np.array_equal(np.concatenate([arr0, arr1]), arr)
    assert np.array_equal(np.concatenate([b0, b1]), b)


def test_physical_sanity_uniaxial(pkg, ol):
    """Homogeneous bar under unit traction: u_x(L) = sigma L / E exactly for Q1 elements."""
    syn = pkg.Synth("S3-hex", 6)
    rs, ci, arr, b = syn.rows()
    S = ol.Sys(3, syn.nb, rs, ci, arr, b)
    ret, x, _ = ol.oracle_cg(S)
    assert ret == 1
    ux = x[0::3].reshape(6, 6, 6)
    assert ux[:, :, 0].max() == 0.0            # symmetry plane
    assert 0.1 < ux[:, :, -1].mean() < 1.0     # E=1 matrix with a stiff E=10 sphere: less than 1/E


@pytest.mark.parametrize("nparts", [1, 2, 3, 8])
def test_partition_balanced_and_halo(pkg, systems, nparts):
    S = systems("S3-hex", 8)
    bounds = pkg.partition_rows(S.row_size, nparts)
    assert bounds[0] == 0 and bounds[-1] == S.nb and np.all(np.diff(bounds.astype(np.int64)) >= 0)
    ptr = np.concatenate([[0], np.cumsum(S.row_size)]).astype(np.int64)
    loads = [ptr[int(bounds[p + 1])] - ptr[int(bounds[p])] for p in range(nparts)]
    assert max(loads) - min(loads) <= 2 * 27 + S.nnzb // (50 * nparts)
    for p in range(nparts):
        r0, r1 = int(bounds[p]), int(bounds[p + 1])
        cols = S.column_index[ptr[r0]:ptr[r1]]
        halo = pkg.partition_halo(r0, r1, S.row_size[r0:r1], cols)
        expect = np.unique(cols[(cols < r0) | (cols >= r1)])
        assert np.array_equal(halo, expect)
        if nparts == 1:
            assert halo.size == 0
        else:
            # slabs along the slowest index: the halo is at most one node plane per side
            assert halo.size <= 2 * 8 * 8 + 2 * 8 + 2
