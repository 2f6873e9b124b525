"""The oracle (oracle/amie_oracle.c) against the reference's own outputs.

* golden vectors in tests/golden/*.npz were produced by the real reference (make_golden.py);
* where oracle/_ref/libamie_ref_oracle.so is present the oracle is also checked live against it.
The restatement keeps the reference's arithmetic order, so both checks are BIT-EXACT at 1 thread.
"""
import glob
import os

import numpy as np
import pytest

# solver fixtures only: *-assembly.npz / bc-rand-*.npz / *-fields.npz / precond-*.npz belong to the assembly, field-recovery and preconditioner tests
GOLDEN = [p for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
          if not (p.endswith(("-assembly.npz", "-fields.npz")) or os.path.basename(p).startswith(("bc-rand-", "precond-", "blockprecond")))]


def _sys(ol, g):
    return ol.Sys(int(g["stride"]), int(g["nb"]), g["row_size"], g["column_index"], g["array"], g["b"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(ol, path):
    g = np.load(path)
    S = _sys(ol, g)
    v = g["v"]
    rs = int(g["rowstart"])
    assert np.array_equal(ol.oracle_assign(S, v), g["assign"])
    assert np.array_equal(ol.oracle_assign(S, v, S.b), g["assign_minus_b"])
    assert np.array_equal(ol.oracle_assign(S, v, S.b, rs, rs), g["assign_minus_b_rowstart"])
    assert np.array_equal(ol.oracle_spmv_serial(S, v), g["serial"])
    assert np.array_equal(ol.oracle_spmv_serial(S, v, S.b), g["serial_minus_b"])
    assert np.array_equal(ol.oracle_inverse_diagonal(S), g["inverse_diagonal"])
    ret, x, info = ol.oracle_cg(S, nssor=32)
    assert (ret, info.nit) == (int(g["cg_ok"]), int(g["cg_nit"]))
    assert np.array_equal(x, g["cg_x"])
    ret, x, info = ol.oracle_cg(S, nssor=32, rowstart=rs, colstart=rs)
    assert (ret, info.nit) == (int(g["cg_rs_ok"]), int(g["cg_rs_nit"]))
    assert np.array_equal(x, g["cg_rs_x"])
    ret, x, info = ol.oracle_cg(S, nssor=0, x0=0.5 * g["cg_x"])
    assert (ret, info.nit) == (int(g["cg_warm_ok"]), int(g["cg_warm_nit"]))
    assert np.array_equal(x, g["cg_warm_x"])
    ret, x, info = ol.oracle_bicgstab(S)
    assert (ret, info.nit) == (int(g["bicg_ok"]), int(g["bicg_nit"]))
    assert np.array_equal(x, g["bicg_x"])


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("preset,n", [("S3-hex", 9), ("S3-tet", 8), ("S2-tri", 20), ("ASR-hex", 8)])
def test_oracle_matches_live_reference(ol, systems, preset, n):
    if ol.ref() is None:
        pytest.skip("oracle/_ref not built here (no /root/reference): golden vectors pin the oracle instead")
    S = systems(preset, n)
    v = np.random.default_rng(3).standard_normal(S.n)
    assert np.array_equal(ol.ref_spmv(S, v, None, mode=0)[0], ol.oracle_assign(S, v))
    assert np.array_equal(ol.ref_spmv(S, v, S.b, mode=3)[0], ol.oracle_spmv_serial(S, v, S.b))
    assert np.array_equal(ol.ref_inverse_diagonal(S), ol.oracle_inverse_diagonal(S))
    for kw in (dict(nssor=32), dict(nssor=128), dict(nssor=0, eps=1e-6), dict(nssor=32, precond=1, eps=1e-4)):
        r_ok, r_x, r_nit, _, _ = ol.ref_cg(S, nthreads=1, **kw)
        o_ok, o_x, o_info = ol.oracle_cg(S, **kw)
        assert (r_ok, r_nit) == (o_ok, o_info.nit), kw
        assert np.array_equal(r_x, o_x), kw
    r_ok, r_x, r_nit, _, _ = ol.ref_bicgstab(S, nthreads=1)
    o_ok, o_x, o_info = ol.oracle_bicgstab(S)
    assert (r_ok, r_nit) == (o_ok, o_info.nit)
    assert np.array_equal(r_x, o_x)


@pytest.mark.parametrize("stride", [1, 4, 6])
def test_oracle_other_strides_match_live_reference(ol, stride):
    """inner_product's stride 1 / 4 / 6 cases (sparse/sparse_matrix.h:222-233, :335-676), including the
    stride-4 compensation quirk (row 3 writes compensate2)."""
    if ol.ref() is None:
        pytest.skip("oracle/_ref not built here: tests/golden/rand-s*.npz pin these strides instead")
    from conftest import random_spd_blocks
    rs, ci, arr, b = random_spd_blocks(stride, 300, 20 + stride)
    S = ol.Sys(stride, 300, rs, ci, arr, b)
    v = np.random.default_rng(2).standard_normal(S.n)
    assert np.array_equal(ol.ref_spmv(S, v, b, mode=1)[0], ol.oracle_assign(S, v, b))
    assert np.array_equal(ol.ref_spmv(S, v, b, mode=3)[0], ol.oracle_spmv_serial(S, v, b))
    assert np.array_equal(ol.ref_inverse_diagonal(S), ol.oracle_inverse_diagonal(S))
    r_ok, r_x, r_nit, _, _ = ol.ref_cg(S, nssor=32, nthreads=1)
    o_ok, o_x, o_info = ol.oracle_cg(S, nssor=32)
    assert (r_ok, r_nit) == (o_ok, o_info.nit) and np.array_equal(r_x, o_x)
    r_ok, r_x, r_nit, _, _ = ol.ref_bicgstab(S, nthreads=1)
    o_ok, o_x, o_info = ol.oracle_bicgstab(S)
    assert (r_ok, r_nit) == (o_ok, o_info.nit) and np.array_equal(r_x, o_x)


def test_oracle_thread_emulation_close_to_reference(ol, systems):
    """OpenMP reductions reorder sums: the reference's own iteration count moves by +-1-2 with the
    thread count.  The oracle's static-chunk emulation stays inside that band."""
    if ol.ref() is None:
        pytest.skip("needs oracle/_ref")
    S = systems("S3-hex", 9)
    r_ok, r_x, r_nit, _, _ = ol.ref_cg(S, nthreads=4, nssor=32)
    o_ok, o_x, o_info = ol.oracle_cg(S, nthreads=4, nssor=32)
    assert r_ok == o_ok == 1
    assert abs(int(r_nit) - int(o_info.nit)) <= 2
    assert np.linalg.norm(r_x - o_x) / np.linalg.norm(r_x) < 1e-8


def test_oracle_edge_cases(ol, systems):
    S = systems("S3-hex", 5)
    # homogeneous right-hand side: returns true and leaves x = 0 (conjugategradient.cpp:74-78)
    ret, x, info = ol.oracle_cg(S, b=np.zeros(S.n))
    assert ret == 1 and info.status == 1 and not x.any()
    # exact initial guess: converges at nit = 0 through the err0 < realeps exit (:170-175)
    _, xs, _ = ol.oracle_cg(S)
    ret, x, info = ol.oracle_cg(S, x0=xs, nssor=0)
    assert ret == 1 and info.nit <= 1
    # short x0 is copied as a prefix (:99-104)
    ret, x, info = ol.oracle_cg(S, x0=xs[:7])
    assert ret == 1
    # NaN in the matrix: the reference exit(0)s; the oracle reports -4
    S2 = ol.Sys(S.stride, S.nb, S.row_size, S.column_index, S.array.copy(), S.b)
    S2.array[5] = np.nan
    ret, _, info = ol.oracle_cg(S2, nssor=0)
    assert ret == -4
