"""The kernel SOURCES of the rows next to the solve (csrc/kernels_setup.cuh, kernels_assemble.cuh, kernels_fields.cuh,
kernels_history.cuh), compiled for the host by the test-only emulation (tests/emu/cuda_emu.h) and run in the launch
sequences of their C-ABI functions, against the oracle and the reference-made fixtures: bit for bit.  This checks
index logic and arithmetic order where no GPU is available; the same kernels run on the device in tests/test_gpu_*.py."""
import glob
import os

import numpy as np
import pytest

import emu_lib as em
from conftest import random_spd_blocks
from test_gpu_assembly import grid_elements

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.mark.parametrize("stride", [1, 2, 3, 4, 6])
def test_repack_and_preconditioner_diagonals(ol, stride):
    rs, ci, arr, b = random_spd_blocks(stride, 70, 60 + stride)
    S = ol.Sys(stride, 70, rs, ci, arr, b)
    vals = em.compact(arr, stride)
    cl = stride + stride % 2
    assert same_bits(vals, arr.reshape(-1, stride, cl)[:, :, :stride].reshape(-1))
    for kind in (0, 2, 3):
        assert same_bits(em.precond_diagonal(kind, stride, rs, ci, vals), ol.oracle_precond_diagonal(S, kind)), kind


@pytest.mark.parametrize("stride", [1, 2, 3, 4, 6])
def test_repack_through_a_block_map(pkg, stride):
    """set_values on a renumbered structure: the values of the caller's array land where the renumbered structure
    stores them (csrc/reorder.cpp gives block_from; block_to is its inverse)."""
    rs, ci, arr, b = random_spd_blocks(stride, 50, 90 + stride)
    perm = pkg.rcm_order(rs, ci)
    rs2, ci2, frm = pkg.permute_structure(rs, ci, perm)
    block_to = np.empty_like(frm)
    block_to[frm] = np.arange(frm.size, dtype=np.uint32)
    cl = stride + stride % 2
    want = em.compact(arr.reshape(-1, stride * cl)[frm].reshape(-1), stride)
    assert same_bits(em.compact_scatter(arr, stride, block_to), want)


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "precond-*.npz"))))
def test_preconditioner_diagonals_against_reference_fixtures(name):
    g = np.load(os.path.join(GOLDEN, name))
    s = int(g["stride"])
    vals = em.compact(g["array"], s)
    for kind in (0, 2, 3):
        assert same_bits(em.precond_diagonal(kind, s, g["row_size"], g["column_index"], vals), g[f"diag{kind}"]), kind


@pytest.mark.parametrize("dims,stride,ragged", [((9, 8), 2, False), ((6, 5, 7), 3, False), ((7, 6, 5), 3, True),
                                                ((12, 11), 1, False), ((5, 4, 4), 4, False), ((4, 4, 3), 6, True)])
def test_assemble_kernels_match_oracle(ol, dims, stride, ragged):
    nb, el = grid_elements(ol, dims, stride, seed=sum(dims) + stride, ragged=ragged)
    rs, ci = el.pattern(nb)
    rc, vals = em.assemble(stride, rs, ci, el.ids, el.ke, el.scales)
    assert rc == 0
    assert same_bits(em.padded(vals, stride), ol.oracle_assemble(stride, nb, rs, ci, el))
    # a damage-like step: some elements change, only their stored blocks are re-accumulated
    rng = np.random.default_rng(1)
    first, count = el.n_elem // 3, max(1, el.n_elem // 5)
    el.ke[first:first + count] *= rng.uniform(0.1, 0.9, (count, 1, 1, 1))
    rc, vals2 = em.assemble(stride, rs, ci, el.ids, el.ke, el.scales, vals=vals, mark=(first, count))
    assert rc == 0
    assert same_bits(em.padded(vals2, stride), ol.oracle_assemble(stride, nb, rs, ci, el))
    # error paths of the map build
    bad = el.ids.copy()
    bad[0, 0] = nb + 5
    assert em.assemble(stride, rs, ci, bad, el.ke, el.scales)[0] == 1
    if el.n_elem > 2:
        far = el.ids.copy()
        far[0, 0], far[0, 1] = el.ids[0, 0], el.ids[-1, -1]
        rc, _ = em.assemble(stride, rs, ci, far, el.ke, el.scales)
        assert rc in (0, 2)          # 2 when that pair is not in the pattern


@pytest.mark.parametrize("name", ["AMIE-2d-s20-assembly.npz", "AMIE-3d-s400-assembly.npz"])
def test_assemble_and_eliminate_reproduce_the_featuretree_matrix(ol, name):
    G = np.load(os.path.join(GOLDEN, name))
    s, nb = int(G["stride"]), int(G["nb"])
    el = ol.Elements(s, G["elem_ids"], G["elem_ke"], G["scales"])
    rc, vals = em.assemble(s, G["row_size"], G["column_index"], el.ids, el.ke, el.scales)
    assert rc == 0
    vals, forces, _, dirty = em.dirichlet(s, G["row_size"], G["column_index"], vals, np.zeros(nb * s), G["fix_ids"], G["fix_values"])
    assert same_bits(em.padded(vals, s), G["array_post"])
    if bool(G["forces_comparable"]):
        assert same_bits(forces, G["forces_post"])
    assert dirty.any()


@pytest.mark.parametrize("name,cuts", [("AMIE-2d-s20-assembly.npz", (0.5,)), ("AMIE-3d-s400-assembly.npz", (0.3, 0.55, 0.9))])
def test_partitioned_assembly_and_elimination_reproduce_the_featuretree_matrix(ol, name, cuts):
    """The same kernels on the parts of a row-partitioned matrix (the numbering of csrc/dist.cu: owned columns, then the
    halo; a row's blocks still in global order): every part sees the whole element list and the global id lists of the
    boundary conditions, builds its own gather map for the rows it owns, and the parts put together give the FeatureTree
    matrix and force vector bit for bit."""
    G = np.load(os.path.join(GOLDEN, name))
    s, nb = int(G["stride"]), int(G["nb"])
    el = ol.Elements(s, G["elem_ids"], G["elem_ke"], G["scales"])
    bounds = [0] + [int(c * nb) for c in cuts] + [nb]
    parts = em.split_rows(G["row_size"], G["column_index"], bounds)
    assert any(p[6].size for p in parts)
    vals_all, forces_all, dirty_any = [], [], False
    for part in parts:
        r0, r1 = part[0], part[1]
        rc, vals = em.assemble_part(s, part, nb, el.ids, el.ke, el.scales)
        assert rc == 0
        vals, forces, _, dirty = em.dirichlet_part(s, part, nb, vals, np.zeros((r1 - r0) * s), G["fix_ids"], G["fix_values"])
        vals_all.append(vals)
        forces_all.append(forces)
        dirty_any |= bool(dirty.any())
    assert same_bits(em.padded(np.concatenate(vals_all), s), G["array_post"])
    if bool(G["forces_comparable"]):
        assert same_bits(np.concatenate(forces_all), G["forces_post"])
    assert dirty_any
    # a node id beyond the GLOBAL matrix is an error on every part; a pair outside the pattern on the part that owns the row
    bad = el.ids.copy()
    bad[0, 0] = nb + 3
    assert all(em.assemble_part(s, part, nb, bad, el.ke, el.scales)[0] == 1 for part in parts)


@pytest.mark.parametrize("stride", [2, 3])
def test_partitioned_dirichlet_kernel_matches_oracle(ol, stride):
    """All of set_boundary_conditions on three parts: imposed displacements, imposed forces, the natural-condition vector
    and the additional forces, ids global, vectors sliced."""
    nb = 90
    rs, ci, arr, b = random_spd_blocks(stride, nb, 900 + stride)
    n = nb * stride
    rng = np.random.default_rng(40 + stride)
    vals = em.compact(arr, stride)
    parts = em.split_rows(rs, ci, [0, 31, 58, nb])
    for nfix in (1, n // 4, n):
        fix = np.sort(rng.choice(n, nfix, replace=False)).astype(np.uint32)
        fv = rng.standard_normal(nfix)
        rest = np.setdiff1d(np.arange(n), fix)
        frc = np.sort(rng.choice(rest, min(9, rest.size), replace=False)).astype(np.uint32)
        frv = rng.standard_normal(frc.size)
        nat, add = rng.standard_normal(n), rng.standard_normal(n)
        a0, f0, n0, _ = ol.oracle_set_bcs(stride, nb, rs, ci, arr, b, fix, fv, frc, frv, nat, add)
        got_v, got_f, got_n = [], [], []
        for part in parts:
            r0, r1, k0, k1 = part[:4]
            sl = slice(r0 * stride, r1 * stride)
            v1, f1, n1, _ = em.dirichlet_part(stride, part, nb, vals[k0 * stride * stride:k1 * stride * stride], b[sl], fix, fv, frc, frv,
                                              nat[sl], add[sl])
            got_v.append(v1) ; got_f.append(f1) ; got_n.append(n1)
        assert same_bits(em.padded(np.concatenate(got_v), stride), a0), nfix
        assert same_bits(np.concatenate(got_f), f0) and same_bits(np.concatenate(got_n), n0), nfix


@pytest.mark.parametrize("stride", [1, 2, 3, 4, 6])
def test_dirichlet_kernel_matches_oracle(ol, stride):
    nb = 70
    rs, ci, arr, b = random_spd_blocks(stride, nb, 500 + stride)
    n = nb * stride
    rng = np.random.default_rng(stride)
    vals = em.compact(arr, stride)
    for nfix in (0, 1, n // 4, n):
        fix = np.sort(rng.choice(n, nfix, replace=False)).astype(np.uint32)
        fv = rng.standard_normal(nfix)
        rest = np.setdiff1d(np.arange(n), fix)
        frc = np.sort(rng.choice(rest, min(9, rest.size), replace=False)).astype(np.uint32)
        frv = rng.standard_normal(frc.size)
        nat, add = rng.standard_normal(n), rng.standard_normal(n)
        a0, f0, n0, _ = ol.oracle_set_bcs(stride, nb, rs, ci, arr, b, fix, fv, frc, frv, nat, add)
        v1, f1, n1, _ = em.dirichlet(stride, rs, ci, vals, b, fix, fv, frc, frv, nat, add)
        assert same_bits(em.padded(v1, stride), a0) and same_bits(f1, f0) and same_bits(n1, n0), nfix


@pytest.mark.parametrize("stride", [2, 3])
def test_dirichlet_kernel_against_reference_fixture(stride):
    G = np.load(os.path.join(GOLDEN, f"bc-rand-s{stride}.npz"))
    vals = em.compact(G["array"], stride)
    v1, f1, n1, _ = em.dirichlet(stride, G["row_size"], G["column_index"], vals, G["forces"], G["fix_ids"], G["fix_values"],
                                 G["force_ids"], G["force_values"], G["natural"], G["add_to_forces"])
    assert same_bits(em.padded(v1, stride), G["array_post"])
    assert same_bits(f1, G["forces_post"]) and same_bits(n1, G["natural_post"])


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*-fields.npz"))))
def test_field_kernel_against_reference_fixtures(ol, name, variant):
    g = np.load(os.path.join(GOLDEN, name))
    dim = int(g["dim"])
    tot, mech, sig = em.element_fields(dim, g["ids"], g["dshape"], g["jinv"], g["u"], g["tensors"], g["imposed_strain"],
                                       g["imposed_stress"], g["tensor_of_elem"], variant=variant)
    assert same_bits(tot, g["total_strain"]) and same_bits(mech, g["mechanical_strain"]) and same_bits(sig, g["real_stress"])
    # per-element behaviours, an unused slot, a vector shorter than the dofs referenced
    ne = g["ids"].shape[0]
    nc = tot.shape[1]
    rng = np.random.default_rng(5)
    toe = g["tensor_of_elem"]
    C = g["tensors"][toe] * rng.uniform(0.2, 1.0, (ne, 1, 1))
    es, ss = 1e-4 * rng.standard_normal((ne, nc)), rng.standard_normal((ne, nc))
    ids5 = np.concatenate([g["ids"], np.full((ne, 1), 0xFFFFFFFF, np.uint32)], axis=1)
    ds5 = np.concatenate([g["dshape"], np.full((ne, 1, dim), 7.0)], axis=1)
    u = g["u"][:g["u"].size // 2 + 1]
    got = em.element_fields(dim, ids5, ds5, g["jinv"], u, C, es, ss, None, variant=variant)
    want = ol.oracle_element_fields(dim, ids5, ds5, g["jinv"], u, C, es, ss, None)
    for a, b in zip(got, want):
        assert same_bits(a, b)
    # unused slots INSIDE the element's own slots, with garbage (NaN) derivatives there: they must not reach the sums
    idsr, dsr = g["ids"].copy(), g["dshape"].copy()
    hole = rng.integers(0, idsr.shape[1], ne)
    pick = rng.random(ne) < 0.3
    idsr[pick, hole[pick]] = 0xFFFFFFFF
    dsr[pick, hole[pick], :] = np.nan
    got = em.element_fields(dim, idsr, dsr, g["jinv"], u, C, es, ss, None, variant=variant)
    want = ol.oracle_element_fields(dim, idsr, dsr, g["jinv"], u, C, es, ss, None)
    for a, b in zip(got, want):
        assert same_bits(a, b) and not np.isnan(a).any()


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*-fields.npz"))))
def test_principal_kernel_against_reference_fixtures(ol, name):
    """On the host the kernel source calls the same C library as the reference: bit for bit in 2D and in 3D."""
    g = np.load(os.path.join(GOLDEN, name))
    dim = int(g["dim"])
    for src, dst, dbl in (("total_strain", "principal_total_strain", 1), ("mechanical_strain", "principal_mechanical_strain", 1),
                          ("real_stress", "principal_real_stress", 0)):
        assert same_bits(em.element_principal(dim, g[src], dbl), g[dst]), dst
    rng = np.random.default_rng(2)
    v = rng.standard_normal((777, 3 if dim == 2 else 6)) * 10.0 ** rng.integers(-3, 4, (777, 1))
    for dbl in (0, 1):
        assert same_bits(em.element_principal(dim, v, dbl), ol.oracle_principal(dim, v, dbl))


def test_history_kernels_match_oracle(ol):
    rng = np.random.default_rng(0)
    n = 1234
    prev, back = rng.standard_normal(n), rng.standard_normal(n)
    back[5] = np.nan
    back[9] = prev[9] = -0.0
    for factor in (1.0, 0.37, -1.5, 0.0):
        x, scrubbed = em.extrapolate(prev, back, factor)
        want, wback = ol.oracle_extrapolate(prev, back, factor)
        assert same_bits(x, want) and same_bits(scrubbed, wback), factor
    v = np.array([1.5, -2.0, 0.0, -0.0, np.inf])
    with np.errstate(invalid="ignore"):
        assert same_bits(em.times_zero(v), v * 0.)
