"""Opt-in kernel variants of the assembly row (options "assemble_variant" = 2 | 3, "dirichlet_variant" = 1) must give the
bits of the default kernels, which the parity tests pin (their logic is also checked on the CPU by
tests/test_emu_kernels.py)."""
import numpy as np
import pytest

from conftest import random_spd_blocks
from test_gpu_assembly import device_array, grid_elements, load

pytestmark = pytest.mark.gpu


def variant_assembly(pkg, stride, rs, ci, variant=2):
    asm = pkg.Assembly(None, None, device=0)
    asm.set_option("assemble_variant", variant)
    asm.set_option("dirichlet_variant", 1)
    asm.set_structure_only(stride, rs, ci)
    return asm


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("name", ["AMIE-2d-s20-assembly.npz", "AMIE-3d-s400-assembly.npz"])
def test_variants_reproduce_featuretree_matrix(pkg, ol, name, variant):
    G = load(name)
    s, nb = int(G["stride"]), int(G["nb"])
    el = ol.Elements(s, G["elem_ids"], G["elem_ke"], G["scales"])
    asm = variant_assembly(pkg, s, G["row_size"], G["column_index"], variant)
    asm.set_elements(el.ids)
    asm.update_elements(0, el.ke, el.scales)
    asm.assemble()
    assert np.array_equal(device_array(asm), ol.oracle_assemble(s, nb, G["row_size"], G["column_index"], el))
    asm.upload_rhs(np.zeros(nb * s))
    asm.set_boundary_conditions(G["fix_ids"], G["fix_values"])
    assert np.array_equal(device_array(asm), G["array_post"])
    if bool(G["forces_comparable"]):
        assert np.array_equal(asm.download_rhs(), G["forces_post"])
    asm.close()


@pytest.mark.parametrize("dims,stride,ragged", [((9, 8), 2, False), ((7, 6, 5), 3, True), ((12, 11), 1, False),
                                                ((5, 4, 4), 4, False), ((4, 4, 3), 6, True), ((23, 19, 17), 3, False)])
@pytest.mark.parametrize("variant", [2, 3])
def test_assemble_variant_matches_oracle_incl_incremental(pkg, ol, dims, stride, ragged, variant):
    nb, el = grid_elements(ol, dims, stride, seed=sum(dims) + stride, ragged=ragged)
    rs, ci = el.pattern(nb)
    asm = variant_assembly(pkg, stride, rs, ci, variant)
    asm.set_elements(el.ids)
    asm.update_elements(0, el.ke, el.scales)
    asm.assemble()
    assert np.array_equal(device_array(asm), ol.oracle_assemble(stride, nb, rs, ci, el))
    rng = np.random.default_rng(1)
    first, count = el.n_elem // 3, max(1, el.n_elem // 5)
    el.ke[first:first + count] *= rng.uniform(0.1, 0.9, (count, 1, 1, 1))
    asm.update_elements(first, el.ke[first:first + count], el.scales[first:first + count])
    asm.assemble()
    assert np.array_equal(device_array(asm), ol.oracle_assemble(stride, nb, rs, ci, el))
    asm.close()


@pytest.mark.parametrize("stride", [1, 2, 3, 4, 6])
def test_dirichlet_variant_matches_oracle(pkg, ol, stride):
    nb = 70
    rs, ci, arr, b = random_spd_blocks(stride, nb, 500 + stride)
    n = nb * stride
    rng = np.random.default_rng(stride)
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(rs, ci, stride, arr), b, device=0)
    asm.set_option("dirichlet_variant", 1)
    for nfix in (0, 1, n // 4, n):
        fix = np.sort(rng.choice(n, nfix, replace=False)).astype(np.uint32)
        fv = rng.standard_normal(nfix)
        rest = np.setdiff1d(np.arange(n), fix)
        frc = np.sort(rng.choice(rest, min(9, rest.size), replace=False)).astype(np.uint32)
        frv = rng.standard_normal(frc.size)
        nat, add = rng.standard_normal(n), rng.standard_normal(n)
        a0, f0, n0, _ = ol.oracle_set_bcs(stride, nb, rs, ci, arr, b, fix, fv, frc, frv, nat, add)
        asm.values_changed()
        asm.sync_matrix()
        asm.upload_rhs(b)
        asm.set_boundary_conditions(fix, fv, frc, frv, add, nat)
        assert np.array_equal(device_array(asm), a0)
        assert np.array_equal(asm.download_rhs(), f0)
        assert np.array_equal(nat, n0)
    asm.close()


@pytest.mark.parametrize("preset,n,graph", [("S3-hex", 12, 1), ("S3-hex", 20, 0), ("S2-tri", 40, 1), ("S3-tet", 14, 0)])
def test_pcg_with_split_dot_matches_oracle(pkg, ol, systems, preset, n, graph):
    """Option "split_dot": q = A p in the plain form and p.q as its own pass.  Same algorithm, another summation order
    of p.q: iteration counts within +-2 of the reference, x within 1e-8 -- with and without CUDA-graph batches, with
    rowstart, and switching back to the fused form on the same context."""
    from conftest import rel_l2
    S = systems(preset, n)
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(S.row_size, S.column_index, S.stride, S.array), S.b, device=0)
    asm.set_option("graph", graph)
    for split in (1, 0, 1):
        asm.set_option("split_dot", split)
        ret, x_ref, info = ol.oracle_cg(S, nssor=32)
        cg = pkg.ConjugateGradient(asm)
        cg.nssor = 32
        assert cg.solve(None, None, 1e-10, -1) == bool(ret)
        assert abs(int(cg.nit) - int(info.nit)) <= 2, (split, cg.nit, info.nit)
        assert rel_l2(cg.x, x_ref) <= 1e-8, split
    rs = (S.n // 4) // S.stride * S.stride
    ret, x_ref, info = ol.oracle_cg(S, nssor=32, rowstart=rs, colstart=rs)
    cg = pkg.ConjugateGradient(asm)
    cg.nssor = 32
    cg.rowstart = cg.colstart = rs
    assert cg.solve(None, None, 1e-10, -1) == bool(ret)
    assert abs(int(cg.nit) - int(info.nit)) <= 2 and rel_l2(cg.x, x_ref) <= 1e-8
    asm.close()


def test_fields_variant_matches_reference_bits(pkg, ol):
    """Option "fields_variant" = 1: the unrolled, phase-split field kernel for linear triangles / tetrahedra."""
    import glob
    import os
    from test_gpu_recovery import context_for, same_bits
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*-fields.npz"))):
        g = np.load(path)
        asm, dim = context_for(pkg, g)
        asm.set_option("fields_variant", 1)
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
