"""Device-side value assembly + Dirichlet elimination against the oracle
(oracle/amie_oracle_assembly.c) and the reference-made golden fixtures, through the C-ABI.  Bar: BIT-EXACT
(the kernels replay the reference's element-order compensated sums; see csrc/assemble.cu)."""
import os

import numpy as np
import pytest

from conftest import rel_l2, random_spd_blocks

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def grid_elements(ol, dims, stride, seed, ragged=False):
    """Structured grid of 2^d-node cells with random (non-symmetric) elementary matrices of wildly different scales,
    in shuffled element order, node numbering permuted: nothing here is favourable to a reordered sum."""
    rng = np.random.default_rng(seed)
    d = len(dims)
    nn = int(np.prod(dims))
    node = rng.permutation(nn).reshape(dims)
    cells = []
    for idx in np.ndindex(*[n - 1 for n in dims]):
        corner = [node[tuple(np.add(idx, off))] for off in np.ndindex(*([2] * d))]
        cells.append(corner)
    ids = np.array(cells, np.uint32)
    ids = ids[rng.permutation(ids.shape[0])]
    npe = ids.shape[1]
    if ragged:                                   # drop the last node slot of every third element
        ids[::3, npe - 1] = 0xFFFFFFFF
    ke = rng.standard_normal((ids.shape[0], npe, npe, stride * stride)) * 10.0 ** rng.integers(-6, 7, (ids.shape[0], 1, 1, 1))
    scales = rng.uniform(0.5, 2.0, ids.shape[0])
    return nn, ol.Elements(stride, ids, ke, scales)


def device_assembly(pkg, stride, row_size, column_index, el):
    asm = pkg.Assembly(None, None, device=0)
    asm.set_structure_only(stride, row_size, column_index)
    asm.set_elements(el.ids)
    asm.update_elements(0, el.ke, el.scales)
    asm.assemble()
    return asm


def device_array(asm):
    return asm.download_matrix()[2]


@pytest.mark.parametrize("name", ["AMIE-2d-s20-assembly.npz", "AMIE-3d-s400-assembly.npz"])
def test_featuretree_matrix_reproduced_bit_for_bit(pkg, ol, name):
    """Elements dumped from the unmodified FeatureTree -> device assembly + elimination == the matrix AMIE solved."""
    G = load(name)
    s, nb = int(G["stride"]), int(G["nb"])
    el = ol.Elements(s, G["elem_ids"], G["elem_ke"], G["scales"])
    asm = device_assembly(pkg, s, G["row_size"], G["column_index"], el)
    assert np.array_equal(device_array(asm), ol.oracle_assemble(s, nb, G["row_size"], G["column_index"], el))
    asm.upload_rhs(np.zeros(nb * s))
    asm.set_boundary_conditions(G["fix_ids"], G["fix_values"])
    assert np.array_equal(device_array(asm), G["array_post"])
    if bool(G["forces_comparable"]):
        f = asm.download_rhs()
        assert np.array_equal(f, G["forces_post"])
        # ... and the solve on the device-assembled system is the solve of the reference-assembled one
        S = ol.Sys(s, nb, G["row_size"], G["column_index"], G["array_post"], G["forces_post"])
        ret, x_ref, info = ol.oracle_cg(S, nssor=32)
        asm.upload_x0(None)
        ok, nit, err, rho = asm.pcg_resident(nssor=32)
        assert ok == bool(ret) and abs(int(nit) - int(info.nit)) <= 2
        assert rel_l2(asm.download_x(), x_ref) <= 1e-8
    asm.close()


@pytest.mark.parametrize("stride", [2, 3])
def test_set_boundary_conditions_golden(pkg, stride):
    G = load(f"bc-rand-s{stride}.npz")
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(G["row_size"], G["column_index"], stride, G["array"]),
                       G["forces"], device=0)
    asm.sync_matrix()
    asm.upload_rhs(G["forces"])
    nat = G["natural"].copy()
    asm.set_boundary_conditions(G["fix_ids"], G["fix_values"], G["force_ids"], G["force_values"], G["add_to_forces"], nat)
    assert np.array_equal(device_array(asm), G["array_post"])
    assert np.array_equal(asm.download_rhs(), G["forces_post"])
    assert np.array_equal(nat, G["natural_post"])
    asm.close()


@pytest.mark.parametrize("stride", [1, 2, 3, 4, 6])
def test_set_boundary_conditions_matches_oracle_all_strides(pkg, ol, stride):
    nb = 70
    rs, ci, arr, b = random_spd_blocks(stride, nb, 500 + stride)
    n = nb * stride
    rng = np.random.default_rng(stride)
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(rs, ci, stride, arr), b, device=0)
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


@pytest.mark.parametrize("dims,stride,ragged", [((9, 8), 2, False), ((6, 5, 7), 3, False), ((7, 6, 5), 3, True),
                                                ((12, 11), 1, False), ((5, 4, 4), 4, False), ((4, 4, 3), 6, True)])
def test_assemble_matches_oracle_bit_for_bit(pkg, ol, dims, stride, ragged):
    nb, el = grid_elements(ol, dims, stride, seed=sum(dims) + stride, ragged=ragged)
    rs, ci = el.pattern(nb)
    asm = device_assembly(pkg, stride, rs, ci, el)
    want = ol.oracle_assemble(stride, nb, rs, ci, el)
    assert np.array_equal(device_array(asm), want)
    # deterministic: a second context gives the same bits
    asm2 = device_assembly(pkg, stride, rs, ci, el)
    assert np.array_equal(device_array(asm2), want)
    asm.close()
    asm2.close()


def test_incremental_reassembly_after_damage_and_elimination(pkg, ol):
    """The damage-step loop: elimination, then some elements change, re-assemble only what they (and the BCs) touched."""
    stride, dims = 3, (8, 7, 6)
    nb, el = grid_elements(ol, dims, stride, seed=3)
    rs, ci = el.pattern(nb)
    n = nb * stride
    rng = np.random.default_rng(9)
    asm = device_assembly(pkg, stride, rs, ci, el)
    fix = np.sort(rng.choice(n, 40, replace=False)).astype(np.uint32)
    fv = rng.standard_normal(40)
    b = rng.standard_normal(n)
    for step in range(3):
        pre = ol.oracle_assemble(stride, nb, rs, ci, el)
        assert np.array_equal(device_array(asm), pre)
        a0, f0, _, _ = ol.oracle_set_bcs(stride, nb, rs, ci, pre, b, fix, fv)
        asm.upload_rhs(b)
        asm.set_boundary_conditions(fix, fv)
        assert np.array_equal(device_array(asm), a0) and np.array_equal(asm.download_rhs(), f0)
        # "damage": two ranges of elements get new matrices and scales
        for first, count in ((5, 17), (el.n_elem - 30, 30)):
            el.ke[first:first + count] *= rng.uniform(0.1, 1.0, (count, 1, 1, 1))
            el.scales[first:first + count] = rng.uniform(0.5, 2.0, count)
            asm.update_elements(first, el.ke[first:first + count], el.scales[first:first + count])
        asm.assemble()
    assert np.array_equal(device_array(asm), ol.oracle_assemble(stride, nb, rs, ci, el))
    asm.close()


def test_assembly_errors(pkg, ol):
    nb, el = grid_elements(ol, (5, 5), 2, seed=1)
    rs, ci = el.pattern(nb)
    asm = pkg.Assembly(None, None, device=0)
    with pytest.raises(pkg.AmieB200Error):
        asm.set_elements(el.ids)                              # before set_structure
    asm.set_structure_only(2, rs, ci)
    with pytest.raises(pkg.AmieB200Error):
        asm.assemble()                                        # before set_elements
    bad = el.ids.copy()
    bad[0, 0], bad[0, 1] = 0, nb - 1                          # far-apart nodes: block not in the pattern
    if not ((ci[:rs[0]] == nb - 1).any()):
        with pytest.raises(pkg.AmieB200Error):
            asm.set_elements(bad)
    bad = el.ids.copy()
    bad[3, 2] = nb + 5
    with pytest.raises(pkg.AmieB200Error):
        asm.set_elements(bad)                                 # node id out of range
    asm.set_elements(el.ids)
    with pytest.raises(pkg.AmieB200Error):
        asm.assemble()                                        # no elementary matrices yet
    with pytest.raises(pkg.AmieB200Error):
        asm.update_elements(el.n_elem - 1, el.ke[:2])         # range out of bounds
    asm.update_elements(0, el.ke, el.scales)
    asm.assemble()
    asm.upload_rhs(np.zeros(nb * 2))
    with pytest.raises(pkg.AmieB200Error):
        asm.set_boundary_conditions([4, 2], [0., 0.])         # ids not ascending
    with pytest.raises(pkg.AmieB200Error):
        asm.set_boundary_conditions([2, 2], [0., 0.])         # duplicate
    with pytest.raises(pkg.AmieB200Error):
        asm.set_boundary_conditions([nb * 2], [0.])           # out of range
    # a new topology drops the gather lists
    asm.set_structure_only(2, rs, ci)
    with pytest.raises(pkg.AmieB200Error):
        asm.assemble()
    asm.close()


def test_assembly_throughput_probe(pkg, ol):
    """Not a pass/fail on speed: records the first device timings of this row (gpurun_out/assembly_probe.json)."""
    import json
    stride, dims = 3, (48, 48, 48)
    rng = np.random.default_rng(0)
    nn = int(np.prod(dims))
    node = np.arange(nn).reshape(dims)
    corners = [node[tuple(slice(o, dims[a] - 1 + o) for a, o in enumerate(off))].reshape(-1) for off in np.ndindex(2, 2, 2)]
    ids = np.stack(corners, 1).astype(np.uint32)
    ke = rng.standard_normal((ids.shape[0], 8, 8, 9))
    el = ol.Elements(stride, ids, ke, None)
    rs, ci = el.pattern(nn)
    asm = device_assembly(pkg, stride, rs, ci, el)
    asm.assemble()
    st = asm.stats()
    n = nn * stride
    fix = np.arange(0, n, 97, dtype=np.uint32)
    asm.upload_rhs(np.zeros(n))
    asm.set_boundary_conditions(fix, np.ones(fix.size))
    st2 = asm.stats()
    nnzb = int(ci.size)
    gather_bytes = el.ke.size * 8 + el.ke.size // 9 * 4 + nnzb * 72 + nnzb * 4
    rec = dict(nodes=nn, elements=int(ids.shape[0]), nnzb=nnzb, elements_ms=st.elements_ms, assemble_ms_incremental_noop=st.assemble_ms,
               bc_ms=st2.bc_ms, gather_algorithmic_bytes=gather_bytes)
    # full re-accumulation timing: mark everything by re-uploading all elements
    asm.update_elements(0, el.ke, el.scales)
    asm.assemble()
    rec["assemble_ms_full"] = asm.stats().assemble_ms
    rec["assemble_GBs_full"] = gather_bytes / (rec["assemble_ms_full"] * 1e-3) / 1e9 if rec["assemble_ms_full"] > 0 else None
    want = ol.oracle_assemble(stride, nn, rs, ci, el)
    assert np.array_equal(device_array(asm), want)
    out = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "assembly_probe.json"), "w") as f:
        json.dump(rec, f)
    print(rec)
    asm.close()
