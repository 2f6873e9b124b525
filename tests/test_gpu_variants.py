"""Opt-in kernel variants must give the bits of the default ones (which the parity tests pin)."""
import os

import numpy as np
import pytest

from test_gpu_assembly import device_array, grid_elements, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["AMIE-2d-s20-assembly.npz", "AMIE-3d-s400-assembly.npz"])
def test_assemble_variant1_reproduces_featuretree_matrix(pkg, ol, name):
    G = load(name)
    s, nb = int(G["stride"]), int(G["nb"])
    el = ol.Elements(s, G["elem_ids"], G["elem_ke"], G["scales"])
    asm = pkg.Assembly(None, None, device=0)
    asm.set_option("assemble_variant", 1)
    asm.set_structure_only(s, G["row_size"], G["column_index"])
    asm.set_elements(el.ids)
    asm.update_elements(0, el.ke, el.scales)
    asm.assemble()
    assert np.array_equal(device_array(asm), ol.oracle_assemble(s, nb, G["row_size"], G["column_index"], el))
    asm.upload_rhs(np.zeros(nb * s))
    asm.set_boundary_conditions(G["fix_ids"], G["fix_values"])
    assert np.array_equal(device_array(asm), G["array_post"])
    asm.close()


@pytest.mark.parametrize("dims,stride,ragged", [((9, 8), 2, False), ((7, 6, 5), 3, True), ((12, 11), 1, False),
                                                ((5, 4, 4), 4, False), ((4, 4, 3), 6, True), ((23, 19, 17), 3, False)])
def test_assemble_variant1_matches_oracle_incl_incremental(pkg, ol, dims, stride, ragged):
    nb, el = grid_elements(ol, dims, stride, seed=sum(dims) + stride, ragged=ragged)
    rs, ci = el.pattern(nb)
    asm = pkg.Assembly(None, None, device=0)
    asm.set_option("assemble_variant", 1)
    asm.set_structure_only(stride, rs, ci)
    asm.set_elements(el.ids)
    asm.update_elements(0, el.ke, el.scales)
    asm.assemble()
    assert np.array_equal(device_array(asm), ol.oracle_assemble(stride, nb, rs, ci, el))
    # a damage-like step: some elements change, only their blocks are re-accumulated
    rng = np.random.default_rng(1)
    first, count = el.n_elem // 3, max(1, el.n_elem // 5)
    el.ke[first:first + count] *= rng.uniform(0.1, 0.9, (count, 1, 1, 1))
    asm.update_elements(first, el.ke[first:first + count], el.scales[first:first + count])
    asm.assemble()
    assert np.array_equal(device_array(asm), ol.oracle_assemble(stride, nb, rs, ci, el))
    asm.close()
