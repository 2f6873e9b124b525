"""Pins oracle/amie_oracle_fields.c (field recovery after the solve) bit for bit against
what ElementState::getField answered inside an unmodified FeatureTree run (tests/golden/AMIE-*-fields.npz, made by
tests/golden/make_golden_fields.py with the compiled reference)."""
import glob
import os

import numpy as np
import pytest

FIELDS = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*-fields.npz")))


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def test_fixtures_exist():
    names = {os.path.basename(p) for p in FIELDS}
    assert {"AMIE-2d-s20-fields.npz", "AMIE-3di-s400-fields.npz"} <= names


@pytest.mark.parametrize("path", FIELDS, ids=[os.path.basename(p)[:-4] for p in FIELDS])
def test_oracle_fields_match_reference(ol, path):
    g = np.load(path)
    tot, mech, sig = ol.oracle_element_fields(int(g["dim"]), g["ids"], g["dshape"], g["jinv"], g["u"], g["tensors"],
                                              g["imposed_strain"], g["imposed_stress"], g["tensor_of_elem"])
    assert np.count_nonzero(g["total_strain"]) > g["ids"].shape[0]
    assert same_bits(tot, g["total_strain"])
    assert same_bits(mech, g["mechanical_strain"])
    assert same_bits(sig, g["real_stress"])


@pytest.mark.parametrize("path", FIELDS, ids=[os.path.basename(p)[:-4] for p in FIELDS])
def test_oracle_principal_values_match_reference(ol, path):
    """toPrincipal restated: against getField(PRINCIPAL_TOTAL_STRAIN | PRINCIPAL_MECHANICAL_STRAIN | PRINCIPAL_REAL_STRESS)
    of the same FeatureTree run, bit for bit (the build container's C library serves both)."""
    g = np.load(path)
    dim = int(g["dim"])
    assert same_bits(ol.oracle_principal(dim, g["total_strain"], True), g["principal_total_strain"])
    assert same_bits(ol.oracle_principal(dim, g["mechanical_strain"], True), g["principal_mechanical_strain"])
    assert same_bits(ol.oracle_principal(dim, g["real_stress"], False), g["principal_real_stress"])
    assert np.count_nonzero(g["principal_real_stress"]) > g["ids"].shape[0]


def test_oracle_fields_options(ol):
    """Per-element behaviours (no index), absent imposed terms, unused node slots and dof ids beyond the vector."""
    g = np.load([p for p in FIELDS if "3di" in p][0])
    ids, toe = g["ids"], g["tensor_of_elem"]
    ref = ol.oracle_element_fields(3, ids, g["dshape"], g["jinv"], g["u"], g["tensors"], g["imposed_strain"],
                                   g["imposed_stress"], toe)
    per = ol.oracle_element_fields(3, ids, g["dshape"], g["jinv"], g["u"], g["tensors"][toe], g["imposed_strain"][toe],
                                   g["imposed_stress"][toe], None)
    for a, b in zip(ref, per):
        assert same_bits(a, b)
    # imposed stress is all zero in the reference (StiffnessWithImposedStrain::getImposedStress returns imposed*0):
    # leaving it out gives the same bits
    nost = ol.oracle_element_fields(3, ids, g["dshape"], g["jinv"], g["u"], g["tensors"], g["imposed_strain"], None, toe)
    assert same_bits(ref[2], nost[2])
    # an unused fifth slot changes nothing
    ne = ids.shape[0]
    ids5 = np.concatenate([ids, np.full((ne, 1), 0xFFFFFFFF, np.uint32)], axis=1)
    ds5 = np.concatenate([g["dshape"], np.full((ne, 1, 3), 7.0)], axis=1)
    pad = ol.oracle_element_fields(3, ids5, ds5, g["jinv"], g["u"], g["tensors"], g["imposed_strain"], g["imposed_stress"], toe)
    for a, b in zip(ref, pad):
        assert same_bits(a, b)
    # dofs beyond the vector read as zero (ElementState::step, elements/integrable_entity.cpp:3641-3648)
    cut = g["u"][:g["u"].size // 2]
    uz = g["u"].copy()
    uz[cut.size:] = 0.
    a = ol.oracle_element_fields(3, ids, g["dshape"], g["jinv"], cut, g["tensors"], g["imposed_strain"], g["imposed_stress"], toe)
    b = ol.oracle_element_fields(3, ids, g["dshape"], g["jinv"], uz, g["tensors"], g["imposed_strain"], g["imposed_stress"], toe)
    for p, q in zip(a, b):
        assert same_bits(p, q)
