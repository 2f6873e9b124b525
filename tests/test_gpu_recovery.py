"""GPU parity of the field recovery after the solve (csrc/fields.cu) through the C-ABI:
bit for bit against what ElementState::getField answered inside an unmodified FeatureTree run
(tests/golden/AMIE-*-fields.npz) and against the oracle restatement on perturbed inputs."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*-fields.npz")))
NO_NODE = 0xFFFFFFFF


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def pattern_of(ids, nb):
    """Block sparsity pattern coupling the nodes of every element (all the context needs here is a structure)."""
    ids = ids.astype(np.int64)
    rows, cols = [np.arange(nb)], [np.arange(nb)]
    for j in range(ids.shape[1]):
        for k in range(ids.shape[1]):
            ok = (ids[:, j] != NO_NODE) & (ids[:, k] != NO_NODE)
            rows.append(ids[ok, j])
            cols.append(ids[ok, k])
    key = np.unique(np.concatenate(rows) * nb + np.concatenate(cols))
    return np.bincount(key // nb, minlength=nb).astype(np.uint32), (key % nb).astype(np.uint32)


def context_for(pkg, g, nb=None):
    dim = int(g["dim"])
    nb = g["u"].size // dim if nb is None else nb
    rs, ci = pattern_of(g["ids"], nb)
    asm = pkg.Assembly(device=0)
    asm.set_structure_only(dim, rs, ci)
    return asm, dim


@pytest.mark.parametrize("path", FIELDS, ids=[os.path.basename(p)[:-4] for p in FIELDS])
def test_fields_match_reference_bits(pkg, path):
    g = np.load(path)
    asm, dim = context_for(pkg, g)
    asm.set_element_kinematics(dim, g["ids"], g["dshape"], g["jinv"])
    asm.set_element_behaviour(g["tensors"], g["imposed_strain"], g["imposed_stress"], g["tensor_of_elem"])
    # host-supplied displacement field
    tot, mech, sig = asm.element_fields(g["u"])
    assert same_bits(tot, g["total_strain"])
    assert same_bits(mech, g["mechanical_strain"])
    assert same_bits(sig, g["real_stress"])
    # the resident solution (what a solve leaves in x): no upload of u in the call
    asm.upload_x0(g["u"])
    tot2, mech2, sig2 = asm.element_fields()
    assert same_bits(tot2, tot) and same_bits(mech2, mech) and same_bits(sig2, sig)
    st = asm.stats()
    assert st.field_elements == g["ids"].shape[0] and st.fields_ms > 0
    asm.close()


@pytest.mark.parametrize("path", FIELDS, ids=[os.path.basename(p)[:-4] for p in FIELDS])
def test_principal_values_match_reference(pkg, path):
    """getField(PRINCIPAL_*_FIELD): 2D has the reference's bits; 3D goes through pow / atan2 / cos / sin of the device
    math library and is held to 1e-12 of the largest principal value of the element."""
    g = np.load(path)
    asm, dim = context_for(pkg, g)
    asm.set_element_kinematics(dim, g["ids"], g["dshape"], g["jinv"])
    asm.set_element_behaviour(g["tensors"], g["imposed_strain"], g["imposed_stress"], g["tensor_of_elem"])
    with pytest.raises(pkg.AmieB200Error) as e:                       # nothing to take principal values of yet
        asm.element_principal(2)
    assert e.value.code == pkg.ERR_STATE
    asm.element_fields(g["u"])
    for field, key in ((0, "principal_total_strain"), (1, "principal_mechanical_strain"), (2, "principal_real_stress")):
        got, want = asm.element_principal(field), g[key]
        if dim == 2:
            assert same_bits(got, want), key
        else:
            scale = np.abs(want).max(axis=1, keepdims=True) + 1e-300
            assert (np.abs(got - want) <= 1e-12 * scale).all(), (key, float((np.abs(got - want) / scale).max()))
    asm.close()


def test_fields_per_element_behaviours_ragged_and_short_vector(pkg, ol):
    """One behaviour per element (damage-like), a random imposed stress, an unused slot, several tiles with a
    partial last one, and a displacement vector shorter than the dofs the elements reference."""
    g = np.load([p for p in FIELDS if "3di" in p][0])
    rng = np.random.default_rng(5)
    ne = g["ids"].shape[0]
    toe = g["tensor_of_elem"]
    C = g["tensors"][toe] * rng.uniform(0.2, 1.0, (ne, 1, 1))         # damaged stiffness per element
    es = g["imposed_strain"][toe] + 1e-4 * rng.standard_normal((ne, 6))
    ss = rng.standard_normal((ne, 6))
    ids5 = np.concatenate([g["ids"], np.full((ne, 1), NO_NODE, np.uint32)], axis=1)
    ds5 = np.concatenate([g["dshape"], np.full((ne, 1, 3), 7.0)], axis=1)
    u = g["u"][: (2 * g["u"].size // 3) // 3 * 3 + 1]                  # cuts inside a node
    asm, dim = context_for(pkg, g)
    asm.set_element_kinematics(3, ids5, ds5, g["jinv"])
    asm.set_element_behaviour(C, es, ss, None)
    got = asm.element_fields(u)
    want = ol.oracle_element_fields(3, ids5, ds5, g["jinv"], u, C, es, ss, None)
    for a, b in zip(got, want):
        assert same_bits(a, b)
    # behaviours replaced without touching the kinematics (a damage step)
    C2 = C * 0.5
    asm.set_element_behaviour(C2, None, None, None)
    got = asm.element_fields(u)
    want = ol.oracle_element_fields(3, ids5, ds5, g["jinv"], u, C2, None, None, None)
    for a, b in zip(got, want):
        assert same_bits(a, b)
    asm.close()


def test_fields_after_a_solve_use_the_resident_solution(pkg, ol):
    """PCG on the device, then strains from the x it left in HBM == strains from the downloaded x."""
    g = np.load([p for p in FIELDS if "2d" in p][0])
    syn = pkg.Synth("S2-tri", 24)
    asm = syn.assembly(device=0)
    cg = pkg.ConjugateGradient(asm)
    cg.nssor = 32
    assert cg.solve(None, None, 1e-10, -1)
    nb = syn.nb
    rng = np.random.default_rng(3)
    ne = 1000
    ids = rng.integers(0, nb, (ne, 3)).astype(np.uint32)
    ds = rng.standard_normal((ne, 3, 2))
    ji = rng.standard_normal((ne, 2, 2))
    asm.set_element_kinematics(2, ids, ds, ji)
    asm.set_element_behaviour(g["tensors"], None, None, np.zeros(ne, np.uint32))
    got = asm.element_fields()
    want = ol.oracle_element_fields(2, ids, ds, ji, cg.x, g["tensors"], None, None, np.zeros(ne, np.uint32))
    for a, b in zip(got, want):
        assert same_bits(a, b)
    asm.close()


def test_fields_error_paths(pkg):
    g = np.load([p for p in FIELDS if "2d" in p][0])
    asm, dim = context_for(pkg, g)
    with pytest.raises(pkg.AmieB200Error) as e:                       # order
        asm._field_shape = (g["ids"].shape[0], 3)
        asm.set_element_behaviour(g["tensors"], None, None, g["tensor_of_elem"])
    assert e.value.code == pkg.ERR_STATE
    with pytest.raises(pkg.AmieB200Error) as e:                       # dim must match the stride
        asm.set_element_kinematics(3, np.zeros((1, 4), np.uint32), np.zeros((1, 4, 3)), np.zeros((1, 3, 3)))
    assert e.value.code == pkg.ERR_UNSUPPORTED
    asm.set_element_kinematics(2, g["ids"], g["dshape"], g["jinv"])
    with pytest.raises(pkg.AmieB200Error) as e:                       # no behaviours yet
        asm.element_fields(g["u"])
    assert e.value.code == pkg.ERR_STATE
    with pytest.raises(pkg.AmieB200Error) as e:                       # index beyond the table
        asm.set_element_behaviour(g["tensors"], None, None, np.full(g["ids"].shape[0], 9, np.uint32))
    assert e.value.code == pkg.ERR_ARG
    with pytest.raises(pkg.AmieB200Error) as e:                       # no index: needs one behaviour per element
        asm.set_element_behaviour(g["tensors"], None, None, None)
    assert e.value.code == pkg.ERR_ARG
    asm.close()


@pytest.mark.parametrize("mode,sampling", [("2d", 20), ("3di", 400)])
def test_fields_inside_a_featuretree_run(tmp_path, mode, sampling):
    """The drop-in build of the FeatureTree driver (oracle/e2e_harness.cpp with -DAMIE_B200_E2E): after F.step() it
    asks ElementState::getField for every element, then the device for the same fields through the context the shim
    keeps for F's Assembly, and counts the values whose bits differ -- in the same process, on the same operands."""
    import re
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "amie_e2e_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref e2e binaries not prebuilt (no /root/reference at build time)")
    p = subprocess.run([exe, mode, str(sampling), "u.bin", "dump.bin", "el.bin", "fields.bin"], cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    m = re.search(r"fields-on-device: (\d+) elements, (\d+) values compared with ElementState::getField, (\d+) differ", p.stderr)
    assert m, p.stderr[-2000:]
    assert int(m.group(1)) > 1000 and int(m.group(3)) == 0, m.group(0)
