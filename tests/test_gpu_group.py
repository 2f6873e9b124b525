"""ONE context over several GPUs, driven from one caller thread: amie_b200_create(devices, ndev > 1) (csrc/group.cu).

The reference's caller is a single thread of a single process (Assembly::cgsolve, solvers/assembly.cpp:1841-1850); a
group context keeps that shape -- global host arrays in and out through the usual C-ABI calls -- and partitions the
block rows inside the library.  Every part runs the per-rank code of csrc/dist.cu (interior SpMV overlapped with peer
halo pushes, mailbox reductions, identical loop decisions everywhere).

`devices` may name the SAME ordinal several times: the parts then live side by side on one GPU, with the same kernels,
the same halo pushes and device-side flags.  That is how the whole multi-part path is covered on a one-GPU box; with
two or more GPUs the same tests also run across distinct devices (NVLink peer stores).

Bars as everywhere (BASELINE.json north_star): PCG iteration count within +-2 of the reference, x within 1e-8 relative
L2; SpMV at 1e-13 relative; BiCGStab on convergence and solution error.
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

X_TOL = 1e-8
NIT_TOL = 2
SPMV_TOL = 1e-13


def device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [[0, 1]]
    if n >= 4:
        sets += [[0, 1, 2, 3]]
    if n >= 8:
        sets += [list(range(8))]
    return sets


def pytest_generate_tests(metafunc):
    if "devices" in metafunc.fixturenames:
        try:
            sets = device_sets()
        except Exception:
            sets = [[0, 0]]
        metafunc.parametrize("devices", sets, ids=["dev" + "".join(map(str, s)) for s in sets])


def group_assembly(pkg, S, devices):
    return pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(S.row_size, S.column_index, S.stride, S.array), S.b, devices=devices)


@pytest.mark.parametrize("preset,n", [("S3-hex", 14), ("S3-tet", 12), ("S2-tri", 40), ("ASR-hex", 12)])
def test_group_pcg_and_bicgstab_parity(pkg, ol, systems, devices, preset, n):
    S = systems(preset, n)
    asm = group_assembly(pkg, S, devices)
    ret, x_ref, info = ol.oracle_cg(S, nssor=32)
    cg = pkg.ConjugateGradient(asm)
    cg.nssor = 32
    ok = cg.solve(None, None, 1e-10, -1)
    assert ok == bool(ret)
    assert abs(int(cg.nit) - int(info.nit)) <= NIT_TOL, (cg.nit, info.nit)
    assert rel_l2(cg.x, x_ref) <= X_TOL
    st = asm.stats()
    assert st.ndof == S.n and st.nnzb == S.column_index.size and st.iterations == cg.nit
    # same matrix, second solve from a warm start shorter than N (conjugategradient.cpp:95-104): values stay resident
    x0 = x_ref[:S.n // 2].copy()
    ret2, x_ref2, info2 = ol.oracle_cg(S, x0=x0, nssor=32)
    ok2 = cg.solve(x0, None, 1e-10, -1)
    assert ok2 == bool(ret2) and abs(int(cg.nit) - int(info2.nit)) <= NIT_TOL
    assert rel_l2(cg.x, x_ref2) <= X_TOL
    # BiCGStab on the same context
    bi = pkg.BiConjugateGradientStabilized(asm)
    okb = bi.solve(None, None, 1e-10, -1)
    retb, xb_ref, _ = ol.oracle_bicgstab(S)
    assert okb == bool(retb) and rel_l2(bi.x, xb_ref) <= X_TOL
    asm.close()


@pytest.mark.parametrize("preset,n", [("S3-hex", 12), ("S2-tri", 30)])
def test_group_spmv_modes_residual_diagonal(pkg, ol, systems, devices, preset, n):
    S = systems(preset, n)
    asm = group_assembly(pkg, S, devices)
    v = np.random.default_rng(11).standard_normal(S.n)
    scale = np.abs(S.to_scipy()).dot(np.abs(v)).max()
    assert np.abs(asm.spmv(v) - ol.oracle_assign(S, v)).max() <= SPMV_TOL * scale
    assert np.abs(asm.spmv(v, minus_b=S.b) - ol.oracle_assign(S, v, S.b)).max() <= SPMV_TOL * (scale + np.abs(S.b).max())
    # rowstart / colstart cut through the partition at different places: inside part 0, on a part boundary's far side
    for frac in (5, 2):
        rs = S.stride * (S.nb // frac)
        y = asm.spmv(v, minus_b=S.b, rowstart=rs, colstart=rs)
        assert not y[:rs].any()
        assert np.abs(y - ol.oracle_assign(S, v, S.b, rs, rs)).max() <= SPMV_TOL * (scale + np.abs(S.b).max())
        y = asm.spmv(v, rowstart=0, colstart=rs)
        assert np.abs(y - ol.oracle_assign(S, v, None, 0, rs)).max() <= SPMV_TOL * scale
    u = np.random.default_rng(4).standard_normal(S.n)
    r, nrm = asm.residual(u)
    ro = ol.oracle_spmv_serial(S, u, S.b)
    assert np.abs(r - ro).max() <= 1e-13 * np.abs(ro).max()
    assert nrm == pytest.approx(np.linalg.norm(ro), rel=1e-12)
    assert np.array_equal(asm.inverse_diagonal(), ol.oracle_inverse_diagonal(S))
    asm.close()


@pytest.mark.parametrize("preset,n,frac", [("S3-hex", 12, 4), ("S2-tri", 40, 3), ("S2-tri", 40, 2)])
def test_group_pcg_rowstart(pkg, ol, systems, devices, preset, n, frac):
    """rowstart = colstart > 0 (space-time planes, solvers/assembly.cpp:327-343) on a partitioned matrix: the cut falls
    inside one part, the parts in front of it own no active row at all."""
    S = systems(preset, n)
    asm = group_assembly(pkg, S, devices)
    rs = S.stride * (S.nb // frac)
    ret, x_ref, info = ol.oracle_cg(S, nssor=32, rowstart=rs, colstart=rs)
    cg = pkg.ConjugateGradient(asm)
    cg.nssor, cg.rowstart, cg.colstart = 32, rs, rs
    ok = cg.solve(None, None, 1e-10, -1)
    assert ok == bool(ret)
    assert abs(int(cg.nit) - int(info.nit)) <= NIT_TOL, (cg.nit, info.nit)
    assert rel_l2(cg.x, x_ref) <= X_TOL
    assert np.array_equal(cg.x[:rs], S.b[:rs])
    asm.close()


def test_group_resident_sequence_and_generator(pkg, ol, systems, devices):
    """The device-resident calls and the generator on a group: synth_to_device == uploaded host arrays, bit for bit;
    cgsolve_resident keeps displacementHistory per part (extrapolation and history shift are per entry)."""
    preset, n = "S3-hex", 14
    S = systems(preset, n)
    syn = pkg.Synth(preset, n)
    a1 = group_assembly(pkg, S, devices)
    a1.sync_matrix()
    a1.upload_rhs(S.b)
    a1.upload_x0(None)
    ok1, nit1, _, _ = a1.pcg_resident(nssor=32)
    x1 = a1.download_x()
    a2 = pkg.Assembly(devices=devices)
    syn.to_device(a2)
    a2.upload_x0(None)
    ok2, nit2, _, _ = a2.pcg_resident(nssor=32)
    x2 = a2.download_x()
    assert ok1 and ok2 and nit1 == nit2 and np.array_equal(x1, x2)
    assert np.array_equal(a2.download_rhs(), S.b)
    # load-step loop without host vectors against the single-device context doing the same
    ref = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(S.row_size, S.column_index, S.stride, S.array), S.b, device=0)
    ref.sync_matrix()
    for a in (a2, ref):
        a.upload_x0(None)               # `displacements` before the first step: zero on both
    for step, f in enumerate((1.0, 1.5, 2.25)):
        for a in (a2, ref):
            a.upload_rhs(S.b * f)
        okg, nitg, _, _ = a2.cgsolve_resident()
        okr, nitr, _, _ = ref.cgsolve_resident()
        assert okg == okr and abs(int(nitg) - int(nitr)) <= NIT_TOL, (step, nitg, nitr)
        assert rel_l2(a2.download_x(), ref.download_x()) <= X_TOL
    x0g, caseg = a2.extrapolate()
    x0r, caser = ref.extrapolate()
    assert caseg == caser == 1 and rel_l2(x0g, x0r) <= 1e-7
    for a in (a1, a2, ref):
        a.close()


def test_group_matches_single_device_iterates(pkg, systems, devices):
    """Same system, the solve cut short by a loose tolerance (maxit bounds only the restarts, conjugategradient.cpp:93,
    :121): the partitioned iterate equals the single-device one up to the rounding of the dot products (the parts sum
    their rows in another order) -- same iteration, same field."""
    S = systems("S3-tet", 14)
    one = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(S.row_size, S.column_index, S.stride, S.array), S.b, device=0)
    grp = group_assembly(pkg, S, devices)
    xs = []
    for a in (one, grp):
        cg = pkg.ConjugateGradient(a)
        cg.nssor = 32
        cg.solve(None, None, 1e-3, -1)
        xs.append((int(cg.nit), cg.x.copy()))
        a.close()
    assert 5 < xs[0][0] == xs[1][0]
    assert rel_l2(xs[1][1], xs[0][1]) <= 1e-9


def test_group_errors(pkg, systems):
    S = systems("S3-hex", 6)
    L = pkg.lib()
    # a device that does not exist
    import ctypes
    bad = (ctypes.c_int * 2)(0, 99)
    assert not L.amie_b200_create(bad, 2)
    assert b"no such CUDA device" in L.amie_b200_global_error()
    asm = group_assembly(pkg, S, [0, 0])
    with pytest.raises(pkg.AmieB200Error) as e:
        asm.upload_rhs(S.b)                 # before set_structure
    assert e.value.code == pkg.ERR_STATE
    asm.sync_matrix()
    # the matrix comes back in the caller's numbering, whatever the parts number their columns like
    rs, ci, arr, _ = asm.download_matrix()
    assert np.array_equal(rs, S.row_size) and np.array_equal(ci, S.column_index) and np.array_equal(arr, S.array)
    with pytest.raises(pkg.AmieB200Error) as e:
        asm.set_elements(np.full((1, 8), S.nb + 7, np.uint32))      # a node beyond the (global) matrix
    assert e.value.code == pkg.ERR_ARG
    asm.close()
    # strides other than 2 and 3 stay on one device
    from conftest import random_spd_blocks
    rs, ci, arr, b = random_spd_blocks(4, 10, 3)
    a4 = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(rs, ci, 4, arr), b, devices=[0, 0])
    with pytest.raises(pkg.AmieB200Error) as e:
        a4.sync_matrix()
    assert e.value.code == pkg.ERR_UNSUPPORTED
    a4.close()


@pytest.mark.parametrize("name", ["AMIE-2d-s20-assembly.npz", "AMIE-3d-s400-assembly.npz"])
def test_group_assembly_and_elimination_reproduce_the_featuretree_matrix(pkg, ol, devices, name):
    """SURVEY section 8 row f1 on a multi-device context: every device sees the whole element list (global node ids) and the
    global id lists of the boundary conditions, and keeps what lands on the block rows it owns.  The parts put together
    are the matrix and the force vector the unmodified FeatureTree solved, bit for bit; the solve that follows runs on
    the device-assembled parts without a host matrix."""
    import os
    G = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", name)))
    s, nb = int(G["stride"]), int(G["nb"])
    el = ol.Elements(s, G["elem_ids"], G["elem_ke"], G["scales"])
    asm = pkg.Assembly(None, None, devices=devices)
    asm.set_structure_only(s, G["row_size"], G["column_index"])
    asm.set_elements(el.ids)
    asm.update_elements(0, el.ke, el.scales)
    asm.assemble()
    assert np.array_equal(asm.download_matrix()[2], ol.oracle_assemble(s, nb, G["row_size"], G["column_index"], el))
    asm.upload_rhs(np.zeros(nb * s))
    asm.set_boundary_conditions(G["fix_ids"], G["fix_values"])
    assert np.array_equal(asm.download_matrix()[2], G["array_post"])
    if bool(G["forces_comparable"]):
        assert np.array_equal(asm.download_rhs(), G["forces_post"])
        S = ol.Sys(s, nb, G["row_size"], G["column_index"], G["array_post"], G["forces_post"])
        ret, x_ref, info = ol.oracle_cg(S, nssor=32)
        asm.upload_x0(None)
        ok, nit, err, rho = asm.pcg_resident(nssor=32)
        assert ok == bool(ret) and abs(int(nit) - int(info.nit)) <= NIT_TOL
        assert rel_l2(asm.download_x(), x_ref) <= X_TOL
    # a damage-like step: some elements change, only their stored blocks are re-accumulated -- on every device
    rng = np.random.default_rng(4)
    first, count = el.n_elem // 4, max(1, el.n_elem // 6)
    el.ke[first:first + count] *= rng.uniform(0.1, 0.9, (count, 1, 1, 1))
    asm.update_elements(first, el.ke[first:first + count], el.scales[first:first + count])
    asm.assemble()
    assert np.array_equal(asm.download_matrix()[2], ol.oracle_assemble(s, nb, G["row_size"], G["column_index"], el))
    asm.close()


@pytest.mark.parametrize("stride", [2, 3])
def test_group_set_boundary_conditions_matches_oracle(pkg, ol, devices, stride):
    """Imposed displacements, imposed forces, the natural-condition vector and the additional forces, all at once."""
    from conftest import random_spd_blocks
    nb = 90
    rs, ci, arr, b = random_spd_blocks(stride, nb, 900 + stride)
    n = nb * stride
    rng = np.random.default_rng(40 + stride)
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(rs, ci, stride, arr), b, devices=devices)
    for nfix in (1, n // 4, n):
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
        assert np.array_equal(asm.download_matrix()[2], a0)
        assert np.array_equal(asm.download_rhs(), f0)
        assert np.array_equal(nat, n0)
    asm.close()


@pytest.mark.parametrize("name", ["AMIE-2d-s20-fields.npz", "AMIE-3di-s400-fields.npz"])
def test_group_field_recovery_matches_reference_bits(pkg, ol, devices, name):
    """SURVEY section 8 row f2 on a multi-device context: the ELEMENTS are split over the devices, each holds the whole
    displacement field.  From a host field and from the resident solution of a solve (gathered from the parts over the
    peer links): the bits ElementState::getField gave inside the FeatureTree run."""
    import os
    from test_gpu_recovery import same_bits
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    dim = int(g["dshape"].shape[2])
    sysname = name.replace("-fields", "").replace("3di", "3d")
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", sysname))
    s, nb = int(G["stride"]), int(G["nb"])
    assert s == dim
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(G["row_size"], G["column_index"], s, G["array"]), G["b"], devices=devices)
    asm.sync_matrix()
    asm.set_element_kinematics(dim, g["ids"], g["dshape"], g["jinv"])
    asm.set_element_behaviour(g["tensors"], g["imposed_strain"], g["imposed_stress"], g["tensor_of_elem"])
    tot, mech, sig = asm.element_fields(g["u"])
    assert same_bits(tot, g["total_strain"]) and same_bits(mech, g["mechanical_strain"]) and same_bits(sig, g["real_stress"])
    for field, key in ((0, "principal_total_strain"), (2, "principal_real_stress")):
        if key in g.files:
            got = asm.element_principal(field)
            # 2D: only sqrt reaches the results -> the reference's bits; 3D goes through pow / atan2 / cos / sin of the
            # device math library and is held to 1e-12 like on one device (tests/test_gpu_recovery.py)
            if dim == 2:
                assert same_bits(got, g[key])
            else:
                assert np.abs(got - g[key]).max() <= 1e-12 * np.abs(g[key]).max()
    # one behaviour per element (cut like the elements), and the field taken from the devices after a solve
    ne = g["ids"].shape[0]
    C = g["tensors"][g["tensor_of_elem"]]
    es, ss = g["imposed_strain"][g["tensor_of_elem"]], g["imposed_stress"][g["tensor_of_elem"]]
    asm.set_element_behaviour(C, es, ss, None)
    asm.upload_rhs(G["b"])
    asm.upload_x0(None)
    ok, nit, err, rho = asm.pcg_resident(nssor=32)
    x = asm.download_x()
    got = asm.element_fields(None)
    want = ol.oracle_element_fields(dim, g["ids"], g["dshape"], g["jinv"], x, C, es, ss, None)
    for a, b in zip(got, want):
        assert same_bits(a, b)
    asm.close()


@pytest.mark.parametrize("name", ["AMIE-3d-s400.npz", "AMIE-2d-s20.npz"])
def test_group_renumbered_device_matrix_gives_the_same_solve(pkg, ol, devices, name):
    """Assembly(devices=..., renumber=True): the devices hold the reverse-Cuthill-McKee numbering (structure permuted on
    the host, amie_b200_set_block_map on the multi-device context: every device's values gathered from the caller's
    array), the caller keeps its own.  Same SpMV, inverse diagonal and PCG answers as the reference on the original
    numbering."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    s, nb = int(G["stride"]), int(G["nb"])
    S = ol.Sys(s, nb, G["row_size"], G["column_index"], G["array"], G["b"])
    asm = pkg.Assembly(pkg.CoordinateIndexedSparseMatrix(G["row_size"], G["column_index"], s, G["array"]), G["b"], devices=devices,
                       renumber=True)
    v = G["v"]
    y = asm.spmv(v)
    assert asm.perm is not None and not np.array_equal(asm.perm, np.arange(nb))
    assert np.abs(y - G["assign"]).max() <= 1e-12 * np.abs(G["assign"]).max()
    cg = pkg.ConjugateGradient(asm)
    cg.nssor = 32
    ok = cg.solve(None, None, 1e-10, -1)
    ret, x_ref, info = ol.oracle_cg(S, nssor=32)
    assert ok == bool(ret) and abs(int(cg.nit) - int(info.nit)) <= NIT_TOL
    assert rel_l2(cg.x, x_ref) <= X_TOL
    asm.close()
