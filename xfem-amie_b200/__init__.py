"""Python mirror of the reference's solver interface over the C-ABI of libamie_b200.so.

The product is the CUDA library (csrc/, C-ABI in include/amie_b200.h).  This module is the thin
host-side mirror used by tests and bench.py: the same names, argument meaning and error behaviour
as the reference's solver classes,

    Amie::CoordinateIndexedSparseMatrix      sparse/sparse_matrix.h:129-136
    Amie::Assembly  (getMatrix / getForces)  solvers/assembly.h:228-450
    Amie::ConjugateGradient                  solvers/conjugategradient.h:22-43
    Amie::BiConjugateGradientStabilized      solvers/biconjugategradientstabilized.h:19-24
    Amie::NullPreconditionner                solvers/preconditionners.h:28-32

There is NO CPU fallback: without the built extension (or without an sm_100 GPU) every compute
call raises.  The directory name has a hyphen, so load it with ``load_package()`` of
``__graft_entry__`` (importlib), which registers it as ``xfem_amie_b200``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libamie_b200.so")

u64 = ctypes.c_uint64
f64 = ctypes.c_double

ERR_CUDA, ERR_ARG, ERR_STATE, ERR_NAN, ERR_UNSUPPORTED, ERR_NCCL = -1, -2, -3, -4, -5, -6
PRECOND_JACOBI, PRECOND_NULL, PRECOND_DIAGONAL_SQUARED, PRECOND_LUMPED, PRECOND_DIAGONAL = 0, 1, 2, 3, 4
PRECOND_BLOCK2X2, PRECOND_BLOCK3X3 = 5, 6
default_solver_precision = 1e-10          # polynomial/variable.h:14


class AmieB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"amie_b200 error {code}: {msg}")
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [("stride", u64), ("nb", u64), ("nnzb", u64), ("ndof", u64),
                ("spmv_launches", u64), ("kernel_launches", u64), ("smoothing_spmv", u64),
                ("iterations", u64), ("restarts", u64),
                ("spmv_ms_total", f64), ("spmv_timed", u64), ("solve_ms", f64),
                ("h2d_ms", f64), ("d2h_ms", f64), ("h2d_bytes", u64), ("d2h_bytes", u64),
                ("structure_ms", f64), ("values_ms", f64),
                ("spmv_algorithmic_bytes", u64), ("device_bytes", u64),
                ("elements_ms", f64), ("assemble_ms", f64), ("bc_ms", f64), ("element_blocks", u64),
                ("fields_ms", f64), ("field_elements", u64), ("early_return", u64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """The CUDA extension.  Fails loudly when it was not built: there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        vp, cp, ci = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int
        L.amie_b200_create.restype = vp
        L.amie_b200_create.argtypes = [vp, ci]
        L.amie_b200_destroy.argtypes = [vp]
        for n in ("amie_b200_last_error", "amie_b200_global_error", "amie_b200_version"):
            getattr(L, n).restype = cp
        L.amie_b200_last_error.argtypes = [vp]
        L.amie_b200_set_structure.argtypes = [vp, ci, u64, vp, vp, u64]
        L.amie_b200_set_values.argtypes = [vp, vp]
        L.amie_b200_set_block_map.argtypes = [vp, vp]
        L.amie_b200_pcg.argtypes = [vp, vp, vp, u64, ci, f64, ci, u64, u64, u64, vp, vp, vp, vp]
        L.amie_b200_bicgstab.argtypes = [vp, vp, vp, u64, ci, f64, ci, vp, vp, vp]
        L.amie_b200_spmv.argtypes = [vp, vp, vp, u64, u64, vp]
        L.amie_b200_inverse_diagonal.argtypes = [vp, vp]
        L.amie_b200_preconditioner_diagonal.argtypes = [vp, ci, vp]
        L.amie_b200_preconditioner_blocks.argtypes = [vp, ci, vp]
        L.amie_b200_set_preconditioner_diagonal.argtypes = [vp, vp]
        L.amie_b200_residual.argtypes = [vp, vp, vp, vp, vp]
        L.amie_b200_upload_rhs.argtypes = [vp, vp]
        L.amie_b200_upload_x0.argtypes = [vp, vp, u64]
        L.amie_b200_download_x.argtypes = [vp, vp]
        L.amie_b200_download_rhs.argtypes = [vp, vp]
        L.amie_b200_download_matrix.argtypes = [vp, vp, vp, vp]
        L.amie_b200_download_vector.argtypes = [vp, ci, vp]
        L.amie_b200_pcg_resident.argtypes = [vp, ci, f64, ci, u64, u64, u64, vp, vp, vp]
        L.amie_b200_bicgstab_resident.argtypes = [vp, ci, f64, ci, vp, vp]
        L.amie_b200_spmv_resident.argtypes = [vp, ci, ci, vp]
        L.amie_b200_cgsolve_resident.argtypes = [vp, ci, f64, u64, u64, u64, f64, vp, vp, vp]
        L.amie_b200_extrapolate.argtypes = [vp, f64, vp, vp]
        L.amie_b200_push_history.argtypes = [vp]
        L.amie_b200_reset_history.argtypes = [vp]
        L.amie_b200_get_stats.argtypes = [vp, vp]
        L.amie_b200_set_elements.argtypes = [vp, u64, ci, vp]
        L.amie_b200_update_elements.argtypes = [vp, u64, u64, vp, vp]
        L.amie_b200_assemble.argtypes = [vp]
        L.amie_b200_set_boundary_conditions.argtypes = [vp, u64, vp, vp, u64, vp, vp, vp, vp]
        L.amie_b200_set_element_kinematics.argtypes = [vp, u64, ci, ci, vp, vp, vp]
        L.amie_b200_set_element_behaviour.argtypes = [vp, u64, vp, vp, vp, vp]
        L.amie_b200_element_fields.argtypes = [vp, vp, u64, vp, vp, vp]
        L.amie_b200_element_principal.argtypes = [vp, ci, vp]
        L.amie_b200_set_option.argtypes = [vp, cp, ctypes.c_int64]
        L.amie_b200_synth_create.restype = vp
        L.amie_b200_synth_create.argtypes = [cp, ci, u64]
        L.amie_b200_synth_destroy.argtypes = [vp]
        L.amie_b200_synth_sizes.argtypes = [vp, vp, vp, vp]
        L.amie_b200_synth_count.argtypes = [vp, u64, u64, vp, vp]
        L.amie_b200_synth_fill.argtypes = [vp, u64, u64, vp, vp, vp]
        L.amie_b200_synth_to_device.argtypes = [vp, vp]
        L.amie_b200_partition_rows.argtypes = [u64, vp, ci, vp]
        L.amie_b200_partition_halo.argtypes = [u64, u64, vp, vp, vp, vp]
        L.amie_b200_rcm_order.argtypes = [u64, vp, vp, vp]
        L.amie_b200_group_rows_by_length.argtypes = [u64, vp, u64, vp]
        L.amie_b200_permute_structure.argtypes = [u64, vp, vp, vp, vp, vp, vp]
        L.amie_b200_nccl_unique_id.argtypes = [vp]
        L.amie_b200_dist_init.argtypes = [vp, ci, ci, vp, vp]
        L.amie_b200_dist_set_structure.argtypes = [vp, ci, u64, vp, vp, u64]
        L.amie_b200_dist_synth_to_device.argtypes = [vp, vp]
        L.amie_b200_dist_info.argtypes = [vp, vp, vp, vp, vp]
        L.amie_b200_dist_transport.argtypes = [vp]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- reference-shaped host objects

class CoordinateIndexedSparseMatrix:
    """Block-CSR in the reference's layout: stride, array (padded column-major blocks),
    column_index, row_size, accumulated_row_size (sparse/sparse_matrix.h:129-136)."""

    def __init__(self, row_size, column_index, stride, array=None):
        self.stride = int(stride)
        self.row_size = np.ascontiguousarray(row_size, np.uint32)
        self.column_index = np.ascontiguousarray(column_index, np.uint32)
        cl = self.stride + self.stride % 2
        n = self.column_index.size * self.stride * cl
        self.array = np.zeros(n) if array is None else np.ascontiguousarray(array, np.float64)
        if self.array.size != n:
            raise ValueError("array size does not match column_index.size*stride*(stride+stride%2)")
        self.accumulated_row_size = np.concatenate([[0], np.cumsum(self.row_size[:-1], dtype=np.uint64)]).astype(np.uint32) \
            if self.row_size.size else np.zeros(0, np.uint32)


class Assembly:
    """The two things the solvers pull from an Assembly: getMatrix() and getForces()
    (solvers/assembly.cpp:90-96), plus the knobs Assembly::cgsolve forwards
    (nssor, rowstart, colstart, epsilon; solvers/assembly.cpp:1829-1850).
    Owns the device context (one per Assembly)."""

    def __init__(self, matrix=None, forces=None, device=None, renumber=False, devices=None):
        """devices=[0, 1, ...]: ONE context over several GPUs (amie_b200_create(devices, ndev > 1)): the same calls with
        the same global host arrays, the block rows partitioned inside the library (csrc/group.cu).
        renumber=True: the device works on the matrix renumbered by reverse Cuthill-McKee (the mesher's numbering has
        no locality); this object keeps the caller's numbering and permutes b, x0, x and the other host vectors at the
        boundary, as host/shim does under AMIE_B200_RENUMBER=1.  The *_resident calls, the assembly and field rows then
        see the DEVICE numbering (use `self.perm`: perm[old node] = new node)."""
        self.renumber = bool(renumber)
        self.perm = None
        self.coordinateIndexedMatrix = matrix
        self.externalForces = None if forces is None else np.ascontiguousarray(forces, np.float64)
        self.displacements = np.zeros(0)
        self.nssor = 32
        self.rowstart = 0
        self.colstart = 0
        self.epsilon = default_solver_precision
        self._device = device
        self._devices = None if devices is None else [int(d) for d in devices]
        self._ctx = None
        self._structure_key = None
        self._values_dirty = True

    # -- device context
    @property
    def ctx(self):
        if self._ctx is None:
            L = lib()
            if self._devices is not None:
                d = (ctypes.c_int * len(self._devices))(*self._devices)
                h = L.amie_b200_create(d, len(self._devices))
            elif self._device is None:
                h = L.amie_b200_create(None, 0)
            else:
                d = (ctypes.c_int * 1)(int(self._device))
                h = L.amie_b200_create(d, 1)
            if not h:
                raise AmieB200Error(ERR_CUDA, L.amie_b200_global_error().decode())
            self._ctx = ctypes.c_void_p(h)
        return self._ctx

    def close(self):
        if self._ctx is not None:
            lib().amie_b200_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc < 0:
            raise AmieB200Error(rc, lib().amie_b200_last_error(self.ctx).decode())
        return rc

    def getMatrix(self):
        return self.coordinateIndexedMatrix

    def getForces(self):
        return self.externalForces

    def setEpsilon(self, e):
        self.epsilon = e

    def set_option(self, key, value):
        self.check(lib().amie_b200_set_option(self.ctx, key.encode(), int(value)))

    def values_changed(self):
        """Tell the context the matrix values were re-assembled (every damage step)."""
        self._values_dirty = True

    def sync_matrix(self):
        """Upload structure once per topology, values whenever they changed."""
        A = self.coordinateIndexedMatrix
        L = lib()
        if A is None:
            return          # matrix was generated on the device (Synth.to_device)
        key = (A.stride, A.row_size.size, A.column_index.size, A.column_index.ctypes.data)
        if key != self._structure_key:
            if self.renumber:
                self.perm = rcm_order(A.row_size, A.column_index)
                rs2, ci2, frm = permute_structure(A.row_size, A.column_index, self.perm)
                block_to = np.empty_like(frm)
                block_to[frm] = np.arange(frm.size, dtype=np.uint32)
                self.check(L.amie_b200_set_structure(self.ctx, A.stride, rs2.size, _ptr(rs2), _ptr(ci2), ci2.size))
                self.check(L.amie_b200_set_block_map(self.ctx, _ptr(block_to)))
            else:
                self.check(L.amie_b200_set_structure(self.ctx, A.stride, A.row_size.size, _ptr(A.row_size),
                                                     _ptr(A.column_index), A.column_index.size))
            self._structure_key = key
            self._values_dirty = True
        if self._values_dirty:
            self.check(L.amie_b200_set_values(self.ctx, _ptr(A.array)))
            self._values_dirty = False

    # ---- host vectors across the renumbering (identity when renumber is off)
    def to_device_order(self, v):
        if v is None or self.perm is None:
            return v
        v = np.ascontiguousarray(v, np.float64)
        s = self.coordinateIndexedMatrix.stride
        n = self.perm.size * s
        full = np.zeros(n)
        full[:min(v.size, n)] = v[:n]          # a shorter x0 is a prefix in the caller's numbering (conjugategradient.cpp:100-104)
        out = np.empty(n)
        out.reshape(-1, s)[self.perm] = full.reshape(-1, s)
        return out

    def from_device_order(self, v):
        if self.perm is None:
            return v
        s = self.coordinateIndexedMatrix.stride
        return np.ascontiguousarray(v.reshape(-1, s)[self.perm].reshape(-1))

    # ---- device-side value assembly + Dirichlet elimination
    def set_structure_only(self, stride, row_size, column_index):
        """Topology without host values: the values will be assembled on the device from the elements."""
        rs = np.ascontiguousarray(row_size, np.uint32)
        ci = np.ascontiguousarray(column_index, np.uint32)
        self.coordinateIndexedMatrix = None
        self.check(lib().amie_b200_set_structure(self.ctx, int(stride), rs.size, _ptr(rs), _ptr(ci), ci.size))

    def set_elements(self, elem_ids):
        """elem_ids[n_elem, npe] node (block-row) ids in Assembly::element2d/element3d order; 0xFFFFFFFF = unused slot."""
        ids = np.ascontiguousarray(elem_ids, np.uint32)
        if ids.ndim != 2:
            raise ValueError("elem_ids must be [n_elem, npe]")
        self._npe = ids.shape[1]
        self.check(lib().amie_b200_set_elements(self.ctx, ids.shape[0], ids.shape[1], _ptr(ids)))

    def update_elements(self, first, ke, scales=None):
        """Elementary matrices of elements [first, first+len(ke)): ke[e, j, k, m*s+n] = Ke_e[j][k][n][m]."""
        ke = np.ascontiguousarray(ke, np.float64)
        count = ke.shape[0]
        scales = None if scales is None else np.ascontiguousarray(scales, np.float64)
        if scales is not None and scales.size != count:
            raise ValueError("one scale per element")
        self.check(lib().amie_b200_update_elements(self.ctx, int(first), int(count), _ptr(ke) if count else None, _ptr(scales)))

    def assemble(self):
        """Assembly::make_final's scatter loops on the device (solvers/assembly.cpp:657-735, :1060-1138)."""
        self.check(lib().amie_b200_assemble(self.ctx))

    def set_boundary_conditions(self, fix_ids, fix_values, force_ids=None, force_values=None, add_to_forces=None,
                                natural=None):
        """Assembly::setBoundaryConditions on the resident matrix and forces (solvers/assembly.cpp:125-330).
        `natural` (naturalBoundaryConditionForces) is updated in place when given."""
        fi = np.ascontiguousarray(fix_ids, np.uint32)
        fv = np.ascontiguousarray(fix_values, np.float64)
        gi = np.ascontiguousarray([] if force_ids is None else force_ids, np.uint32)
        gv = np.ascontiguousarray([] if force_values is None else force_values, np.float64)
        if fi.size != fv.size or gi.size != gv.size:
            raise ValueError("one value per dof id")
        add = None if add_to_forces is None else np.ascontiguousarray(add_to_forces, np.float64)
        if natural is not None and (natural.dtype != np.float64 or not natural.flags.c_contiguous):
            raise ValueError("natural must be a contiguous float64 array (updated in place)")
        self.check(lib().amie_b200_set_boundary_conditions(self.ctx, fi.size, _ptr(fi) if fi.size else None,
                                                           _ptr(fv) if fv.size else None, gi.size,
                                                           _ptr(gi) if gi.size else None, _ptr(gv) if gv.size else None,
                                                           _ptr(add), _ptr(natural)))

    # ---- field recovery after the solve
    def set_element_kinematics(self, dim, elem_ids, dshape, jinv):
        """Once per topology: elem_ids[n_elem, npe] (IntegrableEntity::getDofIds order, 0xFFFFFFFF = unused slot),
        dshape[n_elem, npe, dim] = vm.deval(function j, XI|ETA|ZETA, p), jinv[n_elem, dim, dim] = ElementState::JinvCache."""
        ids = np.ascontiguousarray(elem_ids, np.uint32)
        ds = np.ascontiguousarray(dshape, np.float64)
        ji = np.ascontiguousarray(jinv, np.float64)
        if ids.ndim != 2 or ds.shape != ids.shape + (dim,) or ji.shape != (ids.shape[0], dim, dim):
            raise ValueError("elem_ids [n_elem, npe], dshape [n_elem, npe, dim], jinv [n_elem, dim, dim]")
        self._field_shape = (ids.shape[0], 3 if dim == 2 else 6)
        self.check(lib().amie_b200_set_element_kinematics(self.ctx, ids.shape[0], ids.shape[1], int(dim), _ptr(ids), _ptr(ds), _ptr(ji)))

    def set_element_behaviour(self, tensors, imposed_strain=None, imposed_stress=None, tensor_of_elem=None):
        """A table of behaviours (getTensor / getImposedStrain / getImposedStress at the evaluation point) and the
        table entry of every element (None: one entry per element, in element order)."""
        C = np.ascontiguousarray(tensors, np.float64)
        nc = self._field_shape[1]
        if C.ndim != 3 or C.shape[1:] != (nc, nc):
            raise ValueError(f"tensors must be [n_tensors, {nc}, {nc}]")
        opt = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
        es, ss, toe = opt(imposed_strain, np.float64), opt(imposed_stress, np.float64), opt(tensor_of_elem, np.uint32)
        for a in (es, ss):
            if a is not None and a.shape != (C.shape[0], nc):
                raise ValueError(f"imposed strain / stress must be [n_tensors, {nc}]")
        if toe is not None and toe.shape != (self._field_shape[0],):
            raise ValueError("tensor_of_elem must be [n_elem]")
        self.check(lib().amie_b200_set_element_behaviour(self.ctx, C.shape[0], _ptr(C), _ptr(es), _ptr(ss), _ptr(toe)))

    def element_fields(self, u=None):
        """(TOTAL_STRAIN_FIELD, MECHANICAL_STRAIN_FIELD, REAL_STRESS_FIELD) per element, from the resident solution
        of the last solve (u=None) or from a host vector: ElementState::step + getField
        (elements/integrable_entity.cpp:3607-3667, :964-1104, :1379-1392)."""
        u = None if u is None else np.ascontiguousarray(u, np.float64)
        out = [np.zeros(self._field_shape) for _ in range(3)]
        self.check(lib().amie_b200_element_fields(self.ctx, _ptr(u), 0 if u is None else u.size, *[_ptr(o) for o in out]))
        return tuple(out)

    def element_principal(self, field):
        """Principal values of field 0 (total strain), 1 (mechanical strain) or 2 (real stress) of the last
        element_fields(): getField(PRINCIPAL_*_FIELD), elements/integrable_entity.cpp:475-596."""
        ne, nc = self._field_shape
        out = np.zeros((ne, 2 if nc == 3 else 3))
        self.check(lib().amie_b200_element_principal(self.ctx, int(field), _ptr(out)))
        return out

    def stats(self):
        s = Stats()
        self.check(lib().amie_b200_get_stats(self.ctx, ctypes.byref(s)))
        return s

    def download_matrix(self):
        """(row_size, column_index, array_padded, b) as held on the device (tests)."""
        st = self.stats()
        cl = st.stride + st.stride % 2
        rs = np.zeros(st.nb, np.uint32)
        ci = np.zeros(st.nnzb, np.uint32)
        arr = np.zeros(st.nnzb * st.stride * cl)
        b = np.zeros(st.ndof)
        self.check(lib().amie_b200_download_matrix(self.ctx, _ptr(rs), _ptr(ci), _ptr(arr)))
        self.check(lib().amie_b200_download_rhs(self.ctx, _ptr(b)))
        return rs, ci, arr, b

    # -- device-resident calls (no host<->device copies inside)
    def upload_rhs(self, b):
        b = np.ascontiguousarray(b, np.float64)
        self.check(lib().amie_b200_upload_rhs(self.ctx, _ptr(b)))

    def upload_x0(self, x0=None):
        x0 = np.zeros(0) if x0 is None else np.ascontiguousarray(x0, np.float64)
        self.check(lib().amie_b200_upload_x0(self.ctx, _ptr(x0) if x0.size else None, x0.size))

    def download_vector(self, which):
        v = np.zeros(self.stats().ndof)
        self.check(lib().amie_b200_download_vector(self.ctx, int(which), _ptr(v)))
        return v

    def download_rhs(self):
        b = np.zeros(self.stats().ndof)
        self.check(lib().amie_b200_download_rhs(self.ctx, _ptr(b)))
        return b

    def download_x(self):
        x = np.zeros(self.stats().ndof)
        self.check(lib().amie_b200_download_x(self.ctx, _ptr(x)))
        return x

    def pcg_resident(self, precond=PRECOND_JACOBI, eps=default_solver_precision, maxit=-1, nssor=32, rowstart=0, colstart=0):
        nit, err, rho = u64(), f64(), f64()
        rc = self.check(lib().amie_b200_pcg_resident(self.ctx, precond, eps, int(maxit), int(nssor), int(rowstart), int(colstart),
                                                     ctypes.byref(nit), ctypes.byref(err), ctypes.byref(rho)))
        return bool(rc), nit.value, err.value, rho.value

    def bicgstab_resident(self, precond=PRECOND_JACOBI, eps=default_solver_precision, maxit=-1):
        nit, err = u64(), f64()
        rc = self.check(lib().amie_b200_bicgstab_resident(self.ctx, precond, eps, int(maxit), ctypes.byref(nit), ctypes.byref(err)))
        return bool(rc), nit.value, err.value

    # -- Assembly::cgsolve with displacementHistory in HBM
    def cgsolve_resident(self, precond=PRECOND_JACOBI, factor=1.0):
        """x0 = extrapolate(factor), ConjugateGradient::solve with this assembly's epsilon / nssor / rowstart /
        colstart, history update -- all on the device (solvers/assembly.cpp:1841-1868).  (converged, nit, err, rho)."""
        nit, err, rho = u64(), f64(), f64()
        rc = self.check(lib().amie_b200_cgsolve_resident(self.ctx, int(precond), self.epsilon, int(self.nssor), int(self.rowstart),
                                                         int(self.colstart), float(factor), ctypes.byref(nit), ctypes.byref(err),
                                                         ctypes.byref(rho)))
        return bool(rc), nit.value, err.value, rho.value

    def extrapolate(self, factor=1.0):
        """Assembly::extrapolate (solvers/assembly.cpp:1772-1814) into the resident x: (x0, case) with case 0 = no
        history (x as it was), 1 = extrapolated, 2 = size mismatch (history cleared, zeros)."""
        x0 = np.zeros(self.stats().ndof)
        case = ctypes.c_int()
        self.check(lib().amie_b200_extrapolate(self.ctx, float(factor), _ptr(x0), ctypes.byref(case)))
        return x0, case.value

    def push_history(self):
        """The displacementHistory update of cgsolve (:1859-1868) with the resident x as `displacements`."""
        self.check(lib().amie_b200_push_history(self.ctx))

    def reset_history(self):
        self.check(lib().amie_b200_reset_history(self.ctx))

    def spmv_resident(self, reps=10, variant=0):
        ms = f64()
        self.check(lib().amie_b200_spmv_resident(self.ctx, int(reps), int(variant), ctypes.byref(ms)))
        return ms.value

    # -- row-partitioned context (one process per GPU)
    def dist_init(self, rank, world, id128, bounds):
        _preload_nccl()
        self._bounds = np.ascontiguousarray(bounds, np.uint64)
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(id128))
        self.check(lib().amie_b200_dist_init(self.ctx, int(rank), int(world), ctypes.cast(buf, ctypes.c_void_p), _ptr(self._bounds)))

    def dist_set_structure(self, stride, nb_global, row_size_local, column_index_local):
        rs = np.ascontiguousarray(row_size_local, np.uint32)
        ci = np.ascontiguousarray(column_index_local, np.uint32)
        self.check(lib().amie_b200_dist_set_structure(self.ctx, int(stride), int(nb_global), _ptr(rs), _ptr(ci), ci.size))

    def set_values(self, array_padded):
        arr = np.ascontiguousarray(array_padded, np.float64)
        self.check(lib().amie_b200_set_values(self.ctx, _ptr(arr)))

    def dist_synth_to_device(self, synth):
        self.check(lib().amie_b200_dist_synth_to_device(self.ctx, synth.handle))

    def dist_info(self):
        nh, ns, ni, npeer = u64(), u64(), u64(), ctypes.c_int()
        self.check(lib().amie_b200_dist_info(self.ctx, ctypes.byref(nh), ctypes.byref(ns), ctypes.byref(ni), ctypes.byref(npeer)))
        return dict(halo=nh.value, send=ns.value, interior_rows=ni.value, peers=npeer.value,
                    transport="peer" if lib().amie_b200_dist_transport(self.ctx) == 1 else "nccl")

    def spmv(self, x, minus_b=None, rowstart=0, colstart=0):
        """assign(y, A*x [- b], rowstart, colstart)"""
        self.sync_matrix()
        if self.perm is not None and (rowstart or colstart):
            raise AmieB200Error(ERR_UNSUPPORTED, "rowstart / colstart address the caller's numbering: not with renumber=True")
        x = self.to_device_order(np.ascontiguousarray(x, np.float64))
        y = np.zeros_like(x)
        b = None if minus_b is None else self.to_device_order(np.ascontiguousarray(minus_b, np.float64))
        self.check(lib().amie_b200_spmv(self.ctx, _ptr(x), _ptr(b), rowstart, colstart, _ptr(y)))
        return self.from_device_order(y)

    def residual(self, u, f=None):
        """r = K u - f and |r| (features/features.cpp:4766-4768)."""
        self.sync_matrix()
        u = self.to_device_order(np.ascontiguousarray(u, np.float64))
        f = self.to_device_order(self.getForces() if f is None else np.ascontiguousarray(f, np.float64))
        r = np.zeros_like(u)
        nrm = f64()
        self.check(lib().amie_b200_residual(self.ctx, _ptr(u), _ptr(f), _ptr(r), ctypes.byref(nrm)))
        return self.from_device_order(r), nrm.value

    def inverse_diagonal(self):
        self.sync_matrix()
        d = np.zeros(self.coordinateIndexedMatrix.row_size.size * self.coordinateIndexedMatrix.stride)
        self.check(lib().amie_b200_inverse_diagonal(self.ctx, _ptr(d)))
        return self.from_device_order(d)

    def preconditioner_diagonal(self, kind):
        """The `diagonal` member of the reference's preconditioner class `kind` (PRECOND_JACOBI, _DIAGONAL_SQUARED,
        _LUMPED), built on the device from the resident values."""
        self.sync_matrix()
        d = np.zeros(self.stats().ndof)
        self.check(lib().amie_b200_preconditioner_diagonal(self.ctx, int(kind), _ptr(d)))
        return self.from_device_order(d)

    def preconditioner_blocks(self, kind):
        """The s x s blocks of PRECOND_BLOCK2X2 / PRECOND_BLOCK3X3 ([nb, s, s], row-major: Inverse2x2Diagonal::blocks)."""
        self.sync_matrix()
        st = self.stats()
        B = np.zeros((st.nb, st.stride, st.stride))
        self.check(lib().amie_b200_preconditioner_blocks(self.ctx, int(kind), _ptr(B)))
        return B

    def cgsolve(self, maxit=-1, verbose=False):
        """Assembly::cgsolve for an assembled symmetric system (solvers/assembly.cpp:1841-1858)."""
        cg = ConjugateGradient(self)
        cg.nssor = self.nssor
        if self.rowstart > 0 or self.colstart > 0:
            cg.rowstart, cg.colstart = self.rowstart, self.colstart
        ret = cg.solve(self.displacements, None, self.epsilon, -1, verbose)
        self.displacements = cg.x
        return ret


class Preconditionner:
    """solvers/preconditionners.h:19-26.  `kind` is what the C-ABI is told; `diagonal` (kind PRECOND_DIAGONAL) is
    uploaded before the solve."""
    kind = None
    diagonal = None


class NullPreconditionner(Preconditionner):
    kind = PRECOND_NULL


class InverseDiagonalSquared(Preconditionner):
    """solvers/inversediagonal.h:41-47; built on the device from the assembly's matrix."""
    kind = PRECOND_DIAGONAL_SQUARED

    def __init__(self, A=None):
        pass


class InverseLumpedDiagonal(Preconditionner):
    """solvers/inversediagonal.h:32-38; built on the device from the assembly's matrix."""
    kind = PRECOND_LUMPED

    def __init__(self, A=None):
        pass


class Inverse2x2Diagonal(Preconditionner):
    """solvers/inversediagonal.h:49-55 on a stride-2 system: the inverse of every node's 2x2 diagonal block, built on
    the device from the assembly's matrix with the reference's arithmetic (PCG only)."""
    kind = PRECOND_BLOCK2X2

    def __init__(self, A=None):
        pass


class BlockJacobi3x3(Preconditionner):
    """The 3x3 counterpart for stride-3 systems (det / invert3x3Matrix of utilities/matrixops.cpp).  The reference has
    no such class: opt-in, iteration counts differ from the Jacobi default (PCG only)."""
    kind = PRECOND_BLOCK3X3

    def __init__(self, A=None):
        pass


class DiagonalPreconditionner(Preconditionner):
    """A user-written preconditioner whose precondition(v, t) is t = v * diagonal (also what an explicitly
    constructed InverseDiagonal / InverseDiagonalSquared / InverseLumpedDiagonal object amounts to: the shim
    uploads the object's own `diagonal`)."""
    kind = PRECOND_DIAGONAL

    def __init__(self, diagonal):
        self.diagonal = np.ascontiguousarray(diagonal, np.float64)


class LinearSolver:
    def __init__(self, assembly):
        self.assembly = assembly
        self.rowstart = 0
        self.colstart = 0
        n = 0 if assembly.getForces() is None else assembly.getForces().size
        self.x = np.zeros(n)          # Solver::Solver, solvers/solver.cpp:15

    def _kind(self, precond):
        if precond is None:
            return PRECOND_JACOBI
        if isinstance(precond, Preconditionner) and precond.kind is not None:
            if precond.kind == PRECOND_DIAGONAL:
                d = precond.diagonal
                if d is None or d.size != self.assembly.getForces().size:
                    raise ValueError("DiagonalPreconditionner: one entry per degree of freedom")
                d = self.assembly.to_device_order(d)
                self.assembly.check(lib().amie_b200_set_preconditioner_diagonal(self.assembly.ctx, _ptr(d)))
            return precond.kind
        raise AmieB200Error(ERR_UNSUPPORTED, "only diagonal preconditioners (nullptr -> InverseDiagonal, InverseDiagonalSquared, "
                            "InverseLumpedDiagonal, a user diagonal) and NullPreconditionner run on the device")


class ConjugateGradient(LinearSolver):
    def __init__(self, assembly):
        super().__init__(assembly)
        self.nit = 0
        self.nssor = 128              # conjugategradient.h:34
        self.last_error = 0.0
        self.last_rho = 0.0

    def solve(self, x0=None, precond=None, eps=default_solver_precision, maxit=-1, verbose=False):
        A = self.assembly
        A.sync_matrix()
        if verbose:
            A.set_option("verbose", 1)
        if A.perm is not None and (self.rowstart or self.colstart):
            raise AmieB200Error(ERR_UNSUPPORTED, "rowstart / colstart address the caller's numbering: not with renumber=True")
        b = A.to_device_order(A.getForces())
        x0 = np.zeros(0) if x0 is None else np.ascontiguousarray(x0, np.float64)
        if x0.size:
            x0 = A.to_device_order(x0)
        x = np.zeros(b.size)
        nit, err, rho = u64(), f64(), f64()
        rc = lib().amie_b200_pcg(A.ctx, _ptr(b), _ptr(x0) if x0.size else None, x0.size, self._kind(precond),
                                 eps, int(maxit), int(self.nssor), int(self.rowstart), int(self.colstart),
                                 _ptr(x), ctypes.byref(nit), ctypes.byref(err), ctypes.byref(rho))
        A.check(rc)
        self.x, self.nit, self.last_error, self.last_rho = A.from_device_order(x), nit.value, err.value, rho.value
        return bool(rc)


class BiConjugateGradientStabilized(LinearSolver):
    def __init__(self, assembly):
        super().__init__(assembly)
        self.nit = 0                  # the reference keeps it local; exposed here for the tests
        self.last_error = 0.0

    def solve(self, x0=None, precond=None, eps=default_solver_precision, maxit=-1, verbose=False):
        A = self.assembly
        A.sync_matrix()
        if verbose:
            A.set_option("verbose", 1)
        b = A.to_device_order(A.getForces())
        x0 = np.zeros(0) if x0 is None else np.ascontiguousarray(x0, np.float64)
        if x0.size == b.size:
            x0 = A.to_device_order(x0)         # (any other size is ignored by the reference, :21-24)
        x = np.zeros(b.size)
        nit, err = u64(), f64()
        rc = lib().amie_b200_bicgstab(A.ctx, _ptr(b), _ptr(x0) if x0.size else None, x0.size, self._kind(precond),
                                      eps, int(maxit), _ptr(x), ctypes.byref(nit), ctypes.byref(err))
        A.check(rc)
        self.x, self.nit, self.last_error = A.from_device_order(x), nit.value, err.value
        return bool(rc)


# --------------------------------------------------------------------------- synthetic structured meshes (host-only)

class Synth:
    """S3-hex / S3-tet / S2-tri / ASR-hex systems of SURVEY.md §8(d) in the reference layout."""

    def __init__(self, preset, n, seed=1):
        self.preset, self.n = preset, int(n)
        h = lib().amie_b200_synth_create(preset.encode(), int(n), int(seed))
        if not h:
            raise ValueError(f"unknown synthetic preset {preset!r} or n < 2")
        self.handle = ctypes.c_void_p(h)
        st, nb = ctypes.c_int(), u64()
        lib().amie_b200_synth_sizes(self.handle, ctypes.byref(st), ctypes.byref(nb), None)
        self.stride, self.nb = st.value, nb.value

    def __del__(self):
        try:
            lib().amie_b200_synth_destroy(self.handle)
        except Exception:
            pass

    def row_sizes(self, row0=0, row1=None):
        row1 = self.nb if row1 is None else row1
        rs = np.zeros(row1 - row0, np.uint32)
        tot = u64()
        lib().amie_b200_synth_count(self.handle, row0, row1, _ptr(rs), ctypes.byref(tot))
        return rs, tot.value

    def rows(self, row0=0, row1=None):
        """(row_size, column_index (global), array_padded, b) of block rows [row0,row1)."""
        row1 = self.nb if row1 is None else row1
        rs, nnzb = self.row_sizes(row0, row1)
        cl = self.stride + self.stride % 2
        ci = np.zeros(nnzb, np.uint32)
        arr = np.zeros(nnzb * self.stride * cl)
        b = np.zeros((row1 - row0) * self.stride)
        rc = lib().amie_b200_synth_fill(self.handle, row0, row1, _ptr(ci), _ptr(arr), _ptr(b))
        if rc:
            raise AmieB200Error(rc, "synth_fill")
        return rs, ci, arr, b

    def assembly(self, device=None):
        rs, ci, arr, b = self.rows()
        return Assembly(CoordinateIndexedSparseMatrix(rs, ci, self.stride, arr), b, device=device)

    def to_device(self, assembly):
        """Generate structure + values + rhs directly in HBM (no host arrays)."""
        assembly.check(lib().amie_b200_synth_to_device(assembly.ctx, self.handle))


_nccl_preloaded = False


def _preload_nccl():
    """Python hosts usually carry torch, whose libtorch_cuda.so needs the NCCL wheel it was built with
    (nvidia/nccl/lib/libnccl.so.2).  ld.so shares libraries by SONAME, so whichever libnccl.so.2 enters the
    process first serves both torch and this library: load the wheel's copy first (no torch import needed),
    else leave the choice to the C loader (AMIE_B200_NCCL_LIB, then the system libnccl.so.2)."""
    global _nccl_preloaded
    if _nccl_preloaded or os.environ.get("AMIE_B200_NCCL_LIB"):
        return
    _nccl_preloaded = True
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
            return


def nccl_unique_id():
    _preload_nccl()
    buf = ctypes.create_string_buffer(128)
    rc = lib().amie_b200_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p))
    if rc:
        raise AmieB200Error(rc, "ncclGetUniqueId (is libnccl.so.2 loadable?)")
    return bytes(buf.raw)


def partition_rows(row_size, nparts):
    row_size = np.ascontiguousarray(row_size, np.uint32)
    bounds = np.zeros(nparts + 1, np.uint64)
    rc = lib().amie_b200_partition_rows(row_size.size, _ptr(row_size), nparts, _ptr(bounds))
    if rc:
        raise AmieB200Error(rc, "partition_rows")
    return bounds


def partition_halo(r0, r1, row_size_local, column_index_local):
    rs = np.ascontiguousarray(row_size_local, np.uint32)
    ci = np.ascontiguousarray(column_index_local, np.uint32)
    n = u64()
    lib().amie_b200_partition_halo(int(r0), int(r1), _ptr(rs), _ptr(ci), None, ctypes.byref(n))
    halo = np.zeros(n.value, np.uint32)
    lib().amie_b200_partition_halo(int(r0), int(r1), _ptr(rs), _ptr(ci), _ptr(halo), ctypes.byref(n))
    return halo


def rcm_order(row_size, column_index):
    """perm[old node] = new node: reverse Cuthill-McKee on the block graph (host only)."""
    rs = np.ascontiguousarray(row_size, np.uint32)
    ci = np.ascontiguousarray(column_index, np.uint32)
    perm = np.zeros(rs.size, np.uint32)
    rc = lib().amie_b200_rcm_order(rs.size, _ptr(rs), _ptr(ci), _ptr(perm))
    if rc:
        raise AmieB200Error(rc, "rcm_order")
    return perm


def group_rows_by_length(row_size, perm, window):
    """perm refined: inside windows of `window` consecutive nodes of the numbering perm, nodes ordered by row length
    (longest first); row_size in the original numbering (host only, opt-in: csrc/reorder.cpp)."""
    rs = np.ascontiguousarray(row_size, np.uint32)
    out = np.array(perm, np.uint32)
    rc = lib().amie_b200_group_rows_by_length(rs.size, _ptr(rs), int(window), _ptr(out))
    if rc:
        raise AmieB200Error(rc, "group_rows_by_length: perm is not a permutation of the nodes, or window == 0")
    return out


def permute_structure(row_size, column_index, perm):
    """(row_size, column_index, block_from) of the renumbered structure; block_from[new k] = old k (host only)."""
    rs = np.ascontiguousarray(row_size, np.uint32)
    ci = np.ascontiguousarray(column_index, np.uint32)
    perm = np.ascontiguousarray(perm, np.uint32)
    rs2, ci2, frm = np.zeros_like(rs), np.zeros_like(ci), np.zeros_like(ci)
    rc = lib().amie_b200_permute_structure(rs.size, _ptr(rs), _ptr(ci), _ptr(perm), _ptr(rs2), _ptr(ci2), _ptr(frm))
    if rc:
        raise AmieB200Error(rc, "permute_structure: perm is not a permutation of the nodes")
    return rs2, ci2, frm

