"""ctypes bindings of the CHECKERS (test infrastructure only):

  * oracle/libamie_oracle.so          -- the C restatement (oracle/amie_oracle.c), always available;
  * oracle/_ref/libamie_ref_oracle.so -- the unmodified reference compiled from /root/reference by
                                         oracle/build_ref.py (present where it was prebuilt).

Nothing in the product package imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
u64 = ctypes.c_uint64
f64 = ctypes.c_double


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class CgInfo(ctypes.Structure):
    _fields_ = [("nit", u64), ("err", f64), ("last_rho", f64), ("spmv", u64), ("restarts", u64),
                ("status", ctypes.c_int)]


class BicgInfo(ctypes.Structure):
    _fields_ = [("nit", u64), ("err", f64), ("rho", f64), ("spmv", u64)]


_oracle = None
_ref = None


def oracle():
    """The C restatement; built on demand (gcc only)."""
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "libamie_oracle.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith(".c")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libamie_oracle.so"], stdout=subprocess.DEVNULL)
        _oracle = ctypes.CDLL(so)
        _oracle.amie_oracle_dot.restype = f64
    return _oracle


def ref():
    """The real reference solver, or None when oracle/_ref was not prebuilt."""
    global _ref
    if _ref is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libamie_ref_oracle.so")
        if not os.path.exists(so):
            return None
        _ref = ctypes.CDLL(so)
    return _ref


_synth = None


def synth_lib():
    """oracle/libamie_synth.so: the synthetic-mesh generator as a host-only library (no CUDA library in the process)."""
    global _synth
    if _synth is None:
        so = os.path.join(ORACLE_DIR, "libamie_synth.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libamie_synth.so"], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.amie_b200_synth_create.restype = vp
        L.amie_b200_synth_create.argtypes = [ctypes.c_char_p, ci, u64]
        L.amie_b200_synth_destroy.argtypes = [vp]
        L.amie_b200_synth_sizes.argtypes = [vp, vp, vp, vp]
        L.amie_b200_synth_count.argtypes = [vp, u64, u64, vp, vp]
        L.amie_b200_synth_fill.argtypes = [vp, u64, u64, vp, vp, vp]
        _synth = L
    return _synth


def synth_system(preset, n, seed=1):
    """The S3-hex-n / S3-tet-n / S2-tri-n / ASR-hex-n system of SURVEY.md section 8(d) as a Sys, generated on the host."""
    L = synth_lib()
    h = L.amie_b200_synth_create(preset.encode(), int(n), int(seed))
    if not h:
        raise ValueError(f"unknown synthetic preset {preset!r}")
    h = ctypes.c_void_p(h)
    st, nb, tot = ctypes.c_int(), u64(), u64()
    L.amie_b200_synth_sizes(h, ctypes.byref(st), ctypes.byref(nb), None)
    rs = np.zeros(nb.value, np.uint32)
    L.amie_b200_synth_count(h, 0, nb.value, _vp(rs), ctypes.byref(tot))
    s = st.value
    ci = np.zeros(tot.value, np.uint32)
    arr = np.zeros(tot.value * s * (s + s % 2))
    b = np.zeros(nb.value * s)
    rc = L.amie_b200_synth_fill(h, 0, nb.value, _vp(ci), _vp(arr), _vp(b))
    L.amie_b200_synth_destroy(h)
    assert rc == 0
    return Sys(s, nb.value, rs, ci, arr, b)


class Sys:
    """A block-sparse system in the reference layout (numpy arrays)."""

    def __init__(self, stride, nb, row_size, column_index, array, b):
        self.stride, self.nb = int(stride), int(nb)
        self.row_size = np.ascontiguousarray(row_size, np.uint32)
        self.column_index = np.ascontiguousarray(column_index, np.uint32)
        self.array = np.ascontiguousarray(array, np.float64)
        self.b = np.ascontiguousarray(b, np.float64)
        self.nnzb = int(self.column_index.size)
        self.n = self.nb * self.stride

    def head(self):
        return (self.stride, u64(self.nb), _vp(self.row_size), _vp(self.column_index), u64(self.nnzb), _vp(self.array))

    def to_scipy(self):
        import scipy.sparse as sp
        s = self.stride
        cl = s + s % 2
        blocks = self.array.reshape(self.nnzb, s, cl)[:, :, :s].transpose(0, 2, 1).copy()
        indptr = np.concatenate([[0], np.cumsum(self.row_size, dtype=np.int64)])
        return sp.bsr_matrix((blocks, self.column_index.astype(np.int64), indptr), shape=(self.n, self.n)).tocsr()


# ------------------------------------------------------------------ C restatement

# precond kinds (include/amie_b200.h): 0 nullptr -> InverseDiagonal, 1 NullPreconditionner, 2 InverseDiagonalSquared,
# 3 InverseLumpedDiagonal, 4 a diagonal supplied by the caller (`diag`)
def oracle_cg(S, x0=None, precond=0, eps=1e-10, maxit=-1, nssor=32, rowstart=0, colstart=0, nthreads=1, b=None, diag=None):
    b = S.b if b is None else np.ascontiguousarray(b, np.float64)
    x = np.zeros(S.n)
    info = CgInfo()
    x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
    if precond == 4:
        diag = np.ascontiguousarray(diag, np.float64)
        ret = oracle().amie_oracle_cg_diag(*S.head(), _vp(b), _vp(x0), u64(0 if x0 is None else x0.size), _vp(diag),
                                           f64(eps), int(maxit), u64(nssor), u64(rowstart), u64(colstart), int(nthreads),
                                           _vp(x), ctypes.byref(info))
        return ret, x, info
    ret = oracle().amie_oracle_cg(*S.head(), _vp(b), _vp(x0), u64(0 if x0 is None else x0.size), int(precond),
                                  f64(eps), int(maxit), u64(nssor), u64(rowstart), u64(colstart), int(nthreads),
                                  _vp(x), ctypes.byref(info))
    return ret, x, info


def oracle_bicgstab(S, x0=None, precond=0, eps=1e-10, maxit=-1, nthreads=1, b=None, diag=None):
    b = S.b if b is None else np.ascontiguousarray(b, np.float64)
    x = np.zeros(S.n)
    info = BicgInfo()
    x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
    if precond == 4:
        diag = np.ascontiguousarray(diag, np.float64)
        ret = oracle().amie_oracle_bicgstab_diag(*S.head(), _vp(b), _vp(x0), u64(0 if x0 is None else x0.size), _vp(diag),
                                                 f64(eps), int(maxit), int(nthreads), _vp(x), ctypes.byref(info))
        return ret, x, info
    ret = oracle().amie_oracle_bicgstab(*S.head(), _vp(b), _vp(x0), u64(0 if x0 is None else x0.size), int(precond),
                                        f64(eps), int(maxit), int(nthreads), _vp(x), ctypes.byref(info))
    return ret, x, info


def oracle_precond_diagonal(S, kind):
    d = np.zeros(S.n)
    oracle().amie_oracle_precond_diagonal(*S.head(), int(kind), _vp(d))
    return d


def oracle_precond_blocks(S):
    """The s x s blocks of the block preconditioners (kind 5 on stride 2, kind 6 on stride 3), row-major per node."""
    B = np.zeros(S.nb * S.stride * S.stride)
    oracle().amie_oracle_precond_blocks(*S.head(), _vp(B))
    return B


def ref_precond_blocks2(S):
    """The 2x2 blocks the reference's Inverse2x2Diagonal builds (stride-2 systems: one per node)."""
    B = np.zeros(S.n // 2 * 4)
    n = ref().amie_ref_precond_blocks2(*S.head(), _vp(B))
    assert n == S.n // 2
    return B


def ref_det_invert3x3(m):
    m = np.ascontiguousarray(m, np.float64).reshape(-1, 9)
    det, inv = np.zeros(m.shape[0]), np.zeros_like(m)
    ref().amie_ref_det_invert3x3(_vp(m), u64(m.shape[0]), _vp(det), _vp(inv))
    return det, inv


def oracle_assign(S, v, b=None, rowstart=0, colstart=0):
    y = np.zeros(S.n)
    v = np.ascontiguousarray(v, np.float64)
    b = None if b is None else np.ascontiguousarray(b, np.float64)
    oracle().amie_oracle_assign(*S.head(), _vp(v), _vp(b), u64(rowstart), u64(colstart), _vp(y))
    return y


def oracle_spmv_serial(S, v, b=None):
    y = np.zeros(S.n)
    v = np.ascontiguousarray(v, np.float64)
    b = None if b is None else np.ascontiguousarray(b, np.float64)
    oracle().amie_oracle_spmv_serial(*S.head(), _vp(v), _vp(b), _vp(y))
    return y


def oracle_inverse_diagonal(S):
    d = np.zeros(S.n)
    oracle().amie_oracle_inverse_diagonal(*S.head(), _vp(d))
    return d


def oracle_dot(a, b, nthreads=1):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return oracle().amie_oracle_dot(_vp(a), _vp(b), ctypes.c_int64(a.size), int(nthreads))


# ------------------------------------------------------------------ the real reference (oracle/_ref)

def ref_precond_diagonal(S, kind):
    d = np.zeros(S.n)
    rc = ref().amie_ref_precond_diagonal(*S.head(), int(kind), _vp(d))
    assert rc == 0
    return d


def ref_cg(S, x0=None, precond=0, eps=1e-10, maxit=-1, nssor=32, rowstart=0, colstart=0, nthreads=1, b=None, diag=None):
    R = ref()
    if precond == 4:
        diag = np.ascontiguousarray(diag, np.float64)       # stays alive until the call below returns
        R.amie_ref_set_user_diagonal(_vp(diag), u64(diag.size))
    b = S.b if b is None else np.ascontiguousarray(b, np.float64)
    x = np.zeros(S.n)
    nit = u64()
    wall = f64()
    log = ctypes.create_string_buffer(8192)
    x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
    ret = R.amie_ref_cg(*S.head(), _vp(b), _vp(x0), u64(0 if x0 is None else x0.size), int(precond), f64(eps),
                        int(maxit), u64(nssor), u64(rowstart), u64(colstart), int(nthreads), _vp(x),
                        ctypes.byref(nit), ctypes.byref(wall), log, u64(8192))
    return ret, x, nit.value, wall.value, log.value.decode(errors="replace")


def ref_cg_synth(preset, n, eps=1e-10, maxit=-1, nssor=32, nthreads=1, spmv_reps=0, seed=1, want_x=True, x0=None):
    """The reference's ConjugateGradient::solve on the synthetic system `preset`-n, generated straight into the
    reference's own storage (one copy of the matrix in host memory: benchmark sizes).  Returns
    (ret, x or None, nit, solve seconds, seconds per assign(y, A*b) or None, dict(stride, nb, nnzb))."""
    R, L = ref(), synth_lib()
    h = L.amie_b200_synth_create(preset.encode(), int(n), int(seed))
    if not h:
        raise ValueError(f"unknown synthetic preset {preset!r}")
    h = ctypes.c_void_p(h)
    st, nb, tot = ctypes.c_int(), u64(), u64()
    L.amie_b200_synth_sizes(h, ctypes.byref(st), ctypes.byref(nb), None)
    rs = np.zeros(nb.value, np.uint32)
    L.amie_b200_synth_count(h, 0, nb.value, _vp(rs), ctypes.byref(tot))
    FILL = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)

    def fill(user, ci, arr, forces):
        return L.amie_b200_synth_fill(h, 0, nb.value, ci, arr, forces)
    cb = FILL(fill)
    x = np.zeros(nb.value * st.value) if want_x else None
    nit, wall, spmv_s = u64(), f64(), f64(-1.0)
    log = ctypes.create_string_buffer(8192)
    x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
    R.amie_ref_cg_fill_x0.argtypes = [ctypes.c_int, u64, ctypes.c_void_p, u64, FILL, ctypes.c_void_p, ctypes.c_void_p, u64,
                                      f64, ctypes.c_int, u64,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_char_p, u64]
    ret = R.amie_ref_cg_fill_x0(st.value, nb.value, _vp(rs), tot.value, cb, None, _vp(x0), 0 if x0 is None else x0.size,
                                eps, int(maxit), int(nssor), int(nthreads),
                                int(spmv_reps), _vp(x), ctypes.cast(ctypes.byref(nit), ctypes.c_void_p),
                                ctypes.cast(ctypes.byref(wall), ctypes.c_void_p), ctypes.cast(ctypes.byref(spmv_s), ctypes.c_void_p),
                                log, 8192)
    L.amie_b200_synth_destroy(h)
    if ret < 0:
        raise RuntimeError("amie_ref_cg_fill: the generator failed")
    return ret, x, nit.value, wall.value, (spmv_s.value if spmv_reps > 0 else None), dict(stride=st.value, nb=nb.value, nnzb=tot.value)


def ref_bicgstab(S, x0=None, precond=0, eps=1e-10, maxit=-1, nthreads=1, b=None, diag=None):
    R = ref()
    if precond == 4:
        diag = np.ascontiguousarray(diag, np.float64)
        R.amie_ref_set_user_diagonal(_vp(diag), u64(diag.size))
    b = S.b if b is None else np.ascontiguousarray(b, np.float64)
    x = np.zeros(S.n)
    nit = u64()
    wall = f64()
    log = ctypes.create_string_buffer(8192)
    x0 = None if x0 is None else np.ascontiguousarray(x0, np.float64)
    ret = R.amie_ref_bicgstab(*S.head(), _vp(b), _vp(x0), u64(0 if x0 is None else x0.size), int(precond),
                              f64(eps), int(maxit), int(nthreads), _vp(x), ctypes.byref(nit), ctypes.byref(wall),
                              log, u64(8192))
    return ret, x, nit.value, wall.value, log.value.decode(errors="replace")


def ref_spmv(S, v, b=None, mode=0, rowstart=0, colstart=0, nthreads=1, reps=1):
    R = ref()
    y = np.zeros(S.n)
    wall = f64()
    v = np.ascontiguousarray(v, np.float64)
    b = None if b is None else np.ascontiguousarray(b, np.float64)
    rc = R.amie_ref_spmv(*S.head(), _vp(v), _vp(b), int(mode), u64(rowstart), u64(colstart), int(nthreads),
                         int(reps), _vp(y), ctypes.byref(wall))
    assert rc == 0
    return y, wall.value


def ref_inverse_diagonal(S):
    d = np.zeros(S.n)
    ref().amie_ref_inverse_diagonal(*S.head(), _vp(d))
    return d


def ref_max_threads():
    return ref().amie_ref_max_threads()


# ------------------------------------------------------------------ value assembly + Dirichlet elimination

class Elements:
    """Element connectivity + elementary matrices in the layout of include/amie_b200.h:
    ids u32[n_elem, npe] (0xFFFFFFFF = unused slot); ke f64[n_elem, npe, npe, s*s] with block (j,k) column-major."""

    def __init__(self, stride, ids, ke, scales=None):
        self.stride = int(stride)
        self.ids = np.ascontiguousarray(ids, np.uint32)
        self.n_elem, self.npe = self.ids.shape
        self.ke = np.ascontiguousarray(ke, np.float64).reshape(self.n_elem, self.npe, self.npe, self.stride * self.stride)
        self.scales = np.ones(self.n_elem) if scales is None else np.ascontiguousarray(scales, np.float64)

    def pattern(self, nb):
        """row_size / column_index of the union of node pairs (plus every diagonal), as Assembly::make_final
        builds it from a std::set of pairs (solvers/assembly.cpp:455-521)."""
        ids = self.ids.astype(np.int64)
        rows, cols = [np.arange(nb)], [np.arange(nb)]
        for j in range(self.npe):
            for k in range(self.npe):
                ok = (ids[:, j] != 0xFFFFFFFF) & (ids[:, k] != 0xFFFFFFFF)
                rows.append(ids[ok, j])
                cols.append(ids[ok, k])
        key = np.unique(np.concatenate(rows) * nb + np.concatenate(cols))
        r, c = key // nb, key % nb
        return np.bincount(r, minlength=nb).astype(np.uint32), c.astype(np.uint32)


def oracle_assemble(stride, nb, row_size, column_index, el):
    cl = stride + stride % 2
    array = np.zeros(column_index.size * stride * cl)
    rc = oracle().amie_oracle_assemble(int(stride), u64(nb), _vp(row_size), _vp(column_index), u64(column_index.size),
                                       u64(el.n_elem), int(el.npe), _vp(el.ids), _vp(el.ke), _vp(el.scales), _vp(array))
    assert rc == 0, "element couples nodes outside the sparsity pattern"
    return array


def _bc_args(S_or_tuple, array, forces, natural, add, fix_ids, fix_values, force_ids, force_values):
    stride, nb, row_size, column_index = S_or_tuple
    fix_ids = np.ascontiguousarray(fix_ids, np.uint32)
    fix_values = np.ascontiguousarray(fix_values, np.float64)
    force_ids = np.ascontiguousarray(force_ids if force_ids is not None else [], np.uint32)
    force_values = np.ascontiguousarray(force_values if force_values is not None else [], np.float64)
    keep = (fix_ids, fix_values, force_ids, force_values)
    return keep, (int(stride), u64(nb), _vp(row_size), _vp(column_index), u64(column_index.size), _vp(array), _vp(forces),
                  _vp(natural), _vp(add), u64(fix_ids.size), _vp(fix_ids), _vp(fix_values), u64(force_ids.size),
                  _vp(force_ids), _vp(force_values))


def oracle_set_bcs(stride, nb, row_size, column_index, array, forces, fix_ids, fix_values, force_ids=None,
                   force_values=None, natural=None, add_to_forces=None):
    """In place on copies; returns (array, forces, natural, add_to_forces)."""
    array, forces = array.copy(), forces.copy()
    natural = None if natural is None else natural.copy()
    add = None if add_to_forces is None else add_to_forces.copy()
    keep, args = _bc_args((stride, nb, row_size, column_index), array, forces, natural, add, fix_ids, fix_values,
                          force_ids, force_values)
    oracle().amie_oracle_set_boundary_conditions(*args)
    return array, forces, natural, add


def ref_set_bcs(stride, nb, row_size, column_index, array, forces, fix_ids, fix_values, force_ids=None,
                force_values=None, natural=None, add_to_forces=None):
    array, forces = array.copy(), forces.copy()
    natural = None if natural is None else natural.copy()
    add = None if add_to_forces is None else add_to_forces.copy()
    keep, args = _bc_args((stride, nb, row_size, column_index), array, forces, natural, add, fix_ids, fix_values,
                          force_ids, force_values)
    rc = ref().amie_ref_set_boundary_conditions(*args)
    assert rc == 0
    return array, forces, natural, add


# ------------------------------------------------------------------ field recovery after the solve

def oracle_element_fields(dim, ids, dshape, jinv, u, tensors=None, imposed_strain=None, imposed_stress=None,
                          tensor_of_elem=None):
    """(total strain, mechanical strain, real stress), each [n_elem, nc]: ElementState::getField restated
    (oracle/amie_oracle_fields.c)."""
    ids = np.ascontiguousarray(ids, np.uint32)
    ne, npe = ids.shape
    nc = 3 if dim == 2 else 6
    c = lambda a, t=np.float64: None if a is None else np.ascontiguousarray(a, t)
    dshape, jinv, u, tensors = c(dshape), c(jinv), c(u), c(tensors)
    imposed_strain, imposed_stress, tensor_of_elem = c(imposed_strain), c(imposed_stress), c(tensor_of_elem, np.uint32)
    tot, mech, sig = np.zeros((ne, nc)), np.zeros((ne, nc)), np.zeros((ne, nc))
    rc = oracle().amie_oracle_element_fields(int(dim), u64(ne), int(npe), _vp(ids), _vp(dshape), _vp(jinv), _vp(tensors),
                                             _vp(imposed_strain), _vp(imposed_stress), _vp(tensor_of_elem), _vp(u),
                                             u64(u.size), _vp(tot), _vp(mech), _vp(sig))
    assert rc == 0
    return tot, mech, sig


def oracle_principal(dim, values, double_offdiag):
    """toPrincipal per element: values [n_elem, nc] -> [n_elem, dim] (oracle/amie_oracle_fields.c)."""
    v = np.ascontiguousarray(values, np.float64)
    out = np.zeros((v.shape[0], int(dim)))
    rc = oracle().amie_oracle_principal(int(dim), u64(v.shape[0]), _vp(v), int(bool(double_offdiag)), _vp(out))
    assert rc == 0
    return out


# ------------------------------------------------------------------ Assembly::extrapolate

def oracle_extrapolate(prev, back, factor=1.0):
    """(x0, back after the NaN scrub): Assembly::extrapolate for a two-vector history (oracle/amie_oracle.c)."""
    prev = np.ascontiguousarray(prev, np.float64)
    back = np.array(back, np.float64)
    out = np.zeros_like(back)
    oracle().amie_oracle_extrapolate(_vp(prev), _vp(back), u64(back.size), f64(factor), _vp(out))
    return out, back


def ref_extrapolate(prev, back, disp, factor=1.0, nhist=2):
    """The real Assembly::extrapolate on a hand-filled history: (returned vector, newest history vector afterwards,
    history size afterwards)."""
    R = ref()
    R.amie_ref_extrapolate.restype = u64
    prev = np.ascontiguousarray(prev, np.float64)
    back = np.ascontiguousarray(back, np.float64)
    disp = np.ascontiguousarray(disp, np.float64)
    out = np.zeros(max(back.size, disp.size, 1))
    back_out = back.copy()
    hs = u64()
    n = R.amie_ref_extrapolate(_vp(prev), _vp(back), u64(back.size), int(nhist), _vp(disp), u64(disp.size), f64(factor),
                               _vp(out), _vp(back_out), ctypes.byref(hs))
    return out[:n].copy(), back_out, hs.value
