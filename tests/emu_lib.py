"""ctypes bindings of tests/emu/libamie_emu.so: the product's dependency-light kernel SOURCES compiled for the host
(tests/emu/cuda_emu.h).  Test infrastructure only -- it lets the CPU suite check the kernels' index logic and
arithmetic order against the oracle on a box without a GPU; it is not a path of the product."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "xfem-amie_b200", "csrc")
u64 = ctypes.c_uint64
f64 = ctypes.c_double
_lib = None


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def emu():
    global _lib
    if _lib is None:
        so = os.path.join(EMU_DIR, "libamie_emu.so")
        deps = [os.path.join(EMU_DIR, f) for f in ("emu_driver.cpp", "cuda_emu.h")] + \
               [os.path.join(CSRC, f) for f in ("device_utils.cuh", "kernels_setup.cuh", "kernels_assemble.cuh",
                                                 "kernels_fields.cuh", "kernels_history.cuh")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
            subprocess.check_call(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O1", "-std=c++20",
                                   "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas",
                                   "-I" + CSRC, os.path.join(EMU_DIR, "emu_driver.cpp"), "-o", so])
        _lib = ctypes.CDLL(so)
    return _lib


def compact(array_padded, stride):
    """reference layout (block k at k*s*cl, element (r,c) at c*cl+r) -> compact column-major s x s blocks, by K-Repack"""
    s = int(stride)
    cl = s + s % 2
    arr = np.ascontiguousarray(array_padded, np.float64)
    nblk = arr.size // (s * cl)
    out = np.zeros(nblk * s * s)
    assert emu().emu_repack(s, _vp(arr), _vp(out), u64(nblk)) == 0
    return out


def compact_scatter(array_padded, stride, block_to):
    """K-Repack through a block map: block k of the host array lands on stored block block_to[k]."""
    s = int(stride)
    cl = s + s % 2
    arr = np.ascontiguousarray(array_padded, np.float64)
    bt = np.ascontiguousarray(block_to, np.uint32)
    out = np.zeros(bt.size * s * s)
    assert emu().emu_repack_scatter(s, _vp(arr), _vp(bt), u64(bt.size), _vp(out)) == 0
    return out


def padded(vals_compact, stride):
    s = int(stride)
    cl = s + s % 2
    v = np.asarray(vals_compact).reshape(-1, s, s)
    out = np.zeros((v.shape[0], s, cl))
    out[:, :, :s] = v
    return out.reshape(-1)


def precond_diagonal(kind, stride, row_size, column_index, vals_compact):
    rs = np.ascontiguousarray(row_size, np.uint32)
    ci = np.ascontiguousarray(column_index, np.uint32)
    d = np.zeros(rs.size * int(stride))
    rc = emu().emu_precond_diagonal(int(kind), int(stride), u64(rs.size), _vp(rs), _vp(ci), _vp(vals_compact), _vp(d))
    assert rc == 0, rc
    return d


def assemble(stride, row_size, column_index, ids, ke, scales, vals=None, mark=None):
    """Full assembly (vals None) or the incremental re-accumulation of the blocks touched by elements
    [mark[0], mark[0]+mark[1]) on top of `vals`.  Returns (rc, vals_compact)."""
    rs = np.ascontiguousarray(row_size, np.uint32)
    ci = np.ascontiguousarray(column_index, np.uint32)
    ids = np.ascontiguousarray(ids, np.uint32)
    ke = np.ascontiguousarray(ke, np.float64)
    scales = np.ascontiguousarray(scales, np.float64)
    s = int(stride)
    out = np.zeros(ci.size * s * s) if vals is None else np.array(vals, np.float64)
    first, count = (0, 0) if mark is None else mark
    rc = emu().emu_assemble(s, u64(rs.size), _vp(rs), _vp(ci), u64(ci.size), u64(ids.shape[0]), int(ids.shape[1]), _vp(ids),
                            _vp(ke), _vp(scales), 1 if mark is None else 0, u64(first), u64(count), _vp(out))
    return rc, out


def split_rows(row_size, column_index, parts):
    """The numbering dist.cu gives each part of a row-partitioned matrix: rows [r0, r1) of the global structure, the
    column indices renumbered IN PLACE (owned: c - r0, foreign: nb + position in the sorted halo list), so the blocks of a
    row stay in global ascending order.  Returns [(r0, r1, k0, k1, row_size_local, col_local, halo)] for bounds `parts`."""
    rs = np.ascontiguousarray(row_size, np.uint32)
    ci = np.ascontiguousarray(column_index, np.uint32)
    rp = np.concatenate([[0], np.cumsum(rs, dtype=np.int64)])
    out = []
    for r0, r1 in zip(parts[:-1], parts[1:]):
        k0, k1 = int(rp[r0]), int(rp[r1])
        c = ci[k0:k1].astype(np.int64)
        own = (c >= r0) & (c < r1)
        halo = np.unique(c[~own]).astype(np.uint32)
        loc = np.where(own, c - r0, (r1 - r0) + np.searchsorted(halo, c)).astype(np.uint32)
        out.append((int(r0), int(r1), k0, k1, rs[r0:r1].copy(), loc, halo))
    return out


def assemble_part(stride, part, nb_global, ids, ke, scales=None, vals=None, mark=None):
    r0, r1, k0, k1, rs, loc, halo = part
    ids = np.ascontiguousarray(ids, np.uint32)
    ke = np.ascontiguousarray(ke, np.float64)
    scales = None if scales is None else np.ascontiguousarray(scales, np.float64)
    s = int(stride)
    out = np.zeros(loc.size * s * s) if vals is None else np.array(vals, np.float64)
    first, count = (0, 0) if mark is None else mark
    rc = emu().emu_assemble_part(s, u64(rs.size), _vp(rs), _vp(loc), u64(loc.size), u64(r0), _vp(halo), u64(halo.size), u64(nb_global),
                                 u64(ids.shape[0]), int(ids.shape[1]), _vp(ids), _vp(ke), _vp(scales),
                                 1 if mark is None else 0, u64(first), u64(count), _vp(out))
    return rc, out


def dirichlet_part(stride, part, nb_global, vals_compact, forces, fix_ids, fix_values, force_ids=None, force_values=None,
                   natural=None, add_to_forces=None):
    r0, r1, k0, k1, rs, loc, halo = part
    vals, forces = np.array(vals_compact, np.float64), np.array(forces, np.float64)
    nat = None if natural is None else np.array(natural, np.float64)
    add = None if add_to_forces is None else np.ascontiguousarray(add_to_forces, np.float64)
    fi, fv = np.ascontiguousarray(fix_ids, np.uint32), np.ascontiguousarray(fix_values, np.float64)
    gi = np.ascontiguousarray([] if force_ids is None else force_ids, np.uint32)
    gv = np.ascontiguousarray([] if force_values is None else force_values, np.float64)
    dirty = np.zeros(max(1, loc.size), np.uint8)
    rc = emu().emu_dirichlet_part(int(stride), u64(rs.size), _vp(rs), _vp(loc), u64(loc.size), u64(r0), _vp(halo), u64(halo.size),
                                  u64(nb_global), _vp(vals), _vp(forces), _vp(nat), _vp(add),
                                  u64(fi.size), _vp(fi), _vp(fv), u64(gi.size), _vp(gi), _vp(gv), _vp(dirty))
    assert rc == 0, rc
    return vals, forces, nat, dirty[:loc.size]


def dirichlet(stride, row_size, column_index, vals_compact, forces, fix_ids, fix_values, force_ids=None, force_values=None,
              natural=None, add_to_forces=None):
    rs = np.ascontiguousarray(row_size, np.uint32)
    ci = np.ascontiguousarray(column_index, np.uint32)
    vals, forces = np.array(vals_compact, np.float64), np.array(forces, np.float64)
    nat = None if natural is None else np.array(natural, np.float64)
    add = None if add_to_forces is None else np.ascontiguousarray(add_to_forces, np.float64)
    fi, fv = np.ascontiguousarray(fix_ids, np.uint32), np.ascontiguousarray(fix_values, np.float64)
    gi = np.ascontiguousarray([] if force_ids is None else force_ids, np.uint32)
    gv = np.ascontiguousarray([] if force_values is None else force_values, np.float64)
    dirty = np.zeros(max(1, ci.size), np.uint8)
    rc = emu().emu_dirichlet(int(stride), u64(rs.size), _vp(rs), _vp(ci), u64(ci.size), _vp(vals), _vp(forces), _vp(nat), _vp(add),
                             u64(fi.size), _vp(fi), _vp(fv), u64(gi.size), _vp(gi), _vp(gv), _vp(dirty))
    assert rc == 0, rc
    return vals, forces, nat, dirty[:ci.size]


def element_fields(dim, ids, dshape, jinv, u, tensors, imposed_strain, imposed_stress, tensor_of_elem, variant=0):
    ids = np.ascontiguousarray(ids, np.uint32)
    ne, npe = ids.shape
    nc = 3 if dim == 2 else 6
    c = lambda a, t=np.float64: None if a is None else np.ascontiguousarray(a, t)
    tensors = c(tensors)
    nt = tensors.shape[0]
    # the C-ABI turns absent imposed terms into zeros (fields.cu)
    es = np.zeros((nt, nc)) if imposed_strain is None else c(imposed_strain)
    ss = np.zeros((nt, nc)) if imposed_stress is None else c(imposed_stress)
    u = c(u)
    out = [np.zeros((ne, nc)) for _ in range(3)]
    rc = emu().emu_element_fields(int(dim), u64(ne), int(npe), _vp(ids), _vp(c(dshape)), _vp(c(jinv)), _vp(tensors), _vp(es), _vp(ss),
                                  _vp(c(tensor_of_elem, np.uint32)), _vp(u), u64(u.size), int(variant), *[_vp(o) for o in out])
    assert rc == 0, rc
    return tuple(out)


def element_principal(dim, values, double_offdiag):
    v = np.ascontiguousarray(values, np.float64)
    out = np.zeros((v.shape[0], int(dim)))
    rc = emu().emu_element_principal(int(dim), u64(v.shape[0]), _vp(v), int(bool(double_offdiag)), _vp(out))
    assert rc == 0, rc
    return out


def extrapolate(prev, back, factor):
    prev = np.ascontiguousarray(prev, np.float64)
    back = np.array(back, np.float64)
    x = np.zeros_like(back)
    emu().emu_extrapolate(_vp(prev), _vp(back), _vp(x), u64(back.size), f64(factor))
    return x, back


def times_zero(x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.ones_like(x)
    emu().emu_times_zero(_vp(x), _vp(out), u64(x.size))
    return out
