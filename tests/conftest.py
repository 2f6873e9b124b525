import os
import sys

# multi-device contexts (tests/test_gpu_group.py) need eager CUDA module loading, decided when CUDA initialises:
# before torch or the library touch the driver (xfem-amie_b200/csrc/group.cu)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    so = os.path.join(g.PKG_DIR, "libamie_b200.so")
    if not os.path.exists(so):
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def ol():
    import oracle_lib
    oracle_lib.oracle()
    return oracle_lib


def make_sys(pkg, ol, preset, n, seed=1):
    syn = pkg.Synth(preset, n, seed)
    rs, ci, arr, b = syn.rows()
    return ol.Sys(syn.stride, syn.nb, rs, ci, arr, b)


@pytest.fixture(scope="session")
def systems(pkg, ol):
    cache = {}

    def get(preset, n, seed=1):
        key = (preset, n, seed)
        if key not in cache:
            cache[key] = make_sys(pkg, ol, preset, n, seed)
        return cache[key]
    return get


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb else 1.0)


def random_spd_blocks(stride, nb, seed):
    """Symmetric, block-diagonally-dominant random block matrix in the reference layout."""
    rng = np.random.default_rng(seed)
    nbrs = [set([i]) for i in range(nb)]
    for i in range(nb):
        for j in rng.integers(0, nb, 4):
            nbrs[i].add(int(j))
            nbrs[int(j)].add(i)
    cl = stride + stride % 2
    blocks = {}
    for i in range(nb):
        for j in nbrs[i]:
            if j > i:
                B = 0.1 * rng.standard_normal((stride, stride))
                blocks[(i, j)] = B
                blocks[(j, i)] = B.T
    for i in range(nb):
        D = 0.1 * rng.standard_normal((stride, stride))
        blocks[(i, i)] = 0.5 * (D + D.T) + np.eye(stride) * (2.0 + 0.1 * stride * len(nbrs[i]))
    rs = np.array([len(nbrs[i]) for i in range(nb)], np.uint32)
    ci, arr = [], []
    for i in range(nb):
        for j in sorted(nbrs[i]):
            ci.append(j)
            blk = np.zeros((stride, cl))              # [c][r] with pad
            blk[:, :stride] = blocks[(i, j)].T        # element (r,c) at c*cl + r
            arr.append(blk.ravel())
    return rs, np.array(ci, np.uint32), np.concatenate(arr), rng.standard_normal(nb * stride)


def with_enrichment_like_rows(S, n_extra, seed, ol):
    """What XFEM enrichment does to the assembled system (elements/elements.cpp:3711-3717): extra block rows with
    node-like ids beyond the mesh nodes, coupled to the nodes of the tetrahedra an inclusion surface cuts -- row lengths
    anywhere between a handful and several dozen blocks -- and to one another; the rows of the nodes they touch grow by
    as much.  Built on top of a system S (stride 3): symmetric, diagonally dominant additions, so still SPD."""
    import scipy.sparse as sp
    s, nb = S.stride, S.nb
    rng = np.random.default_rng(seed)
    A = S.to_scipy().tolil()
    n2 = (nb + n_extra) * s
    B = sp.lil_matrix((n2, n2))
    B[:nb * s, :nb * s] = A
    for e in range(n_extra):
        row = nb + e
        centre = int(rng.integers(0, nb))
        k = int(rng.integers(3, 41))
        nbrs = set(int(v) for v in np.clip(centre + rng.integers(-40, 41, k), 0, nb - 1))
        if e and rng.random() < 0.6:
            nbrs.add(nb + int(rng.integers(0, e)))          # enrichment dofs of neighbouring cut elements couple too
        tot = 0.0
        for c in nbrs:
            blk = 0.05 * rng.standard_normal((s, s))
            B[row * s:(row + 1) * s, c * s:(c + 1) * s] = blk
            B[c * s:(c + 1) * s, row * s:(row + 1) * s] = blk.T
            tot += np.abs(blk).sum()
            for m in range(s):
                B[c * s + m, c * s + m] += np.abs(blk).sum()
        D = 0.05 * rng.standard_normal((s, s))
        B[row * s:(row + 1) * s, row * s:(row + 1) * s] = 0.5 * (D + D.T) + np.eye(s) * (1.0 + tot)
    Bb = sp.bsr_matrix(B.tocsr(), blocksize=(s, s))
    Bb.sort_indices()
    cl = s + s % 2
    rs = np.diff(Bb.indptr).astype(np.uint32)
    arr = np.zeros((Bb.indices.size, s, cl))
    arr[:, :, :s] = Bb.data.transpose(0, 2, 1)
    b = np.concatenate([S.b, 0.01 * rng.standard_normal(n_extra * s)])
    return ol.Sys(s, nb + n_extra, rs, Bb.indices.astype(np.uint32), arr.reshape(-1), b)
