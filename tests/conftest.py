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
