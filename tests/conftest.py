import os
import sys

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
