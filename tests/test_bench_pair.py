"""bench.py's search for the same-size sample both sides solve (cpu_baseline.same_size_pair): the tolerance that makes the
reference's solve of the workload-size system a bounded, non-trivial sample -- or the finding that there is none (nit(eps)
jumps over the window on the benchmark system).  Host logic only; the solver behind it here is the oracle."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("amie_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def step_function(table):
    """nit(eps) as a non-increasing step function given by [(eps_threshold, nit)]: the first threshold eps reaches."""
    def solve(eps):
        for thr, nit in table:
            if eps >= thr:
                return nit, float(nit)
        return table[-1][1], float(table[-1][1])
    return solve


def test_window_found(bench):
    cold = step_function([(5e-2, 0), (2e-3, 25), (4e-4, 90), (1e-5, 400), (0, 2500)])
    calls = []
    eps, nit, tried = bench.search_same_size_sample(lambda e: (calls.append(e), cold(e))[1])
    assert 60 <= nit <= 160 and cold(eps)[0] == nit
    assert len(calls) <= 8 + 6 and len(tried) == len(calls)


def test_plateau_gives_no_sample_unless_the_long_one_is_accepted(bench):
    """The benchmark system: a handful of iterations, then ~600 (profiles/r02_notes.md sections 5 and 10).  A degenerate
    sample (0 or 3 iterations time nothing) is never returned."""
    cold = step_function([(5e-2, 0), (4e-3, 3), (1e-6, 600), (0, 2500)])
    assert bench.search_same_size_sample(cold) is None
    eps, nit, tried = bench.search_same_size_sample(cold, long_ok=True)
    assert nit == 600 and cold(eps)[0] == 600
    assert bench.search_same_size_sample(step_function([(0, 2)]), long_ok=True) is None


def test_choose_sample_never_returns_a_trivial_solve(bench):
    assert bench.choose_sample([(1e-1, 0, 0.), (1e-2, 4, 1.)]) is None
    assert bench.choose_sample([(1e-1, 0, 0.), (1e-2, 12, 1.), (1e-3, 600, 9.)])[1] == 12
    assert bench.choose_sample([(1e-2, 12, 1.), (1e-3, 70, 2.), (1e-4, 150, 3.)])[1] == 70
    assert bench.choose_sample([(1e-3, 600, 9.)]) is None
    assert bench.choose_sample([(1e-3, 300, 9.)], max_above=400)[1] == 300


def test_search_with_the_oracle_as_the_solver(bench, pkg, ol):
    """The real thing at a small size: solves of S3-hex-14 through the oracle; what the search returns is reproduced by a
    direct solve with the tolerance it names."""
    from conftest import make_sys
    S = make_sys(pkg, ol, "S3-hex", 14, 1)

    def cold(eps):
        ret, x, info = ol.oracle_cg(S, nssor=32, eps=eps)
        return int(info.nit), 0.
    found = bench.search_same_size_sample(cold)
    assert found is not None
    eps, nit, tried = found
    assert nit >= bench.PAIR_MIN_NIT and cold(eps)[0] == nit


def test_reference_restart_entry_matches_the_oracle(pkg, ol):
    """oracle/ref_harness.cpp amie_ref_cg_fill / amie_ref_cg_fill_x0 (the reference's own ConjugateGradient::solve on the
    system generated into its storage -- what bench.py's CPU leg runs -- from zero and from a starting vector) against the
    oracle restatement: same count, same bits."""
    if ol.ref() is None:
        pytest.skip("oracle/_ref not built")
    from conftest import make_sys
    S = make_sys(pkg, ol, "S3-hex", 10, 1)
    ret_a, xa, nit_a, _, _, dims = ol.ref_cg_synth("S3-hex", 10, eps=1e-3, nssor=32)
    assert dims["nb"] == S.nb
    ret_o, xo, info = ol.oracle_cg(S, nssor=32, eps=1e-3)
    assert int(nit_a) == int(info.nit) and np.array_equal(xa, xo)
    ret_b, xb, nit_b, _, _, _ = ol.ref_cg_synth("S3-hex", 10, eps=1e-9, nssor=32, x0=xa)
    ret_p, xp, info_p = ol.oracle_cg(S, x0=xa, nssor=32, eps=1e-9)
    assert int(nit_b) == int(info_p.nit) > 0 and np.array_equal(xb, xp)
