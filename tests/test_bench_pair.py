"""bench.py's search for the same-size sample both sides solve (cpu_baseline.same_size_pair): the tolerance -- and, when
nit(eps) jumps over the window as on the benchmark system, the starting vector -- that makes the reference's solve of the
workload-size system a bounded, non-trivial sample.  Host logic only; the solver behind it here is the oracle."""
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


def test_window_found_from_a_cold_start(bench):
    cold = step_function([(5e-2, 0), (2e-3, 25), (4e-4, 90), (1e-5, 400), (0, 2500)])
    calls = []
    found = bench.search_same_size_sample(lambda e: (calls.append(e), cold(e))[1], lambda eps_a: pytest.fail("no warm start needed"))
    eps, nit, eps_a, tried_cold, tried_warm = found
    assert eps_a is None and 60 <= nit <= 160 and cold(eps)[0] == nit and not tried_warm
    assert len(calls) <= 8 + 6


def test_plateau_takes_the_warm_start(bench):
    """The benchmark system: a handful of iterations, then ~600 (profiles/r02_notes.md section 5).  No cold sample in
    [10, 160] -> restart from the shortest long-enough solve and tighten."""
    cold = step_function([(5e-2, 0), (4e-3, 3), (1e-6, 600), (0, 2500)])
    started = {}

    def warm_from(eps_a):
        started["eps_a"] = eps_a
        return lambda eps: ((0, 0.) if eps >= eps_a * 0.45 else (int(40 * np.log10(eps_a / eps)), 1.))
    eps, nit, eps_a, tried_cold, tried_warm = bench.search_same_size_sample(cold, warm_from)
    assert eps_a == started["eps_a"] and cold(eps_a)[0] == 600
    assert 60 <= nit <= 160 and eps < eps_a and tried_warm
    # nothing usable anywhere: no sample rather than a degenerate one (0 iterations time nothing)
    assert bench.search_same_size_sample(step_function([(0, 2)]), warm_from) is None
    never = bench.search_same_size_sample(cold, lambda eps_a: (lambda eps: (0, 0.)))
    assert never is None


def test_choose_sample_never_returns_a_trivial_solve(bench):
    assert bench.choose_sample([(1e-1, 0, 0.), (1e-2, 4, 1.)]) is None
    assert bench.choose_sample([(1e-1, 0, 0.), (1e-2, 12, 1.), (1e-3, 600, 9.)])[1] == 12
    assert bench.choose_sample([(1e-2, 12, 1.), (1e-3, 70, 2.), (1e-4, 150, 3.)])[1] == 70
    assert bench.choose_sample([(1e-3, 600, 9.)]) is None
    assert bench.choose_sample([(1e-3, 300, 9.)], max_above=400)[1] == 300


def test_search_with_the_oracle_as_the_solver(bench, pkg, ol):
    """The real thing at a small size: cold and restarted solves of S3-hex-14 through the oracle; what the search returns
    is reproduced by a direct solve with the tolerance and starting vector it names."""
    from conftest import make_sys
    S = make_sys(pkg, ol, "S3-hex", 14, 1)

    def cold(eps):
        ret, x, info = ol.oracle_cg(S, nssor=32, eps=eps)
        return int(info.nit), 0.

    def warm_from(eps_a):
        ret, xa, info = ol.oracle_cg(S, nssor=32, eps=eps_a)

        def warm(eps):
            ret, x, info = ol.oracle_cg(S, x0=xa, nssor=32, eps=eps)
            return int(info.nit), 0.
        return warm
    found = bench.search_same_size_sample(cold, warm_from)
    assert found is not None
    eps, nit, eps_a, tried_cold, tried_warm = found
    assert nit >= bench.PAIR_MIN_NIT
    again = cold(eps)[0] if eps_a is None else warm_from(eps_a)(eps)[0]
    assert again == nit


def test_reference_restart_entry_matches_the_oracle(pkg, ol):
    """oracle/ref_harness.cpp amie_ref_cg_fill_x0 (the reference's own ConjugateGradient::solve(x0, ...) on the generated
    system, what bench.py's CPU leg runs for a restarted sample) against the oracle restatement: same count, same bits."""
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
