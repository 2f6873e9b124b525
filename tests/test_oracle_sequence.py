"""Pins the oracle's Assembly::extrapolate bit for bit against the real method on a
hand-filled displacementHistory (oracle/ref_harness.cpp: amie_ref_extrapolate), including its three cases."""
import numpy as np
import pytest


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def test_extrapolate_matches_live_reference(ol):
    if ol.ref() is None:
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(0)
    n = 777
    prev, back, disp = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    back[5] = np.nan
    back[7] = prev[7] = 0.0
    back[9] = prev[9] = -0.0
    prev[11] = np.inf
    for factor in (1.0, 0.5, -2.0, 1e300, 0.0):
        x0, scrubbed = ol.oracle_extrapolate(prev, back, factor)
        r, rback, hs = ol.ref_extrapolate(prev, back, disp, factor)
        assert hs == 2 and same_bits(x0, r) and same_bits(scrubbed, rback), factor
    # fewer than two vectors: the current displacements come back (solvers/assembly.cpp:1774-1779)
    r, _, hs = ol.ref_extrapolate(prev, back, disp, 1.0, nhist=1)
    assert same_bits(r, disp)
    r, _, hs = ol.ref_extrapolate(prev, back, disp[:0], 1.0, nhist=0)
    assert r.size == 0
    # size mismatch: history cleared, empty vector (:1781-1785)
    r, _, hs = ol.ref_extrapolate(prev, back, disp[:10], 1.0)
    assert r.size == 0 and hs == 0
