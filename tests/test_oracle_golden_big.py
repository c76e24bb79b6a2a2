"""The C oracle against the reference's answers at the operating points (n = 120 / 500 / 2 000 /
50 000): tests/golden/kernels_big.npz holds >= 1 000 distinct (state, node) pairs per likelihood,
computed by the reference's Cython kernels (oracle/make_golden_big.py).  CPU only."""
import numpy as np
import pytest

import pyoracle as O
import workloads as W
from conftest import load_golden
from make_golden_big import state_of, PAIRS


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.fixture(scope="module")
def big():
    return load_golden("kernels_big.npz")


@pytest.mark.parametrize("name", ["cfg2", "cfg4"])
def test_k1_k5_undirected(big, name):
    w = W.make_workload(name)
    X, t, j = state_of(w, int(big[name + "_seed"]))
    assert np.array_equal(t, big[name + "_t"]) and np.array_equal(j, big[name + "_j"])
    assert len(set(zip(t.tolist(), j.tolist()))) == PAIRS >= 1000
    b = float(w["intercept"][0])
    got = [O.partial_loglikelihood(w["Y"][a], X[a], b, int(c)) for a, c in zip(t, j)]
    assert rel(got, big[name + "_k1"]) < 1e-12
    k5 = O.undirected_network_loglikelihood(w["Y"], O.calculate_distances(X), b)
    assert rel(k5, big[name + "_k5"]) < 1e-12


def test_k2_k4_directed_n2000(big):
    w = W.make_workload("cfg3")
    X, t, j = state_of(w, int(big["cfg3_seed"]))
    bi, bo = w["intercept"]
    got = [O.directed_partial_loglikelihood(w["Y"][a], X[a], w["radii"], bi, bo, int(c)) for a, c in zip(t, j)]
    assert rel(got, big["cfg3_k2"]) < 1e-12
    k4 = O.directed_network_loglikelihood(w["Y"], O.calculate_distances(X), w["radii"], bi, bo)
    assert rel(k4, big["cfg3_k4"]) < 1e-11


def test_k3_k6_case_control_n50000(big):
    w = W.make_workload("cfg5")
    X, t, j = state_of(w, int(big["cfg5_seed"]))
    got = [O.approx_directed_partial_loglikelihood(X[a], w["radii"], w["in_edges"][a], w["out_edges"][a],
                                                   w["degrees"][a], w["ctrl_in"][a], w["ctrl_out"][a], 0.3, 0.7, int(c))
           for a, c in zip(t[:200], j[:200])]
    assert rel(got, big["cfg5_k3"][:200]) < 1e-12
    k6 = O.approx_directed_network_loglikelihood(X, w["radii"], w["out_edges"], w["degrees"], w["ctrl_out"], 0.3, 0.7)
    assert rel(k6, big["cfg5_k6"]) < 1e-11
