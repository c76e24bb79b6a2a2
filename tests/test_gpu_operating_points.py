"""Parity AT THE OPERATING POINTS: the workloads bench.py measures (workloads.make_workload: cfg 3
n = 2 000 / T = 20 directed, cfg 4 n = 500 / T = 10 mixture on many chains, cfg 5 n = 50 000
case-control with per-chain control sets), replayed through the oracle with recorded draws --
exact decisions and states -- and the reference's known answers (tests/golden/kernels_big.npz,
>= 1 000 (state, node) pairs per likelihood, 1e-10 relative).  Needs a GPU: ``-m gpu``.
"""
import numpy as np
import pytest

import pyoracle as O
import workloads as W
from conftest import load_golden
from make_golden_big import state_of

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _L():
    from dynetlsm_b200 import _lib
    return _lib


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.fixture(scope="module")
def big():
    return load_golden("kernels_big.npz")


# ------------------------------------------------------------------------------------------
# known answers of the reference's Cython kernels at n = 120 / 500 / 2 000 / 50 000
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg2", "cfg4"])
def test_kat_k1_k5_at_bench_sizes(big, name):
    L = _L()
    w = W.make_workload(name)
    X, t, j = state_of(w, int(big[name + "_seed"]))
    e = L.Engine(T=w["T"], n=w["n"], d=2)
    e.set_network(w["Y"])
    e.set(L.F_X, X[None])
    e.set(L.F_INTERCEPT, np.array([[w["intercept"][0], 0.0]]))
    assert rel(e.loglik_partial()[0][t, j], big[name + "_k1"]) < RTOL
    assert rel(e.loglik_full()[0], big[name + "_k5"]) < RTOL


def test_kat_k2_k4_directed_n2000(big):
    L = _L()
    w = W.make_workload("cfg3")
    X, t, j = state_of(w, int(big["cfg3_seed"]))
    e = L.Engine(T=w["T"], n=w["n"], d=2, is_directed=True)
    e.set_network(w["Y"])
    e.set(L.F_X, X[None])
    e.set(L.F_RADII, w["radii"][None])
    e.set(L.F_INTERCEPT, w["intercept"][None])
    assert rel(e.loglik_partial()[0][t, j], big["cfg3_k2"]) < RTOL
    assert rel(e.loglik_full()[0], big["cfg3_k4"]) < RTOL


def test_kat_k3_k6_case_control_n50000(big):
    L = _L()
    w = W.make_workload("cfg5")
    X, t, j = state_of(w, int(big["cfg5_seed"]))
    e = L.Engine(T=w["T"], n=w["n"], d=2, is_directed=True, case_control=True)
    e.set_edge_lists(w["degrees"], w["in_edges"], w["out_edges"])
    e.set_controls(w["ctrl_in"], w["ctrl_out"])
    e.set(L.F_X, X[None])
    e.set(L.F_RADII, w["radii"][None])
    e.set(L.F_INTERCEPT, np.array([[0.3, 0.7]]))
    assert rel(e.loglik_partial()[0][t, j], big["cfg5_k3"]) < RTOL
    assert rel(e.loglik_full()[0], big["cfg5_k6"]) < RTOL
    assert e.counters()["ub_flags"] == 0


# ------------------------------------------------------------------------------------------
# cfg 3: one full sweep (40 000 node-updates) of the single n = 2 000, T = 20 chain, recorded
# draws, the kernel the heuristic picks for it; then centring, both intercept MH steps and the
# radii MH step
# ------------------------------------------------------------------------------------------
def test_cfg3_full_sweep_replay_vs_oracle():
    L = _L()
    w = W.make_workload("cfg3")
    T, n, d = w["T"], w["n"], w["d"]
    rng = np.random.RandomState(303)
    X0 = W.chain_starts(w, 1, 0)[0]
    step = w["step_X"] / 40.0      # ~25 % acceptance at this state (the bench tunes its step the same way)
    e = L.Engine(T=T, n=n, d=d, is_directed=True, tune=4, tune_interval=2, intercept_tune_interval=(2, 2))
    e.set_network(w["Y"])
    e.set_hyper(tau_sq=w["tau_sq"], sigma_sq=w["sigma_sq"], intercept_prior=w["intercept"],
                intercept_variance_prior=2.0)
    e.set(L.F_X, X0[None])
    e.set(L.F_INTERCEPT, w["intercept"][None])
    e.set(L.F_RADII, w["radii"][None])
    e.set_tuner(step)
    Xo, ic, radii = X0.copy(), w["intercept"].copy(), w["radii"].copy()
    tun = O.TunerState((T, n), step, tune=4, tune_interval=2)
    itun = O.TunerState((2,), 0.1, tune=4, tune_interval=2)
    rtun = O.TunerState((1,), 175000.0, tune=None, tune_interval=100)
    for s in range(2):
        eps, logu = rng.randn(T, n, d), np.log(rng.rand(T, n))
        out = O.sweep_latent(Xo, ic, tun, eps, logu, Y=w["Y"], radii=radii, is_directed=True,
                             tau_sq=w["tau_sq"], sigma_sq=w["sigma_sq"])
        acc, ratio = e.sweep_latent(eps[None], logu[None], want_stats=True)
        assert np.array_equal(acc[0], out["accepted"]), "sweep %d: %d decisions differ" % (
            s, int((acc[0] != out["accepted"]).sum()))
        assert np.array_equal(e.get(L.F_X)[0], Xo)
        assert np.allclose(ratio[0], out["ratio"], rtol=1e-8, atol=1e-8)
        O.center(Xo)
        e.center()
        assert np.array_equal(e.get(L.F_X)[0], Xo)
        dist = O.calculate_distances(Xo)
        ieps, ilogu = rng.randn(2), np.log(rng.rand(2))
        io = O.sample_intercepts(Xo, ic, itun, ieps, ilogu, w["intercept"], 2.0, Y=w["Y"], dist=dist,
                                 radii=radii, is_directed=True)
        iacc, _ = e.sample_intercepts(ieps[None], ilogu[None], want_stats=True)
        assert np.array_equal(iacc[0], io["accepted"])
        assert np.array_equal(e.get(L.F_INTERCEPT)[0], ic)
        prop = rng.dirichlet(175000.0 * radii)
        rlogu = float(np.log(rng.rand()))
        ro = O.sample_radii(Xo, ic, radii, rtun, prop, rlogu, Y=w["Y"], dist=dist)
        racc, _ = e.sample_radii(prop[None], np.array([rlogu]), want_stats=True)
        assert int(racc[0]) == ro["accepted"]
        assert np.array_equal(e.get(L.F_RADII)[0], radii)
    assert 0.01 < out["accepted"].mean() < 0.99
    assert np.array_equal(e.get(L.F_X_STEP)[0], tun.step)
    assert e.counters()["cluster_sweeps"] == 2     # the thread-block-cluster kernel served this chain


# ------------------------------------------------------------------------------------------
# cfg 5: one full sweep of two n = 50 000 chains with their OWN control sets drawn on the device
# (k_resample_controls -> k_cc_deps -> k_sweep_cc), the oracle walks the same lists
# ------------------------------------------------------------------------------------------
def test_cfg5_full_sweep_two_chains_per_chain_controls_vs_oracle():
    L = _L()
    w = W.make_workload("cfg5")
    T, n, d, C_ = w["T"], w["n"], w["d"], 2
    rng = np.random.RandomState(505)
    X0 = W.chain_starts(w, C_, 0)
    e = L.Engine(T=T, n=n, d=d, n_chains=C_, is_directed=True, case_control=True, tune=4, tune_interval=2)
    e.set_edge_lists(w["degrees"], w["in_edges"], w["out_edges"])
    e.set_rng(42)
    e.resample_controls(w["n_control"], per_chain=True)
    ci, co = e.get_controls()
    assert ci.shape == (C_, T, n, w["n_control"]) and not np.array_equal(ci[0], ci[1])
    e.set_hyper(tau_sq=w["tau_sq"], sigma_sq=w["sigma_sq"], intercept_prior=w["intercept"],
                intercept_variance_prior=2.0)
    e.set(L.F_X, X0)
    e.set(L.F_INTERCEPT, np.tile(w["intercept"][None], (C_, 1)))
    e.set(L.F_RADII, np.tile(w["radii"][None], (C_, 1)))
    e.set_tuner(w["step_X"])
    eps, logu = rng.randn(C_, T, n, d), np.log(rng.rand(C_, T, n))
    acc, ratio = e.sweep_latent(eps, logu, want_stats=True)
    got = e.get(L.F_X)
    for c in range(C_):
        Xo = X0[c].copy()
        tun = O.TunerState((T, n), w["step_X"], tune=4, tune_interval=2)
        cc = dict(in_edges=w["in_edges"], out_edges=w["out_edges"], degrees=w["degrees"],
                  ctrl_in=ci[c], ctrl_out=co[c])
        out = O.sweep_latent(Xo, w["intercept"], tun, eps[c], logu[c], radii=w["radii"], is_directed=True,
                             tau_sq=w["tau_sq"], sigma_sq=w["sigma_sq"], case_control=cc)
        assert np.array_equal(acc[c], out["accepted"]), "chain %d: %d decisions differ" % (
            c, int((acc[c] != out["accepted"]).sum()))
        assert np.array_equal(got[c], Xo)
        assert 0.02 < out["accepted"].mean() < 0.98
    assert e.counters()["ub_flags"] == 0
    # the device loop's full-network evaluation (gather-packed 32-byte records, one variant) against
    # the two-variant probe and the oracle
    e.run_sweeps(1, skip_radii=True)
    tracked, fresh = e.get(L.F_LOGLIK), e.loglik_full()
    Xn, icn = e.get(L.F_X), e.get(L.F_INTERCEPT)
    for c in range(C_):
        k6 = O.approx_directed_network_loglikelihood(Xn[c], w["radii"], w["out_edges"], w["degrees"], co[c],
                                                     icn[c, 0], icn[c, 1])
        assert abs(fresh[c] - k6) <= RTOL * abs(k6)
        assert abs(tracked[c] - k6) <= RTOL * abs(k6)


# ------------------------------------------------------------------------------------------
# cfg 4: the mixture prior on the many-chains build (more chains than SMs, ~100 KB of positions
# each: the (320, 2) instantiation), recorded draws, then centring and the label FFBS
# ------------------------------------------------------------------------------------------
def test_cfg4_mixture_many_chains_sweep_and_labels_vs_oracle():
    L = _L()
    w = W.make_workload("cfg4")
    T, n, d, K, C_ = w["T"], w["n"], w["d"], w["K"], 160
    rng = np.random.RandomState(404)
    e = W.build_engine(w, C_, 0, 0)
    X0 = e.get(L.F_X)
    check = (0, 81, 159)
    eps, logu = rng.randn(C_, T, n, d), np.log(rng.rand(C_, T, n))
    U = rng.rand(C_, n, T)
    acc, _ = e.sweep_latent(eps, logu, want_stats=True)
    got = e.get(L.F_X)
    e.center()
    gotc = e.get(L.F_X)
    e.sample_labels(U)
    z, nc, nk = e.get(L.F_Z), e.get(L.F_NCOUNT), e.get(L.F_NK)
    for c in check:
        Xo = X0[c].copy()
        tun = O.TunerState((T, n), w["step_X"], tune=2500, tune_interval=100)
        out = O.sweep_latent(Xo, w["intercept"], tun, eps[c], logu[c], Y=w["Y"], tau_sq=w["tau_sq"],
                             sigma_sq=w["sigma_sq"],
                             mixture=dict(mu=w["mu"], sigma=w["sigma"], lmbda=w["lmbda"], z=w["z"]))
        assert np.array_equal(acc[c], out["accepted"]), "chain %d" % c
        assert np.array_equal(got[c], Xo)
        O.center(Xo)
        assert np.array_equal(gotc[c], Xo)
        zo, nco, nko, _ = O.sample_labels_block(Xo, w["mu"], w["sigma"], w["lmbda"], w["w"], U[c])
        assert np.array_equal(z[c], zo) and np.array_equal(nc[c], nco) and np.array_equal(nk[c], nko)
    assert 0.02 < acc.mean() < 0.98
