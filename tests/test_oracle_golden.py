"""The CPU oracle (oracle/oracle.c) against golden vectors produced by the reference itself
(oracle/make_golden.py, run from /root/reference in the build container).  CPU-only."""
import numpy as np
import pytest

import pyoracle as O
from conftest import load_golden

TAGS = ["a", "b", "c"]


def _ulp_close(a, b, ulps=4):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b))))


@pytest.mark.parametrize("m", [1, 5, 8, 10, 17, 128, 129, 1000, 4099])
def test_numpy_pairwise_sum_bitwise(kernels_golden, m):
    g = kernels_golden
    assert O.np_sum(g["sum_%d_in" % m]) == float(g["sum_%d_out" % m])


def test_center_bitwise(kernels_golden):
    X = kernels_golden["center_in"].copy()
    O.center(X)
    assert np.array_equal(X, kernels_golden["center_out"])


@pytest.mark.parametrize("tag", TAGS)
def test_k1_partial_loglikelihood(kernels_golden, tag):
    g = kernels_golden
    X, Y, b = g[tag + "_X"], g[tag + "_Yu"].astype(np.float64), g[tag + "_b"][0]
    T, n, _ = X.shape
    got = np.array([[O.partial_loglikelihood(Y[t], X[t], b, i) for i in range(n)] for t in range(T)])
    # same libm, same operation order: bitwise
    assert np.array_equal(got, g[tag + "_k1"])
    got = np.array([[O.partial_loglikelihood(Y[t], X[t], b, i, squared=True) for i in range(n)]
                    for t in range(T)])
    assert np.array_equal(got, g[tag + "_k1sq"])


@pytest.mark.parametrize("tag", TAGS)
def test_k2_directed_partial(kernels_golden, tag):
    g = kernels_golden
    X, Y, r = g[tag + "_X"], g[tag + "_Yd"].astype(np.float64), g[tag + "_radii"]
    T, n, _ = X.shape
    Xd = X / n
    _, b_in, b_out = g[tag + "_b"]
    got = np.array([[O.directed_partial_loglikelihood(Y[t], Xd[t], r, b_in, b_out, i)
                     for i in range(n)] for t in range(T)])
    assert np.array_equal(got, g[tag + "_k2"])


@pytest.mark.parametrize("tag", TAGS)
def test_k3_case_control_partial(kernels_golden, tag):
    g = kernels_golden
    X, r = g[tag + "_X"], g[tag + "_radii"]
    T, n, _ = X.shape
    Xd = X / n
    _, b_in, b_out = g[tag + "_b"]
    ok = g[tag + "_cc_ok"]
    assert ok.sum() > 0 and (g[tag + "_ctrl_in"] == -1).any()
    for t in range(T):
        for i in range(n):
            v, ub = O.approx_directed_partial_loglikelihood(
                Xd[t], r, g[tag + "_in_edges"][t], g[tag + "_out_edges"][t], g[tag + "_degrees"][t],
                g[tag + "_ctrl_in"][t], g[tag + "_ctrl_out"][t], b_in, b_out, i, return_ub=True)
            if ok[t, i]:
                assert ub == 0
                ref = g[tag + "_k3"][t, i]
                # a node with no non-neighbours gives 0/0 = NaN in the reference too
                assert v == ref or (np.isnan(v) and np.isnan(ref))
            else:
                assert ub == 1  # the reference's out-of-bounds read: flagged, not replicated


@pytest.mark.parametrize("tag", TAGS)
def test_k4_k5_k6_full_network(kernels_golden, tag):
    g = kernels_golden
    X, r = g[tag + "_X"], g[tag + "_radii"]
    n = X.shape[1]
    b, b_in, b_out = g[tag + "_b"]
    Yd, Yu = g[tag + "_Yd"].astype(np.float64), g[tag + "_Yu"].astype(np.float64)
    # with the reference's own distance cache: bitwise (K4 serial order, K5 numpy pairwise order)
    assert O.directed_network_loglikelihood(Yd, g[tag + "_distd"], r, b_in, b_out) == float(g[tag + "_k4"])
    k5 = O.undirected_network_loglikelihood(Yu, g[tag + "_dist"], b)
    assert _ulp_close(k5, float(g[tag + "_k5"]), 2)  # numpy's SIMD exp/log vs libm: last-ulp terms
    k6 = O.approx_directed_network_loglikelihood(X / n, r, g[tag + "_out_edges"], g[tag + "_degrees"],
                                                 g[tag + "_ctrl_out"], b_in, b_out)
    assert k6 == float(g[tag + "_k6"])
    # the restated sklearn distance formula (BLAS rounding is not reproducible): tolerance
    d = O.calculate_distances(X)
    assert np.allclose(d, g[tag + "_dist"], rtol=0, atol=1e-7)  # cancellation near zero distance
    k5b = O.undirected_network_loglikelihood(Yu, d, b)
    assert abs(k5b - float(g[tag + "_k5"])) <= 1e-10 * abs(float(g[tag + "_k5"]))


@pytest.mark.parametrize("tag", TAGS)
def test_k7_gaussian_likelihood(kernels_golden, tag):
    g = kernels_golden
    X = g[tag + "_X"]
    n = X.shape[1]
    for i in range(n):
        xi = np.ascontiguousarray(X[:, i])
        got = O.compute_gaussian_likelihood(xi, g[tag + "_mu"], g[tag + "_sigma"], g[tag + "_lmbda"],
                                            normalize=False)
        assert _ulp_close(got, g[tag + "_k7"][i], 2)  # numpy np.exp vs libm exp
        got = O.compute_gaussian_likelihood(xi, g[tag + "_mu"], g[tag + "_sigma"], g[tag + "_lmbda"],
                                            normalize=True)
        assert _ulp_close(got, g[tag + "_k7n"][i], 2)


def _tuner_from(g, s, prefix="tuner_", shape=None):
    t = O.TunerState.__new__(O.TunerState)
    t.step = np.ascontiguousarray(g[prefix + "step"][s], dtype=np.float64).copy()
    t.n_accepted = np.ascontiguousarray(g[prefix + "n_accepted"][s], dtype=np.int32).copy()
    t.n_steps = np.ascontiguousarray(g[prefix + "n_steps"][s], dtype=np.int32).copy()
    t.until = np.ascontiguousarray(g[prefix + "until"][s], dtype=np.int32).copy()
    t.tune = int(g["tune"])
    t.tune_interval = int(g["tune_interval"])
    return t


def _cc(g, s):
    if "cc_in_edges" not in g:
        return None
    return dict(in_edges=g["cc_in_edges"], out_edges=g["cc_out_edges"], degrees=g["cc_degrees"],
                ctrl_in=g["ctrl_in"][s], ctrl_out=g["ctrl_out"][s])


@pytest.mark.parametrize("name,directed", [("lsm_undirected_monks.npz", False),
                                           ("lsm_directed_monks.npz", True),
                                           ("lsm_casecontrol_monks.npz", True)])
def test_lsm_sweeps_teacher_forced(name, directed):
    """Every recorded reference sweep, replayed by the oracle from the recorded input state and
    raw draws: identical decisions, bit-identical states and tuner trajectories."""
    g = load_golden(name)
    S = g["X_in"].shape[0]
    Y = g["Y"].astype(np.float64)
    for s in range(S):
        X = g["X_in"][s].copy()
        tun = _tuner_from(g, s)
        out = O.sweep_latent(X, g["intercept_in"][s], tun, g["eps"][s], g["logu"][s], Y=Y,
                             radii=g["radii_in"][s] if directed else None, is_directed=directed,
                             tau_sq=float(g["tau_sq"]), sigma_sq=float(g["sigma_sq"]),
                             case_control=_cc(g, s))
        assert np.array_equal(out["accepted"], g["accepted"][s])
        assert np.array_equal(X, g["X_out"][s])
        assert np.array_equal(out["logp_new"], g["lp_new"][s])
        assert np.array_equal(out["logp_old"], g["lp_old"][s])
        if s + 1 < S:
            assert np.array_equal(tun.step, g["tuner_step"][s + 1])
            assert np.array_equal(tun.n_accepted, g["tuner_n_accepted"][s + 1])
            assert np.array_equal(tun.until, g["tuner_until"][s + 1])
    assert len(np.unique(g["tuner_step"])) > 3  # the tuner really fired in this fixture


@pytest.mark.parametrize("name,directed", [("lsm_undirected_monks.npz", False),
                                           ("lsm_directed_monks.npz", True),
                                           ("lsm_casecontrol_monks.npz", True)])
def test_lsm_intercepts_and_radii(name, directed):
    g = load_golden(name)
    S = g["X_in"].shape[0]
    Y = g["Y"].astype(np.float64)
    m = 2 if directed else 1
    for s in range(S):
        Xc = g["X_centered"][s]
        cc = _cc(g, s)
        dist = None if cc is not None else O.calculate_distances(Xc)
        ic = g["intercept_in"][s].copy()
        tun = _tuner_from(g, s, "itun_")
        # lsm.py:465-467: the undirected intercept sampler ignores tune_interval (default 100)
        tun.intervals = [int(g["tune_interval"])] * 2 if directed else [100]
        radii = g["radii_in"][s] if directed else None
        out = O.sample_intercepts(Xc, ic, tun, g["i_eps"][s], g["i_logu"][s], g["intercept_prior"],
                                  float(g["intercept_variance_prior"]), Y=Y, dist=dist, radii=radii,
                                  is_directed=directed, case_control=cc)
        assert np.array_equal(out["accepted"], g["i_accepted"][s])
        assert np.array_equal(ic, g["intercept_out"][s])
        assert np.allclose(out["ratio"], g["i_ratio"][s], rtol=0, atol=1e-8)
        if s + 1 < S:
            assert np.array_equal(tun.step, g["itun_step"][s + 1])
        if directed:
            r = g["radii_in"][s].copy()
            rt = _tuner_from(g, s, "rtun_")
            rt.tune = -1  # lsm.py:470-472 tune=None
            o2 = O.sample_radii(Xc, ic, r, rt, g["r_proposal"][s], float(g["r_logu"][s]), Y=Y,
                                dist=dist, case_control=cc)
            assert o2["accepted"] == int(g["r_accepted"][s])
            assert abs(o2["ratio"] - float(g["r_ratio"][s])) <= 1e-7 * max(1.0, abs(float(g["r_ratio"][s])))
            assert np.array_equal(r, g["radii_out"][s])


def test_lsm_center_matches_reference_before_procrustes():
    g = load_golden("lsm_undirected_monks.npz")
    n_pre = int(g["tune"]) + int(g["burn"])  # lsm.py:495: Procrustes only when it > tune+burn
    for s in range(min(n_pre, g["X_in"].shape[0])):
        X = g["X_out"][s].copy()
        O.center(X)
        assert np.array_equal(X, g["X_centered"][s])


@pytest.mark.parametrize("name,directed", [("hdp_undirected_split.npz", False),
                                           ("hdp_directed_monks.npz", True)])
def test_hdp_sweeps_teacher_forced(name, directed):
    g = load_golden(name)
    S = g["X_in"].shape[0]
    Y = g["Y"].astype(np.float64)
    for s in range(S):
        X = g["X_in"][s].copy()
        tun = _tuner_from(g, s)
        mix = dict(mu=g["mu"][s], sigma=g["sigma"][s], lmbda=g["lmbda"][s], z=g["z_in"][s])
        out = O.sweep_latent(X, g["intercept_in"][s], tun, g["eps"][s], g["logu"][s], Y=Y,
                             radii=g["radii_in"][s] if directed else None, is_directed=directed,
                             mixture=mix)
        assert np.array_equal(out["accepted"], g["accepted"][s])
        assert np.array_equal(X, g["X_out"][s])
        assert np.array_equal(out["logp_new"], g["lp_new"][s])
        O.center(X)
        assert np.array_equal(X, g["X_centered"][s])
        # labels (FFBS) from the recorded uniforms
        z, nc, nk, resp, pr = O.sample_labels_block(g["X_centered"][s], g["mu"][s], g["sigma"][s],
                                                    g["lmbda"][s], g["w"][s], g["U"][s],
                                                    return_probas=True)
        assert np.array_equal(z, g["z_out"][s])
        assert np.array_equal(nc, g["n_out"][s])
        assert np.array_equal(nk, g["nk_out"][s])
        assert np.allclose(pr, g["probas"][s], rtol=1e-12, atol=0)
