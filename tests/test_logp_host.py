"""The HDP estimator's joint log-posterior on the host (``hdp_updates.hdp_log_prior``) against the
reference's own ``logps_`` (hdp_lpcm.py:1188-1280) recorded in tests/golden/hdp_*.npz: network
log-likelihood from the C oracle (K4 / K5, themselves pinned on the reference's kernels) + the
product's prior terms must reproduce every stored sample's value, rtol 1e-10.  Together with
tests/test_gpu_trace.py (device ``k_logp`` == ``hdp_log_prior``) this anchors the device
log-posterior of the mixture model directly on reference outputs.  No GPU, no reference tree.
"""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))


@pytest.mark.parametrize("name,directed", [("hdp_undirected_split.npz", False),
                                           ("hdp_directed_monks.npz", True)])
def test_hdp_joint_log_posterior_matches_the_reference_logps(name, directed):
    import pyoracle as O
    from dynetlsm_b200.hdp_updates import HDPHyper, hdp_log_prior
    g = load_golden(name)
    S, T, n, d = g["X_centered"].shape
    K = g["sigma_next"].shape[1]
    a, a0, b0, c0, d0, lam0, lamv = [float(v) for v in g["hdp_prior"]]
    Y = g["Y"].astype(np.float64)
    ip, ivp = g["intercept_prior"], float(g["intercept_variance_prior"])
    for s in range(S):
        gamma, alpha_init, alpha, kappa, mvp, b = [float(v) for v in g["hyper_next"][s]]
        hp = HDPHyper(gamma, alpha_init, alpha, kappa, np.array([mvp]), b, a, a0, b0, c0, d0, lam0, lamv,
                      1.0, 0.1, 1.0, 1.0, 5, 0.1, True, True)
        X, ic = g["X_centered"][s], g["intercept_out"][s]
        dist = O.calculate_distances(X)
        if directed:
            radii = g["radii_logp"][s]
            ll = O.directed_network_loglikelihood(Y, dist, radii, ic[0], ic[1])
        else:
            radii = None
            ll = O.undirected_network_loglikelihood(Y, dist, ic[0])
        lp = hdp_log_prior(hp, K, X, ic, ip, ivp, g["mu_next"][s], g["sigma_next"][s],
                           g["z_out"][s].astype(np.int64), g["w_next"][s], g["beta_next"][s],
                           g["lmbda_next"][s], radii=radii)
        got = float(np.ravel(ll + lp)[0])
        assert np.isclose(got, g["logp"][s], rtol=1e-10, atol=0), (s, got, g["logp"][s])
        assert g["logp"][s] == g["logps"][s + 1]


@pytest.mark.parametrize("name", ["hdp_undirected_split.npz", "hdp_directed_monks.npz"])
def test_hdp_conjugate_block_reproduces_the_reference_draws(name):
    """The host restatement of hdp_lpcm.py:881-1023 (``hdp_updates.conjugate_updates``: auxiliary
    tables, beta, initial / transition weights, means, variances, lambda, scale hyper-priors,
    concentration parameters) on the reference's inputs and RandomState of every recorded sweep:
    every draw bit for bit.  (Sweep 0 is skipped: the initial beta is not part of the record.)"""
    from dynetlsm_b200.hdp_updates import HDPHyper, conjugate_updates
    g = load_golden(name)
    S = g["X_centered"].shape[0]
    a, a0, b0, c0, d0, lam0, lamv = [float(v) for v in g["hdp_prior"]]
    for s in range(1, S):
        rng = np.random.RandomState(0)
        rng.set_state(("MT19937", g["rng_keys"][s], int(g["rng_pos"][s]),
                       int(g["rng_has_gauss"][s]), float(g["rng_gauss"][s])))
        gamma, alpha_init, alpha, kappa, mvp, b = [float(v) for v in g["hyper_next"][s - 1]]
        hp = HDPHyper(gamma, alpha_init, alpha, kappa, np.array([mvp]), b, a, a0, b0, c0, d0, lam0, lamv,
                      1.0, 0.1, 1.0, 1.0, 5, 0.1, True, True)
        mu, sigma, w = g["mu"][s].copy(), g["sigma"][s].copy(), g["w"][s].copy()
        beta, lm = conjugate_updates(rng, hp, g["X_centered"][s], g["z_out"][s].astype(np.int64),
                                     g["n_out"][s], g["nk_out"][s].astype(np.int64), mu, sigma,
                                     np.atleast_1d(g["lmbda"][s]).astype(np.float64).copy(),
                                     g["beta_next"][s - 1].copy(), w)
        assert np.array_equal(beta, g["beta_next"][s]), s
        assert np.array_equal(w, g["w_next"][s]), s
        assert np.array_equal(mu, g["mu_next"][s]) and np.array_equal(sigma, g["sigma_next"][s]), s
        assert np.array_equal(np.ravel(lm), np.ravel(g["lmbda_next"][s])), s
        got = np.array([np.ravel(v)[0] for v in (hp.gamma, hp.alpha_init, hp.alpha, hp.kappa,
                                                 hp.mean_variance_prior, hp.b)])
        assert np.array_equal(got, g["hyper_next"][s]), (s, got, g["hyper_next"][s])


@pytest.mark.parametrize("name,directed", [("lsm_undirected_monks.npz", False),
                                           ("lsm_directed_monks.npz", True)])
def test_lsm_joint_log_posterior_matches_the_reference_logps(name, directed):
    """Same pin for the plain latent space model (lsm.py:576-625): C-oracle network term +
    ``DynamicNetworkLSM._log_prior`` (what the replay loop stores in ``logps_``) on the reference's
    stored samples."""
    import pyoracle as O
    from dynetlsm_b200 import DynamicNetworkLSM
    g = load_golden(name)
    Y = g["Y"].astype(np.float64)
    m = DynamicNetworkLSM(is_directed=directed, tau_sq=float(g["tau_sq"]), sigma_sq=float(g["sigma_sq"]),
                          intercept_prior=np.asarray(g["intercept_prior"], dtype=np.float64),
                          intercept_variance_prior=float(g["intercept_variance_prior"]))
    Xs, ics, want = g["Xs"], g["intercepts"], g["logps"]
    for s in range(min(len(want), 40)):
        dist = O.calculate_distances(Xs[s])
        if directed:
            ll = O.directed_network_loglikelihood(Y, dist, g["radiis"][s], ics[s, 0], ics[s, 1])
        else:
            ll = O.undirected_network_loglikelihood(Y, dist, ics[s, 0])
        got = float(np.ravel(ll + m._log_prior(Xs[s], ics[s]))[0])
        assert np.isclose(got, want[s], rtol=1e-10, atol=0), (s, got, want[s])
