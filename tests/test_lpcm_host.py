"""Host block of ``DynamicNetworkLPCM`` (dynetlsm_b200/lpcm.py) against the reference chain recorded by
oracle/make_golden_lpcm.py: at every sweep the reference's inputs to its conjugate block
(lpcm.py:573-656) and its RandomState are restored, the product's block is run, and every draw --
initial / transition weights, means, variances, lambda, the two scale hyper-priors -- must come out
bit for bit; the joint log-posterior's prior part is checked through the stored ``logps``.
No GPU, no reference tree: the golden file carries everything.
"""
import numpy as np
import pytest

from conftest import load_golden


def _hyper(g, directed, n, d=2):
    from dynetlsm_b200.lpcm import MixtureHyper
    mvp0 = (2 * (1. / n) ** (2. / d)) if directed else ((n ** (2. / d)) / 50.)
    b0_ = (float(g["a"]) + 2) * mvp0
    return MixtureHyper(float(g["dirichlet_prior_"]), mvp0, b0_, float(g["a"]), float(g["a0_"]),
                        float(g["b0_"]), float(g["c0_"]), float(g["d0_"]), 0.9, 0.01, True, True)


@pytest.mark.parametrize("name,directed", [("lpcm_undirected_split.npz", False),
                                           ("lpcm_directed_monks.npz", True)])
def test_conjugate_block_reproduces_the_reference_draws(name, directed):
    from dynetlsm_b200.lpcm import lpcm_conjugate_updates
    g = load_golden(name)
    S, T, n, d = g["X_centered"].shape
    K = g["sigma_in"].shape[1]
    hp = _hyper(g, directed, n, d)
    init_prev = trans_prev = None
    for s in range(S):
        rng = np.random.RandomState(0)
        rng.set_state(("MT19937", g["rng_keys"][s], int(g["rng_pos"][s]),
                       int(g["rng_has_gauss"][s]), float(g["rng_gauss"][s])))
        if s > 0:       # the hyper-priors entering sweep s are the ones sweep s-1 drew
            hp.mean_variance_prior = np.array([g["mean_variance_prior"][s - 1]])
            hp.b = float(g["b"][s - 1])
        mu, sigma = g["mu_in"][s].copy(), g["sigma_in"][s].copy()
        init_w = np.zeros(K)
        trans_w = np.zeros((K, K))
        lm = lpcm_conjugate_updates(rng, hp, g["X_centered"][s], g["z_out"][s].astype(np.int64),
                                    g["n_out"][s], g["nk_out"][s].astype(np.int64), mu, sigma,
                                    g["lmbda_in"][s].copy(), init_w, trans_w)
        assert np.array_equal(init_w, g["init_next"][s])
        assert np.array_equal(trans_w, g["trans_next"][s])
        assert np.array_equal(mu, g["mu_next"][s])
        assert np.array_equal(sigma, g["sigma_next"][s])
        assert np.array_equal(np.ravel(lm), np.ravel(g["lmbda_next"][s]))
        assert np.ravel(hp.mean_variance_prior)[0] == g["mean_variance_prior"][s]
        assert float(hp.b) == g["b"][s]


def test_stacked_weights_feed_the_hdp_label_kernel_layout():
    from dynetlsm_b200.lpcm import stacked_weights
    iw = np.array([0.2, 0.8])
    tw = np.array([[0.9, 0.1], [0.3, 0.7]])
    w = stacked_weights(iw, tw, 3)
    assert w.shape == (3, 2, 2)
    assert np.array_equal(w[0, 0], iw) and np.all(w[0, 1] == 0)
    assert np.array_equal(w[1], tw) and np.array_equal(w[2], tw)


@pytest.mark.parametrize("name,directed", [("lpcm_undirected_split.npz", False),
                                           ("lpcm_directed_monks.npz", True)])
def test_joint_log_posterior_matches_the_reference_logps(name, directed):
    """logps_[s] of the reference (lpcm.py:770-856) = network log-likelihood + every prior term.  The
    network term comes from the C oracle (K4 / K5, pinned on the reference's kernels in
    tests/test_oracle_golden.py); adding the product's ``lpcm_log_prior`` must give the recorded
    log-posterior of every stored sample, rtol 1e-10."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyoracle as O
    from dynetlsm_b200.lpcm import lpcm_log_prior
    g = load_golden(name)
    S, T, n, d = g["X_centered"].shape
    hp = _hyper(g, directed, n, d)
    Y = g["Y"].astype(np.float64)
    ip, ivp = g["intercept_prior"], float(g["intercept_variance_prior"])
    for s in range(S):
        hp.mean_variance_prior = np.array([g["mean_variance_prior"][s]])
        hp.b = float(g["b"][s])
        X, ic = g["X_centered"][s], g["intercept_out"][s]
        dist = O.calculate_distances(X)
        if directed:
            ll = O.directed_network_loglikelihood(Y, dist, g["radii_out"][s], ic[0], ic[1])
        else:
            ll = O.undirected_network_loglikelihood(Y, dist, ic[0])
        lp = lpcm_log_prior(hp, X, ic, ip, ivp, g["mu_next"][s], g["sigma_next"][s],
                            g["z_out"][s].astype(np.int64), g["init_next"][s], g["trans_next"][s],
                            g["lmbda_next"][s], radii=g["radii_out"][s] if directed else None)
        got = float(np.ravel(ll + lp)[0])
        assert np.isclose(got, g["logp"][s], rtol=1e-10, atol=0), (s, got, g["logp"][s])
        assert g["logp"][s] == g["logps"][s + 1]
