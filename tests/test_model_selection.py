"""dynetlsm_b200.model_selection against the reference's model_selection package on synthetic
traces (CPU; needs the reference tree, i.e. runs in the build container)."""
import types

import numpy as np
import pytest

import pyoracle as O
import ref_shims
from dynetlsm_b200 import model_selection as MS

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference sources not present (GPU box)")


def _fake_model(directed, seed=0, S=40, T=4, n=30, d=2, K=6, nb=10):
    rng = np.random.RandomState(seed)
    scale = 1.0 / n if directed else 1.0
    m = types.SimpleNamespace()
    Y = (rng.rand(T, n, n) < 0.2).astype(np.float64)
    for t in range(T):
        np.fill_diagonal(Y[t], 0)
    if not directed:
        Y = np.triu(Y, 1); Y = Y + Y.transpose(0, 2, 1)
    m.Y_fit_, m.is_directed, m.n_components, m.n_features, m.n_burn_ = Y, directed, K, d, nb
    m.Xs_ = rng.randn(S, T, n, d) * scale
    m.intercepts_ = rng.rand(S, 2 if directed else 1) + 0.2
    m.radiis_ = rng.dirichlet(np.ones(n) * 5, size=S) if directed else None
    # labels: a few configurations repeated (ties), varying numbers of occupied components
    base = rng.randint(0, 3, (T, n))
    m.zs_ = np.stack([np.where(rng.rand(T, n) < 0.1 * (s % 4), rng.randint(0, K, (T, n)), base) for s in range(S)])
    m.zs_[nb + 3] = m.zs_[nb + 7] = base
    m.logps_ = rng.randn(S) * 10
    m.mus_ = rng.randn(S, K, d) * scale
    m.sigmas_ = rng.gamma(2, 1, (S, K)) * scale ** 2
    m.betas_ = rng.dirichlet(np.ones(K), size=S)
    m.weights_ = rng.dirichlet(np.ones(K), size=(S, T, K))
    m.lambdas_ = rng.rand(S, 1) * 0.5 + 0.4
    m.case_control_sampler_ = None
    eye = np.eye(K)
    m.cooccurrence_probas_ = np.stack([np.mean([eye[z[t]] @ eye[z[t]].T for z in m.zs_[nb:]], axis=0) for t in range(T)])
    return m


def _oracle_loglik(m):
    def fn(idx):
        X, b = m.Xs_[idx], m.intercepts_[idx]
        dist = O.calculate_distances(X)
        if m.is_directed:
            return O.directed_network_loglikelihood(m.Y_fit_, dist, m.radiis_[idx], b[0], b[1])
        return O.undirected_network_loglikelihood(m.Y_fit_, dist, b[0])
    return fn


@pytest.mark.parametrize("directed", [False, True])
def test_vi_and_bic_selection_equal_the_reference(directed):
    ref_shims.load_reference()
    from dynetlsm.model_selection.posterior_vi import (minimize_posterior_expected_vi,
                                                       time_averaged_posterior_expected_vi)
    from dynetlsm.model_selection.approx_bic import select_bic, calculate_cluster_counts_t
    from dynetlsm.label_utils import calculate_posterior_group_counts, renormalize_weights
    m = _fake_model(directed, seed=3 if directed else 1)
    nb = m.n_burn_
    vis = MS.expected_vi_trace(m.zs_[nb:], m.cooccurrence_probas_, m.n_components)
    want = np.array([time_averaged_posterior_expected_vi(z, m.cooccurrence_probas_) for z in m.zs_[nb:]])
    assert np.allclose(vis, want, rtol=1e-12, atol=1e-12)
    assert MS.minimize_posterior_expected_vi(m, _oracle_loglik(m)) == minimize_posterior_expected_vi(m)
    bic, models, counts = MS.select_bic(m, _oracle_loglik(m))
    rbic, rmodels, rcounts = select_bic(m)
    assert np.array_equal(counts, rcounts)
    assert np.array_equal(bic[:, [0, 3]], rbic[:, [0, 3]])
    assert np.allclose(bic[:, 1:3], rbic[:, 1:3], rtol=1e-10)
    for a, b in zip(models, rmodels):
        for f in ("beta", "init_weights", "trans_weights", "mu", "sigma", "X", "z"):
            assert np.allclose(getattr(a, f), getattr(b, f), rtol=1e-13, atol=0)
    ct = MS.cluster_counts_t(m.zs_[nb:], m.n_components)
    assert np.array_equal(ct, calculate_cluster_counts_t(m))
    for t in range(m.Y_fit_.shape[0]):
        gi, gc = MS.posterior_group_counts(ct[t])
        ri, rc = calculate_posterior_group_counts(m, t=t)
        assert np.array_equal(gi, ri) and np.array_equal(gc, rc)
    got = MS.renormalized(m, nb + 2)
    for a, b in zip(got, renormalize_weights(m, sample_id=nb + 2)):
        assert np.allclose(a, b, rtol=1e-13, atol=0)
