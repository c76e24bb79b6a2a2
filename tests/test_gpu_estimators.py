"""The estimator shells (drop-in API of lsm.py / hdp_lpcm.py) on the GPU.

``sampler='replay'`` must reproduce the reference's chains recorded in tests/golden (whole ``fit``,
from the reference's own initialisation through every sweep); ``sampler='device'`` (Philox) must
agree with it in distribution.  Also the reference's own smoke tests, restated
(dynetlsm/tests/test_lsm.py:5-13, test_hdp_lcpm.py:5-15: fit and assert shapes).
"""
import warnings

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")


def _splitting_network(n=50, T=2, seed=42):
    """A small community-structured undirected network (stand-in for the reference generator)."""
    rng = np.random.RandomState(seed)
    z = rng.randint(0, 2, n)
    centers = np.array([[-1.5, 0.0], [1.5, 0.0]])
    Y = np.zeros((T, n, n))
    for t in range(T):
        X = centers[z] + 0.6 * rng.randn(n, 2)
        dist = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1))
        U = np.triu((rng.rand(n, n) < 1 / (1 + np.exp(-(1.0 - dist)))).astype(float), 1)
        Y[t] = U + U.T
    return Y


def test_reference_smoke_test_lsm_shapes():
    from dynetlsm_b200 import DynamicNetworkLSM
    Y = _splitting_network()
    lsm = DynamicNetworkLSM(n_iter=250, burn=250, tune=250, n_features=2, random_state=123).fit(Y)
    assert lsm.X_.shape == (2, 50, 2)
    assert lsm.Xs_.shape == (750, 2, 50, 2) and lsm.intercepts_.shape == (750, 1)
    assert np.isfinite(lsm.logps_).all() and 0.5 < lsm.auc_ <= 1.0
    assert lsm.probas_.shape == (2, 50, 50) and lsm.distances_.shape == (2, 50, 50)


def test_reference_smoke_test_hdp_lpcm_shapes():
    from dynetlsm_b200 import DynamicNetworkHDPLPCM
    Y = _splitting_network()
    m = DynamicNetworkHDPLPCM(n_iter=100, burn=100, tune=100, n_features=2, n_components=10,
                              random_state=123).fit(Y)
    assert m.X_.shape == (2, 50, 2)
    assert m.z_.shape == (2, 50)
    assert m.zs_.shape == (300, 2, 50) and m.weights_.shape == (300, 2, 10, 10)
    assert m.cooccurrence_probas_.shape == (2, 50, 50)
    assert np.allclose(np.diagonal(m.cooccurrence_probas_, axis1=1, axis2=2), 1.0)
    assert np.isfinite(m.logps_).all()


def test_lsm_replay_reproduces_the_reference_fit_undirected():
    """Whole fit(), reference initialisation included: same chain as the reference, draw for draw."""
    from dynetlsm_b200 import DynamicNetworkLSM
    g = load_golden("lsm_undirected_monks.npz")
    Y = g["Y"].astype(np.float64)
    m = DynamicNetworkLSM(n_iter=40, tune=30, burn=20, tune_interval=7, random_state=42,
                          sampler="replay").fit(Y)
    S = g["Xs"].shape[0]
    assert np.array_equal(m.Xs_[0], g["Xs"][0])                 # initial state
    assert np.array_equal(m.intercepts_[:S], g["intercepts"])   # every intercept draw
    assert np.array_equal(m.Xs_[:S], g["Xs"])                   # every position, bit for bit
    assert np.allclose(m.logps_[:S], g["logps"], rtol=1e-9, atol=0)


def test_lsm_replay_case_control_runs_and_matches_reference_control_sets():
    """Directed + case-control: initial values are not comparable (reference UB, K10), but the
    control-set resampling schedule and RNG consumption are: check the chain is self-consistent."""
    from dynetlsm_b200 import DynamicNetworkLSM
    g = load_golden("lsm_casecontrol_monks.npz")
    Y = g["Y"].astype(np.float64)
    m = DynamicNetworkLSM(n_iter=20, tune=20, burn=10, tune_interval=6, is_directed=True,
                          sigma_sq=0.001, tau_sq="auto", step_size_X=0.0075, n_control=5,
                          n_resample_control=8, random_state=11, sampler="replay").fit(Y)
    assert m.Xs_.shape == (50, 3, 18, 2) and m.radiis_.shape == (50, 18)
    assert np.allclose(m.radiis_.sum(axis=1), 1.0)
    assert np.isfinite(m.logps_).all()
    assert np.array_equal(m.case_control_sampler_.in_edges_, g["cc_in_edges"])
    assert m.sampler_counters_["ub_flags"] == 0


def test_hdp_replay_reproduces_the_reference_fit():
    """HDP-LPCM end to end in replay mode: the 999-sweep LSM initialisation, k-means, and the main
    loop (device sweeps + label FFBS, host conjugate updates) follow the reference chain."""
    from dynetlsm_b200 import DynamicNetworkHDPLPCM
    g = load_golden("hdp_undirected_split.npz")
    Y = g["Y"].astype(np.float64)
    m = DynamicNetworkHDPLPCM(n_iter=25, tune=25, burn=10, tune_interval=8, n_components=6,
                              random_state=3, sampler="replay").fit(Y)
    S = g["z_out"].shape[0]
    assert np.array_equal(m.zs_[1:S + 1], g["z_out"])                       # every label draw
    assert np.array_equal(m.intercepts_[1:S + 1], g["intercept_out"])       # every intercept draw
    assert np.array_equal(m.lambdas_[1:S + 1], g["lmbda_next"])             # host conjugate block
    assert np.array_equal(m.sigmas_[1:S + 1], g["sigma_next"])
    assert np.allclose(m.logps_[:S + 1], g["logps"], rtol=1e-9, atol=0)
    # positions are rotated post hoc (hdp_lpcm.py:1141-1146): compare a rotation invariant
    d_ref = np.linalg.norm(g["X_centered"][-1][0][:, None] - g["X_centered"][-1][0][None], axis=-1)
    d_got = np.linalg.norm(m.Xs_[S][0][:, None] - m.Xs_[S][0][None], axis=-1)
    assert np.allclose(d_ref, d_got, rtol=0, atol=1e-10)


def test_device_rng_matches_replay_in_distribution():
    """Native Philox chains vs reference-equivalent replay chains on the monks network: posterior
    means of the intercept and of the pairwise distances agree within Monte-Carlo error."""
    from dynetlsm_b200 import DynamicNetworkLSM
    from dynetlsm_b200.diagnostics import ess, split_rhat
    g = load_golden("lsm_undirected_monks.npz")
    Y = g["Y"].astype(np.float64)
    kw = dict(n_iter=3000, tune=500, burn=500)
    dev = DynamicNetworkLSM(random_state=1, sampler="device", n_chains=4, **kw).fit(Y)
    rep = DynamicNetworkLSM(random_state=2, sampler="replay", **kw).fit(Y)
    nb = 1000
    a = dev.chains_["intercepts"][:, nb:, 0]      # (4, 3000)
    b = rep.intercepts_[nb:, 0]
    se = np.sqrt(a.var() / max(ess(a), 10) + b.var() / max(ess(b[None]), 10))
    assert abs(a.mean() - b.mean()) < 5 * se + 0.02
    assert split_rhat(a) < 1.2
    # distances are invariant to the rotation/translation non-identifiability
    def mean_dist(Xs):
        return np.linalg.norm(Xs[:, :, :, None] - Xs[:, :, None], axis=-1).mean(axis=0)
    da = mean_dist(dev.chains_["Xs"][0, nb:])
    db = mean_dist(rep.Xs_[nb:])
    assert np.abs(da - db).mean() < 0.15 * db.mean()


def test_missing_dyads_are_refused():
    from dynetlsm_b200 import DynamicNetworkLSM
    Y = _splitting_network(n=12)
    Y[0, 1, 2] = Y[0, 2, 1] = -1
    with pytest.raises(NotImplementedError):
        DynamicNetworkLSM(n_iter=5, tune=5, burn=5).fit(Y)


def test_hdp_device_chains_match_replay_in_distribution():
    """HDP-LPCM: device-resident chains (Philox, device conjugate block, device log-posterior and
    traces) against reference-equivalent replay chains (numpy RandomState, host conjugate block) on a
    two-community network: intercept, blending coefficient, number of occupied components and the
    co-clustering matrix agree within Monte-Carlo error (SURVEY 8d parity check iv)."""
    from dynetlsm_b200 import DynamicNetworkHDPLPCM
    from dynetlsm_b200.diagnostics import ess
    Y = _splitting_network(n=36, T=3, seed=7)
    kw = dict(n_iter=700, tune=300, burn=300, n_features=2, n_components=6)
    dev = DynamicNetworkHDPLPCM(random_state=5, sampler="device", n_chains=6, **kw).fit(Y)
    reps = [DynamicNetworkHDPLPCM(random_state=s, sampler="replay", **kw).fit(Y) for s in (11, 12)]
    nb = 600
    a = dev.chains_["intercepts"][:, nb:, 0]
    b = np.stack([r.intercepts_[nb:, 0] for r in reps])
    se = np.sqrt(a.var() / max(ess(a), 10) + b.var() / max(ess(b), 10))
    assert abs(a.mean() - b.mean()) < 5 * se + 0.03, (a.mean(), b.mean(), se)
    la, lb = dev.chains_["lambdas"][:, nb:], np.stack([r.lambdas_[nb:, 0] for r in reps])
    assert abs(la.mean() - lb.mean()) < 0.04, (la.mean(), lb.mean())
    # occupied components: same ballpark (the HDP prior makes this a broad posterior)
    ka = dev.chains_["n_clusters"][:, nb:].mean()
    kb = np.mean([[np.unique(zz).size for zz in r.zs_[nb:]] for r in reps])
    assert abs(ka - kb) < 0.4, (ka, kb)
    # co-clustering probabilities pooled over chains (invariant to label switching)
    def cooc(zs):                                     # (S, T, n) -> (T, n, n)
        return (zs[:, :, :, None] == zs[:, :, None, :]).mean(axis=0)
    ca = np.mean([cooc(dev.chains_["zs"][c, nb:]) for c in range(6)], axis=0)
    cb = np.mean([cooc(r.zs_[nb:]) for r in reps], axis=0)
    assert np.abs(ca - cb).mean() < 0.04, np.abs(ca - cb).mean()
    truth = np.random.RandomState(7).randint(0, 2, 36)   # the generator's communities
    same = truth[:, None] == truth[None, :]
    assert ca[0][same].mean() > ca[0][~same].mean() + 0.2  # and both recover the communities


def test_sparse_network_fit_builds_the_case_control_lists_on_the_device():
    """fit() on a SparseNetwork (ties only, the way a network too large for a dense tensor arrives):
    the degree / edge lists mirrored under the reference's attribute names equal the host
    construction from the dense tensor (case_control_likelihood.py:37-73), the chain runs on the
    case-control kernels and gives a finite, improving log-posterior."""
    from dynetlsm_b200 import DynamicNetworkLSM
    from dynetlsm_b200.case_control_likelihood import DirectedCaseControlSampler, SparseNetwork
    rng = np.random.RandomState(4)
    T, n = 3, 60
    Xtrue = np.cumsum(rng.randn(T, n, 2) * np.array([1.0, 0.05, 0.05])[:, None, None], axis=0) / n
    dist = np.sqrt(((Xtrue[:, :, None] - Xtrue[:, None]) ** 2).sum(-1))
    radii = rng.dirichlet(np.ones(n) * 10)
    eta = 0.5 * (1 - dist / radii[None, None, :]) + 0.5 * (1 - dist / radii[None, :, None])
    Y = (rng.rand(T, n, n) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    for t in range(T):
        np.fill_diagonal(Y[t], 0)
    net = SparseNetwork.from_dense(Y)
    assert np.array_equal(net.toarray(), Y)
    m = DynamicNetworkLSM(n_iter=60, tune=40, burn=40, is_directed=True, n_control=10, n_resample_control=20,
                          step_size_X=0.05 / n, sigma_sq=1e-4, tau_sq="auto", random_state=1, n_chains=2)
    m.fit(net, X_init=Xtrue + 0.1 / n * rng.randn(T, n, 2), radii_init=radii, intercept_init=[0.4, 0.6])
    host = DirectedCaseControlSampler(n_control=10, random_state=0).init(Y, sample=False)
    cc = m.case_control_sampler_
    assert np.array_equal(cc.degrees_, host.degrees_)
    assert np.array_equal(cc.in_edges_, host.in_edges_) and np.array_equal(cc.out_edges_, host.out_edges_)
    assert m.Xs_.shape == (140, T, n, 2) and np.all(np.isfinite(m.logps_))
    assert cc.control_nodes_in_.shape == (T, n, 10)
    assert m.logps_[70:].mean() > m.logps_[:10].mean() - 0.1 * abs(m.logps_[:10].mean())   # (a noisy estimator: controls are redrawn)
    assert not np.array_equal(m.Xs_[-1], m.Xs_[0])
    # a sequence of scipy.sparse matrices is accepted as well; without starting values the chain still runs
    import scipy.sparse as sp
    m2 = DynamicNetworkLSM(n_iter=10, tune=10, burn=10, is_directed=True, n_control=10, step_size_X=0.05 / n,
                           sigma_sq=1e-4, tau_sq="auto", random_state=2)
    m2.fit([sp.csr_matrix(Y[t]) for t in range(T)])
    assert np.all(np.isfinite(m2.logps_)) and np.array_equal(m2.case_control_sampler_.degrees_, host.degrees_)
    with pytest.raises(ValueError):
        DynamicNetworkLSM(n_iter=5, is_directed=True).fit(net)      # no case-control likelihood


@pytest.mark.parametrize("selection,thin", [("vi", None), ("bic", None), ("map", None), ("vi", 4)])
def test_hdp_selection_types_and_thinning(selection, thin):
    """The reference's point-estimate selection on the device traces (hdp_lpcm.py:1089-1170):
    'vi' = minimum posterior-expected VI over the co-clustering probabilities, 'bic' / 'map' through
    the approximate BIC table; burn-in counted in STORED samples when the trace is thinned
    (hdp_lpcm.py:458-465)."""
    from dynetlsm_b200 import DynamicNetworkHDPLPCM
    from dynetlsm_b200 import model_selection as MS
    Y = _splitting_network(n=30, T=3, seed=3)
    m = DynamicNetworkHDPLPCM(n_iter=120, tune=60, burn=60, n_components=5, selection_type=selection, thin=thin,
                              random_state=3).fit(Y)
    S = 240 if thin is None else 60
    nb = 120 if thin is None else 30
    assert m.zs_.shape[0] == S and m.n_burn_ == nb
    assert m.bic_.shape[1] == 4 and len(m.models_) == m.bic_.shape[0] and m.counts_.shape == (S - nb,)
    assert set(m.bic_[:, 0].astype(int)) == set(np.unique(m.counts_))
    assert len(m.posterior_group_ids_) == 3 and all(c.sum() == S - nb for c in m.posterior_group_counts_)
    assert m.cooccurrence_probas_.shape == (3, 30, 30) and np.allclose(np.diagonal(m.cooccurrence_probas_, axis1=1, axis2=2), 1)
    if selection == "vi":
        vis = MS.expected_vi_trace(m.zs_[nb:], m.cooccurrence_probas_, 5)
        assert m.selected_id_ >= nb and vis[m.selected_id_ - nb] == vis.min()
    else:
        k = m.best_k_
        row = m.bic_[m.bic_[:, 0] == k][0]
        assert m.selected_id_ == int(row[3])
        if selection == "bic":
            assert row[1] == m.bic_[:, 1].min()
        else:
            assert k == np.argmax(np.bincount(m.counts_))
    assert m.z_.max() + 1 == m.mu_.shape[0] == m.sigma_.shape[0] == m.init_weights_.shape[0]
    assert np.allclose(m.init_weights_.sum(), 1) and np.allclose(m.trans_weights_[1:].sum(axis=2), 1)
    z, pval = m.logp_geweke_
    assert np.isfinite(z) and 0 <= pval <= 1
    with pytest.raises(ValueError):
        DynamicNetworkHDPLPCM(n_iter=5, selection_type="aic").fit(Y)


@pytest.mark.parametrize("sampler", ["device", "replay"])
def test_hdp_case_control_likelihood_runs_in_both_modes(sampler):
    """Directed HDP-LPCM on the case-control likelihood (hdp_lpcm.py:722-733, 826-829), control sets
    redrawn every 8 sweeps: the chain runs, stays finite and never reads past a control list."""
    from dynetlsm_b200 import DynamicNetworkHDPLPCM
    g = load_golden("lsm_casecontrol_monks.npz")
    Y = g["Y"].astype(np.float64)
    m = DynamicNetworkHDPLPCM(n_iter=20, tune=20, burn=10, tune_interval=6, n_components=4, is_directed=True,
                              n_control=5, n_resample_control=8, random_state=11, sampler=sampler).fit(Y)
    assert m.Xs_.shape == (50, 3, 18, 2) and np.isfinite(m.logps_).all()
    assert m.sampler_counters_["ub_flags"] == 0
