"""``DynamicNetworkLPCM`` (dynetlsm_b200/lpcm.py) on the GPU: whole ``fit`` in replay mode against the
chains the unmodified reference produced (tests/golden/lpcm_*.npz, oracle/make_golden_lpcm.py) --
999-sweep LSM initialisation, k-means, then every sweep of lpcm.py:514-709 with the hot path on the
device (mixture-prior latent sweep, centring, intercept / radii MH, label FFBS through the
stacked-weights layout) and the Dirichlet / conjugate block on the host.  Also the reference's own
smoke test restated (dynetlsm/tests/test_lpcm.py: fit and assert shapes) in device-RNG mode.
"""
import warnings

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")


def _pairwise(X):
    return np.linalg.norm(X[:, :, None] - X[:, None], axis=-1)


def test_lpcm_replay_reproduces_the_reference_fit_undirected():
    from dynetlsm_b200 import DynamicNetworkLPCM
    g = load_golden("lpcm_undirected_split.npz")
    Y = g["Y"].astype(np.float64)
    m = DynamicNetworkLPCM(n_iter=25, tune=25, burn=10, tune_interval=8, n_components=4,
                           random_state=3, sampler="replay").fit(Y)
    S = g["z_out"].shape[0]
    assert m.zs_.shape[0] == S + 1
    assert np.array_equal(m.zs_[1:], g["z_out"])                           # every label draw
    assert np.array_equal(m.intercepts_[1:], g["intercept_out"])           # every intercept draw
    assert np.array_equal(m.init_weights_[1:], g["init_next"])             # host Dirichlet block
    assert np.array_equal(m.trans_weights_[1:], g["trans_next"])
    assert np.array_equal(m.lambdas_[1:], g["lmbda_next"])
    assert np.array_equal(m.sigmas_[1:], g["sigma_next"])
    assert np.allclose(m.logps_, g["logps"], rtol=1e-9, atol=0)
    # positions and means are rotated post hoc (lpcm.py:739-745): compare rotation invariants
    assert np.allclose(_pairwise(m.Xs_[S]), _pairwise(g["X_centered"][-1]), rtol=0, atol=1e-10)
    assert np.allclose(np.linalg.norm(m.mus_[1:], axis=-1), np.linalg.norm(g["mu_next"], axis=-1),
                       rtol=1e-10, atol=1e-12)
    # point estimate: 'map' with the reference's un-offset argmax (lpcm.py:716), co-clustering
    assert m.selected_id_ == int(g["selected_id"])
    assert np.array_equal(m.z_, g["z_hat"])
    assert np.allclose(m.cooccurrence_probas_, g["cooccurrence_probas"], rtol=0, atol=1e-12)
    assert m.mean_variance_prior_ == g["mean_variance_prior"][-1] and m.b_ == g["b"][-1]


def test_lpcm_replay_directed_monks_vi_selection():
    """Directed model with radii, point estimate by posterior-expected VI.  Directed initial values
    are not bit-comparable with the reference (its intercept MLE reads an uninitialised
    accumulator, SURVEY.md K10 / host_init.directed_intercept_mle), so the chain is checked for
    validity and self-consistency; the sweeps themselves are pinned by the teacher-forced replays
    of tests/test_gpu_parity.py and the host block by tests/test_lpcm_host.py."""
    from dynetlsm_b200 import DynamicNetworkLPCM
    g = load_golden("lpcm_directed_monks.npz")
    Y = g["Y"].astype(np.float64)
    m = DynamicNetworkLPCM(n_iter=15, tune=15, burn=10, tune_interval=5, n_components=3,
                           is_directed=True, selection_type="vi", random_state=5,
                           sampler="replay").fit(Y)
    S, T, n = g["z_out"].shape
    assert m.zs_.shape == (S + 1, T, n) and m.radiis_.shape == (S + 1, n)
    assert np.allclose(m.radiis_.sum(axis=1), 1.0)
    assert np.isfinite(m.logps_).all()
    assert np.allclose(m.init_weights_[1:].sum(axis=1), 1.0)
    assert np.allclose(m.trans_weights_[1:].sum(axis=2), 1.0)
    assert m.n_burn_ <= m.selected_id_ <= S                               # VI scans post-burn-in samples
    assert np.allclose(np.diagonal(m.cooccurrence_probas_, axis1=1, axis2=2), 1.0)
    assert m.z_.shape == (T, n) and m.radii_.shape == (n,)
    # the public logp() of the selected sample agrees with the trace (positions were rotated after
    # sampling, the joint density is invariant to it; mu_ was rotated with them)
    lp = m.logp(m.X_, m.intercept_, m.mu_, m.sigma_, m.z_, m.init_weight_, m.trans_weight_, m.lambda_,
                radii=m.radii_)
    assert np.isfinite(lp)


def test_reference_smoke_test_lpcm_shapes_device_rng():
    from dynetlsm_b200 import DynamicNetworkLPCM
    from test_gpu_estimators import _splitting_network
    Y = _splitting_network()
    m = DynamicNetworkLPCM(n_iter=100, burn=100, tune=100, n_features=2, n_components=3,
                           random_state=123).fit(Y)
    assert m.X_.shape == (2, 50, 2) and m.z_.shape == (2, 50)
    assert m.zs_.shape == (300, 2, 50) and m.trans_weights_.shape == (300, 3, 3)
    assert m.init_weights_.shape == (300, 3) and m.mus_.shape == (300, 3, 2)
    assert np.isfinite(m.logps_).all() and 0.5 < m.auc_ <= 1.0
    assert np.allclose(np.diagonal(m.cooccurrence_probas_, axis1=1, axis2=2), 1.0)
    # two planted communities: the sampler separates them better than chance at the point estimate
    rng = np.random.RandomState(42)
    truth = rng.randint(0, 2, 50)
    same_truth = truth[:, None] == truth[None]
    same_hat = m.z_[0][:, None] == m.z_[0][None]
    assert (same_truth == same_hat).mean() > 0.6
    m.delete_traces()
    assert not hasattr(m, "Xs_")


def test_lpcm_thinning_counts_burn_in_in_stored_samples():
    from dynetlsm_b200 import DynamicNetworkLPCM
    from test_gpu_estimators import _splitting_network
    Y = _splitting_network(n=30)
    m = DynamicNetworkLPCM(n_iter=40, burn=20, tune=20, thin=4, n_components=2, random_state=1).fit(Y)
    assert m.n_burn_ == 10 and m.zs_.shape[0] == 20
    assert m.X_mean_.shape == (2, 30, 2)


def test_lpcm_device_rng_matches_replay_in_distribution():
    """Device-RNG fits (Philox hot path in ONE dlsm_run_sweeps call per sweep, label block on the side
    stream) against reference-equivalent replay fits on a two-community network: intercept, blending
    coefficient and the co-clustering matrix agree within Monte-Carlo error (SURVEY 8d parity iv)."""
    from dynetlsm_b200 import DynamicNetworkLPCM
    from dynetlsm_b200.diagnostics import ess
    from test_gpu_estimators import _splitting_network
    Y = _splitting_network(n=36, T=3, seed=7)
    kw = dict(n_iter=700, tune=300, burn=300, n_features=2, n_components=3)
    devs = [DynamicNetworkLPCM(random_state=s, sampler="device", **kw).fit(Y) for s in (5, 6, 7)]
    reps = [DynamicNetworkLPCM(random_state=s, sampler="replay", **kw).fit(Y) for s in (11, 12)]
    nb = 600
    a = np.stack([r.intercepts_[nb:, 0] for r in devs])
    b = np.stack([r.intercepts_[nb:, 0] for r in reps])
    se = np.sqrt(a.var() / max(ess(a), 10) + b.var() / max(ess(b), 10))
    print("lpcm device vs replay: intercept", a.mean(), b.mean(), "se", se)
    assert abs(a.mean() - b.mean()) < 5 * se + 0.05, (a.mean(), b.mean(), se)
    la = np.mean([r.lambdas_[nb:, 0].mean() for r in devs])
    lb = np.mean([r.lambdas_[nb:, 0].mean() for r in reps])
    print("lpcm device vs replay: lambda", la, lb)
    assert abs(la - lb) < 0.05, (la, lb)

    def cooc(zs):                                     # (S, T, n) -> (T, n, n), invariant to label switching
        return (zs[:, :, :, None] == zs[:, :, None, :]).mean(axis=0)
    ca = np.mean([cooc(r.zs_[nb:]) for r in devs], axis=0)
    cb = np.mean([cooc(r.zs_[nb:]) for r in reps], axis=0)
    print("lpcm device vs replay: co-clustering mean abs diff", np.abs(ca - cb).mean())
    assert np.abs(ca - cb).mean() < 0.08, np.abs(ca - cb).mean()
    truth = np.random.RandomState(7).randint(0, 2, 36)   # the generator's communities
    same = truth[:, None] == truth[None, :]
    assert ca[0][same].mean() > ca[0][~same].mean() + 0.2


@pytest.mark.parametrize("sampler", ["replay", "device"])
def test_lpcm_case_control_likelihood_runs_in_both_modes(sampler):
    """Directed finite mixture on the case-control likelihood (lpcm.py:404-415, 525-527): control
    sets redrawn every 8 sweeps (host RandomState in replay mode, device Philox otherwise)."""
    from dynetlsm_b200 import DynamicNetworkLPCM
    g = load_golden("lsm_casecontrol_monks.npz")
    Y = g["Y"].astype(np.float64)
    m = DynamicNetworkLPCM(n_iter=20, tune=20, burn=10, tune_interval=6, n_components=3, is_directed=True,
                           n_control=5, n_resample_control=8, random_state=11, sampler=sampler).fit(Y)
    T, n = Y.shape[:2]
    assert m.Xs_.shape == (50, T, n, 2) and m.radiis_.shape == (50, n)
    assert np.allclose(m.radiis_.sum(axis=1), 1.0) and np.isfinite(m.logps_).all()
    cc = m.case_control_sampler_
    assert np.array_equal(cc.in_edges_, g["cc_in_edges"])
    assert cc.control_nodes_in_.shape[:2] == (T, n) and cc.control_nodes_out_.shape[:2] == (T, n)
    assert m.sampler_counters_["ub_flags"] == 0
    assert not np.array_equal(m.zs_[-1], m.zs_[0]) or not np.array_equal(m.Xs_[-1], m.Xs_[0])
