"""The device-resident remainder of the estimator loops: joint log-posterior
(lsm.py:576-625, hdp_lpcm.py:1188-1280), in-loop Procrustes (lsm.py:495-498, procrustes.py:20-35) and
the trace pipeline of dlsm_run_traced, each against its host statement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _L():
    from dynetlsm_b200 import _lib
    return _lib


def _net(rng, T, n, directed, density=0.2):
    Y = (rng.rand(T, n, n) < density).astype(np.float64)
    for t in range(T):
        np.fill_diagonal(Y[t], 0)
    if not directed:
        Y = np.triu(Y, 1)
        Y = Y + Y.transpose(0, 2, 1)
    return Y


def _lsm_engine(T, n, d, C_, directed, seed=0, chunk=None):
    L = _L()
    rng = np.random.RandomState(seed)
    Y = _net(rng, T, n, directed)
    e = L.Engine(T=T, n=n, d=d, n_chains=C_, is_directed=directed, tune=50, tune_interval=10)
    e.set_network(Y)
    X = rng.randn(C_, T, n, d) * (0.02 if directed else 1.0)
    e.set(L.F_X, X)
    ic = np.zeros((C_, 2)); ic[:, 0] = 0.4 + 0.1 * rng.rand(C_); ic[:, 1] = 0.7 if directed else 0.0
    e.set(L.F_INTERCEPT, ic)
    if directed:
        e.set(L.F_RADII, rng.dirichlet(np.ones(n) * 5, size=C_))
    e.set_hyper(tau_sq=1.7, sigma_sq=0.3 if not directed else 0.001, intercept_prior=(0.3, 0.6),
                intercept_variance_prior=2.5)
    e.set_tuner(0.1 if not directed else 0.005)
    e.set_rng(seed + 11)
    return e


def _hdp_engine(T, n, d, K, C_, directed, seed=0):
    L = _L()
    from dynetlsm_b200.hdp_updates import HDPHyper
    rng = np.random.RandomState(seed)
    Y = _net(rng, T, n, directed)
    e = L.Engine(T=T, n=n, d=d, n_chains=C_, K=K, mixture=True, is_directed=directed, tune=50,
                 tune_interval=10)
    e.set_network(Y)
    e.set(L.F_X, rng.randn(C_, T, n, d) * (0.02 if directed else 1.0))
    ic = np.zeros((C_, 2)); ic[:, 0] = 0.5; ic[:, 1] = 0.6 if directed else 0.0
    e.set(L.F_INTERCEPT, ic)
    if directed:
        e.set(L.F_RADII, rng.dirichlet(np.ones(n) * 5, size=C_))
    e.set(L.F_MU, rng.randn(C_, K, d)); e.set(L.F_SIGMA, rng.gamma(3, 0.3, (C_, K)))
    e.set(L.F_LAMBDA, rng.uniform(0.3, 0.95, C_))
    e.set(L.F_WEIGHTS, rng.dirichlet(np.ones(K), size=(C_, T, K)))
    e.set(L.F_BETA, rng.dirichlet(np.ones(K) * 2, size=C_))
    e.set(L.F_Z, rng.randint(0, K, (C_, T, n)))
    hy = np.zeros((C_, 8))
    hy[:, :6] = np.c_[rng.gamma(2, 1, C_), rng.gamma(2, 1, C_), rng.gamma(2, 1, C_), rng.gamma(2, 2, C_),
                      rng.gamma(3, 1, C_), rng.gamma(2, 1, C_)]
    e.set(L.F_HYPER, hy)
    hp = HDPHyper(gamma=1.0, alpha_init=1.0, alpha=1.0, kappa=1.0, mean_variance_prior=2.0, b=1.0, a=2.0,
                  a0=0.1, b0=0.1, c0=3.0, d0=0.5, lambda_prior=0.9, lambda_variance_prior=0.01,
                  gamma_prior_shape=1.0, gamma_prior_rate=0.1, alpha_init_shape=1.0, alpha_init_rate=0.1,
                  alpha_kappa_shape=1.0, alpha_kappa_rate=0.1, resample_mean_variance=True,
                  resample_b=True)
    e.set_hdp_prior(hp.a, hp.a0, hp.b0, hp.c0, hp.d0, hp.lambda_prior, hp.lambda_variance_prior,
                    hp.gamma_prior_shape, hp.gamma_prior_rate, hp.alpha_init_shape, hp.alpha_init_rate,
                    hp.alpha_kappa_shape, hp.alpha_kappa_rate, hp.resample_mean_variance, hp.resample_b)
    e.set_hyper(intercept_prior=(0.4, 0.5), intercept_variance_prior=2.0)
    e.set_tuner(0.1 if not directed else 0.005)
    e.set_rng(seed + 5)
    return e, hp


def _host_logp_lsm(e, tau_sq, sigma_sq, prior, var):
    L = _L()
    X, ic = e.get(L.F_X), e.get(L.F_INTERCEPT)
    out = e.loglik_full()
    for c in range(e.C):
        lp = -np.sum(0.5 * np.sum(X[c, 0] ** 2, axis=1) / tau_sq)
        for t in range(1, e.T):
            lp -= np.sum(0.5 * np.sum((X[c, t] - X[c, t - 1]) ** 2, axis=1) / sigma_sq)
        for i in range(e.m):
            lp -= 0.5 * (ic[c, i] - prior[i]) ** 2 / var
        out[c] += lp
    return out


def _host_logp_hdp(e, hp):
    import copy
    L = _L()
    from dynetlsm_b200.hdp_updates import hdp_log_prior
    X, ic, z = e.get(L.F_X), e.get(L.F_INTERCEPT), e.get(L.F_Z).astype(np.int64)
    mu, sg, lam = e.get(L.F_MU), e.get(L.F_SIGMA), e.get(L.F_LAMBDA)
    w, be, hy = e.get(L.F_WEIGHTS), e.get(L.F_BETA), e.get(L.F_HYPER)
    ra = e.get(L.F_RADII) if e.is_directed else None
    out = e.loglik_full()
    for c in range(e.C):
        h = copy.copy(hp)
        h.gamma, h.alpha_init, h.alpha, h.kappa, h.mean_variance_prior, h.b = hy[c, :6]
        lp = hdp_log_prior(h, e.K, X[c], ic[c, :e.m], np.array([0.4, 0.5])[:e.m], 2.0, mu[c], sg[c], z[c],
                           w[c], be[c], np.array([lam[c]]), radii=None if ra is None else ra[c])
        out[c] += float(np.ravel(lp)[0])
    return out


@pytest.mark.parametrize("directed", [False, True])
def test_logp_lsm(directed):
    e = _lsm_engine(5, 40, 2, 3, directed)
    want = _host_logp_lsm(e, 1.7, 0.3 if not directed else 0.001, (0.3, 0.6), 2.5)
    assert np.allclose(e.logp(), want, rtol=1e-11, atol=0)


@pytest.mark.parametrize("directed,K,d", [(False, 6, 2), (True, 4, 2), (False, 9, 3)])
def test_logp_hdp(directed, K, d):
    e, hp = _hdp_engine(6, 35, d, K, 3, directed)
    want = _host_logp_hdp(e, hp)
    got = e.logp()
    assert np.all(np.isfinite(got))
    assert np.allclose(got, want, rtol=1e-10, atol=0)


def test_logp_hdp_with_zero_weights():
    """Dirichlet draws that underflow to exactly 0 are clipped like the reference does
    (distributions.py:72-102)."""
    L = _L()
    e, hp = _hdp_engine(4, 20, 2, 5, 2, False)
    w = e.get(L.F_WEIGHTS)
    w[:, 2, 1, 0] += w[:, 2, 1, 3]
    w[:, 2, 1, 3] = 0.0
    e.set(L.F_WEIGHTS, w)
    z = e.get(L.F_Z)
    z[(z == 3)] = 2                       # no transition uses the zero entry
    e.set(L.F_Z, z)
    assert np.allclose(e.logp(), _host_logp_hdp(e, hp), rtol=1e-10, atol=0)


@pytest.mark.parametrize("d", [1, 2, 3, 5, 8])
def test_procrustes_matches_scipy(d):
    from dynetlsm_b200.host_init import longitudinal_procrustes_rotation
    from scipy.stats import ortho_group
    L = _L()
    T, n, C_ = 4, 30, 3
    rng = np.random.RandomState(d)
    e = L.Engine(T=T, n=n, d=d, n_chains=C_)
    ref = rng.randn(C_, T, n, d)
    X = np.empty_like(ref)
    for c in range(C_):
        Q = ortho_group.rvs(d, random_state=rng) if d > 1 else np.array([[-1.0]])
        X[c] = ref[c] @ Q + 0.05 * rng.randn(T, n, d)
    e.set(L.F_X, X)
    e.set_procrustes_ref(ref)
    e.procrustes()
    got = e.get(L.F_X)
    for c in range(C_):
        want, _ = longitudinal_procrustes_rotation(ref[c], X[c])
        assert np.allclose(got[c], want, rtol=0, atol=1e-12)
    e.set_procrustes_ref(None)
    with pytest.raises(L.DlsmError):
        e.procrustes()


@pytest.mark.parametrize("model", ["lsm", "lsm-directed", "hdp"])
@pytest.mark.parametrize("thin,pinned", [(1, False), (3, True)])
def test_run_traced_equals_stepwise(model, thin, pinned, monkeypatch):
    """Records of dlsm_run_traced == the states a sweep-by-sweep loop reads back, across several
    ring chunks (the chunk size is forced down to a few records)."""
    L = _L()
    monkeypatch.setenv("DLSM_TRACE_CHUNK_BYTES", "60000")

    def make():
        if model == "hdp":
            return _hdp_engine(5, 30, 2, 6, 3, False)[0]
        return _lsm_engine(5, 30, 2, 3, model == "lsm-directed")
    a, b = make(), make()
    n_sweeps = 23
    first = (L.F_X,) + ((L.F_MU, L.F_WEIGHTS) if model == "hdp" else ())
    every = (L.F_INTERCEPT,) + ((L.F_Z, L.F_LAMBDA, L.F_HYPER) if model == "hdp" else ()) + \
            ((L.F_RADII,) if model == "lsm-directed" else ())
    tr = a.run_traced(n_sweeps, fields_all=every, fields_first=first, thin=thin, pinned=pinned)
    rec = n_sweeps // thin
    assert tr["logp"].shape == (rec, 3)
    for r in range(rec):
        b.run_sweeps(thin)
        for f in every:
            assert np.array_equal(tr[f][r], b.get(f)), (f, r)
        for f in first:
            assert np.array_equal(tr[f][r, 0], b.get(f)[0]), (f, r)
        assert np.allclose(tr["logp"][r], b.logp(), rtol=1e-10, atol=0)
    b.run_sweeps(n_sweeps - rec * thin)
    assert np.array_equal(a.get(L.F_X), b.get(L.F_X))      # the tail after the last record ran too


def test_run_traced_with_procrustes_reference():
    L = _L()
    a, b = _lsm_engine(4, 25, 2, 2, False), _lsm_engine(4, 25, 2, 2, False)
    ref = a.get(L.F_X)
    a.set_procrustes_ref(ref); b.set_procrustes_ref(ref)
    tr = a.run_traced(6, fields_all=(L.F_X,))
    for r in range(6):
        b.sweep_latent()
        b.procrustes()
        b.center()
        b.sample_intercepts()
        assert np.array_equal(tr[L.F_X][r], b.get(L.F_X))
    assert np.allclose(tr[L.F_X].mean(axis=(2, 3)), 0, atol=1e-12)


def test_run_traced_argument_errors():
    L = _L()
    e = _lsm_engine(3, 12, 2, 1, False)
    with pytest.raises(L.DlsmError):
        e.run_traced(4, fields_all=(L.F_MU,))                 # no mixture fields in an LSM engine
    with pytest.raises(L.DlsmError):
        e._ck(e.L.dlsm_run_traced(e.h, 4, 0, L.TraceSpec(thin=0), None, None))
    out = e.run_traced(0, fields_all=(L.F_X,))
    assert out[L.F_X].shape[0] == 0


def _edge_engine(T, n, C_, density, seed):
    L = _L()
    from dynetlsm_b200.case_control_likelihood import DirectedCaseControlSampler
    rng = np.random.RandomState(seed)
    Y = _net(rng, T, n, True, density)
    cc = DirectedCaseControlSampler(n_control=5, n_resample=10, random_state=rng).init(Y, sample=False)
    e = L.Engine(T=T, n=n, d=2, n_chains=C_, is_directed=True, case_control=True)
    e.set_edge_lists(cc.degrees_, cc.in_edges_, cc.out_edges_)
    return e, Y


@pytest.mark.parametrize("n,m,density,per_chain", [(30, 7, 0.15, False), (12, 20, 0.3, True), (64, 33, 0.05, True)])
def test_device_control_sets_are_valid(n, m, density, per_chain):
    """dlsm_resample_controls (case_control_likelihood.py:75-112 on the device RNG): the right number
    of distinct non-neighbours per node and direction, -1 padded, never the node itself."""
    T, C_ = 3, 3
    e, Y = _edge_engine(T, n, C_, density, seed=n)
    e.set_rng(123)
    e.resample_controls(m, per_chain=per_chain)
    ci, co = e.get_controls()
    assert ci.shape == ((C_ if per_chain else 1), T, n, m)
    for s in range(ci.shape[0]):
        for t in range(T):
            for i in range(n):
                for arr, nbrs in ((ci, np.nonzero(Y[t][:, i])[0]), (co, np.nonzero(Y[t][i])[0])):
                    row = arr[s, t, i]
                    want = min(n - nbrs.size - 1, m)
                    got = row[:want]
                    assert np.all(got >= 0) and np.all(row[want:] == -1)
                    assert np.unique(got).size == want
                    assert i not in got and not np.intersect1d(got, nbrs).size
    first = ci.copy()
    e.resample_controls(m, per_chain=per_chain)
    assert not np.array_equal(first, e.get_controls()[0])          # a new draw
    e2, _ = _edge_engine(T, n, C_, density, seed=n)
    e2.set_rng(123)
    e2.resample_controls(m, per_chain=per_chain)
    assert np.array_equal(first, e2.get_controls()[0])             # same seed, same sets
    if per_chain:
        assert not np.array_equal(first[0], first[1])


def test_device_control_sets_are_uniform():
    """Every eligible node is equally likely to be a control (and equally likely in every slot)."""
    T, n, m, C_ = 1, 10, 3, 4096
    e, Y = _edge_engine(T, n, C_, 0.2, seed=1)
    e.set_rng(9)
    e.resample_controls(m, per_chain=True)
    ci, co = e.get_controls()
    for arr, nb_of in ((co, lambda i: np.nonzero(Y[0][i])[0]), (ci, lambda i: np.nonzero(Y[0][:, i])[0])):
        for i in (0, 4, 9):
            elig = np.setdiff1d(np.arange(n), np.r_[nb_of(i), i])
            want = min(elig.size, m)
            for slot in range(want):
                cnt = np.bincount(arr[:, 0, i, slot], minlength=n)[elig]
                exp = C_ / elig.size
                chi2 = np.sum((cnt - exp) ** 2 / exp)
                assert chi2 < 40.0, (i, slot, cnt)              # dof <= 8: p ~ 1e-6


def test_device_mode_case_control_fit_runs():
    from dynetlsm_b200 import DynamicNetworkLSM
    rng = np.random.RandomState(0)
    Y = _net(rng, 3, 25, True, 0.2)
    m = DynamicNetworkLSM(is_directed=True, n_iter=40, tune=20, burn=20, n_control=6, n_resample_control=15,
                          random_state=1, n_chains=2).fit(Y)
    assert m.Xs_.shape == (80, 3, 25, 2) and np.all(np.isfinite(m.logps_))
    cc = m.case_control_sampler_
    assert cc.control_nodes_in_.shape == (3, 25, 6) and cc.control_nodes_out_.max() < 25


@pytest.mark.parametrize("pinned", [True, False])
def test_run_traced_large_positions_leave_early(pinned):
    """Position records of at least 1 MB into page-locked memory are copied out right after the
    centring, concurrently with the rest of the sweep (pageable destinations go through the ring):
    same records either way, and the same as a sweep-by-sweep loop."""
    L = _L()
    T, n, d, C_ = 4, 1100, 2, 16                    # X = 1.1 MB per record
    rng = np.random.RandomState(4)
    Y = _net(rng, T, n, False, 0.02)

    def make():
        e = L.Engine(T=T, n=n, d=d, n_chains=C_, tune=50, tune_interval=10)
        e.set_network(Y)
        r2 = np.random.RandomState(5)
        e.set(L.F_X, r2.randn(C_, T, n, d))
        e.set(L.F_INTERCEPT, np.tile([[0.5, 0.0]], (C_, 1)))
        e.set_hyper(tau_sq=2.0, sigma_sq=0.2)
        e.set_tuner(0.05)
        e.set_rng(8)
        return e
    a, b = make(), make()
    tr = a.run_traced(7, fields_all=(L.F_X, L.F_INTERCEPT), thin=2, pinned=pinned)
    assert tr[L.F_X].shape == (3, C_, T, n, d)
    for r in range(3):
        b.run_sweeps(2)
        assert np.array_equal(tr[L.F_X][r], b.get(L.F_X)), r
        assert np.array_equal(tr[L.F_INTERCEPT][r], b.get(L.F_INTERCEPT)), r
        assert np.allclose(tr["logp"][r], b.logp(), rtol=1e-10, atol=0)
    b.run_sweeps(1)
    assert np.array_equal(a.get(L.F_X), b.get(L.F_X))
    # positions only, no log-posterior: nothing goes through the ring on the early path
    tr = a.run_traced(3, fields_all=(L.F_X,), logp=False, pinned=pinned)
    b.run_sweeps(3)
    assert np.array_equal(tr[L.F_X][2], b.get(L.F_X))


@pytest.mark.parametrize("directed,d", [(False, 2), (True, 2), (True, 3)])
def test_edge_probas_vs_reference_kernel(directed, d):
    """dlsm_edge_probas against the reference's compiled directed_network_probas (K9,
    directed_likelihoods_fast.pyx:273-294) / expit(beta - dist) for the undirected model."""
    import ref_driver as R
    from scipy.special import expit
    L = _L()
    T, n, C_ = 3, 37, 2
    rng = np.random.RandomState(d)
    X = rng.randn(C_, T, n, d) * (0.02 if directed else 1.0)
    e = L.Engine(T=T, n=n, d=d, n_chains=C_, is_directed=directed)
    e.set(L.F_X, X)
    ic = np.array([[0.4, 0.7], [0.1, 0.9]]) if directed else np.array([[0.8, 0.0], [1.3, 0.0]])
    e.set(L.F_INTERCEPT, ic)
    radii = rng.dirichlet(np.ones(n) * 5, size=C_)
    if directed:
        e.set(L.F_RADII, radii)
    for c in range(C_):
        got = e.edge_probas(c)
        dist = np.linalg.norm(X[c][:, :, None, :] - X[c][:, None, :, :], axis=-1)
        if directed:
            if not R.have_ref():
                pytest.skip("oracle/_ref not built")
            want = R.ref_kernels()["directed_likelihoods_fast"].directed_network_probas(
                np.ascontiguousarray(dist), radii[c], ic[c, 0], ic[c, 1])
        else:
            want = expit(ic[c, 0] - dist)
            idx = np.arange(n)
            want[:, idx, idx] = 0.0
        assert np.allclose(got, want, rtol=1e-12, atol=1e-15)
        assert np.all(np.diagonal(got, axis1=1, axis2=2) == 0.0)


def test_device_cooccurrence_equals_host_counts():
    """Co-clustering probabilities accumulated on the device while sampling (dlsm_run_traced
    cooc_mode; label_utils.py:40-62) equal the host computation from the stored label draws."""
    from dynetlsm_b200 import DynamicNetworkHDPLPCM
    rng = np.random.RandomState(0)
    Y = _net(rng, 3, 28, False, 0.25)
    m = DynamicNetworkHDPLPCM(n_iter=60, tune=30, burn=30, n_components=5, random_state=3, n_chains=2).fit(Y)
    nb = m.n_burn_
    zs = m.zs_[nb:]
    want = (zs[:, :, :, None] == zs[:, :, None, :]).mean(axis=0)
    assert m.cooccurrence_probas_.shape == (3, 28, 28)
    assert np.allclose(m.cooccurrence_probas_, want, rtol=0, atol=1e-15)
    # the engine-level accumulator, pooled over chains
    L = _L()
    e, _ = _hdp_engine(4, 20, 2, 5, 3, False)
    tr = e.run_traced(9, fields_all=(L.F_Z,), cooc=2, cooc_from=4)
    counts, ns = e.cooccurrence(reset=True)
    z = tr[L.F_Z][4:]                                     # (5 records, 3 chains, T, n)
    assert ns == 15
    assert np.array_equal(counts, (z[:, :, :, :, None] == z[:, :, :, None, :]).sum(axis=(0, 1)).astype(np.uint32))
    assert e.cooccurrence()[1] == 0


@pytest.mark.parametrize("name,directed", [("lsm_undirected_monks.npz", False), ("lsm_directed_monks.npz", True)])
def test_logp_against_the_reference_trace(name, directed):
    """k_logp pinned DIRECTLY on the reference: the stored samples of a reference fit (positions,
    intercepts, radii; oracle/make_golden.py) are put on the device and dlsm_logp must return the
    reference's own logps_ (lsm.py:576-625)."""
    import os
    L = _L()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    Y = g["Y"].astype(np.float64)
    T, n, _ = Y.shape
    Xs, ics, want = g["Xs"], g["intercepts"], g["logps"]
    S = min(len(want), 12)
    e = L.Engine(T=T, n=n, d=Xs.shape[-1], n_chains=S, is_directed=directed)
    e.set_network(Y)
    e.set(L.F_X, np.ascontiguousarray(Xs[:S]))
    ic = np.zeros((S, 2)); ic[:, :ics.shape[1]] = ics[:S]
    e.set(L.F_INTERCEPT, ic)
    if directed:
        e.set(L.F_RADII, np.ascontiguousarray(g["radiis"][:S]))
    e.set_hyper(tau_sq=float(g["tau_sq"]), sigma_sq=float(g["sigma_sq"]),
                intercept_prior=np.asarray(g["intercept_prior"], dtype=np.float64),
                intercept_variance_prior=float(g["intercept_variance_prior"]))
    assert np.allclose(e.logp(), want[:S], rtol=1e-10, atol=0)


@pytest.mark.parametrize("name,directed", [("hdp_undirected_split.npz", False), ("hdp_directed_monks.npz", True)])
def test_logp_hdp_against_the_reference_trace(name, directed):
    """k_logp's mixture branch pinned DIRECTLY on the reference: stored samples of a reference
    HDP-LPCM fit (state + the hyper-parameters in force, oracle/make_golden.py) are put on the device,
    one per chain, and dlsm_logp must return the reference's own logps_ (hdp_lpcm.py:1188-1280)."""
    import os
    L = _L()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    Y = g["Y"].astype(np.float64)
    T, n, _ = Y.shape
    S = min(g["logp"].shape[0], 12)
    K, d = g["mu_next"].shape[1:]
    e = L.Engine(T=T, n=n, d=d, n_chains=S, K=K, mixture=True, is_directed=directed)
    e.set_network(Y)
    e.set(L.F_X, np.ascontiguousarray(g["X_centered"][:S]))
    ic = np.zeros((S, 2)); ic[:, :g["intercept_out"].shape[1]] = g["intercept_out"][:S]
    e.set(L.F_INTERCEPT, ic)
    if directed:
        e.set(L.F_RADII, np.ascontiguousarray(g["radii_logp"][:S]))
    e.set(L.F_MU, np.ascontiguousarray(g["mu_next"][:S])); e.set(L.F_SIGMA, np.ascontiguousarray(g["sigma_next"][:S]))
    e.set(L.F_LAMBDA, np.ascontiguousarray(g["lmbda_next"][:S].reshape(S)))
    e.set(L.F_WEIGHTS, np.ascontiguousarray(g["w_next"][:S])); e.set(L.F_BETA, np.ascontiguousarray(g["beta_next"][:S]))
    e.set(L.F_Z, np.ascontiguousarray(g["z_out"][:S]))
    hy = np.zeros((S, 8)); hy[:, :6] = g["hyper_next"][:S]
    e.set(L.F_HYPER, hy)
    a, a0, b0, c0, d0, lam0, lamv = [float(v) for v in g["hdp_prior"]]
    e.set_hdp_prior(a, a0, b0, c0, d0, lam0, lamv, 1.0, 0.1, 1.0, 1.0, 5, 0.1, True, True)
    e.set_hyper(intercept_prior=np.asarray(g["intercept_prior"], dtype=np.float64),
                intercept_variance_prior=float(g["intercept_variance_prior"]))
    got = e.logp()
    e.close()
    assert np.allclose(got, g["logp"][:S], rtol=1e-10, atol=0), (got, g["logp"][:S])
