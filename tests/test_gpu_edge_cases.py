"""Edge cases of the hot path against the CPU oracle: degenerate shapes, row lengths around the
32/64-lane chunking, every latent dimension up to 8 (numpy's 8-accumulator sum kicks in at d = 8),
empty / complete graphs, coincident positions (distance exactly 0), far-apart positions (softplus
tails), single-component mixtures, ragged case-control lists with sentinels."""
import numpy as np
import pytest

import pyoracle as O

pytestmark = pytest.mark.gpu


def _L():
    from dynetlsm_b200 import _lib
    return _lib


def _net(rng, T, n, directed, density):
    Y = (rng.rand(T, n, n) < density).astype(np.float64)
    for t in range(T):
        np.fill_diagonal(Y[t], 0)
    if not directed:
        Y = np.triu(Y, 1)
        Y = Y + Y.transpose(0, 2, 1)
    return Y


def _run_both(T, n, d, directed, Y, X, mode=None, monkeypatch=None, n_sweeps=2, mixture=None,
              tau=2.0, sig=0.1, step=0.1, seed=0):
    L = _L()
    if mode:
        monkeypatch.setenv("DLSM_SWEEP_MODE", mode)
    rng = np.random.RandomState(seed)
    radii = rng.dirichlet(np.ones(n) * 5) if directed else None
    ic = np.array([0.4, 0.7]) if directed else np.array([0.9])
    K = 0 if mixture is None else mixture["sigma"].size
    e = L.Engine(T=T, n=n, d=d, is_directed=directed, K=K, mixture=mixture is not None, tune=3,
                 tune_interval=1)
    e.set_network(Y)
    e.set_hyper(tau_sq=tau, sigma_sq=sig)
    icd = np.zeros((1, 2)); icd[0, :ic.size] = ic
    e.set(L.F_INTERCEPT, icd)
    if directed:
        e.set(L.F_RADII, radii[None])
    if mixture is not None:
        e.set(L.F_MU, mixture["mu"][None]); e.set(L.F_SIGMA, mixture["sigma"][None])
        e.set(L.F_LAMBDA, np.array([mixture["lmbda"]])); e.set(L.F_Z, mixture["z"][None])
    e.set(L.F_X, X[None])
    e.set_tuner(step)
    tun = O.TunerState((T, n), step, tune=3, tune_interval=1)
    Xo = X.copy()
    # per-node and full-network log-likelihoods at the start state
    part = e.loglik_partial()[0]
    for t in range(T):
        for i in range(0, n, max(1, n // 7)):
            ref = (O.directed_partial_loglikelihood(Y[t], Xo[t], radii, ic[0], ic[1], i) if directed
                   else O.partial_loglikelihood(Y[t], Xo[t], ic[0], i))
            assert abs(part[t, i] - ref) <= 1e-10 * max(abs(ref), 1e-300) + 1e-13, (t, i, part[t, i], ref)
    for s in range(n_sweeps):
        eps = rng.randn(T, n, d)
        logu = np.log(rng.rand(T, n))
        out = O.sweep_latent(Xo, ic, tun, eps, logu, Y=Y, radii=radii, is_directed=directed,
                             tau_sq=tau, sigma_sq=sig, mixture=mixture)
        acc, ratio = e.sweep_latent(eps[None], logu[None], want_stats=True)
        assert np.array_equal(acc[0], out["accepted"])
        assert np.array_equal(e.get(L.F_X)[0], Xo)
        O.center(Xo)
        e.center()
        assert np.array_equal(e.get(L.F_X)[0], Xo)
    assert np.array_equal(e.get(L.F_X_STEP)[0], tun.step)
    return e


@pytest.mark.parametrize("mode", ["chain", "chain-dense", "slice"])
@pytest.mark.parametrize("T,n,d", [(1, 2, 1), (1, 5, 2), (2, 31, 2), (2, 32, 2), (2, 33, 3), (3, 63, 2),
                                   (2, 64, 4), (2, 65, 5), (2, 97, 6), (2, 128, 7), (2, 129, 8), (40, 6, 2)])
def test_shapes_around_the_chunking(T, n, d, mode, monkeypatch):
    rng = np.random.RandomState(T * 100 + n)
    _run_both(T, n, d, False, _net(rng, T, n, False, 0.3), rng.randn(T, n, d), mode, monkeypatch)


@pytest.mark.parametrize("mode", ["chain", "slice"])
@pytest.mark.parametrize("T,n,d", [(1, 3, 2), (3, 40, 2), (2, 70, 3)])
def test_directed_shapes(T, n, d, mode, monkeypatch):
    rng = np.random.RandomState(T * 10 + n)
    _run_both(T, n, d, True, _net(rng, T, n, True, 0.2), rng.randn(T, n, d) / n, mode, monkeypatch,
              tau=1.0 / n ** 2, sig=1e-3 / n, step=0.02 / n)


@pytest.mark.parametrize("density", [0.0, 1.0])
def test_empty_and_complete_graphs(density):
    rng = np.random.RandomState(3)
    T, n, d = 2, 40, 2
    _run_both(T, n, d, False, _net(rng, T, n, False, density), rng.randn(T, n, d))
    _run_both(T, n, d, True, _net(rng, T, n, True, density), rng.randn(T, n, d) / n, tau=1e-3, sig=1e-4,
              step=0.01 / n)


def test_coincident_and_far_apart_positions():
    """dist == 0 exactly (several nodes on one point) and |eta| up to ~2000 (softplus tails)."""
    rng = np.random.RandomState(4)
    T, n, d = 2, 50, 2
    X = rng.randn(T, n, d)
    X[:, 5:9] = X[:, 4:5]            # four nodes on top of a fifth
    X[:, 20:25] *= 400.0             # far away: eta ~ -1000
    _run_both(T, n, d, False, _net(rng, T, n, False, 0.2), X)


def test_single_component_mixture_and_labels():
    L = _L()
    rng = np.random.RandomState(5)
    T, n, d, K = 3, 20, 2, 1
    mix = dict(mu=rng.randn(K, d), sigma=np.array([0.7]), lmbda=0.6, z=np.zeros((T, n), np.int64))
    e = _run_both(T, n, d, False, _net(rng, T, n, False, 0.3), rng.randn(T, n, d), mixture=mix)
    w = np.ones((T, K, K))
    e.set(L.F_WEIGHTS, w[None])
    U = rng.rand(n, T)
    e.sample_labels(U[None])
    assert np.all(e.get(L.F_Z) == 0) and np.all(e.get(L.F_NK)[0] == n)


def test_ragged_case_control_lists():
    """Nodes with no edges, with fewer non-neighbours than n_control (sentinel-terminated lists)
    and the reference's out-of-bounds configuration (flagged, never read)."""
    L = _L()
    from dynetlsm_b200 import DirectedCaseControlSampler
    rng = np.random.RandomState(6)
    T, n, d = 2, 30, 2
    Y = _net(rng, T, n, True, 0.15)
    Y[0, 3, :] = 0; Y[0, :, 3] = 0                      # isolated node
    Y[1, 7, :] = 1; Y[1, 7, 7] = 0; Y[1, 7, 10:14] = 0  # out-degree n-5: only 4 out-controls exist
    Y[1, :, 8] = 1; Y[1, 8, 8] = 0; Y[1, 0:3, 8] = 0    # in-degree n-4: only 3 in-controls exist
    cc = DirectedCaseControlSampler(n_control=6, n_resample=None, random_state=np.random.RandomState(1)).init(Y)
    n_in = (cc.control_nodes_in_ != -1).sum(axis=2)
    n_out = (cc.control_nodes_out_ != -1).sum(axis=2)
    assert (n_in < 6).any() and (n_out < n_in).any()    # ragged, and the UB pattern is present
    X = rng.randn(T, n, d) / n
    radii = rng.dirichlet(np.ones(n) * 5)
    e = L.Engine(T=T, n=n, d=d, is_directed=True, case_control=True)
    e.set_edge_lists(cc.degrees_, cc.in_edges_, cc.out_edges_)
    e.set_controls(cc.control_nodes_in_, cc.control_nodes_out_)
    e.set(L.F_X, X[None]); e.set(L.F_RADII, radii[None]); e.set(L.F_INTERCEPT, np.array([[0.3, 0.6]]))
    got = e.loglik_partial()[0]
    n_ub = 0
    for t in range(T):
        for i in range(n):
            ref, ub = O.approx_directed_partial_loglikelihood(
                X[t], radii, cc.in_edges_[t], cc.out_edges_[t], cc.degrees_[t], cc.control_nodes_in_[t],
                cc.control_nodes_out_[t], 0.3, 0.6, i, return_ub=True)
            n_ub += ub
            if np.isnan(ref):
                assert np.isnan(got[t, i])
            else:
                assert abs(got[t, i] - ref) <= 1e-10 * abs(ref) + 1e-13
    full = e.loglik_full()[0]
    ref = O.approx_directed_network_loglikelihood(X, radii, cc.out_edges_, cc.degrees_,
                                                  cc.control_nodes_out_, 0.3, 0.6)
    assert abs(full - ref) <= 1e-10 * abs(ref)
    assert n_ub > 0
    e.set_tuner(0.001)
    try:
        e.sweep_latent(rng.randn(1, T, n, d), np.log(rng.rand(1, T, n)))
    except L.DlsmError as err:       # a node with no non-neighbours gives 0/0 = NaN, as in the reference
        assert err.code == -5
    assert e.counters()["ub_flags"] >= 1


def test_shape_and_argument_errors():
    L = _L()
    e = L.Engine(T=2, n=5, d=2)
    with pytest.raises(ValueError):
        e.set(L.F_X, np.zeros((1, 2, 5, 3)))
    with pytest.raises(L.DlsmError) as ei:
        e.sweep_latent()                      # network not set
    assert ei.value.code == -4
    with pytest.raises(L.DlsmError):
        e.sample_radii()                      # undirected model has no radii
    with pytest.raises(L.DlsmError):
        e.sample_labels()                     # no mixture prior


def _sparse_directed(rng, T, n, out_deg):
    """Sparse directed network as padded edge lists (what dlsm_set_edge_lists takes)."""
    deg = np.zeros((T, n, 2), np.int64)
    outs = [[np.sort(rng.choice(np.delete(np.arange(n), i), size=rng.poisson(out_deg) % (n - 1), replace=False))
             for i in range(n)] for _ in range(T)]
    ins = [[[] for _ in range(n)] for _ in range(T)]
    for t in range(T):
        for i in range(n):
            deg[t, i, 1] = len(outs[t][i])
            for k in outs[t][i]:
                ins[t][k].append(i)
        for i in range(n):
            deg[t, i, 0] = len(ins[t][i])
    in_e = np.zeros((T, n, max(1, deg[:, :, 0].max())), np.int64)
    out_e = np.zeros((T, n, max(1, deg[:, :, 1].max())), np.int64)
    for t in range(T):
        for i in range(n):
            in_e[t, i, :deg[t, i, 0]] = ins[t][i]
            out_e[t, i, :deg[t, i, 1]] = outs[t][i]
    return deg, in_e, out_e


@pytest.mark.parametrize("mode", ["chain", "slice", "slice-v1", "slice-v2", "slice-v3", "slice-plain"])
@pytest.mark.parametrize("T,n,m,per_chain", [(3, 400, 8, False), (2, 150, 5, True), (4, 70, 20, False), (2, 300, 128, True)])
def test_case_control_sweep_all_mappings(T, n, m, per_chain, mode, monkeypatch):
    """The case-control sweep (directed_likelihoods_fast.pyx:83-182 inside
    sample_latent_positions.py:92-146) with recorded draws against the oracle: warp per slice
    ("chain"), batch-parallel CTA per slice ("slice": runs of mutually independent nodes updated
    concurrently, k_sweep_cc / _cc2 / _cc3; auto = the dataflow kernel k_sweep_ccd, which hands nodes to
    warps in index order and waits per dependency) and the serial CTA-per-slice kernel ("slice-plain")
    all make the oracle's decisions and leave its positions."""
    L = _L()
    if mode in ("slice-v1", "slice-v2", "slice-v3"):   # k_sweep_cc / _cc2 / _cc3; "slice" = auto = the dataflow kernel k_sweep_ccd
        monkeypatch.setenv("DLSM_CC_KERNEL", mode[-1])
        mode = "slice"
    monkeypatch.setenv("DLSM_SWEEP_MODE", mode)
    rng = np.random.RandomState(n + m)
    d, C_ = 2, 2
    deg, in_e, out_e = _sparse_directed(rng, T, n, 4.0)
    e = L.Engine(T=T, n=n, d=d, n_chains=C_, is_directed=True, case_control=True, tune=2, tune_interval=1)
    e.set_edge_lists(deg, in_e, out_e)
    e.set_rng(77)
    e.resample_controls(m, per_chain=per_chain)       # device-drawn control sets, read back for the oracle
    ci, co = e.get_controls()
    X = rng.randn(C_, T, n, d) * 0.02
    radii = rng.dirichlet(np.ones(n) * 5, size=C_)
    ic = np.tile([[0.4, 0.7]], (C_, 1))
    e.set(L.F_X, X); e.set(L.F_RADII, radii); e.set(L.F_INTERCEPT, ic)
    e.set_hyper(tau_sq=0.5, sigma_sq=0.002)
    e.set_tuner(0.004)
    Xo = X.copy()
    tuners = [O.TunerState((T, n), 0.004, tune=2, tune_interval=1) for _ in range(C_)]
    for s in range(3):
        eps = rng.randn(C_, T, n, d)
        logu = np.log(rng.rand(C_, T, n))
        acc, _ = e.sweep_latent(eps, logu, want_stats=True)
        got = e.get(L.F_X)
        for c in range(C_):
            k = c if per_chain else 0
            cc = dict(in_edges=in_e, out_edges=out_e, degrees=deg, ctrl_in=ci[k].astype(np.int64),
                      ctrl_out=co[k].astype(np.int64))
            out = O.sweep_latent(Xo[c], ic[c], tuners[c], eps[c], logu[c], radii=radii[c], is_directed=True,
                                 tau_sq=0.5, sigma_sq=0.002, case_control=cc)
            assert np.array_equal(acc[c], out["accepted"]), (mode, s, c)
            assert np.array_equal(got[c], Xo[c])
        assert 0.05 < acc.mean() < 0.98
    assert np.array_equal(e.get(L.F_X_STEP)[0], tuners[0].step)
    # the device loop on top of it: sweeps + intercept / radii MH, log-likelihood kept current
    e.run_sweeps(2)
    assert np.allclose(e.get(L.F_LOGLIK), e.loglik_full(), rtol=1e-10, atol=0)


@pytest.mark.parametrize("T,n,d", [(3, 9000, 2), (2, 12000, 3), (5, 5000, 8)])
def test_centring_of_long_chains(T, n, d, monkeypatch):
    """T n > 16384 rows: dlsm_center keeps numpy's summation order bit for bit (staged serial sum
    + a subtraction spread over many CTAs)."""
    L = _L()
    rng = np.random.RandomState(T * n)
    X = rng.randn(2, T, n, d) + rng.randn(2, 1, 1, d)
    e = L.Engine(T=T, n=n, d=d, n_chains=2)
    e.set(L.F_X, X)
    e.center()
    want = X.copy()
    for c in range(2):
        O.center(want[c])
    assert np.array_equal(e.get(L.F_X), want)


def test_device_loop_centring_tree_vs_exact(monkeypatch):
    """Inside dlsm_run_sweeps long chains are centred with a deterministic tree sum; against the
    numpy-ordered sum (DLSM_CENTER_EXACT=1) the chain differs by rounding only."""
    L = _L()
    rng = np.random.RandomState(3)
    T, n, d, m = 2, 9000, 2, 6
    deg, in_e, out_e = _sparse_directed(rng, T, n, 3.0)
    X0 = rng.randn(1, T, n, d) * 0.01 + 0.5
    radii = rng.dirichlet(np.ones(n) * 5, size=1)
    outs = []
    for exact in (False, True):
        if exact:
            monkeypatch.setenv("DLSM_CENTER_EXACT", "1")
        e = L.Engine(T=T, n=n, d=d, is_directed=True, case_control=True)
        e.set_edge_lists(deg, in_e, out_e)
        e.set_rng(5)
        e.resample_controls(m)
        e.set(L.F_X, X0); e.set(L.F_RADII, radii); e.set(L.F_INTERCEPT, np.array([[0.4, 0.7]]))
        e.set_hyper(tau_sq=0.5, sigma_sq=0.002)
        e.set_tuner(0.002)
        e.run_sweeps(2)
        outs.append(e.get(L.F_X))
    assert np.all(np.abs(outs[0].mean(axis=(1, 2))) < 1e-15)
    assert np.allclose(outs[0], outs[1], rtol=0, atol=1e-14)
    assert not np.allclose(outs[0], X0 - X0.mean(axis=(1, 2), keepdims=True), atol=1e-6)   # it moved


# ------------------------------------------------------------------------------------------
# sparse network input: the case-control edge lists built on the device from the ties
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,n,density,seed", [(3, 40, 0.2, 1), (2, 300, 0.02, 2), (4, 65, 0.6, 3), (1, 5, 0.0, 4)])
def test_device_edge_lists_equal_the_host_sampler_init(T, n, density, seed):
    """dlsm_set_network_edges (count / fill / sort on the device) == DirectedCaseControlSampler.init
    on the dense tensor (the reference's construction, case_control_likelihood.py:37-73)."""
    from dynetlsm_b200 import _lib as L
    from dynetlsm_b200.case_control_likelihood import DirectedCaseControlSampler
    rng = np.random.RandomState(seed)
    Y = (rng.rand(T, n, n) < density).astype(np.float64)
    for t in range(T):
        np.fill_diagonal(Y[t], 0)
    host = DirectedCaseControlSampler(n_control=5, random_state=0).init(Y, sample=False)
    edges = np.argwhere(Y == 1).astype(np.int32)
    edges = edges[rng.permutation(edges.shape[0])]           # any order
    e = L.Engine(T=T, n=n, d=2, is_directed=True, case_control=True)
    e.set_network_edges(edges)
    dg, ie, oe = e.get_edge_lists()
    assert np.array_equal(dg, host.degrees_)
    assert ie.shape == host.in_edges_.shape and np.array_equal(ie, host.in_edges_)
    assert oe.shape == host.out_edges_.shape and np.array_equal(oe, host.out_edges_)
    if edges.shape[0]:
        bad = edges.copy(); bad[0, 2] = bad[0, 1]             # a self tie
        with pytest.raises(L.DlsmError):
            e.set_network_edges(bad)
        with pytest.raises(L.DlsmError):
            e.set_network_edges(np.concatenate([edges, edges[:1]]))   # a tie listed twice
        bad = edges.copy(); bad[0, 0] = T
        with pytest.raises(L.DlsmError):
            e.set_network_edges(bad)


def test_intercept_and_radii_mh_long_chain_case_control_vs_oracle():
    """n = 6 000 > one Dirichlet chunk (k_radii_terms: 2 048 nodes per CTA) and T n / 64 = 282
    likelihood partials per chain (warp-per-chain sums): the intercept and radii MH steps
    (sample_coefficients.py:12-121 around directed_likelihoods_fast.pyx:208-270) replayed against the
    oracle, decisions and ratios."""
    L = _L()
    T, n, d, m, C_ = 3, 6000, 2, 30, 2
    rng = np.random.RandomState(6000)
    deg, in_e, out_e = _sparse_directed(rng, T, n, 5.0)
    e = L.Engine(T=T, n=n, d=d, n_chains=C_, is_directed=True, case_control=True, tune=4, tune_interval=2,
                 intercept_tune_interval=(2, 2), radii_tune=None)
    e.set_edge_lists(deg, in_e, out_e)
    e.set_rng(5)
    e.resample_controls(m, per_chain=True)
    ci, co = e.get_controls()
    X = rng.randn(C_, T, n, d) * 0.02
    radii = rng.dirichlet(np.ones(n) * 5, size=C_)
    ic = np.tile([[0.4, 0.7]], (C_, 1))
    e.set(L.F_X, X); e.set(L.F_RADII, radii); e.set(L.F_INTERCEPT, ic)
    e.set_hyper(tau_sq=0.5, sigma_sq=0.002, intercept_prior=np.array([0.4, 0.7]), intercept_variance_prior=2.0)
    e.set_tuner(0.004, step_intercept=0.003, step_radii=200000.0)
    itun = [O.TunerState((2,), 0.003, tune=4, tune_interval=2) for _ in range(C_)]
    rtun = [O.TunerState((1,), 200000.0, tune=None, tune_interval=100) for _ in range(C_)]
    seen = set()
    for s in range(3):
        ieps, ilogu = rng.randn(C_, 2), np.log(rng.rand(C_, 2))
        iacc, iratio = e.sample_intercepts(ieps, ilogu, want_stats=True)
        prop = np.stack([rng.dirichlet(200000.0 * radii[c]) for c in range(C_)])
        rlogu = np.log(rng.rand(C_))
        racc, rratio = e.sample_radii(prop, rlogu, want_stats=True)
        for c in range(C_):
            cc = dict(in_edges=in_e, out_edges=out_e, degrees=deg, ctrl_in=ci[c].astype(np.int64),
                      ctrl_out=co[c].astype(np.int64))
            io = O.sample_intercepts(X[c], ic[c], itun[c], ieps[c], ilogu[c], np.array([0.4, 0.7]), 2.0,
                                     radii=radii[c], is_directed=True, case_control=cc)
            assert np.array_equal(iacc[c], io["accepted"])
            assert np.allclose(iratio[c], io["ratio"], rtol=1e-8, atol=1e-8)
            ro = O.sample_radii(X[c], ic[c], radii[c], rtun[c], prop[c], float(rlogu[c]), case_control=cc)
            assert int(racc[c]) == ro["accepted"]
            assert abs(rratio[c] - ro["ratio"]) <= 1e-8 * max(1.0, abs(ro["ratio"]))
            seen.update((int(a) for a in io["accepted"]))
            seen.add(2 + ro["accepted"])
        assert np.array_equal(e.get(L.F_INTERCEPT), ic)
        assert np.array_equal(e.get(L.F_RADII), radii)
    assert len(seen) >= 2      # the replay saw both outcomes somewhere


@pytest.mark.parametrize("cc_kernel", ["auto", "1"])
def test_case_control_sweep_hub_node_long_and_ragged_lists(cc_kernel, monkeypatch):
    """A hub with in-degree n - 4 in one slice: its in-list (296 entries) does not fit the dataflow
    kernel's 256-slot register layout (generic list walk), only 3 in-controls exist (sentinel-terminated
    control lists, directed_likelihoods_fast.pyx:137), and every other node has the hub in its out-list.
    Recorded draws, decisions and positions against the oracle for the dataflow kernel (auto) and the
    run-based kernel."""
    L = _L()
    from dynetlsm_b200 import DirectedCaseControlSampler
    if cc_kernel != "auto":
        monkeypatch.setenv("DLSM_CC_KERNEL", cc_kernel)
    monkeypatch.setenv("DLSM_SWEEP_MODE", "slice")
    rng = np.random.RandomState(300)
    T, n, d, C_ = 2, 300, 2, 2
    Y = _net(rng, T, n, True, 0.015)
    Y[1, :, 7] = 1; Y[1, 7, 7] = 0; Y[1, 0:3, 7] = 0      # in-degree n - 4: only 3 in-controls exist
    cc = DirectedCaseControlSampler(n_control=6, n_resample=None, random_state=np.random.RandomState(1)).init(Y)
    n_in = (cc.control_nodes_in_ != -1).sum(axis=2)
    n_out = (cc.control_nodes_out_ != -1).sum(axis=2)
    assert n_in[1, 7] == 3 and (n_out >= n_in).all()      # ragged, without the reference's out-of-bounds pattern
    assert cc.in_edges_.shape[2] >= 296
    e = L.Engine(T=T, n=n, d=d, n_chains=C_, is_directed=True, case_control=True, tune=2, tune_interval=1)
    e.set_edge_lists(cc.degrees_, cc.in_edges_, cc.out_edges_)
    e.set_controls(cc.control_nodes_in_, cc.control_nodes_out_)
    X = rng.randn(C_, T, n, d) * 0.02
    radii = rng.dirichlet(np.ones(n) * 5, size=C_)
    ic = np.tile([[0.4, 0.7]], (C_, 1))
    e.set(L.F_X, X); e.set(L.F_RADII, radii); e.set(L.F_INTERCEPT, ic)
    e.set_hyper(tau_sq=0.5, sigma_sq=0.002)
    e.set_tuner(0.004)
    Xo = X.copy()
    tuners = [O.TunerState((T, n), 0.004, tune=2, tune_interval=1) for _ in range(C_)]
    lists = dict(in_edges=cc.in_edges_, out_edges=cc.out_edges_, degrees=cc.degrees_,
                 ctrl_in=cc.control_nodes_in_.astype(np.int64), ctrl_out=cc.control_nodes_out_.astype(np.int64))
    for s in range(3):
        eps = rng.randn(C_, T, n, d)
        logu = np.log(rng.rand(C_, T, n))
        acc, _ = e.sweep_latent(eps, logu, want_stats=True)
        got = e.get(L.F_X)
        for c in range(C_):
            out = O.sweep_latent(Xo[c], ic[c], tuners[c], eps[c], logu[c], radii=radii[c], is_directed=True,
                                 tau_sq=0.5, sigma_sq=0.002, case_control=lists)
            assert np.array_equal(acc[c], out["accepted"]), (s, c)
            assert np.array_equal(got[c], Xo[c])
        assert 0.05 < acc.mean() < 0.98
    assert e.counters()["ub_flags"] == 0
