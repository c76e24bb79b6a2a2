"""The device loop's row-sum cache (k_sweep with RS, k_rows, k_rows_commit).

With the exact likelihoods the loop keeps rows[c, t, j] = sum_i term(i, j) so that a node-update
evaluates its proposal only and an accepted move trades the old pair terms for the new ones.  The
checks: (1) k_rows' row sums and total equal the per-node / full-network probes (reference
kernels K1/K2/K4/K5, 1e-10); (2) a sweep on the cache makes the oracle's decisions on the device's
own draws, bit for bit; (3) whole device-loop chains equal the chains of the two-variant path
(DLSM_OPT_NO_ROWSUM_CACHE) bit for bit; (4) the cache stays equal to fresh evaluations along a
chain with intercept / radii moves.  Needs a GPU: ``-m gpu``.
"""
import numpy as np
import pytest

import pyoracle as O

pytestmark = pytest.mark.gpu


def _L():
    from dynetlsm_b200 import _lib
    return _lib


def _net(T, n, d, directed, seed, density=0.2):
    rng = np.random.RandomState(seed)
    scale = 1.0 / n if directed else 1.0
    X = rng.randn(T, n, d) * scale
    Y = (rng.rand(T, n, n) < density).astype(np.float64)
    for t in range(T):
        np.fill_diagonal(Y[t], 0)
    if not directed:
        Y = np.triu(Y, 1)
        Y = Y + Y.transpose(0, 2, 1)
    return rng, X, Y


def _engine(T, n, d, directed, C_, X, Y, rng, K=0, tune=4, tune_interval=2):
    L = _L()
    e = L.Engine(T=T, n=n, d=d, n_chains=C_, is_directed=directed, K=K, mixture=K > 0, tune=tune,
                 tune_interval=tune_interval)
    e.set_network(Y)
    Xs = np.stack([X * (1.0 + 0.02 * c) for c in range(C_)])
    e.set(L.F_X, Xs)
    e.set(L.F_INTERCEPT, np.tile([[0.6, 0.35]], (C_, 1)))
    radii = None
    if directed:
        radii = np.random.RandomState(9).dirichlet(np.ones(n) * 4, size=C_)
        e.set(L.F_RADII, radii)
        e.set_hyper(tau_sq=float(np.mean(X[0] * X[0])), sigma_sq=0.001 / n, intercept_prior=(0.6, 0.35))
        e.set_tuner(0.02 / n)
    else:
        e.set_hyper(tau_sq=2.0, sigma_sq=0.1, intercept_prior=(0.6, 0.0))
        e.set_tuner(0.12)
    if K:
        r2 = np.random.RandomState(4)
        e.set(L.F_MU, np.tile(r2.randn(1, K, d), (C_, 1, 1)))
        e.set(L.F_SIGMA, np.tile(r2.gamma(2, 1, (1, K)), (C_, 1)))
        e.set(L.F_LAMBDA, np.full(C_, 0.8))
        e.set(L.F_WEIGHTS, np.tile(r2.dirichlet(np.ones(K), size=(1, T, K)), (C_, 1, 1, 1)))
        e.set(L.F_Z, np.tile(r2.randint(0, K, (1, T, n)), (C_, 1, 1)))
    e.set_rng(77)
    return e, Xs, radii


SHAPES = [(3, 33, 2, False), (4, 64, 2, False), (9, 120, 2, False), (2, 129, 3, False), (3, 500, 2, False),
          (2, 700, 2, False),      # 22 row blocks: two runs of column blocks per row block
          (3, 90, 2, True), (2, 200, 3, True), (2, 600, 2, True)]


@pytest.mark.parametrize("T,n,d,directed", SHAPES)
def test_k_rows_equals_the_per_node_and_full_network_probes(T, n, d, directed):
    L = _L()
    rng, X, Y = _net(T, n, d, directed, seed=n)
    e, _, _ = _engine(T, n, d, directed, 2, X, Y, rng)
    e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_CHAIN)
    e.set_option(L.OPT_CHAIN_KERNEL, L.CHAIN_NODE_ROWSUM)
    rows = e.rowsums()
    fresh = e.loglik_partial()
    assert np.allclose(rows, fresh, rtol=1e-11, atol=0)
    assert np.allclose(0.5 * rows.sum(axis=(1, 2)), e.loglik_full(), rtol=1e-11, atol=0)


@pytest.mark.parametrize("T,n,d,directed", [(6, 64, 2, False), (9, 120, 2, False), (4, 130, 3, False),
                                            (3, 500, 2, False), (5, 90, 2, True), (3, 300, 2, True)])
def test_cached_sweep_makes_the_oracle_decisions_on_the_device_draws(T, n, d, directed):
    L = _L()
    rng, X, Y = _net(T, n, d, directed, seed=7 * n)
    C_ = 3
    e, Xs, radii = _engine(T, n, d, directed, C_, X, Y, rng)
    e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_CHAIN)
    e.set_option(L.OPT_CHAIN_KERNEL, L.CHAIN_NODE_ROWSUM)
    ic = np.array([0.6, 0.35]) if directed else np.array([0.6])
    hy = dict(tau_sq=float(np.mean(X[0] * X[0])), sigma_sq=0.001 / n) if directed else dict(tau_sq=2.0, sigma_sq=0.1)
    tun = [O.TunerState((T, n), 0.02 / n if directed else 0.12, tune=4, tune_interval=2) for _ in range(C_)]
    Xo = [Xs[c].copy() for c in range(C_)]
    for s in range(4):
        eps, logu = e.debug_draws()
        e.run_sweeps(1, skip_center=True, skip_intercepts=True, skip_radii=True)   # the cached path
        got = e.get(L.F_X)
        for c in range(C_):
            out = O.sweep_latent(Xo[c], ic, tun[c], eps[c], logu[c], Y=Y, radii=None if radii is None else radii[c],
                                 is_directed=directed, **hy)
            assert np.array_equal(got[c], Xo[c]), (s, c)
        assert 0.03 < out["accepted"].mean() < 0.97
        assert np.allclose(e.rowsums(), e.loglik_partial(), rtol=1e-11, atol=0)
    assert np.array_equal(e.get(L.F_X_STEP)[0], tun[0].step)


@pytest.mark.parametrize("T,n,d,directed,K", [(9, 120, 2, False, 10), (4, 70, 3, False, 0), (10, 500, 2, False, 10),
                                              (5, 90, 2, True, 0), (3, 257, 2, True, 6), (2, 1500, 2, False, 0)])
def test_device_loop_chain_is_the_two_variant_chain(T, n, d, directed, K):
    """Same Philox streams with and without the cache: the decisions differ only if a ratio lands
    within ~1e-13 of log(u), so whole chains (positions, intercepts, radii, labels) are identical."""
    L = _L()
    outs = []
    for cached in (True, False):
        rng, X, Y = _net(T, n, d, directed, seed=3 * n + T)
        e, _, _ = _engine(T, n, d, directed, 3, X, Y, rng, K=K, tune=500, tune_interval=3)
        e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_CHAIN)
        e.set_option(L.OPT_CHAIN_KERNEL, L.CHAIN_NODE_ROWSUM if cached else L.CHAIN_NODE)
        e.run_sweeps(3 if n >= 500 else 8, skip_hdp=True)
        outs.append([e.get(f) for f in (L.F_X, L.F_INTERCEPT, L.F_LOGLIK)] +
                    ([e.get(L.F_RADII)] if directed else []) + ([e.get(L.F_Z)] if K else []))
        if cached:
            assert np.allclose(e.rowsums(), e.loglik_partial(), rtol=1e-11, atol=0)
            assert np.allclose(e.get(L.F_LOGLIK), e.loglik_full(), rtol=1e-11, atol=0)
    for a, b in zip(*outs):
        if a.dtype == np.float64 and a.ndim == 1:     # tracked log-likelihood: summation order differs
            assert np.allclose(a, b, rtol=1e-11, atol=0)
        else:
            assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------
# thread-block-cluster sweep kernel (k_sweep_slice_cl): a (chain, slice) spread over CS CTAs with
# partial sums through distributed shared memory
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,n,d,directed", [(5, 300, 2, False), (4, 500, 2, True), (3, 257, 3, True),
                                            (7, 1000, 2, False), (2, 2000, 2, True), (20, 640, 2, True),
                                            (70, 256, 2, True), (37, 290, 2, False)])   # clusters of 2 and 4 CTAs
def test_cluster_kernel_vs_oracle_and_vs_cta_per_slice(T, n, d, directed):
    L = _L()
    rng, X, Y = _net(T, n, d, directed, seed=11 * n + T)
    ic = np.array([0.6, 0.35]) if directed else np.array([0.6])
    step = 0.02 / n if directed else 0.12
    hy = dict(tau_sq=float(np.mean(X[0] * X[0])), sigma_sq=0.001 / n) if directed else dict(tau_sq=2.0, sigma_sq=0.1)
    engines = []
    for no_cluster in (0, 3, 2, 1):   # block-speculative cluster kernel with / without the two-block window, per-node cluster kernel, CTA per slice
        e, Xs, radii = _engine(T, n, d, directed, 1, X, Y, rng)
        e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_SLICE)
        e.set_option(L.OPT_NO_CLUSTER, no_cluster)
        engines.append(e)
    tun = O.TunerState((T, n), step, tune=4, tune_interval=2)
    Xo = Xs[0].copy()
    for s in range(3):
        eps, logu = rng.randn(1, T, n, d), np.log(rng.rand(1, T, n))
        out = O.sweep_latent(Xo, ic, tun, eps[0], logu[0], Y=Y, radii=None if radii is None else radii[0],
                             is_directed=directed, **hy)
        for e in engines:
            acc, ratio = e.sweep_latent(eps, logu, want_stats=True)
            assert np.array_equal(acc[0], out["accepted"]), s
            assert np.array_equal(e.get(L.F_X)[0], Xo)
            assert np.allclose(ratio[0], out["ratio"], rtol=1e-8, atol=1e-8)
    assert [e.counters()["cluster_sweeps"] for e in engines] == [3, 3, 3, 0]
    assert np.array_equal(engines[0].get(L.F_X_STEP)[0], tun.step)
    assert 0.03 < out["accepted"].mean() < 0.97


@pytest.mark.parametrize("T,n,K", [(4, 300, 6), (10, 500, 10)])
def test_cluster_kernels_mixture_prior_vs_oracle(T, n, K):
    """The HDP-LPCM sweep (mixture prior, sample_latent_positions.py:149-206) of a single long-row chain
    -- what DynamicNetworkHDPLPCM.fit(n_chains=1) runs for n >= 256 -- on every (chain, slice) kernel,
    recorded draws, the oracle's decisions and positions."""
    L = _L()
    d = 2
    rng, X, Y = _net(T, n, d, False, seed=7 * n + K)
    engines = []
    for no_cluster in (0, 3, 2, 1):
        e, Xs, _ = _engine(T, n, d, False, 1, X, Y, rng, K=K)
        e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_SLICE)
        e.set_option(L.OPT_NO_CLUSTER, no_cluster)
        engines.append(e)
    e0 = engines[0]
    mix = dict(mu=e0.get(L.F_MU)[0], sigma=e0.get(L.F_SIGMA)[0], lmbda=0.8, z=e0.get(L.F_Z)[0].astype(np.int64))
    tun = O.TunerState((T, n), 0.12, tune=4, tune_interval=2)
    Xo = Xs[0].copy()
    ic = np.array([0.6])
    for s in range(2):
        eps, logu = rng.randn(1, T, n, d), np.log(rng.rand(1, T, n))
        out = O.sweep_latent(Xo, ic, tun, eps[0], logu[0], Y=Y, tau_sq=2.0, sigma_sq=0.1, mixture=mix)
        for e in engines:
            acc, ratio = e.sweep_latent(eps, logu, want_stats=True)
            assert np.array_equal(acc[0], out["accepted"]), s
            assert np.array_equal(e.get(L.F_X)[0], Xo)
    assert [e.counters()["cluster_sweeps"] for e in engines] == [2, 2, 2, 0]
    assert 0.03 < out["accepted"].mean() < 0.97


def test_cluster_kernel_native_rng_chain_equals_cta_per_slice_chain():
    L = _L()
    T, n, d = 6, 400, 2
    outs = []
    for no_cluster in (0, 1, 2, 3):
        rng, X, Y = _net(T, n, d, True, seed=5)
        e, _, _ = _engine(T, n, d, True, 2, X, Y, rng, tune=500, tune_interval=3)
        e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_SLICE)
        e.set_option(L.OPT_NO_CLUSTER, no_cluster)
        e.run_sweeps(6)
        outs.append([e.get(f) for f in (L.F_X, L.F_INTERCEPT, L.F_RADII)])
        assert (e.counters()["cluster_sweeps"] > 0) == (no_cluster != 1)
        # (the windowed cluster kernel tracks the log-likelihood itself: dyads {i < j} at the kept states)
        assert np.allclose(e.get(L.F_LOGLIK), e.loglik_full(), rtol=1e-11, atol=0)
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------
# block-speculative chain kernel (k_sweep_cb): one CTA per chain, 32 nodes per step
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,n,d,directed,K,C_", [(9, 120, 2, False, 10, 3), (4, 70, 3, False, 0, 2), (10, 500, 2, False, 10, 150),
                                                 (5, 90, 2, True, 0, 2), (3, 257, 2, True, 6, 2), (2, 1500, 2, False, 0, 1),
                                                 (1, 33, 2, False, 0, 2), (3, 31, 2, True, 0, 200), (17, 40, 2, False, 0, 2)])
@pytest.mark.parametrize("kern", ["block", "block-pair"])
def test_block_chain_kernel_replay_vs_oracle(T, n, d, directed, K, C_, kern):
    """Recorded draws through k_sweep_cb (one warp per slice) and k_sweep_cbp ("block-pair": two warps per
    slice where a chain has an SM to itself: C <= 148, d = 2, T <= 15): the oracle's decisions and
    states, bit for bit."""
    L = _L()
    rng, X, Y = _net(T, n, d, directed, seed=13 * n + T)
    e, Xs, radii = _engine(T, n, d, directed, C_, X, Y, rng, K=K)
    e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_CHAIN if C_ <= 148 else L.SWEEP_CHAIN_DENSE)
    e.set_option(L.OPT_CHAIN_KERNEL, L.CHAIN_BLOCK if kern == "block" else L.CHAIN_BLOCK_PAIR)
    ic = np.array([0.6, 0.35]) if directed else np.array([0.6])
    step = 0.02 / n if directed else 0.12
    hy = dict(tau_sq=float(np.mean(X[0] * X[0])), sigma_sq=0.001 / n) if directed else dict(tau_sq=2.0, sigma_sq=0.1)
    mix = None
    if K:
        mix = dict(mu=e.get(L.F_MU)[0], sigma=e.get(L.F_SIGMA)[0], lmbda=0.8, z=e.get(L.F_Z)[0].astype(np.int64))
    check = sorted(set([0, C_ // 2, C_ - 1]))
    tun = {c: O.TunerState((T, n), step, tune=4, tune_interval=2) for c in check}
    Xo = {c: Xs[c].copy() for c in check}
    for s in range(3 if n < 500 else 2):
        eps, logu = rng.randn(C_, T, n, d), np.log(rng.rand(C_, T, n))
        acc, ratio = e.sweep_latent(eps, logu, want_stats=True)
        got = e.get(L.F_X)
        for c in check:
            out = O.sweep_latent(Xo[c], ic, tun[c], eps[c], logu[c], Y=Y, radii=None if radii is None else radii[c],
                                 is_directed=directed, mixture=mix, **hy)
            assert np.array_equal(acc[c], out["accepted"]), (s, c, int((acc[c] != out["accepted"]).sum()))
            assert np.array_equal(got[c], Xo[c])
            assert np.allclose(ratio[c], out["ratio"], rtol=1e-8, atol=1e-8)
        assert 0.03 < acc.mean() < 0.97
    assert np.array_equal(e.get(L.F_X_STEP)[check[0]], tun[check[0]].step)


@pytest.mark.parametrize("T,n,d,directed,K", [(9, 120, 2, False, 10), (10, 500, 2, False, 10), (5, 90, 2, True, 0)])
def test_block_chain_kernel_device_loop_equals_node_kernel_loop(T, n, d, directed, K):
    L = _L()
    outs = []
    for kern in (L.CHAIN_BLOCK, L.CHAIN_NODE, L.CHAIN_BLOCK_PAIR):
        rng, X, Y = _net(T, n, d, directed, seed=3 * n + T)
        e, _, _ = _engine(T, n, d, directed, 3, X, Y, rng, K=K, tune=500, tune_interval=3)
        e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_CHAIN)
        e.set_option(L.OPT_CHAIN_KERNEL, kern)
        e.run_sweeps(3 if n >= 500 else 8, skip_hdp=True)
        outs.append([e.get(f) for f in (L.F_X, L.F_INTERCEPT)] + ([e.get(L.F_RADII)] if directed else []) +
                    ([e.get(L.F_Z)] if K else []))
        assert np.allclose(e.get(L.F_LOGLIK), e.loglik_full(), rtol=1e-11, atol=0)
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------
# k_full_lr: the one-variant full-network log-likelihood with lanes = rows (intercept / radii MH of the
# device loop) against the folded-rows kernel k_full
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,n,directed,C_", [(3, 33, False, 2), (9, 120, False, 3), (4, 257, True, 2),
                                             (2, 500, False, 5), (3, 1000, True, 1), (2, 2000, True, 1)])
def test_full_network_kernel_lanes_are_rows_equals_folded_rows(T, n, directed, C_):
    L = _L()
    outs = []
    for full_kernel in (2, 1):       # k_full_lr at any n, k_full
        rng, X, Y = _net(T, n, 2, directed, seed=5 * n + T)
        e, _, _ = _engine(T, n, 2, directed, C_, X, Y, rng, tune=500, tune_interval=3)
        e.set_option(L.OPT_FULL_KERNEL, full_kernel)
        e.run_sweeps(4)
        assert np.allclose(e.get(L.F_LOGLIK), e.loglik_full(), rtol=1e-11, atol=0)   # tracked through the MH steps
        outs.append([e.get(f) for f in (L.F_X, L.F_INTERCEPT)] + ([e.get(L.F_RADII)] if directed else []) +
                    [e.get(L.F_LOGLIK)])
    for a, b in zip(*outs):
        if a.ndim == 1:
            assert np.allclose(a, b, rtol=1e-11, atol=0)
        else:
            assert np.array_equal(a, b)      # same decisions: the two kernels differ in summation order only
