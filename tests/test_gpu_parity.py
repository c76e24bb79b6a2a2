"""Parity of the CUDA path (through the C-ABI, dynetlsm_b200._lib.Engine) with the CPU oracle and
with the golden vectors recorded from the reference.  Needs a GPU (B200 box): ``-m gpu``.

Tolerances (BASELINE.json north_star): accept/reject decisions, accepted states, tuner
trajectories and label draws are compared EXACTLY (bit-for-bit); log-likelihood values within
1e-10 relative.
"""
import numpy as np
import pytest

import pyoracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _engine(**kw):
    from dynetlsm_b200 import _lib
    return _lib.Engine(**kw)


def _F():
    from dynetlsm_b200 import _lib
    return _lib


def rel_close(a, b, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + 1e-300)


# ------------------------------------------------------------------------------------------
# kernel known-answer tests against the reference's Cython outputs
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_k1_partial_undirected(kernels_golden, tag):
    g, L = kernels_golden, _F()
    X = g[tag + "_X"]
    T, n, d = X.shape
    e = _engine(T=T, n=n, d=d)
    e.set_network(g[tag + "_Yu"].astype(np.float64))
    e.set(L.F_X, X[None])
    e.set(L.F_INTERCEPT, np.array([[g[tag + "_b"][0], 0.0]]))
    got = e.loglik_partial()[0]
    assert rel_close(got, g[tag + "_k1"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_k2_partial_directed(kernels_golden, tag):
    g, L = kernels_golden, _F()
    X = g[tag + "_X"]
    T, n, d = X.shape
    e = _engine(T=T, n=n, d=d, is_directed=True)
    e.set_network(g[tag + "_Yd"].astype(np.float64))
    e.set(L.F_X, (X / n)[None])
    e.set(L.F_RADII, g[tag + "_radii"][None])
    e.set(L.F_INTERCEPT, g[tag + "_b"][None, 1:])
    got = e.loglik_partial()[0]
    assert rel_close(got, g[tag + "_k2"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_k3_partial_case_control(kernels_golden, tag):
    g, L = kernels_golden, _F()
    X = g[tag + "_X"]
    T, n, d = X.shape
    e = _engine(T=T, n=n, d=d, is_directed=True, case_control=True)
    e.set_edge_lists(g[tag + "_degrees"], g[tag + "_in_edges"], g[tag + "_out_edges"])
    e.set_controls(g[tag + "_ctrl_in"], g[tag + "_ctrl_out"])
    e.set(L.F_X, (X / n)[None])
    e.set(L.F_RADII, g[tag + "_radii"][None])
    e.set(L.F_INTERCEPT, g[tag + "_b"][None, 1:])
    got = e.loglik_partial()[0]
    ok = g[tag + "_cc_ok"]
    assert ok.sum() > 10
    assert rel_close(got[ok], g[tag + "_k3"][ok])
    # K6, the out-side-only full-network estimator
    full = e.loglik_full()[0]
    assert rel_close(full, float(g[tag + "_k6"]))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_k4_k5_full_network(kernels_golden, tag):
    g, L = kernels_golden, _F()
    X = g[tag + "_X"]
    T, n, d = X.shape
    e = _engine(T=T, n=n, d=d)
    e.set_network(g[tag + "_Yu"].astype(np.float64))
    e.set(L.F_X, X[None])
    e.set(L.F_INTERCEPT, np.array([[g[tag + "_b"][0], 0.0]]))
    assert rel_close(e.loglik_full()[0], float(g[tag + "_k5"]))
    e = _engine(T=T, n=n, d=d, is_directed=True)
    e.set_network(g[tag + "_Yd"].astype(np.float64))
    e.set(L.F_X, (X / n)[None])
    e.set(L.F_RADII, g[tag + "_radii"][None])
    e.set(L.F_INTERCEPT, g[tag + "_b"][None, 1:])
    assert rel_close(e.loglik_full()[0], float(g[tag + "_k4"]))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_k7_gaussian_likelihood(kernels_golden, tag):
    g, L = kernels_golden, _F()
    X = g[tag + "_X"]
    T, n, d = X.shape
    K = g[tag + "_sigma"].shape[0]
    e = _engine(T=T, n=n, d=d, K=K, mixture=True)
    e.set(L.F_X, X[None])
    e.set(L.F_MU, g[tag + "_mu"][None])
    e.set(L.F_SIGMA, g[tag + "_sigma"][None])
    e.set(L.F_LAMBDA, np.array([float(g[tag + "_lmbda"])]))
    got = e.gaussian_likelihood()[0]
    assert rel_close(got, g[tag + "_k7"], 1e-12)


def test_nonbinary_network_is_refused():
    e = _engine(T=1, n=4, d=2)
    Y = np.zeros((1, 4, 4))
    Y[0, 1, 2] = -1.0
    with pytest.raises(_F().DlsmError) as ei:
        e.set_network(Y)
    assert ei.value.code == -3


def test_center_bitwise(kernels_golden):
    L = _F()
    Xc = kernels_golden["center_in"]
    T, n, d = Xc.shape
    e = _engine(T=T, n=n, d=d)
    e.set(L.F_X, Xc[None])
    e.center()
    assert np.array_equal(e.get(L.F_X)[0], kernels_golden["center_out"])


# ------------------------------------------------------------------------------------------
# teacher-forced replay of recorded reference sweeps: every recorded sweep becomes one chain
# ------------------------------------------------------------------------------------------
def _load_tuner(e, g, prefix, fields):
    L = _F()
    e.set(fields[0], g[prefix + "step"])
    e.set(fields[1], g[prefix + "n_accepted"])
    e.set(fields[2], g[prefix + "n_steps"])
    e.set(fields[3], g[prefix + "until"])


def _lsm_engine(g, directed, cc):
    L = _F()
    S, T, n, d = g["X_in"].shape
    e = _engine(T=T, n=n, d=d, n_chains=S, is_directed=directed, case_control=cc,
                tune=int(g["tune"]), tune_interval=int(g["tune_interval"]),
                intercept_tune_interval=(int(g["tune_interval"]),) * 2 if directed else (100, 100),
                radii_tune=None)
    if cc:
        e.set_edge_lists(g["cc_degrees"], g["cc_in_edges"], g["cc_out_edges"])
        e.set_controls(g["ctrl_in"], g["ctrl_out"])
    else:
        e.set_network(g["Y"].astype(np.float64))
    e.set_hyper(tau_sq=float(g["tau_sq"]), sigma_sq=float(g["sigma_sq"]),
                intercept_prior=g["intercept_prior"],
                intercept_variance_prior=float(g["intercept_variance_prior"]))
    ic = np.zeros((S, 2))
    ic[:, :g["intercept_in"].shape[1]] = g["intercept_in"]
    e.set(L.F_INTERCEPT, ic)
    if directed:
        e.set(L.F_RADII, g["radii_in"])
    return e


LSM_CASES = [("lsm_undirected_monks.npz", False, False), ("lsm_directed_monks.npz", True, False),
             ("lsm_casecontrol_monks.npz", True, True)]


@pytest.mark.parametrize("mode", ["chain", "chain-dense", "slice"])
@pytest.mark.parametrize("name,directed,cc", LSM_CASES)
def test_lsm_latent_sweep_replay(name, directed, cc, mode, monkeypatch):
    # the sweep kernels: one CTA per chain (warp per slice) in its few-chains build ("chain": all
    # registers) and in its many-chains build ("chain-dense": 72 registers, 3 CTAs per SM), and one
    # CTA per (chain, slice)
    monkeypatch.setenv("DLSM_SWEEP_MODE", mode)
    g, L = load_golden(name), _F()
    S = g["X_in"].shape[0]
    e = _lsm_engine(g, directed, cc)
    e.set(L.F_X, g["X_in"])
    _load_tuner(e, g, "tuner_", (L.F_X_STEP, L.F_X_NACC, L.F_X_NSTEPS, L.F_X_UNTIL))
    acc, ratio = e.sweep_latent(g["eps"], g["logu"], want_stats=True)
    assert np.array_equal(acc, g["accepted"])          # bit-exact decisions
    assert np.array_equal(e.get(L.F_X), g["X_out"])     # bit-exact accepted states
    assert np.allclose(ratio, g["ratio"], rtol=1e-9, atol=1e-9)
    # tuner trajectories: state after sweep s == recorded state before sweep s+1
    assert np.array_equal(e.get(L.F_X_STEP)[:-1], g["tuner_step"][1:])
    assert np.array_equal(e.get(L.F_X_NACC)[:-1], g["tuner_n_accepted"][1:])
    assert np.array_equal(e.get(L.F_X_UNTIL)[:-1], g["tuner_until"][1:])
    assert np.array_equal(e.get(L.F_X_NSTEPS)[:-1], g["tuner_n_steps"][1:])
    assert e.counters()["ub_flags"] == 0
    # per-node log-posteriors the decisions were based on: loglik part within 1e-10
    e.set(L.F_X, g["X_in"])
    assert S > 10


@pytest.mark.parametrize("name,directed,cc", LSM_CASES)
def test_lsm_intercepts_radii_replay(name, directed, cc):
    g, L = load_golden(name), _F()
    S = g["X_in"].shape[0]
    m = 2 if directed else 1
    e = _lsm_engine(g, directed, cc)
    e.set(L.F_X, g["X_centered"])
    st = np.zeros((S, 2)); st[:, :m] = g["itun_step"]
    na = np.zeros((S, 2), np.int32); na[:, :m] = g["itun_n_accepted"]
    ns = np.zeros((S, 2), np.int32); ns[:, :m] = g["itun_n_steps"]
    un = np.zeros((S, 2), np.int32); un[:, :m] = g["itun_until"]
    e.set(L.F_B_STEP, st); e.set(L.F_B_NACC, na); e.set(L.F_B_NSTEPS, ns); e.set(L.F_B_UNTIL, un)
    acc, ratio = e.sample_intercepts(g["i_eps"].reshape(S, m), g["i_logu"].reshape(S, m),
                                     want_stats=True)
    assert np.array_equal(acc, g["i_accepted"].reshape(S, m))
    assert np.array_equal(e.get(L.F_INTERCEPT)[:, :m], g["intercept_out"].reshape(S, m))
    assert np.allclose(ratio, g["i_ratio"].reshape(S, m), rtol=1e-8, atol=1e-7)
    assert np.array_equal(e.get(L.F_B_STEP)[:-1, :m], g["itun_step"][1:].reshape(S - 1, m))
    if directed:
        e.set(L.F_R_STEP, g["rtun_step"].reshape(S))
        e.set(L.F_R_NACC, g["rtun_n_accepted"].reshape(S))
        e.set(L.F_R_NSTEPS, g["rtun_n_steps"].reshape(S))
        e.set(L.F_R_UNTIL, g["rtun_until"].reshape(S))
        acc, ratio = e.sample_radii(g["r_proposal"], g["r_logu"], want_stats=True)
        assert np.array_equal(acc, g["r_accepted"])
        assert np.array_equal(e.get(L.F_RADII), g["radii_out"])
        assert np.allclose(ratio, g["r_ratio"], rtol=1e-7, atol=1e-6)


def test_lsm_center_replay():
    g, L = load_golden("lsm_undirected_monks.npz"), _F()
    n_pre = int(g["tune"]) + int(g["burn"])  # lsm.py:495: no Procrustes before tune+burn
    e = _lsm_engine(g, False, False)
    e.set(L.F_X, g["X_out"])
    e.center()
    assert np.array_equal(e.get(L.F_X)[:n_pre], g["X_centered"][:n_pre])


HDP_CASES = [("hdp_undirected_split.npz", False), ("hdp_directed_monks.npz", True)]


@pytest.mark.parametrize("mode", ["chain", "chain-dense", "slice"])
@pytest.mark.parametrize("name,directed", HDP_CASES)
def test_hdp_sweep_center_labels_replay(name, directed, mode, monkeypatch):
    monkeypatch.setenv("DLSM_SWEEP_MODE", mode)
    g, L = load_golden(name), _F()
    S, T, n, d = g["X_in"].shape
    K = g["sigma"].shape[1]
    e = _engine(T=T, n=n, d=d, n_chains=S, K=K, is_directed=directed, mixture=True,
                tune=int(g["tune"]), tune_interval=int(g["tune_interval"]))
    e.set_network(g["Y"].astype(np.float64))
    ic = np.zeros((S, 2)); ic[:, :g["intercept_in"].shape[1]] = g["intercept_in"]
    e.set(L.F_INTERCEPT, ic)
    if directed:
        e.set(L.F_RADII, g["radii_in"])
    e.set(L.F_MU, g["mu"]); e.set(L.F_SIGMA, g["sigma"]); e.set(L.F_LAMBDA, g["lmbda"].reshape(S))
    e.set(L.F_Z, g["z_in"]); e.set(L.F_WEIGHTS, g["w"])
    e.set(L.F_X, g["X_in"])
    _load_tuner(e, g, "tuner_", (L.F_X_STEP, L.F_X_NACC, L.F_X_NSTEPS, L.F_X_UNTIL))
    acc, ratio = e.sweep_latent(g["eps"], g["logu"], want_stats=True)
    assert np.array_equal(acc, g["accepted"])
    assert np.array_equal(e.get(L.F_X), g["X_out"])
    assert np.array_equal(e.get(L.F_X_STEP)[:-1], g["tuner_step"][1:])
    e.center()
    assert np.array_equal(e.get(L.F_X), g["X_centered"])
    e.sample_labels(g["U"])
    assert np.array_equal(e.get(L.F_Z), g["z_out"])     # bit-exact label draws
    assert np.array_equal(e.get(L.F_NCOUNT), g["n_out"])
    assert np.array_equal(e.get(L.F_NK), g["nk_out"])


# ------------------------------------------------------------------------------------------
# CUDA vs oracle on seeded synthetic inputs at the benchmark shapes (free-running, several sweeps)
# ------------------------------------------------------------------------------------------
def _synthetic(T, n, d, directed, seed, density=0.15):
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


@pytest.mark.parametrize("T,n,d,directed,mode", [
    (9, 120, 2, False, "auto"), (4, 70, 3, False, "auto"), (5, 90, 2, True, "auto"),
    (10, 500, 2, False, "chain"),     # positions in shared memory, 16 chunks per row
    (10, 500, 2, False, "chain-dense"),  # the same with the many-chains register budget
    (9, 120, 2, False, "chain-dense"), (5, 90, 2, True, "chain-dense"),
    (10, 500, 2, False, "slice"),     # CTA per slice, 8 warps per row
    (6, 700, 2, True, "slice"),       # directed, CTA per slice
    (10, 1500, 2, False, "chain"),    # chain too big for shared memory: positions stay in global/L2
    (2, 1900, 8, False, "slice"),     # slice too big for half the shared memory: global positions, d = 8
])
def test_free_running_replay_vs_oracle(T, n, d, directed, mode, monkeypatch):
    if mode != "auto":
        monkeypatch.setenv("DLSM_SWEEP_MODE", mode)
    L = _F()
    rng, X, Y = _synthetic(T, n, d, directed, seed=T * 1000 + n)
    radii = rng.dirichlet(np.ones(n) * 5) if directed else None
    ic = np.array([0.4, 0.7]) if directed else np.array([0.9])
    step0 = 0.0075 / 4 if directed else 0.08
    sig = 0.001 if directed else 0.1
    tau = float(np.mean(X[0] * X[0])) if directed else 2.0
    n_sweeps = 2 if n >= 1500 else (3 if n >= 500 else 6)
    tun = O.TunerState((T, n), step0, tune=4, tune_interval=2)
    e = _engine(T=T, n=n, d=d, is_directed=directed, tune=4, tune_interval=2)
    e.set_network(Y)
    e.set_hyper(tau_sq=tau, sigma_sq=sig)
    icd = np.zeros((1, 2)); icd[0, :ic.size] = ic
    e.set(L.F_INTERCEPT, icd)
    if directed:
        e.set(L.F_RADII, radii[None])
    e.set(L.F_X, X[None])
    e.set_tuner(step0)
    Xo = X.copy()
    for s in range(n_sweeps):
        eps = rng.randn(T, n, d)
        logu = np.log(rng.rand(T, n))
        out = O.sweep_latent(Xo, ic, tun, eps, logu, Y=Y, radii=radii, is_directed=directed,
                             tau_sq=tau, sigma_sq=sig)
        acc, ratio = e.sweep_latent(eps[None], logu[None], want_stats=True)
        assert np.array_equal(acc[0], out["accepted"]), "sweep %d" % s
        assert np.array_equal(e.get(L.F_X)[0], Xo)
        assert np.allclose(ratio[0], out["ratio"], rtol=1e-9, atol=1e-9)
        O.center(Xo)
        e.center()
        assert np.array_equal(e.get(L.F_X)[0], Xo)
    assert np.array_equal(e.get(L.F_X_STEP)[0], tun.step)
    assert 0.05 < out["accepted"].mean() < 0.99


def test_native_philox_sweep_matches_oracle_replay_of_device_draws():
    """Device RNG path: dump the Philox draws of the next sweep, replay them through the oracle."""
    L = _F()
    T, n, d = 6, 64, 2
    rng, X, Y = _synthetic(T, n, d, False, seed=77)
    C_ = 3
    e = _engine(T=T, n=n, d=d, n_chains=C_, tune=500, tune_interval=100)
    e.set_network(Y)
    e.set_hyper(tau_sq=2.0, sigma_sq=0.1)
    e.set(L.F_INTERCEPT, np.tile([[0.5, 0.0]], (C_, 1)))
    Xs = np.stack([X + 0.01 * c for c in range(C_)])
    e.set(L.F_X, Xs)
    e.set_tuner(0.1)
    e.set_rng(seed=1234, chain_offset=5)
    for s in range(3):
        eps, logu = e.debug_draws()
        assert abs(eps.mean()) < 0.1 and abs(eps.std() - 1) < 0.1
        assert np.all(logu < 0)
        e.sweep_latent()
        Xg = e.get(L.F_X)
        for c in range(C_):
            Xo = Xs[c].copy()
            tun = O.TunerState((T, n), 0.1, tune=500, tune_interval=100)
            O.sweep_latent(Xo, np.array([0.5]), tun, eps[c], logu[c], Y=Y, tau_sq=2.0, sigma_sq=0.1)
            assert np.array_equal(Xg[c], Xo)
        Xs = Xg
    # streams: chains differ, and a chain's stream depends only on (seed, chain id)
    assert not np.array_equal(eps[0], eps[1])
    e2 = _engine(T=T, n=n, d=d, n_chains=1)
    e2.set_rng(seed=1234, chain_offset=6, sweep_index=2)
    eps2, _ = e2.debug_draws()
    assert np.array_equal(eps2[0], eps[1])


def test_full_device_sweeps_are_deterministic_and_chain_independent():
    """Size-independent properties at cfg-2 shape: same seed -> same chain; a chain's trajectory
    does not depend on how many other chains share the launch."""
    L = _F()
    T, n, d, K = 9, 120, 2, 10
    rng, X, Y = _synthetic(T, n, d, False, seed=5)

    def run(C_, offset):
        e = _engine(T=T, n=n, d=d, n_chains=C_, K=K, mixture=True)
        e.set_network(Y)
        e.set(L.F_X, np.tile(X[None], (C_, 1, 1, 1)))
        e.set(L.F_INTERCEPT, np.tile([[0.5, 0.0]], (C_, 1)))
        r2 = np.random.RandomState(3)
        e.set(L.F_MU, np.tile(r2.randn(1, K, d), (C_, 1, 1)))
        e.set(L.F_SIGMA, np.tile(r2.gamma(2, 1, (1, K)), (C_, 1)))
        e.set(L.F_LAMBDA, np.full(C_, 0.8))
        w = r2.dirichlet(np.ones(K), size=(1, T, K))
        e.set(L.F_WEIGHTS, np.tile(w, (C_, 1, 1, 1)))
        e.set(L.F_Z, np.tile(r2.randint(0, K, (1, T, n)), (C_, 1, 1)))
        e.set_tuner(0.1)
        e.set_rng(seed=99, chain_offset=offset)
        e.run_sweeps(5)
        return e.get(L.F_X), e.get(L.F_Z), e.get(L.F_INTERCEPT), e.get(L.F_NK)
    Xa, za, ia, nka = run(4, 0)
    Xb, zb, ib, nkb = run(4, 0)
    assert np.array_equal(Xa, Xb) and np.array_equal(za, zb) and np.array_equal(ia, ib)
    Xc, zc, ic_, _ = run(1, 2)
    assert np.array_equal(Xc[0], Xa[2]) and np.array_equal(zc[0], za[2]) and np.array_equal(ic_[0], ia[2])
    assert not np.array_equal(Xa[0], Xa[1])
    assert np.all(nka.sum(axis=2) == n)           # every node has exactly one label per time step
    assert np.allclose(Xa.mean(axis=(1, 2)), 0, atol=1e-12)  # centred


@pytest.mark.parametrize("T,n,d,K", [
    (10, 200, 2, 12),   # thread-per-node, 64 threads per CTA
    (16, 45, 3, 7),     # K < 8 (numpy's serial sum)
    (5, 33, 2, 8),      # K = 8 (numpy's 8-accumulator sum)
    (9, 64, 2, 17),
    (4, 50, 1, 32),
    (3, 70, 2, 1),      # single component
    (12, 70, 3, 40),    # thread-per-node, 32 threads per CTA
    (20, 40, 2, 60),    # warp-per-node fallback (T*K too large for the thread kernel's stage)
    (7, 129, 2, 16),    # register kernel, KC = 16 exactly (two numpy 8-blocks)
    (6, 64, 2, 13),     # KC = 16 with three guarded components
    (8, 65, 3, 4),      # KC = 4, generic latent dimension
    (5, 100, 2, 5),     # KC = 8
    (2, 300, 2, 10),    # several node tiles per chain
    (1, 40, 2, 10)])    # a single time step: no backward pass
@pytest.mark.parametrize("stage", ["default", "global", "shared"])
def test_labels_vs_oracle_all_kernel_variants(T, n, d, K, stage, monkeypatch):
    """FFBS with recorded uniforms on larger random problems vs the oracle (exact labels): the
    register-resident kernel (default for K <= 16), the thread-per-node kernel with its stage in
    L2-resident global memory and in shared memory, the warp-per-node fallback."""
    if stage == "shared":
        monkeypatch.setenv("DLSM_FFBS_SMEM", "1")
    elif stage == "global":
        monkeypatch.setenv("DLSM_FFBS", "thread")
    L = _F()
    rng = np.random.RandomState(8)
    X = rng.randn(T, n, d)
    mu = rng.randn(K, d) * 1.5
    sigma = rng.gamma(3, 0.3, K)
    w = rng.dirichlet(np.ones(K) * 0.7, size=(T, K))
    U = rng.rand(n, T)
    e = _engine(T=T, n=n, d=d, K=K, mixture=True)
    e.set(L.F_X, X[None]); e.set(L.F_MU, mu[None]); e.set(L.F_SIGMA, sigma[None])
    e.set(L.F_LAMBDA, np.array([0.85])); e.set(L.F_WEIGHTS, w[None])
    e.sample_labels(U[None])
    z, nc, nk, _ = O.sample_labels_block(X, mu, sigma, 0.85, w, U)
    assert np.array_equal(e.get(L.F_Z)[0], z)
    assert np.array_equal(e.get(L.F_NCOUNT)[0], nc)
    assert np.array_equal(e.get(L.F_NK)[0], nk)


def test_fused_centring_equals_sweep_then_center():
    """dlsm_run_sweeps centres inside the sweep kernel when the chain lives in shared memory; the
    result must be bit-identical to sweep -> k_center (numpy summation order)."""
    L = _F()
    T, n, d = 7, 90, 2
    rng, X, Y = _synthetic(T, n, d, False, seed=21)

    def make():
        e = _engine(T=T, n=n, d=d, n_chains=3)
        e.set_network(Y)
        e.set(L.F_X, np.stack([X, X + 0.01, X - 0.02]))
        e.set(L.F_INTERCEPT, np.tile([[0.6, 0.0]], (3, 1)))
        e.set_tuner(0.1)
        e.set_rng(5)
        return e
    a, b = make(), make()
    a.run_sweeps(1, skip_intercepts=True)          # fused
    b.sweep_latent()                               # same Philox draws
    b.center()
    assert np.array_equal(a.get(L.F_X), b.get(L.F_X))


@pytest.mark.parametrize("mode", ["chain", "slice"])
@pytest.mark.parametrize("directed", [False, True])
@pytest.mark.parametrize("T,n,d", [(6, 90, 2), (3, 64, 3), (11, 33, 2), (2, 200, 2)])
def test_tracked_loglik_matches_full_kernel(T, n, d, directed, mode, monkeypatch):
    """The chain kernel accumulates the full-network log-likelihood of the state it leaves behind
    (dyad {i<j} taken from node j's update) and the intercept / radii MH keeps it current, so that
    k_full only evaluates proposals inside dlsm_run_sweeps.  The tracked value must equal a fresh
    full-network evaluation (network_likelihoods.py:26-33, directed_likelihoods_fast.pyx:185-205)
    to summation-order accuracy, and the chain must be the one the two-variant path produces."""
    L = _F()
    monkeypatch.setenv("DLSM_SWEEP_MODE", mode)   # "slice": one evaluation of the current state per sweep
    rng, X, Y = _synthetic(T, n, d, directed, seed=31)
    C_ = 3

    def make():
        e = _engine(T=T, n=n, d=d, n_chains=C_, is_directed=directed)
        e.set_network(Y)
        e.set(L.F_X, np.stack([X, X * 1.1, X - 0.03]))
        e.set(L.F_INTERCEPT, np.tile([[0.6, 0.3]], (C_, 1)))
        if directed:
            e.set(L.F_RADII, np.random.RandomState(2).dirichlet(np.ones(n) * 4, size=C_))
        e.set_tuner(0.15)
        e.set_rng(17)
        return e
    a = make()
    for sweeps in (1, 3):
        a.run_sweeps(sweeps)
        tracked = a.get(L.F_LOGLIK)
        fresh = a.loglik_full()
        assert np.all(np.isfinite(tracked))
        assert np.allclose(tracked, fresh, rtol=1e-11, atol=0)
    monkeypatch.setenv("DLSM_NO_LLCUR", "1")
    b = make()
    b.run_sweeps(4)
    # same Philox draws; acceptance tests differ only in the rounding of ll(current)
    assert np.array_equal(a.get(L.F_X), b.get(L.F_X))
    assert np.array_equal(a.get(L.F_INTERCEPT), b.get(L.F_INTERCEPT))
    if directed:
        assert np.array_equal(a.get(L.F_RADII), b.get(L.F_RADII))


def test_native_radii_sampler_targets_the_flat_dirichlet_prior():
    """With both intercepts at 0 the likelihood does not depend on the radii, so the device radii
    MH (gamma-variate Dirichlet proposal on Philox, Hastings correction of metropolis.py:57-82) must
    leave the flat Dirichlet prior invariant: checks proposal, correction and RNG together."""
    L = _F()
    T, n, d, C_ = 2, 5, 2, 4096
    rng = np.random.RandomState(0)
    e = _engine(T=T, n=n, d=d, n_chains=C_, is_directed=True, radii_tune=None)
    e.set_network(np.zeros((T, n, n)))
    e.set(L.F_X, rng.randn(C_, T, n, d))
    e.set(L.F_INTERCEPT, np.zeros((C_, 2)))
    e.set(L.F_RADII, rng.dirichlet(np.ones(n) * 30, size=C_))     # far from the target's spread
    e.set_tuner(0.1, 0.1, 60.0)   # broad proposals Dir(60 r), yet no gamma variate can underflow
    e.set_rng(3)
    acc = 0.0
    for _ in range(400):
        a, _ = e.sample_radii(want_stats=True)
        acc += a.mean()
    assert 0.15 < acc / 400 < 0.95
    r = e.get(L.F_RADII)
    assert np.allclose(r.sum(axis=1), 1.0, atol=1e-12) and np.all(r > 0)
    # Dirichlet(1,...,1): mean 1/n, variance (n-1)/(n^2 (n+1)), P(r_i < x) = 1 - (1-x)^(n-1)
    assert np.all(np.abs(r.mean(axis=0) - 1.0 / n) < 5 * np.sqrt((n - 1) / (n * n * (n + 1.0)) / C_))
    var = (n - 1) / (n * n * (n + 1.0))
    assert np.all(np.abs(r.var(axis=0) / var - 1.0) < 0.12)
    for x in (0.05, 0.2, 0.5):
        want = 1 - (1 - x) ** (n - 1)
        got = (r < x).mean(axis=0)
        assert np.all(np.abs(got - want) < 5 * np.sqrt(want * (1 - want) / C_) + 0.01), (x, got, want)


def test_long_row_many_chains_build_vs_oracle():
    """More chains than SMs with ~100 KB of positions each (cfg-4 shape): the (320, 2) instantiation
    of k_sweep with its unmasked interior trips.  Recorded draws, three of the 150 chains replayed by
    the oracle: identical decisions and positions."""
    L = _F()
    T, n, d, C_ = 10, 500, 2, 150
    rng, X, Y = _synthetic(T, n, d, False, seed=41, density=0.05)
    e = _engine(T=T, n=n, d=d, n_chains=C_, tune=2, tune_interval=1)
    e.set_network(Y)
    Xc = X[None] + 0.05 * rng.randn(C_, T, n, d)
    e.set(L.F_X, Xc)
    e.set(L.F_INTERCEPT, np.tile([[0.7, 0.0]], (C_, 1)))
    e.set_hyper(tau_sq=2.0, sigma_sq=0.1)
    e.set_tuner(0.03)
    check = (0, 77, 149)
    Xo = {c: Xc[c].copy() for c in check}
    tun = {c: O.TunerState((T, n), 0.03, tune=2, tune_interval=1) for c in check}
    for s in range(2):
        eps = rng.randn(C_, T, n, d)
        logu = np.log(rng.rand(C_, T, n))
        acc, _ = e.sweep_latent(eps, logu, want_stats=True)
        got = e.get(L.F_X)
        for c in check:
            out = O.sweep_latent(Xo[c], np.array([0.7]), tun[c], eps[c], logu[c], Y=Y, tau_sq=2.0, sigma_sq=0.1)
            assert np.array_equal(acc[c], out["accepted"]), (s, c)
            assert np.array_equal(got[c], Xo[c])
    assert 0.05 < acc.mean() < 0.95
