"""Benchmark / parity-test inputs of the five BASELINE.json configurations.

Input construction only (numpy on the host): `bench.py` and `tests/` build their networks and start
states here, so that the parity tests run on exactly the workloads the bench measures.

* cfg 1   the Sampson-monks network the reference bundles (T=3, n=18), taken from the golden
          fixture recorded from the reference (`tests/golden/lsm_undirected_monks.npz`).
* cfg 2-4 ``splitting_network``: a restatement of the reference's
          ``simple_splitting_dynamic_network`` (datasets/samples_generator.py:107-260 with
          ``network_from_dynamic_latent_space`` :81-104), drawing from ``numpy.random.RandomState``
          in the reference's order, so that the same seed gives the same ``Y`` and ``z``
          (tests/test_workloads.py checks this bit for bit against the live reference in the build
          container).  Unlike the reference it also returns the generating state (positions, radii,
          group means / spreads / transition weights) -- the chains start there (cfg 3).
* cfg 5   a sparse directed network that cannot exist as a dense tensor (200 GB): latent random
          walk, out-neighbours among the nearest nodes, stored as padded edge lists + control sets.
"""
import os
from math import ceil

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))

WORKLOADS = {
    # name: T, n, d, K, directed, case_control, chains (per GPU unless "total"), description
    "cfg1": dict(T=3, n=18, d=2, K=0, directed=False, chains=1184,
                 desc="DynamicNetworkLSM, Sampson monks (T=3, n=18, d=2)"),
    "cfg2": dict(T=9, n=120, d=2, K=10, directed=False, chains=1332,
                 desc="DynamicNetworkHDPLPCM, simple_splitting_dynamic_network(n=120, T=9), d=2, K=10"),
    "cfg3": dict(T=20, n=2000, d=2, K=0, directed=True, chains=1,
                 desc="directed DynamicNetworkLSM with radii, simple_splitting_dynamic_network(n=2000, "
                      "n_time_steps=19, is_directed=True) -> T=20, single chain"),
    "cfg4": dict(T=10, n=500, d=2, K=10, directed=False, chains=1024, total=True,
                 desc="DynamicNetworkHDPLPCM multi-chain, simple_splitting_dynamic_network(n=500, T=10), "
                      "d=2, K=10, 1024 chains in total"),
    "cfg5": dict(T=10, n=50000, d=2, K=0, directed=True, chains=8, case_control=True, n_control=100,
                 desc="directed LSM, case-control likelihood, sparse network (n=50000, T=10, d=2, "
                      "out-degree ~ Poisson(10), 100 controls), 8 chains per GPU"),
}


def expit(x):
    return 1.0 / (1.0 + np.exp(-x))


# ---------------------------------------------------------------------------------------------
# simple_splitting_dynamic_network, restated (same RandomState consumption as the reference)
# ---------------------------------------------------------------------------------------------
def _sticky_transitions(centres, sticky):
    """Row-stochastic transition matrix ~ 1/distance with `sticky` x the largest off-diagonal
    weight on the diagonal (samples_generator.py:141-149, :225-231)."""
    from sklearn.metrics import pairwise_distances
    with np.errstate(divide="ignore"):
        w = 1.0 / pairwise_distances(centres)
    k = np.arange(w.shape[0])
    w[k, k] = 0
    w[k, k] = sticky * np.max(w, axis=1)
    return w / w.sum(axis=1).reshape(-1, 1)


def _move(rng, z_prev, X_prev, labels, trans, centres, spreads, lmbda, from_labels, left_assoc):
    """One AR(1) step of the generating process: every group draws its members' next labels
    (one `choice` call per source group, in group order), then every destination group draws its
    members' positions (one `randn` call per group, in group order).  `left_assoc`: the reference
    adds noise + lmbda*mu + (1-lmbda)*x left to right after the split (:208-212, :243-247) and as
    noise + (lmbda*mu + (1-lmbda)*x) before it (:171-174); the roundings differ."""
    n = z_prev.shape[0]
    z = np.zeros(n, dtype=int)
    for g, src in enumerate(from_labels):
        m = z_prev == src
        z[m] = rng.choice(labels, p=trans[g, :], size=np.sum(m))
    X = np.zeros((n, 2), dtype=np.float64)
    for g, lab in enumerate(labels):
        m = z == lab
        noise = spreads[g] * rng.randn(np.sum(m), 2)
        if left_assoc:
            X[m, :] = noise + lmbda * centres[g] + (1 - lmbda) * X_prev[m, :]
        else:
            X[m, :] = noise + (lmbda * centres[g] + (1 - lmbda) * X_prev[m, :])
    return z, X


def splitting_network(n_nodes=120, n_time_steps=9, intercept=1.0, lmbda=0.8, sticky_const=20.0,
                      sigma_shape=6, sigma_scale=20, is_directed=False, random_state=42):
    """Two communities that split into four half-way through (the reference's
    ``simple_splitting_dynamic_network``).  Returns a dict with Y (T, n, n), z (T, n) and the
    generating state: X (T, n, 2), radii, intercept, mus (6, 2), sigmas (6,) [standard
    deviations], the two sticky transition matrices and the split matrix.

    T = 2 * ceil(n_time_steps / 2), i.e. 10 for n_time_steps = 9 (the reference
    returns one slice more than asked for odd inputs)."""
    from sklearn.metrics import pairwise_distances
    rng = np.random.RandomState(random_state) if not hasattr(random_state, "randn") else random_state
    half = ceil(n_time_steps / 2)
    mus = np.array([[-1.5, 0.0], [1.5, 0.0], [-1.5, 0.0], [1.5, 0.0], [0.0, 3.0], [0.0, -3.0]])
    if is_directed:
        mus /= 100.0
        sigma_scale, sigma_shape = 1e5, 13
    sigmas = np.sqrt(1.0 / rng.gamma(shape=sigma_shape, scale=sigma_scale, size=6))
    first, second = np.arange(2), 2 + np.arange(4)

    w0 = rng.dirichlet(np.repeat(10, 2))
    w_first = _sticky_transitions(mus[:2], sticky_const)
    z0 = rng.choice(first, p=w0, size=n_nodes)
    X0 = np.zeros((n_nodes, 2), dtype=np.float64)
    for g in first:
        m = z0 == g
        X0[m, :] = sigmas[g] * rng.randn(np.sum(m), 2) + mus[g]
    zs, Xs = [z0], [X0]
    for t in range(1, half):
        z, X = _move(rng, zs[-1], Xs[-1], first, w_first, mus[:2], sigmas[:2], lmbda, first, False)
        zs.append(z); Xs.append(X)

    # the split 2 -> 4: weights ~ 1/distance to the new centres, coincident centres get the row maximum
    with np.errstate(divide="ignore"):
        w_split = 1.0 / pairwise_distances(mus[:2], mus[2:])
    inf = ~np.isfinite(w_split)
    w_split[inf] = 0
    w_split[inf] = np.max(w_split, axis=1)
    w_split /= w_split.sum(axis=1).reshape(-1, 1)
    z, X = _move(rng, zs[-1], Xs[-1], second, w_split, mus[2:], sigmas[2:], lmbda, first, True)
    zs.append(z); Xs.append(X)

    w_second = _sticky_transitions(mus[2:], sticky_const)
    for t in range(half + 1, 2 * half):
        z, X = _move(rng, zs[-1], Xs[-1], second, w_second, mus[2:], sigmas[2:], lmbda, second, True)
        zs.append(z); Xs.append(X)
    X, z = np.stack(Xs, axis=0), np.vstack(zs)

    radii = None
    if is_directed:
        inv_norm = 1.0 / np.linalg.norm(X[0], axis=1)
        inv_norm /= np.max(inv_norm)
        radii = rng.dirichlet(100 * inv_norm)
        intercept = np.array([0.3, 0.7])
    Y = _draw_network(rng, X, intercept, radii)
    return dict(Y=Y, z=z, X=X, radii=radii, intercept=np.atleast_1d(np.asarray(intercept, np.float64)),
                mus=mus, sigmas=sigmas, w0=w0, w_first=w_first, w_split=w_split, w_second=w_second,
                lmbda=lmbda)


def _draw_network(rng, X, intercept, radii):
    """network_from_dynamic_latent_space (samples_generator.py:81-104): one `binomial(1, P_t)` call
    per time step; the undirected network keeps the strict upper triangle, mirrored."""
    from sklearn.metrics import euclidean_distances
    T, n, _ = X.shape
    Y = np.zeros((T, n, n), dtype=np.float64)
    for t in range(T):
        dist = euclidean_distances(X[t])
        if radii is None:
            eta = intercept - 1 * dist
            p = np.exp(eta) / (1 + np.exp(eta))
        else:
            # directed_network_probas (directed_likelihoods_fast.pyx:273-294), zero diagonal
            eta = intercept[0] * (1 - dist / radii[None, :]) + intercept[1] * (1 - dist / radii[:, None])
            p = 1 / (1 + np.exp(-eta))
            np.fill_diagonal(p, 0.0)
        Y[t] = rng.binomial(1, p).astype(int)
        if radii is None:
            Y[t] = np.triu(Y[t], 1)
            Y[t] += Y[t].T
    return Y


# ---------------------------------------------------------------------------------------------
# the five configurations
# ---------------------------------------------------------------------------------------------
def _monks():
    g = np.load(os.path.join(ROOT, "tests", "golden", "lsm_undirected_monks.npz"))
    return np.ascontiguousarray(g["Y"], dtype=np.float64), np.ascontiguousarray(g["X_in"][0])


def make_sparse_workload(name, seed=42):
    from scipy.spatial import cKDTree
    w = dict(WORKLOADS[name])
    T, n, d, nc = w["T"], w["n"], w["d"], w["n_control"]
    rng = np.random.RandomState(seed)
    X = np.empty((T, n, d))
    X[0] = 0.01 * rng.randn(n, d)
    for t in range(1, T):
        X[t] = X[t - 1] + 0.001 * rng.randn(n, d)
    out_lists = []
    for t in range(T):
        deg = np.minimum(rng.poisson(10, n), 30)
        _, nb = cKDTree(X[t]).query(X[t], k=41)
        pick = np.argsort(rng.rand(n, 40), axis=1)          # random subset of the 40 nearest
        out_lists.append((deg, np.take_along_axis(nb[:, 1:], pick, axis=1)))
    max_out = 30
    out_e = np.zeros((T, n, max_out), np.int32)
    degs = np.zeros((T, n, 2), np.int32)
    in_lists = []
    for t, (deg, cand) in enumerate(out_lists):
        keep = np.arange(max_out)[None, :] < deg[:, None]
        out_e[t] = np.where(keep, cand[:, :max_out], 0)
        degs[t, :, 1] = deg
        src = np.repeat(np.arange(n), deg)
        dst = cand[:, :max_out][keep]
        order = np.lexsort((src, dst))
        in_lists.append((dst[order], src[order]))
        degs[t, :, 0] = np.bincount(dst, minlength=n)
    max_in = int(degs[:, :, 0].max())
    in_e = np.zeros((T, n, max_in), np.int32)
    for t, (dst, src) in enumerate(in_lists):
        start = np.searchsorted(dst, dst, side="left")
        in_e[t, dst, np.arange(dst.size) - start] = src

    # control sets: uniform non-neighbours (collisions with neighbours/self are rare at this
    # sparsity and are redrawn; a residual collision only perturbs the estimator's weights)
    def controls(edges, deg_col):
        c = rng.randint(0, n, size=(T, n, nc)).astype(np.int32)
        for _ in range(2):
            bad = c == np.arange(n, dtype=np.int32)[None, :, None]
            for q in range(edges.shape[2]):
                bad |= (c == edges[:, :, q:q + 1]) & (q < degs[:, :, deg_col])[:, :, None]
            c[bad] = rng.randint(0, n, size=int(bad.sum()))
        return c
    radii = rng.dirichlet(np.ones(n) * 20.0)
    w.update(name=name, X=X, Y=None, radii=radii, intercept=np.array([0.3, 0.7]),
             step_X=0.0075 / n * 40, sigma_sq=1e-6, tau_sq=float(np.mean(X[0] * X[0])),
             degrees=degs, in_edges=in_e, out_edges=out_e, ctrl_in=controls(in_e, 0),
             ctrl_out=controls(out_e, 1), density=float(degs[:, :, 1].mean() / n),
             mean_deg=float(degs[:, :, 0].mean() + degs[:, :, 1].mean()))
    return w


_CACHE = {}


def make_workload(name, seed=42):
    """Network + start state of configuration `name` (SURVEY.md 8d's constructions and seeds)."""
    key = (name, seed)
    if key in _CACHE:
        return dict(_CACHE[key])
    if WORKLOADS[name].get("case_control"):
        w = make_sparse_workload(name, seed)
        _CACHE[key] = w
        return dict(w)
    w = dict(WORKLOADS[name])
    T, n, d, K = w["T"], w["n"], w["d"], w["K"]
    if name == "cfg1":
        Y, X = _monks()
        w.update(X=X - X.mean(axis=(0, 1)), Y=Y, radii=None, intercept=np.array([1.0]), step_X=0.1,
                 sigma_sq=0.1, tau_sq=2.0)
    else:
        # cfg 2 / cfg 4: n_time_steps = 9 gives 10 slices, cfg 2 keeps the first 9 (the paper
        # scripts do the same); cfg 3: n_time_steps = 19 gives T = 20
        g = splitting_network(n_nodes=n, n_time_steps=19 if w["directed"] else 9,
                              is_directed=w["directed"], random_state=seed)
        X = np.ascontiguousarray(g["X"][:T])
        Y = np.ascontiguousarray(g["Y"][:T])
        z = g["z"][:T]
        if w["directed"]:
            # the directed settings the reference itself uses (hdp_lpcm.py:59-69)
            w.update(radii=g["radii"], intercept=g["intercept"], step_X=0.0075, sigma_sq=0.001,
                     tau_sq=float(np.mean(X[0] * X[0])))
        else:
            w.update(radii=None, intercept=g["intercept"], step_X=0.1, sigma_sq=0.1, tau_sq=2.0)
        if K:
            mu = np.zeros((K, d)); mu[:6] = g["mus"]
            rs = np.random.RandomState(seed + 1)
            mu[6:] = 2.0 * rs.randn(K - 6, d)
            sigma = np.ones(K); sigma[:6] = g["sigmas"] ** 2          # variances
            wts = np.full((T, K, K), 0.02 / (K - 1)) + np.eye(K) * (0.98 - 0.02 / (K - 1))
            wts[0, 0] = 1.0 / K
            w.update(mu=mu, sigma=sigma, lmbda=float(g["lmbda"]), z=z.astype(np.int64), w=wts)
        w.update(X=X - X.mean(axis=(0, 1)), Y=Y)
    w.update(name=name, density=float(w["Y"].mean()))
    _CACHE[key] = w
    return dict(w)


def bytes_per_node_update(w):
    """Algorithmic bytes of one node-update, model M1."""
    n, d = w["n"], w["d"]
    if w.get("case_control"):
        return (w["mean_deg"] + 2 * w["n_control"]) * (4 + 8.0 * d + 8)
    if w["directed"]:
        return 2 * n / 8.0 + 8.0 * d * n + 8.0 * n
    return n / 8.0 + 8.0 * d * n


def chain_starts(w, chains, chain_offset):
    """Dispersed starts X0 + 0.1 * scale * N(0, I), one stream per global chain id."""
    disp = 0.1 * (1.0 / w["n"] if w["directed"] else 1.0)
    X = np.empty((chains,) + w["X"].shape)
    for c in range(chains):
        X[c] = w["X"] + disp * np.random.RandomState(100000 + chain_offset + c).randn(*w["X"].shape)
    return X


def build_engine(w, chains, device, chain_offset, seed=42):
    """An Engine holding `chains` chains of workload `w` at their dispersed starts."""
    from dynetlsm_b200 import _lib as L
    e = L.Engine(T=w["T"], n=w["n"], d=w["d"], n_chains=chains, K=w["K"], is_directed=w["directed"],
                 mixture=bool(w["K"]), device=device, tune=2500, tune_interval=100,
                 radii_tune=None, case_control=bool(w.get("case_control")))
    if w.get("case_control"):
        e.set_edge_lists(w["degrees"], w["in_edges"], w["out_edges"])
        e.set_controls(w["ctrl_in"], w["ctrl_out"])
    else:
        e.set_network(w["Y"])
    e.set(L.F_X, chain_starts(w, chains, chain_offset))
    ic = np.zeros((chains, 2)); ic[:, :w["intercept"].size] = w["intercept"]
    e.set(L.F_INTERCEPT, ic)
    e.set_hyper(tau_sq=w["tau_sq"], sigma_sq=w["sigma_sq"], intercept_prior=w["intercept"],
                intercept_variance_prior=2.0)
    if w["directed"]:
        e.set(L.F_RADII, np.tile(w["radii"][None], (chains, 1)))
    if w["K"]:
        e.set(L.F_MU, np.tile(w["mu"][None], (chains, 1, 1)))
        e.set(L.F_SIGMA, np.tile(w["sigma"][None], (chains, 1)))
        e.set(L.F_LAMBDA, np.full(chains, w["lmbda"]))
        e.set(L.F_WEIGHTS, np.tile(w["w"][None], (chains, 1, 1, 1)))
        e.set(L.F_Z, np.tile(w["z"][None], (chains, 1, 1)))
        # sticky HDP-HMM hyper state and priors (hdp_lpcm.py defaults, n-dependent 'auto' values)
        K, n, d = w["K"], w["n"], w["d"]
        mvp = (n ** (2.0 / d)) / 50.0
        a, a0 = 2.0, (4.0 ** 2 + 2) * 2
        b0, b_ = (a0 - 2) * mvp * 2, (a + 2) * mvp
        d0 = (4.0 ** 2 / b_) * 2
        e.set(L.F_BETA, np.full((chains, K), 1.0 / K))
        e.set(L.F_HYPER, np.tile(np.array([[1.0, 1.0, 1.0, 4.0, mvp, b_, 0, 0]]), (chains, 1)))
        e.set_hdp_prior(a, a0, b0, b_ * d0, d0, 0.9, 0.01, 1.0, 0.1, 1.0, 1.0, 5, 0.1, True, True)
    e.set_tuner(w["step_X"])
    e.set_rng(seed, chain_offset=chain_offset)
    return e
