"""CPU cost model of the reference sampler: the reference's own compiled Cython kernels
(oracle/_ref, built from /root/reference by oracle/build_ref.py) driven by a Python restatement
of the reference's per-node closure / Metropolis-object loop structure.

TEST INFRASTRUCTURE ONLY -- used by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm,
and cross-checked against the live reference and the golden fixtures in tests/test_ref_driver.py.
The reference's Python files cannot travel to the GPU box (only build outputs may live in
oracle/_ref), so the interpreter-level loop is restated here with the same call structure --
one Python closure call per log-posterior evaluation, one sampler object per (t, i), numpy
scalar arithmetic for the priors, sklearn distances, per-call index construction in the
full-network likelihood -- because that structure, not the arithmetic, is what the reference's
CPU time is made of.

Follows: sample_latent_positions.py:92-206, metropolis.py:40-136, sample_coefficients.py:12-121,
network_likelihoods.py:16-33, array_utils.py:4-8, latent_space.py:19-33, sample_labels.py:134-190.
"""
import glob
import importlib.machinery
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_REF = {}


def have_ref():
    return all(glob.glob(os.path.join(HERE, "_ref", m + ".*.so")) for m in
               ("static_network_fast", "directed_likelihoods_fast", "gaussian_likelihood_fast"))


def ref_kernels():
    """The reference's compiled Cython modules (oracle/_ref)."""
    if not _REF:
        for m in ("static_network_fast", "directed_likelihoods_fast", "gaussian_likelihood_fast"):
            path = glob.glob(os.path.join(HERE, "_ref", m + ".*.so"))
            if not path:
                raise ImportError("oracle/_ref/%s is not built (python oracle/build_ref.py)" % m)
            loader = importlib.machinery.ExtensionFileLoader(m, path[0])
            spec = importlib.util.spec_from_loader(m, loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
            _REF[m] = mod
    return _REF


class MHSampler(object):
    """State machine of metropolis.py:85-136 (one object per sampled block)."""
    RW = ((0.001, 0.1), (0.05, 0.5), (0.25, 0.9))
    RW_HI = ((0.95, 10.0), (0.75, 2.0), (0.4, 1.1))

    def __init__(self, step_size=0.1, tune=500, tune_interval=100, dirichlet=False):
        self.step_size, self.tune, self.tune_interval = step_size, tune, tune_interval
        self.dirichlet = dirichlet
        self.steps_until_tune = tune_interval
        self.n_accepted = 0
        self.n_steps = 0

    def _retune(self, rate):
        lo = [(a, 1.0 / f if self.dirichlet else f) for a, f in self.RW]
        hi = [(a, 1.0 / f if self.dirichlet else f) for a, f in self.RW_HI]
        if self.dirichlet:  # metropolis.py:23-37 uses 10, 2, 1.1 / 0.1, 0.5, 0.9
            lo = [(0.001, 10.0), (0.05, 2), (0.25, 1.1)]
            hi = [(0.95, 0.1), (0.75, 0.5), (0.4, 0.9)]
        for thr, f in lo:
            if rate < thr:
                self.step_size *= f
                return
        for thr, f in hi:
            if rate > thr:
                self.step_size *= f
                return

    def step(self, x, logp, rng):
        if self.dirichlet:
            x_new, accepted = dirichlet_mh(x, logp, self.step_size, rng)
        else:
            x_new, accepted = random_walk_mh(x, logp, self.step_size, rng)
        self.n_accepted += accepted
        self.n_steps += 1
        if self.tune is not None:
            if self.n_steps < self.tune and self.steps_until_tune == 0:
                self._retune(self.n_accepted / self.tune_interval)
                self.n_accepted = 0
                self.steps_until_tune = self.tune_interval
            else:
                self.steps_until_tune -= 1
        return x_new


def random_walk_mh(x0, logp, step_size, rng):
    x = x0 + step_size * rng.randn(x0.shape[0])
    ratio = logp(x) - logp(x0)
    u = rng.rand()
    if np.log(u) >= ratio:
        return x0, 0
    return x, 1


def dirichlet_mh(x0, logp, step_size, rng, reg=1e-5):
    import scipy.stats as stats
    x = rng.dirichlet(step_size * x0)
    if np.any(x == 0.):
        x += reg
        x /= np.sum(x)
    ratio = logp(x) - logp(x0)
    ratio += (stats.dirichlet.logpdf(x0, step_size * x) - stats.dirichlet.logpdf(x, step_size * x0))
    u = rng.rand()
    if np.log(u) >= ratio:
        return x0, 0
    return x, 1


def latent_sweep(Y, X, intercept, samplers, rng, radii=None, is_directed=False, cc=None,
                 tau_sq=2.0, sigma_sq=0.1, mixture=None):
    """sample_latent_positions.py:92-146 (mixture=None) / :149-206 (mixture=(mu, sigma, lmbda, z))."""
    k = ref_kernels()
    k1 = k["static_network_fast"].partial_loglikelihood
    k2 = k["directed_likelihoods_fast"].directed_partial_loglikelihood
    k3 = k["directed_likelihoods_fast"].approx_directed_partial_loglikelihood
    T, n, _ = X.shape
    for t in range(T):
        for j in range(n):
            def logp(x):
                X[t, j] = x
                if is_directed:
                    if cc is not None:
                        ll = k3(X[t], radii=radii, in_edges=cc["in_edges"][t],
                                out_edges=cc["out_edges"][t], degree=cc["degrees"][t],
                                control_nodes_in=cc["ctrl_in"][t],
                                control_nodes_out=cc["ctrl_out"][t], intercept_in=intercept[0],
                                intercept_out=intercept[1], node_id=j, squared=False)
                    else:
                        ll = k2(Y[t], X[t], radii=radii, intercept_in=intercept[0],
                                intercept_out=intercept[1], node_id=j, squared=False)
                else:
                    ll = k1(Y[t], X[t], intercept, j, squared=False)
                if mixture is None:
                    if t == 0:
                        ll -= 0.5 * np.sum(x * x) / tau_sq
                    else:
                        dx = x - X[t - 1, j]
                        ll -= 0.5 * np.sum(dx * dx) / sigma_sq
                    if t < T - 1:
                        dx = X[t + 1, j] - x
                        ll -= 0.5 * np.sum(dx * dx) / sigma_sq
                else:
                    mu, sigma, lmbda, z = mixture
                    if t == 0:
                        dx = x - mu[z[t, j]]
                    else:
                        dx = x - (1 - lmbda) * X[t - 1, j] - lmbda * mu[z[t, j]]
                    ll -= 0.5 * np.sum(dx * dx) / sigma[z[t, j]]
                    if t < T - 1:
                        dx = X[t + 1, j] - (1 - lmbda) * x - lmbda * mu[z[t + 1, j]]
                        ll -= 0.5 * np.sum(dx * dx) / sigma[z[t + 1, j]]
                return ll
            X[t, j] = samplers[t][j].step(X[t, j].copy(), logp, rng)
    return X


def distances(X):
    """latent_space.py:19-33 (sklearn per time slice)."""
    from sklearn.metrics import euclidean_distances
    out = np.empty((X.shape[0], X.shape[1], X.shape[1]))
    for t in range(X.shape[0]):
        out[t] = euclidean_distances(X[t], squared=False)
    return out


def _triu3(Y):
    # array_utils.py:4-8: the index arrays are rebuilt on every call
    return np.nonzero(~np.stack([np.tri(Y.shape[1], Y.shape[2], k=0, dtype=bool)
                                 for _ in range(Y.shape[0])]))


def full_loglik(Y, X, intercept, dist, radii=None, is_directed=False, cc=None):
    """network_likelihoods.py:16-33 and the case-control branch of sample_coefficients.py:24-35."""
    k = ref_kernels()["directed_likelihoods_fast"]
    if is_directed:
        if cc is not None:
            return k.approx_directed_network_loglikelihood(
                X=X, radii=radii, in_edges=cc["in_edges"], out_edges=cc["out_edges"],
                degree=cc["degrees"], control_nodes=cc["ctrl_out"], intercept_in=intercept[0],
                intercept_out=intercept[1], squared=False)
        return k.directed_network_loglikelihood_fast(Y, dist, radii, intercept[0], intercept[1])
    idx = _triu3(dist)
    eta = intercept - dist[idx]
    return np.sum(Y[idx] * eta - np.log(1 + np.exp(eta)))


def intercept_step(Y, X, intercepts, prior_mean, prior_var, samplers, rng, dist, radii=None,
                   is_directed=False, cc=None):
    """sample_coefficients.py:12-88."""
    if is_directed:
        for i in (0, 1):
            def logp(x):
                b = np.array([x[0], intercepts[1]]) if i == 0 else np.array([intercepts[0], x[0]])
                ll = full_loglik(Y, X, b, dist, radii, True, cc)
                return ll - ((x[0] - prior_mean[i]) ** 2 / (2 * prior_var))
            intercepts[i] = samplers[i].step(np.array([intercepts[i]]), logp, rng)[0]
        return intercepts

    def logp(x):
        ll = full_loglik(Y, X, x, dist)
        return ll - ((x - prior_mean) ** 2 / (2 * prior_var))
    return samplers[0].step(intercepts, logp, rng)


def radii_step(Y, X, intercepts, radii, sampler, rng, dist, cc=None):
    """sample_coefficients.py:91-121."""
    def logp(r):
        return full_loglik(Y, X, intercepts, dist, r, True, cc)
    return sampler.step(radii, logp, rng)


def labels_block(X, mu, sigma, lmbda, w, rng):
    """sample_labels.py:134-190 with gaussian_likelihood_fast.pyx:30-54 for the emissions."""
    gl = ref_kernels()["gaussian_likelihood_fast"].compute_gaussian_likelihood
    T, n, _ = X.shape
    K = sigma.shape[0]
    bwd = np.ones((T, K))
    pm = np.zeros((T, K))
    cnt = np.zeros((T, K, K))
    nk = np.zeros((T, K), dtype=np.int64)
    resp = np.zeros((T, n, K), dtype=np.int64)
    z = np.zeros((T, n), dtype=np.int64)
    lm = float(np.ravel(lmbda)[0])
    for i in range(n):
        lik = gl(X[:, i], mu, sigma, lm, normalize=False)
        for t in range(T - 1, 0, -1):
            pm[t] = lik[t] * bwd[t]
            bwd[t - 1] = np.dot(w[t], pm[t].reshape(-1, 1)).ravel()
            bwd[t - 1] /= np.sum(bwd[t - 1])
        pm[0] = lik[0] * bwd[0]
        for t in range(T):
            p = w[0, 0] * pm[0] if t == 0 else w[t, z[t - 1, i]] * pm[t]
            cdf = np.cumsum(p)
            u = rng.uniform(0, cdf[-1])
            z[t, i] = np.sum(u > cdf)
            if t == 0:
                cnt[0, 0, z[t, i]] += 1
            else:
                cnt[t, z[t - 1, i], z[t, i]] += 1
            resp[t, i, z[t, i]] = 1
            nk[t, z[t, i]] += 1
    return z, cnt, nk, resp


def hot_path_sweep(state, rng):
    """One pass of the hot path as the estimator loops run it (lsm.py:483-523 /
    hdp_lpcm.py:840-878): latent sweep -> centre -> distance cache -> intercept MH ->
    [radii MH] -> [label FFBS].  ``state`` is a dict; updated in place."""
    s = state
    X = s["X"]
    mix = (s["mu"], s["sigma"], s["lmbda"], s["z"]) if s.get("mixture") else None
    X = latent_sweep(s["Y"], X, s["intercept"], s["samplers"], rng, radii=s.get("radii"),
                     is_directed=s["is_directed"], cc=s.get("cc"), tau_sq=s.get("tau_sq", 2.0),
                     sigma_sq=s.get("sigma_sq", 0.1), mixture=mix)
    X -= np.mean(X, axis=(0, 1))
    dist = None if s.get("cc") is not None else distances(X)
    s["intercept"] = intercept_step(s["Y"], X, s["intercept"], s["prior_mean"], s["prior_var"],
                                    s["isamplers"], rng, dist, radii=s.get("radii"),
                                    is_directed=s["is_directed"], cc=s.get("cc"))
    if s["is_directed"]:
        s["radii"] = radii_step(s["Y"], X, s["intercept"], s["radii"], s["rsampler"], rng, dist,
                                cc=s.get("cc"))
    if mix is not None:
        s["z"], s["ncount"], s["nk"], _ = labels_block(X, s["mu"], s["sigma"], s["lmbda"], s["w"], rng)
    s["X"] = X
    return s


def make_state(Y, X, intercept, is_directed=False, radii=None, mixture=None, cc=None,
               step_X=0.1, tune=500, tune_interval=100, tau_sq=2.0, sigma_sq=0.1):
    T, n, _ = X.shape
    s = dict(Y=Y, X=X.copy(), intercept=np.array(intercept, dtype=np.float64), is_directed=is_directed,
             tau_sq=tau_sq, sigma_sq=sigma_sq, cc=cc,
             prior_mean=np.array(intercept, dtype=np.float64).copy(), prior_var=2.0,
             samplers=[[MHSampler(step_X, tune, tune_interval) for _ in range(n)] for _ in range(T)],
             isamplers=[MHSampler(0.1, tune, 100) for _ in range(2 if is_directed else 1)])
    if is_directed:
        s["radii"] = radii.copy()
        s["rsampler"] = MHSampler(175000, None, 100, dirichlet=True)
    if mixture is not None:
        s["mixture"] = True
        s["mu"], s["sigma"], s["lmbda"], s["z"], s["w"] = mixture
    return s
