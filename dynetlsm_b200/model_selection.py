"""Point-estimate selection on the traces of ``DynamicNetworkHDPLPCM``: the
reference's ``selection_type`` in {'vi', 'bic', 'map'} and the fitted attributes that go with it
(``bic_``, ``models_``, ``counts_``, ``posterior_group_ids_``, ``posterior_group_counts_``).

Same definitions as the reference (model_selection/posterior_vi.py:27-82,
model_selection/approx_bic.py:29-162, label_utils.py:65-82, hdp_lpcm.py:1089-1167) -- including
its parameter counts -- written for whole traces at once: the expected-VI scan is a few GEMMs per
time step instead of a Python loop over samples, and the network log-likelihoods the BIC and the VI
tie-break need come from the device kernels (K4/K5 through ``Engine.loglik_full``).
"""
import numpy as np

__all__ = ["expected_vi_trace", "minimize_posterior_expected_vi", "cluster_counts", "cluster_counts_t",
           "latent_marginal_loglikelihood", "select_bic", "posterior_group_counts", "MixtureEstimate"]


class MixtureEstimate(object):
    """What ``models_`` holds per model size (approx_bic.py:12-26)."""

    def __init__(self, beta, init_weights, trans_weights, X, mu, sigma, lmbda, z, intercept, radii=None):
        self.beta, self.init_weights, self.trans_weights = beta, init_weights, trans_weights
        self.X, self.mu, self.sigma, self.lmbda = X, mu, sigma, lmbda
        self.z, self.intercept, self.radii = z, intercept, radii


# ---------------------------------------------------------------------------------------------
# posterior expected variation of information (Wade & Ghahramani lower bound)
# ---------------------------------------------------------------------------------------------
def expected_vi_trace(zs, cooc, n_groups):
    """Time-averaged posterior-expected VI of every label configuration of a trace.

    zs (S, T, n) int labels in [0, n_groups); cooc (T, n, n) co-clustering probabilities.
    Returns (S,): for sample s, mean over t of
        [sum_k n_k log2 n_k - 2 sum_i log2 sum_j cooc[t,i,j] 1{z_j = z_i} + sum_i log2 sum_j cooc[t,i,j]] / n
    (posterior_vi.py:27-54)."""
    zs = np.asarray(zs)
    S, T, n = zs.shape
    out = np.zeros(S)
    rows = np.arange(n)
    for t in range(T):
        P = np.asarray(cooc[t], dtype=np.float64)
        row_tot = np.log2(P.sum(axis=1)).sum()
        for s0 in range(0, S, 256):                       # bounded (chunk, n, K) temporaries
            z = zs[s0:s0 + 256, t]                        # (c, n)
            resp = np.zeros(z.shape + (n_groups,))
            np.put_along_axis(resp, z[:, :, None], 1.0, axis=2)
            nk = resp.sum(axis=1)                         # (c, K)
            with np.errstate(divide="ignore", invalid="ignore"):
                size_term = np.where(nk > 0, nk * np.log2(nk), 0.0).sum(axis=1)
            within = np.einsum("ij,cjk->cik", P, resp)    # mass of i's row inside each cluster
            own = np.take_along_axis(within, z[:, :, None], axis=2)[:, :, 0]
            out[s0:s0 + 256] += (size_term - 2.0 * np.log2(own).sum(axis=1) + row_tot) / n
    return out / T


def minimize_posterior_expected_vi(model, loglik_fn=None):
    """Index (into the stored samples) of the post-burn-in label configuration with the smallest
    expected VI; ties go to the sample with the highest network log-likelihood
    (posterior_vi.py:56-82)."""
    nb = model.n_burn_
    vis = expected_vi_trace(model.zs_[nb:], model.cooccurrence_probas_, model.n_components)
    ties = np.where(vis == vis.min())[0]
    if ties.shape[0] == 1:
        return int(nb + ties[0])
    loglik_fn = loglik_fn or _device_loglik(model)
    best, best_ll = None, -np.inf
    for m in ties:
        ll = loglik_fn(nb + int(m))
        if ll > best_ll:
            best, best_ll = nb + int(m), ll
    return int(best)


# ---------------------------------------------------------------------------------------------
# model sizes seen by the chain, approximate BIC per size
# ---------------------------------------------------------------------------------------------
def cluster_counts(zs, n_groups):
    """Occupied components per stored sample, over all time steps (approx_bic.py:44-56)."""
    zs = np.asarray(zs)
    present = np.zeros((zs.shape[0], n_groups), dtype=bool)
    np.put_along_axis(present, zs.reshape(zs.shape[0], -1), True, axis=1)
    return present.sum(axis=1)


def cluster_counts_t(zs, n_groups):
    """Occupied components per time step and stored sample, (T, S) (approx_bic.py:29-41)."""
    zs = np.asarray(zs)
    S, T, n = zs.shape
    present = np.zeros((S, T, n_groups), dtype=bool)
    np.put_along_axis(present, zs, True, axis=2)
    return present.sum(axis=2).T


def posterior_group_counts(counts_t_row):
    """(sizes, frequencies) of the number of groups at one time step (label_utils.py:75-82)."""
    freq = np.bincount(counts_t_row)
    idx = np.where(freq != 0)[0]
    return idx, freq[idx]


def _emission_densities(X, mu, sigma, lmbda):
    """Raw Gaussian densities L[t, i, k] (gaussian_likelihood_fast.pyx:17-54, normalize=False)."""
    T, n, d = X.shape
    lm = float(np.ravel(lmbda)[0])
    mean = np.empty((T, n, mu.shape[0], d))
    mean[0] = mu[None]
    mean[1:] = lm * mu[None, None] + (1.0 - lm) * X[:-1, :, None, :]
    sq = ((X[:, :, None, :] - mean) ** 2).sum(axis=3)
    return np.exp(-0.5 * d * np.log(2.0 * np.pi * sigma)[None, None] - 0.5 * sq / sigma[None, None])


def latent_marginal_loglikelihood(X, init_w, trans_w, mu, sigma, lmbda):
    """log p(X | mixture) with the labels summed out: the HMM forward recursion of every node
    (approx_bic.py:59-76), all nodes at once."""
    L = _emission_densities(X, mu, sigma, lmbda)
    fwd = init_w[None] * L[0]
    c = fwd.sum(axis=1)
    ll = np.log(c).sum()
    fwd /= c[:, None]
    for t in range(1, X.shape[0]):
        fwd = L[t] * (fwd @ trans_w[t])
        c = fwd.sum(axis=1)
        ll += np.log(c).sum()
        fwd /= c[:, None]
    return float(ll)


def renormalized(model, idx):
    """The sample's mixture restricted to its occupied components (label_utils.py:10-37)."""
    T, n = model.Y_fit_.shape[:2]
    active, zi = np.unique(model.zs_[idx].ravel(), return_inverse=True)
    beta = model.betas_[idx, active] / model.betas_[idx, active].sum()
    w = model.weights_[idx]
    init_w = w[0, 0, active] / w[0, 0, active].sum()
    trans_w = np.zeros((T, active.size, active.size))
    for t in range(1, T):
        sub = w[t, active][:, active]
        trans_w[t] = sub / sub.sum(axis=1).reshape(-1, 1)
    return zi.reshape(T, n), beta, init_w, trans_w, model.mus_[idx, active], model.sigmas_[idx, active]


def select_bic(model, loglik_fn=None):
    """(bic (M, 4) rows [k, BIC, network loglik, MAP sample id], models, counts): for every model
    size k visited after burn-in, the MAP sample of that size and its approximate BIC
    (approx_bic.py:79-162, parameter counts as the reference writes them)."""
    T, n, _ = model.Y_fit_.shape
    nb = model.n_burn_
    loglik_fn = loglik_fn or _device_loglik(model)
    counts = cluster_counts(model.zs_[nb:], model.n_components)
    Y = model.Y_fit_
    rows, models = [], []
    for k in np.unique(counts):
        lp = np.where(counts == k, model.logps_[nb:], -np.inf)
        idx = nb + int(np.argmax(lp))
        z, beta, init_w, trans_w, mu, sigma = renormalized(model, idx)
        X, lmbda = model.Xs_[idx], model.lambdas_[idx]
        radii = model.radiis_[idx] if model.is_directed else None
        ll = float(loglik_fn(idx))
        bic = -2.0 * ll
        if model.is_directed:
            off = Y.sum() - np.einsum("tkk->", Y)
            bic += (2 + n) * np.log(off)
        else:
            bic += np.log(0.5 * (Y.sum() - np.einsum("tkk->", Y)))
        bic -= 2.0 * latent_marginal_loglikelihood(X, init_w, trans_w, mu, sigma, lmbda)
        n_params = (model.n_features + 1) * k + (k - 1) + (k - 1) + (T - 1) * k * (k - 1)
        bic += n_params * np.log(n * T)
        rows.append([k, bic, ll, idx])
        models.append(MixtureEstimate(beta, init_w, trans_w, X, mu, sigma, lmbda, model.zs_[idx],
                                      model.intercepts_[idx], radii))
    return np.array(rows), models, counts


def _device_loglik(model):
    """Network log-likelihood log p(Y | X, intercept[, radii]) of stored samples on the device
    (K4 / K5 / K6 through Engine.loglik_full); the engine lives for the duration of the selection."""
    from . import _lib as L
    T, n, d = model.Xs_.shape[1:]
    cc = getattr(model, "case_control_sampler_", None)
    e = L.Engine(T=T, n=n, d=d, n_chains=1, is_directed=model.is_directed, case_control=cc is not None,
                 device=model.device)
    if cc is None:
        e.set_network(model.Y_fit_)
    else:
        e.set_edge_lists(cc.degrees_, cc.in_edges_, cc.out_edges_)
        e.set_controls(cc.control_nodes_in_, cc.control_nodes_out_)

    def fn(idx):
        e.set(L.F_X, model.Xs_[idx][None])
        ic = np.zeros((1, 2))
        ic[0, :np.size(model.intercepts_[idx])] = np.ravel(model.intercepts_[idx])
        e.set(L.F_INTERCEPT, ic)
        if model.is_directed:
            e.set(L.F_RADII, model.radiis_[idx][None])
        return float(e.loglik_full()[0])
    fn.engine = e
    return fn
