"""Host-side conjugate / auxiliary-variable updates of the sticky HDP-HMM mixture that surround the
device hot path in ``DynamicNetworkHDPLPCM`` ("next" for a device
implementation; O(T K^2 + nT) numpy work per sweep today).

Restates, with the reference's random-number consumption order so that a replay-mode fit follows
the reference chain draw for draw:
  sample_tables / sample_mbar         sample_auxillary.py:6-50   (Fox et al. table counts, overrides)
  beta, w0, w[t,k] Dirichlet draws    hdp_lpcm.py:887-898
  mu_k, sigma_k, lambda               hdp_lpcm.py:901-954
  tau^2, b hyper-priors               hdp_lpcm.py:957-972
  gamma, alpha_init (Escobar-West)    sample_concentration.py:6-21, hdp_lpcm.py:977-995
  alpha + kappa, rho                  hdp_lpcm.py:998-1023
and the joint log-posterior of hdp_lpcm.py:1188-1280 (network term supplied by the device).
"""
import numpy as np
import scipy.stats as stats
from scipy.stats import truncnorm

TINY = np.finfo("float64").tiny

__all__ = ["HDPHyper", "conjugate_updates", "mixture_updates", "hdp_log_prior", "mixture_log_prior"]


class HDPHyper(object):
    """Mutable hyper-parameter state the reference keeps on the estimator object."""

    def __init__(self, gamma, alpha_init, alpha, kappa, mean_variance_prior, b, a, a0, b0, c0, d0,
                 lambda_prior, lambda_variance_prior, gamma_prior_shape, gamma_prior_rate,
                 alpha_init_shape, alpha_init_rate, alpha_kappa_shape, alpha_kappa_rate,
                 resample_mean_variance, resample_b):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def _clipped_dirichlet(rng, alphas):
    if np.any(alphas <= 0.):
        alphas = np.clip(alphas, a_min=TINY, a_max=None)
    return rng.dirichlet(alphas)


def _dirichlet_logpdf(x, alphas):
    if np.any(alphas <= 0.):
        alphas = np.clip(alphas, a_min=TINY, a_max=None)
    if np.any(x <= 0):
        x = np.clip(x, a_min=TINY, a_max=None)
    return stats.dirichlet.logpdf(x, alphas)


def _concentration(rng, alpha, n_clusters, n_samples, shape, rate):
    eta = rng.beta(alpha + 1, n_samples)
    m_shape = shape + n_clusters - 1
    m_scale = rate - np.log(eta)
    odds = (m_shape / m_scale) * (1 / n_samples)
    if rng.binomial(1, odds / (1 + odds)):
        m_shape = m_shape + 1
    return rng.gamma(shape=m_shape, scale=1. / m_scale)


def _tables(rng, n, beta, alpha_init, alpha, kappa):
    T, K, _ = n.shape
    m = np.zeros((T, K, K), dtype=np.int64)
    p0 = alpha_init * beta
    for k in range(K):
        m[0, 0, k] = np.sum(rng.binomial(1, p0[k] / (p0[k] + np.arange(n[0, 0, k]))))
    p = alpha * beta + kappa * np.eye(K)
    for t in range(1, T):
        for j in range(K):
            for k in range(K):
                m[t, j, k] = np.sum(rng.binomial(1, p[j, k] / (p[j, k] + np.arange(n[t, j, k]))))
    return m


def _mbar(rng, m, beta, kappa, alpha):
    T, K, _ = m.shape
    w = np.zeros((T - 1, K))
    rho = kappa / (alpha + kappa)
    for t in range(T - 1):
        for j in range(K):
            w[t, j] = rng.binomial(m[t + 1, j, j], rho / (rho + beta[j] * (1 - rho)))
    m_bar = np.zeros((T - 1, K, K))
    for t in range(T - 1):
        m_bar[t] = m[t + 1] - np.diag(w[t])
    return np.sum(m_bar, axis=(0, 1)) + m[0, 0], w


def mixture_updates(rng, hp, X, z, nk, mu, sigma, lmbda):
    """Cluster means, variances, blending coefficient and the two scale hyper-priors: the block
    the HDP sampler (hdp_lpcm.py:899-977) and the finite mixture (lpcm.py:582-656) share, in the
    reference's draw order.  mu, sigma are updated in place; returns lmbda."""
    T, n_nodes, d = X.shape
    K = sigma.shape[0]
    # members of (t, k) gathered once: x_t - (1 - lambda) x_{t-1} serves the mean and the variance draw
    # (same operands in the same order as the reference's masked expressions)
    resid = [[None] * K for _ in range(T)]
    for t in range(T):
        zt = z[t]
        for k in range(K):
            if nk[t, k] > 0:
                idx = np.flatnonzero(zt == k)
                resid[t][k] = X[t][idx] if t == 0 else X[t][idx] - (1 - lmbda) * X[t - 1][idx]
    # cluster means
    for k in range(K):
        prec = 1 / hp.mean_variance_prior
        acc = np.zeros(d)
        for t in range(T):
            if nk[t, k] > 0:
                if t == 0:
                    prec += nk[0, k] / sigma[k]
                    acc += (1 / sigma[k]) * np.sum(resid[t][k], axis=0)
                else:
                    prec += (lmbda ** 2 / sigma[k]) * nk[t, k]
                    acc += (lmbda / sigma[k]) * np.sum(resid[t][k], axis=0)
        var = 1 / prec
        acc *= var
        mu[k] = rng.multivariate_normal(mean=acc, cov=var * np.eye(d))
    # cluster variances
    for k in range(K):
        shape = 0.5 * (np.sum(nk[:, k]) * d + hp.a)
        rate = 0.5 * hp.b
        for t in range(T):
            if nk[t, k] > 0:
                if t == 0:
                    rate += 0.5 * np.sum((resid[t][k] - mu[k]) ** 2)
                else:
                    rate += 0.5 * np.sum((resid[t][k] - lmbda * mu[k]) ** 2)
        sigma[k] = 1. / rng.gamma(shape=shape, scale=1. / rate)
    # blending coefficient lambda ~ truncated normal on (0, 1)
    ml = 0.0
    sl = 1.0 / hp.lambda_variance_prior
    for t in range(1, T):
        sg = sigma[z[t]].reshape(-1, 1)
        dm = (mu[z[t]] - X[t - 1]) / sg
        ml += np.sum(dm * (X[t] - X[t - 1]))
        dm = (mu[z[t]] - X[t - 1]) / np.sqrt(sg)
        sl += np.sum(dm ** 2)
    sl = 1. / sl
    ml += hp.lambda_prior / hp.lambda_variance_prior
    ml *= sl
    std = np.sqrt(sl)
    lmbda = truncnorm.rvs((0 - ml) / std, (1 - ml) / std, size=1, loc=ml, scale=std,
                          random_state=rng)
    # hyper-priors on the scale of the latent space and of the clusters
    if hp.resample_mean_variance:
        bb = 0.5 * hp.b0
        for k in range(K):
            bb += 0.5 * np.sum(mu[k] ** 2)
        hp.mean_variance_prior = 1 / rng.gamma(shape=0.5 * (hp.a0 + K), scale=1. / bb, size=1)
    if hp.resample_b:
        sc = 0.5 * hp.d0
        for k in range(K):
            sc += 0.5 * (1. / sigma[k])
        hp.b = rng.gamma(shape=0.5 * (hp.c0 + K * hp.a), scale=1. / sc)
    return lmbda


def conjugate_updates(rng, hp, X, z, n, nk, mu, sigma, lmbda, beta, weights):
    """Everything between the label draw and the stored sample of one sweep
    (hdp_lpcm.py:881-1023).  ``n`` (T,K,K) transition counts, ``nk`` (T,K) occupancies come from
    the device label kernel.  mu, sigma, weights are updated in place; returns (beta, lmbda)."""
    T, n_nodes, d = X.shape
    K = sigma.shape[0]
    m = _tables(rng, n, beta, hp.alpha_init, hp.alpha, hp.kappa)
    m_bar, w_over = _mbar(rng, m, beta, kappa=hp.kappa, alpha=hp.alpha)

    beta = rng.dirichlet((hp.gamma / K) + m_bar)
    weights[0, 0] = _clipped_dirichlet(rng, hp.alpha_init * beta + nk[0])
    base = hp.alpha * beta + hp.kappa * np.eye(K)
    for t in range(1, T):
        for k in range(K):
            weights[t, k] = _clipped_dirichlet(rng, base[k] + n[t, k])

    lmbda = mixture_updates(rng, hp, X, z, nk, mu, sigma, lmbda)
    # concentration parameters
    hp.gamma = _concentration(rng, hp.gamma, np.sum(m_bar > 0), np.sum(m_bar),
                              hp.gamma_prior_shape, hp.gamma_prior_rate)
    hp.alpha_init = _concentration(rng, hp.alpha_init, np.sum(m[0, 0]), n_nodes,
                                   hp.alpha_init_shape, hp.alpha_init_rate)
    ak = hp.alpha + hp.kappa
    n_dot = np.sum(n[1:], axis=2)
    ok = n_dot > 0
    nd = n_dot[ok]
    s = rng.binomial(1, p=(nd / (nd + ak)))
    r = rng.beta(ak + 1, nd)
    shape = hp.alpha_kappa_shape + np.sum(m[1:], axis=2)[ok].sum() - np.sum(s)
    rate = hp.alpha_kappa_rate - np.sum(np.log(r))
    ak = rng.gamma(shape=shape, scale=1. / rate)
    n_succ = np.sum(w_over)
    rho = rng.beta(a=8 + n_succ, b=np.sum(m[1:]) - n_succ + 2)
    hp.kappa = ak * rho
    hp.alpha = ak - hp.kappa
    return beta, lmbda


def hdp_log_prior(hp, K, X, intercept, intercept_prior, intercept_variance_prior, mu, sigma, z,
                  weights, beta, lmbda, radii=None):
    """All terms of hdp_lpcm.py:1188-1280 except the network log-likelihood."""
    T, n_nodes, _ = X.shape
    lp = _dirichlet_logpdf(beta, np.repeat(hp.gamma / K, K))
    lp += _dirichlet_logpdf(weights[0, 0], hp.alpha_init * beta)
    deltas = hp.kappa * np.eye(K)
    for t in range(1, T):
        for k in range(K):
            lp += _dirichlet_logpdf(weights[t, k], hp.alpha * beta + deltas[k])
    return lp + mixture_log_prior(hp, X, intercept, intercept_prior, intercept_variance_prior, mu, sigma,
                                  z, weights, lmbda, radii=radii)


def mixture_log_prior(hp, X, intercept, intercept_prior, intercept_variance_prior, mu, sigma, z,
                      weights, lmbda, radii=None):
    """The terms hdp_lpcm.py:1208-1280 and lpcm.py:785-852 have in common: label Markov chains under
    ``weights`` (T, K, K; row [0, 0] the initial distribution), intercept prior, latent positions
    given the mixture, cluster means / variances, lambda, radii, scale hyper-priors."""
    T, n_nodes, _ = X.shape
    K = sigma.shape[0]
    lp = 0.0
    # label Markov chains (hdp_lpcm.py:1208-1212), vectorised over nodes
    lp += np.sum(np.log(weights[0, 0, z[0]]))
    for t in range(1, T):
        lp += np.sum(np.log(weights[t, z[t - 1], z[t]]))
    diff = intercept - intercept_prior
    if radii is not None:
        lp -= np.sum(0.5 * (diff * diff) / intercept_variance_prior)
    else:
        lp = lp - 0.5 * (diff * diff) / intercept_variance_prior
    for t in range(T):
        if t == 0:
            df = X[t] - mu[z[t]]
        else:
            df = X[t] - (1 - lmbda) * X[t - 1] - lmbda * mu[z[t]]
        lp = lp + np.sum(-0.5 * np.log(sigma[z[t]]) - 0.5 * np.sum(df * df, axis=1) / sigma[z[t]])
    for k in range(K):
        lp = lp - 0.5 * np.sum(mu[k] ** 2) / hp.mean_variance_prior
    lp = lp + np.sum(-(0.5 * hp.a + 1) * np.log(sigma[z]) - (0.5 * hp.b / sigma[z]))
    std = np.sqrt(hp.lambda_variance_prior)
    lp = lp + truncnorm.logpdf(lmbda, (0 - hp.lambda_prior) / std, (1 - hp.lambda_prior) / std,
                               loc=hp.lambda_prior, scale=std)
    if radii is not None:
        lp = lp + stats.dirichlet.logpdf(radii, np.ones(n_nodes))
    if hp.resample_mean_variance:
        lp = lp + (-(0.5 * hp.a0 + 1) * np.log(hp.mean_variance_prior) -
                   (0.5 * hp.b0 / hp.mean_variance_prior))
    if hp.resample_b:
        lp = lp + (hp.c0 - 1) * np.log(hp.b) - hp.d0 * hp.b
    return lp
