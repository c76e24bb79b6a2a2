"""Host-side, one-off pieces around the hot path: initial values, Procrustes alignment and the
prior terms of the joint log-posterior.  None of this is on the per-sweep critical path of the
device sampler (init OUT, host sklearn/scipy is fine); it exists so that
``fit(Y)`` is a drop-in.  The third-party calls (sklearn MDS / KMeans / euclidean_distances, scipy
BFGS / shortest_path / orthogonal_procrustes) are made with the arguments and in the order the
reference makes them, because they consume the shared ``RandomState`` and therefore define the
chain that follows (reference: latent_space.py:36-153, lsm.py:32-97, procrustes.py:20-35).
"""
import numpy as np
from scipy.linalg import eigh, orthogonal_procrustes
from scipy.optimize import minimize
from scipy.sparse import csgraph
from sklearn.cluster import KMeans
from sklearn.manifold import MDS
from sklearn.metrics import euclidean_distances

__all__ = ["calculate_distances", "generalized_mds", "initialize_radii", "longitudinal_kmeans",
           "scale_intercept_mle", "directed_intercept_mle", "longitudinal_procrustes_rotation",
           "undirected_loglik_host", "directed_loglik_host", "triu_index_3d"]


def calculate_distances(X, squared=False):
    """latent_space.py:19-33 -- per-slice sklearn euclidean_distances."""
    if X.ndim == 2:
        return euclidean_distances(X, squared=squared)
    out = np.empty((X.shape[0], X.shape[1], X.shape[1]))
    for t in range(X.shape[0]):
        out[t] = euclidean_distances(X[t], squared=squared)
    return out


def triu_index_3d(T, n):
    """Index arrays of the strict upper triangles of a (T, n, n) stack, C order
    (array_utils.py:4-8 with k=1)."""
    i, j = np.triu_indices(n, 1)
    t = np.repeat(np.arange(T), i.size)
    return t, np.tile(i, T), np.tile(j, T)


def undirected_loglik_host(Y, dist, intercept, idx=None):
    """network_likelihoods.py:26-33 on the host (initialisation only)."""
    idx = triu_index_3d(*Y.shape[:2]) if idx is None else idx
    eta = intercept - dist[idx]
    return np.sum(Y[idx] * eta - np.log(1 + np.exp(eta)))


def directed_loglik_host(Y, dist, radii, b_in, b_out):
    """directed_likelihoods_fast.pyx:185-205 on the host (initialisation only), vectorised."""
    n = Y.shape[1]
    off = ~np.eye(n, dtype=bool)
    eta = b_in * (1 - dist / radii[None, None, :]) + b_out * (1 - dist / radii[None, :, None])
    return float(np.sum((Y * eta - np.log(1 + np.exp(eta)))[:, off]))


def _rotate_onto(ref, X):
    R, _ = orthogonal_procrustes(X, ref)
    return np.dot(X, R), R


def longitudinal_procrustes_rotation(X_ref, X):
    """One rotation for all time steps (procrustes.py:28-35)."""
    T, n = X.shape[:2]
    out, R = _rotate_onto(X_ref.reshape(T * n, -1), X.reshape(T * n, -1))
    return out.reshape(T, n, -1), R


def generalized_mds(Y, n_features=2, is_directed=False, lmbda=10, random_state=None):
    """Sarkar & Moore's generalised MDS initialisation (latent_space.py:47-95)."""
    T, n, _ = Y.shape
    D = np.empty((T, n, n))
    for t in range(T):
        sp = csgraph.shortest_path(Y[t], directed=False, unweighted=True)
        far = np.isinf(sp)
        sp[far] = np.max(sp[~far]) + 1
        D[t] = sp
    X = np.empty((T, n, n_features))
    X[0] = MDS(dissimilarity="precomputed", n_components=n_features,
               random_state=random_state).fit_transform(D[0])
    H = np.eye(n) - (1. / n) * np.ones((n, n))
    a, b = 1 / (1 + lmbda), lmbda / (1 + lmbda)
    for t in range(1, T):
        G = a * np.dot(H, np.dot(-0.5 * D[t] ** 2, H))
        G = G + b * np.dot(X[t - 1], X[t - 1].T)
        evals, evecs = eigh(G)
        evecs, evals = evecs[:, ::-1], evals[::-1]
        X[t] = evecs[:, :n_features] * np.sqrt(evals[:n_features])
        X[t], _ = _rotate_onto(X[t - 1], X[t])
    if is_directed:
        X /= n
    return X


def initialize_radii(Y, reg=1e-5):
    """latent_space.py:140-153."""
    radii = 0.5 * (Y.sum(axis=(0, 1)) + Y.sum(axis=(0, 2)))
    radii /= Y.sum()
    if np.any(radii == 0.):
        radii += reg
        radii /= np.sum(radii)
    return radii


def scale_intercept_mle(Y, X, tol=1e-4):
    """Joint MLE of a log-scale on the distances and the intercept (lsm.py:47-71), BFGS."""
    dist = calculate_distances(X)
    T, n = Y.shape[:2]
    idx = triu_index_3d(T, n)

    def expit_eta(sd, b):
        e = np.exp(b - sd)
        return e / (1 + e)

    def fun(x):
        return -undirected_loglik_host(Y, np.exp(x[0]) * dist, x[1], idx)

    def jac(x):
        sd = np.exp(x[0]) * dist
        g = Y - expit_eta(sd, x[1])
        g_scale = -sd * g
        return -np.array([np.sum(g_scale) - np.einsum("ikk", g_scale).sum(),
                          0.5 * (np.sum(g) - np.einsum("ikk", g).sum())])

    res = minimize(fun=fun, x0=np.array([0.0, 1.0]), method="BFGS", jac=jac, tol=tol)
    return res.x[0], res.x[1]


def directed_intercept_mle(Y, X, radii, tol=1e-4):
    """Conditional MLE of (beta_in, beta_out) (lsm.py:74-97).  The reference's gradient
    (directed_likelihoods_fast.pyx:20-43) reads an uninitialised accumulator; this one is the
    well-defined gradient, so directed initial values are NOT bit-comparable with the reference
    (SURVEY.md 7, K10)."""
    dist = calculate_distances(X)
    n = Y.shape[1]
    off = ~np.eye(n, dtype=bool)
    d_in = 1 - dist / radii[None, None, :]
    d_out = 1 - dist / radii[None, :, None]

    def fun(x):
        return -directed_loglik_host(Y, dist, radii, x[0], x[1])

    def jac(x):
        eta = x[0] * d_in + x[1] * d_out
        step = (Y - 1. / (1. + np.exp(-eta)))[:, off]
        return -np.array([np.sum(d_in[:, off] * step), np.sum(d_out[:, off] * step)])

    res = minimize(fun=fun, x0=np.array([0.0, 0.0]), method="BFGS", jac=jac, tol=tol)
    return res.x[0], res.x[1]


def longitudinal_kmeans(X, n_clusters=5, var_reg=1e-3, random_state=None):
    """Longitudinal k-means initialisation of the mixture (latent_space.py:98-137)."""
    T, n, d = X.shape
    X_vec = np.moveaxis(X, 0, -1).reshape(n, T * d)
    km = KMeans(n_clusters=n_clusters, random_state=random_state).fit(X_vec)
    labels = np.hstack([km.labels_.reshape(-1, 1)] * T).T
    centers = np.empty((n_clusters, d))
    for k in range(n_clusters):
        centers[k] = km.cluster_centers_[k].reshape(-1, T).T.mean(axis=0)
    variances = np.zeros(n_clusters)
    for k in range(n_clusters):
        for t in range(T):
            variances[k] += np.var(X[t][labels[t] == k], axis=0).mean()
        variances[k] /= T
    variances[variances == 0.] = var_reg
    return centers, variances, labels
