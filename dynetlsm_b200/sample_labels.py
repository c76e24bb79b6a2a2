"""Function-level seam of the HDP-HMM label block sampler (reference: sample_labels.py:134-190)."""
import numpy as np
from sklearn.utils import check_random_state

from . import _lib as L
from ._seams import engine_for

__all__ = ["sample_labels_block"]


def sample_labels_block(X, mu, sigma, lmbda, w, random_state=None):
    """Returns (z, n, nk, resp) with the reference's shapes and dtypes."""
    rng = check_random_state(random_state)
    T, n, d = X.shape
    K = sigma.shape[0]
    e = engine_for(None, X, False, K=K)
    e.set(L.F_X, np.asarray(X, np.float64)[None])
    e.set(L.F_MU, np.asarray(mu, np.float64)[None]); e.set(L.F_SIGMA, np.asarray(sigma, np.float64)[None])
    e.set(L.F_LAMBDA, np.ravel(lmbda)[:1]); e.set(L.F_WEIGHTS, np.asarray(w, np.float64)[None])
    e.sample_labels(rng.random_sample((1, n, T)))
    z = e.get(L.F_Z)[0].astype(np.int64)
    cnt = e.get(L.F_NCOUNT)[0]
    nk = e.get(L.F_NK)[0].astype(np.int64)
    resp = np.zeros((T, n, K), dtype=np.int64)
    resp[np.arange(T)[:, None], np.arange(n)[None, :], z] = 1
    return z, cnt, nk, resp
