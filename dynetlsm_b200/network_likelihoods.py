"""Function-level seams of the likelihood facade (reference: network_likelihoods.py:16-66 and the
Cython kernels it re-exports) on the device probe kernels."""
import numpy as np

from . import _lib as L
from ._seams import engine_for, load_common

__all__ = ["partial_loglikelihood", "directed_partial_loglikelihood",
           "dynamic_network_loglikelihood_undirected", "dynamic_network_loglikelihood_directed"]


def partial_loglikelihood(Y, X, intercept, node_id, squared=False):
    """static_network_fast.pyx:17-44 for one time slice: Y (n, n), X (n, d)."""
    if squared:
        raise NotImplementedError("squared=True is not built for the device")
    e = engine_for(Y, X[None], False)
    load_common(e, X[None], intercept)
    return float(e.loglik_partial()[0, 0, node_id])


def directed_partial_loglikelihood(Y, X, radii, intercept_in, intercept_out, node_id, squared=False):
    """directed_likelihoods_fast.pyx:46-80 for one time slice."""
    if squared:
        raise NotImplementedError("squared=True is not built for the device")
    e = engine_for(Y, X[None], True)
    load_common(e, X[None], [intercept_in, intercept_out], radii)
    return float(e.loglik_partial()[0, 0, node_id])


def dynamic_network_loglikelihood_undirected(Y, X, intercept, squared=False, dist=None):
    e = engine_for(Y, X, False)
    load_common(e, X, intercept)
    return float(e.loglik_full()[0])


def dynamic_network_loglikelihood_directed(Y, X, intercept_in, intercept_out, radii, squared=False,
                                           dist=None):
    e = engine_for(Y, X, True)
    load_common(e, X, [intercept_in, intercept_out], radii)
    return float(e.loglik_full()[0])
