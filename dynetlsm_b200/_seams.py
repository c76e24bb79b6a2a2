"""Shared plumbing of the function-level seams: a small cache of Engines keyed by the network
object and the model flags, so that repeated calls with the same ``Y`` (the way the reference's
estimator loops call the block samplers) do not re-upload the adjacency."""
import numpy as np

from . import _lib as L

_CACHE = {}
_MAX = 4


def clear_cache():
    """Drop the cached engines (device memory, adjacency bits) explicitly."""
    for e, _ in _CACHE.values():
        e.close()
    _CACHE.clear()


def engine_for(Y, X, is_directed, K=0, cc=None, tune=None, tune_interval=100,
               intercept_tune_interval=(100, 100), radii_tune=None, radii_tune_interval=100):
    T, n, d = X.shape
    # content fingerprint: a Y mutated in place between calls (imputed dyads) must not reuse the
    # stale packed adjacency already on the device
    finger = None if Y is None else float(np.dot(np.asarray(Y, dtype=np.float64).ravel()[::7], 1.0 + np.arange(
        np.asarray(Y).size)[::7] % 1021))
    key = (id(Y), finger, None if Y is None else Y.shape, T, n, d, bool(is_directed), int(K), id(cc),
           tune, tune_interval, tuple(intercept_tune_interval), radii_tune, radii_tune_interval)
    hit = _CACHE.get(key)
    if hit is not None and (Y is None or hit[1] is Y):
        e = hit[0]
    else:
        e = L.Engine(T=T, n=n, d=d, n_chains=1, K=K, is_directed=is_directed, case_control=cc is not None,
                     mixture=K > 0, tune=tune, tune_interval=tune_interval,
                     intercept_tune_interval=intercept_tune_interval, radii_tune=radii_tune,
                     radii_tune_interval=radii_tune_interval)
        if cc is None:
            if Y is not None:
                Yd = np.asarray(Y, dtype=np.float64)
                e.set_network(Yd[None] if Yd.ndim == 2 else Yd)
        else:
            e.set_edge_lists(cc.degrees_, cc.in_edges_, cc.out_edges_)
        if len(_CACHE) >= _MAX:
            _CACHE.pop(next(iter(_CACHE)))[0].close()
        _CACHE[key] = (e, Y)
    if cc is not None:
        e.set_controls(cc.control_nodes_in_, cc.control_nodes_out_)
    return e


def load_common(e, X, intercept, radii=None):
    e.set(L.F_X, np.asarray(X, dtype=np.float64)[None])
    ic = np.zeros((1, 2))
    b = np.atleast_1d(np.asarray(intercept, dtype=np.float64))
    ic[0, :b.size] = b
    e.set(L.F_INTERCEPT, ic)
    if radii is not None:
        e.set(L.F_RADII, np.asarray(radii, dtype=np.float64)[None])
