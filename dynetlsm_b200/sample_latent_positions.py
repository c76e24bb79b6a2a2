"""Function-level seams of the latent-position block: the reference's signatures
(sample_latent_positions.py:92 and :149), executed by the device sweep kernel.

The caller keeps the reference's conventions: ``X`` (T, n, d) is updated in place and returned,
``samplers`` is a T x n grid of ``Metropolis`` objects that the call mutates, the random draws
come from ``random_state`` in the reference's order (``randn(d)``, ``rand()`` per node).
"""
import numpy as np
from sklearn.utils import check_random_state

from . import _lib as L
from ._seams import engine_for, load_common
from .metropolis import pack_samplers, unpack_samplers

__all__ = ["sample_latent_positions", "sample_latent_positions_mixture"]


def _run(Y, X, intercept, samplers, radii, is_directed, squared, cc, rng, hyper, mixture):
    if squared:
        raise NotImplementedError("squared=True is never used by the estimators (lsm.py:490, "
                                  "hdp_lpcm.py:847) and is not built for the device")
    T, n, d = X.shape
    st = pack_samplers(samplers)
    K = 0 if mixture is None else mixture[1].shape[0]
    e = engine_for(Y, X, is_directed, K=K, cc=cc, tune=st["tune"], tune_interval=st["tune_interval"])
    load_common(e, X, intercept, radii)
    if hyper is not None:
        e.set_hyper(tau_sq=hyper[0], sigma_sq=hyper[1])
    if mixture is not None:
        mu, sigma, lmbda, z = mixture
        e.set(L.F_MU, mu[None]); e.set(L.F_SIGMA, sigma[None])
        e.set(L.F_LAMBDA, np.ravel(lmbda)[:1]); e.set(L.F_Z, z[None])
    e.set(L.F_X_STEP, st["step"][None]); e.set(L.F_X_NACC, st["n_accepted"][None])
    e.set(L.F_X_NSTEPS, st["n_steps"][None]); e.set(L.F_X_UNTIL, st["until"][None])
    eps, u = np.empty((1, T, n, d)), np.empty((1, T, n))
    for t in range(T):
        for j in range(n):
            eps[0, t, j] = rng.randn(d)
            u[0, t, j] = rng.rand()
    e.sweep_latent(eps, np.log(u))
    X[...] = e.get(L.F_X)[0]
    unpack_samplers(samplers, e.get(L.F_X_STEP)[0], e.get(L.F_X_NACC)[0], e.get(L.F_X_NSTEPS)[0],
                    e.get(L.F_X_UNTIL)[0])
    return X


def sample_latent_positions(Y, X, intercept, tau_sq, sigma_sq, samplers, radii=None,
                            is_directed=False, squared=False, case_control_sampler=None,
                            random_state=None):
    rng = check_random_state(random_state)
    return _run(Y, X, intercept, samplers, radii, is_directed, squared, case_control_sampler, rng,
                (tau_sq, sigma_sq), None)


def sample_latent_positions_mixture(Y, X, intercept, mu, sigma, lmbda, z, samplers, radii=None,
                                    is_directed=False, squared=None, case_control_sampler=None,
                                    random_state=None):
    rng = check_random_state(random_state)
    return _run(Y, X, intercept, samplers, radii, is_directed, False, case_control_sampler, rng,
                None, (np.asarray(mu, np.float64), np.asarray(sigma, np.float64), lmbda,
                       np.asarray(z)))
