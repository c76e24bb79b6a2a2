"""Function-level seams of the intercept / radii MH blocks (reference:
sample_coefficients.py:12-88 and :91-121) on the device full-network likelihood kernel.
``dist`` is accepted for signature compatibility and ignored: the device recomputes distances
from ``X`` (no (T, n, n) fp64 cache exists on this side)."""
import numpy as np
from sklearn.utils import check_random_state

from . import _lib as L
from ._seams import engine_for, load_common
from .metropolis import pack_samplers, unpack_samplers

__all__ = ["sample_intercepts", "sample_radii"]


def sample_intercepts(Y, X, intercepts, intercept_prior, intercept_variance_prior, samplers,
                      radii=None, dist=None, is_directed=False, case_control_sampler=None,
                      squared=False, random_state=None):
    rng = check_random_state(random_state)
    m = 2 if is_directed else 1
    st = pack_samplers(samplers)
    iv = tuple(s.tune_interval for s in samplers) + (100,) * (2 - len(samplers))
    e = engine_for(Y, X, is_directed, cc=case_control_sampler, tune=st["tune"],
                   intercept_tune_interval=iv[:2])
    load_common(e, X, intercepts, radii)
    e.set_hyper(intercept_prior=intercept_prior, intercept_variance_prior=intercept_variance_prior)
    pad = lambda a, dt: np.concatenate([np.asarray(a, dt), np.zeros(2 - m, dt)])[None]
    e.set(L.F_B_STEP, pad(st["step"], np.float64)); e.set(L.F_B_NACC, pad(st["n_accepted"], np.int32))
    e.set(L.F_B_NSTEPS, pad(st["n_steps"], np.int32)); e.set(L.F_B_UNTIL, pad(st["until"], np.int32))
    eps, u = np.empty((1, m)), np.empty((1, m))
    for i in range(m):
        eps[0, i] = rng.randn(1)[0]
        u[0, i] = rng.rand()
    e.sample_intercepts(eps, np.log(u))
    out = e.get(L.F_INTERCEPT)[0, :m]
    unpack_samplers(samplers, e.get(L.F_B_STEP)[0, :m], e.get(L.F_B_NACC)[0, :m],
                    e.get(L.F_B_NSTEPS)[0, :m], e.get(L.F_B_UNTIL)[0, :m])
    if is_directed:
        intercepts[:] = out
        return intercepts
    return out


def sample_radii(Y, X, intercepts, radii, sampler, dist=None, case_control_sampler=None,
                 squared=False, random_state=None):
    rng = check_random_state(random_state)
    e = engine_for(Y, X, True, cc=case_control_sampler, radii_tune=sampler.tune,
                   radii_tune_interval=sampler.tune_interval)
    load_common(e, X, intercepts, radii)
    e.set(L.F_R_STEP, np.array([sampler.step_size], np.float64))
    e.set(L.F_R_NACC, np.array([sampler.n_accepted], np.int32))
    e.set(L.F_R_NSTEPS, np.array([sampler.n_steps], np.int32))
    e.set(L.F_R_UNTIL, np.array([sampler.steps_until_tune], np.int32))
    prop = rng.dirichlet(sampler.step_size * radii)
    if np.any(prop == 0.):
        prop += 1e-5
        prop /= np.sum(prop)
    u = rng.rand()
    e.sample_radii(prop[None], np.array([np.log(u)]))
    unpack_samplers([sampler], e.get(L.F_R_STEP), e.get(L.F_R_NACC), e.get(L.F_R_NSTEPS),
                    e.get(L.F_R_UNTIL))
    return e.get(L.F_RADII)[0]
