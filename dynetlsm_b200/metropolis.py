"""Host mirror of the reference's ``Metropolis`` sampler objects (metropolis.py:85-136).

On the device the state of every sampler lives in SoA arrays (step_size, n_accepted, n_steps,
steps_until_tune); these light objects exist so that the function-level seams
(``sample_latent_positions(..., samplers=...)``) keep the reference's calling convention: the
caller owns a grid of ``Metropolis`` objects, the call mutates them.
"""
import numpy as np

__all__ = ["Metropolis", "pack_samplers", "unpack_samplers"]


class Metropolis(object):
    def __init__(self, step_size=0.1, tune=500, tune_interval=100, proposal_type="random_walk"):
        if proposal_type not in ("random_walk", "dirichlet"):
            raise ValueError("`proposal_type` must be in {'random_walk', 'dirichlet'}, but got "
                             "{}".format(proposal_type))
        self.step_size = step_size
        self.tune = tune
        self.tune_interval = tune_interval
        self.proposal_type = proposal_type
        self.steps_until_tune = tune_interval
        self.n_accepted = 0
        self.n_steps = 0


def pack_samplers(samplers):
    """list (or list of lists) of Metropolis -> dict of arrays with the same leading shape."""
    grid = np.array(samplers, dtype=object)
    flat = grid.ravel()
    out = dict(step=np.array([s.step_size for s in flat], dtype=np.float64).reshape(grid.shape),
               n_accepted=np.array([s.n_accepted for s in flat], dtype=np.int32).reshape(grid.shape),
               n_steps=np.array([s.n_steps for s in flat], dtype=np.int32).reshape(grid.shape),
               until=np.array([s.steps_until_tune for s in flat], dtype=np.int32).reshape(grid.shape))
    first = flat[0]
    out["tune"] = first.tune
    out["tune_interval"] = first.tune_interval
    return out


def unpack_samplers(samplers, step, n_accepted, n_steps, until):
    grid = np.array(samplers, dtype=object)
    for s, a, b, c, d in zip(grid.ravel(), np.ravel(step), np.ravel(n_accepted), np.ravel(n_steps),
                             np.ravel(until)):
        s.step_size = float(a)
        s.n_accepted = int(b)
        s.n_steps = int(c)
        s.steps_until_tune = int(d)
