"""Case-control bookkeeping on the host (reference: case_control_likelihood.py:8-112).

``DirectedCaseControlSampler`` keeps the reference's attribute names (``degrees_``, ``in_edges_``,
``out_edges_``, ``control_nodes_in_``, ``control_nodes_out_``) and its random-number consumption
(one ``rng.choice(..., replace=False)`` per node and direction, out before in), so a fit driven by
the same ``RandomState`` resamples the same control sets.  Besides a dense ``Y`` it accepts edge
lists directly (``init_from_edges``), which is how a network too large for a dense tensor
(n = 50 000) reaches the device.
"""
import numbers

import numpy as np
from sklearn.utils import check_random_state

__all__ = ["DirectedCaseControlSampler"]


class DirectedCaseControlSampler(object):
    def __init__(self, n_control=100, n_resample=100, random_state=None):
        self.n_control = n_control
        self.n_resample = n_resample
        self.random_state = random_state
        self.n_iter = 0

    # -- construction -------------------------------------------------------------------
    def _resolve_n_control(self, n_nodes):
        if isinstance(self.n_control, (numbers.Integral, np.integer)):
            self.n_control_ = int(self.n_control)
        else:
            self.n_control_ = int(self.n_control * n_nodes)

    def init(self, Y, sample=True):
        T, n, _ = Y.shape
        self._resolve_n_control(n)
        deg = np.zeros((T, n, 2), dtype=np.int64)
        deg[:, :, 0] = Y.sum(axis=1)  # in-degree: column sums
        deg[:, :, 1] = Y.sum(axis=2)  # out-degree: row sums
        in_e = np.zeros((T, n, int(deg[:, :, 0].max())), dtype=np.int64)
        out_e = np.zeros((T, n, int(deg[:, :, 1].max())), dtype=np.int64)
        for t in range(T):
            src, dst = np.nonzero(Y[t] == 1)          # row-major: sorted by src, then dst
            order = np.lexsort((src, dst))             # sorted by dst, then src
            _fill_padded(out_e[t], src, dst)
            _fill_padded(in_e[t], dst[order], src[order])
        return self.init_from_edges(deg, in_e, out_e, sample=sample)

    def init_from_edges(self, degrees, in_edges, out_edges, sample=True):
        """``sample=False`` leaves the control sets to the device (``Engine.resample_controls``)."""
        self.degrees_ = np.asarray(degrees, dtype=np.int64)
        self.in_edges_ = np.asarray(in_edges, dtype=np.int64)
        self.out_edges_ = np.asarray(out_edges, dtype=np.int64)
        if not hasattr(self, "n_control_"):
            self._resolve_n_control(self.degrees_.shape[1])
        if sample:
            self.control_nodes_in_, self.control_nodes_out_ = self.sample()
        else:
            self.control_nodes_in_ = self.control_nodes_out_ = None
        self.n_iter += 1
        return self

    # -- sampling -----------------------------------------------------------------------
    def sample(self):
        rng = check_random_state(self.random_state)
        T, n, _ = self.out_edges_.shape
        m = self.n_control_
        ctrl_out = np.full((T, n, m), -1, dtype=np.int64)
        ctrl_in = np.full((T, n, m), -1, dtype=np.int64)
        everyone = set(range(n))
        for t in range(T):
            for i in range(n):
                for col, edges, dst in ((1, self.out_edges_, ctrl_out), (0, self.in_edges_, ctrl_in)):
                    k = int(self.degrees_[t, i, col])
                    n_sample = min(n - k - 1, m)
                    # the candidate ORDER feeds rng.choice, so it is built the way the reference
                    # builds it: a set difference turned into a list
                    cand = list(set.difference(everyone, edges[t, i, :k].tolist() + [i]))
                    dst[t, i, :n_sample] = rng.choice(cand, size=n_sample, replace=False)
        return ctrl_in, ctrl_out

    def resample(self):
        if self.n_resample is not None and self.n_iter % self.n_resample == 0.:
            self.control_nodes_in_, self.control_nodes_out_ = self.sample()
            self.resampled_ = True
        else:
            self.resampled_ = False
        self.n_iter += 1
        return self.control_nodes_in_, self.control_nodes_out_


def _fill_padded(dst, owner, value):
    """dst[o, 0:k_o] = values of rows owned by o (owner sorted ascending)."""
    if owner.size == 0:
        return
    start = np.searchsorted(owner, owner, side="left")
    dst[owner, np.arange(owner.size) - start] = value
