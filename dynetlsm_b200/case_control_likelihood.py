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

__all__ = ["DirectedCaseControlSampler", "SparseNetwork"]


class SparseNetwork(object):
    """A directed dynamic network given by its ties: ``edges`` (E, 3) integer rows (t, sender,
    receiver), every tie once, no self ties.  ``fit`` accepts it in place of the dense (T, n, n)
    tensor when the case-control likelihood is used (``n_control`` set): the degree / edge-list
    bookkeeping of ``DirectedCaseControlSampler.init`` (case_control_likelihood.py:37-73) is then
    built on the device (``dlsm_set_network_edges``), so networks whose dense tensor cannot exist
    (n = 50 000: 200 GB) go through the estimator."""

    def __init__(self, edges, n_time_steps, n_nodes):
        self.edges = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 3)
        self.shape = (int(n_time_steps), int(n_nodes), int(n_nodes))

    @classmethod
    def from_dense(cls, Y):
        Y = np.asarray(Y)
        return cls(np.argwhere(Y == 1), Y.shape[0], Y.shape[1])

    @classmethod
    def from_scipy(cls, matrices):
        """A sequence of T scipy.sparse matrices (n, n); entry (i, j) != 0 is a tie i -> j."""
        rows = []
        for t, m in enumerate(matrices):
            coo = m.tocoo()
            keep = (coo.data != 0) & (coo.row != coo.col)
            rows.append(np.stack([np.full(int(keep.sum()), t), coo.row[keep], coo.col[keep]], axis=1))
        return cls(np.concatenate(rows) if rows else np.zeros((0, 3)), len(matrices), matrices[0].shape[0])

    def toarray(self):
        Y = np.zeros(self.shape)
        Y[self.edges[:, 0], self.edges[:, 1], self.edges[:, 2]] = 1.0
        return Y


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
