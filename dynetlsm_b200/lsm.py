"""``DynamicNetworkLSM`` -- the reference estimator's API (lsm.py:100-625) over the device sampler.

Constructor keywords, ``fit(Y) -> self`` and the fitted attributes (``Xs_``, ``intercepts_``,
``radiis_``, ``logps_``, ``X_``, ``intercept_``, ``radii_``, ``logp_``, ``Y_fit_``,
``case_control_sampler_``, ``n_burn_``, ``distances_``, ``probas_``, ``auc_``) follow the
reference.  Three extra keywords select how the chain is driven:

    sampler='device'   device Philox streams, nothing but traces crosses PCIe (default)
    sampler='replay'   every random draw comes from the numpy ``RandomState`` in the reference's
                       order and is shipped to the device, which then reproduces the reference's
                       chain (accept/reject decisions bit-for-bit)
    n_chains=C         C independent chains on one GPU (``sampler='device'``); the drop-in
                       attributes describe chain 0, ``chains_`` holds the per-chain traces
    device=k           CUDA device ordinal

The hot path (latent-position sweep, full-network likelihood MH steps) runs in libdlsm.so; the
host keeps what the reference also does once per sweep in numpy: Procrustes, MAP bookkeeping,
prior terms of ``logp``.  There is no CPU fallback.
"""
import numpy as np
from scipy.special import expit
from sklearn.utils import check_array, check_random_state

from . import _lib as L
from .case_control_likelihood import DirectedCaseControlSampler, SparseNetwork
from .host_init import (calculate_distances, directed_intercept_mle, generalized_mds,
                        initialize_radii, longitudinal_procrustes_rotation, scale_intercept_mle)

__all__ = ["DynamicNetworkLSM"]


def _philox_seed(rng):
    return int(rng.randint(0, 2 ** 31 - 1)) | (int(rng.randint(0, 2 ** 31 - 1)) << 31)


class _Driver(object):
    """Shared device plumbing of the two estimators: one Engine, draws in either mode."""

    def __init__(self, Y, n_features, n_chains, is_directed, case_control_sampler, mixture_K, tune,
                 tune_interval, intercept_tune_interval, radii_tune, device, replay, rng):
        T, n, _ = Y.shape
        self.T, self.n, self.d, self.C = T, n, n_features, n_chains
        self.replay, self.rng = replay, rng
        self.is_directed, self.cc = is_directed, case_control_sampler
        sparse = isinstance(Y, SparseNetwork)
        self.engine = L.Engine(T=T, n=n, d=n_features, n_chains=n_chains, K=mixture_K,
                               is_directed=is_directed, case_control=self.cc is not None,
                               mixture=mixture_K > 0, device=device, tune=tune,
                               tune_interval=tune_interval,
                               intercept_tune_interval=intercept_tune_interval,
                               radii_tune=radii_tune, radii_tune_interval=100)
        if sparse:
            # ties in, degree / edge lists built on the device and mirrored under the reference's names
            self.engine.set_network_edges(Y.edges)
            self.cc.init_from_edges(*self.engine.get_edge_lists(), sample=False)
        elif self.cc is None:
            self.engine.set_network(Y)
        else:
            self.engine.set_edge_lists(self.cc.degrees_, self.cc.in_edges_, self.cc.out_edges_)
            if self.cc.control_nodes_in_ is not None:
                self.push_controls()
        if not replay:
            self.engine.set_rng(_philox_seed(rng))
            if self.cc is not None and self.cc.control_nodes_in_ is None:
                self.draw_controls()

    def draw_controls(self):
        """Device-side redraw of the control sets, one set per chain."""
        self.engine.resample_controls(self.cc.n_control_, per_chain=self.C > 1)

    def push_controls(self):
        self.engine.set_controls(self.cc.control_nodes_in_, self.cc.control_nodes_out_)

    # one latent sweep for every chain -------------------------------------------------
    def sweep_latent(self):
        e = self.engine
        if not self.replay:
            return e.sweep_latent()
        T, n, d, rng = self.T, self.n, self.d, self.rng
        eps = np.empty((1, T, n, d))
        u = np.empty((1, T, n))
        for t in range(T):           # metropolis.py:44,49: randn(d) then rand(), node by node
            for j in range(n):
                eps[0, t, j] = rng.randn(d)
                u[0, t, j] = rng.rand()
        return e.sweep_latent(eps, np.log(u))

    def sample_intercepts(self):
        e = self.engine
        if not self.replay:
            return e.sample_intercepts()
        m = e.m
        # sample_coefficients.py:20-88: randn(1) when the proposal is made, rand() after both
        # log-posterior evaluations; the second intercept's draws follow the first's
        eps, u = np.empty((1, m)), np.empty((1, m))
        for i in range(m):
            eps[0, i] = self.rng.randn(1)[0]
            u[0, i] = self.rng.rand()
        return e.sample_intercepts(eps, np.log(u))

    def sample_radii(self):
        e = self.engine
        if not self.replay:
            return e.sample_radii()
        radii = e.get(L.F_RADII)[0]
        step = e.get(L.F_R_STEP)[0]
        prop = self.rng.dirichlet(step * radii)      # metropolis.py:61-67
        if np.any(prop == 0.):
            prop += 1e-5
            prop /= np.sum(prop)
        u = self.rng.rand()
        return e.sample_radii(prop[None], np.array([np.log(u)]))

    def sample_labels(self):
        e = self.engine
        if not self.replay:
            return e.sample_labels()
        # sample_labels.py:16-19: one uniform per (node, time), node-major; uniform(0, c) = c * U
        U = self.rng.random_sample((1, self.n, self.T))
        return e.sample_labels(U)


class _FittedNetworkMixin(object):
    """Derived quantities of a fitted estimator (lsm.py:270-317, hdp_lpcm.py:462-495)."""

    @property
    def n_burn_(self):
        return (self.burn or 0) + (self.tune or 0)

    @property
    def distances_(self):
        if not hasattr(self, "X_"):
            raise ValueError("Model not fit.")
        return calculate_distances(self.X_)

    @property
    def probas_(self):
        """Edge probabilities at the point estimate, (T, n, n), zero diagonal: directed
        directed_network_probas (directed_likelihoods_fast.pyx:273-294), undirected
        expit(intercept - dist); evaluated by dlsm_edge_probas on the device."""
        if not hasattr(self, "X_"):
            raise ValueError("Model not fit.")
        # cached per fitted point estimate: auc_ and repeated accesses do not rebuild an engine
        key = (id(self.X_), float(np.sum(self.X_)), tuple(np.ravel(self.intercept_)))
        hit = getattr(self, "_probas_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        T, n, d = self.X_.shape
        e = L.Engine(T=T, n=n, d=d, n_chains=1, is_directed=self.is_directed, device=self.device)
        try:
            e.set(L.F_X, self.X_[None])
            ic = np.zeros((1, 2))
            ic[0, :np.size(self.intercept_)] = np.ravel(self.intercept_)
            e.set(L.F_INTERCEPT, ic)
            if self.is_directed:
                e.set(L.F_RADII, np.asarray(self.radii_)[None])
            out = e.edge_probas(0)
        finally:
            e.close()
        self._probas_cache = (key, out)
        return out

    @property
    def auc_(self):
        from sklearn.metrics import roc_auc_score
        if not hasattr(self, "X_"):
            raise ValueError("Model not fit.")
        n = self.Y_fit_.shape[1]
        Y = self.Y_fit_.toarray() if isinstance(self.Y_fit_, SparseNetwork) else self.Y_fit_
        mask = ~np.eye(n, dtype=bool) if self.is_directed else np.triu(np.ones((n, n), bool), 1)
        return roc_auc_score(Y[:, mask].ravel(), self.probas_[:, mask].ravel())


class DynamicNetworkLSM(_FittedNetworkMixin):
    """Latent space model for dynamic networks (Sewell & Chen 2015), sampled on a B200.

    Parameters follow the reference estimator (lsm.py:103-213); see the module docstring for the
    extra keywords ``sampler``, ``n_chains`` and ``device``.
    """

    def __init__(self, n_features=2, is_directed=False, n_iter=5000, tune=2500, tune_interval=100,
                 burn=2500, intercept_prior="auto", intercept_variance_prior=2.0, tau_sq=2.0,
                 sigma_sq=0.1, step_size_X=0.1, step_size_intercept=0.1, step_size_radii=175000,
                 n_control=None, n_resample_control=100, copy=True, random_state=None,
                 sampler="device", n_chains=1, device=0):
        self.n_iter = n_iter
        self.is_directed = is_directed
        self.n_features = n_features
        self.tau_sq = tau_sq
        self.sigma_sq = sigma_sq
        self.step_size_X = step_size_X
        self.intercept_prior = intercept_prior
        self.intercept_variance_prior = intercept_variance_prior
        self.step_size_intercept = step_size_intercept
        self.step_size_radii = step_size_radii
        self.tune = tune
        self.tune_interval = tune_interval
        self.burn = burn
        self.n_control = n_control
        self.n_resample_control = n_resample_control
        self.copy = copy
        self.random_state = random_state
        self.sampler = sampler
        self.n_chains = n_chains
        self.device = device

    # -- reference properties (lsm.py:270-317) live in _FittedNetworkMixin ----------------------

    # -- joint log-posterior (lsm.py:576-625), network term from the device -------------------
    def _log_prior(self, X, intercept):
        lp = 0.0
        for t in range(X.shape[0]):
            if t == 0:
                lp -= np.sum(0.5 * np.sum(X[t] * X[t], axis=1) / self.tau_sq)
            else:
                diff = X[t] - X[t - 1]
                lp -= np.sum(0.5 * np.sum(diff * diff, axis=1) / self.sigma_sq)
        if self.is_directed:
            diff = intercept - self.intercept_prior
            lp -= np.sum(0.5 * (diff * diff) / self.intercept_variance_prior)
        else:
            diff = intercept[0] - np.ravel(self.intercept_prior)[0]
            lp -= 0.5 * (diff * diff) / self.intercept_variance_prior
        return lp

    def _replay_loop(self, drv, S, n_iter_procrustes, Xs, ics, rads, logps):
        """The reference's loop, block by block, on its own RandomState (lsm.py:474-572)."""
        e, C, m = drv.engine, self.n_chains, (2 if self.is_directed else 1)
        for it in range(1, S):
            if self.case_control_sampler_ is not None:
                self.case_control_sampler_.resample()
                if self.case_control_sampler_.resampled_:
                    drv.push_controls()
            drv.sweep_latent()
            if it > n_iter_procrustes:   # lsm.py:495-498: align with the best pre-burn-in sample
                Xh = e.get(L.F_X)
                for c in range(C):
                    ref = Xs[c, np.argmax(logps[c, :(n_iter_procrustes + 1)])]
                    Xh[c], _ = longitudinal_procrustes_rotation(ref, Xh[c])
                e.set(L.F_X, Xh)
            e.center()
            drv.sample_intercepts()
            if self.is_directed:
                drv.sample_radii()
            Xs[:, it] = e.get(L.F_X)
            ics[:, it] = e.get(L.F_INTERCEPT)[:, :m]
            if self.is_directed:
                rads[:, it] = e.get(L.F_RADII)
            ll = e.loglik_full()
            for c in range(C):
                logps[c, it] = ll[c] + self._log_prior(Xs[c, it], ics[c, it])

    def _device_loop(self, drv, S, n_iter_procrustes, Xs, ics, rads, logps):
        """Device-resident chains: sweeps, centring, Procrustes, log-posterior and traces all stay
        on the GPU (dlsm_run_traced); the host only intervenes where the reference's loop changes
        regime -- the end of burn-in (Procrustes reference) and case-control resampling."""
        e, C, m = drv.engine, self.n_chains, (2 if self.is_directed else 1)
        cc = self.case_control_sampler_
        fields = (L.F_X, L.F_INTERCEPT) + ((L.F_RADII,) if self.is_directed else ())
        rec_bytes = 8 * C * (self.Y_fit_.shape[0] * self.Y_fit_.shape[1] * (self.n_features + 1) + 4)
        seg = int(max(1, min(S, (128 << 20) // rec_bytes)))   # bounded, reused pinned destination
        it, tr = 1, None
        while it < S:
            stop = S
            if it <= n_iter_procrustes:
                stop = min(stop, n_iter_procrustes + 1)
            if cc is not None:                     # lsm.py:478-481, case_control_likelihood.py:27-33
                if cc.n_resample is not None and cc.n_iter % cc.n_resample == 0:
                    drv.draw_controls()            # dlsm_resample_controls: no host sampling
                cc.n_iter += 1
                # iterations whose resample() call is a no-op are batched with this one
                quiet = S if cc.n_resample is None else (cc.n_resample - cc.n_iter % cc.n_resample) % cc.n_resample
                stop = min(stop, it + 1 + quiet)
            if it == n_iter_procrustes + 1:        # lsm.py:495-498
                ref = np.stack([Xs[c, np.argmax(logps[c, :(n_iter_procrustes + 1)])] for c in range(C)])
                e.set_procrustes_ref(ref)
            stop = min(stop, it + seg)
            tr = e.run_traced(stop - it, fields_all=fields, pinned=True, out=tr)
            Xs[:, it:stop] = tr[L.F_X].transpose(1, 0, 2, 3, 4)
            ics[:, it:stop] = tr[L.F_INTERCEPT].transpose(1, 0, 2)[:, :, :m]
            if self.is_directed:
                rads[:, it:stop] = tr[L.F_RADII].transpose(1, 0, 2)
            logps[:, it:stop] = tr["logp"].T
            if cc is not None:
                cc.n_iter += stop - it - 1
            it = stop
        e.set_procrustes_ref(None)
        if cc is not None:                         # expose the last sets under the reference's names
            ci, co = e.get_controls()
            cc.control_nodes_in_, cc.control_nodes_out_ = ci[0].astype(np.int64), co[0].astype(np.int64)

    def fit(self, Y, X_init=None, radii_init=None, intercept_init=None):
        """Sample from the posterior given the dynamic network ``Y``: a dense (T, n, n) array with
        entries 0/1, or -- with the case-control likelihood (``n_control``) -- a ``SparseNetwork`` /
        a sequence of T scipy.sparse matrices.  The reference's starting values (generalised MDS on
        shortest-path distances, lsm.py:385-413) need the dense tensor; a sparse network starts from
        ``X_init`` (T, n, d), ``radii_init`` (n,), ``intercept_init`` (2,) or, where those are not
        given, from a random walk under the model's own prior, degree-proportional radii and
        intercepts (1, 1)."""
        if self.sampler not in ("device", "replay"):
            raise ValueError("`sampler` must be 'device' or 'replay', got {}".format(self.sampler))
        replay = self.sampler == "replay"
        if replay and self.n_chains != 1:
            raise ValueError("sampler='replay' reproduces one reference chain; use n_chains=1")
        if isinstance(Y, (list, tuple)) and len(Y) and hasattr(Y[0], "tocoo"):
            Y = SparseNetwork.from_scipy(Y)
        sparse = isinstance(Y, SparseNetwork)
        if sparse and (self.n_control is None or not self.is_directed or replay):
            raise ValueError("a sparse network needs the case-control likelihood of the directed model "
                             "(is_directed=True, n_control=...) and sampler='device'")
        n_time_steps, n_nodes, _ = Y.shape
        rng = check_random_state(self.random_state)
        if not sparse:
            Y = check_array(Y, dtype=np.float64, ensure_all_finite="allow-nan", ensure_2d=False,
                            allow_nd=True, copy=self.copy)
            if np.any(Y == -1) or np.any(np.isnan(Y)):
                raise NotImplementedError("missing dyads (-1 / NaN) are not supported by the device "
                                          "sampler; impute them first (the reference's "
                                          "SimpleNetworkImputer is outside the accelerated path)")
        self.Y_fit_ = Y
        n_iter_procrustes = 0
        if self.tune is not None:
            self.n_iter += self.tune
            n_iter_procrustes += self.tune
        if self.burn is not None:
            self.n_iter += self.burn
            n_iter_procrustes += self.burn
        S, C, m = self.n_iter, self.n_chains, (2 if self.is_directed else 1)

        # ---- initial values on the host (lsm.py:385-413) ----
        radii = None
        if sparse:
            if X_init is not None:
                X = np.array(X_init, dtype=np.float64)
            else:   # a random walk at the scale the reference initialises directed models at (1 / n)
                X = np.empty((n_time_steps, n_nodes, self.n_features))
                X[0] = rng.randn(n_nodes, self.n_features) / n_nodes
                for t in range(1, n_time_steps):
                    X[t] = X[t - 1] + np.sqrt(self.sigma_sq) / n_nodes * rng.randn(n_nodes, self.n_features)
            if radii_init is not None:
                radii = np.array(radii_init, dtype=np.float64)
            else:   # initialize_radii (latent_space.py): share of the ties a node takes part in
                cnt = np.bincount(Y.edges[:, 1], minlength=n_nodes) + np.bincount(Y.edges[:, 2], minlength=n_nodes)
                radii = (0.5 * cnt + 1e-5 * max(1, Y.edges.shape[0])) / max(1, Y.edges.shape[0])
                radii /= radii.sum()
            intercept = np.array([1.0, 1.0]) if intercept_init is None else np.array(intercept_init, dtype=np.float64)
        else:
            X = generalized_mds(Y, n_features=self.n_features, is_directed=self.is_directed,
                                random_state=rng)
            if self.is_directed:
                radii = initialize_radii(Y)
                intercept = np.array(directed_intercept_mle(Y, X, radii))
            else:
                scale, b = scale_intercept_mle(Y, X)
                intercept = np.array([b])
                X *= np.exp(scale)
        X -= np.mean(X, axis=(0, 1))
        if isinstance(self.tau_sq, str) and self.tau_sq == "auto":
            self.tau_sq = np.mean(X[0] * X[0])
        if isinstance(self.intercept_prior, str) and self.intercept_prior == "auto":
            self.intercept_prior = intercept.copy()

        self.case_control_sampler_ = None
        if self.n_control is not None:
            if not self.is_directed:
                raise ValueError("The case-control likelihood currently only "
                                 "supported for directed networks.")
            self.case_control_sampler_ = DirectedCaseControlSampler(
                n_control=self.n_control, n_resample=self.n_resample_control, random_state=rng)
            if not sparse:   # (a sparse network's lists are built on the device, in _Driver)
                self.case_control_sampler_.init(Y, sample=replay)   # device mode draws the controls on the GPU

        # ---- device state ----
        drv = _Driver(Y, self.n_features, C, self.is_directed, self.case_control_sampler_, 0,
                      self.tune, self.tune_interval,
                      # lsm.py:459-467: only the directed intercept samplers get tune_interval
                      (self.tune_interval,) * 2 if self.is_directed else (100, 100),
                      None, self.device, replay, rng)   # lsm.py:470-472: radii sampler never tunes
        e = drv.engine
        self._engine = e
        disp = np.zeros((C, 1, 1, 1))
        Xc = np.tile(X[None], (C, 1, 1, 1))
        if C > 1:  # dispersed starts for the extra chains
            Xc[1:] += 0.1 * np.std(X) * rng.randn(C - 1, *X.shape)
        e.set(L.F_X, Xc + disp)
        ic = np.zeros((C, 2))
        ic[:, :m] = intercept
        e.set(L.F_INTERCEPT, ic)
        if self.is_directed:
            e.set(L.F_RADII, np.tile(radii[None], (C, 1)))
        e.set_hyper(tau_sq=self.tau_sq, sigma_sq=self.sigma_sq, intercept_prior=self.intercept_prior,
                    intercept_variance_prior=self.intercept_variance_prior)
        e.set_tuner(self.step_size_X, self.step_size_intercept, self.step_size_radii)

        # ---- traces ----
        Xs = np.zeros((C, S, n_time_steps, n_nodes, self.n_features))
        ics = np.zeros((C, S, m))
        rads = np.zeros((C, S, n_nodes)) if self.is_directed else None
        logps = np.zeros((C, S))
        Xs[:, 0] = e.get(L.F_X)
        ics[:, 0] = intercept
        if self.is_directed:
            rads[:, 0] = radii
        logps[:, 0] = e.logp() if not replay else [
            ll + self._log_prior(Xs[c, 0], ics[c, 0]) for c, ll in enumerate(e.loglik_full())]

        if replay:
            self._replay_loop(drv, S, n_iter_procrustes, Xs, ics, rads, logps)
        else:
            self._device_loop(drv, S, n_iter_procrustes, Xs, ics, rads, logps)

        # MAP bookkeeping (lsm.py:554-566): restart at the end of burn-in, then track the max
        best = []
        nb = (self.tune or 0) + (self.burn or 0)
        for c in range(C):
            start = nb if (self.tune and nb < S) else 0
            best.append(dict(it=start + int(np.argmax(logps[c, start:]))))
            best[c]["logp"] = logps[c, best[c]["it"]]

        # ---- drop-in attributes (chain 0) + per-chain traces ----
        self.Xs_, self.intercepts_, self.logps_ = Xs[0], ics[0], logps[0]
        if self.is_directed:
            self.radiis_ = rads[0]
        b0 = best[0]["it"]
        self.logp_ = logps[0, b0]
        self.X_ = Xs[0, b0]
        self.intercept_ = ics[0, b0]
        if self.is_directed:
            self.radii_ = rads[0, b0]
        self.chains_ = dict(Xs=Xs, intercepts=ics, logps=logps, radiis=rads,
                            map_iteration=[b["it"] for b in best])
        self.sampler_counters_ = e.counters()
        e.close()            # chain state, trace rings and pinned buffers are not kept after fit
        self._engine = None
        return self
