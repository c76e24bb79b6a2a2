"""``DynamicNetworkLPCM`` -- the reference's finite-mixture estimator (lpcm.py:134-873) over the
device sampler.

The finite mixture shares the whole hot path with the HDP model: per sweep the device runs the
mixture-prior latent-position sweep, the centring, the intercept / radii MH on the full-network
(or case-control) likelihood and the label forward-filter / backward-sample -- the reference's
``sample_labels_block_lpcm(init_weights, trans_weights)`` (sample_labels.py:73-131) is its
``sample_labels_block(w)`` (:134-188) with ``w[0, 0] = init_weights`` and ``w[t] = trans_weights``
for t >= 1, same arithmetic, so the label kernel serves both.  What differs is the conjugate
block: K + 1 Dirichlet draws under a fixed symmetric prior (lpcm.py:573-580) instead of the HDP's
auxiliary-variable scheme.  It is a few KB of state per sweep and runs on the host in numpy on
the estimator's ``RandomState``, in the reference's draw order.

``sampler='replay'`` draws the hot path's random numbers from the same ``RandomState`` too, so the
chain reproduces the reference's draw for draw (tests/test_gpu_lpcm.py against
tests/golden/lpcm_*.npz); ``sampler='device'`` (default) uses the device Philox streams for the
hot path.  Forecasts (forecast.pyx) and missing dyads are outside the accelerated path.
"""
import numpy as np
from sklearn.utils import check_array, check_random_state

from . import _lib as L
from .case_control_likelihood import DirectedCaseControlSampler
from .hdp_updates import _clipped_dirichlet, _dirichlet_logpdf, mixture_log_prior, mixture_updates
from .host_init import longitudinal_kmeans, longitudinal_procrustes_rotation
from .lsm import DynamicNetworkLSM, _Driver, _FittedNetworkMixin
from . import model_selection as MS

__all__ = ["DynamicNetworkLPCM", "lpcm_conjugate_updates", "lpcm_log_prior", "stacked_weights"]


class _NotOnDevicePath(NotImplementedError, AttributeError):
    """Raised by the forecast properties; an AttributeError too, so hasattr() / inspect stay usable."""


class MixtureHyper(object):
    """Hyper-parameter state the reference keeps on the estimator (lpcm.py:394-471)."""

    def __init__(self, dirichlet_prior, mean_variance_prior, b, a, a0, b0, c0, d0, lambda_prior,
                 lambda_variance_prior, resample_mean_variance, resample_b):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def stacked_weights(init_weights, trans_weights, T):
    """(T, K, K) table of the label kernel: row [0, 0] the initial distribution, [t] the
    transition matrix for every t >= 1 (sample_labels.py:101-116 vs :163-178)."""
    K = init_weights.shape[0]
    w = np.zeros((T, K, K))
    w[0, 0] = init_weights
    w[1:] = trans_weights
    return w


def lpcm_conjugate_updates(rng, hp, X, z, n, nk, mu, sigma, lmbda, init_weights, trans_weights):
    """lpcm.py:573-656: initial and transition distributions, then the block shared with the HDP
    model.  ``n`` (T, K, K) / ``nk`` (T, K) come from the device label kernel.  The weight arrays,
    mu and sigma are updated in place; returns lmbda."""
    K = sigma.shape[0]
    init_weights[:] = _clipped_dirichlet(rng, hp.dirichlet_prior + nk[0])
    for k in range(K):
        trans_weights[k] = _clipped_dirichlet(rng, hp.dirichlet_prior + n[1:, k].sum(axis=0))
    return mixture_updates(rng, hp, X, z, nk, mu, sigma, lmbda)


def lpcm_log_prior(hp, X, intercept, intercept_prior, intercept_variance_prior, mu, sigma, z,
                   init_weights, trans_weights, lmbda, radii=None):
    """All terms of lpcm.py:770-856 except the network log-likelihood."""
    K = sigma.shape[0]
    flat = hp.dirichlet_prior * np.ones(K)
    lp = _dirichlet_logpdf(init_weights, flat)
    for k in range(K):
        lp += _dirichlet_logpdf(trans_weights[k], flat)
    return lp + mixture_log_prior(hp, X, intercept, intercept_prior, intercept_variance_prior, mu,
                                  sigma, z, stacked_weights(init_weights, trans_weights, X.shape[0]),
                                  lmbda, radii=radii)


class DynamicNetworkLPCM(_FittedNetworkMixin):
    """Parameters follow the reference estimator (lpcm.py:135-186) plus ``sampler`` / ``device``."""

    def __init__(self, n_features=2, n_components=5, is_directed=False, selection_type="map",
                 n_iter=5000, tune=2500, tune_interval=100, burn=2500, thin=None,
                 intercept_prior="auto", intercept_variance_prior=2, mean_variance_prior="auto",
                 a=2.0, b="auto", lambda_prior=0.9, lambda_variance_prior=0.01,
                 dirichlet_prior="uniform", sigma_prior_std=4.0, mean_variance_prior_std=4.0,
                 step_size_X="auto", step_size_intercept=0.1, step_size_radii=175000,
                 n_control=None, n_resample_control=100, copy=True, random_state=None,
                 sampler="device", device=0):
        for k, v in list(locals().items()):
            if k != "self":
                setattr(self, k, v)

    @property
    def n_burn_(self):
        # lpcm.py:189-197: burn-in counted in STORED samples when the trace is thinned
        nb = (self.burn or 0) + (self.tune or 0)
        return int(np.ceil(nb / self.thin)) if self.thin else nb

    def fit(self, Y):
        if self.sampler not in ("device", "replay"):
            raise ValueError("`sampler` must be 'device' or 'replay', got {}".format(self.sampler))
        replay = self.sampler == "replay"
        T, n, _ = Y.shape
        K, d = self.n_components, self.n_features
        rng = check_random_state(self.random_state)
        Y = check_array(Y, dtype=np.float64, ensure_all_finite="allow-nan", ensure_2d=False,
                        allow_nd=True, copy=self.copy)
        if np.any(Y == -1) or np.any(np.isnan(Y)):
            raise NotImplementedError("missing dyads (-1 / NaN) are not supported by the device sampler")
        self.Y_fit_ = Y
        if self.burn is not None:
            self.n_iter += self.burn
        if self.tune is not None:
            self.n_iter += self.tune
        S, m = self.n_iter, (2 if self.is_directed else 1)

        # ---- init_sampler (lpcm.py:45-131): a short LSM run, k-means on its point estimate ----
        lsm_kw = dict(n_iter=500, n_features=d, tune=250, burn=250, is_directed=self.is_directed,
                      random_state=rng, sampler=self.sampler, device=self.device)
        if self.is_directed:
            lsm_kw.update(sigma_sq=0.001, tau_sq="auto", step_size_X=0.0075,
                          n_control=self.n_control, n_resample_control=self.n_resample_control)
        else:
            lsm_kw.update(sigma_sq=0.1, tau_sq=2.0, step_size_X=0.1)
        emb = DynamicNetworkLSM(**lsm_kw).fit(Y)
        self.lsm_init_ = emb
        Xs = np.zeros((S, T, n, d)); Xs[0] = emb.X_
        ics = np.zeros((S, m)); ics[0] = emb.intercept_
        rads = None
        if self.is_directed:
            rads = np.zeros((S, n)); rads[0] = emb.radii_
        zs = np.zeros((S, T, n), dtype=np.int64)
        mus = np.zeros((S, K, d)); sigmas = np.zeros((S, K))
        mus[0], sigmas[0], zs[0] = longitudinal_kmeans(Xs[0], n_clusters=K, random_state=rng)
        init_w = np.zeros((S, K))
        init_w[0] = np.bincount(zs[0, 0], minlength=K) / n       # lpcm.py:112-115 (k-means labels are constant in t)
        lambdas = np.zeros((S, 1)); lambdas[0] = self.lambda_prior
        trans_w = np.zeros((S, K, K))
        trans_w[0] = 1. / K                                       # lpcm.py:124-128 (uniform rows)

        self.dirichlet_prior_ = 1. if self.dirichlet_prior == "uniform" else 1. / K
        if self.step_size_X == "auto":
            self.step_size_X = 0.01 if self.is_directed else 0.1
        self.case_control_sampler_ = None
        if self.n_control is not None:
            if not self.is_directed:
                raise ValueError("The case-control likelihood currently only "
                                 "supported for directed networks.")
            self.case_control_sampler_ = DirectedCaseControlSampler(
                n_control=self.n_control, n_resample=self.n_resample_control, random_state=rng)
            self.case_control_sampler_.init(Y, sample=replay)
        if isinstance(self.intercept_prior, str) and self.intercept_prior == "auto":
            self.intercept_prior = ics[0]   # (sic) a view of the trace's first row, as in the reference

        # ---- hyper-priors (lpcm.py:436-471) ----
        if self.mean_variance_prior == "auto":
            mvp = (2 * (1. / n) ** (2. / d)) if self.is_directed else ((n ** (2. / d)) / 50.)
        else:
            mvp = self.mean_variance_prior
        a0 = b0 = c0 = d0 = None
        if self.mean_variance_prior_std is not None:
            a0 = (self.mean_variance_prior_std ** 2 + 2) * 2
            b0 = (a0 - 2) * mvp * 2
        b_ = (self.a + 2) * mvp if self.b == "auto" else self.b
        if self.sigma_prior_std is not None:
            d0 = (self.sigma_prior_std ** 2 / b_) * 2
            c0 = b_ * d0
        hp = MixtureHyper(self.dirichlet_prior_, mvp, b_, self.a, a0, b0, c0, d0, self.lambda_prior,
                          self.lambda_variance_prior, self.mean_variance_prior_std is not None,
                          self.sigma_prior_std is not None)
        self.hyper_ = hp
        self.a0_, self.b0_, self.c0_, self.d0_ = a0, b0, c0, d0

        # ---- device state: one chain, the hot path of every sweep ----
        drv = _Driver(Y, d, 1, self.is_directed, self.case_control_sampler_, K, self.tune,
                      self.tune_interval, (100, 100),   # lpcm.py:417-426: default interval
                      self.tune, self.device, replay, rng)   # lpcm.py:428-431: radii sampler tunes
        e = drv.engine
        self._engine = e
        try:
            e.set(L.F_X, Xs[0][None])
            ic = np.zeros((1, 2)); ic[:, :m] = ics[0]
            e.set(L.F_INTERCEPT, ic)
            if self.is_directed:
                e.set(L.F_RADII, rads[0][None])
            e.set_hyper(intercept_prior=np.array(self.intercept_prior, dtype=np.float64),
                        intercept_variance_prior=self.intercept_variance_prior)
            e.set_tuner(self.step_size_X, self.step_size_intercept, self.step_size_radii)

            def push_mixture(it):
                e.set(L.F_MU, mus[it][None]); e.set(L.F_SIGMA, sigmas[it][None])
                e.set(L.F_LAMBDA, lambdas[it])
                e.set(L.F_WEIGHTS, stacked_weights(init_w[it], trans_w[it], T)[None])
                if it == 0 or replay:      # afterwards the labels on the device are the ones just drawn
                    e.set(L.F_Z, zs[it][None])

            def log_post(it):
                lp = lpcm_log_prior(hp, Xs[it], ics[it], self.intercept_prior,
                                    self.intercept_variance_prior, mus[it], sigmas[it], zs[it],
                                    init_w[it], trans_w[it], lambdas[it],
                                    radii=rads[it] if self.is_directed else None)
                return float(np.ravel(e.loglik_full()[0] + lp)[0])

            logps = np.zeros(S)
            push_mixture(0)
            logps[0] = log_post(0)
            cc = self.case_control_sampler_
            for it in range(1, S):                                    # lpcm.py:514-709
                if cc is not None:
                    if replay:
                        cc.resample()
                        if cc.resampled_:
                            drv.push_controls()
                    else:
                        if cc.n_resample is not None and cc.n_iter % cc.n_resample == 0:
                            drv.draw_controls()
                        cc.n_iter += 1
                if replay:
                    drv.sweep_latent()
                    e.center()
                    drv.sample_intercepts()
                    if self.is_directed:
                        drv.sample_radii()
                    drv.sample_labels()
                else:
                    # one call: sweep -> centre -> {intercept / radii MH || label FFBS on the side
                    # stream}; no HDP prior is set on this handle, so the device loop stops there
                    e.run_sweeps(1)
                X = e.get(L.F_X)[0]
                z = e.get(L.F_Z)[0].astype(np.int64)
                cnt = e.get(L.F_NCOUNT)[0]
                nk = e.get(L.F_NK)[0].astype(np.int64)
                mus[it], sigmas[it] = mus[it - 1], sigmas[it - 1]
                init_w[it], trans_w[it] = init_w[it - 1], trans_w[it - 1]
                lambdas[it] = lpcm_conjugate_updates(rng, hp, X, z, cnt, nk, mus[it], sigmas[it],
                                                     lambdas[it - 1].copy(), init_w[it], trans_w[it])
                Xs[it], zs[it] = X, z
                ics[it] = e.get(L.F_INTERCEPT)[0, :m]
                if self.is_directed:
                    rads[it] = e.get(L.F_RADII)[0]
                push_mixture(it)
                logps[it] = log_post(it)
            if cc is not None and not replay:
                ci, co = e.get_controls()
                cc.control_nodes_in_, cc.control_nodes_out_ = ci[0].astype(np.int64), co[0].astype(np.int64)
            self.sampler_counters_ = e.counters()
        finally:
            e.close()            # chain state is not kept after fit
            self._engine = None

        self.mean_variance_prior_, self.b_ = hp.mean_variance_prior, hp.b
        if self.thin is not None:
            sl = slice(None, None, self.thin)
            Xs, ics, mus, sigmas, zs = Xs[sl], ics[sl], mus[sl], sigmas[sl], zs[sl]
            init_w, trans_w, lambdas, logps = init_w[sl], trans_w[sl], lambdas[sl], logps[sl]
            if self.is_directed:
                rads = rads[sl]
        self.Xs_, self.intercepts_, self.mus_, self.sigmas_, self.zs_ = Xs, ics, mus, sigmas, zs
        self.init_weights_, self.trans_weights_, self.lambdas_, self.logps_ = init_w, trans_w, lambdas, logps
        self.radiis_ = rads
        self._post_process()
        return self

    # ---- point estimate, alignment, posterior means (lpcm.py:711-758) -----------------------
    def _post_process(self):
        nb = self.n_burn_
        T, n = self.Y_fit_.shape[:2]
        K = self.n_components
        self.cooccurrence_probas_ = np.zeros((T, n, n))           # label_utils.py:40-62
        eye = np.eye(K, dtype=np.float32)                         # 0/1 indicators: exact in fp32 BLAS
        for t in range(T):
            ind = eye[self.zs_[nb:, t]]
            flat = ind.transpose(1, 0, 2).reshape(n, -1)
            self.cooccurrence_probas_[t] = (flat @ flat.T).astype(np.float64) / ind.shape[0]
        if self.selection_type == "map":
            best = int(np.argmax(self.logps_[nb:]))               # (sic) lpcm.py:716: no burn-in offset
        else:
            loglik = MS._device_loglik(self)
            try:
                best = MS.minimize_posterior_expected_vi(self, loglik)
            finally:
                loglik.engine.close()
        self.selected_id_ = best
        self.logp_ = self.logps_[best]
        self.X_ = self.Xs_[best]
        self.intercept_ = self.intercepts_[best]
        self.lambda_ = self.lambdas_[best]
        if self.is_directed:
            self.radii_ = self.radiis_[best]
        self.z_ = self.zs_[best]
        self.init_weight_ = self.init_weights_[best]
        self.trans_weight_ = self.trans_weights_[best]
        self.mu_ = self.mus_[best]
        self.sigma_ = self.sigmas_[best]
        ref = self.X_.copy()
        for idx in range(self.Xs_.shape[0]):                      # lpcm.py:739-745
            self.Xs_[idx], R = longitudinal_procrustes_rotation(ref, self.Xs_[idx])
            self.mus_[idx] = np.dot(self.mus_[idx], R)
        self.X_mean_ = self.Xs_[nb:].mean(axis=0)
        self.lambda_mean_ = self.lambdas_[nb:].mean(axis=0)
        self.intercepts_mean_ = self.intercepts_[nb:].mean(axis=0)
        if self.is_directed:
            self.radii_mean_ = self.radiis_[nb:].mean(axis=0)

    def logp(self, X, intercept, mu, sigma, z, init_weights, trans_weights, lmbda, radii=None):
        """Joint log-posterior of one state (lpcm.py:770-856); the network term is evaluated by the
        device kernels."""
        e = _state_engine(self, X, intercept, radii)
        try:
            ll = e.loglik_full()[0]
        finally:
            e.close()
        lp = lpcm_log_prior(self.hyper_, X, np.ravel(intercept), self.intercept_prior,
                            self.intercept_variance_prior, mu, sigma, z, init_weights, trans_weights,
                            lmbda, radii=radii)
        return float(np.ravel(ll + lp)[0])

    def forecast_probas(self, n_samples=5000):
        """lpcm.py:229-318 build on forecast.pyx, which is outside the accelerated path."""
        raise _NotOnDevicePath("one-step-ahead forecasts (forecast.pyx) are not part of the device path")

    forecast_probas_map_ = property(forecast_probas)
    forecast_probas_plugin_ = property(forecast_probas)
    forecast_probas_marginalized_ = property(forecast_probas)

    def delete_traces(self):
        """lpcm.py:858-873."""
        for nm in ("Xs_", "intercepts_", "zs_", "mus_", "sigmas_", "init_weights_", "trans_weights_",
                   "lambdas_", "logps_"):
            delattr(self, nm)
        if self.is_directed:
            del self.radiis_


def _state_engine(model, X, intercept, radii):
    T, n, d = X.shape
    cc = model.case_control_sampler_
    e = L.Engine(T=T, n=n, d=d, n_chains=1, is_directed=model.is_directed, case_control=cc is not None,
                 device=model.device)
    if cc is None:
        e.set_network(model.Y_fit_)
    else:
        e.set_edge_lists(cc.degrees_, cc.in_edges_, cc.out_edges_)
        e.set_controls(cc.control_nodes_in_, cc.control_nodes_out_)
    e.set(L.F_X, X[None])
    ic = np.zeros((1, 2))
    ic[0, :np.size(intercept)] = np.ravel(intercept)
    e.set(L.F_INTERCEPT, ic)
    if model.is_directed:
        e.set(L.F_RADII, np.asarray(radii)[None])
    return e
