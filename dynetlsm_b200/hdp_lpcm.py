"""``DynamicNetworkHDPLPCM`` -- the reference estimator's API (hdp_lpcm.py:144-1330) over the
device sampler.

Per sweep the device runs the hot path -- mixture-prior latent-position sweep, centring,
intercept / radii MH on the full-network likelihood, HDP-HMM label forward-filter /
backward-sample -- and hands the label counts to the host, which performs the reference's
conjugate and auxiliary-variable updates in numpy (``hdp_updates.py``) when ``sampler='replay'``:
every random number then comes from the numpy ``RandomState`` in the reference's order, so the
chain reproduces the reference's.  With ``sampler='device'`` (default) that block, the joint
log-posterior and the stored samples are produced on the device as well (``dlsm_hdp_update``,
``dlsm_run_traced``) and ``fit`` makes no host round trip per sweep.

The point estimate follows the reference's ``selection_type`` (hdp_lpcm.py:1089-1138): 'vi'
(default) minimises the posterior-expected variation of information over the co-clustering
probabilities -- which the device accumulates while sampling --, 'bic' / 'map' go through the
approximate BIC per model size (``model_selection.py``; the network log-likelihoods it needs are
evaluated by the device kernels).  ``bic_``, ``models_``, ``counts_``, ``best_k_``,
``posterior_group_ids_`` / ``posterior_group_counts_``, co-clustering probabilities, posterior
means and the Procrustes alignment of the traces are produced as in the reference
(hdp_lpcm.py:1085-1170, label_utils.py:40-82).
"""
import numpy as np
from sklearn.utils import check_array, check_random_state

from . import _lib as L
from .case_control_likelihood import DirectedCaseControlSampler
from .hdp_updates import HDPHyper, conjugate_updates, hdp_log_prior
from .host_init import longitudinal_kmeans, longitudinal_procrustes_rotation
from .lsm import DynamicNetworkLSM, _Driver, _FittedNetworkMixin
from . import model_selection as MS

__all__ = ["DynamicNetworkHDPLPCM"]


class DynamicNetworkHDPLPCM(_FittedNetworkMixin):
    def __init__(self, n_features=2, n_components=10, is_directed=False, selection_type="vi",
                 n_iter=5000, tune=2500, tune_interval=100, burn=2500, thin=None, gamma=1.0,
                 gamma_prior_shape=1.0, gamma_prior_rate=0.1, alpha_init=1.0, alpha_init_shape=1.,
                 alpha_init_rate=1., alpha=1.0, kappa=4.0, alpha_kappa_shape=5,
                 alpha_kappa_rate=0.1, intercept_prior="auto", intercept_variance_prior=2,
                 mean_variance_prior="auto", a=2.0, b="auto", lambda_prior=0.9,
                 lambda_variance_prior=0.01, sigma_prior_std=4.0, mean_variance_prior_std=4.0,
                 step_size_X="auto", step_size_intercept=0.1, step_size_radii=175000,
                 n_control=None, n_resample_control=100, copy=True, random_state=None,
                 sampler="device", n_chains=1, device=0):
        for k, v in list(locals().items()):
            if k != "self":
                setattr(self, k, v)

    @property
    def n_burn_(self):
        # hdp_lpcm.py:458-465: burn-in counted in STORED samples when the trace is thinned
        nb = (self.burn or 0) + (self.tune or 0)
        return int(np.ceil(nb / self.thin)) if self.thin else nb

    def fit(self, Y):
        replay = self.sampler == "replay"
        if self.sampler not in ("device", "replay"):
            raise ValueError("`sampler` must be 'device' or 'replay', got {}".format(self.sampler))
        if self.selection_type not in ("vi", "bic", "map"):
            raise ValueError("Selection type not recognized")       # hdp_lpcm.py:1122
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

        # ---- init_sampler (hdp_lpcm.py:48-141): a short LSM run, then k-means on its MAP ----
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
        weights = np.zeros((S, T, K, K))
        weights[0, 0, 0] = np.bincount(zs[0, 0], minlength=K) / n
        lambdas = np.zeros((S, 1)); lambdas[0] = self.lambda_prior
        betas = np.zeros((S, K))
        betas[0] = rng.dirichlet(np.repeat(self.gamma / K, K))
        for t in range(1, T):
            for k in range(K):
                weights[0, t, k] = rng.dirichlet(self.alpha * betas[0] + 4.0 * np.eye(K)[k])  # init_sampler default kappa=4.0 (hdp_lpcm.py:50)

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

        # ---- hyper-priors (hdp_lpcm.py:750-793) ----
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
        hp = HDPHyper(self.gamma, self.alpha_init, self.alpha, self.kappa, mvp, b_, self.a, a0, b0,
                      c0, d0, self.lambda_prior, self.lambda_variance_prior, self.gamma_prior_shape,
                      self.gamma_prior_rate, self.alpha_init_shape, self.alpha_init_rate,
                      self.alpha_kappa_shape, self.alpha_kappa_rate,
                      self.mean_variance_prior_std is not None, self.sigma_prior_std is not None)
        self.hyper_ = hp

        # ---- device state ----
        C = 1 if replay else int(self.n_chains)
        if replay and self.n_chains != 1:
            raise ValueError("sampler='replay' reproduces one reference chain; use n_chains=1")
        drv = _Driver(Y, d, C, self.is_directed, self.case_control_sampler_, K, self.tune,
                      self.tune_interval, (100, 100),   # hdp_lpcm.py:735-742: default interval
                      self.tune, self.device, replay, rng)   # hdp_lpcm.py:745-747: radii sampler tunes
        e = drv.engine
        self._engine = e
        tile = lambda a: np.tile(np.asarray(a)[None], (C,) + (1,) * np.ndim(a))
        e.set(L.F_X, tile(Xs[0]))
        ic = np.zeros((C, 2)); ic[:, :m] = ics[0]
        e.set(L.F_INTERCEPT, ic)
        if self.is_directed:
            e.set(L.F_RADII, tile(rads[0]))
        e.set_hyper(intercept_prior=np.array(self.intercept_prior, dtype=np.float64),
                    intercept_variance_prior=self.intercept_variance_prior)
        e.set_tuner(self.step_size_X, self.step_size_intercept, self.step_size_radii)

        def push_mixture(it):
            e.set(L.F_MU, mus[it][None]); e.set(L.F_SIGMA, sigmas[it][None])
            e.set(L.F_LAMBDA, lambdas[it]); e.set(L.F_WEIGHTS, weights[it][None])
            e.set(L.F_Z, zs[it][None])

        def log_post(ll, hpc, X, b, mu, sigma, z, w, beta, lmbda, r):
            lp = hdp_log_prior(hpc, K, X, b, self.intercept_prior, self.intercept_variance_prior,
                               mu, sigma, z, w, beta, lmbda, radii=r)
            return float(np.ravel(ll + lp)[0])

        logps = np.zeros(S)
        if replay:
            push_mixture(0)
            logps[0] = log_post(e.loglik_full()[0], hp, Xs[0], ics[0], mus[0], sigmas[0], zs[0],
                                weights[0], betas[0], lambdas[0], rads[0] if self.is_directed else None)
            for it in range(1, S):
                if self.case_control_sampler_ is not None:
                    self.case_control_sampler_.resample()
                    if self.case_control_sampler_.resampled_:
                        drv.push_controls()
                drv.sweep_latent()
                e.center()
                drv.sample_intercepts()
                if self.is_directed:
                    drv.sample_radii()
                drv.sample_labels()
                X = e.get(L.F_X)[0]
                z = e.get(L.F_Z)[0].astype(np.int64)
                cnt = e.get(L.F_NCOUNT)[0]
                nk = e.get(L.F_NK)[0].astype(np.int64)
                mu, sigma, w = mus[it - 1].copy(), sigmas[it - 1].copy(), weights[it - 1].copy()
                beta, lmbda = conjugate_updates(rng, hp, X, z, cnt, nk, mu, sigma,
                                                lambdas[it - 1].copy(), betas[it - 1].copy(), w)
                Xs[it], zs[it], mus[it], sigmas[it] = X, z, mu, sigma
                betas[it], weights[it], lambdas[it] = beta, w, lmbda
                ics[it] = e.get(L.F_INTERCEPT)[0, :m]
                if self.is_directed:
                    rads[it] = e.get(L.F_RADII)[0]
                push_mixture(it)
                logps[it] = log_post(e.loglik_full()[0], hp, X, ics[it], mu, sigma, z, w, beta, lmbda,
                                     rads[it] if self.is_directed else None)
            chains = None
        else:
            # ---- device-resident chains: the conjugate block runs in k_hdp_update ----
            e.set(L.F_MU, tile(mus[0])); e.set(L.F_SIGMA, tile(sigmas[0]))
            e.set(L.F_LAMBDA, np.full(C, float(self.lambda_prior))); e.set(L.F_WEIGHTS, tile(weights[0]))
            e.set(L.F_Z, tile(zs[0])); e.set(L.F_BETA, tile(betas[0]))
            hy0 = np.array([hp.gamma, hp.alpha_init, hp.alpha, hp.kappa, hp.mean_variance_prior, hp.b, 0, 0],
                           dtype=np.float64)
            e.set(L.F_HYPER, tile(hy0))
            e.set_hdp_prior(hp.a, hp.a0, hp.b0, hp.c0, hp.d0, hp.lambda_prior, hp.lambda_variance_prior,
                            hp.gamma_prior_shape, hp.gamma_prior_rate, hp.alpha_init_shape,
                            hp.alpha_init_rate, hp.alpha_kappa_shape, hp.alpha_kappa_rate,
                            hp.resample_mean_variance, hp.resample_b)
            chains = dict(intercepts=np.zeros((C, S, m)), lambdas=np.zeros((C, S)), logps=np.zeros((C, S)),
                          n_clusters=np.zeros((C, S), dtype=np.int64), zs=np.zeros((C, S, T, n), np.int16))
            chains["intercepts"][:, 0] = ics[0]; chains["lambdas"][:, 0] = self.lambda_prior
            chains["zs"][:, 0] = zs[0]
            chains["n_clusters"][:, 0] = np.unique(zs[0]).size
            chains["logps"][:, 0] = e.logp()
            logps[0] = chains["logps"][0, 0]

            # The whole loop body (hdp_lpcm.py:823-1069) runs on the device, stored samples and the
            # joint log-posterior included (dlsm_run_traced); the host only splits the run where
            # the case-control sets are redrawn and to bound the size of one trace transfer.
            first = (L.F_X, L.F_MU, L.F_SIGMA, L.F_BETA, L.F_WEIGHTS) + ((L.F_RADII,) if self.is_directed else ())
            every = (L.F_INTERCEPT, L.F_LAMBDA, L.F_Z, L.F_HYPER)
            rec_bytes = 8 * (T * n * d + K * d + 2 * K + T * K * K + n) + C * (4 * T * n + 8 * 12)
            seg = int(max(1, min(S, (128 << 20) // rec_bytes)))   # bounded, reused pinned destination
            tr = None
            cc = self.case_control_sampler_
            hya = None
            it = 1
            while it < S:
                stop = min(S, it + seg)
                if cc is not None:              # hdp_lpcm.py:826-829
                    if cc.n_resample is not None and cc.n_iter % cc.n_resample == 0:
                        drv.draw_controls()
                    cc.n_iter += 1
                    quiet = S if cc.n_resample is None else (cc.n_resample - cc.n_iter % cc.n_resample) % cc.n_resample
                    stop = min(stop, it + 1 + quiet)
                nb_ = min((self.burn or 0) + (self.tune or 0), S - 1)
                tr = e.run_traced(stop - it, fields_all=every, fields_first=first, pinned=True, out=tr,
                                  cooc=1, cooc_from=max(0, nb_ - it))   # co-clustering counts of chain 0
                if cc is not None:
                    cc.n_iter += stop - it - 1
                sl = slice(it, stop)
                Xs[sl], mus[sl], sigmas[sl] = tr[L.F_X][:, 0], tr[L.F_MU][:, 0], tr[L.F_SIGMA][:, 0]
                betas[sl], weights[sl] = tr[L.F_BETA][:, 0], tr[L.F_WEIGHTS][:, 0]
                za = tr[L.F_Z]
                zs[sl] = za[:, 0]
                lambdas[sl, 0] = tr[L.F_LAMBDA][:, 0]
                ics[sl] = tr[L.F_INTERCEPT][:, 0, :m]
                if self.is_directed:
                    rads[sl] = tr[L.F_RADII][:, 0]
                logps[sl] = tr["logp"][:, 0]
                chains["logps"][:, sl] = tr["logp"].T
                chains["intercepts"][:, sl] = tr[L.F_INTERCEPT].transpose(1, 0, 2)[:, :, :m]
                chains["lambdas"][:, sl] = tr[L.F_LAMBDA].T
                chains["zs"][:, sl] = za.transpose(1, 0, 2, 3)
                present = np.zeros(za.shape[:2] + (K,), dtype=bool)
                np.put_along_axis(present, za.reshape(za.shape[0], C, -1), True, axis=2)
                chains["n_clusters"][:, sl] = present.sum(axis=2).T
                hya = tr[L.F_HYPER][-1, 0]
                it = stop
            if hya is not None:
                (hp.gamma, hp.alpha_init, hp.alpha, hp.kappa, hp.mean_variance_prior, hp.b) = hya[:6]
            if self.thin is None and min(self.n_burn_, S - 1) >= 1:
                self._device_cooc = e.cooccurrence(reset=True)   # (counts, samples) from the device
            if cc is not None:
                ci, co = e.get_controls()
                cc.control_nodes_in_, cc.control_nodes_out_ = ci[0].astype(np.int64), co[0].astype(np.int64)
        self.chains_ = chains

        # mirror the reference's mutable hyper-parameter attributes
        self.gamma, self.alpha_init, self.alpha, self.kappa = hp.gamma, hp.alpha_init, hp.alpha, hp.kappa
        self.mean_variance_prior_, self.b_ = hp.mean_variance_prior, hp.b

        if self.thin is not None:
            sl = slice(None, None, self.thin)
            Xs, ics, mus, sigmas, zs = Xs[sl], ics[sl], mus[sl], sigmas[sl], zs[sl]
            betas, weights, lambdas, logps = betas[sl], weights[sl], lambdas[sl], logps[sl]
            if self.is_directed:
                rads = rads[sl]
        self.Xs_, self.intercepts_, self.mus_, self.sigmas_, self.zs_ = Xs, ics, mus, sigmas, zs
        self.betas_, self.weights_, self.lambdas_, self.logps_ = betas, weights, lambdas, logps
        self.radiis_ = rads
        self._post_process()
        self.sampler_counters_ = e.counters()
        e.close()            # chain state, trace rings and pinned buffers are not kept after fit
        self._engine = None
        return self

    # ---- point estimate, alignment, posterior summaries -------------------------------------
    def _post_process(self):
        nb = min(self.n_burn_, self.Xs_.shape[0] - 1)
        T, n = self.Y_fit_.shape[:2]
        K = self.n_components
        # co-clustering probabilities over the post-burn-in draws (label_utils.py:40-62)
        self.cooccurrence_probas_ = np.zeros((T, n, n))
        dev = getattr(self, "_device_cooc", None)
        if dev is not None and dev[1] == self.zs_.shape[0] - nb:
            # accumulated on the device while sampling (dlsm_run_traced cooc_mode)
            self.cooccurrence_probas_ = dev[0].astype(np.float64) / dev[1]
        else:
            eye = np.eye(K, dtype=np.float32)                    # 0/1 indicators: exact in fp32 BLAS
            for t in range(T):
                ind = eye[self.zs_[nb:, t]]                      # (S', n, K)
                flat = ind.transpose(1, 0, 2).reshape(n, -1)     # (n, S' K): one GEMM per time step
                self.cooccurrence_probas_[t] = (flat @ flat.T).astype(np.float64) / ind.shape[0]
        self._device_cooc = None

        # model sizes and their approximate BIC (hdp_lpcm.py:1089-1090), then the point estimate
        loglik = MS._device_loglik(self)
        try:
            self.bic_, self.models_, self.counts_ = MS.select_bic(self, loglik)
            if self.selection_type == "vi":
                best = MS.minimize_posterior_expected_vi(self, loglik)
        finally:
            loglik.engine.close()
        if self.selection_type == "vi":
            self.selected_id_ = best
            self.logp_ = self.logps_[best]
            self.X_ = self.Xs_[best]
            self.intercept_ = self.intercepts_[best]
            self.lambda_ = self.lambdas_[best]
            if self.is_directed:
                self.radii_ = self.radiis_[best]
            (self.z_, self.beta_, self.init_weights_, self.trans_weights_, self.mu_,
             self.sigma_) = MS.renormalized(self, best)
        else:
            if self.selection_type == "bic":
                mid = int(np.argmin(self.bic_[:, 1]))
                self.best_k_ = int(self.bic_[mid, 0])
            else:                                               # 'map': the most frequent model size
                self.best_k_ = int(np.argmax(np.bincount(self.counts_)))
                mid = int(np.argwhere(self.bic_[:, 0] == self.best_k_)[0, 0])
            mk = self.models_[mid]
            self.selected_id_ = int(self.bic_[mid, 3])
            self.logp_ = self.logps_[self.selected_id_]
            self.X_, self.intercept_, self.mu_, self.sigma_ = mk.X, mk.intercept, mk.mu, mk.sigma
            if self.is_directed:
                self.radii_ = mk.radii
            self.z_ = np.unique(mk.z.ravel(), return_inverse=True)[1].reshape(T, n)
            self.beta_, self.init_weights_, self.trans_weights_ = mk.beta, mk.init_weights, mk.trans_weights
            self.lambda_ = mk.lmbda
        # rotate every stored sample onto the point estimate (hdp_lpcm.py:1141-1146)
        ref = self.X_.copy()
        for idx in range(self.Xs_.shape[0]):
            self.Xs_[idx], R = longitudinal_procrustes_rotation(ref, self.Xs_[idx])
            self.mus_[idx] = np.dot(self.mus_[idx], R)
        self.X_mean_ = self.Xs_[nb:].mean(axis=0)
        self.lambda_mean_ = self.lambdas_[nb:].mean(axis=0)
        self.intercepts_mean_ = self.intercepts_[nb:].mean(axis=0)
        if self.is_directed:
            self.radii_mean_ = self.radiis_[nb:].mean(axis=0)
        # posterior distribution of the number of groups per time step (hdp_lpcm.py:1161-1167)
        ct = MS.cluster_counts_t(self.zs_[nb:], K)
        self.posterior_group_ids_, self.posterior_group_counts_ = [], []
        for t in range(T):
            ids, freq = MS.posterior_group_counts(ct[t])
            self.posterior_group_ids_.append(ids)
            self.posterior_group_counts_.append(freq)
        # Geweke diagnostics as (z, p) pairs (hdp_lpcm.py:1169-1185, trace_utils.py:59-115)
        from .diagnostics import geweke_zp
        self.logp_geweke_ = geweke_zp(self.logps_, nb)
        self.lambda_geweke_ = geweke_zp(self.lambdas_[:, 0], nb)
        if self.is_directed:
            self.intercept_in_geweke_ = geweke_zp(self.intercepts_[:, 0], nb)
            self.intercept_out_geweke_ = geweke_zp(self.intercepts_[:, 1], nb)
        else:
            self.intercept_geweke_ = geweke_zp(self.intercepts_[:, 0], nb)
