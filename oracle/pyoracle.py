"""ctypes wrapper around oracle/liboracle.so (the CPU restatement in oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
The product package (dynetlsm_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(
                os.path.join(HERE, "oracle.c")):
            build()
        L = C.CDLL(path)
        L.orc_np_sum.restype = C.c_double
        L.orc_partial_loglikelihood.restype = C.c_double
        L.orc_directed_partial_loglikelihood.restype = C.c_double
        L.orc_approx_directed_partial_loglikelihood.restype = C.c_double
        L.orc_directed_network_loglikelihood.restype = C.c_double
        L.orc_undirected_network_loglikelihood.restype = C.c_double
        L.orc_approx_directed_network_loglikelihood.restype = C.c_double
        L.orc_dirichlet_logpdf.restype = C.c_double
        _LIB = L
    return _LIB


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _lp(a):
    return a.ctypes.data_as(c_lp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def np_sum(a):
    a = _f64(a).ravel()
    return lib().orc_np_sum(_dp(a), C.c_int64(a.size))


def partial_loglikelihood(Y, X, intercept, node_id, squared=False):
    Y, X = _f64(Y), _f64(X)
    return lib().orc_partial_loglikelihood(_dp(Y), _dp(X), X.shape[0], X.shape[1],
                                           C.c_double(float(intercept)), int(node_id), int(squared))


def directed_partial_loglikelihood(Y, X, radii, intercept_in, intercept_out, node_id,
                                   squared=False):
    Y, X, radii = _f64(Y), _f64(X), _f64(radii)
    return lib().orc_directed_partial_loglikelihood(
        _dp(Y), _dp(X), _dp(radii), X.shape[0], X.shape[1], C.c_double(float(intercept_in)),
        C.c_double(float(intercept_out)), int(node_id), int(squared))


def approx_directed_partial_loglikelihood(X, radii, in_edges, out_edges, degree, control_nodes_in,
                                          control_nodes_out, intercept_in, intercept_out, node_id,
                                          squared=False, return_ub=False):
    X, radii = _f64(X), _f64(radii)
    ie, oe, dg = _i64(in_edges), _i64(out_edges), _i64(degree)
    ci, co = _i64(control_nodes_in), _i64(control_nodes_out)
    ub = C.c_int(0)
    v = lib().orc_approx_directed_partial_loglikelihood(
        _dp(X), _dp(radii), _lp(ie), ie.shape[1], _lp(oe), oe.shape[1], _lp(dg), _lp(ci), _lp(co),
        ci.shape[1], X.shape[0], X.shape[1], C.c_double(float(intercept_in)),
        C.c_double(float(intercept_out)), int(node_id), int(squared), C.byref(ub))
    return (v, ub.value) if return_ub else v


def calculate_distances(X, squared=False):
    X = _f64(X)
    T, n, d = X.shape
    dist = np.empty((T, n, n))
    lib().orc_calculate_distances(_dp(X), T, n, d, int(squared), _dp(dist))
    return dist


def directed_network_loglikelihood(Y, dist, radii, intercept_in, intercept_out):
    Y, dist, radii = _f64(Y), _f64(dist), _f64(radii)
    T, n, _ = Y.shape
    return lib().orc_directed_network_loglikelihood(
        _dp(Y), _dp(dist), _dp(radii), T, n, C.c_double(float(intercept_in)),
        C.c_double(float(intercept_out)))


def undirected_network_loglikelihood(Y, dist, intercept):
    Y, dist = _f64(Y), _f64(dist)
    T, n, _ = Y.shape
    return lib().orc_undirected_network_loglikelihood(_dp(Y), _dp(dist), T, n,
                                                      C.c_double(float(intercept)))


def approx_directed_network_loglikelihood(X, radii, out_edges, degree, control_nodes,
                                          intercept_in, intercept_out, squared=False):
    X, radii = _f64(X), _f64(radii)
    oe, dg, cn = _i64(out_edges), _i64(degree), _i64(control_nodes)
    T, n, d = X.shape
    return lib().orc_approx_directed_network_loglikelihood(
        _dp(X), _dp(radii), _lp(oe), oe.shape[2], _lp(dg), _lp(cn), cn.shape[2], T, n, d,
        C.c_double(float(intercept_in)), C.c_double(float(intercept_out)), int(squared))


def compute_gaussian_likelihood(X, mu, sigma, lmbda, normalize=True):
    X, mu, sigma = _f64(X), _f64(mu), _f64(sigma)
    T, d = X.shape
    K = sigma.shape[0]
    out = np.empty((T, K))
    lib().orc_compute_gaussian_likelihood(_dp(X), _dp(mu), _dp(sigma), C.c_double(float(lmbda)),
                                          T, K, d, int(normalize), _dp(out))
    return out


def dirichlet_logpdf(x, alpha):
    x, alpha = _f64(x), _f64(alpha)
    return lib().orc_dirichlet_logpdf(_dp(x), _dp(alpha), x.shape[0])


class TunerState(object):
    """SoA mirror of a grid of reference ``Metropolis`` objects (metropolis.py:85-94)."""

    def __init__(self, shape, step_size, tune, tune_interval):
        self.step = np.full(shape, float(step_size), dtype=np.float64)
        self.n_accepted = np.zeros(shape, dtype=np.int32)
        self.n_steps = np.zeros(shape, dtype=np.int32)
        self.until = np.full(shape, int(tune_interval), dtype=np.int32)
        self.tune = -1 if tune is None else int(tune)
        self.tune_interval = int(tune_interval)

    def copy(self):
        o = TunerState.__new__(TunerState)
        o.step, o.n_accepted = self.step.copy(), self.n_accepted.copy()
        o.n_steps, o.until = self.n_steps.copy(), self.until.copy()
        o.tune, o.tune_interval = self.tune, self.tune_interval
        return o


class _SweepArgs(C.Structure):
    _fields_ = [
        ("T", C.c_int32), ("n", C.c_int32), ("d", C.c_int32), ("K", C.c_int32),
        ("is_directed", C.c_int32), ("use_cc", C.c_int32), ("prior_kind", C.c_int32),
        ("squared", C.c_int32),
        ("max_in", C.c_int32), ("max_out", C.c_int32), ("n_control", C.c_int32),
        ("tune", C.c_int32), ("tune_interval", C.c_int32), ("pad0", C.c_int32),
        ("tau_sq", C.c_double), ("sigma_sq", C.c_double), ("lmbda", C.c_double),
        ("Y", c_dp), ("X", c_dp), ("intercept", c_dp), ("radii", c_dp),
        ("in_edges", c_lp), ("out_edges", c_lp), ("degrees", c_lp), ("ctrl_in", c_lp),
        ("ctrl_out", c_lp),
        ("mu", c_dp), ("sigma", c_dp), ("z", c_lp),
        ("step", c_dp), ("n_accepted", c_ip), ("n_steps", c_ip), ("until", c_ip),
        ("eps", c_dp), ("logu", c_dp),
        ("accepted", c_ip), ("ratio", c_dp), ("logp_new", c_dp), ("logp_old", c_dp),
    ]


def sweep_latent(X, intercept, tuner, eps, logu, Y=None, radii=None, is_directed=False,
                 tau_sq=2.0, sigma_sq=0.1, mixture=None, case_control=None, squared=False):
    """One replay-driven latent-position sweep (sample_latent_positions.py:92-206).

    ``mixture`` = dict(mu, sigma, lmbda, z) selects the HDP-LPCM prior; ``case_control`` =
    dict(in_edges, out_edges, degrees, ctrl_in, ctrl_out) selects the case-control likelihood.
    X and ``tuner`` are updated in place.  Returns dict(accepted, ratio, logp_new, logp_old).
    """
    assert X.dtype == np.float64 and X.flags.c_contiguous
    T, n, d = X.shape
    keep = []

    def hold(a):
        keep.append(a)
        return a
    a = _SweepArgs()
    a.T, a.n, a.d = T, n, d
    a.is_directed, a.squared = int(is_directed), int(squared)
    a.tune, a.tune_interval = tuner.tune, tuner.tune_interval
    a.tau_sq, a.sigma_sq = float(tau_sq), float(sigma_sq)
    a.X = _dp(X)
    a.intercept = _dp(hold(_f64(np.atleast_1d(intercept))))
    if Y is not None:
        a.Y = _dp(hold(_f64(Y)))
    if radii is not None:
        a.radii = _dp(hold(_f64(radii)))
    if case_control is not None:
        a.use_cc = 1
        ie, oe = hold(_i64(case_control["in_edges"])), hold(_i64(case_control["out_edges"]))
        a.in_edges, a.out_edges = _lp(ie), _lp(oe)
        a.max_in, a.max_out = ie.shape[2], oe.shape[2]
        a.degrees = _lp(hold(_i64(case_control["degrees"])))
        ci, co = hold(_i64(case_control["ctrl_in"])), hold(_i64(case_control["ctrl_out"]))
        a.ctrl_in, a.ctrl_out, a.n_control = _lp(ci), _lp(co), ci.shape[2]
    if mixture is not None:
        a.prior_kind = 1
        mu, sg = hold(_f64(mixture["mu"])), hold(_f64(mixture["sigma"]))
        a.mu, a.sigma, a.K = _dp(mu), _dp(sg), sg.shape[0]
        a.lmbda = float(np.ravel(mixture["lmbda"])[0])
        a.z = _lp(hold(_i64(mixture["z"])))
    for nm in ("step", "n_accepted", "n_steps", "until"):
        arr = getattr(tuner, nm)
        assert arr.flags.c_contiguous and arr.shape == (T, n)
    a.step, a.n_accepted = _dp(tuner.step), _ip(tuner.n_accepted)
    a.n_steps, a.until = _ip(tuner.n_steps), _ip(tuner.until)
    eps, logu = hold(_f64(eps)), hold(_f64(logu))
    assert eps.shape == (T, n, d) and logu.shape == (T, n)
    a.eps, a.logu = _dp(eps), _dp(logu)
    out = dict(accepted=np.zeros((T, n), np.int32), ratio=np.zeros((T, n)),
               logp_new=np.zeros((T, n)), logp_old=np.zeros((T, n)))
    a.accepted, a.ratio = _ip(out["accepted"]), _dp(out["ratio"])
    a.logp_new, a.logp_old = _dp(out["logp_new"]), _dp(out["logp_old"])
    rc = lib().orc_sweep_latent(C.byref(a))
    if rc != 0:
        raise RuntimeError("orc_sweep_latent failed: %d" % rc)
    return out


def center(X):
    assert X.dtype == np.float64 and X.flags.c_contiguous
    T, n, d = X.shape
    lib().orc_center(_dp(X), T, n, d)
    return X


class _InterceptArgs(C.Structure):
    _fields_ = [
        ("T", C.c_int32), ("n", C.c_int32), ("d", C.c_int32), ("is_directed", C.c_int32),
        ("use_cc", C.c_int32), ("max_out", C.c_int32), ("n_control", C.c_int32),
        ("tune", C.c_int32),
        ("tune_interval", C.c_int32 * 2),
        ("prior_mean", C.c_double * 2), ("prior_var", C.c_double),
        ("Y", c_dp), ("X", c_dp), ("dist", c_dp), ("radii", c_dp),
        ("out_edges", c_lp), ("degrees", c_lp), ("ctrl_out", c_lp),
        ("intercept", c_dp), ("step", c_dp), ("n_accepted", c_ip), ("n_steps", c_ip),
        ("until", c_ip),
        ("eps", c_dp), ("logu", c_dp), ("accepted", c_ip), ("ratio", c_dp),
        ("loglik_new", c_dp), ("loglik_old", c_dp),
    ]


def _fill_full(a, keep, X, Y, dist, radii, is_directed, case_control):
    T, n, d = X.shape
    a.T, a.n, a.d, a.is_directed = T, n, d, int(is_directed)
    keep.append(X)
    a.X = _dp(X)
    if Y is not None:
        Y = _f64(Y); keep.append(Y); a.Y = _dp(Y)
    if dist is not None:
        dist = _f64(dist); keep.append(dist); a.dist = _dp(dist)
    if radii is not None:
        radii = _f64(radii); keep.append(radii); a.radii = _dp(radii)
    if case_control is not None:
        a.use_cc = 1
        oe = _i64(case_control["out_edges"]); dg = _i64(case_control["degrees"])
        co = _i64(case_control["ctrl_out"])
        keep.extend([oe, dg, co])
        a.out_edges, a.degrees, a.ctrl_out = _lp(oe), _lp(dg), _lp(co)
        a.max_out, a.n_control = oe.shape[2], co.shape[2]


def sample_intercepts(X, intercept, tuner, eps, logu, prior_mean, prior_var, Y=None, dist=None,
                      radii=None, is_directed=False, case_control=None):
    """Replay-driven sample_intercepts (sample_coefficients.py:12-88).  ``intercept`` (1,)/(2,)
    and ``tuner`` (shape (m,), per-sampler tune_interval list in ``tuner.intervals``) are
    updated in place."""
    X = _f64(X)
    keep = []
    a = _InterceptArgs()
    _fill_full(a, keep, X, Y, dist, radii, is_directed, case_control)
    m = 2 if is_directed else 1
    assert intercept.dtype == np.float64 and intercept.shape == (m,)
    a.tune = tuner.tune
    iv = getattr(tuner, "intervals", [tuner.tune_interval] * m)
    pm = np.atleast_1d(np.asarray(prior_mean, dtype=np.float64))
    for i in range(m):
        a.tune_interval[i] = int(iv[i])
        a.prior_mean[i] = float(pm[i])
    a.prior_var = float(prior_var)
    a.intercept = _dp(intercept)
    a.step, a.n_accepted = _dp(tuner.step), _ip(tuner.n_accepted)
    a.n_steps, a.until = _ip(tuner.n_steps), _ip(tuner.until)
    eps, logu = _f64(np.atleast_1d(eps)), _f64(np.atleast_1d(logu))
    a.eps, a.logu = _dp(eps), _dp(logu)
    out = dict(accepted=np.zeros(m, np.int32), ratio=np.zeros(m), loglik_new=np.zeros(m),
               loglik_old=np.zeros(m))
    a.accepted, a.ratio = _ip(out["accepted"]), _dp(out["ratio"])
    a.loglik_new, a.loglik_old = _dp(out["loglik_new"]), _dp(out["loglik_old"])
    lib().orc_sample_intercepts(C.byref(a))
    return out


def sample_radii(X, intercept, radii, tuner, proposal, logu, Y=None, dist=None, case_control=None):
    """Replay-driven sample_radii (sample_coefficients.py:91-121); ``radii`` updated in place,
    ``tuner`` has shape (1,)."""
    X = _f64(X)
    keep = []
    a = _InterceptArgs()
    _fill_full(a, keep, X, Y, dist, None, True, case_control)
    ic = _f64(intercept)
    a.intercept = _dp(ic)
    proposal = _f64(proposal)
    acc = C.c_int32(0)
    ratio = C.c_double(0)
    assert radii.dtype == np.float64 and radii.flags.c_contiguous
    lib().orc_sample_radii(C.byref(a), _dp(radii), _dp(proposal), C.c_double(float(logu)),
                           _dp(tuner.step), _ip(tuner.n_accepted), _ip(tuner.n_steps),
                           _ip(tuner.until), tuner.tune, tuner.tune_interval, C.byref(acc),
                           C.byref(ratio))
    return dict(accepted=acc.value, ratio=ratio.value)


def sample_labels_block(X, mu, sigma, lmbda, w, U, return_probas=False):
    """Replay-driven sample_labels_block (sample_labels.py:134-190).  U: (n, T) raw uniforms."""
    X, mu, sigma, w, U = _f64(X), _f64(mu), _f64(sigma), _f64(w), _f64(U)
    T, n, d = X.shape
    K = sigma.shape[0]
    assert U.shape == (n, T) and w.shape == (T, K, K)
    z = np.zeros((T, n), np.int64)
    nc = np.zeros((T, K, K))
    nk = np.zeros((T, K), np.int64)
    pr = np.zeros((n, T, K)) if return_probas else None
    lib().orc_sample_labels_block(_dp(X), _dp(mu), _dp(sigma),
                                  C.c_double(float(np.ravel(lmbda)[0])), _dp(w), _dp(U), T, n, d,
                                  K, _lp(z), _dp(nc), _lp(nk), _dp(pr))
    resp = np.zeros((T, n, K), np.int64)
    resp[np.arange(T)[:, None], np.arange(n)[None, :], z] = 1
    if return_probas:
        return z, nc, nk, resp, pr
    return z, nc, nk, resp
