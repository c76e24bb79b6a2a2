"""ctypes binding of libdlsm.so (include/dlsm.h) and a thin numpy-facing ``Engine``.

There is no CPU fallback: importing this module works anywhere (so host logic can be tested on
CPU), but creating an ``Engine`` without the built extension or without a CUDA device raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("DLSM_LIB") or os.path.join(HERE, "libdlsm.so")   # DLSM_LIB: A/B builds (tools/)
CSRC = os.path.join(HERE, "csrc")
# translation units -> the headers each depends on (mtime-based rebuild of the unit's object file)
HEADERS = [os.path.join(CSRC, f) for f in ("dlsm_kernels.cuh", "dlsm_device.cuh", "dlsm_tables.cuh",
                                            "dlsm_hdp.cuh", "dlsm_trace.cuh", "dlsm_blk.h", "dlsm_graph.h", "dlsm_cc.h", "dlsm_ccd.h", "dlsm_dyad.cuh", "dlsm_fullr.h")]
HEADERS.append(os.path.join(ROOT, "include", "dlsm.h"))
UNITS = [os.path.join(CSRC, f) for f in ("dlsm.cu", "dlsm_blk.cu", "dlsm_graph.cu", "dlsm_cc.cu", "dlsm_ccd.cu", "dlsm_cbp.cu", "dlsm_fullr.cu")]
SRC = UNITS + HEADERS
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


class DlsmError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libdlsm error %d: %s" % (code, msg))
        self.code = code


def build(force=False, verbose=False):
    """Compile libdlsm.so in-tree for sm_100a (nvcc cross-compiles without a GPU): one object file
    per translation unit, compiled concurrently, rebuilt only when the unit or a header changed."""
    newest_hdr = max(os.path.getmtime(s) for s in HEADERS)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    procs, objs = [], []
    for u in UNITS:
        o = os.path.join(OBJ_DIR, os.path.basename(u)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(u), newest_hdr):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", o, u]
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    if procs or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs)
    return LIB_PATH


class Config(C.Structure):
    _fields_ = [("n_chains", C.c_int32), ("T", C.c_int32), ("n", C.c_int32), ("d", C.c_int32),
                ("K", C.c_int32), ("is_directed", C.c_int32), ("likelihood", C.c_int32),
                ("prior", C.c_int32), ("device", C.c_int32), ("tune", C.c_int32),
                ("tune_interval", C.c_int32), ("intercept_tune_interval", C.c_int32 * 2),
                ("radii_tune", C.c_int32), ("radii_tune_interval", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class Hyper(C.Structure):
    _fields_ = [("tau_sq", C.c_double), ("sigma_sq", C.c_double),
                ("intercept_prior", C.c_double * 2), ("intercept_variance_prior", C.c_double)]


class HdpPrior(C.Structure):
    _fields_ = [("a", C.c_double), ("a0", C.c_double), ("b0", C.c_double), ("c0", C.c_double),
                ("d0", C.c_double), ("lambda_prior", C.c_double), ("lambda_variance_prior", C.c_double),
                ("gamma_prior_shape", C.c_double), ("gamma_prior_rate", C.c_double),
                ("alpha_init_shape", C.c_double), ("alpha_init_rate", C.c_double),
                ("alpha_kappa_shape", C.c_double), ("alpha_kappa_rate", C.c_double),
                ("resample_mvp", C.c_int32), ("resample_b", C.c_int32)]


class TraceSpec(C.Structure):
    _fields_ = [("fields_all", C.c_uint32), ("fields_first", C.c_uint32), ("thin", C.c_int32),
                ("want_logp", C.c_int32), ("cooc_mode", C.c_int32), ("cooc_from", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("node_updates", C.c_uint64),
                ("sweeps", C.c_uint64), ("latent_ms", C.c_double), ("other_ms", C.c_double),
                ("ub_flags", C.c_uint64), ("cluster_sweeps", C.c_uint64), ("rowsum_sweeps", C.c_uint64)]


# every symbol include/dlsm.h declares
EXPORTS = [
    "dlsm_abi_version", "dlsm_device_count", "dlsm_create", "dlsm_destroy", "dlsm_last_error",
    "dlsm_set_stream", "dlsm_synchronize", "dlsm_set_network_dense", "dlsm_set_edge_lists",
    "dlsm_set_controls", "dlsm_set_state", "dlsm_get_state", "dlsm_set_hyper", "dlsm_set_rng",
    "dlsm_sweep_latent", "dlsm_center", "dlsm_sample_intercepts", "dlsm_sample_radii",
    "dlsm_sample_labels", "dlsm_set_hdp_prior", "dlsm_hdp_update", "dlsm_run_sweeps", "dlsm_loglik_partial", "dlsm_loglik_full",
    "dlsm_gaussian_likelihood", "dlsm_debug_set_counts", "dlsm_debug_draws", "dlsm_enable_timing", "dlsm_get_counters",
    "dlsm_resample_controls", "dlsm_get_controls", "dlsm_edge_probas", "dlsm_cooccurrence", "dlsm_logp", "dlsm_set_procrustes_ref", "dlsm_procrustes", "dlsm_run_traced", "dlsm_host_alloc", "dlsm_host_free",
    "dlsm_set_option", "dlsm_debug_rowsums", "dlsm_set_network_edges", "dlsm_edge_list_dims",
    "dlsm_get_edge_lists",
]

# dlsm_option / dlsm_sweep_mode / dlsm_ffbs_kernel (include/dlsm.h)
(OPT_SWEEP_MODE, OPT_FFBS_KERNEL, OPT_FFBS_SMEM_STAGE, OPT_FFBS_CTAS_PER_SM, OPT_NO_GATHER_PACK,
 OPT_NO_TRACKED_LOGLIK, OPT_CENTER_EXACT, OPT_HDP_SEGMENTED, OPT_NO_EARLY_X, OPT_TRACE_CHUNK_BYTES,
 OPT_NO_ROWSUM_CACHE, OPT_NO_CLUSTER, OPT_CHAIN_KERNEL, OPT_CC_KERNEL, OPT_FULL_KERNEL, OPT_CCD_GROUP, OPT_FFBS_NO_L2_WINDOW) = range(17)
CHAIN_AUTO, CHAIN_NODE_ROWSUM, CHAIN_NODE, CHAIN_BLOCK, CHAIN_BLOCK_PAIR = range(5)
SWEEP_AUTO, SWEEP_CHAIN, SWEEP_CHAIN_DENSE, SWEEP_SLICE, SWEEP_SLICE_PLAIN = range(5)
FFBS_AUTO, FFBS_THREAD, FFBS_WARP = range(3)

F_X, F_INTERCEPT, F_RADII, F_Z, F_MU, F_SIGMA, F_LAMBDA, F_WEIGHTS = range(8)
F_X_STEP, F_X_NACC, F_X_NSTEPS, F_X_UNTIL = 8, 9, 10, 11
F_B_STEP, F_B_NACC, F_B_NSTEPS, F_B_UNTIL = 12, 13, 14, 15
F_R_STEP, F_R_NACC, F_R_NSTEPS, F_R_UNTIL = 16, 17, 18, 19
F_NCOUNT, F_NK, F_BETA, F_HYPER, F_LOGLIK = 20, 21, 22, 23, 24
_INT_FIELDS = {F_Z, F_X_NACC, F_X_NSTEPS, F_X_UNTIL, F_B_NACC, F_B_NSTEPS, F_B_UNTIL, F_R_NACC,
               F_R_NSTEPS, F_R_UNTIL, F_NK}

_lib = None


def load():
    """Load libdlsm.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("dynetlsm_b200: %s is missing -- build it with `python -c \"import "
                          "__graft_entry__ as g; g.build()\"` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.dlsm_last_error.restype = C.c_char_p
    L.dlsm_last_error.argtypes = [vp]
    L.dlsm_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.dlsm_destroy.argtypes = [vp]
    L.dlsm_destroy.restype = None
    L.dlsm_set_stream.argtypes = [vp, vp]
    L.dlsm_synchronize.argtypes = [vp]
    L.dlsm_set_option.argtypes = [vp, C.c_int, C.c_int64]
    L.dlsm_set_network_dense.argtypes = [vp, dp]
    L.dlsm_set_edge_lists.argtypes = [vp, ip, ip, C.c_int32, ip, C.c_int32]
    L.dlsm_set_controls.argtypes = [vp, ip, ip, C.c_int32, C.c_int32]
    L.dlsm_set_network_edges.argtypes = [vp, ip, C.c_size_t]
    L.dlsm_edge_list_dims.argtypes = [vp, ip, ip]
    L.dlsm_get_edge_lists.argtypes = [vp, ip, ip, ip]
    L.dlsm_set_state.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.dlsm_get_state.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.dlsm_set_hyper.argtypes = [vp, C.POINTER(Hyper)]
    L.dlsm_set_rng.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.dlsm_sweep_latent.argtypes = [vp, dp, dp, ip, dp]
    L.dlsm_center.argtypes = [vp]
    L.dlsm_sample_intercepts.argtypes = [vp, dp, dp, ip, dp]
    L.dlsm_sample_radii.argtypes = [vp, dp, dp, ip, dp]
    L.dlsm_sample_labels.argtypes = [vp, dp]
    L.dlsm_set_hdp_prior.argtypes = [vp, C.POINTER(HdpPrior)]
    L.dlsm_hdp_update.argtypes = [vp]
    L.dlsm_run_sweeps.argtypes = [vp, C.c_int32, C.c_uint32]
    L.dlsm_loglik_partial.argtypes = [vp, dp]
    L.dlsm_loglik_full.argtypes = [vp, dp]
    L.dlsm_gaussian_likelihood.argtypes = [vp, dp]
    L.dlsm_debug_set_counts.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.dlsm_debug_draws.argtypes = [vp, dp, dp]
    L.dlsm_debug_rowsums.argtypes = [vp, dp]
    L.dlsm_enable_timing.argtypes = [vp, C.c_int]
    L.dlsm_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.dlsm_logp.argtypes = [vp, dp]
    L.dlsm_edge_probas.argtypes = [vp, C.c_int32, dp]
    L.dlsm_cooccurrence.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_int32]
    L.dlsm_resample_controls.argtypes = [vp, C.c_int32, C.c_int32]
    L.dlsm_get_controls.argtypes = [vp, ip, ip]
    L.dlsm_set_procrustes_ref.argtypes = [vp, dp]
    L.dlsm_procrustes.argtypes = [vp]
    L.dlsm_run_traced.argtypes = [vp, C.c_int32, C.c_uint32, C.POINTER(TraceSpec), C.POINTER(vp), dp]
    L.dlsm_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.dlsm_host_free.argtypes = [vp]
    _lib = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


def _i32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


N_FIELDS = 25


class _Pinned(object):
    """Page-locked host block (dlsm_host_alloc) exposed through the array interface."""

    def __init__(self, nbytes):
        L = load()
        self._L, self.ptr = L, C.c_void_p()
        if L.dlsm_host_alloc(max(int(nbytes), 1), C.byref(self.ptr)) != 0:
            raise MemoryError("dlsm_host_alloc(%d) failed" % nbytes)
        self.__array_interface__ = {"shape": (max(int(nbytes), 1),), "typestr": "|u1",
                                    "data": (self.ptr.value, False), "version": 3}

    def __del__(self):
        if getattr(self, "ptr", None) and self.ptr.value:
            self._L.dlsm_host_free(self.ptr)
            self.ptr = C.c_void_p()


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked memory (freed when the last view goes away)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    raw = np.asarray(_Pinned(nbytes))
    return raw[:nbytes].view(dtype).reshape(shape)


class _TraceDict(dict):
    """Result of Engine.run_traced; keeps the (possibly larger) destination buffers for reuse."""
    _base = None


class Engine(object):
    """Device-resident sampler state for ``n_chains`` chains sharing one dynamic network."""

    def __init__(self, T, n, d, n_chains=1, K=0, is_directed=False, case_control=False,
                 mixture=False, device=0, tune=500, tune_interval=100,
                 intercept_tune_interval=(100, 100), radii_tune=None, radii_tune_interval=100):
        self.L = load()
        cfg = Config()
        cfg.n_chains, cfg.T, cfg.n, cfg.d, cfg.K = n_chains, T, n, d, K
        cfg.is_directed = int(is_directed)
        cfg.likelihood = 1 if case_control else 0
        cfg.prior = 1 if mixture else 0
        cfg.device = device
        cfg.tune = -1 if tune is None else int(tune)
        cfg.tune_interval = int(tune_interval)
        cfg.intercept_tune_interval[0] = int(intercept_tune_interval[0])
        cfg.intercept_tune_interval[1] = int(intercept_tune_interval[1])
        cfg.radii_tune = -1 if radii_tune is None else int(radii_tune)
        cfg.radii_tune_interval = int(radii_tune_interval)
        self.cfg = cfg
        self.C, self.T, self.n, self.d, self.K = n_chains, T, n, d, K
        self.is_directed, self.case_control, self.mixture = bool(is_directed), bool(case_control), bool(mixture)
        self.m = 2 if is_directed else 1
        self.h = C.c_void_p()
        rc = self.L.dlsm_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise DlsmError(rc, self.L.dlsm_last_error(None).decode())

    # -- plumbing ------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise DlsmError(rc, self.L.dlsm_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.dlsm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shape_of(self, f):
        C_, T, n, d, K = self.C, self.T, self.n, self.d, self.K
        return {F_X: (C_, T, n, d), F_INTERCEPT: (C_, 2), F_RADII: (C_, n), F_Z: (C_, T, n),
                F_MU: (C_, K, d), F_SIGMA: (C_, K), F_LAMBDA: (C_,), F_WEIGHTS: (C_, T, K, K),
                F_X_STEP: (C_, T, n), F_X_NACC: (C_, T, n), F_X_NSTEPS: (C_, T, n),
                F_X_UNTIL: (C_, T, n), F_B_STEP: (C_, 2), F_B_NACC: (C_, 2), F_B_NSTEPS: (C_, 2),
                F_B_UNTIL: (C_, 2), F_R_STEP: (C_,), F_R_NACC: (C_,), F_R_NSTEPS: (C_,),
                F_R_UNTIL: (C_,), F_NCOUNT: (C_, T, K, K), F_NK: (C_, T, K), F_BETA: (C_, K),
                F_HYPER: (C_, 8), F_LOGLIK: (C_,)}[f]

    def set(self, f, a):
        a = (_i32 if f in _INT_FIELDS else _f64)(a, self.shape_of(f))
        self._ck(self.L.dlsm_set_state(self.h, f, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def get(self, f, out=None):
        """Copy a state field to the host; ``out`` (e.g. a pinned buffer) is filled in place."""
        dt = np.int32 if f in _INT_FIELDS else np.float64
        if out is None:
            a = np.empty(self.shape_of(f), dtype=dt)
        else:
            a = out
            if a.dtype != dt or a.shape != self.shape_of(f) or not a.flags.c_contiguous:
                raise ValueError("out must be C-contiguous %s of shape %s" % (dt.__name__, self.shape_of(f)))
        self._ck(self.L.dlsm_get_state(self.h, f, a.ctypes.data_as(C.c_void_p), a.nbytes))
        return a

    def set_option(self, option, value):
        """Typed developer option (dlsm_set_option): which kernel variant serves a step."""
        self._ck(self.L.dlsm_set_option(self.h, int(option), int(value)))

    def set_stream(self, cuda_stream):
        self._ck(self.L.dlsm_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.L.dlsm_synchronize(self.h))

    # -- inputs --------------------------------------------------------------------------
    def set_network(self, Y):
        Y = _f64(Y, (self.T, self.n, self.n))
        self._ck(self.L.dlsm_set_network_dense(self.h, _dp(Y)))

    def set_edge_lists(self, degrees, in_edges, out_edges):
        dg = _i32(degrees, (self.T, self.n, 2))
        ie, oe = _i32(in_edges), _i32(out_edges)
        self._ck(self.L.dlsm_set_edge_lists(self.h, _ip(dg), _ip(ie), ie.shape[2], _ip(oe), oe.shape[2]))

    def set_network_edges(self, edges):
        """Sparse network: (E, 3) int rows (t, sender, receiver); the case-control edge lists are built
        on the device (dlsm_set_network_edges)."""
        ed = _i32(edges)
        if ed.ndim != 2 or ed.shape[1] != 3:
            raise ValueError("edges must have shape (n_edges, 3): (t, sender, receiver)")
        self._ck(self.L.dlsm_set_network_edges(self.h, _ip(ed), ed.shape[0]))

    def get_edge_lists(self):
        """(degrees (T,n,2), in_edges (T,n,max_in), out_edges (T,n,max_out)) as held by the device."""
        mi, mo = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.dlsm_edge_list_dims(self.h, C.byref(mi), C.byref(mo)))
        dg = np.empty((self.T, self.n, 2), np.int32)
        ie = np.zeros((self.T, self.n, mi.value), np.int32)
        oe = np.zeros((self.T, self.n, mo.value), np.int32)
        self._ck(self.L.dlsm_get_edge_lists(self.h, _ip(dg), _ip(ie) if mi.value else None,
                                            _ip(oe) if mo.value else None))
        return dg, ie, oe

    def set_controls(self, ctrl_in, ctrl_out):
        ci, co = _i32(ctrl_in), _i32(ctrl_out)
        if ci.ndim == 3:
            ci, co = ci[None], co[None]
        self._ck(self.L.dlsm_set_controls(self.h, _ip(ci), _ip(co), ci.shape[3], ci.shape[0]))
        self._n_control, self._ctrl_sets = int(ci.shape[3]), int(ci.shape[0])

    def set_hyper(self, tau_sq=2.0, sigma_sq=0.1, intercept_prior=(0.0, 0.0),
                  intercept_variance_prior=2.0):
        hy = Hyper()
        hy.tau_sq, hy.sigma_sq = float(tau_sq), float(sigma_sq)
        ip = np.atleast_1d(np.asarray(intercept_prior, dtype=np.float64))
        hy.intercept_prior[0] = float(ip[0])
        hy.intercept_prior[1] = float(ip[1]) if ip.size > 1 else 0.0
        hy.intercept_variance_prior = float(intercept_variance_prior)
        self._ck(self.L.dlsm_set_hyper(self.h, C.byref(hy)))

    def set_rng(self, seed, chain_offset=0, sweep_index=0):
        self._ck(self.L.dlsm_set_rng(self.h, int(seed), int(chain_offset), int(sweep_index)))

    def set_tuner(self, step_X, step_intercept=0.1, step_radii=175000.0):
        """Fresh Metropolis objects (metropolis.py:86-94): zero counters, until = tune_interval."""
        C_, T, n = self.C, self.T, self.n
        self.set(F_X_STEP, np.full((C_, T, n), float(step_X)))
        self.set(F_X_NACC, np.zeros((C_, T, n), np.int32))
        self.set(F_X_NSTEPS, np.zeros((C_, T, n), np.int32))
        self.set(F_X_UNTIL, np.full((C_, T, n), self.cfg.tune_interval, np.int32))
        self.set(F_B_STEP, np.full((C_, 2), float(step_intercept)))
        self.set(F_B_NACC, np.zeros((C_, 2), np.int32))
        self.set(F_B_NSTEPS, np.zeros((C_, 2), np.int32))
        self.set(F_B_UNTIL, np.tile(np.array(list(self.cfg.intercept_tune_interval), np.int32), (C_, 1)))
        self.set(F_R_STEP, np.full((C_,), float(step_radii)))
        self.set(F_R_NACC, np.zeros((C_,), np.int32))
        self.set(F_R_NSTEPS, np.zeros((C_,), np.int32))
        self.set(F_R_UNTIL, np.full((C_,), self.cfg.radii_tune_interval, np.int32))

    # -- hot path ------------------------------------------------------------------------
    def sweep_latent(self, eps=None, logu=None, want_stats=False):
        C_, T, n, d = self.C, self.T, self.n, self.d
        if eps is not None:
            eps, logu = _f64(eps, (C_, T, n, d)), _f64(logu, (C_, T, n))
        acc = np.empty((C_, T, n), np.int32) if want_stats else None
        rat = np.empty((C_, T, n)) if want_stats else None
        self._ck(self.L.dlsm_sweep_latent(self.h, _dp(eps), _dp(logu), _ip(acc), _dp(rat)))
        return (acc, rat) if want_stats else None

    def center(self):
        self._ck(self.L.dlsm_center(self.h))

    def sample_intercepts(self, eps=None, logu=None, want_stats=False):
        if eps is not None:
            eps, logu = _f64(eps, (self.C, self.m)), _f64(logu, (self.C, self.m))
        acc = np.empty((self.C, self.m), np.int32) if want_stats else None
        rat = np.empty((self.C, self.m)) if want_stats else None
        self._ck(self.L.dlsm_sample_intercepts(self.h, _dp(eps), _dp(logu), _ip(acc), _dp(rat)))
        return (acc, rat) if want_stats else None

    def sample_radii(self, proposal=None, logu=None, want_stats=False):
        if proposal is not None:
            proposal, logu = _f64(proposal, (self.C, self.n)), _f64(logu, (self.C,))
        acc = np.empty((self.C,), np.int32) if want_stats else None
        rat = np.empty((self.C,)) if want_stats else None
        self._ck(self.L.dlsm_sample_radii(self.h, _dp(proposal), _dp(logu), _ip(acc), _dp(rat)))
        return (acc, rat) if want_stats else None

    def sample_labels(self, U=None):
        if U is not None:
            U = _f64(U, (self.C, self.n, self.T))
        self._ck(self.L.dlsm_sample_labels(self.h, _dp(U)))

    def set_hdp_prior(self, a, a0, b0, c0, d0, lambda_prior, lambda_variance_prior,
                      gamma_prior_shape, gamma_prior_rate, alpha_init_shape, alpha_init_rate,
                      alpha_kappa_shape, alpha_kappa_rate, resample_mvp=True, resample_b=True):
        pr = HdpPrior()
        pr.a, pr.a0, pr.b0 = float(a), float(a0 or 0.0), float(b0 or 0.0)
        pr.c0, pr.d0 = float(c0 or 0.0), float(d0 or 0.0)
        pr.lambda_prior, pr.lambda_variance_prior = float(lambda_prior), float(lambda_variance_prior)
        pr.gamma_prior_shape, pr.gamma_prior_rate = float(gamma_prior_shape), float(gamma_prior_rate)
        pr.alpha_init_shape, pr.alpha_init_rate = float(alpha_init_shape), float(alpha_init_rate)
        pr.alpha_kappa_shape, pr.alpha_kappa_rate = float(alpha_kappa_shape), float(alpha_kappa_rate)
        pr.resample_mvp, pr.resample_b = int(bool(resample_mvp)), int(bool(resample_b))
        self._ck(self.L.dlsm_set_hdp_prior(self.h, C.byref(pr)))

    def hdp_update(self):
        self._ck(self.L.dlsm_hdp_update(self.h))

    def run_sweeps(self, n_sweeps, skip_center=False, skip_intercepts=False, skip_radii=False,
                   skip_labels=False, skip_hdp=False):
        flags = (1 if skip_center else 0) | (2 if skip_intercepts else 0) | \
                (4 if skip_radii else 0) | (8 if skip_labels else 0) | (16 if skip_hdp else 0)
        self._ck(self.L.dlsm_run_sweeps(self.h, int(n_sweeps), flags))

    def run_traced(self, n_sweeps, fields_all=(), fields_first=(), thin=1, logp=True, pinned=False,
                   skip_center=False, skip_intercepts=False, skip_radii=False, skip_labels=False,
                   skip_hdp=False, out=None, cooc=0, cooc_from=0):
        """``n_sweeps`` sweeps on the device, recording every ``thin``-th state: returns
        ``{field: array (records, C or 1, ...)}`` plus ``"logp": (records, C)``.  The copies to the
        host overlap the following sweeps.  ``out``: a dict returned by an earlier call with at
        least as many records, reused as the destination (page-locking memory is slow: allocate
        once, then pass it back in); the returned arrays are views of its first ``records`` rows."""
        flags = (1 if skip_center else 0) | (2 if skip_intercepts else 0) | \
                (4 if skip_radii else 0) | (8 if skip_labels else 0) | (16 if skip_hdp else 0)
        rec = int(n_sweeps) // int(thin)
        alloc = pinned_empty if pinned else np.empty
        spec = TraceSpec(thin=int(thin), want_logp=int(bool(logp)), cooc_mode=int(cooc),
                         cooc_from=int(cooc_from))
        dst = (C.c_void_p * N_FIELDS)()
        base = out._base if isinstance(out, _TraceDict) else (out or {})
        res = _TraceDict()
        res._base = base

        def buffer(key, shape, dtype):
            a = base.get(key)
            if a is None or a.shape[0] < rec or a.shape[1:] != shape[1:] or a.dtype != dtype:
                a = base[key] = alloc(shape, dtype)
            return a
        for group, first in ((fields_all, False), (fields_first, True)):
            for f in group:
                shp = self.shape_of(f)
                a = buffer(f, (rec, 1 if first else shp[0]) + tuple(shp[1:]),
                           np.dtype(np.int32 if f in _INT_FIELDS else np.float64))
                res[f] = a[:rec]
                dst[f] = a.ctypes.data
                if first:
                    spec.fields_first |= 1 << f
                else:
                    spec.fields_all |= 1 << f
        lp = buffer("logp", (rec, self.C), np.dtype(np.float64)) if logp else None
        self._ck(self.L.dlsm_run_traced(self.h, int(n_sweeps), flags, C.byref(spec), dst, _dp(lp)))
        if logp:
            res["logp"] = lp[:rec]
        return res

    def resample_controls(self, n_control, per_chain=False):
        """Redraw the case-control sets on the device (uniform without replacement)."""
        self._n_control, self._ctrl_sets = int(n_control), (self.C if per_chain else 1)
        self._ck(self.L.dlsm_resample_controls(self.h, self._n_control, self._ctrl_sets))

    def get_controls(self):
        shp = (self._ctrl_sets, self.T, self.n, self._n_control)
        ci, co = np.empty(shp, np.int32), np.empty(shp, np.int32)
        self._ck(self.L.dlsm_get_controls(self.h, _ip(ci), _ip(co)))
        return ci, co

    def cooccurrence(self, reset=False):
        """(counts (T, n, n) uint32, n_samples) accumulated by run_traced(cooc=1|2)."""
        out = np.zeros((self.T, self.n, self.n), np.uint32)
        ns = C.c_uint64(0)
        self._ck(self.L.dlsm_cooccurrence(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(ns),
                                          int(bool(reset))))
        return out, int(ns.value)

    def edge_probas(self, chain=0):
        """(T, n, n) edge probabilities of one chain at its current state (zero diagonal)."""
        out = np.empty((self.T, self.n, self.n))
        self._ck(self.L.dlsm_edge_probas(self.h, int(chain), _dp(out)))
        return out

    def logp(self):
        out = np.empty((self.C,))
        self._ck(self.L.dlsm_logp(self.h, _dp(out)))
        return out

    def set_procrustes_ref(self, Xref):
        if Xref is None:
            self._ck(self.L.dlsm_set_procrustes_ref(self.h, None))
        else:
            self._ck(self.L.dlsm_set_procrustes_ref(self.h, _dp(_f64(Xref, self.shape_of(F_X)))))

    def procrustes(self):
        self._ck(self.L.dlsm_procrustes(self.h))

    # -- probes --------------------------------------------------------------------------
    def loglik_partial(self):
        out = np.empty((self.C, self.T, self.n))
        self._ck(self.L.dlsm_loglik_partial(self.h, _dp(out)))
        return out

    def loglik_full(self):
        out = np.empty((self.C,))
        self._ck(self.L.dlsm_loglik_full(self.h, _dp(out)))
        return out

    def gaussian_likelihood(self):
        out = np.empty((self.C, self.n, self.T, self.K))
        self._ck(self.L.dlsm_gaussian_likelihood(self.h, _dp(out)))
        return out

    def rowsums(self):
        """The device loop's cached per-node log-likelihoods (C, T, n) (dlsm_debug_rowsums)."""
        out = np.empty((self.C, self.T, self.n))
        self._ck(self.L.dlsm_debug_rowsums(self.h, _dp(out)))
        return out

    def debug_draws(self):
        eps = np.empty((self.C, self.T, self.n, self.d))
        logu = np.empty((self.C, self.T, self.n))
        self._ck(self.L.dlsm_debug_draws(self.h, _dp(eps), _dp(logu)))
        return eps, logu

    def enable_timing(self, on=True):
        self._ck(self.L.dlsm_enable_timing(self.h, int(on)))

    def counters(self):
        c = Counters()
        self._ck(self.L.dlsm_get_counters(self.h, C.byref(c)))
        return dict(kernel_launches=c.kernel_launches, node_updates=c.node_updates, sweeps=c.sweeps,
                    latent_ms=c.latent_ms, other_ms=c.other_ms, ub_flags=c.ub_flags,
                    cluster_sweeps=c.cluster_sweeps, rowsum_sweeps=c.rowsum_sweeps)


def device_count():
    return load().dlsm_device_count()
