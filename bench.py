#!/usr/bin/env python
"""bench.py -- throughput of the MH-within-Gibbs hot path on B200 (see DESIGN.md, section Measurement).

A *step* is one pass of the hot path over every chain resident on the GPU: latent-position sweep
(T*n MH node-updates per chain) -> centring -> intercept MH (full-network likelihood)
[-> radii MH] [-> label FFBS].  The metric is latent-position node-updates/s (whole job);
sweeps/s (chain-sweeps per second) is reported beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg4|cfg3|cfg1]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # the reference's CPU implementation on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: T, n, d, K, directed, case_control, default chains per GPU, description
    "cfg1": dict(T=3, n=18, d=2, K=0, directed=False, chains=1184,
                 desc="DynamicNetworkLSM, Sampson-monks shape (T=3, n=18, d=2)"),
    # 1332 = 148 SMs x 3 resident CTAs x 3 waves of the sweep kernel
    "cfg2": dict(T=9, n=120, d=2, K=10, directed=False, chains=1332,
                 desc="DynamicNetworkHDPLPCM, community-splitting network (n=120, T=9, d=2, K=10)"),
    "cfg3": dict(T=20, n=2000, d=2, K=0, directed=True, chains=1,
                 desc="directed DynamicNetworkLSM with radii (n=2000, T=20, d=2), single chain"),
    "cfg4": dict(T=10, n=500, d=2, K=10, directed=False, chains=296,
                 desc="DynamicNetworkHDPLPCM multi-chain (n=500, T=10, d=2, K=10)"),
    "cfg5": dict(T=10, n=50000, d=2, K=0, directed=True, chains=8, case_control=True, n_control=100,
                 desc="directed LSM, case-control likelihood, sparse network (n=50000, T=10, d=2, "
                      "out-degree ~ Poisson(10), 100 controls)"),
}


def expit(x):
    return 1.0 / (1.0 + np.exp(-x))


def make_sparse_workload(name, seed=42):
    """cfg 5: a sparse directed network that cannot exist as a dense tensor (200 GB): latent random
    walk, out-neighbours drawn among the nearest nodes in the latent space, stored as the padded
    edge lists + control sets of the case-control likelihood."""
    from scipy.spatial import cKDTree
    w = dict(WORKLOADS[name])
    T, n, d, nc = w["T"], w["n"], w["d"], w["n_control"]
    rng = np.random.RandomState(seed)
    X = np.empty((T, n, d))
    X[0] = 0.01 * rng.randn(n, d)
    for t in range(1, T):
        X[t] = X[t - 1] + 0.001 * rng.randn(n, d)
    out_lists = []
    for t in range(T):
        deg = np.minimum(rng.poisson(10, n), 30)
        _, nb = cKDTree(X[t]).query(X[t], k=41)
        pick = np.argsort(rng.rand(n, 40), axis=1)          # random subset of the 40 nearest
        out_lists.append((deg, np.take_along_axis(nb[:, 1:], pick, axis=1)))
    max_out = 30
    out_e = np.zeros((T, n, max_out), np.int32)
    degs = np.zeros((T, n, 2), np.int32)
    in_lists = []
    for t, (deg, cand) in enumerate(out_lists):
        keep = np.arange(max_out)[None, :] < deg[:, None]
        out_e[t] = np.where(keep, cand[:, :max_out], 0)
        degs[t, :, 1] = deg
        src = np.repeat(np.arange(n), deg)
        dst = cand[:, :max_out][keep]
        order = np.lexsort((src, dst))
        in_lists.append((dst[order], src[order]))
        degs[t, :, 0] = np.bincount(dst, minlength=n)
    max_in = int(degs[:, :, 0].max())
    in_e = np.zeros((T, n, max_in), np.int32)
    for t, (dst, src) in enumerate(in_lists):
        start = np.searchsorted(dst, dst, side="left")
        in_e[t, dst, np.arange(dst.size) - start] = src
    # control sets: uniform non-neighbours (collisions with neighbours/self are rare at this
    # sparsity and are redrawn once; a residual collision only perturbs the estimator's weights)
    def controls(edges, deg_col):
        c = rng.randint(0, n, size=(T, n, nc)).astype(np.int32)
        for _ in range(2):
            bad = c == np.arange(n, dtype=np.int32)[None, :, None]
            for q in range(edges.shape[2]):
                bad |= (c == edges[:, :, q:q + 1]) & (q < degs[:, :, deg_col])[:, :, None]
            c[bad] = rng.randint(0, n, size=int(bad.sum()))
        return c
    radii = rng.dirichlet(np.ones(n) * 20.0)
    w.update(name=name, X=X, Y=None, radii=radii, intercept=np.array([0.3, 0.7]),
             step_X=0.0075 / n * 40, sigma_sq=1e-6, tau_sq=float(np.mean(X[0] * X[0])),
             degrees=degs, in_edges=in_e, out_edges=out_e, ctrl_in=controls(in_e, 0),
             ctrl_out=controls(out_e, 1), density=float(degs[:, :, 1].mean() / n),
             mean_deg=float(degs[:, :, 0].mean() + degs[:, :, 1].mean()))
    return w


def make_workload(name, seed=42):
    """Synthetic community-structured dynamic network of the named shape plus a start state."""
    if WORKLOADS[name].get("case_control"):
        return make_sparse_workload(name, seed)
    w = dict(WORKLOADS[name])
    T, n, d, K = w["T"], w["n"], w["d"], max(w["K"], 4)
    rng = np.random.RandomState(seed)
    directed = w["directed"]
    scale = (1.0 / n) if directed else 1.0            # latent_space.py:92-93
    centers = rng.randn(K, d) * 2.0 * scale
    z0 = rng.randint(0, K, n)
    X = np.empty((T, n, d))
    X[0] = centers[z0] + 0.5 * scale * rng.randn(n, d)
    z = np.empty((T, n), np.int64)
    z[0] = z0
    for t in range(1, T):
        move = rng.rand(n) < 0.05
        z[t] = np.where(move, rng.randint(0, K, n), z[t - 1])
        X[t] = 0.8 * centers[z[t]] + 0.2 * X[t - 1] + 0.3 * scale * rng.randn(n, d)
    X -= X.mean(axis=(0, 1))
    Y = np.zeros((T, n, n))
    if directed:
        radii = rng.dirichlet(np.ones(n) * 20.0)
        b_in, b_out = 0.3, 0.7
        for t in range(T):
            dist = np.sqrt(((X[t][:, None, :] - X[t][None, :, :]) ** 2).sum(-1))
            eta = b_in * (1 - dist / radii[None, :]) + b_out * (1 - dist / radii[:, None])
            Y[t] = (rng.rand(n, n) < expit(eta)).astype(np.float64)
            np.fill_diagonal(Y[t], 0)
        w.update(radii=radii, intercept=np.array([b_in, b_out]), step_X=0.0075 / 8,
                 sigma_sq=0.001 * 1e-2, tau_sq=float(np.mean(X[0] * X[0])))
    else:
        beta = 1.0
        for t in range(T):
            dist = np.sqrt(((X[t][:, None, :] - X[t][None, :, :]) ** 2).sum(-1))
            U = np.triu((rng.rand(n, n) < expit(beta - dist)).astype(np.float64), 1)
            Y[t] = U + U.T
        w.update(radii=None, intercept=np.array([beta]), step_X=0.1, sigma_sq=0.1, tau_sq=2.0)
    Kc = w["K"]
    if Kc:
        mu = np.zeros((Kc, d)); mu[:K] = centers[:Kc] if Kc <= K else 0
        mu[:min(K, Kc)] = centers[:min(K, Kc)]
        sigma = np.full(Kc, 0.5)
        zz = np.minimum(z, Kc - 1)
        wts = np.full((T, Kc, Kc), 0.2 / (Kc - 1)) + np.eye(Kc) * (0.8 - 0.2 / (Kc - 1))
        w.update(mu=mu, sigma=sigma, lmbda=0.8, z=zz, w=wts)
    w.update(name=name, X=X, Y=Y, density=float(Y.mean()))
    return w


def bytes_per_node_update(w):
    """Algorithmic bytes of one node-update, model M1 (SURVEY.md 8d / BASELINE.md section 4)."""
    n, d = w["n"], w["d"]
    if w.get("case_control"):
        return (w["mean_deg"] + 2 * w["n_control"]) * (4 + 8.0 * d + 8)
    if w["directed"]:
        return 2 * n / 8.0 + 8.0 * d * n + 8.0 * n
    return n / 8.0 + 8.0 * d * n


def build_engine(w, chains, device, chain_offset, seed=42):
    from dynetlsm_b200 import _lib as L
    e = L.Engine(T=w["T"], n=w["n"], d=w["d"], n_chains=chains, K=w["K"], is_directed=w["directed"],
                 mixture=bool(w["K"]), device=device, tune=2500, tune_interval=100,
                 radii_tune=None, case_control=bool(w.get("case_control")))
    if w.get("case_control"):
        e.set_edge_lists(w["degrees"], w["in_edges"], w["out_edges"])
        e.set_controls(w["ctrl_in"], w["ctrl_out"])
    else:
        e.set_network(w["Y"])
    rng = np.random.RandomState(1000 + chain_offset)
    disp = 0.1 * (1.0 / w["n"] if w["directed"] else 1.0)
    X = w["X"][None] + disp * rng.randn(chains, w["T"], w["n"], w["d"])   # dispersed starts
    e.set(L.F_X, X)
    ic = np.zeros((chains, 2)); ic[:, :w["intercept"].size] = w["intercept"]
    e.set(L.F_INTERCEPT, ic)
    e.set_hyper(tau_sq=w["tau_sq"], sigma_sq=w["sigma_sq"], intercept_prior=w["intercept"],
                intercept_variance_prior=2.0)
    if w["directed"]:
        e.set(L.F_RADII, np.tile(w["radii"][None], (chains, 1)))
    if w["K"]:
        e.set(L.F_MU, np.tile(w["mu"][None], (chains, 1, 1)))
        e.set(L.F_SIGMA, np.tile(w["sigma"][None], (chains, 1)))
        e.set(L.F_LAMBDA, np.full(chains, w["lmbda"]))
        e.set(L.F_WEIGHTS, np.tile(w["w"][None], (chains, 1, 1, 1)))
        e.set(L.F_Z, np.tile(w["z"][None], (chains, 1, 1)))
        # sticky HDP-HMM hyper state and priors (hdp_lpcm.py defaults, n-dependent 'auto' values)
        K, n, d = w["K"], w["n"], w["d"]
        mvp = (n ** (2.0 / d)) / 50.0
        a, a0 = 2.0, (4.0 ** 2 + 2) * 2
        b0, b_ = (a0 - 2) * mvp * 2, (a + 2) * mvp
        d0 = (4.0 ** 2 / b_) * 2
        e.set(L.F_BETA, np.full((chains, K), 1.0 / K))
        e.set(L.F_HYPER, np.tile(np.array([[1.0, 1.0, 1.0, 4.0, mvp, b_, 0, 0]]), (chains, 1)))
        e.set_hdp_prior(a, a0, b0, b_ * d0, d0, 0.9, 0.01, 1.0, 0.1, 1.0, 1.0, 5, 0.1, True, True)
    e.set_tuner(w["step_X"])
    e.set_rng(seed, chain_offset=chain_offset)
    return e


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU legs: the reference's implementation on the host cores, one chain per core
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    name, sweeps, seed, kind = args
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    w = make_workload(name)
    rng = np.random.RandomState(seed)
    T, n = w["T"], w["n"]
    X0 = w["X"] + 0.1 * (1.0 / n if w["directed"] else 1.0) * rng.randn(*w["X"].shape)
    if kind == "reference":
        import ref_driver as R
        mix = (w["mu"], w["sigma"], np.array([w["lmbda"]]), w["z"].copy(), w["w"]) if w["K"] else None
        cc = None
        if w.get("case_control"):
            cc = {k: w[k].astype(np.int64) for k in ("in_edges", "out_edges", "degrees", "ctrl_in", "ctrl_out")}
        st = R.make_state(w["Y"], X0, w["intercept"], is_directed=w["directed"], radii=w["radii"],
                          mixture=mix, cc=cc, step_X=w["step_X"], tune=2500, tau_sq=w["tau_sq"],
                          sigma_sq=w["sigma_sq"])
        R.hot_path_sweep(st, rng)  # warm-up (imports, first-call overheads)
        t0 = time.perf_counter()
        for _ in range(sweeps):
            R.hot_path_sweep(st, rng)
        return time.perf_counter() - t0
    import pyoracle as O
    if w.get("case_control"):
        raise SystemExit("the C-oracle CPU leg does not cover cfg5; build oracle/_ref")
    X = np.ascontiguousarray(X0)
    tun = O.TunerState((T, n), w["step_X"], tune=2500, tune_interval=100)
    itun = O.TunerState((w["intercept"].size,), 0.1, tune=2500, tune_interval=100)
    ic = w["intercept"].copy()
    mix = dict(mu=w["mu"], sigma=w["sigma"], lmbda=w["lmbda"], z=w["z"]) if w["K"] else None

    def one():
        O.sweep_latent(X, ic, tun, rng.randn(T, n, w["d"]), np.log(rng.rand(T, n)), Y=w["Y"],
                       radii=w["radii"], is_directed=w["directed"], tau_sq=w["tau_sq"],
                       sigma_sq=w["sigma_sq"], mixture=mix)
        O.center(X)
        dist = O.calculate_distances(X)
        O.sample_intercepts(X, ic, itun, rng.randn(ic.size), np.log(rng.rand(ic.size)), ic, 2.0,
                            Y=w["Y"], dist=dist, radii=w["radii"], is_directed=w["directed"])
        if mix:
            mix["z"], _, _, _ = O.sample_labels_block(X, w["mu"], w["sigma"], w["lmbda"], w["w"],
                                                      rng.rand(n, T))
    one()
    t0 = time.perf_counter()
    for _ in range(sweeps):
        one()
    return time.perf_counter() - t0


def cpu_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_driver
    return "reference" if ref_driver.have_ref() else "port"


def cpu_hot_path(name, sweeps_per_core, cores=None):
    """Run `sweeps_per_core` hot-path sweeps on each of `cores` processes (one chain per core)."""
    import multiprocessing as mp
    kind = cpu_kind()
    if kind == "port":
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    cores = cores or os.cpu_count() or 1
    w = WORKLOADS[name]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        per = pool.map(_cpu_worker, [(name, sweeps_per_core, 7 + c, kind) for c in range(cores)])
    wall = time.perf_counter() - t0
    loop = max(per)
    updates = cores * sweeps_per_core * w["T"] * w["n"]
    return dict(value=updates / loop, unit="node-updates/s", cores=cores, kind=kind,
                sweeps_per_s=cores * sweeps_per_core / loop, loop_s=loop, wall_s=wall,
                sample="%d hot-path sweeps (latent + centre + distance cache + intercept MH [+ radii MH] "
                       "[+ label FFBS]; the reference's host-side conjugate block is NOT included, which "
                       "favours the CPU side) on each of %d cores, one chain per core, %s" % (
                    sweeps_per_core, cores, WORKLOADS[name]["desc"]))


def cpu_sweeps_for(name, target_s=12.0):
    # rough per-sweep cost of the reference loop (survey-time measurements), to bound the sample
    est = {"cfg1": 0.006, "cfg2": 0.12, "cfg3": 25.0, "cfg4": 0.9, "cfg5": 60.0}[name]
    if cpu_kind() == "port":
        est *= 0.2
    return max(1, int(target_s / est))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    w = WORKLOADS[name]
    per_step = max(1, cpu_sweeps_for(name, 60.0) // max(1, args.steps + args.warmup))
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_hot_path(name, 1, cores)
    t_loop, upd, sw = 0.0, 0, 0
    for _ in range(args.steps):
        r = cpu_hot_path(name, per_step, cores)
        t_loop += r["loop_s"]; upd += cores * per_step * w["T"] * w["n"]; sw += cores * per_step
    val = upd / t_loop
    line = {"impl": "reference", "metric": "latent-position node-updates/sec", "value": val,
            "unit": "node-updates/s", "sweeps_per_s": sw / t_loop, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_loop / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (name, w["desc"]),
                       "chains": cores, "note": "one chain per host core; a step = %d sweeps per core" % per_step},
            "cpu_baseline": {"value": val, "unit": "node-updates/s", "cores": cores, "kind": kind,
                             "sample": "%d steps x %d sweeps on each of %d cores" % (args.steps, per_step, cores)},
            "e2e": {"value": val, "unit": "node-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))  # cfg5: GPU arm only
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dynetlsm_b200 import _lib as L

    w = make_workload(args.workload)
    chains = args.chains_per_gpu or w["chains"]
    e = build_engine(w, chains, local, chain_offset=rank * chains)
    stream = torch.cuda.current_stream()
    e.set_stream(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        e.run_sweeps(1)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    c0 = e.counters()
    e.enable_timing(True)
    clocks = ClockSampler(local)
    clocks.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.zero_()                      # L2 flush between timed iterations (outside the events)
        a.record(stream)
        step()
        b.record(stream)
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    clk = clocks.stop()
    c1 = e.counters()
    e.enable_timing(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    upd_per_step = chains * w["T"] * w["n"]
    total_updates = world * upd_per_step * args.steps
    value = total_updates / (ms_max * 1e-3)
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    latent_ms = c1["latent_ms"] - c0["latent_ms"]
    other_ms = c1["other_ms"] - c0["other_ms"]

    # roofline of the dominant kernel (k_sweep): algorithmic bytes per launch / mean launch time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bpl = bytes_per_node_update(w) * upd_per_step
    lat_per_launch_ms = latent_ms / args.steps
    achieved = bpl / (lat_per_launch_ms * 1e-3) / 1e9
    traffic = None
    try:  # measured DRAM bytes of one k_sweep launch from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tr and tr["chains"] == chains:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_sweep", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": bpl,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "bytes_per_node_update": bytes_per_node_update(w),
                "kernel_ms_per_launch": lat_per_launch_ms,
                "kernel_share_of_step": latent_ms / ms if ms > 0 else None,
                "kernel_node_updates_per_s": upd_per_step / (lat_per_launch_ms * 1e-3)}

    # end-to-end through the public C-ABI with HOST buffers: what one iteration of the estimator's
    # fit loop exchanges with the device when the conjugate HDP updates run on the host
    e2e = None
    if not args.no_e2e:
        def pinned(a):
            tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return tt.numpy()
        # the function-level contract of the reference (sample_latent_positions(Y, X, ...) -> X, then
        # the other blocks): every step takes the positions from (pinned) host memory and returns
        # the new state -- positions, labels, intercepts and the mixture parameters fit() records
        ins = [(L.F_X, pinned(e.get(L.F_X)))]
        outs = [L.F_X, L.F_INTERCEPT] + ([L.F_RADII] if w["directed"] else []) + \
               ([L.F_Z, L.F_MU, L.F_SIGMA, L.F_LAMBDA, L.F_BETA, L.F_WEIGHTS, L.F_HYPER] if w["K"] else [])
        h2d = sum(a.nbytes for _, a in ins)
        nst = max(3, min(args.steps, 10))
        # one call per step returns the whole new state: dlsm_run_traced(1) copies the positions out
        # as soon as they are centred (while the intercept MH / label block still run) and the rest
        # of the state at the end of the sweep
        tr = None
        for it in range(2 + nst):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            for f, a in ins:
                e.set(f, a)
            tr = e.run_traced(1, fields_all=outs, logp=True, pinned=True, out=tr)
            ins[0] = (L.F_X, tr[L.F_X][0])
        barrier()
        dt = time.perf_counter() - t0
        d2h = sum(a.nbytes for a in tr.values())
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * upd_per_step * nst / float(tt.item()), "unit": "node-updates/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": nst,
               "api": "Engine.set(positions) -> Engine.run_traced(1 sweep, whole state + log-posterior to pinned host "
                      "buffers) over the dlsm C-ABI"}
        # the same state leaving the device every sweep through the streaming call fit() uses
        # (dlsm_run_traced: device trace ring drained on a copy stream while the next sweeps run);
        # no per-step host input exists in this mode, so it is reported beside e2e, not as e2e
        ntr = 3 * nst
        tr = None
        tr = e.run_traced(ntr, fields_all=outs, pinned=True)   # allocates the pinned destination
        barrier()
        t0 = time.perf_counter()
        tr = e.run_traced(ntr, fields_all=outs, pinned=True, out=tr)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e["traced"] = {"value": world * upd_per_step * ntr / float(tt.item()), "unit": "node-updates/s",
                         "steps": ntr, "d2h_bytes_per_step": int(sum(a.nbytes for a in tr.values()) // ntr),
                         "api": "Engine.run_traced(n, every state field of every chain, pinned destination)"}
        del tr

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_hot_path(args.workload, cpu_sweeps_for(args.workload))
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "sweeps_per_s")}

    if rank == 0:
        line = {"metric": "latent-position node-updates/sec", "value": value,
                "unit": "node-updates/s", "sweeps_per_s": world * chains * args.steps / (ms_max * 1e-3),
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s: %s" % (args.workload, w["desc"]),
                           "chains_per_gpu": chains, "chains_total": chains * world,
                           "density": w["density"], "rng": "device Philox4x32-10",
                           "step": "latent sweep + centre + intercept MH" +
                                   (" + radii MH" if w["directed"] else "") +
                                   (" + label FFBS + HDP conjugate updates" if w["K"] else ""),
                           "l2": "flushed between timed iterations (256 MiB write)"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clk,
                "gpu_launches": int(launches),
                "phase_ms_per_step": {"latent": latent_ms / args.steps, "other": other_ms / args.steps}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
