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

from workloads import (WORKLOADS, make_workload, make_sparse_workload, bytes_per_node_update,  # noqa: F401,E402
                       build_engine, chain_starts)


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
    make_workload(name)      # built once here, inherited by the forked workers
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
            "higher_is_better": True, "scaling": scaling_of(name), "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (name, w["desc"]),
                       "chains": cores, "note": "one chain per host core; a step = %d sweeps per core" % per_step},
            "cpu_baseline": {"value": val, "unit": "node-updates/s", "cores": cores, "kind": kind,
                             "loop": "restated",
                             "sample": "%d steps x %d sweeps on each of %d cores; the reference's compiled Cython "
                                       "kernels under a restated Python loop (oracle/ref_driver.py)" % (
                                           args.steps, per_step, cores)},
            "e2e": {"value": val, "unit": "node-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def scaling_of(name):
    # cfg 4 is stated as 1 024 chains in total, sharded over the GPUs (BASELINE.json); the other
    # multi-chain configurations keep the chains per GPU fixed
    return "strong" if WORKLOADS[name].get("total") else "weak"


def pin_to_numa_node(local_rank):
    """Bind this rank's host threads (and hence its pinned allocations, first-touch) to the CPUs
    of the NUMA node its GPU hangs off, so that eight ranks do not funnel their PCIe traffic
    through one socket's memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, words)
        cpus = [64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------
class Ctx(object):
    pass


def measure(ctx, name, steps, warmup, chains_override=0, e2e=True, cpu=True, parity=True):
    """Time `steps` hot-path steps of workload `name` on this rank's GPU; returns the result
    dictionary on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from dynetlsm_b200 import _lib as L
    world, rank, local = ctx.world, ctx.rank, ctx.local
    w = make_workload(name)
    spec = WORKLOADS[name]
    if chains_override:
        chains = chains_override
    elif spec.get("total"):
        if spec["chains"] % world:
            raise SystemExit("%d chains do not split over %d GPUs" % (spec["chains"], world))
        chains = spec["chains"] // world
    else:
        chains = spec["chains"]
    e = build_engine(w, chains, local, chain_offset=rank * chains)
    stream = torch.cuda.current_stream()
    e.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_tuned, tune_log = pretune(e, w)
    for _ in range(warmup):
        e.run_sweeps(1)
    torch.cuda.synchronize()
    c0 = e.counters()
    e.enable_timing(True)
    clocks = ClockSampler(local)
    clocks.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(steps)]
    barrier()
    for a, b in evs:
        ctx.flush.zero_()                  # L2 flush between timed iterations (outside the events)
        a.record(stream)
        e.run_sweeps(1)
        b.record(stream)
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    clk = clocks.stop()
    c1 = e.counters()
    e.enable_timing(False)
    ms_max = maxreduce(ms)

    upd_per_step = chains * w["T"] * w["n"]
    value = world * upd_per_step * steps / (ms_max * 1e-3)
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    latent_ms = c1["latent_ms"] - c0["latent_ms"]
    other_ms = c1["other_ms"] - c0["other_ms"]
    acc, _ = e.sweep_latent(want_stats=True)       # one more native sweep, for the acceptance rate
    acc_rate = float(acc.mean())
    del acc

    # roofline of the dominant kernel (the latent sweep): algorithmic bytes per launch / mean launch time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bpl = bytes_per_node_update(w) * upd_per_step
    lat_ms = latent_ms / steps
    achieved = bpl / (lat_ms * 1e-3) / 1e9
    traffic, kname = None, "k_sweep"
    try:  # measured DRAM bytes of one sweep launch from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
        if tr:
            kname = tr.get("kernel", kname)
            if tr["chains"] == chains:
                traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": bpl,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "bytes_per_node_update": bytes_per_node_update(w),
                "kernel_ms_per_launch": lat_ms,
                "kernel_share_of_step": latent_ms / ms if ms > 0 else None,
                "kernel_node_updates_per_s": upd_per_step / (lat_ms * 1e-3)}

    # end-to-end through the public C-ABI with HOST buffers
    e2e_res = None
    if e2e:
        # the function-level contract of the reference (sample_latent_positions(Y, X, ...) -> X, then
        # the other blocks): every step takes the positions from pinned host memory and returns
        # the new state -- positions, labels, intercepts, radii and the mixture parameters
        xin = L.pinned_empty(e.shape_of(L.F_X))
        e.get(L.F_X, out=xin)
        outs = [L.F_X, L.F_INTERCEPT] + ([L.F_RADII] if w["directed"] else []) + \
               ([L.F_Z, L.F_MU, L.F_SIGMA, L.F_LAMBDA, L.F_BETA, L.F_WEIGHTS, L.F_HYPER] if w["K"] else [])
        nst = max(3, min(steps, 10))
        tr = None
        for it in range(2 + nst):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            e.set(L.F_X, xin)
            tr = e.run_traced(1, fields_all=outs, logp=True, pinned=True, out=tr)
            xin = tr[L.F_X][0]
        barrier()
        dt = maxreduce(time.perf_counter() - t0)
        h2d = int(xin.nbytes)
        d2h = int(sum(a.nbytes for a in tr.values()))
        e2e_res = {"value": world * upd_per_step * nst / dt, "unit": "node-updates/s",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": nst,
                   "pcie_gbs_per_rank": (h2d + d2h) * nst / dt / 1e9,
                   "api": "Engine.set(positions) -> Engine.run_traced(1 sweep, whole state + log-posterior to "
                          "pinned host buffers) over the dlsm C-ABI"}
        # the same state leaving the device every sweep through the streaming call fit() uses
        ntr = int(max(4, min(3 * nst, (1 << 30) // max(1, d2h))))   # at most ~1 GB of pinned destination
        tr = e.run_traced(ntr, fields_all=outs, pinned=True)   # allocates the pinned destination
        barrier()
        t0 = time.perf_counter()
        tr = e.run_traced(ntr, fields_all=outs, pinned=True, out=tr)
        barrier()
        dt = maxreduce(time.perf_counter() - t0)
        e2e_res["traced"] = {"value": world * upd_per_step * ntr / dt, "unit": "node-updates/s",
                             "steps": ntr, "d2h_bytes_per_step": int(sum(a.nbytes for a in tr.values()) // ntr),
                             "api": "Engine.run_traced(n, every state field of every chain, pinned destination)"}
        del tr

    # NCCL only pools per-chain scalar traces for R-hat / ESS: do it once, timed
    pooled = None
    if world > 1 and spec["chains"] > 1:
        from dynetlsm_b200 import diagnostics
        tr = e.run_traced(20, fields_all=[L.F_INTERCEPT], logp=True)
        local_tr = np.ascontiguousarray(tr["logp"].T)            # (chains, draws)
        barrier()
        t0 = time.perf_counter()
        allc = diagnostics.pool_traces(local_tr[:, :, None])
        torch.cuda.synchronize()
        dtp = maxreduce(time.perf_counter() - t0)
        pooled = {"chains_pooled": int(allc.shape[0]), "draws": int(allc.shape[1]), "ms": dtp * 1e3,
                  "rhat_logp": float(diagnostics.split_rhat(allc[:, :, 0])), "backend": "nccl all_gather"}

    cpu_res = None
    if cpu and rank == 0 and world == 1:
        cpu_res = cpu_leg(name, e, w, parity)

    e.close()
    if rank != 0:
        return None
    return {"value": value, "unit": "node-updates/s",
            "sweeps_per_s": world * chains * steps / (ms_max * 1e-3),
            "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
            "scaling": scaling_of(name),
            "config": {"workload": "%s: %s" % (name, w["desc"]),
                       "chains_per_gpu": chains, "chains_total": chains * world,
                       "density": w["density"], "rng": "device Philox4x32-10",
                       "inputs": "workloads.make_workload (SURVEY 8d constructions, seed 42), chains start at "
                                 "the generating state + 0.1*scale*N(0, I)",
                       "step": "latent sweep + centre + intercept MH" +
                               (" + radii MH" if w["directed"] else "") +
                               (" + label FFBS + HDP conjugate updates" if w["K"] else ""),
                       "accept_rate": acc_rate, "step_size_X": step_tuned,
                       "step_tuning": "before the warm-up the reference's tuning table (metropolis.py:5-37) is "
                                      "applied to the chain-averaged acceptance rate, one sweep per adjustment, from "
                                      "step_size_X = %g until the rate is in its dead band [0.25, 0.4] (%d sweeps): "
                                      "the state every tuned sampler of a long run sits in" % (w["step_X"], len(tune_log)),
                       "l2": "flushed between timed iterations (256 MiB write)"},
            "roofline": roofline, "cpu_baseline": cpu_res, "e2e": e2e_res, "clocks": clk,
            "gpu_launches": int(launches), "pool_traces": pooled,
            "phase_ms_per_step": {"latent": latent_ms / steps, "other": other_ms / steps}}


def pretune(e, w, max_sweeps=60):
    """Bring the latent samplers to the acceptance band the reference's tuner converges to."""
    from dynetlsm_b200 import _lib as L
    step, log = float(w["step_X"]), []
    for _ in range(max_sweeps):
        acc, _ = e.sweep_latent(want_stats=True)
        rate = float(acc.mean())
        log.append((step, rate))
        if 0.25 <= rate <= 0.4:
            break
        # metropolis.py:5-37
        f = (0.1 if rate < 0.001 else 0.5 if rate < 0.05 else 0.9 if rate < 0.25 else
             10.0 if rate > 0.95 else 2.0 if rate > 0.75 else 1.1)
        step *= f
        e.set(L.F_X_STEP, np.full(e.shape_of(L.F_X_STEP), step))
    e.set_tuner(step)
    return step, log


def cpu_leg(name, e, w, parity):
    """The reference's CPU path on the host cores (bounded sample), and -- the one place bench.py
    may use the checker -- parity check (i) of SURVEY 8(d): per-node log-likelihoods of the state
    the GPU just reached against the reference's Cython kernel (oracle/_ref) or the C oracle."""
    from dynetlsm_b200 import _lib as L
    make_workload(name)
    r = cpu_hot_path(name, cpu_sweeps_for(name))
    res = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "sweeps_per_s")}
    res["loop"] = "restated"
    if parity:
        try:
            res["parity"] = parity_probe(e, w)
        except Exception as ex:  # the checker must never take the measurement down
            res["parity"] = {"error": repr(ex)}
    return res


def parity_probe(e, w, pairs=1000):
    from dynetlsm_b200 import _lib as L
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    rng = np.random.RandomState(0)
    got = e.loglik_partial()
    C_ = min(e.C, 4)
    X = e.get(L.F_X)[:C_]
    ic = e.get(L.F_INTERCEPT)[:C_]
    radii = e.get(L.F_RADII)[:C_] if w["directed"] else None
    worst = 0.0
    for _ in range(pairs):
        c, t, j = rng.randint(C_), rng.randint(w["T"]), rng.randint(w["n"])
        if w.get("case_control"):
            ci, co = e.get_controls()
            s = c if ci.shape[0] > 1 else 0
            ref = O.approx_directed_partial_loglikelihood(
                X[c, t], radii[c], w["in_edges"][t], w["out_edges"][t], w["degrees"][t], ci[s, t], co[s, t],
                ic[c, 0], ic[c, 1], j)
        elif w["directed"]:
            ref = O.directed_partial_loglikelihood(w["Y"][t], X[c, t], radii[c], ic[c, 0], ic[c, 1], j)
        else:
            ref = O.partial_loglikelihood(w["Y"][t], X[c, t], ic[c, 0], j)
        worst = max(worst, abs(got[c, t, j] - ref) / max(abs(ref), 1e-300))
    return {"check": "per-node log-likelihood of the reached state, %d random (chain, t, node) vs the C oracle" % pairs,
            "max_rel_err": worst, "tolerance": 1e-10, "ok": bool(worst <= 1e-10)}


OTHERS = (("cfg2", 60), ("cfg3", 30), ("cfg5", 8))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true",
                    help="only the headline workload (default: cfg2 / cfg3 / cfg5 results nested under 'others')")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    ctx = Ctx()
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    numa_cpus = pin_to_numa_node(ctx.local)
    torch.cuda.set_device(ctx.local)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
    ctx.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    head = measure(ctx, args.workload, args.steps, args.warmup, args.chains_per_gpu,
                   e2e=not args.no_e2e, cpu=not args.no_cpu_baseline)
    others = {}
    if not args.no_others and not args.chains_per_gpu:
        for name, st in OTHERS:
            if name == args.workload:
                continue
            r = measure(ctx, name, min(st, max(3, args.steps)), args.warmup, e2e=not args.no_e2e,
                        cpu=not args.no_cpu_baseline and name != "cfg5" and name != "cfg3")
            if r is not None:
                others[name] = r
    if ctx.rank == 0:
        line = {"metric": "latent-position node-updates/sec", "n_gpus": ctx.world,
                "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        line.update(head)
        line["numa_cpus_bound"] = numa_cpus
        if others:
            line["others"] = others
        print(json.dumps(line))
    if ctx.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
