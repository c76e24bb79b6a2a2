"""Convergence diagnostics over independent chains, and their pooling across GPUs.

The reference has a single-chain notion only (trace_utils.py: ESS by autocorrelation, Geweke's z
with an AR spectral estimate that needs statsmodels).  With chains sharded over the GPUs of one
box the natural diagnostics are split-R-hat and pooled ESS; the only cross-GPU
traffic of the whole sampler is the all-gather of per-chain SCALAR traces done here
(``pool_traces``; O(chains x draws x few doubles), NCCL over NVLink when the process group is
NCCL, gloo on CPU in the tests).
"""
import numpy as np

__all__ = ["autocorr", "ess", "split_rhat", "geweke_z", "geweke_zp", "pool_traces", "summarize"]


def autocorr(x):
    """Autocorrelation function of a 1-D trace via FFT (biased estimator, lag 0 = 1)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    xc = x - x.mean()
    f = np.fft.rfft(xc, 2 * n)
    acov = np.fft.irfft(f * np.conj(f))[:n] / n
    return acov / acov[0] if acov[0] > 0 else np.ones(n)


def ess(chains):
    """Effective sample size of (C, S) draws (Geyer initial positive sequence on the chain-averaged
    autocorrelations, as in Gelman et al. BDA3 11.5)."""
    chains = np.atleast_2d(np.asarray(chains, dtype=np.float64))
    C, S = chains.shape
    if S < 4:
        return float(C * S)
    W = chains.var(axis=1, ddof=1).mean()
    B_over_n = chains.mean(axis=1).var(ddof=1) if C > 1 else 0.0
    var_plus = W * (S - 1) / S + B_over_n
    if var_plus <= 0:
        return float(C * S)
    acov = np.array([autocorr(c) * c.var() for c in chains]).mean(axis=0)
    rho = 1 - (W - acov) / var_plus
    tau, t = -1.0, 0
    while t + 1 < S:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        tau += 2 * pair
        t += 2
    return float(C * S / max(tau, 1e-12)) if tau > 0 else float(C * S)


def split_rhat(chains):
    """Split-R-hat of (C, S) draws: every chain is halved, so it is defined for C = 1 too."""
    chains = np.atleast_2d(np.asarray(chains, dtype=np.float64))
    S = chains.shape[1] // 2
    if S < 2:
        return float("nan")
    halves = np.concatenate([chains[:, :S], chains[:, S:2 * S]], axis=0)
    W = halves.var(axis=1, ddof=1).mean()
    B = S * halves.mean(axis=1).var(ddof=1)
    if W <= 0:
        return float("nan")
    return float(np.sqrt(((S - 1) / S * W + B / S) / W))


def geweke_z(x, n_burn=0, first=0.1, last=0.5):
    """Geweke's z: early-vs-late mean difference over its standard error, with the spectral
    density at zero estimated by a Bartlett-windowed autocovariance (the reference uses an AR fit
    from statsmodels, trace_utils.py:59-115)."""
    x = np.asarray(x, dtype=np.float64)[n_burn:]
    n = x.size
    if n < 20:
        return float("nan")
    a, b = x[:int(first * n)], x[n - int(last * n):]

    def s0(v):
        v = v - v.mean()
        m = v.size
        L = max(1, int(np.floor(4 * (m / 100.0) ** (2.0 / 9.0))))
        acov = np.array([np.dot(v[:m - k], v[k:]) / m for k in range(L + 1)])
        return acov[0] + 2 * np.sum((1 - np.arange(1, L + 1) / (L + 1)) * acov[1:])
    den = s0(a) / a.size + s0(b) / b.size
    return float((a.mean() - b.mean()) / np.sqrt(den)) if den > 0 else float("nan")


def geweke_zp(x, n_burn=0, first=0.1, last=0.5):
    """(z, two-sided p-value) as the reference's ``geweke_diag`` returns them (trace_utils.py:59-115)."""
    from math import erfc, sqrt
    z = geweke_z(x, n_burn, first, last)
    return z, (erfc(abs(z) / sqrt(2.0)) if z == z else float("nan"))


def pool_traces(local, group=None):
    """All-gather per-chain scalar traces over the ranks of a torch.distributed process group.

    ``local``: (C_local, S, P) float64 array (P scalar parameters: logp, intercepts, lambda, ...).
    Returns the (C_total, S, P) array on every rank.  Without an initialised process group the
    input is returned unchanged (single GPU)."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    try:
        import torch
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return local
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(local).to(dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=dev), group=group)
    counts = [int(c.item()) for c in counts]
    cmax = max(counts)
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=torch.float64, device=dev)
    pad[:local.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


def summarize(traces, names=None, n_burn=0):
    """R-hat / ESS / mean / sd per scalar parameter of pooled (C, S, P) traces."""
    traces = np.asarray(traces, dtype=np.float64)[:, n_burn:]
    P = traces.shape[2]
    names = names or ["p%d" % i for i in range(P)]
    return {nm: dict(mean=float(traces[:, :, p].mean()), sd=float(traces[:, :, p].std(ddof=1)),
                     rhat=split_rhat(traces[:, :, p]), ess=ess(traces[:, :, p]))
            for p, nm in enumerate(names)}
