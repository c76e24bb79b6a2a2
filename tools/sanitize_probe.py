"""Small instances of every sweep kernel, run under compute-sanitizer by tools/sanitize.sh."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynetlsm_b200 import _lib as L  # noqa: E402


def net(T, n, directed, seed):
    rng = np.random.RandomState(seed)
    Y = (rng.rand(T, n, n) < 0.2).astype(np.float64)
    for t in range(T):
        np.fill_diagonal(Y[t], 0)
    if not directed:
        Y = np.triu(Y, 1); Y = Y + Y.transpose(0, 2, 1)
    return rng, Y


def run(tag, T, n, C_, directed, opts, sweeps=2, K=0):
    rng, Y = net(T, n, directed, 1)
    e = L.Engine(T=T, n=n, d=2, n_chains=C_, is_directed=directed, K=K, mixture=K > 0)
    for k, v in opts:
        e.set_option(k, v)
    e.set_network(Y)
    scale = 1.0 / n if directed else 1.0
    e.set(L.F_X, rng.randn(C_, T, n, 2) * scale)
    e.set(L.F_INTERCEPT, np.tile([[0.5, 0.3]], (C_, 1)))
    if directed:
        e.set(L.F_RADII, rng.dirichlet(np.ones(n) * 4, size=C_))
        e.set_hyper(tau_sq=scale ** 2, sigma_sq=1e-3 * scale)
    e.set_tuner(0.1 * scale)
    if K:
        e.set(L.F_MU, rng.randn(C_, K, 2)); e.set(L.F_SIGMA, rng.gamma(2, 1, (C_, K)))
        e.set(L.F_LAMBDA, np.full(C_, 0.8)); e.set(L.F_WEIGHTS, rng.dirichlet(np.ones(K), size=(C_, T, K)))
        e.set(L.F_Z, rng.randint(0, K, (C_, T, n)))
    e.set_rng(3)
    e.run_sweeps(sweeps, skip_hdp=True)
    x = e.get(L.F_X)
    print(tag, "ok", float(np.abs(x).sum()), e.counters()["cluster_sweeps"], e.counters()["rowsum_sweeps"])
    e.close()


def run_cc(tag, T, n, C_, cc_kernel=0):
    rng = np.random.RandomState(2)
    deg = np.zeros((T, n, 2), np.int32)
    out_e = np.zeros((T, n, 6), np.int32)
    ins = [[[] for _ in range(n)] for _ in range(T)]
    for t in range(T):
        for i in range(n):
            k = rng.randint(1, 6)
            nb = rng.choice([j for j in range(n) if j != i], k, replace=False)
            out_e[t, i, :k] = nb; deg[t, i, 1] = k
            for j in nb:
                ins[t][j].append(i)
    mi = max(len(v) for r in ins for v in r)
    in_e = np.zeros((T, n, mi), np.int32)
    for t in range(T):
        for j in range(n):
            in_e[t, j, :len(ins[t][j])] = ins[t][j]; deg[t, j, 0] = len(ins[t][j])
    e = L.Engine(T=T, n=n, d=2, n_chains=C_, is_directed=True, case_control=True)
    e.set_option(L.OPT_SWEEP_MODE, L.SWEEP_SLICE)
    e.set_option(L.OPT_CC_KERNEL, cc_kernel)
    e.set_edge_lists(deg, in_e, out_e)
    e.set_rng(5)
    e.resample_controls(8, per_chain=True)
    e.set(L.F_X, rng.randn(C_, T, n, 2) / n); e.set(L.F_INTERCEPT, np.tile([[0.4, 0.6]], (C_, 1)))
    e.set(L.F_RADII, rng.dirichlet(np.ones(n) * 4, size=C_))
    e.set_hyper(tau_sq=1.0 / n ** 2, sigma_sq=1e-3 / n)
    e.set_tuner(0.05 / n)
    e.run_sweeps(2)
    print(tag, "ok", float(np.abs(e.get(L.F_X)).sum()))
    e.close()


if __name__ == "__main__":
    CH, SL = (L.OPT_SWEEP_MODE, L.SWEEP_CHAIN), (L.OPT_SWEEP_MODE, L.SWEEP_SLICE)
    run("k_sweep (node, two-variant)", 4, 70, 3, False, [CH, (L.OPT_CHAIN_KERNEL, L.CHAIN_NODE)], K=4)
    run("k_sweep (row-sum cache) + k_rows", 4, 70, 3, False, [CH, (L.OPT_CHAIN_KERNEL, L.CHAIN_NODE_ROWSUM)])
    run("k_sweep_cb (block chain kernel)", 4, 70, 3, True, [CH, (L.OPT_CHAIN_KERNEL, L.CHAIN_BLOCK)])
    run("k_sweep_blkw (cluster, block-speculative, two-block window)", 3, 130, 1, True, [SL, (L.OPT_NO_CLUSTER, 0)])
    run("k_sweep_blk (cluster, block-speculative)", 3, 130, 1, True, [SL, (L.OPT_NO_CLUSTER, 3)])
    run("k_sweep_slice_cl (cluster, per node)", 3, 130, 1, False, [SL, (L.OPT_NO_CLUSTER, 2)])
    run("k_sweep_slice_ws (CTA per slice)", 2, 200, 1, True, [SL, (L.OPT_NO_CLUSTER, 1)])
    run_cc("k_sweep_cc (case-control batches)", 2, 60, 2, cc_kernel=1)
    run_cc("k_sweep_cc3 (case-control batches, 2-CTA cluster)", 2, 60, 2, cc_kernel=3)
    run_cc("k_sweep_ccd (case-control dataflow) + k_ccd_prep / k_ccd_post", 2, 60, 2, cc_kernel=0)
