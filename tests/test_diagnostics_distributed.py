"""Host-side multi-GPU logic on CPU: chain sharding and the pooling of per-chain scalar traces
(the only collective of the sampler) with world_size=2 over gloo; R-hat / ESS numerics."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from dynetlsm_b200.diagnostics import pool_traces, summarize
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # ragged sharding: rank 0 owns 3 chains, rank 1 owns 2; chain ids are global
    counts = [3, 2]
    first = sum(counts[:rank])
    rng_all = np.random.RandomState(0)
    full = rng_all.randn(sum(counts), 200, 2) + np.arange(sum(counts))[:, None, None] * 0.01
    local = full[first:first + counts[rank]]
    pooled = pool_traces(local)
    assert pooled.shape == full.shape and np.array_equal(pooled, full)
    s = summarize(pooled, names=["logp", "intercept"])
    np.save(os.path.join(out_dir, "rhat_%d.npy" % rank), np.array([s["logp"]["rhat"], s["intercept"]["ess"]]))
    dist.destroy_process_group()


def test_pool_traces_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rhat_0.npy")
    b = np.load(tmp_path / "rhat_1.npy")
    assert np.array_equal(a, b)          # every rank sees the same pooled diagnostics
    assert 0.98 < a[0] < 1.05 and a[1] > 500


def test_pool_traces_without_process_group_is_identity():
    from dynetlsm_b200.diagnostics import pool_traces
    x = np.random.RandomState(1).randn(4, 10, 3)
    assert np.array_equal(pool_traces(x), x)


def test_rhat_and_ess_behave():
    from dynetlsm_b200.diagnostics import ess, geweke_z, split_rhat
    rng = np.random.RandomState(3)
    iid = rng.randn(4, 2000)
    assert abs(split_rhat(iid) - 1.0) < 0.01
    assert 6000 < ess(iid) < 10000
    shifted = iid + np.array([0.0, 0.0, 3.0, 3.0])[:, None]
    assert split_rhat(shifted) > 1.5
    ar = np.zeros((2, 4000))
    for c in range(2):
        e = rng.randn(4000)
        for t in range(1, 4000):
            ar[c, t] = 0.9 * ar[c, t - 1] + e[t]
    assert ess(ar) < 0.15 * ar.size      # AR(1) with rho 0.9: ESS ~ N (1-rho)/(1+rho)
    assert abs(geweke_z(iid[0])) < 4
    drift = iid[0] + np.linspace(0, 5, 2000)
    assert abs(geweke_z(drift)) > 4


def test_chain_ids_are_global_under_sharding():
    """bench.py / the estimators give rank r the chain ids r*C .. r*C+C-1 (Philox counters carry the
    global id), so a chain's stream does not depend on how chains are split over GPUs."""
    C_, world = 5, 2
    ids = [list(range(r * C_, r * C_ + C_)) for r in range(world)]
    assert sorted(sum(ids, [])) == list(range(world * C_))
