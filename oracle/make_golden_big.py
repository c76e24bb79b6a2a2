"""Known answers at the OPERATING POINTS (tests/golden/kernels_big.npz): >= 1 000 (state, node)
pairs per likelihood at n = 500 (cfg 4), n = 2 000 (cfg 3, directed) and n = 50 000 (cfg 5,
case-control), plus the full-network values, computed by the REFERENCE's own Cython kernels
(oracle/_ref, compiled from /root/reference by oracle/build_ref.py) -- SURVEY.md 8(d)(i).

TEST INFRASTRUCTURE ONLY; runs in the build container (python oracle/make_golden_big.py).
The inputs are not stored: they are `workloads.make_workload(cfg)` (deterministic under this
image's numpy / scipy / scikit-learn) perturbed by a seeded normal, so the fixture holds only the
seeds, the sampled (t, node) indices and the reference's answers (a few tens of KB).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "kernels_big.npz")

import ref_shims  # noqa: E402
import workloads as W  # noqa: E402

PAIRS = 1024


def state_of(w, seed):
    """The perturbed state the known answers refer to (also used by the tests)."""
    rng = np.random.RandomState(seed)
    scale = 1.0 / w["n"] if w["directed"] else 1.0
    X = np.ascontiguousarray(w["X"] + 0.05 * scale * rng.randn(*w["X"].shape))
    cells = rng.permutation(w["T"] * w["n"])[:PAIRS]        # distinct (t, node) pairs
    return X, cells // w["n"], cells % w["n"]


def main():
    ref_shims.load_reference()
    from dynetlsm.static_network_fast import partial_loglikelihood
    from dynetlsm.directed_likelihoods_fast import (directed_partial_loglikelihood,
                                                    approx_directed_partial_loglikelihood,
                                                    approx_directed_network_loglikelihood)
    from dynetlsm.network_likelihoods import (dynamic_network_loglikelihood_undirected,
                                              dynamic_network_loglikelihood_directed)
    out = {}
    for name, seed in (("cfg2", 11), ("cfg4", 12)):
        w = W.make_workload(name)
        X, t, j = state_of(w, seed)
        b = float(w["intercept"][0])
        out[name + "_seed"] = seed
        out[name + "_t"], out[name + "_j"] = t, j
        out[name + "_k1"] = np.array([partial_loglikelihood(w["Y"][a], X[a], b, int(c)) for a, c in zip(t, j)])
        out[name + "_k5"] = dynamic_network_loglikelihood_undirected(w["Y"], X, b)
        print(name, "K1", out[name + "_k1"][:3], "K5", out[name + "_k5"])
    w = W.make_workload("cfg3")
    X, t, j = state_of(w, 13)
    bi, bo = float(w["intercept"][0]), float(w["intercept"][1])
    out["cfg3_seed"] = 13
    out["cfg3_t"], out["cfg3_j"] = t, j
    out["cfg3_k2"] = np.array([directed_partial_loglikelihood(w["Y"][a], X[a], w["radii"], bi, bo, int(c))
                               for a, c in zip(t, j)])
    out["cfg3_k4"] = dynamic_network_loglikelihood_directed(w["Y"], X, bi, bo, w["radii"])
    print("cfg3 K2", out["cfg3_k2"][:3], "K4", out["cfg3_k4"])
    w = W.make_workload("cfg5")
    X, t, j = state_of(w, 14)
    ie, oe, dg = (w[k].astype(np.int64) for k in ("in_edges", "out_edges", "degrees"))
    ci, co = w["ctrl_in"].astype(np.int64), w["ctrl_out"].astype(np.int64)
    out["cfg5_seed"] = 14
    out["cfg5_t"], out["cfg5_j"] = t, j
    out["cfg5_k3"] = np.array([approx_directed_partial_loglikelihood(
        X[a], radii=w["radii"], in_edges=ie[a], out_edges=oe[a], degree=dg[a], control_nodes_in=ci[a],
        control_nodes_out=co[a], intercept_in=0.3, intercept_out=0.7, node_id=int(c), squared=False)
        for a, c in zip(t, j)])
    out["cfg5_k6"] = approx_directed_network_loglikelihood(
        X, radii=w["radii"], in_edges=ie, out_edges=oe, degree=dg, control_nodes=co, intercept_in=0.3,
        intercept_out=0.7, squared=False)
    print("cfg5 K3", out["cfg5_k3"][:3], "K6", out["cfg5_k6"])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
