"""Wall-clock of the UNMODIFIED reference's ``DynamicNetworkLPCM.fit`` main loop on the cfg-2 network
(n = 120, T = 9, K = 10), in THIS container (the reference tree does not travel to the GPU box).
TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The 999-sweep LSM initialisation is timed separately.

  python oracle/time_ref_lpcm.py [n_iter]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_shims  # noqa: E402


def main():
    n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ref_shims.load_reference()
    import dynetlsm.lpcm as P
    import dynetlsm.lsm as L
    import workloads
    w = workloads.make_workload("cfg2")
    Y = np.ascontiguousarray(w["Y"], dtype=np.float64)
    L.tqdm = P.tqdm = lambda it, *a, **k: it
    marks = {}
    orig_fit = P.DynamicNetworkLPCM._fit

    def _fit(self, Y_, rng):
        marks["loop0"] = time.perf_counter()
        return orig_fit(self, Y_, rng)
    P.DynamicNetworkLPCM._fit = _fit
    t0 = time.perf_counter()
    m = P.DynamicNetworkLPCM(n_components=10, n_iter=n_iter, tune=n_iter // 2, burn=n_iter // 2,
                             random_state=42).fit(Y)
    t1 = time.perf_counter()
    S = m.n_iter - 1
    T, n = Y.shape[:2]
    loop = t1 - marks["loop0"]
    print("reference LPCM (1 host core, this container): init %.1f s; main loop %d sweeps in %.2f s -> "
          "%.2f sweeps/s, %.3g node-updates/s (incl. post-processing)"
          % (marks["loop0"] - t0, S, loop, S / loop, S * T * n / loop))


if __name__ == "__main__":
    main()
