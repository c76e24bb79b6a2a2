"""Wall-clock of the public estimators' fit() on the bench workloads (sweeps/s through the user API,
traces included).  usage: python tools/fit_timing.py [lsm|hdp|lpcm] [n_iter] [n_chains]"""
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
import bench  # noqa: E402
from dynetlsm_b200 import DynamicNetworkHDPLPCM, DynamicNetworkLPCM, DynamicNetworkLSM  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "lsm"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 1
if which == "lsm":
    w = bench.make_workload("cfg1")
    m = DynamicNetworkLSM(n_iter=n_iter, tune=n_iter // 2, burn=n_iter // 2, random_state=42, n_chains=chains)
elif which == "lpcm":    # finite mixture: device hot path + host Dirichlet / conjugate block per sweep
    w = bench.make_workload("cfg2")
    m = DynamicNetworkLPCM(n_components=10, n_iter=n_iter, tune=n_iter // 2, burn=n_iter // 2, random_state=42)
else:
    w = bench.make_workload("cfg2")
    m = DynamicNetworkHDPLPCM(n_components=10, n_iter=n_iter, tune=n_iter // 2, burn=n_iter // 2,
                              random_state=42, n_chains=chains)
t0 = time.perf_counter()
m.fit(w["Y"])
dt = time.perf_counter() - t0
S = 2 * n_iter
T, n = w["Y"].shape[:2]
print("fit_timing %s: %d sweeps x %d chains in %.2f s -> %.1f sweeps/s, %.3g node-updates/s (logp_=%.4f)"
      % (which, S, chains, dt, S / dt, S * chains * T * n / dt, float(np.ravel(m.logp_)[0])))
