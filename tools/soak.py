"""Long device-resident runs: no non-finite acceptance ratio, log-posterior stays finite, tuner settles."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bench
from dynetlsm_b200 import _lib as L
for name, chains, sweeps in (("cfg2", 8, 20000), ("cfg1", 4, 50000), ("cfg4", 4, 1500), ("cfg5", 2, 60)):
    w = bench.make_workload(name)
    e = bench.build_engine(w, chains, 0, 0)
    t0 = time.perf_counter()
    tr = e.run_traced(sweeps, fields_all=(L.F_INTERCEPT,), thin=max(1, sweeps // 200))
    dt = time.perf_counter() - t0
    lp = tr["logp"]
    acc = e.get(L.F_X_NACC).sum() / max(1, e.get(L.F_X_NSTEPS).sum())
    print("%s: %d sweeps x %d chains in %.1f s, logp finite=%s, logp first/last %.1f / %.1f, latent acceptance %.2f, "
          "intercept %.3f -> %.3f" % (name, sweeps, chains, dt, np.isfinite(lp).all(), lp[0].mean(), lp[-1].mean(),
                                      acc, tr[L.F_INTERCEPT][0, 0, 0], tr[L.F_INTERCEPT][-1, 0, 0]))
    e.close()
