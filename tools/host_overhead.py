"""Wall clock per sweep of dlsm_run_sweeps for single-chain workloads (host launch overhead vs GPU time)."""
import sys, time
sys.path.insert(0, ".")
import bench
for name, n_sw in (("cfg1", 4000), ("cfg2", 2000)):
    w = bench.make_workload(name)
    e = bench.build_engine(w, 1, 0, 0)
    e.run_sweeps(50)
    t0 = time.perf_counter(); e.run_sweeps(n_sw); dt = time.perf_counter() - t0
    e.enable_timing(True); c0 = e.counters(); e.run_sweeps(200); c1 = e.counters(); e.enable_timing(False)
    gpu = (c1["latent_ms"] - c0["latent_ms"] + c1["other_ms"] - c0["other_ms"]) / 200 * 1e3
    print("%s 1 chain: %.1f us/sweep wall, %.1f us/sweep in timed GPU phases, %d launches/sweep"
          % (name, dt / n_sw * 1e6, gpu, (c1["kernel_launches"] - c0["kernel_launches"]) / 200))
