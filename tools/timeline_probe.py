"""Developer probe: per-launch-group device timeline of dlsm_run_sweeps (DLSM_TIMELINE=1 prints it
to stderr).  usage: DLSM_TIMELINE=1 python tools/timeline_probe.py [workload] [chains] [mode]
mode: all | nolabels | nointercepts"""
import sys
sys.path.insert(0, ".")
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name].get("chains_per_gpu", 1332)
mode = sys.argv[3] if len(sys.argv) > 3 else "all"
w = bench.make_workload(name)
e = bench.build_engine(w, chains, 0, 0)
kw = dict(skip_labels=mode == "nolabels", skip_intercepts=mode == "nointercepts")
e.run_sweeps(5, **kw)
sys.stderr.write("---- %s %s ----\n" % (name, mode))
e.run_sweeps(2, **kw)
