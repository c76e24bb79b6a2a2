import sys, cProfile, pstats, warnings
sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
import bench
from dynetlsm_b200 import DynamicNetworkHDPLPCM, DynamicNetworkLSM
w = bench.make_workload("cfg2")
m = DynamicNetworkHDPLPCM(n_components=10, n_iter=1000, tune=500, burn=500, random_state=42)
pr = cProfile.Profile(); pr.enable(); m.fit(w["Y"]); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
