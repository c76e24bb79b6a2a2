#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c48_smoke.log 2>&1
tail -n 1 gpurun_out/c48_smoke.log
( timeout 200 python -m pytest tests/test_gpu_lpcm.py -m gpu -q -k case_control ) > gpurun_out/c48_pytest.log 2>&1
grep -v Warning gpurun_out/c48_pytest.log | tail -30
timeout 100 python - > gpurun_out/c48_hdp_cc.log 2>&1 <<'P'
import sys, warnings
import numpy as np
sys.path.insert(0, "tests"); warnings.filterwarnings("ignore")
from conftest import load_golden
from dynetlsm_b200 import DynamicNetworkHDPLPCM
g = load_golden("lsm_casecontrol_monks.npz")
Y = g["Y"].astype(np.float64)
for s in ("device", "replay"):
    m = DynamicNetworkHDPLPCM(n_iter=20, tune=20, burn=10, tune_interval=6, n_components=4, is_directed=True,
                              n_control=5, n_resample_control=8, random_state=11, sampler=s).fit(Y)
    print("hdp case-control", s, m.Xs_.shape, np.isfinite(m.logps_).all(), m.sampler_counters_["ub_flags"])
P
tail -5 gpurun_out/c48_hdp_cc.log
