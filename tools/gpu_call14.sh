#!/bin/bash
# cfg 5: ncu of the full-network case-control kernel (k_full<2,2,1>) + new scalar-MH test
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c14}
( time timeout 900 python -m pytest tests/test_gpu_edge_cases.py -m gpu -q -x -k "long_chain_case_control" ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
export DLSM_CCD_GROUP=4 DLSM_CCD_LOOKAHEAD=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_full -s 20 -c 1 -o /tmp/prof_full \
   python bench.py --workload cfg5 --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_full.log 2>&1
ncu -i /tmp/prof_full.ncu-rep --page details > gpurun_out/${TAG}_prof_full_details.txt 2>/dev/null
ncu -i /tmp/prof_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_full_raw.csv 2>/dev/null
ncu -i /tmp/prof_full.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_full_source.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_cfg5.csv \
   python bench.py --workload cfg5 --no-others --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_launch.log 2>&1
du -sh gpurun_out
