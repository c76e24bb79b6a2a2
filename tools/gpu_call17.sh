#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c17}
( time timeout 1200 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_operating_points.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 10 --warmup 3"
timeout 300 python bench.py --workload cfg5 $AB > gpurun_out/${TAG}_ab_cfg5.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CCD_THREADS=768 timeout 300 python bench.py --workload cfg5 $AB > gpurun_out/${TAG}_ab_cfg5_t768.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CCD_THREADS=768 DLSM_CCD_GROUP=8 timeout 300 python bench.py --workload cfg5 $AB > gpurun_out/${TAG}_ab_cfg5_t768_g8.json 2>> gpurun_out/${TAG}_ab.err
tail -c 300 gpurun_out/${TAG}_ab_cfg5.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ccd -s 24 -c 1 -o /tmp/prof_cfg5 \
   python bench.py --workload cfg5 --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_cfg5.log 2>&1
ncu -i /tmp/prof_cfg5.ncu-rep --page details > gpurun_out/${TAG}_prof_cfg5_details.txt 2>/dev/null
ncu -i /tmp/prof_cfg5.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_cfg5_raw.csv 2>/dev/null
ncu -i /tmp/prof_cfg5.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_cfg5_source.csv 2>/dev/null
