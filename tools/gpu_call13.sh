#!/bin/bash
# dataflow case-control sweep v2 (self-validating records): parity + A/B at cfg 5
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c13}
( time timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_operating_points.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 10 --warmup 3"
for V in "2 8 -1" "2 4 -1" "2 8 0" "2 4 0" "1 4 0"; do set -- $V
DLSM_CCD_VERSION=$1 DLSM_CCD_GROUP=$2 DLSM_CCD_LOOKAHEAD=$3 timeout 300 python bench.py --workload cfg5 $AB > gpurun_out/${TAG}_ab_cfg5_v$1_g$2_l$3.json 2>> gpurun_out/${TAG}_ab.err
done
DLSM_CCD_GROUP=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ccd2 -s 24 -c 1 -o /tmp/prof_cfg5 \
   python bench.py --workload cfg5 --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_cfg5.log 2>&1
ncu -i /tmp/prof_cfg5.ncu-rep --page details > gpurun_out/${TAG}_prof_cfg5_details.txt 2>/dev/null
ncu -i /tmp/prof_cfg5.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_cfg5_raw.csv 2>/dev/null
ncu -i /tmp/prof_cfg5.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_cfg5_source.csv 2>/dev/null
du -sh gpurun_out
