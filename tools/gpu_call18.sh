#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c18}
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 10 --warmup 3"
timeout 300 python bench.py --workload cfg5 $AB > gpurun_out/${TAG}_ab_cfg5.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CCD_THREADS=768 timeout 300 python bench.py --workload cfg5 $AB > gpurun_out/${TAG}_ab_cfg5_t768.json 2>> gpurun_out/${TAG}_ab.err
tail -c 300 gpurun_out/${TAG}_ab_cfg5.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_cfg5.csv \
   python bench.py --workload cfg5 --no-others --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_launch.log 2>&1
