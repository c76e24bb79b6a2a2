#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c25}
AB="--no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
timeout 300 python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg4 --chains-per-gpu 256 $AB > gpurun_out/${TAG}_ab_cfg4c256.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches_cfg4c128.csv \
   python bench.py --workload cfg4 --chains-per-gpu 128 --no-others --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_launch.log 2>&1
