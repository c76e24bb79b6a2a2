#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c23}
( time timeout 900 python -m pytest tests/test_gpu_rowsum_cache.py tests/test_gpu_operating_points.py tests/test_gpu_estimators.py tests/test_gpu_trace.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 30 --warmup 3"
timeout 300 python bench.py --workload cfg3 $AB > gpurun_out/${TAG}_ab_cfg3.json 2>> gpurun_out/${TAG}_ab.err
tail -c 300 gpurun_out/${TAG}_ab_cfg3.json
