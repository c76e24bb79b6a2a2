#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c28}
( time timeout 900 python -m pytest tests/test_gpu_rowsum_cache.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
timeout 300 python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 --chains-per-gpu 148 $AB > gpurun_out/${TAG}_ab_cfg2c148.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_cfg2c1.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg1 --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_cfg1c1.json 2>> gpurun_out/${TAG}_ab.err
