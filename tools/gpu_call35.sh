#!/bin/bash
# k_full_lr (lanes = rows): parity vs the folded-rows kernel, A/B at cfg 3 and cfg 4 / 128 chains
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c35}
( time timeout 900 python -m pytest tests/test_gpu_rowsum_cache.py tests/test_gpu_parity.py tests/test_gpu_operating_points.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 30 --warmup 3"
timeout 300 python bench.py --workload cfg3 $AB > gpurun_out/${TAG}_ab_cfg3_lr.json 2>> gpurun_out/${TAG}_ab.err
DLSM_FULL_KERNEL=1 timeout 300 python bench.py --workload cfg3 $AB > gpurun_out/${TAG}_ab_cfg3_folded.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_lr.json 2>> gpurun_out/${TAG}_ab.err
DLSM_FULL_KERNEL=1 timeout 300 python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_folded.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 --chains-per-gpu 148 $AB > gpurun_out/${TAG}_ab_cfg2c148_lr.json 2>> gpurun_out/${TAG}_ab.err
DLSM_FULL_KERNEL=1 timeout 300 python bench.py --workload cfg2 --chains-per-gpu 148 $AB > gpurun_out/${TAG}_ab_cfg2c148_folded.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 $AB > gpurun_out/${TAG}_ab_cfg2_lr.json 2>> gpurun_out/${TAG}_ab.err
DLSM_FULL_KERNEL=1 timeout 300 python bench.py --workload cfg2 $AB > gpurun_out/${TAG}_ab_cfg2_folded.json 2>> gpurun_out/${TAG}_ab.err
