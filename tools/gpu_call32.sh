#!/bin/bash
# two warps per slice for chains that have an SM to themselves (k_sweep_cbp): parity, A/B
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c32}
( time timeout 900 python -m pytest tests/test_gpu_rowsum_cache.py -m gpu -q -x -k "block_chain" ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
timeout 300 python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_pair.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CHAIN_KERNEL=block1 timeout 300 python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_single.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 --chains-per-gpu 148 $AB > gpurun_out/${TAG}_ab_cfg2c148_pair.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_cfg2c1_pair.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg1 --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_cfg1c1_pair.json 2>> gpurun_out/${TAG}_ab.err
