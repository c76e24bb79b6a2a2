#!/bin/bash
# two-block-window cluster kernel: parity, A/B at cfg 3
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c21}
( time timeout 600 python -m pytest tests/test_gpu_rowsum_cache.py -m gpu -q -x -k "cluster" ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
timeout 300 python bench.py --workload cfg3 $AB > gpurun_out/${TAG}_ab_cfg3_base.json 2>> gpurun_out/${TAG}_ab.err
DLSM_NO_CLUSTER=3 timeout 300 python bench.py --workload cfg3 $AB > gpurun_out/${TAG}_ab_cfg3_win.json 2>> gpurun_out/${TAG}_ab.err
tail -c 400 gpurun_out/${TAG}_ab_cfg3_win.json
