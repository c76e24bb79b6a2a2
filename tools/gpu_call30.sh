#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c30}
( time timeout 900 python -m pytest tests/test_gpu_rowsum_cache.py -m gpu -q -x -k "cluster" ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log
