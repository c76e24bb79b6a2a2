#!/bin/bash
# dataflow case-control sweep: L2 hints and chains-per-launch A/B at cfg 5; scalar-MH kernels re-check
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c12}
( time timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_operating_points.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 10 --warmup 3"
for H in 1 0; do for G in 8 4 2; do
DLSM_CCD_HINTS=$H DLSM_CCD_GROUP=$G timeout 300 python bench.py --workload cfg5 $AB > gpurun_out/${TAG}_ab_cfg5_h${H}_g${G}.json 2>> gpurun_out/${TAG}_ab.err
done; done
timeout 300 python bench.py --workload cfg3 $AB > gpurun_out/${TAG}_ab_cfg3.json 2>> gpurun_out/${TAG}_ab.err
du -sh gpurun_out
