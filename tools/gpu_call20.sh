#!/bin/bash
# cfg 3: phase timing of the cluster sweep kernel (developer build with clock64 probes)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c20}
AB="--no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3"
DLSM_LIB=$PWD/variants/libdlsm_blktime.so timeout 300 python bench.py --workload cfg3 $AB > gpurun_out/${TAG}_blktime.log 2>&1
grep "blk timing" gpurun_out/${TAG}_blktime.log | tail -8
timeout 300 python bench.py --workload cfg5 --no-others --no-e2e --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_ab_cfg5.json 2>> gpurun_out/${TAG}_ab.err
