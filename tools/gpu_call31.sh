#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c31}
AB="--no-others --no-e2e --no-cpu-baseline --steps 40 --warmup 3"
timeout 300 python bench.py --workload cfg2 $AB > gpurun_out/${TAG}_ab_cfg2_base.json 2>> gpurun_out/${TAG}_ab.err
DLSM_LIB=$PWD/variants/libdlsm_hdp9.so timeout 300 python bench.py --workload cfg2 $AB > gpurun_out/${TAG}_ab_cfg2_hdp9.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 $AB > gpurun_out/${TAG}_ab_cfg2_base2.json 2>> gpurun_out/${TAG}_ab.err
DLSM_LIB=$PWD/variants/libdlsm_hdp9.so timeout 300 python bench.py --workload cfg2 $AB > gpurun_out/${TAG}_ab_cfg2_hdp9b.json 2>> gpurun_out/${TAG}_ab.err
