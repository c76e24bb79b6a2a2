#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c26}
AB="--no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
DLSM_LIB=$PWD/variants/libdlsm_nominb1.so timeout 300 python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_nominb1.json 2>> gpurun_out/${TAG}_ab.err
DLSM_LIB=$PWD/variants/libdlsm_nominb1.so timeout 300 python bench.py --workload cfg2 --chains-per-gpu 148 $AB > gpurun_out/${TAG}_ab_cfg2c148_nominb1.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 --chains-per-gpu 148 $AB > gpurun_out/${TAG}_ab_cfg2c148_base.json 2>> gpurun_out/${TAG}_ab.err
DLSM_LIB=$PWD/variants/libdlsm_nominb1.so timeout 300 python bench.py --workload cfg2 --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_cfg2c1_nominb1.json 2>> gpurun_out/${TAG}_ab.err
timeout 300 python bench.py --workload cfg2 --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_cfg2c1_base.json 2>> gpurun_out/${TAG}_ab.err
