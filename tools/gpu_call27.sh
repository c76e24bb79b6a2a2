#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c27}
AB="--no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
git -C . show 08d6c35:bench.py > /tmp/bench_old.py 2>/dev/null || cp bench.py /tmp/bench_old.py
DLSM_LIB=$PWD/variants/libdlsm_old.so timeout 300 python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_oldlib.json 2>> gpurun_out/${TAG}_ab.err
DLSM_LIB=$PWD/variants/libdlsm_old.so timeout 300 python bench.py --workload cfg2 --chains-per-gpu 148 $AB > gpurun_out/${TAG}_ab_cfg2c148_oldlib.json 2>> gpurun_out/${TAG}_ab.err
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt
