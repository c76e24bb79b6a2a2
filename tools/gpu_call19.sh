#!/bin/bash
# k_rows at three CTAs per SM (A/B build), sanitizer over all sweep kernels incl. the dataflow kernel
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c19}
AB="--no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
timeout 300 python bench.py --workload cfg4 $AB > gpurun_out/${TAG}_ab_cfg4_base.json 2>> gpurun_out/${TAG}_ab.err
DLSM_LIB=$PWD/variants/libdlsm_rows3.so timeout 300 python bench.py --workload cfg4 $AB > gpurun_out/${TAG}_ab_cfg4_rows3.json 2>> gpurun_out/${TAG}_ab.err
tail -c 200 gpurun_out/${TAG}_ab_cfg4_rows3.json
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py > gpurun_out/${TAG}_sanitize_$tool.log 2>&1
  tail -n 4 gpurun_out/${TAG}_sanitize_$tool.log
done
