#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c38}
AB="--no-others --no-e2e --no-cpu-baseline --steps 40 --warmup 3"
for w in cfg2 cfg4; do
timeout 300 python bench.py --workload $w $AB > gpurun_out/${TAG}_ab_${w}_gstage.json 2>> gpurun_out/${TAG}_ab.err
DLSM_FFBS_SMEM=1 timeout 300 python bench.py --workload $w $AB > gpurun_out/${TAG}_ab_${w}_smem.json 2>> gpurun_out/${TAG}_ab.err
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ffbs -c 12 --csv --log-file gpurun_out/${TAG}_ffbs_gstage.csv \
   python bench.py --workload cfg2 --no-others --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_l1.log 2>&1
DLSM_FFBS_SMEM=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ffbs -c 12 --csv --log-file gpurun_out/${TAG}_ffbs_smem.csv \
   python bench.py --workload cfg2 --no-others --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_l2.log 2>&1
