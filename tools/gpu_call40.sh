#!/bin/bash
# persisting-L2 window over the label kernel's global stage: DRAM bytes and time, with / without
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c40}
AB="--no-others --no-e2e --no-cpu-baseline --steps 40 --warmup 3"
for w in cfg2 cfg4; do
timeout 300 python bench.py --workload $w $AB > gpurun_out/${TAG}_ab_${w}_window.json 2>> gpurun_out/${TAG}_ab.err
DLSM_FFBS_NO_L2_WINDOW=1 timeout 300 python bench.py --workload $w $AB > gpurun_out/${TAG}_ab_${w}_nowindow.json 2>> gpurun_out/${TAG}_ab.err
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ffbs -c 12 --csv --log-file gpurun_out/${TAG}_ffbs_window.csv \
   python bench.py --workload cfg2 --no-others --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_l1.log 2>&1
DLSM_FFBS_NO_L2_WINDOW=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ffbs -c 12 --csv --log-file gpurun_out/${TAG}_ffbs_nowindow.csv \
   python bench.py --workload cfg2 --no-others --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_l2.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_estimators.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
