#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c6}
( time timeout 1500 python -m pytest tests/test_gpu_rowsum_cache.py tests/test_gpu_parity.py -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
for wl in cfg4 cfg2; do
  DLSM_CHAIN_KERNEL=block python bench.py --workload $wl $AB > gpurun_out/${TAG}_ab_${wl}_block.json 2>> gpurun_out/${TAG}_ab.err
  DLSM_CHAIN_KERNEL=node python bench.py --workload $wl $AB > gpurun_out/${TAG}_ab_${wl}_node.json 2>> gpurun_out/${TAG}_ab.err
done
DLSM_CHAIN_KERNEL=block python bench.py --workload cfg1 $AB > gpurun_out/${TAG}_ab_cfg1_block.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CHAIN_KERNEL=node python bench.py --workload cfg1 $AB > gpurun_out/${TAG}_ab_cfg1_node.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CHAIN_KERNEL=block python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_block.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CHAIN_KERNEL=block DLSM_SWEEP_MODE=chain python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_blockchain.json 2>> gpurun_out/${TAG}_ab.err
python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_default.json 2>> gpurun_out/${TAG}_ab.err
export DLSM_CHAIN_KERNEL=block
for wl in cfg4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_cb -s 5 -c 1 -o /tmp/prof_$wl \
     python bench.py --workload $wl --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_$wl.log 2>&1
  ncu -i /tmp/prof_$wl.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${wl}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$wl.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_${wl}_source.csv 2>/dev/null
  ncu -i /tmp/prof_$wl.ncu-rep --page details > gpurun_out/${TAG}_prof_${wl}_details.txt 2>/dev/null
done
du -sh gpurun_out
