#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c7}
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
AB="--no-others --no-e2e --no-cpu-baseline --steps 30 --warmup 3"
for wl in cfg2 cfg1; do
  python bench.py --workload $wl --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_${wl}c1_auto.json 2>> gpurun_out/${TAG}_ab.err
  DLSM_CHAIN_KERNEL=node python bench.py --workload $wl --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_${wl}c1_node.json 2>> gpurun_out/${TAG}_ab.err
done
python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_auto.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CHAIN_KERNEL=node python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_node.json 2>> gpurun_out/${TAG}_ab.err
python bench.py --workload cfg4 --chains-per-gpu 296 $AB > gpurun_out/${TAG}_ab_cfg4c296_auto.json 2>> gpurun_out/${TAG}_ab.err
DLSM_CHAIN_KERNEL=rowsum python bench.py --workload cfg4 --chains-per-gpu 296 $AB > gpurun_out/${TAG}_ab_cfg4c296_rowsum.json 2>> gpurun_out/${TAG}_ab.err
python bench.py --workload cfg4 --chains-per-gpu 512 $AB > gpurun_out/${TAG}_ab_cfg4c512_auto.json 2>> gpurun_out/${TAG}_ab.err
python bench.py --workload cfg4 --chains-per-gpu 256 $AB > gpurun_out/${TAG}_ab_cfg4c256_auto.json 2>> gpurun_out/${TAG}_ab.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --workload cfg3 --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_launch_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 17 -c 1 -o /tmp/prof_cfg3 \
   python bench.py --workload cfg3 --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_cfg3.log 2>&1
ncu -i /tmp/prof_cfg3.ncu-rep --page details > gpurun_out/${TAG}_prof_cfg3_details.txt 2>/dev/null
ncu -i /tmp/prof_cfg3.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_cfg3_raw.csv 2>/dev/null
bash tools/sanitize.sh
du -sh gpurun_out
