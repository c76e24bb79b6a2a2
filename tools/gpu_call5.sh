#!/bin/bash
# GPU call: test suite, default bench, A/B of the sweep-kernel variants on cfg4, cfg3 block-speculative kernel profile
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c5}
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
AB="--workload cfg4 --no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
DLSM_NO_ROWSUM=1 python bench.py $AB > gpurun_out/${TAG}_ab_tworow_v2.json 2>> gpurun_out/${TAG}_ab.err
if [ -f variants/libdlsm_v1tab.so ]; then
  DLSM_LIB=$PWD/variants/libdlsm_v1tab.so python bench.py $AB > gpurun_out/${TAG}_ab_rows_v1.json 2>> gpurun_out/${TAG}_ab.err
  DLSM_LIB=$PWD/variants/libdlsm_v1tab.so DLSM_NO_ROWSUM=1 python bench.py $AB > gpurun_out/${TAG}_ab_tworow_v1.json 2>> gpurun_out/${TAG}_ab.err
  DLSM_LIB=$PWD/variants/libdlsm_v1tab.so python bench.py --workload cfg3 --no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/${TAG}_ab_cfg3_v1.json 2>> gpurun_out/${TAG}_ab.err
fi
DLSM_NO_CLUSTER=1 python bench.py --workload cfg3 --no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/${TAG}_ab_cfg3_nocluster.json 2>> gpurun_out/${TAG}_ab.err
DLSM_NO_CLUSTER=2 python bench.py --workload cfg3 --no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/${TAG}_ab_cfg3_pernode.json 2>> gpurun_out/${TAG}_ab.err
for wl in cfg3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
     --log-file gpurun_out/${TAG}_launches_$wl.csv python bench.py --workload $wl --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_launch_$wl.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 17 -c 1 -o /tmp/prof_$wl \
     python bench.py --workload $wl --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_$wl.log 2>&1
  ncu -i /tmp/prof_$wl.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${wl}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$wl.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_${wl}_source.csv 2>/dev/null
  ncu -i /tmp/prof_$wl.ncu-rep --page details > gpurun_out/${TAG}_prof_${wl}_details.txt 2>/dev/null
done
du -sh gpurun_out
