#!/bin/bash
# windowed cluster kernel as the default: full GPU suite, memcheck probe, cfg 3 bench + ncu
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c22}
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/${TAG}_sanitize_memcheck.log 2>&1
tail -n 3 gpurun_out/${TAG}_sanitize_memcheck.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 30 --warmup 3"
timeout 300 python bench.py --workload cfg3 $AB > gpurun_out/${TAG}_ab_cfg3.json 2>> gpurun_out/${TAG}_ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_blkw -s 6 -c 1 -o /tmp/prof_cfg3 \
   python bench.py --workload cfg3 --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_cfg3.log 2>&1
ncu -i /tmp/prof_cfg3.ncu-rep --page details > gpurun_out/${TAG}_prof_cfg3_details.txt 2>/dev/null
ncu -i /tmp/prof_cfg3.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_cfg3_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv \
   python bench.py --workload cfg3 --no-others --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_launch.log 2>&1
