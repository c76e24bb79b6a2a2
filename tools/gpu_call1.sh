#!/bin/bash
# round-2 GPU call: GPU test suite, the default bench line, launch lists and ncu captures
# (ncu reports are reduced to their raw / source CSV pages on the box: gpurun_out/ must stay < 64 MiB)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.json
for wl in cfg4 cfg3 cfg5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
     --log-file gpurun_out/${TAG}_launches_$wl.csv python bench.py --workload $wl --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_launch_$wl.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -o /tmp/prof_$wl \
     python bench.py --workload $wl --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_$wl.log 2>&1
  ncu -i /tmp/prof_$wl.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${wl}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$wl.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_${wl}_source.csv 2>/dev/null
  ncu -i /tmp/prof_$wl.ncu-rep --page details > gpurun_out/${TAG}_prof_${wl}_details.txt 2>/dev/null
done
if [ "$2" = "rows" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_rows$ -s 2 -c 1 -o /tmp/prof_rows \
     python bench.py --workload cfg4 --no-others --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_prof_rows.log 2>&1
  ncu -i /tmp/prof_rows.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_rows_raw.csv 2>/dev/null
  ncu -i /tmp/prof_rows.ncu-rep --page details > gpurun_out/${TAG}_prof_rows_details.txt 2>/dev/null
fi
du -sh gpurun_out
