#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c9}
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
AB="--no-others --no-e2e --no-cpu-baseline --steps 30 --warmup 3"
python bench.py --workload cfg4 --chains-per-gpu 128 $AB > gpurun_out/${TAG}_ab_cfg4c128_auto.json 2>> gpurun_out/${TAG}_ab.err
python bench.py --workload cfg2 --chains-per-gpu 1 $AB > gpurun_out/${TAG}_ab_cfg2c1_auto.json 2>> gpurun_out/${TAG}_ab.err
python bench.py --workload cfg2 --chains-per-gpu 148 $AB > gpurun_out/${TAG}_ab_cfg2c148_auto.json 2>> gpurun_out/${TAG}_ab.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
tail -n 2 gpurun_out/${TAG}_smoke.log
du -sh gpurun_out
