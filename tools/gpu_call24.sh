#!/bin/bash
# the driver's default invocations: reference arm, then the default bench (N = 1)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c24}
( time python bench.py --impl reference ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -3 gpurun_out/${TAG}_bench_reference.err
( time python bench.py ) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -3 gpurun_out/${TAG}_bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
tail -n 2 gpurun_out/${TAG}_smoke.log
