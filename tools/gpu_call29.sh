#!/bin/bash
# final validation of the round: full GPU suite, smoke, build check of the loaded library
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-c29}
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
tail -n 2 gpurun_out/${TAG}_smoke.log
