#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c45_smoke.log 2>&1
tail -n 1 gpurun_out/c45_smoke.log
( timeout 400 python -m pytest tests -m gpu -q -x ) > gpurun_out/c45_pytest.log 2>&1
tail -3 gpurun_out/c45_pytest.log
