#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( timeout 420 python -m pytest tests/test_gpu_lpcm.py "tests/test_gpu_estimators.py::test_hdp_replay_reproduces_the_reference_fit" -m gpu -q -x -s ) > gpurun_out/c44_pytest.log 2>&1
tail -25 gpurun_out/c44_pytest.log
