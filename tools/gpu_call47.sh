#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_lpcm.py -m gpu -q -s ) > gpurun_out/c47_pytest.log 2>&1
grep -v Warning gpurun_out/c47_pytest.log | tail -30
timeout 100 python tools/fit_timing.py lpcm 1000 2>&1 | tail -1 | tee gpurun_out/c47_fit_lpcm.txt
