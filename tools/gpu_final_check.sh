#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c43_smoke.log 2>&1
tail -n 1 gpurun_out/c43_smoke.log
( timeout 600 python -m pytest tests/test_gpu_rowsum_cache.py tests/test_gpu_edge_cases.py -m gpu -q -x ) > gpurun_out/c43_pytest.log 2>&1
tail -2 gpurun_out/c43_pytest.log
timeout 200 python bench.py --workload cfg3 --no-others --no-e2e --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/c43_cfg3.json 2>/dev/null
