#!/bin/bash
# multi-GPU bench: N ranks over NCCL (chains sharded, traces pooled once with an all-gather)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-m$N}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 40 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 400 gpurun_out/${TAG}_bench.json
tail -n 5 gpurun_out/${TAG}_bench.err
