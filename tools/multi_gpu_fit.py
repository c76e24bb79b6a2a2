"""Two (or more) ranks, one GPU each: every rank fits its own HDP-LPCM chains on the same network,
the scalar traces are pooled over NCCL and summarised with split R-hat / ESS.
usage: torchrun --nproc-per-node N tools/multi_gpu_fit.py"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
import bench  # noqa: E402
from dynetlsm_b200 import DynamicNetworkHDPLPCM  # noqa: E402
from dynetlsm_b200.diagnostics import pool_traces, summarize  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
w = bench.make_workload("cfg2")
m = DynamicNetworkHDPLPCM(n_components=10, n_iter=300, tune=150, burn=150, random_state=100 + rank,
                          n_chains=4, device=local).fit(w["Y"])
tr = np.stack([m.chains_["logps"], m.chains_["intercepts"][:, :, 0], m.chains_["lambdas"]], axis=2)
pooled = pool_traces(tr)
if rank == 0:
    s = summarize(pooled, ["logp", "intercept", "lambda"], n_burn=300)
    print("pooled traces", pooled.shape, {k: {a: round(b, 3) for a, b in v.items()} for k, v in s.items()})
dist.destroy_process_group()
