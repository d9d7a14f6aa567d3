#!/usr/bin/env python
"""cProfile of the plugin call (bench.py's e2e leg) on rank 0 under torchrun: where the host time of a multi-GPU call goes.
usage: python -m torch.distributed.run --nproc-per-node N tools/e2e_profile.py"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
import bench
from spearmint_b200.backend import DeviceBackend
from spearmint_b200.chooser import GPEIOptChooserB200 as plugin

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
D, N, M, S = bench.WORKLOADS["headline"]
comp, cand, vals, hs = bench.synth(D, N, M, S)
pend = np.zeros((0, D))
backend = DeviceBackend(device="cuda:%d" % local)
ch = plugin.init(tempfile.mkdtemp(), "mcmc_iters=%d,burnin=0,noiseless=1" % S)
ch._backend = backend
ch.D, ch.hyper_samples = D, list(hs)
for _ in range(3):
    ch.ei_over_hypers(comp, pend, cand, vals)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(3):
    ch.ei_over_hypers(comp, pend, cand, vals)
pr.disable()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
if rank == 0:
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print("world %d: %.1f ms per call" % (world, dt * 1e3))
    print(s.getvalue()[:6000])
