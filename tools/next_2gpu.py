#!/usr/bin/env python
"""torchrun --nproc-per-node 2 tools/next_2gpu.py : the drop-in chooser's next() with hyper-samples sharded over ranks
(one NCCL all-reduce of the EI sum per grid pass) must propose what the reference proposed on the golden runs."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import hypers, load  # noqa: E402

if __name__ == "__main__":
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    from spearmint_b200.chooser import GPEIOptChooserB200 as mod
    import spearmint_b200.locker as lk
    mod.log = lk.log = lambda *a: None
    ok = True
    for name in ("opt_d8_m52", "opt_d8_m52_pend", "opt_d5_ardse", "opt_branin2d"):
        g = load(name)
        args = "covar=%s,mcmc_iters=%d,burnin=%d,noiseless=%d,grid_subset=5" % (
            str(g["kind"]), int(g["S"]), int(g["burnin"]), int(g["noiseless"]))
        ch = mod.init(tempfile.mkdtemp(), args)
        np.random.seed(int(g["seed"]))
        ret = ch.next(g["grid"], g["values"], g["durations"], g["candidates"], g["pending"], g["complete"])
        ref_pt = g["next_point"] if int(g["next_is_tuple"]) else g["grid"][int(g["next_index"])]
        got_pt = ret[1] if isinstance(ret, tuple) else g["grid"][ret]
        err = float(np.abs(got_pt - ref_pt).max())
        chain = max(float(np.abs(np.hstack(a) - np.hstack(b)).max()) for a, b in zip(ch.hyper_samples, hypers(g)))
        good = err < 2e-4 and chain < 1e-6
        ok = ok and good
        print("rank %d %s: proposal err %.2e chain err %.2e %s" % (rank, name, err, chain, "OK" if good else "MISMATCH"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
