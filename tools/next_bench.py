#!/usr/bin/env python
"""chooser.next() wall-clock (the second half of BASELINE.json's metric): the drop-in chooser on the GPU vs the same
host logic with the CPU oracle numerics (= a faithful port of the reference's next(): identical RNG order, identical
chain -- tests/test_chooser_host.py).  Usage: python tools/next_bench.py --workload c2 [--cpu] [--calls 2]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def problem(D, N, M):
    grid = np.random.RandomState(0).rand(N + M, D)
    perm = np.random.RandomState(0).permutation(N + M)
    complete, candidates = np.sort(perm[:N]), np.sort(perm[N:])
    values = np.full(N + M, np.nan)
    y = np.sin(3 * grid[complete]).sum(1) + 0.01 * np.random.RandomState(1).randn(N)
    values[complete] = (y - y.mean()) / y.std()
    return grid, values, candidates, complete


def run(backend_name, D, N, M, S, burnin, calls, grid_subset, backend=None):
    from spearmint_b200.chooser import GPEIOptChooserB200 as mod
    import spearmint_b200.locker as lk
    lk.log = lambda *a: None
    mod.log = lambda *a: None
    grid, values, candidates, complete = problem(D, N, M)
    ch = mod.init(tempfile.mkdtemp(), "mcmc_iters=%d,burnin=%d,noiseless=1,grid_subset=%d" % (S, burnin, grid_subset))
    if backend_name == "cpu":
        from tests.oracle_backend import OracleBackend
        ch._backend = OracleBackend()
    elif backend is not None:
        ch._backend = backend
    np.random.seed(0)
    out = []
    for c in range(calls):
        if backend_name != "cpu":
            import torch
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = ch.next(grid, values, None, candidates, np.array([], dtype=int), complete)
        if backend_name != "cpu":
            import torch
            torch.cuda.synchronize()
        out.append(dict(call=c, ms=1e3 * (time.perf_counter() - t0), ret=int(r[0] if isinstance(r, tuple) else r),
                        refine_evals=ch.stats.get("refine_evals"), loglik_evals=ch.stats.get("loglik_evals"),
                        loglik_batches=ch.stats.get("loglik_batches"), phase_ms=ch.stats.get("phase_ms")))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--calls", type=int, default=2)
    ap.add_argument("--burnin", type=int, default=10)
    ap.add_argument("--S", type=int, default=None)
    ap.add_argument("--grid-subset", type=int, default=20)
    a = ap.parse_args()
    D, N, M, S = bench.WORKLOADS[a.workload]
    S = a.S or S
    res = run("cpu" if a.cpu else "gpu", D, N, M, S, a.burnin, a.calls, a.grid_subset)
    print(json.dumps(dict(workload=a.workload, D=D, N=N, M=M, S=S, burnin=a.burnin, impl="cpu-port" if a.cpu else "b200",
                          cores=os.cpu_count(), calls=res)))
