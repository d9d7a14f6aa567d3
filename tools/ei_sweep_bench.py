#!/usr/bin/env python
"""Achieved HBM bandwidth of the EI sweep kernel (north_star: 'EI scan over candidates as a coalesced HBM sweep').
Algorithmic bytes per (candidate, sample[, fantasy]) evaluation: read mu (4 B) [+ var 4 B when F == 1] + write ei (8 B / F: EI is stored in double).
The headline grid (M=100k, S=40, F=1) is 48 MB -- launch-latency sized -- so the asymptotic figure is measured on
larger sweeps (more candidates, and the pending-fantasy shape F=100)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spearmint_b200.engine import GPEIEngine  # noqa: E402


def run(M, S, F, iters=20):
    eng = GPEIEngine(dtype=torch.float32)
    ldm = ((M + 127) // 128) * 128
    g = torch.Generator(device="cuda").manual_seed(0)
    mu = torch.randn((S, F, ldm), device="cuda", generator=g)
    var = torch.rand((S, ldm), device="cuda", generator=g) + 0.01
    best = torch.full((S, F), -0.5, device="cuda")
    for _ in range(3):
        eng.ei_sweep(M, S, F, mu, var, ldm, best)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                      # > L2: every timed sweep reads from HBM
        e0.record()
        eng.ei_sweep(M, S, F, mu, var, ldm, best)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    nbytes = 4.0 * M * S * F + 4.0 * M * S + 8.0 * M * S + 16.0 * M     # mu + var + ei (double) + ei_sum (read + write)
    return dict(M=M, S=S, F=F, ms=ms, GBps=nbytes / ms / 1e6, bytes=nbytes)


if __name__ == "__main__":
    peak = 6570.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    for M, S, F in ((100000, 40, 1), (1000000, 40, 1), (4000000, 40, 1), (100000, 40, 100), (1000000, 8, 100)):
        r = run(M, S, F)
        r["frac_of_measured_hbm"] = r["GBps"] / peak
        print(json.dumps(r))
