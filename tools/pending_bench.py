#!/usr/bin/env python
"""Grid pass WITH pending fantasies (OPT:558-619) at a given workload: P pending points, F fantasies per sample."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spearmint_b200.engine import GPEIEngine  # noqa: E402

if __name__ == "__main__":
    w = sys.argv[1] if len(sys.argv) > 1 else "c3"
    P, F = 3, 100
    D, N, M, S = bench.WORKLOADS[w]
    comp, cand, vals, hs = bench.synth(D, N, M, S)
    pend = np.random.RandomState(5).rand(P, D)
    normals = np.random.RandomState(6).randn(P, F)
    eng = GPEIEngine(dtype=torch.float32)
    for impl in ("tc", "simt"):
        eng.predict_impl = impl
        eng.ei_over_hypers_device("Matern52", hs, comp, pend, cand, vals, normals, want_matrix=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, ei_sum, _ = eng.ei_over_hypers_device("Matern52", hs, comp, pend, cand, vals, normals, want_matrix=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps(dict(workload=w, P=P, F=F, impl=impl, ms=1e3 * dt, cand_per_s=M / dt,
                              argmax=int(torch.argmax(ei_sum[:M])))))
