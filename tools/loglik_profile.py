#!/usr/bin/env python
"""Times the float64 GP log-likelihood (slice-sampler logprob, f2) per call and per batch: python tools/loglik_profile.py N D"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from spearmint_b200.backend import DeviceBackend

N, D = int(sys.argv[1]), int(sys.argv[2])
b = DeviceBackend()
rs = np.random.RandomState(0)
X = rs.rand(N, D)
y = rs.randn(N)
ll = b.loglik("Matern52", X, y)
out = {"N": N, "D": D}
for B in (1, 2, 4, 6, 8):
    hs = [(0.0, 1e-3, 1.0 + 0.01 * i, np.ones(D)) for i in range(B)]
    for _ in range(2):
        ll.batch(hs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ll.batch(hs)
    torch.cuda.synchronize()
    out["batch%d_ms" % B] = 1e3 * (time.perf_counter() - t0) / reps
print(json.dumps(out))
