#!/usr/bin/env python
"""Stage split of one float64 log-likelihood batch (LogLik.batch): host prep, covariance, factorisation, finish + read-back.
usage: python tools/loglik_stages.py N D [B]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from spearmint_b200 import _lib
from spearmint_b200.backend import DeviceBackend
from spearmint_b200.engine import KINDS, check, fn, ptr

N, D = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
b = DeviceBackend()
rs = np.random.RandomState(0)
X = rs.rand(N, D)
y = rs.randn(N)
ll = b.loglik("Matern52", X, y)
eng, dt, Npad = ll.eng, ll.eng.dtype, ll.Npad
hs = [(0.0, 1e-3, 1.0 + 0.01 * i, np.ones(D)) for i in range(B)]
for _ in range(3):
    ll.batch(hs)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
acc = np.zeros(6)
reps = 10
for _ in range(reps):
    t0 = time.perf_counter()
    hb = eng.hypers([(h[0], h[1], h[2], np.asarray(h[3], dtype=float)) for h in hs], ll.kind)
    st = eng.stream()
    t1 = time.perf_counter()
    ev[0].record()
    check(fn("smk_cov_build_lower", dt)(KINDS[ll.kind], N, D, B, ptr(ll.X), ptr(hb.inv_ls), ptr(hb.amp2), ptr(hb.noise),
                                        ptr(ll.L), Npad, st), "cov_build")
    check(fn("smk_loglik_set_rhs", dt)(N, Npad, B, ptr(ll.y), ptr(hb.mean), ptr(ll.L), st), "set_rhs")
    ev[1].record()
    check(_lib.lib().smk_potrf_loglik_f64(Npad, B, ptr(ll.L), ptr(ll.winv), ll.ws_bytes, ptr(ll.info), ll.use_graph, st), "potrf")
    ev[2].record()
    check(fn("smk_loglik_finish", dt)(N, Npad, B, ptr(ll.L), ptr(ll.out[0]), ptr(ll.out[1]), st), "finish")
    ev[3].record()
    r = torch.cat([ll.out[0, :B].double(), ll.out[1, :B].double(), ll.info[:B].double()]).cpu().numpy()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    acc += [1e3 * (t1 - t0), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), 1e3 * (t2 - t1),
            1e3 * (t2 - t0)]
acc /= reps
print(json.dumps({"N": N, "D": D, "B": B, "host_prep_ms": acc[0], "cov_ms": acc[1], "potrf_ms": acc[2], "finish_ms": acc[3],
                  "enqueue_to_result_ms": acc[4], "total_ms": acc[5]}))
