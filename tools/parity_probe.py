#!/usr/bin/env python
"""max|dEI|/max EI per hyper-sample column of the float32 GPU path vs the float64 oracle, for one bench workload subset.
usage: python tools/parity_probe.py WORKLOAD S_SUB M_SUB   (env SMK_FACTOR_IMPL / SMK_PREDICT_IMPL select the kernels)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, '.')
import bench
from oracle import gp_oracle as O

wl, S_sub, M_sub = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
D, N, M, S = bench.WORKLOADS[wl]
comp, cand, vals, hs = bench.synth(D, N, M, S)
hs = hs[:S_sub]
rs = np.random.RandomState(7)
pick = np.sort(rs.permutation(M)[:M_sub])
cand = np.vstack([rs.randn(10, D) * 0.001 + comp[np.argmin(vals)], cand[pick]])
pend = np.zeros((0, D))
ref = O.ei_over_hypers(bench.KIND, hs, comp, pend, cand, vals)
import torch
from spearmint_b200.engine import GPEIEngine
out = {"workload": wl, "factor": os.environ.get("SMK_FACTOR_IMPL", "tc"), "predict": os.environ.get("SMK_PREDICT_IMPL", "tc")}
for prec in ("f32", "f64"):
    eng = GPEIEngine(dtype=torch.float32 if prec == "f32" else torch.float64)
    ei = eng.ei_over_hypers(bench.KIND, hs, comp, pend, cand, vals)
    out[prec] = [float(np.abs(ei[:, s] - ref[:, s]).max() / ref[:, s].max()) for s in range(len(hs))]
    out[prec + "_argmax_ok"] = bool(np.argmax(ei.mean(1)) == np.argmax(ref.mean(1)))
    if prec == "f32":
        out["guard"] = eng.last_guard
        out["guard_threshold"] = eng.guard_threshold
print(json.dumps(out))
