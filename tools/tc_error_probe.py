#!/usr/bin/env python
"""Where does the tensor-core predict path lose accuracy on an ill-conditioned problem (C2: D=8 N=512, worst hyper-sample)?
Each variant swaps ONE GPU-produced operand into an otherwise float64 computation of var = amp2(1+1e-6) - |Linv Kx|^2."""
import json
import sys

import numpy as np
import scipy.linalg as spla
import torch

sys.path.insert(0, '.')
import bench
from oracle import gp_oracle as O
from spearmint_b200 import _lib
from spearmint_b200.engine import GPEIEngine, ptr, check, KINDS

wl, s_idx, M_sub = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
D, N, M, S = bench.WORKLOADS[wl]
comp, cand, vals, hs = bench.synth(D, N, M, S)
cand = np.vstack([np.random.RandomState(7).randn(10, D) * 0.001 + comp[np.argmin(vals)], cand[:M_sub - 10]])   # jitter cloud first
h = hs[s_idx]
mean, noise, amp2, ls = h
kind = bench.KIND
eng = GPEIEngine(dtype=torch.float32)
L_ = _lib.lib()
hb = eng.hypers([h], kind)
Xd, Cd, yd = eng.to_dev(comp), eng.to_dev(cand), eng.to_dev(vals)
fac = eng.factor(kind, Xd, hb)
fac.check_pd()
alpha = fac.alpha_via_linv(yd)
hi, lo, Np = fac.linv()
h16, l16, exps, _ = fac.linv16()
Mc = ((M_sub + 127) // 128) * 128
dbg = torch.zeros((1, Mc, Np), dtype=torch.float32, device=eng.device)
mu, var, ldm = eng.predict(kind, fac, Cd, alpha.view(1, fac.Npad), impl="tc", dbg_beta=dbg)
# generator output
kh = torch.zeros((1, Mc, Np), dtype=torch.float16, device=eng.device)
kl = torch.zeros((1, Mc, Np), dtype=torch.float16, device=eng.device)
mu2 = torch.zeros((1, ldm), dtype=torch.float32, device=eng.device)
nb = L_.smk_kxt_pack_workspace_bytes(Np, M_sub, 1)
ws = torch.empty((nb,), dtype=torch.uint8, device=eng.device)
a_pad = torch.zeros((1, Np), dtype=torch.float32, device=eng.device)
check(L_.smk_kxt_pack_f16(0, KINDS[kind], N, Np, M_sub, D, 1, ptr(Xd), ptr(Cd), ptr(hb.inv_ls), ptr(hb.amp2), ptr(hb.mean),
                          ptr(a_pad), Np, ptr(kh), ptr(kl), ptr(mu2), ldm, ptr(ws), nb, eng.stream()), "kxt")
torch.cuda.synchronize()
ea = 15 - np.frexp(np.float32(np.float32(amp2) * np.float32(1.000001)) * np.float32(1.00001))[1]
Kx_gpu = ((kh.double() + kl.double()).cpu().numpy()[0, :M_sub, :N] * 2.0 ** -ea).T            # (N, M)
Linv32 = (hi.double() + lo.double()).cpu().numpy()[0, :N, :N]
eb = int(exps.cpu().numpy()[0])
Linv16 = (h16.double() + l16.double()).cpu().numpy()[0, :N, :N] * 2.0 ** -eb
beta_dbg = dbg.double().cpu().numpy()[0, :M_sub, :N].T
var_gpu = var.double().cpu().numpy()[0, :M_sub]
mu_gpu = mu.double().cpu().numpy()[0, :M_sub]

K = O.cov(kind, amp2, ls, comp) + noise * np.eye(N)
Kx = O.cov(kind, amp2, ls, comp, cand)
L = spla.cholesky(K, lower=True)
Linv = spla.solve_triangular(L, np.eye(N), lower=True)
a_ref = spla.cho_solve((L, True), vals - mean)
m_ref = Kx.T @ a_ref + mean
best = vals.min()
c = amp2 * (1 + 1e-6)
v_ref = c - np.sum((Linv @ Kx) ** 2, axis=0)
ei_ref = O._ei_from_moments(best, m_ref, np.sqrt(v_ref))


def rep(name, v, m=None):
    m = m_ref if m is None else m
    ei = O._ei_from_moments(best, m, np.sqrt(np.maximum(v, 1e-300)))
    print("%-58s max|dv| %.3e  mean dv %+.3e  max|dEI|/maxEI %.3e" % (name, np.abs(v - v_ref).max(), (v - v_ref).mean(),
                                                                     np.abs(ei - ei_ref).max() / ei_ref.max()))


print(wl, "sample", s_idx, "N", N, "cond %.2e" % np.linalg.cond(K), "min var %.3e" % v_ref.min(), "max|Linv| %.1f" % np.abs(Linv).max())
rep("exact operands (sanity)", v_ref)
rep("GPU Kx (generator, fp16 pair), exact Linv", c - np.sum((Linv @ Kx_gpu) ** 2, axis=0))
rep("exact Kx, GPU Linv (fp32 factor + inverse)", c - np.sum((Linv32 @ Kx) ** 2, axis=0))
rep("exact Kx, GPU Linv fp16 pair", c - np.sum((Linv16 @ Kx) ** 2, axis=0))
rep("exact Kx rounded to fp32, exact Linv rounded to fp32", c - np.sum((Linv.astype(np.float32).astype(float) @ Kx.astype(np.float32).astype(float)) ** 2, axis=0))
b_ab = Linv16 @ Kx_gpu
rep("GPU operands (both fp16 pairs), float64 GEMM", c - np.sum(b_ab ** 2, axis=0))
rep("GPU beta dump (tensor-core GEMM), float64 square-sum", c - np.sum(beta_dbg ** 2, axis=0))
rep("GPU var", var_gpu)
rep("GPU var + GPU mu", var_gpu, mu_gpu)
rep("exact var + GPU mu", v_ref, mu_gpu)
print("cloud only (first 10 candidates):")
for nm, vv, mm in (("GPU var, exact mu", var_gpu, m_ref), ("exact var, GPU mu", v_ref, mu_gpu), ("GPU var + mu", var_gpu, mu_gpu)):
    ei = O._ei_from_moments(best, mm, np.sqrt(np.maximum(vv, 1e-300)))
    print("   %-22s cloud max|dEI|/maxEI %.3e   rest %.3e   (EI cloud max %.3e, EI max %.3e)" % (
        nm, np.abs(ei - ei_ref)[:10].max() / ei_ref.max(), np.abs(ei - ei_ref)[10:].max() / ei_ref.max(), ei_ref[:10].max(), ei_ref.max()))
d = beta_dbg - b_ab
print("GEMM error: max|d| %.3e  mean d %+.3e  mean |d| %.3e  ; mean(d*sign(b)) %+.3e (negative = truncation toward zero) ; max|beta| %.3e"
      % (np.abs(d).max(), d.mean(), np.abs(d).mean(), (d * np.sign(b_ab)).mean(), np.abs(b_ab).max()))
# partial-sum magnitudes: how much bigger than the result are the running sums?
big = np.abs(Linv16)[:, :, None] * np.abs(Kx_gpu)[None, :, :64]
print("sum|terms| / |beta| median %.1f" % np.median(big.sum(1) / np.maximum(np.abs(b_ab[:, :64]), 1e-30)))
