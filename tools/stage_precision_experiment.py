#!/usr/bin/env python
"""Which stage of the float32 grid path costs the accuracy at the ill-conditioned BASELINE configs (C2: D=8 N=512, C4: D=8 N=1024)?
CPU emulation: each variant rounds ONE group of intermediates to float32 and keeps the rest in float64."""
import sys
import numpy as np
import scipy.linalg as spla
sys.path.insert(0, '.')
import bench
from oracle import gp_oracle as O

f32 = lambda a: np.asarray(a, dtype=np.float32)
f64 = lambda a: np.asarray(a, dtype=np.float64)


def run(workload, s_idx, M_sub=3000):
    D, N, M, S = bench.WORKLOADS[workload]
    comp, cand, vals, hs = bench.synth(D, N, M, S)
    cand = cand[:M_sub]
    mean, noise, amp2, ls = hs[s_idx]
    best = vals.min()
    K = O.cov(bench.KIND, amp2, ls, comp) + noise * np.eye(N)
    Kx = O.cov(bench.KIND, amp2, ls, comp, cand)
    L = spla.cholesky(K, lower=True)
    print(workload, "sample", s_idx, "cond(K) %.2e" % np.linalg.cond(K), "|K|2 %.1f" % np.linalg.norm(K, 2))

    def ei_of(L_, Linv_, Kx_, alpha_):
        beta = Linv_ @ Kx_ if Linv_ is not None else spla.solve_triangular(L_, Kx_, lower=True)
        m = Kx_.T @ alpha_ + mean
        v = amp2 * (1 + 1e-6) - np.sum(f64(beta) ** 2, axis=0)
        return O._ei_from_moments(best, m, np.sqrt(np.maximum(v, 1e-300))), m, v

    alpha = spla.cho_solve((L, True), vals - mean)
    ref, mref, vref = ei_of(L, None, Kx, alpha)
    Linv = spla.solve_triangular(L, np.eye(N), lower=True)

    def rep(name, ei, m, v):
        print("   %-46s max|dEI|/maxEI %.2e   max|dm| %.2e  max|dv| %.2e  argmax_ok %d" %
              (name, np.abs(ei - ref).max() / ref.max(), np.abs(m - mref).max(), np.abs(v - vref).max(),
               int(np.argmax(ei) == np.argmax(ref))))

    # 1. fp32 K + fp32 Cholesky, everything after in fp64 with that factor
    L32 = f64(spla.cholesky(f32(K), lower=True))
    a1 = spla.cho_solve((L32, True), vals - mean)
    rep("fp32 K+chol, rest fp64 (alpha from L32)", *ei_of(L32, None, Kx, a1))
    rep("fp32 K+chol, rest fp64 (alpha exact)", *ei_of(L32, None, Kx, alpha))
    # 2. fp64 factor and inverse, Linv rounded to fp32, Kx rounded to fp32
    rep("fp64 chol+inv, operands rounded to fp32", *ei_of(L, f64(f32(Linv)), f64(f32(Kx)), alpha))
    # 3. fp64 factor, inverse computed in fp32 from L rounded to fp32
    Li32 = f64(spla.solve_triangular(f32(L), np.eye(N, dtype=np.float32), lower=True))
    rep("fp64 chol -> fp32 L -> fp32 inverse", *ei_of(L, Li32, f64(f32(Kx)), alpha))
    # 4. all fp32: K, chol, inverse, alpha via inverse
    Li32b = f64(spla.solve_triangular(f32(L32), np.eye(N, dtype=np.float32), lower=True))
    a4 = f64(f32(Li32b.T) @ (f32(Li32b) @ f32(vals - mean)))
    rep("all fp32 (K, chol, inverse, alpha)", *ei_of(L32, Li32b, f64(f32(Kx)), a4))
    rep("all fp32 but alpha exact", *ei_of(L32, Li32b, f64(f32(Kx)), alpha))
    # 5. fp32 chol + one refinement of the factor?  (K - L32 L32^T residual correction is O(N^3) fp64: not cheaper)
    # 6. K built in fp64, rounded to fp32, fp32 chol
    L32c = f64(spla.cholesky(f32(K), lower=True))
    # 7. fp64 chol, L rounded to fp32, fp64 rest
    rep("fp64 chol, L rounded to fp32, rest fp64", *ei_of(f64(f32(L)), None, Kx, alpha))


run("c2", 4)
run("c4", 1)
run("headline", 0, 1500)
