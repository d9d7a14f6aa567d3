"""CPU emulation of the tensor-core cross-covariance generator: r2[s, c, n] = sum_d w[s, d] * q[c, n, d],
q = (x - c)^2 (float32, difference form -> no cancellation), w = 1 / ls^2, both operands as scaled fp16 (hi, lo) pairs,
3 products, fp32 accumulation (emulated by float32 matmuls).  Reports the relative error of r2 and of the Matern52
value against float64, next to the plain float32 FMA chain the SIMT generator uses."""
import numpy as np

def split16(x32):
    hi = x32.astype(np.float16)
    lo = (x32 - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)

def sexp(m):
    return 15 - int(np.frexp(np.float32(m) * np.float32(1.00001))[1])

def m52(r2):
    r = np.sqrt(r2)
    return (1 + np.sqrt(5.) * r + 5. / 3. * r2) * np.exp(-np.sqrt(5.) * r)

rs = np.random.RandomState(0)
for D, lsr in [(32, (0.3, 2.0)), (8, (0.3, 2.0)), (3, (0.05, 1.0)), (20, (0.5, 5.0))]:
    N, M = 512, 256
    X = rs.rand(N, D); C = rs.rand(M, D); C[:10] = X[0] + 1e-3 * rs.randn(10, D)
    ls = rs.uniform(*lsr, size=D)
    w64 = 1.0 / ls ** 2
    q64 = (X[None, :, :] - C[:, None, :]) ** 2                        # (M, N, D)
    r2_64 = q64 @ w64
    # float32 SIMT chain: sum_d ((x - c) * ils)^2
    ils32 = (1.0 / ls).astype(np.float32)
    df = (X.astype(np.float32)[None] * ils32 - C.astype(np.float32)[:, None] * ils32)
    r2_simt = np.zeros((M, N), np.float32)
    for d in range(D):
        r2_simt = r2_simt + df[..., d] * df[..., d]
    # tensor path
    x32, c32 = X.astype(np.float32), C.astype(np.float32)
    eq = sexp((np.abs(x32).max() + np.abs(c32).max()) ** 2) & ~1
    h = np.float32(2.0 ** (eq // 2))
    q32 = ((x32 * h)[None] - (c32 * h)[:, None]) ** 2                 # float32, exactly scaled by 2^eq
    w32 = (ils32 * ils32)
    ew = sexp(w32.max())
    qh, ql = split16(q32)
    wh, wl = split16(w32 * np.float32(2.0 ** ew))
    acc = (ql @ wh).astype(np.float32) + (qh @ wl).astype(np.float32) + (qh @ wh).astype(np.float32)   # fp32 accumulate
    r2_tc = acc * np.float32(2.0 ** -(eq + ew))
    def rel(a):
        return np.abs(a.astype(np.float64) - r2_64) / r2_64
    k64 = m52(r2_64)
    print("D=%2d ls in %s: r2 rel err  simt max %.2e mean %.2e | tc max %.2e mean %.2e ;  |dk| simt %.2e tc %.2e (near pairs r2<0.05: tc %.2e)" % (
        D, lsr, rel(r2_simt).max(), rel(r2_simt).mean(), rel(r2_tc).max(), rel(r2_tc).mean(),
        np.abs(m52(r2_simt.astype(np.float64)) - k64).max(), np.abs(m52(r2_tc.astype(np.float64)) - k64).max(),
        np.abs(m52(r2_tc.astype(np.float64)) - k64)[r2_64 < 0.05].max() if (r2_64 < 0.05).any() else 0.0))
