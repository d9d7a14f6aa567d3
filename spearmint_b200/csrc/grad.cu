// grad.cu -- EI value + input-gradient terms at a few query points with CACHED factors.
//
// Reference: GPEIOptChooser.grad_optimize_ei (OPT:391-525) rebuilds K and re-factors it on every L-BFGS
// function evaluation (OPT:397-402; ~160 evaluations x S Choleskys per next()).  Here the S factors are
// computed once; one evaluation is  cov_build(cross, N x Q) -> chol_solve (gamma = K^-1 kx) -> this kernel.
// It also retires the reference's only native snippet, the weave loop of grad_dist2 (GP:69-79): the
// N x M x D gradient tensor is never materialised, it is contracted on the fly.
//
// For sample s and query point q this kernel forms the small product  OUT = A * T  with
//     A[f][n] = alpha[s][f][n]  (f < F),   A[F][n] = gamma[s][q][n]
//     T[n][d] = gk[n][d]        (d < D),   T[n][D] = kx[n] = amp2 * k(r2_n)
//     gk[n][d] = dk/dr2(r2_n) * (2 * inv_ls_d) * (X[n][d] - x_q[d]) * inv_ls_d        (GP:56-85, GP:102-132)
// so that (host side, float64):
//     OUT[f][D]  = kx' alpha_f           -> func_m - mean            (OPT:417, 508)
//     OUT[f][d]  = grad_xp_m[f][d]                                    (OPT:433, 516)
//     OUT[F][d]  = -0.5 * grad_xp_v[d]                                (OPT:434-435, 517-518)
//     OUT[F][D]  = kx' K^-1 kx = sum(beta^2)                          (OPT:418, 509)
#include "common.cuh"

namespace smk {

template <typename T>
__device__ __forceinline__ T dk_dr2(int kind, T r2) {
  if (kind <= 1) return T(-0.5) * smk_exp(T(-0.5) * r2);                              // GP:102-105
  T r = smk_sqrt(r2);
  if (kind == 2) return T(-1.5) * smk_exp(-T(1.7320508075688772) * r);                // GP:115-118
  T a = T(2.23606797749979) * r;
  return T(-5.0 / 6.0) * smk_exp(-a) * (T(1) + a);                                    // GP:129-132
}

constexpr int kGC = 64;  // rows of X per staged chunk

template <typename T>
__global__ void __launch_bounds__(256) ei_grad_terms_kernel(int kind, int N, int Npad, int D, int Q, int F,
                                                             const T* __restrict__ X, const T* __restrict__ xq,
                                                             const T* __restrict__ inv_ls,
                                                             const T* __restrict__ amp2,
                                                             const T* __restrict__ alpha,
                                                             const T* __restrict__ gamma, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D1 = D + 1, F1 = F + 1;
  T* Tt = reinterpret_cast<T*>(smem_raw);   // [kGC][D1]
  T* At = Tt + kGC * D1;                    // [F1][kGC]
  T* xs = At + F1 * kGC;                    // [D] scaled query point
  T* il = xs + D;                           // [D]
  const int q = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  for (int d = tid; d < D; d += 256) {
    T sc = inv_ls[(long)s * D + d];
    il[d] = sc;
    xs[d] = xq[(long)q * D + d] * sc;
  }
  __syncthreads();
  const T a2 = amp2[s];
  const T* al = alpha + (long)s * F * Npad;
  const T* ga = gamma + ((long)s * Q + q) * Npad;
  T* o = out + ((long)s * Q + q) * F1 * D1;
  const int npairs = F1 * D1;
  constexpr int R = 8;

  for (int p0 = 0; p0 < npairs; p0 += 256 * R) {
    T acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = T(0);
    for (int n0 = 0; n0 < N; n0 += kGC) {
      __syncthreads();
      // stage T chunk: one thread per row computes r2 and the kernel derivatives, then all fill columns
      for (int i = tid; i < kGC; i += 256) {
        int n = n0 + i;
        T r2 = T(0);
        if (n < N)
          for (int d = 0; d < D; ++d) {
            T df = X[(long)n * D + d] * il[d] - xs[d];
            r2 = fma(df, df, r2);
          }
        T kv = (n < N) ? a2 * kernel_of_r2<T>(kind, r2) : T(0);
        T w = (n < N) ? dk_dr2<T>(kind, r2) : T(0);
        Tt[i * D1 + D] = kv;
        // park w in the first column slot temporarily (overwritten below after the barrier)
        At[i] = w;
      }
      __syncthreads();
      for (int e = tid; e < kGC * D; e += 256) {
        int i = e / D, d = e % D, n = n0 + i;
        T w = At[i];
        T g = (n < N) ? w * (T(2) * il[d]) * (X[(long)n * D + d] * il[d] - xs[d]) : T(0);
        Tt[i * D1 + d] = g;
      }
      __syncthreads();
      for (int e = tid; e < F1 * kGC; e += 256) {
        int f = e / kGC, i = e % kGC, n = n0 + i;
        At[f * kGC + i] = (n < N) ? (f < F ? al[(long)f * Npad + n] : ga[n]) : T(0);
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        int p = p0 + r * 256 + tid;
        if (p < npairs) {
          int f = p / D1, d = p % D1;
          T a = acc[r];
          for (int i = 0; i < kGC; ++i) a = fma(At[f * kGC + i], Tt[i * D1 + d], a);
          acc[r] = a;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int p = p0 + r * 256 + tid;
      if (p < npairs) o[p] = acc[r];
    }
  }
}

template <typename T>
int ei_grad_terms(int kind, int N, int Npad, int D, int S, int Q, int F, const T* X, const T* xq, const T* inv_ls,
                  const T* amp2, const T* alpha, const T* gamma, T* out, cudaStream_t st) {
  if (kind < 0 || kind > 3) return -1;
  if (N <= 0 || Npad < N) return -2;
  if (D <= 0) return -4;
  if (S <= 0) return -5;
  if (Q <= 0) return -6;
  if (F <= 0) return -7;
  if (!X || !xq || !inv_ls || !amp2 || !alpha || !gamma || !out) return -8;
  const size_t dsm = sizeof(T) * ((size_t)kGC * (D + 1) + (size_t)(F + 1) * kGC + 2 * (size_t)D);
  if (dsm > 200 * 1024) return -4;
  static size_t attr_set = 48 * 1024;
  if (dsm > attr_set) {
    cudaFuncSetAttribute(ei_grad_terms_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    attr_set = dsm;
  }
  ei_grad_terms_kernel<T><<<dim3(Q, S), 256, dsm, st>>>(kind, N, Npad, D, Q, F, X, xq, inv_ls, amp2, alpha, gamma, out);
  count_launch();
  return check_launch("ei_grad_terms");
}

template int ei_grad_terms<float>(int, int, int, int, int, int, int, const float*, const float*, const float*,
                                  const float*, const float*, const float*, float*, cudaStream_t);
template int ei_grad_terms<double>(int, int, int, int, int, int, int, const double*, const double*, const double*,
                                   const double*, const double*, const double*, double*, cudaStream_t);

}  // namespace smk
