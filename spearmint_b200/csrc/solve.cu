// solve.cu -- alpha = K^-1 (y - mean) by forward + backward substitution against the blocked factor,
// plus the two scalars of the GP log marginal likelihood.
// Reference: spla.cho_solve((L, True), vals - mean) OPT:543 (vector) / OPT:603 ((N+P) x F fantasies);
// -sum(log(diag(chol))) - 0.5 * dot(vals - mean, solve)  OPT:637-640, 659-661, 690-692.
//
// One block per (sample, group of RB right-hand sides).  The right-hand sides live in shared memory;
// the factor is streamed once per direction (HBM/L2-bound: N^2/2 elements each way).  Diagonal blocks
// are applied as multiplications by the stored inverses W_II (potrf.cu).
#include "common.cuh"

namespace smk {

template <typename T> struct SolveCfg;
template <> struct SolveCfg<float>  { static constexpr int RB = 4; };
template <> struct SolveCfg<double> { static constexpr int RB = 2; };

__device__ __forceinline__ float  smk_log(float x)  { return logf(x); }
__device__ __forceinline__ double smk_log(double x) { return log(x); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ T block_sum(T v, T* red /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T t = T(0);
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

template <typename T>
__global__ void __launch_bounds__(256) chol_solve_kernel(int N, int Npad, int F, const T* __restrict__ L,
                                                          const T* __restrict__ winv, const T* __restrict__ y,
                                                          long long y_stride, int ldy,
                                                          const T* __restrict__ mean, T* __restrict__ alpha,
                                                          T* __restrict__ sum_log_diag, T* __restrict__ quad,
                                                          int do_backward) {
  constexpr int NB = Cfg<T>::NB, RB = SolveCfg<T>::RB, NG = 256 / NB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* x = reinterpret_cast<T*>(smem_raw);  // [RB][Npad]
  T* tt = x + (long)RB * Npad;            // [RB][NB]
  T* red = tt + RB * NB;                  // [NG][RB][NB]
  __shared__ T red8[8];

  const int s = blockIdx.y, f0 = blockIdx.x * RB, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const T* Ls = L + (long)s * Npad * Npad;
  const T* Ws = winv + (long)s * (Npad / NB) * NB * NB;
  const int nblk = Npad / NB;
  const T mu = mean ? mean[s] : T(0);

  for (int n = tid; n < Npad; n += 256)
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      int f = f0 + r;
      x[r * Npad + n] = (n < N && f < F) ? y[(long long)s * y_stride + (long)f * ldy + n] - mu : T(0);
    }
  __syncthreads();

  // ---------------- forward: L t = r
  for (int I = 0; I < nblk; ++I) {
    const int base = I * NB;
    const T* Wb = Ws + (long)I * NB * NB;
    // phase A: tt[i] = r[base+i] - L[base+i, 0:base] . t[0:base]   (a warp owns 4 rows at a time)
    for (int i0 = warp * 4; i0 < NB; i0 += 32) {
      T p[4][RB];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < RB; ++r) p[q][r] = T(0);
      const T* Lr = Ls + (long)(base + i0) * Npad;
      for (int k = lane; k < base; k += 32) {
        T l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) l[q] = Lr[(long)q * Npad + k];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          T xv = x[r * Npad + k];
#pragma unroll
          for (int q = 0; q < 4; ++q) p[q][r] = fma(l[q], xv, p[q][r]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          T v = warp_sum(p[q][r]);
          if (lane == 0) tt[r * NB + i0 + q] = x[r * Npad + base + i0 + q] - v;
        }
    }
    __syncthreads();
    // phase B: t[base+i] = W_II[i, 0:i+1] . tt
    for (int i = warp; i < NB; i += 8) {
      T p[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) p[r] = T(0);
      for (int k = lane; k <= i; k += 32) {
        T w = Wb[(long)i * NB + k];
#pragma unroll
        for (int r = 0; r < RB; ++r) p[r] = fma(w, tt[r * NB + k], p[r]);
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        T v = warp_sum(p[r]);
        if (lane == 0) x[r * Npad + base + i] = v;
      }
    }
    __syncthreads();
  }

  // Rows >= N are outside the (leading) system being solved: when L is the factor of a larger joint
  // matrix (observed + pending, OPT:574 "use the sub-Cholesky") they hold joint-factor rows, not the
  // identity, so their forward values must not leak into the quadratic form or the backward pass.
  for (int n = N + tid; n < Npad; n += 256)
#pragma unroll
    for (int r = 0; r < RB; ++r) x[r * Npad + n] = T(0);
  __syncthreads();

  if (quad) {
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      T a = T(0);
      for (int n = tid; n < Npad; n += 256) a = fma(x[r * Npad + n], x[r * Npad + n], a);
      a = block_sum(a, red8);
      if (tid == 0 && f0 + r < F) quad[(long)s * F + f0 + r] = a;
    }
  }
  if (sum_log_diag && blockIdx.x == 0) {
    T a = T(0);
    for (int n = tid; n < N; n += 256) a += smk_log(Ls[(long)n * Npad + n]);
    a = block_sum(a, red8);
    if (tid == 0) sum_log_diag[s] = a;
  }
  if (!do_backward) return;

  // ---------------- backward: L^T a = t
  const int i = tid % NB, g = tid / NB;
  for (int I = nblk - 1; I >= 0; --I) {
    const int base = I * NB;
    const T* Wb = Ws + (long)I * NB * NB;
    T p[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) p[r] = T(0);
    // phase A: sum over rows k below the block of L[k, base+i] * a[k]
    const T* Lc = Ls + base + i;
    int k = base + NB + g;
    for (; k + 3 * NG < Npad; k += 4 * NG) {
      T l0 = Lc[(long)k * Npad], l1 = Lc[(long)(k + NG) * Npad], l2 = Lc[(long)(k + 2 * NG) * Npad],
        l3 = Lc[(long)(k + 3 * NG) * Npad];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        p[r] = fma(l0, x[r * Npad + k], p[r]);
        p[r] = fma(l1, x[r * Npad + k + NG], p[r]);
        p[r] = fma(l2, x[r * Npad + k + 2 * NG], p[r]);
        p[r] = fma(l3, x[r * Npad + k + 3 * NG], p[r]);
      }
    }
    for (; k < Npad; k += NG) {
      T l0 = Lc[(long)k * Npad];
#pragma unroll
      for (int r = 0; r < RB; ++r) p[r] = fma(l0, x[r * Npad + k], p[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) red[(g * RB + r) * NB + i] = p[r];
    __syncthreads();
    if (g == 0) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        T v = x[r * Npad + base + i];
        for (int gg = 0; gg < NG; ++gg) v -= red[(gg * RB + r) * NB + i];
        tt[r * NB + i] = v;
      }
    }
    __syncthreads();
    // phase B: a[base+i] = sum_{k>=i} W_II[k, i] * tt[k]
#pragma unroll
    for (int r = 0; r < RB; ++r) p[r] = T(0);
    for (int kk = i + g; kk < NB; kk += NG) {
      T w = Wb[(long)kk * NB + i];
#pragma unroll
      for (int r = 0; r < RB; ++r) p[r] = fma(w, tt[r * NB + kk], p[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) red[(g * RB + r) * NB + i] = p[r];
    __syncthreads();
    if (g == 0) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        T v = T(0);
        for (int gg = 0; gg < NG; ++gg) v += red[(gg * RB + r) * NB + i];
        x[r * Npad + base + i] = v;
      }
    }
    __syncthreads();
  }
  for (int n = tid; n < Npad; n += 256)
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (f0 + r < F) alpha[((long)s * F + f0 + r) * Npad + n] = x[r * Npad + n];
}

template <typename T>
int chol_solve(int N, int Npad, int S, int F, const T* L, const T* winv, const T* y, long long y_stride, int ldy,
               const T* mean, T* alpha, T* sum_log_diag, T* quad, cudaStream_t st) {
  constexpr int NB = Cfg<T>::NB, RB = SolveCfg<T>::RB, NG = 256 / NB;
  if (N <= 0) return -1;
  if (Npad < N || Npad % kNpadMult) return -2;
  if (S <= 0) return -3;
  if (F <= 0) return -4;
  if (!L) return -5;
  if (!winv) return -6;
  if (!y) return -7;
  if (ldy < N) return -9;
  const size_t dsm = sizeof(T) * ((size_t)RB * Npad + RB * NB + (size_t)NG * RB * NB);
  if (dsm > 227 * 1024) return -2;
  static size_t attr_set = 0;
  if (dsm > attr_set) {
    cudaFuncSetAttribute(chol_solve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    attr_set = dsm;
  }
  dim3 grid((F + RB - 1) / RB, S);
  chol_solve_kernel<T><<<grid, 256, dsm, st>>>(N, Npad, F, L, winv, y, y_stride, ldy, mean, alpha, sum_log_diag,
                                              quad, alpha != nullptr);
  count_launch();
  return check_launch("chol_solve");
}

// ---------------------------------------------------------------------------------------------------------------
// Log-likelihood by augmentation: put r = y - mean in row N of the (padded) covariance before factoring it,
//   [ K  . ]   [ L        0 ] [ L^T  L^-1 r ]
//   [ r' c ] = [ (L^-1r)' * ] [ 0    *      ]
// so the Cholesky panel/trailing kernels perform the forward substitution as part of the factorisation (one extra
// row in a block row that exists anyway) and   quad = |L[N, 0:N]|^2,   sum_log_diag = sum_{i<N} log L_ii.
// Needs Npad > N (the caller pads to ceil128(N + 1)).  Replaces the serial chol_solve on the slice-sampler path.
template <typename T>
__global__ void loglik_set_rhs_kernel(int N, int Npad, const T* __restrict__ y, const T* __restrict__ mean, T* A) {
  const int s = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
  T* row = A + (long)s * Npad * Npad + (long)N * Npad;
  if (n < N) row[n] = y[n] - mean[s];
  else if (n == N) row[n] = T(1e30);            // pivot of the augmented row: c - quad must stay positive
}

template <typename T>
__global__ void __launch_bounds__(256) loglik_finish_kernel(int N, int Npad, const T* __restrict__ L, T* __restrict__ sld,
                                                             T* __restrict__ quad) {
  __shared__ T red8[8];
  const int s = blockIdx.x, tid = threadIdx.x;
  const T* Ls = L + (long)s * Npad * Npad;
  T a = T(0), q = T(0);
  for (int n = tid; n < N; n += 256) {
    a += smk_log(Ls[(long)n * Npad + n]);
    T v = Ls[(long)N * Npad + n];
    q = fma(v, v, q);
  }
  a = block_sum(a, red8);
  q = block_sum(q, red8);
  if (tid == 0) { sld[s] = a; quad[s] = q; }
}

template <typename T>
int loglik_set_rhs(int N, int Npad, int S, const T* y, const T* mean, T* A, cudaStream_t st) {
  if (N <= 0 || Npad <= N || Npad % kNpadMult) return -2;
  if (S <= 0) return -3;
  if (!y || !mean || !A) return -4;
  loglik_set_rhs_kernel<T><<<dim3((N + 256) / 256, S), 256, 0, st>>>(N, Npad, y, mean, A);
  count_launch();
  return check_launch("loglik_set_rhs");
}
template <typename T>
int loglik_finish(int N, int Npad, int S, const T* L, T* sld, T* quad, cudaStream_t st) {
  if (N <= 0 || Npad <= N) return -2;
  if (S <= 0) return -3;
  if (!L || !sld || !quad) return -4;
  loglik_finish_kernel<T><<<S, 256, 0, st>>>(N, Npad, L, sld, quad);
  count_launch();
  return check_launch("loglik_finish");
}
template int loglik_set_rhs<float>(int, int, int, const float*, const float*, float*, cudaStream_t);
template int loglik_set_rhs<double>(int, int, int, const double*, const double*, double*, cudaStream_t);
template int loglik_finish<float>(int, int, int, const float*, float*, float*, cudaStream_t);
template int loglik_finish<double>(int, int, int, const double*, double*, double*, cudaStream_t);

template int chol_solve<float>(int, int, int, int, const float*, const float*, const float*, long long, int,
                               const float*, float*, float*, float*, cudaStream_t);
template int chol_solve<double>(int, int, int, int, const double*, const double*, const double*, long long, int,
                                const double*, double*, double*, double*, cudaStream_t);

}  // namespace smk
