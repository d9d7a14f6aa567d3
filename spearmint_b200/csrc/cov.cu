// cov.cu -- batched covariance build (reference: gp.dist2 GP:34-54, kernels GP:87-127,
// chooser.cov OPT:207-212 / PSEC:145-150, caller's noise*I OPT:539).
//
// One launch builds the matrices of ALL hyper-samples (grid.z = sample).  The scaled squared
// distance is accumulated from direct differences ((x_i - y_j) * inv_ls)^2 rather than the
// reference's expanded -2xy + |x|^2 + |y|^2 form: in float the expanded form loses ~|x/ls|^2 * eps
// absolutely, which is fatal for the 1e-3-wide jitter cloud around the incumbent (OPT:236-238).
// The kernel is output-bound (D <= 32 flops-equivalent per 4 or 8 byte store); roofline = HBM write.
#include "common.cuh"

namespace smk {

constexpr int kCT = 32;   // output tile edge
constexpr int kDC = 32;   // D chunk staged in shared memory

template <typename T>
__global__ void __launch_bounds__(256) cov_build_kernel(int kind, int N, int M, int D, const T* __restrict__ X,
                                                         const T* __restrict__ Y, const T* __restrict__ inv_ls,
                                                         const T* __restrict__ amp2, const T* __restrict__ diag_add,
                                                         T* __restrict__ out, int ld, int self, int lower) {
  if (lower && blockIdx.x > blockIdx.y) return;            // tile strictly above the diagonal: the factorisations never read it
  __shared__ T xs[kCT][kDC + 1];
  __shared__ T ys[kCT][kDC + 1];
  const int s = blockIdx.z;
  const int i0 = blockIdx.y * kCT, j0 = blockIdx.x * kCT;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const T* ils = inv_ls + (long)s * D;
  const T* Yp = self ? X : Y;
  const int rowsX = N, rowsY = self ? N : M;

  T r2[4] = {T(0), T(0), T(0), T(0)};
  for (int d0 = 0; d0 < D; d0 += kDC) {
    // stage scaled rows: thread t loads element (row = t/32 + 8k, dd = t%32)
    for (int k = 0; k < 4; ++k) {
      int row = ty + 8 * k, dd = tx, d = d0 + dd;
      int gi = i0 + row, gj = j0 + row;
      T sc = (d < D) ? ils[d] : T(0);
      xs[row][dd] = (d < D && gi < rowsX) ? X[(long)gi * D + d] * sc : T(0);
      ys[row][dd] = (d < D && gj < rowsY) ? Yp[(long)gj * D + d] * sc : T(0);
    }
    __syncthreads();
    const int dmax = min(kDC, D - d0);
    for (int dd = 0; dd < dmax; ++dd) {
      T y = ys[tx][dd];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        T df = xs[ty + 8 * k][dd] - y;
        r2[k] = fma(df, df, r2[k]);
      }
    }
    __syncthreads();
  }
  const T a2 = amp2[s];
  const T dg = self ? (a2 * T(1e-6) + (diag_add ? diag_add[s] : T(0))) : T(0);
  const int j = j0 + tx;
  const int nrows = self ? ld : N, ncols = self ? ld : M;
  T* o = out + (long)s * (self ? (long)ld * ld : (long)N * ld);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int i = i0 + ty + 8 * k;
    if (i >= nrows || j >= ncols) continue;
    T val;
    if (self && (i >= N || j >= N)) {
      val = (i == j) ? T(1) : T(0);  // identity on the padding
    } else {
      val = a2 * kernel_of_r2<T>(kind, r2[k]);
      if (self && i == j) val += dg;
    }
    o[(long)i * ld + j] = val;
  }
}

template <typename T>
int cov_build(int kind, int N, int M, int D, int S, const T* X, const T* Y, const T* inv_ls, const T* amp2,
              const T* diag_add, T* out, int ld, cudaStream_t st, int lower) {
  if (kind < 0 || kind > 3) return -1;
  if (N <= 0) return -2;
  if (D <= 0) return -4;
  if (S <= 0) return -5;
  if (!X) return -6;
  if (!inv_ls) return -8;
  if (!amp2) return -9;
  if (!out) return -11;
  const int self = (Y == nullptr);
  if (lower && !self) return -13;
  if (!self && M <= 0) return -3;
  if (ld < (self ? N : M)) return -12;
  const int rows = self ? ld : N, cols = self ? ld : M;
  dim3 grid((cols + kCT - 1) / kCT, (rows + kCT - 1) / kCT, S);
  cov_build_kernel<T><<<grid, 256, 0, st>>>(kind, N, M, D, X, Y, inv_ls, amp2, diag_add, out, ld, self, lower);
  count_launch();
  return check_launch("cov_build");
}

template int cov_build<float>(int, int, int, int, int, const float*, const float*, const float*, const float*,
                              const float*, float*, int, cudaStream_t, int);
template int cov_build<double>(int, int, int, int, int, const double*, const double*, const double*,
                               const double*, const double*, double*, int, cudaStream_t, int);

}  // namespace smk
