// potrf.cu -- batched blocked right-looking lower Cholesky (reference: spla.cholesky(K, lower=True)
// at OPT:540, OPT:567, OPT:585 and inside every slice-sampler logprob OPT:637, 659, 690).
//
// Per block column j (NB = 128 for float, 64 for double), for ALL hyper-samples at once (grid.z):
//   diag     : L_jj = chol(A_jj) in shared memory, W_jj = L_jj^-1 (kept: every later triangular
//              solve multiplies by W_jj instead of substituting)
//   panel    : L_Ij = A_Ij * W_jj^T                       (block GEMM, I > j)
//   trailing : A_IK -= L_Ij * L_Kj^T   for I >= K > j     (block SYRK/GEMM)
// The trailing update carries N^3/3 of the flops and is the tensor-core candidate.
#include "common.cuh"

namespace smk {

// ------------------------------------------------------------------------------------------ diag
// Factor one NB x NB diagonal block and invert its factor, entirely on one SM (this kernel is the serial spine of
// the factorisation: nblk launches per matrix, so its latency -- not its flops -- is what matters, above all for
// the one-matrix-at-a-time log-likelihood calls of the slice sampler).
//   for each 32-wide sub-block:  (a) factor the 32 x 32 diagonal piece and invert it (block_chol_inv_32);
//   (b) rows below: X = A_sub * Wdd^T;  (c) rank-32 update of what is left.
//   Then W = L^-1 is assembled from the 32 x 32 inverses by block distance (d = 1, 2, ...).
template <typename T>
__device__ __forceinline__ T shfl_t(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// The whole block (256 threads) factors the 32x32 SPD piece at `a` (row stride lda, lower part) in place and writes its
// inverse into `w`.  One column (resp. one row of the inverse) per step, a few elements per thread, block barriers in
// between: the per-step latency is a barrier plus one shared-memory round trip (~100-150 cycles) instead of a
// 30-iteration dependent loop in a single warp.  `red` is a [8][32] scratch.
template <typename T>
__device__ __forceinline__ void block_chol_inv_32(T* a, int lda, T* w, int ldw, T* red, int tid, int& bad) {
  for (int j = 0; j < 32; ++j) {
    T d = a[j * lda + j];
    if (!(d > T(0))) { if (bad < 0) bad = j; d = T(1); }
    const T piv = smk_sqrt(d), ipiv = T(1) / piv;
    __syncthreads();                               // everybody has read the pivot
    if (tid == j) a[j * lda + j] = piv;
    else if (tid > j && tid < 32) a[tid * lda + j] *= ipiv;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {                  // rank-1 update of the trailing lower triangle (<= 496 elements)
      const int e = tid + q * 256, i = e >> 5, k = e & 31;
      if (k > j && k <= i) a[i * lda + k] = fma(-a[i * lda + j], a[k * lda + j], a[i * lda + k]);
    }
    __syncthreads();
  }
  // inverse by rows: W[i][c] = (delta_ic - sum_{c<=k<i} L[i][k] W[k][c]) / L[i][i], k split over 8 thread groups
  const int c = tid & 31, part = tid >> 5;
  for (int i = 0; i < 32; ++i) {
    T acc = T(0);
    for (int k = c + part; k < i; k += 8) acc = fma(a[i * lda + k], w[k * ldw + c], acc);
    red[part * 32 + c] = acc;
    __syncthreads();
    if (part == 0) {
      T sum = T(0);
#pragma unroll
      for (int g = 0; g < 8; ++g) sum += red[g * 32 + c];
      w[i * ldw + c] = (c <= i) ? (((c == i) ? T(1) : T(0)) - sum) / a[i * lda + i] : T(0);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256) potrf_diag_kernel(int Npad, int jb, T* __restrict__ A,
                                                          T* __restrict__ winv, int* __restrict__ info) {
  constexpr int NB = Cfg<T>::NB, SB = 32, NSB = NB / SB;
  constexpr int LDS = NB + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* a = reinterpret_cast<T*>(smem_raw);   // [NB][LDS] block being factored (lower)
  T* w = a + NB * LDS;                     // [NB][LDS] its inverse (lower)
  T* t = w + NB * LDS;                     // [NB][SB+1] scratch for the panel / inverse assembly
  constexpr int LDT = SB + 1;
  const int s = blockIdx.x, tid = threadIdx.x;
  T* Ab = A + (long)s * Npad * Npad + (long)jb * NB * Npad + (long)jb * NB;

  for (int e0 = tid; e0 < NB * NB; e0 += 256 * 8) {
    T v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int e = e0 + q * 256, i = e / NB, k = e % NB;
      v[q] = (e < NB * NB && k <= i) ? Ab[(long)i * Npad + k] : T(0);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int e = e0 + q * 256, i = e / NB, k = e % NB;
      if (e < NB * NB) { a[i * LDS + k] = v[q]; w[i * LDS + k] = T(0); }
    }
  }
  __syncthreads();

  for (int sb = 0; sb < NSB; ++sb) {
    const int o = sb * SB;
    // (a) 32x32 diagonal piece: factor + invert
    {
      int bad = -1;
      block_chol_inv_32<T>(a + o * LDS + o, LDS, w + o * LDS + o, LDS, t, tid, bad);
      if (bad >= 0 && tid == 0 && info[s] == 0) info[s] = jb * NB + o + bad + 1;
    }
    __syncthreads();
    const int rows = NB - o - SB;                    // rows below the diagonal piece
    if (rows > 0) {
      // (b) X[r][k] = sum_{m<=k} A[r][o+m] * Wdd[k][m]   -> scratch, then back into a
      for (int e = tid; e < rows * SB; e += 256) {
        int rr = e / SB, k = e % SB;
        const T* ar = a + (o + SB + rr) * LDS + o;
        const T* wk = w + (o + k) * LDS + o;
        T acc = T(0);
        for (int m = 0; m <= k; ++m) acc = fma(ar[m], wk[m], acc);
        t[rr * LDT + k] = acc;
      }
      __syncthreads();
      for (int e = tid; e < rows * SB; e += 256) {
        int rr = e / SB, k = e % SB;
        a[(o + SB + rr) * LDS + o + k] = t[rr * LDT + k];
      }
      // (c) rank-32 update of the remaining lower triangle: a[r][c] -= X[r] . X[c]   (r >= c)
      for (int e = tid; e < rows * rows; e += 256) {
        int rr = e / rows, cc = e % rows;
        if (cc > rr) continue;
        const T* xr = t + rr * LDT;
        const T* xc = t + cc * LDT;
        T acc = a[(o + SB + rr) * LDS + o + SB + cc];
#pragma unroll 8
        for (int k = 0; k < SB; ++k) acc = fma(-xr[k], xc[k], acc);
        a[(o + SB + rr) * LDS + o + SB + cc] = acc;
      }
      __syncthreads();
    }
  }

  // W = L^-1: off-diagonal 32x32 blocks by block distance d:  W_ij = -W_ii * (sum_{k=j}^{i-1} L_ik W_kj)
  for (int d = 1; d < NSB; ++d) {
    const int npair = NSB - d;                       // (i, j) = (j + d, j)
    // phase 1: T_ij = sum_k L_ik W_kj  into scratch t[(pair*32 + r)][c]
    for (int e = tid; e < npair * SB * SB; e += 256) {
      int pr = e / (SB * SB), rr = (e / SB) % SB, cc = e % SB;
      int j = pr, i = pr + d;
      T acc = T(0);
      for (int kb = j; kb < i; ++kb) {
        const T* lrow = a + (i * SB + rr) * LDS + kb * SB;
#pragma unroll 8
        for (int m = 0; m < SB; ++m) acc = fma(lrow[m], w[(kb * SB + m) * LDS + j * SB + cc], acc);
      }
      t[(pr * SB + rr) * LDT + cc] = acc;
    }
    __syncthreads();
    // phase 2: W_ij = -W_ii * T_ij
    for (int e = tid; e < npair * SB * SB; e += 256) {
      int pr = e / (SB * SB), rr = (e / SB) % SB, cc = e % SB;
      int j = pr, i = pr + d;
      const T* wrow = w + (i * SB + rr) * LDS + i * SB;
      T acc = T(0);
      for (int m = 0; m <= rr; ++m) acc = fma(wrow[m], t[(pr * SB + m) * LDT + cc], acc);
      w[(i * SB + rr) * LDS + j * SB + cc] = -acc;
    }
    __syncthreads();
  }

  T* Wb = winv + ((long)s * (Npad / NB) + jb) * NB * NB;
  for (int e = tid; e < NB * NB; e += 256) {
    int i = e / NB, k = e % NB;
    if (k <= i) Ab[(long)i * Npad + k] = a[i * LDS + k];
    Wb[e] = w[i * LDS + k];
  }
}

// ------------------------------------------------------------------------------------------ panel
// L_Ij = A_Ij * W_jj^T  (one block per row tile I > jb)
template <typename T>
__global__ void __launch_bounds__(256) potrf_panel_kernel(int Npad, int jb, T* __restrict__ A,
                                                           const T* __restrict__ winv, T* __restrict__ hi = nullptr,
                                                           T* __restrict__ lo = nullptr) {
  using C = Cfg<T>;
  constexpr int NB = C::NB;
  __shared__ TileSmem<T> sm;
  const int s = blockIdx.z, I = jb + 1 + blockIdx.x;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  T* Aij = A + (long)s * Npad * Npad + (long)I * NB * Npad + (long)jb * NB;
  const T* W = winv + ((long)s * (Npad / NB) + jb) * NB * NB;
  T acc[C::TM][C::TN];
#pragma unroll
  for (int r = 0; r < C::TM; ++r)
#pragma unroll
    for (int c = 0; c < C::TN; ++c) acc[r][c] = T(0);
  TileGemm<T, Lay::KContig, Lay::KContig, false>::run(acc, Aij, Npad, W, NB, NB, sm);
  const long off = (long)s * Npad * Npad + (long)I * NB * Npad + (long)jb * NB;
#pragma unroll
  for (int r = 0; r < C::TM; ++r)
#pragma unroll
    for (int g = 0; g < C::TN / 4; ++g) {
      V4<T> v;
#pragma unroll
      for (int e = 0; e < 4; ++e) v.v[e] = acc[r][g * 4 + e];
      const long o = (long)tile_row(ty, r) * Npad + tile_col(tx, g * 4);
      st4(Aij + o, v);
      if (sizeof(T) == 4 && hi != nullptr) {     // tf32 hi/lo copies of the finished panel (operands of the tcgen05 update)
        V4<T> h, l;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = (float)v.v[e];
          float xh = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
          h.v[e] = (T)xh;
          l.v[e] = (T)(x - xh);
        }
        st4(hi + off + o, h);
        st4(lo + off + o, l);
      }
    }
}

// --------------------------------------------------------------------------------------- trailing
// A_IK -= L[I, jb:jb+kd] * L[K, jb:jb+kd]^T   for c0 <= K <= I   (grid.x = I - c0, grid.y = K - c0), kd = depth in
// block columns: the driver updates block columns in PAIRS (rank 2*NB), halving the accumulator load/store traffic.
template <typename T>
__global__ void __launch_bounds__(256, 2) potrf_trailing_kernel(int Npad, int jb, int kd, int c0, T* __restrict__ A) {
  using C = Cfg<T>;
  constexpr int NB = C::NB;
  if (blockIdx.y > blockIdx.x) return;
  __shared__ TileSmem<T> sm;
  const int s = blockIdx.z, I = c0 + blockIdx.x, K = c0 + blockIdx.y;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  T* As = A + (long)s * Npad * Npad;
  T* Aik = As + (long)I * NB * Npad + (long)K * NB;
  const T* Lij = As + (long)I * NB * Npad + (long)jb * NB;
  const T* Lkj = As + (long)K * NB * Npad + (long)jb * NB;
  T acc[C::TM][C::TN];
#pragma unroll
  for (int r = 0; r < C::TM; ++r)
#pragma unroll
    for (int g = 0; g < C::TN / 4; ++g) {
      V4<T> v = ld4(Aik + (long)tile_row(ty, r) * Npad + tile_col(tx, g * 4));
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[r][g * 4 + e] = v.v[e];
    }
  TileGemm<T, Lay::KContig, Lay::KContig, true>::run(acc, Lij, Npad, Lkj, Npad, kd * NB, sm);
#pragma unroll
  for (int r = 0; r < C::TM; ++r)
#pragma unroll
    for (int g = 0; g < C::TN / 4; ++g) {
      V4<T> v;
#pragma unroll
      for (int e = 0; e < 4; ++e) v.v[e] = acc[r][g * 4 + e];
      st4(Aik + (long)tile_row(ty, r) * Npad + tile_col(tx, g * 4), v);
    }
}

template <typename T>
int potrf_lower_batched(int Npad, int S, T* A, T* winv, int* info, cudaStream_t st) {
  constexpr int NB = Cfg<T>::NB;
  if (Npad <= 0 || Npad % kNpadMult) return -1;
  if (S <= 0) return -2;
  if (!A) return -3;
  if (!winv) return -4;
  if (!info) return -5;
  const int nblk = Npad / NB;
  const size_t dsm = (2 * (size_t)NB * (NB + 1) + (size_t)NB * 33) * sizeof(T);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(potrf_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    attr_done = true;
  }
  cudaMemsetAsync(info, 0, sizeof(int) * S, st);
  // Block columns in pairs (jb, jb+1): factor jb, bring only column jb+1 up to date (rank NB), factor jb+1, then ONE
  // rank-2*NB update of everything to the right.
  for (int jb = 0; jb < nblk; jb += 2) {
    potrf_diag_kernel<T><<<S, 256, dsm, st>>>(Npad, jb, A, winv, info);
    count_launch();
    int rem = nblk - jb - 1;
    if (rem <= 0) break;
    potrf_panel_kernel<T><<<dim3(rem, 1, S), 256, 0, st>>>(Npad, jb, A, winv);
    potrf_trailing_kernel<T><<<dim3(rem, 1, S), 256, 0, st>>>(Npad, jb, 1, jb + 1, A);      // column jb+1 only
    potrf_diag_kernel<T><<<S, 256, dsm, st>>>(Npad, jb + 1, A, winv, info);
    count_launch(3);
    rem = nblk - jb - 2;
    if (rem <= 0) break;
    potrf_panel_kernel<T><<<dim3(rem, 1, S), 256, 0, st>>>(Npad, jb + 1, A, winv);
    potrf_trailing_kernel<T><<<dim3(rem, rem, S), 256, 0, st>>>(Npad, jb, 2, jb + 2, A);   // rank 2*NB, columns >= jb+2
    count_launch(2);
  }
  return check_launch("potrf_lower_batched");
}

// ---- left-looking variant with the rank-(jb*NB) update of each block-column pair on the tensor cores (float32):
//   update(jb)  : A[rows >= jb][cols jb, jb+1] -= L[rows, 0:jb] L[jb:jb+2, 0:jb]^T     tcgen05 3xTF32 (predict_tc.cu)
//   diag, panel : as above; the panel kernel also writes the tf32 hi/lo copies the next updates read through TMA
//   column jb+1 is brought up to date w.r.t. column jb by one rank-NB SIMT update.
int tc_chol_update(int Npad, int S, int jb, int ncols, float* A, const float* lhi, const float* llo, cudaStream_t st);

int potrf_lower_batched_tc(int Npad, int S, float* A, float* winv, int* info, float* lhi, float* llo, cudaStream_t st) {
  constexpr int NB = Cfg<float>::NB;
  if (Npad <= 0 || Npad % kNpadMult) return -1;
  if (S <= 0) return -2;
  if (!A || !winv || !info || !lhi || !llo) return -3;
  const int nblk = Npad / NB;
  const size_t dsm = (2 * (size_t)NB * (NB + 1) + (size_t)NB * 33) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(potrf_diag_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    attr_done = true;
  }
  cudaMemsetAsync(info, 0, sizeof(int) * S, st);
  for (int jb = 0; jb < nblk; jb += 2) {
    if (jb > 0) {
      int rc = tc_chol_update(Npad, S, jb, (jb + 1 < nblk) ? 2 * NB : NB, A, lhi, llo, st);
      if (rc) return rc;
    }
    potrf_diag_kernel<float><<<S, 256, dsm, st>>>(Npad, jb, A, winv, info);
    count_launch();
    int rem = nblk - jb - 1;
    if (rem <= 0) break;
    potrf_panel_kernel<float><<<dim3(rem, 1, S), 256, 0, st>>>(Npad, jb, A, winv, lhi, llo);
    potrf_trailing_kernel<float><<<dim3(rem, 1, S), 256, 0, st>>>(Npad, jb, 1, jb + 1, A);
    potrf_diag_kernel<float><<<S, 256, dsm, st>>>(Npad, jb + 1, A, winv, info);
    count_launch(3);
    rem = nblk - jb - 2;
    if (rem <= 0) break;
    potrf_panel_kernel<float><<<dim3(rem, 1, S), 256, 0, st>>>(Npad, jb + 1, A, winv, lhi, llo);
    count_launch();
  }
  return check_launch("potrf_lower_batched_tc");
}

template int potrf_lower_batched<float>(int, int, float*, float*, int*, cudaStream_t);
template int potrf_lower_batched<double>(int, int, double*, double*, int*, cudaStream_t);

}  // namespace smk
