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
template <typename T>
__global__ void __launch_bounds__(256) potrf_diag_kernel(int Npad, int jb, T* __restrict__ A,
                                                          T* __restrict__ winv, int* __restrict__ info) {
  constexpr int NB = Cfg<T>::NB;
  constexpr int LDS = NB + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* a = reinterpret_cast<T*>(smem_raw);   // [NB][LDS] block being factored
  T* w = a + NB * LDS;                     // [NB][LDS] its inverse
  const int s = blockIdx.x, tid = threadIdx.x;
  T* Ab = A + (long)s * Npad * Npad + (long)jb * NB * Npad + (long)jb * NB;

  for (int e = tid; e < NB * NB; e += 256) {
    int i = e / NB, k = e % NB;
    a[i * LDS + k] = (k <= i) ? Ab[(long)i * Npad + k] : T(0);
    w[i * LDS + k] = T(0);
  }
  __syncthreads();

  const int ty = tid >> 4, tx = tid & 15;
  for (int j = 0; j < NB; ++j) {
    if (tid == 0) {
      T d = a[j * LDS + j];
      if (!(d > T(0))) {               // also catches NaN
        if (info[s] == 0) info[s] = jb * NB + j + 1;
        d = T(1);
      }
      a[j * LDS + j] = smk_sqrt(d);
    }
    __syncthreads();
    const T piv = T(1) / a[j * LDS + j];
    for (int i = j + 1 + tid; i < NB; i += 256) a[i * LDS + j] *= piv;
    __syncthreads();
    // rank-1 update of the trailing lower triangle
    for (int i = j + 1 + ty; i < NB; i += 16) {
      const T lij = a[i * LDS + j];
      for (int k = j + 1 + tx; k <= i; k += 16) a[i * LDS + k] = fma(-lij, a[k * LDS + j], a[i * LDS + k]);
    }
    __syncthreads();
  }

  // W = L^-1 by forward substitution, one column per thread (columns are independent).
  if (tid < NB) {
    const int c = tid;
    for (int i = c; i < NB; ++i) {
      T sacc = (i == c) ? T(1) : T(0);
      for (int k = c; k < i; ++k) sacc = fma(-a[i * LDS + k], w[k * LDS + c], sacc);
      w[i * LDS + c] = sacc / a[i * LDS + i];
    }
  }
  __syncthreads();

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
                                                           const T* __restrict__ winv) {
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
#pragma unroll
  for (int r = 0; r < C::TM; ++r)
#pragma unroll
    for (int g = 0; g < C::TN / 4; ++g) {
      V4<T> v;
#pragma unroll
      for (int e = 0; e < 4; ++e) v.v[e] = acc[r][g * 4 + e];
      st4(Aij + (long)tile_row(ty, r) * Npad + tile_col(tx, g * 4), v);
    }
}

// --------------------------------------------------------------------------------------- trailing
// A_IK -= L_Ij * L_Kj^T for jb < K <= I   (grid.x = I - jb - 1, grid.y = K - jb - 1)
template <typename T>
__global__ void __launch_bounds__(256, 2) potrf_trailing_kernel(int Npad, int jb, T* __restrict__ A) {
  using C = Cfg<T>;
  constexpr int NB = C::NB;
  if (blockIdx.y > blockIdx.x) return;
  __shared__ TileSmem<T> sm;
  const int s = blockIdx.z, I = jb + 1 + blockIdx.x, K = jb + 1 + blockIdx.y;
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
  TileGemm<T, Lay::KContig, Lay::KContig, true>::run(acc, Lij, Npad, Lkj, Npad, NB, sm);
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
  const size_t dsm = 2 * (size_t)NB * (NB + 1) * sizeof(T);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(potrf_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    attr_done = true;
  }
  cudaMemsetAsync(info, 0, sizeof(int) * S, st);
  for (int jb = 0; jb < nblk; ++jb) {
    potrf_diag_kernel<T><<<S, 256, dsm, st>>>(Npad, jb, A, winv, info);
    count_launch();
    const int rem = nblk - jb - 1;
    if (rem > 0) {
      potrf_panel_kernel<T><<<dim3(rem, 1, S), 256, 0, st>>>(Npad, jb, A, winv);
      potrf_trailing_kernel<T><<<dim3(rem, rem, S), 256, 0, st>>>(Npad, jb, A);
      count_launch(2);
    }
  }
  return check_launch("potrf_lower_batched");
}

template int potrf_lower_batched<float>(int, int, float*, float*, int*, cudaStream_t);
template int potrf_lower_batched<double>(int, int, double*, double*, int*, cudaStream_t);

}  // namespace smk
