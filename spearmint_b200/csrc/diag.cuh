// diag.cuh -- factor one NB x NB diagonal block of the blocked Cholesky on ONE SM and invert its factor.
//
// This kernel is the serial spine of every factorisation in the library (reference: spla.cholesky at OPT:540, 567, 585
// and inside every slice-sampler log-probability OPT:637, 659, 690): Npad / NB launches per matrix, strictly one after
// the other, so what matters is its latency.  Round 1 used block-wide barriers around every column of every 32 x 32
// piece (155 us for a 128 x 128 float block).  Here:
//   * the block lives in shared memory as 32 x 32 tiles (row stride 33: conflict-free by row and by column);
//   * a 32 x 32 diagonal piece is factored by ONE warp, warp-synchronously: lane = row, the row sits in registers, a
//     column step is one shuffle (pivot), one rsqrt, a 32-entry broadcast buffer and 31 - j fused multiply-adds --
//     no block barrier inside the piece;
//   * rows below the piece are solved by substitution with one thread per row (registers), not via the inverse;
//   * the rank-32 update of what is left of the block is register-tiled 4 x 4 over all 256 threads;
//   * W = L^-1 is assembled at the end: the four 32 x 32 diagonal inverses concurrently (one warp each, lane = column,
//     column in registers), then the off-diagonal tiles by block distance.
#pragma once
#include "common.cuh"

namespace smk {

template <typename T> __device__ __forceinline__ T smk_rsqrt(T x);
template <> __device__ __forceinline__ float smk_rsqrt<float>(float x) { return 1.0f / sqrtf(x); }
template <> __device__ __forceinline__ double smk_rsqrt<double>(double x) { return rsqrt(x); }

template <typename T, int NB>
struct DiagSmem {
  static constexpr int SB = 32, LDT = 33, NP = NB / SB, NT = NP * (NP + 1) / 2;
  static constexpr int TILE = SB * LDT;
  // a tiles | w tiles | broadcast column [32] | inverse pivots [NB]
  static constexpr size_t bytes = sizeof(T) * ((size_t)2 * NT * TILE + 32 + NB);
  static __device__ __forceinline__ int tile(int bi, int bj) { return (bi * (bi + 1) / 2 + bj) * TILE; }
};

// Lower Cholesky of the 32 x 32 piece at `a` (tile, row stride 33) by the calling warp; lane = row.
// `col` is a 32-entry broadcast buffer, ipv receives 1 / L_jj.  Returns the first bad pivot (or -1) in `bad`.
template <typename T>
__device__ __forceinline__ void warp_chol32(T* a, T* col, T* ipv, int lane, int& bad) {
  constexpr int LDT = 33;
  T r[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) r[k] = (k <= lane) ? a[lane * LDT + k] : T(0);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    T d = __shfl_sync(0xffffffffu, r[j], j);
    if (!(d > T(0))) { if (bad < 0) bad = j; d = T(1); }
    const T ip = smk_rsqrt<T>(d);
    const T l = r[j] * ip;                       // lane j: d / sqrt(d) = sqrt(d)
    if (lane >= j) r[j] = l;
    if (lane == j) ipv[j] = ip;
    col[lane] = l;
    __syncwarp();
#pragma unroll
    for (int k = j + 1; k < 32; ++k) r[k] = fma(-l, col[k], r[k]);   // meaningful for j < k <= lane
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k <= lane) a[lane * LDT + k] = r[k];
}

// W = L^-1 of a 32 x 32 lower-triangular tile by the calling warp; lane = column c, the column lives in registers.
template <typename T>
__device__ __forceinline__ void warp_trinv32(const T* l, const T* ipv, T* w, int lane) {
  constexpr int LDT = 33;
  T x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    T t0 = (i == lane) ? T(1) : T(0), t1 = T(0), t2 = T(0), t3 = T(0);
#pragma unroll
    for (int k = 0; k < i; ++k) {                // x[k] == 0 for k < lane: no lane-dependent bounds
      const T lv = l[i * LDT + k];
      if ((k & 3) == 0) t0 = fma(-lv, x[k], t0);
      else if ((k & 3) == 1) t1 = fma(-lv, x[k], t1);
      else if ((k & 3) == 2) t2 = fma(-lv, x[k], t2);
      else t3 = fma(-lv, x[k], t3);
    }
    x[i] = (i >= lane) ? ((t0 + t1) + (t2 + t3)) * ipv[i] : T(0);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) w[i * LDT + lane] = x[i];
}

// A: [..][ld] matrix of this sample, block (jb, jb) is factored in place (lower triangle; the strict upper triangle of
// the block is left untouched); Wb: NB x NB row-major, receives L_jj^-1 (zeros above the diagonal).
// info (may be NULL): 1-based index of the first non-positive pivot, written once.
template <typename T, int NB>
__device__ void diag_factor_block(T* __restrict__ Ab, int ld, T* __restrict__ Wb, int* info, int info_base, T* sm) {
  using DS = DiagSmem<T, NB>;
  constexpr int SB = 32, LDT = 33, NP = DS::NP, TILE = DS::TILE;
  T* a = sm;
  T* w = a + DS::NT * TILE;
  T* col = w + DS::NT * TILE;
  T* ipv = col + 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- load the lower triangle as tiles (rows of 32 contiguous elements per warp: coalesced)
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e / NB, k = e % NB;
    if ((k >> 5) <= (i >> 5)) a[DS::tile(i >> 5, k >> 5) + (i & 31) * LDT + (k & 31)] = (k <= i) ? Ab[(long)i * ld + k] : T(0);
  }
  __syncthreads();

  for (int p = 0; p < NP; ++p) {
    // (a) diagonal piece
    if (warp == 0) {
      int bad = -1;
      warp_chol32<T>(a + DS::tile(p, p), col, ipv + p * SB, lane, bad);
      if (bad >= 0 && lane == 0 && info && *info == 0) *info = info_base + p * SB + bad + 1;
    }
    __syncthreads();
    const int rows = NB - (p + 1) * SB;
    if (rows == 0) break;
    // (b) rows below: X L_pp^T = A_sub by substitution, one thread per row (x_k needs x_0..x_{k-1}: registers)
    if (tid < rows) {
      const int bi = p + 1 + (tid >> 5);
      T* ar = a + DS::tile(bi, p) + (tid & 31) * LDT;
      const T* lp = a + DS::tile(p, p);
      const T* ip = ipv + p * SB;
      T x[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = ar[k];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        T t0 = x[k], t1 = T(0), t2 = T(0), t3 = T(0);
#pragma unroll
        for (int m = 0; m < k; ++m) {
          const T lv = lp[k * LDT + m];          // same address in every lane: broadcast
          if ((m & 3) == 0) t0 = fma(-lv, x[m], t0);
          else if ((m & 3) == 1) t1 = fma(-lv, x[m], t1);
          else if ((m & 3) == 2) t2 = fma(-lv, x[m], t2);
          else t3 = fma(-lv, x[m], t3);
        }
        x[k] = ((t0 + t1) + (t2 + t3)) * ip[k];
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) ar[k] = x[k];
    }
    __syncthreads();
    // (c) rank-32 update of the remaining lower triangle: C[r][c] -= X[r] . X[c] for r >= c.  A thread owns the 4 x 4
    //     elements (gr + i nt, gc + j nt): consecutive lanes read consecutive rows of X (conflict-free, stride 33) and
    //     share the other operand (broadcast).  i > j is always below the diagonal, i == j iff gr >= gc, i < j never.
    const int nt = rows / 4;
    for (int e = tid; e < nt * nt; e += 256) {
      const int gr = e / nt, gc = e % nt;
      const bool dg = gr >= gc;
      const T* xr[4];
      const T* xc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = gr + i * nt, c = gc + i * nt;
        xr[i] = a + DS::tile(p + 1 + (r >> 5), p) + (r & 31) * LDT;
        xc[i] = a + DS::tile(p + 1 + (c >> 5), p) + (c & 31) * LDT;
      }
      T acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
#pragma unroll 8
      for (int k = 0; k < SB; ++k) {
        T xa[4], xb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { xa[i] = xr[i][k]; xb[i] = xc[i][k]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) acc[i][j] = fma(xa[i], xb[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          if (i == j && !dg) continue;
          const int r = gr + i * nt, c = gc + j * nt;
          a[DS::tile(p + 1 + (r >> 5), p + 1 + (c >> 5)) + (r & 31) * LDT + (c & 31)] -= acc[i][j];
        }
    }
    __syncthreads();
  }

  // ---- W = L^-1: diagonal tiles (one warp each), then off-diagonal tiles by block distance
  if (warp < NP) warp_trinv32<T>(a + DS::tile(warp, warp), ipv + warp * SB, w + DS::tile(warp, warp), lane);
  __syncthreads();
  for (int d = 1; d < NP; ++d) {
    // W_ij = -W_ii * (sum_{k=j}^{i-1} L_ik W_kj), i = j + d.  T_ij goes through the (free) strictly-lower tile of w.
    const int npair = NP - d;
    for (int e = tid; e < npair * SB * SB; e += 256) {
      const int pr = e / (SB * SB), rr = (e / SB) % SB, cc = e % SB;
      const int j = pr, i = pr + d;
      T acc = T(0);
      for (int kb = j; kb < i; ++kb) {
        const T* lrow = a + DS::tile(i, kb) + rr * LDT;
        const T* wk = w + DS::tile(kb, j) + cc;
#pragma unroll 8
        for (int m = 0; m < SB; ++m) acc = fma(lrow[m], wk[m * LDT], acc);
      }
      w[DS::tile(i, j) + rr * LDT + cc] = acc;
    }
    __syncthreads();
    // in-place multiply by -W_ii: column cc of T_ij is read fully before it is overwritten (thread = (pair, column))
    for (int e = tid; e < npair * SB; e += 256) {
      const int pr = e / SB, cc = e % SB;
      const int j = pr, i = pr + d;
      T* tcol = w + DS::tile(i, j) + cc;
      const T* wii = w + DS::tile(i, i);
      T tv[32];
#pragma unroll
      for (int m = 0; m < 32; ++m) tv[m] = tcol[m * LDT];
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        T acc = T(0);
#pragma unroll
        for (int m = 0; m <= rr; ++m) acc = fma(wii[rr * LDT + m], tv[m], acc);
        tcol[rr * LDT] = -acc;
      }
    }
    __syncthreads();
  }

  // ---- store L (lower triangle only) and W (full block, zeros above the diagonal)
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e / NB, k = e % NB;
    const bool low = (k >> 5) <= (i >> 5);
    const int o = low ? DS::tile(i >> 5, k >> 5) + (i & 31) * LDT + (k & 31) : 0;
    if (k <= i) Ab[(long)i * ld + k] = a[o];
    Wb[e] = (k <= i) ? w[o] : T(0);
  }
}

}  // namespace smk
