// potrf_ll.cu -- float64 batched Cholesky for the slice-sampler log-likelihood (f2; reference: the spla.cholesky inside
// every logprob closure, OPT:637, 659, 690, ~50 D/8 calls per hyper-sample, thousands per next() at N = 4096).
//
// Round 1 ran this on the generic SIMT factorisation (NB = 64, 4 x 4 register tiles): 8.3 ms per N = 4096 matrix,
// 2.8 TFLOP/s.  This file is the dedicated path:
//   * NB = 128 block columns in pairs, right-looking, look-ahead on two streams: while the rank-256 update of the
//     trailing matrix (pair j) runs, the next pair's diagonal blocks and panels are already being factored, so the
//     serial spine  diag -> panel -> column update  hides behind the N^3/3 flops;
//   * diagonal blocks by the warp-synchronous single-SM kernel of diag.cuh, panels by its block substitution;
//   * the trailing updates (A_IK -= L_Ij L_Kj^T) as one tiled GEMM kernel on the fp64 tensor
//     path: mma.sync.m8n8k4.f64 (DMMA), operands staged by cp.async through a 3-stage shared-memory ring
//     (row stride 20 doubles: conflict-free 8-byte fragment loads), 128 x 128 output tiles (32 x 64 per warp) for the
//     update and 32-row tiles for the thin panel / look-ahead column so that even one matrix spreads over the SMs;
//   * the whole launch sequence of one batch is captured once into a CUDA graph per (pointer, shape) and replayed.
// Only L (lower triangle) is produced; W_jj (inverse diagonal blocks) is internal workspace.
#include <cuda_runtime.h>

#include <map>
#include <tuple>

#include "common.cuh"
#include "diag.cuh"

namespace smk {
namespace ll {

constexpr int NB = 128;
constexpr int KC = 16;         // k per stage
constexpr int LDS = 20;        // shared row stride in doubles (16 + 4): fragment loads hit 16 distinct 8-byte banks
constexpr int STAGES = 3;

constexpr int WD = NB * 32;     // compact diagonal inverses of one block: [4][32][32]

__global__ void __launch_bounds__(256) diag_kernel(int ld, int jb, double* __restrict__ A, long a_stride,
                                                    double* __restrict__ W, long w_stride, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int s = blockIdx.x;
  double* Ab = A + (long)s * a_stride + (long)jb * NB * ld + (long)jb * NB;
  diag_factor_block<double, NB>(Ab, ld, W + (long)s * w_stride + (long)jb * WD, info ? info + s : nullptr, jb * NB,
                                reinterpret_cast<double*>(smem_raw));
}

// L_Ij = A_Ij L_jj^-T by block substitution, 32 rows per CTA; grid = (rows below / 32, 1, S)
__global__ void __launch_bounds__(256) panel_kernel(int ld, int jb, double* __restrict__ A, long a_stride,
                                                     const double* __restrict__ W, long w_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int s = blockIdx.z;
  double* As = A + (long)s * a_stride;
  panel_sub_block<double, NB>(As + ((long)(jb + 1) * NB + (long)blockIdx.x * 32) * ld + (long)jb * NB, ld,
                              As + (long)jb * NB * ld + (long)jb * NB, W + (long)s * w_stride + (long)jb * WD, nullptr,
                              nullptr, reinterpret_cast<double*>(smem_raw));
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) { dmma_884(c, a, b); }

// C[r0 + 0..BM)[c0 + 0..128) (op)= A[r0 + ..][0..K) * B[c0 + ..][0..K)^T, all row-major (k contiguous).
//   sub = 1 : C -= A B^T (trailing / column update);  sub = 0 : C = A B^T (panel; C may alias A: each CTA reads its
//   whole A tile before it writes).  tri = 1: tiles strictly above the block diagonal are skipped (BM == 128 only).
struct GemmArgs {
  const double* A; const double* B; double* C;
  long a_stride, b_stride, c_stride;     // per matrix (blockIdx.z)
  int lda, ldb, ldc, K;
  int row0, col0;                        // first output row / column (elements) of tile (0, 0)
  int brow0;                             // first row of B for column tile 0
  int sub, tri;
};

template <int BM>
__global__ void __launch_bounds__(256) dgemm_nt_kernel(GemmArgs g) {
  constexpr int WM = BM / 32, WN = 8 / WM;          // warp grid
  constexpr int WTN = 128 / WN;                     // columns per warp: 64 / 32 / 16
  constexpr int NT = WTN / 8;                       // 8-column DMMA tiles per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);           // [STAGES][BM][LDS]
  double* Bs = As + STAGES * BM * LDS;                         // [STAGES][128][LDS]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp / WN, wn = warp % WN, gq = lane >> 2, t4 = lane & 3;
  const int r0 = g.row0 + blockIdx.x * BM, c0 = g.col0 + blockIdx.y * 128;
  if (g.tri && c0 > r0 + BM - 1) return;            // the whole tile lies above the diagonal
  const double* A = g.A + (long)blockIdx.z * g.a_stride + (long)r0 * g.lda;
  const double* B = g.B + (long)blockIdx.z * g.b_stride + (long)(g.brow0 + blockIdx.y * 128) * g.ldb;
  double* C = g.C + (long)blockIdx.z * g.c_stride + (long)r0 * g.ldc + c0;

  auto load_stage = [&](int st, int k0) {
    // 16-byte chunks: row = chunk / 8, piece = chunk % 8 (2 doubles each)
#pragma unroll
    for (int q = 0; q < BM * 8 / 256; ++q) {
      const int ch = tid + q * 256, row = ch >> 3, pc = ch & 7;
      cp_async16(As + ((size_t)st * BM + row) * LDS + pc * 2, A + (long)row * g.lda + k0 + pc * 2);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ch = tid + q * 256, row = ch >> 3, pc = ch & 7;
      cp_async16(Bs + ((size_t)st * 128 + row) * LDS + pc * 2, B + (long)row * g.ldb + k0 + pc * 2);
    }
  };

  double acc[4][NT][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }

  const int nk = g.K / KC;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s * KC);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < nk) load_stage((kt + STAGES - 1) % STAGES, (kt + STAGES - 1) * KC);
    cp_async_commit();
    const double* as = As + ((size_t)(kt % STAGES) * BM + wm * 32 + gq) * LDS + t4;
    const double* bs = Bs + ((size_t)(kt % STAGES) * 128 + wn * WTN + gq) * LDS + t4;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      double af[4], bf[NT];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) af[mi] = as[mi * 8 * LDS + kk];
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) bf[ni] = bs[ni * 8 * LDS + kk];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) dmma(acc[mi][ni], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();          // panel mode writes over its own A tile: every warp is done reading

#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      double2* cp = reinterpret_cast<double2*>(C + (long)(wm * 32 + mi * 8 + gq) * g.ldc + wn * WTN + ni * 8 + t4 * 2);
      double2 v;
      if (g.sub) { v = *cp; v.x -= acc[mi][ni][0]; v.y -= acc[mi][ni][1]; }
      else { v.x = acc[mi][ni][0]; v.y = acc[mi][ni][1]; }
      *cp = v;
    }
}

template <int BM>
static void launch_gemm(const GemmArgs& g, int tiles_m, int tiles_n, int S, cudaStream_t st) {
  const size_t smem = sizeof(double) * STAGES * (BM + 128) * LDS;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(dgemm_nt_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  dgemm_nt_kernel<BM><<<dim3(tiles_m, tiles_n, S), 256, smem, st>>>(g);
}

struct Streams {
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t panel[2] = {nullptr, nullptr}, rest[2] = {nullptr, nullptr}, join = nullptr;
};
static Streams& streams() {
  static Streams s;
  if (!s.main) {
    cudaStreamCreateWithFlags(&s.main, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking);
    for (int i = 0; i < 2; ++i) {
      cudaEventCreateWithFlags(&s.panel[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&s.rest[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming);
  }
  return s;
}

// The launch sequence on (main, side); called directly or under stream capture.
static int enqueue(int Npad, int S, double* A, double* W, int* info, Streams& ss) {
  const int nblk = Npad / NB;
  const long as = (long)Npad * Npad, ws = (long)nblk * WD;
  const size_t dsm = DiagSmem<double, NB>::bytes, psm = PanelSmem<double, NB>::bytes;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
    attr = true;
  }
  cudaStream_t m = ss.main, sd = ss.side;
  bool side_used = false;
  // Block columns in PAIRS (j, j+1).  Main stream (the spine): diag j, panel j, rank-128 update of column j+1, diag j+1,
  // panel j+1, then the look-ahead: the rank-256 update of the NEXT pair's two columns.  Side stream: the rank-256 update
  // of everything to the right of the next pair, overlapped with the next pair's spine.  Rank 256 halves the read-modify-
  // write traffic of the trailing matrix per flop.
  auto gemm_args = [&](int jcol, int K, int row_blk, int col_blk, int tri) {
    GemmArgs g{};
    g.A = A + (long)jcol * NB; g.lda = Npad; g.a_stride = as;
    g.B = A + (long)jcol * NB; g.ldb = Npad; g.b_stride = as; g.brow0 = col_blk * NB;
    g.C = A; g.ldc = Npad; g.c_stride = as;
    g.K = K; g.row0 = row_blk * NB; g.col0 = col_blk * NB; g.sub = 1; g.tri = tri;
    return g;
  };
  for (int j = 0, pair = 0; j < nblk; j += 2, ++pair) {
    diag_kernel<<<S, 256, dsm, m>>>(Npad, j, A, as, W, ws, info);
    count_launch();
    const int rem = nblk - j - 1;                // block rows below column j
    if (rem <= 0) break;
    panel_kernel<<<dim3(rem * 4, 1, S), 256, psm, m>>>(Npad, j, A, as, W, ws);           // L_Ij = A_Ij L_jj^-T
    launch_gemm<32>(gemm_args(j, NB, j + 1, j + 1, 0), rem * 4, 1, S, m);                 // column j+1 -= L_Ij L_(j+1)j^T
    diag_kernel<<<S, 256, dsm, m>>>(Npad, j + 1, A, as, W, ws, info);
    count_launch(3);
    const int rem2 = nblk - j - 2;               // block rows below column j+1
    if (rem2 <= 0) break;
    panel_kernel<<<dim3(rem2 * 4, 1, S), 256, psm, m>>>(Npad, j + 1, A, as, W, ws);
    cudaEventRecord(ss.panel[pair & 1], m);
    // look-ahead: the next pair's columns (j+2, j+3) with respect to panels j and j+1; the side stream's update of the
    // previous pair also wrote those columns, so wait for it first
    if (side_used) cudaStreamWaitEvent(m, ss.rest[(pair - 1) & 1], 0);
    launch_gemm<32>(gemm_args(j, 2 * NB, j + 2, j + 2, 1), rem2 * 4, rem2 >= 2 ? 2 : 1, S, m);
    count_launch(2);
    if (rem2 > 2) {                              // everything to the right of the next pair, on the side stream
      cudaStreamWaitEvent(sd, ss.panel[pair & 1], 0);
      launch_gemm<128>(gemm_args(j, 2 * NB, j + 4, j + 4, 1), rem2 - 2, rem2 - 2, S, sd);
      cudaEventRecord(ss.rest[pair & 1], sd);
      count_launch();
      side_used = true;
    }
  }
  if (side_used) {
    cudaEventRecord(ss.join, sd);
    cudaStreamWaitEvent(m, ss.join, 0);
  }
  return 0;
}

}  // namespace ll

static std::map<cudaGraphExec_t, int>& launches_per_graph() {
  static std::map<cudaGraphExec_t, int> m;
  return m;
}

size_t potrf_ll_workspace_bytes(int Npad, int S) { return (size_t)S * (Npad / ll::NB) * ll::WD * sizeof(double); }

// A: [S][Npad][Npad] (lower triangle in/out), W: workspace, info[S].  Npad % 128 == 0.
// use_graph: capture the sequence once per (A, W, info, Npad, S) on internal streams and replay it behind `st`.
int potrf_ll_f64(int Npad, int S, double* A, double* W, int* info, int use_graph, cudaStream_t st) {
  if (Npad <= 0 || Npad % ll::NB) return -1;
  if (S <= 0) return -2;
  if (!A || !W || !info) return -3;
  ll::Streams& ss = ll::streams();
  static cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  if (!ev_in) {
    cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming);
  }
  cudaMemsetAsync(info, 0, sizeof(int) * S, st);
  if (!use_graph) {
    cudaEventRecord(ev_in, st);
    cudaStreamWaitEvent(ss.main, ev_in, 0);
    ll::enqueue(Npad, S, A, W, info, ss);
    cudaEventRecord(ev_out, ss.main);
    cudaStreamWaitEvent(st, ev_out, 0);
    return check_launch("potrf_ll");
  }
  typedef std::tuple<double*, double*, int*, int, int> Key;
  static std::map<Key, cudaGraphExec_t> cache;
  const Key key(A, W, info, Npad, S);
  auto it = cache.find(key);
  if (it == cache.end()) {
    // first use: run it once un-captured (sets the kernel attributes outside capture), then capture
    cudaEventRecord(ev_in, st);
    cudaStreamWaitEvent(ss.main, ev_in, 0);
    const long long before = launch_count();
    ll::enqueue(Npad, S, A, W, info, ss);
    const int per_run = (int)(launch_count() - before);
    cudaEventRecord(ev_out, ss.main);
    cudaStreamWaitEvent(st, ev_out, 0);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (cudaStreamBeginCapture(ss.main, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      ll::enqueue(Npad, S, A, W, info, ss);
      count_launch(-per_run);                     // the captured enqueue launched nothing
      if (cudaStreamEndCapture(ss.main, &graph) == cudaSuccess && graph &&
          cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        if (cache.size() > 64) { for (auto& kv : cache) cudaGraphExecDestroy(kv.second); cache.clear(); }
        cache[key] = exec;
        launches_per_graph()[exec] = per_run;
      }
      if (graph) cudaGraphDestroy(graph);
    }
    cudaGetLastError();
    return check_launch("potrf_ll");
  }
  cudaGraphLaunch(it->second, st);
  count_launch(launches_per_graph()[it->second]);
  return check_launch("potrf_ll(graph)");
}

}  // namespace smk
