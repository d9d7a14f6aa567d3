// tc_common.cuh -- device helpers shared by the tcgen05 kernels (predict_tc.cu, kxt_tc.cu): fp16 operand scaling /
// splitting, fast stationary kernels on packed float32, and thin wrappers over the mbarrier / TMA / tcgen05 PTX.
#pragma once
#include <cuda_fp16.h>
#include <cuda.h>

#include <cstdint>

namespace smk {

// ---------------------------------------------------------------------------------------------- fp16 operand packing
// The predict GEMMs run as 3 x FP16 tensor products (hi*hi + hi*lo + lo*hi, fp32 accumulate).  fp16 carries the same
// 11-bit significand as tf32 at twice the tensor rate but only 5 exponent bits, so every operand matrix is multiplied
// by an exact power of two that puts its largest |entry| in [2^14, 2^15) before the round-to-nearest split
//     hi = fp16(x * 2^e),   lo = fp16(x * 2^e - hi)
// (representation error <= 2^-23 |x| for entries within 2^-18 of the maximum, <= 2^-40 max|x| absolute below that;
// tools/fp16_split_experiment.py, profiles/r01_fp16_split_experiment.md).  The epilogue multiplies the accumulator by
// 2^-(ea + eb); the scaling is exact, so it changes nothing but the representable range.
__host__ __device__ __forceinline__ int scale_exp(float amax) {
  int e;
  frexpf(amax * 1.00001f, &e);          // amax * 1.00001 < 2^e
  return 15 - e;
}
__device__ __forceinline__ int kx_exp(float amp2) { return scale_exp(amp2 * 1.000001f); }   // cross-covariance <= amp2 (1 + 1e-6)
__device__ __forceinline__ void split16(float x, __half& h, __half& l) {
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}
__device__ __forceinline__ uint2 pack4(const __half (&v)[4]) {
  __half2 a = __halves2half2(v[0], v[1]), b = __halves2half2(v[2], v[3]);
  uint2 r;
  r.x = *reinterpret_cast<unsigned*>(&a);
  r.y = *reinterpret_cast<unsigned*>(&b);
  return r;
}

// Fast stationary kernels for the generator (float32): sqrt.approx / ex2.approx (MUFU, ~1-2 ulp) instead of the accurate
// sqrtf/expf sequences -- their relative error (~2e-7) is at the level of the float32 rounding already carried by r2.
// The kernel is issue-bound (ncu r01: 110 instructions per element, 73% issue-slot utilisation, FMA pipe 53%), so the
// arithmetic is written with Blackwell's packed float32 instructions (FADD2 / FFMA2 / FMUL2: two lanes per issue slot).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 dup2(float x) { return make_float2(x, x); }
template <int KIND>
__device__ __forceinline__ float2 kernel_pair_fast_t(float2 r2) {      // same arithmetic, kind resolved at compile time
  constexpr float kL2E = 1.4426950408889634f;
  if (KIND <= 1) {
    const float2 t = __fmul2_rn(r2, dup2(-0.5f * kL2E));
    return make_float2(ex2_approx(t.x), ex2_approx(t.y));
  }
  const float2 r = make_float2(sqrt_approx(r2.x), sqrt_approx(r2.y));
  constexpr float c = (KIND == 2) ? 1.7320508075688772f : 2.23606797749979f;
  const float2 t = __fmul2_rn(r, dup2(-c * kL2E));
  const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
  float2 p = __ffma2_rn(r, dup2(c), dup2(1.f));
  if (KIND == 3) p = __ffma2_rn(r2, dup2(5.0f / 3.0f), p);
  return __fmul2_rn(p, e);
}
__device__ __forceinline__ float2 kernel_pair_fast(int kind, float2 r2) {
  constexpr float kL2E = 1.4426950408889634f;
  if (kind <= 1) {                                            // SE / ARDSE: exp(-r2 / 2)
    const float2 t = __fmul2_rn(r2, dup2(-0.5f * kL2E));
    return make_float2(ex2_approx(t.x), ex2_approx(t.y));
  }
  const float2 r = make_float2(sqrt_approx(r2.x), sqrt_approx(r2.y));
  const float c = (kind == 2) ? 1.7320508075688772f : 2.23606797749979f;
  const float2 t = __fmul2_rn(r, dup2(-c * kL2E));
  const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
  float2 p = __ffma2_rn(r, dup2(c), dup2(1.f));               // Matern32: (1 + sqrt3 r) e^-sqrt3 r
  if (kind == 3) p = __ffma2_rn(r2, dup2(5.0f / 3.0f), p);    // Matern52: (1 + sqrt5 r + 5/3 r2) e^-sqrt5 r
  return __fmul2_rn(p, e);
}

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Same wait for threads that expect to wait LONG (a producer on a free stage, an epilogue warp on its next accumulator):
// poll, then SLEEP for `ns` before the next poll, so the waiting warp stays out of the issue slots (and the power budget) of
// the warps that are doing the work.  ncu of the generator (r02): with a bare try_wait loop SYNCS + YIELD + BRA were ~20 % of
// all executed instructions; with try_wait's suspend-time hint the compiler's loop (SYNCS, NANOSLEEP, BRA) still ran 4.8e8
// iterations per launch = 29 % of the executed instructions -- the hint does not keep the thread asleep.  An explicit sleep
// of a few hundred nanoseconds costs at most that much latency per tile and removes the polling.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns = 256) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                            uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

}  // namespace tc
}  // namespace smk
