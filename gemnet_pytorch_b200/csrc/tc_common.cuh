// PTX helpers shared by the round-2 tcgen05 kernels (chain_tc.cu, segbilinear_bwd_tc.cu): mbarrier, bulk copy,
// TMEM allocation / load, UMMA issue, shared-memory matrix descriptors, exact bf16 splits.  sm_100a only.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace gemnet {
namespace tcx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// a wait that lasts longer than this is a protocol bug: trap instead of hanging the device.  (Wall time, not a spin
// count: under compute-sanitizer the other warps run orders of magnitude slower than a failing try_wait.)
constexpr uint64_t kWaitLimitNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (spin & 0xFFFu) == 0xFFFu) {
      const uint64_t t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > kWaitLimitNs) __trap();
    }
  }
}
// the same for a warp with nothing else to do (issuer / mover roles): back off between polls so that the polling
// instructions do not take issue slots from the warps that work on the same scheduler
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) {
      __nanosleep(40);
      if ((spin & 0xFFFu) == 0xFFFu) {
        const uint64_t t = globaltimer_ns();
        if (t0 == 0) t0 = t;
        else if (t - t0 > kWaitLimitNs) __trap();
      }
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 x bf16 -> fp32), single CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (elect.sync): tcgen05.mma is issued by it with WARP-UNIFORM operands — then the
// instruction takes its descriptors from uniform registers directly; issued under a thread-variant predicate
// (`if (tid == 0)`) ptxas wraps every MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~66 cycles per instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// the same with the A operand in tensor memory (M lanes x K/2 32-bit columns, two consecutive K elements per column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B operand tile (rows of 128 bytes, 8-row groups of 1024 bytes): shared-memory matrix descriptor
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);   // start address
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), 16 B units
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128 (or 64), N multiple of 16 (8 for M = 64)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 16 consecutive fp32 columns of this thread's TMEM lane (32x32b shape: lane i of the warp <-> TMEM lane base + i)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// registers -> 16 consecutive fp32 columns of this thread's TMEM lane (complete when it returns)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st4_nowait(uint32_t taddr, const uint4& r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r.x), "r"(r.y), "r"(r.z),
               "r"(r.w)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <int NV>
__device__ __forceinline__ void tmem_ld_nowait(uint32_t taddr, uint32_t (&r)[NV]) {
  static_assert(NV == 16 || NV == 8 || NV == 4, "tcgen05.ld shape");
  if constexpr (NV == 16) tmem_ld16_nowait(taddr, r);
  else if constexpr (NV == 8) tmem_ld8_nowait(taddr, r);
  else tmem_ld4_nowait(taddr, r);
}
template <int NV>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[NV]) {
  if constexpr (NV == 16) tmem_st16(taddr, r);
  else if constexpr (NV == 8) tmem_st8(taddr, r);
  else tmem_st4(taddr, r);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// exact three-way bf16 split of two fp32 values: v = h + m + l (each packed as bf16x2, low half = first value)
__device__ __forceinline__ void split3(float v0, float v1, uint32_t& h, uint32_t& m, uint32_t& l) {
  const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
  h = *reinterpret_cast<const uint32_t*>(&hh);
  const float r0 = v0 - __uint_as_float(h << 16), r1 = v1 - __uint_as_float(h & 0xFFFF0000u);
  const __nv_bfloat162 mm = __floats2bfloat162_rn(r0, r1);
  m = *reinterpret_cast<const uint32_t*>(&mm);
  const float q0 = r0 - __uint_as_float(m << 16), q1 = r1 - __uint_as_float(m & 0xFFFF0000u);
  const __nv_bfloat162 ll = __floats2bfloat162_rn(q0, q1);
  l = *reinterpret_cast<const uint32_t*>(&ll);
}
// the inverse (exact): two fp32 values from their three packed bf16x2 terms
__device__ __forceinline__ void join3(uint32_t h, uint32_t m, uint32_t l, float& v0, float& v1) {
  v0 = (__uint_as_float(h << 16) + __uint_as_float(m << 16)) + __uint_as_float(l << 16);
  v1 = (__uint_as_float(h & 0xFFFF0000u) + __uint_as_float(m & 0xFFFF0000u)) + __uint_as_float(l & 0xFFFF0000u);
}

constexpr float kSiluScale = 1.0f / 0.6f;
constexpr float kShrinkPerStep = 1.25e-8f;  // mean relative toward-zero rounding loss of one tcgen05 accumulate

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// k * silu(z) and k * d/dz silu(z) (k = 1/0.6 times the layer's output scale, folded by the caller); sigmoid and
// 1 - sigmoid from exp(-|z|): no overflow / cancellation; ex2.approx and rcp.approx (2 ulp), no slow paths
// exp(-|z|) as ONE multiply + ex2.approx.ftz (|z| * log2 e in (−inf, 0]: no range handling needed)
__device__ __forceinline__ float exp_neg_abs(float z) {
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(z) * -1.4426950408889634f));
  return t;
}
__device__ __forceinline__ void scaled_silu(float z, float k, float& v, float& d) {
  const float t = exp_neg_abs(z);
  const float rr = rcp_approx(1.0f + t);      // 1 + t in (1, 2]
  const float p = t * rr;
  const float sg = z >= 0.f ? rr : p, oms = z >= 0.f ? p : rr;
  const float sk = sg * k;
  v = z * sk;
  d = sk * fmaf(z, oms, 1.0f);
}
__device__ __forceinline__ float scaled_silu_value(float z, float k) {
  const float t = exp_neg_abs(z);
  const float rr = rcp_approx(1.0f + t);
  return z * ((z >= 0.f ? rr : t * rr) * k);
}

}  // namespace tcx
}  // namespace gemnet
