// N1: fused dense layer on the 5th-generation tensor cores — fp32 in / fp32 out GEMM with the surrounding pointwise
// work fused in:
//
//     A      = x              (M x K)      or   x .* xmul          (prologue Hadamard: the activation backward)
//     acc    = A @ B^T,   B = (N x K) weight operand  (pre-packed bf16 splits, K-major SWIZZLE_128B, resident in smem)
//     y      = s_acc * f(acc) + s_res * res            f = identity | ScaledSiLU  (silu(z)/0.6, base_layers.py:51-58)
//     dact   = s_acc * f'(acc)                         (optional; what the backward multiplies the gradient with)
//
// which covers Dense (base_layers.py:45-48), the tail of ResidualLayer ((x + act(z))/sqrt(2), :86-89) and their
// input-gradients ((g .* dact) @ W [+ s*g]).  The reference runs these as cuBLAS SIMT-fp32 GEMMs + 2-4 pointwise kernels.
//
// Numerics (measured on B200, profiles/time_dense.py): tcgen05 TRUNCATES when it adds a product into the fp32 TMEM
// accumulator — a mean signed error of -1.6e-8 per accumulate.  A naive error-compensated GEMM that chains all its
// MMAs into one accumulator is therefore BIASED (-1e-6 for K = 128), and the bias compounds over GemNet's ~40 dense
// layers to ~1e-4 on energies/forces.  This kernel keeps fp32-grade, unbiased results by
//   * splitting every fp32 operand exactly into three bf16 terms  x = h + m + l  (8 + 8 + 8 mantissa bits),
//   * sending the dominant h*h products of each 64-wide K chunk to their OWN accumulator (<= 4 accumulates each),
//   * sending the five correction products (h*m, m*h, m*m, h*l, l*h; the dropped ones are < 2^-24) to a separate
//     accumulator, where truncation is relative to a 2^-8 smaller magnitude,
//   * and adding the accumulators in the epilogue with round-to-nearest fp32 adds.
//
// Persistent warp-specialised kernel, one CTA per SM, 128-row tiles:
//   warps 0-7   producers: coalesced 128-bit loads of the fp32 rows, three-way bf16 split, written into the K-major
//               SWIZZLE_128B stage (64 k per chunk = one 128-byte row per operand row and split)
//   warp 8      one thread issues tcgen05.mma.kind::f16 (M128, N, K16) into the TMEM accumulators
//   warp 9      one thread loads the whole packed weight image once per CTA with cp.async.bulk
//   warps 10-17 epilogue: tcgen05.ld, accumulator sum, activation / scale / residual, 128-bit stores
// Launch-bound sizes (every CTA has at most one tile; N = 128 is then also split over two CTAs): the roles cooperate
// instead of waiting for each other — warps 10-17 produce K chunk 1 while warps 0-7 produce chunk 0, and warps 0-7
// take half of the epilogue columns.  The kernel is a programmatic dependent launch: barrier setup, TMEM allocation
// and the weight-image fetch run while the preceding kernel drains (griddepcontrol.wait guards every activation).
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"

namespace gemnet {
namespace dtc {

constexpr int TILE_M = 128;
constexpr int KB = 64;                          // k per chunk (bf16: 64 x 2 B = one 128-byte swizzle row)
constexpr int A_SPLIT_BYTES = TILE_M * KB * 2;  // 16384: one of the three bf16 splits
constexpr int STAGE_BYTES = 3 * A_SPLIT_BYTES;  // 49152
constexpr int kStages = 2;
constexpr int PRODUCER_WARPS = 8, PRODUCER_THREADS = PRODUCER_WARPS * 32;
constexpr int EPI_WARP0 = PRODUCER_WARPS + 2, EPI_WARPS = 8;
constexpr int THREADS = (PRODUCER_WARPS + 2 + EPI_WARPS) * 32;  // 576

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 26)) __trap();
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
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (elect.sync).  tcgen05.mma / commit issued by it with warp-uniform operands take their
// descriptors from uniform registers directly; under a thread-variant predicate (`if (lane == 0)`) ptxas wraps every
// instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~66 cycles per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, 8-row groups of 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int CW>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[CW]) {
  if constexpr (CW == 32) tmem_ld32(taddr, r);
  else tmem_ld16(taddr, r);
}

// exact three-way bf16 split of two fp32 values: v = h + m + l (each packed as bf16x2, .x = first value)
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
constexpr float kSiluScale = 1.0f / 0.6f;
constexpr float kShrinkPerStep = 1.25e-8f;  // mean relative toward-zero rounding loss of one tcgen05 accumulate

// ---------------------------------------------------------------------------------------------- weight packing
// image: [kb][split h|m|l][n][64 bf16, 16-byte chunks XOR-swizzled by (n & 7)];  element (n, k) = src[n*sn + k*sk]
// Second image, behind the shared-memory one: the layout chain_tc.cu copies into tensor memory (its A operand).
// u32 [split][K / 8 column groups][128 features][4 columns]; column c of a split holds the bf16 pair (k = 2c, 2c + 1).
// Always 128 feature slots per group (rows >= N of a 64-row image stay unwritten: those TMEM lanes are never read).
__device__ __forceinline__ void tmem_image_store(__nv_bfloat16* img, int Kp, int n, int k, __nv_bfloat16 h,
                                                 __nv_bfloat16 m, __nv_bfloat16 l) {
  const int groups = Kp >> 3;
  const int64_t u = ((int64_t)(k >> 3) * 128 + n) * 4 + ((k >> 1) & 3);
  const int64_t split = (int64_t)groups * 128 * 4;
  img[2 * u + (k & 1)] = h;
  img[2 * (u + split) + (k & 1)] = m;
  img[2 * (u + 2 * split) + (k & 1)] = l;
}

__global__ void pack_weight_kernel(const float* __restrict__ src, int64_t sn, int64_t sk, int N_src, int K_src, int N,
                                   int K, __nv_bfloat16* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * K) return;
  const int k = idx % K, n = idx / K;
  const float w = (n < N_src && k < K_src) ? src[(int64_t)n * sn + (int64_t)k * sk] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(w);
  const float r = w - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(r);
  const __nv_bfloat16 l = __float2bfloat16_rn(r - __bfloat162float(m));
  const int kb = k / KB, kl = k % KB;
  const int phys = n * KB + (((kl >> 3) ^ (n & 7)) << 3) + (kl & 7);
  dst[(int64_t)(kb * 3 + 0) * N * KB + phys] = h;
  dst[(int64_t)(kb * 3 + 1) * N * KB + phys] = m;
  dst[(int64_t)(kb * 3 + 2) * N * KB + phys] = l;
  tmem_image_store(dst + (int64_t)(K / KB) * 3 * N * KB, K, n, k, h, m, l);
}

// many weights in one launch (training: every image changes every optimiser step): blockIdx.y = table entry
__global__ void pack_many_kernel(const gemnet_pack_desc* __restrict__ table) {
  const gemnet_pack_desc d = table[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d.Np * d.Kp) return;
  const int k = idx % d.Kp, n = idx / d.Kp;
  const float w = (n < d.N && k < d.K) ? d.src[(int64_t)n * d.stride_n + (int64_t)k * d.stride_k] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(w);
  const float r = w - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(r);
  const __nv_bfloat16 l = __float2bfloat16_rn(r - __bfloat162float(m));
  const int kb = k / KB, kl = k % KB;
  const int phys = n * KB + (((kl >> 3) ^ (n & 7)) << 3) + (kl & 7);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(d.dst);
  dst[(int64_t)(kb * 3 + 0) * d.Np * KB + phys] = h;
  dst[(int64_t)(kb * 3 + 1) * d.Np * KB + phys] = m;
  dst[(int64_t)(kb * 3 + 2) * d.Np * KB + phys] = l;
  tmem_image_store(dst + (int64_t)(d.Kp / KB) * 3 * d.Np * KB, d.Kp, n, k, h, m, l);
}

static inline int64_t ceil_div_i(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct Params {
  const float* x;      // (M, ldx) rows, K_src valid columns
  const float* xmul;   // optional (M, ldx) prologue multiplier
  const void* Wp;      // packed image: NKB * 3 * N * 64 bf16
  const float* res;    // optional (M, N_out) residual
  float* y;            // (M, N_out)
  float* dact;         // optional (M, N_out)
  int64_t M;
  int ldx, K_src, N_out;
  int act;             // 0 identity, 1 ScaledSiLU
  float s_acc, s_res;
  int N_img;           // rows per (kb, split) block of the packed image (64 or 128)
  int coop;            // allow the epilogue warps to produce the second K chunk of single-tile CTAs
};

template <int NKB, int N>
__global__ void __launch_bounds__(THREADS, 1) dense_tc_kernel(const Params p, int n_tiles) {
  constexpr int W_BYTES = NKB * 3 * N * KB * 2;
  constexpr int NACC = NKB + 1;                        // one main accumulator per K chunk + the correction accumulator
  constexpr uint32_t TMEM_COLS = NACC * N <= 128 ? 128 : (NACC * N <= 256 ? 256 : 512);
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_base = base;                       // resident weight image
  const uint32_t stage_base = base + W_BYTES;         // kStages * STAGE_BYTES
  const uint32_t bar_base = stage_base + kStages * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * kStages);
  const uint32_t acc_empty = bar_base + 8u * (2 * kStages + 1);
  const uint32_t w_bar = bar_base + 8u * (2 * kStages + 2);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 3);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(sm + W_BYTES + kStages * STAGE_BYTES + 8 * (2 * kStages + 3));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // (as a warp-uniform value: see elect_one)
  const int n_off = blockIdx.y * N;  // column slab of this CTA (small-M launches split N = 128 over two CTAs)
  // Programmatic dependent launch: let the NEXT kernel of the stream start its prologue (barriers, TMEM, weight
  // image) while this one runs.  Everything here that touches activations waits in griddep_wait() for the
  // PREVIOUS kernel to have completed and flushed; the weight image is older than that (see the flags argument of
  // gemnet_dense_tc) and is fetched before the wait.  Both instructions are no-ops in a plain launch.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), PRODUCER_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EPI_WARPS);
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == PRODUCER_WARPS) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  // One K chunk (kb) of one row tile: 128 rows x 64 columns of x (.* xmul) -> exact bf16 triple -> operand stage.
  // `ptid` = index of the thread in its 256-thread producer group.
  auto produce = [&](int64_t r0, int kb, uint32_t g, int ptid) {
    const int stage = g % kStages;
    const uint32_t parity = (g / kStages) & 1u;
    // 128 rows x 8 pieces of 8 elements per chunk = 1024 pieces, 4 per thread; 8 consecutive threads = one row
    float4 v[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int piece = ptid + q * PRODUCER_THREADS;
      const int row = piece >> 3, j = piece & 7;
      const int64_t r = r0 + row;
      const int k = kb * KB + j * 8;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        v[q][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < p.M && k + 4 * u < p.K_src) {
          v[q][u] = __ldg(reinterpret_cast<const float4*>(p.x + r * p.ldx + k + 4 * u));
          if (p.xmul) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(p.xmul + r * p.ldx + k + 4 * u));
            v[q][u] = make_float4(v[q][u].x * m.x, v[q][u].y * m.y, v[q][u].z * m.z, v[q][u].w * m.w);
          }
        }
      }
    }
    mbar_wait(empty_bar(stage), parity ^ 1u);
    uint8_t* st = sm + W_BYTES + stage * STAGE_BYTES;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int piece = ptid + q * PRODUCER_THREADS;
      const int row = piece >> 3, j = piece & 7;
      const uint32_t off = row * 128 + ((j ^ (row & 7)) << 4);
      uint32_t h[4], m[4], l[4];
      split3(v[q][0].x, v[q][0].y, h[0], m[0], l[0]);
      split3(v[q][0].z, v[q][0].w, h[1], m[1], l[1]);
      split3(v[q][1].x, v[q][1].y, h[2], m[2], l[2]);
      split3(v[q][1].z, v[q][1].w, h[3], m[3], l[3]);
      *reinterpret_cast<uint4*>(st + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(st + A_SPLIT_BYTES + off) = make_uint4(m[0], m[1], m[2], m[3]);
      *reinterpret_cast<uint4*>(st + 2 * A_SPLIT_BYTES + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(full_bar(stage));
  };
  // Launch-bound sizes give every CTA at most ONE tile: the epilogue warps would idle until the accumulator is
  // complete, so with two K chunks they produce chunk 1 while the producer warps produce chunk 0 — both chunks'
  // loads are in flight together and the operand conversion takes half as long.
  const bool coop = NKB == 2 && p.coop && n_tiles <= (int)gridDim.x;

  // Epilogue of columns [c_begin, c_end) of this CTA's slab for the 32 rows of the warp's TMEM lane quadrant, CW columns
  // at a time: accumulators -> compensated sum -> activation / residual -> y (and dact).
  auto epilogue_cols = [&](auto cw_tag, int tile, int c_begin, int c_end) {
    constexpr int CW = decltype(cw_tag)::value;
    const int ew = warp & 3;                          // a warp can only touch TMEM lanes [32*(warp%4), +32)
    const int64_t r = (int64_t)tile * TILE_M + ew * 32 + lane;
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += CW) {
      uint32_t acc[CW], t2[CW];
      const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16) + c0;
      // Rounding-bias compensation: every tcgen05 accumulate rounds its fp32 result TOWARD ZERO, which shrinks a
      // chain of n non-trivial K16 steps by a mean factor (1 - 1.25e-8 n) (measured: profiles/diag_shrink.py).
      // The mean loss c*main is below half an ulp of main, so it is added to the (2^-8 smaller, finer-grained)
      // correction sum first; the final main + correction add then rounds once, to nearest.  What remains is
      // zero-mean (rms 7e-8; cuBLAS SIMT fp32: 2e-7).
      const float c0f = kShrinkPerStep * (float)min(4, (p.K_src + 15) / 16);
      tmem_ld_cols<CW>(lane_base, acc);                  // main accumulator of chunk 0
      tmem_ld_cols<CW>(lane_base + NKB * N, t2);         // correction accumulator
#pragma unroll
      for (int v = 0; v < CW; ++v) t2[v] = __float_as_uint(fmaf(__uint_as_float(acc[v]), c0f, __uint_as_float(t2[v])));
      if (NKB == 2) {
        uint32_t t3[CW];
        const float c1f = kShrinkPerStep * (float)max(0, min(4, (p.K_src - KB + 15) / 16));
        tmem_ld_cols<CW>(lane_base + N, t3);             // main accumulator of chunk 1
#pragma unroll
        for (int v = 0; v < CW; ++v) {
          t2[v] = __float_as_uint(fmaf(__uint_as_float(t3[v]), c1f, __uint_as_float(t2[v])));
          acc[v] = __float_as_uint(__uint_as_float(acc[v]) + __uint_as_float(t3[v]));
        }
      }
#pragma unroll
      for (int v = 0; v < CW; ++v) acc[v] = __float_as_uint(__uint_as_float(acc[v]) + __uint_as_float(t2[v]));
      const int col0 = n_off + c0;
      if (r < p.M && col0 < p.N_out) {
        float* yrow = p.y + r * p.N_out + col0;
        const float* rrow = p.res ? p.res + r * p.N_out + col0 : nullptr;
        float* drow = p.dact ? p.dact + r * p.N_out + col0 : nullptr;
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
          if (col0 + 4 * q >= p.N_out) break;
          float o[4], d[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float z = __uint_as_float(acc[4 * q + u]);
            if (p.act) {
              // silu(z)/0.6 and its derivative; sigmoid and 1 - sigmoid from exp(-|z|): no overflow / cancellation
              const float t = expf(-fabsf(z));   // in (0, 1]: no overflow
              const float rr = 1.0f / (1.0f + t);
              const float sg = z >= 0.f ? rr : t * rr, oms = z >= 0.f ? t * rr : rr;
              o[u] = z * sg * kSiluScale * p.s_acc;
              d[u] = sg * fmaf(z, oms, 1.0f) * kSiluScale * p.s_acc;
            } else {
              o[u] = z * p.s_acc;
              d[u] = p.s_acc;
            }
          }
          if (rrow) {
            const float4 rv = __ldg(reinterpret_cast<const float4*>(rrow + 4 * q));
            o[0] += p.s_res * rv.x; o[1] += p.s_res * rv.y; o[2] += p.s_res * rv.z; o[3] += p.s_res * rv.w;
          }
          *reinterpret_cast<float4*>(yrow + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
          if (drow) *reinterpret_cast<float4*>(drow + 4 * q) = make_float4(d[0], d[1], d[2], d[3]);
        }
      }
    }
  };

  if (warp < PRODUCER_WARPS) {
    // ============================================================ producers
    griddep_wait();
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t r0 = (int64_t)tile * TILE_M;
      for (int kb = 0; kb < (coop ? 1 : NKB); ++kb, ++g) produce(r0, kb, g, tid);
    }
    if (coop && (int)blockIdx.x < n_tiles) {  // ... and take the other half of the epilogue columns
      mbar_wait(acc_full, 0u);
      tc_fence_after();
      const int cq = 2 + (warp >> 2);
      epilogue_cols(std::integral_constant<int, 16>{}, blockIdx.x, cq * (N / 4), (cq + 1) * (N / 4));
      tc_fence_before();
    }
  } else if (warp == PRODUCER_WARPS) {
    // ============================================================ MMA issuer
    if (elect_one()) {
      mbar_wait(w_bar, 0);  // weight image resident
      uint32_t g = 0;
      for (int it = 0, tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(acc_empty, ((uint32_t)it & 1u) ^ 1u);  // epilogue of the previous tile has drained TMEM
        const uint32_t d_corr = tmem_base + NKB * N;
        for (int kb = 0; kb < NKB; ++kb, ++g) {
          const int stage = g % kStages;
          mbar_wait(full_bar(stage), (g / kStages) & 1u);
          tc_fence_after();
          const uint32_t sa = stage_base + stage * STAGE_BYTES;
          const uint64_t a_h = make_desc(sa), a_m = make_desc(sa + A_SPLIT_BYTES), a_l = make_desc(sa + 2 * A_SPLIT_BYTES);
          const uint32_t wb = w_base + kb * 3 * N * KB * 2;
          const uint64_t w_h = make_desc(wb), w_m = make_desc(wb + N * KB * 2), w_l = make_desc(wb + 2 * N * KB * 2);
          const uint32_t d_main = tmem_base + kb * N;  // this chunk's own accumulator: 4 accumulates in total
#pragma unroll
          for (int kk = 0; kk < KB / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);  // 16 bf16 = 32 bytes along K
            umma_bf16(d_main, a_h + adv, w_h + adv, kIdesc, kk != 0);
            umma_bf16(d_corr, a_h + adv, w_m + adv, kIdesc, (kb | kk) != 0);
            umma_bf16(d_corr, a_m + adv, w_h + adv, kIdesc, 1u);
            umma_bf16(d_corr, a_m + adv, w_m + adv, kIdesc, 1u);
            umma_bf16(d_corr, a_h + adv, w_l + adv, kIdesc, 1u);
            umma_bf16(d_corr, a_l + adv, w_h + adv, kIdesc, 1u);
          }
          umma_commit(empty_bar(stage));
        }
        umma_commit(acc_full);
      }
    }
  } else if (warp == PRODUCER_WARPS + 1) {
    // ============================================================ weight loader (once per CTA)
    if (lane == 0) {
      mbar_arrive_expect_tx(w_bar, W_BYTES);
      constexpr int BLK = N * KB * 2;  // one (kb, split) block of this CTA's slab: N rows x 128 B
      for (int b = 0; b < NKB * 3; ++b)
        bulk_g2s(w_base + b * BLK,
                 reinterpret_cast<const uint8_t*>(p.Wp) + ((size_t)b * p.N_img + n_off) * (KB * 2), BLK, w_bar);
    }
  } else {
    // ============================================================ epilogue: 8 warps = 4 TMEM lane quadrants x 2 column halves
    const int ch = (warp - EPI_WARP0) >> 2;           // column half (quarter pair in cooperative mode)
    griddep_wait();
    if (coop && (int)blockIdx.x < n_tiles) produce((int64_t)blockIdx.x * TILE_M, 1, 1u, tid - EPI_WARP0 * 32);
    for (int it = 0, tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      mbar_wait(acc_full, (uint32_t)it & 1u);
      tc_fence_after();
      if (coop) epilogue_cols(std::integral_constant<int, 16>{}, tile, ch * (N / 4), (ch + 1) * (N / 4));
      else epilogue_cols(std::integral_constant<int, 32>{}, tile, ch * (N / 2), (ch + 1) * (N / 2));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int NKB, int N>
static int launch(const Params& p, cudaStream_t st, int n_slabs, bool pdl) {
  constexpr int W_BYTES = NKB * 3 * N * KB * 2;
  constexpr int SMEM = 1024 + W_BYTES + kStages * STAGE_BYTES + 512;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel<NKB, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("dense_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const int n_tiles = (int)ceil_div(p.M, TILE_M);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_tiles < sms ? n_tiles : sms, n_slabs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, dense_tc_kernel<NKB, N>, p, n_tiles);
  if (e != cudaSuccess) {
    set_error("gemnet_dense_tc: launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return check_launch("gemnet_dense_tc");
}

}  // namespace dtc
}  // namespace gemnet

using namespace gemnet;

static inline int round_k(int K) { return K <= 64 ? 64 : 128; }
static inline int round_n(int N) { return N <= 64 ? 64 : 128; }

GEMNET_API int64_t gemnet_dense_pack_floats(int32_t N, int32_t K) {
  if (N < 1 || N > 128 || K < 1 || K > 128) return -1;
  // the shared-memory image (dense_tc) + the tensor-memory image (chain_tc), bf16, counted in 4-byte units
  return (int64_t)(round_k(K) / dtc::KB) * 3 * round_n(N) * dtc::KB / 2 + (int64_t)3 * (round_k(K) / 2) * 128;
}

GEMNET_API int gemnet_dense_pack_weight(const float* W, int64_t stride_n, int64_t stride_k, int32_t N, int32_t K,
                                        float* packed, void* stream) {
  GEMNET_REQUIRE(W && packed, "null pointer");
  GEMNET_REQUIRE(N >= 1 && N <= 128 && K >= 1 && K <= 128, "dense_tc supports N, K in [1, 128]");
  const int Np = round_n(N), Kp = round_k(K);
  dtc::pack_weight_kernel<<<(Np * Kp + 255) / 256, 256, 0, as_stream(stream)>>>(
      W, stride_n, stride_k, N, K, Np, Kp, reinterpret_cast<__nv_bfloat16*>(packed));
  return check_launch("gemnet_dense_pack_weight");
}

GEMNET_API int gemnet_dense_pack_many(const gemnet_pack_desc* table_device, int32_t n_entries, void* stream) {
  GEMNET_REQUIRE(n_entries >= 0, "negative size");
  if (n_entries == 0) return 0;
  GEMNET_REQUIRE(table_device, "null pointer");
  dtc::pack_many_kernel<<<dim3((128 * 128 + 255) / 256, n_entries), 256, 0, as_stream(stream)>>>(table_device);
  return check_launch("gemnet_dense_pack_many");
}

GEMNET_API int gemnet_dense_tc(const float* x, const float* xmul, int64_t M, int32_t K, int32_t ldx,
                               const float* packed, int32_t N, int32_t act, float s_acc, const float* res, float s_res,
                               float* y, float* dact, int32_t flags, void* stream) {
  GEMNET_REQUIRE(M >= 0, "negative size");
  GEMNET_REQUIRE(N >= 4 && N <= 128 && N % 4 == 0 && K >= 4 && K <= 128 && K % 4 == 0 && ldx % 4 == 0 && ldx >= K,
                 "dense_tc needs N, K multiples of 4 in [4, 128] and a row stride that keeps 16-byte alignment");
  if (M == 0) return 0;
  GEMNET_REQUIRE(x && packed && y, "null pointer");
  GEMNET_REQUIRE((((uintptr_t)x | (uintptr_t)xmul | (uintptr_t)packed | (uintptr_t)res | (uintptr_t)y | (uintptr_t)dact) % 16) == 0,
                 "pointers must be 16-byte aligned");
  const int Kp = round_k(K), Np = round_n(N);
  static const int coop = [] {
    const char* e = getenv("GEMNET_B200_DENSE_COOP");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  dtc::Params p{x, xmul, packed, res, y, dact, M, ldx, K, N, act, s_acc, s_res, Np, coop};
  cudaStream_t st = as_stream(stream);
  const bool pdl = (flags & GEMNET_DENSE_WEIGHTS_SETTLED) != 0;
  if (Kp == 64 && Np == 64) return dtc::launch<1, 64>(p, st, 1, pdl);
  if (Kp == 128 && Np == 64) return dtc::launch<2, 64>(p, st, 1, pdl);
  // N = 128: with few row tiles (launch-bound sizes) two CTAs share a tile, 64 output columns each, so that twice as
  // many SMs work and each loads half of the weight image; large M keeps one CTA per tile (A is read only once)
  const bool split = dtc::ceil_div_i(M, dtc::TILE_M) <= 74 && N > 64;
  if (Kp == 64) return split ? dtc::launch<1, 64>(p, st, 2, pdl) : dtc::launch<1, 128>(p, st, 1, pdl);
  return split ? dtc::launch<2, 64>(p, st, 2, pdl) : dtc::launch<2, 128>(p, st, 1, pdl);
}
