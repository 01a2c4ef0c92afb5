// H7 forward on the 5th-generation tensor cores (algo 1): tcgen05.mma with the accumulator in TMEM.
//
//   out[e,:] = scale * B_e (1 x K) @ W' (K x O),   K = C*I = 1024, O = 64
//   B_e[k = c*I + i] = sum_s rbf_W1[e,i,s] * A_e[s,c],   A_e[s,c] = sum_{t in seg(e)} sph[t,s] * x[expand[t], c]
//
// Persistent warp-specialised kernel, one CTA per SM, tile = 64 target edges (UMMA M = 64, N = 64, K = 16):
//   phase 1  (16 producer warps, one warp per edge, lanes over c): the CSR segment is gathered ONCE per edge with
//            fully coalesced 256-byte row loads; A_e (7 x 64 fp32) goes to shared memory (As, 64 edges = 113 KB).
//   phase 2  (same 512 threads, 8 per edge row): for each of the 16 K-chunks (4 x-columns x 16 i = 64 k) the thread
//            (row, column cq, i-half ih) contracts its 7 A-values with its 8 x 7 register-resident rbf_W1 block and
//            writes 8 k-values, split into bf16 hi + bf16 lo, into the K-major SWIZZLE_128B operand tile of the stage.
//   warp 16  lane 0 issues tcgen05.mma: D += Ahi*Whi + Ahi*Wlo + Alo*Whi (bf16x3: the dropped lo*lo term is 2^-16
//            relative, so this is an fp32-parity mode, not a reduced-precision one).  Main products and correction
//            products go to separate TMEM accumulators (two slots of each); the epilogue undoes the mean
//            toward-zero rounding loss of the 64-step main chain.
//   warp 17  lane 0 streams the pre-packed (already swizzled) W chunk image with one cp.async.bulk per stage.
//   warps 0-3 also run the epilogue of the previous accumulator: tcgen05.ld, scale, 128-bit stores.
// Why 64-edge tiles: the gather side wants (edge-major, all columns) — coalesced, once — while the MMA wants
// (all edges, one K-chunk); the transposition has to live in shared memory and 128 edges x 7 x 64 fp32 does not fit.
// The first version of this kernel (128-edge tiles, per-chunk re-gather by one thread per edge row) spent 20 L1
// wavefronts per request on uncoalesced gathers (profiles/ncu_r01_segbil_fwd_tc_v1_large.txt).
//
// Shapes: only the published triplet shape (S,C,I,O) = (7,64,16,64).
#include <cuda_bf16.h>

#include "common.cuh"

namespace gemnet {
namespace tc {

constexpr int S = 7, C = 64, I = 16, O = 64;
constexpr int K = C * I;                 // 1024
constexpr int TILE_M = 64;
constexpr int KC = 64;                   // k per chunk (4 columns of x times I)
constexpr int NCHUNK = K / KC;           // 16
constexpr int kStages = 3;
constexpr int PRODUCER_WARPS = 16;
constexpr int PRODUCER_THREADS = PRODUCER_WARPS * 32;  // 512
constexpr int THREADS = PRODUCER_THREADS + 64;         // + MMA warp + loader warp

constexpr int A_TILE_BYTES = TILE_M * KC * 2;          // 8192 (bf16)
constexpr int W_TILE_BYTES = O * KC * 2;               // 8192
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * W_TILE_BYTES;  // 32768
constexpr int AS_LD = S * C + 4;                       // padded A_e row (floats): conflict-free phase-2 reads
constexpr int AS_BYTES = TILE_M * AS_LD * 4;           // 115712
constexpr int SMEM_BYTES = 1024 /*align slack*/ + kStages * STAGE_BYTES + AS_BYTES + 512 /*barriers*/;
constexpr uint32_t TMEM_COLS = 256;                    // two slots x (main + correction) 64-column accumulators
constexpr float kShrinkPerStep = 1.25e-8f;             // mean toward-zero rounding loss of one tcgen05 accumulate

// packed weight image: [chunk][hi|lo][o][k_local] bf16, rows already XOR-swizzled (SWIZZLE_128B)
constexpr int WPACK_BYTES = NCHUNK * 2 * W_TILE_BYTES;  // 262144

// ---------------------------------------------------------------------------------------------- PTX helpers
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
    if (!done && spin > (1u << 26)) __trap();  // never hang the device on a protocol bug
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
// instruction descriptor: D fp32, A/B bf16, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(O >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

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

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// ---------------------------------------------------------------------------------------------- weight packing
// Wp[chunk][h][o][chunk16 ^ (o & 7)][8 bf16]  with  k_local = cl*16 + i  <->  W[c = 4*chunk + cl][i][o]
__global__ void pack_weight_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (chunk, o, k_local)
  if (idx >= NCHUNK * O * KC) return;
  const int kl = idx % KC, o = (idx / KC) % O, chunk = idx / (KC * O);
  const int cl = kl / I, i = kl % I;
  const int c = chunk * (KC / I) + cl;
  __nv_bfloat16 hi, lo;
  split_bf16(W[((int64_t)c * I + i) * O + o], hi, lo);
  const int j = kl / 8;                       // 16-byte chunk within the 128-byte row
  const int phys = o * KC + ((j ^ (o & 7)) * 8) + (kl % 8);
  Wp[(int64_t)(chunk * 2 + 0) * (O * KC) + phys] = hi;
  Wp[(int64_t)(chunk * 2 + 1) * (O * KC) + phys] = lo;
}

// ---------------------------------------------------------------------------------------------- main kernel
__global__ void __launch_bounds__(THREADS, 1)
segbil_fwd_tc_kernel(const float* __restrict__ x, const int32_t* __restrict__ expand,
                     const int32_t* __restrict__ rowptr, const float* __restrict__ sph,
                     const float* __restrict__ rbf_W1, const __nv_bfloat16* __restrict__ Wp, float scale,
                     int64_t n_edges, int n_tiles, float* __restrict__ out, long long* __restrict__ prof) {
  pdl_enter();
  // diagnostics (gemnet_chain_set_profile_buffer): thread 0 of CTA 0 records clock64() at the phase boundaries
  int prof_n = 0;
  auto stamp = [&]() {
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && prof_n < 250) prof[prof_n++] = clock64();
  };
  stamp();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage_base = base;                                    // kStages * STAGE_BYTES
  float* As = reinterpret_cast<float*>(sm + kStages * STAGE_BYTES);    // [TILE_M][AS_LD]: A_e[s][c] at s*C + c
  const uint32_t bar_base = base + kStages * STAGE_BYTES + AS_BYTES;
  // every producer warp produces every chunk, in order: a plain stage ring (full/empty per stage) is safe
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(sm + kStages * STAGE_BYTES + AS_BYTES + 8 * (2 * kStages + 4));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // (as a warp-uniform value: see elect_one)

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), PRODUCER_WARPS + 1);  // 16 producer warps + the weight loader's expect_tx
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == PRODUCER_WARPS) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  if (warp < PRODUCER_WARPS) {
    // ============================================================ producers (+ epilogue on warps 0-3)
    const int el = tid >> 3;          // phase-2 edge row of this thread
    const int q = tid & 7;
    const int cq = q & 3, ih = q >> 2;  // column within the chunk, half of the 16 i
    uint32_t g = 0;                   // chunk counter = position in the stage ring

    // epilogue of accumulator slot `slot` for tile `tile` (warps 0-3; M=64 layout: row r -> TMEM lane (r%16)+32*(r/16))
    auto epilogue = [&](int it, int tile) {
      const int slot = it & 1;
      mbar_wait(acc_full(slot), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      const int64_t er = (int64_t)tile * TILE_M + warp * 16 + lane;  // valid for lane < 16
      // The 64 hi*hi accumulates of the main accumulator each round toward zero (mean loss 1.25e-8 per step, see
      // dense_tc.cu / profiles/diag_shrink.py): add the mean loss back through the small correction accumulator.
      constexpr float kComp = kShrinkPerStep * (float)(NCHUNK * (KC / 16));
#pragma unroll 1
      for (int part = 0; part < 4; ++part) {  // 16 columns at a time keeps the epilogue out of the producers' registers
        uint32_t r[16], c[16];
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + slot * 2 * O + part * 16;
        tmem_ld16(ta, r);
        tmem_ld16(ta + O, c);
        if (lane < 16 && er < n_edges) {
          float4* dst = reinterpret_cast<float4*>(out + er * O + part * 16);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            float o4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float mainv = __uint_as_float(r[4 * v + u]);
              o4[u] = (mainv + fmaf(mainv, kComp, __uint_as_float(c[4 * v + u]))) * scale;
            }
            dst[v] = make_float4(o4[0], o4[1], o4[2], o4[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(slot));
    };

    int prev_tile = -1, it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int64_t e0 = (int64_t)tile * TILE_M;
      stamp();  // tile: start
      asm volatile("bar.sync 1, %0;" ::"n"(PRODUCER_THREADS) : "memory");  // As of the previous tile is no longer read

      // ---------------- phase 1: one HALF-warp per edge (two independent segments per warp), 4 columns per lane.
      // The gather indices of up to 16 triplets are fetched with one coalesced load per half-warp and broadcast by
      // shuffles, so the row gathers do not wait for an index load, and they are issued four triplets at a time:
      // 8 independent 128-bit requests in flight per warp instead of one dependent pair (the long-scoreboard stall of
      // the round-1 kernel, profiles/ncu_r01_final_segbil_fwd_tc_large.txt).
      for (int rp = warp * 2; rp < TILE_M; rp += PRODUCER_WARPS * 2) {
        const int hl = lane & 15, half = lane >> 4;
        const int r = rp + half;
        const int64_t e = e0 + r;
        float A[S][4];
#pragma unroll
        for (int s = 0; s < S; ++s) A[s][0] = A[s][1] = A[s][2] = A[s][3] = 0.f;
        int32_t beg = 0, end = 0;
        if (e < n_edges) {
          beg = __ldg(rowptr + e);
          end = __ldg(rowptr + e + 1);
        }
        // both halves run the same number of 16-triplet chunks (the longer segment decides): shuffles stay convergent
        const int len = end - beg;
        const int len_max = max(len, __shfl_xor_sync(0xffffffffu, len, 16));
        for (int base = 0; base < len_max; base += 16) {
          const int n = min(16, len - base);                       // triplets of this half in the chunk (may be <= 0)
          const int32_t my_idx = hl < n ? __ldg(expand + beg + base + hl) : 0;
          const int n_max = min(16, len_max - base);
          for (int k0 = 0; k0 < n_max; k0 += 4) {
            float4 xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int32_t ix = __shfl_sync(0xffffffffu, my_idx, (half << 4) + ((k0 + u) & 15));
              xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (k0 + u < n) xv[u] = __ldg(reinterpret_cast<const float4*>(x + (int64_t)ix * C) + hl);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (k0 + u < n) {
                const float* sp = sph + (int64_t)(beg + base + k0 + u) * S;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                  const float w = __ldg(sp + s);
                  A[s][0] = fmaf(w, xv[u].x, A[s][0]);
                  A[s][1] = fmaf(w, xv[u].y, A[s][1]);
                  A[s][2] = fmaf(w, xv[u].z, A[s][2]);
                  A[s][3] = fmaf(w, xv[u].w, A[s][3]);
                }
              }
            }
          }
        }
        float* arow = As + r * AS_LD + 4 * hl;
#pragma unroll
        for (int s = 0; s < S; ++s) *reinterpret_cast<float4*>(arow + s * C) = make_float4(A[s][0], A[s][1], A[s][2], A[s][3]);
      }

      // rbf_W1 block of this thread: rows i = 8*ih .. 8*ih+7, all 7 s  (56 contiguous floats, 16-byte aligned)
      float rw[8][S];
      {
        const int64_t e = e0 + el;
        if (e < n_edges) {
          const float4* src = reinterpret_cast<const float4*>(rbf_W1 + e * (I * S) + ih * 8 * S);
          float tmp[56];
#pragma unroll
          for (int v = 0; v < 14; ++v) {
            const float4 f = __ldg(src + v);
            tmp[4 * v] = f.x; tmp[4 * v + 1] = f.y; tmp[4 * v + 2] = f.z; tmp[4 * v + 3] = f.w;
          }
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8)
#pragma unroll
            for (int s = 0; s < S; ++s) rw[i8][s] = tmp[i8 * S + s];
        } else {
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8)
#pragma unroll
            for (int s = 0; s < S; ++s) rw[i8][s] = 0.f;
        }
      }
      stamp();  // tile: phase 1 of this warp done (+ rbf_W1 registers loaded)
      asm volatile("bar.sync 1, %0;" ::"n"(PRODUCER_THREADS) : "memory");  // As complete
      stamp();  // tile: As complete (all warps)

      // epilogue of the PREVIOUS tile here: its MMAs had the whole phase 1 of this tile to finish
      if (warp < 4 && prev_tile >= 0) epilogue(it - 1, prev_tile);
      stamp();  // tile: epilogue of the previous tile done

      // ---------------- phase 2: 16 chunks of 64 k; thread = (row el, column cq, i-half ih) -> 8 k-values
      const float* arow = As + el * AS_LD + cq;
      for (int chunk = 0; chunk < NCHUNK; ++chunk, ++g) {
        const int stage = g % kStages;
        const uint32_t parity = (g / kStages) & 1u;
        float a[S];
#pragma unroll
        for (int s = 0; s < S; ++s) a[s] = arow[s * C + chunk * 4];
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float b0 = 0.f, b1 = 0.f;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            b0 = fmaf(rw[2 * v][s], a[s], b0);
            b1 = fmaf(rw[2 * v + 1][s], a[s], b1);
          }
          // packed round-to-nearest conversions (one F2FP per pair instead of two scalar F2F + a PRMT)
          const __nv_bfloat162 h = __floats2bfloat162_rn(b0, b1);  // .x (low half) = b0 -> k, .y = b1 -> k + 1
          const float2 hf = __bfloat1622float2(h);
          const __nv_bfloat162 l = __floats2bfloat162_rn(b0 - hf.x, b1 - hf.y);
          hw[v] = *reinterpret_cast<const uint32_t*>(&h);
          lw[v] = *reinterpret_cast<const uint32_t*>(&l);
        }
        mbar_wait(empty_bar(stage), parity ^ 1u);  // the MMAs that read this stage have completed
        uint8_t* a_hi = sm + stage * STAGE_BYTES + el * 128;
        const uint32_t off = (uint32_t)(((cq * 2 + ih) ^ (el & 7)) << 4);  // k_local = cq*16 + ih*8 + (0..7)
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(a_hi + A_TILE_BYTES + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(stage));
      }
      stamp();  // tile: phase 2 done (16 chunks handed to the MMA)
      prev_tile = tile;
    }
    if (warp < 4 && prev_tile >= 0) epilogue(it - 1, prev_tile);
    stamp();  // last epilogue done
  } else if (warp == PRODUCER_WARPS) {
    // ============================================================ MMA issuer
    if (elect_one()) {
      uint32_t g = 0;
      for (int it = 0, tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int slot = it & 1;
        mbar_wait(acc_empty(slot), (((uint32_t)it >> 1) & 1u) ^ 1u);  // epilogue of the tile two back is done
        const uint32_t d_main = tmem_base + slot * 2 * O, d_corr = d_main + O;
        for (int chunk = 0; chunk < NCHUNK; ++chunk, ++g) {
          const int stage = g % kStages;
          mbar_wait(full_bar(stage), (g / kStages) & 1u);
          tc_fence_after();
          const uint32_t sa = stage_base + stage * STAGE_BYTES;
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_TILE_BYTES);
          const uint64_t w_hi = make_desc(sa + 2 * A_TILE_BYTES), w_lo = make_desc(sa + 2 * A_TILE_BYTES + W_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);  // 16 bf16 = 32 bytes along K inside the swizzle atom
            umma_bf16(d_main, a_hi + adv, w_hi + adv, kIdesc, (chunk | kk) != 0);
            umma_bf16(d_corr, a_hi + adv, w_lo + adv, kIdesc, (chunk | kk) != 0);
            umma_bf16(d_corr, a_lo + adv, w_hi + adv, kIdesc, 1u);
          }
          umma_commit(empty_bar(stage));
        }
        umma_commit(acc_full(slot));
      }
    }
  } else {
    // ============================================================ W loader
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < NCHUNK; ++chunk, ++g) {
          const int stage = g % kStages;
          mbar_wait(empty_bar(stage), ((g / kStages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), 2 * W_TILE_BYTES);
          bulk_g2s(stage_base + stage * STAGE_BYTES + 2 * A_TILE_BYTES,
                   reinterpret_cast<const uint8_t*>(Wp) + (size_t)chunk * 2 * W_TILE_BYTES, 2 * W_TILE_BYTES,
                   full_bar(stage));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc

int64_t segbil_tc_workspace_bytes() { return tc::WPACK_BYTES; }

static long long* g_h7_prof = nullptr;
void segbil_fwd_tc_set_profile_buffer(long long* p) { g_h7_prof = p; }

int launch_segbil_fwd_pack_tc(const float* W, void* packed, cudaStream_t st) {
  using namespace tc;
  pack_weight_kernel<<<(NCHUNK * O * KC + 255) / 256, 256, 0, st>>>(W, reinterpret_cast<__nv_bfloat16*>(packed));
  return check_launch("gemnet_segbilinear_pack(fwd)");
}

int launch_segbil_fwd_tc_packed(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                                const float* rbf_W1, const void* packed, float scale, int64_t n_edges, float* out,
                                cudaStream_t st) {
  using namespace tc;
  if (!packed || ((uintptr_t)packed % 16) != 0) {
    set_error("gemnet_segbilinear_fwd(tcgen05): packed weight image missing / misaligned");
    return 1;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(segbil_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("segbilinear tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const int n_tiles = (int)ceil_div(n_edges, TILE_M);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = n_tiles < sms ? n_tiles : sms;
  launch_pdl(segbil_fwd_tc_kernel, grid, THREADS, SMEM_BYTES, st, x, expand, rowptr, sph, rbf_W1,
             reinterpret_cast<const __nv_bfloat16*>(packed), scale, n_edges, n_tiles, out, g_h7_prof);
  return check_launch("gemnet_segbilinear_fwd(algo=1)");
}

int launch_segbil_fwd_tc(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                         const float* rbf_W1, const float* W, float scale, int64_t n_edges, float* out,
                         void* workspace, cudaStream_t st) {
  if (!workspace || ((uintptr_t)workspace % 16) != 0) {
    set_error("gemnet_segbilinear_fwd(algo=1): workspace of gemnet_segbilinear_workspace_bytes() bytes required");
    return 1;
  }
  if (launch_segbil_fwd_pack_tc(W, workspace, st)) return 1;
  return launch_segbil_fwd_tc_packed(x, expand, rowptr, sph, rbf_W1, workspace, scale, n_edges, out, st);
}

}  // namespace gemnet
