// H7 backward, edge part, on the 5th-generation tensor cores (triplet shape (S,C,I,O) = (7,64,16,64)).
//
//   dB[e, c*I+i] = sum_o (scale * g[e,o]) W[c,i,o]                 <- tcgen05: (E x 64) @ (64 x 1024), bf16x3
//   dA[e, c, s]  = sum_i rbf_W1[e,i,s] dB[e,c,i]                   <- CUDA cores, straight out of TMEM
//   g_rbf_W1[e,i,s] = sum_c dB[e,c,i] A[e,c,s]                     <- CUDA cores, straight out of TMEM
//
// (efficient.py:159-189 differentiated: dB is the gradient of the (E,C,I) intermediate `rbf_W1_sum_k`; the fp32
// kernel segbil_bwd_edges_kernel spent 47 % of its instructions on the dB GEMM.)  dB never exists in memory: the
// 1024-column accumulator is produced 256 columns (16 values of c x all 16 i) at a time into one of two TMEM slots
// while the other slot is consumed.
//
// One CTA per SM, persistent over 128-edge tiles, 320 threads:
//   warps 0-3  role dA:    thread = edge row (TMEM lane); rbf_W1[e] (16x7) in registers; per c one tcgen05.ld of the
//              16 i-columns -> 7 values of dA[e,c,:] -> two 128-bit stores (layout (E, C, 8): s contiguous, padded)
//   warps 4-7  role g_rbf: thread = edge row; accumulates the 16x7 block of g_rbf_W1[e] over all 64 c in registers
//              from dB (TMEM) and A[e,c,:] (two 128-bit loads per c)
//   all 8      produce the A operand of the tile first: g rows -> scale -> exact bf16 hi/lo split -> swizzled tile
//   warp 8     lane 0 issues tcgen05.mma (M128 N256 K16, 3 products: hi*hi, hi*lo, lo*hi; the dropped lo*lo term is
//              2^-16 relative, the same fp32-parity mode as the forward kernel)
//   warp 9     lane 0 streams the packed weight image (W viewed as a (1024 x 64) K-major matrix: no transposition
//              needed, the pack kernel only splits and swizzles) through a 2-slot ring, 64 KB per 256-column chunk
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gemnet {
namespace btc {

using namespace tcx;

constexpr int S = 7, C = 64, I = 16, O = 64, SP = 8;
constexpr int NCOLS = C * I;                 // 1024 accumulator columns per edge
constexpr int TM = 128;
constexpr int NCH = 256;                     // columns per chunk = 16 values of c
constexpr int NCHUNK = NCOLS / NCH;          // 4
constexpr int CPC = NCH / I;                 // c values per chunk: 16
constexpr int A_SPLIT = TM * 128;            // 16384: one bf16 split of the g tile (128 rows x 64 k)
constexpr int A_BYTES = 2 * A_SPLIT;
constexpr int W_SPLIT = NCH * 128;           // 32768: one split of one chunk (256 rows x 64 k)
constexpr int W_SLOT = 2 * W_SPLIT;          // 65536
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int SMEM_BYTES = 1024 + A_BYTES + 2 * W_SLOT + 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr int64_t WPACK_BYTES = (int64_t)NCHUNK * W_SLOT;  // 262144

// image: [chunk][hi|lo][n_local (256)][64 bf16, 16-byte pieces XOR-swizzled by (n & 7)];  row n = c*I + i, k = o
__global__ void pack_weight_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (n, k)
  if (idx >= NCOLS * O) return;
  const int k = idx % O, n = idx / O;
  const float w = W[idx];                                  // W is (C, I, O) row-major == (n, k) row-major
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const int chunk = n / NCH, nl = n % NCH;
  const int phys = nl * 64 + (((k >> 3) ^ (nl & 7)) << 3) + (k & 7);
  Wp[(int64_t)(chunk * 2 + 0) * (NCH * 64) + phys] = hi;
  Wp[(int64_t)(chunk * 2 + 1) * (NCH * 64) + phys] = lo;
}

__global__ void __launch_bounds__(THREADS, 1)
segbil_bwd_edges_tc_kernel(const float* __restrict__ rbf_W1, const __nv_bfloat16* __restrict__ Wp, float scale,
                           int64_t n_edges, int n_tiles, const float* __restrict__ g_out,
                           const float* __restrict__ A_in /* (E, C, 8) */, float* __restrict__ dA_out /* (E, C, 8) */,
                           float* __restrict__ g_rbf_W1 /* (E, I, S) or NULL */) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_base = base;
  const uint32_t w_base = base + A_BYTES;
  const uint32_t bar_base = w_base + 2 * W_SLOT;
  const uint32_t a_ready = bar_base;
  auto w_full = [&](int s) { return bar_base + 8u + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 24u + 8u * s; };
  auto acc_full = [&](int s) { return bar_base + 40u + 8u * s; };
  auto acc_empty = [&](int s) { return bar_base + 56u + 8u * s; };
  const uint32_t tmem_slot = bar_base + 72;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + A_BYTES + 2 * W_SLOT + 72);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // (as a warp-uniform value: see tcx::elect_one)
  griddep_launch();
  if (tid == 0) {
    mbar_init(a_ready, EPI_WARPS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  if (warp < EPI_WARPS) {
    // ===================================================================== operand producers + epilogue roles
    griddep_wait();
    const bool role_da = warp < 4;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t g = 0;  // chunk counter (ring position of the accumulator slots)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t e0 = (int64_t)tile * TM;
      const int64_t e = e0 + row;
      const bool ev = e < n_edges;
      const int64_t ec = ev ? e : n_edges - 1;
      // ---- A operand: 128 rows x 8 granules of 8 k; every MMA of the previous tile has completed (this thread saw
      //      the acc_full of its last chunk), so the tile may be overwritten
      for (int p = tid; p < TM * 8; p += EPI_WARPS * 32) {
        const int pr = p >> 3, j = p & 7;
        const int64_t er = e0 + pr;
        float v[8];
        if (er < n_edges) {
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(g_out + er * O + j * 8));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(g_out + er * O + j * 8 + 4));
          v[0] = a0.x * scale; v[1] = a0.y * scale; v[2] = a0.z * scale; v[3] = a0.w * scale;
          v[4] = a1.x * scale; v[5] = a1.y * scale; v[6] = a1.z * scale; v[7] = a1.w * scale;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = 0.f;
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
          const float2 hf = __bfloat1622float2(hh);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * u] - hf.x, v[2 * u + 1] - hf.y);
          h[u] = *reinterpret_cast<const uint32_t*>(&hh);
          l[u] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        uint8_t* dst = sm + pr * 128 + ((j ^ (pr & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dst + A_SPLIT) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);

      // ---- per-edge register state of the role
      float rw[I][S];   // role dA: rbf_W1[e];  role g_rbf: the accumulator of g_rbf_W1[e]
      if (role_da) {
        const float4* src = reinterpret_cast<const float4*>(rbf_W1 + ec * (I * S));
        float tmp[I * S];
#pragma unroll
        for (int v4 = 0; v4 < I * S / 4; ++v4) {
          const float4 f = __ldg(src + v4);
          tmp[4 * v4] = f.x; tmp[4 * v4 + 1] = f.y; tmp[4 * v4 + 2] = f.z; tmp[4 * v4 + 3] = f.w;
        }
#pragma unroll
        for (int i = 0; i < I; ++i)
#pragma unroll
          for (int s = 0; s < S; ++s) rw[i][s] = tmp[i * S + s];
      } else {
#pragma unroll
        for (int i = 0; i < I; ++i)
#pragma unroll
          for (int s = 0; s < S; ++s) rw[i][s] = 0.f;
      }

      for (int chunk = 0; chunk < NCHUNK; ++chunk, ++g) {
        const int slot = g & 1u;
        mbar_wait(acc_full(slot), (g >> 1) & 1u);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * NCH);
        if (role_da || g_rbf_W1 != nullptr) {
#pragma unroll 2
          for (int cl = 0; cl < CPC; ++cl) {
            const int c = chunk * CPC + cl;
            uint32_t raw[16];
            tmem_ld16_nowait(ta + cl * I, raw);
            float4 a_lo = make_float4(0.f, 0.f, 0.f, 0.f), a_hi = a_lo;
            if (!role_da) {
              const float4* ap = reinterpret_cast<const float4*>(A_in + (ec * C + c) * SP);
              a_lo = __ldg(ap);
              a_hi = __ldg(ap + 1);
            }
            tmem_wait_ld();
            if (role_da) {
              float d[S];
#pragma unroll
              for (int s = 0; s < S; ++s) d[s] = 0.f;
#pragma unroll
              for (int i = 0; i < I; ++i) {
                const float b = __uint_as_float(raw[i]);
#pragma unroll
                for (int s = 0; s < S; ++s) d[s] = fmaf(rw[i][s], b, d[s]);
              }
              if (ev) {
                float4* dp = reinterpret_cast<float4*>(dA_out + (e * C + c) * SP);
                dp[0] = make_float4(d[0], d[1], d[2], d[3]);
                dp[1] = make_float4(d[4], d[5], d[6], 0.f);
              }
            } else {
              const float a[S] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z};
#pragma unroll
              for (int i = 0; i < I; ++i) {
                const float b = __uint_as_float(raw[i]);
#pragma unroll
                for (int s = 0; s < S; ++s) rw[i][s] = fmaf(b, a[s], rw[i][s]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(slot));
      }
      if (!role_da && g_rbf_W1 != nullptr && ev) {
        float4* dst = reinterpret_cast<float4*>(g_rbf_W1 + e * (I * S));
        float tmp[I * S];
#pragma unroll
        for (int i = 0; i < I; ++i)
#pragma unroll
          for (int s = 0; s < S; ++s) tmp[i * S + s] = rw[i][s];
#pragma unroll
        for (int v4 = 0; v4 < I * S / 4; ++v4) dst[v4] = make_float4(tmp[4 * v4], tmp[4 * v4 + 1], tmp[4 * v4 + 2], tmp[4 * v4 + 3]);
      }
    }
    tc_fence_before();
  } else if (warp == EPI_WARPS) {
    // ===================================================================== MMA issuer
    if (elect_one()) {   // (tc_common.cuh: uniform operands, no per-instruction election loop)
      const uint32_t idesc = make_idesc(TM, NCH);
      uint32_t g = 0;
      for (int it = 0, tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(a_ready, (uint32_t)it & 1u);
        tc_fence_after();
        const uint64_t a_hi = make_desc(a_base), a_lo = make_desc(a_base + A_SPLIT);
        for (int chunk = 0; chunk < NCHUNK; ++chunk, ++g) {
          const int slot = g & 1u;
          mbar_wait(acc_empty(slot), ((g >> 1) & 1u) ^ 1u);
          mbar_wait(w_full(slot), (g >> 1) & 1u);
          tc_fence_after();
          const uint32_t wb = w_base + slot * W_SLOT;
          const uint64_t w_hi = make_desc(wb), w_lo = make_desc(wb + W_SPLIT);
          const uint32_t d = tmem_base + slot * NCH;
#pragma unroll
          for (int kk = 0; kk < O / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);
            umma_bf16(d, a_hi + adv, w_hi + adv, idesc, kk != 0);
            umma_bf16(d, a_hi + adv, w_lo + adv, idesc, 1u);
            umma_bf16(d, a_lo + adv, w_hi + adv, idesc, 1u);
          }
          umma_commit(w_empty(slot));
          umma_commit(acc_full(slot));
        }
      }
    }
  } else {
    // ===================================================================== weight loader
    if (lane == 0) {
      griddep_wait();  // the image may have been packed by the kernel launched just before
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < NCHUNK; ++chunk, ++g) {
          const int slot = g & 1u;
          mbar_wait(w_empty(slot), ((g >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(w_full(slot), W_SLOT);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(Wp) + (size_t)chunk * W_SLOT;
          bulk_g2s(w_base + slot * W_SLOT, src, W_SPLIT, w_full(slot));
          bulk_g2s(w_base + slot * W_SLOT + W_SPLIT, src + W_SPLIT, W_SPLIT, w_full(slot));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace btc

int64_t segbil_bwd_tc_pack_bytes() { return btc::WPACK_BYTES; }

int launch_segbil_bwd_pack_tc(const float* W, void* packed, cudaStream_t st) {
  btc::pack_weight_kernel<<<(btc::NCOLS * btc::O + 255) / 256, 256, 0, st>>>(W, reinterpret_cast<__nv_bfloat16*>(packed));
  return check_launch("gemnet_segbilinear_pack(bwd)");
}

int launch_segbil_bwd_edges_tc(const float* rbf_W1, const void* packed, float scale, int64_t n_edges, const float* g_out,
                               const float* A_in, float* dA_out, float* g_rbf_W1, cudaStream_t st) {
  using namespace btc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(segbil_bwd_edges_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("segbilinear bwd tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const int n_tiles = (int)ceil_div(n_edges, TM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = n_tiles < sms ? n_tiles : sms;
  launch_pdl(segbil_bwd_edges_tc_kernel, grid, THREADS, SMEM_BYTES, st, rbf_W1,
             reinterpret_cast<const __nv_bfloat16*>(packed), scale, n_edges, n_tiles, g_out, A_in, dA_out, g_rbf_W1);
  return check_launch("gemnet_segbilinear_bwd(tcgen05 edges)");
}

}  // namespace gemnet
