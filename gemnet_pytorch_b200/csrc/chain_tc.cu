// N1 chain fusion: a whole stack of dense layers with the pointwise work between them in ONE tcgen05 kernel
// (include/gemnet_b200.h, gemnet_chain_tc).  Batch-64 GemNet is latency bound: ~25 dependent launches per
// interaction block in each direction, every one a 128x128 GEMM that finishes in microseconds.  Here a 128-row tile of
// the activation stays on chip for the whole chain:
//
//   operand A      shared memory, exact three-way bf16 split (h + m + l == fp32 value), K-major SWIZZLE_128B, 96 KB
//   weights        streamed through a 2-slot ring (one 64-k chunk = 48 KB per slot) by cp.async.bulk, one thread;
//                  the next layer's weights arrive while the current layer's epilogue runs
//   accumulators   TMEM: one main accumulator per K chunk + one correction accumulator (dense_tc.cu's numerics:
//                  six significant products, the toward-zero rounding of the accumulator compensated)
//   skip value     "carry": the last 128 TMEM columns (tcgen05.st / tcgen05.ld of the thread's row x 32-column patch):
//                  it costs no registers and no shared memory
//
// 512 threads (16 warps: 128 registers per thread), one CTA per SM, persistent over row tiles.  Every thread is a
// worker: operand fill from global memory (coalesced; gathers and several summed terms allowed) and the epilogue of
// every step — tcgen05.ld of the thread's (row, 32 columns) patch, the pointwise program (activation, Hadamard
// products, skip connections, scales), the stores the backward pass needs, and the next layer's operand written back
// into shared memory (so the GEMM chain never touches HBM).  Thread 0 additionally issues the tcgen05.mma of each
// step (6 products per K16 step) once the operand is complete (one __syncthreads per GEMM) and, as soon as a GEMM
// has completed (mbarrier acc_full), the cp.async.bulk of the NEXT GEMM's packed weight image into the two chunk
// slots — the weights stream in while the epilogue runs.  The step program is a kernel parameter (__grid_constant__).
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace gemnet {
namespace chain {

using namespace tcx;

constexpr int TM = 128;
constexpr int SPLIT_BYTES = TM * 128;          // one bf16 split of one 64-k chunk: 128 rows x 128 B
constexpr int CHUNK_BYTES = 3 * SPLIT_BYTES;   // 49152
constexpr int A_BYTES = 2 * CHUNK_BYTES;       // 98304
constexpr int WSLOT_BYTES = 3 * 128 * 128;     // 49152: one 64-k chunk of a (<=128)-row weight image, three splits
constexpr int W_SLOTS = 2;
constexpr int THREADS = 512;
constexpr int MISC_BYTES = 1024;               // barriers, TMEM slot, dot buffer
constexpr int SMEM_BYTES = 1024 + A_BYTES + W_SLOTS * WSLOT_BYTES + MISC_BYTES;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC_STRIDE = 128, CORR_COL = 256, CARRY_COL = 384;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// coherent load: for tensors that an earlier step of the SAME kernel (the same thread) may have written
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void store_granule(uint8_t* A, int row, int g, const float (&v)[8]) {
  // granule g (8 k-values) of operand row `row`: chunk g / 8, 16-byte slot (g % 8) XOR-swizzled by the row
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) split3(v[2 * u], v[2 * u + 1], h[u], m[u], l[u]);
  uint8_t* p = A + (g >> 3) * CHUNK_BYTES + row * 128 + (((g & 7) ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(p + SPLIT_BYTES) = make_uint4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint4*>(p + 2 * SPLIT_BYTES) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void load_granule(const uint8_t* A, int row, int g, float (&v)[8]) {
  const uint8_t* p = A + (g >> 3) * CHUNK_BYTES + row * 128 + (((g & 7) ^ (row & 7)) << 4);
  const uint4 h = *reinterpret_cast<const uint4*>(p);
  const uint4 m = *reinterpret_cast<const uint4*>(p + SPLIT_BYTES);
  const uint4 l = *reinterpret_cast<const uint4*>(p + 2 * SPLIT_BYTES);
  join3(h.x, m.x, l.x, v[0], v[1]);
  join3(h.y, m.y, l.y, v[2], v[3]);
  join3(h.z, m.z, l.z, v[4], v[5]);
  join3(h.w, m.w, l.w, v[6], v[7]);
}

__global__ void __launch_bounds__(THREADS, 1)
chain_tc_kernel(const __grid_constant__ gemnet_chain_program prog, int n_tiles, int weights_settled,
                long long* __restrict__ prof) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* A = sm;                                      // operand: [chunk][split][row][128 B]
  const uint32_t a_base = base;
  const uint32_t w_base = base + A_BYTES;               // weight slots: [chunk][split][n][128 B]
  const uint32_t bar_base = w_base + W_SLOTS * WSLOT_BYTES;
  const uint32_t acc_full = bar_base;
  auto w_full = [&](int s) { return bar_base + 8u + 8u * s; };
  const uint32_t tmem_slot = bar_base + 32;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + A_BYTES + W_SLOTS * WSLOT_BYTES + 32);
  float* dotbuf = reinterpret_cast<float*>(sm + A_BYTES + W_SLOTS * WSLOT_BYTES + 64);  // 128 floats

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // diagnostics (gemnet_chain_set_profile_buffer): thread 0 of CTA 0 records clock64() at the phase boundaries
  int prof_n = 0;
  auto stamp = [&]() {
    if (prof != nullptr && blockIdx.x == 0 && tid == 0 && prof_n < 250) prof[prof_n++] = clock64();
  };
  stamp();
  const int64_t M = prog.M;
  const int n_steps = prog.n_steps;

  // Programmatic dependent launch: the next kernel of the stream may start its prologue now; everything here that
  // touches activations waits in griddep_wait() for the PREVIOUS kernel to have completed and flushed.
  griddep_launch();
  if (tid == 0) {
    mbar_init(acc_full, 1);
    for (int s = 0; s < W_SLOTS; ++s) mbar_init(w_full(s), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  stamp();  // 1: prologue done

  // weight image of one GEMM -> the chunk slots (thread 0 only; the slots are free: every MMA that read them is done)
  auto fetch_weights = [&](const gemnet_chain_step& st) {
    const int NKB = st.K > 64 ? 2 : 1;
    const uint32_t blk = (uint32_t)(st.N > 64 ? 128 : 64) * 128u;   // one (chunk, split) block of the image
    for (int kb = 0; kb < NKB; ++kb) {
      mbar_arrive_expect_tx(w_full(kb), 3 * blk);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(st.Wp) + (size_t)kb * 3 * blk;
      for (int sp = 0; sp < 3; ++sp) bulk_g2s(w_base + kb * WSLOT_BYTES + sp * blk, src + (size_t)sp * blk, blk, w_full(kb));
    }
  };
  if (tid == 0 && (int)blockIdx.x < n_tiles) {
    if (!weights_settled) griddep_wait();  // the packed images may come from the kernel launched just before
    fetch_weights(prog.step[0]);
  }
  griddep_wait();
  stamp();  // 2: predecessor complete

  const int q = warp & 3, cg = warp >> 2;       // TMEM lane quadrant (rows 32q..32q+31), column group
  const int row = q * 32 + lane;                // epilogue row of this thread inside the tile
  // carry: 16 columns of this thread's row at TMEM column CARRY_COL + col
  auto carry_ld = [&](int col, float (&c)[16]) {
    uint32_t raw[16];
    tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + CARRY_COL + (uint32_t)col, raw);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = __uint_as_float(raw[j]);
  };
  auto carry_st = [&](int col, const float (&c)[16]) {
    uint32_t raw[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) raw[j] = __float_as_uint(c[j]);
    tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + CARRY_COL + (uint32_t)col, raw);
  };
  uint32_t gemm = 0;                            // GEMMs issued by this CTA (parity of acc_full)
  uint32_t wuse0 = 0, wuse1 = 0;                // fills of weight slot 0 / 1 consumed so far (thread 0)

  // the thread's epilogue patch of the NEXT operand (width W = Kpad / 4 columns starting at W * cg) from the carry
  auto patch_from_carry = [&](int Knext, float s_a, const float* amul, int64_t rcl, bool rv) {
    const int Wd = Knext > 64 ? 32 : 16;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h * 16 >= Wd) break;
      float c[16];
      carry_ld(Wd * cg + h * 16, c);
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {
        const int col = Wd * cg + h * 16 + g2 * 8;
        const bool gv = rv && col < Knext;
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = gv ? c[g2 * 8 + j] * s_a : 0.f;
        if (amul) {
          const int cn = min(col, Knext - 8);
          const float4 m0 = ldg4(amul + rcl * Knext + cn), m1 = ldg4(amul + rcl * Knext + cn + 4);
          w[0] *= m0.x; w[1] *= m0.y; w[2] *= m0.z; w[3] *= m0.w;
          w[4] *= m1.x; w[5] *= m1.y; w[6] *= m1.z; w[7] *= m1.w;
        }
        store_granule(A, row, col >> 3, w);
      }
    }
  };

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = (int64_t)tile * TM;
    const int64_t r = r0 + row;
    const bool rv = r < M;
#pragma unroll 1
    for (int s = 0; s < n_steps; ++s) {
      const gemnet_chain_step& st = prog.step[s];
      const int K = st.K, N = st.N;
      const int NKB = K > 64 ? 2 : 1;
      // ------------------------------------------------------------------ operand fill from global memory
      if (st.n_terms > 0) {
        const int gpr = 8 * NKB;                                   // granules (8 columns) per operand row
        const bool fold_amul = st.amul != nullptr && !st.load_set_carry;
        // PG (row, granule) pieces per thread and pass: all loads of a pass are in flight together.  The common fill
        // (one plain or gathered term) takes its 4 pieces in ONE pass; fills with Hadamard factors or several terms
        // take 2 per pass (registers).
        auto fill = [&](auto pg_tag) {
          constexpr int PG = decltype(pg_tag)::value;
#pragma unroll 1
          for (int p0 = tid; p0 < TM * gpr; p0 += PG * THREADS) {
            float v[PG][8];
            int prow[PG], pgr[PG];
#pragma unroll
            for (int h = 0; h < PG; ++h) {
              const int p = p0 + h * THREADS;
              prow[h] = p / gpr;
              pgr[h] = p % gpr;
#pragma unroll
              for (int j = 0; j < 8; ++j) v[h][j] = 0.f;
            }
            for (int t = 0; t < st.n_terms; ++t) {
              const gemnet_chain_term& tm = st.term[t];
              float4 a0[PG], a1[PG];
              if (tm.kind == 0) {
#pragma unroll
                for (int h = 0; h < PG; ++h) {
                  // out-of-range pieces read a clamped (valid) address and are zeroed below: no divergent branches
                  const int64_t rr = min(r0 + prow[h], M - 1);
                  const int col = min(pgr[h] * 8, K - 8);
                  const int64_t sr = tm.idx ? (int64_t)__ldg(tm.idx + rr) : rr;
                  a0[h] = ldg4(tm.src + sr * K + col);
                  a1[h] = ldg4(tm.src + sr * K + col + 4);
                }
                if (tm.mul) {
#pragma unroll
                  for (int h = 0; h < PG; ++h) {
                    const int64_t rr = min(r0 + prow[h], M - 1);
                    const int col = min(pgr[h] * 8, K - 8);
                    const float4 m0 = ldg4(tm.mul + rr * K + col), m1 = ldg4(tm.mul + rr * K + col + 4);
                    a0[h].x *= m0.x; a0[h].y *= m0.y; a0[h].z *= m0.z; a0[h].w *= m0.w;
                    a1[h].x *= m1.x; a1[h].y *= m1.y; a1[h].z *= m1.z; a1[h].w *= m1.w;
                  }
                }
              } else {  // outer product src[row] * mul[col]
#pragma unroll
                for (int h = 0; h < PG; ++h) {
                  const int64_t rr = min(r0 + prow[h], M - 1);
                  const int col = min(pgr[h] * 8, K - 8);
                  const float a = __ldg(tm.src + rr);
                  const float4 m0 = ldg4(tm.mul + col), m1 = ldg4(tm.mul + col + 4);
                  a0[h] = make_float4(a * m0.x, a * m0.y, a * m0.z, a * m0.w);
                  a1[h] = make_float4(a * m1.x, a * m1.y, a * m1.z, a * m1.w);
                }
              }
              const float sc = tm.scale;
#pragma unroll
              for (int h = 0; h < PG; ++h) {
                v[h][0] = fmaf(sc, a0[h].x, v[h][0]); v[h][1] = fmaf(sc, a0[h].y, v[h][1]);
                v[h][2] = fmaf(sc, a0[h].z, v[h][2]); v[h][3] = fmaf(sc, a0[h].w, v[h][3]);
                v[h][4] = fmaf(sc, a1[h].x, v[h][4]); v[h][5] = fmaf(sc, a1[h].y, v[h][5]);
                v[h][6] = fmaf(sc, a1[h].z, v[h][6]); v[h][7] = fmaf(sc, a1[h].w, v[h][7]);
              }
            }
#pragma unroll
            for (int h = 0; h < PG; ++h) {
              if (fold_amul) {
                const int64_t rr = min(r0 + prow[h], M - 1);
                const int col = min(pgr[h] * 8, K - 8);
                const float4 m0 = ldg4(st.amul + rr * K + col), m1 = ldg4(st.amul + rr * K + col + 4);
                v[h][0] *= m0.x; v[h][1] *= m0.y; v[h][2] *= m0.z; v[h][3] *= m0.w;
                v[h][4] *= m1.x; v[h][5] *= m1.y; v[h][6] *= m1.z; v[h][7] *= m1.w;
              }
              const bool valid = r0 + prow[h] < M && pgr[h] * 8 < K;
#pragma unroll
              for (int j = 0; j < 8; ++j) v[h][j] = valid ? v[h][j] : 0.f;
              store_granule(A, prow[h], pgr[h], v[h]);
            }
          }
        };
        const bool simple = st.n_terms == 1 && st.term[0].mul == nullptr && !fold_amul && NKB == 2;
        if (simple) fill(std::integral_constant<int, 4>{});
        else fill(std::integral_constant<int, 2>{});
        if (st.load_set_carry) {
          // the skip value is wanted in the epilogue's thread mapping: read it back from the (exact) operand split
          __syncthreads();
          const int Wd = K > 64 ? 32 : 16;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h * 16 >= Wd) break;
            float c[16], lo[8], hi[8];
            load_granule(A, row, (Wd * cg + h * 16) >> 3, lo);
            load_granule(A, row, ((Wd * cg + h * 16) >> 3) + 1, hi);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              c[j] = lo[j];
              c[8 + j] = hi[j];
            }
            carry_st(Wd * cg + h * 16, c);
          }
          if (st.amul) patch_from_carry(K, 1.0f, st.amul, rv ? r : M - 1, rv);
        }
      }
      stamp();  // step: operand written by this thread
      // ------------------------------------------------------------------ GEMM: operand complete -> thread 0 issues
      fence_proxy_async();   // generic-proxy writes of A -> visible to the tensor-core (async) proxy
      tc_fence_before();     // this thread's tcgen05.ld of the previous accumulators are done
      __syncthreads();
      stamp();  // step: all threads' operand parts visible
      if (tid == 0) {
        tc_fence_after();
        const int Npad = N > 64 ? 128 : 64;
        const uint32_t idesc = make_idesc(TM, Npad);
        const uint32_t d_corr = tmem_base + CORR_COL;
        for (int kb = 0; kb < NKB; ++kb) {
          uint32_t& use = kb ? wuse1 : wuse0;
          mbar_wait(w_full(kb), use & 1u);
          ++use;
          stamp();  // step: weight chunk kb resident
          tc_fence_after();
          const uint32_t sa = a_base + kb * CHUNK_BYTES;
          const uint64_t a_h = make_desc(sa), a_m = make_desc(sa + SPLIT_BYTES), a_l = make_desc(sa + 2 * SPLIT_BYTES);
          const uint32_t wb = w_base + kb * WSLOT_BYTES;
          const uint64_t w_h = make_desc(wb), w_m = make_desc(wb + Npad * 128), w_l = make_desc(wb + 2 * Npad * 128);
          const uint32_t d_main = tmem_base + kb * ACC_STRIDE;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);  // 16 bf16 = 32 bytes along K inside the swizzle atom
            umma_bf16(d_main, a_h + adv, w_h + adv, idesc, kk != 0);
            umma_bf16(d_corr, a_h + adv, w_m + adv, idesc, (kb | kk) != 0);
            umma_bf16(d_corr, a_m + adv, w_h + adv, idesc, 1u);
            umma_bf16(d_corr, a_m + adv, w_m + adv, idesc, 1u);
            umma_bf16(d_corr, a_h + adv, w_l + adv, idesc, 1u);
            umma_bf16(d_corr, a_l + adv, w_h + adv, idesc, 1u);
          }
        }
        umma_commit(acc_full);
        stamp();  // step: MMAs issued
      }
      __syncwarp();
      mbar_wait(acc_full, gemm & 1u);
      stamp();  // step: accumulators complete
      ++gemm;
      tc_fence_after();
      const bool last = s + 1 == n_steps;
      if (tid == 0) {  // both weight slots are free now: stream the next GEMM's image in while the epilogue runs
        if (!last) fetch_weights(prog.step[s + 1]);
        else if (tile + (int)gridDim.x < n_tiles) fetch_weights(prog.step[0]);
      }

      // ------------------------------------------------------------------ epilogue of this step's GEMM
      const int Wd = N > 64 ? 32 : 16;                  // columns of this thread's patch
      const int flags = st.flags;
      const int Knext = last ? 0 : prog.step[s + 1].K;
      const float c0f = kShrinkPerStep * (float)min(4, (K + 15) / 16);
      const float c1f = kShrinkPerStep * (float)max(0, min(4, (K - 64 + 15) / 16));
      const int64_t rc = rv ? r : M - 1;   // loads of rows past the end read the last row (and are discarded)
      const float* const pre0 = st.pre0; const float* const pre1 = st.pre1;
      const float* const mul_early = st.mul_early; const float* const mul_late = st.mul_late;
      const float* const post0 = st.post0; const float* const post1 = st.post1;
      const float s_acc = st.s_acc, s_carry = st.s_carry;
      const int64_t pr0 = pre0 ? (st.pre0_idx ? (int64_t)__ldg(st.pre0_idx + rc) : rc) : 0;
      const int64_t pr1 = pre1 ? (st.pre1_idx ? (int64_t)__ldg(st.pre1_idx + rc) : rc) : 0;
      const int64_t orow = st.out ? (st.out_idx ? (int64_t)__ldg(st.out_idx + rc) : rc) : 0;
      float dot = 0.f;
      if (st.dot_w) {
        if (cg == 0) dotbuf[row] = 0.f;
        __syncthreads();
      }
      const bool need_d = st.dact != nullptr;
#pragma unroll
      for (int piece = 0; piece < 2; ++piece) {
        if (piece * 16 >= Wd) break;
        // The pointwise program runs STAGE BY STAGE over the 16 values of the piece (feature tests outside the element
        // loops): 16 independent dependency chains per stage keep the FMA / MUFU pipes busy with only 4 warps per
        // scheduler, and a disabled feature costs one uniform branch instead of one per float4.
        const int c0 = Wd * cg + piece * 16;            // first column of this 16-column piece
        float v[16], d[16];
        {
          uint32_t a0[16], a1[16], cc[16];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          tmem_ld16_nowait(ta, a0);
          tmem_ld16_nowait(ta + CORR_COL, cc);
          if (NKB == 2) tmem_ld16_nowait(ta + ACC_STRIDE, a1);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // compensated sum (dense_tc.cu): the mean truncation loss of the main chains goes through the small
            // correction sum, then ONE round-to-nearest add per accumulator
            float corr = fmaf(__uint_as_float(a0[j]), c0f, __uint_as_float(cc[j]));
            float mainv = __uint_as_float(a0[j]);
            if (NKB == 2) {
              corr = fmaf(__uint_as_float(a1[j]), c1f, corr);
              mainv += __uint_as_float(a1[j]);
            }
            v[j] = mainv + corr;
          }
        }
        bool ok[4];
        int cl[4];                                      // clamped column for loads (N is a multiple of 8)
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          ok[f] = rv && c0 + 4 * f < N;
          cl[f] = min(c0 + 4 * f, N - 4);
        }
        if (pre0) {
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float4 t = ld4(pre0 + pr0 * N + cl[f]);
            v[4 * f] += t.x; v[4 * f + 1] += t.y; v[4 * f + 2] += t.z; v[4 * f + 3] += t.w;
          }
        }
        if (pre1) {
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float4 t = ld4(pre1 + pr1 * N + cl[f]);
            v[4 * f] += t.x; v[4 * f + 1] += t.y; v[4 * f + 2] += t.z; v[4 * f + 3] += t.w;
          }
        }
        if (flags & GEMNET_CHAIN_ACT) {
          if (need_d) {
#pragma unroll
            for (int j = 0; j < 16; ++j) scaled_silu(v[j], v[j], d[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = scaled_silu_value(v[j]);
          }
        } else if (need_d) {
#pragma unroll
          for (int j = 0; j < 16; ++j) d[j] = 1.f;
        }
        if (mul_early) {
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float4 t = ldg4(mul_early + rc * N + cl[f]);
            if (st.aux && ok[f])
              *reinterpret_cast<float4*>(st.aux + r * N + c0 + 4 * f) =
                  make_float4(v[4 * f] * s_acc, v[4 * f + 1] * s_acc, v[4 * f + 2] * s_acc, v[4 * f + 3] * s_acc);
            v[4 * f] *= t.x; v[4 * f + 1] *= t.y; v[4 * f + 2] *= t.z; v[4 * f + 3] *= t.w;
            if (need_d) {
              d[4 * f] *= t.x; d[4 * f + 1] *= t.y; d[4 * f + 2] *= t.z; d[4 * f + 3] *= t.w;
            }
          }
        }
        if (need_d) {  // stored here: the derivative's registers are free for the rest of the program
#pragma unroll
          for (int f = 0; f < 4; ++f)
            if (ok[f])
              *reinterpret_cast<float4*>(st.dact + r * N + c0 + 4 * f) =
                  make_float4(d[4 * f] * s_acc, d[4 * f + 1] * s_acc, d[4 * f + 2] * s_acc, d[4 * f + 3] * s_acc);
        }
        if (flags & GEMNET_CHAIN_CARRY_ADD) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= s_acc;
          float c[16];
          carry_ld(c0, c);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(c[j], s_carry, v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= s_acc;
        }
        if (post0) {
          const float sp = st.s_post0;
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float4 t = ld4(post0 + rc * N + cl[f]);
            v[4 * f] = fmaf(t.x, sp, v[4 * f]); v[4 * f + 1] = fmaf(t.y, sp, v[4 * f + 1]);
            v[4 * f + 2] = fmaf(t.z, sp, v[4 * f + 2]); v[4 * f + 3] = fmaf(t.w, sp, v[4 * f + 3]);
          }
        }
        if (post1) {
          const float sp = st.s_post1;
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float4 t = ld4(post1 + rc * N + cl[f]);
            v[4 * f] = fmaf(t.x, sp, v[4 * f]); v[4 * f + 1] = fmaf(t.y, sp, v[4 * f + 1]);
            v[4 * f + 2] = fmaf(t.z, sp, v[4 * f + 2]); v[4 * f + 3] = fmaf(t.w, sp, v[4 * f + 3]);
          }
        }
        if (mul_late) {
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float4 t = ldg4(mul_late + rc * N + cl[f]);
            v[4 * f] *= t.x; v[4 * f + 1] *= t.y; v[4 * f + 2] *= t.z; v[4 * f + 3] *= t.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = ok[j >> 2] ? v[j] : 0.f;
        if (st.out) {
          const float ss = st.s_store;
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            float4 o = make_float4(v[4 * f] * ss, v[4 * f + 1] * ss, v[4 * f + 2] * ss, v[4 * f + 3] * ss);
            if (st.smul) {
              const float4 t = ldg4(st.smul + rc * N + cl[f]);
              o.x *= t.x; o.y *= t.y; o.z *= t.z; o.w *= t.w;
            }
            if (ok[f]) *reinterpret_cast<float4*>(st.out + orow * N + c0 + 4 * f) = o;
          }
        }
        if (st.dot_w) {
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float4 t = ldg4(st.dot_w + cl[f]);
            dot = fmaf(v[4 * f], t.x, fmaf(v[4 * f + 1], t.y, fmaf(v[4 * f + 2], t.z, fmaf(v[4 * f + 3], t.w, dot))));
          }
        }
        if (flags & GEMNET_CHAIN_SET_CARRY) {
          const float sc = st.s_set_carry;
          float c[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) c[j] = v[j] * sc;
          carry_st(c0, c);
        }
        if ((flags & GEMNET_CHAIN_TO_A) && !last) {
          // two granules of the next operand; its width equals this step's (Kpad_next == Npad)
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            const int col = c0 + 8 * g2;
            const bool gv = rv && col < Knext;
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = gv ? v[8 * g2 + j] * st.s_a : 0.f;
            if (st.next_amul) {
              const int cn = min(col, Knext - 8);
              const float4 m0 = ldg4(st.next_amul + rc * Knext + cn), m1 = ldg4(st.next_amul + rc * Knext + cn + 4);
              w[0] *= m0.x; w[1] *= m0.y; w[2] *= m0.z; w[3] *= m0.w;
              w[4] *= m1.x; w[5] *= m1.y; w[6] *= m1.z; w[7] *= m1.w;
            }
            store_granule(A, row, col >> 3, w);
          }
        }
      }
      if ((flags & GEMNET_CHAIN_A_FROM_CARRY) && !last) patch_from_carry(Knext, st.s_a, st.next_amul, rc, rv);
      stamp();  // step: epilogue of this thread done
      if (st.dot_w) {
        atomicAdd(&dotbuf[row], dot);
        __syncthreads();
        if (cg == 0 && rv) st.dot_out[r] = dotbuf[row] + (st.dot_in ? st.dot_in[r] : 0.f);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace chain
}  // namespace gemnet

using namespace gemnet;

static long long* g_chain_prof = nullptr;
GEMNET_API void gemnet_chain_set_profile_buffer(void* device_int64_x256) {
  g_chain_prof = static_cast<long long*>(device_int64_x256);
}

GEMNET_API int gemnet_chain_tc(const gemnet_chain_program* program_host, int32_t flags, void* stream) {
  GEMNET_REQUIRE(program_host, "null program");
  const gemnet_chain_program& p = *program_host;
  GEMNET_REQUIRE(p.M >= 0, "negative size");
  GEMNET_REQUIRE(p.n_steps >= 1 && p.n_steps <= GEMNET_CHAIN_MAX_STEPS, "chain: 1..GEMNET_CHAIN_MAX_STEPS steps");
  if (p.M == 0) return 0;
  GEMNET_REQUIRE(p.step[0].n_terms > 0, "chain: the first step must load its operand");
  for (int s = 0; s < p.n_steps; ++s) {
    const gemnet_chain_step& st = p.step[s];
    GEMNET_REQUIRE(st.K >= 8 && st.K <= 128 && st.K % 8 == 0 && st.N >= 8 && st.N <= 128 && st.N % 8 == 0,
                   "chain: K, N must be multiples of 8 in [8, 128]");
    GEMNET_REQUIRE(st.Wp && ((uintptr_t)st.Wp % 16) == 0, "chain: packed weight image missing / misaligned");
    GEMNET_REQUIRE(st.n_terms >= 0 && st.n_terms <= GEMNET_CHAIN_MAX_TERMS, "chain: too many operand terms");
    for (int t = 0; t < st.n_terms; ++t)
      GEMNET_REQUIRE(st.term[t].src && (st.term[t].kind == 0 || st.term[t].mul), "chain: operand term without source");
    if (s + 1 < p.n_steps && p.step[s + 1].n_terms == 0 &&
        (st.flags & (GEMNET_CHAIN_TO_A | GEMNET_CHAIN_A_FROM_CARRY))) {
      const int kn = p.step[s + 1].K;
      GEMNET_REQUIRE((st.flags & GEMNET_CHAIN_A_FROM_CARRY) || (kn > 64) == (st.N > 64),
                     "chain: TO_A needs the next step's K to pad like this step's N");
    }
    GEMNET_REQUIRE(!st.dot_w || st.dot_out, "chain: dot_w without dot_out");
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(chain::chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         chain::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("chain_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const int n_tiles = (int)ceil_div(p.M, chain::TM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_tiles < sms ? n_tiles : sms);
  cfg.blockDim = dim3(chain::THREADS);
  cfg.dynamicSmemBytes = chain::SMEM_BYTES;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  const int settled = (flags & GEMNET_DENSE_WEIGHTS_SETTLED) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, chain::chain_tc_kernel, p, n_tiles, settled, g_chain_prof);
  if (e != cudaSuccess) {
    set_error("gemnet_chain_tc: launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return check_launch("gemnet_chain_tc");
}
