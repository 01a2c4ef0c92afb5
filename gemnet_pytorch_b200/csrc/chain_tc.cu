// N1 chain fusion: a whole stack of dense layers with the pointwise work between them in ONE tcgen05 kernel
// (include/gemnet_b200.h, gemnet_chain_tc).  Batch-64 GemNet is latency bound: ~25 dependent launches per
// interaction block in each direction, every one a 128x128 GEMM that finishes in microseconds.  Here a 64-row tile of
// the activation stays on chip for the whole chain.
//
// The GEMM is issued TRANSPOSED: D^T (features x rows) = W (features x K) @ X^T, i.e. the weight image is the M-side
// operand (M = 128 output features = the 128 TMEM lanes) and the activation tile is the N-side operand.  With the
// accumulator in that orientation a warp's 32 lanes are 32 consecutive FEATURES of one row, so every global access of
// the epilogue (stores of the activation derivative and of the outputs, loads of skip / Hadamard / gathered terms) is
// a fully coalesced 128-byte line per warp instruction, and batch 64 (8186 edges) runs on 128 CTAs.
//
//   operand X      shared memory, exact three-way bf16 split (h + m + l == fp32 value), K-major SWIZZLE_128B, 48 KB
//   weights        A operand of the MMA in TENSOR MEMORY (192 columns: 3 splits x K / 2): the tensor pipe then reads
//                  only the activation tile from shared memory and runs at its N / 2-cycle floor per instruction.
//                  The next layer's image streams into a 96 KB staging buffer (cp.async.bulk) while the current
//                  layer runs and is moved to TMEM by four dedicated warps
//   accumulators   TMEM: one main accumulator per K chunk + one correction accumulator (dense_tc.cu's numerics:
//                  six significant products, the toward-zero rounding of the accumulator compensated)
//   skip value     "carry": TMEM columns next to the accumulators (tcgen05.st / tcgen05.ld)
//
// Software pipeline.  A 64-row tile is processed as two 32-row halves: while the worker warps run the epilogue of one
// half (and write that half's rows of the next operand), the tensor pipe runs the other half's GEMM.  Warp roles
// (register budgets re-balanced with setmaxnreg):
//   16 worker warps   operand fill from global memory (coalesced; gathers, Hadamard factors and several summed terms)
//                     and the epilogue of every (step, half): thread (feature f = 32 q + lane, 8 rows) reads its
//                     accumulator values with one tcgen05.ld per accumulator, runs the pointwise program stage by
//                     stage over them, stores what the backward pass needs and writes the next layer's operand back
//                     into shared memory; then the warp arrives on the half's `op_ready` mbarrier.  No block-wide
//                     barrier inside the chain.
//   1 issuer warp     waits for op_ready(half) (16 workers [+ the 4 movers for half 0]), issues the 48 tcgen05.mma of
//                     the (step, half) from ONE ELECTED LANE WITH WARP-UNIFORM OPERANDS (under a thread-variant
//                     predicate ptxas wraps every MMA in an election loop of ~66 cycles), commits to acc_full(half),
//                     and starts the cp.async.bulk of the next weight image.
//   4 mover warps     one TMEM lane quadrant each: staged weight image -> tensor memory as soon as the GEMM that reads
//                     the previous image has completed (they observe acc_full of the last half).
// A 16-row variant (one half, 4 values per thread) serves atom-level chains (762 atoms at batch 64 = 48 tiles).
// The step program is a kernel parameter (__grid_constant__).  One CTA per SM, persistent over row tiles.
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace gemnet {
namespace chain {

using namespace tcx;

// rows per tile: 64 (two pipelined 32-row halves), or 16 when that still gives at most one tile per SM (the atom-level
// chains: 762 atoms at batch 64).  Template parameter TR of the kernel; the shared-memory map is sized for 64.
constexpr int X_BYTES = 2 * 3 * 64 * 128;      // 49152: operand, [chunk][split][row][128 B]
// Warp roles (warp-specialised; register budgets re-balanced with setmaxnreg, warpgroup = 4 warps):
//   warps  0..15  workers: operand fill + epilogue                                  (4 warpgroups, 104 registers)
//   warp   16     issuer: tcgen05.mma of every (step, half) + cp.async.bulk of the weight images
//   warps  20..23 weight movers: staged image -> tensor memory, one TMEM lane quadrant each   (warps 17..19 idle)
constexpr int WORKERS = 512;
constexpr int ISSUER_WARP = WORKERS / 32, MOVER_WARP0 = ISSUER_WARP + 4;
constexpr int THREADS = WORKERS + 256;
constexpr int MISC_BYTES = 2048;               // barriers, TMEM slot, dot buffer (4 quadrants x TR floats)
constexpr int WSTAGE_BYTES = 3 * 64 * 128 * 4;  // 98304: staging buffer of one weight image on its way to tensor memory
constexpr int SMEM_BYTES = 1024 + X_BYTES + WSTAGE_BYTES + MISC_BYTES;
constexpr uint32_t W_COLS = 192;               // weight image in TMEM: 3 splits x (K / 2 <= 64) columns, 128 lanes
// 1: instantiate the FULL epilogue for N / K = 128 and 64 with immediate row strides (fewer address instructions,
// but 1.5x the code: measured slower on B200 — instruction fetch)
#ifndef GEMNET_CHAIN_STRIDE_SPEC
#define GEMNET_CHAIN_STRIDE_SPEC 0
#endif

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// byte offset (inside one split) of element (row, k) of the K-major SWIZZLE_128B operand: chunk k / 64, 16-byte granule
// (k % 64) / 8 XOR-swizzled by the row
template <int TR>
__device__ __forceinline__ uint32_t x_offset(int row, int k) {
  constexpr int CHUNK_BYTES = 3 * TR * 128;
  return (uint32_t)((k >> 6) * CHUNK_BYTES + row * 128 + (((((k & 63) >> 3)) ^ (row & 7)) << 4) + (k & 7) * 2);
}
template <int TR>
__device__ __forceinline__ void store_granule(uint8_t* X, int row, int g, const float (&v)[8]) {
  constexpr int SPLIT_BYTES = TR * 128, CHUNK_BYTES = 3 * SPLIT_BYTES;
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) split3(v[2 * u], v[2 * u + 1], h[u], m[u], l[u]);
  uint8_t* p = X + (g >> 3) * CHUNK_BYTES + row * 128 + (((g & 7) ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(p + SPLIT_BYTES) = make_uint4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint4*>(p + 2 * SPLIT_BYTES) = make_uint4(l[0], l[1], l[2], l[3]);
}
template <int TR>
__device__ __forceinline__ float load_elem(const uint8_t* X, int row, int k) {
  constexpr int SPLIT_BYTES = TR * 128;
  const uint8_t* p = X + x_offset<TR>(row, k);
  const uint32_t h = *reinterpret_cast<const uint16_t*>(p);
  const uint32_t m = *reinterpret_cast<const uint16_t*>(p + SPLIT_BYTES);
  const uint32_t l = *reinterpret_cast<const uint16_t*>(p + 2 * SPLIT_BYTES);
  return (__uint_as_float(h << 16) + __uint_as_float(m << 16)) + __uint_as_float(l << 16);
}

// kSave: the training variant (steps may store their operand, `save_x`); the inference variant compiles it away — the
// extra pointers cost the epilogue ~1k cycles per step in register pressure even when they are NULL.
// what one thread of the epilogue owns: feature f (accumulator row = TMEM lane), tile rows e_loc .. e_loc + VPT - 1
struct ThreadTile {
  int f, nv, jmax, e_loc, xf_g;   // nv: rows of them inside M; jmax = max(nv - 1, 0)
  int64_t rb, rbc;                // first row; the same clamped into [0, M)
  uint8_t* xf_base;               // operand element (tile row e_loc + j, k = f): xf_base + j * 128 + ((xf_g ^ ((e_loc + j) & 7)) << 4)
};

// values of feature f for rows rb .. rb + VPT - 1 -> the next operand (element (row, k = f)); zero beyond Knext / M
// KC: the operand width when it is 128 or 64 (compile-time row stride: every access of the thread is then one base
// pointer + an immediate), 0 = run-time stride
template <bool kSave, int TR, int VPT, int KC>
__device__ __forceinline__ void operand_from(const ThreadTile& t, const float (&vals)[VPT], int Knext_rt, float s_a,
                                             const float* amul, bool full, float* save) {
  constexpr int SPLIT_BYTES = TR * 128;
  const int f = t.f, nv = t.nv, jmax = t.jmax, e_loc = t.e_loc, xf_g = t.xf_g;
  const int64_t rb = t.rb, rbc = t.rbc;
  uint8_t* const xf_base = t.xf_base;
  const int Knext = KC ? KC : Knext_rt;
  const bool kv = f < Knext;
  const int fc = min(f, Knext - 1);
  float w[VPT];
  if (full) {
    if (s_a == 1.0f) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) w[j] = vals[j];
    } else {
#pragma unroll
      for (int j = 0; j < VPT; ++j) w[j] = vals[j] * s_a;
    }
    if (amul) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) w[j] *= __ldg(amul + (rb * Knext + fc) + j * Knext);
    }
    if (!kv) {                                             // (warp-uniform for operand widths of 32, 64, 96, 128)
#pragma unroll
      for (int j = 0; j < VPT; ++j) w[j] = 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < VPT; ++j) w[j] = (kv && j < nv) ? vals[j] * s_a : 0.f;
    if (amul) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) w[j] *= __ldg(amul + (rbc + min(j, jmax)) * Knext + fc);
    }
  }
  if (f < (Knext > 64 ? 128 : 64)) {                    // inside the padded operand width
#pragma unroll
    for (int j = 0; j < VPT; j += 2) {
      uint32_t h, m, l;
      split3(w[j], w[j + 1], h, m, l);
      // (the first row of a thread is a multiple of VPT: for VPT = 8 the row's swizzle term is the constant j & 7)
      const int s0 = VPT % 8 == 0 ? (j & 7) : ((e_loc + j) & 7), s1 = VPT % 8 == 0 ? ((j + 1) & 7) : ((e_loc + j + 1) & 7);
      uint8_t* p0 = xf_base + j * 128 + ((xf_g ^ s0) << 4);
      uint8_t* p1 = xf_base + (j + 1) * 128 + ((xf_g ^ s1) << 4);
      *reinterpret_cast<uint16_t*>(p0) = (uint16_t)h;
      *reinterpret_cast<uint16_t*>(p1) = (uint16_t)(h >> 16);
      *reinterpret_cast<uint16_t*>(p0 + SPLIT_BYTES) = (uint16_t)m;
      *reinterpret_cast<uint16_t*>(p1 + SPLIT_BYTES) = (uint16_t)(m >> 16);
      *reinterpret_cast<uint16_t*>(p0 + 2 * SPLIT_BYTES) = (uint16_t)l;
      *reinterpret_cast<uint16_t*>(p1 + 2 * SPLIT_BYTES) = (uint16_t)(l >> 16);
    }
  }
  if (kSave && save != nullptr && kv) {                 // training: the operand itself (coalesced over f)
#pragma unroll
    for (int j = 0; j < VPT; ++j)
      if (j < nv) save[(rb * Knext + f) + j * Knext] = w[j];
  }
}

template <bool kSave, int TR>
__global__ void __launch_bounds__(THREADS, 1)
chain_tc_kernel(const __grid_constant__ gemnet_chain_program prog, int n_tiles, int weights_settled,
                long long* __restrict__ prof) {
  // A 64-row tile is worked on as two 32-row HALVES, software-pipelined: while the threads run the epilogue of one
  // half (and write its rows of the next operand), the tensor pipe runs the other half's GEMM.
  constexpr int HALVES = TR >= 64 ? 2 : 1;
  constexpr int HR = TR / HALVES;                               // rows per half = N of one MMA
  constexpr int VPT = HR / 4;                                   // values per thread and half: rows of one feature
  static_assert(HR * 16 <= WORKERS, "one operand-fill piece per thread");
  constexpr int SPLIT_BYTES = TR * 128, CHUNK_BYTES = 3 * SPLIT_BYTES;
  // TMEM columns of one half: main accumulator of K chunk 0 / 1, correction accumulator, carry; then the weight image
  constexpr uint32_t ACC_STRIDE = HR, CORR_COL = 2 * HR, CARRY_COL = 3 * HR, HCOLS = 4 * HR, W_COL = 4 * TR;
  constexpr uint32_t TMEM_COLS = 4 * TR + W_COLS <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* X = sm;                                      // operand: [chunk][split][row][128 B]
  const uint32_t x_base = base;
  const uint32_t w_stage = base + X_BYTES;              // weight image in flight: u32 [column group][128 features][4]
  const uint32_t bar_base = w_stage + WSTAGE_BYTES;
  auto acc_full = [&](int hf) { return bar_base + 8u * hf; };   // GEMM of half hf complete (tcgen05.commit)
  const uint32_t w_full = bar_base + 16;                        // weight image landed in the staging buffer
  auto op_ready = [&](int hf) { return bar_base + 40u + 8u * hf; };   // the 16 worker warps finished operand (+ weights)
  const uint32_t tmem_slot = bar_base + 32;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + X_BYTES + WSTAGE_BYTES + 32);
  float* dotbuf = reinterpret_cast<float*>(sm + X_BYTES + WSTAGE_BYTES + 64);  // [4 quadrants][TR] floats

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // the warp index as a warp-uniform value
  // diagnostics (gemnet_chain_set_profile_buffer): thread 0 of CTA 0 records clock64() at the phase boundaries
  int prof_n = 0;
  const bool profiling = prof != nullptr && blockIdx.x == 0 && tid == 0;
  auto stamp = [&]() {
    if (profiling && prof_n < 250) prof[prof_n++] = clock64();
  };
  stamp();
  const int64_t M = prog.M;
  const int n_steps = prog.n_steps;

  // Programmatic dependent launch: the next kernel of the stream may start its prologue now; everything here that
  // touches activations waits in griddep_wait() for the PREVIOUS kernel to have completed and flushed.
  griddep_launch();
  if (tid == 0) {
    mbar_init(acc_full(0), 1);
    mbar_init(acc_full(1), 1);
    mbar_init(w_full, 1);
    mbar_init(op_ready(0), WORKERS / 32 + 4);   // + the four weight movers: this GEMM's image is in TMEM
    mbar_init(op_ready(1), WORKERS / 32);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  stamp();  // 1: prologue done

  const int q = warp & 3, cg = warp >> 2;       // TMEM lane quadrant (features 32q..32q+31), row group (VPT rows of the half)
  const int f = q * 32 + lane;                  // feature (accumulator row = TMEM lane) of this thread

  // Weight image of one GEMM -> tensor memory, where the MMA reads its A operand (M = 128 features = the lanes): the
  // tensor pipe then fetches only the activation tile from shared memory per instruction and runs at its N / 2-cycle
  // floor.  TMEM image (dense_tc.cu pack kernels): u32 [split][K / 8 column groups][128 features][4 columns], column c
  // of a split = bf16 pair (k = 2c, 2c + 1).  The issuer streams the NEXT GEMM's image into the staging buffer with
  // cp.async.bulk as soon as the current image has been moved out of it; the four mover warps copy it into TMEM
  // (conflict-free 16-byte shared loads + tcgen05.st, lane = feature) once the GEMM that read the old image is complete
  // — under the workers' epilogues, off their critical path.
  auto fetch_weights = [&](const gemnet_chain_step& st) {    // (one thread)
    const int NKB = st.K > 64 ? 2 : 1;
    const size_t smem_image = (size_t)NKB * 3 * (st.N > 64 ? 128 : 64) * 128;   // the dense_tc image comes first
    const uint8_t* src = reinterpret_cast<const uint8_t*>(st.Wp) + smem_image;
    const uint32_t split_bytes = 16384u * NKB;
    fence_proxy_async();   // (the movers' generic-proxy reads of the old image were ordered before this by op_ready)
    mbar_arrive_expect_tx(w_full, 3 * split_bytes);
    for (int sp = 0; sp < 3; ++sp) bulk_g2s(w_stage + sp * split_bytes, src + (size_t)sp * split_bytes, split_bytes, w_full);
  };
  if (warp_u == ISSUER_WARP && elect_one()) {
    if (!weights_settled) griddep_wait();  // the packed images may come from the kernel launched just before
    fetch_weights(prog.step[0]);
  }
  griddep_wait();
  auto worker_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory"); };   // (the 16 worker warps)
  stamp();  // 2: predecessor complete
  const int xf_g = (f & 63) >> 3;
  uint32_t done0 = 0, done1 = 0;   // GEMMs of half 0 / 1 whose completion has been consumed (parity of acc_full(half))

  if (warp_u < ISSUER_WARP) {
  // ====================================================================== worker warps
  // (the pool holds only what this CTA's other warps release: 256 threads x (80 - 32) = 512 x 24 registers)
  asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = (int64_t)tile * TR;
    // Software pipeline over (step s, half hf): [epilogue of (s - 1, hf), which writes the half's rows of operand s]
    // -> [operand fill from global memory, if step s loads one] -> barrier -> elected lane issues the MMAs of (s, hf).
    // The GEMM of the other half runs on the tensor pipe meanwhile.  s == n_steps only drains the last epilogues.
#pragma unroll 1
    for (int s = 0; s <= n_steps; ++s) {
#pragma unroll 1
      for (int hf = 0; hf < HALVES; ++hf) {
      // ------------------------------------------------------------------ this thread's part of the half
      const int e_loc = hf * HR + cg * VPT;         // first tile row of this thread's VPT values
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * HCOLS + cg * VPT);
      // element (tile row e_loc + j, k = f) of the operand: xf_base + j * 128 + ((xf_g ^ ((e_loc + j) & 7)) << 4)
      uint8_t* const xf_base = X + (f >> 6) * CHUNK_BYTES + e_loc * 128 + (f & 7) * 2;
      const int64_t rb = r0 + e_loc;                          // first row of this thread's values
      const int nv = (int)max((int64_t)0, min((int64_t)VPT, M - rb));   // valid rows among them
      const int64_t rbc = nv > 0 ? rb : M - 1;                // clamped row base / index for loads (always in range)
      const int jmax = max(nv, 1) - 1;
      const ThreadTile tt{f, nv, jmax, e_loc, xf_g, rb, rbc, xf_base};
      auto carry_ld = [&](float (&c)[VPT]) {
        uint32_t raw[VPT];
        tmem_ld_nowait<VPT>(t_lane + CARRY_COL, raw);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < VPT; ++j) c[j] = __uint_as_float(raw[j]);
      };
      auto carry_st = [&](const float (&c)[VPT]) {
        uint32_t raw[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) raw[j] = __float_as_uint(c[j]);
        tmem_st<VPT>(t_lane + CARRY_COL, raw);
      };

      if (s > 0) {
      // ================================================================== epilogue of GEMM (s - 1, hf)
      const gemnet_chain_step& st = prog.step[s - 1];
      const int K = st.K, N = st.N;
      const int NKB = K > 64 ? 2 : 1;
      const bool last = s == n_steps;
      {
        uint32_t& done = hf ? done1 : done0;
        mbar_wait(acc_full(hf), done & 1u);
        ++done;
      }
      tc_fence_after();
      stamp();  // accumulators of (s - 1, hf) complete
      // ------------------------------------------------------------------ epilogue of this step's GEMM
      // thread = feature f, rows rb .. rb + VPT - 1.  The pointwise program runs STAGE BY STAGE over the VPT values
      // (feature tests outside the element loops): VPT independent dependency chains per stage, one uniform branch
      // per disabled feature, and every global access is a coalesced line per warp (lanes = consecutive features).
      const int flags = st.flags;
      const int Knext = last ? 0 : prog.step[s].K;
      const float c0f = kShrinkPerStep * (float)min(4, (K + 15) / 16);
      const float c1f = kShrinkPerStep * (float)max(0, min(4, (K - 64 + 15) / 16));
      const float s_acc = st.s_acc;
      const bool need_d = st.dact != nullptr;
      if (st.dot_w) {   // idle / partial warps must not leave stale partial sums (this half's rows)
        if (tid < 4 * HR) dotbuf[(tid / HR) * TR + hf * HR + tid % HR] = 0.f;
        worker_sync();
      }
      // a warp whose 32 features lie beyond this step's output AND the next operand has nothing to do
      const bool warp_idle = 32 * q >= N && (last || 32 * q >= (Knext > 64 ? 128 : 64));
      // FULL (warp-uniform): all VPT rows and all 32 features of the warp are in range -> no clamps, no predicates
      // n_tag: N when it is 128 or 64 (FULL only) — compile-time row stride, so the 16 accesses of an array are one
      // base pointer + immediates instead of four address instructions each; 0 = run-time stride
      auto epilogue = [&](auto full_tag, auto n_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        constexpr int NC = decltype(n_tag)::value;
        const int NS = NC ? NC : st.N;     // this step's NS (row stride of its arrays)
        const bool fv = FULL || f < NS;
        const int fc = FULL ? f : min(f, NS - 1);
        const int64_t RB = FULL ? rb : rbc;
        const int64_t off0 = RB * NS + fc;                  // element j at off0 + JJ(j) * N
        auto OKJ = [&](int j) { return FULL || (fv && j < nv); };
        auto JJ = [&](int j) { return FULL ? j : min(j, jmax); };

        float v[VPT], d[VPT];
        {
          uint32_t a0[VPT], a1[VPT], cc[VPT];
          tmem_ld_nowait<VPT>(t_lane, a0);
          tmem_ld_nowait<VPT>(t_lane + CORR_COL, cc);
          if (NKB == 2) tmem_ld_nowait<VPT>(t_lane + ACC_STRIDE, a1);
          tmem_wait_ld();
  #pragma unroll
          for (int j = 0; j < VPT; ++j) {
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
        if (st.pre0) {
  #pragma unroll
          for (int j = 0; j < VPT; ++j) {
            const int64_t rr = RB + JJ(j);
            const int64_t pr = st.pre0_idx ? (int64_t)__ldg(st.pre0_idx + rr) : rr;
            v[j] += __ldg(st.pre0 + pr * NS + fc);
          }
        }
        if (st.pre1) {
  #pragma unroll
          for (int j = 0; j < VPT; ++j) {
            const int64_t rr = RB + JJ(j);
            const int64_t pr = st.pre1_idx ? (int64_t)__ldg(st.pre1_idx + rr) : rr;
            v[j] += __ldg(st.pre1 + pr * NS + fc);
          }
        }
        // the layer's output scale s_acc is folded into the activation's constant (no separate multiplies)
        const bool act = (flags & GEMNET_CHAIN_ACT) != 0;
        if (act) {
          const float kk = kSiluScale * s_acc;
          if (need_d) {
  #pragma unroll
            for (int j = 0; j < VPT; ++j) scaled_silu(v[j], kk, v[j], d[j]);
          } else {
  #pragma unroll
            for (int j = 0; j < VPT; ++j) v[j] = scaled_silu_value(v[j], kk);
          }
        } else {
  #pragma unroll
          for (int j = 0; j < VPT; ++j) v[j] *= s_acc;
          if (need_d) {
  #pragma unroll
            for (int j = 0; j < VPT; ++j) d[j] = s_acc;
          }
        }
        if (st.mul_early) {
          const float* pm = st.mul_early + off0;
          if (st.aux) {
            float* pa = st.aux + off0;
  #pragma unroll
            for (int j = 0; j < VPT; ++j)
              if (OKJ(j)) pa[j * NS] = v[j];             // (already times s_acc)
          }
  #pragma unroll
          for (int j = 0; j < VPT; ++j) {
            const float t = __ldg(pm + JJ(j) * NS);
            v[j] *= t;
            if (need_d) d[j] *= t;
          }
        }
        if (need_d) {  // stored here: the derivative's registers are free for the rest of the program
          float* pd = st.dact + off0;
  #pragma unroll
          for (int j = 0; j < VPT; ++j)
            if (OKJ(j)) pd[j * NS] = d[j];
        }
        if (flags & GEMNET_CHAIN_CARRY_ADD) {
          float c[VPT];
          carry_ld(c);
          const float sc = st.s_carry;
  #pragma unroll
          for (int j = 0; j < VPT; ++j) v[j] = fmaf(c[j], sc, v[j]);
        }
        if (st.post0) {
          const float sp = st.s_post0;
          const float* pp = st.post0 + off0;
  #pragma unroll
          for (int j = 0; j < VPT; ++j) v[j] = fmaf(__ldg(pp + JJ(j) * NS), sp, v[j]);
        }
        if (st.post1) {
          const float sp = st.s_post1;
          const float* pp = st.post1 + off0;
  #pragma unroll
          for (int j = 0; j < VPT; ++j) v[j] = fmaf(__ldg(pp + JJ(j) * NS), sp, v[j]);
        }
        if (st.mul_late) {
          const float* pm = st.mul_late + off0;
  #pragma unroll
          for (int j = 0; j < VPT; ++j) v[j] *= __ldg(pm + JJ(j) * NS);
        }
  #pragma unroll
        for (int j = 0; j < VPT; ++j) v[j] = (OKJ(j)) ? v[j] : 0.f;
        if (st.out) {
          const float ss = st.s_store;
          float o[VPT];
  #pragma unroll
          for (int j = 0; j < VPT; ++j) o[j] = v[j] * ss;
          if (st.smul) {
            const float* pm = st.smul + off0;
  #pragma unroll
            for (int j = 0; j < VPT; ++j) o[j] *= __ldg(pm + JJ(j) * NS);
          }
          if (st.out_idx) {
            const int* pi = st.out_idx + RB;
            float* po = st.out + f;
  #pragma unroll
            for (int j = 0; j < VPT; ++j) {
              const int64_t orow = (int64_t)__ldg(pi + JJ(j));
              if (OKJ(j)) po[orow * NS] = o[j];
            }
          } else {
            float* po = st.out + off0;
  #pragma unroll
            for (int j = 0; j < VPT; ++j)
              if (OKJ(j)) po[j * NS] = o[j];
          }
        }
        if (st.dot_w) {
          const float wv = __ldg(st.dot_w + fc);
  #pragma unroll
          for (int j = 0; j < VPT; ++j) {
            float p = v[j] * wv;                       // (0 for invalid features / rows)
  #pragma unroll
            for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
            if (lane == 0) dotbuf[q * TR + e_loc + j] = p;   // one writer per (quadrant, row): deterministic
          }
        }
        if (flags & GEMNET_CHAIN_SET_CARRY) {
          const float sc = st.s_set_carry;
          if (sc == 1.0f) {
            carry_st(v);
          } else {
            float c[VPT];
  #pragma unroll
            for (int j = 0; j < VPT; ++j) c[j] = v[j] * sc;
            carry_st(c);
          }
        }
        if (!last && (flags & (GEMNET_CHAIN_TO_A | GEMNET_CHAIN_A_FROM_CARRY))) {
          if (!(flags & GEMNET_CHAIN_TO_A)) carry_ld(v);        // (v is dead after this point)
          float* save = kSave ? prog.step[s].save_x : nullptr;
          if constexpr (NC != 0) {
            if (Knext == 128) operand_from<kSave, TR, VPT, 128>(tt, v, Knext, st.s_a, st.next_amul, true, save);
            else if (Knext == 64) operand_from<kSave, TR, VPT, 64>(tt, v, Knext, st.s_a, st.next_amul, true, save);
            else operand_from<kSave, TR, VPT, 0>(tt, v, Knext, st.s_a, st.next_amul, true, save);
          } else {
            operand_from<kSave, TR, VPT, 0>(tt, v, Knext, st.s_a, st.next_amul, FULL, save);
          }
        }
      };
      if (!warp_idle) {
        const bool full = nv == VPT && 32 * q + 32 <= N;
#if GEMNET_CHAIN_STRIDE_SPEC
        if (full && N == 128) epilogue(std::true_type{}, std::integral_constant<int, 128>{});
        else if (full && N == 64) epilogue(std::true_type{}, std::integral_constant<int, 64>{});
        else epilogue(std::false_type{}, std::integral_constant<int, 0>{});
#else
        if (full) epilogue(std::true_type{}, std::integral_constant<int, 0>{});
        else epilogue(std::false_type{}, std::integral_constant<int, 0>{});
#endif
      }
      stamp();  // step: epilogue of this thread done
      if (st.dot_w) {
        worker_sync();
        const int dr = hf * HR + tid;                       // (tile row)
        if (tid < HR && r0 + dr < M)
          st.dot_out[r0 + dr] = ((dotbuf[dr] + dotbuf[TR + dr]) + (dotbuf[2 * TR + dr] + dotbuf[3 * TR + dr])) +
                                (st.dot_in ? st.dot_in[r0 + dr] : 0.f);
      }
      }
      if (s == n_steps) continue;

      const gemnet_chain_step& st = prog.step[s];
      const int K = st.K;
      const int NKB = K > 64 ? 2 : 1;
      // ------------------------------------------------------------------ operand fill from global memory
      if (st.n_terms > 0) {
        const int gpr = 8 * NKB;                                   // granules (8 columns) per operand row
        const bool fold_amul = st.amul != nullptr && !st.load_set_carry;
        // the fill IS the operand unless a carry read-back multiplies it afterwards (then operand_from stores it)
        const bool save_here = kSave && st.save_x != nullptr && !(st.load_set_carry && st.amul != nullptr);
        // (row, granule) pieces: 32 rows x 8|16 granules = 256|512, one per thread, all loads in flight together
        auto fill = [&](auto pg_tag) {
          constexpr int PG = decltype(pg_tag)::value;
          float v[PG][8];
          int prow[PG], pgr[PG];
#pragma unroll
          for (int h = 0; h < PG; ++h) {
            const int p = min(tid + h * WORKERS, HR * gpr - 1);   // (threads beyond the half's pieces: see `pin`)
            prow[h] = hf * HR + p / gpr;
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
            const bool pin = tid + h * WORKERS < HR * gpr;       // this thread owns a piece of the half
            const bool valid = r0 + prow[h] < M && pgr[h] * 8 < K;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[h][j] = valid ? v[h][j] : 0.f;
            if (pin) store_granule<TR>(X, prow[h], pgr[h], v[h]);
            if (save_here && valid && pin) {
              float4* dst = reinterpret_cast<float4*>(st.save_x + (r0 + prow[h]) * K + pgr[h] * 8);
              dst[0] = make_float4(v[h][0], v[h][1], v[h][2], v[h][3]);
              dst[1] = make_float4(v[h][4], v[h][5], v[h][6], v[h][7]);
            }
          }
        };
        fill(std::integral_constant<int, 1>{});
        if (st.load_set_carry) {
          // the skip value is wanted in the epilogue's thread mapping: read it back from the (exact) operand split
          worker_sync();
          float c[VPT];
          const bool kin = f < 64 * NKB;
#pragma unroll
          for (int j = 0; j < VPT; ++j) c[j] = kin ? load_elem<TR>(X, e_loc + j, min(f, 64 * NKB - 1)) : 0.f;
          carry_st(c);
          if (st.amul) operand_from<kSave, TR, VPT, 0>(tt, c, K, 1.0f, st.amul, false, kSave ? st.save_x : nullptr);   // (each thread rewrites exactly the elements it read)
        }
      }
      stamp();  // step: operand written by this thread
      // ------------------------------------------------------------------ hand the half to the issuer warp
      fence_proxy_async();   // generic-proxy writes of X -> visible to the tensor-core (async) proxy
      tc_fence_before();     // this thread's tcgen05.ld / st are done
      __syncwarp();
      if (lane == 0) mbar_arrive(op_ready(hf));
      stamp();  // operand of (s, hf) handed over
      }
    }
  }
  } else {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
  if (warp_u == ISSUER_WARP) {
  // ====================================================================== issuer warp: MMAs and weight fetches
  // (converged; every operand is warp-uniform, one elected lane issues: see elect_one)
  const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem_base, 0);
  uint32_t ready0 = 0, ready1 = 0;   // operands of half 0 / 1 consumed (parity of op_ready(half))
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
    for (int s = 0; s < n_steps; ++s) {
      const gemnet_chain_step& st = prog.step[s];
      const int NKB = st.K > 64 ? 2 : 1;
      // the NEXT GEMM (of this tile, or the first of the CTA's next tile)
      const bool more = s + 1 < n_steps || tile + (int)gridDim.x < n_tiles;
      const gemnet_chain_step& nst = prog.step[s + 1 < n_steps ? s + 1 : 0];
#pragma unroll 1
      for (int hf = 0; hf < HALVES; ++hf) {
        {
          uint32_t& ready = hf ? ready1 : ready0;
          mbar_wait_idle(op_ready(hf), ready & 1u);     // every worker warp: operand rows (+ weight image) written
          ++ready;
        }
        tc_fence_after();
        const uint32_t idesc = make_idesc(128, HR);     // D^T of this half: 128 features x HR rows
        const uint32_t d_corr = tm_u + hf * HCOLS + CORR_COL;
        if (elect_one()) {
          const uint32_t wsp = 32u * NKB;                 // TMEM columns per split of the weight image (K / 2)
          for (int kb = 0; kb < NKB; ++kb) {
            const uint32_t xa = x_base + kb * CHUNK_BYTES + hf * (HR * 128);
            const uint64_t x_h = make_desc(xa), x_m = make_desc(xa + SPLIT_BYTES), x_l = make_desc(xa + 2 * SPLIT_BYTES);
            const uint32_t w_h = tm_u + W_COL + kb * 32u, w_m = w_h + wsp, w_l = w_m + wsp;
            const uint32_t d_main = tm_u + hf * HCOLS + kb * ACC_STRIDE;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 32 >> 4);  // 16 bf16 = 32 bytes along K inside the swizzle atom
              const uint32_t wadv = 8u * kk;                  // ... = 8 TMEM columns of the weight image
              umma_bf16_ts(d_main, w_h + wadv, x_h + adv, idesc, kk != 0);
              umma_bf16_ts(d_corr, w_h + wadv, x_m + adv, idesc, (kb | kk) != 0);
              umma_bf16_ts(d_corr, w_m + wadv, x_h + adv, idesc, 1u);
              umma_bf16_ts(d_corr, w_m + wadv, x_m + adv, idesc, 1u);
              umma_bf16_ts(d_corr, w_h + wadv, x_l + adv, idesc, 1u);
              umma_bf16_ts(d_corr, w_l + wadv, x_h + adv, idesc, 1u);
            }
          }
          umma_commit(acc_full(hf));
          // the staging buffer is free (the movers arrived on op_ready(0) after copying this GEMM's image out)
          if (hf == 0 && more) fetch_weights(nst);
        }
        __syncwarp();
      }
    }
  }
  } else if (warp_u >= MOVER_WARP0) {
  // ====================================================================== weight movers: staging buffer -> TMEM
  const int qd = warp_u & 3;                                   // TMEM lane quadrant of this warp
  const uint4* src0 = reinterpret_cast<const uint4*>(sm + X_BYTES) + (qd * 32 + lane);
  const uint32_t t0 = tmem_base + ((uint32_t)(qd * 32) << 16) + W_COL;
  uint32_t g = 0;                                              // GEMM images moved so far (parities of w_full / acc_full)
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
    for (int s = 0; s < n_steps; ++s, ++g) {
      const int groups = prog.step[s].K > 64 ? 48 : 24;        // column groups of 4 (3 splits x K / 8)
      if (g > 0) {   // the GEMM that reads the previous image: its last half must be complete
        mbar_wait_idle(acc_full(HALVES - 1), (g - 1) & 1u);
        tc_fence_after();
      }
      mbar_wait_idle(w_full, g & 1u);
#pragma unroll 1
      for (int i = 0; i < groups; i += 8) {
        uint4 r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = src0[(size_t)(i + j) * 128];
#pragma unroll
        for (int j = 0; j < 8; ++j) tmem_st4_nowait(t0 + 4u * (i + j), r[j]);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(op_ready(0));
    }
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

static int g_chain_rows = 0;
static long long* g_chain_prof = nullptr;
namespace gemnet { void segbil_fwd_tc_set_profile_buffer(long long* p); }
GEMNET_API int gemnet_chain_set_tile_rows(int rows) {
  GEMNET_REQUIRE(rows == 0 || rows == 16 || rows == 64, "chain: tile rows must be 0 (automatic), 16 or 64");
  g_chain_rows = rows;
  return 0;
}

GEMNET_API void gemnet_chain_set_profile_buffer(void* device_int64_x256) {
  g_chain_prof = static_cast<long long*>(device_int64_x256);
  gemnet::segbil_fwd_tc_set_profile_buffer(g_chain_prof);   // the same buffer serves the H7 forward kernel's stamps
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
    GEMNET_REQUIRE(!st.dot_w || st.dot_out, "chain: dot_w without dot_out");
  }
  bool save = false;
  for (int s = 0; s < p.n_steps; ++s) save |= p.step[s].save_x != nullptr;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(chain::chain_tc_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         chain::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(chain::chain_tc_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(chain::chain_tc_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(chain::chain_tc_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("chain_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  // 16-row tiles while they still fit one per SM (atom-level chains): same MMA time per step, a quarter of the epilogue
  const bool small = g_chain_rows ? g_chain_rows == 16 : ceil_div(p.M, 16) <= sms;
  const int n_tiles = (int)ceil_div(p.M, small ? 16 : 64);
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
  cudaError_t e;
  if (small)
    e = save ? cudaLaunchKernelEx(&cfg, chain::chain_tc_kernel<true, 16>, p, n_tiles, settled, g_chain_prof)
             : cudaLaunchKernelEx(&cfg, chain::chain_tc_kernel<false, 16>, p, n_tiles, settled, g_chain_prof);
  else
    e = save ? cudaLaunchKernelEx(&cfg, chain::chain_tc_kernel<true, 64>, p, n_tiles, settled, g_chain_prof)
             : cudaLaunchKernelEx(&cfg, chain::chain_tc_kernel<false, 64>, p, n_tiles, settled, g_chain_prof);
  if (e != cudaSuccess) {
    set_error("gemnet_chain_tc: launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return check_launch("gemnet_chain_tc");
}
