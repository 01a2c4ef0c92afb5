// H7: the ragged (CSR) segmented bilinear aggregation — fp32 CUDA-core implementation (algo 0).
//
//   out[e,o] = scale * sum_{t in seg(e)} sum_{c,i,s} x[expand[t],c] sph[t,s] rbf_W1[e,i,s] W[c,i,o]
//
// evaluated in the reference's contraction order (efficient.py:159-189) but without padding and without
// materialising the gathered (T,C) operand:
//   phase 1  A_e[s,c]  = sum_t sph[t,s] x[expand[t],c]            one warp per edge, lanes over c, registers
//   phase 2  B_e[i,c]  = sum_s rbf_W1[e,i,s] A_e[s,c]             registers -> shared memory row  Bs[e][i*C+c]
//   phase 3  out[e,:]  = scale * Bs[e,:] (1 x K) @ W' (K x O)     K = C*I, CTA-level tile GEMM (EB edges per CTA)
// The tcgen05 variant (algo 1, segbilinear_tc.cu) replaces phase 3 by a tensor-core GEMM.
#include "common.cuh"

namespace gemnet {

template <int S_, int C_, int I_, int O_, int EB_>
struct SegCfg {
  static constexpr int S = S_, C = C_, I = I_, O = O_, EB = EB_;
  static constexpr int K = C * I;
  static constexpr int CC = C / 32;      // columns of x per lane
  static constexpr int LD = K + 4;       // padded row of Bs (keeps float4 alignment, staggers banks)
  static constexpr int TX = O / 4;       // threads across o (float4 each) in phase 3
  static constexpr int THREADS = 256;
  static constexpr int WARPS = THREADS / 32;
  static constexpr int SP = (S + 3) / 4 * 4;    // basis row padded to a multiple of 4 floats in shared memory
  static constexpr int SB = 8;                  // units staged per batch and warp
  static constexpr int STAGE_FLOATS = WARPS * SB * (SP + 1);  // + the row index of the unit
  static_assert(C % 32 == 0, "C must be a multiple of 32");
  static_assert(O % 4 == 0 && (THREADS / TX) == EB, "EB must equal 256 / (O/4)");
  static_assert((C & (C - 1)) == 0, "C must be a power of two");
  static_assert(I % SB == 0, "I must be a multiple of the stage depth");
};

// Stage the basis rows (and gather indices) of up to SB units of this warp's segment in shared memory: the rows of
// consecutive units are contiguous in global memory, so the stage is one coalesced copy, and the per-unit broadcast
// reads become 128-bit shared loads (4 values / instruction) instead of one global load (+ descriptor moves) per
// value.  The copy is software-pipelined through registers: fetch() issues all loads of the NEXT batch (independent,
// so they overlap each other and the FMAs of the current batch), commit() parks them in shared memory.
// stage layout per warp: [SB][SP] basis values, then [SB] row indices (as float bit patterns).
template <typename Cfg>
struct UnitStager {
  static constexpr int NL = (Cfg::SB * Cfg::S + 31) / 32;
  float v[NL];
  int32_t idx;

  __device__ __forceinline__ void fetch(const int32_t* __restrict__ expand, const float* __restrict__ sph, int32_t t0,
                                        int32_t end, int lane) {
    const int n = min(Cfg::SB, end - t0) * Cfg::S;
    const float* src = sph + (int64_t)t0 * Cfg::S;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int i = lane + 32 * j;
      v[j] = i < n ? __ldg(src + i) : 0.f;
    }
    idx = (expand != nullptr && lane < Cfg::SB && t0 + lane < end) ? __ldg(expand + t0 + lane) : 0;
  }

  __device__ __forceinline__ void commit(float* __restrict__ st, int lane) const {
    __syncwarp();  // the previous batch has been consumed
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int i = lane + 32 * j;
      if (i < Cfg::SB * Cfg::S) {
        const int u = i / Cfg::S, s = i - u * Cfg::S;
        st[u * Cfg::SP + s] = v[j];
      }
    }
    if (lane < Cfg::SB) st[Cfg::SB * Cfg::SP + lane] = __int_as_float(idx);
    __syncwarp();
  }
};

// phase 1: A[s][cc] for edge segment [beg, end)
template <typename Cfg>
__device__ __forceinline__ void segment_sum(const float* __restrict__ x, const int32_t* __restrict__ expand,
                                            const float* __restrict__ sph, int32_t beg, int32_t end, int lane,
                                            float* __restrict__ st, float (&A)[Cfg::S][Cfg::CC]) {
#pragma unroll
  for (int s = 0; s < Cfg::S; ++s)
#pragma unroll
    for (int cc = 0; cc < Cfg::CC; ++cc) A[s][cc] = 0.f;
  UnitStager<Cfg> stg;
  if (beg < end) stg.fetch(expand, sph, beg, end, lane);
  for (int32_t t0 = beg; t0 < end; t0 += Cfg::SB) {
    const int nb = min(Cfg::SB, end - t0);
    stg.commit(st, lane);
    if (t0 + Cfg::SB < end) stg.fetch(expand, sph, t0 + Cfg::SB, end, lane);
    // issue the gathers of the whole batch first (SB independent L2 requests in flight per lane), then the FMAs
    float xb[Cfg::SB][Cfg::CC];
#pragma unroll
    for (int u = 0; u < Cfg::SB; ++u) {
      if (u < nb) {
        const float* xr = x + (int64_t)__float_as_int(st[Cfg::SB * Cfg::SP + u]) * Cfg::C + lane;
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) xb[u][cc] = __ldg(xr + 32 * cc);
      }
    }
#pragma unroll
    for (int u = 0; u < Cfg::SB; ++u) {
      if (u < nb) {
        const float4* sp4 = reinterpret_cast<const float4*>(st + u * Cfg::SP);
#pragma unroll
        for (int s4 = 0; s4 < Cfg::SP / 4; ++s4) {
          const float4 w = sp4[s4];
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (4 * s4 + j < Cfg::S) {
#pragma unroll
              for (int cc = 0; cc < Cfg::CC; ++cc)
                A[4 * s4 + j][cc] = fmaf(wv[j], xb[u][cc], A[4 * s4 + j][cc]);
            }
          }
        }
      }
    }
  }
}

template <typename Cfg, bool kFromA>
__global__ void __launch_bounds__(256) segbil_fwd_kernel(const float* __restrict__ x,
                                                         const int32_t* __restrict__ expand,
                                                         const int32_t* __restrict__ rowptr,
                                                         const float* __restrict__ sph,
                                                         const float* __restrict__ rbf_W1,
                                                         const float* __restrict__ W, float scale, int64_t n_edges,
                                                         const float* __restrict__ A_in, float* __restrict__ out) {
  pdl_enter();
  extern __shared__ __align__(16) float smem[];
  float* Bs = smem;  // [EB][LD]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* st = smem + Cfg::EB * Cfg::LD + warp * (Cfg::SB * (Cfg::SP + 1));  // per-warp unit stage
  const int64_t e0 = (int64_t)blockIdx.x * Cfg::EB;

  for (int el = warp; el < Cfg::EB; el += Cfg::WARPS) {
    const int64_t e = e0 + el;
    float* brow = Bs + el * Cfg::LD;
    if (e >= n_edges) {
      for (int k = lane; k < Cfg::K; k += 32) brow[k] = 0.f;
      continue;
    }
    float A[Cfg::S][Cfg::CC];
    if constexpr (kFromA) {  // split pipeline: phase 1 was done by segsum_kernel
#pragma unroll
      for (int s = 0; s < Cfg::S; ++s)
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) A[s][cc] = __ldg(A_in + e * (Cfg::S * Cfg::C) + s * Cfg::C + lane + 32 * cc);
    } else {
      segment_sum<Cfg>(x, expand, sph, rowptr[e], rowptr[e + 1], lane, st, A);
    }
    // phase 2 with the (I,S) slab of rbf_W1 streamed through the same per-warp stage, SB rows of i at a time
    const float* rw = rbf_W1 + e * (Cfg::I * Cfg::S);
    UnitStager<Cfg> rs;
    rs.fetch(nullptr, rw, 0, Cfg::I, lane);
#pragma unroll 1
    for (int i0 = 0; i0 < Cfg::I; i0 += Cfg::SB) {
      rs.commit(st, lane);
      if (i0 + Cfg::SB < Cfg::I) rs.fetch(nullptr, rw, i0 + Cfg::SB, Cfg::I, lane);
#pragma unroll
      for (int u = 0; u < Cfg::SB; ++u) {
        float b[Cfg::CC];
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) b[cc] = 0.f;
        const float4* w4 = reinterpret_cast<const float4*>(st + u * Cfg::SP);
#pragma unroll
        for (int s4 = 0; s4 < Cfg::SP / 4; ++s4) {
          const float4 w = w4[s4];
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (4 * s4 + j < Cfg::S) {
#pragma unroll
              for (int cc = 0; cc < Cfg::CC; ++cc) b[cc] = fmaf(wv[j], A[4 * s4 + j][cc], b[cc]);
            }
          }
        }
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) brow[(i0 + u) * Cfg::C + lane + 32 * cc] = b[cc];
      }
    }
  }
  __syncthreads();

  // phase 3: thread (el, oq) computes out[e0+el][4*oq .. 4*oq+3]
  const int el = threadIdx.x / Cfg::TX, oq = threadIdx.x % Cfg::TX;
  const float* brow = Bs + el * Cfg::LD;
  const float4* W4 = reinterpret_cast<const float4*>(W);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int i = 0; i < Cfg::I; ++i) {
#pragma unroll 8
    for (int c = 0; c < Cfg::C; ++c) {
      const float a = brow[i * Cfg::C + c];
      const float4 w = __ldg(W4 + (int64_t)(c * Cfg::I + i) * Cfg::TX + oq);
      acc.x = fmaf(a, w.x, acc.x);
      acc.y = fmaf(a, w.y, acc.y);
      acc.z = fmaf(a, w.z, acc.z);
      acc.w = fmaf(a, w.w, acc.w);
    }
  }
  const int64_t e = e0 + el;
  if (e < n_edges)
    reinterpret_cast<float4*>(out)[e * Cfg::TX + oq] =
        make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
}

// Reduce N (power of two <= 32) per-lane values across the warp with log-depth "halving" exchanges:
// afterwards v[0] of lane L holds the full sum of value index (L >> (5 - log2 N)) & (N-1).
template <int N>
__device__ __forceinline__ float warp_multi_reduce(float (&v)[N], int lane) {
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < n / 2; ++k) {
      const float lo = v[k], hi = v[k + n / 2];
      const float send = up ? lo : hi;
      const float keep = up ? hi : lo;
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

// backward w.r.t. x (atomics), sph, rbf_W1.  Same tiling as the forward kernel.
template <typename Cfg>
__global__ void __launch_bounds__(256) segbil_bwd_kernel(const float* __restrict__ x,
                                                         const int32_t* __restrict__ expand,
                                                         const int32_t* __restrict__ rowptr,
                                                         const float* __restrict__ sph,
                                                         const float* __restrict__ rbf_W1,
                                                         const float* __restrict__ W, float scale, int64_t n_edges,
                                                         const float* __restrict__ g_out, float* __restrict__ g_x,
                                                         float* __restrict__ g_sph, float* __restrict__ g_rbf_W1) {
  pdl_enter();
  extern __shared__ __align__(16) float smem[];
  float* dBs = smem;                        // [EB][LD]
  float* gs = smem + Cfg::EB * Cfg::LD;     // [EB][O]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* st = gs + Cfg::EB * Cfg::O + warp * (Cfg::SB * (Cfg::SP + 1));  // per-warp unit stage
  const int64_t e0 = (int64_t)blockIdx.x * Cfg::EB;

  for (int idx = threadIdx.x; idx < Cfg::EB * Cfg::O; idx += Cfg::THREADS) {
    const int64_t e = e0 + idx / Cfg::O;
    gs[idx] = e < n_edges ? g_out[e * Cfg::O + (idx % Cfg::O)] * scale : 0.f;
  }
  __syncthreads();

  // phase A: dB[e][i*C+c] = sum_o g[e,o] W[c,i,o]; thread owns k' = tid + 256 j for all EB edges
  {
    const float4* W4 = reinterpret_cast<const float4*>(W);
    const float4* gs4 = reinterpret_cast<const float4*>(gs);
    for (int kp = threadIdx.x; kp < Cfg::K; kp += Cfg::THREADS) {
      const int c = kp & (Cfg::C - 1), i = kp / Cfg::C;
      const float4* wrow = W4 + (int64_t)(c * Cfg::I + i) * Cfg::TX;
      float acc[Cfg::EB];
#pragma unroll
      for (int el = 0; el < Cfg::EB; ++el) acc[el] = 0.f;
#pragma unroll 2
      for (int oq = 0; oq < Cfg::TX; ++oq) {
        const float4 w = __ldg(wrow + oq);
#pragma unroll
        for (int el = 0; el < Cfg::EB; ++el) {
          const float4 g = gs4[el * Cfg::TX + oq];
          acc[el] = fmaf(g.x, w.x, fmaf(g.y, w.y, fmaf(g.z, w.z, fmaf(g.w, w.w, acc[el]))));
        }
      }
#pragma unroll
      for (int el = 0; el < Cfg::EB; ++el) dBs[el * Cfg::LD + kp] = acc[el];
    }
  }
  __syncthreads();

  // phase B: per edge (one warp each)
  constexpr int IS = Cfg::I * Cfg::S;
  constexpr int SP = Cfg::S <= 8 ? 8 : (Cfg::S <= 16 ? 16 : 32);   // group size for the g_sph reductions
  for (int el = warp; el < Cfg::EB; el += Cfg::WARPS) {
    const int64_t e = e0 + el;
    if (e >= n_edges) continue;
    const int32_t beg = rowptr[e], end = rowptr[e + 1];
    const float* drow = dBs + el * Cfg::LD;
    const float* rw = rbf_W1 + e * IS;
    float A[Cfg::S][Cfg::CC];
    segment_sum<Cfg>(x, expand, sph, beg, end, lane, st, A);

    // dA[s,c] = sum_i rbf_W1[e,i,s] dB[i,c];   g_rbf_W1[e,i,s] = sum_c dB[i,c] A[s,c]
    float dA[Cfg::S][Cfg::CC];
#pragma unroll
    for (int s = 0; s < Cfg::S; ++s)
#pragma unroll
      for (int cc = 0; cc < Cfg::CC; ++cc) dA[s][cc] = 0.f;
    constexpr int SHIFT = SP == 8 ? 2 : (SP == 16 ? 1 : 0);
    UnitStager<Cfg> rs;
    rs.fetch(nullptr, rw, 0, Cfg::I, lane);
#pragma unroll 1
    for (int i0 = 0; i0 < Cfg::I; i0 += Cfg::SB) {
     rs.commit(st, lane);
     if (i0 + Cfg::SB < Cfg::I) rs.fetch(nullptr, rw, i0 + Cfg::SB, Cfg::I, lane);
#pragma unroll 1
     for (int u = 0; u < Cfg::SB; ++u) {
      const int i = i0 + u;
      const float* wrow = st + u * Cfg::SP;
      float d[Cfg::CC];
#pragma unroll
      for (int cc = 0; cc < Cfg::CC; ++cc) d[cc] = drow[i * Cfg::C + lane + 32 * cc];
#pragma unroll
      for (int s = 0; s < Cfg::S; ++s) {
        const float w = wrow[s];
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) dA[s][cc] = fmaf(w, d[cc], dA[s][cc]);
      }
      if (g_rbf_W1) {
#pragma unroll
        for (int s0 = 0; s0 < Cfg::S; s0 += SP) {
          float v[SP];
#pragma unroll
          for (int k = 0; k < SP; ++k) {
            float p = 0.f;
            if (s0 + k < Cfg::S) {
#pragma unroll
              for (int cc = 0; cc < Cfg::CC; ++cc) p = fmaf(d[cc], A[(s0 + k) < Cfg::S ? (s0 + k) : 0][cc], p);
            }
            v[k] = p;
          }
          const float tot = warp_multi_reduce<SP>(v, lane);
          const int k = lane >> SHIFT;
          if ((lane & ((1 << SHIFT) - 1)) == 0 && s0 + k < Cfg::S) g_rbf_W1[e * IS + i * Cfg::S + s0 + k] = tot;
        }
      }
     }
    }

    // per unit t of the segment: g_sph[t,s] = sum_c dA[s,c] x[exp,c];  g_x[exp,c] += sum_s sph[t,s] dA[s,c]
    UnitStager<Cfg> stg;
    if (beg < end) stg.fetch(expand, sph, beg, end, lane);
    for (int32_t t0 = beg; t0 < end; t0 += Cfg::SB) {
     const int nb = min(Cfg::SB, end - t0);
     stg.commit(st, lane);
     if (t0 + Cfg::SB < end) stg.fetch(expand, sph, t0 + Cfg::SB, end, lane);
     float xb[Cfg::SB][Cfg::CC];
#pragma unroll
     for (int u = 0; u < Cfg::SB; ++u) {
       if (u < nb) {
         const float* xr = x + (int64_t)__float_as_int(st[Cfg::SB * Cfg::SP + u]) * Cfg::C + lane;
#pragma unroll
         for (int cc = 0; cc < Cfg::CC; ++cc) xb[u][cc] = __ldg(xr + 32 * cc);
       }
     }
#pragma unroll
     for (int u = 0; u < Cfg::SB; ++u) {
      if (u >= nb) break;
      const int32_t t = t0 + u;
      const int64_t row = (int64_t)__float_as_int(st[Cfg::SB * Cfg::SP + u]) * Cfg::C + lane;
      float xv[Cfg::CC], gx[Cfg::CC];
#pragma unroll
      for (int cc = 0; cc < Cfg::CC; ++cc) {
        xv[cc] = xb[u][cc];
        gx[cc] = 0.f;
      }
      const float* sp = st + u * Cfg::SP;
#pragma unroll
      for (int s0 = 0; s0 < Cfg::S; s0 += SP) {
        float v[SP];
#pragma unroll
        for (int k = 0; k < SP; ++k) {
          float p = 0.f;
          if (s0 + k < Cfg::S) {
            const float w = sp[s0 + k];
#pragma unroll
            for (int cc = 0; cc < Cfg::CC; ++cc) {
              p = fmaf(dA[(s0 + k) < Cfg::S ? (s0 + k) : 0][cc], xv[cc], p);
              gx[cc] = fmaf(w, dA[(s0 + k) < Cfg::S ? (s0 + k) : 0][cc], gx[cc]);
            }
          }
          v[k] = p;
        }
        if (g_sph) {
          const float tot = warp_multi_reduce<SP>(v, lane);
          const int k = lane >> SHIFT;
          if ((lane & ((1 << SHIFT) - 1)) == 0 && s0 + k < Cfg::S) g_sph[(int64_t)t * Cfg::S + s0 + k] = tot;
        }
      }
      if (g_x) {
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) atomicAdd(g_x + row + 32 * cc, gx[cc]);
      }
     }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Split pipeline (workspace given): the gather-bound parts run in their own small-footprint kernels — one warp per
// edge, 4 warps per CTA, no big shared-memory tile — so that ~4x more warps are resident to hide the gather latency
// and the block scheduler balances the very uneven segment lengths; the GEMM-like parts keep the EB-edge tiling.
//   forward :  segsum (A -> workspace)  ->  segbil_fwd_kernel<Cfg, true>  (phases 2 + 3 from A)
//   backward:  segsum (A)  ->  segbil_bwd_edges_kernel (dB, dA -> workspace, g_rbf_W1)  ->  segbil_bwd_units_kernel
constexpr int kSplitWarps = 4;

// kCS8: output layout (E, C, 8) — the S values of one column contiguous (padded to 8) — for the tcgen05 backward kernel,
// whose edge-per-thread roles read / write whole columns with 128-bit accesses; default (E, S, C).
template <typename Cfg, bool kCS8 = false>
__global__ void __launch_bounds__(kSplitWarps * 32) segsum_kernel(const float* __restrict__ x,
                                                                   const int32_t* __restrict__ expand,
                                                                   const int32_t* __restrict__ rowptr,
                                                                   const float* __restrict__ sph, int64_t n_edges,
                                                                   float* __restrict__ A_out) {
  pdl_enter();
  __shared__ __align__(16) float stage[kSplitWarps][Cfg::SB * (Cfg::SP + 1)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * kSplitWarps + warp;
  if (e >= n_edges) return;
  float A[Cfg::S][Cfg::CC];
  segment_sum<Cfg>(x, expand, sph, rowptr[e], rowptr[e + 1], lane, stage[warp], A);
  if constexpr (kCS8) {
    static_assert(!kCS8 || Cfg::S <= 8, "the (E, C, 8) layout holds up to 8 basis values per column");
#pragma unroll
    for (int cc = 0; cc < Cfg::CC; ++cc) {
      float v[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) v[s] = s < Cfg::S ? A[s < Cfg::S ? s : 0][cc] : 0.f;
      float4* dst = reinterpret_cast<float4*>(A_out + (e * Cfg::C + lane + 32 * cc) * 8);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  } else {
    float* dst = A_out + e * (Cfg::S * Cfg::C) + lane;
#pragma unroll
    for (int s = 0; s < Cfg::S; ++s)
#pragma unroll
      for (int cc = 0; cc < Cfg::CC; ++cc) dst[s * Cfg::C + 32 * cc] = A[s][cc];
  }
}

// rows of an (S, C) per-edge matrix with lanes over s: lane L holds row L (and row 32 + L when S > 32), all C columns
template <typename Cfg>
__device__ __forceinline__ void load_rows_by_lane(const float* __restrict__ M, int lane,
                                                  float (&R)[(Cfg::S + 31) / 32][Cfg::C]) {
#pragma unroll
  for (int h = 0; h < (Cfg::S + 31) / 32; ++h) {
    const int s = lane + 32 * h;
    const float4* src = reinterpret_cast<const float4*>(M + (s < Cfg::S ? s : 0) * Cfg::C);
#pragma unroll
    for (int q = 0; q < Cfg::C / 4; ++q) {
      const float4 v = s < Cfg::S ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      R[h][4 * q] = v.x; R[h][4 * q + 1] = v.y; R[h][4 * q + 2] = v.z; R[h][4 * q + 3] = v.w;
    }
  }
}

// backward, edge part: dB = g W'^T (CTA tile GEMM), then per edge dA = rbf_W1^T dB (-> workspace) and g_rbf_W1 = dB A^T
template <typename Cfg>
__global__ void __launch_bounds__(256) segbil_bwd_edges_kernel(const float* __restrict__ rbf_W1,
                                                               const float* __restrict__ W, float scale,
                                                               int64_t n_edges, const float* __restrict__ g_out,
                                                               const float* __restrict__ A_in,
                                                               float* __restrict__ dA_out,
                                                               float* __restrict__ g_rbf_W1) {
  pdl_enter();
  extern __shared__ __align__(16) float smem[];
  float* dBs = smem;                        // [EB][LD]
  float* gs = smem + Cfg::EB * Cfg::LD;     // [EB][O]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* st = gs + Cfg::EB * Cfg::O + warp * (Cfg::SB * (Cfg::SP + 1));
  const int64_t e0 = (int64_t)blockIdx.x * Cfg::EB;

  for (int idx = threadIdx.x; idx < Cfg::EB * Cfg::O; idx += Cfg::THREADS) {
    const int64_t e = e0 + idx / Cfg::O;
    gs[idx] = e < n_edges ? g_out[e * Cfg::O + (idx % Cfg::O)] * scale : 0.f;
  }
  __syncthreads();
  {  // dB[e][i*C+c] = sum_o g[e,o] W[c,i,o]; thread owns k' = tid + 256 j for all EB edges
    const float4* W4 = reinterpret_cast<const float4*>(W);
    const float4* gs4 = reinterpret_cast<const float4*>(gs);
    for (int kp = threadIdx.x; kp < Cfg::K; kp += Cfg::THREADS) {
      const int c = kp & (Cfg::C - 1), i = kp / Cfg::C;
      const float4* wrow = W4 + (int64_t)(c * Cfg::I + i) * Cfg::TX;
      float acc[Cfg::EB];
#pragma unroll
      for (int el = 0; el < Cfg::EB; ++el) acc[el] = 0.f;
#pragma unroll 2
      for (int oq = 0; oq < Cfg::TX; ++oq) {
        const float4 w = __ldg(wrow + oq);
#pragma unroll
        for (int el = 0; el < Cfg::EB; ++el) {
          const float4 g = gs4[el * Cfg::TX + oq];
          acc[el] = fmaf(g.x, w.x, fmaf(g.y, w.y, fmaf(g.z, w.z, fmaf(g.w, w.w, acc[el]))));
        }
      }
#pragma unroll
      for (int el = 0; el < Cfg::EB; ++el) dBs[el * Cfg::LD + kp] = acc[el];
    }
  }
  __syncthreads();

  constexpr int IS = Cfg::I * Cfg::S;
  constexpr bool kBySLane = Cfg::S > 16;  // wide bases: g_rbf_W1 with lanes over s (no cross-lane reductions)
  constexpr int RSP = Cfg::S <= 8 ? 8 : 16;
  constexpr int SHIFT = RSP == 8 ? 2 : 1;
  for (int el = warp; el < Cfg::EB; el += Cfg::WARPS) {
    const int64_t e = e0 + el;
    if (e >= n_edges) continue;
    const float* drow = dBs + el * Cfg::LD;
    const float* rw = rbf_W1 + e * IS;
    const float* Ae = A_in + e * (Cfg::S * Cfg::C);
    float dA[Cfg::S][Cfg::CC];
#pragma unroll
    for (int s = 0; s < Cfg::S; ++s)
#pragma unroll
      for (int cc = 0; cc < Cfg::CC; ++cc) dA[s][cc] = 0.f;
    float As[kBySLane ? (Cfg::S + 31) / 32 : 1][kBySLane ? Cfg::C : 1];   // A rows by lane (wide bases)
    float Ac[kBySLane ? 1 : Cfg::S][kBySLane ? 1 : Cfg::CC];              // A columns by lane (narrow bases)
    if constexpr (kBySLane) {
      load_rows_by_lane<Cfg>(Ae, lane, As);
    } else {
#pragma unroll
      for (int s = 0; s < Cfg::S; ++s)
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) Ac[s][cc] = __ldg(Ae + s * Cfg::C + lane + 32 * cc);
    }
    UnitStager<Cfg> rs;
    rs.fetch(nullptr, rw, 0, Cfg::I, lane);
#pragma unroll 1
    for (int i0 = 0; i0 < Cfg::I; i0 += Cfg::SB) {
      rs.commit(st, lane);
      if (i0 + Cfg::SB < Cfg::I) rs.fetch(nullptr, rw, i0 + Cfg::SB, Cfg::I, lane);
#pragma unroll 1
      for (int u = 0; u < Cfg::SB; ++u) {
        const int i = i0 + u;
        const float* wrow = st + u * Cfg::SP;
        float d[Cfg::CC];
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) d[cc] = drow[i * Cfg::C + lane + 32 * cc];
#pragma unroll
        for (int s = 0; s < Cfg::S; ++s) {
          const float w = wrow[s];
#pragma unroll
          for (int cc = 0; cc < Cfg::CC; ++cc) dA[s][cc] = fmaf(w, d[cc], dA[s][cc]);
        }
        if (g_rbf_W1) {
          if constexpr (kBySLane) {
            // g_rbf_W1[e,i,s] = sum_c dB[i,c] A[s,c]: dB row broadcast from shared memory, A row in registers
            float p[(Cfg::S + 31) / 32];
#pragma unroll
            for (int h = 0; h < (Cfg::S + 31) / 32; ++h) p[h] = 0.f;
            const float4* d4 = reinterpret_cast<const float4*>(drow + i * Cfg::C);
#pragma unroll
            for (int q = 0; q < Cfg::C / 4; ++q) {
              const float4 dv = d4[q];
#pragma unroll
              for (int h = 0; h < (Cfg::S + 31) / 32; ++h)
                p[h] = fmaf(dv.x, As[h][4 * q], fmaf(dv.y, As[h][4 * q + 1],
                       fmaf(dv.z, As[h][4 * q + 2], fmaf(dv.w, As[h][4 * q + 3], p[h]))));
            }
#pragma unroll
            for (int h = 0; h < (Cfg::S + 31) / 32; ++h)
              if (lane + 32 * h < Cfg::S) g_rbf_W1[e * IS + i * Cfg::S + lane + 32 * h] = p[h];
          } else {
#pragma unroll
            for (int s0 = 0; s0 < Cfg::S; s0 += RSP) {
              float v[RSP];
#pragma unroll
              for (int k = 0; k < RSP; ++k) {
                float p = 0.f;
                if (s0 + k < Cfg::S) {
#pragma unroll
                  for (int cc = 0; cc < Cfg::CC; ++cc) p = fmaf(d[cc], Ac[(s0 + k) < Cfg::S ? (s0 + k) : 0][cc], p);
                }
                v[k] = p;
              }
              const float tot = warp_multi_reduce<RSP>(v, lane);
              const int k = lane >> SHIFT;
              if ((lane & ((1 << SHIFT) - 1)) == 0 && s0 + k < Cfg::S) g_rbf_W1[e * IS + i * Cfg::S + s0 + k] = tot;
            }
          }
        }
      }
    }
    float* dst = dA_out + e * (Cfg::S * Cfg::C) + lane;
#pragma unroll
    for (int s = 0; s < Cfg::S; ++s)
#pragma unroll
      for (int cc = 0; cc < Cfg::CC; ++cc) dst[s * Cfg::C + 32 * cc] = dA[s][cc];
  }
}

// backward, unit part (one warp per edge): g_sph[t,s] = sum_c dA[s,c] x[exp,c];  g_x[exp,c] += sum_s sph[t,s] dA[s,c]
template <typename Cfg, bool kCS8 = false>
__global__ void __launch_bounds__(kSplitWarps * 32) segbil_bwd_units_kernel(const float* __restrict__ x,
                                                                             const int32_t* __restrict__ expand,
                                                                             const int32_t* __restrict__ rowptr,
                                                                             const float* __restrict__ sph,
                                                                             int64_t n_edges,
                                                                             const float* __restrict__ dA_in,
                                                                             float* __restrict__ g_x,
                                                                             float* __restrict__ g_sph) {
  pdl_enter();
  constexpr bool kBySLane = Cfg::S > 16;
  __shared__ __align__(16) float stage[kSplitWarps][Cfg::SB * (Cfg::SP + 1)];
  __shared__ __align__(16) float xrows[kBySLane ? kSplitWarps : 1][kBySLane ? Cfg::SB * Cfg::C : 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * kSplitWarps + warp;
  if (e >= n_edges) return;
  const int32_t beg = rowptr[e], end = rowptr[e + 1];
  if (beg >= end) return;
  float* st = stage[warp];
  const float* dAe = dA_in + e * (kCS8 ? Cfg::C * 8 : Cfg::S * Cfg::C);
  float dA[Cfg::S][Cfg::CC];
  if constexpr (kCS8) {  // (E, C, 8) layout written by the tcgen05 edge kernel
#pragma unroll
    for (int cc = 0; cc < Cfg::CC; ++cc) {
      const float4* src = reinterpret_cast<const float4*>(dAe + (lane + 32 * cc) * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int s = 0; s < Cfg::S; ++s) dA[s][cc] = v[s < 8 ? s : 0];
    }
  } else {
#pragma unroll
    for (int s = 0; s < Cfg::S; ++s)
#pragma unroll
      for (int cc = 0; cc < Cfg::CC; ++cc) dA[s][cc] = __ldg(dAe + s * Cfg::C + lane + 32 * cc);
  }
  float dAs[kBySLane ? (Cfg::S + 31) / 32 : 1][kBySLane ? Cfg::C : 1];
  if constexpr (kBySLane) load_rows_by_lane<Cfg>(dAe, lane, dAs);
  constexpr int RSP = Cfg::S <= 8 ? 8 : 16;
  constexpr int SHIFT = RSP == 8 ? 2 : 1;

  UnitStager<Cfg> stg;
  stg.fetch(expand, sph, beg, end, lane);
  for (int32_t t0 = beg; t0 < end; t0 += Cfg::SB) {
    const int nb = min(Cfg::SB, end - t0);
    stg.commit(st, lane);
    if (t0 + Cfg::SB < end) stg.fetch(expand, sph, t0 + Cfg::SB, end, lane);
    float xb[Cfg::SB][Cfg::CC];
#pragma unroll
    for (int u = 0; u < Cfg::SB; ++u) {
      if (u < nb) {
        const float* xr = x + (int64_t)__float_as_int(st[Cfg::SB * Cfg::SP + u]) * Cfg::C + lane;
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) xb[u][cc] = __ldg(xr + 32 * cc);
      }
    }
    if constexpr (kBySLane) {  // park the gathered rows in shared memory: the g_sph product reads them as broadcasts
      if (g_sph) {
        __syncwarp();
#pragma unroll
        for (int u = 0; u < Cfg::SB; ++u)
          if (u < nb) {
#pragma unroll
            for (int cc = 0; cc < Cfg::CC; ++cc) xrows[warp][u * Cfg::C + lane + 32 * cc] = xb[u][cc];
          }
        __syncwarp();
      }
    }
#pragma unroll
    for (int u = 0; u < Cfg::SB; ++u) {
      if (u >= nb) break;
      const int32_t t = t0 + u;
      const float* sp = st + u * Cfg::SP;
      if (g_x) {
        float gx[Cfg::CC];
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) gx[cc] = 0.f;
        const float4* sp4 = reinterpret_cast<const float4*>(sp);
#pragma unroll
        for (int s4 = 0; s4 < Cfg::SP / 4; ++s4) {
          const float4 w = sp4[s4];
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (4 * s4 + j < Cfg::S) {
#pragma unroll
              for (int cc = 0; cc < Cfg::CC; ++cc) gx[cc] = fmaf(wv[j], dA[4 * s4 + j][cc], gx[cc]);
            }
        }
        float* gr = g_x + (int64_t)__float_as_int(st[Cfg::SB * Cfg::SP + u]) * Cfg::C + lane;
#pragma unroll
        for (int cc = 0; cc < Cfg::CC; ++cc) atomicAdd(gr + 32 * cc, gx[cc]);
      }
      if (g_sph) {
        if constexpr (kBySLane) {
          float p[(Cfg::S + 31) / 32];
#pragma unroll
          for (int h = 0; h < (Cfg::S + 31) / 32; ++h) p[h] = 0.f;
          const float4* x4 = reinterpret_cast<const float4*>(xrows[warp] + u * Cfg::C);
#pragma unroll
          for (int q = 0; q < Cfg::C / 4; ++q) {
            const float4 xv = x4[q];
#pragma unroll
            for (int h = 0; h < (Cfg::S + 31) / 32; ++h)
              p[h] = fmaf(xv.x, dAs[h][4 * q], fmaf(xv.y, dAs[h][4 * q + 1],
                     fmaf(xv.z, dAs[h][4 * q + 2], fmaf(xv.w, dAs[h][4 * q + 3], p[h]))));
          }
#pragma unroll
          for (int h = 0; h < (Cfg::S + 31) / 32; ++h)
            if (lane + 32 * h < Cfg::S) g_sph[(int64_t)t * Cfg::S + lane + 32 * h] = p[h];
        } else {
#pragma unroll
          for (int s0 = 0; s0 < Cfg::S; s0 += RSP) {
            float v[RSP];
#pragma unroll
            for (int k = 0; k < RSP; ++k) {
              float p = 0.f;
              if (s0 + k < Cfg::S) {
#pragma unroll
                for (int cc = 0; cc < Cfg::CC; ++cc) p = fmaf(dA[(s0 + k) < Cfg::S ? (s0 + k) : 0][cc], xb[u][cc], p);
              }
              v[k] = p;
            }
            const float tot = warp_multi_reduce<RSP>(v, lane);
            const int k = lane >> SHIFT;
            if ((lane & ((1 << SHIFT) - 1)) == 0 && s0 + k < Cfg::S) g_sph[(int64_t)t * Cfg::S + s0 + k] = tot;
          }
        }
      }
    }
  }
}

template <typename Kernel>
static int set_smem(Kernel kernel, size_t smem, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// workspace (floats): forward E*S*C (A), backward 2*E*S*C (A, dA); nullptr = fused single-kernel variants
template <typename Cfg>
static int launch_fwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                      const float* rbf_W1, const float* W, float scale, int64_t n_edges, float* out,
                      float* workspace, cudaStream_t st) {
  const size_t smem = sizeof(float) * (Cfg::EB * Cfg::LD + Cfg::STAGE_FLOATS);
  static bool configured = false;
  if (!configured) {
    if (set_smem(segbil_fwd_kernel<Cfg, false>, smem, "segbilinear fwd")) return 1;
    if (set_smem(segbil_fwd_kernel<Cfg, true>, smem, "segbilinear fwd")) return 1;
    configured = true;
  }
  const unsigned grid = (unsigned)ceil_div(n_edges, Cfg::EB);
  if (workspace) {
    launch_pdl(segsum_kernel<Cfg>, (unsigned)ceil_div(n_edges, kSplitWarps), kSplitWarps * 32, 0, st, x, expand, rowptr,
               sph, n_edges, workspace);
    launch_pdl(segbil_fwd_kernel<Cfg, true>, grid, Cfg::THREADS, smem, st, x, expand, rowptr, sph, rbf_W1, W, scale,
               n_edges, workspace, out);
  } else {
    launch_pdl(segbil_fwd_kernel<Cfg, false>, grid, Cfg::THREADS, smem, st, x, expand, rowptr, sph, rbf_W1, W, scale,
               n_edges, nullptr, out);
  }
  return check_launch("gemnet_segbilinear_fwd");
}

template <typename Cfg>
static int launch_bwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                      const float* rbf_W1, const float* W, float scale, int64_t n_edges, const float* g_out,
                      float* g_x, float* g_sph, float* g_rbf_W1, float* workspace, cudaStream_t st,
                      bool a_saved = false) {
  const size_t smem = sizeof(float) * (Cfg::EB * Cfg::LD + Cfg::EB * Cfg::O + Cfg::STAGE_FLOATS);
  static bool configured = false;
  if (!configured) {
    if (set_smem(segbil_bwd_kernel<Cfg>, smem, "segbilinear bwd")) return 1;
    if (set_smem(segbil_bwd_edges_kernel<Cfg>, smem, "segbilinear bwd")) return 1;
    configured = true;
  }
  const unsigned grid = (unsigned)ceil_div(n_edges, Cfg::EB);
  if (workspace) {
    const unsigned wgrid = (unsigned)ceil_div(n_edges, kSplitWarps);
    float* A = workspace;
    float* dA = workspace + n_edges * (Cfg::S * Cfg::C);
    if (g_rbf_W1 && !a_saved)   // (a_saved: the forward's segment sums are still in the workspace head)
      launch_pdl(segsum_kernel<Cfg>, wgrid, kSplitWarps * 32, 0, st, x, expand, rowptr, sph, n_edges, A);
    launch_pdl(segbil_bwd_edges_kernel<Cfg>, grid, Cfg::THREADS, smem, st, rbf_W1, W, scale, n_edges, g_out, A, dA,
               g_rbf_W1);
    if (g_x || g_sph)
      launch_pdl(segbil_bwd_units_kernel<Cfg>, wgrid, kSplitWarps * 32, 0, st, x, expand, rowptr, sph, n_edges, dA, g_x,
                 g_sph);
  } else {
    launch_pdl(segbil_bwd_kernel<Cfg>, grid, Cfg::THREADS, smem, st, x, expand, rowptr, sph, rbf_W1, W, scale, n_edges,
               g_out, g_x, g_sph, g_rbf_W1);
  }
  return check_launch("gemnet_segbilinear_bwd");
}

// tcgen05 path (segbilinear_tc.cu)
int64_t segbil_tc_workspace_bytes();
int launch_segbil_fwd_tc(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                         const float* rbf_W1, const float* W, float scale, int64_t n_edges, float* out,
                         void* workspace, cudaStream_t st);

// tcgen05 backward, edge part (segbilinear_bwd_tc.cu)
int64_t segbil_bwd_tc_pack_bytes();
int launch_segbil_bwd_pack_tc(const float* W, void* packed, cudaStream_t st);
int launch_segbil_bwd_edges_tc(const float* rbf_W1, const void* packed, float scale, int64_t n_edges, const float* g_out,
                               const float* A_in, float* dA_out, float* g_rbf_W1, cudaStream_t st);
int launch_segbil_fwd_pack_tc(const float* W, void* packed, cudaStream_t st);
int launch_segbil_fwd_tc_packed(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                                const float* rbf_W1, const void* packed, float scale, int64_t n_edges, float* out,
                                cudaStream_t st);

using TripCfg = SegCfg<7, 64, 16, 64, 16>;
using QuadCfg = SegCfg<49, 32, 32, 32, 32>;

// backward of the triplet shape with the dB GEMM on tensor cores: segsum (A, layout (E,C,8)) -> tcgen05 edge kernel
// (dA, g_rbf_W1) -> unit kernel (g_x, g_sph).  ws: A and dA, E*C*8 floats each.
static int launch_bwd_tc(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                         const float* rbf_W1, const void* packed, float scale, int64_t n_edges, const float* g_out,
                         float* g_x, float* g_sph, float* g_rbf_W1, float* ws, cudaStream_t st) {
  using Cfg = TripCfg;
  const unsigned wgrid = (unsigned)ceil_div(n_edges, kSplitWarps);
  float* A = ws;
  float* dA = ws + n_edges * (Cfg::C * 8);
  if (g_rbf_W1)
    launch_pdl(segsum_kernel<Cfg, true>, wgrid, kSplitWarps * 32, 0, st, x, expand, rowptr, sph, n_edges, A);
  if (launch_segbil_bwd_edges_tc(rbf_W1, packed, scale, n_edges, g_out, A, dA, g_rbf_W1, st)) return 1;
  if (g_x || g_sph)
    launch_pdl(segbil_bwd_units_kernel<Cfg, true>, wgrid, kSplitWarps * 32, 0, st, x, expand, rowptr, sph, n_edges, dA,
               g_x, g_sph);
  return check_launch("gemnet_segbilinear_bwd(tcgen05)");
}

}  // namespace gemnet

using namespace gemnet;

static int shape_id(int32_t S, int32_t C, int32_t I, int32_t O) {
  if (S == 7 && C == 64 && I == 16 && O == 64) return 1;
  if (S == 49 && C == 32 && I == 32 && O == 32) return 2;
  return 0;
}

static bool split_enabled() {
  static const bool on = [] {
    const char* e = getenv("GEMNET_B200_H7_SPLIT");
    return !(e && e[0] == '0');
  }();
  return on;
}

GEMNET_API int64_t gemnet_segbilinear_pack_bytes(int32_t S, int32_t C, int32_t I, int32_t O, int32_t backward) {
  if (shape_id(S, C, I, O) != 1) return 0;
  return backward ? segbil_bwd_tc_pack_bytes() : segbil_tc_workspace_bytes();
}

GEMNET_API int gemnet_segbilinear_pack(const float* W, int32_t S, int32_t C, int32_t I, int32_t O, int32_t backward,
                                       void* packed, void* stream) {
  GEMNET_REQUIRE(shape_id(S, C, I, O) == 1, "the tcgen05 kernels are built for the triplet shape (7,64,16,64) only");
  GEMNET_REQUIRE(W && packed && ((uintptr_t)packed % 16) == 0, "null / misaligned pointer");
  return backward ? launch_segbil_bwd_pack_tc(W, packed, as_stream(stream)) : launch_segbil_fwd_pack_tc(W, packed, as_stream(stream));
}

/* A[e,s,c] = sum_{t in seg(e)} sph[t,s] x[expand[t],c]  (phase 1 of the layer on its own: the weight gradient
 * dL/dW[c,i,o] = sum_e (sum_s rbf_W1[e,i,s] A[e,s,c]) g[e,o] is two library contractions on top of it) */
GEMNET_API int gemnet_segbilinear_segsum(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                                         int64_t n_edges, int32_t S, int32_t C, int32_t I, int32_t O, float* A_out,
                                         void* stream) {
  const int sid = shape_id(S, C, I, O);
  GEMNET_REQUIRE(sid != 0, "unsupported (S,C,I,O): built for (7,64,16,64) and (49,32,32,32)");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(x && expand && rowptr && sph && A_out, "null pointer");
  const unsigned wgrid = (unsigned)ceil_div(n_edges, kSplitWarps);
  if (sid == 1)
    launch_pdl(segsum_kernel<TripCfg, false>, wgrid, kSplitWarps * 32, 0, as_stream(stream), x, expand, rowptr, sph, n_edges, A_out);
  else
    launch_pdl(segsum_kernel<QuadCfg, false>, wgrid, kSplitWarps * 32, 0, as_stream(stream), x, expand, rowptr, sph, n_edges, A_out);
  return check_launch("gemnet_segbilinear_segsum");
}

GEMNET_API int64_t gemnet_segbilinear_workspace_bytes(int64_t n_edges, int32_t S, int32_t C, int32_t I, int32_t O,
                                                      int32_t algo, int32_t backward) {
  const bool tc = (algo & GEMNET_SEGBIL_TC) && shape_id(S, C, I, O) == 1;
  const bool prepacked = (algo & GEMNET_SEGBIL_PREPACKED) != 0;
  if (tc && !backward) return prepacked ? 0 : segbil_tc_workspace_bytes();
  // tcgen05 backward: [packed image unless prepacked][A (E,C,8)][dA (E,C,8)]
  if (tc && backward)
    return (prepacked ? 0 : segbil_bwd_tc_pack_bytes()) + (int64_t)sizeof(float) * n_edges * C * 8 * 2;
  if (!split_enabled()) return 0;
  // split pipeline of the fp32 kernels: A (E,S,C) for the forward, A and dA for the backward.  The narrow triplet
  // forward stays fused (measured: 83 us fused vs 140 us split at batch 64; its segments are short and balanced).
  if (!backward && S <= 16) return 0;
  return (int64_t)sizeof(float) * n_edges * S * C * (backward ? 2 : 1);
}

GEMNET_API int gemnet_segbilinear_fwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                                      const float* rbf_W1, const float* W, float scale, int64_t n_edges,
                                      int64_t n_units, int32_t S, int32_t C, int32_t I, int32_t O, int32_t algo,
                                      float* out, void* workspace, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && n_units >= 0, "negative size");
  GEMNET_REQUIRE(algo >= 0 && algo <= 3 && algo != 2, "algo: 0 fp32 CUDA cores, 1 tcgen05, 3 tcgen05 with a pre-packed weight image");
  const int sid = shape_id(S, C, I, O);
  GEMNET_REQUIRE(sid != 0, "unsupported (S,C,I,O): built for (7,64,16,64) and (49,32,32,32)");
  GEMNET_REQUIRE(algo == 0 || sid == 1, "algo 1 (tcgen05) is built for the triplet shape (7,64,16,64) only");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(rowptr && rbf_W1 && W && out && (n_units == 0 || (x && expand && sph)), "null pointer");
  GEMNET_REQUIRE(((uintptr_t)W | (uintptr_t)out) % 16 == 0, "W/out must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (algo == 3) return launch_segbil_fwd_tc_packed(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, out, st);
  if (algo == 1) return launch_segbil_fwd_tc(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, out, workspace, st);
  float* ws = split_enabled() ? static_cast<float*>(workspace) : nullptr;
  if (sid == 1) return launch_fwd<TripCfg>(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, out, nullptr, st);
  return launch_fwd<QuadCfg>(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, out, ws, st);
}

GEMNET_API int gemnet_segbilinear_bwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                                      const float* rbf_W1, const float* W, float scale, int64_t n_edges,
                                      int64_t n_units, int32_t S, int32_t C, int32_t I, int32_t O, int32_t algo,
                                      const float* g_out, float* g_x, float* g_sph, float* g_rbf_W1, float* g_W,
                                      void* workspace, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && n_units >= 0, "negative size");
  const bool a_saved = (algo & GEMNET_SEGBIL_A_SAVED) != 0;
  algo &= ~GEMNET_SEGBIL_A_SAVED;
  GEMNET_REQUIRE(algo >= 0 && algo <= 3 && algo != 2, "algo: 0 fp32 CUDA cores, 1 tcgen05, 3 tcgen05 with a pre-packed weight image");
  GEMNET_REQUIRE(g_W == nullptr, "g_W is computed by the host wrapper in this library version; pass NULL");
  const int sid = shape_id(S, C, I, O);
  GEMNET_REQUIRE(sid != 0, "unsupported (S,C,I,O): built for (7,64,16,64) and (49,32,32,32)");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(rowptr && rbf_W1 && W && g_out && (n_units == 0 || (x && expand && sph)), "null pointer");
  GEMNET_REQUIRE(((uintptr_t)W) % 16 == 0, "W must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if ((algo & GEMNET_SEGBIL_TC) && sid == 1) {
    // tensor-core backward (triplet shape): dB = g W'^T on tcgen05; the other shapes keep the fp32 kernels
    GEMNET_REQUIRE(workspace && ((uintptr_t)workspace % 16) == 0, "workspace of gemnet_segbilinear_workspace_bytes() bytes required");
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    const void* packed = W;
    if (!(algo & GEMNET_SEGBIL_PREPACKED)) {
      if (launch_segbil_bwd_pack_tc(W, ws, st)) return 1;
      packed = ws;
      ws += segbil_bwd_tc_pack_bytes();
    }
    return launch_bwd_tc(x, expand, rowptr, sph, rbf_W1, packed, scale, n_edges, g_out, g_x, g_sph, g_rbf_W1,
                         reinterpret_cast<float*>(ws), st);
  }
  float* ws = split_enabled() ? static_cast<float*>(workspace) : nullptr;
  if (sid == 1)
    return launch_bwd<TripCfg>(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, g_out, g_x, g_sph, g_rbf_W1, ws, st);
  return launch_bwd<QuadCfg>(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, g_out, g_x, g_sph, g_rbf_W1, ws, st,
                             a_saved && ws != nullptr);
}
