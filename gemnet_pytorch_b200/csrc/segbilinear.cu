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

template <typename Cfg>
__global__ void __launch_bounds__(256) segbil_fwd_kernel(const float* __restrict__ x,
                                                         const int32_t* __restrict__ expand,
                                                         const int32_t* __restrict__ rowptr,
                                                         const float* __restrict__ sph,
                                                         const float* __restrict__ rbf_W1,
                                                         const float* __restrict__ W, float scale, int64_t n_edges,
                                                         float* __restrict__ out) {
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
    segment_sum<Cfg>(x, expand, sph, rowptr[e], rowptr[e + 1], lane, st, A);
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

template <typename Cfg>
static int launch_fwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                      const float* rbf_W1, const float* W, float scale, int64_t n_edges, float* out,
                      cudaStream_t st) {
  const size_t smem = sizeof(float) * (Cfg::EB * Cfg::LD + Cfg::STAGE_FLOATS);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(segbil_fwd_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("segbilinear fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const unsigned grid = (unsigned)ceil_div(n_edges, Cfg::EB);
  launch_pdl(segbil_fwd_kernel<Cfg>, grid, Cfg::THREADS, smem, st, x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, out);
  return check_launch("gemnet_segbilinear_fwd");
}

template <typename Cfg>
static int launch_bwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                      const float* rbf_W1, const float* W, float scale, int64_t n_edges, const float* g_out,
                      float* g_x, float* g_sph, float* g_rbf_W1, cudaStream_t st) {
  const size_t smem = sizeof(float) * (Cfg::EB * Cfg::LD + Cfg::EB * Cfg::O + Cfg::STAGE_FLOATS);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(segbil_bwd_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("segbilinear bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const unsigned grid = (unsigned)ceil_div(n_edges, Cfg::EB);
  launch_pdl(segbil_bwd_kernel<Cfg>, grid, Cfg::THREADS, smem, st, x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, g_out,
                                                           g_x, g_sph, g_rbf_W1);
  return check_launch("gemnet_segbilinear_bwd");
}

// tcgen05 path (segbilinear_tc.cu)
int64_t segbil_tc_workspace_bytes();
int launch_segbil_fwd_tc(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                         const float* rbf_W1, const float* W, float scale, int64_t n_edges, float* out,
                         void* workspace, cudaStream_t st);

using TripCfg = SegCfg<7, 64, 16, 64, 16>;
using QuadCfg = SegCfg<49, 32, 32, 32, 32>;

}  // namespace gemnet

using namespace gemnet;

GEMNET_API int64_t gemnet_segbilinear_workspace_bytes(int64_t n_edges, int32_t S, int32_t C, int32_t I, int32_t O,
                                                      int32_t algo, int32_t backward) {
  (void)n_edges;
  if (algo == 1 && !backward && S == 7 && C == 64 && I == 16 && O == 64) return segbil_tc_workspace_bytes();
  return 0;
}

static int shape_id(int32_t S, int32_t C, int32_t I, int32_t O) {
  if (S == 7 && C == 64 && I == 16 && O == 64) return 1;
  if (S == 49 && C == 32 && I == 32 && O == 32) return 2;
  return 0;
}

GEMNET_API int gemnet_segbilinear_fwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                                      const float* rbf_W1, const float* W, float scale, int64_t n_edges,
                                      int64_t n_units, int32_t S, int32_t C, int32_t I, int32_t O, int32_t algo,
                                      float* out, void* workspace, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && n_units >= 0, "negative size");
  GEMNET_REQUIRE(algo == 0 || algo == 1, "algo must be 0 (fp32 CUDA cores) or 1 (tcgen05)");
  const int sid = shape_id(S, C, I, O);
  GEMNET_REQUIRE(sid != 0, "unsupported (S,C,I,O): built for (7,64,16,64) and (49,32,32,32)");
  GEMNET_REQUIRE(algo == 0 || sid == 1, "algo 1 (tcgen05) is built for the triplet shape (7,64,16,64) only");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(rowptr && rbf_W1 && W && out && (n_units == 0 || (x && expand && sph)), "null pointer");
  GEMNET_REQUIRE(((uintptr_t)W | (uintptr_t)out) % 16 == 0, "W/out must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (algo == 1) return launch_segbil_fwd_tc(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, out, workspace, st);
  if (sid == 1) return launch_fwd<TripCfg>(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, out, st);
  return launch_fwd<QuadCfg>(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, out, st);
}

GEMNET_API int gemnet_segbilinear_bwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                                      const float* rbf_W1, const float* W, float scale, int64_t n_edges,
                                      int64_t n_units, int32_t S, int32_t C, int32_t I, int32_t O, int32_t algo,
                                      const float* g_out, float* g_x, float* g_sph, float* g_rbf_W1, float* g_W,
                                      void* workspace, void* stream) {
  (void)workspace;
  (void)algo;  // the backward runs on the fp32 CUDA-core kernel for every algo in this library version
  GEMNET_REQUIRE(n_edges >= 0 && n_units >= 0, "negative size");
  GEMNET_REQUIRE(g_W == nullptr, "g_W is computed by the host wrapper in this library version; pass NULL");
  const int sid = shape_id(S, C, I, O);
  GEMNET_REQUIRE(sid != 0, "unsupported (S,C,I,O): built for (7,64,16,64) and (49,32,32,32)");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(rowptr && rbf_W1 && W && g_out && (n_units == 0 || (x && expand && sph)), "null pointer");
  GEMNET_REQUIRE(((uintptr_t)W) % 16 == 0, "W must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (sid == 1)
    return launch_bwd<TripCfg>(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, g_out, g_x, g_sph, g_rbf_W1, st);
  return launch_bwd<QuadCfg>(x, expand, rowptr, sph, rbf_W1, W, scale, n_edges, g_out, g_x, g_sph, g_rbf_W1, st);
}
