// N2 / H0 on the device: index construction kernels (see graph_device_bodies.h for the algorithm and the per-item
// bodies, which are shared with the host emulation used by the CPU tests).  This file adds the grid-stride loops, the
// block-wide exclusive scans between the counting and the emission passes, and the C ABI (include/gemnet_b200.h, N2).
//
// Launches: stage 1 = 3 (molecule scan, per-atom counts, atom scans), stage 2 = 4 for triplet models (edges, incoming
// lists, edge scans, triplets) + 2 for quadruplet models (counts per reduce edge, their scan), stage 3 = 1, pad = 1.
// Every size is read from the device header, so the same launches replay inside a captured CUDA graph for batches of
// any size up to the capacities; a stage whose header carries an overflow bit does nothing.
#include "common.cuh"
#include "graph_device_bodies.h"

namespace gemnet {
namespace gd {

constexpr int kScan = 1024;  // threads of the single-CTA scan kernels (32 warps: one shuffle scan over the warp totals)

// Exclusive scan of f(0..n) by one 1024-thread CTA, tile by tile with coalesced accesses; out (n + 1 entries, may alias
// the array f reads element-wise, may be NULL) and the total (the same value in every thread).
template <typename F>
__device__ long long block_scan(F f, int32_t n, int32_t* out) {
  __shared__ long long warp_tot[32];
  __shared__ long long tile_tot;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  long long carry = 0;
  for (int32_t base = 0; base < n; base += kScan) {
    const int32_t i = base + tid;
    const long long v = i < n ? (long long)f(i) : 0;
    long long s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_tot[w] = s;
    __syncthreads();
    if (w == 0) {
      const long long t = warp_tot[lane];
      long long ts = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, ts, o);
        if (lane >= o) ts += u;
      }
      warp_tot[lane] = ts - t;  // exclusive over the warps
      if (lane == 31) tile_tot = ts;
    }
    __syncthreads();
    if (out != nullptr && i < n) out[i] = (int32_t)(carry + warp_tot[w] + s - v);
    carry += tile_tot;
    __syncthreads();  // warp_tot / tile_tot are rewritten by the next tile
  }
  if (out != nullptr && tid == 0) out[n] = (int32_t)carry;
  return carry;
}

__global__ void __launch_bounds__(kScan) mol_scan_kernel(gemnet_device_graph g) {
  __shared__ int bad_s;
  int32_t* hdr = g.hdr;
  if (threadIdx.x == 0) bad_s = 0;
  __syncthreads();
  int32_t n_mol = hdr[H_MOLS];
  const int32_t n_atoms = hdr[H_ATOMS];
  const bool bad = n_mol < 0 || n_mol > g.cap_molecules || n_atoms < 0 || n_atoms > g.cap_atoms;
  if (bad) n_mol = 0;
  const long long tot = block_scan(
      [&](int32_t m) {
        const int32_t v = g.n_per_mol[m];
        if (v < 0) bad_s = 1;
        return v > 0 ? v : 0;
      },
      n_mol, g.mol_off);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = H_UPPER; k < GEMNET_GRAPH_HDR_WORDS; ++k) hdr[k] = 0;
    if (bad || bad_s || tot != (long long)n_atoms) hdr[H_OVERFLOW] = 1;
  }
}

__global__ void atom_count_kernel(gemnet_device_graph g) {
  const int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const int32_t n = hdr[H_ATOMS];
  const int32_t stride = gridDim.x * blockDim.x;
  float amax = 0.f;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) amax = fmaxf(amax, atom_count(g, hdr, i));
  // non-negative floats order like their bit patterns
  if (amax > 0.f) atomicMax(&g.hdr[H_AMAX], __float_as_int(amax));
}

__global__ void __launch_bounds__(kScan) atom_scan_kernel(gemnet_device_graph g) {
  int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const int32_t n = hdr[H_ATOMS];
  const int32_t* up = g.atom_cnt;
  const int32_t* deg = g.atom_cnt + g.cap_atoms;
  const int32_t* ideg = g.atom_cnt + 2 * g.cap_atoms;
  const long long nu = block_scan([&](int32_t i) { return up[i]; }, n, g.upper_off);
  block_scan([&](int32_t i) { return deg[i]; }, n, g.in_ptr);
  const long long nt = block_scan([&](int32_t i) { return (long long)deg[i] * (deg[i] - 1); }, n, nullptr);
  long long ni = 0, nintm = 0;
  if (!g.triplets_only) {
    ni = block_scan([&](int32_t i) { return ideg[i]; }, n, g.int_off);
    nintm = block_scan([&](int32_t i) { return (long long)ideg[i] * deg[i]; }, n, nullptr);
  }
  if (threadIdx.x == 0) {
    const long long lim = 0x7fffffffLL;
    int32_t ov = 0;
    if (2 * nu > g.cap_edges || 2 * nu > lim) ov |= 2;
    if (nt > g.cap_triplets || nt > lim) ov |= 4;
    if (ni > g.cap_int_edges || ni > lim) ov |= 8;
    if (nintm > g.cap_intm || nintm > lim) ov |= 16;
    hdr[H_UPPER] = (int32_t)nu;
    hdr[H_EDGES] = (int32_t)(2 * nu);
    hdr[H_TRIPLETS] = (int32_t)nt;
    hdr[H_INT_EDGES] = (int32_t)ni;
    hdr[H_INTM] = (int32_t)nintm;
    hdr[H_OVERFLOW] = ov;
  }
}

__global__ void edges_kernel(gemnet_device_graph g) {
  const int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const int32_t n = hdr[H_ATOMS];
  const int32_t stride = gridDim.x * blockDim.x;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) edges(g, hdr, i);
}

__global__ void incoming_kernel(gemnet_device_graph g) {
  const int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const int32_t n = hdr[H_ATOMS];
  const int32_t stride = gridDim.x * blockDim.x;
  for (int32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) incoming(g, hdr, x);
}

__global__ void __launch_bounds__(kScan) edge_scan_kernel(gemnet_device_graph g) {
  const int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  block_scan([&](int32_t r) { return in_degree(g, g.id_a[r]) - 1; }, hdr[H_EDGES], g.rowptr3);
  if (!g.triplets_only) {
    const int32_t ni = hdr[H_INT_EDGES];
    block_scan([&](int32_t i) { return in_degree(g, g.id4_int_a[i]); }, ni, g.intm_off_a);
    block_scan([&](int32_t i) { return in_degree(g, g.id4_int_b[i]); }, ni, g.intm_off_b);
  }
}

__global__ void triplets_kernel(gemnet_device_graph g) {
  const int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const int32_t stride = gridDim.x * blockDim.x;
  const int32_t t0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t ne = hdr[H_EDGES];
  for (int32_t r = t0; r < ne; r += stride) triplets(g, r);
  if (!g.triplets_only) {
    const int32_t ni = hdr[H_INT_EDGES];
    for (int32_t i = t0; i < ni; i += stride) intermediate(g, i);
  }
}

__global__ void quad_count_kernel(gemnet_device_graph g) {
  const int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const int32_t ne = hdr[H_EDGES];
  const int32_t stride = gridDim.x * blockDim.x;
  for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < ne; r += stride) g.quad_ptr[r] = quad_scan<false>(g, r);
}

__global__ void __launch_bounds__(kScan) quad_ptr_kernel(gemnet_device_graph g) {
  int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const int32_t* cnt = g.quad_ptr;
  const long long nq = block_scan([&](int32_t r) { return cnt[r]; }, hdr[H_EDGES], g.quad_ptr);  // in place
  if (threadIdx.x == 0) {
    hdr[H_QUADS] = nq > 0x7fffffffLL ? 0x7fffffff : (int32_t)nq;
    if (nq > g.cap_quads || nq > 0x7fffffffLL) hdr[H_OVERFLOW] = 32;
  }
}

__global__ void quad_emit_kernel(gemnet_device_graph g) {
  const int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const int32_t ne = hdr[H_EDGES];
  const int32_t stride = gridDim.x * blockDim.x;
  for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < ne; r += stride) quad_scan<true>(g, r);
}

__global__ void pad_kernel(gemnet_device_graph g, int32_t* Z, int32_t* N_out) {
  const int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return;
  const PadSizes s = pad_sizes(g, hdr);
  const int32_t stride = gridDim.x * blockDim.x;
  const int32_t t0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (s.n_pad < 3 || s.n_pe < 6 || s.nt < 0) {  // no room for the pad molecule: the capacity bucket was too small
    if (t0 == 0) atomicOr(&g.hdr[H_OVERFLOW], 64);
    return;
  }
  for (int32_t k = t0; k < s.n_pad; k += stride) pad_atom(g, s, Z, k);
  for (int32_t m = t0; m <= (int32_t)g.cap_molecules; m += stride) pad_molecule(g, s, N_out, m);
  for (int32_t j = t0; j < s.n_pe; j += stride) pad_edge(g, s, j);
  for (int32_t j = t0; j < s.nt; j += stride) pad_triplet(g, s, j);
}

inline bool fits_i32(int64_t v) { return v >= 0 && v <= 0x7fffffffLL; }

}  // namespace gd
}  // namespace gemnet

using namespace gemnet;
using namespace gemnet::gd;

#define GD_REQUIRE_COMMON(g)                                                                            \
  GEMNET_REQUIRE((g) != nullptr, "null graph descriptor");                                              \
  GEMNET_REQUIRE((g)->hdr != nullptr, "null header");                                                   \
  GEMNET_REQUIRE(fits_i32((g)->cap_atoms) && fits_i32((g)->cap_molecules) && fits_i32((g)->cap_edges) && \
                     fits_i32((g)->cap_triplets) && fits_i32((g)->cap_int_edges) && fits_i32((g)->cap_intm) && \
                     fits_i32((g)->cap_quads),                                                          \
                 "capacities must be in [0, 2^31)")

GEMNET_API int gemnet_graph_device_count(const gemnet_device_graph* g, void* stream) {
  GD_REQUIRE_COMMON(g);
  GEMNET_REQUIRE(g->mol_off && g->upper_off && g->in_ptr, "null workspace");
  GEMNET_REQUIRE(g->cap_atoms == 0 || (g->R && g->atom_cnt && g->batch_seg), "null atom-level array");
  GEMNET_REQUIRE(g->cap_molecules == 0 || g->n_per_mol != nullptr, "null n_per_mol");
  GEMNET_REQUIRE(g->triplets_only || g->int_off != nullptr, "null int_off");
  cudaStream_t st = as_stream(stream);
  mol_scan_kernel<<<1, kScan, 0, st>>>(*g);
  atom_count_kernel<<<flat_grid(g->cap_atoms, 128), 128, 0, st>>>(*g);
  atom_scan_kernel<<<1, kScan, 0, st>>>(*g);
  return check_launch("gemnet_graph_device_count");
}

GEMNET_API int gemnet_graph_device_edges(const gemnet_device_graph* g, void* stream) {
  GD_REQUIRE_COMMON(g);
  GEMNET_REQUIRE(g->mol_off && g->upper_off && g->in_ptr && g->rowptr3, "null workspace / output");
  GEMNET_REQUIRE(g->cap_atoms == 0 || (g->R && g->batch_seg), "null atom-level array");
  if (g->cap_edges > 0)
    GEMNET_REQUIRE(g->incoming && g->incoming_pos && g->id_c && g->id_a && g->id_undir && g->id_swap, "null edge array");
  if (g->cap_triplets > 0) GEMNET_REQUIRE(g->id3_expand_ba && g->id3_reduce_ca, "null triplet array");
  if (!g->triplets_only) {
    GEMNET_REQUIRE(g->int_off && g->intm_off_a && g->intm_off_b && g->quad_ptr, "null quadruplet workspace");
    if (g->cap_int_edges > 0) GEMNET_REQUIRE(g->id4_int_a && g->id4_int_b, "null interaction-edge array");
    if (g->cap_intm > 0)
      GEMNET_REQUIRE(g->id4_reduce_intm_ca && g->id4_reduce_intm_ab && g->id4_expand_intm_db && g->id4_expand_intm_ab,
                     "null intermediate-triplet array");
  }
  cudaStream_t st = as_stream(stream);
  edges_kernel<<<flat_grid(g->cap_atoms, 128), 128, 0, st>>>(*g);
  incoming_kernel<<<flat_grid(g->cap_atoms, 128), 128, 0, st>>>(*g);
  edge_scan_kernel<<<1, kScan, 0, st>>>(*g);
  const int64_t items = g->cap_edges > g->cap_int_edges ? g->cap_edges : g->cap_int_edges;
  triplets_kernel<<<flat_grid(items, 128), 128, 0, st>>>(*g);
  if (!g->triplets_only) {
    quad_count_kernel<<<flat_grid(g->cap_edges, 128), 128, 0, st>>>(*g);
    quad_ptr_kernel<<<1, kScan, 0, st>>>(*g);
  }
  return check_launch("gemnet_graph_device_edges");
}

GEMNET_API int gemnet_graph_device_quads(const gemnet_device_graph* g, void* stream) {
  GD_REQUIRE_COMMON(g);
  GEMNET_REQUIRE(!g->triplets_only, "quadruplets were not requested (triplets_only)");
  GEMNET_REQUIRE(g->int_off && g->intm_off_a && g->intm_off_b && g->quad_ptr, "null workspace");
  GEMNET_REQUIRE(g->cap_edges == 0 || (g->incoming_pos && g->id_a && g->id_c), "null edge array");
  if (g->cap_quads > 0)
    GEMNET_REQUIRE(g->id4_reduce_ca && g->id4_expand_db && g->id4_reduce_cab && g->id4_expand_abd && g->Kidx4,
                   "null quadruplet array");
  quad_emit_kernel<<<flat_grid(g->cap_edges, 128), 128, 0, as_stream(stream)>>>(*g);
  return check_launch("gemnet_graph_device_quads");
}

GEMNET_API int gemnet_graph_device_pad(const gemnet_device_graph* g, int32_t* Z, int32_t* N_out, void* stream) {
  GD_REQUIRE_COMMON(g);
  GEMNET_REQUIRE(g->triplets_only, "the pad pattern is defined for triplet models");
  GEMNET_REQUIRE(Z && N_out && g->R32 && g->batch_seg && g->n_per_mol, "null atom-level array");
  GEMNET_REQUIRE(g->id_c && g->id_a && g->id_undir && g->id_swap && g->id3_expand_ba && g->id3_reduce_ca,
                 "null index array");
  GEMNET_REQUIRE((g->cap_edges & 1) == 0, "edge capacity must be even");
  int64_t items = g->cap_edges > g->cap_triplets ? g->cap_edges : g->cap_triplets;
  if (g->cap_atoms > items) items = g->cap_atoms;
  pad_kernel<<<flat_grid(items, 256), 256, 0, as_stream(stream)>>>(*g, Z, N_out);
  return check_launch("gemnet_graph_device_pad");
}
