// N2 / H0 on the device: per-item bodies of the index-construction kernels (graph_device.cu).
//
// Replaces DataContainer.__getitem__ / get_triplets / get_quadruplets (reference
// gemnet/training/data_container.py:156-408, 411-489) for batches that are already on the GPU, with the SAME output as
// the native host builder (graph_host.cpp) element for element: every array is produced at its final position
// (counting -> exclusive scan -> emission), there are no atomics and no sort, so the order is deterministic.
//
// The bodies are plain functions of the work-item index so that the identical code also compiles as host C++:
// tests/emu/graph_device_emu.cpp runs them sequentially and the CPU test suite compares that run with the host builder
// and the reference-generated golden graphs (the kernels in graph_device.cu only add the grid-stride loops and the
// block-wide scans).
//
//   item            body                produces
//   atom i          atom_count          batch_seg, R32, #upper / #incoming / #interaction neighbours of i
//   atom i          edges               id_a, id_c, id_undir, id_swap (both halves, data_container.py:289-308), id4_int_a/b
//   atom x          incoming            edges into x by ascending source atom (row x of the reference's edge_ids CSR,
//                                       :311-315) + the position of every edge in its target's list
//   edge r          triplets            id3_expand_ba, id3_reduce_ca, Kidx3 (:411-425, :318-332), grouped by r
//   int. edge i     intermediate        id4_{reduce_intm_ca, reduce_intm_ab, expand_intm_db, expand_intm_ab} (:428-453)
//   edge r          quad_scan<emit>     count / emit id4_{reduce_ca, expand_db, reduce_cab, expand_abd}, Kidx4 (:455-489,
//                                       :354-379): for the reduce edge c->a the candidates are the incoming edges d->b of
//                                       the interaction neighbours b of a — one contiguous range of the intermediate list
//   pad item        pad_*               the inert pad pattern of graph.pad_batch (static-shape serving)
#pragma once
#include <stdint.h>

#include "../../include/gemnet_b200.h"

#if defined(__CUDACC__)
#define GD_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define GD_HD inline
#endif

namespace gemnet {
namespace gd {

// header words (device int32[GEMNET_GRAPH_HDR_WORDS]); the first two are written by the caller
enum {
  H_ATOMS = 0, H_MOLS = 1, H_UPPER = 2, H_EDGES = 3, H_TRIPLETS = 4, H_INT_EDGES = 5, H_INTM = 6, H_QUADS = 7,
  H_OVERFLOW = 8, H_AMAX = 9
};

// |R_i - R_j| exactly as np.linalg.norm(R[:,None]-R[None,:], axis=-1) evaluates it in float64
// (data_container.py:258): sqrt((dx*dx + dy*dy) + dz*dz), every operation rounded on its own (no FMA contraction).
GD_HD double pair_distance(const double* R, int32_t i, int32_t j) {
  const double* a = R + 3 * (int64_t)i;
  const double* b = R + 3 * (int64_t)j;
#if defined(__CUDA_ARCH__)
  const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
#else
  const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  volatile double sx = dx * dx, sy = dy * dy, sz = dz * dz;
  volatile double s = sx + sy;
  s = s + sz;
  return std::sqrt(s);
#endif
}

// molecule of atom i: the largest m with mol_off[m] <= i (empty molecules are skipped over)
GD_HD int32_t molecule_of(const int32_t* mol_off, int32_t n_mol, int32_t i) {
  int32_t lo = 0, hi = n_mol + 1;  // first m with mol_off[m] > i
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (mol_off[mid] <= i) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

// atom i: neighbour counts.  Returns max |float(R_i)| (the pad geometry sits beyond the real atoms).
GD_HD float atom_count(const gemnet_device_graph& g, const int32_t* hdr, int32_t i) {
  const int32_t m = molecule_of(g.mol_off, hdr[H_MOLS], i);
  const int32_t b = g.mol_off[m], e = g.mol_off[m + 1];
  g.batch_seg[i] = m;
  int32_t up = 0, deg = 0, ideg = 0;
  for (int32_t j = b; j < e; ++j) {
    if (j == i) continue;
    const double d = pair_distance(g.R, i, j);
    if (d <= g.cutoff) {
      ++deg;
      up += j > i;
    }
    if (!g.triplets_only && d <= g.int_cutoff) ++ideg;
  }
  g.atom_cnt[i] = up;
  g.atom_cnt[g.cap_atoms + i] = deg;
  g.atom_cnt[2 * g.cap_atoms + i] = ideg;
  float amax = 0.f;
  for (int k = 0; k < 3; ++k) {
    const float v = (float)g.R[3 * (int64_t)i + k];
    if (g.R32) g.R32[3 * (int64_t)i + k] = v;
    const float av = v < 0.f ? -v : v;
    amax = av > amax ? av : amax;
  }
  return amax;
}

// atom i: the embedding edges whose smaller end is i — first half (target i < source j, row-major) and the reversed
// copy in the second half (data_container.py:289-308) — and the interaction edges with target i (:236-241).
GD_HD void edges(const gemnet_device_graph& g, const int32_t* hdr, int32_t i) {
  const int32_t m = g.batch_seg[i];
  const int32_t b = g.mol_off[m], e = g.mol_off[m + 1];
  const int32_t nu = hdr[H_UPPER];
  int32_t p = g.upper_off[i];
  int32_t q = g.triplets_only ? 0 : g.int_off[i];
  for (int32_t j = b; j < e; ++j) {
    if (j == i) continue;
    const double d = pair_distance(g.R, i, j);
    if (j > i && d <= g.cutoff) {
      g.id_a[p] = i;
      g.id_c[p] = j;
      g.id_a[nu + p] = j;
      g.id_c[nu + p] = i;
      g.id_undir[p] = p;
      g.id_undir[nu + p] = p;
      g.id_swap[p] = nu + p;
      g.id_swap[nu + p] = p;
      ++p;
    }
    if (!g.triplets_only && d <= g.int_cutoff) {
      g.id4_int_a[q] = i;
      g.id4_int_b[q] = j;
      ++q;
    }
  }
}

// atom x: the edges k -> x by ascending k.  Sources k < x are second-half edges (found in k's first-half row, whose
// sources ascend, by binary search); sources k > x are x's own first-half row.
GD_HD void incoming(const gemnet_device_graph& g, const int32_t* hdr, int32_t x) {
  const int32_t nu = hdr[H_UPPER];
  const int32_t first = g.mol_off[g.batch_seg[x]];
  const int32_t q0 = g.in_ptr[x];
  int32_t q = q0;
  for (int32_t k = first; k < x; ++k) {
    int32_t lo = g.upper_off[k];
    const int32_t end = g.upper_off[k + 1];
    int32_t hi = end;
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if (g.id_c[mid] < x) lo = mid + 1; else hi = mid;
    }
    if (lo < end && g.id_c[lo] == x) {
      g.incoming[q] = nu + lo;
      g.incoming_pos[nu + lo] = q - q0;
      ++q;
    }
  }
  for (int32_t p = g.upper_off[x]; p < g.upper_off[x + 1]; ++p) {
    g.incoming[q] = p;
    g.incoming_pos[p] = q - q0;
    ++q;
  }
}

GD_HD int32_t in_degree(const gemnet_device_graph& g, int32_t atom) { return g.in_ptr[atom + 1] - g.in_ptr[atom]; }

// edge r = (c -> a): the triplets c -> a <- b, b != c, in the order of a's incoming list
GD_HD void triplets(const gemnet_device_graph& g, int32_t r) {
  const int32_t a = g.id_a[r];
  int32_t pos = g.rowptr3[r];
  int32_t k = 0;
  for (int32_t q = g.in_ptr[a]; q < g.in_ptr[a + 1]; ++q) {
    const int32_t eb = g.incoming[q];
    if (eb == r) continue;  // b == c: the reduce edge itself (edges are unique per ordered atom pair)
    g.id3_expand_ba[pos] = eb;
    g.id3_reduce_ca[pos] = r;
    g.Kidx3[pos] = k++;
    ++pos;
  }
}

// interaction edge i = (a, b): incoming edges of a (c -> a) and of b (d -> b)
GD_HD void intermediate(const gemnet_device_graph& g, int32_t i) {
  const int32_t a = g.id4_int_a[i], b = g.id4_int_b[i];
  int32_t o = g.intm_off_a[i];
  for (int32_t q = g.in_ptr[a]; q < g.in_ptr[a + 1]; ++q, ++o) {
    g.id4_reduce_intm_ca[o] = g.incoming[q];
    g.id4_reduce_intm_ab[o] = i;
  }
  o = g.intm_off_b[i];
  for (int32_t q = g.in_ptr[b]; q < g.in_ptr[b + 1]; ++q, ++o) {
    g.id4_expand_intm_db[o] = g.incoming[q];
    g.id4_expand_intm_ab[o] = i;
  }
}

// reduce edge r = (c -> a): quadruplets c -> a - b <- d with c != b, a != d, c != d (data_container.py:470-479), in
// (interaction edge, d -> b) order.  EMIT = false: returns the count; EMIT = true: writes them at quad_ptr[r].
template <bool EMIT>
GD_HD int32_t quad_scan(const gemnet_device_graph& g, int32_t r) {
  const int32_t a = g.id_a[r], c = g.id_c[r];
  const int32_t i0 = g.int_off[a], i1 = g.int_off[a + 1];
  const int32_t jd0 = g.intm_off_b[i0], jd1 = g.intm_off_b[i1];
  const int32_t base = EMIT ? g.quad_ptr[r] : 0;
  const int32_t my_pos = EMIT ? g.incoming_pos[r] : 0;
  int32_t k = 0;
  for (int32_t jd = jd0; jd < jd1; ++jd) {
    const int32_t ed = g.id4_expand_intm_db[jd];
    const int32_t i = g.id4_expand_intm_ab[jd];
    const int32_t b = g.id4_int_b[i];
    const int32_t d = g.id_c[ed];
    if (c == b || a == d || c == d) continue;
    if (EMIT) {
      const int32_t pos = base + k;
      g.id4_reduce_ca[pos] = r;
      g.id4_expand_db[pos] = ed;
      g.id4_reduce_cab[pos] = g.intm_off_a[i] + my_pos;
      g.id4_expand_abd[pos] = jd;
      g.Kidx4[pos] = k;
    }
    ++k;
  }
  return k;
}

// ---- static-shape serving: the inert pad pattern of graph.pad_batch / gemnet_pad_batch_host, written on the device ----
struct PadSizes {
  int32_t A, M, E, T;          // real atoms, molecules, edges, triplets (device header)
  int32_t n_pad, n_tri, n_pe, nt, per;
  float far;
};

GD_HD PadSizes pad_sizes(const gemnet_device_graph& g, const int32_t* hdr) {
  PadSizes s;
  s.A = hdr[H_ATOMS]; s.M = hdr[H_MOLS]; s.E = hdr[H_EDGES]; s.T = hdr[H_TRIPLETS];
  s.n_pad = (int32_t)g.cap_atoms - s.A;
  s.n_tri = s.n_pad / 3 > 0 ? s.n_pad / 3 : 1;
  s.n_pe = (int32_t)g.cap_edges - s.E;
  s.nt = (int32_t)g.cap_triplets - s.T;
  const int32_t n_full = (s.n_pe / 6) * 6;
  s.per = n_full > 0 ? (s.nt + n_full - 1) / n_full : 1;
  if (s.per < 1) s.per = 1;
  union { int32_t i; float f; } u;
  u.i = hdr[H_AMAX];
  s.far = (s.A ? u.f : 0.f) + 100.0f;
  return s;
}

GD_HD void pad_atom(const gemnet_device_graph& g, const PadSizes& s, int32_t* Z, int32_t k) {  // k in [0, n_pad)
  const float xyz[3][3] = {{0.0f, 0.0f, 0.0f}, {1.2f, 0.0f, 0.0f}, {0.5f, 1.1f, 0.0f}};
  const int32_t at = s.A + k;
  Z[at] = 1;
  g.batch_seg[at] = (int32_t)g.cap_molecules;
  const int v = k < 3 * s.n_tri ? k % 3 : 0;
  const int32_t tri = k / 3 < s.n_tri - 1 ? k / 3 : s.n_tri - 1;
  g.R32[3 * (int64_t)at + 0] = (xyz[v][0] + s.far) + 10.0f * (float)tri;
  g.R32[3 * (int64_t)at + 1] = xyz[v][1] + s.far;
  g.R32[3 * (int64_t)at + 2] = xyz[v][2] + s.far;
}

GD_HD void pad_molecule(const gemnet_device_graph& g, const PadSizes& s, int32_t* N_out, int32_t m) {  // m in [0, cap_molecules]
  N_out[m] = m < s.M ? g.n_per_mol[m] : (m == (int32_t)g.cap_molecules ? s.n_pad : 0);
}

GD_HD void pad_edge(const gemnet_device_graph& g, const PadSizes& s, int32_t j) {  // j in [0, n_pe)
  const int32_t kC[6] = {0, 1, 0, 2, 1, 2}, kA[6] = {1, 0, 2, 0, 2, 1};
  const int32_t base = s.A + 3 * ((j / 6) % s.n_tri);
  g.id_c[s.E + j] = base + kC[j % 6];
  g.id_a[s.E + j] = base + kA[j % 6];
  g.id_swap[s.E + j] = s.E + (j ^ 1);
  g.id_undir[s.E + j] = s.E / 2 + j / 2;
}

GD_HD void pad_triplet(const gemnet_device_graph& g, const PadSizes& s, int32_t j) {  // j in [0, nt)
  const int32_t kPartner[6] = {5, 3, 4, 1, 2, 0};
  const int32_t r = j / s.per;
  g.id3_reduce_ca[s.T + j] = s.E + r;
  g.id3_expand_ba[s.T + j] = s.E + (r / 6) * 6 + kPartner[r % 6];
  if (g.Kidx3) g.Kidx3[s.T + j] = j % s.per;
}

}  // namespace gd
}  // namespace gemnet
