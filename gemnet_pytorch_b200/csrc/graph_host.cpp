// H0: native host index builder — replaces DataContainer.__getitem__ / get_triplets / get_quadruplets
// (reference gemnet/training/data_container.py:156-408, 411-489; numba helpers :520-565) which go through
// scipy.sparse fancy indexing, np.argsort and numba.  Plain C++: per-molecule dense distance test in fp64
// (as the reference, :258-267), then counting-sort style emission that is already grouped by the reduce edge,
// so no argsort is needed.  Semantics documented in include/gemnet_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gemnet_b200.h"

namespace gemnet {
void set_error(const char* fmt, ...);
}

struct gemnet_graph {
  std::map<std::string, std::vector<int32_t>> arrays;
};

namespace {

using ivec = std::vector<int32_t>;

// all directed pairs (t, s), t != s, |R_t - R_s| <= cutoff, molecule by molecule, row-major
void directed_pairs(const double* R, const int32_t* n_per_mol, int64_t n_mol, double cutoff, ivec& T, ivec& S) {
  int64_t off = 0;
  for (int64_t m = 0; m < n_mol; ++m) {
    const int32_t n = n_per_mol[m];
    for (int32_t i = 0; i < n; ++i) {
      const double* ri = R + 3 * (off + i);
      for (int32_t j = 0; j < n; ++j) {
        if (i == j) continue;
        const double* rj = R + 3 * (off + j);
        const double dx = ri[0] - rj[0], dy = ri[1] - rj[1], dz = ri[2] - rj[2];
        // same operation order as np.linalg.norm(..., axis=-1): sqrt((dx^2 + dy^2) + dz^2), no FMA contraction
        volatile double sx = dx * dx, sy = dy * dy, sz = dz * dz;
        const double d = std::sqrt((sx + sy) + sz);
        if (d <= cutoff) {
          T.push_back((int32_t)(off + i));
          S.push_back((int32_t)(off + j));
        }
      }
    }
    off += n;
  }
}

// Run fn(chunk) for chunk in [0, n_chunks) on up to `max_threads` host threads (chunk c on thread c % T).
template <typename Fn>
void parallel_chunks(int n_chunks, int max_threads, Fn fn) {
  const int T = std::max(1, std::min(n_chunks, max_threads));
  if (T == 1) {
    for (int c = 0; c < n_chunks; ++c) fn(c);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(T);
  for (int t = 0; t < T; ++t)
    th.emplace_back([=] {
      for (int c = t; c < n_chunks; c += T) fn(c);
    });
  for (auto& x : th) x.join();
}

int host_threads() {
  static const int n = [] {
    if (const char* e = getenv("GEMNET_B200_HOST_THREADS")) return std::max(1, atoi(e));
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hc ? hc : 1u, 16u));
  }();
  return n;
}

// One retired graph object is kept for reuse: its vectors keep their capacity, so the next batch of similar size
// writes into already-mapped memory instead of page-faulting through ~20 MB of fresh allocations (10 ms of a 30 ms
// quadruplet build).
std::mutex g_pool_mutex;
gemnet_graph* g_pool = nullptr;

gemnet_graph* acquire_graph() {
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pool) {
      gemnet_graph* g = g_pool;
      g_pool = nullptr;
      for (auto& kv : g->arrays) kv.second.clear();
      return g;
    }
  }
  return new gemnet_graph();
}

void release_graph(gemnet_graph* g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pool) {
      g_pool = g;
      return;
    }
  }
  delete g;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int gemnet_graph_build_host(const double* R, const int32_t* n_per_mol,
                                                                               int64_t n_mol, double cutoff,
                                                                               double int_cutoff, int32_t triplets_only,
                                                                               gemnet_graph** out) {
  if (!out || n_mol < 0 || (n_mol > 0 && (!R || !n_per_mol))) {
    gemnet::set_error("gemnet_graph_build_host: bad arguments");
    return 1;
  }
  auto* g = acquire_graph();
  auto& A = g->arrays;
  int64_t n_atoms = 0;
  for (int64_t m = 0; m < n_mol; ++m) {
    if (n_per_mol[m] < 0) {
      delete g;
      gemnet::set_error("gemnet_graph_build_host: negative molecule size");
      return 1;
    }
    n_atoms += n_per_mol[m];
  }
  ivec& batch_seg = A["batch_seg"];
  batch_seg.reserve(n_atoms);
  for (int64_t m = 0; m < n_mol; ++m)
    for (int32_t i = 0; i < n_per_mol[m]; ++i) batch_seg.push_back((int32_t)m);

  const char* keys3[] = {"id_undir", "id_swap", "id_c", "id_a", "id3_expand_ba", "id3_reduce_ca", "Kidx3"};
  const char* keys4[] = {"id4_int_b", "id4_int_a", "id4_reduce_ca", "id4_expand_db", "id4_reduce_cab",
                         "id4_expand_abd", "Kidx4", "id4_reduce_intm_ca", "id4_expand_intm_db",
                         "id4_reduce_intm_ab", "id4_expand_intm_ab"};
  for (const char* k : keys3) A[k];
  if (!triplets_only)
    for (const char* k : keys4) A[k];

  ivec pt, ps;
  directed_pairs(R, n_per_mol, n_mol, cutoff, pt, ps);
  if (pt.empty()) {  // data_container.py:282-285: every index array empty
    *out = g;
    return 0;
  }

  // ---- edges: [upper triangle (t < s), row-major] ++ [reversed]   (:289-308)
  ivec ut, us;
  for (size_t p = 0; p < pt.size(); ++p)
    if (pt[p] < ps[p]) {
      ut.push_back(pt[p]);
      us.push_back(ps[p]);
    }
  const int32_t nu = (int32_t)ut.size();
  const int32_t nE = 2 * nu;
  ivec& id_a = A["id_a"];
  ivec& id_c = A["id_c"];
  ivec& id_undir = A["id_undir"];
  ivec& id_swap = A["id_swap"];
  id_a.resize(nE); id_c.resize(nE); id_undir.resize(nE); id_swap.resize(nE);
  for (int32_t p = 0; p < nu; ++p) {
    id_a[p] = ut[p]; id_c[p] = us[p];
    id_a[nu + p] = us[p]; id_c[nu + p] = ut[p];
    id_undir[p] = p; id_undir[nu + p] = p;
    id_swap[p] = nu + p; id_swap[nu + p] = p;
  }

  // ---- incoming edge lists per atom, ascending source atom (row x of the reference's edge_ids CSR, :311-315)
  std::vector<int32_t> ptr(n_atoms + 1, 0);
  for (int32_t e = 0; e < nE; ++e) ptr[id_a[e] + 1]++;
  for (int64_t x = 0; x < n_atoms; ++x) ptr[x + 1] += ptr[x];
  ivec inc(nE);
  {
    std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
    for (int32_t p = 0; p < nu; ++p) inc[cur[us[p]]++] = nu + p;  // sources k < x (second-half edges), ascending k
    for (int32_t p = 0; p < nu; ++p) inc[cur[ut[p]]++] = p;       // sources k > x (first-half edges), ascending k
  }

  // ---- triplets c -> a <- b, b != c, grouped by the reduce edge (c -> a)   (:411-425, :318-332)
  {
    ivec& exp = A["id3_expand_ba"];
    ivec& red = A["id3_reduce_ca"];
    ivec& kidx = A["Kidx3"];
    size_t total = 0;
    for (int32_t r = 0; r < nE; ++r) total += (size_t)(ptr[id_a[r] + 1] - ptr[id_a[r]]) - 1;
    exp.reserve(total); red.reserve(total); kidx.reserve(total);
    for (int32_t r = 0; r < nE; ++r) {
      const int32_t a = id_a[r], c = id_c[r];
      int32_t k = 0;
      for (int32_t q = ptr[a]; q < ptr[a + 1]; ++q) {
        const int32_t eb = inc[q];
        if (id_c[eb] == c) continue;
        exp.push_back(eb); red.push_back(r); kidx.push_back(k++);
      }
    }
  }
  if (triplets_only) {
    *out = g;
    return 0;
  }

  // ---- quadruplets c -> a - b <- d   (:428-489, :354-379)
  ivec it, is;
  directed_pairs(R, n_per_mol, n_mol, int_cutoff, it, is);
  const int64_t nI = (int64_t)it.size();
  A["id4_int_a"] = it;
  A["id4_int_b"] = is;
  ivec& intm_ca = A["id4_reduce_intm_ca"];
  ivec& intm_db = A["id4_expand_intm_db"];
  ivec& red_intm_ab = A["id4_reduce_intm_ab"];
  ivec& exp_intm_ab = A["id4_expand_intm_ab"];
  std::vector<int64_t> off_t(nI + 1, 0), off_s(nI + 1, 0);
  for (int64_t i = 0; i < nI; ++i) {
    off_t[i + 1] = off_t[i] + (ptr[it[i] + 1] - ptr[it[i]]);
    off_s[i + 1] = off_s[i] + (ptr[is[i] + 1] - ptr[is[i]]);
  }
  intm_ca.resize(off_t[nI]); red_intm_ab.resize(off_t[nI]);
  intm_db.resize(off_s[nI]); exp_intm_ab.resize(off_s[nI]);
  for (int64_t i = 0; i < nI; ++i) {
    for (int32_t q = ptr[it[i]], j = 0; q < ptr[it[i] + 1]; ++q, ++j) {
      intm_ca[off_t[i] + j] = inc[q];
      red_intm_ab[off_t[i] + j] = (int32_t)i;
    }
    for (int32_t q = ptr[is[i]], j = 0; q < ptr[is[i] + 1]; ++q, ++j) {
      intm_db[off_s[i] + j] = inc[q];
      exp_intm_ab[off_s[i] + j] = (int32_t)i;
    }
  }
  // Quadruplets of different molecules touch disjoint reduce edges, so the interaction edges are cut at molecule
  // boundaries into chunks of similar work and the two passes (count, emit) run chunk-parallel on the host threads.
  auto valid = [&](int64_t i, int32_t ec, int32_t ed) {
    const int32_t c = id_c[ec], a = it[i], b = is[i], d = id_c[ed];
    return c != b && a != d && c != d;
  };
  std::vector<int64_t> chunk_begin{0};
  {
    const int64_t total_work = [&] {
      int64_t w = 0;
      for (int64_t i = 0; i < nI; ++i) w += (off_t[i + 1] - off_t[i]) * (off_s[i + 1] - off_s[i]);
      return w;
    }();
    const int64_t target = total_work / (8 * host_threads()) + 1;
    int64_t acc = 0;
    for (int64_t i = 0; i < nI; ++i) {
      if (i > 0 && batch_seg[it[i]] != batch_seg[it[i - 1]] && acc >= target) {
        chunk_begin.push_back(i);
        acc = 0;
      }
      acc += (off_t[i + 1] - off_t[i]) * (off_s[i + 1] - off_s[i]);
    }
    chunk_begin.push_back(nI);
  }
  const int n_chunks = (int)chunk_begin.size() - 1;
  // pass 1: count quadruplets per reduce edge
  std::vector<int64_t> qptr(nE + 1, 0);
  parallel_chunks(n_chunks, host_threads(), [&](int ch) {
    for (int64_t i = chunk_begin[ch]; i < chunk_begin[ch + 1]; ++i)
      for (int64_t jd = off_s[i]; jd < off_s[i + 1]; ++jd)
        for (int64_t jc = off_t[i]; jc < off_t[i + 1]; ++jc)
          if (valid(i, intm_ca[jc], intm_db[jd])) qptr[intm_ca[jc] + 1]++;
  });
  for (int32_t r = 0; r < nE; ++r) qptr[r + 1] += qptr[r];
  const int64_t nQ = qptr[nE];
  if (nQ > INT32_MAX) {
    delete g;
    gemnet::set_error("gemnet_graph_build_host: more than 2^31-1 quadruplets in one batch");
    return 1;
  }
  ivec& red_ca = A["id4_reduce_ca"];
  ivec& exp_db = A["id4_expand_db"];
  ivec& red_cab = A["id4_reduce_cab"];
  ivec& exp_abd = A["id4_expand_abd"];
  ivec& kidx4 = A["Kidx4"];
  red_ca.resize(nQ); exp_db.resize(nQ); red_cab.resize(nQ); exp_abd.resize(nQ); kidx4.resize(nQ);
  {
    // pass 2: emit, grouped by the reduce edge; within a group the order is the reference's (i, jd, jc) scan order
    std::vector<int64_t> cur(qptr.begin(), qptr.end() - 1);
    parallel_chunks(n_chunks, host_threads(), [&](int ch) {
      for (int64_t i = chunk_begin[ch]; i < chunk_begin[ch + 1]; ++i)
        for (int64_t jd = off_s[i]; jd < off_s[i + 1]; ++jd)
          for (int64_t jc = off_t[i]; jc < off_t[i + 1]; ++jc) {
            const int32_t ec = intm_ca[jc], ed = intm_db[jd];
            if (!valid(i, ec, ed)) continue;
            const int64_t pos = cur[ec]++;
            red_ca[pos] = ec; exp_db[pos] = ed;
            red_cab[pos] = (int32_t)jc; exp_abd[pos] = (int32_t)jd;
            kidx4[pos] = (int32_t)(pos - qptr[ec]);
          }
    });
  }
  *out = g;
  return 0;
}

// Sizes of a batch without building it: per-atom in-degree from the same distance test as directed_pairs; every
// directed edge r = (c -> a) carries deg(a) - 1 triplets.
extern "C" __attribute__((visibility("default"))) int gemnet_graph_count_host(const double* R, const int32_t* n_per_mol,
                                                                               int64_t n_mol, double cutoff,
                                                                               int64_t* n_edges, int64_t* n_triplets) {
  if (n_mol < 0 || (n_mol > 0 && (!R || !n_per_mol)) || !n_edges || !n_triplets) {
    gemnet::set_error("gemnet_graph_count_host: bad arguments");
    return 1;
  }
  int64_t off = 0, ne = 0, nt = 0;
  for (int64_t m = 0; m < n_mol; ++m) {
    const int32_t n = n_per_mol[m];
    if (n < 0) {
      gemnet::set_error("gemnet_graph_count_host: negative molecule size");
      return 1;
    }
    for (int32_t i = 0; i < n; ++i) {
      const double* ri = R + 3 * (off + i);
      int64_t deg = 0;
      for (int32_t j = 0; j < n; ++j) {
        if (i == j) continue;
        const double* rj = R + 3 * (off + j);
        const double dx = ri[0] - rj[0], dy = ri[1] - rj[1], dz = ri[2] - rj[2];
        volatile double sx = dx * dx, sy = dy * dy, sz = dz * dz;
        deg += std::sqrt((sx + sy) + sz) <= cutoff;
      }
      ne += deg;
      nt += deg * (deg - 1);
    }
    off += n;
  }
  *n_edges = ne;
  *n_triplets = nt;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int64_t gemnet_graph_size(const gemnet_graph* g, const char* name) {
  if (!g || !name) return -1;
  auto it = g->arrays.find(name);
  return it == g->arrays.end() ? -1 : (int64_t)it->second.size();
}

extern "C" __attribute__((visibility("default"))) int gemnet_graph_copy_host(const gemnet_graph* g, const char* name,
                                                                              int32_t* dst) {
  if (!g || !name) {
    gemnet::set_error("gemnet_graph_copy_host: null argument");
    return 1;
  }
  auto it = g->arrays.find(name);
  if (it == g->arrays.end()) {
    gemnet::set_error("gemnet_graph_copy_host: unknown array '%s'", name);
    return 1;
  }
  if (!it->second.empty()) {
    if (!dst) {
      gemnet::set_error("gemnet_graph_copy_host: null destination");
      return 1;
    }
    std::memcpy(dst, it->second.data(), it->second.size() * sizeof(int32_t));
  }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int gemnet_graph_copy_host_i64(const gemnet_graph* g, const char* name,
                                                                                  int64_t* dst_host) {
  if (!g || !name) {
    gemnet::set_error("gemnet_graph_copy_host_i64: null argument");
    return 1;
  }
  auto it = g->arrays.find(name);
  if (it == g->arrays.end()) {
    gemnet::set_error("gemnet_graph_copy_host_i64: unknown array '%s'", name);
    return 1;
  }
  const std::vector<int32_t>& v = it->second;
  const int64_t n = (int64_t)v.size();
  if (n > 0 && !dst_host) {
    gemnet::set_error("gemnet_graph_copy_host_i64: null destination");
    return 1;
  }
  const int64_t kChunk = 1 << 18;  // widen in parallel: the reference's tensors are int64 (data_container.py:505-512)
  const int n_chunks = (int)((n + kChunk - 1) / kChunk);
  const int32_t* src = v.data();
  parallel_chunks(n_chunks, host_threads(), [=](int c) {
    const int64_t b = (int64_t)c * kChunk, e = std::min(n, b + kChunk);
    for (int64_t i = b; i < e; ++i) dst_host[i] = src[i];
  });
  return 0;
}

extern "C" __attribute__((visibility("default"))) void gemnet_graph_free(gemnet_graph* g) { release_graph(g); }

// ------------------------------------------------------------------------------------------------------------------
// Static-shape serving: write a reference-format (int64 / float32) triplet batch into capacity-sized int32 / float32
// staging arrays and fill the rest with the inert pad pattern (graph.py: pad molecule of far-away triangles whose
// atoms / edges / triplets only reference each other).  One pass per array; the Python implementation
// (graph.pad_batch) is the specification and the test oracle for this function.
extern "C" __attribute__((visibility("default"))) int gemnet_pad_batch_host(
    const int64_t* Z, const float* R, const int64_t* batch_seg, const int64_t* N, const int64_t* id_c, const int64_t* id_a,
    const int64_t* id_swap, const int64_t* id_undir, const int64_t* id3_expand, const int64_t* id3_reduce, int64_t A,
    int64_t E, int64_t T, int64_t M, int64_t capA, int64_t capE, int64_t capT, int64_t capM, int32_t* oZ, float* oR,
    int32_t* oSeg, int32_t* oN, int32_t* o_c, int32_t* o_a, int32_t* o_swap, int32_t* o_undir, int32_t* o_expand,
    int32_t* o_reduce) {
  if (A < 0 || E < 0 || T < 0 || M < 0 || A + 3 > capA || E + 6 > capE || T > capT || M > capM || (capE & 1)) {
    gemnet::set_error("gemnet_pad_batch_host: batch (atoms=%lld, edges=%lld, triplets=%lld, molecules=%lld) exceeds the "
                      "capacity (%lld, %lld, %lld, %lld)", (long long)A, (long long)E, (long long)T, (long long)M,
                      (long long)capA, (long long)capE, (long long)capT, (long long)capM);
    return 1;
  }
  static const int32_t kC[6] = {0, 1, 0, 2, 1, 2}, kA[6] = {1, 0, 2, 0, 2, 1}, kPartner[6] = {5, 3, 4, 1, 2, 0};
  static const float kXYZ[3][3] = {{0.0f, 0.0f, 0.0f}, {1.2f, 0.0f, 0.0f}, {0.5f, 1.1f, 0.0f}};
  // atoms
  float amax = 0.f;
  for (int64_t i = 0; i < A; ++i) {
    oZ[i] = (int32_t)Z[i];
    oSeg[i] = (int32_t)batch_seg[i];
    for (int k = 0; k < 3; ++k) {
      const float v = R[3 * i + k];
      oR[3 * i + k] = v;
      amax = std::fmax(amax, std::fabs(v));
    }
  }
  const int64_t n_pad = capA - A;
  const int64_t n_tri = n_pad / 3 > 0 ? n_pad / 3 : 1;
  const float far = (A ? amax : 0.f) + 100.0f;
  for (int64_t k = 0; k < n_pad; ++k) {
    oZ[A + k] = 1;
    oSeg[A + k] = (int32_t)capM;
    const int v = k < 3 * n_tri ? (int)(k % 3) : 0;
    const int64_t tri = k / 3 < n_tri - 1 ? k / 3 : n_tri - 1;
    oR[3 * (A + k) + 0] = (kXYZ[v][0] + far) + 10.0f * (float)tri;
    oR[3 * (A + k) + 1] = kXYZ[v][1] + far;
    oR[3 * (A + k) + 2] = kXYZ[v][2] + far;
  }
  for (int64_t m = 0; m < M; ++m) oN[m] = (int32_t)N[m];
  for (int64_t m = M; m < capM; ++m) oN[m] = 0;
  oN[capM] = (int32_t)n_pad;
  // edges
  for (int64_t e = 0; e < E; ++e) {
    o_c[e] = (int32_t)id_c[e];
    o_a[e] = (int32_t)id_a[e];
    o_swap[e] = (int32_t)id_swap[e];
    o_undir[e] = (int32_t)id_undir[e];
  }
  const int64_t n_pe = capE - E;
  for (int64_t j = 0; j < n_pe; ++j) {
    const int32_t base = (int32_t)(A + 3 * ((j / 6) % n_tri));
    o_c[E + j] = base + kC[j % 6];
    o_a[E + j] = base + kA[j % 6];
    o_swap[E + j] = (int32_t)(E + (j ^ 1));
    o_undir[E + j] = (int32_t)(E / 2 + j / 2);
  }
  // triplets
  for (int64_t t = 0; t < T; ++t) {
    o_expand[t] = (int32_t)id3_expand[t];
    o_reduce[t] = (int32_t)id3_reduce[t];
  }
  const int64_t nt = capT - T;
  if (nt > 0) {
    const int64_t n_full = (n_pe / 6) * 6;
    const int64_t per = (nt + n_full - 1) / n_full;
    for (int64_t j = 0; j < nt; ++j) {
      const int64_t r = j / per;
      o_reduce[T + j] = (int32_t)(E + r);
      o_expand[T + j] = (int32_t)(E + (r / 6) * 6 + kPartner[r % 6]);
    }
  }
  return 0;
}
