// TEST INFRASTRUCTURE.  Host emulation of the device index builder (gemnet_pytorch_b200/csrc/graph_device.cu): the
// per-item bodies are the SAME code (graph_device_bodies.h compiled as host C++), run here sequentially with plain
// loops where the kernels have grid-stride loops and with sequential prefix sums where they have block scans.  Same
// descriptor, same stages, host pointers.  Built by tests/test_graph_device_cpu.py with g++; never loaded by the package.
#include <cstdint>
#include <cstring>

#include "../../gemnet_pytorch_b200/csrc/graph_device_bodies.h"

using namespace gemnet::gd;

namespace {

template <typename F>
long long scan(F f, int32_t n, int32_t* out) {
  long long carry = 0;
  for (int32_t i = 0; i < n; ++i) {
    const long long v = (long long)f(i);  // read before the (possibly aliasing) write
    if (out) out[i] = (int32_t)carry;
    carry += v;
  }
  if (out) out[n] = (int32_t)carry;
  return carry;
}

}  // namespace

extern "C" int emu_graph_device_count(const gemnet_device_graph* gp) {
  const gemnet_device_graph& g = *gp;
  int32_t* hdr = g.hdr;
  int32_t n_mol = hdr[H_MOLS];
  const int32_t n_atoms = hdr[H_ATOMS];
  const bool bad = n_mol < 0 || n_mol > g.cap_molecules || n_atoms < 0 || n_atoms > g.cap_atoms;
  if (bad) n_mol = 0;
  int neg = 0;
  const long long tot = scan(
      [&](int32_t m) {
        const int32_t v = g.n_per_mol[m];
        if (v < 0) neg = 1;
        return v > 0 ? v : 0;
      },
      n_mol, g.mol_off);
  for (int k = H_UPPER; k < GEMNET_GRAPH_HDR_WORDS; ++k) hdr[k] = 0;
  if (bad || neg || tot != (long long)n_atoms) hdr[H_OVERFLOW] = 1;
  if (hdr[H_OVERFLOW]) return 0;

  float amax = 0.f;
  for (int32_t i = 0; i < n_atoms; ++i) {
    const float a = atom_count(g, hdr, i);
    amax = a > amax ? a : amax;
  }
  std::memcpy(&hdr[H_AMAX], &amax, 4);

  const int32_t n = n_atoms;
  const int32_t* up = g.atom_cnt;
  const int32_t* deg = g.atom_cnt + g.cap_atoms;
  const int32_t* ideg = g.atom_cnt + 2 * g.cap_atoms;
  const long long nu = scan([&](int32_t i) { return up[i]; }, n, g.upper_off);
  scan([&](int32_t i) { return deg[i]; }, n, g.in_ptr);
  const long long nt = scan([&](int32_t i) { return (long long)deg[i] * (deg[i] - 1); }, n, nullptr);
  long long ni = 0, nintm = 0;
  if (!g.triplets_only) {
    ni = scan([&](int32_t i) { return ideg[i]; }, n, g.int_off);
    nintm = scan([&](int32_t i) { return (long long)ideg[i] * deg[i]; }, n, nullptr);
  }
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
  return 0;
}

extern "C" int emu_graph_device_edges(const gemnet_device_graph* gp) {
  const gemnet_device_graph& g = *gp;
  int32_t* hdr = g.hdr;
  if (hdr[H_OVERFLOW]) return 0;
  const int32_t n = hdr[H_ATOMS];
  for (int32_t i = 0; i < n; ++i) edges(g, hdr, i);
  for (int32_t x = 0; x < n; ++x) incoming(g, hdr, x);
  const int32_t ne = hdr[H_EDGES], ni = hdr[H_INT_EDGES];
  scan([&](int32_t r) { return in_degree(g, g.id_a[r]) - 1; }, ne, g.rowptr3);
  if (!g.triplets_only) {
    scan([&](int32_t i) { return in_degree(g, g.id4_int_a[i]); }, ni, g.intm_off_a);
    scan([&](int32_t i) { return in_degree(g, g.id4_int_b[i]); }, ni, g.intm_off_b);
  }
  for (int32_t r = 0; r < ne; ++r) triplets(g, r);
  if (!g.triplets_only) {
    for (int32_t i = 0; i < ni; ++i) intermediate(g, i);
    for (int32_t r = 0; r < ne; ++r) g.quad_ptr[r] = quad_scan<false>(g, r);
    const int32_t* cnt = g.quad_ptr;
    const long long nq = scan([&](int32_t r) { return cnt[r]; }, ne, g.quad_ptr);
    hdr[H_QUADS] = nq > 0x7fffffffLL ? 0x7fffffff : (int32_t)nq;
    if (nq > g.cap_quads || nq > 0x7fffffffLL) hdr[H_OVERFLOW] = 32;
  }
  return 0;
}

extern "C" int emu_graph_device_quads(const gemnet_device_graph* gp) {
  const gemnet_device_graph& g = *gp;
  if (g.hdr[H_OVERFLOW]) return 0;
  const int32_t ne = g.hdr[H_EDGES];
  for (int32_t r = 0; r < ne; ++r) quad_scan<true>(g, r);
  return 0;
}

extern "C" int emu_graph_device_pad(const gemnet_device_graph* gp, int32_t* Z, int32_t* N_out) {
  const gemnet_device_graph& g = *gp;
  if (g.hdr[H_OVERFLOW]) return 0;
  const PadSizes s = pad_sizes(g, g.hdr);
  if (s.n_pad < 3 || s.n_pe < 6 || s.nt < 0) {
    g.hdr[H_OVERFLOW] |= 64;
    return 0;
  }
  for (int32_t k = 0; k < s.n_pad; ++k) pad_atom(g, s, Z, k);
  for (int32_t m = 0; m <= (int32_t)g.cap_molecules; ++m) pad_molecule(g, s, N_out, m);
  for (int32_t j = 0; j < s.n_pe; ++j) pad_edge(g, s, j);
  for (int32_t j = 0; j < s.nt; ++j) pad_triplet(g, s, j);
  return 0;
}
