// H10/H11 (edge -> atom Hadamard + segment sum) and H12 (molecule energy sum, direct-force read-out).
// HBM-bound: rows of F floats are streamed with 128-bit loads, one warp per atom, deterministic order.
#include "common.cuh"

namespace gemnet {

// out[a,:] = scale * sum_{k in [rowptr[a],rowptr[a+1])} m[perm[k],:] * r[perm[k],:]   (F % 4 == 0)
// one warp per atom, lanes stride over float4 columns.
__global__ void edge_to_atom_fwd_kernel(const float4* __restrict__ m, const float4* __restrict__ r,
                                        const int32_t* __restrict__ perm, const int32_t* __restrict__ rowptr,
                                        int64_t n_atoms, int F4, float scale, float4* __restrict__ out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t a = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; a < n_atoms; a += warps) {
    const int32_t b = rowptr[a], e = rowptr[a + 1];
    for (int c = lane; c < F4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int32_t k = b; k < e; ++k) {
        const int64_t row = (int64_t)__ldg(perm + k) * F4 + c;
        const float4 mv = __ldg(m + row), rv = __ldg(r + row);
        acc.x += mv.x * rv.x;
        acc.y += mv.y * rv.y;
        acc.z += mv.z * rv.z;
        acc.w += mv.w * rv.w;
      }
      out[a * F4 + c] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
    }
  }
}

__global__ void hadamard_kernel(const float4* __restrict__ m, const float4* __restrict__ r, int64_t n4, float s,
                                float4* __restrict__ out) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(m + i), b = __ldg(r + i);
    out[i] = make_float4(a.x * b.x * s, a.y * b.y * s, a.z * b.z * s, a.w * b.w * s);
  }
}

// g_m[e,:] = (scale*g_out[id_a[e],:] + had_scale*g_had[e,:]) * r[e,:]; g_r likewise with m
__global__ void edge_to_atom_bwd_kernel(const float4* __restrict__ m, const float4* __restrict__ r,
                                        const int32_t* __restrict__ id_a, int64_t n_edges, int F4, float scale,
                                        const float4* __restrict__ g_out, float had_scale,
                                        const float4* __restrict__ g_had, float4* __restrict__ g_m,
                                        float4* __restrict__ g_r) {
  pdl_enter();
  const int64_t total = n_edges * F4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t e = i / F4;
    const int c = (int)(i - e * F4);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g_out) {
      const float4 t = __ldg(g_out + (int64_t)__ldg(id_a + e) * F4 + c);
      g = make_float4(t.x * scale, t.y * scale, t.z * scale, t.w * scale);
    }
    if (g_had) {
      const float4 t = __ldg(g_had + i);
      g.x += t.x * had_scale;
      g.y += t.y * had_scale;
      g.z += t.z * had_scale;
      g.w += t.w * had_scale;
    }
    const float4 mv = __ldg(m + i), rv = __ldg(r + i);
    if (g_m) g_m[i] = make_float4(g.x * rv.x, g.y * rv.y, g.z * rv.z, g.w * rv.w);
    if (g_r) g_r[i] = make_float4(g.x * mv.x, g.y * mv.y, g.z * mv.z, g.w * mv.w);
  }
}

// ------------------------------------------------------------------------------------------ H12
__global__ void molecule_sum_fwd_kernel(const float* __restrict__ E_atom, const int32_t* __restrict__ rowptr,
                                        int64_t n_mol, int nT, int mean, float* __restrict__ E_mol) {
  pdl_enter();
  const int64_t total = n_mol * nT;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t mol = i / nT;
    const int t = (int)(i - mol * nT);
    const int32_t b = rowptr[mol], e = rowptr[mol + 1];
    float acc = 0.f;
    for (int32_t a = b; a < e; ++a) acc += E_atom[(int64_t)a * nT + t];
    if (mean) acc /= (float)max(e - b, 1);
    E_mol[i] = acc;
  }
}

__global__ void molecule_sum_bwd_kernel(const float* __restrict__ g_E_mol, const int32_t* __restrict__ batch_seg,
                                        const int32_t* __restrict__ rowptr, int64_t n_atoms, int nT, int mean,
                                        float* __restrict__ g_E_atom) {
  pdl_enter();
  const int64_t total = n_atoms * nT;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t a = i / nT;
    const int t = (int)(i - a * nT);
    const int32_t mol = batch_seg[a];
    float g = g_E_mol[(int64_t)mol * nT + t];
    if (mean) g /= (float)max(rowptr[mol + 1] - rowptr[mol], 1);
    g_E_atom[i] = g;
  }
}

// F_atom[a,t,k] = sum_{e in in(a)} F_ca[e,t] * V[e,k]
__global__ void force_readout_fwd_kernel(const float* __restrict__ F_ca, const float* __restrict__ V,
                                         const int32_t* __restrict__ perm, const int32_t* __restrict__ rowptr,
                                         int64_t n_atoms, int nT, float* __restrict__ F_atom) {
  pdl_enter();
  const int64_t total = n_atoms * nT;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t a = i / nT;
    const int t = (int)(i - a * nT);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    for (int32_t k = rowptr[a]; k < rowptr[a + 1]; ++k) {
      const int64_t e = perm[k];
      const float f = F_ca[e * nT + t];
      fx += f * V[3 * e];
      fy += f * V[3 * e + 1];
      fz += f * V[3 * e + 2];
    }
    F_atom[3 * i] = fx;
    F_atom[3 * i + 1] = fy;
    F_atom[3 * i + 2] = fz;
  }
}

__global__ void force_readout_bwd_kernel(const float* __restrict__ F_ca, const float* __restrict__ V,
                                         const int32_t* __restrict__ id_a, int64_t n_edges, int nT,
                                         const float* __restrict__ g_F_atom, float* __restrict__ g_F_ca,
                                         float* __restrict__ g_V) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int64_t a = id_a[e];
    const float vx = V[3 * e], vy = V[3 * e + 1], vz = V[3 * e + 2];
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int t = 0; t < nT; ++t) {
      const float* g = g_F_atom + (a * nT + t) * 3;
      if (g_F_ca) g_F_ca[e * nT + t] = g[0] * vx + g[1] * vy + g[2] * vz;
      const float f = F_ca[e * nT + t];
      gx += g[0] * f;
      gy += g[1] * f;
      gz += g[2] * f;
    }
    if (g_V) {
      g_V[3 * e] = gx;
      g_V[3 * e + 1] = gy;
      g_V[3 * e + 2] = gz;
    }
  }
}

}  // namespace gemnet

using namespace gemnet;

GEMNET_API int gemnet_edge_to_atom_fwd(const float* m, const float* r, const int32_t* perm, const int32_t* rowptr,
                                       int64_t n_atoms, int64_t n_edges, int32_t F, float scale, float* out,
                                       float had_scale, float* had, void* stream) {
  GEMNET_REQUIRE(n_atoms >= 0 && n_edges >= 0 && F > 0 && F % 4 == 0, "bad size (F must be a multiple of 4)");
  GEMNET_REQUIRE(((uintptr_t)m | (uintptr_t)r | (uintptr_t)out | (uintptr_t)had) % 16 == 0, "pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (n_atoms > 0) {
    GEMNET_REQUIRE(rowptr && out && (n_edges == 0 || (m && r && perm)), "null pointer");
    const int64_t threads = n_atoms * 32;
    launch_pdl(edge_to_atom_fwd_kernel, flat_grid(threads, 256), 256, 0, st, 
        (const float4*)m, (const float4*)r, perm, rowptr, n_atoms, F / 4, scale, (float4*)out);
  }
  if (had && n_edges > 0)
    launch_pdl(hadamard_kernel, flat_grid(n_edges * (F / 4), 256), 256, 0, st, (const float4*)m, (const float4*)r,
                                                                       n_edges * (F / 4), had_scale, (float4*)had);
  return check_launch("gemnet_edge_to_atom_fwd");
}

GEMNET_API int gemnet_edge_to_atom_bwd(const float* m, const float* r, const int32_t* id_a, int64_t n_edges,
                                       int32_t F, float scale, const float* g_out, float had_scale,
                                       const float* g_had, float* g_m, float* g_r, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && F > 0 && F % 4 == 0, "bad size (F must be a multiple of 4)");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(m && r && id_a, "null pointer");
  GEMNET_REQUIRE(((uintptr_t)m | (uintptr_t)r | (uintptr_t)g_out | (uintptr_t)g_had | (uintptr_t)g_m | (uintptr_t)g_r) % 16 == 0,
                 "pointers must be 16-byte aligned");
  launch_pdl(edge_to_atom_bwd_kernel, flat_grid(n_edges * (F / 4), 256), 256, 0, as_stream(stream), 
      (const float4*)m, (const float4*)r, id_a, n_edges, F / 4, scale, (const float4*)g_out, had_scale,
      (const float4*)g_had, (float4*)g_m, (float4*)g_r);
  return check_launch("gemnet_edge_to_atom_bwd");
}

GEMNET_API int gemnet_molecule_sum_fwd(const float* E_atom, const int32_t* mol_rowptr, int64_t n_mol,
                                       int32_t n_targets, int32_t mean, float* E_mol, void* stream) {
  GEMNET_REQUIRE(n_mol >= 0 && n_targets >= 1, "bad size");
  if (n_mol == 0) return 0;
  GEMNET_REQUIRE(E_atom && mol_rowptr && E_mol, "null pointer");
  launch_pdl(molecule_sum_fwd_kernel, flat_grid(n_mol * n_targets, 128), 128, 0, as_stream(stream), E_atom, mol_rowptr, n_mol,
                                                                                           n_targets, mean, E_mol);
  return check_launch("gemnet_molecule_sum_fwd");
}

GEMNET_API int gemnet_molecule_sum_bwd(const float* g_E_mol, const int32_t* batch_seg, const int32_t* mol_rowptr,
                                       int64_t n_atoms, int32_t n_targets, int32_t mean, float* g_E_atom,
                                       void* stream) {
  GEMNET_REQUIRE(n_atoms >= 0 && n_targets >= 1, "bad size");
  if (n_atoms == 0) return 0;
  GEMNET_REQUIRE(g_E_mol && batch_seg && mol_rowptr && g_E_atom, "null pointer");
  launch_pdl(molecule_sum_bwd_kernel, flat_grid(n_atoms * n_targets, 256), 256, 0, as_stream(stream), 
      g_E_mol, batch_seg, mol_rowptr, n_atoms, n_targets, mean, g_E_atom);
  return check_launch("gemnet_molecule_sum_bwd");
}

GEMNET_API int gemnet_force_readout_fwd(const float* F_ca, const float* V, const int32_t* perm, const int32_t* rowptr,
                                        int64_t n_atoms, int32_t n_targets, float* F_atom, void* stream) {
  GEMNET_REQUIRE(n_atoms >= 0 && n_targets >= 1, "bad size");
  if (n_atoms == 0) return 0;
  GEMNET_REQUIRE(rowptr && F_atom, "null pointer");
  launch_pdl(force_readout_fwd_kernel, flat_grid(n_atoms * n_targets, 128), 128, 0, as_stream(stream), F_ca, V, perm, rowptr,
                                                                                              n_atoms, n_targets,
                                                                                              F_atom);
  return check_launch("gemnet_force_readout_fwd");
}

GEMNET_API int gemnet_force_readout_bwd(const float* F_ca, const float* V, const int32_t* id_a, int64_t n_edges,
                                        int32_t n_targets, const float* g_F_atom, float* g_F_ca, float* g_V,
                                        void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && n_targets >= 1, "bad size");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(F_ca && V && id_a && g_F_atom, "null pointer");
  launch_pdl(force_readout_bwd_kernel, flat_grid(n_edges, 256), 256, 0, as_stream(stream), F_ca, V, id_a, n_edges, n_targets,
                                                                                  g_F_atom, g_F_ca, g_V);
  return check_launch("gemnet_force_readout_bwd");
}
