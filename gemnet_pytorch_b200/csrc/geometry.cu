// H1 (edge vectors) and H2+H4-angular (triplet angle -> Y_l0) kernels, forward and backward.
// HBM-bound gather/pointwise work: one thread per edge / triplet, R (A x 12 B) stays L1/L2 resident.
#include "common.cuh"
#include "vec3.cuh"

namespace gemnet {

// ------------------------------------------------------------------------------------------ H1
__global__ void edge_geometry_fwd_kernel(const float* __restrict__ R, const int32_t* __restrict__ id_c,
                                         const int32_t* __restrict__ id_a, int64_t n, float* __restrict__ D,
                                         float* __restrict__ V) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    f3 x = sub(ld3(R, id_a[e]), ld3(R, id_c[e]));
    float d = sqrtf(dot(x, x));
    D[e] = d;
    V[3 * e] = x.x / d;
    V[3 * e + 1] = x.y / d;
    V[3 * e + 2] = x.z / d;
  }
}

__global__ void edge_geometry_bwd_kernel(const float* __restrict__ D, const float* __restrict__ V,
                                         const int32_t* __restrict__ id_c, const int32_t* __restrict__ id_a,
                                         int64_t n, const float* __restrict__ gD, const float* __restrict__ gV,
                                         float* __restrict__ gR) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    f3 v = ld3(V, e);
    f3 g = {0.f, 0.f, 0.f};
    if (gD) g = mul(v, gD[e]);
    if (gV) {
      f3 gv = ld3(gV, e);
      float inv = 1.0f / D[e];
      g = add(g, mul(sub(gv, mul(v, dot(gv, v))), inv));
    }
    atomic_add3(gR, id_a[e], g);
    atomic_add3(gR, id_c[e], mul(g, -1.0f));
  }
}

// Legendre P_l(z) and P_l'(z), l < S
template <int MAXS>
__device__ __forceinline__ void legendre(float z, int S, float* P, float* dP) {
  P[0] = 1.f;
  dP[0] = 0.f;
  if (S > 1) {
    P[1] = z;
    dP[1] = 1.f;
  }
  for (int l = 2; l < S; ++l) {
    P[l] = ((2 * l - 1) * z * P[l - 1] - (l - 1) * P[l - 2]) / l;
    dP[l] = dP[l - 2] + (2 * l - 1) * P[l - 1];
  }
}

// ------------------------------------------------------------------------------------------ H2 + H4 angular
__global__ void triplet_sph_fwd_kernel(const float* __restrict__ R, const int32_t* __restrict__ id_c,
                                       const int32_t* __restrict__ id_a, const int32_t* __restrict__ red,
                                       const int32_t* __restrict__ exp, int64_t n, int S,
                                       const float* __restrict__ pref, float* __restrict__ sph,
                                       float* __restrict__ angle) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int32_t er = red[t], ee = exp[t];
    f3 ra = ld3(R, id_a[er]);
    f3 u = sub(ld3(R, id_c[er]), ra);
    f3 v = sub(ld3(R, id_c[ee]), ra);
    AnglePair a = angle_pair(u, v);
    float P[kMaxSpherical], dP[kMaxSpherical];
    legendre<kMaxSpherical>(a.x / a.r, S, P, dP);
    for (int l = 0; l < S; ++l) sph[t * S + l] = __ldg(pref + l) * P[l];
    if (angle) angle[t] = atan2f(a.y, a.x);
  }
}

__global__ void triplet_sph_bwd_kernel(const float* __restrict__ R, const int32_t* __restrict__ id_c,
                                       const int32_t* __restrict__ id_a, const int32_t* __restrict__ red,
                                       const int32_t* __restrict__ exp, int64_t n, int S,
                                       const float* __restrict__ pref, const float* __restrict__ g_sph,
                                       float* __restrict__ gR) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int32_t er = red[t], ee = exp[t];
    const int32_t ia = id_a[er], ic = id_c[er], ib = id_c[ee];
    f3 ra = ld3(R, ia);
    f3 u = sub(ld3(R, ic), ra);
    f3 v = sub(ld3(R, ib), ra);
    AnglePair a = angle_pair(u, v);
    float P[kMaxSpherical], dP[kMaxSpherical];
    legendre<kMaxSpherical>(a.x / a.r, S, P, dP);
    float gz = 0.f;
    for (int l = 0; l < S; ++l) gz += g_sph[t * S + l] * __ldg(pref + l) * dP[l];
    f3 gu, gv;
    angle_pair_bwd(a, u, v, gz, 0.f, gu, gv);
    atomic_add3(gR, ic, gu);
    atomic_add3(gR, ib, gv);
    atomic_add3(gR, ia, mul(add(gu, gv), -1.0f));
  }
}

// ------------------------------------------------------------------------------------------ quadruplet intermediates
__global__ void quad_intm_fwd_kernel(const float* __restrict__ R, const int32_t* __restrict__ ce,
                                     const int32_t* __restrict__ fi, const int32_t* __restrict__ se, int64_t n, int S,
                                     const float* __restrict__ pref0, float* __restrict__ sph,
                                     float* __restrict__ cossin, float* __restrict__ proj) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    f3 rc = ld3(R, ce[j]);
    f3 u = sub(ld3(R, fi[j]), rc);
    f3 v = sub(ld3(R, se[j]), rc);
    AnglePair a = angle_pair(u, v);
    const float cz = a.x / a.r;
    if (sph) {
      float P[kMaxSpherical], dP[kMaxSpherical];
      legendre<kMaxSpherical>(cz, S, P, dP);
      for (int l = 0; l < S; ++l) sph[j * S + l] = __ldg(pref0 + l) * P[l];
    }
    if (cossin) {
      cossin[2 * j] = cz;
      cossin[2 * j + 1] = a.y / a.r;
    }
    if (proj) {  // vector_rejection(v, u), gemnet.py:314-332
      const float alpha = dot(v, u) / dot(u, u);
      proj[3 * j] = v.x - alpha * u.x;
      proj[3 * j + 1] = v.y - alpha * u.y;
      proj[3 * j + 2] = v.z - alpha * u.z;
    }
  }
}

__global__ void quad_intm_bwd_kernel(const float* __restrict__ R, const int32_t* __restrict__ ce,
                                     const int32_t* __restrict__ fi, const int32_t* __restrict__ se, int64_t n, int S,
                                     const float* __restrict__ pref0, const float* __restrict__ g_sph,
                                     const float* __restrict__ g_cossin, const float* __restrict__ g_proj,
                                     float* __restrict__ gR) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const int32_t ic = ce[j], i1 = fi[j], i2 = se[j];
    f3 rc = ld3(R, ic);
    f3 u = sub(ld3(R, i1), rc);
    f3 v = sub(ld3(R, i2), rc);
    AnglePair a = angle_pair(u, v);
    float g_cos = 0.f, g_sin = 0.f;
    if (g_sph) {
      float P[kMaxSpherical], dP[kMaxSpherical];
      legendre<kMaxSpherical>(a.x / a.r, S, P, dP);
      for (int l = 0; l < S; ++l) g_cos += g_sph[j * S + l] * __ldg(pref0 + l) * dP[l];
    }
    if (g_cossin) {
      g_cos += g_cossin[2 * j];
      g_sin += g_cossin[2 * j + 1];
    }
    f3 gu, gv;
    angle_pair_bwd(a, u, v, g_cos, g_sin, gu, gv);
    if (g_proj) {
      f3 gp = ld3(g_proj, j);
      const float q = dot(u, u), d = dot(v, u), gpu = dot(gp, u);
      gv = add(gv, sub(gp, mul(u, gpu / q)));
      gu = add(gu, add(mul(gp, -d / q), add(mul(v, -gpu / q), mul(u, 2.f * gpu * d / (q * q)))));
    }
    atomic_add3(gR, i1, gu);
    atomic_add3(gR, i2, gv);
    atomic_add3(gR, ic, mul(add(gu, gv), -1.0f));
  }
}

}  // namespace gemnet

using namespace gemnet;

GEMNET_API int gemnet_edge_geometry_fwd(const float* R, const int32_t* id_c, const int32_t* id_a, int64_t n_edges,
                                        float* D, float* V, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0, "negative size");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(R && id_c && id_a && D && V, "null pointer");
  edge_geometry_fwd_kernel<<<flat_grid(n_edges, 256), 256, 0, as_stream(stream)>>>(R, id_c, id_a, n_edges, D, V);
  return check_launch("gemnet_edge_geometry_fwd");
}

GEMNET_API int gemnet_edge_geometry_bwd(const float* D, const float* V, const int32_t* id_c, const int32_t* id_a,
                                        int64_t n_edges, const float* gD, const float* gV, float* gR, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0, "negative size");
  if (n_edges == 0 || (!gD && !gV)) return 0;
  GEMNET_REQUIRE(D && V && id_c && id_a && gR, "null pointer");
  edge_geometry_bwd_kernel<<<flat_grid(n_edges, 256), 256, 0, as_stream(stream)>>>(D, V, id_c, id_a, n_edges, gD, gV,
                                                                                     gR);
  return check_launch("gemnet_edge_geometry_bwd");
}

GEMNET_API int gemnet_triplet_sph_fwd(const float* R, const int32_t* id_c, const int32_t* id_a,
                                      const int32_t* id3_reduce_ca, const int32_t* id3_expand_ba, int64_t n_triplets,
                                      int32_t num_spherical, const float* pref, float* sph, float* angle,
                                      void* stream) {
  GEMNET_REQUIRE(n_triplets >= 0, "negative size");
  GEMNET_REQUIRE(num_spherical >= 1 && num_spherical <= kMaxSpherical, "num_spherical must be in [1, 8]");
  if (n_triplets == 0) return 0;
  GEMNET_REQUIRE(R && id_c && id_a && id3_reduce_ca && id3_expand_ba && pref && sph, "null pointer");
  triplet_sph_fwd_kernel<<<flat_grid(n_triplets, 256), 256, 0, as_stream(stream)>>>(
      R, id_c, id_a, id3_reduce_ca, id3_expand_ba, n_triplets, num_spherical, pref, sph, angle);
  return check_launch("gemnet_triplet_sph_fwd");
}

GEMNET_API int gemnet_triplet_sph_bwd(const float* R, const int32_t* id_c, const int32_t* id_a,
                                      const int32_t* id3_reduce_ca, const int32_t* id3_expand_ba, int64_t n_triplets,
                                      int32_t num_spherical, const float* pref, const float* g_sph, float* gR,
                                      void* stream) {
  GEMNET_REQUIRE(n_triplets >= 0, "negative size");
  GEMNET_REQUIRE(num_spherical >= 1 && num_spherical <= kMaxSpherical, "num_spherical must be in [1, 8]");
  if (n_triplets == 0) return 0;
  GEMNET_REQUIRE(R && id_c && id_a && id3_reduce_ca && id3_expand_ba && pref && g_sph && gR, "null pointer");
  triplet_sph_bwd_kernel<<<flat_grid(n_triplets, 256), 256, 0, as_stream(stream)>>>(
      R, id_c, id_a, id3_reduce_ca, id3_expand_ba, n_triplets, num_spherical, pref, g_sph, gR);
  return check_launch("gemnet_triplet_sph_bwd");
}

GEMNET_API int gemnet_quad_intm_fwd(const float* R, const int32_t* atom_center, const int32_t* atom_first,
                                    const int32_t* atom_second, int64_t n_intm, int32_t num_spherical,
                                    const float* pref0, float* sph, float* cossin, float* proj, void* stream) {
  GEMNET_REQUIRE(n_intm >= 0, "negative size");
  GEMNET_REQUIRE(num_spherical >= 1 && num_spherical <= kMaxSpherical, "num_spherical must be in [1, 8]");
  if (n_intm == 0) return 0;
  GEMNET_REQUIRE(R && atom_center && atom_first && atom_second, "null pointer");
  GEMNET_REQUIRE(!sph || pref0, "sph requested without prefactors");
  quad_intm_fwd_kernel<<<flat_grid(n_intm, 256), 256, 0, as_stream(stream)>>>(
      R, atom_center, atom_first, atom_second, n_intm, num_spherical, pref0, sph, cossin, proj);
  return check_launch("gemnet_quad_intm_fwd");
}

GEMNET_API int gemnet_quad_intm_bwd(const float* R, const int32_t* atom_center, const int32_t* atom_first,
                                    const int32_t* atom_second, int64_t n_intm, int32_t num_spherical,
                                    const float* pref0, const float* g_sph, const float* g_cossin,
                                    const float* g_proj, float* gR, void* stream) {
  GEMNET_REQUIRE(n_intm >= 0, "negative size");
  GEMNET_REQUIRE(num_spherical >= 1 && num_spherical <= kMaxSpherical, "num_spherical must be in [1, 8]");
  if (n_intm == 0) return 0;
  GEMNET_REQUIRE(R && atom_center && atom_first && atom_second && gR, "null pointer");
  GEMNET_REQUIRE(!g_sph || pref0, "g_sph given without prefactors");
  quad_intm_bwd_kernel<<<flat_grid(n_intm, 256), 256, 0, as_stream(stream)>>>(
      R, atom_center, atom_first, atom_second, n_intm, num_spherical, pref0, g_sph, g_cossin, g_proj, gR);
  return check_launch("gemnet_quad_intm_bwd");
}
