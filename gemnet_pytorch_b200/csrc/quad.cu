// H2+H5: per-quadruplet dihedral geometry fused with the real spherical harmonics Y_lm(theta, phi) of
// TensorBasisLayer (reference basis_layers.py:269-270, formulas basis_utils.py:139-171,217-243), forward and
// backward.  theta = Phi_cab comes in as (cos, sin) from the c->a-b intermediate triplet; phi = Theta_cabd is the
// clamped angle between the two plane-rejected vectors (gemnet.py:397-416).  No atan2 / sincos: cos(m phi) and
// sin(m phi) come from the Chebyshev recurrences, P_l^m from the reference's own recurrences.
#include "common.cuh"
#include "vec3.cuh"

namespace gemnet {

constexpr int kMaxL = kMaxSpherical;
constexpr int kTri = kMaxL * (kMaxL + 1) / 2;
__device__ __forceinline__ constexpr int tri(int l, int m) { return l * (l + 1) / 2 + m; }

// associated Legendre P_l^m(z) with (1 - z^2)^(1/2) := st, optionally with forward-mode derivatives
template <bool GRAD>
__device__ __forceinline__ void assoc_legendre(float z, float st, int S, float* P, float* dPz, float* dPs) {
  P[0] = 1.f;
  if (GRAD) {
    dPz[0] = 0.f;
    dPs[0] = 0.f;
  }
#pragma unroll
  for (int l = 1; l < kMaxL; ++l) {
    if (l < S) {  // P_l^l = (1 - 2l) st P_{l-1}^{l-1}                                  (basis_utils.py:141-144)
      const float f = (float)(1 - 2 * l);
      P[tri(l, l)] = f * st * P[tri(l - 1, l - 1)];
      if (GRAD) {
        dPz[tri(l, l)] = f * st * dPz[tri(l - 1, l - 1)];
        dPs[tri(l, l)] = f * (P[tri(l - 1, l - 1)] + st * dPs[tri(l - 1, l - 1)]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < kMaxL - 1; ++m) {
    if (m + 1 < S) {  // P_{m+1}^m = (2m+1) z P_m^m                                      (:146-149)
      const float f = (float)(2 * m + 1);
      P[tri(m + 1, m)] = f * z * P[tri(m, m)];
      if (GRAD) {
        dPz[tri(m + 1, m)] = f * (P[tri(m, m)] + z * dPz[tri(m, m)]);
        dPs[tri(m + 1, m)] = f * z * dPs[tri(m, m)];
      }
    }
  }
#pragma unroll
  for (int l = 2; l < kMaxL; ++l) {
#pragma unroll
    for (int m = 0; m < kMaxL - 2; ++m) {
      if (l < S && m < l - 1) {  // P_l^m = ((2l-1) z P_{l-1}^m - (l+m-1) P_{l-2}^m) / (l-m)   (:151-159)
        const float a = (float)(2 * l - 1) / (float)(l - m), b = (float)(l + m - 1) / (float)(l - m);
        P[tri(l, m)] = a * z * P[tri(l - 1, m)] - b * P[tri(l - 2, m)];
        if (GRAD) {
          dPz[tri(l, m)] = a * (P[tri(l - 1, m)] + z * dPz[tri(l - 1, m)]) - b * dPz[tri(l - 2, m)];
          dPs[tri(l, m)] = a * z * dPs[tri(l - 1, m)] - b * dPs[tri(l - 2, m)];
        }
      }
    }
  }
}

constexpr int kYlmThreads = 128;

__global__ void quad_ylm_fwd_kernel(const float* __restrict__ cossin, const float* __restrict__ ac_proj,
                                    const float* __restrict__ bd_proj, const int32_t* __restrict__ red_cab,
                                    const int32_t* __restrict__ exp_abd, int64_t n, int S,
                                    const float* __restrict__ pref, float* __restrict__ ylm,
                                    float* __restrict__ theta_phi) {
  // A thread owns one quadruplet = one row of S^2 values; the rows of the block's kYlmThreads consecutive quadruplets
  // are contiguous in memory, so they are assembled in shared memory (row stride S^2 is odd: conflict-free) and
  // written out as one coalesced block instead of 32 scattered 196-byte rows per store instruction.
  pdl_enter();
  __shared__ float rows[kYlmThreads * kMaxL * kMaxL];
  const int S2 = S * S;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x; q0 < n; q0 += stride) {
    const int64_t q = q0 + threadIdx.x;
    const bool valid = q < n;
    const int32_t i1 = valid ? red_cab[q] : 0, i2 = valid ? exp_abd[q] : 0;
    const float z = __ldg(cossin + 2 * i1), st = __ldg(cossin + 2 * i1 + 1);
    const f3 p1 = ld3(ac_proj, i1), p2 = ld3(bd_proj, i2);
    const AnglePair a = angle_pair(p1, p2);
    const float cp = a.x / a.r, sp = a.y / a.r;
    float P[kTri];
    assoc_legendre<false>(z, st, S, P, nullptr, nullptr);
    float cm[kMaxL], sm[kMaxL];
    cm[0] = 1.f;
    sm[0] = 0.f;
    cm[1] = cp;
    sm[1] = sp;
#pragma unroll
    for (int m = 2; m < kMaxL; ++m) {
      cm[m] = 2.f * cp * cm[m - 1] - cm[m - 2];
      sm[m] = 2.f * cp * sm[m - 1] - sm[m - 2];
    }
    float* o = rows + threadIdx.x * S2;
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      if (l < S) {
        const int base = l * l;
        o[base] = __ldg(pref + base) * P[tri(l, 0)];
#pragma unroll
        for (int m = 1; m < kMaxL; ++m) {
          if (m <= l) {
            o[base + m] = __ldg(pref + base + m) * P[tri(l, m)] * cm[m];
            o[base + 2 * l + 1 - m] = __ldg(pref + base + 2 * l + 1 - m) * P[tri(l, m)] * sm[m];
          }
        }
      }
    }
    if (theta_phi && valid) {
      theta_phi[2 * q] = atan2f(st, z);
      theta_phi[2 * q + 1] = atan2f(a.y, a.x);
    }
    __syncthreads();
    const int64_t cnt = min((int64_t)blockDim.x, n - q0) * S2;
    float* dst = ylm + q0 * S2;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = rows[i];
    __syncthreads();
  }
}

__global__ void quad_ylm_bwd_kernel(const float* __restrict__ cossin, const float* __restrict__ ac_proj,
                                    const float* __restrict__ bd_proj, const int32_t* __restrict__ red_cab,
                                    const int32_t* __restrict__ exp_abd, int64_t n, int S,
                                    const float* __restrict__ pref, const float* __restrict__ g_ylm,
                                    float* __restrict__ g_cossin, float* __restrict__ g_ac, float* __restrict__ g_bd) {
  pdl_enter();
  __shared__ float rows[kYlmThreads * kMaxL * kMaxL];  // the block's gradient rows, fetched as one coalesced block
  const int S2 = S * S;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x; q0 < n; q0 += stride) {
    __syncthreads();
    {
      const int64_t cnt = min((int64_t)blockDim.x, n - q0) * S2;
      const float* src = g_ylm + q0 * S2;
      for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) rows[i] = __ldg(src + i);
    }
    __syncthreads();
    const int64_t q = q0 + threadIdx.x;
    if (q >= n) continue;
    const int32_t i1 = red_cab[q], i2 = exp_abd[q];
    const float z = __ldg(cossin + 2 * i1), st = __ldg(cossin + 2 * i1 + 1);
    const f3 p1 = ld3(ac_proj, i1), p2 = ld3(bd_proj, i2);
    const AnglePair a = angle_pair(p1, p2);
    const float cp = a.x / a.r, sp = a.y / a.r;
    float P[kTri], dPz[kTri], dPs[kTri];
    assoc_legendre<true>(z, st, S, P, dPz, dPs);
    // cos(m phi), sin(m phi) and derivatives w.r.t. (cp, sp)
    float cm[kMaxL], sm[kMaxL], dcm[kMaxL], dsm_c[kMaxL], dsm_s[kMaxL];
    cm[0] = 1.f; sm[0] = 0.f; dcm[0] = 0.f; dsm_c[0] = 0.f; dsm_s[0] = 0.f;
    cm[1] = cp;  sm[1] = sp;  dcm[1] = 1.f; dsm_c[1] = 0.f; dsm_s[1] = 1.f;
#pragma unroll
    for (int m = 2; m < kMaxL; ++m) {
      cm[m] = 2.f * cp * cm[m - 1] - cm[m - 2];
      sm[m] = 2.f * cp * sm[m - 1] - sm[m - 2];
      dcm[m] = 2.f * cm[m - 1] + 2.f * cp * dcm[m - 1] - dcm[m - 2];
      dsm_c[m] = 2.f * sm[m - 1] + 2.f * cp * dsm_c[m - 1] - dsm_c[m - 2];
      dsm_s[m] = 2.f * cp * dsm_s[m - 1] - dsm_s[m - 2];
    }
    float gz = 0.f, gst = 0.f, gcp = 0.f, gsp = 0.f;
    const float* g = rows + threadIdx.x * S2;
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      if (l < S) {
        const int base = l * l;
        {
          const float w = g[base] * __ldg(pref + base);
          gz += w * dPz[tri(l, 0)];
          gst += w * dPs[tri(l, 0)];
        }
#pragma unroll
        for (int m = 1; m < kMaxL; ++m) {
          if (m <= l) {
            const float wc = g[base + m] * __ldg(pref + base + m);
            const float ws = g[base + 2 * l + 1 - m] * __ldg(pref + base + 2 * l + 1 - m);
            const float t = wc * cm[m] + ws * sm[m];
            gz += t * dPz[tri(l, m)];
            gst += t * dPs[tri(l, m)];
            gcp += P[tri(l, m)] * (wc * dcm[m] + ws * dsm_c[m]);
            gsp += P[tri(l, m)] * ws * dsm_s[m];
          }
        }
      }
    }
    f3 g1, g2;
    angle_pair_bwd(a, p1, p2, gcp, gsp, g1, g2);
    atomicAdd(g_cossin + 2 * i1, gz);
    atomicAdd(g_cossin + 2 * i1 + 1, gst);
    atomic_add3(g_ac, i1, g1);
    atomic_add3(g_bd, i2, g2);
  }
}

}  // namespace gemnet

using namespace gemnet;

GEMNET_API int gemnet_quad_ylm_fwd(const float* cossin_cab, const float* ac_proj, const float* bd_proj,
                                   const int32_t* id4_reduce_cab, const int32_t* id4_expand_abd, int64_t n_quads,
                                   int32_t num_spherical, const float* pref_lm, float* ylm, float* theta_phi,
                                   void* stream) {
  GEMNET_REQUIRE(n_quads >= 0, "negative size");
  GEMNET_REQUIRE(num_spherical >= 2 && num_spherical <= kMaxSpherical, "num_spherical must be in [2, 8]");
  if (n_quads == 0) return 0;
  GEMNET_REQUIRE(cossin_cab && ac_proj && bd_proj && id4_reduce_cab && id4_expand_abd && pref_lm && ylm, "null pointer");
  launch_pdl(quad_ylm_fwd_kernel, flat_grid(n_quads, kYlmThreads), kYlmThreads, 0, as_stream(stream),
      cossin_cab, ac_proj, bd_proj, id4_reduce_cab, id4_expand_abd, n_quads, num_spherical, pref_lm, ylm, theta_phi);
  return check_launch("gemnet_quad_ylm_fwd");
}

GEMNET_API int gemnet_quad_ylm_bwd(const float* cossin_cab, const float* ac_proj, const float* bd_proj,
                                   const int32_t* id4_reduce_cab, const int32_t* id4_expand_abd, int64_t n_quads,
                                   int32_t num_spherical, const float* pref_lm, const float* g_ylm,
                                   float* g_cossin_cab, float* g_ac_proj, float* g_bd_proj, void* stream) {
  GEMNET_REQUIRE(n_quads >= 0, "negative size");
  GEMNET_REQUIRE(num_spherical >= 2 && num_spherical <= kMaxSpherical, "num_spherical must be in [2, 8]");
  if (n_quads == 0) return 0;
  GEMNET_REQUIRE(cossin_cab && ac_proj && bd_proj && id4_reduce_cab && id4_expand_abd && pref_lm && g_ylm &&
                     g_cossin_cab && g_ac_proj && g_bd_proj, "null pointer");
  launch_pdl(quad_ylm_bwd_kernel, flat_grid(n_quads, kYlmThreads), kYlmThreads, 0, as_stream(stream),
      cossin_cab, ac_proj, bd_proj, id4_reduce_cab, id4_expand_abd, n_quads, num_spherical, pref_lm, g_ylm,
      g_cossin_cab, g_ac_proj, g_bd_proj);
  return check_launch("gemnet_quad_ylm_bwd");
}
