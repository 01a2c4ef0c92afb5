// H3 (Bessel radial basis), H4/H5 radial part (normalised spherical Bessel functions) and H6 (down projection).
// All HBM-bound, E-sized.  The spherical Bessel functions are evaluated in fp64 (the reference's closed-form fp32
// expressions cancel catastrophically for small arguments / high l — SURVEY §7 hard part 2); results are fp32.
#include "common.cuh"

namespace gemnet {

// ------------------------------------------------------------------------------------------ H3
// rbf[e,n] = u(x) * sqrt(2/c) * sin(f_n x) / d,  x = d/c
__global__ void bessel_rbf_fwd_kernel(const float* __restrict__ D, const float* __restrict__ freq, int64_t n_edges,
                                      int nR, float inv_cutoff, float norm_const, int p, float* __restrict__ out) {
  const int64_t total = n_edges * nR;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t e = i / nR;
    const int n = (int)(i - e * nR);
    const float d = D[e];
    const float x = d * inv_cutoff;
    float u, du;
    envelope_val_grad<float>(x, p, u, du);
    out[i] = u * norm_const * sinf(__ldg(freq + n) * x) / d;
  }
}

// one warp per 32 edges; gfreq reduced per warp, then one atomic per (warp, n)
__global__ void bessel_rbf_bwd_kernel(const float* __restrict__ D, const float* __restrict__ freq, int64_t n_edges,
                                      int nR, float inv_cutoff, float norm_const, int p,
                                      const float* __restrict__ g, float* __restrict__ gD,
                                      float* __restrict__ gfreq) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (n_edges + 31) / 32 * 32;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_round; e += stride) {
    const bool live = e < n_edges;
    const float d = live ? D[e] : 1.f;
    const float x = d * inv_cutoff;
    float u, du;
    envelope_val_grad<float>(x, p, u, du);
    float acc = 0.f;
    for (int n = 0; n < nR; ++n) {
      const float f = __ldg(freq + n);
      float s, c;
      sincosf(f * x, &s, &c);
      const float go = live ? g[e * nR + n] : 0.f;
      // d/dd [ u(x) k sin(f x) / d ]
      const float dval = norm_const * ((du * inv_cutoff * s + u * f * inv_cutoff * c) / d - u * s / (d * d));
      acc += go * dval;
      if (gfreq) {
        float gf = warp_sum(go * norm_const * u * x * c / d);
        if ((threadIdx.x & 31) == 0 && gf != 0.f) atomicAdd(gfreq + n, gf);
      }
    }
    if (live) gD[e] = acc;
  }
}

// ------------------------------------------------------------------------------------------ spherical Bessel (fp64)
// j_0..j_{L-1}(t) and derivatives.  Upward recurrence for t >= 1 (error ~ (2l-1)!!/t^l eps, negligible in fp64
// for l <= 7); ascending power series below.
__device__ __forceinline__ void sph_bessel_all(double t, int L, double* j, double* dj) {
  if (t < 1.0) {
    double tl = 1.0;      // t^l
    double dfact = 1.0;   // (2l+1)!!
    const double t2 = t * t;
    for (int l = 0; l < L; ++l) {
      dfact *= (2 * l + 1);
      // j_l = t^l/(2l+1)!! * sum_k (-t^2/2)^k / (k! (2l+3)(2l+5)...(2l+2k+1))
      double term = 1.0, sum = 1.0;
      for (int k = 1; k < 12; ++k) {
        term *= -0.5 * t2 / (k * (2.0 * l + 2.0 * k + 1.0));
        sum += term;
      }
      j[l] = tl / dfact * sum;
      tl *= t;
    }
    // one extra order for the derivative of the highest l: j_l' = l/t j_l - j_{l+1}
    dfact *= (2 * L + 1);
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 12; ++k) {
      term *= -0.5 * t2 / (k * (2.0 * L + 2.0 * k + 1.0));
      sum += term;
    }
    const double jL = tl / dfact * sum;
    for (int l = 0; l < L; ++l) {
      const double jn = (l + 1 < L) ? j[l + 1] : jL;
      dj[l] = (t > 0.0 ? l / t * j[l] : (l == 1 ? 1.0 / 3.0 : 0.0)) - jn;
    }
    return;
  }
  double s, c;
  sincos(t, &s, &c);
  const double inv = 1.0 / t;
  j[0] = s * inv;
  if (L > 1) j[1] = (s * inv - c) * inv;
  for (int l = 1; l + 1 < L; ++l) j[l + 1] = (2 * l + 1) * inv * j[l] - j[l - 1];
  dj[0] = (c - s * inv) * inv;  // = -j_1
  for (int l = 1; l < L; ++l) dj[l] = j[l - 1] - (l + 1) * inv * j[l];
}

// rad[e,l,n] = u(x) c^-1.5 N_ln j_l(z_ln x); one thread per (e, l*nR+n)
__global__ void sph_bessel_radial_fwd_kernel(const float* __restrict__ D, int64_t n_edges, int S, int nR,
                                             const double* __restrict__ zeros, const double* __restrict__ norm,
                                             double inv_cutoff, double norm_const, int p, float* __restrict__ out) {
  const int SR = S * nR;
  const int64_t total = n_edges * SR;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t e = i / SR;
    const int ln = (int)(i - e * SR);
    const int l = ln / nR;
    const double x = (double)D[e] * inv_cutoff;
    double u, du;
    envelope_val_grad<double>(x, p, u, du);
    double j[kMaxSpherical], dj[kMaxSpherical];
    sph_bessel_all(zeros[ln] * x, l + 1, j, dj);
    out[i] = (float)(u * norm_const * norm[ln] * j[l]);
  }
}

// gD[e] = sum_{l,n} g[e,l,n] d rad / d d; one thread per edge
__global__ void sph_bessel_radial_bwd_kernel(const float* __restrict__ D, int64_t n_edges, int S, int nR,
                                             const double* __restrict__ zeros, const double* __restrict__ norm,
                                             double inv_cutoff, double norm_const, int p,
                                             const float* __restrict__ g, float* __restrict__ gD) {
  const int SR = S * nR;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const double x = (double)D[e] * inv_cutoff;
    double u, du;
    envelope_val_grad<double>(x, p, u, du);
    double acc = 0.0;
    for (int n = 0; n < nR; ++n) {
      for (int l = 0; l < S; ++l) {
        const int ln = l * nR + n;
        const double z = zeros[ln];
        double j[kMaxSpherical], dj[kMaxSpherical];
        sph_bessel_all(z * x, l + 1, j, dj);
        acc += (double)g[e * SR + ln] * norm_const * norm[ln] * (du * j[l] + u * z * dj[l]) * inv_cutoff;
      }
    }
    gD[e] = (float)acc;
  }
}

// ------------------------------------------------------------------------------------------ H6
// out[e,i,s] = sum_n rad[e, l_of_s[s], n] * Wd[s,n,i]; one thread per (e,i,s), s fastest (coalesced store)
__global__ void downproj_fwd_kernel(const float* __restrict__ rad, const float* __restrict__ Wd,
                                    const int32_t* __restrict__ l_of_s, int64_t n_edges, int L, int nR, int S, int I,
                                    float* __restrict__ out) {
  const int IS = I * S;
  const int64_t total = n_edges * IS;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t e = idx / IS;
    const int is = (int)(idx - e * IS);
    const int i = is / S, s = is - i * S;
    const float* r = rad + (e * L + __ldg(l_of_s + s)) * nR;
    const float* w = Wd + (int64_t)s * nR * I + i;
    float acc = 0.f;
    for (int n = 0; n < nR; ++n) acc += r[n] * __ldg(w + n * I);
    out[idx] = acc;
  }
}

// g_rad[e,l,n] = sum_{s: l_of_s[s]==l} sum_i g[e,i,s] Wd[s,n,i]; one thread per (e,l,n)
__global__ void downproj_bwd_rad_kernel(const float* __restrict__ Wd, const int32_t* __restrict__ l_of_s,
                                        int64_t n_edges, int L, int nR, int S, int I, const float* __restrict__ g,
                                        float* __restrict__ g_rad) {
  const int LR = L * nR;
  const int64_t total = n_edges * LR;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t e = idx / LR;
    const int ln = (int)(idx - e * LR);
    const int l = ln / nR, n = ln - l * nR;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) {
      if (__ldg(l_of_s + s) != l) continue;
      const float* w = Wd + ((int64_t)s * nR + n) * I;
      const float* ge = g + e * I * S + s;
      for (int i = 0; i < I; ++i) acc += ge[i * S] * __ldg(w + i);
    }
    g_rad[idx] = acc;
  }
}

// g_Wd[s,n,i] += sum_e g[e,i,s] rad[e,l(s),n]; CTA owns a slab of edges, thread owns (s,n,i) entries
__global__ void downproj_bwd_weight_kernel(const float* __restrict__ rad, const int32_t* __restrict__ l_of_s,
                                           int64_t n_edges, int L, int nR, int S, int I,
                                           const float* __restrict__ g, float* __restrict__ g_Wd,
                                           int64_t edges_per_cta) {
  const int64_t e0 = (int64_t)blockIdx.x * edges_per_cta;
  const int64_t e1 = min(e0 + edges_per_cta, n_edges);
  const int total = S * nR * I;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    const int s = w / (nR * I);
    const int rem = w - s * nR * I;
    const int n = rem / I, i = rem - n * I;
    const int l = __ldg(l_of_s + s);
    float acc = 0.f;
    for (int64_t e = e0; e < e1; ++e) acc += g[(e * I + i) * S + s] * rad[(e * L + l) * nR + n];
    atomicAdd(g_Wd + w, acc);
  }
}

}  // namespace gemnet

using namespace gemnet;

GEMNET_API int gemnet_bessel_rbf_fwd(const float* D, const float* freq, int64_t n_edges, int32_t num_radial,
                                     float cutoff, int32_t envelope_p, float* rbf, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && num_radial >= 1 && num_radial <= 64, "bad size (num_radial must be <= 64)");
  GEMNET_REQUIRE(envelope_p > 0 && cutoff > 0.f, "bad envelope/cutoff");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(D && freq && rbf, "null pointer");
  const float inv_c = 1.0f / cutoff;
  bessel_rbf_fwd_kernel<<<flat_grid(n_edges * num_radial, 256), 256, 0, as_stream(stream)>>>(
      D, freq, n_edges, num_radial, inv_c, sqrtf(2.0f * inv_c), envelope_p, rbf);
  return check_launch("gemnet_bessel_rbf_fwd");
}

GEMNET_API int gemnet_bessel_rbf_bwd(const float* D, const float* freq, int64_t n_edges, int32_t num_radial,
                                     float cutoff, int32_t envelope_p, const float* g_rbf, float* gD, float* gfreq,
                                     void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && num_radial >= 1 && num_radial <= 64, "bad size (num_radial must be <= 64)");
  GEMNET_REQUIRE(envelope_p > 0 && cutoff > 0.f, "bad envelope/cutoff");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(D && freq && g_rbf && gD, "null pointer");
  const float inv_c = 1.0f / cutoff;
  bessel_rbf_bwd_kernel<<<flat_grid(n_edges, 256), 256, 0, as_stream(stream)>>>(
      D, freq, n_edges, num_radial, inv_c, sqrtf(2.0f * inv_c), envelope_p, g_rbf, gD, gfreq);
  return check_launch("gemnet_bessel_rbf_bwd");
}

GEMNET_API int gemnet_sph_bessel_radial_fwd(const float* D, int64_t n_edges, int32_t num_spherical,
                                            int32_t num_radial, const double* zeros, const double* norm,
                                            float cutoff, int32_t envelope_p, float* rad, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && num_radial >= 1 && num_radial <= 64, "bad size (num_radial must be <= 64)");
  GEMNET_REQUIRE(num_spherical >= 1 && num_spherical <= kMaxSpherical, "num_spherical must be in [1, 8]");
  GEMNET_REQUIRE(envelope_p > 0 && cutoff > 0.f, "bad envelope/cutoff");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(D && zeros && norm && rad, "null pointer");
  const double inv_c = 1.0 / (double)cutoff;
  sph_bessel_radial_fwd_kernel<<<flat_grid(n_edges * num_spherical * num_radial, 256), 256, 0, as_stream(stream)>>>(
      D, n_edges, num_spherical, num_radial, zeros, norm, inv_c, inv_c * sqrt(inv_c), envelope_p, rad);
  return check_launch("gemnet_sph_bessel_radial_fwd");
}

GEMNET_API int gemnet_sph_bessel_radial_bwd(const float* D, int64_t n_edges, int32_t num_spherical,
                                            int32_t num_radial, const double* zeros, const double* norm,
                                            float cutoff, int32_t envelope_p, const float* g_rad, float* gD,
                                            void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && num_radial >= 1 && num_radial <= 64, "bad size (num_radial must be <= 64)");
  GEMNET_REQUIRE(num_spherical >= 1 && num_spherical <= kMaxSpherical, "num_spherical must be in [1, 8]");
  GEMNET_REQUIRE(envelope_p > 0 && cutoff > 0.f, "bad envelope/cutoff");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(D && zeros && norm && g_rad && gD, "null pointer");
  const double inv_c = 1.0 / (double)cutoff;
  sph_bessel_radial_bwd_kernel<<<flat_grid(n_edges, 128), 128, 0, as_stream(stream)>>>(
      D, n_edges, num_spherical, num_radial, zeros, norm, inv_c, inv_c * sqrt(inv_c), envelope_p, g_rad, gD);
  return check_launch("gemnet_sph_bessel_radial_bwd");
}

GEMNET_API int gemnet_downproj_fwd(const float* rad, const float* Wd, const int32_t* l_of_s, int64_t n_edges,
                                   int32_t L, int32_t num_radial, int32_t S, int32_t I, float* rbf_W1,
                                   void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && L >= 1 && num_radial >= 1 && S >= 1 && I >= 1, "bad size");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(rad && Wd && l_of_s && rbf_W1, "null pointer");
  downproj_fwd_kernel<<<flat_grid(n_edges * I * S, 256), 256, 0, as_stream(stream)>>>(rad, Wd, l_of_s, n_edges, L,
                                                                                        num_radial, S, I, rbf_W1);
  return check_launch("gemnet_downproj_fwd");
}

GEMNET_API int gemnet_downproj_bwd(const float* rad, const float* Wd, const int32_t* l_of_s, int64_t n_edges,
                                   int32_t L, int32_t num_radial, int32_t S, int32_t I, const float* g_rbf_W1,
                                   float* g_rad, float* g_Wd, void* stream) {
  GEMNET_REQUIRE(n_edges >= 0 && L >= 1 && num_radial >= 1 && S >= 1 && I >= 1, "bad size");
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(rad && Wd && l_of_s && g_rbf_W1, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (g_rad)
    downproj_bwd_rad_kernel<<<flat_grid(n_edges * L * num_radial, 256), 256, 0, st>>>(Wd, l_of_s, n_edges, L,
                                                                                      num_radial, S, I, g_rbf_W1,
                                                                                      g_rad);
  if (g_Wd) {
    const int64_t ctas = n_edges < 148 * 4 ? (n_edges + 15) / 16 : 148 * 4;
    const int64_t per = ceil_div(n_edges, ctas);
    downproj_bwd_weight_kernel<<<(unsigned)ceil_div(n_edges, per), 256, 0, st>>>(rad, l_of_s, n_edges, L, num_radial,
                                                                                  S, I, g_rbf_W1, g_Wd, per);
  }
  return check_launch("gemnet_downproj_bwd");
}
