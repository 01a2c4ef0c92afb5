// Fused pointwise kernels around the dense layers (ScaledSiLU, residual tails, Hadamard*scale) and row gathers
// with CSR (atomic-free, deterministic) backward.  The reference runs each of these as 2-4 separate ATen kernels
// (base_layers.py:51-89, interaction_block.py:202-233,661-664,693-695; embedding_block.py:70-71).
// HBM-bound elementwise work: 128-bit loads/stores, grid-stride over a multiple of the SM count.
#include "common.cuh"

namespace gemnet {

constexpr float kSiluScale = 1.0f / 0.6f;

__device__ __forceinline__ float ssilu(float x) { return x / (1.0f + expf(-x)) * kSiluScale; }
__device__ __forceinline__ float dssilu(float x) {
  // sigmoid s and 1 - s from t = exp(-|x|) in (0, 1]: no overflow, no cancellation
  const float t = expf(-fabsf(x));
  const float r = 1.0f / (1.0f + t);
  const float s = x >= 0.f ? r : t * r, oms = x >= 0.f ? t * r : r;
  return s * fmaf(x, oms, 1.0f) * kSiluScale;
}

// mode 0: out = ssilu(z)
// mode 1: out = (x + ssilu(z)) * s
__global__ void silu_fwd_kernel(const float4* __restrict__ z, const float4* __restrict__ x, float s, int64_t n4,
                                int mode, float4* __restrict__ out) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(z + i);
    float4 r = make_float4(ssilu(a.x), ssilu(a.y), ssilu(a.z), ssilu(a.w));
    if (mode == 1) {
      const float4 b = __ldg(x + i);
      r = make_float4((b.x + r.x) * s, (b.y + r.y) * s, (b.z + r.z) * s, (b.w + r.w) * s);
    }
    out[i] = r;
  }
}

// gz = g * s * dssilu(z)   (s = 1 for mode 0);  the x-branch gradient of mode 1 is g * s (scale_kernel)
__global__ void silu_bwd_kernel(const float4* __restrict__ z, const float4* __restrict__ g, float s, int64_t n4,
                                float4* __restrict__ gz) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(z + i), b = __ldg(g + i);
    gz[i] = make_float4(b.x * s * dssilu(a.x), b.y * s * dssilu(a.y), b.z * s * dssilu(a.z), b.w * s * dssilu(a.w));
  }
}

// out = sa*a [+ sb*b [+ sc*c]]   (b, c optional; any n: the last n % 4 elements are done by one thread)
__global__ void lincomb3_kernel(const float* __restrict__ a, float sa, const float* __restrict__ b, float sb,
                                const float* __restrict__ c, float sc, int64_t n, float* __restrict__ out) {
  pdl_enter();
  const int64_t n4 = n / 4, stride = (int64_t)gridDim.x * blockDim.x;
  const float4 *a4 = reinterpret_cast<const float4*>(a), *b4 = reinterpret_cast<const float4*>(b),
               *c4 = reinterpret_cast<const float4*>(c);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 r = __ldg(a4 + i);
    r = make_float4(r.x * sa, r.y * sa, r.z * sa, r.w * sa);
    if (b) {
      const float4 t = __ldg(b4 + i);
      r = make_float4(fmaf(sb, t.x, r.x), fmaf(sb, t.y, r.y), fmaf(sb, t.z, r.z), fmaf(sb, t.w, r.w));
    }
    if (c) {
      const float4 t = __ldg(c4 + i);
      r = make_float4(fmaf(sc, t.x, r.x), fmaf(sc, t.y, r.y), fmaf(sc, t.z, r.z), fmaf(sc, t.w, r.w));
    }
    reinterpret_cast<float4*>(out)[i] = r;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = 4 * n4; i < n; ++i) {
      float r = a[i] * sa;
      if (b) r = fmaf(sb, b[i], r);
      if (c) r = fmaf(sc, c[i], r);
      out[i] = r;
    }
}

// out = (a [+ b [+ c]]) * s      (b, c optional)
__global__ void add_scale_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                 const float4* __restrict__ c, float s, int64_t n4, float4* __restrict__ out) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 r = __ldg(a + i);
    if (b) {
      const float4 t = __ldg(b + i);
      r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
    }
    if (c) {
      const float4 t = __ldg(c + i);
      r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
    }
    out[i] = make_float4(r.x * s, r.y * s, r.z * s, r.w * s);
  }
}

// out = a * b * s ; backward: ga = g * b * s, gb = g * a * s
__global__ void mul_scale_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float s, int64_t n4,
                                 float4* __restrict__ out) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 p = __ldg(a + i), q = __ldg(b + i);
    out[i] = make_float4(p.x * q.x * s, p.y * q.y * s, p.z * q.z * s, p.w * q.w * s);
  }
}
__global__ void mul_scale_bwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                     const float4* __restrict__ g, float s, int64_t n4, float4* __restrict__ ga,
                                     float4* __restrict__ gb) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 p = __ldg(a + i), q = __ldg(b + i), t = __ldg(g + i);
    if (ga) ga[i] = make_float4(t.x * q.x * s, t.y * q.y * s, t.z * q.z * s, t.w * q.w * s);
    if (gb) gb[i] = make_float4(t.x * p.x * s, t.y * p.y * s, t.z * p.z * s, t.w * p.w * s);
  }
}

// scalar tails (n % 4 elements) of the vector kernels above: op 0 silu mode0, 1 silu mode1, 2 silu_bwd, 3 add_scale,
// 4 mul_scale, 5 mul_scale_bwd
__global__ void tail_kernel(int op, const float* a, const float* b, const float* c, float s, int64_t begin, int64_t n,
                            float* o1, float* o2) {
  pdl_enter();
  const int64_t i = begin + threadIdx.x;
  if (i >= n) return;
  switch (op) {
    case 0: o1[i] = ssilu(a[i]); break;
    case 1: o1[i] = (b[i] + ssilu(a[i])) * s; break;
    case 2: o1[i] = b[i] * s * dssilu(a[i]); break;
    case 3: o1[i] = (a[i] + (b ? b[i] : 0.f) + (c ? c[i] : 0.f)) * s; break;
    case 4: o1[i] = a[i] * b[i] * s; break;
    case 5:
      if (o1) o1[i] = c[i] * b[i] * s;
      if (o2) o2[i] = c[i] * a[i] * s;
      break;
  }
}

// out[j,:] = x[idx[j],:]  (F % 4 == 0)
__global__ void gather_rows_kernel(const float4* __restrict__ x, const int32_t* __restrict__ idx, int64_t n, int F4,
                                   float4* __restrict__ out) {
  pdl_enter();
  const int64_t total = n * F4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t j = i / F4;
    const int c = (int)(i - j * F4);
    out[i] = __ldg(x + (int64_t)__ldg(idx + j) * F4 + c);
  }
}

// out[r,:] = sum_{k in [rowptr[r], rowptr[r+1])} src[perm[k],:]  — deterministic segment sum (gather backward)
__global__ void segment_sum_rows_kernel(const float4* __restrict__ src, const int32_t* __restrict__ perm,
                                        const int32_t* __restrict__ rowptr, int64_t n_rows, int F4,
                                        float4* __restrict__ out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const int32_t b = rowptr[r], e = rowptr[r + 1];
    for (int c = lane; c < F4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int32_t k = b; k < e; ++k) {
        const float4 v = __ldg(src + (int64_t)__ldg(perm + k) * F4 + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      out[r * F4 + c] = acc;
    }
  }
}

}  // namespace gemnet

using namespace gemnet;

static inline bool aligned16(const void* p) { return ((uintptr_t)p % 16) == 0; }

GEMNET_API int gemnet_scaled_silu_fwd(const float* z, const float* x, float s, int64_t n, int32_t mode, float* out,
                                      void* stream) {
  GEMNET_REQUIRE(n >= 0, "negative size");
  GEMNET_REQUIRE(mode == 0 || (mode == 1 && x), "mode 1 needs x");
  if (n == 0) return 0;
  GEMNET_REQUIRE(z && out && aligned16(z) && aligned16(x) && aligned16(out), "null or misaligned pointer");
  if (n / 4)
    launch_pdl(silu_fwd_kernel, flat_grid(n / 4, 256), 256, 0, as_stream(stream), (const float4*)z, (const float4*)x, s,
                                                                         n / 4, mode, (float4*)out);
  if (n % 4) launch_pdl(tail_kernel, 1, 4, 0, as_stream(stream), mode, z, x, nullptr, s, n - n % 4, n, out, nullptr);
  return check_launch("gemnet_scaled_silu_fwd");
}

GEMNET_API int gemnet_scaled_silu_bwd(const float* z, const float* g, float s, int64_t n, float* gz, void* stream) {
  GEMNET_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  GEMNET_REQUIRE(z && g && gz && aligned16(z) && aligned16(g) && aligned16(gz), "null or misaligned pointer");
  if (n / 4)
    launch_pdl(silu_bwd_kernel, flat_grid(n / 4, 256), 256, 0, as_stream(stream), (const float4*)z, (const float4*)g, s,
                                                                         n / 4, (float4*)gz);
  if (n % 4) launch_pdl(tail_kernel, 1, 4, 0, as_stream(stream), 2, z, g, nullptr, s, n - n % 4, n, gz, nullptr);
  return check_launch("gemnet_scaled_silu_bwd");
}

GEMNET_API int gemnet_lincomb3(const float* a, float sa, const float* b, float sb, const float* c, float sc, int64_t n,
                               float* out, void* stream) {
  GEMNET_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  GEMNET_REQUIRE(a && out && aligned16(a) && aligned16(b) && aligned16(c) && aligned16(out), "null or misaligned pointer");
  launch_pdl(lincomb3_kernel, flat_grid(n / 4 + 1, 256), 256, 0, as_stream(stream), a, sa, b, sb, c, sc, n, out);
  return check_launch("gemnet_lincomb3");
}

GEMNET_API int gemnet_add_scale(const float* a, const float* b, const float* c, float s, int64_t n, float* out,
                                void* stream) {
  GEMNET_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  GEMNET_REQUIRE(a && out && aligned16(a) && aligned16(b) && aligned16(c) && aligned16(out), "null or misaligned pointer");
  if (n / 4)
    launch_pdl(add_scale_kernel, flat_grid(n / 4, 256), 256, 0, as_stream(stream), (const float4*)a, (const float4*)b,
                                                                          (const float4*)c, s, n / 4, (float4*)out);
  if (n % 4) launch_pdl(tail_kernel, 1, 4, 0, as_stream(stream), 3, a, b, c, s, n - n % 4, n, out, nullptr);
  return check_launch("gemnet_add_scale");
}

GEMNET_API int gemnet_mul_scale_fwd(const float* a, const float* b, float s, int64_t n, float* out, void* stream) {
  GEMNET_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  GEMNET_REQUIRE(a && b && out && aligned16(a) && aligned16(b) && aligned16(out), "null or misaligned pointer");
  if (n / 4)
    launch_pdl(mul_scale_kernel, flat_grid(n / 4, 256), 256, 0, as_stream(stream), (const float4*)a, (const float4*)b, s,
                                                                          n / 4, (float4*)out);
  if (n % 4) launch_pdl(tail_kernel, 1, 4, 0, as_stream(stream), 4, a, b, nullptr, s, n - n % 4, n, out, nullptr);
  return check_launch("gemnet_mul_scale_fwd");
}

GEMNET_API int gemnet_mul_scale_bwd(const float* a, const float* b, const float* g, float s, int64_t n, float* ga,
                                    float* gb, void* stream) {
  GEMNET_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  GEMNET_REQUIRE(a && b && g && aligned16(a) && aligned16(b) && aligned16(g) && aligned16(ga) && aligned16(gb),
                 "null or misaligned pointer");
  if (n / 4)
    launch_pdl(mul_scale_bwd_kernel, flat_grid(n / 4, 256), 256, 0, as_stream(stream), 
        (const float4*)a, (const float4*)b, (const float4*)g, s, n / 4, (float4*)ga, (float4*)gb);
  if (n % 4) launch_pdl(tail_kernel, 1, 4, 0, as_stream(stream), 5, a, b, g, s, n - n % 4, n, ga, gb);
  return check_launch("gemnet_mul_scale_bwd");
}

GEMNET_API int gemnet_gather_rows(const float* x, const int32_t* idx, int64_t n, int32_t F, float* out, void* stream) {
  GEMNET_REQUIRE(n >= 0 && F > 0 && F % 4 == 0, "bad size (F must be a multiple of 4)");
  if (n == 0) return 0;
  GEMNET_REQUIRE(x && idx && out && aligned16(x) && aligned16(out), "null or misaligned pointer");
  launch_pdl(gather_rows_kernel, flat_grid(n * (F / 4), 256), 256, 0, as_stream(stream), (const float4*)x, idx, n, F / 4,
                                                                                (float4*)out);
  return check_launch("gemnet_gather_rows");
}

GEMNET_API int gemnet_segment_sum_rows(const float* src, const int32_t* perm, const int32_t* rowptr, int64_t n_rows,
                                       int32_t F, float* out, void* stream) {
  GEMNET_REQUIRE(n_rows >= 0 && F > 0 && F % 4 == 0, "bad size (F must be a multiple of 4)");
  if (n_rows == 0) return 0;
  GEMNET_REQUIRE(rowptr && out && aligned16(src) && aligned16(out), "null or misaligned pointer");
  launch_pdl(segment_sum_rows_kernel, flat_grid(n_rows * 32, 256), 256, 0, as_stream(stream), 
      (const float4*)src, perm, rowptr, n_rows, F / 4, (float4*)out);
  return check_launch("gemnet_segment_sum_rows");
}
