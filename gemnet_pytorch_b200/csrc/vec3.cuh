// float3 helpers and the clamped angle primitive shared by the geometry kernels.
#pragma once
#include "common.cuh"

namespace gemnet {

struct f3 {
  float x, y, z;
};
__device__ __forceinline__ f3 ld3(const float* __restrict__ p, int64_t i) {
  return {__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)};
}
__device__ __forceinline__ f3 sub(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 add(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 mul(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ void atomic_add3(float* __restrict__ p, int64_t i, f3 v) {
  atomicAdd(p + 3 * i, v.x);
  atomicAdd(p + 3 * i + 1, v.y);
  atomicAdd(p + 3 * i + 2, v.z);
}

// ------------------------------------------------------------------------------------------ angles
// shared by the triplet and the quadruplet-intermediate kernels
struct AnglePair {
  float x, y, yraw, r;  // x = u.v, y = max(|u x v|, 1e-9), r = sqrt(x^2 + y^2)
  f3 w;                 // u x v
};
__device__ __forceinline__ AnglePair angle_pair(f3 u, f3 v) {
  AnglePair a;
  a.x = dot(u, v);
  a.w = cross(u, v);
  a.yraw = sqrtf(dot(a.w, a.w));
  a.y = fmaxf(a.yraw, 1e-9f);  // gemnet.py:309
  a.r = sqrtf(a.x * a.x + a.y * a.y);
  return a;
}
// gradient of (cos, sin) = (x/r, y/r) w.r.t. u and v given upstream (g_cos, g_sin)
__device__ __forceinline__ void angle_pair_bwd(const AnglePair& a, f3 u, f3 v, float g_cos, float g_sin, f3& gu,
                                               f3& gv) {
  const float inv_r3 = 1.0f / (a.r * a.r * a.r);
  const float gx = (g_cos * a.y * a.y - g_sin * a.x * a.y) * inv_r3;
  float gy = (-g_cos * a.x * a.y + g_sin * a.x * a.x) * inv_r3;
  if (!(a.yraw > 1e-9f)) gy = 0.f;  // clamped: torch.max routes the gradient to the constant
  f3 gw = a.yraw > 0.f ? mul(a.w, gy / a.yraw) : f3{0.f, 0.f, 0.f};
  gu = add(mul(v, gx), cross(v, gw));
  gv = add(mul(u, gx), cross(gw, u));
}


}  // namespace gemnet
