// Shared host/device helpers for the gemnet_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "../../include/gemnet_b200.h"

#define GEMNET_API extern "C" __attribute__((visibility("default")))

namespace gemnet {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for a flat elementwise kernel: enough CTAs to cover n, capped so that huge inputs use a
// grid-stride loop over a multiple of the SM count (148 SMs x 8 resident CTAs of 256 threads).
inline unsigned flat_grid(int64_t n, int block) {
  int64_t g = ceil_div(n, block);
  const int64_t cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

constexpr int kMaxSpherical = 8;   // Legendre / Y_lm orders supported by the register-resident recurrences
constexpr int kMaxRadial = 16;

// Programmatic dependent launch: allow the next kernel of the stream — if it was launched as a programmatic dependent
// (the fused dense layers are) — to start its prologue now.  It still waits for this grid to complete and flush
// before touching any activation.  A no-op otherwise.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// First statement of every kernel launched through launch_pdl(): let the successor start, then wait until the
// predecessor grid has completed and its writes are visible.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("GEMNET_B200_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Launch `kernel` as a programmatic dependent of the preceding kernel of the stream: its blocks may become resident
// while that kernel drains (this removes most of the kernel-to-kernel launch gap, also inside CUDA graphs); the
// kernel must call pdl_enter() before it touches global memory.  Errors surface through check_launch().
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// polynomial envelope u(x) and its derivative (reference envelope.py:14-29), x = d / cutoff
template <typename T>
__device__ __forceinline__ void envelope_val_grad(T x, int p, T& u, T& du) {
  if (!(x < T(1))) {
    u = T(0);
    du = T(0);
    return;
  }
  const T a = -T(p + 1) * T(p + 2) / T(2);
  const T b = T(p) * T(p + 2);
  const T c = -T(p) * T(p + 1) / T(2);
  T xp1 = T(1);  // x^(p-1)
  for (int i = 0; i < p - 1; ++i) xp1 *= x;
  const T xp = xp1 * x;
  u = T(1) + xp * (a + x * (b + x * c));
  du = xp1 * (a * T(p) + x * (b * T(p + 1) + x * c * T(p + 2)));
}

}  // namespace gemnet

#define GEMNET_REQUIRE(cond, msg)                 \
  do {                                            \
    if (!(cond)) {                                \
      gemnet::set_error("%s: %s", __func__, msg); \
      return 1;                                   \
    }                                             \
  } while (0)
