// N1: every projection of the radial (Bessel) basis of a GemNet evaluation in ONE kernel.
//
// The model applies a two-level bias-free linear map to rbf (E, R = num_radial) once per consumer:
//   level 1 (shared)   t_g = rbf W1_g^T            mlp_rbf3 / mlp_rbf_h / mlp_rbf_out [/ mlp_rbf4]   (gemnet.py:206-247)
//   level 2 (per use)  out_j = t_{g(j)} W2_j^T     TripletInteraction.mlp_rbf / QuadrupletInteraction.mlp_rbf
//                                                  (interaction_block.py:458-463,601-606), AtomUpdateBlock.dense_rbf
//                                                  (atom_update_block.py:36-41), OutputBlock.dense_rbf (:128)
// i.e. 13 (GemNet-T) or 18 (GemNet-Q) launches of a K = 16 GEMM per direction that sit on the critical path of the
// blocks that consume them.  K = 16 is no tensor-core shape worth a launch each: here one CUDA-core kernel writes
// all J outputs (grid = edge tiles x J), the reference's rounding order (level 1 rounded to fp32, then level 2) kept.
// Backward: tg_j = g_out_j W2_j (grid = edge tiles x J), then g_rbf = sum_j tg_j W1_{g(j)} in a fixed order
// (deterministic: no atomics).
#include "common.cuh"

namespace gemnet {
namespace fan {

constexpr int MAXJ = GEMNET_RBF_FAN_MAX_OUT, MAXG = GEMNET_RBF_FAN_MAX_GROUPS;
constexpr int EMB = 16;      // emb_size_rbf (level-1 width) supported
constexpr int TE = 32;       // edges per block

struct Params {
  const float* rbf;          // (E, R)
  const float* W1[MAXG];     // (EMB, R) each
  const float* W2[MAXJ];     // (H, EMB) each
  const float* gout[MAXJ];   // backward: (E, H) gradient of output j or NULL
  int32_t group[MAXJ];
  int64_t E;
  int32_t R, H, J, G;
};

// grid (ceil(E/TE), J), block H threads (thread = output column)
__global__ void rbf_fan_fwd_kernel(const __grid_constant__ Params p, float* __restrict__ out /* (J, E, H) */) {
  pdl_enter();
  __shared__ float t[TE][EMB];
  const int j = blockIdx.y, c = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * TE;
  const float* W1 = p.W1[p.group[j]];
  for (int idx = threadIdx.x; idx < TE * EMB; idx += blockDim.x) {
    const int el = idx / EMB, i = idx % EMB;
    const int64_t e = e0 + el;
    float acc = 0.f;
    if (e < p.E)
      for (int k = 0; k < p.R; ++k) acc = fmaf(__ldg(p.rbf + e * p.R + k), __ldg(W1 + i * p.R + k), acc);
    t[el][i] = acc;
  }
  float w[EMB];
#pragma unroll
  for (int i = 0; i < EMB; ++i) w[i] = __ldg(p.W2[j] + c * EMB + i);
  __syncthreads();
  float* dst = out + ((int64_t)j * p.E + e0) * p.H + c;
  const int n = (int)min((int64_t)TE, p.E - e0);
  for (int el = 0; el < n; ++el) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < EMB; ++i) acc = fmaf(t[el][i], w[i], acc);
    dst[(int64_t)el * p.H] = acc;
  }
}

// stage 1: tg[j][e][i] = sum_c g[j][e][c] W2_j[c][i];  grid (ceil(E/TE), J), block TE * EMB = 512 threads
__global__ void rbf_fan_bwd_stage1_kernel(const __grid_constant__ Params p, float* __restrict__ tg /* (J, E, EMB) */) {
  pdl_enter();
  extern __shared__ float sm[];
  float* gs = sm;                      // [TE][H]
  float* ws = sm + TE * p.H;           // [H][EMB]
  const int j = blockIdx.y;
  const int64_t e0 = (int64_t)blockIdx.x * TE;
  const float* g = p.gout[j];
  const int el = threadIdx.x / EMB, i = threadIdx.x % EMB;
  if (g == nullptr) {                  // this output had no consumer in the backward pass
    if (e0 + el < p.E) tg[((int64_t)j * p.E + e0 + el) * EMB + i] = 0.f;
    return;
  }
  for (int idx = threadIdx.x; idx < TE * p.H; idx += blockDim.x) {
    const int64_t e = e0 + idx / p.H;
    gs[idx] = e < p.E ? __ldg(g + e * p.H + idx % p.H) : 0.f;
  }
  for (int idx = threadIdx.x; idx < p.H * EMB; idx += blockDim.x) ws[idx] = __ldg(p.W2[j] + idx);
  __syncthreads();
  float acc = 0.f;
  const float* grow = gs + el * p.H;
  for (int c = 0; c < p.H; ++c) acc = fmaf(grow[c], ws[c * EMB + i], acc);
  if (e0 + el < p.E) tg[((int64_t)j * p.E + e0 + el) * EMB + i] = acc;
}

// stage 2: g_rbf[e][k] = sum_j sum_i tg[j][e][i] W1_{g(j)}[i][k]   (fixed order over j: deterministic)
__global__ void rbf_fan_bwd_stage2_kernel(const __grid_constant__ Params p, const float* __restrict__ tg,
                                          float* __restrict__ g_rbf /* (E, R) */) {
  pdl_enter();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.E * p.R) return;
  const int64_t e = idx / p.R;
  const int k = (int)(idx % p.R);
  float acc = 0.f;
  for (int j = 0; j < p.J; ++j) {
    const float* W1 = p.W1[p.group[j]];
    const float* t = tg + ((int64_t)j * p.E + e) * EMB;
#pragma unroll
    for (int i = 0; i < EMB; ++i) acc = fmaf(__ldg(t + i), __ldg(W1 + i * p.R + k), acc);
  }
  g_rbf[idx] = acc;
}

static int fill(Params& p, const float* rbf, int64_t E, int32_t R, const float* const* W1, int32_t G,
                const float* const* W2, const int32_t* group, int32_t J, int32_t H) {
  if (!(E >= 0 && R >= 1 && R <= 64 && G >= 1 && G <= MAXG && J >= 1 && J <= MAXJ && H >= 32 && H <= 512 && H % 32 == 0)) {
    set_error("rbf_fan: unsupported sizes (R <= 64, groups <= %d, outputs <= %d, H a multiple of 32 up to 512)", MAXG, MAXJ);
    return 1;
  }
  p.rbf = rbf; p.E = E; p.R = R; p.H = H; p.J = J; p.G = G;
  for (int g = 0; g < G; ++g) {
    if (!W1[g]) { set_error("rbf_fan: null level-1 weight"); return 1; }
    p.W1[g] = W1[g];
  }
  for (int j = 0; j < J; ++j) {
    if (!W2[j] || group[j] < 0 || group[j] >= G) { set_error("rbf_fan: bad level-2 weight / group"); return 1; }
    p.W2[j] = W2[j];
    p.group[j] = group[j];
  }
  return 0;
}

}  // namespace fan
}  // namespace gemnet

using namespace gemnet;

GEMNET_API int gemnet_rbf_fan_fwd(const float* rbf, int64_t n_edges, int32_t num_radial, const float* const* W1_host,
                                  int32_t n_groups, const float* const* W2_host, const int32_t* group_host,
                                  int32_t n_out, int32_t H, float* out, void* stream) {
  fan::Params p{};
  if (fan::fill(p, rbf, n_edges, num_radial, W1_host, n_groups, W2_host, group_host, n_out, H)) return 1;
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(rbf && out, "null pointer");
  launch_pdl(fan::rbf_fan_fwd_kernel, dim3((unsigned)ceil_div(n_edges, fan::TE), n_out), dim3(H), 0, as_stream(stream), p, out);
  return check_launch("gemnet_rbf_fan_fwd");
}

GEMNET_API int gemnet_rbf_fan_bwd(int64_t n_edges, int32_t num_radial, const float* const* W1_host, int32_t n_groups,
                                  const float* const* W2_host, const int32_t* group_host, int32_t n_out, int32_t H,
                                  const float* const* g_out_host, float* workspace, float* g_rbf, void* stream) {
  fan::Params p{};
  if (fan::fill(p, nullptr, n_edges, num_radial, W1_host, n_groups, W2_host, group_host, n_out, H)) return 1;
  if (n_edges == 0) return 0;
  GEMNET_REQUIRE(g_out_host && workspace && g_rbf, "null pointer");
  for (int j = 0; j < n_out; ++j) p.gout[j] = g_out_host[j];
  const size_t smem = sizeof(float) * ((size_t)fan::TE * H + (size_t)H * fan::EMB);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(fan::rbf_fan_bwd_stage1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("rbf_fan: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    configured = smem;
  }
  cudaStream_t st = as_stream(stream);
  launch_pdl(fan::rbf_fan_bwd_stage1_kernel, dim3((unsigned)ceil_div(n_edges, fan::TE), n_out), dim3(fan::TE * fan::EMB), smem, st,
             p, workspace);
  launch_pdl(fan::rbf_fan_bwd_stage2_kernel, dim3((unsigned)ceil_div(n_edges * num_radial, 256)), dim3(256), 0, st, p,
             (const float*)workspace, g_rbf);
  return check_launch("gemnet_rbf_fan_bwd");
}
