// Error plumbing + index preparation kernels (CSR row pointers, stable counting sort).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace gemnet {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// rowptr[r] = lower_bound(ids, r); also validates sortedness / range.
__global__ void csr_from_sorted_kernel(const int32_t* __restrict__ ids, int64_t n, int64_t n_rows,
                                       int32_t* __restrict__ rowptr, int32_t* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_rows; r += stride) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (ids[mid] < r) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = (int32_t)lo;
  }
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    int32_t v = ids[p];
    bool bad = v < 0 || v >= n_rows || (p + 1 < n && ids[p + 1] < v);
    if (bad) *flag = 1;
  }
}

__global__ void zero_i32_kernel(int32_t* p, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0;
}

// ids outside [0, n_rows) are not counted (and not placed): they raise `flag` instead of writing out of bounds
__global__ void histogram_kernel(const int32_t* __restrict__ ids, int64_t n, int64_t n_rows, int32_t* __restrict__ count,
                                 int32_t* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t r = ids[i];
    if (r < 0 || r >= n_rows) {
      if (flag) *flag = 1;
      continue;
    }
    atomicAdd(&count[r], 1);
  }
}

// single-CTA exclusive scan of count[0..n_rows) into rowptr[0..n_rows]; count is left zeroed (reused as cursor).
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(int32_t* __restrict__ count, int64_t n_rows,
                                                             int32_t* __restrict__ rowptr) {
  __shared__ int32_t partial[1024];
  const int tid = threadIdx.x;
  const int64_t chunk = (n_rows + 1023) / 1024;
  const int64_t lo = min((int64_t)tid * chunk, n_rows), hi = min(lo + chunk, n_rows);
  int32_t s = 0;
  for (int64_t i = lo; i < hi; ++i) s += count[i];
  partial[tid] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (int off = 1; off < 1024; off <<= 1) {
    int32_t v = tid >= off ? partial[tid - off] : 0;
    __syncthreads();
    partial[tid] += v;
    __syncthreads();
  }
  int32_t run = tid == 0 ? 0 : partial[tid - 1];
  for (int64_t i = lo; i < hi; ++i) {
    int32_t c = count[i];
    rowptr[i] = run;
    count[i] = 0;
    run += c;
  }
  if (tid == 1023) rowptr[n_rows] = partial[1023];
}

__global__ void place_kernel(const int32_t* __restrict__ ids, int64_t n, int64_t n_rows,
                             const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                             int32_t* __restrict__ perm) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int32_t r = ids[i];
    if (r < 0 || r >= n_rows) continue;
    int32_t slot = atomicAdd(&cursor[r], 1);
    perm[rowptr[r] + slot] = (int32_t)i;
  }
}

// make the placement deterministic / stable: ascending positions inside each row.  One warp per row ranks the
// row's (distinct) positions by counting the smaller ones — O(n^2 / 32) with independent loads, so a long row (a hub
// atom, or the pad atoms of a static-shape batch) costs microseconds where a per-thread insertion sort cost
// milliseconds.  `unordered` is the atomics-ordered placement, `perm` the final one.
__global__ void rank_rows_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows,
                                 const int32_t* __restrict__ unordered, int32_t* __restrict__ perm) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const int32_t b = rowptr[r], n = rowptr[r + 1] - b;
    if (n == 1) {
      if (lane == 0) perm[b] = unordered[b];
      continue;
    }
    for (int32_t i = lane; i < n; i += 32) {
      const int32_t v = unordered[b + i];
      int32_t rank = 0;
      for (int32_t j = 0; j < n; ++j) rank += unordered[b + j] < v;
      perm[b + rank] = v;
    }
  }
}

struct RangeLists {
  const int32_t* ids[GEMNET_RANGE_CHECK_MAX_LISTS];
  int64_t n[GEMNET_RANGE_CHECK_MAX_LISTS];
  int64_t bound[GEMNET_RANGE_CHECK_MAX_LISTS];
  int32_t count;
};

// every id of every list must lie in [0, bound): the gather / scatter kernels index their tables with them unchecked
__global__ void range_check_kernel(RangeLists L, int32_t* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int32_t l = 0; l < L.count; ++l) {
    const int32_t* __restrict__ ids = L.ids[l];
    const int64_t n = L.n[l], bound = L.bound[l];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int32_t v = ids[i];
      bad |= v < 0 || v >= bound;
    }
  }
  if (bad) *flag = 1;
}

}  // namespace gemnet

using namespace gemnet;

GEMNET_API int gemnet_abi_version(void) { return GEMNET_B200_ABI_VERSION; }
GEMNET_API const char* gemnet_last_error(void) { return gemnet::g_err; }

GEMNET_API int gemnet_csr_from_sorted(const int32_t* sorted_ids, int64_t n, int64_t n_rows, int32_t* rowptr,
                                      int32_t* flag, void* stream) {
  GEMNET_REQUIRE(n >= 0 && n_rows >= 0, "negative size");
  GEMNET_REQUIRE(rowptr != nullptr && flag != nullptr, "null output");
  GEMNET_REQUIRE(n == 0 || sorted_ids != nullptr, "null ids");
  int64_t work = n > n_rows + 1 ? n : n_rows + 1;
  csr_from_sorted_kernel<<<flat_grid(work, 256), 256, 0, as_stream(stream)>>>(sorted_ids, n, n_rows, rowptr, flag);
  return check_launch("gemnet_csr_from_sorted");
}

GEMNET_API int gemnet_csr_by_key(const int32_t* ids, int64_t n, int64_t n_rows, int32_t* rowptr, int32_t* perm,
                                 int32_t* workspace, int32_t* flag, void* stream) {
  GEMNET_REQUIRE(n >= 0 && n_rows >= 0, "negative size");
  GEMNET_REQUIRE(rowptr != nullptr && workspace != nullptr, "null output");
  GEMNET_REQUIRE(n == 0 || (ids != nullptr && perm != nullptr), "null ids/perm");
  cudaStream_t st = as_stream(stream);
  zero_i32_kernel<<<flat_grid(n_rows + 1, 256), 256, 0, st>>>(workspace, n_rows + 1);
  if (n > 0) histogram_kernel<<<flat_grid(n, 256), 256, 0, st>>>(ids, n, n_rows, workspace, flag);
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(workspace, n_rows, rowptr);
  if (n > 0) {
    int32_t* unordered = workspace + n_rows + 1;
    place_kernel<<<flat_grid(n, 256), 256, 0, st>>>(ids, n, n_rows, rowptr, workspace, unordered);
    rank_rows_kernel<<<flat_grid(n_rows * 32, 256), 256, 0, st>>>(rowptr, n_rows, unordered, perm);
  }
  return check_launch("gemnet_csr_by_key");
}

GEMNET_API int gemnet_index_range_check(const int32_t* const* lists_host, const int64_t* sizes_host,
                                        const int64_t* bounds_host, int32_t n_lists, int32_t* flag, void* stream) {
  GEMNET_REQUIRE(n_lists >= 0 && n_lists <= GEMNET_RANGE_CHECK_MAX_LISTS, "at most GEMNET_RANGE_CHECK_MAX_LISTS lists per call");
  GEMNET_REQUIRE(flag != nullptr, "null flag");
  GEMNET_REQUIRE(n_lists == 0 || (lists_host && sizes_host && bounds_host), "null argument");
  RangeLists L;
  L.count = 0;
  int64_t longest = 0;
  for (int32_t l = 0; l < n_lists; ++l) {
    GEMNET_REQUIRE(sizes_host[l] >= 0 && bounds_host[l] >= 0, "negative size");
    if (sizes_host[l] == 0) continue;
    GEMNET_REQUIRE(lists_host[l] != nullptr, "null list");
    L.ids[L.count] = lists_host[l];
    L.n[L.count] = sizes_host[l];
    L.bound[L.count] = bounds_host[l];
    ++L.count;
    if (sizes_host[l] > longest) longest = sizes_host[l];
  }
  if (L.count == 0) return 0;
  range_check_kernel<<<flat_grid(longest, 256), 256, 0, as_stream(stream)>>>(L, flag);
  return check_launch("gemnet_index_range_check");
}
