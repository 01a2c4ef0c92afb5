// Probe for compute-sanitizer racecheck: ONE correctly synchronised cp.async.bulk (global -> shared, completion on an
// mbarrier via complete_tx) followed by plain shared loads of every thread after mbarrier.try_wait.  If racecheck
// reports a hazard here, its reports of the same shape for chain_tc.cu's weight staging buffer (bulk copy by the issuer
// warp, shared loads by the mover warps after waiting on w_full) are a limitation of the tool, not a race.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o racecheck_tma_probe racecheck_tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(const float* __restrict__ src, float* __restrict__ dst, int rounds) {
  __shared__ alignas(128) float buf[1024];
  __shared__ alignas(8) uint64_t bar, freed;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), f = (uint32_t)__cvta_generic_to_shared(&freed);
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f), "r"(blockDim.x / 32 - 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x < 32) {                      // producer warp (like the issuer): wait until the readers released the buffer
      if (r > 0) {
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(f), "r"((uint32_t)(r - 1) & 1u) : "memory");
      }
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4096u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s), "l"(src + (size_t)r * 1024), "r"(4096u), "r"(b) : "memory");
      }
    } else {                                     // reader warps (like the movers)
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b), "r"((uint32_t)r & 1u) : "memory");
      for (int i = threadIdx.x - 32; i < 1024; i += blockDim.x - 32) acc += buf[i];
      __syncwarp();
      if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(f) : "memory");
    }
  }
  dst[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const int rounds = 8;
  float *src, *dst;
  cudaMalloc(&src, rounds * 4096);
  cudaMalloc(&dst, 4 * 160 * sizeof(float));
  cudaMemset(src, 0, rounds * 4096);
  probe<<<4, 160>>>(src, dst, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
