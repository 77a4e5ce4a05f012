// Micro-benchmark: issue rate of tcgen05.mma.cta_group::2 kind::f16 for M = 128 vs M = 256 (N = 256, K = 16),
// and cta_group::1 M = 128 / M = 64, on one CTA pair.  Operands are zero-filled smem; only timing matters.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../tdmpc2_b200/csrc/ptx.cuh"

__global__ void __cluster_dims__(2, 1, 1) rate_kernel(long long* out, int nmma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int rank = ptx::cluster_ctarank();
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async_smem();
  __syncthreads();
  ptx::cluster_sync();
  if (threadIdx.x < 32) ptx::tmem_alloc_2sm(&tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const uint32_t sa = ptx::smem_u32(smem), sb = sa + 32768;
  for (int cfg = 0; cfg < 2; ++cfg) {
    const uint32_t M = cfg == 0 ? 256 : 128;
    if (rank == 0 && threadIdx.x == 0) {
      const uint32_t idesc = ptx::make_idesc_f16(M, 256);
      const long long t0 = clock64();
      for (int i = 0; i < nmma; ++i) {
        const uint64_t ad = ptx::make_sw128_kmajor_desc(sa + (i & 3) * 32);
        const uint64_t bd = ptx::make_sw128_kmajor_desc(sb + (i & 3) * 32);
        ptx::umma_f16_2sm(tm, ad, bd, idesc, i > 0);
      }
      ptx::umma_commit_2sm(&bar);
      const long long t1 = clock64();
      ptx::mbar_wait(&bar, cfg & 1);
      const long long t2 = clock64();
      out[cfg * 2] = t1 - t0;
      out[cfg * 2 + 1] = t2 - t0;
    } else if (threadIdx.x == 0) {
      ptx::mbar_wait(&bar, cfg & 1);
    }
    __syncthreads();
    ptx::cluster_sync();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (threadIdx.x < 32) ptx::tmem_dealloc_2sm(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int nmma = 2048;
  rate_kernel<<<2, 128, 100 * 1024>>>(d, nmma);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  printf("cta_group::2 M=256 N=256 K=16: issue %.1f cyc/mma, complete %.1f cyc/mma\n", h[0] / double(nmma), h[1] / double(nmma));
  printf("cta_group::2 M=128 N=256 K=16: issue %.1f cyc/mma, complete %.1f cyc/mma\n", h[2] / double(nmma), h[3] / double(nmma));
  return 0;
}
