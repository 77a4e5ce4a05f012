// Micro-benchmark: L2 -> shared-memory TMA throughput when every CTA streams the SAME weight tensor (the access
// pattern of the planning kernels: all SMs read one layer's 1 MiB of packed weights at about the same time).
//   mode 0: unicast, every CTA loads every 16 KiB box itself
//   mode 1: multicast, the C CTAs of a cluster each load 1/C of a box and multicast it to all C
//   mode 2: unicast, every CTA streams its own private region (the activation-plane pattern)
//   mode 3: unicast, every box fetched as C separate row slices by the same CTA (isolates the request-size effect)
// Reports delivered bytes per SM clock per SM, wall-clock TB/s and how many clusters of each size fit the chip.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <vector>
#include <cuda_runtime.h>
#include "../../tdmpc2_b200/csrc/ptx.cuh"

#ifndef TB_SLOTS
#define TB_SLOTS 12
#endif
constexpr int kSlots = TB_SLOTS;   // ring capacity; P.nslots of them are used
constexpr int kBox = 64 * 128 * 2;   // 16 KiB: 64 fp16 columns x 128 rows

struct Params {
  CUtensorMap tm[4];   // box rows 128, 64, 32, 16
  int mode, csize, nbox, rows_total, private_rows, nslots;
  long long* out;
};

__device__ __forceinline__ void arrive_remote(uint64_t* bar, int rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(ptx::smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(640, 1) bw_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots], empty[kSlots];
  const int rank = ptx::cluster_ctarank();
  const int C = P.mode == 1 ? P.csize : 1;   // CTAs that write into (and must free) each slot
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], C); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (P.csize > 1 && P.mode != 3) ptx::cluster_sync();
  const int ncols = 512 / 64;   // boxes per row block
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    // producer
    for (int i = 0; i < P.nbox; ++i) {
      const int s = i % P.nslots, ph = (i / P.nslots) & 1;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      ptx::mbar_expect_tx(&full[s], kBox);
      int rb = (i / ncols) % (P.rows_total / 128);
      const int cb = i % ncols;
      if (P.mode == 2) rb = (blockIdx.x * (P.private_rows / 128) + (i / ncols) % (P.private_rows / 128));
      if (P.mode == 1 && C > 1) {
        const int rows = 128 / C;
        const int tmi = C == 2 ? 1 : C == 4 ? 2 : 3;
        tma_load_2d_mc(&P.tm[tmi], &full[s], smem + s * kBox + rank * rows * 128, cb * 64, rb * 128 + rank * rows,
                       static_cast<uint16_t>((1u << C) - 1));
      } else if (P.mode == 3) {
        const int S = P.csize, rows = 128 / S;
        const int tmi = S == 2 ? 1 : S == 4 ? 2 : 3;
        for (int q = 0; q < S; ++q)
          ptx::tma_load_2d(&P.tm[tmi], &full[s], smem + s * kBox + q * rows * 128, cb * 64, rb * 128 + q * rows);
      } else {
        ptx::tma_load_2d(&P.tm[0], &full[s], smem + s * kBox, cb * 64, rb * 128);
      }
    }
  } else if ((threadIdx.x & 31) == 0) {
    // consumers: warp (r, k) frees, in CTA r, the slots s = k (mod K).  One warp owns a slot for the whole run (a
    // parity wait must never lag a barrier by two phases); remote arrives are slow, so K of them are in flight per target.
#ifdef TB_CONS3
    const int w = (threadIdx.x >> 5) - 1, K = 3, r = w / K, k = w % K;
#else
    const int w = (threadIdx.x >> 5) - 1, K = C == 8 ? 2 : 12 / C, r = w / K, k = w % K;
#endif
    if (r < C) {
      for (int i = 0; i < P.nbox; ++i) {
        const int s = i % P.nslots, ph = (i / P.nslots) & 1;
        if (s % K != k) continue;
        ptx::mbar_wait(&full[s], ph);
        if (C > 1) arrive_remote(&empty[s], r);
        else ptx::mbar_arrive(&empty[s]);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (P.csize > 1 && P.mode != 3) ptx::cluster_sync();
  if (threadIdx.x == 0) P.out[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows_shared = 2048;            // 2048 x 512 fp16 = 2 MiB: hi + lo planes of a 512 x 512 layer, twice
  const int private_rows = 512;            // per CTA: 512 x 512 fp16 = 512 KiB
  const int grid_max = 160;
  const size_t rows_all = rows_shared + static_cast<size_t>(grid_max) * private_rows;
  void* buf; cudaMalloc(&buf, rows_all * 512 * 2); cudaMemset(buf, 0, rows_all * 512 * 2);
  long long* out; cudaMalloc(&out, grid_max * 8);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  Params P{};
  for (int i = 0; i < 4; ++i) {
    cuuint64_t dims[2] = {512, rows_all}; cuuint64_t strides[1] = {1024};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(128 >> i)}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&P.tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  P.out = out; P.rows_total = rows_shared; P.private_rows = private_rows; P.nbox = 4096;   // 64 MiB per CTA
  const int smem = kSlots * kBox + 1024;
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int cs : {2, 4, 8}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = 220 * 1024;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {unsigned(cs), 1, 1};
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, bw_kernel, &cfg);
    printf("clusters of %d (640 threads, 220 KiB smem): max active clusters %d -> %d CTAs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  struct Case { int mode, csize; const char* name; };
  const Case cases[] = {{0, 1, "unicast shared"}, {2, 1, "unicast private"}, {0, 2, "unicast shared, cluster 2"},
                        {3, 4, "unicast shared, 4 KiB slices"}, {1, 2, "multicast 2"}, {1, 4, "multicast 4"},
                        {1, 8, "multicast 8"}};
  // partial grids (unicast shared): does the per-SM delivery rate rise when fewer SMs stream at once (chip-level L2 bound)
  // or stay put (per-SM ingest port bound)?  Decides whether de-phasing the CTAs' GEMM phases can help.
  struct Run { Case c; int grid; };
  std::vector<Run> runs;
  for (const Case& c : cases) {
    const int cl = c.mode == 3 ? 1 : c.csize;
    // whole clusters that are co-resident (see the occupancy lines above): 148 / 132 / 120 CTAs
    runs.push_back({c, cl == 4 ? 132 : cl == 8 ? 120 : 148});
  }
  for (int g : {111, 74, 37, 16, 4}) { runs.push_back({cases[0], g}); runs.push_back({cases[2], g & ~1}); }
  for (int nslots : {kSlots})
  for (const Run& r : runs) {
    const Case& c = r.c;
    P.mode = c.mode; P.csize = c.csize; P.nslots = nslots;
    const int cl = c.mode == 3 ? 1 : c.csize;
    const int grid = r.grid;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(544); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {unsigned(cl), 1, 1};
#ifdef TB_NOCLUSTER
    cfg.attrs = at; cfg.numAttrs = (cl > 1) ? 1 : 0;
#else
    cfg.attrs = at; cfg.numAttrs = 1;
#endif
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f; long long cyc = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      cudaError_t le = cudaLaunchKernelEx(&cfg, bw_kernel, P);
      cudaEventRecord(e1);
      cudaError_t se = cudaDeviceSynchronize();
      if (le != cudaSuccess || se != cudaSuccess) { printf("%s: %s / %s\n", c.name, cudaGetErrorString(le), cudaGetErrorString(se)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) {
        best = ms;
        long long h[grid_max]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        cyc = 0; for (int i = 0; i < grid; ++i) cyc = h[i] > cyc ? h[i] : cyc;
      }
    }
    const double bytes = double(P.nbox) * kBox;
    printf("%2d slots  %-28s grid %3d: %6.2f B/clk/SM delivered (slowest CTA), %6.2f TB/s delivered chip-wide, %.3f ms\n", nslots, c.name, grid,
           bytes / double(cyc), bytes * grid / (best * 1e-3) / 1e12, best);
  }
  return 0;
}
