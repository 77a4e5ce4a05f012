// Micro-benchmark: is the ~21-27 B/clk/SM at which TMA delivers L2-resident operands a limit of the SM's input port or of
// the TMA path?  Every CTA streams the SAME 2 MiB region (L2-resident after the first touch) three ways:
//   T : cp.async.bulk.tensor (TMA) boxes of 16 KiB into a 9-slot smem ring (one producer thread, like the planner)
//   L : ld.global.nc.v4 by 16 warps, 8 loads in flight per thread (data xor-reduced, never stored)
//   C : cp.async.cg 16 B (LDGSTS) by 16 warps into a private smem ring, 8 groups in flight
// alone and together (T+L, T+C).  Reports bytes per SM clock per SM for each path.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../tdmpc2_b200/csrc/ptx.cuh"

constexpr int kSlots = 9;
constexpr int kBox = 64 * 128 * 2;          // 16 KiB
constexpr int kThreads = 26 * 32;           // warp 0 producer, warps 1-9 TMA consumers, warps 10-25 streaming / polling warps
constexpr int kStreamWarps = 16;
constexpr int kCpStages = 8;

struct Params {
  CUtensorMap tm;
  const uint4* region;       // 2 MiB
  int region_vec;            // uint4 elements in the region
  int do_tma, do_ldg, do_cp;
  int ncons;                 // TMA consumer warps (lane 0 each): 3 or 9
  int npoll;                 // streaming warps that instead POLL an mbarrier (all 32 lanes, try_wait loop) until the TMA stream ends
  int poll_sleep;            // nanosleep between polls (0 = tight loop)
  int nbox;                  // TMA boxes per CTA
  int niter;                 // streaming iterations per thread (each 8 x 16 B)
  long long* out;            // [grid][3]: cycles TMA path, cycles streaming path, xor sink
};

__global__ void __launch_bounds__(kThreads, 1) ingest_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots], empty[kSlots], done_bar;
  uint8_t* cp_area = smem + kSlots * kBox;  // kStreamWarps * kCpStages * 512 B = 64 KiB
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0 && P.do_tma) {
      for (int i = 0; i < P.nbox; ++i) {
        const int s = i % kSlots, ph = (i / kSlots) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&full[s], kBox);
        ptx::tma_load_2d(&P.tm, &full[s], smem + s * kBox, (i % 8) * 64, ((i / 8) % 16) * 128);
      }
    }
  } else if (warp < 1 + P.ncons) {
    if (lane == 0 && P.do_tma) {
      const int k = warp - 1;                     // this consumer frees slots s = k (mod ncons)
      for (int i = 0; i < P.nbox; ++i) {
        const int s = i % kSlots, ph = (i / kSlots) & 1;
        if (s % P.ncons != k) continue;
        ptx::mbar_wait(&full[s], ph);
        ptx::mbar_arrive(&empty[s]);
      }
      if (k == (P.nbox - 1) % kSlots % P.ncons) { P.out[blockIdx.x * 3 + 0] = clock64() - t0; ptx::mbar_arrive(&done_bar); }
    }
  } else if (warp >= 10) {
    const int sw = warp - 10;
    if (sw < P.npoll) {                           // like the planner's epilogue warps waiting for the accumulator
      while (!ptx::mbar_try_wait(&done_bar, 0)) { if (P.poll_sleep) __nanosleep(P.poll_sleep); }
    } else if (P.do_ldg) {
      uint4 acc = make_uint4(0, 0, 0, 0);
      size_t idx = (static_cast<size_t>(sw) * 32 + lane);
      for (int it = 0; it < P.niter; ++it) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const size_t j = (idx + static_cast<size_t>(u) * kStreamWarps * 32) % P.region_vec;
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                       : "l"(P.region + j));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
        idx += 8 * kStreamWarps * 32;
      }
      if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) P.out[blockIdx.x * 3 + 2] = 1;
      if (lane == 0 && sw == 0) P.out[blockIdx.x * 3 + 1] = clock64() - t0;
    } else if (P.do_cp) {
      uint8_t* mine = cp_area + (sw * kCpStages) * 512 + lane * 16;
      size_t idx = (static_cast<size_t>(sw) * 32 + lane);
      for (int it = 0; it < P.niter; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const size_t j = (idx + static_cast<size_t>(u) * kStreamWarps * 32) % P.region_vec;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(ptx::smem_u32(mine + u * 512)), "l"(P.region + j) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");      // two groups (16 x 16 B per thread) in flight
        idx += 8 * kStreamWarps * 32;
      }
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      if (lane == 0 && sw == 0) P.out[blockIdx.x * 3 + 1] = clock64() - t0;
    }
  }
  __syncthreads();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t region_bytes = 2u << 20;          // 2048 rows x 512 fp16
  void* buf; cudaMalloc(&buf, region_bytes); cudaMemset(buf, 1, region_bytes);
  long long* out; cudaMalloc(&out, 160 * 3 * 8);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  Params P{};
  cuuint64_t dims[2] = {512, 2048}; cuuint64_t strides[1] = {1024};
  cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  if (enc(&P.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n"); return 1;
  }
  P.region = static_cast<const uint4*>(buf); P.region_vec = static_cast<int>(region_bytes / 16); P.out = out;
  const int smem = kSlots * kBox + kStreamWarps * kCpStages * 512 + 1024;
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Case { int t, l, c, ncons, npoll, sleep; const char* name; };
  const Case cases[] = {{1, 0, 0, 3, 0, 0, "TMA alone"}, {0, 1, 0, 3, 0, 0, "LDG.128 alone"}, {0, 0, 1, 3, 0, 0, "cp.async alone"},
                        {1, 1, 0, 3, 0, 0, "TMA + LDG.128"}, {1, 0, 1, 3, 0, 0, "TMA + cp.async"},
                        {1, 0, 0, 9, 0, 0, "TMA, 9 consumers"}, {1, 0, 0, 3, 4, 0, "TMA, 4 warps poll"},
                        {1, 0, 0, 3, 16, 0, "TMA, 16 warps poll"}, {1, 0, 0, 3, 16, 100, "TMA, 16 poll+sleep100"},
                        {1, 0, 0, 3, 16, 1000, "TMA, 16 poll+sleep1k"}};
  for (int grid : {148}) {
    for (const Case& c : cases) {
      P.do_tma = c.t; P.do_ldg = c.l; P.do_cp = c.c; P.ncons = c.ncons; P.npoll = c.npoll; P.poll_sleep = c.sleep;
      P.nbox = 2048;                              // 32 MiB per CTA through TMA
      P.niter = 2048;                             // 16 warps x 32 lanes x 8 x 16 B x 2048 = 128 MiB per CTA through LDG / cp.async
      float best = 1e30f; long long ct = 0, cs = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(out, 0, 160 * 3 * 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        ingest_kernel<<<grid, kThreads, smem>>>(P);
        cudaEventRecord(e1);
        cudaError_t le = cudaGetLastError();
        cudaError_t se = cudaDeviceSynchronize();
        if (le != cudaSuccess) { printf("%s: launch %s\n", c.name, cudaGetErrorString(le)); return 1; }
        if (se != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(se)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) {
          best = ms;
          long long h[160 * 3]; cudaMemcpy(h, out, grid * 3 * 8, cudaMemcpyDeviceToHost);
          ct = cs = 0;
          for (int i = 0; i < grid; ++i) { ct = h[i * 3] > ct ? h[i * 3] : ct; cs = h[i * 3 + 1] > cs ? h[i * 3 + 1] : cs; }
        }
      }
      const double bt = double(P.nbox) * kBox, bs = double(P.niter) * kStreamWarps * 32 * 8 * 16;
      printf("grid %3d  %-22s: TMA %6.2f B/clk/SM   LDG/cp.async %6.2f B/clk/SM   (%.3f ms)\n", grid, c.name,
             c.t ? bt / double(ct) : 0.0, (c.l || c.c) ? bs / double(cs) : 0.0, best);
    }
  }
  return 0;
}
