// Ping-pong CEM-iteration kernel for sm_100a: GEMM and epilogue of one SM overlap.
//
// Why: with a 128-row tile the fp32 accumulator of a 512-wide layer fills all 512 TMEM columns, so the tensor
// pipe idles while the LayerNorm / Mish epilogue runs (profiles/README.md).  Here every CTA splits its 128-row
// tile into two 64-row halves.  CTA pairs issue tcgen05.mma.cta_group::2 with M = 128 (64 rows from each CTA),
// which runs at the full MMA rate (scripts/micro/mma_rate.cu: 64 cycles per N=256,K=16) and stores a 64 x N
// accumulator as 128 lanes x N/2 columns ("2x2" layout: lanes [0,64) hold the first half of every 256-column
// chunk, lanes [64,128) the second half).  Two halves therefore fit TMEM side by side, and while the MMA warp
// works on one half the 16 epilogue warps drain the other:
//
//     MMA      : G(A,0) G(B,0) G(A,1) G(B,1) G(A,2) ...
//     epilogue :        E(A,0) E(B,0) E(A,1) E(B,1) ...
//
// Roles are decoupled with mbarriers: act_ready[h] (epilogue -> producer/MMA: the planes of half h are
// published), facc[h] (MMA -> epilogue: accumulator complete, multicast to both CTAs), acc_free[h] (both CTAs'
// epilogues -> leader's MMA thread: TMEM of half h may be overwritten).  Everything the reference does per CEM
// iteration (tdmpc2.py:173-197) happens in this one launch, as in plan_kernel MODE_ITER.
//
// Scope: MODE_ITER, models whose LayerNorm layers are 256- or 512-wide and whose heads fit 128 columns (the 1M /
// 5M presets), num_samples a multiple of 128 with an even number of tiles per environment.  Anything else runs on
// plan_kernel.
#pragma once
#include "plan_kernels.cuh"

namespace tdmpc2 {

// This kernel keeps the 20-warp layout it was validated with (plan_kernel moved to 18 warps for register headroom).
constexpr int kPPThreads = 640;
constexpr int kPPEpiWarp0 = 4;                  // warps 4..19: epilogue
struct GroupEpiPP { static constexpr int N = kEpiThreads; __device__ static int tid() { return threadIdx.x - kPPEpiWarp0 * 32; } __device__ static void sync() { epi_bar_sync(); } };
constexpr int kPPHalf = 64;
constexpr int kPPAPlane = kPPHalf * 128;               // 8 KiB: 64 rows x 64 fp16
constexpr int kPPASlot = 2 * kPPAPlane;                // hi | lo
constexpr int kPPARing = 3;
constexpr int kPPWPlane = 128 * 128;                   // 16 KiB: this CTA's 128 weight rows of a 256-column chunk
constexpr int kPPWSlot = 2 * kPPWPlane;
// Operand rings: what bounds a half's GEMM is the number of operand bytes in flight (a TMA request takes ~2.8 k cycles
// from issue to a released slot; the hardware delivers > 65 B/clk/SM when enough requests are outstanding:
// profiles/r02_nomma_*.txt, r02_micro_ingest_paths.txt), and a half needs 640 KB per 12.3 k cycles of MMA.  Four weight
// slots instead of three; the 32 KiB come out of the epilogue's staging tiles, which shrink to 64 rows x 16 columns.
constexpr int kPPWRing = 4;
constexpr int kPPWOff = kPPARing * kPPASlot;           // 48 KiB
constexpr int kPPStgOff = kPPWOff + kPPWRing * kPPWSlot;   // 176 KiB
constexpr int kPPStgCols = 16;                         // columns per staged block
constexpr int kPPStgPlane = kPPHalf * kPPStgCols * 2;  // 2 KiB: 64 rows x 16 fp16, row-major (32-byte rows, no swizzle)
constexpr int kPPStgBuf = 2 * kPPStgPlane;             // hi | lo
constexpr int kPPGroups = 8;                           // block groups: 2 column halves x 4 column groups, 2 warps each
constexpr int kPPOperandBytes = kPPStgOff + kPPGroups * kPPStgBuf;   // 208 KiB
constexpr int kPPCtrl = 2048;
constexpr int kPPVec = 3 * kFusedMaxN * 4;             // bias | ln_g | ln_b (third slot doubles as exchange buffer)
constexpr int kPPPart = 2 * kPPGroups * kPPHalf * 4;   // two partial arrays [8][64]
constexpr int kPPAct = 3 * kMaxHeadCols * 4;           // mean | std | mask of one time step
constexpr int kPPMaxSteps = 48;
constexpr int kPPProg = 2 * kPPMaxSteps * 32;          // double-buffered layer program
constexpr int kPPSmemBytes = kPPOperandBytes + kPPCtrl + kPPVec + kPPPart + kPPAct + kPPProg + 1024;
static_assert(kPPSmemBytes <= 232448, "ping-pong kernel shared memory");
static_assert(kPPGroups * kPPStgBuf >= kPPHalf * 128 * 4, "the pi head's exchange buffer aliases the staging area");

// Diagnostics: clock stamps of CTA 0's first tile, [step][event] after the per-CTA counters (see TDMPC2_TRACE).
#define PP_TRACE(P_, on_, s_, ev_)                                                           \
  do {                                                                                      \
    if (kProf && (P_).prof && (on_) && (s_) < 32) (P_).prof[static_cast<size_t>((P_).prof_slots) * 4 * 12 + (s_) * 16 + (ev_)] = prof_clock(); \
  } while (0)

struct PPStep {
  int li, src, kind, dstbuf, head, t_act;   // t_act >= 0: X action columns must hold a_t before this step's GEMM
  float disc;
  int pad;
};
static_assert(sizeof(PPStep) == 32, "PPStep size");

struct PPCtx {
  uint8_t* base;
  uint64_t *a_full, *a_empty, *w_full, *w_empty, *facc, *acc_free, *act_ready;
  uint32_t* tmem_ptr;
  int* flags;
  float *G, *q1, *vec, *part, *actv;
  PPStep* prog;
  uint32_t tmem_base;
  int slot, warp, lane, rank;
};

__device__ __forceinline__ void pp_arrive_leader(uint64_t* bar, int rank) {
  // arrive on the barrier at this smem offset in the LEADER CTA (rank 0) of the pair
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(ptx::smem_u32(bar)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(raddr) : "memory");
  (void)rank;
}
__device__ __forceinline__ void pp_group_sync(int bg) {   // the 2 warps that share a 64 x 32 output block
  asm volatile("bar.sync %0, 64;\n" ::"r"(2 + bg) : "memory");
}

// value bookkeeping of _estimate_value for one row (see head_commit in plan_kernels.cuh)
__device__ __forceinline__ void pp_head_commit(const PlanParams& P, PPCtx& c, const PPStep& st, int tile, int r, float val) {
  if (st.head == HEAD_REWARD) {
    c.G[r] = __fadd_rn(c.G[r], __fmul_rn(st.disc, val));
  } else if (st.head == HEAD_Q1) {
    c.q1[r] = val;
  } else {
    const float qavg = __fmul_rn(__fadd_rn(c.q1[r], val), 0.5f);
    const float v = nan_to_num0(__fadd_rn(c.G[r], __fmul_rn(st.disc, qavg)));
    const RowMap rm = map_row(P, tile, r);
    if (rm.env >= 0) P.values[static_cast<size_t>(rm.env) * P.N + rm.idx] = v;
  }
}

// X action columns of rows [r0, r0 + nrows) <- a_t (tdmpc2.py:176-181); epilogue threads only
__device__ __forceinline__ void pp_write_actions(const PlanParams& P, PPCtx& c, int tile, int env, int task, int t, int r0,
                                                 int nrows, bool stage) {
  const int tid = threadIdx.x - kPPEpiWarp0 * 32;
  float* sm_mean = c.actv; float* sm_std = c.actv + kMaxHeadCols; float* sm_mask = c.actv + 2 * kMaxHeadCols;
  if (stage) {
    for (int a = tid; a < P.A; a += kEpiThreads) {
      const size_t sa = (static_cast<size_t>(env) * P.H + t) * P.A + a;
      sm_mean[a] = P.mean[sa]; sm_std[a] = P.std[sa];
      sm_mask[a] = P.masks ? P.masks[static_cast<size_t>(task) * P.A + a] : 1.f;
    }
    epi_bar_sync();
  }
  __half* xhi = plane_ptr(P, c.slot, BUF_X, 0);
  __half* xlo = plane_ptr(P, c.slot, BUF_X, 1);
  const int n0 = (tile % P.tiles_per_env) * kTileM;
  const float* nz = P.noise_r + (static_cast<size_t>(env) * P.H + t) * (P.N - P.P) * P.A;
  const float* pa = P.pi_actions + (static_cast<size_t>(env) * P.H + t) * P.P * P.A;
  const int ng = (P.A + 7) >> 3;
  if (P.L + P.T + 8 * ng <= P.KpadX) {            // ((L + T) % 8 == 0 is a condition of this kernel: pp_eligible)
    // one (row, 8 action columns) item per thread, two 16-byte stores per item (same pass as plan_kernel's)
    for (int i = tid; i < nrows * ng; i += kEpiThreads) {
      const int r = r0 + i / ng, g8 = (i % ng) * 8;
      const int n = n0 + r;
      const float* src = (n < P.P) ? pa + static_cast<size_t>(n) * P.A : nz + static_cast<size_t>(n - P.P) * P.A;
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = (g8 + u < P.A) ? (n < P.P ? src[g8 + u] : (P.rng_state ? 0.f : __ldcs(src + g8 + u))) : 0.f;
      if (n >= P.P && P.rng_state) {                                 // in-kernel noise (rng.cuh): two groups of four
        const float4 g0 = rng_normal4(P.rng_state, 2u * P.rng_iter, rng_group_r(P, env, t, n, g8 >> 2));
        const float4 g1 = rng_normal4(P.rng_state, 2u * P.rng_iter, rng_group_r(P, env, t, n, (g8 >> 2) + 1));
        x[0] = g0.x; x[1] = g0.y; x[2] = g0.z; x[3] = g0.w; x[4] = g1.x; x[5] = g1.y; x[6] = g1.z; x[7] = g1.w;
      }
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        __half h[2], l[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int a = min(g8 + u + w, P.A - 1);
          float v = x[u + w];
          if (n >= P.P) v = fminf(fmaxf(__fadd_rn(sm_mean[a], __fmul_rn(sm_std[a], v)), -1.f), 1.f);
          v = (g8 + u + w < P.A) ? v * sm_mask[a] : 0.f;
          split_f(v, h[w], l[w]);
        }
        hw[u >> 1] = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
        lw[u >> 1] = static_cast<uint32_t>(__half_as_ushort(l[0])) | (static_cast<uint32_t>(__half_as_ushort(l[1])) << 16);
      }
      const size_t o = static_cast<size_t>(r) * P.KpadX + P.L + P.T + g8;
      __stcg(reinterpret_cast<uint4*>(xhi + o), make_uint4(hw[0], hw[1], hw[2], hw[3]));
      __stcg(reinterpret_cast<uint4*>(xlo + o), make_uint4(lw[0], lw[1], lw[2], lw[3]));
    }
    return;
  }
#pragma unroll 4
  for (int i = tid; i < nrows * P.A; i += kEpiThreads) {
    const int r = r0 + i / P.A, a = i % P.A;
    const int n = n0 + r;
    float v;
    if (n < P.P) v = pa[static_cast<size_t>(n) * P.A + a];
    else {
      v = __fadd_rn(sm_mean[a], __fmul_rn(sm_std[a], noise_r_at(P, env, t, n, a)));
      v = fminf(fmaxf(v, -1.f), 1.f);
    }
    v *= sm_mask[a];
    const size_t o = static_cast<size_t>(r) * P.KpadX + P.L + P.T + a;
    split_store(xhi + o, xlo + o, v);
  }
}

// ---------------------------------------------------------------------------------------------- epilogues
struct PPThread {
  int q, grp, row, colhalf, bg;
  uint32_t taddr;      // TMEM address of this thread's lane, column 0
};

// LayerNorm (+ Mish | SimNorm) of one 64-row half; Npad a multiple of 256, N == Npad.
template <int KIND>
__device__ __forceinline__ void pp_epi_ln(const PlanParams& P, PPCtx& c, const PPThread& et, const LayerDev& ly, const PPStep& st,
                                          int h) {
  const float* sb = c.vec; const float* sg = c.vec + kFusedMaxN; const float* sbe = c.vec + 2 * kFusedMaxN;
  const int nnc = ly.Npad / kNch;
  const float inv_scale = ly.inv_scale;
  const float2 inv2 = f2s(inv_scale);
  const uint32_t tcol0 = c.tmem_base + static_cast<uint32_t>(h * 256 + et.grp * 32);
  // ---- pass 1: shifted moments over this thread's 32 columns of every chunk
  float x0 = 0.f;
  float2 sa = f2s(0.f), sb2 = f2s(0.f), qa = f2s(0.f), qb = f2s(0.f);
  for (int nc = 0; nc < nnc; ++nc) {
    const int c0 = nc * kNch + et.colhalf * 128 + et.grp * 32;
    uint32_t v[32];
    ptx::tmem_ld_32x32(et.taddr + tcol0 + nc * 128, v);
    ptx::tmem_ld_wait();
    if (nc == 0) x0 = fmaf(__uint_as_float(v[0]), inv_scale, sb[c0]);
    const float2 nx0 = f2s(-x0);
#pragma unroll
    for (int i4 = 0; i4 < 32; i4 += 4) {
      const float4 b4 = lds128(sb + c0 + i4);
      const float2 xa = __ffma2_rn(f2(__uint_as_float(v[i4]), __uint_as_float(v[i4 + 1])), inv2, __fadd2_rn(f2(b4.x, b4.y), nx0));
      const float2 xb = __ffma2_rn(f2(__uint_as_float(v[i4 + 2]), __uint_as_float(v[i4 + 3])), inv2, __fadd2_rn(f2(b4.z, b4.w), nx0));
      sa = __fadd2_rn(sa, xa); qa = __ffma2_rn(xa, xa, qa);
      sb2 = __fadd2_rn(sb2, xb); qb = __ffma2_rn(xb, xb, qb);
    }
  }
  {
    const float cnt = static_cast<float>(32 * nnc);
    const float s = (sa.x + sa.y) + (sb2.x + sb2.y), q = (qa.x + qa.y) + (qb.x + qb.y);
    c.part[et.bg * kPPHalf + et.row] = x0 + s / cnt;                              // partial mean
    c.part[(kPPGroups + et.bg) * kPPHalf + et.row] = q - s * s / cnt;             // partial M2
  }
  epi_bar_sync();
  float mean = 0.f, rstd;
  {
    float m2 = 0.f;
#pragma unroll
    for (int g = 0; g < kPPGroups; ++g) mean += c.part[g * kPPHalf + et.row];
    mean *= (1.f / kPPGroups);
    const float cnt = static_cast<float>(32 * nnc);
#pragma unroll
    for (int g = 0; g < kPPGroups; ++g) {
      const float d = c.part[g * kPPHalf + et.row] - mean;
      m2 += c.part[(kPPGroups + g) * kPPHalf + et.row] + cnt * d * d;
    }
    rstd = rsqrtf(m2 / static_cast<float>(ly.N) + 1e-5f);       // nn.LayerNorm eps, biased variance
  }
  const float2 rstd2 = f2s(rstd), nmr2 = f2s(-mean * rstd);
  // ---- pass 2: normalise, activate, split, stage, TMA-store one 64 x 16 block per 16 columns
  uint8_t* buf = c.base + kPPStgOff + et.bg * kPPStgBuf;
  const uint32_t rowaddr = ptx::smem_u32(buf) + static_cast<uint32_t>(et.row) * (kPPStgCols * 2u);
  const bool lead_warp = (et.q & 1) == 0;              // of the two warps that share a block; one elected lane of it issues
  const CUtensorMap* tmD = (st.dstbuf == BUF_X) ? &P.tmXs64 : &P.tmHs64;
  const int row_hi = plane_row0(P, c.slot, st.dstbuf, 0) + h * kPPHalf, row_lo = plane_row0(P, c.slot, st.dstbuf, 1) + h * kPPHalf;
  for (int nc = 0; nc < nnc; ++nc) {
    const int c0 = nc * kNch + et.colhalf * 128 + et.grp * 32;
#pragma unroll
    for (int sub = 0; sub < 32; sub += 16) {
      uint32_t v[16];
      ptx::tmem_ld_32x16(et.taddr + tcol0 + nc * 128 + sub, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i4 = 0; i4 < 16; i4 += 4) {
        const float4 b4 = lds128(sb + c0 + sub + i4);
        const float4 g4 = lds128(sg + c0 + sub + i4);
        const float4 e4 = lds128(sbe + c0 + sub + i4);
        float2 t0 = __ffma2_rn(__ffma2_rn(__ffma2_rn(f2(__uint_as_float(v[i4]), __uint_as_float(v[i4 + 1])), inv2, f2(b4.x, b4.y)),
                                          rstd2, nmr2), f2(g4.x, g4.y), f2(e4.x, e4.y));
        float2 t1 = __ffma2_rn(__ffma2_rn(__ffma2_rn(f2(__uint_as_float(v[i4 + 2]), __uint_as_float(v[i4 + 3])), inv2, f2(b4.z, b4.w)),
                                          rstd2, nmr2), f2(g4.z, g4.w), f2(e4.z, e4.w));
        if (KIND == EPI_LN_MISH) { t0 = mish_fast2(t0); t1 = mish_fast2(t1); }
        v[i4] = __float_as_uint(t0.x); v[i4 + 1] = __float_as_uint(t0.y);
        v[i4 + 2] = __float_as_uint(t1.x); v[i4 + 3] = __float_as_uint(t1.y);
      }
      if (KIND == EPI_LN_SIMNORM) {
#pragma unroll
        for (int g0 = 0; g0 < 16; g0 += 8) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(v[g0 + i]);
          float m = y[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) m = fmaxf(m, y[i]);
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { y[i] = exp_fast(y[i] - m); t += y[i]; }
          const float rt = rcp_ftz(t);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[g0 + i] = __float_as_uint(y[i] * rt);
        }
      }
      uint32_t hw[8], lw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a0 = __uint_as_float(v[2 * i]), a1 = __uint_as_float(v[2 * i + 1]);
        const __half2 h2 = __floats2half2_rn(a0, a1);
        const float2 hf = __half22float2(h2);
        const float2 df = __fadd2_rn(f2(a0, a1), f2(-hf.x, -hf.y));
        const __half2 l2 = __floats2half2_rn(df.x, df.y);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      if (nc > 0 || sub > 0) {
        // single staging block per group: its previous store (issued a whole block of math ago) must have read it
        if (lead_warp && ptx::elect_one()) ptx::bulk_wait_read<0>();   // (the same lane issues, commits and waits)
        pp_group_sync(et.bg);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        ptx::st_shared_v4(rowaddr + 16u * i, hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
        ptx::st_shared_v4(rowaddr + kPPStgPlane + 16u * i, lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
      }
      ptx::fence_proxy_async_smem();
      pp_group_sync(et.bg);
      if (lead_warp && ptx::elect_one()) {
        if (P.l2hint & 2) {                                   // evict-last: re-read by the next layer (see plan_kernels.cuh)
          ptx::tma_store_2d_hint(tmD, buf, c0 + sub, row_hi, ptx::kL2EvictLast);   // LN outputs start at column 0 of their buffer
          ptx::tma_store_2d_hint(tmD, buf + kPPStgPlane, c0 + sub, row_lo, ptx::kL2EvictLast);
        } else {
          ptx::tma_store_2d(tmD, buf, c0 + sub, row_hi);
          ptx::tma_store_2d(tmD, buf + kPPStgPlane, c0 + sub, row_lo);
        }
        ptx::bulk_commit();
      }
    }
  }
}

// two_hot_inv (math.py:74-83) for a head of <= 128 bins; accumulator chunk is 128 columns wide ("2x2": lanes < 64
// hold bins [0,64), lanes >= 64 hold bins [64,128)); every thread owns 16 bins of its row.
__device__ __forceinline__ void pp_epi_twohot(const PlanParams& P, PPCtx& c, const PPThread& et, const LayerDev& ly, const PPStep& st,
                                              int h, int tile) {
  const float* sb = c.vec;
  const float* bins = c.vec + kFusedMaxN;
  float* xch = c.vec + 2 * kFusedMaxN;                                 // [8][64]
  const int B = P.B, c0 = et.colhalf * 64 + et.grp * 16;
  uint32_t v[16];
  ptx::tmem_ld_32x16(et.taddr + c.tmem_base + static_cast<uint32_t>(h * 256 + et.grp * 16), v);
  ptx::tmem_ld_wait();
  float x[16];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    x[i] = (c0 + i < B) ? fmaf(__uint_as_float(v[i]), ly.inv_scale, sb[min(c0 + i, B - 1)]) : -CUDART_INF_F;
    m = fmaxf(m, x[i]);
  }
  c.part[et.bg * kPPHalf + et.row] = m;
  epi_bar_sync();
#pragma unroll
  for (int g = 0; g < kPPGroups; ++g) m = fmaxf(m, c.part[g * kPPHalf + et.row]);
  float ssum = 0.f, acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float e = (c0 + i < B) ? exp_fast(x[i] - m) : 0.f;
    ssum += e;
    acc = fmaf(e, bins[min(c0 + i, B - 1)], acc);
  }
  c.part[(kPPGroups + et.bg) * kPPHalf + et.row] = ssum;
  xch[et.bg * kPPHalf + et.row] = acc;
  epi_bar_sync();
  if (et.bg == 0) {
    float S = 0.f, Acc = 0.f;
#pragma unroll
    for (int g = 0; g < kPPGroups; ++g) { S += c.part[(kPPGroups + g) * kPPHalf + et.row]; Acc += xch[g * kPPHalf + et.row]; }
    pp_head_commit(P, c, st, tile, h * kPPHalf + et.row, symexp_f(__fdiv_rn(Acc, S)));
  }
}

// pi head (world_model.py:144-174): Npad = 128, so lanes < 64 hold logical columns [0, 64) and lanes >= 64 columns
// [64, 128).  Mean logits are columns [0, A), log-std logits columns [Apad, Apad + A): every thread drops its 16
// biased logits into an smem row, then the threads owning mean columns finish the action.
// exchange buffer of the pi head: [64 rows][128] fp32 = the 32 KiB of the staging area; element (row, col) lives at
// column col ^ (row & 31) of its row, so that the 32 rows of a warp touching one column hit 32 different banks
constexpr int kPPPiPitch = 128;
__device__ __forceinline__ int pp_pi_idx(int row, int col) { return row * kPPPiPitch + (col ^ (row & 31)); }
__device__ __forceinline__ void pp_epi_pi(const PlanParams& P, PPCtx& c, const PPThread& et, const LayerDev& ly, int h, int tile,
                                          int env, int task) {
  const float* sb = c.vec;
  float* xch = reinterpret_cast<float*>(c.base + kPPStgOff);           // [64 rows][128] fp32: see pp_pi_idx
  const int c0 = et.colhalf * 64 + et.grp * 16;
  uint32_t v[16];
  ptx::tmem_ld_32x16(et.taddr + c.tmem_base + static_cast<uint32_t>(h * 256 + et.grp * 16), v);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) xch[pp_pi_idx(et.row, c0 + i)] = fmaf(__uint_as_float(v[i]), ly.inv_scale, sb[c0 + i]);
  epi_bar_sync();
  const int a0 = c0;
  if (a0 < P.A) {
    const int r = h * kPPHalf + et.row;
    const int n = (tile % P.tiles_per_env) * kTileM + r;
    __half* xhi = plane_ptr(P, c.slot, BUF_X, 0) + static_cast<size_t>(r) * P.KpadX + P.L + P.T;
    __half* xlo = plane_ptr(P, c.slot, BUF_X, 1) + static_cast<size_t>(r) * P.KpadX + P.L + P.T;
    const float* eps = P.noise_pi + (static_cast<size_t>(env) * P.N + n) * P.A;
    const float* xr = xch;
    const bool vec = (((P.L + P.T) & 7) == 0) && (P.L + P.T + a0 + 16 <= P.KpadX);
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float act2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int a = a0 + i + u;
        act2[u] = 0.f;
        if (a < P.A) {
          act2[u] = pi_action(P, xr[pp_pi_idx(et.row, a)], xr[pp_pi_idx(et.row, P.Apad + a)],
                              P.rng_state ? noise_pi_at(P, env, n, a) : eps[a], task, a);
          if (!vec) split_store(xhi + a, xlo + a, act2[u]);
        }
      }
      __half h0, l0, h1, l1;
      split_f(act2[0], h0, l0);
      split_f(act2[1], h1, l1);
      hw[i >> 1] = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
      lw[i >> 1] = static_cast<uint32_t>(__half_as_ushort(l0)) | (static_cast<uint32_t>(__half_as_ushort(l1)) << 16);
    }
    if (vec) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __stcg(reinterpret_cast<uint4*>(xhi + a0) + i, make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]));
        __stcg(reinterpret_cast<uint4*>(xlo + a0) + i, make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kPPThreads, 1) plan_pp_kernel(const __grid_constant__ PlanParams P) {
  extern __shared__ uint8_t smem_raw[];
  PPCtx c;
  {
    uintptr_t b = (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023);
    c.base = reinterpret_cast<uint8_t*>(b);
    uint8_t* ctrl = c.base + kPPOperandBytes;
    c.a_full = reinterpret_cast<uint64_t*>(ctrl);
    c.a_empty = c.a_full + kPPARing;
    c.w_full = c.a_empty + kPPARing;
    c.w_empty = c.w_full + kPPWRing;
    c.facc = c.w_empty + kPPWRing;
    c.acc_free = c.facc + 2;
    c.act_ready = c.acc_free + 2;
    c.tmem_ptr = reinterpret_cast<uint32_t*>(c.act_ready + 2);      // 18 barriers = 144 B
    c.flags = reinterpret_cast<int*>(c.tmem_ptr + 1);
    c.G = reinterpret_cast<float*>(ctrl + 256);
    c.q1 = c.G + kTileM;
    c.vec = reinterpret_cast<float*>(ctrl + kPPCtrl);
    c.part = c.vec + 3 * kFusedMaxN;
    c.actv = c.part + 2 * kPPGroups * kPPHalf;
    c.prog = reinterpret_cast<PPStep*>(c.actv + 3 * kMaxHeadCols);
  }
  c.slot = blockIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.rank = static_cast<int>(ptx::cluster_ctarank());
  if (threadIdx.x == 0) {
    for (int s = 0; s < kPPARing; ++s) { ptx::mbar_init(&c.a_full[s], 1); ptx::mbar_init(&c.a_empty[s], 1); }
    for (int s = 0; s < kPPWRing; ++s) { ptx::mbar_init(&c.w_full[s], 1); ptx::mbar_init(&c.w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&c.facc[s], 1); ptx::mbar_init(&c.acc_free[s], 2); ptx::mbar_init(&c.act_ready[s], 1); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  ptx::cluster_sync();
  if (c.warp == 2) ptx::tmem_alloc_2sm(c.tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  c.tmem_base = *c.tmem_ptr;

  const LayerDev* LY = P.layers;
  const int nsteps = 6 * P.H + 9;
  const int tile0 = 2 * (static_cast<int>(blockIdx.x) >> 1) + c.rank;

  if (c.warp == 0) {
    // =================================================================== TMA producer (both CTAs)
    // The WHOLE warp runs the loops and one elected lane issues: in warp-convergent code the compiler keeps counters,
    // coordinates and descriptors in uniform registers, whereas inside a single-lane branch every cp.async.bulk.tensor /
    // tcgen05.mma is wrapped in a ~13-instruction "elect + R2UR + retry" waterfall (the MMA issuer then needs ~125 cycles
    // per MMA, twice what the M = 128 pair MMA takes: profiles/r02_timeline_pp.txt before / after).
    {
      uint32_t pa_it = 0, pw_it = 0, rdy_it[2] = {0, 0};
      int tcount = 0;
      for (int tile = tile0; tile < P.ntiles; tile += gridDim.x, ++tcount) {
        const PPStep* prog = c.prog + (tcount & 1) * kPPMaxSteps;
        for (int s = 0; s < nsteps; ++s) {
          for (int h = 0; h < 2; ++h) {
            int npre = 0;
#ifdef TDMPC2_PP_PREFETCH
            // Experiment (off by default: measured slower, profiles/README.md): the weights of (h, s) do not depend on the
            // epilogue that is still producing its activation planes, so while waiting for act_ready refill every weight
            // slot the MMAs release with the head of this half's weight stream (not for the first half-step of a tile,
            // whose layer program is published by that very barrier).  Non-blocking probes: try_wait may suspend.
            if (s > 0 || h > 0) {
              const PPStep st0 = prog[s];
              const LayerDev& ly0 = LY[st0.li];
              const CUtensorMap* tmW0 = &P.tmW[ly0.wmap];
              const int nnc0 = (ly0.Npad + kNch - 1) / kNch, maxpre = min(kPPWRing, (ly0.Kpad / kKch) * nnc0);
              while (true) {
                if (npre < maxpre) {
                  const uint32_t sl = pw_it % kPPWRing, ph = (pw_it / kPPWRing) & 1;
                  if (__any_sync(0xffffffffu, ptx::mbar_test_wait(&c.w_empty[sl], ph ^ 1))) {
                    const int kc = npre / nnc0, nc = npre % nnc0;
                    const int ncols = min(kNch, ly0.Npad - nc * kNch);
                    uint8_t* dst = c.base + kPPWOff + sl * kPPWSlot;
                    const int wr = ly0.wrow + nc * kNch + c.rank * (ncols / 2);
                    if (ptx::elect_one()) {
                      if (c.rank == 0) ptx::mbar_expect_tx(&c.w_full[sl], 2 * kPPWSlot);
                      ptx::tma_load_2d_2sm(tmW0, &c.w_full[sl], dst, kc * kKch, wr);
                      ptx::tma_load_2d_2sm(tmW0, &c.w_full[sl], dst + kPPWPlane, kc * kKch, wr + ly0.Npad);
                    }
                    ++pw_it; ++npre;
                    continue;
                  }
                }
                if (__any_sync(0xffffffffu, ptx::mbar_test_wait(&c.act_ready[h], rdy_it[h] & 1))) break;
              }
            } else
#endif
            ptx::mbar_wait(&c.act_ready[h], rdy_it[h] & 1);          // planes of (h, s) are published
            ++rdy_it[h];
            const PPStep st = prog[s];
            const LayerDev& ly = LY[st.li];
            const CUtensorMap* tmA = (st.src == BUF_X) ? &P.tmX64 : &P.tmH64;
            const CUtensorMap* tmW = &P.tmW[ly.wmap];
            const int nkc = ly.Kpad / kKch, nnc = (ly.Npad + kNch - 1) / kNch;
            const int arow_hi = plane_row0(P, c.slot, st.src, 0) + h * kPPHalf, arow_lo = plane_row0(P, c.slot, st.src, 1) + h * kPPHalf;
            for (int kc = 0; kc < nkc; ++kc) {
              {
                const uint32_t sl = pa_it % kPPARing, ph = (pa_it / kPPARing) & 1;
                ptx::mbar_wait(&c.a_empty[sl], ph ^ 1);
                uint8_t* dst = c.base + sl * kPPASlot;
                if (ptx::elect_one()) {
                  if (c.rank == 0) ptx::mbar_expect_tx(&c.a_full[sl], 2 * kPPASlot);
                  ptx::tma_load_2d_2sm(tmA, &c.a_full[sl], dst, kc * kKch, arow_hi);
                  ptx::tma_load_2d_2sm(tmA, &c.a_full[sl], dst + kPPAPlane, kc * kKch, arow_lo);
                }
                ++pa_it;
              }
              for (int nc = 0; nc < nnc; ++nc) {
                if (kc * nnc + nc < npre) continue;                  // already requested while waiting for act_ready
                const int ncols = min(kNch, ly.Npad - nc * kNch);
                const uint32_t sl = pw_it % kPPWRing, ph = (pw_it / kPPWRing) & 1;
                ptx::mbar_wait(&c.w_empty[sl], ph ^ 1);
                uint8_t* dst = c.base + kPPWOff + sl * kPPWSlot;
                const int wr = ly.wrow + nc * kNch + c.rank * (ncols / 2);
                if (ptx::elect_one()) {
                  if (c.rank == 0) ptx::mbar_expect_tx(&c.w_full[sl], 2 * kPPWSlot);
                  ptx::tma_load_2d_2sm(tmW, &c.w_full[sl], dst, kc * kKch, wr);
                  ptx::tma_load_2d_2sm(tmW, &c.w_full[sl], dst + kPPWPlane, kc * kKch, wr + ly.Npad);
                }
                ++pw_it;
              }
            }
          }
        }
      }
    }
  } else if (c.warp == 1) {
    // =================================================================== MMA issuer (leader CTA only; whole warp, elected issue)
    if (c.rank == 0) {
      uint32_t ma_it = 0, mw_it = 0, rdy_it[2] = {0, 0}, free_it[2] = {0, 0};
      const uint32_t sbase = ptx::smem_u32(c.base);
      int tcount = 0;
      for (int tile = tile0; tile < P.ntiles; tile += gridDim.x, ++tcount) {
        const PPStep* prog = c.prog + (tcount & 1) * kPPMaxSteps;
        for (int s = 0; s < nsteps; ++s) {
          for (int h = 0; h < 2; ++h) {
            ptx::mbar_wait(&c.act_ready[h], rdy_it[h] & 1);          // also orders the read of prog[]
            ++rdy_it[h];
            const PPStep st = prog[s];
            const LayerDev& ly = LY[st.li];
            const int nkc = ly.Kpad / kKch, nnc = (ly.Npad + kNch - 1) / kNch;
            ptx::mbar_wait(&c.acc_free[h], (free_it[h] & 1) ^ 1);    // both CTAs drained the previous accumulator of h
            ++free_it[h];
            ptx::tc_fence_after();
            PP_TRACE(P, blockIdx.x == 0 && tcount == 0, s, 1 + 2 * h);
            for (int kc = 0; kc < nkc; ++kc) {
              const uint32_t as = ma_it % kPPARing, aph = (ma_it / kPPARing) & 1;
              ptx::mbar_wait(&c.a_full[as], aph);
              for (int nc = 0; nc < nnc; ++nc) {
                const int ncols = min(kNch, ly.Npad - nc * kNch);
                const uint32_t ws = mw_it % kPPWRing, wph = (mw_it / kPPWRing) & 1;
                ptx::mbar_wait(&c.w_full[ws], wph);
                ptx::tc_fence_after();
                const uint32_t d = c.tmem_base + static_cast<uint32_t>(h * 256 + nc * 128);
                const uint32_t sa = sbase + as * kPPASlot, sw = sbase + kPPWOff + ws * kPPWSlot;
                const uint32_t idesc = ptx::make_idesc_f16(2 * kPPHalf, ncols);
                if (ptx::elect_one()) {
#pragma unroll
                  for (int ks = 0; ks < kKch / 16; ++ks) {
                    const uint64_t a_hi = ptx::make_sw128_kmajor_desc(sa + ks * 32);
                    const uint64_t a_lo = ptx::make_sw128_kmajor_desc(sa + kPPAPlane + ks * 32);
                    const uint64_t w_hi = ptx::make_sw128_kmajor_desc(sw + ks * 32);
                    const uint64_t w_lo = ptx::make_sw128_kmajor_desc(sw + kPPWPlane + ks * 32);
#ifndef TDMPC2_EXP_NOMMA   // (measurement build: see mma_stage in plan_kernels.cuh)
                    ptx::umma_f16_2sm(d, a_lo, w_hi, idesc, !(kc == 0 && ks == 0));
                    ptx::umma_f16_2sm(d, a_hi, w_lo, idesc, 1);
                    ptx::umma_f16_2sm(d, a_hi, w_hi, idesc, 1);
#else
                    (void)a_hi; (void)a_lo; (void)w_hi; (void)w_lo; (void)d; (void)idesc;
#endif
                  }
                  ptx::umma_commit_2sm(&c.w_empty[ws]);
                }
                ++mw_it;
              }
              if (ptx::elect_one()) ptx::umma_commit_2sm(&c.a_empty[as]);
              ++ma_it;
            }
            if (ptx::elect_one()) ptx::umma_commit_2sm(&c.facc[h]);
            PP_TRACE(P, blockIdx.x == 0 && tcount == 0, s, 2 + 2 * h);
          }
        }
      }
    }
  } else if (c.warp >= kPPEpiWarp0) {
    // =================================================================== epilogue + glue (16 warps, both CTAs)
    PPThread et;
    {
      const int e = c.warp - kPPEpiWarp0;
      et.q = e & 3; et.grp = e >> 2; et.row = (et.q & 1) * 32 + c.lane; et.colhalf = et.q >> 1;
      et.bg = et.colhalf * 4 + et.grp;
      et.taddr = static_cast<uint32_t>(et.q * 32) << 16;
    }
    const int tid = threadIdx.x - kPPEpiWarp0 * 32;
    uint32_t fph[2] = {0, 0};
    int tcount = 0;
    for (int tile = tile0; tile < P.ntiles; tile += gridDim.x, ++tcount) {
      PPStep* prog = c.prog + (tcount & 1) * kPPMaxSteps;
      const int env = tile / P.tiles_per_env;
      const int task = P.task ? P.task[env] : 0;
      // ---------------- tile set-up: layer program, value accumulators, X planes [z | emb | a_0]
      if (tid < nsteps) {
        const float* dpow = P.disc_pow + static_cast<size_t>(task) * (P.H + 1);
        const int* qi = P.qidx + static_cast<size_t>(env) * 2;
        PPStep st;
        int mlp, l, t = 0;                       // mlp: 0 reward, 1 dynamics, 2 pi, 3 q_a, 4 q_b
        if (tid < 6 * P.H) { t = tid / 6; l = tid % 6; mlp = l < 3 ? 0 : 1; l %= 3; }
        else { const int u = tid - 6 * P.H; mlp = 2 + u / 3; l = u % 3; }
        const int base = mlp == 0 ? P.li_rew : mlp == 1 ? P.li_dyn : mlp == 2 ? P.li_pi : P.li_q + 3 * qi[mlp - 3];
        st.li = base + l;
        st.src = l == 0 ? BUF_X : BUF_H1;
        st.t_act = (mlp == 0 && l == 0) ? t : -1;
        st.head = 0; st.disc = 0.f; st.dstbuf = -1; st.pad = 0;
        if (l < 2) { st.kind = EPI_LN_MISH; st.dstbuf = BUF_H1; }
        else if (mlp == 0) { st.kind = EPI_TWOHOT; st.head = HEAD_REWARD; st.disc = dpow[t]; }
        else if (mlp == 1) { st.kind = EPI_LN_SIMNORM; st.dstbuf = BUF_X; }
        else if (mlp == 2) { st.kind = EPI_PI; }
        else { st.kind = EPI_TWOHOT; st.head = (mlp == 3) ? HEAD_Q1 : HEAD_Q2; st.disc = dpow[P.H]; }
        prog[tid] = st;
      }
      for (int r = tid; r < kTileM; r += kEpiThreads) { c.G[r] = 0.f; c.q1[r] = 0.f; }
      if (tid <= P.H && !P.rng_state) {
        // this tile's slabs of the read-once noise tensors: ask for them now (L2) instead of paying DRAM latency per step
        const int n0 = (tile % P.tiles_per_env) * kTileM, n1 = n0 + kTileM;
        if (tid < P.H) {
          const int r0 = max(n0, P.P) - P.P, r1 = n1 - P.P;
          if (r1 > r0)
            ptx::bulk_prefetch_l2(P.noise_r + ((static_cast<size_t>(env) * P.H + tid) * (P.N - P.P) + r0) * P.A,
                                  static_cast<size_t>(r1 - r0) * P.A * sizeof(float));
        } else {
          ptx::bulk_prefetch_l2(P.noise_pi + (static_cast<size_t>(env) * P.N + n0) * P.A, static_cast<size_t>(kTileM) * P.A * sizeof(float));
        }
      }
      {
        __half* xhi = plane_ptr(P, c.slot, BUF_X, 0);
        __half* xlo = plane_ptr(P, c.slot, BUF_X, 1);
        const int nch = (P.L + P.T) / 8;
        const int ngrp = kEpiThreads / nch;
        const int ch = tid % nch, rg = tid / nch;
        if (rg < ngrp) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int col = ch * 8 + 2 * j + u;
              x[u] = col < P.L ? P.z[static_cast<size_t>(env) * P.L + col] : P.emb[static_cast<size_t>(task) * P.T + (col - P.L)];
              x[u] = (fabsf(x[u]) <= 3.0e38f) ? fminf(fmaxf(x[u], -65000.f), 65000.f) : CUDART_NAN_F;
            }
            const __half2 h2 = __floats2half2_rn(x[0], x[1]);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(x[0] - hf.x, x[1] - hf.y);
            hw[j] = *reinterpret_cast<const uint32_t*>(&h2);
            lw[j] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          for (int r = rg; r < kTileM; r += ngrp) {
            __stcg(reinterpret_cast<uint4*>(xhi + static_cast<size_t>(r) * P.KpadX + ch * 8), make_uint4(hw[0], hw[1], hw[2], hw[3]));
            __stcg(reinterpret_cast<uint4*>(xlo + static_cast<size_t>(r) * P.KpadX + ch * 8), make_uint4(lw[0], lw[1], lw[2], lw[3]));
          }
        }
      }
      pp_write_actions(P, c, tile, env, task, 0, 0, kTileM, true);
      __threadfence();
      ptx::fence_proxy_async_all();
      epi_bar_sync();
      if (tid == 0) { ptx::mbar_arrive(&c.act_ready[0]); ptx::mbar_arrive(&c.act_ready[1]); }

      // ---------------- the layer program, two halves per step
      for (int s = 0; s < nsteps; ++s) {
        const PPStep st = prog[s];
        const LayerDev& ly = LY[st.li];
        const bool is_ln = (st.kind == EPI_LN_MISH || st.kind == EPI_LN_SIMNORM);
        // stage this layer's vectors (both halves use them)
        for (int i = tid; i < ly.Npad; i += kEpiThreads) {
          c.vec[i] = ly.bias[i];
          if (is_ln) { c.vec[kFusedMaxN + i] = ly.ln_g[i]; c.vec[2 * kFusedMaxN + i] = ly.ln_b[i]; }
        }
        if (st.kind == EPI_TWOHOT) for (int i = tid; i < P.B; i += kEpiThreads) c.vec[kFusedMaxN + i] = P.bins[i];
        const int t_next = (s + 1 < nsteps) ? prog[s + 1].t_act : -1;
        epi_bar_sync();
        const bool tr_on = (blockIdx.x == 0 && tcount == 0 && tid == 0);
        PP_TRACE(P, tr_on, s, 0);
        for (int h = 0; h < 2; ++h) {
          ptx::mbar_wait(&c.facc[h], fph[h]);
          fph[h] ^= 1;
          ptx::tc_fence_after();
          PP_TRACE(P, tr_on, s, 5 + 3 * h);
          if (st.kind == EPI_LN_MISH) pp_epi_ln<EPI_LN_MISH>(P, c, et, ly, st, h);
          else if (st.kind == EPI_LN_SIMNORM) pp_epi_ln<EPI_LN_SIMNORM>(P, c, et, ly, st, h);
          else if (st.kind == EPI_TWOHOT) pp_epi_twohot(P, c, et, ly, st, h, tile);
          else pp_epi_pi(P, c, et, ly, h, tile, env, task);
          ptx::tc_fence_before();
          PP_TRACE(P, tr_on, s, 6 + 3 * h);
          if (t_next >= 0) pp_write_actions(P, c, tile, env, task, t_next, h * kPPHalf, kPPHalf, h == 0);
          if (is_ln && ((et.q & 1) == 0)) {
            __syncwarp();
            if (ptx::elect_one()) ptx::bulk_wait<0>();                 // this block group's stores are performed
          }
          if (!is_ln || t_next >= 0) {
            // heads and the action pass wrote global memory through the generic proxy; LayerNorm outputs left through
            // TMA stores only (the leaders have waited for them), for which the barrier below is all the next loads need
            __threadfence();
            ptx::fence_proxy_async_all();
          }
          epi_bar_sync();
          if (tid == 0) {
            pp_arrive_leader(&c.acc_free[h], c.rank);                 // TMEM of half h may be overwritten
            if (s + 1 < nsteps) ptx::mbar_arrive(&c.act_ready[h]);    // planes of (h, s+1) are published
          }
          PP_TRACE(P, tr_on, s, 7 + 3 * h);
        }
      }

      // ---------------- last CTA of this environment: top-k, MPPI weights, mean/std refit (tdmpc2.py:184-197)
      __threadfence();
      epi_bar_sync();
      if (tid == 0) {
        const unsigned old = atomicAdd(&P.env_counter[env], 1u);
        c.flags[0] = (old == static_cast<unsigned>(P.tiles_per_env - 1));
        if (c.flags[0]) P.env_counter[env] = 0;
      }
      epi_bar_sync();
      if (c.flags[0]) {
        __threadfence();
        refit_env<GroupEpiPP>(P, c.base, static_cast<size_t>(kPPOperandBytes), env, task);
      }
      ptx::fence_proxy_async_all();         // refit wrote ring / staging smem through the generic proxy
      epi_bar_sync();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (c.warp == 2) ptx::tmem_dealloc_2sm(c.tmem_base, 512);
}

}  // namespace tdmpc2
