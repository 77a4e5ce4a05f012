// Fused planning kernels for sm_100a (B200).
//
// One persistent kernel template, `plan_kernel<ENGINE>`, runs the MLP chains of
// the TD-MPC2 planner on 128-row tiles.  Three modes share all device code:
//
//   MODE_ENCODE : rows = environments.      z = encode(obs, task)        (reference world_model.py:103-112)
//   MODE_PRIOR  : rows = (env, pi-traj).    the P policy-prior rollouts  (reference tdmpc2.py:154-160)
//   MODE_ITER   : rows = (env, sample).     ONE CEM iteration:           (reference tdmpc2.py:173-197)
//                   sample actions -> H x (reward, dynamics) -> terminal pi -> 2 Q heads
//                   -> value -> [last tile of an env] top-k, MPPI weights, mean/std refit.
//   MODE_VALUE  : MODE_ITER's rollout on caller-given z / actions, no refit (reference tdmpc2.py:122-136)
//   MODE_LAYER  : one layer, for diagnostics.
//
// Every dense layer is `raw = A[128, Kpad] * W[Npad, Kpad]^T` with both operands
// stored as two fp16 planes (hi, lo; x ~= hi + lo to ~22 bits).  The tcgen05
// engine accumulates A_lo*W_hi + A_hi*W_lo + A_hi*W_hi in fp32 in TMEM
// (kind::f16, M=128, N<=256, K=16), operands staged by TMA (128-byte swizzle)
// through a 2-stage mbarrier pipeline; warp 0 = TMA producer, warp 1 = MMA
// issuer, warps 4-7 drain TMEM.  The row phase (all warps, one warp per row)
// then applies bias, LayerNorm, Mish / SimNorm / two-hot-inverse / tanh-Gaussian
// sampling and writes the next layer's fp16 planes.  Activations live in a
// per-CTA scratch slot that stays L2-resident.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "ptx.cuh"

namespace tdmpc2 {

constexpr int kTileM = 128;       // rows per tile (UMMA M)
constexpr int kKch = 64;          // K elements per pipeline stage (128 B of fp16: one swizzle row)
constexpr int kNch = 256;         // N columns per accumulator chunk (UMMA N max)
constexpr int kStages = 2;
constexpr int kAPlane = kTileM * 128;           // 16 KiB: one A plane of a stage
constexpr int kWPlane = kNch * 128;             // 32 KiB: one W plane of a stage
constexpr int kStageBytes = 2 * kAPlane + 2 * kWPlane;   // 96 KiB
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxWMaps = 8;
constexpr int kMaxHeadCols = 256;  // widest head output (2A or num_bins), padded to 128s
constexpr int kSmemCtrl = 2048;    // barriers, tmem ptr, flags (128 B) + G[128] + q1[128]
constexpr int kSmemRowBuf = kWarps * kMaxHeadCols * 4;  // 8 KiB: one head-output row per warp
constexpr int kSmemRowEnv = kTileM * 4;                 // env index of each tile row
constexpr int kSmemBytes = kStages * kStageBytes + kSmemCtrl + kSmemRowBuf + kSmemRowEnv + 1024 /*align slack*/;

enum Mode { MODE_ENCODE = 0, MODE_PRIOR = 1, MODE_ITER = 2, MODE_VALUE = 3, MODE_LAYER = 4 };
enum Engine { ENGINE_TC = 0, ENGINE_SIMT = 1 };
enum Buf { BUF_X = 0, BUF_H1 = 1, BUF_H2 = 2 };

struct LayerDev {
  int K, Kpad, N, Npad;
  int wmap;          // which weight tensor map (Kpad class)
  int wrow;          // row of the hi plane inside that class tensor; lo plane at wrow + Npad
  int has_ln;
  float inv_scale;   // 2^-k; the packed planes hold W * 2^k (written by the pack kernel)
  const float* bias; // [Npad] zero padded
  const float* ln_g; // [Npad]
  const float* ln_b; // [Npad]
  const __half* w_hi;  // direct pointers (SIMT engine): [Npad][Kpad]
  const __half* w_lo;
};

struct PlanParams {
  CUtensorMap tmX;                 // [slots*2*128, KpadX] fp16, box 64 x 128
  CUtensorMap tmH;                 // [slots*4*128, KpadH]
  CUtensorMap tmW[kMaxWMaps];      // weights, one map per Kpad class, box 64 x 128
  const LayerDev* layers;
  int E, N, P, Ppad, K, H, obs_dim, A, L, M, T, B, num_q, simnorm, num_enc;
  int tiles_per_env, ntiles, KpadX, KpadH, NpadMax;
  int li_enc, li_dyn, li_rew, li_pi, li_q;
  float temperature, min_std, max_std, log_std_min, log_std_dif;
  __half* X; __half* Hb; float* raw;       // scratch, indexed by slot = blockIdx.x
  const float* emb; const float* masks; const float* disc_pow; const float* bins;
  int mode;
  // per-call inputs
  const float* obs; const int* task; const float* noise_prior; const float* noise_r; const float* noise_pi;
  const int* qidx; const float* z_rows; const float* actions_explicit;
  // planner state
  float* z; float* pi_actions; float* mean; float* std; float* values; unsigned* env_counter;
  float* score; float* elite_act0; int* elite_idx32;
  long long* elite_idx_out; float* values_out;
  // MODE_LAYER
  int dbg_layer, dbg_mode, dbg_rows; const float* dbg_x; float* dbg_y;
};

// ------------------------------------------------------------------------------------ small math
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Mish(x) = x * tanh(softplus(x)) (layers.py:103; softplus threshold 20).
// tanh(log(1+e^x)) = n / (n + 2) with n = e^x (e^x + 2): one exp, one divide, no cancellation.
__device__ __forceinline__ float mish_f(float x) {
  if (x > 20.f) return x;
  const float e = expf(x);
  const float n = e * (e + 2.f);
  return x * __fdiv_rn(n, n + 2.f);
}
__device__ __forceinline__ float symexp_f(float x) {   // math.py:50-55
  const float m = expf(fabsf(x)) - 1.f;
  return x > 0.f ? m : (x < 0.f ? -m : 0.f);
}
__device__ __forceinline__ void split_store(__half* hi, __half* lo, float x) {
  x = fminf(fmaxf(x, -65000.f), 65000.f);
  const __half h = __float2half_rn(x);
  *hi = h;
  *lo = __float2half_rn(x - __half2float(h));
}
__device__ __forceinline__ float nan_to_num0(float v) {   // torch.nan_to_num(0): nan->0, +-inf -> +-FLT_MAX
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}

// ------------------------------------------------------------------------------------ CTA context
struct Ctx {
  uint8_t* stage_base;      // kStages * kStageBytes, 1024-aligned
  uint64_t* full;           // [kStages]
  uint64_t* empty;          // [kStages]
  uint64_t* acc_full;       // [2]
  uint64_t* acc_empty;      // [2]
  uint32_t* tmem_ptr;
  float* rowbuf;            // [kWarps][kMaxHeadCols]
  float* G;                 // [128] discounted reward sum
  float* q1;                // [128] first Q head
  int* flags;               // small ints
  uint32_t tmem_base;
  int slot, warp, lane;
  // pipeline counters (each role keeps its own; persist across layers / tiles)
  uint32_t p_it, m_it, a_it, d_it;
};

__device__ __forceinline__ __half* plane_ptr(const PlanParams& P, int slot, int buf, int plane) {
  if (buf == BUF_X) return P.X + (static_cast<size_t>(slot) * 2 + plane) * kTileM * P.KpadX;
  return P.Hb + ((static_cast<size_t>(slot) * 2 + (buf - BUF_H1)) * 2 + plane) * kTileM * P.KpadH;
}
__device__ __forceinline__ int plane_pitch(const PlanParams& P, int buf) { return buf == BUF_X ? P.KpadX : P.KpadH; }
__device__ __forceinline__ int plane_row0(const PlanParams& P, int slot, int buf, int plane) {  // TMA row coord
  if (buf == BUF_X) return (slot * 2 + plane) * kTileM;
  return ((slot * 2 + (buf - BUF_H1)) * 2 + plane) * kTileM;
}
__device__ __forceinline__ float* raw_ptr(const PlanParams& P, int slot) {
  return P.raw + static_cast<size_t>(slot) * kTileM * P.NpadMax;
}

// ------------------------------------------------------------------------------------ GEMM, tcgen05 engine
// raw[128, Npad] = (A_hi + A_lo)[128, Kpad] * (W_hi + W_lo)[Npad, Kpad]^T   (scaled by 2^k; row phase unscales)
__device__ void gemm_tc(const PlanParams& P, Ctx& c, const LayerDev& ly, int srcbuf) {
  const int nkc = ly.Kpad / kKch;
  const int nnc = (ly.Npad + kNch - 1) / kNch;
  if (c.warp == 0) {
    // ===== TMA producer =====
    if (c.lane == 0) {
      const CUtensorMap* tmA = (srcbuf == BUF_X) ? &P.tmX : &P.tmH;
      const CUtensorMap* tmW = &P.tmW[ly.wmap];
      const int arow_hi = plane_row0(P, c.slot, srcbuf, 0), arow_lo = plane_row0(P, c.slot, srcbuf, 1);
      for (int nc = 0; nc < nnc; ++nc) {
        const int ncols = min(kNch, ly.Npad - nc * kNch);   // 128 or 256
        for (int kc = 0; kc < nkc; ++kc) {
          const uint32_t s = c.p_it % kStages, ph = (c.p_it / kStages) & 1;
          ptx::mbar_wait(&c.empty[s], ph ^ 1);
          uint8_t* st = c.stage_base + s * kStageBytes;
          ptx::mbar_expect_tx(&c.full[s], 2 * kAPlane + 2 * ncols * 128);
          ptx::tma_load_2d(tmA, &c.full[s], st, kc * kKch, arow_hi);
          ptx::tma_load_2d(tmA, &c.full[s], st + kAPlane, kc * kKch, arow_lo);
          for (int b = 0; b < ncols / 128; ++b) {
            const int wr = ly.wrow + nc * kNch + b * 128;
            ptx::tma_load_2d(tmW, &c.full[s], st + 2 * kAPlane + b * (128 * 128), kc * kKch, wr);
            ptx::tma_load_2d(tmW, &c.full[s], st + 2 * kAPlane + kWPlane + b * (128 * 128), kc * kKch, wr + ly.Npad);
          }
          ++c.p_it;
        }
      }
    }
  } else if (c.warp == 1) {
    // ===== MMA issuer =====
    if (c.lane == 0) {
      for (int nc = 0; nc < nnc; ++nc) {
        const int ncols = min(kNch, ly.Npad - nc * kNch);
        const uint32_t idesc = ptx::make_idesc_f16(kTileM, ncols);
        const uint32_t slot = c.a_it & 1, aph = (c.a_it >> 1) & 1;
        ptx::mbar_wait(&c.acc_empty[slot], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d = c.tmem_base + slot * kNch;
        for (int kc = 0; kc < nkc; ++kc) {
          const uint32_t s = c.m_it % kStages, ph = (c.m_it / kStages) & 1;
          ptx::mbar_wait(&c.full[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(c.stage_base + s * kStageBytes);
#pragma unroll
          for (int ks = 0; ks < kKch / 16; ++ks) {
            const uint64_t a_hi = ptx::make_sw128_kmajor_desc(sa + ks * 32);
            const uint64_t a_lo = ptx::make_sw128_kmajor_desc(sa + kAPlane + ks * 32);
            const uint64_t w_hi = ptx::make_sw128_kmajor_desc(sa + 2 * kAPlane + ks * 32);
            const uint64_t w_lo = ptx::make_sw128_kmajor_desc(sa + 2 * kAPlane + kWPlane + ks * 32);
            ptx::umma_f16(d, a_lo, w_hi, idesc, (kc | ks) != 0);   // small terms first
            ptx::umma_f16(d, a_hi, w_lo, idesc, 1);
            ptx::umma_f16(d, a_hi, w_hi, idesc, 1);
          }
          ptx::umma_commit(&c.empty[s]);    // frees the smem stage when these MMAs retire
          ++c.m_it;
        }
        ptx::umma_commit(&c.acc_full[slot]);
        ++c.a_it;
      }
    }
  } else if (c.warp >= 4) {
    // ===== TMEM drain: accumulator chunk -> raw scratch (fp32) =====
    const int q = c.warp & 3;                  // TMEM lane quarter this warp may touch
    const int row = q * 32 + c.lane;
    float* rawrow = raw_ptr(P, c.slot) + static_cast<size_t>(row) * P.NpadMax;
    for (int nc = 0; nc < nnc; ++nc) {
      const int ncols = min(kNch, ly.Npad - nc * kNch);
      const uint32_t slot = c.d_it & 1, dph = (c.d_it >> 1) & 1;
      ptx::mbar_wait(&c.acc_full[slot], dph);
      ptx::tc_fence_after();
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(c.tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * kNch + c0, v);
        ptx::tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(rawrow + nc * kNch + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          __stcg(dst + i, make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                      __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&c.acc_empty[slot]);
      ++c.d_it;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------ GEMM, SIMT engine
// Same operands, plain fp32 FFMA on CUDA cores (exact products of the split operands).
__device__ void gemm_simt(const PlanParams& P, Ctx& c, const LayerDev& ly, int srcbuf) {
  constexpr int BN = 64, BK = 32;
  float* sA = reinterpret_cast<float*>(c.stage_base);          // [BK][128+4]
  float* sW = sA + BK * (kTileM + 4);                          // [BK][BN+4]
  const __half* a_hi = plane_ptr(P, c.slot, srcbuf, 0);
  const __half* a_lo = plane_ptr(P, c.slot, srcbuf, 1);
  const int pitch = plane_pitch(P, srcbuf);
  const int tid = threadIdx.x;
  const int tr = (tid / 16) * 8, tc = (tid % 16) * 4;          // 8 rows x 4 cols per thread
  float* rawbase = raw_ptr(P, c.slot);
  for (int n0 = 0; n0 < ly.Npad; n0 += BN) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < ly.Kpad; k0 += BK) {
      for (int i = tid; i < kTileM * BK; i += kThreads) {
        const int r = i / BK, k = i % BK;
        const size_t o = static_cast<size_t>(r) * pitch + k0 + k;
        sA[k * (kTileM + 4) + r] = __half2float(__ldcg(a_hi + o)) + __half2float(__ldcg(a_lo + o));
      }
      for (int i = tid; i < BN * BK; i += kThreads) {
        const int n = i / BK, k = i % BK;
        const size_t o = static_cast<size_t>(n0 + n) * ly.Kpad + k0 + k;
        sW[k * (BN + 4) + n] = __half2float(ly.w_hi[o]) + __half2float(ly.w_lo[o]);
      }
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < BK; ++k) {
        float a[8], w[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = sA[k * (kTileM + 4) + tr + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = sW[k * (BN + 4) + tc + j];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        __stcg(rawbase + static_cast<size_t>(tr + i) * P.NpadMax + n0 + tc + j, acc[i][j]);
  }
  __syncthreads();
}

template <int ENGINE>
__device__ __forceinline__ void gemm(const PlanParams& P, Ctx& c, const LayerDev& ly, int srcbuf) {
  if (ENGINE == ENGINE_TC) gemm_tc(P, c, ly, srcbuf);
  else gemm_simt(P, c, ly, srcbuf);
}

// Make generic-proxy global writes (activation planes) visible to the TMA unit
// (async proxy) before the next layer's loads, and sync the CTA.
__device__ __forceinline__ void publish_planes() {
  __threadfence();
  ptx::fence_proxy_async_all();
  __syncthreads();
}

// ------------------------------------------------------------------------------------ row phases
enum Act { ACT_MISH = 0, ACT_SIMNORM = 1 };

// LayerNorm (+ Mish | SimNorm) over raw rows; one warp per row, lane-strided columns.
// Writes fp16 planes into dstbuf columns [dst_col0, dst_col0 + N) and/or fp32 rows to out_f32.
// REGS: the row (Npad <= 512) is held in 16 registers per lane; otherwise re-read from L2.
__device__ __forceinline__ float ln_act_one(float y, bool valid, int act) {
  if (act == ACT_MISH) return mish_f(y);
  // SimNorm (layers.py:74-88): softmax over groups of 8 consecutive columns = 8 adjacent lanes.
  float m = valid ? y : -CUDART_INF_F;
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
  const float e = valid ? expf(y - m) : 0.f;
  float t = e;
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 4);
  return valid ? __fdiv_rn(e, t) : 0.f;
}

template <bool REGS>
__device__ void rows_ln_act_impl(const PlanParams& P, Ctx& c, const LayerDev& ly, int act, int dstbuf, int dst_col0,
                                 float* out_f32, int out_pitch, const int* out_rowmap) {
  const float* rawbase = raw_ptr(P, c.slot);
  const int N = ly.N;
  const float invN = 1.f / static_cast<float>(N);
  __half* dhi = dstbuf >= 0 ? plane_ptr(P, c.slot, dstbuf, 0) : nullptr;
  __half* dlo = dstbuf >= 0 ? plane_ptr(P, c.slot, dstbuf, 1) : nullptr;
  const int pitch = dstbuf >= 0 ? plane_pitch(P, dstbuf) : 0;
  const float inv_scale = ly.inv_scale;
  const int ncolj = (N + 31) / 32;
  for (int r = c.warp; r < kTileM; r += kWarps) {
    const float* rr = rawbase + static_cast<size_t>(r) * P.NpadMax;
    float v[16];
    float s = 0.f;
    if (REGS) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c.lane + 32 * j;
        v[j] = 0.f;
        if (col < N) { v[j] = fmaf(__ldcg(rr + col), inv_scale, ly.bias[col]); s += v[j]; }
      }
    } else {
      for (int col = c.lane; col < N; col += 32) s += fmaf(__ldcg(rr + col), inv_scale, ly.bias[col]);
    }
    const float mean = warp_sum(s) * invN;
    float sq = 0.f;
    if (REGS) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c.lane + 32 * j;
        if (col < N) { const float d = v[j] - mean; sq = fmaf(d, d, sq); }
      }
    } else {
      for (int col = c.lane; col < N; col += 32) {
        const float d = fmaf(__ldcg(rr + col), inv_scale, ly.bias[col]) - mean;
        sq = fmaf(d, d, sq);
      }
    }
    const float var = warp_sum(sq) * invN;
    const float rstd = 1.f / sqrtf(var + 1e-5f);   // nn.LayerNorm eps (layers.py:101)
    const int orow = out_rowmap ? out_rowmap[r] : r;
    auto emit = [&](int col, float x) {
      const bool valid = col < N;
      float y = valid ? (x - mean) * rstd * ly.ln_g[col] + ly.ln_b[col] : 0.f;
      y = ln_act_one(y, valid, act);
      if (valid) {
        if (dhi) split_store(dhi + static_cast<size_t>(r) * pitch + dst_col0 + col,
                             dlo + static_cast<size_t>(r) * pitch + dst_col0 + col, y);
        if (out_f32 && orow >= 0) out_f32[static_cast<size_t>(orow) * out_pitch + col] = y;
      }
    };
    if (REGS) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < ncolj) emit(c.lane + 32 * j, v[j]);
    } else {
      for (int j = 0; j < ncolj; ++j) {
        const int col = c.lane + 32 * j;
        emit(col, col < N ? fmaf(__ldcg(rr + col), inv_scale, ly.bias[col]) : 0.f);
      }
    }
  }
}
__device__ __forceinline__ void rows_ln_act(const PlanParams& P, Ctx& c, const LayerDev& ly, int act, int dstbuf,
                                            int dst_col0, float* out_f32, int out_pitch, const int* out_rowmap) {
  if (ly.Npad <= 512) rows_ln_act_impl<true>(P, c, ly, act, dstbuf, dst_col0, out_f32, out_pitch, out_rowmap);
  else rows_ln_act_impl<false>(P, c, ly, act, dstbuf, dst_col0, out_f32, out_pitch, out_rowmap);
}

// Head output row -> smem row buffer: out[col] = raw*inv_scale + bias (plain Linear, no LN).
__device__ __forceinline__ void head_row_to_smem(const PlanParams& P, Ctx& c, const LayerDev& ly, int r, float* buf) {
  const float* rr = raw_ptr(P, c.slot) + static_cast<size_t>(r) * P.NpadMax;
  for (int col = c.lane; col < ly.N; col += 32) buf[col] = fmaf(__ldcg(rr + col), ly.inv_scale, ly.bias[col]);
  __syncwarp();
}

// two_hot_inv (math.py:74-83): softmax over the bins, expectation under linspace(vmin,vmax,B), symexp.
__device__ __forceinline__ float two_hot_inv_row(const PlanParams& P, Ctx& c, const float* buf) {
  float m = -CUDART_INF_F;
  for (int col = c.lane; col < P.B; col += 32) m = fmaxf(m, buf[col]);
  m = warp_max(m);
  float s = 0.f;
  for (int col = c.lane; col < P.B; col += 32) s += expf(buf[col] - m);
  s = warp_sum(s);
  float acc = 0.f;
  for (int col = c.lane; col < P.B; col += 32) acc = fmaf(__fdiv_rn(expf(buf[col] - m), s), P.bins[col], acc);
  acc = warp_sum(acc);
  return symexp_f(acc);
}

// ------------------------------------------------------------------------------------ row -> (env, sample) mapping
struct RowMap {
  int env;     // -1 if the row is padding
  int idx;     // sample index n (ITER/VALUE), pi-trajectory p (PRIOR), unused (ENCODE)
};
__device__ __forceinline__ RowMap map_row(const PlanParams& P, int tile, int r) {
  RowMap m;
  if (P.mode == MODE_ENCODE) {
    m.env = tile * kTileM + r; m.idx = 0;
    if (m.env >= P.E) m.env = -1;
  } else if (P.mode == MODE_PRIOR) {
    const int per = kTileM / P.Ppad;
    m.env = tile * per + r / P.Ppad; m.idx = r % P.Ppad;
    if (m.env >= P.E || m.idx >= P.P) m.env = -1;
  } else {
    m.env = tile / P.tiles_per_env; m.idx = (tile % P.tiles_per_env) * kTileM + r;
    if (m.idx >= P.N) m.env = -1;
  }
  return m;
}

// action a_t of sample n of env e at CEM time (tdmpc2.py:168-181)
__device__ __forceinline__ float sample_action(const PlanParams& P, int e, int t, int n, int a, int task) {
  float v;
  if (P.actions_explicit) {
    return P.actions_explicit[((static_cast<size_t>(e) * P.H + t) * P.N + n) * P.A + a];
  } else if (n < P.P) {
    v = P.pi_actions[((static_cast<size_t>(e) * P.H + t) * P.P + n) * P.A + a];
  } else {
    const size_t sa = (static_cast<size_t>(e) * P.H + t) * P.A + a;
    const float r = P.noise_r[((static_cast<size_t>(e) * P.H + t) * (P.N - P.P) + (n - P.P)) * P.A + a];
    v = __fadd_rn(P.mean[sa], __fmul_rn(P.std[sa], r));       // mean + std * r, two roundings like eager torch
    v = fminf(fmaxf(v, -1.f), 1.f);
  }
  if (P.masks) v *= P.masks[static_cast<size_t>(task) * P.A + a];
  return v;
}

// ------------------------------------------------------------------------------------ top-k + MPPI refit
// Runs in the LAST CTA to finish a tile of environment e (tdmpc2.py:184-197).
__device__ void refit_env(const PlanParams& P, Ctx& c, int e, int task) {
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(c.stage_base);
  int nsort = 1;
  while (nsort < P.N) nsort <<= 1;
  const float* vals = P.values + static_cast<size_t>(e) * P.N;
  for (int i = threadIdx.x; i < nsort; i += kThreads) {
    unsigned long long k = 0ull;   // sentinel: sorts last
    if (i < P.N) {
      const float v = __ldcg(vals + i);
      unsigned u = __float_as_uint(v);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      k = (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned>(i));
      if (P.values_out) P.values_out[static_cast<size_t>(e) * P.N + i] = v;
    }
    keys[i] = k;
  }
  __syncthreads();
  // bitonic sort, descending (value desc, index asc on ties)
  for (int k2 = 2; k2 <= nsort; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < nsort; i += kThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool desc = ((i & k2) == 0);
          if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  float* escore = reinterpret_cast<float*>(keys + nsort);     // [K]
  int* eidx = reinterpret_cast<int*>(escore + P.K);           // [K]
  float* red = reinterpret_cast<float*>(eidx + P.K);          // [2]
  const float vmax = __ldcg(vals + (0xFFFFFFFFu - static_cast<unsigned>(keys[0] & 0xFFFFFFFFull)));
  for (int k = threadIdx.x; k < P.K; k += kThreads) {
    const int idx = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(keys[k] & 0xFFFFFFFFull));
    eidx[k] = idx;
    escore[k] = expf(__fmul_rn(P.temperature, __ldcg(vals + idx) - vmax));   // exp(T * (v - max))
    P.elite_idx32[static_cast<size_t>(e) * P.K + k] = idx;
    if (P.elite_idx_out) P.elite_idx_out[static_cast<size_t>(e) * P.K + k] = idx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < P.K; ++k) s += escore[k];
    red[0] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < P.K; k += kThreads) escore[k] = __fdiv_rn(escore[k], red[0]);   // score /= score.sum(0)
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < P.K; ++k) s += escore[k];
    red[1] = s + 1e-9f;                                        // score.sum(0) + 1e-9
  }
  __syncthreads();
  const float denom = red[1];
  for (int k = threadIdx.x; k < P.K; k += kThreads) P.score[static_cast<size_t>(e) * P.K + k] = escore[k];
  for (int i = threadIdx.x; i < P.H * P.A; i += kThreads) {
    const int t = i / P.A, a = i % P.A;
    float m = 0.f;
    for (int k = 0; k < P.K; ++k) m = fmaf(escore[k], sample_action(P, e, t, eidx[k], a, task), m);
    m = __fdiv_rn(m, denom);
    float var = 0.f;
    for (int k = 0; k < P.K; ++k) {
      const float act = sample_action(P, e, t, eidx[k], a, task);
      const float d = act - m;
      var = fmaf(escore[k], d * d, var);
      if (t == 0) P.elite_act0[(static_cast<size_t>(e) * P.K + k) * P.A + a] = act;
    }
    float sd = sqrtf(__fdiv_rn(var, denom));
    sd = fminf(fmaxf(sd, P.min_std), P.max_std);
    if (P.masks) { const float mk = P.masks[static_cast<size_t>(task) * P.A + a]; m *= mk; sd *= mk; }
    // sample_action() above read the OLD mean/std of this (t, a) only: safe to overwrite now.
    const size_t sa = (static_cast<size_t>(e) * P.H + t) * P.A + a;
    P.mean[sa] = m;
    P.std[sa] = sd;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------ the kernel
template <int ENGINE>
__global__ void __launch_bounds__(kThreads, 1) plan_kernel(const __grid_constant__ PlanParams P) {
  extern __shared__ uint8_t smem_raw[];
  Ctx c;
  {
    uintptr_t base = (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023);
    c.stage_base = reinterpret_cast<uint8_t*>(base);
    uint8_t* ctrl = c.stage_base + kStages * kStageBytes;
    c.full = reinterpret_cast<uint64_t*>(ctrl);
    c.empty = c.full + kStages;
    c.acc_full = c.empty + kStages;
    c.acc_empty = c.acc_full + 2;
    c.tmem_ptr = reinterpret_cast<uint32_t*>(c.acc_empty + 2);
    c.flags = reinterpret_cast<int*>(c.tmem_ptr + 1);          // [8]
    c.G = reinterpret_cast<float*>(ctrl + 128);                 // [128]
    c.q1 = c.G + kTileM;                                        // [128]  (ends at ctrl + 1152 <= kSmemCtrl)
    c.rowbuf = reinterpret_cast<float*>(ctrl + kSmemCtrl);
  }
  c.slot = blockIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.p_it = c.m_it = c.a_it = c.d_it = 0;
  c.tmem_base = 0;
  int* rowenv = reinterpret_cast<int*>(c.rowbuf + kWarps * kMaxHeadCols);   // [128] env of each row (or -1)

  if (ENGINE == ENGINE_TC) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&c.full[s], 1); ptx::mbar_init(&c.empty[s], 1); }
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&c.acc_full[s], 1); ptx::mbar_init(&c.acc_empty[s], 4 * 32); }
      ptx::fence_barrier_init();
      ptx::prefetch_tensormap(&P.tmX);
      ptx::prefetch_tensormap(&P.tmH);
    }
    if (c.warp == 2) ptx::tmem_alloc(c.tmem_ptr, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    c.tmem_base = *c.tmem_ptr;
  }

  const LayerDev* LY = P.layers;
  float* myrow = c.rowbuf + c.warp * kMaxHeadCols;

  for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    // ---------------- tile set-up: fill the input planes of X ----------------
    for (int r = threadIdx.x; r < kTileM; r += kThreads) { rowenv[r] = map_row(P, tile, r).env; c.G[r] = 0.f; c.q1[r] = 0.f; }
    __syncthreads();
    __half* xhi = plane_ptr(P, c.slot, BUF_X, 0);
    __half* xlo = plane_ptr(P, c.slot, BUF_X, 1);
    if (P.mode == MODE_LAYER) {
      const LayerDev& ly = LY[P.dbg_layer];
      for (int r = c.warp; r < kTileM; r += kWarps)
        for (int col = c.lane; col < ly.Kpad; col += 32) {
          const float x = (r < P.dbg_rows && col < ly.K) ? P.dbg_x[static_cast<size_t>(r) * ly.K + col] : 0.f;
          split_store(xhi + static_cast<size_t>(r) * P.KpadX + col, xlo + static_cast<size_t>(r) * P.KpadX + col, x);
        }
      publish_planes();
      gemm<ENGINE>(P, c, ly, BUF_X);
      if (P.dbg_mode == 0) {
        for (int r = c.warp; r < P.dbg_rows; r += kWarps)
          for (int col = c.lane; col < ly.N; col += 32)
            P.dbg_y[static_cast<size_t>(r) * ly.N + col] =
                fmaf(__ldcg(raw_ptr(P, c.slot) + static_cast<size_t>(r) * P.NpadMax + col), ly.inv_scale, ly.bias[col]);
      } else {
        for (int r = threadIdx.x; r < kTileM; r += kThreads) rowenv[r] = r < P.dbg_rows ? r : -1;
        __syncthreads();
        rows_ln_act(P, c, ly, P.dbg_mode == 1 ? ACT_MISH : ACT_SIMNORM, -1, 0, P.dbg_y, ly.N, rowenv);
      }
      __syncthreads();
      continue;
    }

    for (int r = c.warp; r < kTileM; r += kWarps) {
      const RowMap rm = map_row(P, tile, r);
      const int e = rm.env < 0 ? 0 : rm.env;          // padding rows compute on env 0's data, results dropped
      const int task = P.task ? P.task[e] : 0;
      __half* rhi = xhi + static_cast<size_t>(r) * P.KpadX;
      __half* rlo = xlo + static_cast<size_t>(r) * P.KpadX;
      if (P.mode == MODE_ENCODE) {
        // [obs | task_emb]  (world_model.py:108-109)
        for (int col = c.lane; col < P.obs_dim; col += 32)
          split_store(rhi + col, rlo + col, P.obs[static_cast<size_t>(e) * P.obs_dim + col]);
        for (int col = c.lane; col < P.T; col += 32)
          split_store(rhi + P.obs_dim + col, rlo + P.obs_dim + col, P.emb[static_cast<size_t>(task) * P.T + col]);
      } else {
        // [z | task_emb | a]  (world_model.py:119-120)
        const float* zsrc = P.z_rows ? P.z_rows + (static_cast<size_t>(e) * P.N + (rm.env < 0 ? 0 : rm.idx)) * P.L
                                     : P.z + static_cast<size_t>(e) * P.L;
        for (int col = c.lane; col < P.L; col += 32) split_store(rhi + col, rlo + col, zsrc[col]);
        for (int col = c.lane; col < P.T; col += 32)
          split_store(rhi + P.L + col, rlo + P.L + col, P.emb[static_cast<size_t>(task) * P.T + col]);
      }
    }

    if (P.mode == MODE_ENCODE) {
      publish_planes();
      int src = BUF_X;
      for (int l = 0; l < P.num_enc; ++l) {
        const LayerDev& ly = LY[P.li_enc + l];
        gemm<ENGINE>(P, c, ly, src);
        const bool last = (l == P.num_enc - 1);
        const int dst = (src == BUF_H1) ? BUF_H2 : BUF_H1;
        if (last) rows_ln_act(P, c, ly, ACT_SIMNORM, -1, 0, P.z, P.L, rowenv);
        else rows_ln_act(P, c, ly, ACT_MISH, dst, 0, nullptr, 0, nullptr);
        publish_planes();
        src = dst;
      }
      continue;
    }

    // helper lambdas -----------------------------------------------------------
    auto run_mlp_hidden = [&](int li0) {   // layers 0 and 1: X -> H1 -> H2 (LN + Mish)
      gemm<ENGINE>(P, c, LY[li0], BUF_X);
      rows_ln_act(P, c, LY[li0], ACT_MISH, BUF_H1, 0, nullptr, 0, nullptr);
      publish_planes();
      gemm<ENGINE>(P, c, LY[li0 + 1], BUF_H1);
      rows_ln_act(P, c, LY[li0 + 1], ACT_MISH, BUF_H2, 0, nullptr, 0, nullptr);
      publish_planes();
      gemm<ENGINE>(P, c, LY[li0 + 2], BUF_H2);
    };
    auto write_actions = [&](int t) {      // X action columns <- a_t  (tdmpc2.py:176-181)
      for (int r = c.warp; r < kTileM; r += kWarps) {
        const RowMap rm = map_row(P, tile, r);
        const int e = rm.env < 0 ? 0 : rm.env, n = rm.env < 0 ? 0 : rm.idx;
        const int task = P.task ? P.task[e] : 0;
        for (int a = c.lane; a < P.A; a += 32) {
          const float v = sample_action(P, e, t, n, a, task);
          const size_t o = static_cast<size_t>(r) * P.KpadX + P.L + P.T + a;
          split_store(xhi + o, xlo + o, v);
        }
      }
    };
    auto dynamics_step = [&]() {           // z <- next(z, a)  (world_model.py:114-121)
      run_mlp_hidden(P.li_dyn);
      rows_ln_act(P, c, LY[P.li_dyn + 2], ACT_SIMNORM, BUF_X, 0, nullptr, 0, nullptr);
      publish_planes();
    };
    auto pi_step = [&](const float* eps_base, int eps_rows_per_env, float* act_out, int t_out) {
      // a = tanh(mean + eps * exp(log_std))  (world_model.py:144-174); writes X action columns.
      run_mlp_hidden(P.li_pi);
      const LayerDev& ly = LY[P.li_pi + 2];
      for (int r = c.warp; r < kTileM; r += kWarps) {
        const RowMap rm = map_row(P, tile, r);
        const int e = rm.env < 0 ? 0 : rm.env, idx = rm.env < 0 ? 0 : rm.idx;
        const int task = P.task ? P.task[e] : 0;
        head_row_to_smem(P, c, ly, r, myrow);
        for (int a = c.lane; a < P.A; a += 32) {
          float mu = myrow[a];
          float ls = myrow[P.A + a];
          // log_std = low + 0.5 * dif * (tanh(x) + 1)   (math.py:12-13)
          ls = __fadd_rn(P.log_std_min, __fmul_rn(__fmul_rn(0.5f, P.log_std_dif), __fadd_rn(tanhf(ls), 1.f)));
          float eps = eps_base[(static_cast<size_t>(e) * eps_rows_per_env + idx) * P.A + a];
          if (P.masks) { const float mk = P.masks[static_cast<size_t>(task) * P.A + a]; mu *= mk; ls *= mk; eps *= mk; }
          const float act = tanhf(__fadd_rn(mu, __fmul_rn(eps, expf(ls))));
          const size_t o = static_cast<size_t>(r) * P.KpadX + P.L + P.T + a;
          split_store(xhi + o, xlo + o, act);
          if (act_out && rm.env >= 0)
            act_out[((static_cast<size_t>(e) * P.H + t_out) * P.P + idx) * P.A + a] = act;
        }
        __syncwarp();
      }
      publish_planes();
    };

    if (P.mode == MODE_PRIOR) {
      publish_planes();
      for (int t = 0; t < P.H; ++t) {
        // eps for step t: noise_prior[e, t, p, :]
        pi_step(P.noise_prior + static_cast<size_t>(t) * P.P * P.A, P.H * P.P, P.pi_actions, t);
        if (t < P.H - 1) dynamics_step();
      }
      continue;
    }

    // ---------------- MODE_ITER / MODE_VALUE: _estimate_value (tdmpc2.py:122-136) ----------------
    const int env = tile / P.tiles_per_env;
    const int task = P.task ? P.task[env] : 0;
    const float* dpow = P.disc_pow + static_cast<size_t>(task) * (P.H + 1);
    for (int t = 0; t < P.H; ++t) {
      write_actions(t);
      publish_planes();
      // reward (world_model.py:123-130) + two_hot_inv
      run_mlp_hidden(P.li_rew);
      {
        const LayerDev& ly = LY[P.li_rew + 2];
        const float disc = dpow[t];
        for (int r = c.warp; r < kTileM; r += kWarps) {
          head_row_to_smem(P, c, ly, r, myrow);
          const float rew = two_hot_inv_row(P, c, myrow);
          if (c.lane == 0) c.G[r] = __fadd_rn(c.G[r], __fmul_rn(disc, rew));   // G + discount * reward
          __syncwarp();
        }
      }
      __syncthreads();
      dynamics_step();
    }
    pi_step(P.noise_pi, P.N, nullptr, 0);
    const int* qi = P.qidx + static_cast<size_t>(env) * 2;
    for (int h = 0; h < 2; ++h) {
      const int li = P.li_q + 3 * qi[h];
      run_mlp_hidden(li);
      const LayerDev& ly = LY[li + 2];
      for (int r = c.warp; r < kTileM; r += kWarps) {
        head_row_to_smem(P, c, ly, r, myrow);
        const float q = two_hot_inv_row(P, c, myrow);
        if (c.lane == 0) {
          if (h == 0) c.q1[r] = q;
          else {
            const float qavg = __fmul_rn(__fadd_rn(c.q1[r], q), 0.5f);          // Q.sum(0) / 2
            float v = __fadd_rn(c.G[r], __fmul_rn(dpow[P.H], qavg));
            if (P.mode == MODE_ITER) v = nan_to_num0(v);                          // tdmpc2.py:184
            const RowMap rm = map_row(P, tile, r);
            if (rm.env >= 0) {
              float* dst = (P.mode == MODE_ITER) ? P.values : P.values_out;
              dst[static_cast<size_t>(rm.env) * P.N + rm.idx] = v;
            }
          }
        }
        __syncwarp();
      }
      __syncthreads();
    }
    if (P.mode == MODE_ITER) {
      // last CTA to finish a tile of this environment refits its mean/std
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(&P.env_counter[env], 1u);
        c.flags[0] = (old == static_cast<unsigned>(P.tiles_per_env - 1));
        if (c.flags[0]) P.env_counter[env] = 0;
      }
      __syncthreads();
      if (c.flags[0]) {
        __threadfence();
        refit_env(P, c, env, task);
      }
      __syncthreads();
    }
  }

  if (ENGINE == ENGINE_TC) {
    ptx::tc_fence_before();
    __syncthreads();
    if (c.warp == 2) ptx::tmem_dealloc(c.tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ small kernels
// mean/std initialisation (tdmpc2.py:164-167)
__global__ void init_state_kernel(float* mean, float* std, unsigned* env_counter, const float* prev_mean,
                                  const uint8_t* t0, int E, int H, int A, float max_std) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) env_counter[i] = 0;
  if (i >= E * H * A) return;
  const int a = i % A, t = (i / A) % H, e = i / (A * H);
  float m = 0.f;
  if (!t0[e] && t < H - 1) m = prev_mean[(static_cast<size_t>(e) * H + t + 1) * A + a];
  mean[i] = m;
  std[i] = max_std;
}

// final action (tdmpc2.py:199-206, math.py:86-94); one warp per environment
__global__ void pick_kernel(const float* score, const float* elite_act0, const float* mean, const float* std,
                            const float* expo, const float* noise_final, float* action, float* prev_mean_out,
                            int* pick_out, int E, int K, int H, int A) {
  const int e = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  // argmax_k softmax(log(score_k) - log(expo_k)) == argmax of the logits (first max on ties)
  float best = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float g = __fadd_rn(logf(score[static_cast<size_t>(e) * K + k]), -logf(expo[static_cast<size_t>(e) * K + k]));
    if (g > best || (g == best && k < bi)) { best = g; bi = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0 && pick_out) pick_out[e] = bi;
  for (int a = lane; a < A; a += 32) {
    float v = elite_act0[(static_cast<size_t>(e) * K + bi) * A + a];
    if (noise_final) v = __fadd_rn(v, __fmul_rn(std[(static_cast<size_t>(e) * H) * A + a], noise_final[static_cast<size_t>(e) * A + a]));
    action[static_cast<size_t>(e) * A + a] = fminf(fmaxf(v, -1.f), 1.f);
  }
  for (int i = lane; i < H * A; i += 32) prev_mean_out[static_cast<size_t>(e) * H * A + i] = mean[static_cast<size_t>(e) * H * A + i];
}

}  // namespace tdmpc2
