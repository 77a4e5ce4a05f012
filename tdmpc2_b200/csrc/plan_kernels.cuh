// Fused planning kernels for sm_100a (B200).
//
// One persistent kernel template, `plan_kernel<ENGINE, PAIR, EPISODIC, WPF>`, runs the MLP chains of
// the TD-MPC2 planner on 128-row tiles (the flags are compile-time so that an unused feature costs no code:
// PAIR = CTA pairs / cta_group::2, EPISODIC = termination head in the rollout, WPF = weight prefetch during the
// epilogue).  All modes share the device code:
//
//   MODE_ENCODE : rows = environments.      z = encode(obs, task)        (reference world_model.py:103-112)
//   MODE_PRIOR  : rows = (env, pi-traj).    the P policy-prior rollouts  (reference tdmpc2.py:154-160)
//   MODE_ITER   : rows = (env, sample).     ONE CEM iteration:           (reference tdmpc2.py:173-197)
//                   sample actions -> H x (reward, dynamics) -> terminal pi -> 2 Q heads
//                   -> value -> [last tile of an env] top-k, MPPI weights, mean/std refit.
//   MODE_VALUE  : MODE_ITER's rollout on caller-given z / actions, no refit (reference tdmpc2.py:122-136)
//   MODE_LAYER  : one layer, for diagnostics.
//
// Every dense layer is `acc = A[128, Kpad] * W[Npad, Kpad]^T` with both operands
// stored as two fp16 planes (hi, lo; x ~= hi + lo to ~22 bits).  The tcgen05
// engine accumulates A_lo*W_hi + A_hi*W_lo + A_hi*W_hi in fp32 in TMEM
// (kind::f16, K=16), operands staged by TMA (128-byte swizzle) through two
// mbarrier rings (A: 2 x 32 KiB, W: 2 x 64 KiB | 4 x 32 KiB) so that an A K-chunk
// is streamed once per layer: warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 4-19 = epilogue (4 TMEM lane quarters x 4 column groups).
//
//   * PAIR mode (CEM iterations): clusters of two CTAs issue cta_group::2 MMAs with
//     M = 256 (each CTA's own 128-row tile); each CTA streams half of every W tile.
//     plan_pp.cuh holds the ping-pong variant of this mode (two 64-row halves per CTA).
//   * Layers with Npad <= 512 (the whole accumulator fits TMEM) use the FUSED
//     epilogue: one thread per row reads its accumulator row straight from TMEM
//     (tcgen05.ld), applies bias + LayerNorm + Mish / SimNorm / two-hot-inverse /
//     tanh-Gaussian sampling in packed fp32x2 math, and emits the next layer's fp16
//     planes as swizzled smem tiles that leave through TMA stores.
//   * Wider layers (48M / 317M presets) drain N-chunks of 256 columns to an fp32
//     scratch row buffer and run the same math warp-per-row afterwards.
//
// Activations live in a per-CTA scratch slot (global memory, L2-resident: X planes +
// one in-place hidden buffer, 0.56 MB per slot for the 5M model).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "ptx.cuh"
#include "rng.cuh"

namespace tdmpc2 {

constexpr int kTileM = 128;       // rows per tile (UMMA M)
constexpr int kKch = 64;          // K elements per pipeline stage (128 B of fp16: one swizzle row)
constexpr int kNch = 256;         // N columns per accumulator chunk (UMMA N max)
constexpr int kStages = 2;
constexpr int kAPlane = kTileM * 128;           // 16 KiB: one A plane of a stage
constexpr int kWPlane = kNch * 128;             // 32 KiB: one W plane of a stage
constexpr int kStageBytes = 2 * kAPlane + 2 * kWPlane;   // 96 KiB (the operand smem is kStages * kStageBytes = 192 KiB)
// The 192 KiB are split into two independent rings so that an A K-chunk is loaded ONCE per layer and re-used by
// every N-chunk of weights: A ring = 2 x (hi 16 KiB | lo 16 KiB), W ring = 2 x (hi 32 KiB | lo 32 KiB).
constexpr int kASlotBytes = 2 * kAPlane;                  // 32 KiB
constexpr int kWSlotBytes = 2 * kWPlane;                  // 64 KiB
constexpr int kARing = 2, kWRing = 2;
constexpr int kWRingPair = 4;                            // pair mode: 4 half-size W slots (hi 16 KiB | lo 16 KiB) in the same 128 KiB
constexpr int kWRingOff = kARing * kASlotBytes;           // 64 KiB
static_assert(kARing * kASlotBytes + kWRing * kWSlotBytes == kStages * kStageBytes, "operand smem layout");
// Fused-epilogue staging for TMA stores: per column group 2 buffers x (hi 8 KiB | lo 8 KiB) = one 32-column block
// each, 64-byte-swizzled; the 128 KiB alias the W ring (idle once every MMA of the layer has retired).
constexpr int kStgPlane = kTileM * 64;                    // 8 KiB
constexpr int kStgBuf = 2 * kStgPlane;                    // 16 KiB
constexpr int kThreads = 576;                   // 18 warps: 65536 / 576 = 113 -> 112 registers per thread
constexpr int kWarps = kThreads / 32;
constexpr int kEpiWarp0 = 2;                    // warps 2..17 are the epilogue warps: 4 column groups x 4 lane quarters
constexpr int kEpiGroups = 4;
constexpr int kEpiThreads = kEpiGroups * 128;
constexpr int kMaxWMaps = 8;
constexpr int kMaxHeadCols = 256;  // widest head output (2A or num_bins)
constexpr int kFusedMaxN = 512;    // TMEM columns
constexpr int kSmemCtrl = 2048;    // barriers, tmem ptr, flags (128 B) + G[128] + q1[128]
constexpr int kSmemRowBuf = kWarps * kMaxHeadCols * 4;  // one head-output row per warp (wide path)
constexpr int kSmemRowEnv = kTileM * 4;                 // env index of each tile row
constexpr int kSmemVec = 3 * kFusedMaxN * 4;            // bias / ln_g / ln_b of the current layer
constexpr int kSmemPart = 2 * kEpiGroups * kTileM * 4;  // LayerNorm partials [mean | M2][group][128]
constexpr int kSmemBytes = kStages * kStageBytes + kSmemCtrl + kSmemRowBuf + kSmemRowEnv + kSmemVec + kSmemPart +
                           1024 /*align slack*/;

enum Mode { MODE_ENCODE = 0, MODE_PRIOR = 1, MODE_ITER = 2, MODE_VALUE = 3, MODE_LAYER = 4 };
enum Engine { ENGINE_TC = 0, ENGINE_SIMT = 1 };
// X = [z | emb | a] input planes; H = hidden planes.  ONE hidden buffer is enough: a layer's epilogue only runs after
// every MMA of its GEMM has consumed the A operand, so layer 1 overwrites its own input in place.
enum Buf { BUF_X = 0, BUF_H1 = 1 };
enum Epi { EPI_LN_MISH = 0, EPI_LN_SIMNORM = 1, EPI_TWOHOT = 2, EPI_PI = 3, EPI_RAW = 4, EPI_TERM = 5 };
enum HeadKind { HEAD_REWARD = 0, HEAD_Q1 = 1, HEAD_Q2 = 2 };

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
  CUtensorMap tmH;                 // [slots*2*128, KpadH]
  CUtensorMap tmXs, tmHs;          // same tensors, box 32 x 128, 64-byte swizzle: the epilogue's TMA stores
  CUtensorMap tmX64, tmH64;        // 64-row boxes (ping-pong engine, plan_pp.cuh): loads 64 x 64, 128-byte swizzle
  CUtensorMap tmXs64, tmHs64;      //                                             : stores 16 x 64, 32-byte rows, no swizzle
  CUtensorMap tmW[kMaxWMaps];      // weights, one map per Kpad class, box 64 x 128
  const LayerDev* layers;
  int E, N, P, Ppad, K, H, obs_dim, A, Apad, L, M, T, B, num_q, simnorm, num_enc;   // Apad = pad32(A): the pi head's
                                                                                  // log_std logits start at column Apad
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
  long long* prof;   // optional [prof_slots][4][12] cycle counters + [32][16] trace stamps of CTA 0 (diagnostics), or nullptr
  int prof_slots;    // = number of scratch slots (SM count)
  int li_term;       // first of the 3 termination-head layers (cfg.episodic, world_model.py:28), or -1
  int head_kseg;     // head layers: K-chunks per accumulator hand-off (0 = whole K in one accumulation)
  unsigned wide_sleep_ns;   // wide layers: nanosleep between polls of the 16 epilogue warps' accumulator wait (0 = spin)
  int kseg;          // wide layers: K-chunks (of 64) accumulated in TMEM before the partial sum is flushed to the fp32 raw
                     // scratch and added there with round-to-nearest (0 = the whole K in one go); see epi_wide
  // Shared-latent fold (MODE_ITER, rollout step t = 0): every sample row of an environment carries the SAME [z | emb]
  // (z.repeat(N), tdmpc2.py:163), so the [z | emb] part of reward.0 / dynamics.0 is a per-environment vector.  The
  // prologue computes zbias[mlp][e][n] = bias[n] + sum_{k < 64 zb_kc0} [z | emb]_e[k] W[n][k] once per plan()
  // (zbias_kernel) and the t = 0 GEMMs of those two layers start at K-chunk zb_kc0 (the action columns) with that
  // vector as their bias.  zb_kc0 == 0: fold off.
  const float* zbias;   // [2 (0 reward, 1 dynamics)][E][zb_pitch]
  int zb_kc0, zb_pitch;
  // 3 = fp32-parity arithmetic (A_lo W_hi + A_hi W_lo + A_hi W_hi); 1 = the DECLARED NON-PARITY fast mode: hi planes only
  // (one fp16 MMA per product, fp32 accumulate), half the operand bytes, no lo planes written.  tdmpc2_planner_set_passes.
  int passes;
  // MODE_ITER: every second CTA (pair) starts this many clock cycles late, so that neighbouring SMs are not all in their
  // GEMM phase (L2 -> SM ingest, tensor-pipe power) and all in their epilogue at the same instants.  0 = off.
  unsigned stagger;
  // Wide models (activation planes + weights exceed L2): 1 = operand loads carry L2 eviction hints -- the per-CTA activation
  // planes, re-read once per 512-column super-chunk with a reuse distance far beyond L2, are loaded evict-first so that
  // they stop displacing the weight chunks that all 148 CTAs read within a short window (evict-last).  Bit 1 (value 2):
  // evict-last on the fused epilogues' activation-plane stores (default on).  0 = evict-normal everywhere.
  int l2hint;
  // Declared non-parity throughput mode (rng.cuh): non-null = the CEM iteration generates noise_r / noise_pi itself
  // (noise_r / noise_pi are null then); rng_iter = index of this iteration within the plan (selects the Philox stream)
  const unsigned long long* rng_state;
  int rng_iter;
};

// The layer table lives in global memory; role loops are full of asm volatile(... "memory") (TMA issue, mbarrier waits,
// fences), each of which would force the compiler to re-read any field it needs afterwards -- a dependent global load
// in the single-thread producer / MMA loops and in the epilogue's inner loops.  Roles therefore work on a by-value copy.
struct LayerRec {
  int K, Kpad, N, Npad, wmap, wrow;
  float inv_scale;
  const float* bias; const float* ln_g; const float* ln_b;
  const __half* w_hi; const __half* w_lo;
};
__device__ __forceinline__ LayerRec layer_rec(const LayerDev& l) {
  LayerRec r;
  r.K = l.K; r.Kpad = l.Kpad; r.N = l.N; r.Npad = l.Npad; r.wmap = l.wmap; r.wrow = l.wrow; r.inv_scale = l.inv_scale;
  r.bias = l.bias; r.ln_g = l.ln_g; r.ln_b = l.ln_b; r.w_hi = l.w_hi; r.w_lo = l.w_lo;
  return r;
}

// What a layer's epilogue has to do besides the activation itself.
struct EpiArgs {
  int kind;                 // Epi
  int dstbuf;               // LN kinds: plane destination buffer (or -1)
  int dst_col0;
  float* out_f32;           // LN kinds / RAW: optional fp32 row output
  int out_pitch;
  const int* rowmap;        // smem [128]: output row of each tile row (or -1), for out_f32
  int head;                 // TWOHOT: HeadKind
  float disc;               // TWOHOT: discount factor applied to this head's value
  int tile;                 // TWOHOT / PI: tile index (row -> env mapping)
  const float* eps_base;    // PI: noise tensor, element (e*eps_rows + idx)*A + a
  int eps_rows;
  float* act_out;           // PI: optional pi_actions output
  int t_out;
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
__device__ __forceinline__ void split_f(float x, __half& h, __half& l) {
  // finite values are clamped into fp16 range; NaN / +-inf stay non-finite so that they poison the row exactly
  // as they do in the reference (whose planner then zeroes the value: nan_to_num, tdmpc2.py:184)
  const bool finite = fabsf(x) <= 3.0e38f;
  x = fminf(fmaxf(x, -65000.f), 65000.f);
  h = finite ? __float2half_rn(x) : __ushort_as_half(static_cast<unsigned short>(0x7fff));
  l = __float2half_rn(finite ? x - __half2float(h) : 0.f);
}
__device__ __forceinline__ void split_store(__half* hi, __half* lo, float x) {
  __half h, l;
  split_f(x, h, l);
  *hi = h;
  *lo = l;
}
__device__ __forceinline__ float nan_to_num0(float v) {   // torch.nan_to_num(0): nan->0, +-inf -> +-FLT_MAX
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}
__device__ __forceinline__ float4 lds128(const float* p) {   // p must point into shared memory, 16-byte aligned
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(ptx::smem_u32(p)));
  return r;
}
__device__ __forceinline__ void group_bar_sync(int grp) {   // named barrier among the 4 warps of one column group
  asm volatile("bar.sync %0, 128;\n" ::"r"(2 + grp) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() {   // named barrier among the 8 epilogue warps
  asm volatile("bar.sync 1, %0;\n" ::"n"(kEpiThreads) : "memory");
}

// Diagnostics (cycle counters per role, clock stamps per layer of CTA 0) exist only in builds with -DTDMPC2_PROF
// (tdmpc2_b200.build.build_variant("prof", ["TDMPC2_PROF=1"]), loaded through TDMPC2_B200_LIB): in the product build
// they compile to nothing -- the 8 per-thread 64-bit counters alone cost 16 registers of a 96-register budget.
#ifdef TDMPC2_PROF
constexpr bool kProf = true;
#else
constexpr bool kProf = false;
#endif
__device__ __forceinline__ long long prof_clock() { return kProf ? clock64() : 0ll; }
// Raw clock stamps of CTA 0 (events x layers) appended after the per-CTA counters.
#define TDMPC2_TRACE(P_, c_, ev_)                                                                         \
  do {                                                                                                    \
    if (kProf && (P_).prof && blockIdx.x == 0 && (c_).trace_step < 32)                                              \
      (P_).prof[static_cast<size_t>((P_).prof_slots) * 4 * 12 + (c_).trace_step * 16 + (ev_)] = prof_clock();                                  \
  } while (0)

// ------------------------------------------------------------------------------------ CTA context
struct Ctx {
  uint8_t* stage_base;      // kStages * kStageBytes, 1024-aligned
  uint64_t* a_full;         // [kARing]
  uint64_t* a_empty;        // [kARing]
  uint64_t* w_full;         // [kWRing]
  uint64_t* w_empty;        // [kWRing]
  uint64_t* acc_full;       // [2]  wide path: accumulator slot ready
  uint64_t* acc_empty;      // [2]  wide path: accumulator slot drained
  uint64_t* facc;           // [2]  fused path: accumulator chunk ready
  uint64_t* rawb;           // [4 groups][2] wide layers, normalise pass: raw block landed in the group's input buffer
  uint32_t* tmem_ptr;
  float* rowbuf;            // [kWarps][kMaxHeadCols]
  float* vec;               // [3][kFusedMaxN]  bias, ln_g, ln_b (fused path)
  float* part;              // [2][2][128]      LayerNorm partials (fused path)
  float* G;                 // [128] discounted reward sum
  float* q1;                // [128] first Q head
  float* term;              // [128] sticky termination flag of the row (episodic models, tdmpc2.py:126-134)
  int* flags;               // small ints
  int* rowenv;              // [128]
  uint32_t tmem_base;
  int slot, warp, lane;
  int cg2, rank;             // CTA-pair mode (tcgen05 cta_group::2): pair rank 0 = leader issues the MMAs
  int wpf;                   // W prefetch: the producer streams the NEXT layer's first weight chunks during the epilogue
  uint32_t w_pref;           // (producer thread) weight chunks of the coming layer that are already in flight
  uint32_t w_ring, w_stride, w_lo_off;   // W ring geometry: 2 x 64 KiB (lo plane at +32 KiB) or, in pair mode, 4 x 32 KiB (+16 KiB)
  // pipeline counters (each role keeps its own; persist across layers / tiles)
  uint32_t pa_it, pw_it, ma_it, mw_it, a_it, d_it;
  uint32_t rb_it;           // wide layers: raw blocks consumed so far by this thread's column group (buffer = it & 1)
  uint32_t fph0, fph1;      // fused path: phase parity of facc[0|1] (tracked identically by every thread)
  long long pf0, pf1, pf2, pf3, pf4, pf5, pf6, pf7;   // per-thread cycle accumulators (diagnostics)
  int trace_step;
};

__device__ __forceinline__ __half* plane_ptr(const PlanParams& P, int slot, int buf, int plane) {
  if (buf == BUF_X) return P.X + (static_cast<size_t>(slot) * 2 + plane) * kTileM * P.KpadX;
  return P.Hb + (static_cast<size_t>(slot) * 2 + plane) * kTileM * P.KpadH;
}
__device__ __forceinline__ int plane_pitch(const PlanParams& P, int buf) { return buf == BUF_X ? P.KpadX : P.KpadH; }
__device__ __forceinline__ int plane_row0(const PlanParams& P, int slot, int buf, int plane) {  // TMA row coord
  if (buf == BUF_X) return (slot * 2 + plane) * kTileM;
  return (slot * 2 + plane) * kTileM;
}
__device__ __forceinline__ float* raw_ptr(const PlanParams& P, int slot) {
  return P.raw + static_cast<size_t>(slot) * kTileM * P.NpadMax;
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

// In-kernel noise (rng.cuh): group indices.  noise_r element (e, t, n, a), n >= P: group ((e H + t) N + n) A4 + a / 4 of
// stream 2 iter; noise_pi element (e, n, a): group (e N + n) A4 + a / 4 of stream 2 iter + 1; A4 = ceil(A / 4).
__device__ __forceinline__ unsigned long long rng_group_r(const PlanParams& P, int e, int t, int n, int a4) {
  return ((static_cast<unsigned long long>(e) * P.H + t) * P.N + n) * static_cast<unsigned>((P.A + 3) >> 2) + a4;
}
__device__ __forceinline__ unsigned long long rng_group_pi(const PlanParams& P, int e, int n, int a4) {
  return (static_cast<unsigned long long>(e) * P.N + n) * static_cast<unsigned>((P.A + 3) >> 2) + a4;
}
__device__ __forceinline__ float noise_r_at(const PlanParams& P, int e, int t, int n, int a) {
  if (P.rng_state) return rng_pick(rng_normal4(P.rng_state, 2u * P.rng_iter, rng_group_r(P, e, t, n, a >> 2)), a & 3);
  return __ldcs(&P.noise_r[((static_cast<size_t>(e) * P.H + t) * (P.N - P.P) + (n - P.P)) * P.A + a]);
}
__device__ __forceinline__ float noise_pi_at(const PlanParams& P, int e, int n, int a) {      // MODE_ITER / VALUE terminal policy sample
  if (P.rng_state) return rng_pick(rng_normal4(P.rng_state, 2u * P.rng_iter + 1u, rng_group_pi(P, e, n, a >> 2)), a & 3);
  return __ldcs(&P.noise_pi[(static_cast<size_t>(e) * P.N + n) * P.A + a]);
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
    const float r = noise_r_at(P, e, t, n, a);
    v = __fadd_rn(P.mean[sa], __fmul_rn(P.std[sa], r));       // mean + std * r, two roundings like eager torch
    v = fminf(fmaxf(v, -1.f), 1.f);
  }
  if (P.masks) v *= P.masks[static_cast<size_t>(task) * P.A + a];
  return v;
}

// ------------------------------------------------------------------------------------ shared per-row commits
// Value bookkeeping of _estimate_value (tdmpc2.py:128-136) for one row; called by exactly one thread per row.
template <bool EPISODIC>
__device__ __forceinline__ void head_commit(const PlanParams& P, Ctx& c, const EpiArgs& ea, int r, float val) {
  // discount * (1 - termination), tdmpc2.py:130,136 (termination is identically 0 unless cfg.episodic)
  const float disc = EPISODIC ? __fmul_rn(ea.disc, __fsub_rn(1.f, c.term[r])) : ea.disc;
  if (ea.head == HEAD_REWARD) {
    c.G[r] = __fadd_rn(c.G[r], __fmul_rn(disc, val));                 // G + discount * reward
  } else if (ea.head == HEAD_Q1) {
    c.q1[r] = val;
  } else {
    const float qavg = __fmul_rn(__fadd_rn(c.q1[r], val), 0.5f);      // Q.sum(0) / 2
    float v = __fadd_rn(c.G[r], __fmul_rn(disc, qavg));               // G + discount * Q
    if (P.mode == MODE_ITER) v = nan_to_num0(v);                        // tdmpc2.py:184
    const RowMap rm = map_row(P, ea.tile, r);
    if (rm.env >= 0) {
      float* dst = (P.mode == MODE_ITER) ? P.values : P.values_out;
      dst[static_cast<size_t>(rm.env) * P.N + rm.idx] = v;
    }
  }
}
// termination = clip(termination + (sigmoid(logit) > 0.5), max=1)  (tdmpc2.py:133-134, world_model.py:132-141).
// torch's fp32 sigmoid returns exactly 0.5 for |x| < ~3e-8, so "> 0.5" is evaluated on the same expression.
__device__ __forceinline__ void term_commit(Ctx& c, int r, float logit) {
  const float sg = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-logit)));
  c.term[r] = fminf(__fadd_rn(c.term[r], sg > 0.5f ? 1.f : 0.f), 1.f);
}
// a = tanh(mean + eps * exp(log_std)) for one (row, action dim)  (world_model.py:151-174)
__device__ __forceinline__ float pi_action(const PlanParams& P, float mu, float ls, float eps, int task, int a) {
  // log_std = low + 0.5 * dif * (tanh(x) + 1)   (math.py:12-13)
  ls = __fadd_rn(P.log_std_min, __fmul_rn(__fmul_rn(0.5f, P.log_std_dif), __fadd_rn(tanhf(ls), 1.f)));
  if (P.masks) { const float mk = P.masks[static_cast<size_t>(task) * P.A + a]; mu *= mk; ls *= mk; eps *= mk; }
  return tanhf(__fadd_rn(mu, __fmul_rn(eps, expf(ls))));
}

// ------------------------------------------------------------------------------------ TMA producer / MMA issuer
// FUSED == true : the whole accumulator (Npad <= 512) lives in TMEM: K-chunks outermost, each A chunk is loaded
//                 once and multiplied with every N-chunk of weights; facc[0] fires when all chunks are complete.
// FUSED == false: wide path, N-chunks outermost, accumulator chunks alternate between two 256-column TMEM slots
//                 (acc_full / acc_empty) and A is re-streamed per chunk.
__device__ __forceinline__ void prod_load_a(const PlanParams& P, Ctx& c, const CUtensorMap* tmA, int kc, int arow_hi, int arow_lo) {
  const uint32_t s = c.pa_it % kARing, ph = (c.pa_it / kARing) & 1;
  const long long tw = prof_clock();
  ptx::mbar_wait(&c.a_empty[s], ph ^ 1);
  c.pf0 += prof_clock() - tw;
  uint8_t* st = c.stage_base + s * kASlotBytes;
  const bool lo = P.passes != 1;                 // fast mode streams the hi planes only
  if (ptx::elect_one()) {                        // the whole warp runs the role loop, one lane issues (see tc_producer)
    if (c.cg2) {
      // both CTAs of the pair stream their own 128 rows; the bytes of both land on the leader's barrier
      if (c.rank == 0) ptx::mbar_expect_tx(&c.a_full[s], lo ? 2 * kASlotBytes : 2 * kAPlane);
      const uint64_t pol = (P.l2hint & 1) ? ptx::kL2EvictFirst : ptx::kL2EvictNormal;
      ptx::tma_load_2d_2sm(tmA, &c.a_full[s], st, kc * kKch, arow_hi, pol);
      if (lo) ptx::tma_load_2d_2sm(tmA, &c.a_full[s], st + kAPlane, kc * kKch, arow_lo, pol);
    } else {
      ptx::mbar_expect_tx(&c.a_full[s], lo ? kASlotBytes : kAPlane);
      ptx::tma_load_2d(tmA, &c.a_full[s], st, kc * kKch, arow_hi);
      if (lo) ptx::tma_load_2d(tmA, &c.a_full[s], st + kAPlane, kc * kKch, arow_lo);
    }
  }
  ++c.pa_it;
}
__device__ __forceinline__ void prod_load_w(Ctx& c, const CUtensorMap* tmW, int Npad, int wrow, int kc, int nc, bool lo = true,
                                            bool l2hint = false) {
  const int ncols = min(kNch, Npad - nc * kNch);   // 128 or 256
  const uint32_t s = c.pw_it % c.w_ring, ph = (c.pw_it / c.w_ring) & 1;
  const long long tw = prof_clock();
  ptx::mbar_wait(&c.w_empty[s], ph ^ 1);
  c.pf0 += prof_clock() - tw;
  uint8_t* st = c.stage_base + kWRingOff + s * c.w_stride;
  if (!ptx::elect_one()) { ++c.pw_it; return; }
  if (c.cg2) {
    // each CTA streams HALF of the N-chunk's weight rows (one 128-row box per plane; for a 128-column chunk only
    // its first 64 rows are consumed): the pair MMA reads B rows [0, N/2) from the leader and [N/2, N) from the peer
    if (c.rank == 0) ptx::mbar_expect_tx(&c.w_full[s], (lo ? 2 : 1) * (2 * 128 * 128));
    const int wr = wrow + nc * kNch + c.rank * (ncols / 2);
    const uint64_t pol = l2hint ? ptx::kL2EvictLast : ptx::kL2EvictNormal;
    ptx::tma_load_2d_2sm(tmW, &c.w_full[s], st, kc * kKch, wr, pol);
    if (lo) ptx::tma_load_2d_2sm(tmW, &c.w_full[s], st + c.w_lo_off, kc * kKch, wr + Npad, pol);
  } else {
    ptx::mbar_expect_tx(&c.w_full[s], (lo ? 2 : 1) * ncols * 128);
    for (int b = 0; b < ncols / 128; ++b) {
      const int wr = wrow + nc * kNch + b * 128;
      ptx::tma_load_2d(tmW, &c.w_full[s], st + b * (128 * 128), kc * kKch, wr);
      if (lo) ptx::tma_load_2d(tmW, &c.w_full[s], st + kWPlane + b * (128 * 128), kc * kKch, wr + Npad);
    }
  }
  ++c.pw_it;
}

// N-chunks [nc0, nc0 + nnc_lim) of the layer (default: all of them): K-chunks outermost, each A chunk is loaded once
// and multiplied with every N-chunk of the range; layers wider than TMEM call this once per 512-column super-chunk.
// Role loops (TMA producer, MMA issuer) are run by the WHOLE warp, with one elected lane issuing the asynchronous
// instructions: in warp-convergent code the compiler keeps ring counters, coordinates and operand descriptors in uniform
// registers (12 UTCHMMA back to back), whereas inside an `if (lane == 0)` branch every cp.async.bulk.tensor / tcgen05.mma
// gets a ~13-instruction "elect + R2UR + retry" waterfall (~120 cycles per MMA: more than an N = 128 MMA takes).
__device__ __forceinline__ void tc_producer(const PlanParams& P, Ctx& c, const LayerRec& ly, int srcbuf, const LayerDev* next,
                                            int nc0 = 0, int nnc_lim = 1 << 30, int kc0 = 0, int kc_lim = 1 << 30, int next_kc0 = 0) {
  const int nkc = min(ly.Kpad / kKch, kc0 + kc_lim);
  const int nnc = min((ly.Npad + kNch - 1) / kNch - nc0, nnc_lim);
  const CUtensorMap* tmA = (srcbuf == BUF_X) ? &P.tmX : &P.tmH;
  const CUtensorMap* tmW = &P.tmW[ly.wmap];
  const int arow_hi = plane_row0(P, c.slot, srcbuf, 0), arow_lo = plane_row0(P, c.slot, srcbuf, 1);
  TDMPC2_TRACE(P, c, 1);
  uint32_t skip = c.wpf ? c.w_pref : 0u;     // chunks the previous layer's producer pass already requested
  for (int kc = kc0; kc < nkc; ++kc) {
    prod_load_a(P, c, tmA, kc, arow_hi, arow_lo);
    for (int nc = nc0; nc < nc0 + nnc; ++nc) {
      if (skip) --skip;
      else prod_load_w(c, tmW, ly.Npad, ly.wrow, kc, nc, P.passes != 1, (P.l2hint & 1) != 0);
    }
  }
  if (c.wpf) {
    // The W ring drains while this layer's last MMAs retire and then idles through the whole epilogue (which
    // stages its output in the A ring in this mode): fill it with the head of the next layer's weight stream.
    uint32_t n = 0;
    if (next) {
      const CUtensorMap* tmW2 = &P.tmW[next->wmap];
      const int nkc2 = next->Kpad / kKch, nnc2 = (next->Npad + kNch - 1) / kNch;
      for (int kc = next_kc0; kc < nkc2 && n < c.w_ring; ++kc)      // next_kc0: the next layer's first K-chunk (shared-latent fold)
        for (int nc = 0; nc < nnc2 && n < c.w_ring; ++nc) { prod_load_w(c, tmW2, next->Npad, next->wrow, kc, nc, P.passes != 1); ++n; }
    }
    c.w_pref = n;
  }
}

// 12 MMAs of one (A K-chunk, W K-chunk x N-chunk) pair: A_lo*W_hi + A_hi*W_lo + A_hi*W_hi, 4 K-steps of 16.
__device__ __forceinline__ void mma_stage(uint32_t d, uint32_t sa, uint32_t sw, uint32_t w_lo_off, uint32_t idesc, bool first, bool cg2,
                                          bool one_pass = false) {
#ifdef TDMPC2_EXP_NOMMA   // measurement build: no MMAs are issued, the commits release the operand slots at once -> the GEMM
  return;                 // phases last exactly as long as the operand ingest (results are garbage; scripts/gpu_r2l.sh)
#endif
  if (!ptx::elect_one()) return;   // whole-warp role loop, one lane issues (tcgen05.commit must come from the same lane:
                                   // elect.sync with a full mask always picks the same one)
  if (one_pass) {                                   // declared non-parity fast mode: A_hi * W_hi only
#pragma unroll
    for (int ks = 0; ks < kKch / 16; ++ks) {
      const uint64_t a_hi = ptx::make_sw128_kmajor_desc(sa + ks * 32);
      const uint64_t w_hi = ptx::make_sw128_kmajor_desc(sw + ks * 32);
      if (cg2) ptx::umma_f16_2sm(d, a_hi, w_hi, idesc, !(first && ks == 0));
      else ptx::umma_f16(d, a_hi, w_hi, idesc, !(first && ks == 0));
    }
    return;
  }
#pragma unroll
  for (int ks = 0; ks < kKch / 16; ++ks) {
    const uint64_t a_hi = ptx::make_sw128_kmajor_desc(sa + ks * 32);
    const uint64_t a_lo = ptx::make_sw128_kmajor_desc(sa + kAPlane + ks * 32);
    const uint64_t w_hi = ptx::make_sw128_kmajor_desc(sw + ks * 32);
    const uint64_t w_lo = ptx::make_sw128_kmajor_desc(sw + w_lo_off + ks * 32);
    if (cg2) {
      ptx::umma_f16_2sm(d, a_lo, w_hi, idesc, !(first && ks == 0));
      ptx::umma_f16_2sm(d, a_hi, w_lo, idesc, 1);
      ptx::umma_f16_2sm(d, a_hi, w_hi, idesc, 1);
    } else {
      ptx::umma_f16(d, a_lo, w_hi, idesc, !(first && ks == 0));   // small terms first
      ptx::umma_f16(d, a_hi, w_lo, idesc, 1);
      ptx::umma_f16(d, a_hi, w_hi, idesc, 1);
    }
  }
}
__device__ __forceinline__ uint32_t mma_wait_a(Ctx& c) {
  const uint32_t s = c.ma_it % kARing, ph = (c.ma_it / kARing) & 1;
  const long long tw = prof_clock();
  ptx::mbar_wait(&c.a_full[s], ph);
  c.pf0 += prof_clock() - tw;
  return s;
}
__device__ __forceinline__ uint32_t mma_wait_w(Ctx& c) {
  const uint32_t s = c.mw_it % c.w_ring, ph = (c.mw_it / c.w_ring) & 1;
  const long long tw = prof_clock();
  ptx::mbar_wait(&c.w_full[s], ph);
  c.pf0 += prof_clock() - tw;
  return s;
}

// Accumulates N-chunks [nc0, nc0 + nnc_lim) of the layer into TMEM columns [0, 256 * nnc); facc[0] fires when all of
// them are complete.
__device__ __forceinline__ void tc_mma(const PlanParams& P, Ctx& c, const LayerRec& ly, int nc0 = 0, int nnc_lim = 1 << 30,
                                       int kc0 = 0, int kc_lim = 1 << 30) {
  const int nkc = min(ly.Kpad / kKch, kc0 + kc_lim);
  const int nnc = min((ly.Npad + kNch - 1) / kNch - nc0, nnc_lim);
  const uint32_t sbase = ptx::smem_u32(c.stage_base);
  for (int kc = kc0; kc < nkc; ++kc) {
    const uint32_t as = mma_wait_a(c);
    if (kc == kc0) TDMPC2_TRACE(P, c, 2);
    for (int nc = nc0; nc < nc0 + nnc; ++nc) {
      const uint32_t ws = mma_wait_w(c);
      ptx::tc_fence_after();
      const int ncols = min(kNch, ly.Npad - nc * kNch);
      mma_stage(c.tmem_base + (nc - nc0) * kNch, sbase + as * kASlotBytes, sbase + kWRingOff + ws * c.w_stride, c.w_lo_off,
                ptx::make_idesc_f16(c.cg2 ? 2 * kTileM : kTileM, ncols), kc == kc0, c.cg2 != 0, P.passes == 1);
      if (ptx::elect_one()) {
        if (c.cg2) ptx::umma_commit_2sm(&c.w_empty[ws]);
        else ptx::umma_commit(&c.w_empty[ws]);     // frees the W slot when these MMAs retire
      }
      ++c.mw_it;
    }
    if (ptx::elect_one()) {
      if (c.cg2) ptx::umma_commit_2sm(&c.a_empty[as]);
      else ptx::umma_commit(&c.a_empty[as]);
    }
    ++c.ma_it;
  }
  if (ptx::elect_one()) {
    if (c.cg2) ptx::umma_commit_2sm(&c.facc[0]);
    else ptx::umma_commit(&c.facc[0]);
  }
  TDMPC2_TRACE(P, c, 3);
}

// Head layers (plain Linear outputs, Npad <= 256) with a long reduction: the accumulator is handed to the epilogue every
// `kseg` K-chunks, alternating between two TMEM buffers (columns [0,256) and [256,512)), and the epilogue adds the
// segments in fp32 registers with round-to-nearest.  Why: tcgen05's accumulate step rounds TOWARD ZERO (measured:
// scripts/micro/mma_rounding.py), which shrinks a long accumulation by ~n_MMA * 2^-25 relative; LayerNorm removes a
// uniform shrink from the hidden layers, but the heads' logits have no LayerNorm behind them.
__device__ __forceinline__ int head_segments(const PlanParams& P, const LayerRec& ly) {
  const int nkc = ly.Kpad / kKch;
  return (P.head_kseg > 0 && nkc > P.head_kseg) ? (nkc + P.head_kseg - 1) / P.head_kseg : 1;
}
__device__ __forceinline__ void tc_mma_head_seg(const PlanParams& P, Ctx& c, const LayerRec& ly, int nseg) {
  const int nkc = ly.Kpad / kKch;
  const uint32_t sbase = ptx::smem_u32(c.stage_base);
  const uint32_t idesc = ptx::make_idesc_f16(c.cg2 ? 2 * kTileM : kTileM, ly.Npad);
  for (int seg = 0; seg < nseg; ++seg) {
    const int b = seg & 1;
    if (seg >= 2) {                                    // buffer b was drained (segment seg - 2)
      uint32_t& it = b ? c.d_it : c.a_it;
      ptx::mbar_wait(&c.acc_empty[b], it & 1);
      ++it;
      ptx::tc_fence_after();
    }
    const int kc0 = seg * P.head_kseg, kc1 = min(nkc, kc0 + P.head_kseg);
    for (int kc = kc0; kc < kc1; ++kc) {
      const uint32_t as = mma_wait_a(c);
      const uint32_t ws = mma_wait_w(c);
      ptx::tc_fence_after();
      mma_stage(c.tmem_base + b * kNch, sbase + as * kASlotBytes, sbase + kWRingOff + ws * c.w_stride, c.w_lo_off, idesc,
                kc == kc0, c.cg2 != 0, P.passes == 1);
      if (ptx::elect_one()) {
        if (c.cg2) { ptx::umma_commit_2sm(&c.w_empty[ws]); ptx::umma_commit_2sm(&c.a_empty[as]); }
        else { ptx::umma_commit(&c.w_empty[ws]); ptx::umma_commit(&c.a_empty[as]); }
      }
      ++c.mw_it; ++c.ma_it;
    }
    if (ptx::elect_one()) {
      if (c.cg2) ptx::umma_commit_2sm(&c.facc[b]);
      else ptx::umma_commit(&c.facc[b]);
    }
  }
}

// ------------------------------------------------------------------------------------ SIMT engine: GEMM -> raw scratch
// Same operands, plain fp32 FFMA on CUDA cores (exact products of the split operands).
__device__ __forceinline__ void gemm_simt(const PlanParams& P, Ctx& c, const LayerRec& ly, int srcbuf, int kc0 = 0) {
  constexpr int BN = 64, BK = 32;
  float* sA = reinterpret_cast<float*>(c.stage_base);          // [BK][128+4]
  float* sW = sA + BK * (kTileM + 4);                          // [BK][BN+4]
  const __half* a_hi = plane_ptr(P, c.slot, srcbuf, 0);
  const __half* a_lo = plane_ptr(P, c.slot, srcbuf, 1);
  const int pitch = plane_pitch(P, srcbuf);
  const int tid = threadIdx.x;
  const int tr = ((tid & 255) / 16) * 8, tc = (tid % 16) * 4;   // 8 rows x 4 cols per thread (threads 0..255)
  float* rawbase = raw_ptr(P, c.slot);
  for (int n0 = 0; n0 < ly.Npad; n0 += BN) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = kc0 * kKch; k0 < ly.Kpad; k0 += BK) {
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
      if (tid < 256) {
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
      }
      __syncthreads();
    }
    if (tid < 256) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          __stcg(rawbase + static_cast<size_t>(tr + i) * P.NpadMax + n0 + tc + j, acc[i][j]);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------ wide path: row phases (warp per row)
__device__ __forceinline__ float ln_act_lane(float y, bool valid, int act) {
  if (act == EPI_LN_MISH) return mish_f(y);
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

// LayerNorm (+ Mish | SimNorm) over raw rows; one warp per row, lane-strided columns; raw re-read from L2 per pass.
__device__ __forceinline__ void rows_ln_act(const PlanParams& P, Ctx& c, const LayerRec& ly, const EpiArgs& ea) {
  const float* rawbase = raw_ptr(P, c.slot);
  const int N = ly.N;
  const float invN = 1.f / static_cast<float>(N);
  __half* dhi = ea.dstbuf >= 0 ? plane_ptr(P, c.slot, ea.dstbuf, 0) : nullptr;
  __half* dlo = ea.dstbuf >= 0 ? plane_ptr(P, c.slot, ea.dstbuf, 1) : nullptr;
  const int pitch = ea.dstbuf >= 0 ? plane_pitch(P, ea.dstbuf) : 0;
  const float inv_scale = ly.inv_scale;
  const float* bias = ly.bias; const float* lg = ly.ln_g; const float* lb = ly.ln_b;
  const int ncolj = (N + 31) / 32;
  for (int r = c.warp; r < kTileM; r += kWarps) {
    const float* rr = rawbase + static_cast<size_t>(r) * P.NpadMax;
    float s = 0.f;
    for (int col = c.lane; col < N; col += 32) s += fmaf(__ldcg(rr + col), inv_scale, bias[col]);
    const float mean = warp_sum(s) * invN;
    float sq = 0.f;
    for (int col = c.lane; col < N; col += 32) {
      const float d = fmaf(__ldcg(rr + col), inv_scale, bias[col]) - mean;
      sq = fmaf(d, d, sq);
    }
    const float var = warp_sum(sq) * invN;
    const float rstd = 1.f / sqrtf(var + 1e-5f);   // nn.LayerNorm eps (layers.py:101)
    const int orow = ea.rowmap ? ea.rowmap[r] : r;
    for (int j = 0; j < ncolj; ++j) {
      const int col = c.lane + 32 * j;
      const bool valid = col < N;
      float y = 0.f;
      if (valid) y = (fmaf(__ldcg(rr + col), inv_scale, bias[col]) - mean) * rstd * lg[col] + lb[col];
      y = ln_act_lane(y, valid, ea.kind);
      if (valid) {
        if (dhi) split_store(dhi + static_cast<size_t>(r) * pitch + ea.dst_col0 + col,
                             dlo + static_cast<size_t>(r) * pitch + ea.dst_col0 + col, y);
        if (ea.out_f32 && orow >= 0) ea.out_f32[static_cast<size_t>(orow) * ea.out_pitch + col] = y;
      }
    }
  }
}

// Head output row -> smem row buffer: out[col] = raw*inv_scale + bias (plain Linear, no LN).
__device__ __forceinline__ void head_row_to_smem(const PlanParams& P, Ctx& c, const LayerRec& ly, int r, float* buf) {
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

template <bool EPISODIC>
__device__ __forceinline__ void rows_head(const PlanParams& P, Ctx& c, const LayerRec& ly, const EpiArgs& ea) {
  float* myrow = c.rowbuf + c.warp * kMaxHeadCols;
  __half* xhi = plane_ptr(P, c.slot, BUF_X, 0);
  __half* xlo = plane_ptr(P, c.slot, BUF_X, 1);
  for (int r = c.warp; r < kTileM; r += kWarps) {
    if (ea.kind == EPI_RAW) {
      const int orow = ea.rowmap ? ea.rowmap[r] : r;
      const float* rr = raw_ptr(P, c.slot) + static_cast<size_t>(r) * P.NpadMax;
      if (orow >= 0)
        for (int col = c.lane; col < ly.N; col += 32)
          ea.out_f32[static_cast<size_t>(orow) * ea.out_pitch + col] = fmaf(__ldcg(rr + col), ly.inv_scale, ly.bias[col]);
      continue;
    }
    head_row_to_smem(P, c, ly, r, myrow);
    if (EPISODIC && ea.kind == EPI_TERM) {
      if (c.lane == 0) term_commit(c, r, myrow[0]);
    } else if (ea.kind == EPI_TWOHOT) {
      const float v = two_hot_inv_row(P, c, myrow);
      if (c.lane == 0) head_commit<EPISODIC>(P, c, ea, r, v);
    } else if (ea.kind == EPI_PI) {
      const RowMap rm = map_row(P, ea.tile, r);
      const int e = rm.env < 0 ? 0 : rm.env, idx = rm.env < 0 ? 0 : rm.idx;
      const int task = P.task ? P.task[e] : 0;
      for (int a = c.lane; a < P.A; a += 32) {
        const float eps = ea.eps_base ? __ldcs(&ea.eps_base[(static_cast<size_t>(e) * ea.eps_rows + idx) * P.A + a])
                                      : noise_pi_at(P, e, idx, a);        // eps_base == nullptr: in-kernel noise (ITER)
        const float act = pi_action(P, myrow[a], myrow[P.Apad + a], eps, task, a);
        const size_t o = static_cast<size_t>(r) * P.KpadX + P.L + P.T + a;
        split_store(xhi + o, xlo + o, act);
        if (ea.act_out && rm.env >= 0)
          ea.act_out[((static_cast<size_t>(e) * P.H + ea.t_out) * P.P + idx) * P.A + a] = act;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------ fused path: TMEM epilogues
// Thread-per-row: epilogue warp e (0..15) owns TMEM lanes 32*(e&3).. and the column group (e>>2).
struct EpiThread {
  int q, grp, row;
  uint32_t taddr;     // TMEM address of this thread's lane, column 0
};
__device__ __forceinline__ EpiThread epi_thread(const Ctx& c) {
  EpiThread t;
  const int e = c.warp - kEpiWarp0;
  t.q = c.warp & 3;                   // a warp may only touch the TMEM lane quarter (warp id % 4)
  t.grp = e >> 2; t.row = t.q * 32 + c.lane;
  t.taddr = c.tmem_base + (static_cast<uint32_t>(t.q * 32) << 16);
  return t;
}
__device__ __forceinline__ void epi_stage_vectors(Ctx& c, const LayerRec& ly, bool with_ln) {
  const int t = threadIdx.x - kEpiWarp0 * 32;
  for (int i = t; i < ly.Npad; i += kEpiThreads) {
    c.vec[i] = ly.bias[i];
    if (with_ln) { c.vec[kFusedMaxN + i] = ly.ln_g[i]; c.vec[2 * kFusedMaxN + i] = ly.ln_b[i]; }
  }
  epi_bar_sync();
}

// Fast-path activation math for the fused epilogue.  __expf = ex2.approx(x*log2e) (rel. error ~2^-22 + |x|*6e-8),
// __fdividef = rcp.approx * n (~1.5 ulp): Mish stays within ~1e-6 relative of the exact value.
__device__ __forceinline__ float ex2_ftz(float x) {   // single MUFU.EX2
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {   // single MUFU.RCP
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float exp_fast(float x) { return ex2_ftz(x * 1.4426950408889634f); }
// Packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100): two elements per instruction on the FP32 pipe,
// which is what bounds the fused epilogue.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
// Two Mish values; the two divisions share one reciprocal, 1/d.x = d.y / (d.x d.y) (halves the MUFU.RCP count; it
// measured neutral, the epilogue is latency-bound).  x is clamped at 10: n/(n+2) already rounds to 1 there, and
// d.x * d.y stays below 2.4e17.
__device__ __forceinline__ float2 mish_fast2(float2 x) {
  const float2 a = f2(fminf(x.x, 10.f), fminf(x.y, 10.f));
  const float2 z = __fmul2_rn(a, f2s(1.4426950408889634f));
  const float2 e = f2(ex2_ftz(z.x), ex2_ftz(z.y));
  const float2 n = __fmul2_rn(e, __fadd2_rn(e, f2s(2.f)));
  const float2 d = __fadd2_rn(n, f2s(2.f));
  const float2 r = __fmul2_rn(f2s(rcp_ftz(d.x * d.y)), f2(d.y, d.x));
  return __fmul2_rn(x, __fmul2_rn(n, r));
}
__device__ __forceinline__ float mish_fast(float x) {
  const float e = exp_fast(fminf(x, 30.f));        // x > 30: n/(n+2) == 1 in fp32 already; keeps e*e finite
  const float n = e * (e + 2.f);
  return x * (n * rcp_ftz(n + 2.f));
}

// Pass 2 of the fused LayerNorm epilogue, specialised for the common case (whole 32-column blocks, planes out
// through TMA stores, no fp32 side output): one block = two x16 TMEM loads, packed fp32x2 math, four 16-byte
// swizzled smem stores per plane, then the block is handed to the TMA unit.
template <int KIND, bool FAST = false>   // FAST: single-pass mode (PlanParams::passes == 1), the lo plane is neither computed nor stored
__device__ __forceinline__ void epi_pass2_fast(const PlanParams& P, Ctx& c, const EpiThread& et, const EpiArgs& ea, int cb,
                                               int nvalid, float inv_scale, float rstd, float nmr) {
  const float* sb = c.vec; const float* sg = c.vec + kFusedMaxN; const float* sbe = c.vec + 2 * kFusedMaxN;
  // staging tiles: two per group aliasing the idle W ring -- or, when the W ring is busy prefetching the next layer
  // (c.wpf), ONE per group in the idle A ring, re-used once the previous store has read it
  uint8_t* stg = c.wpf ? c.stage_base + et.grp * kStgBuf : c.stage_base + kWRingOff + et.grp * (2 * kStgBuf);
  const bool lead_warp = (et.q == 0);   // one elected lane of the group's first warp issues, commits and waits (cheap in SASS: no waterfall)
  const CUtensorMap* tmD = (ea.dstbuf == BUF_X) ? &P.tmXs : &P.tmHs;
  const uint32_t swz = static_cast<uint32_t>((et.row >> 1) & 3);
  const int row_hi = plane_row0(P, c.slot, ea.dstbuf, 0), row_lo = plane_row0(P, c.slot, ea.dstbuf, 1);
  const float2 inv2 = f2s(inv_scale), rstd2 = f2s(rstd), nmr2 = f2s(nmr);
  const int nblk = nvalid >> 5;
  for (int blk = 0; blk < nblk; ++blk) {
    const int c0 = cb + (blk << 5);
    uint8_t* buf = c.wpf ? stg : stg + (blk & 1) * kStgBuf;
    const uint32_t rowaddr = ptx::smem_u32(buf) + static_cast<uint32_t>(et.row) * 64u;
    if (!c.wpf && blk >= 2) {                                     // buffer reuse: its previous store must have read it
      if (lead_warp && ptx::elect_one()) ptx::bulk_wait_read<1>();
      group_bar_sync(et.grp);
    }
#pragma unroll
    for (int sub = 0; sub < 32; sub += 16) {
      uint32_t v[16];
      ptx::tmem_ld_32x16(et.taddr + c0 + sub, v);
      ptx::tmem_ld_wait();
      uint32_t hw[8], lw[8];
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
        // SimNorm: softmax over groups of 8 consecutive columns (layers.py:84-88)
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
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a0 = __uint_as_float(v[2 * i]), a1 = __uint_as_float(v[2 * i + 1]);
        const __half2 h2 = __floats2half2_rn(a0, a1);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
        if (!FAST) {
          const float2 hf = __half22float2(h2);
          const float2 df = __fadd2_rn(f2(a0, a1), f2(-hf.x, -hf.y));
          const __half2 l2 = __floats2half2_rn(df.x, df.y);
          lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
        }
      }
      if (c.wpf && sub == 0 && blk >= 1) {
        // single staging tile: the previous block's store was issued a whole block of math ago, so this wait is short
        if (lead_warp && ptx::elect_one()) ptx::bulk_wait_read<0>();
        group_bar_sync(et.grp);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t off = ((static_cast<uint32_t>((sub >> 3) + i) ^ swz) << 4);
        ptx::st_shared_v4(rowaddr + off, hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
        if (!FAST) ptx::st_shared_v4(rowaddr + kStgPlane + off, lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
      }
    }
    ptx::fence_proxy_async_smem();
    group_bar_sync(et.grp);
    if (lead_warp && ptx::elect_one()) {
      if (P.l2hint & 2) {     // evict-last on the activation-plane stores (re-read by the next layer, then overwritten):
                              // DRAM write-back of the scratch 2.11 -> 1.43 GB per c2 iteration, -0.9 % per plan under the power cap
        ptx::tma_store_2d_hint(tmD, buf, ea.dst_col0 + c0, row_hi, ptx::kL2EvictLast);
        if (!FAST) ptx::tma_store_2d_hint(tmD, buf + kStgPlane, ea.dst_col0 + c0, row_lo, ptx::kL2EvictLast);
      } else {
        ptx::tma_store_2d(tmD, buf, ea.dst_col0 + c0, row_hi);
        if (!FAST) ptx::tma_store_2d(tmD, buf + kStgPlane, ea.dst_col0 + c0, row_lo);
      }
      ptx::bulk_commit();
    }
  }
  if (lead_warp && ptx::elect_one()) ptx::bulk_wait<0>();                               // stores performed before the layer is published
}

// bias + LayerNorm + (Mish | SimNorm); planes and/or fp32 rows out.  16 epilogue warps: 4 lane quarters x 4
// column groups; a group owns whole 64-column blocks (the TMA-store granule).
// Pass 1 reads the accumulator row once for shifted first/second moments (groups merged with Chan's
// parallel-variance formula), pass 2 re-reads it 16 columns at a time, normalises, activates and emits.
__device__ __forceinline__ void epi_ln_fused(const PlanParams& P, Ctx& c, const LayerRec& ly, const EpiArgs& ea) {
  const EpiThread et = epi_thread(c);
  const int N = ly.N;
  const int nblocks = ly.Npad / 64;
  const int bpg = (nblocks + kEpiGroups - 1) / kEpiGroups;        // 64-column blocks per group
  const int cb = et.grp * bpg * 64;                               // this thread's columns: [cb, cb + ncols) & < N
  const int ncols = max(0, min(bpg * 64, ly.Npad - cb));
  const int nvalid = max(0, min(N - cb, ncols));
  const float inv_scale = ly.inv_scale;
  epi_stage_vectors(c, ly, true);
  const float* sb = c.vec; const float* sg = c.vec + kFusedMaxN; const float* sbe = c.vec + 2 * kFusedMaxN;
  if (nvalid > 0) {
    const long long tw = prof_clock();
    ptx::mbar_wait_long(&c.facc[0], c.fph0);
    c.pf2 += prof_clock() - tw;
  }
  const bool tr0 = (et.grp == 0 && et.q == 0 && c.lane == 0), tr3 = (et.grp == 3 && et.q == 0 && c.lane == 0);
  if (tr0) TDMPC2_TRACE(P, c, 4);
  if (tr3) TDMPC2_TRACE(P, c, 10);
  ptx::tc_fence_after();
  // ---- pass 1: shifted moments of this group's columns
  float x0 = 0.f, s = 0.f, q = 0.f;
  float2 s2 = f2s(0.f), q2 = f2s(0.f);
  const bool p1_fast = (nvalid > 0) && ((nvalid & 31) == 0);
  if (p1_fast) {
    // whole 32-column chunks: x32 loads, four independent packed accumulator chains
    float2 sa = f2s(0.f), sbb = f2s(0.f), qa = f2s(0.f), qb = f2s(0.f);
    const float2 inv2 = f2s(inv_scale);
    for (int c0 = cb; c0 < cb + nvalid; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(et.taddr + c0, v);
      ptx::tmem_ld_wait();
      if (c0 == cb) x0 = fmaf(__uint_as_float(v[0]), inv_scale, sb[c0]);
      const float2 nx0 = f2s(-x0);
#pragma unroll
      for (int i4 = 0; i4 < 32; i4 += 4) {
        const float4 b4 = lds128(sb + c0 + i4);
        const float2 xa = __ffma2_rn(f2(__uint_as_float(v[i4]), __uint_as_float(v[i4 + 1])), inv2, __fadd2_rn(f2(b4.x, b4.y), nx0));
        const float2 xb = __ffma2_rn(f2(__uint_as_float(v[i4 + 2]), __uint_as_float(v[i4 + 3])), inv2, __fadd2_rn(f2(b4.z, b4.w), nx0));
        sa = __fadd2_rn(sa, xa); qa = __ffma2_rn(xa, xa, qa);
        sbb = __fadd2_rn(sbb, xb); qb = __ffma2_rn(xb, xb, qb);
      }
    }
    s2 = __fadd2_rn(sa, sbb); q2 = __fadd2_rn(qa, qb);
  }
  for (int c0 = cb; c0 < cb + (p1_fast ? 0 : nvalid); c0 += 16) {
    uint32_t v[16];
    ptx::tmem_ld_32x16(et.taddr + c0, v);
    ptx::tmem_ld_wait();
    if (c0 == cb) x0 = fmaf(__uint_as_float(v[0]), inv_scale, sb[c0]);
    const bool full = (c0 + 16 <= N);
#pragma unroll
    for (int i4 = 0; i4 < 16; i4 += 4) {
      const float4 b4 = lds128(sb + c0 + i4);
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
      if (full) {
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const float2 xv = __ffma2_rn(f2(__uint_as_float(v[i4 + j]), __uint_as_float(v[i4 + j + 1])), f2s(inv_scale),
                                       f2(bb[j] - x0, bb[j + 1] - x0));
          s2 = __fadd2_rn(s2, xv);
          q2 = __ffma2_rn(xv, xv, q2);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = fmaf(__uint_as_float(v[i4 + j]), inv_scale, bb[j]) - x0;
          if (c0 + i4 + j < N) { s += d; q = fmaf(d, d, q); }
        }
      }
    }
  }
  s += s2.x + s2.y;
  q += q2.x + q2.y;
  {
    const float n_g = static_cast<float>(nvalid);
    c.part[et.grp * kTileM + et.row] = nvalid > 0 ? x0 + s / n_g : 0.f;                       // group mean
    c.part[(kEpiGroups + et.grp) * kTileM + et.row] = nvalid > 0 ? q - s * s / n_g : 0.f;     // group M2
  }
  if (tr0) TDMPC2_TRACE(P, c, 5);
  epi_bar_sync();
  if (tr0) TDMPC2_TRACE(P, c, 6);
  float mean = 0.f, rstd;
  {
    float m2 = 0.f, cnt = 0.f;
#pragma unroll
    for (int g = 0; g < kEpiGroups; ++g) {
      const int gb = g * bpg * 64;
      const float ng = static_cast<float>(max(0, min(N - gb, min(bpg * 64, ly.Npad - gb))));
      if (ng > 0.f) {
        const float mg = c.part[g * kTileM + et.row], m2g = c.part[(kEpiGroups + g) * kTileM + et.row];
        const float tot = cnt + ng, delta = mg - mean;
        mean += delta * (ng / tot);
        m2 += m2g + delta * delta * (cnt * ng / tot);
        cnt = tot;
      }
    }
    rstd = rsqrtf(m2 / static_cast<float>(N) + 1e-5f);            // nn.LayerNorm eps (layers.py:101), biased variance
  }
  const float nmr = -mean * rstd;
  if (ea.dstbuf >= 0 && !ea.out_f32 && (N % 32 == 0) && (ea.dst_col0 % 32 == 0)) {
    if (P.passes == 1) {
      if (ea.kind == EPI_LN_MISH) epi_pass2_fast<EPI_LN_MISH, true>(P, c, et, ea, cb, nvalid, inv_scale, rstd, nmr);
      else epi_pass2_fast<EPI_LN_SIMNORM, true>(P, c, et, ea, cb, nvalid, inv_scale, rstd, nmr);
    } else {
      if (ea.kind == EPI_LN_MISH) epi_pass2_fast<EPI_LN_MISH>(P, c, et, ea, cb, nvalid, inv_scale, rstd, nmr);
      else epi_pass2_fast<EPI_LN_SIMNORM>(P, c, et, ea, cb, nvalid, inv_scale, rstd, nmr);
    }
    return;
  }
  // ---- pass 2 (general path): normalise, activate, emit
  __half* dhi = ea.dstbuf >= 0 ? plane_ptr(P, c.slot, ea.dstbuf, 0) : nullptr;
  __half* dlo = ea.dstbuf >= 0 ? plane_ptr(P, c.slot, ea.dstbuf, 1) : nullptr;
  const int pitch = ea.dstbuf >= 0 ? plane_pitch(P, ea.dstbuf) : 0;
  const int orow = ea.rowmap ? ea.rowmap[et.row] : et.row;
  // Plane output goes through a 128B-swizzled smem tile [128 rows x 64 cols] per plane and a TMA store
  // (thread-per-row global stores would touch 32 cache lines per instruction).  The staging tiles alias
  // the operand pipeline stages, which are idle here: every MMA of this layer has retired.
  const bool use_tma = (dhi != nullptr) && (N % 32 == 0) && (ea.dst_col0 % 32 == 0);
  uint8_t* stg = c.stage_base + kWRingOff + et.grp * (2 * kStgBuf);   // per group: 2 buffers x (hi 8 KiB | lo 8 KiB)
  const bool leader = (et.q == 0) && (c.lane == 0);
  const CUtensorMap* tmD = (ea.dstbuf == BUF_X) ? &P.tmXs : &P.tmHs;
  const uint32_t swz = static_cast<uint32_t>((et.row >> 1) & 3);     // 64-byte swizzle: chunk ^= (row / 2) % 4
  for (int c0 = cb; c0 < cb + nvalid; c0 += 16) {
    const int sub = (c0 - cb) & 31;                               // 0 | 16 within the 32-column block
    const int blk = (c0 - cb) >> 5;
    uint8_t* buf = stg + (blk & 1) * kStgBuf;
    const uint32_t rowaddr = ptx::smem_u32(buf) + static_cast<uint32_t>(et.row) * 64u;
    if (use_tma && sub == 0 && blk >= 2) {                        // buffer reuse: its previous store must have read it
      if (leader) ptx::bulk_wait_read<1>();
      group_bar_sync(et.grp);
    }
    uint32_t v[16];
    ptx::tmem_ld_32x16(et.taddr + c0, v);
    ptx::tmem_ld_wait();
    const bool full = (c0 + 16 <= N);
    float y[16];
#pragma unroll
    for (int i4 = 0; i4 < 16; i4 += 4) {
      const float4 b4 = lds128(sb + c0 + i4);
      const float4 g4 = lds128(sg + c0 + i4);
      const float4 e4 = lds128(sbe + c0 + i4);
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, ee[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const float2 x = __ffma2_rn(f2(__uint_as_float(v[i4 + j]), __uint_as_float(v[i4 + j + 1])), f2s(inv_scale), f2(bb[j], bb[j + 1]));
        const float2 u = __ffma2_rn(x, f2s(rstd), f2s(nmr));
        float2 t = __ffma2_rn(u, f2(gg[j], gg[j + 1]), f2(ee[j], ee[j + 1]));
        if (ea.kind == EPI_LN_MISH) t = mish_fast2(t);
        y[i4 + j] = t.x; y[i4 + j + 1] = t.y;
      }
    }
    if (!full) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i >= N) y[i] = (ea.kind == EPI_LN_MISH) ? 0.f : -CUDART_INF_F;
    }
    if (ea.kind == EPI_LN_MISH) {
    } else {
      // SimNorm: softmax over groups of 8 consecutive columns (layers.py:84-88)
#pragma unroll
      for (int g0 = 0; g0 < 16; g0 += 8) {
        float m = y[g0];
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmaxf(m, y[g0 + i]);
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { y[g0 + i] = exp_fast(y[g0 + i] - m); t += y[g0 + i]; }
        const float rt = rcp_ftz(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[g0 + i] *= rt;
      }
    }
    if (dhi) {
      if (use_tma || (full && ((ea.dst_col0 & 7) == 0))) {
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // |y| <= sqrt(N) max|g| + max|b| after LayerNorm (Mish and SimNorm only shrink it): far inside fp16 range
          const float a0 = y[2 * i], a1 = y[2 * i + 1];
          const __half2 h2 = __floats2half2_rn(a0, a1);
          const float2 hf = __half22float2(h2);
          const float2 df = __fadd2_rn(f2(a0, a1), f2(-hf.x, -hf.y));
          const __half2 l2 = __floats2half2_rn(df.x, df.y);
          hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
          lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        if (use_tma) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t chunk = static_cast<uint32_t>((sub >> 3) + i);          // 16-byte chunk index in the 64 B row
            const uint32_t off = ((chunk ^ swz) << 4);
            ptx::st_shared_v4(rowaddr + off, hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
            ptx::st_shared_v4(rowaddr + kStgPlane + off, lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
          }
          if (sub == 16) {                                        // 32-column block complete: hand it to the TMA unit
            ptx::fence_proxy_async_smem();
            group_bar_sync(et.grp);
            if (leader) {
              const int col = ea.dst_col0 + c0 - 16;
              ptx::tma_store_2d(tmD, buf, col, plane_row0(P, c.slot, ea.dstbuf, 0));
              ptx::tma_store_2d(tmD, buf + kStgPlane, col, plane_row0(P, c.slot, ea.dstbuf, 1));
              ptx::bulk_commit();
            }
          }
        } else {
          __half* ph = dhi + static_cast<size_t>(et.row) * pitch + ea.dst_col0 + c0;
          __half* pl = dlo + static_cast<size_t>(et.row) * pitch + ea.dst_col0 + c0;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            __stcg(reinterpret_cast<uint4*>(ph) + i, make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]));
            __stcg(reinterpret_cast<uint4*>(pl) + i, make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]));
          }
        }
      } else {
        __half* ph = dhi + static_cast<size_t>(et.row) * pitch + ea.dst_col0 + c0;
        __half* pl = dlo + static_cast<size_t>(et.row) * pitch + ea.dst_col0 + c0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < N) split_store(ph + i, pl + i, y[i]);
      }
    }
    if (ea.out_f32 && orow >= 0) {
      float* po = ea.out_f32 + static_cast<size_t>(orow) * ea.out_pitch + c0;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < N) po[i] = y[i];
    }
  }
  if (tr0) TDMPC2_TRACE(P, c, 7);
  if (tr3) TDMPC2_TRACE(P, c, 11);
  if (use_tma && leader) ptx::bulk_wait<0>();                     // stores performed before the layer is published
  if (tr0) TDMPC2_TRACE(P, c, 8);
}

// Head epilogues (plain Linear outputs, Npad <= 256 so the accumulator is chunk 0 only).
// Two-hot heads with <= 128 bins and pi heads with <= 64 action dims are spread over all four column groups
// (32 bins / 16 action dims per group); anything larger runs on column group 0 alone.
// Can this head's epilogue take K-segmented accumulators (segments summed in registers)?  The common shapes only:
// two-hot heads with <= 128 bins, pi heads with <= 64 action dims, the termination head.
__device__ __forceinline__ bool head_seg_ok(const PlanParams& P, int kind) {
  return (kind == EPI_TWOHOT && P.B <= 32 * kEpiGroups) || (kind == EPI_PI && P.A <= 16 * kEpiGroups) || kind == EPI_TERM;
}

// `nseg` > 1: the accumulator arrives in K-segments alternating between TMEM buffers 0 / 1 (tc_mma_head_seg); each
// thread adds the segments of its own columns in registers.  v[0..32): this thread's 32 columns at column offset col0
// (two-hot: bins 32*grp..; pi: 16 mean logits then 16 log_std logits; termination: column 0).
template <bool EPISODIC>
__device__ __forceinline__ void epi_head_fused(const PlanParams& P, Ctx& c, const LayerRec& ly, const EpiArgs& ea, int nseg) {
  const EpiThread et = epi_thread(c);
  const float inv_scale = ly.inv_scale;
  epi_stage_vectors(c, ly, false);
  const float* sb = c.vec;
  if (ea.kind == EPI_TWOHOT) {
    // bins -> smem (second vector slot)
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < P.B; i += kEpiThreads) c.vec[kFusedMaxN + i] = P.bins[i];
    epi_bar_sync();
  }
  const bool wide_twohot = (ea.kind == EPI_TWOHOT) && (P.B <= 32 * kEpiGroups);
  const bool wide_pi = (ea.kind == EPI_PI) && (P.A <= 16 * kEpiGroups);
  float v[32];                                         // this thread's accumulator values (scaled domain)
  if (nseg > 1) {
    // every epilogue warp takes part in the hand-off protocol, whether or not it owns columns of this head
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    const bool mine = wide_twohot || wide_pi || et.grp == 0;
    for (int seg = 0; seg < nseg; ++seg) {
      const int b = seg & 1;
      const uint32_t ph = (b ? c.fph1 : c.fph0) ^ static_cast<uint32_t>((seg >> 1) & 1);
      ptx::mbar_wait_long(&c.facc[b], ph);
      ptx::tc_fence_after();
      if (mine) {
        const uint32_t ta = et.taddr + b * kNch;
        // two 16-column loads (keeps the live registers at acc[32] + 16): columns [c_lo, +16) and [c_hi, +16)
        const uint32_t c_lo = wide_pi ? 16 * et.grp : (wide_twohot ? 32 * et.grp : 0);
        const uint32_t c_hi = wide_pi ? P.Apad + 16 * et.grp : c_lo + 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t a0[16];
          ptx::tmem_ld_32x16(ta + (h ? c_hi : c_lo), a0);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[16 * h + i] += __uint_as_float(a0[i]);
        }
      }
      if (seg + 2 < nseg) {                            // the buffer is reused: tell the MMA issuer it has been read
        ptx::tc_fence_before();
        epi_bar_sync();
        if (threadIdx.x == kEpiWarp0 * 32) {
          if (c.cg2) ptx::mbar_arrive_leader(&c.acc_empty[b]);
          else ptx::mbar_arrive(&c.acc_empty[b]);
        }
      }
    }
    if (!mine) return;
  } else {
    if (!wide_twohot && !wide_pi && et.grp != 0) return;
    {
      const long long tw = prof_clock();
      ptx::mbar_wait_long(&c.facc[0], c.fph0);
      c.pf2 += prof_clock() - tw;
    }
    ptx::tc_fence_after();
  }
  if (wide_twohot) {
    // two_hot_inv (math.py:74-83) with the 4 groups each owning 32 bins: max, then exp-sum and bin-weighted sum,
    // exchanged through smem (part: [0,4) max, [4,8) sum; third array in the unused LN-beta vector slot).
    const float* bins = c.vec + kFusedMaxN;
    float* xch = c.vec + 2 * kFusedMaxN;                // [kEpiGroups][128]
    const int B = P.B, c0 = 32 * et.grp;
    if (nseg == 1) {
      uint32_t t32[32];
      ptx::tmem_ld_32x32(et.taddr + c0, t32);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(t32[i]);
    }
    float (&x)[32] = v;                                // logits, in place
    float m = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      x[i] = (c0 + i < B) ? fmaf(v[i], inv_scale, sb[min(c0 + i, B - 1)]) : -CUDART_INF_F;
      m = fmaxf(m, x[i]);
    }
    c.part[et.grp * kTileM + et.row] = m;
    epi_bar_sync();
#pragma unroll
    for (int g = 0; g < kEpiGroups; ++g) m = fmaxf(m, c.part[g * kTileM + et.row]);
    float ssum = 0.f, acc = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float e = (c0 + i < B) ? exp_fast(x[i] - m) : 0.f;     // exp(-inf - m) would be 0 as well; keep NaN rows NaN
      ssum += e;
      acc = fmaf(e, bins[min(c0 + i, B - 1)], acc);
    }
    c.part[(kEpiGroups + et.grp) * kTileM + et.row] = ssum;
    xch[et.grp * kTileM + et.row] = acc;
    epi_bar_sync();
    if (et.grp == 0) {
      float S = 0.f, Acc = 0.f;
#pragma unroll
      for (int g = 0; g < kEpiGroups; ++g) { S += c.part[(kEpiGroups + g) * kTileM + et.row]; Acc += xch[g * kTileM + et.row]; }
      head_commit<EPISODIC>(P, c, ea, et.row, symexp_f(__fdiv_rn(Acc, S)));
    }
  } else if (EPISODIC && ea.kind == EPI_TERM) {
    // termination head (one output column): group 0's thread of each row updates the row's sticky flag
    if (nseg == 1) {
      uint32_t t16[16];
      ptx::tmem_ld_32x16(et.taddr, t16);
      ptx::tmem_ld_wait();
      v[0] = __uint_as_float(t16[0]);
    }
    term_commit(c, et.row, fmaf(v[0], inv_scale, sb[0]));
  } else if (ea.kind == EPI_TWOHOT) {
    const float* bins = c.vec + kFusedMaxN;
    const int B = P.B;
    float m = -CUDART_INF_F;
    for (int c0 = 0; c0 < B; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(et.taddr + c0, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < B) m = fmaxf(m, fmaf(__uint_as_float(v[i]), inv_scale, sb[c0 + i]));
    }
    float ssum = 0.f, acc = 0.f;
    for (int c0 = 0; c0 < B; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(et.taddr + c0, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < B) {
          const float e = exp_fast(fmaf(__uint_as_float(v[i]), inv_scale, sb[c0 + i]) - m);
          ssum += e;
          acc = fmaf(e, bins[c0 + i], acc);
        }
    }
    head_commit<EPISODIC>(P, c, ea, et.row, symexp_f(__fdiv_rn(acc, ssum)));
  } else if (ea.kind == EPI_PI) {
    const RowMap rm = map_row(P, ea.tile, et.row);
    const int e = rm.env < 0 ? 0 : rm.env, idx = rm.env < 0 ? 0 : rm.idx;
    const int task = P.task ? P.task[e] : 0;
    __half* xhi = plane_ptr(P, c.slot, BUF_X, 0) + static_cast<size_t>(et.row) * P.KpadX + P.L + P.T;
    __half* xlo = plane_ptr(P, c.slot, BUF_X, 1) + static_cast<size_t>(et.row) * P.KpadX + P.L + P.T;
    const float* eps = ea.eps_base ? ea.eps_base + (static_cast<size_t>(e) * ea.eps_rows + idx) * P.A : nullptr;   // nullptr: in-kernel noise
    const int a_begin = wide_pi ? 16 * et.grp : 0;
    const int a_end = wide_pi ? min(P.A, a_begin + 16) : P.A;
    for (int a0 = a_begin; a0 < a_end; a0 += 16) {
      uint32_t vm[16], vs[16];
      if (nseg > 1) {                                   // wide_pi only: one 16-column block per thread, already summed
#pragma unroll
        for (int i = 0; i < 16; ++i) { vm[i] = __float_as_uint(v[i]); vs[i] = __float_as_uint(v[16 + i]); }
      } else {
        ptx::tmem_ld_32x16(et.taddr + a0, vm);            // mean logits, columns [a0, a0+16)
        ptx::tmem_ld_32x16(et.taddr + P.Apad + a0, vs);   // log_std logits, columns [Apad+a0, Apad+a0+16) (aligned)
        ptx::tmem_ld_wait();
      }
      // 16 action columns = 32 B per plane: two 16-byte stores when the whole span lies inside the row
      // (columns past A are zero-weight padding of X, so writing zeros there is harmless)
      const bool vec = (((P.L + P.T) & 7) == 0) && (P.L + P.T + a0 + 16 <= P.KpadX);
      uint32_t hw[8], lw[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float act2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int a = a0 + i + u;
          act2[u] = 0.f;
          if (a < a_end) {
            const float mu = fmaf(__uint_as_float(vm[i + u]), inv_scale, sb[a]);
            const float ls = fmaf(__uint_as_float(vs[i + u]), inv_scale, sb[P.Apad + a]);
            act2[u] = pi_action(P, mu, ls, eps ? __ldcs(&eps[a]) : noise_pi_at(P, e, idx, a), task, a);
            if (!vec) split_store(xhi + a, xlo + a, act2[u]);
            if (ea.act_out && rm.env >= 0)
              ea.act_out[((static_cast<size_t>(e) * P.H + ea.t_out) * P.P + idx) * P.A + a] = act2[u];
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
  } else {  // EPI_RAW
    const int orow = ea.rowmap ? ea.rowmap[et.row] : et.row;
    for (int c0 = 0; c0 < ly.N; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(et.taddr + c0, v);
      ptx::tmem_ld_wait();
      if (orow >= 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < ly.N)
            ea.out_f32[static_cast<size_t>(orow) * ea.out_pitch + c0 + i] = fmaf(__uint_as_float(v[i]), inv_scale, sb[c0 + i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ layers wider than TMEM (48M / 317M presets)
// A LayerNorm layer with Npad > 512 is produced in SUPER-CHUNKS of 512 output columns (the whole TMEM): each is one
// fused-style GEMM -- K outermost, the A K-chunks streamed once per super-chunk, weights split over the CTA pair --
// after which all 16 epilogue warps DRAIN the accumulator: x = acc * 2^-k + bias goes to the slot's fp32 raw scratch,
// stored column-major ([col][128 rows]: a warp's 32 rows of one column are one 128-byte line) while the thread that
// owns (row, column group) keeps running shifted moments of its row.  The MMA issuer waits for the drain (acc_empty)
// before it overwrites TMEM; the TMA producer meanwhile refills the operand rings for the next super-chunk.  After the
// last super-chunk the row statistics are merged across the 4 column groups (Chan) and ONE pass over the raw scratch
// normalises, activates, splits to fp16 hi/lo and sends the planes out through swizzled smem tiles + TMA stores.
// Every thread reads back exactly the raw elements it wrote itself.
struct WideCols { int cb, ncols; };
__device__ __forceinline__ WideCols wide_cols(const LayerRec& ly, int sc, int grp) {
  const int wsc = min(kFusedMaxN, ly.Npad - sc * kFusedMaxN);     // 128 .. 512, multiple of 128
  const int nblocks = wsc / 64;
  const int bpg = (nblocks + kEpiGroups - 1) / kEpiGroups;
  WideCols w;
  w.cb = grp * bpg * 64;
  w.ncols = max(0, min(bpg * 64, wsc - w.cb));
  return w;
}

// LayerNorm affine + activation of 16 consecutive columns (x already holds acc * 2^-k + bias), in place.
template <int KIND>
__device__ __forceinline__ void wide_norm16(float (&y)[16], const float* __restrict__ g, const float* __restrict__ be,
                                            float2 rstd2, float2 nmr2, int nvalid) {
#pragma unroll
  for (int i4 = 0; i4 < 16; i4 += 4) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + i4));
    const float4 e4 = __ldg(reinterpret_cast<const float4*>(be + i4));
    float2 t0 = __ffma2_rn(__ffma2_rn(f2(y[i4], y[i4 + 1]), rstd2, nmr2), f2(g4.x, g4.y), f2(e4.x, e4.y));
    float2 t1 = __ffma2_rn(__ffma2_rn(f2(y[i4 + 2], y[i4 + 3]), rstd2, nmr2), f2(g4.z, g4.w), f2(e4.z, e4.w));
    if (KIND == EPI_LN_MISH) { t0 = mish_fast2(t0); t1 = mish_fast2(t1); }
    y[i4] = t0.x; y[i4 + 1] = t0.y; y[i4 + 2] = t1.x; y[i4 + 3] = t1.y;
  }
  if (nvalid < 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i >= nvalid) y[i] = (KIND == EPI_LN_MISH) ? 0.f : -CUDART_INF_F;
  }
  if (KIND == EPI_LN_SIMNORM) {
    // SimNorm: softmax over groups of 8 consecutive columns (layers.py:84-88)
#pragma unroll
    for (int g0 = 0; g0 < 16; g0 += 8) {
      float m = y[g0];
#pragma unroll
      for (int i = 1; i < 8; ++i) m = fmaxf(m, y[g0 + i]);
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { y[g0 + i] = exp_fast(y[g0 + i] - m); t += y[g0 + i]; }
      const float rt = rcp_ftz(t);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[g0 + i] *= rt;
    }
  }
}

// The normalise pass of a wide layer, common case (planes out through TMA stores, whole 32-column blocks).  Both operand
// rings are idle by now (every MMA of the layer has retired), so the whole 192 KiB of operand smem serve as staging:
// per column group two 16 KiB INPUT buffers in the W ring -- 32 raw columns x 128 rows fp32, exactly one contiguous
// block of the column-major raw scratch, fetched with one bulk copy (cp.async.bulk, mbarrier completion) two blocks
// ahead of its use, which hides the HBM latency of the 317M preset's raw scratch (2 MB per slot) -- and one 16 KiB
// OUTPUT tile (hi | lo planes, 64-byte swizzle) in the A ring that leaves through TMA stores.
// Block b (columns [32 b, 32 b + 32)) belongs to column group b % 4: the raw scratch is global memory, so the
// assignment need not follow the drain's (the stats-merge barrier has made every drained value visible).
template <int KIND>   // EPI_LN_MISH | EPI_LN_SIMNORM
__device__ __forceinline__ void wide_pass2_tma(const PlanParams& P, Ctx& c, const EpiThread& et, const LayerRec& ly, const EpiArgs& ea,
                                               float rstd, float nmr) {
  const int nblk = ly.N >> 5;                                         // N % 32 == 0 on this path
  const float* rawT = raw_ptr(P, c.slot);                            // block b at rawT + b * 32 * 128
  uint8_t* inb = c.stage_base + kWRingOff + et.grp * (2 * kStgBuf);  // 2 x 16 KiB raw blocks
  uint8_t* stg = c.stage_base + et.grp * kStgBuf;                    // 16 KiB output tile (hi 8 KiB | lo 8 KiB)
  uint64_t* bars = c.rawb + et.grp * 2;
  const bool lead_warp = (et.q == 0);   // one elected lane of the group's first warp issues, commits and waits (cheap in SASS: no waterfall)
  const CUtensorMap* tmD = (ea.dstbuf == BUF_X) ? &P.tmXs : &P.tmHs;
  const uint32_t swz = static_cast<uint32_t>((et.row >> 1) & 3);
  const int row_hi = plane_row0(P, c.slot, ea.dstbuf, 0), row_lo = plane_row0(P, c.slot, ea.dstbuf, 1);
  const float2 rstd2 = f2s(rstd), nmr2 = f2s(nmr);
  const uint32_t rowaddr = ptx::smem_u32(stg) + static_cast<uint32_t>(et.row) * 64u;
  uint32_t it = c.rb_it;                              // blocks consumed by this group so far: buffer = it & 1
  if (lead_warp && ptx::elect_one()) {
    for (int j = 0; j < 2; ++j) {
      const int b = et.grp + 4 * j;
      if (b < nblk) {
        const uint32_t ib = (it + j) & 1u;
        ptx::mbar_expect_tx(&bars[ib], kStgBuf);
        ptx::bulk_load(inb + ib * kStgBuf, rawT + static_cast<size_t>(b) * 32 * kTileM, kStgBuf, &bars[ib]);
      }
    }
  }
  for (int b = et.grp; b < nblk; b += kEpiGroups, ++it) {
    const uint32_t ib = it & 1u;
    // the previous block's TMA store has finished reading the output tile
    if (lead_warp && ptx::elect_one()) ptx::bulk_wait_read<0>();
    group_bar_sync(et.grp);
    ptx::mbar_wait(&bars[ib], (it >> 1) & 1u);
    const float* blk = reinterpret_cast<const float*>(inb + ib * kStgBuf) + et.row;     // element (col, row) at blk[col * 128]
#pragma unroll
    for (int sub = 0; sub < 32; sub += 16) {
      float y[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = blk[(sub + i) * kTileM];
      wide_norm16<KIND>(y, ly.ln_g + b * 32 + sub, ly.ln_b + b * 32 + sub, rstd2, nmr2, 16);
      uint32_t hw[8], lw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a0 = y[2 * i], a1 = y[2 * i + 1];
        const __half2 h2 = __floats2half2_rn(a0, a1);
        const float2 hf = __half22float2(h2);
        const float2 df = __fadd2_rn(f2(a0, a1), f2(-hf.x, -hf.y));
        const __half2 l2 = __floats2half2_rn(df.x, df.y);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t off = ((static_cast<uint32_t>((sub >> 3) + i) ^ swz) << 4);
        ptx::st_shared_v4(rowaddr + off, hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
        ptx::st_shared_v4(rowaddr + kStgPlane + off, lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
      }
    }
    ptx::fence_proxy_async_smem();
    group_bar_sync(et.grp);                            // tile complete; every thread of the group is done with raw buffer ib
    if (lead_warp && ptx::elect_one()) {
      ptx::tma_store_2d(tmD, stg, ea.dst_col0 + b * 32, row_hi);
      ptx::tma_store_2d(tmD, stg + kStgPlane, ea.dst_col0 + b * 32, row_lo);
      ptx::bulk_commit();
      const int nb = b + 2 * kEpiGroups;               // refill the buffer just released, two blocks ahead
      if (nb < nblk) {
        ptx::mbar_expect_tx(&bars[ib], kStgBuf);
        ptx::bulk_load(inb + ib * kStgBuf, rawT + static_cast<size_t>(nb) * 32 * kTileM, kStgBuf, &bars[ib]);
      }
    }
  }
  c.rb_it = it;
  if (lead_warp && ptx::elect_one()) ptx::bulk_wait<0>();                                 // stores performed before the layer is published
}

// General fallback of the normalise pass (fp32 row output, ragged N, unaligned destination): plain loads of the raw
// scratch, scalar stores.  Only the encoder's last layer and the diagnostic mode come here.
template <int KIND>
__device__ __forceinline__ void wide_pass2_general(const PlanParams& P, Ctx& c, const EpiThread& et, const LayerRec& ly, const EpiArgs& ea,
                                                   float rstd, float nmr) {
  const int N = ly.N;
  const float* rawT = raw_ptr(P, c.slot) + et.row;
  const bool planes = ea.dstbuf >= 0;
  __half* dhi = planes ? plane_ptr(P, c.slot, ea.dstbuf, 0) + static_cast<size_t>(et.row) * plane_pitch(P, ea.dstbuf) + ea.dst_col0 : nullptr;
  __half* dlo = planes ? plane_ptr(P, c.slot, ea.dstbuf, 1) + static_cast<size_t>(et.row) * plane_pitch(P, ea.dstbuf) + ea.dst_col0 : nullptr;
  const int orow = ea.rowmap ? ea.rowmap[et.row] : et.row;
  float* po = (ea.out_f32 && orow >= 0) ? ea.out_f32 + static_cast<size_t>(orow) * ea.out_pitch : nullptr;
  const float2 rstd2 = f2s(rstd), nmr2 = f2s(nmr);
  for (int c0 = 16 * et.grp; c0 < N; c0 += 16 * kEpiGroups) {        // N <= Npad, the parameter vectors are padded to Npad
    float y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = (c0 + i < ly.Npad) ? __ldcg(rawT + static_cast<size_t>(c0 + i) * kTileM) : 0.f;
    wide_norm16<KIND>(y, ly.ln_g + c0, ly.ln_b + c0, rstd2, nmr2, min(16, N - c0));
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < N) {
        if (po) po[c0 + i] = y[i];
        if (planes) split_store(dhi + c0 + i, dlo + c0 + i, y[i]);
      }
  }
}

// Epilogue warps of a wide layer (EPI_LN_MISH | EPI_LN_SIMNORM | EPI_RAW).
__device__ __forceinline__ void epi_wide(const PlanParams& P, Ctx& c, const LayerRec& ly, const EpiArgs& ea, int nsc, int nseg) {
  const EpiThread et = epi_thread(c);
  const int N = ly.N;
  const float inv_scale = ly.inv_scale;
  const bool is_ln = (ea.kind != EPI_RAW);
  float* rawT = raw_ptr(P, c.slot) + et.row;                        // element (col, row) at rawT[col * 128]
  // running shifted moments of this thread's row over the columns of its group (all super-chunks)
  float x0 = 0.f, s = 0.f, q = 0.f;
  float2 sa = f2s(0.f), sbb = f2s(0.f), qa = f2s(0.f), qb = f2s(0.f);
  bool have_x0 = false;
  const float2 inv2 = f2s(inv_scale);
  const bool tr0 = (et.grp == 0 && et.q == 0 && c.lane == 0);
  for (int sc = 0; sc < nsc; ++sc) {
   const WideCols wc = wide_cols(ly, sc, et.grp);
   for (int seg = 0; seg < nseg; ++seg) {
    // K-segments: the tensor core's fp32 accumulator rounds toward zero on every K = 16 step, an error that grows
    // linearly with the reduction length; flushing the partial sum every P.kseg K-chunks and adding the segments here
    // with round-to-nearest bounds it (at the price of one more drain per segment).
    const bool last_seg = (seg == nseg - 1);
    {
      const long long tw = prof_clock();
      ptx::mbar_wait_sleep(&c.facc[0], c.fph0 ^ static_cast<uint32_t>((sc * nseg + seg) & 1), P.wide_sleep_ns);
      c.pf2 += prof_clock() - tw;
    }
    ptx::tc_fence_after();
    if (tr0 && sc == 0 && seg == 0) TDMPC2_TRACE(P, c, 4);          // first accumulator ready
    if (tr0 && sc == nsc - 1 && last_seg) TDMPC2_TRACE(P, c, 10);   // last accumulator ready
    for (int c0 = wc.cb; c0 < wc.cb + wc.ncols; c0 += 32) {
      const int gcol = sc * kFusedMaxN + c0;
      uint32_t v[32];
      ptx::tmem_ld_32x32(et.taddr + c0, v);
      if (seg == 0) {
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i4 = 0; i4 < 32; i4 += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(ly.bias + gcol + i4));
          const float2 xa = __ffma2_rn(f2(__uint_as_float(v[i4]), __uint_as_float(v[i4 + 1])), inv2, f2(b4.x, b4.y));
          const float2 xb = __ffma2_rn(f2(__uint_as_float(v[i4 + 2]), __uint_as_float(v[i4 + 3])), inv2, f2(b4.z, b4.w));
          v[i4] = __float_as_uint(xa.x); v[i4 + 1] = __float_as_uint(xa.y);
          v[i4 + 2] = __float_as_uint(xb.x); v[i4 + 3] = __float_as_uint(xb.y);
        }
      } else {
        ptx::tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 32; h += 8) {                          // this thread's own partial sums of the earlier segments
          float prev[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) prev[i] = __ldcg(rawT + static_cast<size_t>(gcol + h + i) * kTileM);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[h + i] = __float_as_uint(fmaf(__uint_as_float(v[h + i]), inv_scale, prev[i]));
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) __stcg(rawT + static_cast<size_t>(gcol + i) * kTileM, __uint_as_float(v[i]));
      const int nv = last_seg ? max(0, min(32, N - gcol)) : 0;
      if (is_ln && nv > 0) {
        if (!have_x0) { x0 = __uint_as_float(v[0]); have_x0 = true; }
        if (nv == 32) {
          const float2 nx0 = f2s(-x0);
#pragma unroll
          for (int i4 = 0; i4 < 32; i4 += 4) {
            const float2 da = __fadd2_rn(f2(__uint_as_float(v[i4]), __uint_as_float(v[i4 + 1])), nx0);
            const float2 db = __fadd2_rn(f2(__uint_as_float(v[i4 + 2]), __uint_as_float(v[i4 + 3])), nx0);
            sa = __fadd2_rn(sa, da); qa = __ffma2_rn(da, da, qa);
            sbb = __fadd2_rn(sbb, db); qb = __ffma2_rn(db, db, qb);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nv) { const float d = __uint_as_float(v[i]) - x0; s += d; q = fmaf(d, d, q); }
        }
      }
    }
    if (sc + 1 < nsc || !last_seg) {
      // TMEM may be overwritten by the next segment's / super-chunk's MMAs once every epilogue warp (of both CTAs of a
      // pair) has read its part
      ptx::tc_fence_before();
      epi_bar_sync();
      if (threadIdx.x == kEpiWarp0 * 32) {
        if (c.cg2) ptx::mbar_arrive_leader(&c.acc_empty[0]);
        else ptx::mbar_arrive(&c.acc_empty[0]);
      }
    }
   }
  }
  if (!is_ln) {
    // EPI_RAW (diagnostics): plain Linear output rows; every thread copies out the columns it drained itself
    const int orow = ea.rowmap ? ea.rowmap[et.row] : et.row;
    if (orow >= 0)
      for (int sc = 0; sc < nsc; ++sc) {
        const WideCols wc = wide_cols(ly, sc, et.grp);
        for (int c0 = wc.cb; c0 < wc.cb + wc.ncols; ++c0) {
          const int gcol = sc * kFusedMaxN + c0;
          if (gcol < N) ea.out_f32[static_cast<size_t>(orow) * ea.out_pitch + gcol] = __ldcg(rawT + static_cast<size_t>(gcol) * kTileM);
        }
      }
    return;
  }
  if (tr0) TDMPC2_TRACE(P, c, 5);                                   // every super-chunk drained
  // ---- merge the 4 column groups' statistics (Chan), as in epi_ln_fused
  s += (sa.x + sa.y) + (sbb.x + sbb.y);
  q += (qa.x + qa.y) + (qb.x + qb.y);
  int cnt_g[kEpiGroups];
#pragma unroll
  for (int g = 0; g < kEpiGroups; ++g) {
    int n = 0;
    for (int sc = 0; sc < nsc; ++sc) {
      const WideCols wc = wide_cols(ly, sc, g);
      n += max(0, min(wc.ncols, N - (sc * kFusedMaxN + wc.cb)));
    }
    cnt_g[g] = n;
  }
  {
    const float n_g = static_cast<float>(cnt_g[et.grp]);
    c.part[et.grp * kTileM + et.row] = cnt_g[et.grp] > 0 ? x0 + s / n_g : 0.f;                       // group mean
    c.part[(kEpiGroups + et.grp) * kTileM + et.row] = cnt_g[et.grp] > 0 ? q - s * s / n_g : 0.f;     // group M2
  }
  // the raw scratch was written through the generic proxy; pass 2 fetches it with bulk copies (async proxy)
  __threadfence();
  ptx::fence_proxy_async_all();
  epi_bar_sync();
  float mean = 0.f, rstd;
  {
    float m2 = 0.f, cnt = 0.f;
#pragma unroll
    for (int g = 0; g < kEpiGroups; ++g) {
      const float ng = static_cast<float>(cnt_g[g]);
      if (ng > 0.f) {
        const float mg = c.part[g * kTileM + et.row], m2g = c.part[(kEpiGroups + g) * kTileM + et.row];
        const float tot = cnt + ng, delta = mg - mean;
        mean += delta * (ng / tot);
        m2 += m2g + delta * delta * (cnt * ng / tot);
        cnt = tot;
      }
    }
    rstd = rsqrtf(m2 / static_cast<float>(N) + 1e-5f);            // nn.LayerNorm eps (layers.py:101), biased variance
  }
  const float nmr = -mean * rstd;
  if (tr0) TDMPC2_TRACE(P, c, 6);
  const bool tma_out = ea.dstbuf >= 0 && !ea.out_f32 && (N % 32 == 0) && (ea.dst_col0 % 32 == 0);   // == run_layer's tma_only
  if (tma_out) {
    if (ea.kind == EPI_LN_MISH) wide_pass2_tma<EPI_LN_MISH>(P, c, et, ly, ea, rstd, nmr);
    else wide_pass2_tma<EPI_LN_SIMNORM>(P, c, et, ly, ea, rstd, nmr);
  } else {
    if (ea.kind == EPI_LN_MISH) wide_pass2_general<EPI_LN_MISH>(P, c, et, ly, ea, rstd, nmr);
    else wide_pass2_general<EPI_LN_SIMNORM>(P, c, et, ly, ea, rstd, nmr);
  }
  if (tr0) TDMPC2_TRACE(P, c, 8);                                   // normalise pass done, stores performed
}

// GEMM roles + epilogue of a wide layer; returns the number of accumulator hand-offs (= facc phases consumed).
__device__ __forceinline__ int wide_layer_tc(const PlanParams& P, Ctx& c, const LayerRec& ly, int srcbuf, const EpiArgs& ea, int kc_begin) {
  const int nnc_all = (ly.Npad + kNch - 1) / kNch;
  const int nsc = (nnc_all + 1) / 2;
  const int nkc = ly.Kpad / kKch - kc_begin;          // K-chunks [kc_begin, Kpad / 64): see PlanParams::zbias
  const int kseg = (P.kseg > 0 && P.kseg < nkc) ? P.kseg : nkc;
  const int nseg = (nkc + kseg - 1) / kseg;
  if (c.warp == 0) {
    for (int sc = 0; sc < nsc; ++sc)
      for (int seg = 0; seg < nseg; ++seg) tc_producer(P, c, ly, srcbuf, nullptr, 2 * sc, 2, kc_begin + seg * kseg, kseg);
  } else if (c.warp == 1) {
    if (!c.cg2 || c.rank == 0)
      for (int sc = 0; sc < nsc; ++sc)
        for (int seg = 0; seg < nseg; ++seg) {
          if (sc + seg > 0) {                           // the previous accumulator has been drained
            const long long tw = prof_clock();
            ptx::mbar_wait(&c.acc_empty[0], c.a_it & 1);
            c.pf0 += prof_clock() - tw;
            ++c.a_it;
            ptx::tc_fence_after();
          }
          tc_mma(P, c, ly, 2 * sc, 2, kc_begin + seg * kseg, kseg);
        }
  } else if (c.warp >= kEpiWarp0) {
    epi_wide(P, c, ly, ea, nsc, nseg);
    ptx::tc_fence_before();
  }
  return nsc * nseg;
}

// Make generic-proxy global writes (activation planes) visible to the TMA unit
// (async proxy) before the next layer's loads, and sync the CTA.
__device__ __forceinline__ void publish_planes() {
  __threadfence();
  ptx::fence_proxy_async_all();
  __syncthreads();
}

// ------------------------------------------------------------------------------------ one layer
// GEMM + epilogue; on return the epilogue's outputs are published (CTA-synchronised, TMA-visible).
template <int ENGINE, bool EPISODIC, bool WIDE>
__device__ __forceinline__ void run_layer(const PlanParams& P, Ctx& c, const LayerDev& ly_global, int srcbuf, const EpiArgs& ea,
                                          const LayerDev* next, int kc0 = 0, const float* bias_override = nullptr, int next_kc0 = 0) {
  LayerRec ly = layer_rec(ly_global);                // registers: see LayerRec
  if (bias_override) ly.bias = bias_override;        // shared-latent fold: see PlanParams::zbias
  const bool is_ln = (ea.kind == EPI_LN_MISH || ea.kind == EPI_LN_SIMNORM);
  const bool fused = (ENGINE == ENGINE_TC) && (ly.Npad <= kFusedMaxN) && (is_ln || ly.Npad <= kNch) &&
                     (ea.kind != EPI_RAW || ly.Npad <= kNch);
  if (fused) {
    const long long tl = prof_clock();
    if (threadIdx.x == 0) TDMPC2_TRACE(P, c, 0);
    const int hseg = (WIDE && !is_ln && head_seg_ok(P, ea.kind)) ? head_segments(P, ly) : 1;   // K <= 512 unless the model is wide
    if (c.warp == 0) {
      tc_producer(P, c, ly, srcbuf, next, 0, 1 << 30, kc0, 1 << 30, next_kc0);
    } else if (c.warp == 1) {
      if (!c.cg2 || c.rank == 0) {
        if (hseg > 1) tc_mma_head_seg(P, c, ly, hseg);
        else tc_mma(P, c, ly, 0, 1 << 30, kc0);
      }
    } else if (c.warp >= kEpiWarp0) {
      if (is_ln) epi_ln_fused(P, c, ly, ea);
      else epi_head_fused<EPISODIC>(P, c, ly, ea, hseg);
      ptx::tc_fence_before();
    }
    c.pf1 += prof_clock() - tl;
    // every thread tracks the facc phases: buffer 0 completed ceil(hseg / 2) times, buffer 1 floor(hseg / 2) times
    c.fph0 ^= static_cast<uint32_t>(((hseg + 1) >> 1) & 1);
    c.fph1 ^= static_cast<uint32_t>((hseg >> 1) & 1);
  } else if (ENGINE == ENGINE_TC && WIDE) {
    // LayerNorm layers (and the diagnostic raw mode) wider than TMEM; heads are never wider than one N-chunk
    const long long tl = prof_clock();
    if (threadIdx.x == 0) TDMPC2_TRACE(P, c, 0);
    const int nph = wide_layer_tc(P, c, ly, srcbuf, ea, kc0);
    c.pf1 += prof_clock() - tl;
    c.fph0 ^= static_cast<uint32_t>(nph & 1);         // one facc phase per (super-chunk, K-segment)
  } else if (ENGINE == ENGINE_SIMT) {
    gemm_simt(P, c, ly, srcbuf, kc0);
    if (is_ln) rows_ln_act(P, c, ly, ea);
    else rows_head<EPISODIC>(P, c, ly, ea);
  }                                                   // (a tcgen05 kernel without WIDE is never launched on a model with wide layers)
  const long long tp = prof_clock();
  // LN layers that went out through TMA stores wrote nothing through the generic proxy: the leaders have
  // waited for their bulk groups, so a CTA barrier is all the next layer's TMA loads need.
  const bool tma_only = (ENGINE == ENGINE_TC) && is_ln && ea.dstbuf >= 0 && (ly.N % 32 == 0) && (ea.dst_col0 % 32 == 0) && !ea.out_f32;
  if (tma_only) __syncthreads();
  else publish_planes();
  c.pf3 += prof_clock() - tp;
  if (threadIdx.x == 64 && ENGINE == ENGINE_TC) TDMPC2_TRACE(P, c, 9);
  if (ENGINE == ENGINE_TC) ptx::tc_fence_after();
}

// ------------------------------------------------------------------------------------ top-k + MPPI refit
// Runs in the LAST CTA to finish a tile of environment e (tdmpc2.py:184-197).
// Thread groups that can run the CTA-wide glue phases: the whole CTA (bar 0) or only the 16 epilogue warps (bar 1).
struct GroupAll { static constexpr int N = kThreads; __device__ static int tid() { return threadIdx.x; } __device__ static void sync() { __syncthreads(); } };
struct GroupEpi { static constexpr int N = kEpiThreads; __device__ static int tid() { return threadIdx.x - kEpiWarp0 * 32; } __device__ static void sync() { epi_bar_sync(); } };

template <class G>
__device__ __forceinline__ void refit_env(const PlanParams& P, uint8_t* scratch, size_t scratch_bytes, int e, int task) {
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(scratch);
  int nsort = 1;
  while (nsort < P.N) nsort <<= 1;
  const float* vals = P.values + static_cast<size_t>(e) * P.N;
  for (int i = G::tid(); i < nsort; i += G::N) {
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
  G::sync();
  // bitonic sort, descending (value desc, index asc on ties)
  for (int k2 = 2; k2 <= nsort; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = G::tid(); i < nsort; i += G::N) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool desc = ((i & k2) == 0);
          if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      G::sync();
    }
  }
  float* escore = reinterpret_cast<float*>(keys + nsort);     // [K]
  int* eidx = reinterpret_cast<int*>(escore + P.K);           // [K]
  float* red = reinterpret_cast<float*>(eidx + P.K);          // [2]
  const float vmax = __ldcg(vals + (0xFFFFFFFFu - static_cast<unsigned>(keys[0] & 0xFFFFFFFFull)));
  for (int k = G::tid(); k < P.K; k += G::N) {
    const int idx = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(keys[k] & 0xFFFFFFFFull));
    eidx[k] = idx;
    escore[k] = expf(__fmul_rn(P.temperature, __ldcg(vals + idx) - vmax));   // exp(T * (v - max))
    P.elite_idx32[static_cast<size_t>(e) * P.K + k] = idx;
    if (P.elite_idx_out) P.elite_idx_out[static_cast<size_t>(e) * P.K + k] = idx;
  }
  G::sync();
  if (G::tid() == 0) {
    float s = 0.f;
    for (int k = 0; k < P.K; ++k) s += escore[k];
    red[0] = s;
  }
  G::sync();
  for (int k = G::tid(); k < P.K; k += G::N) escore[k] = __fdiv_rn(escore[k], red[0]);   // score /= score.sum(0)
  G::sync();
  if (G::tid() == 0) {
    float s = 0.f;
    for (int k = 0; k < P.K; ++k) s += escore[k];
    red[1] = s + 1e-9f;                                        // score.sum(0) + 1e-9
  }
  G::sync();
  const float denom = red[1];
  for (int k = G::tid(); k < P.K; k += G::N) P.score[static_cast<size_t>(e) * P.K + k] = escore[k];
  // Gather the elites' actions once, all loads independent (the weighted sums below then run out of smem).
  const int HA = P.H * P.A;
  float* eact = reinterpret_cast<float*>(scratch + 65536);          // [K][H*A]; keys/score/idx stay below 64 KiB
  const bool staged = static_cast<size_t>(P.K) * HA * 4 <= (scratch_bytes - 65536) &&
                      static_cast<size_t>(nsort) * 8 + static_cast<size_t>(P.K) * 8 + 16 <= 65536;
  if (staged) {
    for (int i = G::tid(); i < P.K * HA; i += G::N) {
      const int k = i / HA, ta = i % HA;
      eact[i] = sample_action(P, e, ta / P.A, eidx[k], ta % P.A, task);
    }
    G::sync();
  }
  for (int i = G::tid(); i < HA; i += G::N) {
    const int t = i / P.A, a = i % P.A;
    float m = 0.f;
    for (int k = 0; k < P.K; ++k) m = fmaf(escore[k], staged ? eact[k * HA + i] : sample_action(P, e, t, eidx[k], a, task), m);
    m = __fdiv_rn(m, denom);
    float var = 0.f;
    for (int k = 0; k < P.K; ++k) {
      const float act = staged ? eact[k * HA + i] : sample_action(P, e, t, eidx[k], a, task);
      const float d = act - m;
      var = fmaf(escore[k], d * d, var);
      if (t == 0) P.elite_act0[(static_cast<size_t>(e) * P.K + k) * P.A + a] = act;
    }
    float sd = sqrtf(__fdiv_rn(var, denom));
    sd = fminf(fmaxf(sd, P.min_std), P.max_std);
    if (P.masks) { const float mk = P.masks[static_cast<size_t>(task) * P.A + a]; m *= mk; sd *= mk; }
    // every read of the OLD mean/std of this environment happened above (eact gather, or this thread's own
    // (t, a) in the unstaged path): safe to overwrite now.
    const size_t sa = (static_cast<size_t>(e) * P.H + t) * P.A + a;
    P.mean[sa] = m;
    P.std[sa] = sd;
  }
  G::sync();
}

// ------------------------------------------------------------------------------------ the kernel
// CG2: CTA pairs (cluster of 2) run the GEMMs as tcgen05 cta_group::2 MMAs (M = 256: 128 rows of each CTA's own
// tile; each CTA streams only half of every weight tile).  Only MODE_ITER with an even number of tiles per
// environment and every layer on the fused path is launched this way.
// EPISODIC (cfg.episodic, single-task models): every rollout step gains the 3-layer termination head on z_{t+1}
// and the value bookkeeping its sticky (1 - termination) factor (tdmpc2.py:126-136); compiled out otherwise.
// WIDE: the model has layers wider than the 512 TMEM columns (48M / 317M presets): adds the super-chunked wide-layer path and
// the K-segmented heads; compiled out of the kernels the 5M preset runs (their register allocation is untouched by it).
// WPF (CEM iterations of fully fused models): the TMA producer prefetches the next layer's first weight chunks into
// the W ring while the epilogue runs; the epilogue stages its output in the A ring instead.  Same arithmetic.
template <int ENGINE, bool CG2 = false, bool EPISODIC = false, bool WPF = false, bool WIDE = false>
__global__ void __launch_bounds__(kThreads, 1) plan_kernel(const __grid_constant__ PlanParams P) {
  constexpr int SPT = EPISODIC ? 9 : 6;      // layer steps per rollout time step (ITER / VALUE)
  extern __shared__ uint8_t smem_raw[];
  Ctx c;
  {
    // The operand tiles need 1024-byte alignment (128-byte swizzle atoms).  The kernel has no static shared memory, so
    // the dynamic window starts right after the 1 KiB the system reserves per CTA: it IS 1024-aligned, which is checked
    // here instead of being fixed up -- every smem pointer below is then the array symbol plus a compile-time constant
    // (no register, nothing to spill; the round-up through uintptr_t used before cost a 64-bit base that ptxas spilled
    // and reloaded at ~100 sites).
    if ((ptx::smem_u32(smem_raw) & 1023u) != 0u) {
      if (threadIdx.x == 0) printf("tdmpc2_b200: dynamic shared memory is not 1024-byte aligned (0x%x)\n", ptx::smem_u32(smem_raw));
      __trap();
    }
    c.stage_base = smem_raw;
    uint8_t* ctrl = c.stage_base + kStages * kStageBytes;
    c.a_full = reinterpret_cast<uint64_t*>(ctrl);
    c.a_empty = c.a_full + kARing;
    c.w_full = c.a_empty + kARing;
    c.w_empty = c.w_full + kWRingPair;
    c.acc_full = c.w_empty + kWRingPair;
    c.acc_empty = c.acc_full + 2;
    c.facc = c.acc_empty + 2;
    c.tmem_ptr = reinterpret_cast<uint32_t*>(c.facc + 2);
    c.flags = reinterpret_cast<int*>(c.tmem_ptr + 1);          // [8]
    c.rawb = reinterpret_cast<uint64_t*>(ctrl + 184);           // [8]  (ends at ctrl + 248)
    c.G = reinterpret_cast<float*>(ctrl + 256);                 // [128]  (18 mbarriers + tmem ptr + flags live below 256)
    c.q1 = c.G + kTileM;                                        // [128]  (ends at ctrl + 1280)
    c.term = c.q1 + kTileM;                                     // [128]  (ends at ctrl + 1792 <= kSmemCtrl)
    c.rowbuf = reinterpret_cast<float*>(ctrl + kSmemCtrl);
    c.rowenv = reinterpret_cast<int*>(ctrl + kSmemCtrl + kSmemRowBuf);
    c.vec = reinterpret_cast<float*>(ctrl + kSmemCtrl + kSmemRowBuf + kSmemRowEnv);
    c.part = c.vec + 3 * kFusedMaxN;
  }
  c.slot = blockIdx.x;
  c.cg2 = CG2 ? 1 : 0;
  c.wpf = WPF ? 1 : 0;
  c.w_pref = 0;
  c.w_ring = CG2 ? kWRingPair : kWRing;
  c.w_stride = CG2 ? kWSlotBytes / 2 : kWSlotBytes;
  c.w_lo_off = CG2 ? kAPlane : kWPlane;
  c.rank = CG2 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.pa_it = c.pw_it = c.ma_it = c.mw_it = c.a_it = c.d_it = 0;
  c.fph0 = c.fph1 = 0;
  c.rb_it = 0;
  c.pf0 = c.pf1 = c.pf2 = c.pf3 = c.pf4 = c.pf5 = c.pf6 = c.pf7 = 0;
  c.trace_step = 1 << 30;
  const long long t_kernel0 = prof_clock();
  c.tmem_base = 0;
  int* rowenv = c.rowenv;

  if (ENGINE == ENGINE_TC) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < kARing; ++s) { ptx::mbar_init(&c.a_full[s], 1); ptx::mbar_init(&c.a_empty[s], 1); }
      for (int s = 0; s < kWRingPair; ++s) { ptx::mbar_init(&c.w_full[s], 1); ptx::mbar_init(&c.w_empty[s], 1); }
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&c.acc_full[s], 1);
        ptx::mbar_init(&c.acc_empty[s], CG2 ? 2 : 1);   // wide layers: one arrival per CTA of the pair (its drain is done)
        ptx::mbar_init(&c.facc[s], 1);
      }
      for (int s = 0; s < 2 * kEpiGroups; ++s) ptx::mbar_init(&c.rawb[s], 1);
      ptx::fence_barrier_init();
      ptx::prefetch_tensormap(&P.tmX);
      ptx::prefetch_tensormap(&P.tmH);
    }
    if (CG2) ptx::cluster_sync();          // the peer's barriers are initialised before anything can signal them
    if (c.warp == 1) { if (CG2) ptx::tmem_alloc_2sm(c.tmem_ptr, 512); else ptx::tmem_alloc(c.tmem_ptr, 512); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    c.tmem_base = *c.tmem_ptr;
  }

  const LayerDev* LY = P.layers;

  if (P.stagger != 0u && P.mode == MODE_ITER && ((static_cast<unsigned>(blockIdx.x) >> (CG2 ? 1 : 0)) & 1u)) {
    const long long t_start = clock64();
    while (clock64() - t_start < static_cast<long long>(P.stagger)) __nanosleep(500);
  }

  // pair mode: the two CTAs of a pair take tiles (2p, 2p+1) and run the same number of loop trips
  for (int tile = CG2 ? 2 * (static_cast<int>(blockIdx.x) >> 1) + c.rank : static_cast<int>(blockIdx.x); tile < P.ntiles;
       tile += gridDim.x) {
    // ---------------- tile set-up: fill the input planes of X ----------------
    const long long t_setup = prof_clock();
    for (int r = threadIdx.x; r < kTileM; r += kThreads) {
      rowenv[r] = map_row(P, tile, r).env; c.G[r] = 0.f; c.q1[r] = 0.f;
      if (EPISODIC) c.term[r] = 0.f;
    }
    __syncthreads();
    __half* xhi = plane_ptr(P, c.slot, BUF_X, 0);
    __half* xlo = plane_ptr(P, c.slot, BUF_X, 1);
    const int env_tile = (P.mode == MODE_ITER || P.mode == MODE_VALUE) ? tile / P.tiles_per_env : 0;
    const int task_tile = (P.task && (P.mode == MODE_ITER || P.mode == MODE_VALUE)) ? P.task[env_tile] : 0;

    if (P.mode == MODE_LAYER) {
      const LayerDev& ly = LY[P.dbg_layer];
      for (int i = threadIdx.x; i < kTileM * ly.Kpad; i += kThreads) {
        const int r = i / ly.Kpad, col = i % ly.Kpad;
        const float x = (r < P.dbg_rows && col < ly.K) ? P.dbg_x[static_cast<size_t>(r) * ly.K + col] : 0.f;
        split_store(xhi + static_cast<size_t>(r) * P.KpadX + col, xlo + static_cast<size_t>(r) * P.KpadX + col, x);
      }
      for (int r = threadIdx.x; r < kTileM; r += kThreads) rowenv[r] = r < P.dbg_rows ? r : -1;
    } else if (P.mode == MODE_ENCODE) {
      // [obs | task_emb]  (world_model.py:108-109)
      const int W = P.obs_dim + P.T;
      for (int i = threadIdx.x; i < kTileM * W; i += kThreads) {
        const int r = i / W, col = i % W;
        const RowMap rm = map_row(P, tile, r);
        const int e = rm.env < 0 ? 0 : rm.env;
        const float x = col < P.obs_dim ? P.obs[static_cast<size_t>(e) * P.obs_dim + col]
                                        : P.emb[static_cast<size_t>(P.task ? P.task[e] : 0) * P.T + (col - P.obs_dim)];
        split_store(xhi + static_cast<size_t>(r) * P.KpadX + col, xlo + static_cast<size_t>(r) * P.KpadX + col, x);
      }
    } else if (P.mode == MODE_ITER && !P.z_rows && ((P.L & 7) == 0) && ((P.T & 7) == 0) && (P.L + P.T) / 8 <= kThreads) {
      // [z | task_emb | .]: every row of an ITER tile carries the SAME latent (z.repeat(N), tdmpc2.py:163):
      // split 8 columns once, then broadcast them down the rows with 16-byte stores.
      const int nch = (P.L + P.T) / 8;
      const int ngrp = kThreads / nch;
      const int ch = threadIdx.x % nch, rg = threadIdx.x / nch;
      if (rg < ngrp) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = ch * 8 + 2 * j + u;
            x[u] = col < P.L ? P.z[static_cast<size_t>(env_tile) * P.L + col]
                             : P.emb[static_cast<size_t>(task_tile) * P.T + (col - P.L)];
            x[u] = (fabsf(x[u]) <= 3.0e38f) ? fminf(fmaxf(x[u], -65000.f), 65000.f) : CUDART_NAN_F;   // keep NaN/inf poisonous
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
    } else {
      // [z | task_emb | a]  (world_model.py:119-120); the action columns are written per step
      const int W = P.L + P.T;
      for (int i = threadIdx.x; i < kTileM * W; i += kThreads) {
        const int r = i / W, col = i % W;
        const RowMap rm = map_row(P, tile, r);
        const int e = rm.env < 0 ? 0 : rm.env;
        float x;
        if (col < P.L) {
          x = P.z_rows ? P.z_rows[(static_cast<size_t>(e) * P.N + (rm.env < 0 ? 0 : rm.idx)) * P.L + col]
                       : P.z[static_cast<size_t>(e) * P.L + col];
        } else {
          x = P.emb[static_cast<size_t>(P.task ? P.task[e] : 0) * P.T + (col - P.L)];
        }
        split_store(xhi + static_cast<size_t>(r) * P.KpadX + col, xlo + static_cast<size_t>(r) * P.KpadX + col, x);
      }
    }
    if (P.mode == MODE_ITER && !P.actions_explicit && !P.rng_state && threadIdx.x <= P.H) {
      // This tile's slabs of the (HBM-resident, read-once) noise tensors are contiguous: ask for them now, so that the
      // per-step action pass and the terminal policy sample find them in L2 instead of paying DRAM latency.
      const int n0 = (tile % P.tiles_per_env) * kTileM, n1 = min(n0 + kTileM, P.N);
      const int t = threadIdx.x;
      if (t < P.H) {
        const int r0 = max(n0, P.P) - P.P, r1 = n1 - P.P;
        if (r1 > r0)
          ptx::bulk_prefetch_l2(P.noise_r + ((static_cast<size_t>(env_tile) * P.H + t) * (P.N - P.P) + r0) * P.A,
                                static_cast<size_t>(r1 - r0) * P.A * sizeof(float));
      } else {
        ptx::bulk_prefetch_l2(P.noise_pi + (static_cast<size_t>(env_tile) * P.N + n0) * P.A, static_cast<size_t>(n1 - n0) * P.A * sizeof(float));
      }
    }
    publish_planes();
    c.pf4 += prof_clock() - t_setup;

    // ---------------- the tile's layer program: ONE run_layer call site ----------------
    //   ENCODE : enc.0 .. enc.(n-1)
    //   PRIOR  : per t: pi.0-2 [, dyn.0-2 if t < H-1]
    //   ITER   : per t: [a_t -> X] rew.0-2, dyn.0-2 [, term.0-2 if EPISODIC] ; then pi.0-2, q_a.0-2, q_b.0-2
    int nsteps;
    if (P.mode == MODE_LAYER) nsteps = 1;
    else if (P.mode == MODE_ENCODE) nsteps = P.num_enc;
    else if (P.mode == MODE_PRIOR) nsteps = 6 * (P.H - 1) + 3;
    else nsteps = SPT * P.H + 9;
    const float* dpow = P.disc_pow + static_cast<size_t>(task_tile) * (P.H + 1);
    const int* qi = (P.mode == MODE_ITER || P.mode == MODE_VALUE) ? P.qidx + static_cast<size_t>(env_tile) * 2 : nullptr;

    for (int sidx = 0; sidx < nsteps; ++sidx) {
      EpiArgs ea;
      ea.kind = EPI_LN_MISH; ea.dstbuf = -1; ea.dst_col0 = 0; ea.out_f32 = nullptr; ea.out_pitch = 0; ea.rowmap = nullptr;
      ea.head = 0; ea.disc = 0.f; ea.tile = tile; ea.eps_base = nullptr; ea.eps_rows = 0; ea.act_out = nullptr; ea.t_out = 0;
      int li, src;
      int kc0 = 0;                       // first K-chunk of the GEMM and bias vector: the shared-latent fold (PlanParams::zbias)
      const float* bias_ov = nullptr;
      if (P.mode == MODE_LAYER) {
        li = P.dbg_layer; src = BUF_X;
        ea.kind = P.dbg_mode == 0 ? EPI_RAW : (P.dbg_mode == 1 ? EPI_LN_MISH : EPI_LN_SIMNORM);
        ea.out_f32 = P.dbg_y; ea.out_pitch = LY[li].N; ea.rowmap = rowenv;
      } else if (P.mode == MODE_ENCODE) {
        li = P.li_enc + sidx;
        src = sidx == 0 ? BUF_X : BUF_H1;
        if (sidx == P.num_enc - 1) { ea.kind = EPI_LN_SIMNORM; ea.out_f32 = P.z; ea.out_pitch = P.L; ea.rowmap = rowenv; }
        else { ea.kind = EPI_LN_MISH; ea.dstbuf = BUF_H1; }
      } else {
        // which MLP, which of its 3 layers
        int mlp, l, t = 0;       // mlp: 0 reward, 1 dynamics, 2 pi, 3 q_a, 4 q_b, 5 termination
        if (P.mode == MODE_PRIOR) {
          t = sidx / 6; l = sidx % 6;
          mlp = l < 3 ? 2 : 1; l %= 3;
        } else if (sidx < SPT * P.H) {
          t = sidx / SPT; l = sidx % SPT;
          mlp = l < 3 ? 0 : ((EPISODIC && l >= 6) ? 5 : 1); l %= 3;
          if (mlp == 0 && l == 0) {
            // X action columns <- a_t  (tdmpc2.py:176-181)
            const long long t_act = prof_clock();
            if (P.actions_explicit) {
              for (int i = threadIdx.x; i < kTileM * P.A; i += kThreads) {
                const int r = i / P.A, a = i % P.A;
                const RowMap rm = map_row(P, tile, r);
                const float v = sample_action(P, env_tile, t, rm.env < 0 ? 0 : rm.idx, a, task_tile);
                const size_t o = static_cast<size_t>(r) * P.KpadX + P.L + P.T + a;
                split_store(xhi + o, xlo + o, v);
              }
            } else {
              // mean/std/mask of step t staged in smem, then one pass of independent, coalesced noise loads
              float* sm_mean = c.rowbuf; float* sm_std = c.rowbuf + kMaxHeadCols; float* sm_mask = c.rowbuf + 2 * kMaxHeadCols;
              for (int a = threadIdx.x; a < P.A; a += kThreads) {
                const size_t sa = (static_cast<size_t>(env_tile) * P.H + t) * P.A + a;
                sm_mean[a] = P.mean[sa]; sm_std[a] = P.std[sa];
                sm_mask[a] = P.masks ? P.masks[static_cast<size_t>(task_tile) * P.A + a] : 1.f;
              }
              __syncthreads();
              const int n0 = (tile % P.tiles_per_env) * kTileM;
              const float* nz = P.noise_r + (static_cast<size_t>(env_tile) * P.H + t) * (P.N - P.P) * P.A;
              const float* pa = P.pi_actions + (static_cast<size_t>(env_tile) * P.H + t) * P.P * P.A;
              const int ng = (P.A + 7) >> 3;
              if ((((P.L + P.T) & 7) == 0) && (P.L + P.T + 8 * ng <= P.KpadX)) {
                // one (row, 8 action columns) item per thread: every load of the pass is in flight at once, two 16-byte
                // stores per item (columns past A are zero-weight padding of X: zeros there are harmless)
                for (int i = threadIdx.x; i < kTileM * ng; i += kThreads) {
                  const int r = i / ng, g8 = (i % ng) * 8;
                  const int n = min(n0 + r, P.N - 1);
                  const float* src = (n < P.P) ? pa + static_cast<size_t>(n) * P.A : nz + static_cast<size_t>(n - P.P) * P.A;
                  float x[8];
                  if (n >= P.P && P.rng_state) {                               // in-kernel noise: two groups of four
                    const float4 g0 = rng_normal4(P.rng_state, 2u * P.rng_iter, rng_group_r(P, env_tile, t, n, g8 >> 2));
                    const float4 g1 = rng_normal4(P.rng_state, 2u * P.rng_iter, rng_group_r(P, env_tile, t, n, (g8 >> 2) + 1));
                    x[0] = g0.x; x[1] = g0.y; x[2] = g0.z; x[3] = g0.w; x[4] = g1.x; x[5] = g1.y; x[6] = g1.z; x[7] = g1.w;
                  } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u)                                // noise is one-shot: evict-first
                      x[u] = (g8 + u < P.A) ? (n < P.P ? src[g8 + u] : __ldcs(src + g8 + u)) : 0.f;
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
              } else {
#pragma unroll 4
                for (int i = threadIdx.x; i < kTileM * P.A; i += kThreads) {
                  const int r = i / P.A, a = i % P.A;
                  const int n = min(n0 + r, P.N - 1);
                  float v;
                  if (n < P.P) v = pa[static_cast<size_t>(n) * P.A + a];
                  else {
                    v = __fadd_rn(sm_mean[a], __fmul_rn(sm_std[a], noise_r_at(P, env_tile, t, n, a)));   // one-shot: evict-first
                    v = fminf(fmaxf(v, -1.f), 1.f);
                  }
                  v *= sm_mask[a];
                  const size_t o = static_cast<size_t>(r) * P.KpadX + P.L + P.T + a;
                  split_store(xhi + o, xlo + o, v);
                }
              }
            }
            publish_planes();
            c.pf5 += prof_clock() - t_act;
          }
        } else {
          const int u = sidx - SPT * P.H;
          mlp = 2 + u / 3; l = u % 3;
        }
        src = l == 0 ? BUF_X : BUF_H1;
        const int base = (EPISODIC && mlp == 5) ? P.li_term
                         : mlp == 0 ? P.li_rew : mlp == 1 ? P.li_dyn : mlp == 2 ? P.li_pi : P.li_q + 3 * qi[mlp - 3];
        li = base + l;
        if (P.mode == MODE_ITER && P.zb_kc0 > 0 && sidx < SPT * P.H && t == 0 && l == 0 && mlp <= 1) {
          // t = 0: the rows of the tile share [z | emb]; its product with reward.0 / dynamics.0 was folded into a
          // per-environment bias by the prologue, the GEMM only covers the action columns
          kc0 = P.zb_kc0;
          bias_ov = P.zbias + (static_cast<size_t>(mlp) * P.E + env_tile) * P.zb_pitch;
        }
        if (l < 2) {
          ea.kind = EPI_LN_MISH; ea.dstbuf = BUF_H1;
        } else if (mlp == 0) {                 // reward (world_model.py:123-130) + two_hot_inv
          ea.kind = EPI_TWOHOT; ea.head = HEAD_REWARD; ea.disc = dpow[t];
        } else if (mlp == 1) {                 // z <- next(z, a)  (world_model.py:114-121)
          ea.kind = EPI_LN_SIMNORM; ea.dstbuf = BUF_X; ea.dst_col0 = 0;
        } else if (mlp == 2) {                 // a = pi(z)  (world_model.py:144-174) -> X action columns
          ea.kind = EPI_PI;
          if (P.mode == MODE_PRIOR) {          // eps = noise_prior[e, t, p, :]
            ea.eps_base = P.noise_prior + static_cast<size_t>(t) * P.P * P.A; ea.eps_rows = P.H * P.P;
            ea.act_out = P.pi_actions; ea.t_out = t;
          } else {                             // eps = noise_pi[e, n, :]
            ea.eps_base = P.rng_state ? nullptr : P.noise_pi; ea.eps_rows = P.N;
          }
        } else if (EPISODIC && mlp == 5) {     // termination(z_{t+1})  (world_model.py:132-141), input [z | emb]
          ea.kind = EPI_TERM;
        } else {                               // Q heads (world_model.py:186-216)
          ea.kind = EPI_TWOHOT; ea.head = (mlp == 3) ? HEAD_Q1 : HEAD_Q2; ea.disc = dpow[P.H];
        }
      }
      c.trace_step = (tile == static_cast<int>(blockIdx.x)) ? sidx : (1 << 30);
      const LayerDev* next = nullptr;
      int next_kc0 = 0;
      if (WPF && P.mode == MODE_ITER && sidx + 1 < nsteps) {
        // layer of the step after this one (same mapping as above, without its side effects); none after the
        // tile's last step: the refit and the next tile's set-up use the operand smem as scratch
        const int s1 = sidx + 1;
        int mlp1, l1;
        if (s1 < SPT * P.H) { l1 = s1 % SPT; mlp1 = l1 < 3 ? 0 : ((EPISODIC && l1 >= 6) ? 5 : 1); l1 %= 3; }
        else { const int u1 = s1 - SPT * P.H; mlp1 = 2 + u1 / 3; l1 = u1 % 3; }
        const int base1 = (EPISODIC && mlp1 == 5) ? P.li_term
                          : mlp1 == 0 ? P.li_rew : mlp1 == 1 ? P.li_dyn : mlp1 == 2 ? P.li_pi : P.li_q + 3 * qi[mlp1 - 3];
        next = &LY[base1 + l1];
        if (P.zb_kc0 > 0 && s1 < SPT && l1 == 0 && mlp1 <= 1) next_kc0 = P.zb_kc0;   // the step after this one is a folded t = 0 layer
      }
      run_layer<ENGINE, EPISODIC, WIDE>(P, c, LY[li], src, ea, next, kc0, bias_ov, next_kc0);
    }

    const long long t_refit = prof_clock();
    if (P.mode == MODE_ITER) {
      // last CTA to finish a tile of this environment refits its mean/std
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(&P.env_counter[env_tile], 1u);
        c.flags[0] = (old == static_cast<unsigned>(P.tiles_per_env - 1));
        if (c.flags[0]) P.env_counter[env_tile] = 0;
      }
      __syncthreads();
      if (c.flags[0]) {
        __threadfence();
        refit_env<GroupAll>(P, c.stage_base, static_cast<size_t>(kStages) * kStageBytes, env_tile, task_tile);
      }
      publish_planes();     // refit_env wrote the stage smem through the generic proxy; TMA reuses it next tile
    }
    c.pf6 += prof_clock() - t_refit;
  }

  if (kProf && P.prof) {
    // rows: 0 producer (warp 0 lane 0), 1 MMA issuer (warp 1 lane 0), 2 epilogue thread (warp 4 lane 0), 3 idle warp 2
    // cols: 0 barrier-wait cycles (empty | full), 1 cycles inside fused layers, 2 facc wait, 3 publish, 5 whole kernel
    const int who = (threadIdx.x == 0) ? 0 : (threadIdx.x == 32) ? 1 : (threadIdx.x == kEpiWarp0 * 32) ? 2
                    : (threadIdx.x == 64) ? 3 : -1;
    if (who >= 0) {
      long long* o = P.prof + (static_cast<size_t>(blockIdx.x) * 4 + who) * 12;
      o[0] = c.pf0; o[1] = c.pf1; o[2] = c.pf2; o[3] = c.pf3; o[4] = c.pf4; o[5] = prof_clock() - t_kernel0;
      o[6] = c.pf5; o[7] = c.pf6; o[8] = c.pf7;
    }
  }
  if (ENGINE == ENGINE_TC) {
    ptx::tc_fence_before();
    __syncthreads();
    if (CG2) ptx::cluster_sync();          // neither CTA may retire while the pair's MMAs / barriers are in use
    if (c.warp == 1) { if (CG2) ptx::tmem_dealloc_2sm(c.tmem_base, 512); else ptx::tmem_dealloc(c.tmem_base, 512); }
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

// Shared-latent fold (PlanParams::zbias): zbias[mlp][e][n] = bias[n] + 2^-k * sum_{k < Kz} x_e[k] * (W_hi + W_lo)[n][k], x_e = [z_e | emb_task(e)],
// for mlp 0 = reward.0, 1 = dynamics.0 (world_model.py:114-130: both take [z | emb | a]).  fp32 FFMA over the packed
// planes (whose sum is W * 2^k to ~22 bits, the operand the tensor-core path multiplies with).  Block = 4 environments
// x 64 output columns; a warp owns 8 columns, its lanes stride over k (coalesced half2 loads of the K-major rows).
constexpr int kZbEnvs = 4, kZbCols = 64, kZbKc = 1024;
__global__ void __launch_bounds__(256) zbias_kernel(const LayerDev* __restrict__ layers, int li_rew, int li_dyn, const float* __restrict__ z,
                                                    const float* __restrict__ emb, const int* __restrict__ task, int E, int L, int T,
                                                    int Kz, float* __restrict__ out, int pitch) {
  __shared__ float xs[kZbEnvs][kZbKc];
  const int mlp = blockIdx.z;
  const LayerDev& ly = layers[mlp == 0 ? li_rew : li_dyn];
  const int e0 = blockIdx.y * kZbEnvs, n0 = blockIdx.x * kZbCols;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8][kZbEnvs];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < kZbEnvs; ++i) acc[j][i] = 0.f;
  for (int kb = 0; kb < Kz; kb += kZbKc) {
    const int kn = min(kZbKc, Kz - kb);
    __syncthreads();
    for (int i = threadIdx.x; i < kZbEnvs * kn; i += blockDim.x) {
      const int ei = i / kn, k = kb + i % kn, e = min(e0 + ei, E - 1);
      xs[ei][i % kn] = k < L ? z[static_cast<size_t>(e) * L + k] : emb[static_cast<size_t>(task ? task[e] : 0) * T + (k - L)];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + warp * 8 + j;
      if (n >= ly.Npad) continue;
      const __half2* wh = reinterpret_cast<const __half2*>(ly.w_hi + static_cast<size_t>(n) * ly.Kpad + kb);
      const __half2* wl = reinterpret_cast<const __half2*>(ly.w_lo + static_cast<size_t>(n) * ly.Kpad + kb);
      for (int k2 = lane; k2 < kn / 2; k2 += 32) {          // Kz and kZbKc are multiples of 64
        const float2 h = __half22float2(wh[k2]), l = __half22float2(wl[k2]);
        const float w0 = h.x + l.x, w1 = h.y + l.y;
#pragma unroll
        for (int i = 0; i < kZbEnvs; ++i) acc[j][i] = fmaf(xs[i][2 * k2 + 1], w1, fmaf(xs[i][2 * k2], w0, acc[j][i]));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + warp * 8 + j;
#pragma unroll
    for (int i = 0; i < kZbEnvs; ++i) {
      const float v = warp_sum(acc[j][i]);
      if (lane == 0 && n < ly.Npad && e0 + i < E)
        out[(static_cast<size_t>(mlp) * E + e0 + i) * pitch + n] = fmaf(v, ly.inv_scale, ly.bias[n]);
    }
  }
}

// final action (tdmpc2.py:199-206, math.py:86-94); one warp per environment
__global__ void pick_kernel(const float* score, const float* elite_act0, const float* mean, const float* std,
                            const float* expo, const float* noise_final, float* action, float* prev_mean_out,
                            int* pick_out, int E, int K, int H, int A) {
  const int e = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  // argmax_k softmax(log(score_k) - log(expo_k)) == argmax of the logits (first max on ties)
  // NaN logits (score == 0 together with expo == 0, or a NaN temperature) never win a comparison: the pick then stays
  // at elite 0 (in range), where torch.argmax would return the first NaN position -- degenerate input either way.
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
  if (bi < 0 || bi >= K) bi = 0;
  if (lane == 0 && pick_out) pick_out[e] = bi;
  for (int a = lane; a < A; a += 32) {
    float v = elite_act0[(static_cast<size_t>(e) * K + bi) * A + a];
    if (noise_final) v = __fadd_rn(v, __fmul_rn(std[(static_cast<size_t>(e) * H) * A + a], noise_final[static_cast<size_t>(e) * A + a]));
    action[static_cast<size_t>(e) * A + a] = fminf(fmaxf(v, -1.f), 1.f);
  }
  for (int i = lane; i < H * A; i += 32) prev_mean_out[static_cast<size_t>(e) * H * A + i] = mean[static_cast<size_t>(e) * H * A + i];
}

}  // namespace tdmpc2
