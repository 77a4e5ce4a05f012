// Host side of the C ABI declared in include/tdmpc2_b200.h: planner object,
// memory layout of the packed weights and the workspace, TMA descriptor set-up,
// weight packing kernels and the launch sequence of the fused planning kernels.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tdmpc2_b200.h"
#include "plan_kernels.cuh"
#include "plan_pp.cuh"
#include "pixel_encoder.cuh"

using namespace tdmpc2;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(TDMPC2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static unsigned env_uint(const char* name, unsigned dflt) {
  const char* v = getenv(name);
  return (v && *v) ? static_cast<unsigned>(strtoul(v, nullptr, 10)) : dflt;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int pad_to(int x, int a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------ pack kernels
__global__ void absmax_kernel(const float* __restrict__ w, size_t n, unsigned* slot) {
  float m = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));   // non-negative floats order like uints
}

// W[N, K] fp32 row-major -> two fp16 planes [Npad, Kpad] of W * 2^k, zero padded.
// 2^k puts max|W| in [128, 256): fp16 hi keeps 11 bits, lo the next 11, both in the normal range.
// Output row n takes source row n (n < split_at) or split_at + (n - split_to) (n >= split_to): the pi head's
// log_std rows are moved to a 32-aligned column so the fused epilogue reads them with aligned tcgen05.ld.
__device__ __forceinline__ int src_index(int n, int src_n, int split_at, int split_to) {
  if (n < split_at) return n;
  if (n >= split_to && n - split_to + split_at < src_n) return n - split_to + split_at;
  return -1;
}
__global__ void split_weight_kernel(const float* __restrict__ w, int N, int K, int Npad, int Kpad, __half* hi,
                                    __half* lo, const unsigned* absmax_slot, LayerDev* entry, int src_n, int split_at,
                                    int split_to) {
  const float amax = __uint_as_float(*absmax_slot);
  int ex = 0;
  float scale = 1.f;
  if (amax > 0.f && isfinite(amax)) {
    frexpf(amax, &ex);                 // amax = m * 2^ex, m in [0.5, 1)
    scale = ldexpf(1.f, 8 - ex);       // amax * scale in [128, 256)
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) entry->inv_scale = 1.f / scale;
  const size_t total = static_cast<size_t>(Npad) * Kpad;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / Kpad), k = static_cast<int>(i % Kpad);
    float x = 0.f;
    const int sn = src_index(n, src_n, split_at, split_to);
    if (n < N && k < K && sn >= 0) x = w[static_cast<size_t>(sn) * K + k] * scale;
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn(x - __half2float(h));
  }
}

__global__ void pad_vector_kernel(const float* __restrict__ src, int n, int npad, float* dst, float fill, int src_n,
                                  int split_at, int split_to) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  const int si = src_index(i, src_n, split_at, split_to);
  dst[i] = (i < n && src && si >= 0) ? src[si] : fill;
}

// nn.Embedding(max_norm=1) (world_model.py:21): rows with ||w|| > 1 are scaled by 1/(||w|| + 1e-7) at lookup.
__global__ void emb_renorm_kernel(const float* __restrict__ emb, int T, float* out) {
  const int row = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < T; i += 32) { const float v = emb[static_cast<size_t>(row) * T + i]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  const float nrm = sqrtf(s);
  const float sc = nrm > 1.f ? 1.f / (nrm + 1e-7f) : 1.f;
  for (int i = threadIdx.x; i < T; i += 32) out[static_cast<size_t>(row) * T + i] = emb[static_cast<size_t>(row) * T + i] * sc;
}

// ------------------------------------------------------------------------------------ planner
struct LayerHost {
  int K, Kpad, N, Npad, wmap, wrow, has_ln;
  int src_n, split_at, split_to;                   // source rows and the optional row remap (pi head)
  size_t off_hi, off_lo, off_bias, off_g, off_b;   // byte offsets into the packed blob
};

struct tdmpc2_planner {
  tdmpc2_dims d;
  int num_sms = 0, nslots = 0;
  std::vector<LayerHost> layers;
  int li_enc = 0, li_dyn = 0, li_rew = 0, li_pi = 0, li_q = 0, li_term = -1, num_enc = 0;
  int nmaps = 0;
  int map_kpad[kMaxWMaps];
  int map_rows[kMaxWMaps];
  size_t map_off[kMaxWMaps];
  int KpadX = 0, KpadH = 0, NpadMax = 0, Ppad = 1, tiles_per_env = 0;
  // packed blob offsets
  size_t off_table = 0, off_absmax = 0, off_emb = 0, off_masks = 0, off_disc = 0, off_bins = 0, packed_bytes = 0;
  // workspace offsets
  size_t off_X = 0, off_H = 0, off_raw = 0, off_z = 0, off_pia = 0, off_mean = 0, off_std = 0, off_values = 0,
         off_counter = 0, off_score = 0, off_eact = 0, off_eidx = 0, off_zbias = 0, ws_bytes = 0;
  uint8_t* packed = nullptr;
  uint8_t* ws = nullptr;
  PlanParams base;
  int engine = TDMPC2_ENGINE_TCGEN05;
  bool bound = false, weights_ok = false, smem_attr_pp = false, all_fused = true;
  bool attr_done[16] = {};          // dynamic-smem opt-in done, per plan_kernel instantiation
  int64_t launches = 0;
  unsigned wide_sleep_ns = 0;
  int head_kseg = 8;                // heads: 512 elements of K per TMEM accumulation (the 5M preset's whole K)
  int kseg = 32;                    // wide layers: K-chunks per TMEM accumulation segment (2048 elements; 0 = whole K)
  size_t l2_window_bytes = 0;       // > 0: launches carry a persisting-L2 access-policy window over the activation scratch
  float l2_hit_ratio = 1.f;
  bool pair_ok = true;              // every layer of the CEM iteration can run as cta_group::2
  int l2hint = 2;                   // PlanParams::l2hint (TDMPC2_B200_L2HINT): 2 = evict-last on fused activation stores (default), 1 = operand-load hints (experiment)
  unsigned stagger = 0;             // PlanParams::stagger (TDMPC2_B200_STAGGER, clock cycles; experiment knob)
  int passes = 3;                   // 3 = fp32-parity arithmetic, 1 = declared non-parity fast mode (PlanParams::passes)
  int zb_kc0 = 0, zb_pitch = 0;     // shared-latent fold (PlanParams::zbias): K-chunks of [z | emb] folded into a per-env bias
  const int32_t* cur_task = nullptr;
  long long* prof = nullptr;
};

static int add_layer(tdmpc2_planner* p, int K, int N, bool has_ln) {
  LayerHost l{};
  l.K = K; l.Kpad = pad_to(K, kKch); l.N = N; l.Npad = pad_to(N, 128); l.has_ln = has_ln ? 1 : 0;
  l.src_n = N; l.split_at = N; l.split_to = N;
  int m = -1;
  for (int i = 0; i < p->nmaps; ++i) if (p->map_kpad[i] == l.Kpad) m = i;
  if (m < 0) {
    if (p->nmaps == kMaxWMaps) return -1;
    m = p->nmaps++;
    p->map_kpad[m] = l.Kpad;
    p->map_rows[m] = 0;
  }
  l.wmap = m;
  l.wrow = p->map_rows[m];
  p->map_rows[m] += 2 * l.Npad;
  p->layers.push_back(l);
  return static_cast<int>(p->layers.size()) - 1;
}

extern "C" int tdmpc2_abi_version(void) { return TDMPC2_B200_ABI_VERSION; }
extern "C" const char* tdmpc2_last_error(void) { return g_err.c_str(); }

static int check_device(int* num_sms) {
  int dev = 0, ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TDMPC2_ERR_NO_DEVICE, "no CUDA device: the B200 planner has no CPU fallback");
  }
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(TDMPC2_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev, prop.name,
                prop.major, prop.minor);
  *num_sms = prop.multiProcessorCount;
  return 0;
}

extern "C" int tdmpc2_planner_create(const tdmpc2_dims* dims, tdmpc2_planner** out) {
  if (!dims || !out) return fail(TDMPC2_ERR_INVALID, "null argument");
  const tdmpc2_dims& d = *dims;
  if (d.episodic != 0 && d.episodic != 1) return fail(TDMPC2_ERR_INVALID, "episodic must be 0 or 1");
  if (d.episodic && d.task_dim > 0)   // WorldModel.termination asserts task is None (world_model.py:136)
    return fail(TDMPC2_ERR_UNSUPPORTED, "episodic (termination head) models are single-task in the reference");
  if (d.num_envs < 1 || d.num_samples < 1 || d.horizon < 1 || d.iterations < 1 || d.obs_dim < 1 || d.action_dim < 1 ||
      d.latent_dim < 1 || d.mlp_dim < 1 || d.enc_dim < 1 || d.num_enc_layers < 0 || d.num_q < 2 || d.num_bins < 0 ||
      d.task_dim < 0 || d.num_tasks < 1)
    return fail(TDMPC2_ERR_INVALID, "non-positive dimension");
  if (d.num_bins < 2)   // math.two_hot_inv's num_bins == 0 (raw) / == 1 (symexp only) regression heads (math.py:76-79)
    return fail(TDMPC2_ERR_UNSUPPORTED, "num_bins=%d: only the discrete-regression heads (num_bins >= 2) are built; the "
                "reference's num_bins 0 / 1 scalar heads are not supported by the fused planner", d.num_bins);
  if (d.num_elites < 1 || d.num_elites > d.num_samples) return fail(TDMPC2_ERR_INVALID, "num_elites must be in [1, num_samples]");
  if (d.num_pi_trajs < 0 || d.num_pi_trajs > 128 || d.num_pi_trajs > d.num_samples)
    return fail(TDMPC2_ERR_INVALID, "num_pi_trajs must be in [0, min(128, num_samples)]");
  if (d.num_samples > 4096) return fail(TDMPC2_ERR_INVALID, "num_samples > 4096 unsupported");
  if (d.num_elites > 1024) return fail(TDMPC2_ERR_INVALID, "num_elites > 1024 unsupported");
  if (pad_to(d.action_dim, 32) + d.action_dim > kMaxHeadCols || d.num_bins > kMaxHeadCols)
    return fail(TDMPC2_ERR_INVALID, "pad32(action_dim)+action_dim and num_bins must be <= %d", kMaxHeadCols);
  if (d.simnorm_dim != 8 || d.latent_dim % 8 != 0)
    return fail(TDMPC2_ERR_INVALID, "simnorm_dim must be 8 and divide latent_dim");
  const int n_hidden = d.num_enc_layers - 1 > 1 ? d.num_enc_layers - 1 : 1;   // layers.py:157
  if (n_hidden + 1 > TDMPC2_MAX_ENC_LAYERS) return fail(TDMPC2_ERR_INVALID, "too many encoder layers");
  int num_sms = 0;
  int rc = check_device(&num_sms);
  if (rc) return rc;

  tdmpc2_planner* p = new tdmpc2_planner();
  p->d = d;
  p->num_sms = num_sms;
  p->nslots = num_sms;
  const int L = d.latent_dim, M = d.mlp_dim, A = d.action_dim, T = d.task_dim, B = d.num_bins;
  const int D = L + T + A;
  bool ok = true;
  auto add = [&](int K, int N, bool ln) { int i = add_layer(p, K, N, ln); if (i < 0) ok = false; return i; };
  // encoder: mlp(obs+T, n_hidden*[enc_dim], L, act=SimNorm)  (layers.py:157-159); num_enc_layers == 0: the model has no
  // state encoder (pixel observations: tdmpc2_pixel_encode + tdmpc2_plan_prologue_latent supply the latent)
  p->num_enc = d.num_enc_layers == 0 ? 0 : n_hidden + 1;
  p->li_enc = static_cast<int>(p->layers.size());
  if (p->num_enc > 0) { int k = d.obs_dim + T; for (int i = 0; i < n_hidden; ++i) { add(k, d.enc_dim, true); k = d.enc_dim; } add(k, L, true); }
  p->li_dyn = static_cast<int>(p->layers.size()); add(D, M, true); add(M, M, true); add(M, L, true);
  p->li_rew = static_cast<int>(p->layers.size()); add(D, M, true); add(M, M, true); add(M, B, false);
  p->li_pi = static_cast<int>(p->layers.size()); add(L + T, M, true); add(M, M, true);
  {
    // pi head: rows [0, A) = mean logits, rows [A, 2A) = log_std logits, the latter moved to column pad32(A)
    const int Apad = pad_to(A, 32);
    const int i = add(M, Apad + A, false);
    if (i >= 0) { p->layers[i].src_n = 2 * A; p->layers[i].split_at = A; p->layers[i].split_to = Apad; }
  }
  p->li_q = static_cast<int>(p->layers.size());
  for (int h = 0; h < d.num_q; ++h) { add(D, M, true); add(M, M, true); add(M, B, false); }
  // termination head: mlp(L+T, 2*[M], 1) on z_{t+1}  (world_model.py:28); appended so the other layer indices stay put
  p->li_term = -1;
  if (d.episodic) { p->li_term = static_cast<int>(p->layers.size()); add(L + T, M, true); add(M, M, true); add(M, 1, false); }
  if (!ok) { delete p; return fail(TDMPC2_ERR_INVALID, "more than %d distinct padded input widths", kMaxWMaps); }

  p->KpadX = std::max(pad_to(D, kKch), pad_to(d.obs_dim + T, kKch));
  p->KpadH = std::max(pad_to(M, kKch), pad_to(d.enc_dim, kKch));
  p->NpadMax = 0;
  for (auto& l : p->layers) { p->NpadMax = std::max(p->NpadMax, l.Npad); if (l.Npad > kFusedMaxN) p->all_fused = false; }
  p->Ppad = 1;
  while (p->Ppad < d.num_pi_trajs) p->Ppad <<= 1;
  p->tiles_per_env = (d.num_samples + kTileM - 1) / kTileM;
  p->wide_sleep_ns = env_uint("TDMPC2_B200_WIDE_SLEEP_NS", 0);   // experiment knob
  p->stagger = env_uint("TDMPC2_B200_STAGGER", 0);
  p->l2hint = static_cast<int>(env_uint("TDMPC2_B200_L2HINT", 2));   // bit 1 (value 2) on by default: see PlanParams::l2hint
  p->pair_ok = true;   // fused layers and the super-chunked wide layers both run as cta_group::2

  // ---- packed blob layout
  size_t off = 0;
  p->off_table = off; off = align_up(off + p->layers.size() * sizeof(LayerDev), 256);
  p->off_absmax = off; off = align_up(off + p->layers.size() * sizeof(unsigned), 256);
  for (auto& l : p->layers) {
    l.off_bias = off; off = align_up(off + l.Npad * 4, 256);
    l.off_g = off; off = align_up(off + l.Npad * 4, 256);
    l.off_b = off; off = align_up(off + l.Npad * 4, 256);
  }
  p->off_emb = off; off = align_up(off + static_cast<size_t>(d.num_tasks) * std::max(T, 1) * 4, 256);
  p->off_masks = off; off = align_up(off + static_cast<size_t>(d.num_tasks) * A * 4, 256);
  p->off_disc = off; off = align_up(off + static_cast<size_t>(d.num_tasks) * (d.horizon + 1) * 4, 256);
  p->off_bins = off; off = align_up(off + static_cast<size_t>(B) * 4, 256);
  for (int m = 0; m < p->nmaps; ++m) {
    off = align_up(off, 1024);
    p->map_off[m] = off;
    off += static_cast<size_t>(p->map_rows[m]) * p->map_kpad[m] * 2;
  }
  p->packed_bytes = align_up(off, 1024);
  for (auto& l : p->layers) {
    l.off_hi = p->map_off[l.wmap] + static_cast<size_t>(l.wrow) * l.Kpad * 2;
    l.off_lo = l.off_hi + static_cast<size_t>(l.Npad) * l.Kpad * 2;
  }
  // ---- workspace layout
  const size_t E = d.num_envs, H = d.horizon, N = d.num_samples, K = d.num_elites, P = d.num_pi_trajs;
  off = 0;
  p->off_X = off; off = align_up(off + static_cast<size_t>(p->nslots) * 2 * kTileM * p->KpadX * 2, 1024);
  p->off_H = off; off = align_up(off + static_cast<size_t>(p->nslots) * 2 * kTileM * p->KpadH * 2, 1024);
  p->off_raw = off; off = align_up(off + static_cast<size_t>(p->nslots) * kTileM * p->NpadMax * 4, 1024);
  p->off_z = off; off = align_up(off + E * L * 4, 256);
  p->off_pia = off; off = align_up(off + E * H * std::max<size_t>(P, 1) * A * 4, 256);
  p->off_mean = off; off = align_up(off + E * H * A * 4, 256);
  p->off_std = off; off = align_up(off + E * H * A * 4, 256);
  p->off_values = off; off = align_up(off + E * N * 4, 256);
  p->off_counter = off; off = align_up(off + E * 4, 256);
  p->off_score = off; off = align_up(off + E * K * 4, 256);
  p->off_eact = off; off = align_up(off + E * K * A * 4, 256);
  p->off_eidx = off; off = align_up(off + E * K * 4, 256);
  // shared-latent fold: whole 64-column chunks of [z | emb] (TDMPC2_B200_ZFOLD=0 turns it off: A/B knob)
  p->zb_pitch = p->layers[p->li_rew].Npad;
  p->zb_kc0 = env_uint("TDMPC2_B200_ZFOLD", 1) ? (L + T) / kKch : 0;
  p->off_zbias = off; off = align_up(off + 2 * E * static_cast<size_t>(p->zb_pitch) * 4, 256);
  p->ws_bytes = align_up(off, 1024);
  *out = p;
  return 0;
}

extern "C" void tdmpc2_planner_destroy(tdmpc2_planner* p) { delete p; }
extern "C" int tdmpc2_planner_packed_bytes(const tdmpc2_planner* p, size_t* out) {
  if (!p || !out) return fail(TDMPC2_ERR_INVALID, "null argument");
  *out = p->packed_bytes;
  return 0;
}
extern "C" int tdmpc2_planner_workspace_bytes(const tdmpc2_planner* p, size_t* out) {
  if (!p || !out) return fail(TDMPC2_ERR_INVALID, "null argument");
  *out = p->ws_bytes;
  return 0;
}
extern "C" int tdmpc2_planner_layer_count(const tdmpc2_planner* p) { return p ? static_cast<int>(p->layers.size()) : -1; }
extern "C" int64_t tdmpc2_planner_launch_count(const tdmpc2_planner* p) { return p ? p->launches : -1; }
extern "C" int tdmpc2_planner_set_profile(tdmpc2_planner* p, long long* device_buf) {
  if (!p) return fail(TDMPC2_ERR_INVALID, "null planner");
  p->prof = device_buf;
  return 0;
}
extern "C" int tdmpc2_planner_set_engine(tdmpc2_planner* p, int engine) {
  if (!p || (engine != TDMPC2_ENGINE_TCGEN05 && engine != TDMPC2_ENGINE_SIMT && engine != TDMPC2_ENGINE_TCGEN05_2SM &&
             engine != TDMPC2_ENGINE_TCGEN05_PP && engine != TDMPC2_ENGINE_TCGEN05_2SM_PF))
    return fail(TDMPC2_ERR_INVALID, "bad engine");
  p->engine = engine;
  return 0;
}

// Wide layers (output wider than the 512 TMEM columns): flush the TMEM partial sums to fp32 every `k_elems` elements of
// the reduction dimension (rounded to 64-element chunks) and add the segments with round-to-nearest.  0 = accumulate the
// whole reduction in TMEM (fastest).
extern "C" int tdmpc2_planner_set_kseg(tdmpc2_planner* p, int k_elems) {
  if (!p || k_elems < 0) return fail(TDMPC2_ERR_INVALID, "bad kseg");
  p->kseg = (k_elems + kKch - 1) / kKch;
  return 0;
}
// The same for the head layers (reward / Q / pi / termination outputs): their partial sums alternate between two TMEM
// buffers and are added in registers, at no measurable cost.  Default 512; 0 = whole K in one accumulation.
extern "C" int tdmpc2_planner_set_head_kseg(tdmpc2_planner* p, int k_elems) {
  if (!p || k_elems < 0) return fail(TDMPC2_ERR_INVALID, "bad head kseg");
  p->head_kseg = (k_elems + kKch - 1) / kKch;
  return 0;
}

extern "C" int tdmpc2_planner_set_passes(tdmpc2_planner* p, int passes) {
  if (!p || (passes != 1 && passes != 3)) return fail(TDMPC2_ERR_INVALID, "passes must be 3 (fp32 parity) or 1 (non-parity fast mode)");
  p->passes = passes;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn enc, CUtensorMap* m, void* base, uint64_t kpad, uint64_t rows, int box_cols = kKch,
                    int box_rows = 128) {
  cuuint64_t dims[2] = {kpad, rows};
  cuuint64_t strides[1] = {kpad * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  // 64-column boxes (operand loads): 128-byte rows, 128B swizzle; 32-column boxes (epilogue stores): 64B swizzle;
  // 16-column boxes (ping-pong epilogue stores): 32-byte rows, no swizzle
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols == kKch ? CU_TENSOR_MAP_SWIZZLE_128B : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TDMPC2_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) kpad=%llu rows=%llu", (int)r,
                                     (unsigned long long)kpad, (unsigned long long)rows);
  return 0;
}

extern "C" int tdmpc2_planner_bind(tdmpc2_planner* p, void* packed, void* workspace) {
  if (!p || !packed || !workspace) return fail(TDMPC2_ERR_INVALID, "null argument");
  if ((reinterpret_cast<uintptr_t>(packed) & 255) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return fail(TDMPC2_ERR_INVALID, "buffers must be 256-byte aligned");
  p->packed = static_cast<uint8_t*>(packed);
  p->ws = static_cast<uint8_t*>(workspace);
  CUDA_TRY(cudaMemset(p->ws, 0, p->ws_bytes));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(TDMPC2_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);

  const tdmpc2_dims& d = p->d;
  PlanParams& B = p->base;
  memset(&B, 0, sizeof(B));
  int rc;
  if ((rc = make_map(enc, &B.tmX, p->ws + p->off_X, p->KpadX, static_cast<uint64_t>(p->nslots) * 2 * kTileM))) return rc;
  if ((rc = make_map(enc, &B.tmH, p->ws + p->off_H, p->KpadH, static_cast<uint64_t>(p->nslots) * 2 * kTileM))) return rc;
  if ((rc = make_map(enc, &B.tmXs, p->ws + p->off_X, p->KpadX, static_cast<uint64_t>(p->nslots) * 2 * kTileM, 32))) return rc;
  if ((rc = make_map(enc, &B.tmHs, p->ws + p->off_H, p->KpadH, static_cast<uint64_t>(p->nslots) * 2 * kTileM, 32))) return rc;
  {
    const uint64_t rows = static_cast<uint64_t>(p->nslots) * 2 * kTileM;
    if ((rc = make_map(enc, &B.tmX64, p->ws + p->off_X, p->KpadX, rows, kKch, kPPHalf))) return rc;
    if ((rc = make_map(enc, &B.tmH64, p->ws + p->off_H, p->KpadH, rows, kKch, kPPHalf))) return rc;
    if ((rc = make_map(enc, &B.tmXs64, p->ws + p->off_X, p->KpadX, rows, kPPStgCols, kPPHalf))) return rc;
    if ((rc = make_map(enc, &B.tmHs64, p->ws + p->off_H, p->KpadH, rows, kPPStgCols, kPPHalf))) return rc;
  }
  for (int m = 0; m < p->nmaps; ++m)
    if ((rc = make_map(enc, &B.tmW[m], p->packed + p->map_off[m], p->map_kpad[m], p->map_rows[m]))) return rc;

  // layer table (host part; inv_scale is filled by the pack kernel)
  std::vector<LayerDev> tab(p->layers.size());
  for (size_t i = 0; i < tab.size(); ++i) {
    const LayerHost& l = p->layers[i];
    LayerDev& t = tab[i];
    t.K = l.K; t.Kpad = l.Kpad; t.N = l.N; t.Npad = l.Npad; t.wmap = l.wmap; t.wrow = l.wrow; t.has_ln = l.has_ln;
    t.inv_scale = 1.f;
    t.bias = reinterpret_cast<const float*>(p->packed + l.off_bias);
    t.ln_g = reinterpret_cast<const float*>(p->packed + l.off_g);
    t.ln_b = reinterpret_cast<const float*>(p->packed + l.off_b);
    t.w_hi = reinterpret_cast<const __half*>(p->packed + l.off_hi);
    t.w_lo = reinterpret_cast<const __half*>(p->packed + l.off_lo);
  }
  CUDA_TRY(cudaMemcpy(p->packed + p->off_table, tab.data(), tab.size() * sizeof(LayerDev), cudaMemcpyHostToDevice));

  B.layers = reinterpret_cast<const LayerDev*>(p->packed + p->off_table);
  B.E = d.num_envs; B.N = d.num_samples; B.P = d.num_pi_trajs; B.Ppad = p->Ppad; B.K = d.num_elites; B.H = d.horizon;
  B.obs_dim = d.obs_dim; B.A = d.action_dim; B.Apad = pad_to(d.action_dim, 32); B.L = d.latent_dim; B.M = d.mlp_dim; B.T = d.task_dim; B.B = d.num_bins;
  B.num_q = d.num_q; B.simnorm = d.simnorm_dim; B.num_enc = p->num_enc;
  B.tiles_per_env = p->tiles_per_env; B.KpadX = p->KpadX; B.KpadH = p->KpadH; B.NpadMax = p->NpadMax;
  B.li_enc = p->li_enc; B.li_dyn = p->li_dyn; B.li_rew = p->li_rew; B.li_pi = p->li_pi; B.li_q = p->li_q; B.li_term = p->li_term;
  B.temperature = d.temperature; B.min_std = d.min_std; B.max_std = d.max_std;
  B.log_std_min = d.log_std_min; B.log_std_dif = d.log_std_dif;
  B.X = reinterpret_cast<__half*>(p->ws + p->off_X);
  B.Hb = reinterpret_cast<__half*>(p->ws + p->off_H);
  B.raw = reinterpret_cast<float*>(p->ws + p->off_raw);
  B.emb = d.task_dim > 0 ? reinterpret_cast<const float*>(p->packed + p->off_emb) : nullptr;
  B.masks = nullptr;   // set by pack_weights when the model is multi-task
  B.disc_pow = reinterpret_cast<const float*>(p->packed + p->off_disc);
  B.bins = reinterpret_cast<const float*>(p->packed + p->off_bins);
  B.z = reinterpret_cast<float*>(p->ws + p->off_z);
  B.pi_actions = reinterpret_cast<float*>(p->ws + p->off_pia);
  B.mean = reinterpret_cast<float*>(p->ws + p->off_mean);
  B.std = reinterpret_cast<float*>(p->ws + p->off_std);
  B.values = reinterpret_cast<float*>(p->ws + p->off_values);
  B.env_counter = reinterpret_cast<unsigned*>(p->ws + p->off_counter);
  B.score = reinterpret_cast<float*>(p->ws + p->off_score);
  B.elite_act0 = reinterpret_cast<float*>(p->ws + p->off_eact);
  B.elite_idx32 = reinterpret_cast<int*>(p->ws + p->off_eidx);
  B.zbias = reinterpret_cast<const float*>(p->ws + p->off_zbias);
  B.zb_kc0 = 0; B.zb_pitch = p->zb_pitch;       // zb_kc0 is set on the CEM-iteration launches only
  p->bound = true;
  p->weights_ok = false;
  return 0;
}

extern "C" int tdmpc2_pack_weights(tdmpc2_planner* p, const tdmpc2_weights* w, void* stream_) {
  if (!p || !w) return fail(TDMPC2_ERR_INVALID, "null argument");
  if (!p->bound) return fail(TDMPC2_ERR_STATE, "tdmpc2_planner_bind must be called first");
  if (w->num_enc != p->num_enc) return fail(TDMPC2_ERR_INVALID, "expected %d encoder layers, got %d", p->num_enc, w->num_enc);
  const tdmpc2_dims& d = p->d;
  if (d.task_dim > 0 && (!w->task_emb || !w->action_masks)) return fail(TDMPC2_ERR_INVALID, "multi-task model needs task_emb and action_masks");
  if (!w->discount_pow || !w->bins) return fail(TDMPC2_ERR_INVALID, "discount_pow and bins are required");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaMemsetAsync(p->packed + p->off_absmax, 0, p->layers.size() * sizeof(unsigned), st));
  LayerDev* table = reinterpret_cast<LayerDev*>(p->packed + p->off_table);
  unsigned* absmax = reinterpret_cast<unsigned*>(p->packed + p->off_absmax);
  auto pack_one = [&](int li, const tdmpc2_linear& lin, size_t head) -> int {
    const LayerHost& l = p->layers[li];
    if (!lin.weight || !lin.bias) return fail(TDMPC2_ERR_INVALID, "layer %d: null weight/bias", li);
    if (l.has_ln && (!lin.ln_weight || !lin.ln_bias)) return fail(TDMPC2_ERR_INVALID, "layer %d: missing LayerNorm tensors", li);
    const float* W = lin.weight + head * static_cast<size_t>(l.src_n) * l.K;
    const size_t n = static_cast<size_t>(l.src_n) * l.K;
    const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 1024));
    absmax_kernel<<<blocks, 256, 0, st>>>(W, n, absmax + li);
    const size_t tot = static_cast<size_t>(l.Npad) * l.Kpad;
    split_weight_kernel<<<static_cast<int>(std::min<size_t>((tot + 255) / 256, 2048)), 256, 0, st>>>(
        W, l.N, l.K, l.Npad, l.Kpad, reinterpret_cast<__half*>(p->packed + l.off_hi),
        reinterpret_cast<__half*>(p->packed + l.off_lo), absmax + li, table + li, l.src_n, l.split_at, l.split_to);
    const int vb = (l.Npad + 255) / 256;
    pad_vector_kernel<<<vb, 256, 0, st>>>(lin.bias + head * l.src_n, l.N, l.Npad, reinterpret_cast<float*>(p->packed + l.off_bias),
                                          0.f, l.src_n, l.split_at, l.split_to);
    pad_vector_kernel<<<vb, 256, 0, st>>>(l.has_ln ? lin.ln_weight + head * l.src_n : nullptr, l.N, l.Npad,
                                          reinterpret_cast<float*>(p->packed + l.off_g), 1.f, l.src_n, l.split_at, l.split_to);
    pad_vector_kernel<<<vb, 256, 0, st>>>(l.has_ln ? lin.ln_bias + head * l.src_n : nullptr, l.N, l.Npad,
                                          reinterpret_cast<float*>(p->packed + l.off_b), 0.f, l.src_n, l.split_at, l.split_to);
    p->launches += 5;
    return 0;
  };
  int rc;
  for (int i = 0; i < p->num_enc; ++i) if ((rc = pack_one(p->li_enc + i, w->enc[i], 0))) return rc;
  for (int i = 0; i < 3; ++i) {
    if ((rc = pack_one(p->li_dyn + i, w->dynamics[i], 0))) return rc;
    if ((rc = pack_one(p->li_rew + i, w->reward[i], 0))) return rc;
    if ((rc = pack_one(p->li_pi + i, w->pi[i], 0))) return rc;
    for (int h = 0; h < d.num_q; ++h) if ((rc = pack_one(p->li_q + 3 * h + i, w->qs[i], h))) return rc;
    if (d.episodic && (rc = pack_one(p->li_term + i, w->termination[i], 0))) return rc;
  }
  if (d.task_dim > 0) {
    emb_renorm_kernel<<<d.num_tasks, 32, 0, st>>>(w->task_emb, d.task_dim, reinterpret_cast<float*>(p->packed + p->off_emb));
    CUDA_TRY(cudaMemcpyAsync(p->packed + p->off_masks, w->action_masks, static_cast<size_t>(d.num_tasks) * d.action_dim * 4,
                             cudaMemcpyDeviceToDevice, st));
    p->base.masks = reinterpret_cast<const float*>(p->packed + p->off_masks);
    p->launches += 1;
  } else {
    p->base.masks = nullptr;
  }
  CUDA_TRY(cudaMemcpyAsync(p->packed + p->off_disc, w->discount_pow, static_cast<size_t>(d.num_tasks) * (d.horizon + 1) * 4,
                           cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(p->packed + p->off_bins, w->bins, static_cast<size_t>(d.num_bins) * 4, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaGetLastError());
  p->weights_ok = true;
  return 0;
}

// ------------------------------------------------------------------------------------ launches
// plan_pp.cuh covers models whose trunk layers are 256 / 512 wide and whose heads fit one 128-column chunk
static bool pp_eligible(const tdmpc2_planner* p) {
  const tdmpc2_dims& d = p->d;
  if (d.episodic || !p->all_fused || d.action_dim > 64 || d.num_bins > 128 || d.num_bins < 1 || 6 * d.horizon + 9 > kPPMaxSteps) return false;
  if (d.num_samples % kTileM != 0 || (d.latent_dim + d.task_dim) % 8 != 0 || (d.latent_dim + d.task_dim) / 8 > kEpiThreads) return false;
  if (d.simnorm_dim != 8) return false;
  for (size_t i = static_cast<size_t>(p->li_dyn); i < p->layers.size(); ++i) {
    const LayerHost& l = p->layers[i];
    if (l.has_ln) { if (l.N != l.Npad || (l.Npad != 256 && l.Npad != 512)) return false; }
    else if (l.Npad != 128) return false;
  }
  return true;
}

// The engine the CEM-iteration launches of this planner actually run (the requested one falls back when a model or
// shape does not fit it: ping-pong -> CTA pairs -> single CTAs).
extern "C" int tdmpc2_planner_iter_engine(const tdmpc2_planner* p) {
  if (!p) return -1;
  if (p->engine == TDMPC2_ENGINE_SIMT || p->engine == TDMPC2_ENGINE_TCGEN05) return p->engine;
  const bool pairs = (p->tiles_per_env % 2 == 0) && p->pair_ok;
  if (!pairs) return TDMPC2_ENGINE_TCGEN05;
  if (p->engine == TDMPC2_ENGINE_TCGEN05_PP) return (p->passes == 3 && pp_eligible(p)) ? TDMPC2_ENGINE_TCGEN05_PP : TDMPC2_ENGINE_TCGEN05_2SM;
  return p->engine;
}

// W prefetch (plan_kernel<..., WPF>): every LayerNorm layer of the CEM iteration must take the epilogue's fast path
// (whole 32-column blocks out through TMA stores), which is the only one that stages in the A ring
static bool wpf_eligible(const tdmpc2_planner* p) {
  if (p->d.episodic || !p->all_fused) return false;
  for (size_t i = static_cast<size_t>(p->li_dyn); i < p->layers.size(); ++i)
    if (p->layers[i].has_ln && p->layers[i].N % 32 != 0) return false;
  return true;
}

// Launch attributes shared by every plan_kernel launch: the cluster dimension of the CTA-pair engines and -- when
// enabled (tdmpc2_planner_set_l2_persist) -- an access-policy window that keeps the per-CTA activation scratch
// (X + H planes, contiguous in the workspace) resident in the persisting part of L2, so that its dirty lines are not
// written back to HBM while noise and weights stream through.
struct LaunchCfg {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  LaunchCfg(tdmpc2_planner* p, int grid, size_t smem, cudaStream_t st, bool cluster2);
};

template <class K>
static int launch_one(tdmpc2_planner* p, K kernel, int grid, size_t smem, cudaStream_t st, bool cluster2, const PlanParams& prm,
                      int threads = kThreads) {
  LaunchCfg lc(p, grid, smem, st, cluster2);
  lc.cfg.blockDim = dim3(threads);
  CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, kernel, prm));
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  return 0;
}

template <class K>
static int launch_big(tdmpc2_planner* p, K kernel, bool* attr_done, int grid, cudaStream_t st, bool cluster2, const PlanParams& prm) {
  if (!*attr_done) {     // > 48 KiB of dynamic shared memory needs the opt-in, once per kernel instantiation
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    *attr_done = true;
  }
  return launch_one(p, kernel, grid, kSmemBytes, st, cluster2, prm);
}

static int launch_plan(tdmpc2_planner* p, const PlanParams& prm, int ntiles, cudaStream_t st) {
  const bool simt = p->engine == TDMPC2_ENGINE_SIMT;
  // episodic models: the rollout modes run the instantiations that carry the termination head
  const bool epi = p->d.episodic && (prm.mode == MODE_ITER || prm.mode == MODE_VALUE);
  // models with layers wider than TMEM (and the one-layer diagnostic mode, whose raw output may be) take the WIDE kernels
  const bool wide = !p->all_fused || prm.mode == MODE_LAYER;
  int grid = std::min(ntiles, p->nslots);
  PlanParams prm2 = prm;
  prm2.prof = p->prof;
  prm2.prof_slots = p->nslots;
  prm2.kseg = p->kseg;
  prm2.head_kseg = p->head_kseg;
  prm2.wide_sleep_ns = p->wide_sleep_ns;
  prm2.passes = p->passes;
  prm2.stagger = p->stagger;
  prm2.l2hint = p->l2hint;
  static_assert(sizeof(p->attr_done) / sizeof(bool) >= 16, "attr_done slots");
  // CTA-pair (cta_group::2) launch: CEM iterations only, whole pairs of tiles of one environment
  const bool pair_engine = (p->engine == TDMPC2_ENGINE_TCGEN05_2SM || p->engine == TDMPC2_ENGINE_TCGEN05_PP ||
                            p->engine == TDMPC2_ENGINE_TCGEN05_2SM_PF);
  if (p->engine == TDMPC2_ENGINE_TCGEN05_PP && prm.mode == MODE_ITER && (p->tiles_per_env % 2 == 0) && (ntiles % 2 == 0) &&
      p->passes == 3 && pp_eligible(p)) {
    // ping-pong kernel (plan_pp.cuh): GEMM of one 64-row half overlaps the epilogue of the other
    if (!p->smem_attr_pp) {
      CUDA_TRY(cudaFuncSetAttribute(plan_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPPSmemBytes));
      p->smem_attr_pp = true;
    }
    return launch_one(p, plan_pp_kernel, grid & ~1, kPPSmemBytes, st, true, prm2, kPPThreads);
  }
  bool* ad = p->attr_done;
  if (simt) {
    if (epi) return launch_big(p, plan_kernel<ENGINE_SIMT, false, true>, &ad[0], grid, st, false, prm2);
    return launch_big(p, plan_kernel<ENGINE_SIMT>, &ad[1], grid, st, false, prm2);
  }
  // CTA pairs take tiles (2p, 2p + 1): CEM iterations with an even number of tiles per environment (so that a pair never
  // straddles the refit hand-off asymmetrically), and the policy-prior rollouts whenever the tile count is even
  const bool pair = pair_engine && p->pair_ok && (ntiles % 2 == 0) &&
                    ((prm.mode == MODE_ITER && p->tiles_per_env % 2 == 0) || prm.mode == MODE_PRIOR);
  if (pair) {
    grid &= ~1;
    if (wide) {
      if (epi) return launch_big(p, plan_kernel<ENGINE_TC, true, true, false, true>, &ad[2], grid, st, true, prm2);
      return launch_big(p, plan_kernel<ENGINE_TC, true, false, false, true>, &ad[3], grid, st, true, prm2);
    }
    if (epi) return launch_big(p, plan_kernel<ENGINE_TC, true, true>, &ad[4], grid, st, true, prm2);
    if (p->engine == TDMPC2_ENGINE_TCGEN05_2SM_PF && wpf_eligible(p))
      return launch_big(p, plan_kernel<ENGINE_TC, true, false, true>, &ad[5], grid, st, true, prm2);
    return launch_big(p, plan_kernel<ENGINE_TC, true>, &ad[6], grid, st, true, prm2);
  }
  if (wide) {
    if (epi) return launch_big(p, plan_kernel<ENGINE_TC, false, true, false, true>, &ad[7], grid, st, false, prm2);
    return launch_big(p, plan_kernel<ENGINE_TC, false, false, false, true>, &ad[8], grid, st, false, prm2);
  }
  if (epi) return launch_big(p, plan_kernel<ENGINE_TC, false, true>, &ad[9], grid, st, false, prm2);
  return launch_big(p, plan_kernel<ENGINE_TC>, &ad[10], grid, st, false, prm2);
}

LaunchCfg::LaunchCfg(tdmpc2_planner* p, int grid, size_t smem, cudaStream_t st, bool cluster2) {
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  int n = 0;
  if (cluster2) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (p->l2_window_bytes > 0) {
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow.base_ptr = p->ws + p->off_X;
    attr[n].val.accessPolicyWindow.num_bytes = p->l2_window_bytes;
    attr[n].val.accessPolicyWindow.hitRatio = p->l2_hit_ratio;
    attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
}

// Keep the activation scratch in the persisting part of L2 (0 = off).  Sets the device's persisting-L2 carve-out to
// what the scratch needs (capped by the device maximum) -- a device-wide setting, hence opt-in.
extern "C" int tdmpc2_planner_set_l2_persist(tdmpc2_planner* p, int enable) {
  if (!p) return fail(TDMPC2_ERR_INVALID, "null planner");
  if (!p->bound) return fail(TDMPC2_ERR_STATE, "tdmpc2_planner_bind must be called first");
  if (!enable) { p->l2_window_bytes = 0; return 0; }
  int dev = 0, max_persist = 0, max_window = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  const size_t scratch = p->off_raw - p->off_X;              // X planes + H planes of all slots
  if (max_persist <= 0 || max_window <= 0) return fail(TDMPC2_ERR_UNSUPPORTED, "device has no persisting L2");
  const size_t carve = std::min<size_t>(scratch, static_cast<size_t>(max_persist));
  CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
  p->l2_window_bytes = std::min<size_t>(scratch, static_cast<size_t>(max_window));
  p->l2_hit_ratio = std::min(1.0f, static_cast<float>(carve) / static_cast<float>(p->l2_window_bytes));
  return 0;
}

static int ready(tdmpc2_planner* p) {
  if (!p) return fail(TDMPC2_ERR_INVALID, "null planner");
  if (!p->bound || !p->weights_ok) return fail(TDMPC2_ERR_STATE, "planner needs bind() and pack_weights() first");
  return 0;
}

// obs != nullptr: z = encode(obs, task) through the state encoder; z_in != nullptr: the latent is given (pixel models)
static int prologue_impl(tdmpc2_planner* p, const float* obs, const float* z_in, const int32_t* task, const uint8_t* t0,
                         const float* prev_mean, const float* noise_prior, void* stream_) {
  int rc = ready(p);
  if (rc) return rc;
  const tdmpc2_dims& d = p->d;
  if ((!obs && !z_in) || !t0 || !prev_mean) return fail(TDMPC2_ERR_INVALID, "null argument");
  if (obs && p->num_enc == 0) return fail(TDMPC2_ERR_STATE, "this planner was created without a state encoder (num_enc_layers = 0): use tdmpc2_plan_prologue_latent");
  if (d.task_dim > 0 && !task) return fail(TDMPC2_ERR_INVALID, "multi-task model needs task indices");
  if (d.num_pi_trajs > 0 && !noise_prior) return fail(TDMPC2_ERR_INVALID, "noise_prior is required when num_pi_trajs > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  p->cur_task = d.task_dim > 0 ? task : nullptr;
  const int n = d.num_envs * d.horizon * d.action_dim;
  init_state_kernel<<<(std::max(n, d.num_envs) + 255) / 256, 256, 0, st>>>(p->base.mean, p->base.std, p->base.env_counter,
                                                                          prev_mean, t0, d.num_envs, d.horizon, d.action_dim,
                                                                          d.max_std);
  p->launches += 1;
  PlanParams prm = p->base;
  prm.task = p->cur_task;
  if (obs) {
    prm.obs = obs;
    prm.mode = MODE_ENCODE;
    prm.ntiles = (d.num_envs + kTileM - 1) / kTileM;
    if ((rc = launch_plan(p, prm, prm.ntiles, st))) return rc;
  } else {
    CUDA_TRY(cudaMemcpyAsync(p->base.z, z_in, static_cast<size_t>(d.num_envs) * d.latent_dim * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (p->zb_kc0 > 0) {
    // shared-latent fold: [z | emb] . W of reward.0 / dynamics.0, once per plan() (z is the same in every CEM iteration)
    const dim3 grid((p->zb_pitch + kZbCols - 1) / kZbCols, (d.num_envs + kZbEnvs - 1) / kZbEnvs, 2);
    zbias_kernel<<<grid, 256, 0, st>>>(p->base.layers, p->li_rew, p->li_dyn, p->base.z, p->base.emb, p->cur_task, d.num_envs,
                                       d.latent_dim, d.task_dim, p->zb_kc0 * kKch, reinterpret_cast<float*>(p->ws + p->off_zbias),
                                       p->zb_pitch);
    CUDA_TRY(cudaGetLastError());
    p->launches += 1;
  }
  if (d.num_pi_trajs > 0) {
    prm.mode = MODE_PRIOR;
    prm.noise_prior = noise_prior;
    const int per = kTileM / p->Ppad;
    prm.ntiles = (d.num_envs + per - 1) / per;
    if ((rc = launch_plan(p, prm, prm.ntiles, st))) return rc;
  }
  return 0;
}

extern "C" int tdmpc2_plan_prologue(tdmpc2_planner* p, const float* obs, const int32_t* task, const uint8_t* t0,
                                    const float* prev_mean, const float* noise_prior, void* stream_) {
  if (!obs) return fail(TDMPC2_ERR_INVALID, "null argument");
  return prologue_impl(p, obs, nullptr, task, t0, prev_mean, noise_prior, stream_);
}
extern "C" int tdmpc2_plan_prologue_latent(tdmpc2_planner* p, const float* z, const int32_t* task, const uint8_t* t0,
                                           const float* prev_mean, const float* noise_prior, void* stream_) {
  if (!z) return fail(TDMPC2_ERR_INVALID, "null argument");
  return prologue_impl(p, nullptr, z, task, t0, prev_mean, noise_prior, stream_);
}

// ------------------------------------------------------------------------------------ pixel encoder (cfg.obs == 'rgb')
struct tdmpc2_pixel_encoder {
  tdmpc2_pixel_dims d;
  size_t ws_bytes = 0, smem = 0;
  bool attr_done = false;
};

extern "C" int tdmpc2_pixel_encoder_create(const tdmpc2_pixel_dims* dims, tdmpc2_pixel_encoder** out) {
  if (!dims || !out) return fail(TDMPC2_ERR_INVALID, "null argument");
  const tdmpc2_pixel_dims& d = *dims;
  if (d.num_envs < 1 || d.in_channels < 1 || d.num_channels < 8 || d.num_channels % 8 != 0 || d.simnorm_dim < 1 ||
      (16 * d.num_channels) % d.simnorm_dim != 0)
    return fail(TDMPC2_ERR_INVALID, "pixel encoder: num_channels must be a positive multiple of 8 and simnorm_dim must divide 16 * num_channels");
  int num_sms = 0;
  int rc = check_device(&num_sms);
  if (rc) return rc;
  tdmpc2_pixel_encoder* e = new tdmpc2_pixel_encoder();
  e->d = d;
  e->ws_bytes = align_up(static_cast<size_t>(d.num_envs) * d.num_channels * kPixO1 * kPixO1 * 4, 256);
  const size_t stage = static_cast<size_t>(d.in_channels) * kPixHW * kPixHW;
  const size_t maps = static_cast<size_t>(d.num_channels) * (kPixO2 * kPixO2 + kPixO3 * kPixO3 + kPixO4 * kPixO4);
  e->smem = std::max(stage, maps) * 4;
  if (e->smem > 227 * 1024) { delete e; return fail(TDMPC2_ERR_INVALID, "pixel encoder: %d input channels do not fit shared memory", d.in_channels); }
  *out = e;
  return 0;
}
extern "C" void tdmpc2_pixel_encoder_destroy(tdmpc2_pixel_encoder* e) { delete e; }
extern "C" int tdmpc2_pixel_encoder_workspace_bytes(const tdmpc2_pixel_encoder* e, size_t* out) {
  if (!e || !out) return fail(TDMPC2_ERR_INVALID, "null argument");
  *out = e->ws_bytes;
  return 0;
}
extern "C" int tdmpc2_pixel_encode(tdmpc2_pixel_encoder* e, void* workspace, const tdmpc2_conv_weights* w, const float* frames,
                                   const float* shift, const float* grid_base, float* z_out, void* stream_) {
  if (!e || !workspace || !w || !frames || !shift || !grid_base || !z_out) return fail(TDMPC2_ERR_INVALID, "null argument");
  for (int i = 0; i < 4; ++i) if (!w->weight[i] || !w->bias[i]) return fail(TDMPC2_ERR_INVALID, "pixel encoder: null conv weight");
  PixelParams P{};
  P.frames = frames; P.shift = shift; P.grid = grid_base; P.scratch = static_cast<float*>(workspace); P.z = z_out;
  for (int i = 0; i < 4; ++i) { P.w[i] = w->weight[i]; P.b[i] = w->bias[i]; }
  P.E = e->d.num_envs; P.C = e->d.in_channels; P.nc = e->d.num_channels; P.simnorm = e->d.simnorm_dim;
  if (!e->attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(pixel_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(e->smem)));
    e->attr_done = true;
  }
  pixel_encode_kernel<<<P.E, kPixThreads, e->smem, static_cast<cudaStream_t>(stream_)>>>(P);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int tdmpc2_plan_iter(tdmpc2_planner* p, const float* noise_r, const float* noise_pi, const int32_t* qidx,
                                float* values_out, int64_t* elite_idx_out, void* stream_) {
  int rc = ready(p);
  if (rc) return rc;
  const tdmpc2_dims& d = p->d;
  if (!noise_pi || !qidx || (d.num_samples > d.num_pi_trajs && !noise_r)) return fail(TDMPC2_ERR_INVALID, "null argument");
  PlanParams prm = p->base;
  prm.task = p->cur_task;
  prm.mode = MODE_ITER;
  prm.noise_r = noise_r; prm.noise_pi = noise_pi; prm.qidx = qidx;
  prm.values_out = values_out;
  prm.elite_idx_out = reinterpret_cast<long long*>(elite_idx_out);
  prm.ntiles = d.num_envs * p->tiles_per_env;
  prm.zb_kc0 = p->zb_kc0;
  return launch_plan(p, prm, prm.ntiles, static_cast<cudaStream_t>(stream_));
}

// Declared non-parity throughput mode: the iteration generates its two large noise tensors itself (rng.cuh).
extern "C" int tdmpc2_plan_iter_rng(tdmpc2_planner* p, const uint64_t* rng_state, int iteration, const int32_t* qidx,
                                    float* values_out, int64_t* elite_idx_out, void* stream_) {
  int rc = ready(p);
  if (rc) return rc;
  const tdmpc2_dims& d = p->d;
  if (!rng_state || !qidx || iteration < 0) return fail(TDMPC2_ERR_INVALID, "null argument");
  if (p->engine == TDMPC2_ENGINE_SIMT) return fail(TDMPC2_ERR_UNSUPPORTED, "in-kernel noise needs a tcgen05 engine");
  PlanParams prm = p->base;
  prm.task = p->cur_task;
  prm.mode = MODE_ITER;
  prm.noise_r = nullptr; prm.noise_pi = nullptr; prm.qidx = qidx;
  prm.rng_state = reinterpret_cast<const unsigned long long*>(rng_state);
  prm.rng_iter = iteration;
  prm.values_out = values_out;
  prm.elite_idx_out = reinterpret_cast<long long*>(elite_idx_out);
  prm.ntiles = d.num_envs * p->tiles_per_env;
  prm.zb_kc0 = p->zb_kc0;
  return launch_plan(p, prm, prm.ntiles, static_cast<cudaStream_t>(stream_));
}
// Diagnostics / tests: the normals of `ngroups` consecutive groups of one stream (4 per group).
extern "C" int tdmpc2_debug_rng(const uint64_t* rng_state, uint32_t stream, uint64_t group0, int ngroups, float* out, void* stream_) {
  if (!rng_state || !out || ngroups < 1) return fail(TDMPC2_ERR_INVALID, "bad debug_rng arguments");
  rng_debug_kernel<<<(ngroups + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const unsigned long long*>(rng_state), stream, group0, ngroups, out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int tdmpc2_plan_epilogue(tdmpc2_planner* p, const float* expo, const float* noise_final, float* action_out,
                                    float* prev_mean_out, int32_t* pick_out, void* stream_) {
  int rc = ready(p);
  if (rc) return rc;
  if (!expo || !action_out || !prev_mean_out) return fail(TDMPC2_ERR_INVALID, "null argument");
  const tdmpc2_dims& d = p->d;
  const int warps_per_block = 4;
  pick_kernel<<<(d.num_envs + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, static_cast<cudaStream_t>(stream_)>>>(
      p->base.score, p->base.elite_act0, p->base.mean, p->base.std, expo, noise_final, action_out, prev_mean_out, pick_out,
      d.num_envs, d.num_elites, d.horizon, d.action_dim);
  CUDA_TRY(cudaGetLastError());
  p->launches += 1;
  return 0;
}

extern "C" int tdmpc2_plan_get_state(tdmpc2_planner* p, float* mean, float* std, float* z, float* pi_actions, float* score,
                                     void* stream_) {
  int rc = ready(p);
  if (rc) return rc;
  const tdmpc2_dims& d = p->d;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const size_t E = d.num_envs, H = d.horizon, A = d.action_dim;
  if (mean) CUDA_TRY(cudaMemcpyAsync(mean, p->base.mean, E * H * A * 4, cudaMemcpyDeviceToDevice, st));
  if (std) CUDA_TRY(cudaMemcpyAsync(std, p->base.std, E * H * A * 4, cudaMemcpyDeviceToDevice, st));
  if (z) CUDA_TRY(cudaMemcpyAsync(z, p->base.z, E * d.latent_dim * 4, cudaMemcpyDeviceToDevice, st));
  if (pi_actions && d.num_pi_trajs > 0)
    CUDA_TRY(cudaMemcpyAsync(pi_actions, p->base.pi_actions, E * H * d.num_pi_trajs * A * 4, cudaMemcpyDeviceToDevice, st));
  if (score) CUDA_TRY(cudaMemcpyAsync(score, p->base.score, E * d.num_elites * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int tdmpc2_estimate_value(tdmpc2_planner* p, const float* z, const float* actions, const int32_t* task,
                                     const float* noise_pi, const int32_t* qidx, float* value_out, void* stream_) {
  int rc = ready(p);
  if (rc) return rc;
  const tdmpc2_dims& d = p->d;
  if (!z || !actions || !noise_pi || !qidx || !value_out) return fail(TDMPC2_ERR_INVALID, "null argument");
  if (d.task_dim > 0 && !task) return fail(TDMPC2_ERR_INVALID, "multi-task model needs task indices");
  PlanParams prm = p->base;
  prm.task = d.task_dim > 0 ? task : nullptr;
  prm.mode = MODE_VALUE;
  prm.z_rows = z; prm.actions_explicit = actions; prm.noise_pi = noise_pi; prm.qidx = qidx;
  prm.values_out = value_out;
  prm.ntiles = d.num_envs * p->tiles_per_env;
  return launch_plan(p, prm, prm.ntiles, static_cast<cudaStream_t>(stream_));
}

extern "C" int tdmpc2_debug_layer(tdmpc2_planner* p, int layer, int mode, const float* x, int rows, float* y, void* stream_) {
  int rc = ready(p);
  if (rc) return rc;
  if (layer < 0 || layer >= static_cast<int>(p->layers.size()) || rows < 1 || rows > kTileM || !x || !y || mode < 0 || mode > 2)
    return fail(TDMPC2_ERR_INVALID, "bad debug_layer arguments");
  if (mode != 0 && !p->layers[layer].has_ln) return fail(TDMPC2_ERR_INVALID, "layer %d has no LayerNorm", layer);
  if (p->layers[layer].Kpad > p->KpadX) return fail(TDMPC2_ERR_INVALID, "layer %d input wider than the X scratch", layer);
  PlanParams prm = p->base;
  prm.mode = MODE_LAYER;
  prm.dbg_layer = layer; prm.dbg_mode = mode; prm.dbg_rows = rows; prm.dbg_x = x; prm.dbg_y = y;
  prm.ntiles = 1;
  return launch_plan(p, prm, 1, static_cast<cudaStream_t>(stream_));
}
