// Pixel-observation encoder of TD-MPC2 (cfg.obs == 'rgb'), forward only, for the planner's prologue.
//
// Reference: layers.conv (common/layers.py:136-150): ShiftAug (:36-59, applied at inference too -- it sits inside the
// nn.Sequential), PixelPreprocess (:62-71: x / 255 - 0.5), Conv2d(C, nc, 7, stride 2) + ReLU, Conv2d(nc, nc, 5, stride 2)
// + ReLU, Conv2d(nc, nc, 3, stride 2) + ReLU, Conv2d(nc, nc, 3, stride 1), Flatten, SimNorm (:74-88).  64 x 64 inputs
// (asserted at :141) give 29 -> 13 -> 6 -> 4 feature maps, so latent_dim == 16 * nc.
//
// One CTA per environment: the augmented, normalised frame stack is staged in shared memory (C x 64 x 64 fp32), conv1
// goes to a global scratch (it does not fit next to its input), conv2-4 and SimNorm stay in shared memory.  Plain fp32
// FFMA -- the encoder runs once per plan() on E images and is < 2 % of a plan's time; the products are exact fp32, the
// summation order differs from ATen's (1e-6 relative).
//
// ShiftAug restated: pad 3 with edge replication, sample the padded 70 x 70 image bilinearly (zeros outside, align_corners
// = False) at base + shift * 2/70, where base is linspace(-1 + 1/70, 1 - 1/70, 70)[:64] (computed by the host with
// torch.linspace so that the coordinates are bit-identical to the reference's) and shift is an integer pair in [0, 6]
// drawn by the host (torch.randint, the first random draw of a reference _plan call on pixel observations).  The sample
// points are pixel centres up to fp32 rounding, so the result is the shifted crop plus O(1e-5) of its neighbours -- kept,
// because the reference has it.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace tdmpc2 {

constexpr int kPixHW = 64, kPixPad = 3, kPixPadded = kPixHW + 2 * kPixPad;
constexpr int kPixO1 = 29, kPixO2 = 13, kPixO3 = 6, kPixO4 = 4;
constexpr int kPixThreads = 512;

struct PixelParams {
  const float* frames;   // [E, C, 64, 64], 0 .. 255
  const float* shift;    // [E, 2]: (x, y), integral values in [0, 6]  (layers.py:55)
  const float* grid;     // [64]: linspace(-1 + eps, 1 - eps, 70)[:64]  (layers.py:51)
  const float* w[4];     // Conv2d weights [out, in, k, k] as nn.Conv2d stores them
  const float* b[4];
  float* scratch;        // [E][nc * 29 * 29]
  float* z;              // [E, 16 * nc]
  int E, C, nc, simnorm;
};

// ATen's CPU grid sampler (GridSamplerKernel.cpp, align_corners = false): (x + 1) * (size / 2) - 0.5, separate roundings.
__device__ __forceinline__ float pix_unnormalize(float g) {
  return __fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f * kPixPadded), 0.5f);
}
__device__ __forceinline__ float pix_padded(const float* img, int py, int px) {   // replicate-padded image, zeros outside
  if (py < 0 || py >= kPixPadded || px < 0 || px >= kPixPadded) return 0.f;
  const int y = min(max(py - kPixPad, 0), kPixHW - 1), x = min(max(px - kPixPad, 0), kPixHW - 1);
  return img[y * kPixHW + x];
}

// out[oc][oy][ox] = act(b[oc] + sum_{ic,ky,kx} in[ic][oy*S+ky][ox*S+kx] * w[oc][ic][ky][kx]); 8 output channels per item
template <bool RELU, bool IN_GLOBAL>
__device__ __forceinline__ void pix_conv(const float* in, int IC, int IH, int IW, const float* __restrict__ w,
                                         const float* __restrict__ b, int OC, int K, int S, float* out, int OH, int OW) {
  const int npos = OH * OW, ngrp = OC / 8;
  for (int i = threadIdx.x; i < ngrp * npos; i += kPixThreads) {
    const int pos = i % npos, oc0 = (i / npos) * 8;
    const int oy = pos / OW, ox = pos % OW;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __ldg(b + oc0 + j);
    const int wstride = IC * K * K;
    for (int ic = 0; ic < IC; ++ic)
      for (int ky = 0; ky < K; ++ky) {
        const float* irow = in + (static_cast<size_t>(ic) * IH + oy * S + ky) * IW + ox * S;
        const float* wrow = w + (static_cast<size_t>(oc0) * IC + ic) * K * K + ky * K;
        for (int kx = 0; kx < K; ++kx) {
          const float v = IN_GLOBAL ? __ldcg(irow + kx) : irow[kx];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, __ldg(wrow + j * wstride + kx), acc[j]);
        }
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) out[static_cast<size_t>(oc0 + j) * npos + pos] = RELU ? fmaxf(acc[j], 0.f) : acc[j];
  }
}

__global__ void __launch_bounds__(kPixThreads, 1) pixel_encode_kernel(const PixelParams P) {
  extern __shared__ float pix_smem[];
  const int e = blockIdx.x;
  const float* img = P.frames + static_cast<size_t>(e) * P.C * kPixHW * kPixHW;
  // ---- ShiftAug + PixelPreprocess -> smem [C][64][64]
  const float sx = __fmul_rn(P.shift[e * 2 + 0], 2.0f / kPixPadded), sy = __fmul_rn(P.shift[e * 2 + 1], 2.0f / kPixPadded);
  for (int i = threadIdx.x; i < P.C * kPixHW * kPixHW; i += kPixThreads) {
    const int c = i / (kPixHW * kPixHW), y = (i / kPixHW) % kPixHW, x = i % kPixHW;
    const float fx = pix_unnormalize(__fadd_rn(P.grid[x], sx)), fy = pix_unnormalize(__fadd_rn(P.grid[y], sy));
    const float wx = floorf(fx), ny = floorf(fy);
    const float ex = wx + 1.f, sy1 = ny + 1.f;
    const int ix = static_cast<int>(wx), iy = static_cast<int>(ny);
    const float* ch = img + static_cast<size_t>(c) * kPixHW * kPixHW;
    float v = pix_padded(ch, iy, ix) * ((ex - fx) * (sy1 - fy));
    v += pix_padded(ch, iy, ix + 1) * ((fx - wx) * (sy1 - fy));
    v += pix_padded(ch, iy + 1, ix) * ((ex - fx) * (fy - ny));
    v += pix_padded(ch, iy + 1, ix + 1) * ((fx - wx) * (fy - ny));
    pix_smem[i] = __fsub_rn(__fdiv_rn(v, 255.f), 0.5f);
  }
  __syncthreads();
  float* s1 = P.scratch + static_cast<size_t>(e) * P.nc * kPixO1 * kPixO1;
  pix_conv<true, false>(pix_smem, P.C, kPixHW, kPixHW, P.w[0], P.b[0], P.nc, 7, 2, s1, kPixO1, kPixO1);
  __syncthreads();                               // conv1 output (global, written by this CTA) is visible to this CTA
  float* s2 = pix_smem;                          // the staged input is dead now
  float* s3 = s2 + P.nc * kPixO2 * kPixO2;
  float* s4 = s3 + P.nc * kPixO3 * kPixO3;
  pix_conv<true, true>(s1, P.nc, kPixO1, kPixO1, P.w[1], P.b[1], P.nc, 5, 2, s2, kPixO2, kPixO2);
  __syncthreads();
  pix_conv<true, false>(s2, P.nc, kPixO2, kPixO2, P.w[2], P.b[2], P.nc, 3, 2, s3, kPixO3, kPixO3);
  __syncthreads();
  pix_conv<false, false>(s3, P.nc, kPixO3, kPixO3, P.w[3], P.b[3], P.nc, 3, 1, s4, kPixO4, kPixO4);
  __syncthreads();
  // ---- Flatten ([nc][4][4] is already the flattened order) + SimNorm: softmax over groups of `simnorm` consecutive values
  const int L = P.nc * kPixO4 * kPixO4;
  for (int g0 = threadIdx.x * P.simnorm; g0 < L; g0 += kPixThreads * P.simnorm) {
    float m = -CUDART_INF_F;
    for (int i = 0; i < P.simnorm; ++i) m = fmaxf(m, s4[g0 + i]);
    float t = 0.f;
    for (int i = 0; i < P.simnorm; ++i) t += expf(s4[g0 + i] - m);
    for (int i = 0; i < P.simnorm; ++i) P.z[static_cast<size_t>(e) * L + g0 + i] = __fdiv_rn(expf(s4[g0 + i] - m), t);
  }
}

}  // namespace tdmpc2
