// f-2: loss epilogues of LVD training over the path's low-res outputs.
// Reference: models/synthesizer.py:1114-1118 `blur` (torchvision GaussianBlur, reflect padding) and
// :886-892 the layer-entropy regulariser, :933 `fg_mask`; both forward and backward.
#pragma once
#include "wb_common.cuh"
#include "../../include/waldo_b200.h"

// ------------------------------------------------------------------------------------------------ Gaussian blur
// blur(vid, sigma, kernel_size) = conv2d(reflect_pad(vid), outer(k, k)) per plane, k = normalised Gaussian taps
// (torchvision.transforms.functional.gaussian_blur).  The 2-D kernel is an outer product, so the kernel here runs the two
// 1-D passes on a shared-memory tile (32 x 32 outputs + a halo of `r` on every side): one global read per input sample and
// tile, no intermediate plane in HBM.  ADJ = the adjoint (backward): zero padding instead of reflection on the way in and
// the reflected margins folded back onto the taps:
//   forward   y[j] = sum_t k[t] x[refl(j + t - r)]
//   adjoint  dx[i] = sum_j dy[j] (k[i - j + r] + k[-i - j + r] [i >= 1] + k[2(n-1) - i - j + r] [i <= n-2])   (taps outside 0..2r are 0)
#define WB_BLUR_MAX_K 31
#define WB_BLUR_T 32
#define WB_BLUR_IN (WB_BLUR_T + WB_BLUR_MAX_K - 1)

WB_DEV int wb_reflect(int p, int n) {   // torch 'reflect': -1 -> 1, n -> n-2 (pad < n)
  p = p < 0 ? -p : p;
  return p > n - 1 ? 2 * (n - 1) - p : p;
}
// tap weight of the adjoint along one axis: output i, input j = i + t - r  (see above); kk = taps, K = 2r + 1
WB_DEV float wb_blur_adj_w(const float* kk, int K, int r, int i, int t, int n) {
  float w = kk[t];                                  // k[i - j + r] = k[2r - t] = k[t] (symmetric)
  const int tl = 2 * r - 2 * i - t;                 // -i - j + r
  if (i >= 1 && tl >= 0 && tl < K) w += kk[tl];
  const int tr = 2 * (n - 1) - 2 * i - t + 2 * r;   // 2(n-1) - i - j + r
  if (i <= n - 2 && tr >= 0 && tr < K) w += kk[tr];
  return w;
}

template <bool ADJ>
__global__ void __launch_bounds__(256) k_blur(waldo_blur_t p) {
  const int K = p.ksize, r = K / 2, H = p.H, W = p.W;
  __shared__ float s_k[WB_BLUR_MAX_K];
  __shared__ float s_in[WB_BLUR_IN][WB_BLUR_IN + 1];
  __shared__ float s_h[WB_BLUR_IN][WB_BLUR_T + 1];   // after the horizontal pass: rows of the tile + halo, columns of the tile
  const int tid = wb_tid(), nthr = wb_nthr();
  if (tid == 0) {   // the taps exactly as torchvision builds them: linspace(-r, r, K), exp(-0.5 (x / sigma)^2), normalised
    float sum = 0.f;
    for (int t = 0; t < K; ++t) {
      const float x = (float)(t - r) / p.sigma;
      s_k[t] = expf(-0.5f * x * x);
      sum += s_k[t];
    }
    for (int t = 0; t < K; ++t) s_k[t] = s_k[t] / sum;
  }
  const int tiles_x = (W + WB_BLUR_T - 1) / WB_BLUR_T, tiles_y = (H + WB_BLUR_T - 1) / WB_BLUR_T;
  const int plane = blockIdx.y;
  const float* in = p.in + (size_t)plane * H * W;
  float* out = p.out + (size_t)plane * H * W;
  for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
    const int ty0 = (tile / tiles_x) * WB_BLUR_T, tx0 = (tile % tiles_x) * WB_BLUR_T;
    const int IN = WB_BLUR_T + 2 * r;
    __syncthreads();   // (taps ready; previous tile's passes done)
    for (int i = tid; i < IN * IN; i += nthr) {
      const int ly = i / IN, lx = i - ly * IN;
      const int gy = ty0 + ly - r, gx = tx0 + lx - r;
      float v;
      if (ADJ) v = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(in + (size_t)gy * W + gx) : 0.f;
      else v = __ldg(in + (size_t)wb_reflect(gy, H) * W + wb_reflect(gx, W));   // (positions beyond the halo of the image are never used)
      s_in[ly][lx] = v;
    }
    __syncthreads();
    // Each thread produces 4 consecutive outputs of a pass from one sliding window (K + 3 shared-memory reads for 4 K
    // multiply-adds).  The adjoint's folded weights are only needed by tiles within r of an image border.
    const bool fold_x = ADJ && (tx0 <= r || tx0 + WB_BLUR_T - 1 >= W - 1 - r);
    const bool fold_y = ADJ && (ty0 <= r || ty0 + WB_BLUR_T - 1 >= H - 1 - r);
    for (int i = tid; i < IN * (WB_BLUR_T / 4); i += nthr) {   // horizontal pass
      const int ly = i / (WB_BLUR_T / 4), ox = (i - ly * (WB_BLUR_T / 4)) * 4;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (fold_x) {
        WB_UNROLL for (int j = 0; j < 4; ++j)
          for (int t = 0; t < K; ++t) acc[j] += wb_blur_adj_w(s_k, K, r, tx0 + ox + j, t, W) * s_in[ly][ox + j + t];
      } else {
        for (int t = 0; t < K + 3; ++t) {
          const float v = s_in[ly][ox + t];
          WB_UNROLL for (int j = 0; j < 4; ++j) { const int tt = t - j; if (tt >= 0 && tt < K) acc[j] += s_k[tt] * v; }
        }
      }
      WB_UNROLL for (int j = 0; j < 4; ++j) s_h[ly][ox + j] = acc[j];
    }
    __syncthreads();
    for (int i = tid; i < (WB_BLUR_T / 4) * WB_BLUR_T; i += nthr) {   // vertical pass
      const int oyq = i / WB_BLUR_T, ox = i - oyq * WB_BLUR_T, oy = oyq * 4;
      const int gx = tx0 + ox;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (fold_y) {
        WB_UNROLL for (int j = 0; j < 4; ++j)
          if (ty0 + oy + j < H) for (int t = 0; t < K; ++t) acc[j] += wb_blur_adj_w(s_k, K, r, ty0 + oy + j, t, H) * s_h[oy + j + t][ox];
      } else {
        for (int t = 0; t < K + 3; ++t) {
          const float v = s_h[oy + t][ox];
          WB_UNROLL for (int j = 0; j < 4; ++j) { const int tt = t - j; if (tt >= 0 && tt < K) acc[j] += s_k[tt] * v; }
        }
      }
      if (gx < W) { WB_UNROLL for (int j = 0; j < 4; ++j) if (ty0 + oy + j < H) out[(size_t)(ty0 + oy + j) * W + gx] = acc[j]; }
    }
  }
}

// ------------------------------------------------------------------------------------------------ layer entropy
// synthesizer.py:886-889:  x_k = (alpha_k + 1) / 2 + 1e-6;  p = F.normalize(x, p=1, dim=layers) = x / max(sum_k |x_k|, 1e-12);
//                          entropy = -sum_k p_k log(p_k + 1e-6) / 0.37
// synthesizer.py:933:      fg_mask = sum_{k >= 1} (alpha_k + 1) / 2
// One thread per pixel walks the L layer planes (coalesced 128 B lines); HBM-bound streaming, (L + 2) floats per pixel.
__global__ void __launch_bounds__(256) k_layer_entropy_fwd(waldo_layer_entropy_t p) {
  const size_t HW = (size_t)p.HW;
  const int f = blockIdx.y;
  const float* a = p.alpha + (size_t)f * p.L * HW;
  for (size_t q = (size_t)blockIdx.x * wb_nthr() + wb_tid(); q < HW; q += (size_t)gridDim.x * wb_nthr()) {
    float s = 0.f, fg = 0.f;
    for (int k = 0; k < p.L; ++k) {
      const float h = (__ldg(a + (size_t)k * HW + q) + 1.f) * 0.5f;
      s += fabsf(h + 1e-6f);
      if (k >= 1) fg += h;
    }
    const float inv = 1.f / fmaxf(s, 1e-12f);
    float e = 0.f;
    for (int k = 0; k < p.L; ++k) {
      const float pk = ((__ldg(a + (size_t)k * HW + q) + 1.f) * 0.5f + 1e-6f) * inv;
      e += pk * logf(pk + 1e-6f);
    }
    if (p.entropy) p.entropy[(size_t)f * HW + q] = -e / 0.37f;
    if (p.fg) p.fg[(size_t)f * HW + q] = fg;
  }
}
// d alpha_j = 0.5 * ( (q_j - sgn(x_j) sum_k q_k p_k) / s * d entropy  +  [j >= 1] d fg ),
//   q_k = d e / d p_k = -(log(p_k + 1e-6) + p_k / (p_k + 1e-6)) / 0.37
__global__ void __launch_bounds__(256) k_layer_entropy_bwd(waldo_layer_entropy_bwd_t pb) {
  const waldo_layer_entropy_t& p = pb.f;
  const size_t HW = (size_t)p.HW;
  const int f = blockIdx.y;
  const float* a = p.alpha + (size_t)f * p.L * HW;
  float* da = pb.d_alpha + (size_t)f * p.L * HW;
  for (size_t q = (size_t)blockIdx.x * wb_nthr() + wb_tid(); q < HW; q += (size_t)gridDim.x * wb_nthr()) {
    const float ge = pb.d_entropy ? __ldg(pb.d_entropy + (size_t)f * HW + q) : 0.f;
    const float gf = pb.d_fg ? __ldg(pb.d_fg + (size_t)f * HW + q) : 0.f;
    float s = 0.f;
    for (int k = 0; k < p.L; ++k) s += fabsf((__ldg(a + (size_t)k * HW + q) + 1.f) * 0.5f + 1e-6f);
    const bool clamped = s < 1e-12f;   // F.normalize's eps branch: p = x / eps, no dependence of the norm on x
    const float inv = 1.f / fmaxf(s, 1e-12f);
    float dot = 0.f;
    for (int k = 0; k < p.L; ++k) {
      const float pk = ((__ldg(a + (size_t)k * HW + q) + 1.f) * 0.5f + 1e-6f) * inv;
      const float qk = -(logf(pk + 1e-6f) + pk / (pk + 1e-6f)) / 0.37f;
      dot += qk * pk;
    }
    if (clamped) dot = 0.f;
    for (int k = 0; k < p.L; ++k) {
      const float x = (__ldg(a + (size_t)k * HW + q) + 1.f) * 0.5f + 1e-6f;
      const float pk = x * inv;
      const float qk = -(logf(pk + 1e-6f) + pk / (pk + 1e-6f)) / 0.37f;
      const float sg = x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
      da[(size_t)k * HW + q] = 0.5f * ((qk - sg * dot) * inv * ge + (k >= 1 ? gf : 0.f));
    }
  }
}
