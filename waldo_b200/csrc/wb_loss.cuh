// f-2: loss epilogues of LVD training over the path's low-res outputs.
// Reference: models/synthesizer.py:1114-1118 `blur` (torchvision GaussianBlur, reflect padding),
// :886-892 the layer-entropy regulariser, :933 `fg_mask`, :965-979 `cell_dis` / `center_dis`; all forward and backward.
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

// ------------------------------------------------------------------------------------------------ pose distances
// synthesizer.py:965-979 (`cell_dis`, `center_dis`): squared distances between every lattice point g of the low-res grid and
//   c_{o,k} = the centre of cell k of object o's control-point lattice (mean of its 4 corners), summed over the cells, and
//   m_o     = the mean of all control points of object o,
// each in the expanded form the reference uses (|g|^2 + |c|^2 - 2 c.g), weighted per pixel and minimised over the objects:
//   cell_min   = min_o ( (mov + eps) (1 - fg) * sum_k d(g, c_{o,k}) )        center_min = min_o ( mov * d(g, m_o) )
// The reference materialises the (B, T, No, cells, H, W) distance tensor through a K = 2 matmul (755 MB at B = 8, T = 5,
// 16 objects, 9 cells, 128 x 256); here one thread per pixel walks the objects with the centres staged in shared memory:
// 4 floats in, 2 floats + 2 bytes out per pixel.  The argmin objects are saved for the backward (first index on ties).
#define WB_PD_MAX_NO 32
#define WB_PD_MAX_CELLS 1024   // No * (ho - 1) * (wo - 1)
#define WB_PD_THREADS 256

// cell centres (x, y, |c|^2) and object means (x, y, |m|^2) of frame f into shared memory
WB_DEV void wb_pd_stage(const waldo_pose_dis_t& p, int f, float* s_c, float* s_m) {
  const int K = p.ho * p.wo, ncell = (p.ho - 1) * (p.wo - 1);
  const float* pose = p.pose + (size_t)f * p.No * K * 2;
  for (int i = wb_tid(); i < p.No * ncell; i += wb_nthr()) {
    const int o = i / ncell, k = i - o * ncell, cy = k / (p.wo - 1), cx = k - cy * (p.wo - 1);
    const float* q = pose + ((size_t)o * K + (size_t)cy * p.wo + cx) * 2;   // corner (cy, cx); the reference adds (1,1) + (1,0) + (0,1) + (0,0)
    const float x = (((q[(p.wo + 1) * 2] + q[p.wo * 2]) + q[2]) + q[0]) / 4.f;
    const float y = (((q[(p.wo + 1) * 2 + 1] + q[p.wo * 2 + 1]) + q[3]) + q[1]) / 4.f;
    s_c[i * 3] = x; s_c[i * 3 + 1] = y; s_c[i * 3 + 2] = x * x + y * y;
  }
  for (int o = wb_tid(); o < p.No; o += wb_nthr()) {
    float sx = 0.f, sy = 0.f;
    for (int k = 0; k < K; ++k) { sx += pose[((size_t)o * K + k) * 2]; sy += pose[((size_t)o * K + k) * 2 + 1]; }
    const float x = sx / (float)K, y = sy / (float)K;
    s_m[o * 3] = x; s_m[o * 3 + 1] = y; s_m[o * 3 + 2] = x * x + y * y;
  }
}
WB_DEV float wb_pd_cell_sum(const float* s_c, int o, int ncell, float gx, float gy, float g2) {
  float d = 0.f;
  for (int k = 0; k < ncell; ++k) {
    const float* c = s_c + (o * ncell + k) * 3;
    d += (g2 + c[2]) - 2.f * (c[0] * gx + c[1] * gy);
  }
  return d;
}
WB_DEV float wb_pd_center(const float* s_m, int o, float gx, float gy, float g2) {
  return (g2 + s_m[o * 3 + 2]) - 2.f * (s_m[o * 3] * gx + s_m[o * 3 + 1] * gy);
}

__global__ void __launch_bounds__(WB_PD_THREADS) k_pose_dis_fwd(waldo_pose_dis_t p) {
  __shared__ float s_c[WB_PD_MAX_CELLS * 3];
  __shared__ float s_m[WB_PD_MAX_NO * 3];
  const int f = blockIdx.y, ncell = (p.ho - 1) * (p.wo - 1);
  wb_pd_stage(p, f, s_c, s_m);
  __syncthreads();
  const size_t base = (size_t)f * p.HW;
  for (int q = blockIdx.x * wb_nthr() + wb_tid(); q < p.HW; q += gridDim.x * wb_nthr()) {
    const float gx = __ldg(p.grid + 2 * q), gy = __ldg(p.grid + 2 * q + 1), g2 = gx * gx + gy * gy;
    const float mov = __ldg(p.mov + base + q), fg = __ldg(p.fg + base + q);
    const float w = (mov + p.eps) * (1.f - fg);
    float bc = 0.f, bm = 0.f;
    int ac = 0, am = 0;
    for (int o = 0; o < p.No; ++o) {
      const float vc = w * wb_pd_cell_sum(s_c, o, ncell, gx, gy, g2);
      const float vm = mov * wb_pd_center(s_m, o, gx, gy, g2);
      if (o == 0 || vc < bc) { bc = vc; ac = o; }
      if (o == 0 || vm < bm) { bm = vm; am = o; }
    }
    p.cell_min[base + q] = bc; p.center_min[base + q] = bm;
    p.cell_arg[base + q] = (unsigned char)ac; p.center_arg[base + q] = (unsigned char)am;
  }
}

// Backward, pixel pass.  With o* / o' the saved argmins, Dc = sum_k d(g, c_{o*,k}), Dm = d(g, m_{o'}):
//   d fg  = -d cell_min (mov + eps) Dc          d mov = d cell_min (1 - fg) Dc + d center_min Dm
//   d c_{o*,k} += wc (2 c_{o*,k} - 2 g),  wc = d cell_min (mov + eps)(1 - fg)      d m_{o'} += wm (2 m_{o'} - 2 g),  wm = d center_min mov
// The sums over pixels only need S0 = sum w and S1 = sum w g per (frame, object) and term: six floats.  Each warp adds its lanes'
// values object by object (only the objects present in the warp) into its own shared-memory row, the rows are added in warp order
// and leave the CTA as one partial per (frame, CTA, object): no atomics, the same bits on every run.
__global__ void __launch_bounds__(WB_PD_THREADS) k_pose_dis_bwd(waldo_pose_dis_bwd_t b) {
  const waldo_pose_dis_t& p = b.f;
  __shared__ float s_c[WB_PD_MAX_CELLS * 3];
  __shared__ float s_m[WB_PD_MAX_NO * 3];
  __shared__ float s_acc[WB_PD_THREADS / 32][WB_PD_MAX_NO * 6];
  const int f = blockIdx.y, ncell = (p.ho - 1) * (p.wo - 1);
  wb_pd_stage(p, f, s_c, s_m);
  for (int i = wb_tid(); i < (WB_PD_THREADS / 32) * WB_PD_MAX_NO * 6; i += wb_nthr()) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  float* acc = s_acc[wb_warp()];
  const size_t base = (size_t)f * p.HW;
  const int span = gridDim.x * wb_nthr();
  for (int q0 = blockIdx.x * wb_nthr(); q0 < p.HW; q0 += span) {   // whole warps stay in the loop: the reductions need every lane
    const int q = q0 + wb_tid();
    const bool on = q < p.HW;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int oc = 0, om = 0;
    if (on) {
      const float gx = __ldg(p.grid + 2 * q), gy = __ldg(p.grid + 2 * q + 1), g2 = gx * gx + gy * gy;
      const float mov = __ldg(p.mov + base + q), fg = __ldg(p.fg + base + q);
      const float gc = b.d_cell ? __ldg(b.d_cell + base + q) : 0.f, gm = b.d_center ? __ldg(b.d_center + base + q) : 0.f;
      oc = p.cell_arg[base + q]; om = p.center_arg[base + q];
      const float Dc = wb_pd_cell_sum(s_c, oc, ncell, gx, gy, g2), Dm = wb_pd_center(s_m, om, gx, gy, g2);
      if (b.d_fg) b.d_fg[base + q] = -gc * (mov + p.eps) * Dc;
      if (b.d_mov) b.d_mov[base + q] = gc * (1.f - fg) * Dc + gm * Dm;
      const float wc = gc * ((mov + p.eps) * (1.f - fg)), wm = gm * mov;
      v[0] = wc; v[1] = wc * gx; v[2] = wc * gy; v[3] = wm; v[4] = wm * gx; v[5] = wm * gy;
    }
    unsigned present = wb_warp_or(on ? ((1u << oc) | (1u << om)) : 0u);
    while (present) {
      const int o = __ffs((int)present) - 1;
      present &= present - 1u;
      WB_UNROLL for (int j = 0; j < 6; ++j) {
        const bool mine = on && (j < 3 ? oc == o : om == o);
        const float s = wb_warp_sum(mine ? v[j] : 0.f);
        if (wb_lane() == 0) acc[o * 6 + j] += s;
      }
    }
  }
  __syncthreads();
  float* part = b.part + ((size_t)f * gridDim.x + blockIdx.x) * p.No * 6;
  for (int i = wb_tid(); i < p.No * 6; i += wb_nthr()) {
    float s = 0.f;
    WB_UNROLL for (int w = 0; w < WB_PD_THREADS / 32; ++w) s += s_acc[w][i];   // (emulation: one warp, the other rows stay zero)
    part[i] = s;
  }
}
// one thread per (frame, object): partials added in CTA order, then spread over the control points
//   d pose_{o,(i,j)} = sum_{cells k that have (i,j) as a corner} (2 c_k S0c - 2 S1c) / 4  +  (2 m S0m - 2 S1m) / (ho wo)
__global__ void __launch_bounds__(128) k_pose_dis_bwd_final(waldo_pose_dis_bwd_t b) {
  const waldo_pose_dis_t& p = b.f;
  const int K = p.ho * p.wo;
  for (int i = blockIdx.x * wb_nthr() + wb_tid(); i < p.n * p.No; i += gridDim.x * wb_nthr()) {
    const int f = i / p.No, o = i - f * p.No;
    float S[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < b.ctas; ++c) {
      const float* part = b.part + (((size_t)f * b.ctas + c) * p.No + o) * 6;
      WB_UNROLL for (int j = 0; j < 6; ++j) S[j] += part[j];
    }
    const float* pose = p.pose + ((size_t)f * p.No + o) * K * 2;
    float* dp = b.d_pose + ((size_t)f * p.No + o) * K * 2;
    float mx = 0.f, my = 0.f;
    for (int k = 0; k < K; ++k) { mx += pose[k * 2]; my += pose[k * 2 + 1]; }
    mx /= (float)K; my /= (float)K;
    const float cmx = (2.f * mx * S[3] - 2.f * S[4]) / (float)K, cmy = (2.f * my * S[3] - 2.f * S[5]) / (float)K;
    for (int y = 0; y < p.ho; ++y)
      for (int x = 0; x < p.wo; ++x) {
        float gx = cmx, gy = cmy;
        for (int cy = y - 1; cy <= y; ++cy)
          for (int cx = x - 1; cx <= x; ++cx) {
            if (cy < 0 || cx < 0 || cy >= p.ho - 1 || cx >= p.wo - 1) continue;
            const float* q = pose + ((size_t)cy * p.wo + cx) * 2;
            const float ccx = (((q[(p.wo + 1) * 2] + q[p.wo * 2]) + q[2]) + q[0]) / 4.f;
            const float ccy = (((q[(p.wo + 1) * 2 + 1] + q[p.wo * 2 + 1]) + q[3]) + q[1]) / 4.f;
            gx += (2.f * ccx * S[0] - 2.f * S[1]) * 0.25f;
            gy += (2.f * ccy * S[0] - 2.f * S[2]) * 0.25f;
          }
        dp[(y * p.wo + x) * 2] = gx; dp[(y * p.wo + x) * 2 + 1] = gy;
      }
  }
}

// ------------------------------------------------------------------------------------------------ obj_flow
// synthesizer.py:864-868 ("same mean motion in layers"):  a_o = (alpha_{o+1} + 1) / 2 + 1e-6  (o = 0 .. L-2, the object layers),
//   S_o = sum_p a_o,  m_o = sum_p a_o f / S_o  (f = real_flow, 2 components),  obj_flow = mean_{o,p} a_o (|fx - mx_o| + |fy - my_o|).
// The reference builds (B, T, No, 2, H, W) products for the means and again for the deviations; here
//   k_of_moments (+ final): per (frame, object) S, sum a fx, sum a fy   -- per-CTA partials, added in CTA order
//   k_of_map:               per pixel  V = sum_o a_o (|fx - mx_o| + |fy - my_o|)          (obj_flow = sum V / (n No HW))
// Backward, with gV the upstream gradient of V:
//   d a_o(p) = gV(p) D_o(p) - ((fx(p) - mx_o) Tx_o + (fy(p) - my_o) Ty_o) / S_o,   T_o = sum_q gV(q) a_o(q) sign(f(q) - m_o)
//   k_of_tsum (+ final): T;   k_of_dalpha: d alpha_{o+1} = d a_o / 2, d alpha_0 = 0.
#define WB_OF_THREADS 256
#define WB_OF_MAX_L 33

// sum of `nv` values per thread over the CTA, in a fixed order (lanes by shuffles, warps one after the other); result valid in thread 0
template <int NV>
WB_DEV void wb_of_block_sum(float (&v)[NV], float (*s_w)[4]) {
  WB_UNROLL for (int j = 0; j < NV; ++j) v[j] = wb_warp_sum(v[j]);
  __syncthreads();   // (the previous use of s_w is over)
  if (wb_lane() == 0) { WB_UNROLL for (int j = 0; j < NV; ++j) s_w[wb_warp()][j] = v[j]; }
  __syncthreads();
  if (wb_tid() == 0) {
    const int nw = (wb_nthr() + 31) / 32;
    WB_UNROLL for (int j = 0; j < NV; ++j) { float s = 0.f; for (int w = 0; w < nw; ++w) s += s_w[w][j]; v[j] = s; }
  }
}

// MODE 0: moments (a, a fx, a fy) -> part (n, ctas, L-1, 3);  MODE 1: T sums (gV a sign(fx - mx), gV a sign(fy - my)) -> part (n, ctas, L-1, 2)
template <int MODE>
__global__ void __launch_bounds__(WB_OF_THREADS) k_of_reduce(waldo_obj_flow_t p, const float* d_map) {
  __shared__ float s_w[WB_OF_THREADS / 32][4];
  const int f = blockIdx.y, No = p.L - 1, NV = MODE == 0 ? 3 : 2;
  const float* fx = p.flow + (size_t)f * 2 * p.HW;
  const float* fy = fx + p.HW;
  const float* gv = MODE == 1 ? d_map + (size_t)f * p.HW : nullptr;
  float* part = p.part + ((size_t)f * gridDim.x + blockIdx.x) * No * NV;
  for (int o = 0; o < No; ++o) {
    const float* al = p.alpha + ((size_t)f * p.L + o + 1) * p.HW;
    float mx = 0.f, my = 0.f;
    if (MODE == 1) { const float* m = p.mom + ((size_t)f * No + o) * 3; mx = m[1] / m[0]; my = m[2] / m[0]; }
    float v[3] = {0.f, 0.f, 0.f};
    for (int q = blockIdx.x * wb_nthr() + wb_tid(); q < p.HW; q += gridDim.x * wb_nthr()) {
      const float a = (__ldg(al + q) + 1.f) * 0.5f + 1e-6f, x = __ldg(fx + q), y = __ldg(fy + q);
      if (MODE == 0) { v[0] += a; v[1] += a * x; v[2] += a * y; }
      else {
        const float g = __ldg(gv + q) * a, dx = x - mx, dy = y - my;
        v[0] += g * (dx > 0.f ? 1.f : (dx < 0.f ? -1.f : 0.f));
        v[1] += g * (dy > 0.f ? 1.f : (dy < 0.f ? -1.f : 0.f));
      }
    }
    wb_of_block_sum<3>(v, s_w);
    if (wb_tid() == 0) { for (int j = 0; j < NV; ++j) part[o * NV + j] = v[j]; }
  }
}
// one thread per (frame, object, value): partials added in CTA order
__global__ void __launch_bounds__(128) k_of_reduce_final(const float* part, float* out, int n, int ctas, int per_frame) {
  for (int i = blockIdx.x * wb_nthr() + wb_tid(); i < n * per_frame; i += gridDim.x * wb_nthr()) {
    const int f = i / per_frame, e = i - f * per_frame;
    float s = 0.f;
    for (int c = 0; c < ctas; ++c) s += part[((size_t)f * ctas + c) * per_frame + e];
    out[i] = s;
  }
}
__global__ void __launch_bounds__(WB_OF_THREADS) k_of_map(waldo_obj_flow_t p) {
  __shared__ float s_m[WB_OF_MAX_L * 2];
  const int f = blockIdx.y, No = p.L - 1;
  for (int o = wb_tid(); o < No; o += wb_nthr()) {
    const float* m = p.mom + ((size_t)f * No + o) * 3;
    s_m[o * 2] = m[1] / m[0]; s_m[o * 2 + 1] = m[2] / m[0];
  }
  __syncthreads();
  const float* fx = p.flow + (size_t)f * 2 * p.HW;
  const float* fy = fx + p.HW;
  for (int q = blockIdx.x * wb_nthr() + wb_tid(); q < p.HW; q += gridDim.x * wb_nthr()) {
    const float x = __ldg(fx + q), y = __ldg(fy + q);
    float v = 0.f;
    for (int o = 0; o < No; ++o) {
      const float a = (__ldg(p.alpha + ((size_t)f * p.L + o + 1) * p.HW + q) + 1.f) * 0.5f + 1e-6f;
      v += a * (fabsf(x - s_m[o * 2]) + fabsf(y - s_m[o * 2 + 1]));
    }
    p.dev_map[(size_t)f * p.HW + q] = v;
  }
}
__global__ void __launch_bounds__(WB_OF_THREADS) k_of_dalpha(waldo_obj_flow_bwd_t b) {
  const waldo_obj_flow_t& p = b.f;
  __shared__ float s_m[WB_OF_MAX_L * 5];   // mx, my, Tx / S, Ty / S per object
  const int f = blockIdx.y, No = p.L - 1;
  for (int o = wb_tid(); o < No; o += wb_nthr()) {
    const float* m = p.mom + ((size_t)f * No + o) * 3;
    const float* t = b.tsum + ((size_t)f * No + o) * 2;
    s_m[o * 4] = m[1] / m[0]; s_m[o * 4 + 1] = m[2] / m[0]; s_m[o * 4 + 2] = t[0] / m[0]; s_m[o * 4 + 3] = t[1] / m[0];
  }
  __syncthreads();
  const float* fx = p.flow + (size_t)f * 2 * p.HW;
  const float* fy = fx + p.HW;
  float* da = b.d_alpha + (size_t)f * p.L * p.HW;
  for (int q = blockIdx.x * wb_nthr() + wb_tid(); q < p.HW; q += gridDim.x * wb_nthr()) {
    const float x = __ldg(fx + q), y = __ldg(fy + q), g = __ldg(b.d_map + (size_t)f * p.HW + q);
    da[q] = 0.f;
    for (int o = 0; o < No; ++o) {
      const float dx = x - s_m[o * 4], dy = y - s_m[o * 4 + 1];
      da[(size_t)(o + 1) * p.HW + q] = 0.5f * (g * (fabsf(dx) + fabsf(dy)) - (dx * s_m[o * 4 + 2] + dy * s_m[o * 4 + 3]));
    }
  }
}
