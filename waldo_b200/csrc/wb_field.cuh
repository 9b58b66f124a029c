// a-5 / a-11: the generic "warp a per-layer field into the image" primitive and the bilinear `scale` helper.
// Reference: Warper.obj_to_output / bg_to_output (models/nets/lvd.py:538-559): grid_sample(x + delta, src_grid) - delta,
// bilinear / zeros padding / align_corners=False; `scale` (lvd.py:175-179): F.interpolate(bilinear, scale_factor).
// Forward only: inside decode_output these steps are fused into the decode kernels (which carry the gradients); the
// stand-alone forms serve the MAT propagation flows (lvd.py:575-600) and direct callers at inference.
#pragma once
#include "wb_common.cuh"
#include "../../include/waldo_b200.h"

// one thread per (item, output cell): the four taps are shared by the c channels
__global__ void __launch_bounds__(256) k_warp_field(waldo_warp_field_t p) {
  const int HW = p.H * p.W, hw = p.h * p.w;
  const long long total = (long long)p.n * HW;
  for (long long e = (long long)blockIdx.x * wb_nthr() + wb_tid(); e < total; e += (long long)gridDim.x * wb_nthr()) {
    const int q = (int)(e % HW);
    const long long item = e / HW;
    const float* g = p.grid + e * 2;
    const WbTaps t = wb_taps(__ldg(g), __ldg(g + 1), p.w, p.h);
    const int m = wb_tap_mask(t, p.w, p.h);
    const long long off = (long long)t.y0 * p.w + t.x0;
    for (int ch = 0; ch < p.c; ++ch) {
      const float* pl = p.field + (item * p.c + ch) * hw + off;
      const float vnw = (m & 1) ? __fadd_rn(__ldg(pl), p.delta) : 0.f, vne = (m & 2) ? __fadd_rn(__ldg(pl + 1), p.delta) : 0.f;
      const float vsw = (m & 4) ? __fadd_rn(__ldg(pl + p.w), p.delta) : 0.f, vse = (m & 8) ? __fadd_rn(__ldg(pl + p.w + 1), p.delta) : 0.f;
      p.out[(item * p.c + ch) * HW + q] = __fsub_rn(wb_chain(vnw, vne, vsw, vse, t), p.delta);
    }
  }
}

// upsample_bilinear2d(align_corners=False) of n planes h x w -> H x W, source coordinate ratio h/H (ATen with a scale factor)
__global__ void __launch_bounds__(256) k_resize_bilinear(waldo_resize_t p) {
  const int HW = p.H * p.W, hw = p.h * p.w;
  const float rh = (float)p.h / (float)p.H, rw = (float)p.w / (float)p.W;
  const long long total = (long long)p.n * HW;
  for (long long e = (long long)blockIdx.x * wb_nthr() + wb_tid(); e < total; e += (long long)gridDim.x * wb_nthr()) {
    const int q = (int)(e % HW), Y = q / p.W, X = q - Y * p.W;
    const float* pl = p.in + (e / HW) * hw;
    if (p.h == p.H && p.w == p.W) { p.out[e] = __ldg(pl + q); continue; }
    const WbAxis ay = wb_axis(Y, rh, p.h), ax = wb_axis(X, rw, p.w);
    p.out[e] = wb_lerp2(__ldg(pl + ay.i0 * p.w + ax.i0), __ldg(pl + ay.i0 * p.w + ax.i1), __ldg(pl + ay.i1 * p.w + ax.i0),
                        __ldg(pl + ay.i1 * p.w + ax.i1), ax, ay);
  }
}
