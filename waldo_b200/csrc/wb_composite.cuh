// The fused HD kernel: per-layer flow up-sampling, warp of the context opacities, occlusion-aware
// compositing, flow reduction, warp of the context frame and fusion over contexts -- B5(up)..B9 + stage C.
// Reference: models/nets/lvd.py:794-818 (/ :671-695) and :830-853.  No per-layer HD tensor ever reaches HBM.
#pragma once
#include "wb_common.cuh"
#include "wb_prep.cuh"

// Everything one (b,tp) pixel needs that does not depend on the context.
struct WbPix {
  int Y, X;
  float gx, gy;            // identity grid (buffer src_grid_hd)
  WbAxis ax, ay;           // up-sampling taps into the low-res lattice
  int o00, o01, o10, o11;
  unsigned isobj;          // bit k set <=> layer k may show at this pixel (lvd.py:788-791), bit 0 always
};

WB_DEV WbPix wb_pix(const WbDec& d, int b, int tp, size_t q) {
  const waldo_geom_t& g = d.g;
  WbPix px;
  px.Y = (int)(q / g.Wd); px.X = (int)(q - (size_t)px.Y * g.Wd);
  px.gx = __ldg(d.xs_hd + px.X); px.gy = __ldg(d.ys_hd + px.Y);
  const float r = (float)g.H / (float)g.Hd;
  px.ay = wb_axis(px.Y, r, g.H); px.ax = wb_axis(px.X, r, g.W);
  px.o00 = px.ay.i0 * g.W + px.ax.i0; px.o01 = px.ay.i0 * g.W + px.ax.i1;
  px.o10 = px.ay.i1 * g.W + px.ax.i0; px.o11 = px.ay.i1 * g.W + px.ax.i1;
  px.isobj = 0xffffffffu;
  if (g.flags & WALDO_F_IS_OBJ) {
    const int HW = g.H * g.W;
    px.isobj = 1u;
    for (int k = 1; k <= g.No; ++k) {
      const float* s = d.s_lo + (((size_t)b * g.Tp + tp) * g.No + (k - 1)) * HW;
      float v = (g.Hd == g.H) ? __ldg(s + px.o00)
                              : wb_lerp2(__ldg(s + px.o00), __ldg(s + px.o01), __ldg(s + px.o10), __ldg(s + px.o11), px.ax, px.ay);
      if (v > 0.9f) px.isobj |= 1u << k;
    }
  }
  return px;
}

// Layers of one (b,tc,tp) pixel: per-layer flow F, warped context opacity R, composited opacity Actx, reduced flow.
struct WbLayers {
  float Fx[WB_MAX_L], Fy[WB_MAX_L];
  float R[WB_MAX_L];
  float A[WB_MAX_L];
  float flow_x, flow_y, score, disocc;
};

// alpha_plane = stored context alpha (2A-1) of frame (b,c): (L, Hd, Wd)
WB_DEV void wb_layers_fwd(const WbDec& d, const WbPix& px, const float* __restrict__ f_lo /* (L,H,W,2) of this pair */,
                          const float* __restrict__ alpha_c, const float* __restrict__ s_occ, WbLayers& ly) {
  const waldo_geom_t& g = d.g;
  const int L = g.No + 1, HW = g.H * g.W;
  const size_t HWd = (size_t)g.Hd * g.Wd;
  float mx = -INFINITY;
  WB_UNROLL for (int k = 0; k < WB_MAX_L; ++k) {
    if (k < L) {
      const float2* fl = reinterpret_cast<const float2*>(f_lo) + (size_t)k * HW;
      float fx, fy;
      if (g.Hd == g.H) { float2 v = __ldg(fl + px.o00); fx = v.x; fy = v.y; }
      else {
        float2 v00 = __ldg(fl + px.o00), v01 = __ldg(fl + px.o01), v10 = __ldg(fl + px.o10), v11 = __ldg(fl + px.o11);
        fx = wb_lerp2(v00.x, v01.x, v10.x, v11.x, px.ax, px.ay);
        fy = wb_lerp2(v00.y, v01.y, v10.y, v11.y, px.ax, px.ay);
      }
      ly.Fx[k] = fx; ly.Fy[k] = fy;
      float r = 0.f;
      if ((px.isobj >> k) & 1u) {
        WbTaps t = wb_taps(__fadd_rn(px.gx, fx), __fadd_rn(px.gy, fy), g.Wd, g.Hd);
        int m = wb_tap_mask(t, g.Wd, g.Hd);
        const float* p = alpha_c + (size_t)k * HWd + (long long)t.y0 * g.Wd + t.x0;
        // stored value is 2A-1; zero padding applies to A, so out-of-range taps contribute 0
        float vnw = (m & 1) ? (__ldg(p) + 1.f) * 0.5f : 0.f;
        float vne = (m & 2) ? (__ldg(p + 1) + 1.f) * 0.5f : 0.f;
        float vsw = (m & 4) ? (__ldg(p + g.Wd) + 1.f) * 0.5f : 0.f;
        float vse = (m & 8) ? (__ldg(p + g.Wd + 1) + 1.f) * 0.5f : 0.f;
        r = wb_chain(vnw, vne, vsw, vse, t);
      }
      ly.R[k] = r;
      mx = fmaxf(mx, r);
    }
  }
  ly.disocc = mx;
  float fx = 0.f, fy = 0.f, sc = 0.f;
  WB_UNROLL for (int i = 0; i < WB_MAX_L; ++i) {
    if (i < L) {
      float vis = 1.f;
      WB_UNROLL for (int j = 0; j < WB_MAX_L; ++j) if (j < L) vis *= 1.f - ly.R[j] * s_occ[j * L + i];
      float a = vis * ly.R[i];
      ly.A[i] = a;
      fx += a * ly.Fx[i]; fy += a * ly.Fy[i]; sc += a;
    }
  }
  ly.flow_x = fx; ly.flow_y = fy; ly.score = sc;
}

// grid = (pixel chunks, B*Tp); one thread per HD pixel, contexts looped inside so that the fused `output`
// (lvd.py:850-851) never leaves registers.
__global__ void __launch_bounds__(256) k_warp_composite_fwd(WbDec d) {
  const waldo_geom_t g = d.g;
  const int L = g.No + 1, HW = g.H * g.W, C = g.C;
  const size_t HWd = (size_t)g.Hd * g.Wd;
  const int btp = blockIdx.y, b = btp / g.Tp, tp = btp - b * g.Tp;
  const int u = (int)d.pred_ts[tp];
  const bool self = (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
  const bool disocc_ch = (g.flags & WALDO_F_USE_DISOCC) != 0;
  const int TcR = g.Tc + (self ? 1 : 0), CR = C + L + (disocc_ch ? 1 : 0);
  __shared__ float s_occ[WB_MAX_L * WB_MAX_L];
  for (int i = wb_tid(); i < L * L; i += wb_nthr()) s_occ[i] = __ldg(d.occ + ((size_t)b * g.T + u) * L * L + i);
  __syncthreads();
  for (size_t q = (size_t)blockIdx.x * wb_nthr() + wb_tid(); q < HWd; q += (size_t)gridDim.x * wb_nthr()) {
    WbPix px = wb_pix(d, b, tp, q);
    float acc[WB_MAX_C + 1];
    WB_UNROLL for (int c = 0; c <= WB_MAX_C; ++c) acc[c] = 0.f;
    float den = 0.f;
    for (int tc = 0; tc < g.Tc; ++tc) {
      const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
      const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
      const float* f_lo = d.f_lo + pair * L * HW * 2;
      const float* alpha_c = d.alpha + ((size_t)b * g.Tw + c_t) * L * HWd;
      WbLayers ly;
      wb_layers_fwd(d, px, f_lo, alpha_c, s_occ, ly);
      float* raw = d.raw_output + (((size_t)b * TcR + tc) * g.Tp + tp) * CR * HWd + q;
      WB_UNROLL for (int k = 0; k < WB_MAX_L; ++k) if (k < L) raw[(size_t)(C + k) * HWd] = ly.A[k] * 2.f - 1.f;
      if (disocc_ch) raw[(size_t)(C + L) * HWd] = ly.disocc;
      float* fl = d.flow + pair * 2 * HWd + q;
      fl[0] = ly.flow_x; fl[HWd] = ly.flow_y;
      // stage C: warp the context frame by the reduced flow
      WbTaps t = wb_taps(__fadd_rn(px.gx, ly.flow_x), __fadd_rn(px.gy, ly.flow_y), g.Wd, g.Hd);
      int m = wb_tap_mask(t, g.Wd, g.Hd);
      const float* src = d.input + ((size_t)b * g.T + c_t) * C * HWd;
      const float wgt = ly.score + 1e-6f;
      WB_UNROLL for (int c = 0; c < WB_MAX_C; ++c) {
        if (c < C) {
          float v = wb_sample(src + (size_t)c * HWd, t, m, g.Wd);
          raw[(size_t)c * HWd] = v;
          acc[c] += wgt * v;
        }
      }
      acc[WB_MAX_C] += wgt * (ly.score * 2.f - 1.f);
      den += wgt;
    }
    if (self) {   // lvd.py:842-845: the target frame itself, fully opaque, score 1
      float* raw = d.raw_output + (((size_t)b * TcR + g.Tc) * g.Tp + tp) * CR * HWd + q;
      const float* src = d.input + ((size_t)b * g.T + tp) * C * HWd + q;
      const float wgt = 1.f + 1e-6f;
      WB_UNROLL for (int c = 0; c < WB_MAX_C; ++c) {
        if (c < C) { float v = __ldg(src + (size_t)c * HWd); raw[(size_t)c * HWd] = v; acc[c] += wgt * v; }
      }
      for (int k = 0; k < L; ++k) raw[(size_t)(C + k) * HWd] = 1.f;
      if (disocc_ch) raw[(size_t)(C + L) * HWd] = 1.f;
      acc[WB_MAX_C] += wgt * 1.f;
      den += wgt;
    }
    const float inv = 1.f / fmaxf(den, 1e-12f);
    float* of = d.out_full + ((size_t)b * g.Tp + tp) * (C + 1) * HWd + q;
    WB_UNROLL for (int c = 0; c < WB_MAX_C; ++c) if (c < C) of[(size_t)c * HWd] = acc[c] * inv;
    of[(size_t)C * HWd] = acc[WB_MAX_C] * inv;
    if (d.norm) d.norm[((size_t)b * g.Tp + tp) * HWd + q] = den;
  }
}
