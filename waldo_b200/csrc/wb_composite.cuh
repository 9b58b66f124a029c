// The fused HD kernel: per-layer flow up-sampling, warp of the context opacities, occlusion-aware
// compositing, flow reduction, warp of the context frame and fusion over contexts -- B5(up)..B9 + stage C.
// Reference: models/nets/lvd.py:794-818 (/ :671-695) and :830-853.  No per-layer HD tensor ever reaches HBM.
//
// Layer sparsity: with is_obj (lvd.py:788-791) a layer can only show where the up-sampled object support exceeds
// 0.9, i.e. at a handful of the 17 layers per pixel.  Each warp gathers the union of its pixels' live layers into a
// compact slot list (WbIdx) and every layer loop runs over those slots only (warp-uniform trip counts, statically
// indexed register arrays of NA slots; NA in {4, 8, 17} picked per warp).  Skipped layers have R = A = 0 exactly and
// their occlusion factors are exactly 1, so the result is bit-identical to the dense evaluation.
#pragma once
#include "wb_common.cuh"
#include "wb_prep.cuh"

// Everything one (b,tp) pixel needs that does not depend on the context.
struct WbPix {
  float gx, gy;            // identity grid (buffer src_grid_hd)
  WbAxis ax, ay;           // up-sampling taps into the low-res lattice
  int o00, o01, o10, o11;
  unsigned isobj;          // bit k set <=> layer k may show at this pixel (lvd.py:788-791), bit 0 always
};

WB_DEV WbPix wb_pix(const WbDec& d, int b, int tp, int X, int Y) {
  const waldo_geom_t& g = d.g;
  WbPix px;
  px.gx = __ldg(d.xs_hd + X); px.gy = __ldg(d.ys_hd + Y);
  const float r = (float)g.H / (float)g.Hd;
  px.ay = wb_axis(Y, r, g.H); px.ax = wb_axis(X, r, g.W);
  px.o00 = px.ay.i0 * g.W + px.ax.i0; px.o01 = px.ay.i0 * g.W + px.ax.i1;
  px.o10 = px.ay.i1 * g.W + px.ax.i0; px.o11 = px.ay.i1 * g.W + px.ax.i1;
  const int L = g.No + 1;
  px.isobj = (1u << L) - 1u;
  if (g.flags & WALDO_F_IS_OBJ) {
    const int HW = g.H * g.W;
    unsigned cand = wb_live4(d.live_pred + ((size_t)b * g.Tp + tp) * HW, px.o00, px.o01, px.o10, px.o11) & px.isobj;
    px.isobj = 1u;
    cand &= ~1u;
    while (cand) {
      const int k = __ffs((int)cand) - 1;
      cand &= cand - 1u;
      const float* s = d.s_lo + (((size_t)b * g.Tp + tp) * g.No + (k - 1)) * HW;
      float v = (g.Hd == g.H) ? __ldg(s + px.o00)
                              : wb_lerp2(__ldg(s + px.o00), __ldg(s + px.o01), __ldg(s + px.o10), __ldg(s + px.o11), px.ax, px.ay);
      if (v > 0.9f) px.isobj |= 1u << k;
    }
  }
  return px;
}

// Layers of one (b,tc,tp) pixel in slot order: per-layer flow F, warped context opacity R, composited opacity A.
template <int NA> struct WbLay {
  float Fx[NA], Fy[NA], R[NA], A[NA];
  float flow_x, flow_y, score, disocc;
};

// alpha_c = stored context alpha (2A-1) of frame (b,c): (L, Hd, Wd)
template <int NA>
WB_DEV void wb_layers_fwd(const WbDec& d, const WbPix& px, const WbIdx<NA>& ix, const float* __restrict__ f_lo /* (L,H,W,2) of this pair */,
                          const float* __restrict__ alpha_c, const float* __restrict__ s_occ, WbLay<NA>& ly) {
  const waldo_geom_t& g = d.g;
  const int L = g.No + 1, HW = g.H * g.W;
  const size_t HWd = (size_t)g.Hd * g.Wd;
  float mx = 0.f;   // layers outside the union have R = 0, and R >= 0 always
  WB_UNROLL_NA for (int s = 0; s < NA; ++s) {
    ly.R[s] = 0.f; ly.A[s] = 0.f; ly.Fx[s] = 0.f; ly.Fy[s] = 0.f;
    if (s < ix.n) {
      const int k = ix.k[s];
      const float2* fl = reinterpret_cast<const float2*>(f_lo) + (size_t)k * HW;
      float fx, fy;
      if (g.Hd == g.H) { float2 v = __ldg(fl + px.o00); fx = v.x; fy = v.y; }
      else {
        float2 v00 = __ldg(fl + px.o00), v01 = __ldg(fl + px.o01), v10 = __ldg(fl + px.o10), v11 = __ldg(fl + px.o11);
        fx = wb_lerp2(v00.x, v01.x, v10.x, v11.x, px.ax, px.ay);
        fy = wb_lerp2(v00.y, v01.y, v10.y, v11.y, px.ax, px.ay);
      }
      ly.Fx[s] = fx; ly.Fy[s] = fy;
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
      ly.R[s] = r;
      mx = fmaxf(mx, r);
    }
  }
  ly.disocc = mx;
  float fx = 0.f, fy = 0.f, sc = 0.f;
  WB_UNROLL_NA for (int i = 0; i < NA; ++i) {
    if (i < ix.n) {
      float vis = 1.f;
      WB_UNROLL_NA for (int j = 0; j < NA; ++j) if (j < ix.n) vis *= 1.f - ly.R[j] * s_occ[ix.k[j] * L + ix.k[i]];
      float a = vis * ly.R[i];
      ly.A[i] = a;
      fx += a * ly.Fx[i]; fy += a * ly.Fy[i]; sc += a;
    }
  }
  ly.flow_x = fx; ly.flow_y = fy; ly.score = sc;
}

struct WbFwdCtx {   // per-CTA constants of the fused forward
  int b, tp, L, C, TcR, CR, HW;
  size_t HWd;
  bool self, disocc_ch;
  const float* s_occ;
};

template <int NA>
WB_DEV void wb_fwd_pixel(const WbDec& d, const WbFwdCtx& c, const WbPix& px, unsigned wm, bool active, size_t q) {
  const waldo_geom_t& g = d.g;
  const int L = c.L, C = c.C, b = c.b, tp = c.tp;
  const size_t HWd = c.HWd;
  const WbIdx<NA> ix = wb_idx<NA>(wm);
  float acc[WB_MAX_C + 1];
  WB_UNROLL for (int ch = 0; ch <= WB_MAX_C; ++ch) acc[ch] = 0.f;
  float den = 0.f;
  for (int tc = 0; tc < g.Tc; ++tc) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
    const float* f_lo = d.f_lo + pair * L * c.HW * 2;
    const float* alpha_c = d.alpha + ((size_t)b * g.Tw + c_t) * L * HWd;
    WbLay<NA> ly;
    wb_layers_fwd<NA>(d, px, ix, f_lo, alpha_c, c.s_occ, ly);
    // stage C: warp the context frame by the reduced flow
    WbTaps t = wb_taps(__fadd_rn(px.gx, ly.flow_x), __fadd_rn(px.gy, ly.flow_y), g.Wd, g.Hd);
    int m = wb_tap_mask(t, g.Wd, g.Hd);
    const float* src = d.input + ((size_t)b * g.T + c_t) * C * HWd;
    const float wgt = ly.score + 1e-6f;
    float* raw = d.raw_output + (((size_t)b * c.TcR + tc) * g.Tp + tp) * c.CR * HWd + q;
    WB_UNROLL for (int ch = 0; ch < WB_MAX_C; ++ch) {
      if (ch < C) {
        float v = wb_sample(src + (size_t)ch * HWd, t, m, g.Wd);
        if (active) raw[(size_t)ch * HWd] = v;
        acc[ch] += wgt * v;
      }
    }
    acc[WB_MAX_C] += wgt * (ly.score * 2.f - 1.f);
    den += wgt;
    if (active) {
      WB_UNROLL for (int k = 0; k < WB_MAX_L; ++k) if (k < L && !((wm >> k) & 1u)) raw[(size_t)(C + k) * HWd] = -1.f;
      WB_UNROLL_NA for (int s = 0; s < NA; ++s) if (s < ix.n) raw[(size_t)(C + ix.k[s]) * HWd] = ly.A[s] * 2.f - 1.f;
      if (c.disocc_ch) raw[(size_t)(C + L) * HWd] = ly.disocc;
      float* fl = d.flow + pair * 2 * HWd + q;
      fl[0] = ly.flow_x; fl[HWd] = ly.flow_y;
    }
  }
  if (c.self) {   // lvd.py:842-845: the target frame itself, fully opaque, score 1
    float* raw = d.raw_output + (((size_t)b * c.TcR + g.Tc) * g.Tp + tp) * c.CR * HWd + q;
    const float* src = d.input + ((size_t)b * g.T + tp) * C * HWd + q;
    const float wgt = 1.f + 1e-6f;
    WB_UNROLL for (int ch = 0; ch < WB_MAX_C; ++ch) {
      if (ch < C) { float v = __ldg(src + (size_t)ch * HWd); if (active) raw[(size_t)ch * HWd] = v; acc[ch] += wgt * v; }
    }
    if (active) {
      for (int k = 0; k < L; ++k) raw[(size_t)(C + k) * HWd] = 1.f;
      if (c.disocc_ch) raw[(size_t)(C + L) * HWd] = 1.f;
    }
    acc[WB_MAX_C] += wgt * 1.f;
    den += wgt;
  }
  if (active) {
    const float inv = 1.f / fmaxf(den, 1e-12f);
    float* of = d.out_full + ((size_t)b * g.Tp + tp) * (C + 1) * HWd + q;
    WB_UNROLL for (int ch = 0; ch < WB_MAX_C; ++ch) if (ch < C) of[(size_t)ch * HWd] = acc[ch] * inv;
    of[(size_t)C * HWd] = acc[WB_MAX_C] * inv;
    if (d.norm) d.norm[((size_t)b * g.Tp + tp) * HWd + q] = den;
  }
}

// grid = (CTAs, B*Tp); one thread per HD pixel (32x8 tiles), contexts looped inside so that the fused `output`
// (lvd.py:850-851) never leaves registers.
__global__ void __launch_bounds__(WB_TILE_PX, 2) k_warp_composite_fwd(WbDec d) {
  const waldo_geom_t g = d.g;
  WbFwdCtx c;
  c.L = g.No + 1; c.HW = g.H * g.W; c.C = g.C; c.HWd = (size_t)g.Hd * g.Wd;
  const int btp = blockIdx.y;
  c.b = btp / g.Tp; c.tp = btp - c.b * g.Tp;
  const int u = (int)d.pred_ts[c.tp];
  c.self = (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
  c.disocc_ch = (g.flags & WALDO_F_USE_DISOCC) != 0;
  c.TcR = g.Tc + (c.self ? 1 : 0); c.CR = c.C + c.L + (c.disocc_ch ? 1 : 0);
  __shared__ float s_occ[WB_MAX_L * WB_MAX_L];
  for (int i = wb_tid(); i < c.L * c.L; i += wb_nthr()) s_occ[i] = __ldg(d.occ + ((size_t)c.b * g.T + u) * c.L * c.L + i);
  __syncthreads();
  c.s_occ = s_occ;
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    for (int it = wb_tid(); it < WB_TILE_PX; it += wb_nthr()) {
      const int X = tx0 + (it & (WB_TILE_W - 1)), Y = ty0 + it / WB_TILE_W;
      const bool active = X < g.Wd && Y < g.Hd;
      const size_t q = active ? (size_t)Y * g.Wd + X : 0;
      WbPix px = wb_pix(d, c.b, c.tp, active ? X : 0, active ? Y : 0);
      const unsigned wm = wb_warp_or(active ? px.isobj : 1u);
      const int n = __popc(wm);
      if (n <= 4) wb_fwd_pixel<4>(d, c, px, wm, active, q);
      else if (n <= 8) wb_fwd_pixel<8>(d, c, px, wm, active, q);
      else wb_fwd_pixel<WB_MAX_L>(d, c, px, wm, active, q);
    }
  }
}
