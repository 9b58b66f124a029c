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

// alpha_c = stored context alpha (2A-1) of frame (b,c): (L, Hd, Wd).  All per-thread addressing is 32-bit element
// offsets against warp-uniform 64-bit bases (a frame never exceeds 2^31 elements).
template <int NA>
WB_DEV void wb_layers_fwd(const WbDec& d, const WbPix& px, const WbIdx<NA>& ix, const float* __restrict__ f_lo /* (L,H,W,2) of this pair */,
                          const float* __restrict__ alpha_c, const float* __restrict__ s_occ, WbLay<NA>& ly) {
  const waldo_geom_t& g = d.g;
  const int L = g.No + 1;
  const unsigned HW = (unsigned)(g.H * g.W), HWd = (unsigned)(g.Hd * g.Wd);
  float mx = 0.f;   // layers outside the union have R = 0, and R >= 0 always
  WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s) {
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
        const WbTaps t = wb_taps(__fadd_rn(px.gx, fx), __fadd_rn(px.gy, fy), g.Wd, g.Hd);
        const WbTap2 t2 = wb_tap2(t, g.Wd, g.Hd);
        const float* pl = alpha_c + (size_t)k * HWd;
        r = wb_gather2_01(pl + t2.o0, pl + t2.o1, t2.w);
      }
      ly.R[s] = r;
      mx = fmaxf(mx, r);
    }
  }
  ly.disocc = mx;
  float fx = 0.f, fy = 0.f, sc = 0.f;
  WB_UNROLL_NA for (int i = 0; i < WB_NEND; ++i) {
    if (i < ix.n) {
      const float* oc = s_occ + ix.k[i];
      float vis = 1.f;
      WB_UNROLL_NA for (int j = 0; j < WB_NEND; ++j) if (j < ix.n) vis *= 1.f - ly.R[j] * oc[ix.k[j] * L];
      float a = vis * ly.R[i];
      ly.A[i] = a;
      fx += a * ly.Fx[i]; fy += a * ly.Fy[i]; sc += a;
    }
  }
  ly.flow_x = fx; ly.flow_y = fy; ly.score = sc;
}

struct WbFwdCtx {   // per-CTA constants of the fused forward
  int b, tp, L, C, TcR, CR, HW;
  unsigned HWd;
  bool self, disocc_ch;
  const float* s_occ;
};

// Layer part of one (pixel, context): evaluates the live layers, writes the alpha channels of raw_output (+ disocc),
// the reduced flow, and returns (flow, score) for the channel part.
template <int NA, typename ST>
WB_DEV void wb_fwd_layers(const WbDec& d, const WbFwdCtx& c, const WbPix& px, unsigned wm, const WbIdx<NA>& ix, unsigned q, int c_t, size_t pair,
                          ST* __restrict__ raw, float& flow_x, float& flow_y, float& score) {
  const waldo_geom_t& g = d.g;
  const int L = c.L, C = c.C;
  const unsigned HWd = c.HWd;
  const float* f_lo = d.f_lo + pair * L * c.HW * 2;
  const float* alpha_c = d.alpha + ((size_t)c.b * g.Tw + c_t) * L * HWd;   // (fp32 in every storage variant)
  WbLay<NA> ly;
  wb_layers_fwd<NA>(d, px, ix, f_lo, alpha_c, c.s_occ, ly);
  ST* ra = raw + (size_t)C * HWd + q;
  WB_UNROLL for (int k = 0; k < WB_MAX_L; ++k) { if (k < L && !((wm >> k) & 1u)) wb_sts(ra, -1.f); ra += HWd; }
  ra = raw + (size_t)C * HWd + q;
  WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s) if (s < ix.n) wb_sts(ra + (size_t)ix.k[s] * HWd, ly.A[s] * 2.f - 1.f);
  if (c.disocc_ch) wb_sts(ra + (size_t)L * HWd, ly.disocc);
  float* fl = d.flow + pair * 2 * HWd + q;
  fl[0] = ly.flow_x; fl[HWd] = ly.flow_y;
  flow_x = ly.flow_x; flow_y = ly.flow_y; score = ly.score;
}

// all contexts of one pixel: the slot list is built once
template <int NA, typename ST>
WB_DEV void wb_fwd_layers_ctxs(const WbDec& d, const WbFwdCtx& c, const WbPix& px, unsigned wm, unsigned q) {
  const waldo_geom_t& g = d.g;
  const int b = c.b, tp = c.tp;
  const unsigned HWd = c.HWd;
  const WbIdx<NA> ix = wb_idx<NA>(wm);
  for (int tc = 0; tc < g.Tc; ++tc) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
    ST* raw = reinterpret_cast<ST*>(d.raw_output) + (((size_t)b * c.TcR + tc) * g.Tp + tp) * c.CR * HWd;
    float flow_x, flow_y, score;
    wb_fwd_layers<NA, ST>(d, c, px, wm, ix, q, c_t, pair, raw, flow_x, flow_y, score);
    d.score[pair * HWd + q] = score;
  }
}

#ifndef WB_HOST_EMU
// ------------------------------------------------------------------------------------------------------------------
// Lanes-per-layer form of the layer forward (rows whose union of live layers has n <= 8 members): LP = 1|2|4|8 >= n lanes
// per pixel, lane = slot * (32/LP) + pixel, LP passes over the 32 pixels of the row.  Every lane owns ONE (pixel, layer):
// the layers of a pixel load in parallel, the occlusion product and the reductions over layers run on warp shuffles,
// and there is no per-thread array indexed by a layer.  Results are those of wb_layers_fwd up to the order of the sums
// over layers in flow / score (each A_k, R_k is bit-identical).
template <int LP, typename ST>
WB_DEV void wb_lanes_layers_fwd(const WbDec& d, const WbFwdCtx& c, unsigned wm, int n, unsigned isobj_lane, int tx0, int Y,
                                const WbAxis& ay, float gy) {
  constexpr int PPW = 32 / LP;
  const waldo_geom_t& g = d.g;
  const int L = c.L, C = c.C, HW = c.HW, b = c.b, tp = c.tp;
  const unsigned HWd = c.HWd;
  const int lane = wb_lane(), pl = lane % PPW, slot = lane / PPW;
  const bool valid = slot < n;
  const int k = valid ? wb_nth_bit(wm, slot) : 0;
  float oc[LP];
  WB_UNROLL for (int j = 0; j < LP; ++j) oc[j] = (valid && j < n) ? c.s_occ[wb_nth_bit(wm, j) * L + k] : 0.f;
  const float r_lo = (float)g.H / (float)g.Hd;
  const bool direct = g.Hd == g.H;
  const int Xl = min(tx0 + lane, g.Wd - 1);
  const unsigned ql = (unsigned)(Y * g.Wd + Xl);   // this lane's own pixel (lane = pixel layout, dead-layer stores)
  for (int tc = 0; tc < g.Tc; ++tc) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
    const float2* fl = reinterpret_cast<const float2*>(d.f_lo) + (pair * L + k) * HW;
    const float* alpha_k = d.alpha + (((size_t)b * g.Tw + c_t) * L + k) * HWd;
    ST* ra = reinterpret_cast<ST*>(d.raw_output) + ((((size_t)b * c.TcR + tc) * g.Tp + tp) * c.CR + C) * HWd;   // alpha channels of this pair
    float* flo = d.flow + pair * 2 * HWd;
    float* sco = d.score + pair * HWd;
    // layers outside the row's union are fully transparent
    { ST* o = ra + ql; WB_UNROLL for (int kk = 0; kk < WB_MAX_L; ++kk) { if (kk < L && !((wm >> kk) & 1u)) wb_sts(o, -1.f); o += HWd; } }
#pragma unroll 1
    for (int r = 0; r < LP; ++r) {
      const int p = r * PPW + pl, X = min(tx0 + p, g.Wd - 1);
      const unsigned q = (unsigned)(Y * g.Wd + X);
      const WbAxis ax = wb_axis(X, r_lo, g.W);
      const float gx = __ldg(d.xs_hd + X);
      const int o00 = ay.i0 * g.W + ax.i0, o01 = ay.i0 * g.W + ax.i1, o10 = ay.i1 * g.W + ax.i0, o11 = ay.i1 * g.W + ax.i1;
      const unsigned isobj = __shfl_sync(0xffffffffu, isobj_lane, p);
      float2 f00 = __ldg(fl + o00), f01 = f00, f10 = f00, f11 = f00;
      if (!direct) { f01 = __ldg(fl + o01); f10 = __ldg(fl + o10); f11 = __ldg(fl + o11); }
#if WB_PF_FLO
      if (tc + 1 < g.Tc) {   // the same cells of the next context: pair + Tp
        const float2* fn = fl + (size_t)g.Tp * L * HW;
        wb_prefetch_l1(fn + o00);
        if (!direct) wb_prefetch_l1(fn + o10);
      }
#endif
      float Fx = 0.f, Fy = 0.f, rr = 0.f;
      if (valid) {
        if (direct) { Fx = f00.x; Fy = f00.y; }
        else {
          Fx = wb_lerp2(f00.x, f01.x, f10.x, f11.x, ax, ay);
          Fy = wb_lerp2(f00.y, f01.y, f10.y, f11.y, ax, ay);
        }
        if ((isobj >> k) & 1u) {
          const WbTaps t = wb_taps(__fadd_rn(gx, Fx), __fadd_rn(gy, Fy), g.Wd, g.Hd);
          const WbTap2 t2 = wb_tap2(t, g.Wd, g.Hd);
          rr = wb_gather2_01(alpha_k + t2.o0, alpha_k + t2.o1, t2.w);
        }
      }
      float run = 1.f;
      WB_UNROLL for (int j = 0; j < LP; ++j) run *= 1.f - __shfl_sync(0xffffffffu, rr, pl + j * PPW) * oc[j];
      const float A = run * rr;
      float fx = A * Fx, fy = A * Fy, sc = A, mx = rr;
      WB_UNROLL for (int o = PPW; o < 32; o <<= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o);
        sc += __shfl_xor_sync(0xffffffffu, sc, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      if (valid) wb_sts(ra + ((size_t)k * HWd + q), A * 2.f - 1.f);
      if (slot == 0) {
        flo[q] = fx; flo[HWd + q] = fy; sco[q] = sc;
        if (c.disocc_ch) wb_sts(ra + ((size_t)L * HWd + q), mx);
      }
    }
  }
}
#endif  // !WB_HOST_EMU

// ------------------------------------------------------------------------------------------------------------------
// The forward runs as two kernels per (b, tp) so that each gets the register budget it needs:
//   k_layers_fwd : the irregular layer part (B5up..B9) -> alpha channels of raw_output, flow, score
//   k_gather_fwd : the streaming part (stage C)        -> image channels of raw_output, fused output, norm
// Only 3 floats per (pixel, context) pass between them (flow is an output of the path anyway, score is kept for the
// backward); no per-layer HD tensor ever reaches HBM.
// ------------------------------------------------------------------------------------------------------------------

// grid = (CTAs, B*Tp), 32x8 pixel tiles, one thread per HD pixel, rolled loop over the contexts.
template <typename ST>
__global__ void __launch_bounds__(WB_TILE_PX, WB_OCC_LAYERS_FWD) k_layers_fwd(WbDec d) {
  const waldo_geom_t g = d.g;
  WbFwdCtx c;
  c.L = g.No + 1; c.HW = g.H * g.W; c.C = g.C; c.HWd = (unsigned)(g.Hd * g.Wd);
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
  const int b = c.b, tp = c.tp;
  const unsigned HWd = c.HWd;
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    for (int it = wb_tid(); it < WB_TILE_PX; it += wb_nthr()) {
      // threads beyond the image edge recompute (and re-store, identically) the nearest valid pixel: no predicates
      const int X = min(tx0 + (it & (WB_TILE_W - 1)), g.Wd - 1), Y = min(ty0 + it / WB_TILE_W, g.Hd - 1);
      const unsigned q = (unsigned)(Y * g.Wd + X);
      WbPix px = wb_pix(d, b, tp, X, Y);
      const unsigned wm = wb_warp_or(px.isobj);
      const int n = __popc(wm);
#if !defined(WB_HOST_EMU) && !defined(WB_NO_LANES) && !defined(WB_NO_LANES_FWD)
      if (n == 1) wb_lanes_layers_fwd<1, ST>(d, c, wm, n, px.isobj, tx0, Y, px.ay, px.gy);
      else if (n == 2) wb_lanes_layers_fwd<2, ST>(d, c, wm, n, px.isobj, tx0, Y, px.ay, px.gy);
      else if (n <= 4) wb_lanes_layers_fwd<4, ST>(d, c, wm, n, px.isobj, tx0, Y, px.ay, px.gy);
      else if (n <= 8) wb_lanes_layers_fwd<8, ST>(d, c, wm, n, px.isobj, tx0, Y, px.ay, px.gy);
      else wb_fwd_layers_ctxs<WB_MAX_L, ST>(d, c, px, wm, q);
#else
      if (WB_NA_VARIANTS_FWD >= 2 && n <= 4) wb_fwd_layers_ctxs<4, ST>(d, c, px, wm, q);
      else if (WB_NA_VARIANTS_FWD >= 3 && n <= 8) wb_fwd_layers_ctxs<8, ST>(d, c, px, wm, q);
      else wb_fwd_layers_ctxs<WB_MAX_L, ST>(d, c, px, wm, q);
#endif
      if (c.self) {   // lvd.py:842-845: the target frame itself is fully opaque
        ST* raw = reinterpret_cast<ST*>(d.raw_output) + (((size_t)b * c.TcR + g.Tc) * g.Tp + tp) * c.CR * HWd + q;
        for (int k = 0; k < c.L; ++k) wb_sts(raw + (size_t)(c.C + k) * HWd, 1.f);
        if (c.disocc_ch) wb_sts(raw + (size_t)(c.C + c.L) * HWd, 1.f);
      }
    }
  }
}

// grid = (CTAs, B*Tp), 32x8 pixel tiles, one thread per HD pixel.  The taps of the TCAP (>= Tc) contexts live in
// registers; ONE rolled loop walks the C image channels with the contexts unrolled inside: every channel of every
// context frame is gathered, stored to raw_output and fused into `output` (lvd.py:850-851) on the fly.
// FAST = exactly TCAP contexts and no include_self context, resolved at compile time (no predicates in the channel loop).
template <int TCAP, bool FAST, typename ST>
__global__ void __launch_bounds__(WB_TILE_PX, WB_OCC_GATHER_FWD) k_gather_fwd(WbDec d) {
  const waldo_geom_t g = d.g;
  const int C = g.C, L = g.No + 1;
  const unsigned HWd = (unsigned)(g.Hd * g.Wd);
  const int btp = blockIdx.y, b = btp / g.Tp, tp = btp - b * g.Tp;
  const bool self = !FAST && (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
  const int TcR = g.Tc + (self ? 1 : 0), CR = C + L + ((g.flags & WALDO_F_USE_DISOCC) ? 1 : 0);
  __shared__ const ST* s_src[TCAP];   // context frame of every context (CTA-uniform)
  __shared__ ST* s_raw[TCAP];         // raw_output block of every context
  for (int tc = wb_tid(); tc < g.Tc; tc += wb_nthr()) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    s_src[tc] = reinterpret_cast<const ST*>(d.input) + ((size_t)b * g.T + c_t) * C * HWd;
    s_raw[tc] = reinterpret_cast<ST*>(d.raw_output) + (((size_t)b * TcR + tc) * g.Tp + tp) * CR * HWd;
  }
  __syncthreads();
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    for (int it = wb_tid(); it < WB_TILE_PX; it += wb_nthr()) {
      const int X = min(tx0 + (it & (WB_TILE_W - 1)), g.Wd - 1), Y = min(ty0 + it / WB_TILE_W, g.Hd - 1);
      const unsigned q = (unsigned)(Y * g.Wd + X);
      const float gx = __ldg(d.xs_hd + X), gy = __ldg(d.ys_hd + Y);
      unsigned o0[TCAP], o1[TCAP];
      float w[TCAP][4], wgt[TCAP];
      float den = 0.f, accs = 0.f;
      WB_UNROLL for (int tc = 0; tc < TCAP; ++tc) {
        o0[tc] = 0u; o1[tc] = 0u; wgt[tc] = 0.f;
        WB_UNROLL for (int j = 0; j < 4; ++j) w[tc][j] = 0.f;
        if (FAST || tc < g.Tc) {
          const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
          const float* fl = d.flow + pair * 2 * HWd + q;
          const float score = __ldg(d.score + pair * HWd + q);
          const WbTaps t = wb_taps(__fadd_rn(gx, __ldg(fl)), __fadd_rn(gy, __ldg(fl + HWd)), g.Wd, g.Hd);
          const WbTap2 t2 = wb_tap2(t, g.Wd, g.Hd);
          o0[tc] = t2.o0; o1[tc] = t2.o1;
          WB_UNROLL for (int j = 0; j < 4; ++j) w[tc][j] = t2.w[j];
          wgt[tc] = score + 1e-6f;
          den += wgt[tc];
          accs += wgt[tc] * (score * 2.f - 1.f);
        }
      }
      const ST* self_src = nullptr;
      ST* self_raw = nullptr;
      float wself = 0.f;
      if (self) {   // lvd.py:842-845: the target frame itself, score 1
        self_raw = reinterpret_cast<ST*>(d.raw_output) + (((size_t)b * TcR + g.Tc) * g.Tp + tp) * CR * HWd + q;
        self_src = reinterpret_cast<const ST*>(d.input) + ((size_t)b * g.T + tp) * C * HWd + q;
        wself = 1.f + 1e-6f;
        den += wself; accs += wself;
      }
      const float inv = 1.f / fmaxf(den, 1e-12f);
      ST* of = reinterpret_cast<ST*>(d.out_full) + ((size_t)b * g.Tp + tp) * (C + 1) * HWd + q;
      unsigned choff = 0u;   // ch * HWd
#ifndef WB_HOST_EMU
#pragma unroll 2
#endif
      for (int ch = 0; ch < C; ++ch) {
        // all 4 x TCAP loads of this channel first (read-only path), then the arithmetic and the stores
        float v[TCAP][4];
        WB_UNROLL for (int tc = 0; tc < TCAP; ++tc) {
          if (FAST || tc < g.Tc) {
            const ST* pl = s_src[tc] + choff;
            const ST* p0 = pl + o0[tc];
            const ST* p1 = pl + o1[tc];
            v[tc][0] = wb_lds(p0); v[tc][1] = wb_lds(p0 + 1); v[tc][2] = wb_lds(p1); v[tc][3] = wb_lds(p1 + 1);
          }
        }
        float acc = 0.f;
        WB_UNROLL for (int tc = 0; tc < TCAP; ++tc) {
          if (FAST || tc < g.Tc) {
            const float r = __fmaf_rn(v[tc][3], w[tc][3], __fmaf_rn(v[tc][2], w[tc][2], __fmaf_rn(v[tc][1], w[tc][1], __fmul_rn(v[tc][0], w[tc][0]))));
            wb_sts(s_raw[tc] + (choff + q), r);
            acc += wgt[tc] * r;
          }
        }
        if (self) { const float v = wb_lds(self_src + choff); wb_sts(self_raw + choff, v); acc += wself * v; }
        wb_sts(of + choff, acc * inv);
        choff += HWd;
      }
      wb_sts(of + choff, accs * inv);
      d.norm[((size_t)b * g.Tp + tp) * HWd + q] = den;
    }
  }
}

