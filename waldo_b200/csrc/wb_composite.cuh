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
#define WB_ROW_AL 20   // floats of the alpha part of a raw_output record staged per pixel: CRp - ceil4(C) <= 17 + 1 + pad
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
  int CRp, npass;     // floats per raw_output record; alpha channels handed to the gather kernel through `apass`
  unsigned HWd;
  bool self, disocc_ch;
  const float* s_occ;
  float* s_row;       // this warp's staging row of the alpha part of 32 records (lanes form)
};

// alpha channel k of pixel q of one (b,tc,tp) pair: the first `npass` layers travel through `apass` (the gather kernel
// writes them together with the last image channels as one 16-byte chunk), the others go straight into the record
WB_DEV void wb_store_alpha(const WbFwdCtx& c, float* __restrict__ raw, float* __restrict__ apass, int k, unsigned q, float v) {
  if (k < c.npass) apass[(size_t)k * c.HWd + q] = v;
  else raw[(size_t)q * c.CRp + c.C + k] = v;
}

// Layer part of one (pixel, context): evaluates the live layers, writes the alpha channels of raw_output (+ disocc),
// the reduced flow, and returns (flow, score) for the channel part.
template <int NA>
WB_DEV void wb_fwd_layers(const WbDec& d, const WbFwdCtx& c, const WbPix& px, unsigned wm, const WbIdx<NA>& ix, unsigned q, int c_t, size_t pair,
                          float* __restrict__ raw, float& flow_x, float& flow_y, float& score) {
  const waldo_geom_t& g = d.g;
  const int L = c.L, C = c.C;
  const unsigned HWd = c.HWd;
  const float* f_lo = d.f_lo + pair * L * c.HW * 2;
  const float* alpha_c = d.alpha + ((size_t)c.b * g.Tw + c_t) * L * HWd;
  WbLay<NA> ly;
  wb_layers_fwd<NA>(d, px, ix, f_lo, alpha_c, c.s_occ, ly);
  float* ap = d.apass + pair * c.npass * HWd;
  WB_UNROLL for (int k = 0; k < WB_MAX_L; ++k) if (k < L && !((wm >> k) & 1u)) wb_store_alpha(c, raw, ap, k, q, -1.f);
  WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s) if (s < ix.n) wb_store_alpha(c, raw, ap, ix.k[s], q, ly.A[s] * 2.f - 1.f);
  float* rec = raw + (size_t)q * c.CRp;
  if (c.disocc_ch) rec[C + L] = ly.disocc;
  for (int ch = c.CR; ch < c.CRp; ++ch) rec[ch] = 0.f;   // record padding
  float* fl = d.flow + pair * 2 * HWd + q;
  fl[0] = ly.flow_x; fl[HWd] = ly.flow_y;
  flow_x = ly.flow_x; flow_y = ly.flow_y; score = ly.score;
}

// all contexts of one pixel: the slot list is built once
template <int NA>
WB_DEV void wb_fwd_layers_ctxs(const WbDec& d, const WbFwdCtx& c, const WbPix& px, unsigned wm, unsigned q) {
  const waldo_geom_t& g = d.g;
  const int b = c.b, tp = c.tp;
  const unsigned HWd = c.HWd;
  const WbIdx<NA> ix = wb_idx<NA>(wm);
  for (int tc = 0; tc < g.Tc; ++tc) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
    float* raw = d.raw_output + (((size_t)b * c.TcR + tc) * g.Tp + tp) * (size_t)HWd * c.CRp;
    float flow_x, flow_y, score;
    wb_fwd_layers<NA>(d, c, px, wm, ix, q, c_t, pair, raw, flow_x, flow_y, score);
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
template <int LP>
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
    float* raw = d.raw_output + (((size_t)b * c.TcR + tc) * g.Tp + tp) * (size_t)HWd * c.CRp;   // records of this pair
    float* ap = d.apass + pair * c.npass * HWd;
    float* flo = d.flow + pair * 2 * HWd;
    float* sco = d.score + pair * HWd;
    // The alpha part of the 32 records of this row (floats Csplit .. CRp of each) is assembled in shared memory and leaves
    // as 128-bit stores.  Layers outside the row's union are fully transparent (-1), record padding is 0.
    const int Csplit = C + c.npass, nal = c.CRp - Csplit;   // nal: a multiple of 4, <= WB_ROW_AL
    {
      float* mine = c.s_row + lane * WB_ROW_AL;
      for (int e = 0; e < nal; ++e) mine[e] = (Csplit + e < C + L) ? -1.f : 0.f;
      WB_UNROLL for (int kk = 0; kk < 3; ++kk) if (kk < c.npass && !((wm >> kk) & 1u)) ap[(size_t)kk * HWd + ql] = -1.f;
    }
    __syncwarp();
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
      if (valid) {
        if (k < c.npass) ap[(size_t)k * HWd + q] = A * 2.f - 1.f;
        else c.s_row[p * WB_ROW_AL + (k - c.npass)] = A * 2.f - 1.f;
      }
      if (slot == 0) {
        flo[q] = fx; flo[HWd + q] = fy; sco[q] = sc;
        if (c.disocc_ch) c.s_row[p * WB_ROW_AL + (L - c.npass)] = mx;
      }
    }
    __syncwarp();
    for (int idx = lane; idx < 32 * (nal / 4); idx += 32) {   // (pixel, chunk): consecutive lanes -> consecutive 16-byte chunks
      const int p = idx / (nal / 4), ch = idx - p * (nal / 4);
      const int X = min(tx0 + p, g.Wd - 1);
      const float4 v = *reinterpret_cast<const float4*>(c.s_row + p * WB_ROW_AL + 4 * ch);
      wb_st4(raw + (size_t)(Y * g.Wd + X) * c.CRp + Csplit + 4 * ch, v);
    }
    __syncwarp();
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
  c.CRp = g.CRp; c.npass = ((c.C + 3) & ~3) - c.C;
  __shared__ float s_occ[WB_MAX_L * WB_MAX_L];
  __shared__ __align__(16) float s_rows[WB_TILE_PX / 32][32 * WB_ROW_AL];
  for (int i = wb_tid(); i < c.L * c.L; i += wb_nthr()) s_occ[i] = __ldg(d.occ + ((size_t)c.b * g.T + u) * c.L * c.L + i);
  __syncthreads();
  c.s_occ = s_occ;
  c.s_row = s_rows[wb_warp()];
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
      if (n == 1) wb_lanes_layers_fwd<1>(d, c, wm, n, px.isobj, tx0, Y, px.ay, px.gy);
      else if (n == 2) wb_lanes_layers_fwd<2>(d, c, wm, n, px.isobj, tx0, Y, px.ay, px.gy);
      else if (n <= 4) wb_lanes_layers_fwd<4>(d, c, wm, n, px.isobj, tx0, Y, px.ay, px.gy);
      else if (n <= 8) wb_lanes_layers_fwd<8>(d, c, wm, n, px.isobj, tx0, Y, px.ay, px.gy);
      else wb_fwd_layers_ctxs<WB_MAX_L>(d, c, px, wm, q);
#else
      if (WB_NA_VARIANTS_FWD >= 2 && n <= 4) wb_fwd_layers_ctxs<4>(d, c, px, wm, q);
      else if (WB_NA_VARIANTS_FWD >= 3 && n <= 8) wb_fwd_layers_ctxs<8>(d, c, px, wm, q);
      else wb_fwd_layers_ctxs<WB_MAX_L>(d, c, px, wm, q);
#endif
      if (c.self) {   // lvd.py:842-845: the target frame itself is fully opaque (the first npass layers: gather kernel)
        float* rec = d.raw_output + ((((size_t)b * c.TcR + g.Tc) * g.Tp + tp) * (size_t)HWd + q) * c.CRp;
        for (int k = c.npass; k < c.L; ++k) rec[c.C + k] = 1.f;
        if (c.disocc_ch) rec[c.C + c.L] = 1.f;
        for (int ch = c.CR; ch < c.CRp; ++ch) rec[ch] = 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_gather_fwd: stage C (lvd.py:830-853) on channels-last records.
// grid = (CTAs, B*Tp), 32x8 pixel tiles.  A PIXEL GROUP of WB_GRP = 8 lanes owns one pixel; lane j owns the 16-byte chunk j
// (channels 4j .. 4j+3) of every record of that pixel, so each of the four bilinear taps of a context frame is ONE 128-bit
// load per lane (LDG.E.128), warp-wide a run of whole consecutive records, and the warped channels leave as one 128-bit store
// per lane into the raw_output record.  The C image channels fill the first Csplit = ceil4(C) floats of that record; the
// Csplit - C leftover floats of the last image chunk are the first alpha channels (layer 0 ..), which the layer kernel hands
// over through `apass` so that this kernel writes whole chunks.  The fused `output` (lvd.py:850-851) accumulates in registers
// over the contexts.  Bit-identical arithmetic per channel to the reference association (wb_chain).
__global__ void __launch_bounds__(WB_TILE_PX, WB_OCC_GATHER_FWD) k_gather_fwd(WbDec d) {
  const waldo_geom_t g = d.g;
  const int C = g.C, L = g.No + 1, Cp = g.Cp, CRp = g.CRp;
  const unsigned HWd = (unsigned)(g.Hd * g.Wd);
  const int btp = blockIdx.y, b = btp / g.Tp, tp = btp - b * g.Tp;
  const bool self = (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
  const int TcR = g.Tc + (self ? 1 : 0);
  const int Csplit = (C + 3) & ~3, npass = Csplit - C;
  const int nchi = Csplit / 4;          // chunks of a raw_output record written here
  const int ncho = Cp / 4;              // chunks of an out_full record
  __shared__ const float* s_src[8];     // context frame of every context (CTA-uniform)
  __shared__ float* s_raw[8];           // raw_output block of every context
  for (int tc = wb_tid(); tc < g.Tc; tc += wb_nthr()) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    s_src[tc] = d.input + ((size_t)b * g.T + c_t) * HWd * Cp;
    s_raw[tc] = d.raw_output + (((size_t)b * TcR + tc) * g.Tp + tp) * (size_t)HWd * CRp;
  }
  __syncthreads();
  const int j0 = wb_tid() % WB_GRP, ppi = max(wb_nthr() / WB_GRP, 1);
  float* of = d.out_full + ((size_t)b * g.Tp + tp) * (size_t)HWd * Cp;
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    for (int pi = wb_tid() / WB_GRP; pi < WB_TILE_PX; pi += ppi) {
      // pixels beyond the image edge recompute (and re-store, identically) the nearest valid pixel: no predicates
      const int X = min(tx0 + (pi & (WB_TILE_W - 1)), g.Wd - 1), Y = min(ty0 + pi / WB_TILE_W, g.Hd - 1);
      const unsigned q = (unsigned)(Y * g.Wd + X);
      const float gx = __ldg(d.xs_hd + X), gy = __ldg(d.ys_hd + Y);
      float4 acc[WB_GCH];
      WB_CHUNKS acc[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
      float den = 0.f, accs = 0.f;
      for (int tc = 0; tc < g.Tc; ++tc) {
        const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
        const float* fl = d.flow + pair * 2 * HWd + q;
        const float score = __ldg(d.score + pair * HWd + q);
        const WbTaps t = wb_taps(__fadd_rn(gx, __ldg(fl)), __fadd_rn(gy, __ldg(fl + HWd)), g.Wd, g.Hd);
        const WbTap2 t2 = wb_tap2(t, g.Wd, g.Hd);
        const float wgt = score + 1e-6f;
        den += wgt;
        accs += wgt * (score * 2.f - 1.f);
        const float* r0 = s_src[tc] + (size_t)t2.o0 * Cp;
        const float* r1 = s_src[tc] + (size_t)t2.o1 * Cp;
        float* rw = s_raw[tc] + (size_t)q * CRp;
        WB_CHUNKS {
          const int j = j0 + ci;
          if (j < nchi) {
            const float4 v0 = wb_ld4(r0 + 4 * j), v1 = wb_ld4(r0 + Cp + 4 * j), v2 = wb_ld4(r1 + 4 * j), v3 = wb_ld4(r1 + Cp + 4 * j);
            float4 r;
            r.x = __fmaf_rn(v3.x, t2.w[3], __fmaf_rn(v2.x, t2.w[2], __fmaf_rn(v1.x, t2.w[1], __fmul_rn(v0.x, t2.w[0]))));
            r.y = __fmaf_rn(v3.y, t2.w[3], __fmaf_rn(v2.y, t2.w[2], __fmaf_rn(v1.y, t2.w[1], __fmul_rn(v0.y, t2.w[0]))));
            r.z = __fmaf_rn(v3.z, t2.w[3], __fmaf_rn(v2.z, t2.w[2], __fmaf_rn(v1.z, t2.w[1], __fmul_rn(v0.z, t2.w[0]))));
            r.w = __fmaf_rn(v3.w, t2.w[3], __fmaf_rn(v2.w, t2.w[2], __fmaf_rn(v1.w, t2.w[1], __fmul_rn(v0.w, t2.w[0]))));
            acc[ci].x += wgt * r.x; acc[ci].y += wgt * r.y; acc[ci].z += wgt * r.z; acc[ci].w += wgt * r.w;
            if (4 * j + 3 >= C) {   // the chunk that holds the last image channels: its tail = the first alpha channels
              const float* ap = d.apass + pair * npass * HWd + q;
              WB_UNROLL for (int e = 0; e < 4; ++e)
                if (4 * j + e >= C) wb_set(r, e, __ldg(ap + (size_t)(4 * j + e - C) * HWd));
            }
            wb_st4(rw + 4 * j, r);
          }
        }
      }
      if (self) {   // lvd.py:842-845: the target frame itself, score 1, every layer fully opaque
        const float wself = 1.f + 1e-6f;
        den += wself; accs += wself;
        const float* sr = d.input + (((size_t)b * g.T + tp) * HWd + q) * Cp;
        float* rw = d.raw_output + ((((size_t)b * TcR + g.Tc) * g.Tp + tp) * (size_t)HWd + q) * CRp;
        WB_CHUNKS {
          const int j = j0 + ci;
          if (j < nchi) {
            float4 v = wb_ld4(sr + 4 * j);
            acc[ci].x += wself * v.x; acc[ci].y += wself * v.y; acc[ci].z += wself * v.z; acc[ci].w += wself * v.w;
            WB_UNROLL for (int e = 0; e < 4; ++e) if (4 * j + e >= C) wb_set(v, e, 1.f);
            wb_st4(rw + 4 * j, v);
          }
        }
      }
      const float inv = 1.f / fmaxf(den, 1e-12f);
      WB_CHUNKS {
        const int j = j0 + ci;
        if (j < ncho) {
          float4 o = j < nchi ? make_float4(acc[ci].x * inv, acc[ci].y * inv, acc[ci].z * inv, acc[ci].w * inv) : make_float4(0.f, 0.f, 0.f, 0.f);
          WB_UNROLL for (int e = 0; e < 4; ++e) {
            const int ch = 4 * j + e;
            if (ch == C) wb_set(o, e, accs * inv);       // the fused score channel = raw_alpha (lvd.py:147)
            else if (ch > C) wb_set(o, e, 0.f);
          }
          wb_st4(of + (size_t)q * Cp + 4 * j, o);
        }
        if (j == 0) d.norm[((size_t)b * g.Tp + tp) * HWd + q] = den;
      }
    }
  }
}
