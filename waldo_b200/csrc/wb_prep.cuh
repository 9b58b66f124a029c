// Stage B low-res + context-alpha kernels (forward).
// Reference: models/nets/lvd.py:707-766 (grid_to_flow_ctx B1-B4), :617-653 (grid_to_flow), :771-794 (B5).
#pragma once
#include "wb_common.cuh"
#include "../../include/waldo_b200.h"

typedef waldo_decode_fwd_t WbDec;

WB_DEV int wb_L(const waldo_geom_t& g) { return g.No + 1; }

// ------------------------------------------------------------------ B1: project opacities (lvd.py:723-728)
// a_lo[b,t,k,p] = bil0((alpha_k+1)/2, src_grid_k[b,t,p]); k = 0 background, k >= 1 objects.  One thread per low-res
// cell walks the layers and also records which of them have any in-range tap there (live_ctx): everywhere else the
// layer's opacity AND every gradient path through it are exactly zero, which the HD kernels use to skip layers.
__global__ void k_project_alpha(WbDec d) {
  const waldo_geom_t g = d.g;
  const int L = wb_L(g), HW = g.H * g.W;
  const long long total = (long long)g.B * g.Tw * HW;
  for (long long e = (long long)blockIdx.x * wb_nthr() + wb_tid(); e < total; e += (long long)gridDim.x * wb_nthr()) {
    int p = (int)(e % HW);
    int t = (int)((e / HW) % g.Tw);
    int b = (int)(e / ((long long)HW * g.Tw));
    unsigned live = 0u;
    for (int k = 0; k < L; ++k) {
      const float* sg; const float* plane; int w, h;
      if (k == 0) {
        sg = d.src_grid_bg + (((size_t)b * g.T + t) * HW + p) * 2;
        plane = d.bg_alpha + (size_t)b * HW; w = g.W; h = g.H;
      } else {
        sg = d.src_grid_obj + ((((size_t)b * g.T + t) * g.No + (k - 1)) * HW + p) * 2;
        plane = d.obj_alpha + ((size_t)b * g.No + (k - 1)) * g.Ho * g.Wo; w = g.Wo; h = g.Ho;
      }
      WbTaps tp = wb_taps(__ldg(sg), __ldg(sg + 1), w, h);
      int m = wb_tap_mask(tp, w, h);
      if (m) live |= 1u << k;
      const float* q = plane + (long long)tp.y0 * w + tp.x0;
      float vnw = (m & 1) ? (__ldg(q) + 1.f) * 0.5f : 0.f;
      float vne = (m & 2) ? (__ldg(q + 1) + 1.f) * 0.5f : 0.f;
      float vsw = (m & 4) ? (__ldg(q + w) + 1.f) * 0.5f : 0.f;
      float vse = (m & 8) ? (__ldg(q + w + 1) + 1.f) * 0.5f : 0.f;
      d.a_lo[(((size_t)b * g.Tw + t) * L + k) * HW + p] = wb_chain(vnw, vne, vsw, vse, tp);
    }
    d.live_ctx[e] = live;
  }
}

// low-res layout logits of one (b,t,p): bilinear down-sample of the HD input (lvd.py:716 `scale`)
WB_DEV void wb_lyt_lo(const WbDec& d, int b, int t, int p, float* lyt /*[Nl]*/) {
  const waldo_geom_t& g = d.g;
  int y = p / g.W, x = p - y * g.W;
  const float r = (float)g.Hd / (float)g.H;   // = 1 / (1/scale_hd)
  WbAxis ay = wb_axis(y, r, g.Hd), ax = wb_axis(x, r, g.Wd);
  const size_t HWd = (size_t)g.Hd * g.Wd;
  const float* base = d.input + (((size_t)b * g.T + t) * g.C + 3) * HWd;
  for (int c = 0; c < g.Nl; ++c) {
    const float* pl = base + c * HWd;
    lyt[c] = wb_lerp2(__ldg(pl + (size_t)ay.i0 * g.Wd + ax.i0), __ldg(pl + (size_t)ay.i0 * g.Wd + ax.i1),
                      __ldg(pl + (size_t)ay.i1 * g.Wd + ax.i0), __ldg(pl + (size_t)ay.i1 * g.Wd + ax.i1), ax, ay);
  }
}

WB_DEV void wb_softmax(const float* x, float* y, int n) {
  float mx = x[0];
  for (int c = 1; c < n; ++c) mx = fmaxf(mx, x[c]);
  float s = 0.f;
  for (int c = 0; c < n; ++c) { y[c] = expf(x[c] - mx); s += y[c]; }
  float inv = 1.f / s;
  for (int c = 0; c < n; ++c) y[c] *= inv;
}

// ------------------------------------------------------------------ B2: class profile partial sums (lvd.py:735-742)
// Deterministic two-step reduction: every CTA stages WB_PROF_BATCH samples (logits + per-object weights) in
// shared memory, then thread <-> (object, class) accumulates them in sample order; CTA partials are reduced
// in CTA order by k_profile_final.  grid = (prof_ctas, B).
#define WB_PROF_BATCH 256
// The class count is a template parameter (20 / 19 / generic) and the object loop is unrolled over the compiled maximum:
// the per-class vectors stay in registers (run-time trip counts put them in local memory) and the rows of `cls` are read
// from shared memory at compile-time offsets.  Same operations in the same order as the generic loops.
template <int NLC, typename ST>
__global__ void __launch_bounds__(256) k_class_profile(WbDec d) {
  constexpr int NN = NLC > 0 ? NLC : WB_MAX_NL;
  constexpr int NO = WB_MAX_L - 1;
  const waldo_geom_t g = d.g;
  const int No = g.No, Nl = NLC > 0 ? NLC : g.Nl, HW = g.H * g.W, L = No + 1;
  const int b = blockIdx.y;
  const int nsamp = g.Tw * HW;
  const int nout = No * Nl + No;
  __shared__ float s_lyt[WB_PROF_BATCH][WB_MAX_NL + 1];
  __shared__ float s_w[WB_PROF_BATCH][WB_MAX_L];
  __shared__ __align__(16) float s_cls[(WB_MAX_L - 1) * WB_MAX_NL];
  const bool wcls = (g.flags & WALDO_F_WEIGHT_CLS) != 0;
  if (wcls)
    for (int i = wb_tid(); i < No * Nl; i += wb_nthr()) s_cls[i] = __ldg(d.cls + (size_t)b * No * Nl + i) + g.min_cls;
  // accumulators: outputs o = tid, tid + nthr, ... (at most 2 per thread with 256 threads; generic loop for emu)
  float* part = d.prof_part + ((size_t)b * gridDim.x + blockIdx.x) * nout;
  for (int o = wb_tid(); o < nout; o += wb_nthr()) part[o] = 0.f;
  __syncthreads();
  const float r = (float)g.Hd / (float)g.H;   // = 1 / (1/scale_hd)
  const size_t HWd = (size_t)g.Hd * g.Wd;
  for (int s0 = blockIdx.x * WB_PROF_BATCH; s0 < nsamp; s0 += gridDim.x * WB_PROF_BATCH) {
    const int ns = min(WB_PROF_BATCH, nsamp - s0);
    for (int i = wb_tid(); i < ns; i += wb_nthr()) {
      const int s = s0 + i, t = s / HW, p = s - t * HW;
      float al[NO];
      WB_UNROLL for (int k = 0; k < NO; ++k)
        al[k] = k < No ? __ldg(d.a_lo + (((size_t)b * g.Tw + t) * L + k + 1) * HW + p) + 1e-6f : 0.f;
      float lyt[NN], sm[NN];
      {   // same arithmetic as wb_lyt_lo
        const int y = p / g.W, x = p - y * g.W;
        const WbAxis ay = wb_axis(y, r, g.Hd), ax = wb_axis(x, r, g.Wd);
        const ST* base = reinterpret_cast<const ST*>(d.input) + (((size_t)b * g.T + t) * g.C + 3) * HWd;
        const size_t h00 = (size_t)ay.i0 * g.Wd + ax.i0, h01 = (size_t)ay.i0 * g.Wd + ax.i1;
        const size_t h10 = (size_t)ay.i1 * g.Wd + ax.i0, h11 = (size_t)ay.i1 * g.Wd + ax.i1;
        WB_UNROLL for (int c = 0; c < NN; ++c)
          if (NLC > 0 || c < Nl) {
            const ST* pl = base + c * HWd;
            lyt[c] = wb_lerp2(wb_lds(pl + h00), wb_lds(pl + h01), wb_lds(pl + h10), wb_lds(pl + h11), ax, ay);
          }
      }
      if (wcls) {   // same arithmetic as wb_softmax
        float mx = lyt[0];
        WB_UNROLL for (int c = 1; c < NN; ++c) if (NLC > 0 || c < Nl) mx = fmaxf(mx, lyt[c]);
        float ssum = 0.f;
        WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) { sm[c] = expf(lyt[c] - mx); ssum += sm[c]; }
        const float inv = 1.f / ssum;
        WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) sm[c] *= inv;
      }
      WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) s_lyt[i][c] = lyt[c];
      if (d.lyt_lo) {
        float* o = d.lyt_lo + ((size_t)b * g.Tw + t) * Nl * HW + p;
        WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) o[(size_t)c * HW] = lyt[c];
      }
      WB_UNROLL for (int k = 0; k < NO; ++k) {
        if (k < No) {
          float w = al[k];
          if (wcls) {
            const float* Ck = s_cls + k * Nl;
            float acc = 0.f;
            WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) acc += Ck[c] * sm[c];
            w *= acc;
          }
          s_w[i][k] = w;
        }
      }
    }
    __syncthreads();
    for (int o = wb_tid(); o < nout; o += wb_nthr()) {
      float acc = part[o];
      if (o < No * Nl) {
        int k = o / Nl, c = o - k * Nl;
        for (int i = 0; i < ns; ++i) acc += s_lyt[i][c] * s_w[i][k];
      } else {
        int k = o - No * Nl;
        for (int i = 0; i < ns; ++i) acc += s_w[i][k];
      }
      part[o] = acc;
    }
    __syncthreads();
  }
}

// reduce CTA partials in order, then P = softmax_c(num / den)  (lvd.py:742,744); or P = cls (lvd.py:753)
__global__ void k_profile_final(WbDec d) {
  const waldo_geom_t g = d.g;
  const int No = g.No, Nl = g.Nl, nout = No * Nl + No;
  const int b = blockIdx.x;
  const bool from_cls = (g.flags & WALDO_F_HAS_CLS) && !(g.flags & WALDO_F_WEIGHT_CLS);
  if (!from_cls) {
    for (int o = wb_tid(); o < nout; o += wb_nthr()) {
      float acc = 0.f;
      WB_UNROLL_N(8) for (int c = 0; c < d.prof_ctas; ++c) acc += d.prof_part[((size_t)b * d.prof_ctas + c) * nout + o];
      d.prof_sum[(size_t)b * nout + o] = acc;
    }
    __syncthreads();
  }
  for (int k = wb_tid(); k < No; k += wb_nthr()) {
    float* P = d.prof_p + ((size_t)b * No + k) * Nl;
    if (from_cls) {
      for (int c = 0; c < Nl; ++c) P[c] = __ldg(d.cls + ((size_t)b * No + k) * Nl + c);
    } else {
      float mean[WB_MAX_NL];
      float den = d.prof_sum[(size_t)b * nout + No * Nl + k];
      for (int c = 0; c < Nl; ++c) mean[c] = d.prof_sum[(size_t)b * nout + k * Nl + c] / den;
      wb_softmax(mean, P, Nl);
    }
  }
}

// ------------------------------------------------------------------ B2b-B4: context opacity at HD (lvd.py:744-765)
// One thread per HD pixel of one (b,t), 32x8 pixel tiles; grid = (CTAs, B*Tw).
//   l_k   = 1 - 0.5 * sum_c |P[b,k,c] - softmax(hd_lyt)[c]|
//   a_k   = up(a_lo[k]) * (k >= 1 ? l_k : 1)
//   A_i   = a_i * prod_j (1 - a_j * occ[b,t,j,i]);     alpha = 2A - 1
// Layers that are dead at all four low-res taps of a pixel (live_ctx) have a_k == 0 exactly: their factors are exactly
// 1 and their output exactly -1, so the loops run over the warp-wide union of live layers only (bit-identical result).
#define WB_TILE_W 32
#define WB_TILE_H 8
#define WB_TILE_PX (WB_TILE_W * WB_TILE_H)

struct WbTileIter {
  int tiles_x, ntiles;
  __device__ WbTileIter(int Hd, int Wd) { tiles_x = (Wd + WB_TILE_W - 1) / WB_TILE_W; ntiles = tiles_x * ((Hd + WB_TILE_H - 1) / WB_TILE_H); }
};

WB_DEV unsigned wb_live4(const uint32_t* __restrict__ live, int o00, int o01, int o10, int o11) {
  return __ldg(live + o00) | __ldg(live + o01) | __ldg(live + o10) | __ldg(live + o11);
}

// compact slot list of the layers in a (warp-uniform) mask, ascending
template <int NA> struct WbIdx { int k[NA]; int n; };
template <int NA> WB_DEV WbIdx<NA> wb_idx(unsigned wm) {
  WbIdx<NA> r;
  r.n = __popc(wm);
  unsigned m = wm;
  // (rolled form: only the live slots are ever read)
  WB_UNROLL_NA for (int s = 0; s < (NA <= 8 ? NA : r.n); ++s) { r.k[s] = m ? __ffs((int)m) - 1 : 0; m &= m - 1u; }
  return r;
}

template <typename ST> struct WbPrepCtx {
  int b, t, L, Nl, HW;
  size_t HWd;
  bool filt;
  const float *s_P, *s_occ, *alo;
  const ST* lyt_base;
  float* out;   // alpha stays fp32 in every storage variant: the flow is computed from it
};

// softmax of the HD layout logits of pixel q (lvd.py:744).  NLC = compile-time class count (0: run-time Nl).
template <int NLC, typename ST = float>
WB_DEV void wb_softmax_hd(const ST* __restrict__ lyt_base, unsigned HWd, unsigned q, int Nl, float* sm) {
  constexpr int NN = NLC > 0 ? NLC : WB_MAX_NL;
  float lyt[NN];
  const ST* p = lyt_base + q;
  WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) { lyt[c] = wb_lds(p); p += HWd; }
  float mx = lyt[0];
  WB_UNROLL for (int c = 1; c < NN; ++c) if (NLC > 0 || c < Nl) mx = fmaxf(mx, lyt[c]);
  float s = 0.f;
  WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) { sm[c] = expf(lyt[c] - mx); s += sm[c]; }
  const float inv = 1.f / s;
  WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) sm[c] *= inv;
}

template <int NA, int NLC, typename ST>
WB_DEV void wb_prep_pixel(const WbDec& d, const WbPrepCtx<ST>& c, unsigned wm, unsigned q, const WbAxis& ax, const WbAxis& ay,
                          int o00, int o01, int o10, int o11) {
  constexpr int NN = NLC > 0 ? NLC : WB_MAX_NL;
  const waldo_geom_t& g = d.g;
  const int L = c.L, Nl = NLC > 0 ? NLC : c.Nl;
  const unsigned HWd = (unsigned)c.HWd;
  const WbIdx<NA> ix = wb_idx<NA>(wm);
  float sm[NN];
  if (c.filt && (wm >> 1)) wb_softmax_hd<NLC, ST>(c.lyt_base, HWd, q, Nl, sm);
  float a[NA];
  WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s) {
    a[s] = 0.f;
    if (s < ix.n) {
      const int k = ix.k[s];
      const float* pl = c.alo + (size_t)k * c.HW;
      float v = (g.Hd == g.H) ? __ldg(pl + o00)
                              : wb_lerp2(__ldg(pl + o00), __ldg(pl + o01), __ldg(pl + o10), __ldg(pl + o11), ax, ay);
      if (c.filt && k >= 1) {
        const float* P = c.s_P + (k - 1) * Nl;
        float dist = 0.f;
        WB_UNROLL for (int cc = 0; cc < NN; ++cc) if (NLC > 0 || cc < Nl) dist += fabsf(P[cc] - sm[cc]);
        v *= 1.f - dist * 0.5f;
      }
      a[s] = v;
    }
  }
  float* o = c.out + q;
  WB_UNROLL for (int k = 0; k < WB_MAX_L; ++k) { if (k < L && !((wm >> k) & 1u)) wb_sts(o, -1.f); o += HWd; }
  o = c.out + q;
  WB_UNROLL_NA for (int i = 0; i < WB_NEND; ++i) {
    if (i < ix.n) {
      const float* oc = c.s_occ + ix.k[i];
      float vis = 1.f;
      WB_UNROLL_NA for (int j = 0; j < WB_NEND; ++j) if (j < ix.n) vis *= 1.f - a[j] * oc[ix.k[j] * L];
      wb_sts(o + (size_t)ix.k[i] * HWd, (vis * a[i]) * 2.f - 1.f);
    }
  }
}

template <int NLC, typename ST>
__global__ void __launch_bounds__(WB_TILE_PX, WB_OCC_PREP_FWD) k_alpha_prep(WbDec d) {
  const waldo_geom_t g = d.g;
  WbPrepCtx<ST> c;
  c.L = g.No + 1; c.Nl = g.Nl; c.HW = g.H * g.W; c.HWd = (size_t)g.Hd * g.Wd;
  const int bt = blockIdx.y;
  c.b = bt / g.Tw; c.t = bt - c.b * g.Tw;
  c.filt = (g.flags & WALDO_F_FILTER) != 0;
  __shared__ float s_P[(WB_MAX_L - 1) * WB_MAX_NL];
  __shared__ float s_occ[WB_MAX_L * WB_MAX_L];
  if (c.filt) for (int i = wb_tid(); i < g.No * g.Nl; i += wb_nthr()) s_P[i] = d.prof_p[(size_t)c.b * g.No * g.Nl + i];
  for (int i = wb_tid(); i < c.L * c.L; i += wb_nthr()) s_occ[i] = __ldg(d.occ + ((size_t)c.b * g.T + c.t) * c.L * c.L + i);
  __syncthreads();
  c.s_P = s_P; c.s_occ = s_occ;
  const float r = (float)g.H / (float)g.Hd;   // 1 / scale_hd
  c.lyt_base = reinterpret_cast<const ST*>(d.input) + (((size_t)c.b * g.T + c.t) * g.C + 3) * c.HWd;
  c.alo = d.a_lo + ((size_t)c.b * g.Tw + c.t) * c.L * c.HW;
  const uint32_t* live = d.live_ctx + ((size_t)c.b * g.Tw + c.t) * c.HW;
  c.out = d.alpha + ((size_t)c.b * g.Tw + c.t) * c.L * c.HWd;
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    for (int it = wb_tid(); it < WB_TILE_PX; it += wb_nthr()) {
      // threads beyond the image edge recompute (and re-store, identically) the nearest valid pixel: no predicates
      const int X = min(tx0 + (it & (WB_TILE_W - 1)), g.Wd - 1), Y = min(ty0 + it / WB_TILE_W, g.Hd - 1);
      const unsigned q = (unsigned)(Y * g.Wd + X);
      WbAxis ay = wb_axis(Y, r, g.H), ax = wb_axis(X, r, g.W);
      const int o00 = ay.i0 * g.W + ax.i0, o01 = ay.i0 * g.W + ax.i1, o10 = ay.i1 * g.W + ax.i0, o11 = ay.i1 * g.W + ax.i1;
      const unsigned wm = wb_warp_or(wb_live4(live, o00, o01, o10, o11));
      const int n = __popc(wm);
      if (WB_NA_VARIANTS_FWD >= 2 && n <= 4) wb_prep_pixel<4, NLC, ST>(d, c, wm, q, ax, ay, o00, o01, o10, o11);
      else if (WB_NA_VARIANTS_FWD >= 3 && n <= 8) wb_prep_pixel<8, NLC, ST>(d, c, wm, q, ax, ay, o00, o01, o10, o11);
      else wb_prep_pixel<WB_MAX_L, NLC, ST>(d, c, wm, q, ax, ay, o00, o01, o10, o11);
    }
  }
}

// ------------------------------------------------------------------ B5: per-layer flow on the low-res lattice (lvd.py:771-792)
// One thread per (b,tp,p) walks the layers: the bilinear taps of src_grid[b,u,k,p] in the layer's canonical frame are
// shared by all Tc contexts; value sampled = tgt_grid[b,c,k] - tgt_grid[b,u,k].  Also the object support
// s_lo = bil0(1) (lvd.py:788) and the live-layer mask of the target frame.
__global__ void k_layer_flow_lo(WbDec d) {
  const waldo_geom_t g = d.g;
  const int L = wb_L(g), HW = g.H * g.W;
  const long long total = (long long)g.B * g.Tp * HW;
  for (long long e = (long long)blockIdx.x * wb_nthr() + wb_tid(); e < total; e += (long long)gridDim.x * wb_nthr()) {
    int p = (int)(e % HW);
    int tp = (int)((e / HW) % g.Tp);
    int b = (int)(e / ((long long)HW * g.Tp));
    int u = (int)d.pred_ts[tp];
    unsigned live = 1u;
    for (int k = 0; k < L; ++k) {
      const float* sg; const float* tg_u; int w, h; size_t frame_stride, layer_off;
      if (k == 0) {
        w = g.W; h = g.H; frame_stride = (size_t)HW * 2; layer_off = 0;
        sg = d.src_grid_bg + (((size_t)b * g.T + u) * HW + p) * 2;
        tg_u = d.tgt_grid_bg + ((size_t)b * g.T + u) * frame_stride;
      } else {
        w = g.Wo; h = g.Ho; frame_stride = (size_t)g.No * g.Ho * g.Wo * 2; layer_off = (size_t)(k - 1) * g.Ho * g.Wo * 2;
        sg = d.src_grid_obj + ((((size_t)b * g.T + u) * g.No + (k - 1)) * HW + p) * 2;
        tg_u = d.tgt_grid_obj + ((size_t)b * g.T + u) * frame_stride + layer_off;
      }
      const float* tg_b = (k == 0 ? d.tgt_grid_bg : d.tgt_grid_obj) + (size_t)b * g.T * frame_stride + layer_off;
      WbTaps t = wb_taps(__ldg(sg), __ldg(sg + 1), w, h);
      int m = wb_tap_mask(t, w, h);
      if (m) live |= 1u << k;
      long long o = ((long long)t.y0 * w + t.x0) * 2;
      float un[4][2];
      WB_UNROLL for (int c = 0; c < 2; ++c) {
        un[0][c] = (m & 1) ? __ldg(tg_u + o + c) : 0.f;
        un[1][c] = (m & 2) ? __ldg(tg_u + o + 2 + c) : 0.f;
        un[2][c] = (m & 4) ? __ldg(tg_u + o + 2 * w + c) : 0.f;
        un[3][c] = (m & 8) ? __ldg(tg_u + o + 2 * w + 2 + c) : 0.f;
      }
      if (k >= 1 && d.s_lo)
        d.s_lo[(((size_t)b * g.Tp + tp) * g.No + (k - 1)) * HW + p] =
            wb_chain((m & 1) ? 1.f : 0.f, (m & 2) ? 1.f : 0.f, (m & 4) ? 1.f : 0.f, (m & 8) ? 1.f : 0.f, t);
      for (int tc = 0; tc < g.Tc; ++tc) {
        int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
        const float* tg_c = tg_b + (size_t)c_t * frame_stride;
        float f[2];
        WB_UNROLL for (int c = 0; c < 2; ++c) {
          float vnw = (m & 1) ? __fsub_rn(__ldg(tg_c + o + c), un[0][c]) : 0.f;
          float vne = (m & 2) ? __fsub_rn(__ldg(tg_c + o + 2 + c), un[1][c]) : 0.f;
          float vsw = (m & 4) ? __fsub_rn(__ldg(tg_c + o + 2 * w + c), un[2][c]) : 0.f;
          float vse = (m & 8) ? __fsub_rn(__ldg(tg_c + o + 2 * w + 2 + c), un[3][c]) : 0.f;
          f[c] = wb_chain(vnw, vne, vsw, vse, t);
        }
        float* out = d.f_lo + (((((size_t)b * g.Tc + tc) * g.Tp + tp) * L + k) * HW + p) * 2;
        out[0] = f[0]; out[1] = f[1];
      }
    }
    d.live_pred[e] = live;
  }
}
