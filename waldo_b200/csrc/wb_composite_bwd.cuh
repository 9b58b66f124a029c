// Backward of decode_output: the fused HD backward kernel, the context-alpha
// backward, and the low-res chain down to grids / opacities / class scores.
//
// Accumulation strategy:
//   * small reductions (d occ, d class profile, d cls): per-warp shuffle tree -> per-CTA shared slot ->
//     per-CTA partial in global scratch -> reduced in CTA order by a second kernel: deterministic, no atomics;
//   * low-res scatter targets of the HD kernels (d f_lo, d a_lo): the 32 pixels of a row are staged in shared memory and
//     each low-res column is reduced by one lane (transpose of the up-sampling), then two red.global per column;
//   * HD scatter targets (d input, d context opacity): red.global.add.f32 (wb_red), neighbouring lanes merged by shuffle.
// All layer loops run over the warp-wide union of live layers (see wb_composite.cuh); rows with <= 8 live layers use the
// lanes-per-layer form of the layer kernel.
//
// This file is included TWICE by waldo_abi.cu: with WB_DET 0 (namespace wb_plain: fire-and-forget float reductions, the
// default) and with WB_DET 1 (namespace wb_fixed: 64-bit fixed-point accumulation of every scatter target, run-to-run
// bit-identical gradients; include/waldo_b200.h det_* fields).  Two instantiations instead of a run-time switch, so that
// the default kernels carry no trace of the deterministic path (same SASS as before it existed).
#include "wb_common.cuh"
#include "wb_prep.cuh"
#include "wb_composite.cuh"

#ifndef WB_BWD_SHARED_DECLS
#define WB_BWD_SHARED_DECLS
typedef waldo_decode_bwd_t WbDecB;
#ifndef WB_HOST_EMU
static int wb_check_launch(const char* file, int line);
#endif
static int wb_fail(int code, const char* fmt, ...);
#endif

#undef WB_RED
#undef WB_RED_NZ
#if WB_DET
namespace wb_fixed {
// every call site has the kernel argument `a` in scope
#define WB_RED(p, v) wb_red_fixed(a.det_base, a.det_shadow, a.det_scale, (p), (v))
#define WB_RED_NZ(p, v) wb_atomic_add(a, (p), (v))
WB_DEV void wb_atomic_add(const WbDecB& a, float* p, float v) { if (v != 0.f) WB_RED(p, v); }
#else
namespace wb_plain {
#define WB_RED(p, v) wb_red((p), (v))
#define WB_RED_NZ(p, v) wb_atomic_add((p), (v))
WB_DEV void wb_atomic_add(float* p, float v) { if (v != 0.f) wb_red(p, v); }
#endif

#define WB_NWARP (WB_TILE_PX / 32)
#ifndef WB_PF_GB
#define WB_PF_GB 0   // L1 prefetch of the next context group's flow / score lines in k_gather_bwd_async: no effect (2.172 vs 2.176 ms)
#endif
#ifndef WB_PREP_OCC8
#define WB_PREP_OCC8 1   // context-alpha backward: 8-slot register form of the occlusion backward for rows with 5..8 live layers
#endif

// ---------------------------------------------------------------------------- transpose of the bilinear up-sampling
// A warp is one row of 32 HD pixels, so all its lanes share the two low-res rows and touch a short run of low-res
// columns.  Instead of 4 atomics per lane and value, the lanes stage their values in shared memory and the first
// `ncols` lanes each own one low-res column: they sum the (<= 8, scale_hd <= 4) lanes that touch it in lane order and
// issue two global reductions (one per low-res row).  No shared-memory atomics, no block barrier.
#ifdef WB_HOST_EMU
#define WB_WARP 1
#define WB_CPL 2          // low-res columns owned per lane (one emulated lane touches 2 columns)
#define __syncwarp() ((void)0)
#else
#define WB_WARP 32
#define WB_CPL 1
#endif
#define WB_COL_TAPS 8     // max lanes touching one low-res column
#define WB_STAGE_SLOTS 8  // values staged per round and lane
#define WB_STAGE_ROW (WB_WARP + 2 * WB_COL_TAPS)   // one staged value: 8 zeros | the lanes | 8 zeros (no bounds checks in the flush)
#define WB_STAGE_AT(v, lane) ((v) * WB_STAGE_ROW + WB_COL_TAPS + (lane))

struct WbColRed {
  int col[WB_CPL];                 // absolute low-res column owned by this lane (-1: none)
  int lo[WB_CPL];                  // lane of the first contributing HD column (may be negative: previous warp's pixel)
  float w[WB_CPL][WB_COL_TAPS];    // x-weights of lanes lo .. lo+7
  int row0, row1;                  // the two low-res rows (same for the whole warp)
  float wy0, wy1;
};

// up_tab[col] = { first HD column X whose up-sampling taps touch low-res column col with non-zero weight,
//                 the x-weights of X, X+1, ..., X+7 }   -- depends on (W, Wd) only
__global__ void k_up_tab(int W, int Wd, float* __restrict__ tab) {
  const float r = (float)W / (float)Wd;
  for (int col = blockIdx.x * wb_nthr() + wb_tid(); col < W; col += gridDim.x * wb_nthr()) {
    int lo = -1;
    float w[WB_COL_TAPS];
    for (int t = 0; t < WB_COL_TAPS; ++t) w[t] = 0.f;
    // only HD columns whose source position lies within (col - 1, col + 1) can touch `col` (two columns of margin)
    const float inv_r = (float)Wd / (float)W;
    const int Xa = max(0, (int)(((float)col - 1.f) * inv_r) - 2), Xb = min(Wd, (int)(((float)col + 2.f) * inv_r) + 3);
    for (int X = Xa; X < Xb; ++X) {
      const WbAxis ax = wb_axis(X, r, W);
      const float wv = (ax.i0 == col ? ax.l0 : 0.f) + (ax.i1 == col ? ax.l1 : 0.f);
      if (wv != 0.f) {
        if (lo < 0) lo = X;
        if (X - lo < WB_COL_TAPS) w[X - lo] = wv;
      }
    }
    tab[col * 9] = (float)(lo < 0 ? 0 : lo);
    for (int t = 0; t < WB_COL_TAPS; ++t) tab[col * 9 + 1 + t] = w[t];
  }
}

// x0 = HD column of lane 0 of this warp; (i0_first, i1_last) = low-res columns of the first / last lane
WB_DEV WbColRed wb_colred_setup(const float* __restrict__ tab, int x0, int i0_first, int i1_last, const WbAxis& ay) {
  const int lane = wb_lane();
  WbColRed cr;
  cr.row0 = ay.i0; cr.row1 = ay.i1; cr.wy0 = ay.l0; cr.wy1 = ay.l1;
  const int ncols = i1_last - i0_first + 1;
  WB_UNROLL for (int cpl = 0; cpl < WB_CPL; ++cpl) {
    const int jj = lane + cpl * WB_WARP;
    cr.col[cpl] = -1; cr.lo[cpl] = 0;
    WB_UNROLL for (int t = 0; t < WB_COL_TAPS; ++t) cr.w[cpl][t] = 0.f;
    if (jj < ncols) {
      const int col = i0_first + jj;
      cr.col[cpl] = col;
      const float* e = tab + col * 9;
      const int lo = (int)__ldg(e) - x0;
      if (lo < -(WB_COL_TAPS - 1) || lo > WB_WARP - 1) cr.col[cpl] = -1;   // none of this warp's lanes touches the column
      else {
        cr.lo[cpl] = lo;
        WB_UNROLL for (int t = 0; t < WB_COL_TAPS; ++t) cr.w[cpl][t] = __ldg(e + 1 + t);
      }
    }
  }
  return cr;
}

// reduce `nv` staged values per lane (s_stage[v * WB_WARP + lane]) into dst[(row * W + col) * stride + v * vstride]
WB_DEV void wb_colred_flush(const WbDecB& a, const WbColRed& cr, const float* s_stage, int nv, float* const* dst, int W, int stride) {
  WB_UNROLL for (int cpl = 0; cpl < WB_CPL; ++cpl) {
    if (cr.col[cpl] >= 0) {
      for (int v = 0; v < nv; ++v) {
        const float* sv = s_stage + WB_STAGE_AT(v, cr.lo[cpl]);   // lo in [-7, WARP-1]: always inside the padded row
        float acc = 0.f;
        WB_UNROLL for (int t = 0; t < WB_COL_TAPS; ++t) acc += cr.w[cpl][t] * sv[t];
        if (acc != 0.f) {
          WB_RED(dst[v] + ((size_t)cr.row0 * W + cr.col[cpl]) * stride, acc * cr.wy0);
          WB_RED(dst[v] + ((size_t)cr.row1 * W + cr.col[cpl]) * stride, acc * cr.wy1);
        }
      }
    }
  }
}

// exclusive-product backward of  A_i = R_i * prod_j (1 - R_j occ[j,i])  over the slots of `ix`.
// gR (+=) gets d/dR; when s_acc != nullptr, d/d occ[j,i] summed over the warp is added to s_acc[j*L+i] by lane 0.
// pairs_only: d occ is consumed by waldo_occ_bwd alone, which never reads row 0, column 0 or the diagonal (those
// entries of occ are constants, lvd.py:63-66) -- skip their warp reductions.
template <int NA>
WB_DEV void wb_occlude_bwd(const float* R, const float* gA, const float* s_occ, int L, const WbIdx<NA>& ix, float* gR, float* s_acc,
                           bool pairs_only) {
  const int lane = wb_lane();
  WB_UNROLL_NA for (int i = 0; i < WB_NEND; ++i) {
    if (i < ix.n) {
      float pre[NA];
      float run = 1.f;
      WB_UNROLL_NA for (int j = 0; j < WB_NEND; ++j) if (j < ix.n) { pre[j] = run; run *= 1.f - R[j] * s_occ[ix.k[j] * L + ix.k[i]]; }
      gR[i] += gA[i] * run;
      const float gV = gA[i] * R[i];
      float suf = 1.f;
      WB_UNROLL_NA for (int j = WB_NEND - 1; j >= 0; --j) {
        if (j < ix.n) {
          const float oc = s_occ[ix.k[j] * L + ix.k[i]];
          const float excl = pre[j] * suf;
          suf *= 1.f - R[j] * oc;
          gR[j] -= gV * oc * excl;
          if (s_acc && !(pairs_only && (ix.k[i] == 0 || ix.k[j] == 0 || ix.k[i] == ix.k[j]))) {
            float v = wb_warp_sum(-gV * R[j] * excl);
            if (lane == 0) s_acc[ix.k[j] * L + ix.k[i]] += v;
          }
        }
      }
    }
  }
}

// ============================================================================ fused HD backward
struct WbBwdCtx {   // per-CTA constants of the fused backward
  int b, tp, L, C, TcR, CR, HW;
  unsigned HWd;
  bool self, disocc_ch, need_layers, lowres_direct, pairs_only;
  const float* s_occ;
  float* s_acc;      // this warp's d occ accumulators (or null)
  float* s_stage;    // this warp's staging area: WB_STAGE_SLOTS * 2 * WB_WARP floats
};

// backward of the layer part: B9, B8, B7, B6, B5(up) of one (pixel, context).  gs = d/d score, (dfx, dfy) = d/d flow,
// draw = this pixel's upstream d raw_output (null = zero).
template <int NA>
WB_DEV void wb_bwd_layers_bwd(const WbDecB& a, const WbBwdCtx& c, const WbPix& px, const WbColRed& cr, const WbIdx<NA>& ix, int tc, int c_t,
                              size_t pair, const float* __restrict__ draw, float actf, float gs, float dfx, float dfy) {
  const WbDec& d = a.f;
  const waldo_geom_t& g = d.g;
  const int L = c.L, C = c.C, HW = c.HW;
  const unsigned HWd = c.HWd;
  const float* alpha_c = d.alpha + ((size_t)c.b * g.Tw + c_t) * L * HWd;
  WbLay<NA> ly;
  wb_layers_fwd<NA>(d, px, ix, d.f_lo + pair * L * HW * 2, alpha_c, c.s_occ, ly);
  float gA[NA], gR[NA], gFx[NA], gFy[NA];
  WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s) {
    gR[s] = 0.f; gA[s] = 0.f; gFx[s] = 0.f; gFy[s] = 0.f;
    if (s < ix.n) {
      gA[s] = gs + 2.f * (draw ? actf * __ldg(draw + (size_t)(C + ix.k[s]) * HWd) : 0.f) + dfx * ly.Fx[s] + dfy * ly.Fy[s];
      gFx[s] = ly.A[s] * dfx; gFy[s] = ly.A[s] * dfy;
    }
  }
  wb_occlude_bwd<NA>(ly.R, gA, c.s_occ, L, ix, gR, c.s_acc, c.pairs_only);
  // ---- B7 backward: disocc = max_k R_k (first maximal layer takes the gradient)
  if (c.disocc_ch && draw) {
    const float gd = actf * __ldg(draw + (size_t)(C + L) * HWd);
    bool done = false;
    WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s)
      if (s < ix.n && !done && ly.R[s] == ly.disocc) { gR[s] += gd; done = true; }
  }
  // ---- B6 backward: bilinear sample of the context opacity through layer k's flow
  float* dal = a.d_alpha_acc ? a.d_alpha_acc + ((size_t)c.b * g.Tw + c_t) * L * HWd : nullptr;
  WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s) {
    if (s < ix.n) {
      const int k = ix.k[s];
      if (((px.isobj >> k) & 1u) && gR[s] != 0.f) {
        const WbTaps tk = wb_taps(__fadd_rn(px.gx, ly.Fx[s]), __fadd_rn(px.gy, ly.Fy[s]), g.Wd, g.Hd);
        const WbTap2 t2 = wb_tap2(tk, g.Wd, g.Hd);
        float cx[4], cy[4];
        wb_pos4(t2, -tk.wy0, tk.wy0, -tk.wy1, tk.wy1, cx);
        wb_pos4(t2, -tk.wx0, -tk.wx1, tk.wx0, tk.wx1, cy);
        const float* p0 = alpha_c + (size_t)k * HWd + t2.o0;
        const float* p1 = alpha_c + (size_t)k * HWd + t2.o1;
        const float v0 = (__ldg(p0) + 1.f) * 0.5f, v1 = (__ldg(p0 + 1) + 1.f) * 0.5f;
        const float v2 = (__ldg(p1) + 1.f) * 0.5f, v3 = (__ldg(p1 + 1) + 1.f) * 0.5f;
        const float gr = gR[s];
        gFx[s] += gr * (v0 * cx[0] + v1 * cx[1] + v2 * cx[2] + v3 * cx[3]) * (0.5f * (float)g.Wd);
        gFy[s] += gr * (v0 * cy[0] + v1 * cy[1] + v2 * cy[2] + v3 * cy[3]) * (0.5f * (float)g.Hd);
        if (dal) {
          float* o0 = dal + (size_t)k * HWd + t2.o0;
          float* o1 = dal + (size_t)k * HWd + t2.o1;
          WB_RED_NZ(o0, t2.w[0] * gr); WB_RED_NZ(o0 + 1, t2.w[1] * gr);
          WB_RED_NZ(o1, t2.w[2] * gr); WB_RED_NZ(o1 + 1, t2.w[3] * gr);
        }
      }
    }
  }
  // ---- B5(up) backward: transpose of the bilinear up-sampling of the layer flows
  if (a.d_f_lo) {
    if (c.lowres_direct) {
      WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s)
        if (s < ix.n) {
          float* o = a.d_f_lo + (pair * L + ix.k[s]) * HW * 2 + (size_t)px.o00 * 2;
          WB_RED_NZ(o, gFx[s]); WB_RED_NZ(o + 1, gFy[s]);
        }
    } else {
      const int lane = wb_lane();
      if constexpr (NA <= WB_STAGE_SLOTS) {
        float* dst[2 * NA];
        WB_UNROLL for (int s = 0; s < NA; ++s) {
          c.s_stage[WB_STAGE_AT(2 * s, lane)] = gFx[s];
          c.s_stage[WB_STAGE_AT(2 * s + 1, lane)] = gFy[s];
          float* base = a.d_f_lo + (pair * L + ix.k[s]) * HW * 2;
          dst[2 * s] = base; dst[2 * s + 1] = base + 1;
        }
        __syncwarp();
        wb_colred_flush(a, cr, c.s_stage, 2 * ix.n, dst, g.W, 2);
        __syncwarp();
      } else {
        for (int s0 = 0; s0 < ix.n; s0 += WB_STAGE_SLOTS) {
          float* dst[2 * WB_STAGE_SLOTS];
          const int ns = min(WB_STAGE_SLOTS, ix.n - s0);
          for (int j = 0; j < ns; ++j) {
            c.s_stage[WB_STAGE_AT(2 * j, lane)] = gFx[s0 + j];
            c.s_stage[WB_STAGE_AT(2 * j + 1, lane)] = gFy[s0 + j];
            float* base = a.d_f_lo + (pair * L + ix.k[s0 + j]) * HW * 2;
            dst[2 * j] = base; dst[2 * j + 1] = base + 1;
          }
          __syncwarp();
          wb_colred_flush(a, cr, c.s_stage, 2 * ns, dst, g.W, 2);
          __syncwarp();
        }
      }
    }
  }
}

// all contexts of one pixel: the slot list is built once
template <int NA>
WB_DEV void wb_bwd_layers_ctxs(const WbDecB& a, const WbBwdCtx& c, const WbPix& px, const WbColRed& cr, unsigned wm, unsigned q, float actf) {
  const WbDec& d = a.f;
  const waldo_geom_t& g = d.g;
  const int b = c.b, tp = c.tp;
  const unsigned HWd = c.HWd;
  const WbIdx<NA> ix = wb_idx<NA>(wm);
  for (int tc = 0; tc < g.Tc; ++tc) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
    const float* gl = a.glue + pair * 3 * HWd + q;
    const float gs = actf * __ldg(gl), dfx = actf * __ldg(gl + HWd), dfy = actf * __ldg(gl + 2 * HWd);
    const float* draw = a.d_raw_output ? a.d_raw_output + (((size_t)b * c.TcR + tc) * g.Tp + tp) * c.CR * HWd + q : nullptr;
    wb_bwd_layers_bwd<NA>(a, c, px, cr, ix, tc, c_t, pair, draw, actf, gs, dfx, dfy);
  }
}

#ifndef WB_HOST_EMU
// ============================================================================ lanes-per-layer form of the layer backward
// A warp owns a row of 32 pixels whose union of live layers has n <= 8 members.  Instead of one lane per pixel looping
// over the n layers (serial dependent loads, per-thread arrays), the warp makes LP = 1|2|4|8 >= n passes over the row
// with LP lanes per pixel: lane = slot * (32/LP) + pixel, every lane owns ONE (pixel, layer).  All layers of a pixel
// load in parallel, the occlusion products run over warp shuffles, and there is no per-thread array indexed by a layer.
// n > 8 (dense rows, or training without is_obj) keeps the one-lane-per-pixel form above.
// v[j] (j < LP) of lane (pixel, slot i)  ->  v[0] of lane (pixel, slot j) = sum_i v_i[j]   (LP - 1 shuffles)
template <int LP>
WB_DEV void wb_slot_transpose_sum(float* v, int slot) {
  constexpr int PPW = 32 / LP;
  WB_UNROLL for (int o = LP / 2; o >= 1; o >>= 1) {
    const bool up = (slot & o) != 0;
    WB_UNROLL for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o * PPW);
    }
  }
}

// flush the staged per-(slot, pixel) values of one row: value rows 2*s (x) and 2*s+1 (y) of slot s go to layer k_s of d_f_lo
WB_DEV void wb_colred_flush_slots(const WbDecB& a, const WbColRed& cr, const float* s_stage, unsigned wm, int ncomp, float* base, size_t layer_stride,
                                  int W, int stride) {
  if (cr.col[0] < 0) return;
  int s = 0;
  for (unsigned m = wm; m; m &= m - 1u, ++s) {
    const int k = __ffs((int)m) - 1;
    for (int comp = 0; comp < ncomp; ++comp) {
      const float* sv = s_stage + WB_STAGE_AT(ncomp * s + comp, cr.lo[0]);
      float acc = 0.f;
      WB_UNROLL for (int t = 0; t < WB_COL_TAPS; ++t) acc += cr.w[0][t] * sv[t];
      if (acc != 0.f) {
        float* dst = base + (size_t)k * layer_stride + comp;
        WB_RED(dst + ((size_t)cr.row0 * W + cr.col[0]) * stride, acc * cr.wy0);
        WB_RED(dst + ((size_t)cr.row1 * W + cr.col[0]) * stride, acc * cr.wy1);
      }
    }
  }
}

template <int LP>
WB_DEV void wb_lanes_layers_bwd(const WbDecB& a, const WbBwdCtx& c, const WbColRed& cr, unsigned wm, int n, unsigned isobj_lane,
                                int tx0, bool rowact, int Y, const WbAxis& ay, float gy) {
  constexpr int PPW = 32 / LP;
  const WbDec& d = a.f;
  const waldo_geom_t& g = d.g;
  const int L = c.L, C = c.C, HW = c.HW, b = c.b, tp = c.tp;
  const unsigned HWd = c.HWd;
  const int lane = wb_lane(), pl = lane % PPW, slot = lane / PPW;
  const bool valid = slot < n;
  const int k = valid ? wb_nth_bit(wm, slot) : 0;
  float oc[LP], accj[LP];
  WB_UNROLL for (int j = 0; j < LP; ++j) {
    accj[j] = 0.f;
    oc[j] = (valid && j < n) ? c.s_occ[wb_nth_bit(wm, j) * L + k] : 0.f;
  }
  const float r_lo = (float)g.H / (float)g.Hd;
  const bool stage = a.d_f_lo && !c.lowres_direct;
  for (int tc = 0; tc < g.Tc; ++tc) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    const size_t pair = ((size_t)b * g.Tc + tc) * g.Tp + tp;
    const float2* fl = reinterpret_cast<const float2*>(d.f_lo) + (pair * L + k) * HW;
    const float* alpha_k = d.alpha + (((size_t)b * g.Tw + c_t) * L + k) * HWd;
    float* dal_k = a.d_alpha_acc ? a.d_alpha_acc + (((size_t)b * g.Tw + c_t) * L + k) * HWd : nullptr;
    const float* gl = a.glue + pair * 3 * HWd;
    const float* draw = a.d_raw_output ? a.d_raw_output + (((size_t)b * c.TcR + tc) * g.Tp + tp) * c.CR * HWd : nullptr;
#pragma unroll 1
    for (int r = 0; r < LP; ++r) {
      const int p = r * PPW + pl, Xr = tx0 + p, X = min(Xr, g.Wd - 1);
      const float actf = (rowact && Xr < g.Wd) ? 1.f : 0.f;
      const unsigned q = (unsigned)(Y * g.Wd + X);
      const WbAxis ax = wb_axis(X, r_lo, g.W);
      const float gx = __ldg(d.xs_hd + X);
      const int o00 = ay.i0 * g.W + ax.i0, o01 = ay.i0 * g.W + ax.i1, o10 = ay.i1 * g.W + ax.i0, o11 = ay.i1 * g.W + ax.i1;
      const unsigned isobj = __shfl_sync(0xffffffffu, isobj_lane, p);
      // ---- every load that does not depend on another load is issued first (invalid lanes read layer 0: harmless)
      const float gs_l = __ldg(gl + q), dfx_l = __ldg(gl + HWd + q), dfy_l = __ldg(gl + 2 * HWd + q);
      const float dr_l = draw ? __ldg(draw + (size_t)(C + k) * HWd + q) : 0.f;
      float2 f00 = __ldg(fl + o00), f01 = f00, f10 = f00, f11 = f00;
      if (!c.lowres_direct) { f01 = __ldg(fl + o01); f10 = __ldg(fl + o10); f11 = __ldg(fl + o11); }
#if WB_PF_FLO
      if (tc + 1 < g.Tc) {   // the same cells of the next context: pair + Tp
        const float2* fn = fl + (size_t)g.Tp * L * HW;
        wb_prefetch_l1(fn + o00);
        if (!c.lowres_direct) wb_prefetch_l1(fn + o10);
      }
#endif
      // ---- forward of this (pixel, layer), same arithmetic as wb_layers_fwd
      float Fx = 0.f, Fy = 0.f, rr = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      bool samp = false;
      WbTaps tk; WbTap2 t2;
      if (valid) {
        if (c.lowres_direct) { Fx = f00.x; Fy = f00.y; }
        else {
          Fx = wb_lerp2(f00.x, f01.x, f10.x, f11.x, ax, ay);
          Fy = wb_lerp2(f00.y, f01.y, f10.y, f11.y, ax, ay);
        }
        if ((isobj >> k) & 1u) {
          samp = true;
          tk = wb_taps(__fadd_rn(gx, Fx), __fadd_rn(gy, Fy), g.Wd, g.Hd);
          t2 = wb_tap2(tk, g.Wd, g.Hd);
          const float* p0 = alpha_k + t2.o0;
          const float* p1 = alpha_k + t2.o1;
          v0 = (__ldg(p0) + 1.f) * 0.5f; v1 = (__ldg(p0 + 1) + 1.f) * 0.5f;
          v2 = (__ldg(p1) + 1.f) * 0.5f; v3 = (__ldg(p1 + 1) + 1.f) * 0.5f;
          rr = __fmaf_rn(v3, t2.w[3], __fmaf_rn(v2, t2.w[2], __fmaf_rn(v1, t2.w[1], __fmul_rn(v0, t2.w[0]))));
        }
      }
      float Rj[LP], pre[LP];
      float run = 1.f;
      WB_UNROLL for (int j = 0; j < LP; ++j) {
        Rj[j] = __shfl_sync(0xffffffffu, rr, pl + j * PPW);
        pre[j] = run;
        run *= 1.f - Rj[j] * oc[j];
      }
      const float A = run * rr;
      // ---- upstream of this (pixel, layer)
      const float gs = actf * gs_l, dfx = actf * dfx_l, dfy = actf * dfy_l;
      const float gA = valid ? gs + 2.f * (actf * dr_l) + dfx * Fx + dfy * Fy : 0.f;
      float gFx = A * dfx, gFy = A * dfy;
      // ---- exclusive-product backward over the slot lanes
      float gR = gA * run;
      const float gV = gA * rr;
      float term[LP];
      float suf = 1.f;
      WB_UNROLL for (int j = LP - 1; j >= 0; --j) {
        const float excl = pre[j] * suf;
        suf *= 1.f - Rj[j] * oc[j];
        term[j] = -gV * oc[j] * excl;
        accj[j] += -gV * Rj[j] * excl;
      }
      wb_slot_transpose_sum<LP>(term, slot);
      gR += term[0];
      if (c.disocc_ch && draw) {   // B7 backward: the first maximal layer takes the gradient
        float mx = rr;
        WB_UNROLL for (int o = PPW; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        int first = (valid && rr == mx) ? slot : 99;
        WB_UNROLL for (int o = PPW; o < 32; o <<= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        if (slot == first) gR += actf * __ldg(draw + (size_t)(C + L) * HWd + q);
      }
      // ---- B6 backward: bilinear sample of the context opacity through this layer's flow
      if (samp && gR != 0.f) {
        float cx[4], cy[4];
        wb_pos4(t2, -tk.wy0, tk.wy0, -tk.wy1, tk.wy1, cx);
        wb_pos4(t2, -tk.wx0, -tk.wx1, tk.wx0, tk.wx1, cy);
        gFx += gR * (v0 * cx[0] + v1 * cx[1] + v2 * cx[2] + v3 * cx[3]) * (0.5f * (float)g.Wd);
        gFy += gR * (v0 * cy[0] + v1 * cy[1] + v2 * cy[2] + v3 * cy[3]) * (0.5f * (float)g.Hd);
        if (dal_k) {
          float* q0 = dal_k + t2.o0;
          float* q1 = dal_k + t2.o1;
          WB_RED_NZ(q0, t2.w[0] * gR); WB_RED_NZ(q0 + 1, t2.w[1] * gR);
          WB_RED_NZ(q1, t2.w[2] * gR); WB_RED_NZ(q1 + 1, t2.w[3] * gR);
        }
      }
      // ---- B5(up) backward
      if (a.d_f_lo && valid) {
        if (c.lowres_direct) {
          float* o = a.d_f_lo + (pair * L + k) * HW * 2 + (size_t)o00 * 2;
          WB_RED_NZ(o, gFx); WB_RED_NZ(o + 1, gFy);
        } else {
          c.s_stage[WB_STAGE_AT(2 * slot, p)] = gFx;
          c.s_stage[WB_STAGE_AT(2 * slot + 1, p)] = gFy;
        }
      }
    }
    if (stage) {
      __syncwarp();
      wb_colred_flush_slots(a, cr, c.s_stage, wm, 2, a.d_f_lo + pair * L * HW * 2, (size_t)HW * 2, g.W, 2);
      __syncwarp();
    }
  }
  if (c.s_acc) {   // d occ: pixels of the row summed over the pixel lanes, one lane per (j, i) pair adds to the warp's slot
    WB_UNROLL for (int j = 0; j < LP; ++j) {
      float v = accj[j];
      WB_UNROLL for (int o = PPW / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (pl == 0 && valid && j < n) {
        const int kj = wb_nth_bit(wm, j);
        if (!(c.pairs_only && (k == 0 || kj == 0 || k == kj))) c.s_acc[kj * L + k] += v;
      }
    }
  }
}
#endif  // !WB_HOST_EMU

// ------------------------------------------------------------------------------------------------------------------
// The backward of the two HD kernels, in reverse order:
//   k_gather_bwd : stage C backward.  Scatters d input and reduces, per (pixel, context), the upstream gradients
//                  of the image channels to three numbers: d score, d flow x, d flow y (`glue`).
//   k_layers_bwd : backward of the layer part (B9..B5up) driven by `glue` and the alpha channels of d raw_output.
// ------------------------------------------------------------------------------------------------------------------

// grid = (CTAs, B*Tp), 32x8 pixel tiles.  The contexts are processed TG at a time (small register state -> high
// occupancy); per group ONE rolled loop over the image channels with the TG contexts unrolled inside.  Since the
// gathered value is bilinear in the four taps, d score and d flow follow from the tap moments
//   U_j = sum_ch dOut_ch * v_j,   T_j = sum_ch dRaw_ch * v_j     (j = the four tap positions).
// FAST = the common full case, resolved at compile time (no predicates, no divergence bookkeeping in the channel loop):
// Tc a multiple of TG, every upstream gradient and d_input present, no include_self context.
template <int TG, bool FAST>
__global__ void __launch_bounds__(WB_TILE_PX, WB_OCC_GATHER_BWD) k_gather_bwd(WbDecB a) {
  const WbDec& d = a.f;
  const waldo_geom_t g = d.g;
  const int C = g.C, L = g.No + 1;
  const unsigned HWd = (unsigned)(g.Hd * g.Wd);
  const int btp = blockIdx.y, b = btp / g.Tp, tp = btp - b * g.Tp;
  const bool self = !FAST && (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
  const int TcR = g.Tc + (self ? 1 : 0), CR = C + L + ((g.flags & WALDO_F_USE_DISOCC) ? 1 : 0);
  __shared__ const float* s_src[8];    // context frame of every context (CTA-uniform)
  __shared__ float* s_dsrc[8];         // its gradient
  __shared__ const float* s_draw[8];   // upstream d raw_output block of every context (or null)
  for (int tc = wb_tid(); tc < g.Tc; tc += wb_nthr()) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    s_src[tc] = d.input + ((size_t)b * g.T + c_t) * C * HWd;
    s_dsrc[tc] = a.d_input ? a.d_input + ((size_t)b * g.T + c_t) * C * HWd : nullptr;
    s_draw[tc] = a.d_raw_output ? a.d_raw_output + (((size_t)b * TcR + tc) * g.Tp + tp) * CR * HWd : nullptr;
  }
  __syncthreads();
  const bool has_din = FAST || a.d_input != nullptr, has_draw = FAST || a.d_raw_output != nullptr;
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    for (int it = wb_tid(); it < WB_TILE_PX; it += wb_nthr()) {
      const int Xr = tx0 + (it & (WB_TILE_W - 1)), Yr = ty0 + it / WB_TILE_W;
      const bool active = Xr < g.Wd && Yr < g.Hd;
      const float actf = active ? 1.f : 0.f;    // threads beyond the edge run on the nearest valid pixel with zero upstream
      const int X = min(Xr, g.Wd - 1), Y = min(Yr, g.Hd - 1);
      const unsigned q = (unsigned)(Y * g.Wd + X);
      const float gx = __ldg(d.xs_hd + X), gy = __ldg(d.ys_hd + Y);
      const float D = fmaxf(__ldg(d.norm + ((size_t)b * g.Tp + tp) * HWd + q), 1e-12f);
      const float* dof = a.d_output ? a.d_output + ((size_t)b * g.Tp + tp) * C * HWd + q : nullptr;
      const float* dra = a.d_raw_alpha ? a.d_raw_alpha + ((size_t)b * g.Tp + tp) * HWd + q : nullptr;
      const float* of = d.out_full + ((size_t)b * g.Tp + tp) * (C + 1) * HWd + q;
      float S = 0.f;   // sum_ch dOut * out, complete after the first group
      for (int tc0 = 0; tc0 < g.Tc; tc0 += TG) {
        unsigned o0[TG], o1[TG];
        float w[TG][4], nrm[TG], U[TG][4], Tq[TG][4];
        WB_UNROLL for (int i = 0; i < TG; ++i) {
          o0[i] = 0u; o1[i] = 0u; nrm[i] = 0.f;
          WB_UNROLL for (int j = 0; j < 4; ++j) { w[i][j] = 0.f; U[i][j] = 0.f; Tq[i][j] = 0.f; }
          if (FAST || tc0 + i < g.Tc) {
            const size_t pair = ((size_t)b * g.Tc + tc0 + i) * g.Tp + tp;
            const float* fl = d.flow + pair * 2 * HWd + q;
            const WbTaps t = wb_taps(__fadd_rn(gx, __ldg(fl)), __fadd_rn(gy, __ldg(fl + HWd)), g.Wd, g.Hd);
            const WbTap2 t2 = wb_tap2(t, g.Wd, g.Hd);
            o0[i] = t2.o0; o1[i] = t2.o1;
            WB_UNROLL for (int j = 0; j < 4; ++j) w[i][j] = t2.w[j];
            nrm[i] = (__ldg(d.score + pair * HWd + q) + 1e-6f) / D;
          }
        }
        // Neighbouring pixels of a row mostly hit neighbouring source cells: where lane l+1's left taps are lane l's right
        // taps (o[l+1] == o[l] + 1), lane l hands its right-tap contributions to lane l+1 (one shuffle each) instead of
        // issuing its own reductions -- about half of the global reductions in smooth-flow regions.
        bool skipR0[TG], skipR1[TG], mergeL0[TG], mergeL1[TG];
        WB_UNROLL for (int i = 0; i < TG; ++i) {
          skipR0[i] = skipR1[i] = mergeL0[i] = mergeL1[i] = false;
#if !defined(WB_HOST_EMU) && WB_GB_MERGE
          if (FAST) {
            const int lane = wb_lane();
            const unsigned n0 = __shfl_down_sync(0xffffffffu, o0[i], 1), n1 = __shfl_down_sync(0xffffffffu, o1[i], 1);
            skipR0[i] = lane < 31 && n0 == o0[i] + 1u;
            skipR1[i] = lane < 31 && n1 == o1[i] + 1u;
            mergeL0[i] = __shfl_up_sync(0xffffffffu, (int)skipR0[i], 1) != 0 && lane > 0;
            mergeL1[i] = __shfl_up_sync(0xffffffffu, (int)skipR1[i], 1) != 0 && lane > 0;
          }
#endif
        }
        const bool first = tc0 == 0;
        float* dself = (first && self && has_din) ? a.d_input + ((size_t)b * g.T + tp) * C * HWd + q : nullptr;
        const float* drself = (self && has_draw) ? a.d_raw_output + (((size_t)b * TcR + g.Tc) * g.Tp + tp) * CR * HWd + q : nullptr;
        unsigned choff = 0u;   // ch * HWd
        WB_UNROLL_N(WB_GB_UNROLL)
        for (int ch = 0; ch < C; ++ch) {
          // all loads of this channel first (read-only path), then the arithmetic and the reductions
          float v[TG][4], gd[TG];
          const float gO = (FAST || dof) ? actf * __ldg(dof + choff) : 0.f;
          WB_UNROLL for (int i = 0; i < TG; ++i) {
            if (FAST || tc0 + i < g.Tc) {
              const float* pl = s_src[tc0 + i] + choff;
              const float* p0 = pl + o0[i];
              const float* p1 = pl + o1[i];
              v[i][0] = __ldg(p0); v[i][1] = __ldg(p0 + 1); v[i][2] = __ldg(p1); v[i][3] = __ldg(p1 + 1);
              gd[i] = has_draw ? __ldg(s_draw[tc0 + i] + choff + q) : 0.f;
            }
          }
          if (first && (FAST || dof)) S += gO * __ldg(of + choff);
          WB_UNROLL for (int i = 0; i < TG; ++i) {
            if (FAST || tc0 + i < g.Tc) {
              const float gdt = actf * gd[i];
              const float go = gdt + nrm[i] * gO;
              WB_UNROLL for (int j = 0; j < 4; ++j) { U[i][j] += gO * v[i][j]; Tq[i][j] += gdt * v[i][j]; }
              if (has_din) {
                float* dl = s_dsrc[tc0 + i] + choff;
#if !defined(WB_HOST_EMU) && WB_GB_MERGE
                if (FAST) {
                  const float c1 = w[i][1] * go, c3 = w[i][3] * go;
                  const float r1 = __shfl_up_sync(0xffffffffu, c1, 1), r3 = __shfl_up_sync(0xffffffffu, c3, 1);
                  WB_RED(dl + o0[i], w[i][0] * go + (mergeL0[i] ? r1 : 0.f));
                  if (!skipR0[i]) WB_RED(dl + o0[i] + 1, c1);
                  WB_RED(dl + o1[i], w[i][2] * go + (mergeL1[i] ? r3 : 0.f));
                  if (!skipR1[i]) WB_RED(dl + o1[i] + 1, c3);
                } else
#endif
                {
                  WB_RED(dl + o0[i], w[i][0] * go); WB_RED(dl + o0[i] + 1, w[i][1] * go);
                  WB_RED(dl + o1[i], w[i][2] * go); WB_RED(dl + o1[i] + 1, w[i][3] * go);
                }
              }
            }
          }
          if (dself && active) {   // lvd.py:845: the target frame passes straight through
            const float nself = (1.f + 1e-6f) / D;
            WB_RED_NZ(dself + choff, (drself ? __ldg(drself + choff) : 0.f) + nself * gO);
          }
          choff += HWd;
        }
        if (!a.glue || !active) continue;   // (threads beyond the edge must not overwrite the pixel they mirror)
        const float gOs = dra ? __ldg(dra) : 0.f;   // d / d (fused score channel), index C of out_full
        if (first && dra) S += gOs * __ldg(of + choff);
        WB_UNROLL for (int i = 0; i < TG; ++i) {
          if (FAST || tc0 + i < g.Tc) {
            const size_t pair = ((size_t)b * g.Tc + tc0 + i) * g.Tp + tp;
            const float* fl = d.flow + pair * 2 * HWd + q;
            const WbTaps t = wb_taps(__fadd_rn(gx, __ldg(fl)), __fadd_rn(gy, __ldg(fl + HWd)), g.Wd, g.Hd);
            const WbTap2 t2 = wb_tap2(t, g.Wd, g.Hd);
            float cx[4], cy[4];
            wb_pos4(t2, -t.wy0, t.wy0, -t.wy1, t.wy1, cx);
            wb_pos4(t2, -t.wx0, -t.wx1, t.wx0, t.wx1, cy);
            const float sc = nrm[i] * D - 1e-6f;
            float G = U[i][0] * w[i][0] + U[i][1] * w[i][1] + U[i][2] * w[i][2] + U[i][3] * w[i][3];
            float gix = 0.f, giy = 0.f;
            WB_UNROLL for (int j = 0; j < 4; ++j) {
              const float tj = Tq[i][j] + nrm[i] * U[i][j];
              gix += tj * cx[j]; giy += tj * cy[j];
            }
            G += gOs * (sc * 2.f - 1.f);
            const float* dfl = a.d_flow ? a.d_flow + pair * 2 * HWd + q : nullptr;
            float* gl = a.glue + pair * 3 * HWd + q;
            gl[0] = 2.f * nrm[i] * gOs + (G - S) / D;
            gl[HWd] = (dfl ? __ldg(dfl) : 0.f) + gix * (0.5f * (float)g.Wd);
            gl[2 * HWd] = (dfl ? __ldg(dfl + HWd) : 0.f) + giy * (0.5f * (float)g.Hd);
          }
        }
      }
    }
  }
}

#ifndef WB_HOST_EMU
// ------------------------------------------------------------------------------------------------------------------
// k_gather_bwd with an asynchronous-copy pipeline (the FAST case only).  Every load of the channel loop has an address
// that is known before the loop starts (fixed taps + ch * HWd), so the kernel is limited by how many loads a thread
// keeps in flight, i.e. by registers.  cp.async (LDGSTS) moves each thread's operands of the next WB_GB_DEPTH - 1
// channels into its own shared-memory slots without holding registers: (5 TG + 2) x (DEPTH - 1) loads in flight per
// thread instead of the ~12 the register file allows.  A thread only ever reads its own slots: no barrier, just
// cp.async.wait_group.

#ifndef WB_GB_DEPTH
#define WB_GB_DEPTH 3
#endif

template <int TG>
__global__ void __launch_bounds__(WB_TILE_PX, WB_OCC_GATHER_BWD) k_gather_bwd_async(WbDecB a) {
  constexpr int NF = 5 * TG + 2;            // per channel: TG x (4 taps + d raw) + d out + out
  constexpr int DEPTH = WB_GB_DEPTH;
  const WbDec& d = a.f;
  const waldo_geom_t g = d.g;
  const int C = g.C, L = g.No + 1;
  const unsigned HWd = (unsigned)(g.Hd * g.Wd);
  const int btp = blockIdx.y, b = btp / g.Tp, tp = btp - b * g.Tp;
  const int CR = C + L + ((g.flags & WALDO_F_USE_DISOCC) ? 1 : 0);
  extern __shared__ __align__(16) float s_ring[];   // [DEPTH][NF][WB_TILE_PX]
  __shared__ const float* s_src[8];
  __shared__ float* s_dsrc[8];
  __shared__ const float* s_draw[8];
  for (int tc = wb_tid(); tc < g.Tc; tc += wb_nthr()) {
    const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
    s_src[tc] = d.input + ((size_t)b * g.T + c_t) * C * HWd;
    s_dsrc[tc] = a.d_input + ((size_t)b * g.T + c_t) * C * HWd;
    s_draw[tc] = a.d_raw_output + (((size_t)b * g.Tc + tc) * g.Tp + tp) * CR * HWd;
  }
  __syncthreads();
  float* my = s_ring + wb_tid();
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    const int it = wb_tid();
    const int Xr = tx0 + (it & (WB_TILE_W - 1)), Yr = ty0 + it / WB_TILE_W;
    const bool active = Xr < g.Wd && Yr < g.Hd;
    const float actf = active ? 1.f : 0.f;
    const int X = min(Xr, g.Wd - 1), Y = min(Yr, g.Hd - 1);
    const unsigned q = (unsigned)(Y * g.Wd + X);
    const float gx = __ldg(d.xs_hd + X), gy = __ldg(d.ys_hd + Y);
    const float D = fmaxf(__ldg(d.norm + ((size_t)b * g.Tp + tp) * HWd + q), 1e-12f);
    const float* dof = a.d_output + ((size_t)b * g.Tp + tp) * C * HWd + q;
    const float* dra = a.d_raw_alpha ? a.d_raw_alpha + ((size_t)b * g.Tp + tp) * HWd + q : nullptr;
    const float* of = d.out_full + ((size_t)b * g.Tp + tp) * (C + 1) * HWd + q;
    float S = 0.f;
    for (int tc0 = 0; tc0 < g.Tc; tc0 += TG) {
#if WB_PF_GB
      if (tc0 + TG < g.Tc) {   // this pixel's flow / score lines of the next context group: their latency heads its prologue
        WB_UNROLL for (int i = 0; i < TG; ++i) {
          const size_t pn = ((size_t)b * g.Tc + tc0 + TG + i) * g.Tp + tp;
          wb_prefetch_l1(d.flow + pn * 2 * HWd + q); wb_prefetch_l1(d.flow + pn * 2 * HWd + HWd + q);
          wb_prefetch_l1(d.score + pn * HWd + q);
        }
      }
#endif
      unsigned o0[TG], o1[TG];
      float w[TG][4], nrm[TG], U[TG][4], Tq[TG][4];
      const float* src[TG];
      const float* drw[TG];
      float* dsr[TG];
      WB_UNROLL for (int i = 0; i < TG; ++i) {
        WB_UNROLL for (int j = 0; j < 4; ++j) { U[i][j] = 0.f; Tq[i][j] = 0.f; }
        const size_t pair = ((size_t)b * g.Tc + tc0 + i) * g.Tp + tp;
        const float* fl = d.flow + pair * 2 * HWd + q;
        const WbTaps t = wb_taps(__fadd_rn(gx, __ldg(fl)), __fadd_rn(gy, __ldg(fl + HWd)), g.Wd, g.Hd);
        const WbTap2 t2 = wb_tap2(t, g.Wd, g.Hd);
        o0[i] = t2.o0; o1[i] = t2.o1;
        WB_UNROLL for (int j = 0; j < 4; ++j) w[i][j] = t2.w[j];
        nrm[i] = (__ldg(d.score + pair * HWd + q) + 1e-6f) / D;
        src[i] = s_src[tc0 + i]; drw[i] = s_draw[tc0 + i] + q; dsr[i] = s_dsrc[tc0 + i];
      }
      bool skipR0[TG], skipR1[TG], mergeL0[TG], mergeL1[TG];   // see k_gather_bwd
      WB_UNROLL for (int i = 0; i < TG; ++i) {
        const int lane = wb_lane();
        const unsigned n0 = __shfl_down_sync(0xffffffffu, o0[i], 1), n1 = __shfl_down_sync(0xffffffffu, o1[i], 1);
        skipR0[i] = WB_GB_MERGE && lane < 31 && n0 == o0[i] + 1u;
        skipR1[i] = WB_GB_MERGE && lane < 31 && n1 == o1[i] + 1u;
        mergeL0[i] = __shfl_up_sync(0xffffffffu, (int)skipR0[i], 1) != 0 && lane > 0;
        mergeL1[i] = __shfl_up_sync(0xffffffffu, (int)skipR1[i], 1) != 0 && lane > 0;
      }
      const bool first = tc0 == 0;
      // stage `ch` -> ring slot ch % DEPTH
#define WB_GB_ISSUE(ch_)                                                                          \
      do {                                                                                          \
        const unsigned off_ = (unsigned)(ch_) * HWd;                                                \
        float* base_ = my + ((ch_) % DEPTH) * NF * WB_TILE_PX;                                      \
        WB_UNROLL for (int i = 0; i < TG; ++i) {                                                    \
          const float* p0_ = src[i] + off_ + o0[i];                                                 \
          const float* p1_ = src[i] + off_ + o1[i];                                                 \
          wb_cp4(base_ + (5 * i + 0) * WB_TILE_PX, p0_); wb_cp4(base_ + (5 * i + 1) * WB_TILE_PX, p0_ + 1); \
          wb_cp4(base_ + (5 * i + 2) * WB_TILE_PX, p1_); wb_cp4(base_ + (5 * i + 3) * WB_TILE_PX, p1_ + 1); \
          wb_cp4(base_ + (5 * i + 4) * WB_TILE_PX, drw[i] + off_);                                  \
        }                                                                                           \
        wb_cp4(base_ + (5 * TG) * WB_TILE_PX, dof + off_);                                          \
        if (first) wb_cp4(base_ + (5 * TG + 1) * WB_TILE_PX, of + off_);                            \
      } while (0)
      WB_UNROLL for (int ch = 0; ch < DEPTH - 1; ++ch) { if (ch < C) WB_GB_ISSUE(ch); wb_cp_commit(); }
      unsigned choff = 0u;
#pragma unroll 1
      for (int ch = 0; ch < C; ++ch) {
        if (ch + DEPTH - 1 < C) WB_GB_ISSUE(ch + DEPTH - 1);
        wb_cp_commit();
        wb_cp_wait<DEPTH - 1>();
        const float* base = my + (ch % DEPTH) * NF * WB_TILE_PX;
        const float gO = actf * base[(5 * TG) * WB_TILE_PX];
        if (first) S += gO * base[(5 * TG + 1) * WB_TILE_PX];
        WB_UNROLL for (int i = 0; i < TG; ++i) {
          const float v0 = base[(5 * i + 0) * WB_TILE_PX], v1 = base[(5 * i + 1) * WB_TILE_PX];
          const float v2 = base[(5 * i + 2) * WB_TILE_PX], v3 = base[(5 * i + 3) * WB_TILE_PX];
          const float gdt = actf * base[(5 * i + 4) * WB_TILE_PX];
          const float go = gdt + nrm[i] * gO;
          U[i][0] += gO * v0; U[i][1] += gO * v1; U[i][2] += gO * v2; U[i][3] += gO * v3;
          Tq[i][0] += gdt * v0; Tq[i][1] += gdt * v1; Tq[i][2] += gdt * v2; Tq[i][3] += gdt * v3;
          float* dl = dsr[i] + choff;
          const float c1 = w[i][1] * go, c3 = w[i][3] * go;
          const float r1 = __shfl_up_sync(0xffffffffu, c1, 1), r3 = __shfl_up_sync(0xffffffffu, c3, 1);
          WB_RED(dl + o0[i], w[i][0] * go + (mergeL0[i] ? r1 : 0.f));
          if (!skipR0[i]) WB_RED(dl + o0[i] + 1, c1);
          WB_RED(dl + o1[i], w[i][2] * go + (mergeL1[i] ? r3 : 0.f));
          if (!skipR1[i]) WB_RED(dl + o1[i] + 1, c3);
        }
        choff += HWd;
      }
#undef WB_GB_ISSUE
      wb_cp_wait<0>();
      if (!a.glue || !active) continue;
      const float gOs = dra ? __ldg(dra) : 0.f;
      if (first && dra) S += gOs * __ldg(of + choff);
      WB_UNROLL for (int i = 0; i < TG; ++i) {
        const size_t pair = ((size_t)b * g.Tc + tc0 + i) * g.Tp + tp;
        const float* fl = d.flow + pair * 2 * HWd + q;
        const WbTaps t = wb_taps(__fadd_rn(gx, __ldg(fl)), __fadd_rn(gy, __ldg(fl + HWd)), g.Wd, g.Hd);
        const WbTap2 t2 = wb_tap2(t, g.Wd, g.Hd);
        float cx[4], cy[4];
        wb_pos4(t2, -t.wy0, t.wy0, -t.wy1, t.wy1, cx);
        wb_pos4(t2, -t.wx0, -t.wx1, t.wx0, t.wx1, cy);
        const float sc = nrm[i] * D - 1e-6f;
        float G = U[i][0] * w[i][0] + U[i][1] * w[i][1] + U[i][2] * w[i][2] + U[i][3] * w[i][3];
        float gix = 0.f, giy = 0.f;
        WB_UNROLL for (int j = 0; j < 4; ++j) {
          const float tj = Tq[i][j] + nrm[i] * U[i][j];
          gix += tj * cx[j]; giy += tj * cy[j];
        }
        G += gOs * (sc * 2.f - 1.f);
        const float* dfl = a.d_flow ? a.d_flow + pair * 2 * HWd + q : nullptr;
        float* gl = a.glue + pair * 3 * HWd + q;
        gl[0] = 2.f * nrm[i] * gOs + (G - S) / D;
        gl[HWd] = (dfl ? __ldg(dfl) : 0.f) + gix * (0.5f * (float)g.Wd);
        gl[2 * HWd] = (dfl ? __ldg(dfl + HWd) : 0.f) + giy * (0.5f * (float)g.Hd);
      }
    }
  }
}
#endif  // !WB_HOST_EMU

// grid = (red_ctas, B*Tp), block = 256 (one 32x8 pixel tile per iteration), rolled loop over the contexts.
__global__ void __launch_bounds__(WB_TILE_PX, WB_OCC_LAYERS_BWD) k_layers_bwd(WbDecB a) {
  const WbDec& d = a.f;
  const waldo_geom_t g = d.g;
  WbBwdCtx c;
  c.L = g.No + 1; c.HW = g.H * g.W; c.C = g.C; c.HWd = (unsigned)(g.Hd * g.Wd);
  const int btp = blockIdx.y;
  c.b = btp / g.Tp; c.tp = btp - c.b * g.Tp;
  const int u = (int)d.pred_ts[c.tp];
  c.self = (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
  c.disocc_ch = (g.flags & WALDO_F_USE_DISOCC) != 0;
  c.TcR = g.Tc + (c.self ? 1 : 0); c.CR = c.C + c.L + (c.disocc_ch ? 1 : 0);
  c.need_layers = true;
  c.lowres_direct = (g.Hd == g.H);
  c.pairs_only = (g.flags & WALDO_F_OCC_PAIRS) != 0;
  const int L = c.L, b = c.b, tp = c.tp;
  const unsigned HWd = c.HWd;
  __shared__ float s_occ[WB_MAX_L * WB_MAX_L];
  __shared__ float s_red[WB_NWARP][WB_MAX_L * WB_MAX_L];
  __shared__ float s_stage[WB_NWARP][2 * WB_STAGE_SLOTS * WB_STAGE_ROW];
  for (int i = wb_tid(); i < L * L; i += wb_nthr()) s_occ[i] = __ldg(d.occ + ((size_t)b * g.T + u) * L * L + i);
  for (int i = wb_tid(); i < WB_NWARP * WB_MAX_L * WB_MAX_L; i += wb_nthr()) (&s_red[0][0])[i] = 0.f;
  for (int i = wb_tid(); i < WB_NWARP * 2 * WB_STAGE_SLOTS * WB_STAGE_ROW; i += wb_nthr()) (&s_stage[0][0])[i] = 0.f;
  __syncthreads();
  c.s_occ = s_occ; c.s_stage = s_stage[wb_warp()];
  c.s_acc = a.d_occ ? s_red[wb_warp()] : nullptr;
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    for (int it = wb_tid(); it < WB_TILE_PX; it += wb_nthr()) {
      const int Xr = tx0 + (it & (WB_TILE_W - 1)), Yr = ty0 + it / WB_TILE_W;
      const float actf = (Xr < g.Wd && Yr < g.Hd) ? 1.f : 0.f;
      const int X = min(Xr, g.Wd - 1), Y = min(Yr, g.Hd - 1);
      const unsigned q = (unsigned)(Y * g.Wd + X);
      WbPix px = wb_pix(d, b, tp, X, Y);
      const unsigned wm = wb_warp_or(px.isobj);
      const int n = __popc(wm);
      WbColRed cr;
      if (a.d_f_lo && !c.lowres_direct)
        cr = wb_colred_setup(a.up_tab, wb_shfl(X, 0), wb_shfl(px.ax.i0, 0), wb_shfl(px.ax.i1, WB_WARP - 1), px.ay);
#if !defined(WB_HOST_EMU) && !defined(WB_NO_LANES)
      if (n <= WB_LANES_MAX_BWD) {
        const bool rowact = Yr < g.Hd;
        const unsigned iso = px.isobj;
        if (n == 1) wb_lanes_layers_bwd<1>(a, c, cr, wm, n, iso, tx0, rowact, Y, px.ay, px.gy);
        else if (n == 2) wb_lanes_layers_bwd<2>(a, c, cr, wm, n, iso, tx0, rowact, Y, px.ay, px.gy);
        else if (n <= 4) wb_lanes_layers_bwd<4>(a, c, cr, wm, n, iso, tx0, rowact, Y, px.ay, px.gy);
#if WB_LANES_MAX_BWD > 4
        else wb_lanes_layers_bwd<8>(a, c, cr, wm, n, iso, tx0, rowact, Y, px.ay, px.gy);
#endif
        continue;
      }
#endif
      if (WB_NA_VARIANTS_BWD >= 2 && n <= 4) wb_bwd_layers_ctxs<4>(a, c, px, cr, wm, q, actf);
      else if (WB_NA_VARIANTS_BWD >= 3 && n <= 8) wb_bwd_layers_ctxs<8>(a, c, px, cr, wm, q, actf);
      else wb_bwd_layers_ctxs<WB_MAX_L>(a, c, px, cr, wm, q, actf);
    }
  }
  if (a.d_occ) {
    __syncthreads();
    float* part = a.occ_part + ((size_t)btp * gridDim.x + blockIdx.x) * L * L;
    for (int e = wb_tid(); e < L * L; e += wb_nthr()) {
      float acc = 0.f;
      for (int w = 0; w < WB_NWARP; ++w) acc += s_red[w][e];
      part[e] = acc;
    }
  }
}

// d_occ[b, frame(group), :] += sum over CTAs (in order) of the partials.  mode 0: group = (b,tp) -> frame pred_ts[tp];
// mode 1: group = (b,t) -> frame t.
// grid = (groups, slices of the L*L entries): the serial loop over the CTA partials is the critical path, so the entries
// are spread over many CTAs instead of one CTA per group.
__global__ void k_occ_reduce(const float* __restrict__ part, int groups, int per_b, int ctas, int LL, int T,
                             const int64_t* __restrict__ pred_ts, int mode, float* __restrict__ d_occ) {
  const int grp = blockIdx.x, b = grp / per_b, j = grp - b * per_b;
  const int frame = mode == 0 ? (int)pred_ts[j] : j;
  for (int e = blockIdx.y * wb_nthr() + wb_tid(); e < LL; e += gridDim.y * wb_nthr()) {
    float acc = 0.f;
    WB_UNROLL_N(8) for (int c = 0; c < ctas; ++c) acc += __ldg(part + ((size_t)grp * ctas + c) * LL + e);   // fixed order
    d_occ[((size_t)b * T + frame) * LL + e] += acc;
  }
}

// ============================================================================ context-alpha backward (B4..B2b)
struct WbPrepBwdCtx {
  int b, t, L, Nl, HW;
  unsigned HWd;
  bool filt, lowres_direct, need_p, pairs_only;
  const float *s_P, *s_occ, *lyt_base, *alo;
  float *s_acc, *s_accp;
  float* s_stage;    // this warp's staging area, WB_STAGE_SLOTS * WB_WARP floats
};

// Sum over the warp of 32 per-lane values v[0..31]: afterwards lane l holds sum_lanes v[l] in v[0] (a fixed butterfly:
// 31 shuffles instead of 32 x 5, deterministic).  The emulation build (one lane) leaves v untouched.
WB_DEV void wb_warp_transpose_sum(float* v) {
#ifndef WB_HOST_EMU
  const int lane = wb_lane();
  WB_UNROLL for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
    WB_UNROLL for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#endif
}

template <int NA, int NLC>
WB_DEV void wb_prep_bwd_pixel(const WbDecB& a, const WbPrepBwdCtx& c, const WbColRed& cr, unsigned wm, float actf, unsigned q,
                              const WbAxis& ax, const WbAxis& ay, int o00, int o01, int o10, int o11) {
  constexpr int NN = NLC > 0 ? NLC : WB_MAX_NL;
  const WbDec& d = a.f;
  const waldo_geom_t& g = d.g;
  const int L = c.L, Nl = NLC > 0 ? NLC : c.Nl, HW = c.HW, b = c.b, t = c.t;
  const unsigned HWd = c.HWd;
  const int lane = wb_lane();
  const WbIdx<NA> ix = wb_idx<NA>(wm);
  const bool any_obj = (wm >> 1) != 0u;
  // the scatter-accumulated d/dA of the first four slots is requested up front (most rows have <= 4 live layers): four
  // independent HBM loads in flight while the forward is recomputed, instead of one exposed load per trip of the rolled loop
  float pf[4] = {0.f, 0.f, 0.f, 0.f};
  if (NA > 8 && a.d_alpha_acc) {
    const float* base = a.d_alpha_acc + ((size_t)b * g.Tw + t) * L * HWd + q;
    WB_UNROLL for (int u = 0; u < 4; ++u) if (u < ix.n) pf[u] = base[(size_t)ix.k[u] * HWd];
  }
  // ---- recompute the forward of this pixel
  float sm[NN];
  if (c.filt && any_obj) wb_softmax_hd<NLC>(c.lyt_base, HWd, q, Nl, sm);
  float aup[NA], av[NA], ell[NA];
  WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s) {
    av[s] = 0.f; aup[s] = 0.f; ell[s] = 1.f;
    if (s < ix.n) {
      const int k = ix.k[s];
      const float* pl = c.alo + (size_t)k * HW;
      float v = c.lowres_direct ? __ldg(pl + o00)
                                : wb_lerp2(__ldg(pl + o00), __ldg(pl + o01), __ldg(pl + o10), __ldg(pl + o11), ax, ay);
      aup[s] = v;
      if (c.filt && k >= 1) {
        const float* P = c.s_P + (k - 1) * Nl;
        float dist = 0.f;
        WB_UNROLL for (int cc = 0; cc < NN; ++cc) if (NLC > 0 || cc < Nl) dist += fabsf(P[cc] - sm[cc]);
        ell[s] = 1.f - dist * 0.5f;
      }
      av[s] = v * ell[s];
    }
  }
  // ---- upstream: scatter-accumulated d/dA plus the returned alpha = 2A - 1 (zero beyond the image edge)
  float gA[NA], ga[NA];
  WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s) {
    ga[s] = 0.f; gA[s] = 0.f;
    if (s < ix.n) {
      const size_t o = (((size_t)b * g.Tw + t) * L + ix.k[s]) * HWd + q;
      float v = 0.f;
      if (a.d_alpha_acc) {
        if (NA > 8 && s < 4) v += s == 0 ? pf[0] : (s == 1 ? pf[1] : (s == 2 ? pf[2] : pf[3]));
        else v += a.d_alpha_acc[o];
      }
      if (a.d_alpha) v += 2.f * __ldg(a.d_alpha + o);
      gA[s] = v * actf;
    }
  }
  if (NA > 8 && ix.n <= 4) {
    // most rows have <= 4 live layers: run the O(n^2) exclusive-product backward on registers, fully unrolled, instead of
    // the rolled double loop over local-memory arrays
    WbIdx<4> ix4;
    float R4[4], gA4[4], ga4[4];
    ix4.n = ix.n;
    WB_UNROLL for (int u = 0; u < 4; ++u) {
      const bool on = u < ix.n;
      ix4.k[u] = on ? ix.k[u] : 0; R4[u] = on ? av[u] : 0.f; gA4[u] = on ? gA[u] : 0.f; ga4[u] = 0.f;
    }
    wb_occlude_bwd<4>(R4, gA4, c.s_occ, L, ix4, ga4, c.s_acc, c.pairs_only);
    WB_UNROLL for (int u = 0; u < 4; ++u) if (u < ix.n) ga[u] = ga4[u];
  }
#if WB_PREP_OCC8
  else if (NA > 8 && ix.n <= 8) {   // 5..8 live layers (30 % of the rows at the benchmark shape): same on 8 register slots
    WbIdx<8> ix8;
    float R8[8], gA8[8], ga8[8];
    ix8.n = ix.n;
    WB_UNROLL for (int u = 0; u < 8; ++u) {
      const bool on = u < ix.n;
      ix8.k[u] = on ? ix.k[u] : 0; R8[u] = on ? av[u] : 0.f; gA8[u] = on ? gA[u] : 0.f; ga8[u] = 0.f;
    }
    wb_occlude_bwd<8>(R8, gA8, c.s_occ, L, ix8, ga8, c.s_acc, c.pairs_only);
    WB_UNROLL for (int u = 0; u < 8; ++u) if (u < ix.n) ga[u] = ga8[u];
  }
#endif
  else {
    wb_occlude_bwd<NA>(av, gA, c.s_occ, L, ix, ga, c.s_acc, c.pairs_only);
  }
  // ---- filter backward: l_k = 1 - 0.5 sum_c |P_kc - sm_c|.  One ROLLED loop over the object slots (a single copy of
  // the class loop in the code); d P_kc = sum over pixels of -0.5 sign(P_kc - sm_c) d l_k is reduced over the warp with
  // a transpose butterfly and accumulated by lane c.
  float gsm[NN];
  WB_UNROLL for (int cc = 0; cc < NN; ++cc) gsm[cc] = 0.f;
  if (c.filt && any_obj) {
    for (int s = 0; s < ix.n; ++s) {
      float gl = 0.f;
      int k = 0;
      if (NA > 8) { gl = ga[s] * aup[s]; k = ix.k[s]; }   // rolled form: the arrays live in local memory, index them directly
      else { WB_UNROLL_NA for (int ss = 0; ss < WB_NEND; ++ss) if (ss == s) { gl = ga[ss] * aup[ss]; k = ix.k[ss]; } }   // d / d l_k
      if (k < 1) continue;   // warp-uniform
      const float* P = c.s_P + (k - 1) * Nl;
      float v[32];
      WB_UNROLL for (int cc = 0; cc < 32; ++cc) {
        v[cc] = 0.f;
        if (cc < NN && (NLC > 0 || cc < Nl)) {
          const float df = P[cc] - sm[cc];
          const float sg = df > 0.f ? 0.5f : (df < 0.f ? -0.5f : 0.f);
          gsm[cc] += sg * gl;
          v[cc] = -sg * gl;
        }
      }
      if (c.need_p) {
        wb_warp_transpose_sum(v);
#ifdef WB_HOST_EMU
        for (int cc = 0; cc < Nl; ++cc) c.s_accp[(k - 1) * Nl + cc] += v[cc];
#else
        if (lane < Nl) c.s_accp[(k - 1) * Nl + lane] += v[0];
#endif
      }
    }
  }
  // ---- up-sampling backward: d a_lo
  if (a.d_a_lo) {
    if (c.lowres_direct) {
      WB_UNROLL_NA for (int s = 0; s < WB_NEND; ++s)
        if (s < ix.n) WB_RED_NZ(a.d_a_lo + (((size_t)b * g.Tw + t) * L + ix.k[s]) * HW + o00, ga[s] * ell[s]);
    } else if constexpr (NA <= WB_STAGE_SLOTS) {
      float* dst[NA];
      WB_UNROLL for (int s = 0; s < NA; ++s) {
        c.s_stage[WB_STAGE_AT(s, lane)] = ga[s] * ell[s];
        dst[s] = a.d_a_lo + (((size_t)b * g.Tw + t) * L + ix.k[s]) * HW;
      }
      __syncwarp();
      wb_colred_flush(a, cr, c.s_stage, ix.n, dst, g.W, 1);
      __syncwarp();
    }
#ifndef WB_HOST_EMU
    else if (ix.n <= WB_STAGE_SLOTS) {   // the usual case: all live slots staged at once, destinations derived from the mask
      for (int s = 0; s < ix.n; ++s) c.s_stage[WB_STAGE_AT(s, lane)] = ga[s] * ell[s];
      __syncwarp();
      wb_colred_flush_slots(a, cr, c.s_stage, wm, 1, a.d_a_lo + ((size_t)b * g.Tw + t) * L * HW, (size_t)HW, g.W, 1);
      __syncwarp();
    }
#endif
    else {
      for (int s0 = 0; s0 < ix.n; s0 += WB_STAGE_SLOTS) {
        float* dst[WB_STAGE_SLOTS];
        const int ns = min(WB_STAGE_SLOTS, ix.n - s0);
        for (int j = 0; j < ns; ++j) {
          c.s_stage[WB_STAGE_AT(j, lane)] = ga[s0 + j] * ell[s0 + j];
          dst[j] = a.d_a_lo + (((size_t)b * g.Tw + t) * L + ix.k[s0 + j]) * HW;
        }
        __syncwarp();
        wb_colred_flush(a, cr, c.s_stage, ns, dst, g.W, 1);
        __syncwarp();
      }
    }
  }
  if (c.filt && any_obj && a.d_input && actf != 0.f) {   // softmax backward into the layout logits of this frame
    float dot = 0.f;
    WB_UNROLL for (int cc = 0; cc < NN; ++cc) if (NLC > 0 || cc < Nl) dot += gsm[cc] * sm[cc];
    float* o = a.d_input + (((size_t)b * g.T + t) * g.C + 3) * HWd + q;
    WB_UNROLL for (int cc = 0; cc < NN; ++cc)
      if (NLC > 0 || cc < Nl) { WB_RED(o, sm[cc] * (gsm[cc] - dot)); o += HWd; }   // fire-and-forget reduction
  }
}

#ifndef WB_HOST_EMU
// ============================================================================ lanes-per-layer form of the context-alpha backward
// NOT the default (WB_LANES_PREP_BWD = 0): parity-green, but measured SLOWER on B200 than the one-lane-per-pixel form
// (4.2 ms vs 3.2 ms per launch, profiles/r1_v10): unlike the layer kernels this one is dominated by the per-object loops
// over the 20 classes, and with one layer per lane the background / padding lanes of every pixel idle through them.
// Same layout as wb_lanes_layers_bwd: LP lanes per pixel, one (pixel, layer) per lane, LP passes over the row.
// Class-indexed quantities (20 classes padded to 32) are spread over lanes with partial transpose-sums:

// over the SLOT lanes of a pixel: afterwards v[0 .. 32/LP) of slot lane s holds the sums (over the LP lanes) of the classes
// cbase + i, with cbase returned  (= sum over stages t of bit_t(s) * (16 >> t))
template <int LP>
WB_DEV int wb_class_split_slots(float* v, int slot) {
  constexpr int PPW = 32 / LP;
  int cbase = 0;
  WB_UNROLL for (int t = 0; (1 << t) < LP; ++t) {
    const int o = 16 >> t, bit = 1 << t;
    const bool up = (slot & bit) != 0;
    WB_UNROLL for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit * PPW);
    }
    if (up) cbase += o;
  }
  return cbase;
}
// over the PPW pixel lanes that share a slot: v[0 .. 32/PPW) of pixel lane pl holds the row sums of the classes cbase + i
template <int PPW>
WB_DEV int wb_class_split_pixels(float* v, int pl) {
  int cbase = 0;
  WB_UNROLL for (int t = 0; (1 << t) < PPW; ++t) {
    const int o = 16 >> t, bit = 1 << t;
    const bool up = (pl & bit) != 0;
    WB_UNROLL for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
    if (up) cbase += o;
  }
  return cbase;
}

#define WB_SM_ROW 32   // s_sm[class][pixel of the row]

template <int LP, int NLC>
WB_DEV void wb_lanes_prep_bwd(const WbDecB& a, const WbPrepBwdCtx& c, const WbColRed& cr, unsigned wm, int n, int tx0, bool rowact, int Y,
                              const WbAxis& ay, float* __restrict__ s_sm) {
  constexpr int PPW = 32 / LP;
  constexpr int NN = NLC > 0 ? NLC : WB_MAX_NL;
  constexpr int NS = 32 / LP;     // classes per slot lane after the split
  constexpr int NPX = 32 / PPW;   // classes per pixel lane after the row split (= LP)
  const WbDec& d = a.f;
  const waldo_geom_t& g = d.g;
  const int L = c.L, Nl = NLC > 0 ? NLC : c.Nl, HW = c.HW, b = c.b, t = c.t;
  const unsigned HWd = c.HWd;
  const int lane = wb_lane(), pl = lane % PPW, slot = lane / PPW;
  const bool valid = slot < n;
  const int k = valid ? wb_nth_bit(wm, slot) : 0;
  const bool isobjk = valid && k >= 1;
  float oc[LP], accj[LP];
  WB_UNROLL for (int j = 0; j < LP; ++j) {
    accj[j] = 0.f;
    oc[j] = (valid && j < n) ? c.s_occ[wb_nth_bit(wm, j) * L + k] : 0.f;
  }
  const bool filt_row = LP > 1 && c.filt && (wm >> 1) != 0u;   // some object is live in this row
  // ---- phase A (lane = pixel): softmax of the HD layout logits, staged for the slot lanes
  if (filt_row) {
    const int X = min(tx0 + lane, g.Wd - 1);
    float sm[NN];
    wb_softmax_hd<NLC>(c.lyt_base, HWd, (unsigned)(Y * g.Wd + X), Nl, sm);
    WB_UNROLL for (int cc = 0; cc < NN; ++cc) if (NLC > 0 || cc < Nl) s_sm[cc * WB_SM_ROW + lane] = sm[cc];
    __syncwarp();
  }
  float accP[NN];   // d P[k, :] of this lane's layer, summed over the passes
  WB_UNROLL for (int cc = 0; cc < NN; ++cc) accP[cc] = 0.f;
  const float* P = c.s_P + (isobjk ? (k - 1) * Nl : 0);
  const float* alo_k = c.alo + (size_t)k * HW;
  const size_t plane = (((size_t)b * g.Tw + t) * L + k) * HWd;
  const float r_lo = (float)g.H / (float)g.Hd;
  const bool stage = a.d_a_lo && !c.lowres_direct;
#pragma unroll 1
  for (int r = 0; r < LP; ++r) {
    const int p = r * PPW + pl, Xr = tx0 + p, X = min(Xr, g.Wd - 1);
    const float actf = (rowact && Xr < g.Wd) ? 1.f : 0.f;
    const unsigned q = (unsigned)(Y * g.Wd + X);
    const WbAxis ax = wb_axis(X, r_lo, g.W);
    const int o00 = ay.i0 * g.W + ax.i0, o01 = ay.i0 * g.W + ax.i1, o10 = ay.i1 * g.W + ax.i0, o11 = ay.i1 * g.W + ax.i1;
    // ---- loads first (invalid lanes read layer 0: harmless)
    float a00 = __ldg(alo_k + o00), a01 = a00, a10 = a00, a11 = a00;
    if (!c.lowres_direct) { a01 = __ldg(alo_k + o01); a10 = __ldg(alo_k + o10); a11 = __ldg(alo_k + o11); }
    const float gacc = a.d_alpha_acc ? a.d_alpha_acc[plane + q] : 0.f;
    const float gal = a.d_alpha ? __ldg(a.d_alpha + plane + q) : 0.f;
    // ---- forward of this (pixel, layer), same arithmetic as wb_prep_pixel
    const float aup = valid ? (c.lowres_direct ? a00 : wb_lerp2(a00, a01, a10, a11, ax, ay)) : 0.f;
    float ell = 1.f;
    if (filt_row && isobjk) {
      float dist = 0.f;
      WB_UNROLL for (int cc = 0; cc < NN; ++cc) if (NLC > 0 || cc < Nl) dist += fabsf(P[cc] - s_sm[cc * WB_SM_ROW + p]);
      ell = 1.f - dist * 0.5f;
    }
    const float av = aup * ell;
    const float gA = valid ? (gacc + 2.f * gal) * actf : 0.f;
    // ---- exclusive-product backward over the slot lanes
    float Rj[LP], pre[LP], term[LP];
    float run = 1.f;
    WB_UNROLL for (int j = 0; j < LP; ++j) {
      Rj[j] = __shfl_sync(0xffffffffu, av, pl + j * PPW);
      pre[j] = run;
      run *= 1.f - Rj[j] * oc[j];
    }
    float ga = gA * run;
    const float gV = gA * av;
    float suf = 1.f;
    WB_UNROLL for (int j = LP - 1; j >= 0; --j) {
      const float excl = pre[j] * suf;
      suf *= 1.f - Rj[j] * oc[j];
      term[j] = -gV * oc[j] * excl;
      accj[j] += -gV * Rj[j] * excl;
    }
    wb_slot_transpose_sum<LP>(term, slot);
    ga += term[0];
    // ---- filter backward: l_k = 1 - 0.5 sum_c |P_kc - sm_c|
    if (filt_row) {
      const float gl = isobjk ? ga * aup : 0.f;   // d / d l_k
      float v[32];
      WB_UNROLL for (int cc = 0; cc < 32; ++cc) {
        v[cc] = 0.f;
        if (cc < NN && (NLC > 0 || cc < Nl)) {
          const float df = P[cc] - s_sm[cc * WB_SM_ROW + p];
          const float sg = df > 0.f ? 0.5f : (df < 0.f ? -0.5f : 0.f);
          v[cc] = sg * gl;          // this layer's share of d / d sm_c
          accP[cc] -= sg * gl;      // d / d P_kc
        }
      }
      if (a.d_input) {   // softmax backward into the layout logits: d lyt_c = sm_c (gsm_c - sum_c' gsm_c' sm_c')
        const int cbase = wb_class_split_slots<LP>(v, slot);
        float smc[NS];
        float dot = 0.f;
        WB_UNROLL for (int i = 0; i < NS; ++i) {
          const int cls = cbase + i;
          smc[i] = cls < Nl ? s_sm[cls * WB_SM_ROW + p] : 0.f;
          dot += v[i] * smc[i];
        }
        WB_UNROLL for (int o = PPW; o < 32; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (actf != 0.f) {
          float* o = a.d_input + (((size_t)b * g.T + t) * g.C + 3 + cbase) * HWd + q;
          WB_UNROLL for (int i = 0; i < NS; ++i) {
            if (cbase + i < Nl) WB_RED(o, smc[i] * (v[i] - dot));   // fire-and-forget reduction
            o += HWd;
          }
        }
      }
    }
    // ---- up-sampling backward: d a_lo
    if (a.d_a_lo && valid) {
      if (c.lowres_direct) WB_RED_NZ(a.d_a_lo + (((size_t)b * g.Tw + t) * L + k) * HW + o00, ga * ell);
      else c.s_stage[WB_STAGE_AT(slot, p)] = ga * ell;
    }
  }
  if (stage) {
    __syncwarp();
    wb_colred_flush_slots(a, cr, c.s_stage, wm, 1, a.d_a_lo + ((size_t)b * g.Tw + t) * L * HW, (size_t)HW, g.W, 1);
    __syncwarp();
  }
  if (c.s_acc) {
    WB_UNROLL for (int j = 0; j < LP; ++j) {
      float v = accj[j];
      WB_UNROLL for (int o = PPW / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (pl == 0 && valid && j < n) {
        const int kj = wb_nth_bit(wm, j);
        if (!(c.pairs_only && (k == 0 || kj == 0 || k == kj))) c.s_acc[kj * L + k] += v;
      }
    }
  }
  if (filt_row) {
    if (c.need_p) {   // d P[k, :]: row sums over the pixel lanes, one lane per (layer, class)
      float v[32];
      WB_UNROLL for (int cc = 0; cc < 32; ++cc) v[cc] = cc < NN ? accP[cc] : 0.f;
      const int cbase = wb_class_split_pixels<PPW>(v, pl);
      WB_UNROLL for (int i = 0; i < NPX; ++i)
        if (isobjk && cbase + i < Nl) c.s_accp[(k - 1) * Nl + cbase + i] += v[i];
    }
    __syncwarp();   // s_sm is rewritten by the warp's next row
  }
}
#endif  // !WB_HOST_EMU

// grid = (red_ctas, B*Tw), block = 256.
template <int NLC>
__global__ void __launch_bounds__(WB_TILE_PX, WB_OCC_PREP_BWD) k_alpha_prep_bwd(WbDecB a) {
  const WbDec& d = a.f;
  const waldo_geom_t g = d.g;
  WbPrepBwdCtx c;
  const int No = g.No, Nl = g.Nl, L = No + 1;
  c.L = L; c.Nl = Nl; c.HW = g.H * g.W; c.HWd = (unsigned)(g.Hd * g.Wd);
  const int bt = blockIdx.y;
  c.b = bt / g.Tw; c.t = bt - c.b * g.Tw;
  c.filt = (g.flags & WALDO_F_FILTER) != 0;
  c.lowres_direct = (g.Hd == g.H);
  c.need_p = c.filt && a.d_prof_p;
  c.pairs_only = (g.flags & WALDO_F_OCC_PAIRS) != 0;
  __shared__ float s_P[(WB_MAX_L - 1) * WB_MAX_NL];
  __shared__ float s_occ[WB_MAX_L * WB_MAX_L];
  __shared__ float s_red[WB_NWARP][WB_MAX_L * WB_MAX_L];
  __shared__ float s_redp[WB_NWARP][(WB_MAX_L - 1) * WB_MAX_NL];
  __shared__ float s_stage[WB_NWARP][WB_STAGE_SLOTS * WB_STAGE_ROW];
  WB_DYN_SMEM(s_dyn);   // WB_NWARP x [WB_MAX_NL][32] softmax staging of the lanes-per-layer path
  if (c.filt) for (int i = wb_tid(); i < No * Nl; i += wb_nthr()) s_P[i] = d.prof_p[(size_t)c.b * No * Nl + i];
  for (int i = wb_tid(); i < L * L; i += wb_nthr()) s_occ[i] = __ldg(d.occ + ((size_t)c.b * g.T + c.t) * L * L + i);
  for (int i = wb_tid(); i < WB_NWARP * WB_MAX_L * WB_MAX_L; i += wb_nthr()) (&s_red[0][0])[i] = 0.f;
  for (int i = wb_tid(); i < WB_NWARP * (WB_MAX_L - 1) * WB_MAX_NL; i += wb_nthr()) (&s_redp[0][0])[i] = 0.f;
  for (int i = wb_tid(); i < WB_NWARP * WB_STAGE_SLOTS * WB_STAGE_ROW; i += wb_nthr()) (&s_stage[0][0])[i] = 0.f;
  __syncthreads();
  c.s_P = s_P; c.s_occ = s_occ; c.s_stage = s_stage[wb_warp()];
  c.s_acc = a.d_occ ? s_red[wb_warp()] : nullptr;
  c.s_accp = s_redp[wb_warp()];
  const float rlo = (float)g.H / (float)g.Hd;
  c.lyt_base = d.input + (((size_t)c.b * g.T + c.t) * g.C + 3) * c.HWd;
  c.alo = d.a_lo + ((size_t)c.b * g.Tw + c.t) * L * c.HW;
  const uint32_t* live = d.live_ctx + ((size_t)c.b * g.Tw + c.t) * c.HW;
  const WbTileIter ti(g.Hd, g.Wd);
  for (int tile = blockIdx.x; tile < ti.ntiles; tile += gridDim.x) {
    const int ty0 = (tile / ti.tiles_x) * WB_TILE_H, tx0 = (tile % ti.tiles_x) * WB_TILE_W;
    for (int it = wb_tid(); it < WB_TILE_PX; it += wb_nthr()) {
      const int Xr = tx0 + (it & (WB_TILE_W - 1)), Yr = ty0 + it / WB_TILE_W;
      const float actf = (Xr < g.Wd && Yr < g.Hd) ? 1.f : 0.f;   // beyond the edge: nearest valid pixel, zero upstream
      const int X = min(Xr, g.Wd - 1), Y = min(Yr, g.Hd - 1);
      const unsigned q = (unsigned)(Y * g.Wd + X);
      WbAxis ay = wb_axis(Y, rlo, g.H), ax = wb_axis(X, rlo, g.W);
      const int o00 = ay.i0 * g.W + ax.i0, o01 = ay.i0 * g.W + ax.i1, o10 = ay.i1 * g.W + ax.i0, o11 = ay.i1 * g.W + ax.i1;
      const unsigned wm = wb_warp_or(wb_live4(live, o00, o01, o10, o11));
      const int n = __popc(wm);
      if (n == 0) continue;   // warp-uniform: nothing live here, every gradient path is dead
      WbColRed cr;
      if (a.d_a_lo && !c.lowres_direct) cr = wb_colred_setup(a.up_tab, wb_shfl(X, 0), wb_shfl(ax.i0, 0), wb_shfl(ax.i1, WB_WARP - 1), ay);
#if !defined(WB_HOST_EMU) && !defined(WB_NO_LANES) && WB_LANES_PREP_BWD
      if (n <= 8) {
        float* s_sm = s_dyn + wb_warp() * (WB_MAX_NL * WB_SM_ROW);
        const bool rowact = Yr < g.Hd;
        if (n == 1) wb_lanes_prep_bwd<1, NLC>(a, c, cr, wm, n, tx0, rowact, Y, ay, s_sm);
        else if (n == 2) wb_lanes_prep_bwd<2, NLC>(a, c, cr, wm, n, tx0, rowact, Y, ay, s_sm);
        else if (n <= 4) wb_lanes_prep_bwd<4, NLC>(a, c, cr, wm, n, tx0, rowact, Y, ay, s_sm);
        else wb_lanes_prep_bwd<8, NLC>(a, c, cr, wm, n, tx0, rowact, Y, ay, s_sm);
        continue;
      }
#endif
      if (WB_NA_VARIANTS_BWD >= 2 && n <= 4) wb_prep_bwd_pixel<4, NLC>(a, c, cr, wm, actf, q, ax, ay, o00, o01, o10, o11);
      else if (WB_NA_VARIANTS_BWD >= 3 && n <= 8) wb_prep_bwd_pixel<8, NLC>(a, c, cr, wm, actf, q, ax, ay, o00, o01, o10, o11);
      else wb_prep_bwd_pixel<WB_MAX_L, NLC>(a, c, cr, wm, actf, q, ax, ay, o00, o01, o10, o11);
    }
  }
  __syncthreads();
  if (a.d_occ) {
    float* part = a.occ_part + ((size_t)bt * gridDim.x + blockIdx.x) * L * L;
    for (int e = wb_tid(); e < L * L; e += wb_nthr()) {
      float acc = 0.f;
      for (int w = 0; w < WB_NWARP; ++w) acc += s_red[w][e];
      part[e] = acc;
    }
  }
  if (c.need_p) {
    float* part = a.prof_p_part + ((size_t)bt * gridDim.x + blockIdx.x) * No * Nl;
    for (int e = wb_tid(); e < No * Nl; e += wb_nthr()) {
      float acc = 0.f;
      for (int w = 0; w < WB_NWARP; ++w) acc += s_redp[w][e];
      part[e] = acc;
    }
  }
}

// d_prof_p[b,:] = sum_t sum_cta partial (ordered); then through P = softmax(num/den) (or P = cls).
__global__ void __launch_bounds__(1024) k_profile_final_bwd(WbDecB a) {
  const WbDec& d = a.f;
  const waldo_geom_t g = d.g;
  const int No = g.No, Nl = g.Nl, nout = No * Nl + No;
  const int b = blockIdx.x;
  const bool from_cls = (g.flags & WALDO_F_HAS_CLS) && !(g.flags & WALDO_F_WEIGHT_CLS);
  // warp w sums slice w of the (t, cta) partials with its lanes on consecutive elements (coalesced, independent loads);
  // the slices are then added in slice order.  (One thread per element looping over all 592 partials was the whole
  // duration of this kernel.)
  {
    const int np = g.Tw * a.red_ctas, ne = No * Nl;
    const int nw = wb_nthr() / WB_WARP, w = wb_warp();
    const int per = (np + nw - 1) / nw, c0 = min(np, w * per), c1 = min(np, c0 + per);
    const float* pp = a.prof_p_part + (size_t)b * np * ne;
    __shared__ float s_slice[32][(WB_MAX_L - 1) * WB_MAX_NL];
    for (int e = wb_lane(); e < ne; e += WB_WARP) {
      float acc = 0.f;
      WB_UNROLL_N(4) for (int c = c0; c < c1; ++c) acc += __ldg(pp + (size_t)c * ne + e);
      s_slice[w][e] = acc;
    }
    __syncthreads();
    for (int e = wb_tid(); e < ne; e += wb_nthr()) {
      float acc = 0.f;
      for (int ww = 0; ww < nw; ++ww) acc += s_slice[ww][e];
      a.d_prof_p[(size_t)b * ne + e] = acc;
    }
  }
  __syncthreads();
  for (int k = wb_tid(); k < No; k += wb_nthr()) {
    const float* gP = a.d_prof_p + ((size_t)b * No + k) * Nl;
    if (from_cls) {
      if (a.d_cls) for (int c = 0; c < Nl; ++c) a.d_cls[((size_t)b * No + k) * Nl + c] += gP[c];
      continue;
    }
    const float* P = d.prof_p + ((size_t)b * No + k) * Nl;
    const float den = d.prof_sum[(size_t)b * nout + No * Nl + k];
    float dot = 0.f;
    for (int c = 0; c < Nl; ++c) dot += gP[c] * P[c];
    float gden = 0.f;
    for (int c = 0; c < Nl; ++c) {
      const float gM = P[c] * (gP[c] - dot);                  // softmax backward
      const float num = d.prof_sum[(size_t)b * nout + k * Nl + c];
      a.d_prof_sum[(size_t)b * nout + k * Nl + c] = gM / den; // M = num / den
      gden -= gM * num / (den * den);
    }
    a.d_prof_sum[(size_t)b * nout + No * Nl + k] = gden;
  }
}

// backward of the class-profile sums (lvd.py:735-742) on the low-res lattice; grid = (prof_ctas, B), block 256.
// d cls uses the same staged, ordered reduction as the forward.  One thread per low-res sample; the class count is a
// template parameter (20 / 19 / generic) and the object loop is unrolled over the compiled maximum, so that the four
// per-class vectors live in registers (with run-time trip counts they sat in local memory: 336 bytes of stack and half of
// the kernel's stall samples, profiles/r1_v37) and the rows of the two shared-memory tables are read as 128-bit words at
// compile-time offsets.  The per-object loads (a_lo, d a_lo) are issued before the arithmetic instead of one dependent
// load per trip.  Same operations in the same order as before: results are bit-identical.
#ifndef WB_OCC_PROF_BWD
#define WB_OCC_PROF_BWD 2   // 128 registers, 2 CTAs/SM: 1.05 ms vs 1.13 ms for the whole low-res tail at 1 CTA/SM (B200 A/B, r1_v42)
#endif
template <int NLC>
__global__ void __launch_bounds__(256, WB_OCC_PROF_BWD) k_class_profile_bwd(WbDecB a) {
  constexpr int NN = NLC > 0 ? NLC : WB_MAX_NL;
  constexpr int NO = WB_MAX_L - 1;
  const WbDec& d = a.f;
  const waldo_geom_t g = d.g;
  const int No = g.No, Nl = NLC > 0 ? NLC : g.Nl, HW = g.H * g.W, L = No + 1, nout = No * Nl + No;
  const size_t HWd = (size_t)g.Hd * g.Wd;
  const int b = blockIdx.y;
  const int nsamp = g.Tw * HW;
  const bool wcls = (g.flags & WALDO_F_WEIGHT_CLS) != 0;
  __shared__ float s_sm[WB_PROF_BATCH][WB_MAX_NL + 1];
  __shared__ float s_gq[WB_PROF_BATCH][WB_MAX_L];
  __shared__ __align__(16) float s_cls[(WB_MAX_L - 1) * WB_MAX_NL];
  __shared__ __align__(16) float s_gsum[(WB_MAX_L - 1) * WB_MAX_NL + WB_MAX_L];
  if (wcls)
    for (int i = wb_tid(); i < No * Nl; i += wb_nthr()) s_cls[i] = __ldg(d.cls + (size_t)b * No * Nl + i) + g.min_cls;
  for (int i = wb_tid(); i < nout; i += wb_nthr()) s_gsum[i] = a.d_prof_sum[(size_t)b * nout + i];
  float* part = (wcls && a.d_cls) ? a.cls_part + ((size_t)b * gridDim.x + blockIdx.x) * No * Nl : nullptr;
  if (part) for (int o = wb_tid(); o < No * Nl; o += wb_nthr()) part[o] = 0.f;
  __syncthreads();
  const float r = (float)g.Hd / (float)g.H;
  for (int s0 = blockIdx.x * WB_PROF_BATCH; s0 < nsamp; s0 += gridDim.x * WB_PROF_BATCH) {
    const int ns = min(WB_PROF_BATCH, nsamp - s0);
    for (int i = wb_tid(); i < ns; i += wb_nthr()) {
      const int s = s0 + i, t = s / HW, p = s - t * HW;
      const int y = p / g.W, x = p - y * g.W;
      const WbAxis ay = wb_axis(y, r, g.Hd), ax = wb_axis(x, r, g.Wd);
      const size_t hd00 = (size_t)ay.i0 * g.Wd + ax.i0, hd01 = (size_t)ay.i0 * g.Wd + ax.i1;
      const size_t hd10 = (size_t)ay.i1 * g.Wd + ax.i0, hd11 = (size_t)ay.i1 * g.Wd + ax.i1;
      // ---- every load whose address is known up front
      const size_t ia0 = (((size_t)b * g.Tw + t) * L + 1) * HW + p;   // object k: ia0 + k * HW
      float al[NO], dal[NO];
      WB_UNROLL for (int k = 0; k < NO; ++k) {
        al[k] = k < No ? __ldg(d.a_lo + ia0 + (size_t)k * HW) + 1e-6f : 0.f;
        dal[k] = (k < No && a.d_a_lo) ? a.d_a_lo[ia0 + (size_t)k * HW] : 0.f;
      }
      float lyt[NN], sm[NN], glyt[NN], gsmx[NN];
      if (d.lyt_lo) {
        const float* pl = d.lyt_lo + ((size_t)b * g.Tw + t) * Nl * HW + p;
        WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) lyt[c] = __ldg(pl + (size_t)c * HW);
      } else {   // same arithmetic as wb_lyt_lo
        const float* base = d.input + (((size_t)b * g.T + t) * g.C + 3) * HWd;
        WB_UNROLL for (int c = 0; c < NN; ++c)
          if (NLC > 0 || c < Nl) {
            const float* pl = base + c * HWd;
            lyt[c] = wb_lerp2(__ldg(pl + hd00), __ldg(pl + hd01), __ldg(pl + hd10), __ldg(pl + hd11), ax, ay);
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
      WB_UNROLL for (int c = 0; c < NN; ++c) { glyt[c] = 0.f; gsmx[c] = 0.f; }
      WB_UNROLL for (int k = 0; k < NO; ++k) {
        if (k < No) {
          const float* G = s_gsum + k * Nl;
          const float* Ck = s_cls + k * Nl;
          float qk = 1.f;
          if (wcls) { qk = 0.f; WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) qk += Ck[c] * sm[c]; }
          const float w = al[k] * qk;
          float gw = s_gsum[No * Nl + k];
          WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) { gw += G[c] * lyt[c]; glyt[c] += G[c] * w; }
          if (a.d_a_lo) a.d_a_lo[ia0 + (size_t)k * HW] = dal[k] + gw * qk;
          const float gq = gw * al[k];
          s_gq[i][k] = gq;
          if (wcls) { WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) gsmx[c] += gq * Ck[c]; }
        }
      }
      if (wcls) {
        float dot = 0.f;
        WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) dot += gsmx[c] * sm[c];
        WB_UNROLL for (int c = 0; c < NN; ++c) if (NLC > 0 || c < Nl) { glyt[c] += sm[c] * (gsmx[c] - dot); s_sm[i][c] = sm[c]; }
      }
      if (a.d_input) {   // transpose of the bilinear down-sampling
        float* base = a.d_input + (((size_t)b * g.T + t) * g.C + 3) * HWd;
        WB_UNROLL for (int c = 0; c < NN; ++c) {
          if (NLC > 0 || c < Nl) {
            float* pl = base + (size_t)c * HWd;
            const float gv = glyt[c];
            WB_RED_NZ(pl + hd00, gv * ax.l0 * ay.l0);
            WB_RED_NZ(pl + hd01, gv * ax.l1 * ay.l0);
            WB_RED_NZ(pl + hd10, gv * ax.l0 * ay.l1);
            WB_RED_NZ(pl + hd11, gv * ax.l1 * ay.l1);
          }
        }
      }
    }
    __syncthreads();
    if (part) {
      for (int o = wb_tid(); o < No * Nl; o += wb_nthr()) {
        const int k = o / Nl, c = o - k * Nl;
        float acc = part[o];
        for (int i = 0; i < ns; ++i) acc += s_gq[i][k] * s_sm[i][c];
        part[o] = acc;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) k_cls_reduce(WbDecB a) {
  const waldo_geom_t g = a.f.g;
  const int n = g.No * g.Nl, b = blockIdx.x;
  for (int e = wb_warp(); e < n; e += wb_nthr() / WB_WARP) {   // one warp per element, lanes over the CTA partials
    float acc = 0.f;
    for (int c = wb_lane(); c < a.f.prof_ctas; c += WB_WARP) acc += a.cls_part[((size_t)b * a.f.prof_ctas + c) * n + e];
    acc = wb_warp_sum(acc);
    if (wb_lane() == 0) a.d_cls[(size_t)b * n + e] += acc;
  }
}

// ============================================================================ low-res chain
// B1 backward: d a_lo -> d (obj|bg) alpha (scatter) and d src_grid (grid gradient).
__global__ void k_project_alpha_bwd(WbDecB a) {
  const WbDec& d = a.f;
  const waldo_geom_t g = d.g;
  const int L = wb_L(g), HW = g.H * g.W;
  const long long total = (long long)g.B * g.Tw * L * HW;
  for (long long e = (long long)blockIdx.x * wb_nthr() + wb_tid(); e < total; e += (long long)gridDim.x * wb_nthr()) {
    const float gv = a.d_a_lo[e];
    if (gv == 0.f) continue;
    int p = (int)(e % HW);
    int k = (int)((e / HW) % L);
    int t = (int)((e / ((long long)HW * L)) % g.Tw);
    int b = (int)(e / ((long long)HW * L * g.Tw));
    const float* sg; const float* plane; float* dplane; float* dsg; int w, h;
    if (k == 0) {
      size_t o = (((size_t)b * g.T + t) * HW + p) * 2;
      sg = d.src_grid_bg + o; dsg = a.d_src_grid_bg ? a.d_src_grid_bg + o : nullptr;
      plane = d.bg_alpha + (size_t)b * HW; dplane = a.d_bg_alpha ? a.d_bg_alpha + (size_t)b * HW : nullptr; w = g.W; h = g.H;
    } else {
      size_t o = ((((size_t)b * g.T + t) * g.No + (k - 1)) * HW + p) * 2;
      sg = d.src_grid_obj + o; dsg = a.d_src_grid_obj ? a.d_src_grid_obj + o : nullptr;
      size_t po = ((size_t)b * g.No + (k - 1)) * g.Ho * g.Wo;
      plane = d.obj_alpha + po; dplane = a.d_obj_alpha ? a.d_obj_alpha + po : nullptr; w = g.Wo; h = g.Ho;
    }
    WbTaps tp = wb_taps(__ldg(sg), __ldg(sg + 1), w, h);
    const int m = wb_tap_mask(tp, w, h);
    const long long off = (long long)tp.y0 * w + tp.x0;
    const float* q = plane + off;
    const float vnw = (m & 1) ? (__ldg(q) + 1.f) * 0.5f : 0.f, vne = (m & 2) ? (__ldg(q + 1) + 1.f) * 0.5f : 0.f;
    const float vsw = (m & 4) ? (__ldg(q + w) + 1.f) * 0.5f : 0.f, vse = (m & 8) ? (__ldg(q + w + 1) + 1.f) * 0.5f : 0.f;
    if (dplane) {
      float* o = dplane + off;
      if (m & 1) WB_RED_NZ(o, 0.5f * tp.nw * gv);
      if (m & 2) WB_RED_NZ(o + 1, 0.5f * tp.ne * gv);
      if (m & 4) WB_RED_NZ(o + w, 0.5f * tp.sw * gv);
      if (m & 8) WB_RED_NZ(o + w + 1, 0.5f * tp.se * gv);
    }
    if (dsg) {
      WB_RED_NZ(dsg, gv * ((vne - vnw) * tp.wy0 + (vse - vsw) * tp.wy1) * (0.5f * (float)w));
      WB_RED_NZ(dsg + 1, gv * ((vsw - vnw) * tp.wx0 + (vse - vne) * tp.wx1) * (0.5f * (float)h));
    }
  }
}

// B5 backward: d f_lo -> d tgt_grid (context +, target -) and d src_grid of the target frame.
__global__ void k_layer_flow_lo_bwd(WbDecB a) {
  const WbDec& d = a.f;
  const waldo_geom_t g = d.g;
  const int L = wb_L(g), HW = g.H * g.W;
  const long long total = (long long)g.B * g.Tp * L * HW;
  for (long long e = (long long)blockIdx.x * wb_nthr() + wb_tid(); e < total; e += (long long)gridDim.x * wb_nthr()) {
    int p = (int)(e % HW);
    int k = (int)((e / HW) % L);
    int tp = (int)((e / ((long long)HW * L)) % g.Tp);
    int b = (int)(e / ((long long)HW * L * g.Tp));
    int u = (int)d.pred_ts[tp];
    const float* sg; float* dsg; int w, h; size_t frame_stride, layer_off;
    const float* tg; float* dtg;
    if (k == 0) {
      w = g.W; h = g.H; frame_stride = (size_t)HW * 2; layer_off = 0;
      size_t o = (((size_t)b * g.T + u) * HW + p) * 2;
      sg = d.src_grid_bg + o; dsg = a.d_src_grid_bg ? a.d_src_grid_bg + o : nullptr;
      tg = d.tgt_grid_bg; dtg = a.d_tgt_grid_bg;
    } else {
      w = g.Wo; h = g.Ho; frame_stride = (size_t)g.No * g.Ho * g.Wo * 2; layer_off = (size_t)(k - 1) * g.Ho * g.Wo * 2;
      size_t o = ((((size_t)b * g.T + u) * g.No + (k - 1)) * HW + p) * 2;
      sg = d.src_grid_obj + o; dsg = a.d_src_grid_obj ? a.d_src_grid_obj + o : nullptr;
      tg = d.tgt_grid_obj; dtg = a.d_tgt_grid_obj;
    }
    WbTaps t = wb_taps(__ldg(sg), __ldg(sg + 1), w, h);
    const int m = wb_tap_mask(t, w, h);
    if (m == 0) continue;
    const long long o = ((long long)t.y0 * w + t.x0) * 2;
    const size_t base_u = ((size_t)b * g.T + u) * frame_stride + layer_off;
    float gsx = 0.f, gsy = 0.f;
    for (int tc = 0; tc < g.Tc; ++tc) {
      const int c_t = (int)d.ctx_ts[((size_t)b * g.Tc + tc) * g.Tp + tp];
      const float* gf = a.d_f_lo + (((((size_t)b * g.Tc + tc) * g.Tp + tp) * L + k) * HW + p) * 2;
      const float g0 = gf[0], g1 = gf[1];
      if (g0 == 0.f && g1 == 0.f) continue;
      const size_t base_c = ((size_t)b * g.T + c_t) * frame_stride + layer_off;
      WB_UNROLL for (int c = 0; c < 2; ++c) {
        const float gc = c == 0 ? g0 : g1;
        const float vnw = (m & 1) ? tg[base_c + o + c] - tg[base_u + o + c] : 0.f;
        const float vne = (m & 2) ? tg[base_c + o + 2 + c] - tg[base_u + o + 2 + c] : 0.f;
        const float vsw = (m & 4) ? tg[base_c + o + 2 * w + c] - tg[base_u + o + 2 * w + c] : 0.f;
        const float vse = (m & 8) ? tg[base_c + o + 2 * w + 2 + c] - tg[base_u + o + 2 * w + 2 + c] : 0.f;
        gsx += gc * ((vne - vnw) * t.wy0 + (vse - vsw) * t.wy1);
        gsy += gc * ((vsw - vnw) * t.wx0 + (vse - vne) * t.wx1);
        if (dtg && c_t != u) {
          if (m & 1) { WB_RED_NZ(dtg + base_c + o + c, t.nw * gc); WB_RED_NZ(dtg + base_u + o + c, -t.nw * gc); }
          if (m & 2) { WB_RED_NZ(dtg + base_c + o + 2 + c, t.ne * gc); WB_RED_NZ(dtg + base_u + o + 2 + c, -t.ne * gc); }
          if (m & 4) { WB_RED_NZ(dtg + base_c + o + 2 * w + c, t.sw * gc); WB_RED_NZ(dtg + base_u + o + 2 * w + c, -t.sw * gc); }
          if (m & 8) { WB_RED_NZ(dtg + base_c + o + 2 * w + 2 + c, t.se * gc); WB_RED_NZ(dtg + base_u + o + 2 * w + 2 + c, -t.se * gc); }
        }
      }
    }
    if (dsg) {
      WB_RED_NZ(dsg, gsx * (0.5f * (float)w));
      WB_RED_NZ(dsg + 1, gsy * (0.5f * (float)h));
    }
  }
}

// ============================================================================ launcher
static inline unsigned wb_blocks_b(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (1 << 20)) b = 1 << 20;
  return (unsigned)b;
}

static int wb_decode_bwd_launch(const WbDecB& a, waldo_stream_t st) {
  const WbDec& d = a.f;
  const waldo_geom_t& g = d.g;
#define WB_BREQ(cond, msg) do { if (!(cond)) return wb_fail(WALDO_EINVAL, "decode_bwd: %s", msg); } while (0)
#define WB_BLAUNCHED() do { int rc_ = WB_CHECK_LAUNCH(); if (rc_) return rc_; } while (0)
  WB_BREQ(g.B > 0 && g.No >= 1 && g.No + 1 <= WB_MAX_L && g.Nl <= WB_MAX_NL && g.C <= WB_MAX_C, "bad geometry");
  WB_BREQ(d.input && d.alpha && d.f_lo && d.a_lo && d.out_full && d.norm && d.occ && d.ctx_ts && d.pred_ts, "forward state missing");
  WB_BREQ(a.red_ctas > 0, "red_ctas must be positive");
  WB_BREQ(d.storage == WALDO_ST_F32, "the backward needs fp32 storage (bf16 storage is forward / inference only)");
  const int L = g.No + 1, HW = g.H * g.W;
  const bool filt = (g.flags & WALDO_F_FILTER) != 0;
  const bool from_cls = (g.flags & WALDO_F_HAS_CLS) && !(g.flags & WALDO_F_WEIGHT_CLS);
  const bool geom = a.d_tgt_grid_obj || a.d_src_grid_obj || a.d_tgt_grid_bg || a.d_src_grid_bg;
  const bool need_alpha_chain = geom || a.d_obj_alpha || a.d_bg_alpha || a.d_cls || a.d_occ || (filt && a.d_input) || a.d_alpha;
  if (a.d_occ) WB_BREQ(a.occ_part, "occ_part scratch missing");
  if (geom) WB_BREQ(a.d_f_lo && a.d_a_lo && a.d_alpha_acc, "geometry gradients need d_f_lo, d_a_lo, d_alpha_acc scratch");
  if (a.d_obj_alpha || a.d_bg_alpha || a.d_cls) WB_BREQ(a.d_a_lo && a.d_alpha_acc, "alpha gradients need d_a_lo, d_alpha_acc scratch");
  if (g.Hd != g.H) WB_BREQ(g.Hd >= 2 * g.H && g.Hd <= 4 * g.H, "backward supports scale_hd in {1} or [2, 4]");
  if (g.Hd != g.H && (a.d_f_lo || a.d_a_lo)) {
    WB_BREQ(a.up_tab, "up_tab scratch missing");
    WB_LAUNCH(k_up_tab, dim3((g.W + 63) / 64), dim3(64), 0, st, g.W, g.Wd, a.up_tab);
    WB_BLAUNCHED();
  }
  const bool st_gather = a.stages == 0 || (a.stages & 1), st_layers = a.stages == 0 || (a.stages & 2), st_rest = a.stages == 0 || (a.stages & 4);
  const bool st_aprep = a.stages == 0 || (a.stages & 8);
  const bool need_layers = a.d_alpha_acc || a.d_f_lo || a.d_occ;
  if (need_layers) WB_BREQ(a.glue && d.score, "glue / score buffers missing");
#if WB_DET
  // ---- deterministic accumulation: sizes of the scatter targets, arena checks, conversion helper
  const long long HWdl = (long long)g.Hd * g.Wd;
  const bool self_ctx = (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
  const long long n_input = (long long)g.B * g.T * g.C * HWdl, n_alpha = (long long)g.B * g.Tw * L * HWdl;
  const long long n_f_lo = (long long)g.B * g.Tc * g.Tp * L * HW * 2, n_a_lo = (long long)g.B * g.Tw * L * HW;
  const long long n_tgo = (long long)g.B * g.T * g.No * g.Ho * g.Wo * 2, n_sgo = (long long)g.B * g.T * g.No * HW * 2;
  const long long n_gbg = (long long)g.B * g.T * HW * 2, n_oa = (long long)g.B * g.No * g.Ho * g.Wo, n_ba = (long long)g.B * HW;
  WB_BREQ(a.det_base && a.det_shadow && a.det_scale && a.det_n > 0, "deterministic mode needs det_base, det_shadow, det_scale");
  {
    const float* tg[10] = {a.d_input, a.d_alpha_acc, a.d_f_lo, a.d_a_lo, a.d_tgt_grid_obj, a.d_src_grid_obj, a.d_tgt_grid_bg,
                           a.d_src_grid_bg, a.d_obj_alpha, a.d_bg_alpha};
    const long long tn[10] = {n_input, n_alpha, n_f_lo, n_a_lo, n_tgo, n_sgo, n_gbg, n_gbg, n_oa, n_ba};
    for (int i = 0; i < 10; ++i)
      WB_BREQ(!tg[i] || (tg[i] >= a.det_base && tg[i] + tn[i] <= a.det_base + a.det_n), "a scatter target lies outside the det arena");
  }
  auto det_convert = [&](float* dst, long long n) -> int {
    if (!dst) return 0;
    WB_LAUNCH(k_det_convert, dim3(wb_blocks_b(n, 256 * 4) > 1184 ? 1184 : wb_blocks_b(n, 256 * 4)), dim3(256), 0, st,
              a.det_shadow + (dst - a.det_base), dst, n, a.det_scale);
    return WB_CHECK_LAUNCH();
  };
  auto det_finish = [&]() -> int {   // the targets whose last addend comes from the low-res chain / the context-alpha kernels
    float* tg[7] = {a.d_input, a.d_tgt_grid_obj, a.d_src_grid_obj, a.d_tgt_grid_bg, a.d_src_grid_bg, a.d_obj_alpha, a.d_bg_alpha};
    const long long tn[7] = {n_input, n_tgo, n_sgo, n_gbg, n_gbg, n_oa, n_ba};
    for (int i = 0; i < 7; ++i) { int rc_ = det_convert(tg[i], tn[i]); if (rc_) return rc_; }
    return 0;
  };
  if (st_gather) {   // fixed-point unit from the largest upstream gradient magnitude (a maximum is order-independent)
    WB_LAUNCH(k_det_clear, dim3(1), dim3(32), 0, st, a.det_scale);
    WB_BLAUNCHED();
    const int TcRd = g.Tc + (self_ctx ? 1 : 0), CRd = g.C + L + ((g.flags & WALDO_F_USE_DISOCC) ? 1 : 0);
    const float* up[5] = {a.d_output, a.d_raw_alpha, a.d_raw_output, a.d_flow, a.d_alpha};
    const long long un[5] = {(long long)g.B * g.Tp * g.C * HWdl, (long long)g.B * g.Tp * HWdl,
                             (long long)g.B * TcRd * g.Tp * CRd * HWdl, (long long)g.B * g.Tc * g.Tp * 2 * HWdl, n_alpha};
    for (int i = 0; i < 5; ++i) {
      if (!up[i]) continue;
      WB_LAUNCH(k_det_absmax, dim3(wb_blocks_b(un[i], 256 * 8) > 1184 ? 1184 : wb_blocks_b(un[i], 256 * 8)), dim3(256), 0, st, up[i], un[i], a.det_scale);
      WB_BLAUNCHED();
    }
    WB_LAUNCH(k_det_scale, dim3(1), dim3(32), 0, st, a.det_scale);
    WB_BLAUNCHED();
  }
#define WB_DET_DO(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)
#else
#define WB_DET_DO(expr) ((void)0)
#endif
  // 1. HD gather backward
  if (st_gather) {
    WbDecB ag = a;
    if (!need_layers) ag.glue = nullptr;
    const dim3 ggrid(wb_blocks_b((long long)g.Hd * g.Wd, WB_TILE_PX) > 1024 ? 1024 : wb_blocks_b((long long)g.Hd * g.Wd, WB_TILE_PX), g.B * g.Tp);
    const bool self = (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
    const bool fast = g.Tc % WB_GB_TG == 0 && !self && a.d_input && a.d_raw_output && a.d_output;
#if !defined(WB_HOST_EMU) && WB_GB_ASYNC
    if (fast) {
      const size_t ring = (size_t)WB_GB_DEPTH * (5 * WB_GB_TG + 2) * WB_TILE_PX * sizeof(float);
      cudaFuncSetAttribute(k_gather_bwd_async<WB_GB_TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);   // > 48 KB: opt in
      WB_LAUNCH((k_gather_bwd_async<WB_GB_TG>), ggrid, dim3(WB_TILE_PX), ring, st, ag);
    } else
#endif
    if (fast) WB_LAUNCH((k_gather_bwd<WB_GB_TG, true>), ggrid, dim3(WB_TILE_PX), 0, st, ag);
    else WB_LAUNCH((k_gather_bwd<2, false>), ggrid, dim3(WB_TILE_PX), 0, st, ag);
    WB_BLAUNCHED();
  }
  // 1b. HD layer backward
  if (st_layers && need_layers) {
    WB_LAUNCH(k_layers_bwd, dim3(a.red_ctas, g.B * g.Tp), dim3(WB_TILE_PX), 0, st, a);
    WB_BLAUNCHED();
    if (a.d_occ) {
      WB_LAUNCH(k_occ_reduce, dim3(g.B * g.Tp, (L * L + 31) / 32), dim3(32), 0, st, a.occ_part, g.B * g.Tp, g.Tp, a.red_ctas, L * L, g.T, d.pred_ts, 0, a.d_occ);
      WB_BLAUNCHED();
    }
    // (deterministic mode) the two targets of the layer kernel are complete: the kernels below read them as fp32
    WB_DET_DO(det_convert(a.d_alpha_acc, n_alpha));
    WB_DET_DO(det_convert(a.d_f_lo, n_f_lo));
  }
  if (!need_alpha_chain || !(st_rest || st_aprep)) {
    if (st_rest) WB_DET_DO(det_finish());
    return 0;
  }
  // 2. context-alpha backward
  if (st_aprep && (a.d_alpha_acc || a.d_alpha)) {
    if (filt && a.d_prof_p) WB_BREQ(a.prof_p_part, "prof_p_part scratch missing");
    const dim3 pgrid(a.red_ctas, g.B * g.Tw);
    const size_t dyn = WB_LANES_PREP_BWD ? (size_t)WB_NWARP * WB_MAX_NL * 32 * sizeof(float) : 0;   // softmax staging of the lanes form only
#ifndef WB_HOST_EMU
    // static + dynamic shared memory exceeds the 48 KB default: opt in (cheap, idempotent)
    if (g.Nl == 20) cudaFuncSetAttribute(k_alpha_prep_bwd<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    else if (g.Nl == 19) cudaFuncSetAttribute(k_alpha_prep_bwd<19>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    else cudaFuncSetAttribute(k_alpha_prep_bwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
#endif
    if (g.Nl == 20) WB_LAUNCH(k_alpha_prep_bwd<20>, pgrid, dim3(WB_TILE_PX), dyn, st, a);
    else if (g.Nl == 19) WB_LAUNCH(k_alpha_prep_bwd<19>, pgrid, dim3(WB_TILE_PX), dyn, st, a);
    else WB_LAUNCH(k_alpha_prep_bwd<0>, pgrid, dim3(WB_TILE_PX), dyn, st, a);
    WB_BLAUNCHED();
    if (a.d_occ) {
      WB_LAUNCH(k_occ_reduce, dim3(g.B * g.Tw, (L * L + 31) / 32), dim3(32), 0, st, a.occ_part, g.B * g.Tw, g.Tw, a.red_ctas, L * L, g.T, d.pred_ts, 1, a.d_occ);
      WB_BLAUNCHED();
    }
    WB_DET_DO(det_convert(a.d_a_lo, n_a_lo));   // k_class_profile_bwd / k_project_alpha_bwd read (and update in place) fp32
  }
  if (!st_rest) return 0;
  if (a.d_alpha_acc || a.d_alpha) {
    if (filt && a.d_prof_p) {
      WB_BREQ(a.d_prof_sum, "d_prof_sum scratch missing");
      WB_LAUNCH(k_profile_final_bwd, dim3(g.B), dim3(1024), 0, st, a);
      WB_BLAUNCHED();
      if (!from_cls) {
        if ((g.flags & WALDO_F_WEIGHT_CLS) && a.d_cls) WB_BREQ(a.cls_part, "cls_part scratch missing");
        if (g.Nl == 20) WB_LAUNCH(k_class_profile_bwd<20>, dim3(d.prof_ctas, g.B), dim3(256), 0, st, a);
        else if (g.Nl == 19) WB_LAUNCH(k_class_profile_bwd<19>, dim3(d.prof_ctas, g.B), dim3(256), 0, st, a);
        else WB_LAUNCH(k_class_profile_bwd<0>, dim3(d.prof_ctas, g.B), dim3(256), 0, st, a);
        WB_BLAUNCHED();
        if ((g.flags & WALDO_F_WEIGHT_CLS) && a.d_cls) {
          WB_LAUNCH(k_cls_reduce, dim3(g.B), dim3(1024), 0, st, a);
          WB_BLAUNCHED();
        }
      }
    }
  }
  // 3. low-res chain
  if (a.d_a_lo && (a.d_obj_alpha || a.d_bg_alpha || a.d_src_grid_obj || a.d_src_grid_bg)) {
    WB_LAUNCH(k_project_alpha_bwd, dim3(wb_blocks_b((long long)g.B * g.Tw * L * HW, 256)), dim3(256), 0, st, a);
    WB_BLAUNCHED();
  }
  if (a.d_f_lo && geom) {
    WB_LAUNCH(k_layer_flow_lo_bwd, dim3(wb_blocks_b((long long)g.B * g.Tp * L * HW, 256)), dim3(256), 0, st, a);
    WB_BLAUNCHED();
  }
  WB_DET_DO(det_finish());
  return 0;
#undef WB_BREQ
#undef WB_BLAUNCHED
#undef WB_DET_DO
}

}   // namespace wb_plain / wb_fixed
