// Stage A kernels: TPS grids from control points, inverse warp by splatting, occlusion matrix.
// Reference: models/modules/warp.py (TPSWarp :21-55, InverseWarp :58-174), models/nets/lvd.py:59-68.
#pragma once
#include "wb_common.cuh"

// =========================================================================================== TPS
// mapping[n] = inverse_kernel @ [pts[n]; 0]   (warp.py:52-53), fp64 accumulation
__global__ void k_tps_mapping(int n, int N, const float* __restrict__ inv, const float* __restrict__ pts,
                              double* __restrict__ mapping) {
  const int K = N + 3;
  const int item = blockIdx.x;
  const float* p = pts + (size_t)item * N * 2;
  for (int j = wb_tid(); j < K; j += wb_nthr()) {
    double ax = 0.0, ay = 0.0;
    const float* row = inv + (size_t)j * K;
    for (int i = 0; i < N; ++i) {
      double w = (double)__ldg(row + i);
      ax += w * (double)__ldg(p + 2 * i);
      ay += w * (double)__ldg(p + 2 * i + 1);
    }
    mapping[((size_t)item * K + j) * 2] = ax;
    mapping[((size_t)item * K + j) * 2 + 1] = ay;
  }
}

// grid[n,p] = tgt_grid_repr[p,:] @ mapping[n]   (warp.py:54).  One thread per lattice point, WB_TPS_NI
// items per pass so that every repr element loaded feeds 2*NI fp64 FMAs; mappings staged in smem.
#define WB_TPS_NI 8
__global__ void __launch_bounds__(128) k_tps_eval(int n, int N, int P, const float* __restrict__ repr,
                                                  const double* __restrict__ mapping, float* __restrict__ grid) {
  const int K = N + 3;
  __shared__ double s_map[WB_TPS_NI * WB_MAX_K * 2];
  const int item0 = blockIdx.y * WB_TPS_NI;
  const int ni = min(WB_TPS_NI, n - item0);
  for (int i = wb_tid(); i < ni * K * 2; i += wb_nthr()) s_map[i] = mapping[(size_t)item0 * K * 2 + i];
  __syncthreads();
  for (int p = blockIdx.x * wb_nthr() + wb_tid(); p < P; p += gridDim.x * wb_nthr()) {
    double ax[WB_TPS_NI], ay[WB_TPS_NI];
    WB_UNROLL for (int i = 0; i < WB_TPS_NI; ++i) { ax[i] = 0.0; ay[i] = 0.0; }
    const float* row = repr + (size_t)p * K;
    for (int j = 0; j < K; ++j) {
      double r = (double)__ldg(row + j);
      WB_UNROLL for (int i = 0; i < WB_TPS_NI; ++i) {
        ax[i] += r * s_map[(i * K + j) * 2];
        ay[i] += r * s_map[(i * K + j) * 2 + 1];
      }
    }
    WB_UNROLL for (int i = 0; i < WB_TPS_NI; ++i)
      if (i < ni) {
        float* o = grid + ((size_t)(item0 + i) * P + p) * 2;
        o[0] = (float)ax[i]; o[1] = (float)ay[i];
      }
  }
}

// backward: partial[n,chunk,j,:] = sum_{p in chunk} repr[p,j] * dgrid[n,p,:]  (ordered, one thread per j)
__global__ void k_tps_bwd_partial(int n, int N, int P, int chunks, const float* __restrict__ repr,
                                  const float* __restrict__ dgrid, double* __restrict__ partial) {
  const int K = N + 3;
  const int item = blockIdx.x, ch = blockIdx.y;
  const int per = (P + chunks - 1) / chunks;
  const int p0 = ch * per, p1 = min(P, p0 + per);
  for (int j = wb_tid(); j < K; j += wb_nthr()) {
    double ax = 0.0, ay = 0.0;
    WB_UNROLL_N(8) for (int p = p0; p < p1; ++p) {
      double r = (double)__ldg(repr + (size_t)p * K + j);
      ax += r * (double)__ldg(dgrid + ((size_t)item * P + p) * 2);
      ay += r * (double)__ldg(dgrid + ((size_t)item * P + p) * 2 + 1);
    }
    double* o = partial + (((size_t)item * chunks + ch) * K + j) * 2;
    o[0] = ax; o[1] = ay;
  }
}
// (Measured alternatives on B200, profiles/r1_v48_tps_bwd_ab.md: one warp per (item, slice) with lanes over the points and the
//  K x 2 sums in registers -- rows of repr read straight from global memory 0.18 ms, staged through shared memory 0.43 ms per
//  launch -- both slower than this form's 0.14 ms.)
// dpts[n,i,:] = sum_j inverse_kernel[j,i] * (sum_chunks partial[n,chunk,j,:])
__global__ void k_tps_bwd_final(int n, int N, int chunks, const float* __restrict__ inv,
                                const double* __restrict__ partial, float* __restrict__ dpts) {
  const int K = N + 3;
  __shared__ double s_tot[WB_MAX_K * 2];
  const int item = blockIdx.x;
  for (int j = wb_tid(); j < K; j += wb_nthr()) {
    double ax = 0.0, ay = 0.0;
    for (int c = 0; c < chunks; ++c) {
      const double* q = partial + (((size_t)item * chunks + c) * K + j) * 2;
      ax += q[0]; ay += q[1];
    }
    s_tot[2 * j] = ax; s_tot[2 * j + 1] = ay;
  }
  __syncthreads();
  for (int i = wb_tid(); i < N; i += wb_nthr()) {
    double ax = 0.0, ay = 0.0;
    for (int j = 0; j < K; ++j) {
      double w = (double)__ldg(inv + (size_t)j * K + i);
      ax += w * s_tot[2 * j]; ay += w * s_tot[2 * j + 1];
    }
    dpts[((size_t)item * N + i) * 2] = (float)ax;
    dpts[((size_t)item * N + i) * 2 + 1] = (float)ay;
  }
}

// ================================================================================ inverse warp
// One CTA per item; block-stride loops over the (padded) lattice separated by __syncthreads().
// The two byte maps `level` / `eroded` make every phase race-free without ping-pong buffers:
//   * dilation k reads only cells with level <= k-1 (stable) and writes level = k, val of its own cell;
//   * erosion k reads only `eroded` marks < k and writes mark k on its own cell.
struct WbInvArgs {
  int n, Hs, Ws, Ht, Wt, niter, erode;
  const float* fwd; const float* id_src; const float* id_tgt; const float* gauss;
  float* out; int32_t* field; int32_t* winner; uint8_t* level; uint8_t* eroded; float* val; int32_t* bbox;
};

WB_DEV bool wb_level_known_before(uint8_t lv, int it) { return lv != 255 && (int)lv < it; }

// The forward runs as a sequence of flat kernels, one per phase, each one thread per (item, cell): every phase is
// embarrassingly parallel over cells, and a kernel boundary is the only ordering it needs (the one-CTA-per-item form
// with __syncthreads() between phases kept 40 background items on 40 SMs).  Cells outside the bounding box of an
// item's hit cells grown by the iteration count cannot change in the dilation / erosion phases and exit at once.
struct WbInvItem {
  const float* fwd; int32_t* field; int32_t* winner; uint8_t* level; uint8_t* eroded; float* vx; float* vy; int32_t* bbox;
};
WB_DEV WbInvItem wb_inv_item(const WbInvArgs& a, int item, int P, int PP) {
  WbInvItem r;
  r.fwd = a.fwd + (size_t)item * a.Hs * a.Ws * 2;
  r.field = a.field + (size_t)item * P; r.winner = a.winner + (size_t)item * P;
  r.level = a.level + (size_t)item * PP; r.eroded = a.eroded + (size_t)item * PP;
  r.vx = a.val + (size_t)item * 2 * PP; r.vy = r.vx + PP;
  r.bbox = a.bbox + (size_t)item * 4;
  return r;
}
#define WB_INV_GEOM \
  const int Ht = a.Ht, Wt = a.Wt, P = Ht * Wt, m = a.niter + 1; \
  const int Hp = Ht + 2 * m, Wp = Wt + 2 * m, PP = Hp * Wp; (void)P; (void)PP; (void)Hp; (void)Wp

// displacement of lattice sample (X, Y) in pixels: interpolate (fwd - identity) to the target size   warp.py:76-79
WB_DEV void wb_inv_disp(const WbInvArgs& a, const float* __restrict__ fwd, int X, int Y, float& dx, float& dy) {
  const float rh = (float)a.Hs / (float)a.Ht, rw = (float)a.Ws / (float)a.Wt;   // area_pixel_compute_scale
  WbAxis ay = wb_axis(Y, rh, a.Hs), ax = wb_axis(X, rw, a.Ws);
  float d[2];
  WB_UNROLL for (int c = 0; c < 2; ++c) {
    int i00 = (ay.i0 * a.Ws + ax.i0) * 2 + c, i01 = (ay.i0 * a.Ws + ax.i1) * 2 + c;
    int i10 = (ay.i1 * a.Ws + ax.i0) * 2 + c, i11 = (ay.i1 * a.Ws + ax.i1) * 2 + c;
    float v00 = __fsub_rn(__ldg(fwd + i00), __ldg(a.id_src + i00));
    float v01 = __fsub_rn(__ldg(fwd + i01), __ldg(a.id_src + i01));
    float v10 = __fsub_rn(__ldg(fwd + i10), __ldg(a.id_src + i10));
    float v11 = __fsub_rn(__ldg(fwd + i11), __ldg(a.id_src + i11));
    d[c] = wb_lerp2(v00, v01, v10, v11, ax, ay);
  }
  dx = __fdiv_rn(__fmul_rn(d[0], (float)a.Wt), 2.f); dy = __fdiv_rn(__fmul_rn(d[1], (float)a.Ht), 2.f);
}

// phase 0: clear.  grid = (ctas, n)
__global__ void __launch_bounds__(256) k_inv_clear(WbInvArgs a) {
  WB_INV_GEOM;
  const WbInvItem it = wb_inv_item(a, blockIdx.y, P, PP);
  for (int i = blockIdx.x * wb_nthr() + wb_tid(); i < PP; i += gridDim.x * wb_nthr()) {
    it.level[i] = 255; it.eroded[i] = 0; it.vx[i] = 0.f; it.vy[i] = 0.f;
    if (i < P) it.winner[i] = INT_MAX;
    if (i < 4) it.bbox[i] = (i < 2) ? INT_MAX : -1;   // x0, y0, x1, y1 of the hit cells (padded coordinates)
  }
}
// phase 1: landing cell of every lattice sample; lowest sample index claims the cell   warp.py:76-88,113-117
__global__ void __launch_bounds__(256) k_inv_claim(WbInvArgs a) {
  WB_INV_GEOM;
  const WbInvItem it = wb_inv_item(a, blockIdx.y, P, PP);
  for (int s = blockIdx.x * wb_nthr() + wb_tid(); s < P; s += gridDim.x * wb_nthr()) {
    int Y = s / Wt, X = s - Y * Wt;
    float dx, dy;
    wb_inv_disp(a, it.fwd, X, Y, dx, dy);
    float fx = rintf(__fadd_rn((float)X, dx)), fy = rintf(__fadd_rn((float)Y, dy));   // half-to-even
    int cell = -1;
    if (fx >= 0.f && fy >= 0.f && fx <= (float)(Wt - 1) && fy <= (float)(Ht - 1)) {
      cell = (int)fy * Wt + (int)fx;
      atomicMin(&it.winner[cell], s);
    }
    it.field[s] = cell;
  }
}
// phase 2: winners deposit the negated displacement                                       warp.py:121-123
__global__ void __launch_bounds__(256) k_inv_deposit(WbInvArgs a) {
  WB_INV_GEOM;
  const WbInvItem it = wb_inv_item(a, blockIdx.y, P, PP);
  __shared__ int s_bb[4];   // the CTA's share of the bounding box: one global atomic per CTA and bound, not per sample
  for (int i = wb_tid(); i < 4; i += wb_nthr()) s_bb[i] = (i < 2) ? INT_MAX : -1;
  __syncthreads();
  for (int s = blockIdx.x * wb_nthr() + wb_tid(); s < P; s += gridDim.x * wb_nthr()) {
    int cell = it.field[s];
    if (cell < 0 || it.winner[cell] != s) continue;
    int Y = s / Wt, X = s - Y * Wt;
    float dx, dy;
    wb_inv_disp(a, it.fwd, X, Y, dx, dy);
    int cy = cell / Wt, cx = cell - cy * Wt;
    int pc = (cy + m) * Wp + cx + m;
    it.vx[pc] = -dx; it.vy[pc] = -dy; it.level[pc] = 0;
    atomicMin(&s_bb[0], cx + m); atomicMin(&s_bb[1], cy + m);
    atomicMax(&s_bb[2], cx + m); atomicMax(&s_bb[3], cy + m);
  }
  __syncthreads();
  for (int i = wb_tid(); i < 4; i += wb_nthr()) {
    if (i < 2) { if (s_bb[i] != INT_MAX) atomicMin(&it.bbox[i], s_bb[i]); }
    else if (s_bb[i] >= 0) atomicMax(&it.bbox[i], s_bb[i]);
  }
}
// The padded-lattice phases run on grid = (bands of WB_INV_ROWS rows, n): a CTA walks only the cells of its band that lie
// inside the hit cells' bounding box grown by `grow` (nothing can change outside it); most CTAs of an object exit at once.
#define WB_INV_ROWS 8
struct WbInvBand { int x0, y0, w, cells; };
WB_DEV WbInvBand wb_inv_band(const int32_t* __restrict__ bbox, int grow, int Hp, int Wp) {
  WbInvBand r;
  const int ya = blockIdx.x * WB_INV_ROWS, yb = min(Hp, ya + WB_INV_ROWS) - 1;
  const int bx0 = bbox[0], by0 = bbox[1], bx1 = bbox[2], by1 = bbox[3];
  r.x0 = 0; r.y0 = 0; r.w = 0; r.cells = 0;
  if (bx1 < 0) return r;                                  // no hit cell at all
  const int x0 = max(0, bx0 - grow), x1 = min(Wp - 1, bx1 + grow);
  const int y0 = max(ya, by0 - grow), y1 = min(yb, by1 + grow);
  if (x0 > x1 || y0 > y1) return r;
  r.x0 = x0; r.y0 = y0; r.w = x1 - x0 + 1; r.cells = r.w * (y1 - y0 + 1);
  return r;
}
// true when padded cell (x, y) lies outside the hit cells' bounding box grown by `grow` (nothing can change there)
WB_DEV bool wb_inv_outside(const int32_t* __restrict__ bbox, int x, int y, int grow) {
  return x < bbox[0] - grow || x > bbox[2] + grow || y < bbox[1] - grow || y > bbox[3] + grow;
}
// phase 3, iteration `iter`: a frontier cell takes the normalised Gaussian mean of the known cells   warp.py:135-151
WB_DEV void wb_inv_dilate_cell(const WbInvItem& it, const float* g, int x, int y, int iter, int Hp, int Wp) {
  const uint8_t* level = it.level;
  const int c = y * Wp + x;
  if (level[c] != 255) return;
  bool front = (y > 0 && wb_level_known_before(level[c - Wp], iter)) || (y < Hp - 1 && wb_level_known_before(level[c + Wp], iter)) ||
               (x > 0 && wb_level_known_before(level[c - 1], iter)) || (x < Wp - 1 && wb_level_known_before(level[c + 1], iter));
  if (!front) return;
  float sx = 0.f, sy = 0.f, sw = 0.f;
  WB_UNROLL for (int dy = -1; dy <= 1; ++dy)
    WB_UNROLL for (int dx = -1; dx <= 1; ++dx) {
      int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= Hp || xx < 0 || xx >= Wp) continue;
      int q = yy * Wp + xx;
      if (!wb_level_known_before(level[q], iter)) continue;
      float w = g[(dy + 1) * 3 + dx + 1];
      sx += w * it.vx[q]; sy += w * it.vy[q]; sw += w;
    }
  it.vx[c] = sx / sw; it.vy[c] = sy / sw; it.level[c] = (uint8_t)iter;
}
__global__ void __launch_bounds__(256) k_inv_dilate(WbInvArgs a, int iter) {
  WB_INV_GEOM;
  const WbInvItem it = wb_inv_item(a, blockIdx.y, P, PP);
  float g[9];
  WB_UNROLL for (int i = 0; i < 9; ++i) g[i] = __ldg(a.gauss + i);
  const WbInvBand band = wb_inv_band(it.bbox, iter, Hp, Wp);
  for (int i = wb_tid(); i < band.cells; i += wb_nthr()) wb_inv_dilate_cell(it, g, band.x0 + i % band.w, band.y0 + i / band.w, iter, Hp, Wp);
}
// phase 4, iteration `iter`: erosion of the known set (objects only)                      warp.py:153-162
WB_DEV void wb_inv_erode_cell(const WbInvItem& it, int x, int y, int iter, int Hp, int Wp) {
  const uint8_t* level = it.level;
  const uint8_t* eroded = it.eroded;
  const int c = y * Wp + x;
  if (level[c] == 255 || eroded[c] != 0) return;
#define WB_GONE(q) (level[q] == 255 || (eroded[q] != 0 && (int)eroded[q] < iter))
  bool edge = (y > 0 && WB_GONE(c - Wp)) || (y < Hp - 1 && WB_GONE(c + Wp)) ||
              (x > 0 && WB_GONE(c - 1)) || (x < Wp - 1 && WB_GONE(c + 1));
#undef WB_GONE
  if (edge) it.eroded[c] = (uint8_t)iter;
}
__global__ void __launch_bounds__(256) k_inv_erode(WbInvArgs a, int iter) {
  WB_INV_GEOM;
  const WbInvItem it = wb_inv_item(a, blockIdx.y, P, PP);
  const WbInvBand band = wb_inv_band(it.bbox, a.niter, Hp, Wp);
  for (int i = wb_tid(); i < band.cells; i += wb_nthr()) wb_inv_erode_cell(it, band.x0 + i % band.w, band.y0 + i / band.w, iter, Hp, Wp);
}
// Items whose hit cells span a small box (objects: a 64x64 canvas splatted into the 128x256 image) run ALL dilation and
// erosion iterations in one CTA, a __syncthreads() per iteration, over the grown box only.  grid = (n).
WB_DEV void wb_inv_box(const int32_t* bbox, int grow, int Hp, int Wp, int& x0, int& y0, int& w, int& cells) {
  x0 = y0 = w = cells = 0;
  if (bbox[2] < 0) return;
  x0 = max(0, bbox[0] - grow); y0 = max(0, bbox[1] - grow);
  const int x1 = min(Wp - 1, bbox[2] + grow), y1 = min(Hp - 1, bbox[3] + grow);
  w = x1 - x0 + 1; cells = w * (y1 - y0 + 1);
}
__global__ void __launch_bounds__(512) k_inv_grow_fused(WbInvArgs a) {
  WB_INV_GEOM;
  const WbInvItem it = wb_inv_item(a, blockIdx.x, P, PP);
  float g[9];
  WB_UNROLL for (int i = 0; i < 9; ++i) g[i] = __ldg(a.gauss + i);
  int x0, y0, w, cells;
  for (int iter = 1; iter <= a.niter; ++iter) {
    wb_inv_box(it.bbox, iter, Hp, Wp, x0, y0, w, cells);
    for (int i = wb_tid(); i < cells; i += wb_nthr()) wb_inv_dilate_cell(it, g, x0 + i % w, y0 + i / w, iter, Hp, Wp);
    __syncthreads();
  }
  if (a.erode) {
    wb_inv_box(it.bbox, a.niter, Hp, Wp, x0, y0, w, cells);
    for (int iter = 1; iter <= a.niter; ++iter) {
      for (int i = wb_tid(); i < cells; i += wb_nthr()) wb_inv_erode_cell(it, x0 + i % w, y0 + i / w, iter, Hp, Wp);
      __syncthreads();
    }
  }
}
// ------------------------------------------------------------------------------------------------------------------
// Objects, all phases in ONE kernel: one CTA per item, working set in shared memory.
// An object canvas (Hs x Ws lattice, 64 x 64 at the benchmark shape) lands in a small box of the Ht x Wt image.  The
// CTA stages the lattice displacements (fwd - identity) in shared memory, bounds the landing box from them, and runs
// clear -> claim -> deposit -> dilations -> erosions -> final over that box (grown by niter + 1) entirely in shared
// memory: `val` never reaches HBM, the displacement of a sample is interpolated from shared memory, and five launches
// (each a pass over n x Ht x Wt cells in global memory) become one.  The maps the backward and the parity checks read
// (field, winner, level, eroded, bbox) are written once at the end.  A box that does not fit the shared-memory budget,
// or a sample that lands outside the predicted box (the claim phase checks every sample, so the bound is never trusted),
// sends the item down the same code with the work area in global memory.  Same arithmetic, same tie rule, same results.
#define WB_INVF_THREADS 1024
#define WB_INVF_MAX_SRC 4096   // lattice points staged per item
struct WbInvArea {             // work area in padded coordinates: cell (x, y) -> (y - y0) * w + (x - x0)
  int x0, y0, w, h;
  int wx0, wy0, ww;            // the same for `winner` (unpadded image layout when the area lives in global memory)
  int* winner; float* vx; float* vy; uint8_t* level; uint8_t* eroded;
};
WB_DEV int wb_warp_min(int v) {
#ifndef WB_HOST_EMU
  v = __reduce_min_sync(0xffffffffu, v);
#endif
  return v;
}
WB_DEV int wb_warp_max(int v) {
#ifndef WB_HOST_EMU
  v = __reduce_max_sync(0xffffffffu, v);
#endif
  return v;
}
// walk s = tid, tid + nthr, ... over a W-wide lattice as (X, Y) without a division per step
struct WbWalk { int X, Y, sx, sy, W; };
WB_DEV WbWalk wb_walk(int tid, int nthr, int W) {
  WbWalk w;
  w.Y = tid / W; w.X = tid - w.Y * W; w.sy = nthr / W; w.sx = nthr - w.sy * W; w.W = W;
  return w;
}
WB_DEV void wb_walk_next(WbWalk& w) { w.X += w.sx; w.Y += w.sy; if (w.X >= w.W) { w.X -= w.W; ++w.Y; } }
// c / W for 0 <= c < 2^24 through a float reciprocal, corrected to the exact quotient
WB_DEV int wb_div_small(int c, int W, float invW) {
  int q = (int)((float)c * invW);
  if (q * W > c) --q; else if ((q + 1) * W <= c) ++q;
  return q;
}
// displacement of a sample in pixels from the staged lattice displacements: the arithmetic of wb_inv_disp
WB_DEV void wb_inv_disp_s(const WbInvArgs& a, const float2* s_d, const WbAxis& ax, const WbAxis& ay, float& dx, float& dy) {
  const float2 v00 = s_d[ay.i0 * a.Ws + ax.i0], v01 = s_d[ay.i0 * a.Ws + ax.i1];
  const float2 v10 = s_d[ay.i1 * a.Ws + ax.i0], v11 = s_d[ay.i1 * a.Ws + ax.i1];
  const float d0 = wb_lerp2(v00.x, v01.x, v10.x, v11.x, ax, ay), d1 = wb_lerp2(v00.y, v01.y, v10.y, v11.y, ax, ay);
  dx = __fdiv_rn(__fmul_rn(d0, (float)a.Wt), 2.f); dy = __fdiv_rn(__fmul_rn(d1, (float)a.Ht), 2.f);
}
// wb_inv_dilate_cell / wb_inv_erode_cell on a work area (bounds are those of the padded frame, Hp x Wp)
WB_DEV void wb_inv_dilate_area(const WbInvArea& A, const float* g, int x, int y, int iter, int Hp, int Wp) {
  const uint8_t* level = A.level;
  const int c = (y - A.y0) * A.w + (x - A.x0), S = A.w;
  if (level[c] != 255) return;
  bool front = (y > 0 && wb_level_known_before(level[c - S], iter)) || (y < Hp - 1 && wb_level_known_before(level[c + S], iter)) ||
               (x > 0 && wb_level_known_before(level[c - 1], iter)) || (x < Wp - 1 && wb_level_known_before(level[c + 1], iter));
  if (!front) return;
  float sx = 0.f, sy = 0.f, sw = 0.f;
  WB_UNROLL for (int dy = -1; dy <= 1; ++dy)
    WB_UNROLL for (int dx = -1; dx <= 1; ++dx) {
      int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= Hp || xx < 0 || xx >= Wp) continue;
      int q = c + dy * S + dx;
      if (!wb_level_known_before(level[q], iter)) continue;
      float w = g[(dy + 1) * 3 + dx + 1];
      sx += w * A.vx[q]; sy += w * A.vy[q]; sw += w;
    }
  A.vx[c] = sx / sw; A.vy[c] = sy / sw; A.level[c] = (uint8_t)iter;
}
WB_DEV void wb_inv_erode_area(const WbInvArea& A, int x, int y, int iter, int Hp, int Wp) {
  const uint8_t* level = A.level;
  const uint8_t* eroded = A.eroded;
  const int c = (y - A.y0) * A.w + (x - A.x0), S = A.w;
  if (level[c] == 255 || eroded[c] != 0) return;
#define WB_GONE(q) (level[q] == 255 || (eroded[q] != 0 && (int)eroded[q] < iter))
  bool edge = (y > 0 && WB_GONE(c - S)) || (y < Hp - 1 && WB_GONE(c + S)) ||
              (x > 0 && WB_GONE(c - 1)) || (x < Wp - 1 && WB_GONE(c + 1));
#undef WB_GONE
  if (edge) A.eroded[c] = (uint8_t)iter;
}
// cells of the hit box grown by `grow`, clipped to the work area
WB_DEV void wb_inv_box_area(const int* bb, int grow, const WbInvArea& A, int& x0, int& y0, int& w, int& cells) {
  x0 = y0 = w = cells = 0;
  if (bb[2] < 0) return;
  x0 = max(A.x0, bb[0] - grow); y0 = max(A.y0, bb[1] - grow);
  const int x1 = min(A.x0 + A.w - 1, bb[2] + grow), y1 = min(A.y0 + A.h - 1, bb[3] + grow);
  if (x1 < x0 || y1 < y0) return;
  w = x1 - x0 + 1; cells = w * (y1 - y0 + 1);
}
__global__ void __launch_bounds__(WB_INVF_THREADS, 1) k_inv_fused(WbInvArgs a, int cap, int mg_adj) {
  WB_INV_GEOM;
  const WbInvItem it = wb_inv_item(a, blockIdx.x, P, PP);
  const int NS = a.Hs * a.Ws, tid = wb_tid(), nthr = wb_nthr();
  WB_DYN_SMEM(smem);
  float2* s_d = reinterpret_cast<float2*>(smem);          // lattice displacements fwd - identity   warp.py:76
  int* s_win = reinterpret_cast<int*>(smem + 2 * NS);
  float* s_vx = smem + 2 * NS + cap;
  float* s_vy = s_vx + cap;
  uint8_t* s_lv = reinterpret_cast<uint8_t*>(s_vy + cap);
  uint8_t* s_er = s_lv + cap;
  __shared__ int s_bb[4], s_hit[4], s_over;
  float g[9];
  WB_UNROLL for (int i = 0; i < 9; ++i) g[i] = __ldg(a.gauss + i);
  for (int i = tid; i < 4; i += nthr) { s_bb[i] = i < 2 ? INT_MAX : INT_MIN; s_hit[i] = i < 2 ? INT_MAX : -1; }   // (block-stride: the
  if (tid == 0) s_over = 0;                                                                                      //  emulation runs one thread)
  __syncthreads();
  {   // stage the lattice, bound where its points land (pixels)
    int bx0 = INT_MAX, by0 = INT_MAX, bx1 = INT_MIN, by1 = INT_MIN;
    for (int i = tid; i < NS; i += nthr) {
      const float fx = __ldg(it.fwd + 2 * i), fy = __ldg(it.fwd + 2 * i + 1);
      s_d[i] = make_float2(__fsub_rn(fx, __ldg(a.id_src + 2 * i)), __fsub_rn(fy, __ldg(a.id_src + 2 * i + 1)));
      const float px = fminf(fmaxf(((fx + 1.f) * (float)Wt - 1.f) * 0.5f, -8.f), (float)Wt + 8.f);
      const float py = fminf(fmaxf(((fy + 1.f) * (float)Ht - 1.f) * 0.5f, -8.f), (float)Ht + 8.f);
      bx0 = min(bx0, (int)floorf(px)); bx1 = max(bx1, (int)ceilf(px));
      by0 = min(by0, (int)floorf(py)); by1 = max(by1, (int)ceilf(py));
    }
    bx0 = wb_warp_min(bx0); by0 = wb_warp_min(by0); bx1 = wb_warp_max(bx1); by1 = wb_warp_max(by1);
    if (wb_lane() == 0) { atomicMin(&s_bb[0], bx0); atomicMin(&s_bb[1], by0); atomicMax(&s_bb[2], bx1); atomicMax(&s_bb[3], by1); }
  }
  __syncthreads();
  // samples between lattice points land between the lattice points' own landing positions; beyond the outermost lattice
  // points the displacement is extended as a constant over half a lattice step (+ rounding): the margin
  const int mgx = (Wt + a.Ws - 1) / a.Ws / 2 + 2 + mg_adj, mgy = (Ht + a.Hs - 1) / a.Hs / 2 + 2 + mg_adj;   // (mg_adj: tests only)
  const int ix0 = max(0, s_bb[0] - mgx), ix1 = min(Wt - 1, s_bb[2] + mgx), iy0 = max(0, s_bb[1] - mgy), iy1 = min(Ht - 1, s_bb[3] + mgy);
  const bool some = ix0 <= ix1 && iy0 <= iy1;
  const int aw = some ? ix1 - ix0 + 1 + 2 * m : 0, ah = some ? iy1 - iy0 + 1 + 2 * m : 0;   // padded coordinates: + m on both sides
  bool fits = (long long)aw * ah <= (long long)cap;
  const float rh = (float)a.Hs / (float)Ht, rw = (float)a.Ws / (float)Wt, invWt = 1.f / (float)Wt;
  const bool xconst = nthr % Wt == 0;        // every thread stays in one column: its x taps are computed once
  WbAxis ax = wb_axis(tid % Wt, rw, a.Ws);
  for (int attempt = 0; attempt < 2; ++attempt) {
    WbInvArea A;
    if (fits) {
      A.x0 = ix0; A.y0 = iy0; A.w = aw; A.h = ah; A.wx0 = ix0; A.wy0 = iy0; A.ww = aw;
      A.winner = s_win; A.vx = s_vx; A.vy = s_vy; A.level = s_lv; A.eroded = s_er;
      for (int i = tid; i < aw * ah; i += nthr) { s_lv[i] = 255; s_er[i] = 0; s_win[i] = INT_MAX; }
    } else {
      A.x0 = 0; A.y0 = 0; A.w = Wp; A.h = Hp; A.wx0 = m; A.wy0 = m; A.ww = Wt;
      A.winner = it.winner; A.vx = it.vx; A.vy = it.vy; A.level = it.level; A.eroded = it.eroded;
      for (int i = tid; i < PP; i += nthr) { it.level[i] = 255; it.eroded[i] = 0; it.vx[i] = 0.f; it.vy[i] = 0.f; if (i < P) it.winner[i] = INT_MAX; }
    }
    __syncthreads();
    // claim: landing cell of every sample; the lowest sample index takes the cell            warp.py:76-88,113-117
    WbWalk wk = wb_walk(tid, nthr, Wt);
    for (int s = tid; s < P; s += nthr, wb_walk_next(wk)) {
      const int Y = wk.Y, X = wk.X;
      if (!xconst) ax = wb_axis(X, rw, a.Ws);
      float dx, dy;
      wb_inv_disp_s(a, s_d, ax, wb_axis(Y, rh, a.Hs), dx, dy);
      const float fx = rintf(__fadd_rn((float)X, dx)), fy = rintf(__fadd_rn((float)Y, dy));   // half-to-even
      int cell = -1;
      if (fx >= 0.f && fy >= 0.f && fx <= (float)(Wt - 1) && fy <= (float)(Ht - 1)) {
        const int cx = (int)fx, cy = (int)fy, px = cx + m, py = cy + m;
        cell = cy * Wt + cx;
        if (px < A.x0 || px >= A.x0 + A.w || py < A.y0 || py >= A.y0 + A.h) s_over = 1;   // outside the predicted box
        else atomicMin(&A.winner[(py - A.wy0) * A.ww + (px - A.wx0)], s);
      }
      it.field[s] = cell;
    }
    __syncthreads();
    if (!(fits && s_over)) {
      // deposit: winners write the negated displacement                                       warp.py:121-123
      int hx0 = INT_MAX, hy0 = INT_MAX, hx1 = -1, hy1 = -1;
      {   // (walks the winner map -- the landing box -- not the samples: a quarter of the trips at the benchmark shape)
        const int nwin = fits ? aw * ah : P;
        WbWalk ww = wb_walk(tid, nthr, A.ww);
        for (int i = tid; i < nwin; i += nthr, wb_walk_next(ww)) {
          const int s = A.winner[i];
          if (s == INT_MAX) continue;
          const int px = A.wx0 + ww.X, py = A.wy0 + ww.Y;
          const int Y = wb_div_small(s, Wt, invWt), X = s - Y * Wt;
          float dx, dy;
          wb_inv_disp_s(a, s_d, wb_axis(X, rw, a.Ws), wb_axis(Y, rh, a.Hs), dx, dy);
          const int c = (py - A.y0) * A.w + (px - A.x0);
          A.vx[c] = -dx; A.vy[c] = -dy; A.level[c] = 0;
          hx0 = min(hx0, px); hy0 = min(hy0, py); hx1 = max(hx1, px); hy1 = max(hy1, py);
        }
      }
      hx0 = wb_warp_min(hx0); hy0 = wb_warp_min(hy0); hx1 = wb_warp_max(hx1); hy1 = wb_warp_max(hy1);
      if (wb_lane() == 0 && hx1 >= 0) { atomicMin(&s_hit[0], hx0); atomicMin(&s_hit[1], hy0); atomicMax(&s_hit[2], hx1); atomicMax(&s_hit[3], hy1); }
      __syncthreads();
      for (int i = tid; i < 4; i += nthr) it.bbox[i] = s_hit[i];
      // dilations, erosions over the hit box grown by the iteration count                     warp.py:135-162
      int x0, y0, w, cells;
      for (int iter = 1; iter <= a.niter; ++iter) {
        wb_inv_box_area(s_hit, iter, A, x0, y0, w, cells);
        if (cells) { WbWalk wc = wb_walk(tid, nthr, w); for (int i = tid; i < cells; i += nthr, wb_walk_next(wc)) wb_inv_dilate_area(A, g, x0 + wc.X, y0 + wc.Y, iter, Hp, Wp); }
        __syncthreads();
      }
      if (a.erode) {
        wb_inv_box_area(s_hit, a.niter, A, x0, y0, w, cells);
        for (int iter = 1; iter <= a.niter; ++iter) {
          if (cells) { WbWalk wc = wb_walk(tid, nthr, w); for (int i = tid; i < cells; i += nthr, wb_walk_next(wc)) wb_inv_erode_area(A, x0 + wc.X, y0 + wc.Y, iter, Hp, Wp); }
          __syncthreads();
        }
      }
      // final: sentinel for unknown cells, crop, back to normalised coordinates                warp.py:164-174
      float* out = a.out + (size_t)blockIdx.x * P * 2;
      wk = wb_walk(tid, nthr, Wt);
      for (int s = tid; s < P; s += nthr, wb_walk_next(wk)) {
        const int px = wk.X + m, py = wk.Y + m;
        bool known = false;
        int c = 0;
        const bool in = px >= A.x0 && px < A.x0 + A.w && py >= A.y0 && py < A.y0 + A.h;
        if (in) {
          c = (py - A.y0) * A.w + (px - A.x0);
          known = A.level[c] != 255 && A.eroded[c] == 0;
        }
        const float ix = known ? A.vx[c] : (float)(2 * Wt), iy = known ? A.vy[c] : (float)(2 * Ht);
        out[2 * s] = __fadd_rn(__ldg(a.id_tgt + 2 * s), __fdiv_rn(__fmul_rn(ix, 2.f), (float)Wt));
        out[2 * s + 1] = __fadd_rn(__ldg(a.id_tgt + 2 * s + 1), __fdiv_rn(__fmul_rn(iy, 2.f), (float)Ht));
        if (fits) it.winner[s] = in ? s_win[c] : INT_MAX;   // (shared-memory winner map: same indexing as the area)
      }
      if (fits) {   // the byte maps the backward (and the index-map parity checks) read
        WbWalk wp = wb_walk(tid, nthr, Wp);
        for (int i = tid; i < PP; i += nthr, wb_walk_next(wp)) {
          const int y = wp.Y, x = wp.X;
          const bool in = x >= A.x0 && x < A.x0 + A.w && y >= A.y0 && y < A.y0 + A.h;
          const int c = in ? (y - A.y0) * A.w + (x - A.x0) : 0;
          it.level[i] = in ? s_lv[c] : (uint8_t)255;
          it.eroded[i] = in ? s_er[c] : (uint8_t)0;
        }
      }
      return;
    }
    // a sample landed outside the predicted box: redo the item with the work area in global memory
    fits = false;
    __syncthreads();
    if (tid == 0) s_over = 0;
    __syncthreads();
  }
}
// phase 5: sentinel for unknown cells, crop, back to normalised coordinates              warp.py:164-174
__global__ void __launch_bounds__(256) k_inv_final(WbInvArgs a) {
  WB_INV_GEOM;
  const WbInvItem it = wb_inv_item(a, blockIdx.y, P, PP);
  float* out = a.out + (size_t)blockIdx.y * P * 2;
  for (int s = blockIdx.x * wb_nthr() + wb_tid(); s < P; s += gridDim.x * wb_nthr()) {
    int Y = s / Wt, X = s - Y * Wt;
    int pc = (Y + m) * Wp + X + m;
    bool known = it.level[pc] != 255 && it.eroded[pc] == 0;
    float ix = known ? it.vx[pc] : (float)(2 * Wt), iy = known ? it.vy[pc] : (float)(2 * Ht);
    out[2 * s] = __fadd_rn(__ldg(a.id_tgt + 2 * s), __fdiv_rn(__fmul_rn(ix, 2.f), (float)Wt));
    out[2 * s + 1] = __fadd_rn(__ldg(a.id_tgt + 2 * s + 1), __fdiv_rn(__fmul_rn(iy, 2.f), (float)Ht));
  }
}

// Backward of the values scattered and filled by the inverse warp.
struct WbInvBwdArgs {
  int n, Hs, Ws, Ht, Wt, niter;
  const float* gauss; const float* dout; const int32_t* field; const int32_t* winner;
  const uint8_t* level; const uint8_t* eroded; const int32_t* bbox;
  float* gval; float* inv_sw; float* gdisp; float* dfwd;
};

// Same phase-per-kernel structure as the forward; grid = (ctas, n).
struct WbInvBItem {
  const uint8_t* level; const uint8_t* eroded; const int32_t* field; const int32_t* winner; const int32_t* bbox;
  float* gx; float* gy; float* isw; float* gdisp; const float* dout;
};
WB_DEV WbInvBItem wb_invb_item(const WbInvBwdArgs& a, int item, int P, int PP) {
  WbInvBItem r;
  r.level = a.level + (size_t)item * PP; r.eroded = a.eroded + (size_t)item * PP;
  r.field = a.field + (size_t)item * P; r.winner = a.winner + (size_t)item * P;
  r.bbox = a.bbox + (size_t)item * 4;
  r.gx = a.gval + (size_t)item * 2 * PP; r.gy = r.gx + PP;
  r.isw = a.inv_sw + (size_t)item * PP;
  r.gdisp = a.gdisp + (size_t)item * P * 2;
  r.dout = a.dout + (size_t)item * P * 2;
  return r;
}
// own gradient of every padded cell + 1/sum-of-weights of the filled ones.  grid = (bands of WB_INV_ROWS padded rows, n): a CTA
// walks only the cells of its band inside the hit cells' bounding box grown by niter -- the level kernels and the handoff never
// read gx / gy / isw outside it (a level-lv cell lies within lv of a hit cell and reads neighbours one further out).
__global__ void __launch_bounds__(256) k_invb_init(WbInvBwdArgs a) {
  WB_INV_GEOM;
  const WbInvBItem it = wb_invb_item(a, blockIdx.y, P, PP);
  float g[9];
  WB_UNROLL for (int i = 0; i < 9; ++i) g[i] = __ldg(a.gauss + i);
  const WbInvBand band = wb_inv_band(it.bbox, a.niter, Hp, Wp);
  for (int i = wb_tid(); i < band.cells; i += wb_nthr()) {
    const int y = band.y0 + i / band.w, x = band.x0 + i % band.w, c = y * Wp + x;
    float ox = 0.f, oy = 0.f, sw = 0.f;
    const int Y = y - m, X = x - m;
    const int lv = it.level[c];
    if (Y >= 0 && Y < Ht && X >= 0 && X < Wt && lv != 255 && it.eroded[c] == 0) {
      ox = it.dout[2 * (Y * Wt + X)] * 2.f / (float)Wt;
      oy = it.dout[2 * (Y * Wt + X) + 1] * 2.f / (float)Ht;
    }
    if (lv != 255 && lv > 0) {
      WB_UNROLL for (int dy = -1; dy <= 1; ++dy)
        WB_UNROLL for (int dx = -1; dx <= 1; ++dx) {
          int yy = y + dy, xx = x + dx;
          if (yy < 0 || yy >= Hp || xx < 0 || xx >= Wp) continue;
          if (wb_level_known_before(it.level[yy * Wp + xx], lv)) sw += g[(dy + 1) * 3 + dx + 1];
        }
    }
    it.gx[c] = ox; it.gy[c] = oy;
    it.isw[c] = sw > 0.f ? 1.f / sw : 0.f;
  }
}
// level `lv` gathers from the (complete) totals of the later-filled neighbours; run for lv = niter-1 .. 0
WB_DEV void wb_invb_level_cell(const WbInvBItem& it, const float* g, int x, int y, int lv, int Hp, int Wp) {
  const int c = y * Wp + x;
  if ((int)it.level[c] != lv) return;
  float ax = 0.f, ay = 0.f;
  WB_UNROLL for (int dy = -1; dy <= 1; ++dy)
    WB_UNROLL for (int dx = -1; dx <= 1; ++dx) {
      int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= Hp || xx < 0 || xx >= Wp) continue;
      int q = yy * Wp + xx;
      int lq = it.level[q];
      if (lq == 255 || lq <= lv) continue;
      // cell q (filled at lq) read cell c through kernel tap (c - q) = (-dy,-dx)
      float w = g[(1 - dy) * 3 + (1 - dx)] * it.isw[q];
      ax += w * it.gx[q]; ay += w * it.gy[q];
    }
  it.gx[c] += ax; it.gy[c] += ay;
}
__global__ void __launch_bounds__(256) k_invb_level(WbInvBwdArgs a, int lv) {
  WB_INV_GEOM;
  const WbInvBItem it = wb_invb_item(a, blockIdx.y, P, PP);
  float g[9];
  WB_UNROLL for (int i = 0; i < 9; ++i) g[i] = __ldg(a.gauss + i);
  const WbInvBand band = wb_inv_band(it.bbox, lv, Hp, Wp);
  for (int i = wb_tid(); i < band.cells; i += wb_nthr()) wb_invb_level_cell(it, g, band.x0 + i % band.w, band.y0 + i / band.w, lv, Hp, Wp);
}
// small-box items: all levels in one CTA (see k_inv_grow_fused).  grid = (n)
__global__ void __launch_bounds__(512) k_invb_levels_fused(WbInvBwdArgs a) {
  WB_INV_GEOM;
  const WbInvBItem it = wb_invb_item(a, blockIdx.x, P, PP);
  float g[9];
  WB_UNROLL for (int i = 0; i < 9; ++i) g[i] = __ldg(a.gauss + i);
  int x0, y0, w, cells;
  for (int lv = a.niter - 1; lv >= 0; --lv) {
    wb_inv_box(it.bbox, lv, Hp, Wp, x0, y0, w, cells);
    for (int i = wb_tid(); i < cells; i += wb_nthr()) wb_invb_level_cell(it, g, x0 + i % w, y0 + i / w, lv, Hp, Wp);
    __syncthreads();
  }
}
// hit cells hand their total to the winning sample: val = -dx, dx = disp * Wt / 2
__global__ void __launch_bounds__(256) k_invb_handoff(WbInvBwdArgs a) {
  WB_INV_GEOM;
  const WbInvBItem it = wb_invb_item(a, blockIdx.y, P, PP);
  for (int s = blockIdx.x * wb_nthr() + wb_tid(); s < P; s += gridDim.x * wb_nthr()) {
    int cell = it.field[s];
    float ox = 0.f, oy = 0.f;
    if (cell >= 0 && it.winner[cell] == s) {
      int cy = cell / Wt, cx = cell - cy * Wt;
      int pc = (cy + m) * Wp + cx + m;
      ox = -it.gx[pc] * (float)Wt * 0.5f; oy = -it.gy[pc] * (float)Ht * 0.5f;
    }
    it.gdisp[2 * s] = ox; it.gdisp[2 * s + 1] = oy;
  }
}
// transpose of the bilinear resize Hs x Ws -> Ht x Wt, as an ordered gather per source point
__global__ void __launch_bounds__(256) k_invb_resize_t(WbInvBwdArgs a) {
  WB_INV_GEOM;
  const WbInvBItem it = wb_invb_item(a, blockIdx.y, P, PP);
  const float* gdisp = it.gdisp;
  const float rh = (float)a.Hs / (float)Ht, rw = (float)a.Ws / (float)Wt;
  float* dfwd = a.dfwd + (size_t)blockIdx.y * a.Hs * a.Ws * 2;
  for (int i = blockIdx.x * wb_nthr() + wb_tid(); i < a.Hs * a.Ws; i += gridDim.x * wb_nthr()) {
    int sy = i / a.Ws, sx = i - sy * a.Ws;
    int ylo = max(0, (int)floorf(((float)sy - 0.5f) / rh - 0.5f) - 2), yhi = min(Ht - 1, (int)ceilf(((float)sy + 1.5f) / rh - 0.5f) + 2);
    int xlo = max(0, (int)floorf(((float)sx - 0.5f) / rw - 0.5f) - 2), xhi = min(Wt - 1, (int)ceilf(((float)sx + 1.5f) / rw - 0.5f) + 2);
    // the stencil is separable: the x-weights of the window are computed once, not once per row
    constexpr int XW = 24;
    float wxv[XW];
    xhi = min(xhi, xlo + XW - 1);   // window width is 2/rw + 5 <= 21 for scale factors up to 8 (checked by the launcher)
    WB_UNROLL for (int j = 0; j < XW; ++j) {
      wxv[j] = 0.f;
      if (xlo + j <= xhi) {
        WbAxis xa = wb_axis(xlo + j, rw, a.Ws);
        wxv[j] = (xa.i0 == sx ? xa.l0 : 0.f) + (xa.i1 == sx ? xa.l1 : 0.f);
      }
    }
    float ax = 0.f, ay = 0.f;
    for (int Y = ylo; Y <= yhi; ++Y) {
      WbAxis ya = wb_axis(Y, rh, a.Hs);
      float wy = (ya.i0 == sy ? ya.l0 : 0.f) + (ya.i1 == sy ? ya.l1 : 0.f);
      if (wy == 0.f) continue;
      const float* grow = gdisp + 2 * (Y * Wt + xlo);
      WB_UNROLL for (int j = 0; j < XW; ++j) {
        if (wxv[j] != 0.f) {   // (zero weights are skipped exactly as the non-separable form skipped them)
          float w = wxv[j] * wy;
          ax += w * grow[2 * j]; ay += w * grow[2 * j + 1];
        }
      }
    }
    dfwd[2 * i] = ax; dfwd[2 * i + 1] = ay;
  }
}

// ================================================================================ occlusion matrix
// lvd.py:59-68.  One thread per (bt, j, i) entry of the (L x L) matrix.
__global__ void k_occ_fwd(int BT, int No, const float* __restrict__ score, float* __restrict__ occ) {
  const int L = No + 1;
  const long long total = (long long)BT * L * L;
  for (long long e = (long long)blockIdx.x * wb_nthr() + wb_tid(); e < total; e += (long long)gridDim.x * wb_nthr()) {
    int i = (int)(e % L), j = (int)((e / L) % L);
    long long bt = e / (L * L);
    float v;
    if (j == 0) v = 0.f;
    else if (i == 0) v = 1.f;
    else {
      float sj = score[bt * No + j - 1], si = score[bt * No + i - 1];
      float ej = expf(-sj * sj) + 1e-6f, ei = expf(-si * si) + 1e-6f;
      v = ej / (ej + ei) - (i == j ? 0.5f : 0.f);
    }
    occ[e] = v;
  }
}
__global__ void k_occ_bwd(int BT, int No, const float* __restrict__ score, const float* __restrict__ docc,
                          float* __restrict__ dscore) {
  const int L = No + 1;
  const long long total = (long long)BT * No;
  for (long long e = (long long)blockIdx.x * wb_nthr() + wb_tid(); e < total; e += (long long)gridDim.x * wb_nthr()) {
    int j = (int)(e % No);
    long long bt = e / No;
    float sj = score[bt * No + j];
    float xj = expf(-sj * sj), ej = xj + 1e-6f;
    const float* d = docc + bt * L * L;
    float acc = 0.f;
    for (int i = 0; i < No; ++i) {
      float si = score[bt * No + i];
      float ei = expf(-si * si) + 1e-6f;
      float den = (ej + ei) * (ej + ei);
      // occ[j+1][i+1] = ej/(ej+ei): d/dej = ei/den ; occ[i+1][j+1] = ei/(ei+ej): d/dej = -ei/den
      acc += (d[(j + 1) * L + i + 1] - d[(i + 1) * L + j + 1]) * ei / den;
    }
    dscore[e] = acc * (-2.f * sj) * xj;
  }
}
