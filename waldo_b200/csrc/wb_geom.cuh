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
    for (int p = p0; p < p1; ++p) {
      double r = (double)__ldg(repr + (size_t)p * K + j);
      ax += r * (double)__ldg(dgrid + ((size_t)item * P + p) * 2);
      ay += r * (double)__ldg(dgrid + ((size_t)item * P + p) * 2 + 1);
    }
    double* o = partial + (((size_t)item * chunks + ch) * K + j) * 2;
    o[0] = ax; o[1] = ay;
  }
}
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
  float* out; int32_t* field; int32_t* winner; uint8_t* level; uint8_t* eroded; float* val;
};

WB_DEV bool wb_level_known_before(uint8_t lv, int it) { return lv != 255 && (int)lv < it; }

__global__ void __launch_bounds__(1024) k_invwarp_fwd(WbInvArgs a) {
  const int item = blockIdx.x;
  const int Ht = a.Ht, Wt = a.Wt, P = Ht * Wt, m = a.niter + 1;
  const int Hp = Ht + 2 * m, Wp = Wt + 2 * m, PP = Hp * Wp;
  const float* fwd = a.fwd + (size_t)item * a.Hs * a.Ws * 2;
  int32_t* field = a.field + (size_t)item * P;
  int32_t* winner = a.winner + (size_t)item * P;
  uint8_t* level = a.level + (size_t)item * PP;
  uint8_t* eroded = a.eroded + (size_t)item * PP;
  float* vx = a.val + (size_t)item * 2 * PP;
  float* vy = vx + PP;
  const float rh = (float)a.Hs / (float)Ht, rw = (float)a.Ws / (float)Wt;   // area_pixel_compute_scale

  // phase 0: clear
  for (int i = wb_tid(); i < PP; i += wb_nthr()) { level[i] = 255; eroded[i] = 0; vx[i] = 0.f; vy[i] = 0.f; }
  for (int i = wb_tid(); i < P; i += wb_nthr()) winner[i] = INT_MAX;
  __syncthreads();
  // phase 1: landing cell of every lattice sample; lowest sample index claims the cell   warp.py:76-88,113-117
  for (int s = wb_tid(); s < P; s += wb_nthr()) {
    int Y = s / Wt, X = s - Y * Wt;
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
    float dx = __fdiv_rn(__fmul_rn(d[0], (float)Wt), 2.f), dy = __fdiv_rn(__fmul_rn(d[1], (float)Ht), 2.f);
    float fx = rintf(__fadd_rn((float)X, dx)), fy = rintf(__fadd_rn((float)Y, dy));   // half-to-even
    int cell = -1;
    if (fx >= 0.f && fy >= 0.f && fx <= (float)(Wt - 1) && fy <= (float)(Ht - 1)) {
      cell = (int)fy * Wt + (int)fx;
      atomicMin(&winner[cell], s);
    }
    field[s] = cell;
    // park -dx,-dy at the sample's own padded slot?  no: recomputed below from the winner only
  }
  __syncthreads();
  // phase 2: winners deposit the negated displacement                                       warp.py:121-123
  for (int s = wb_tid(); s < P; s += wb_nthr()) {
    int cell = field[s];
    if (cell < 0 || winner[cell] != s) continue;
    int Y = s / Wt, X = s - Y * Wt;
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
    float dx = __fdiv_rn(__fmul_rn(d[0], (float)Wt), 2.f), dy = __fdiv_rn(__fmul_rn(d[1], (float)Ht), 2.f);
    int cy = cell / Wt, cx = cell - cy * Wt;
    int pc = (cy + m) * Wp + cx + m;
    vx[pc] = -dx; vy[pc] = -dy; level[pc] = 0;
  }
  __syncthreads();
  // phase 3: grow `niter` rings; a frontier cell takes the normalised Gaussian mean of known cells   warp.py:135-151
  float g[9];
  WB_UNROLL for (int i = 0; i < 9; ++i) g[i] = __ldg(a.gauss + i);
  for (int it = 1; it <= a.niter; ++it) {
    for (int c = wb_tid(); c < PP; c += wb_nthr()) {
      if (level[c] != 255) continue;
      int y = c / Wp, x = c - y * Wp;
      bool front = (y > 0 && wb_level_known_before(level[c - Wp], it)) || (y < Hp - 1 && wb_level_known_before(level[c + Wp], it)) ||
                   (x > 0 && wb_level_known_before(level[c - 1], it)) || (x < Wp - 1 && wb_level_known_before(level[c + 1], it));
      if (!front) continue;
      float sx = 0.f, sy = 0.f, sw = 0.f;
      WB_UNROLL for (int dy = -1; dy <= 1; ++dy)
        WB_UNROLL for (int dx = -1; dx <= 1; ++dx) {
          int yy = y + dy, xx = x + dx;
          if (yy < 0 || yy >= Hp || xx < 0 || xx >= Wp) continue;
          int q = yy * Wp + xx;
          if (!wb_level_known_before(level[q], it)) continue;
          float w = g[(dy + 1) * 3 + dx + 1];
          sx += w * vx[q]; sy += w * vy[q]; sw += w;
        }
      vx[c] = sx / sw; vy[c] = sy / sw; level[c] = (uint8_t)it;
    }
    __syncthreads();
  }
  // phase 4: erosion of the known set (objects only)                                       warp.py:153-162
  if (a.erode) {
    for (int it = 1; it <= a.niter; ++it) {
      for (int c = wb_tid(); c < PP; c += wb_nthr()) {
        if (level[c] == 255 || eroded[c] != 0) continue;
        int y = c / Wp, x = c - y * Wp;
#define WB_GONE(q) (level[q] == 255 || (eroded[q] != 0 && (int)eroded[q] < it))
        bool edge = (y > 0 && WB_GONE(c - Wp)) || (y < Hp - 1 && WB_GONE(c + Wp)) ||
                    (x > 0 && WB_GONE(c - 1)) || (x < Wp - 1 && WB_GONE(c + 1));
#undef WB_GONE
        if (edge) eroded[c] = (uint8_t)it;
      }
      __syncthreads();
    }
  }
  // phase 5: sentinel for unknown cells, crop, back to normalised coordinates              warp.py:164-174
  float* out = a.out + (size_t)item * P * 2;
  for (int s = wb_tid(); s < P; s += wb_nthr()) {
    int Y = s / Wt, X = s - Y * Wt;
    int pc = (Y + m) * Wp + X + m;
    bool known = level[pc] != 255 && eroded[pc] == 0;
    float ix = known ? vx[pc] : (float)(2 * Wt), iy = known ? vy[pc] : (float)(2 * Ht);
    out[2 * s] = __fadd_rn(__ldg(a.id_tgt + 2 * s), __fdiv_rn(__fmul_rn(ix, 2.f), (float)Wt));
    out[2 * s + 1] = __fadd_rn(__ldg(a.id_tgt + 2 * s + 1), __fdiv_rn(__fmul_rn(iy, 2.f), (float)Ht));
  }
}

// Backward of the values scattered and filled by the inverse warp.
struct WbInvBwdArgs {
  int n, Hs, Ws, Ht, Wt, niter;
  const float* gauss; const float* dout; const int32_t* field; const int32_t* winner;
  const uint8_t* level; const uint8_t* eroded;
  float* gval; float* inv_sw; float* gdisp; float* dfwd;
};

__global__ void __launch_bounds__(1024) k_invwarp_bwd(WbInvBwdArgs a) {
  const int item = blockIdx.x;
  const int Ht = a.Ht, Wt = a.Wt, P = Ht * Wt, m = a.niter + 1;
  const int Hp = Ht + 2 * m, Wp = Wt + 2 * m, PP = Hp * Wp;
  const uint8_t* level = a.level + (size_t)item * PP;
  const uint8_t* eroded = a.eroded + (size_t)item * PP;
  const int32_t* field = a.field + (size_t)item * P;
  const int32_t* winner = a.winner + (size_t)item * P;
  float* gx = a.gval + (size_t)item * 2 * PP;
  float* gy = gx + PP;
  float* isw = a.inv_sw + (size_t)item * PP;
  float* gdisp = a.gdisp + (size_t)item * P * 2;
  const float* dout = a.dout + (size_t)item * P * 2;
  float g[9];
  WB_UNROLL for (int i = 0; i < 9; ++i) g[i] = __ldg(a.gauss + i);
  // own gradient of every padded cell + 1/sum-of-weights of the filled ones
  for (int c = wb_tid(); c < PP; c += wb_nthr()) {
    int y = c / Wp, x = c - y * Wp;
    float ox = 0.f, oy = 0.f;
    int Y = y - m, X = x - m;
    if (Y >= 0 && Y < Ht && X >= 0 && X < Wt && level[c] != 255 && eroded[c] == 0) {
      ox = dout[2 * (Y * Wt + X)] * 2.f / (float)Wt;
      oy = dout[2 * (Y * Wt + X) + 1] * 2.f / (float)Ht;
    }
    gx[c] = ox; gy[c] = oy;
    float sw = 0.f;
    int lv = level[c];
    if (lv != 255 && lv > 0) {
      WB_UNROLL for (int dy = -1; dy <= 1; ++dy)
        WB_UNROLL for (int dx = -1; dx <= 1; ++dx) {
          int yy = y + dy, xx = x + dx;
          if (yy < 0 || yy >= Hp || xx < 0 || xx >= Wp) continue;
          if (wb_level_known_before(level[yy * Wp + xx], lv)) sw += g[(dy + 1) * 3 + dx + 1];
        }
    }
    isw[c] = sw > 0.f ? 1.f / sw : 0.f;
  }
  __syncthreads();
  // levels niter-1 .. 0 gather from the (complete) totals of the later-filled neighbours
  for (int lv = a.niter - 1; lv >= 0; --lv) {
    for (int c = wb_tid(); c < PP; c += wb_nthr()) {
      if ((int)level[c] != lv) continue;
      int y = c / Wp, x = c - y * Wp;
      float ax = 0.f, ay = 0.f;
      WB_UNROLL for (int dy = -1; dy <= 1; ++dy)
        WB_UNROLL for (int dx = -1; dx <= 1; ++dx) {
          int yy = y + dy, xx = x + dx;
          if (yy < 0 || yy >= Hp || xx < 0 || xx >= Wp) continue;
          int q = yy * Wp + xx;
          int lq = level[q];
          if (lq == 255 || lq <= lv) continue;
          // cell q (filled at lq) read cell c through kernel tap (c - q) = (-dy,-dx)
          float w = g[(1 - dy) * 3 + (1 - dx)] * isw[q];
          ax += w * gx[q]; ay += w * gy[q];
        }
      gx[c] += ax; gy[c] += ay;
    }
    __syncthreads();
  }
  // hit cells hand their total to the winning sample: val = -dx, dx = disp * Wt / 2
  for (int s = wb_tid(); s < P; s += wb_nthr()) {
    int cell = field[s];
    float ox = 0.f, oy = 0.f;
    if (cell >= 0 && winner[cell] == s) {
      int cy = cell / Wt, cx = cell - cy * Wt;
      int pc = (cy + m) * Wp + cx + m;
      ox = -gx[pc] * (float)Wt * 0.5f; oy = -gy[pc] * (float)Ht * 0.5f;
    }
    gdisp[2 * s] = ox; gdisp[2 * s + 1] = oy;
  }
  __syncthreads();
  // transpose of the bilinear resize Hs x Ws -> Ht x Wt, as an ordered gather per source point
  const float rh = (float)a.Hs / (float)Ht, rw = (float)a.Ws / (float)Wt;
  float* dfwd = a.dfwd + (size_t)item * a.Hs * a.Ws * 2;
  for (int i = wb_tid(); i < a.Hs * a.Ws; i += wb_nthr()) {
    int sy = i / a.Ws, sx = i - sy * a.Ws;
    int ylo = max(0, (int)floorf(((float)sy - 0.5f) / rh - 0.5f) - 2), yhi = min(Ht - 1, (int)ceilf(((float)sy + 1.5f) / rh - 0.5f) + 2);
    int xlo = max(0, (int)floorf(((float)sx - 0.5f) / rw - 0.5f) - 2), xhi = min(Wt - 1, (int)ceilf(((float)sx + 1.5f) / rw - 0.5f) + 2);
    float ax = 0.f, ay = 0.f;
    for (int Y = ylo; Y <= yhi; ++Y) {
      WbAxis ya = wb_axis(Y, rh, a.Hs);
      float wy = (ya.i0 == sy ? ya.l0 : 0.f) + (ya.i1 == sy ? ya.l1 : 0.f);
      if (wy == 0.f) continue;
      for (int X = xlo; X <= xhi; ++X) {
        WbAxis xa = wb_axis(X, rw, a.Ws);
        float wx = (xa.i0 == sx ? xa.l0 : 0.f) + (xa.i1 == sx ? xa.l1 : 0.f);
        if (wx == 0.f) continue;
        float w = wx * wy;
        ax += w * gdisp[2 * (Y * Wt + X)]; ay += w * gdisp[2 * (Y * Wt + X) + 1];
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
