// f-1, first layer: the consumer of the path's `raw_output`.
// WIF's UNet starts with  to_emb = conv3x3(Cin = 3 + num_lyt + num_obj + 1 [+ disocc] -> embed_dim / 2^(depth-1) = 16), stride 1, padding 1,
// no bias (models/modules/conv.py:9-11, :36, :54), applied to raw_output permuted to (B, Tp, Tc, C, H, W) and flattened to
// B*Tp*Tc images (models/nets/wif.py:33-38).  At 512 x 1024 this layer reads the whole of raw_output (2.7 GB per step) and
// writes 16 planes per image: 3.75 GB for 193 GFLOP, i.e. bound by HBM, not by the tensor pipes -- so it is an implicit GEMM
// on warp-level tensor-core instructions (TF32 mma.sync m16n8k8, fp32 accumulate; TF32 is also what the reference's own
// cuDNN convolution uses on this GPU with torch's default allow_tf32), not a tcgen05 pipeline: M = pixels, N = Cout = 16,
// K = 9 * Cin.  The (b, tc, tp) -> (b, tp, tc) permute of wif.py:33 is folded into the addressing (no 2.7 GB copy).
//
// CTA = 256 threads = 8 warps, output tile 32 x 8 pixels; warp w owns tile row w = two 16-pixel M tiles x Cout / 8 N tiles.
// Shared memory: the input tile + halo, copied as is -- by TMA bulk copies (cp.async.bulk + mbarrier: one copy per (channel, row),
// no load/store-pipe instruction per element) when the rows are 16-byte aligned, else by cp.async -- (the tensor core ignores the
// low 13 mantissa bits of an fp32 pattern: round toward zero; the weights are rounded to nearest once, when staged), [Cin padded to 8][10 rows][36 floats] (channel stride 360 = 8 mod 32:
// the four A-fragment loads of a warp each hit 32 different banks), and all weights [tap][cin][Cout pitch 24 | 40] (the two
// B-fragment loads likewise).  Persistent CTAs: the weights are staged once per CTA.
#pragma once
#include "wb_common.cuh"
#include "../../include/waldo_b200.h"

#define WB_CV_TW 32
#define WB_CV_TH 8
#define WB_CV_ROWS (WB_CV_TH + 2)
// Two staging forms.  VEC (W a multiple of 4, 16-byte aligned planes): the tile's columns [tx0 - 4, tx0 + 36) as ten 16-byte
// cp.async per (channel, row); row pitch 44 floats, channel stride 440 = 24 (mod 32).  Otherwise columns [tx0 - 1, tx0 + 33) with
// 4-byte cp.async; pitch 36, channel stride 360 = 8 (mod 32).  Either way the four A-fragment loads of a warp hit 32 banks.
#define WB_CV_PITCH_V 44
#define WB_CV_PITCH_S 36
#define WB_CV_MAX_CIN 48
#ifndef WB_CV_TMA
#define WB_CV_TMA 1   // 16-byte-aligned tiles are staged by TMA bulk copies + an mbarrier instead of per-thread cp.async
#endif
#define WB_CV_MAX_COUT 48

WB_DEV float wb_tf32(float v) {   // round to nearest, ties away from zero, 10-bit mantissa (cvt.rna.tf32.f32)
#ifdef WB_HOST_EMU
  union { float f; unsigned u; } c;
  c.f = v;
  if ((c.u & 0x7f800000u) == 0x7f800000u) return v;
  c.u = (c.u + 0x1000u) & 0xffffe000u;
  return c.f;
#else
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
#endif
}

// what the tensor core makes of an fp32 bit pattern handed to it as TF32: the low 13 mantissa bits are ignored
WB_DEV float wb_tf32_rz(float v) {
  union { float f; unsigned u; } c;
  c.f = v;
  c.u &= 0xffffe000u;
  return c.f;
}

// out-channel pitch of the staged weights: >= Cout and = 8 or 24 (mod 32), so that the B-fragment loads of a warp hit 32 banks
WB_DEV int wb_cv_wpitch(int Cout) { return Cout <= 24 ? 24 : (Cout <= 40 ? 40 : 56); }

// image of `in` that output image i reads: identity, or (b, tp, tc) <- (b, tc, tp)
WB_DEV int wb_cv_src_image(const waldo_conv3x3_t& p, int i) {
  if (p.Tc <= 0) return i;
  const int tc = i % p.Tc, tp = (i / p.Tc) % p.Tp, b = i / (p.Tc * p.Tp);
  return (b * p.Tc + tc) * p.Tp + tp;
}

// CPT: Cin padded to a multiple of 8, compile-time (0: run-time); NTT: Cout / 8, compile-time (0: run-time)
template <int CPT, int NTT, bool VEC>
__global__ void __launch_bounds__(256, 2) k_conv3x3_fwd(waldo_conv3x3_t p) {
  WB_DYN_SMEM(smem);
  constexpr int WB_CV_PITCH = VEC ? WB_CV_PITCH_V : WB_CV_PITCH_S, WB_CV_CH = WB_CV_ROWS * WB_CV_PITCH;
  constexpr int COL0 = VEC ? 3 : 0;            // staged column of pixel x under tap dx: x + dx + COL0
  const int Cin = p.Cin, Cout = p.Cout, H = p.H, W = p.W;
  const int Cp = CPT > 0 ? CPT : ((Cin + 7) & ~7), WP = wb_cv_wpitch(Cout), NT = NTT > 0 ? NTT : (Cout + 7) / 8;
  float* s_in = smem;                          // [Cp][10][36]
  float* s_w = smem + Cp * WB_CV_CH;           // [9][Cp][WP]
  const int tid = wb_tid(), nthr = wb_nthr();
  for (int i = tid; i < 9 * Cp * WP; i += nthr) {
    const int n = i % WP, ch = (i / WP) % Cp, tap = i / (WP * Cp);
    s_w[i] = (n < Cout && ch < Cin) ? wb_tf32(__ldg(p.weight + ((size_t)n * Cin + ch) * 9 + tap)) : 0.f;
  }
  const int tiles_x = (W + WB_CV_TW - 1) / WB_CV_TW, tiles_y = (H + WB_CV_TH - 1) / WB_CV_TH;
  const long long ntiles = (long long)p.n * tiles_x * tiles_y;
  const size_t HW = (size_t)H * W;
#if !defined(WB_HOST_EMU) && WB_CV_TMA
  __shared__ __align__(8) unsigned long long s_bar;
  unsigned phase = 0u;
  if (VEC && tid == 0) wb_mbar_init(&s_bar, 1);
#endif
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int img = (int)(tile / (tiles_x * tiles_y)), tt = (int)(tile - (long long)img * tiles_x * tiles_y);
    const int ty0 = (tt / tiles_x) * WB_CV_TH, tx0 = (tt % tiles_x) * WB_CV_TW;
    const float* in = p.in + (size_t)wb_cv_src_image(p, img) * Cin * HW;
    float* out = p.out + (size_t)img * Cout * HW;
    __syncthreads();   // weights staged / the previous tile's MMAs done
    if (VEC) {
#if !defined(WB_HOST_EMU) && WB_CV_TMA
      // staging by TMA bulk copies: one copy per (channel, row) of the in-image part of the 40-float row (every offset and length is
      // a multiple of 16 bytes: W % 4 == 0), counted in bytes on an mbarrier; the out-of-image chunks / rows are zeroed by hand.
      // No load/store-pipe instruction per element: the staging of this kernel competes with its fragment loads for that pipe.
      {
        const int gx_lo = max(tx0 - 4, 0), gx_hi = min(tx0 + WB_CV_TW + 4, W);
        const unsigned rowbytes = (unsigned)(gx_hi - gx_lo) * 4u;
        const int d_lo = gx_lo - (tx0 - 4);                       // first in-image float of the staged row
        const int r_lo = max(0, 1 - ty0), r_hi = min(WB_CV_ROWS, H - ty0 + 1);   // rows r with 0 <= ty0 + r - 1 < H
        wb_fence_async_smem();                                     // the previous tile's reads (and zero stores) before the async writes
        if (tid == 0) wb_mbar_expect_tx(&s_bar, (unsigned)(Cin * (r_hi - r_lo)) * rowbytes);
        for (int pr = tid; pr < Cp * WB_CV_ROWS; pr += nthr) {
          const int ch = pr / WB_CV_ROWS, r = pr - ch * WB_CV_ROWS;
          float* dst = s_in + ch * WB_CV_CH + r * WB_CV_PITCH;
          if (ch < Cin && r >= r_lo && r < r_hi) {
            wb_bulk_g2s(dst + d_lo, in + (size_t)ch * HW + (size_t)(ty0 + r - 1) * W + gx_lo, rowbytes, &s_bar);
            for (int c = 0; c < d_lo; ++c) dst[c] = 0.f;                                   // left of the image
            for (int c = d_lo + (gx_hi - gx_lo); c < WB_CV_TW + 8; ++c) dst[c] = 0.f;      // right of the image
          } else {
            for (int c = 0; c < WB_CV_TW + 8; ++c) dst[c] = 0.f;                           // padding channel / row outside the image
          }
        }
        wb_mbar_wait(&s_bar, phase);
        phase ^= 1u;
      }
#else
      // staging: 16-byte chunks, 10 per (channel, row); a chunk is wholly inside or wholly outside the image (W % 4 == 0)
      for (int id = tid; id < Cp * WB_CV_ROWS * 10; id += nthr) {
        const int ch = id / (WB_CV_ROWS * 10), rem = id - ch * (WB_CV_ROWS * 10), r = rem / 10, j4 = rem - r * 10;
        const int gy = ty0 + r - 1, gx = tx0 - 4 + 4 * j4;
        const bool ok = ch < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W;
        float* dst = s_in + ch * WB_CV_CH + r * WB_CV_PITCH + 4 * j4;
#ifdef WB_HOST_EMU
        for (int e = 0; e < 4; ++e) dst[e] = ok ? wb_tf32_rz(in[(size_t)ch * HW + (size_t)gy * W + gx + e]) : 0.f;
#else
        wb_cp16z(dst, ok ? in + (size_t)ch * HW + (size_t)gy * W + gx : in, ok);
#endif
      }
#endif
    } else {
    // staging: one (channel, row) of 34 floats per warp and trip (lanes = columns; lanes 0, 1 also take columns 32, 33)
    {
      const int ws = nthr < 32 ? nthr : 32;   // (the host emulation runs one thread per CTA)
      const int nw = (nthr + ws - 1) / ws, wi = tid / ws, ln = tid % ws;
      for (int pr = wi; pr < Cp * WB_CV_ROWS; pr += nw) {
        const int ch = pr / WB_CV_ROWS, r = pr - ch * WB_CV_ROWS;
        const int gy = ty0 + r - 1;
        const bool rok = ch < Cin && gy >= 0 && gy < H;
        const float* src = in + (size_t)ch * HW + (size_t)(rok ? gy : 0) * W;
        float* dst = s_in + ch * WB_CV_CH + r * WB_CV_PITCH;
        for (int c = ln; c < WB_CV_TW + 2; c += ws) {
          const int gx = tx0 + c - 1;
          const bool ok = rok && gx >= 0 && gx < W;
#ifdef WB_HOST_EMU
          dst[c] = ok ? wb_tf32_rz(src[gx]) : 0.f;
#else
          wb_cp4z(dst + c, ok ? src + gx : in, ok);   // LDGSTS: no register staging, every copy of the tile in flight at once
#endif
        }
      }
    }
    }
#ifndef WB_HOST_EMU
    wb_cp_commit();
    wb_cp_wait<0>();
#endif
    __syncthreads();
#ifdef WB_HOST_EMU
    // one host thread per CTA: the same sums (TF32-rounded operands, fp32 accumulation), pixel by pixel
    for (int y = 0; y < WB_CV_TH; ++y)
      for (int x = 0; x < WB_CV_TW; ++x) {
        if (ty0 + y >= H || tx0 + x >= W) continue;
        for (int n = 0; n < Cout; ++n) {
          float acc = 0.f;
          for (int tap = 0; tap < 9; ++tap)
            for (int ch = 0; ch < Cp; ++ch)
              acc += s_in[ch * WB_CV_CH + (y + tap / 3) * WB_CV_PITCH + x + tap % 3 + COL0] * s_w[(tap * Cp + ch) * WP + n];
          out[(size_t)n * HW + (size_t)(ty0 + y) * W + tx0 + x] = acc;
        }
      }
#else
    const int lane = tid & 31, w = tid >> 5, gid = lane >> 2, tig = lane & 3;
    float acc[2][WB_CV_MAX_COUT / 8][4];
    WB_UNROLL for (int mt = 0; mt < 2; ++mt)
      WB_UNROLL for (int nt = 0; nt < WB_CV_MAX_COUT / 8; ++nt)
        WB_UNROLL for (int j = 0; j < 4; ++j) acc[mt][nt][j] = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap - dy * 3;
      const float* arow = s_in + tig * WB_CV_CH + (w + dy) * WB_CV_PITCH + dx + gid + COL0;
      const float* brow = s_w + (tap * Cp + tig) * WP + gid;
      WB_PRAGMA(unroll (CPT > 0 ? CPT / 8 : 1))
      for (int c0 = 0; c0 < Cp; c0 += 8) {
        unsigned a[2][4];
        WB_UNROLL for (int mt = 0; mt < 2; ++mt) {
          const float* q = arow + c0 * WB_CV_CH + mt * 16;
          a[mt][0] = __float_as_uint(q[0]); a[mt][1] = __float_as_uint(q[8]);
          a[mt][2] = __float_as_uint(q[4 * WB_CV_CH]); a[mt][3] = __float_as_uint(q[4 * WB_CV_CH + 8]);
        }
        WB_UNROLL for (int nt = 0; nt < WB_CV_MAX_COUT / 8; ++nt) {
          if (nt < NT) {
            const unsigned b0 = __float_as_uint(brow[c0 * WP + nt * 8]), b1 = __float_as_uint(brow[(c0 + 4) * WP + nt * 8]);
            WB_UNROLL for (int mt = 0; mt < 2; ++mt)
              asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                           : "+f"(acc[mt][nt][0]), "+f"(acc[mt][nt][1]), "+f"(acc[mt][nt][2]), "+f"(acc[mt][nt][3])
                           : "r"(a[mt][0]), "r"(a[mt][1]), "r"(a[mt][2]), "r"(a[mt][3]), "r"(b0), "r"(b1));
          }
        }
      }
    }
    const int y = ty0 + w;
    if (y < H) {
      WB_UNROLL for (int mt = 0; mt < 2; ++mt)
        WB_UNROLL for (int nt = 0; nt < WB_CV_MAX_COUT / 8; ++nt) {
          if (nt < NT) {
            const int x0 = tx0 + mt * 16 + gid, n0 = nt * 8 + 2 * tig;
            float* o = out + (size_t)n0 * HW + (size_t)y * W;
            const bool c0ok = n0 < Cout, c1ok = n0 + 1 < Cout;   // (Cout need not be a multiple of 8: the staged weights beyond it are 0)
            if (x0 < W) { if (c0ok) o[x0] = acc[mt][nt][0]; if (c1ok) o[HW + x0] = acc[mt][nt][1]; }
            if (x0 + 8 < W) { if (c0ok) o[x0 + 8] = acc[mt][nt][2]; if (c1ok) o[HW + x0 + 8] = acc[mt][nt][3]; }
          }
        }
    }
#endif
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Weight gradient of the same layer:  dW[co][ci][tap] = sum over images and pixels of dY[co][y][x] * X[ci][y + dy - 1][x + dx - 1].
// As a GEMM: M = Cout = 16 (one M tile), N = 9 * Cin columns (ci, tap), K = all pixels of all images.  Persistent CTAs walk the
// same 32 x 8 tiles as the forward: X tile + halo and the dY tile are staged by cp.async, each warp owns N tiles j = warp,
// warp + 8, ... (8 columns = 8 consecutive input channels under one tap) and keeps their 16 x 8 accumulators in registers over
// ALL tiles of the CTA; K advances 8 pixels of one tile row per MMA.  Plane strides = 4 (mod 32) floats: the A (dY) and B (X)
// fragment loads of a warp each hit 32 banks.  Each CTA writes its partial dW once; k_conv3x3_wgrad_final adds the partials in
// CTA order (deterministic).  TF32 products (both operands truncated by the tensor core), fp32 accumulation.
#define WB_CW_XS (WB_CV_ROWS * WB_CV_PITCH_V + 12)   // 452: X plane stride
#define WB_CW_YS (WB_CV_TH * WB_CV_TW + 4)          // 260: dY plane stride
#define WB_CW_MAX_NT 6                               // N tiles per warp: 9 * 48 / 8 = 54 tiles over 8 warps -> at most 7; capped by Cin <= 40 here

template <int CPT>
__global__ void __launch_bounds__(256, 2) k_conv3x3_wgrad(waldo_conv3x3_wgrad_t p) {
  WB_DYN_SMEM(smem);
  const int Cin = p.c.Cin, Cout = p.c.Cout, H = p.c.H, W = p.c.W;
  const int Cp = CPT > 0 ? CPT : ((Cin + 7) & ~7), NTt = 9 * (Cp / 8);
  float* s_x = smem;                       // [Cp][452]
  float* s_y = smem + Cp * WB_CW_XS;       // [16][260]
  const int tid = wb_tid(), nthr = wb_nthr();
  const int tiles_x = (W + WB_CV_TW - 1) / WB_CV_TW, tiles_y = (H + WB_CV_TH - 1) / WB_CV_TH, tpi = tiles_x * tiles_y;
  const long long ntiles = (long long)p.c.n * tpi;
  const size_t HW = (size_t)H * W;
  float* part = p.part + (size_t)blockIdx.x * Cout * Cin * 9;
#if !defined(WB_HOST_EMU) && WB_CV_TMA
  __shared__ __align__(8) unsigned long long s_bar;
  unsigned phase = 0u;
  if (tid == 0) wb_mbar_init(&s_bar, 1);
#endif
#ifdef WB_HOST_EMU
  for (int i = 0; i < Cout * Cin * 9; ++i) part[i] = 0.f;
#else
  const int lane = tid & 31, wp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  float acc[WB_CW_MAX_NT][4];
  WB_UNROLL for (int j = 0; j < WB_CW_MAX_NT; ++j) WB_UNROLL for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#endif
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int img = (int)(tile / tpi), tt = (int)(tile - (long long)img * tpi);
    const int ty0 = (tt / tiles_x) * WB_CV_TH, tx0 = (tt % tiles_x) * WB_CV_TW;
    const float* in = p.c.in + (size_t)wb_cv_src_image(p.c, img) * Cin * HW;
    const float* dy = p.dout + (size_t)img * Cout * HW;
    __syncthreads();   // the previous tile's MMAs are done
#if !defined(WB_HOST_EMU) && WB_CV_TMA
    {   // X tile + halo and the dY tile by TMA bulk copies, one per (plane, row), counted on one mbarrier (see k_conv3x3_fwd)
      const int gx_lo = max(tx0 - 4, 0), gx_hi = min(tx0 + WB_CV_TW + 4, W);
      const unsigned xbytes = (unsigned)(gx_hi - gx_lo) * 4u;
      const int d_lo = gx_lo - (tx0 - 4);
      const int r_lo = max(0, 1 - ty0), r_hi = min(WB_CV_ROWS, H - ty0 + 1);
      const int yw = min(WB_CV_TW, W - tx0), yr = min(WB_CV_TH, H - ty0);   // in-image part of the dY tile
      const unsigned ybytes = (unsigned)yw * 4u;
      wb_fence_async_smem();
      if (tid == 0) wb_mbar_expect_tx(&s_bar, (unsigned)(Cin * (r_hi - r_lo)) * xbytes + (unsigned)(Cout * yr) * ybytes);
      for (int pr = tid; pr < Cp * WB_CV_ROWS; pr += nthr) {
        const int ch = pr / WB_CV_ROWS, r = pr - ch * WB_CV_ROWS;
        float* dst = s_x + ch * WB_CW_XS + r * WB_CV_PITCH_V;
        if (ch < Cin && r >= r_lo && r < r_hi) {
          wb_bulk_g2s(dst + d_lo, in + (size_t)ch * HW + (size_t)(ty0 + r - 1) * W + gx_lo, xbytes, &s_bar);
          for (int c = 0; c < d_lo; ++c) dst[c] = 0.f;
          for (int c = d_lo + (gx_hi - gx_lo); c < WB_CV_TW + 8; ++c) dst[c] = 0.f;
        } else {
          for (int c = 0; c < WB_CV_TW + 8; ++c) dst[c] = 0.f;
        }
      }
      for (int pr = tid; pr < 16 * WB_CV_TH; pr += nthr) {
        const int co = pr / WB_CV_TH, r = pr - co * WB_CV_TH;
        float* dst = s_y + co * WB_CW_YS + r * WB_CV_TW;
        if (co < Cout && r < yr) {
          wb_bulk_g2s(dst, dy + (size_t)co * HW + (size_t)(ty0 + r) * W + tx0, ybytes, &s_bar);
          for (int c = yw; c < WB_CV_TW; ++c) dst[c] = 0.f;
        } else {
          for (int c = 0; c < WB_CV_TW; ++c) dst[c] = 0.f;
        }
      }
      wb_mbar_wait(&s_bar, phase);
      phase ^= 1u;
    }
    __syncthreads();
#else
    for (int id = tid; id < Cp * WB_CV_ROWS * 10; id += nthr) {   // X tile + halo, 16-byte chunks (W % 4 == 0)
      const int ch = id / (WB_CV_ROWS * 10), rem = id - ch * (WB_CV_ROWS * 10), r = rem / 10, j4 = rem - r * 10;
      const int gy = ty0 + r - 1, gx = tx0 - 4 + 4 * j4;
      const bool ok = ch < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W;
      float* dst = s_x + ch * WB_CW_XS + r * WB_CV_PITCH_V + 4 * j4;
#ifdef WB_HOST_EMU
      for (int e = 0; e < 4; ++e) dst[e] = ok ? wb_tf32_rz(in[(size_t)ch * HW + (size_t)gy * W + gx + e]) : 0.f;
#else
      wb_cp16z(dst, ok ? in + (size_t)ch * HW + (size_t)gy * W + gx : in, ok);
#endif
    }
    for (int id = tid; id < 16 * WB_CV_TH * 8; id += nthr) {       // dY tile (16 planes: zeros beyond Cout), 16-byte chunks
      const int co = id / (WB_CV_TH * 8), rem = id - co * (WB_CV_TH * 8), r = rem / 8, j4 = rem - r * 8;
      const int gy = ty0 + r, gx = tx0 + 4 * j4;
      const bool ok = co < Cout && gy < H && gx < W;
      float* dst = s_y + co * WB_CW_YS + r * WB_CV_TW + 4 * j4;
#ifdef WB_HOST_EMU
      for (int e = 0; e < 4; ++e) dst[e] = ok ? wb_tf32_rz(dy[(size_t)co * HW + (size_t)gy * W + gx + e]) : 0.f;
#else
      wb_cp16z(dst, ok ? dy + (size_t)co * HW + (size_t)gy * W + gx : dy, ok);
#endif
    }
#ifndef WB_HOST_EMU
    wb_cp_commit();
    wb_cp_wait<0>();
#endif
    __syncthreads();
#endif
#ifdef WB_HOST_EMU
    for (int co = 0; co < Cout; ++co)
      for (int ci = 0; ci < Cin; ++ci)
        for (int tap = 0; tap < 9; ++tap) {
          float a = 0.f;
          for (int y = 0; y < WB_CV_TH; ++y)
            for (int x = 0; x < WB_CV_TW; ++x)
              a += s_y[co * WB_CW_YS + y * WB_CV_TW + x] * s_x[ci * WB_CW_XS + (y + tap / 3) * WB_CV_PITCH_V + x + tap % 3 + 3];
          part[((size_t)co * Cin + ci) * 9 + tap] += a;
        }
#else
#pragma unroll 1
    for (int ks = 0; ks < WB_CV_TH * 4; ++ks) {   // 8 pixels of one tile row per step
      const int y = ks >> 2, x0 = (ks & 3) * 8;
      const float* ay = s_y + gid * WB_CW_YS + y * WB_CV_TW + x0 + tig;
      unsigned a[4];
      a[0] = __float_as_uint(ay[0]); a[1] = __float_as_uint(ay[8 * WB_CW_YS]);
      a[2] = __float_as_uint(ay[4]); a[3] = __float_as_uint(ay[8 * WB_CW_YS + 4]);
      WB_UNROLL for (int j = 0; j < WB_CW_MAX_NT; ++j) {
        const int nt = wp + 8 * j;
        if (nt < NTt) {
          const int tap = nt / (Cp / 8), c0 = (nt - tap * (Cp / 8)) * 8, dyy = tap / 3, dxx = tap - dyy * 3;
          const float* bx = s_x + (c0 + gid) * WB_CW_XS + (y + dyy) * WB_CV_PITCH_V + x0 + tig + dxx + 3;
          const unsigned b0 = __float_as_uint(bx[0]), b1 = __float_as_uint(bx[4]);
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
                       : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
        }
      }
    }
#endif
  }
#ifndef WB_HOST_EMU
  WB_UNROLL for (int j = 0; j < WB_CW_MAX_NT; ++j) {
    const int nt = wp + 8 * j;
    if (nt < NTt) {
      const int tap = nt / (Cp / 8), c0 = (nt - tap * (Cp / 8)) * 8;
      WB_UNROLL for (int e = 0; e < 4; ++e) {
        const int co = gid + (e >> 1) * 8, ci = c0 + 2 * tig + (e & 1);
        if (co < Cout && ci < Cin) part[((size_t)co * Cin + ci) * 9 + tap] = acc[j][e];
      }
    }
  }
#endif
}

// dW = sum of the CTA partials, in CTA order
__global__ void __launch_bounds__(256) k_conv3x3_wgrad_final(waldo_conv3x3_wgrad_t p) {
  const int n = p.c.Cout * p.c.Cin * 9;
  for (int i = blockIdx.x * wb_nthr() + wb_tid(); i < n; i += gridDim.x * wb_nthr()) {
    float acc = 0.f;
    for (int c = 0; c < p.ctas; ++c) acc += p.part[(size_t)c * n + i];
    p.dweight[i] = acc;
  }
}
