// Input packing: build the path's `input` tensor on the device from what the dataset really
// holds -- 8-bit RGB and an 8-bit label map -- instead of shipping 3 + Nl fp32 planes over PCIe.
// Reference: data/base_dataset.py:173-183 (label -> one-hot -> 5 * (2x - 1)), :355-372 (ToTensor + Normalize(0.5, 0.5)),
// models/synthesizer.py:444 (input = cat([vid, lyt], dim=2)).
//   input[f, 0:3]   = ((rgb / 255) - 0.5) / 0.5          (or a copy of an fp32 frame that is already normalised)
//   input[f, 3 + c] = label == c ? on : off               (on = 5, off = -5; labels >= Nl give all-off)
// One thread handles 4 consecutive pixels: one 32-bit load per byte plane, one 128-bit store per output plane.
#pragma once
#include "wb_common.cuh"
#include "../../include/waldo_b200.h"

WB_DEV float wb_norm_u8(unsigned v) { return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), 0.5f), 0.5f); }

// four consecutive elements in one store: STG.E.128 (fp32) / STG.E.64 (bf16)
WB_DEV void wb_st4(float* o, float a, float b, float c, float d) { *reinterpret_cast<float4*>(o) = make_float4(a, b, c, d); }
WB_DEV void wb_st4(wb_bf16* o, float a, float b, float c, float d) {
  wb_bf16 t[4];
  wb_sts(t + 0, a); wb_sts(t + 1, b); wb_sts(t + 2, c); wb_sts(t + 3, d);
  uint2 w;
  w.x = (unsigned)t[0].x | ((unsigned)t[1].x << 16); w.y = (unsigned)t[2].x | ((unsigned)t[3].x << 16);
  *reinterpret_cast<uint2*>(o) = w;
}

template <typename ST>
__global__ void __launch_bounds__(256) k_pack_input(waldo_pack_input_t p) {
  const int C = 3 + p.Nl;
  const size_t HW = (size_t)p.HW;
  const size_t nq = (HW + 3) / 4;                     // groups of 4 pixels per frame
  const bool vec = (HW & 3) == 0;
  const int f = blockIdx.y;
  for (size_t gq = (size_t)blockIdx.x * wb_nthr() + wb_tid(); gq < nq; gq += (size_t)gridDim.x * wb_nthr()) {
    const size_t q = gq * 4;
    const int npx = (int)(HW - q < 4 ? HW - q : 4);
    ST* out = reinterpret_cast<ST*>(p.input) + (size_t)f * C * HW + q;
    unsigned lab[4] = {255u, 255u, 255u, 255u};
    if (vec) {
      const unsigned w = *reinterpret_cast<const unsigned*>(p.label + (size_t)f * HW + q);
      WB_UNROLL for (int i = 0; i < 4; ++i) lab[i] = (w >> (8 * i)) & 255u;
    } else {
      for (int i = 0; i < npx; ++i) lab[i] = p.label[(size_t)f * HW + q + i];
    }
    for (int c = 0; c < 3; ++c) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.rgb_u8) {
        const uint8_t* s = p.rgb_u8 + ((size_t)f * 3 + c) * HW + q;
        if (vec) {
          const unsigned w = *reinterpret_cast<const unsigned*>(s);
          WB_UNROLL for (int i = 0; i < 4; ++i) v[i] = wb_norm_u8((w >> (8 * i)) & 255u);
        } else {
          for (int i = 0; i < npx; ++i) v[i] = wb_norm_u8(s[i]);
        }
      } else {
        const float* s = p.rgb_f32 + ((size_t)f * 3 + c) * HW + q;
        for (int i = 0; i < npx; ++i) v[i] = s[i];
      }
      ST* o = out + (size_t)c * HW;
      if (vec) wb_st4(o, v[0], v[1], v[2], v[3]);
      else for (int i = 0; i < npx; ++i) wb_sts(o + i, v[i]);
    }
    for (int c = 0; c < p.Nl; ++c) {
      ST* o = out + (size_t)(3 + c) * HW;
      const float v0 = lab[0] == (unsigned)c ? p.on : p.off, v1 = lab[1] == (unsigned)c ? p.on : p.off;
      const float v2 = lab[2] == (unsigned)c ? p.on : p.off, v3 = lab[3] == (unsigned)c ? p.on : p.off;
      if (vec) wb_st4(o, v0, v1, v2, v3);
      else { const float vv[4] = {v0, v1, v2, v3}; for (int i = 0; i < npx; ++i) wb_sts(o + i, vv[i]); }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// f-4, the output side: tools/utils.py:246-249 `normalize` + :258-264 `dump_video`
//     tensor.clamp(lo, hi) -> (tensor - lo) / (hi - lo) -> permute(0, 2, 3, 1) * 255 -> uint8 (truncation)
// done on the device, so that 1 byte per sample crosses PCIe instead of 4 and the (T, H, W, 3) layout the video writer wants
// is produced by the kernel.  frames (n, 3, HW) fp32 planar -> out (n, HW, 3) uint8.  One thread per 4 consecutive
// pixels: three 128-bit loads (LDG.E.128, one per colour plane), twelve bytes out as three 32-bit stores.
WB_DEV unsigned wb_to_u8(float v, float lo, float hi) {
  // the reference's operation order, one IEEE operation each (no contraction)
  const float c = fminf(fmaxf(v, lo), hi);
  const float u = __fmul_rn(__fdiv_rn(__fsub_rn(c, lo), __fsub_rn(hi, lo)), 255.f);
  return (unsigned)u;   // float -> uint8: truncation toward zero, as Tensor.to(torch.uint8); u is within [0, 255]
}
__global__ void __launch_bounds__(256) k_frames_to_u8(waldo_frames_u8_t p) {
  const size_t HW = (size_t)p.HW;
  const size_t nq = (HW + 3) / 4;
  const bool vec = (HW & 3) == 0;
  const int f = blockIdx.y;
  const float* base = p.frames + (size_t)f * 3 * HW;
  uint8_t* out = p.out + (size_t)f * HW * 3;
  for (size_t gq = (size_t)blockIdx.x * wb_nthr() + wb_tid(); gq < nq; gq += (size_t)gridDim.x * wb_nthr()) {
    const size_t q = gq * 4;
    if (vec) {
      const float4 r = wb_ld4f(base + q), g = wb_ld4f(base + HW + q), b = wb_ld4f(base + 2 * HW + q);
      const unsigned r0 = wb_to_u8(r.x, p.lo, p.hi), g0 = wb_to_u8(g.x, p.lo, p.hi), b0 = wb_to_u8(b.x, p.lo, p.hi);
      const unsigned r1 = wb_to_u8(r.y, p.lo, p.hi), g1 = wb_to_u8(g.y, p.lo, p.hi), b1 = wb_to_u8(b.y, p.lo, p.hi);
      const unsigned r2 = wb_to_u8(r.z, p.lo, p.hi), g2 = wb_to_u8(g.z, p.lo, p.hi), b2 = wb_to_u8(b.z, p.lo, p.hi);
      const unsigned r3 = wb_to_u8(r.w, p.lo, p.hi), g3 = wb_to_u8(g.w, p.lo, p.hi), b3 = wb_to_u8(b.w, p.lo, p.hi);
      unsigned* o = reinterpret_cast<unsigned*>(out + q * 3);   // 12 bytes, 4-byte aligned (q is a multiple of 4)
      o[0] = r0 | (g0 << 8) | (b0 << 16) | (r1 << 24);
      o[1] = g1 | (b1 << 8) | (r2 << 16) | (g2 << 24);
      o[2] = b2 | (r3 << 8) | (g3 << 16) | (b3 << 24);
    } else {
      for (size_t i = q; i < HW && i < q + 4; ++i)
        for (int c = 0; c < 3; ++c) out[i * 3 + c] = (uint8_t)wb_to_u8(base[(size_t)c * HW + i], p.lo, p.hi);
    }
  }
}
