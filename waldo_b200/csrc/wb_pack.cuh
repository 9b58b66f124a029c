// Input packing: build the path's `input` tensor on the device from what the dataset really
// holds -- 8-bit RGB and an 8-bit label map -- instead of shipping 3 + Nl fp32 planes over PCIe.
// Reference: data/base_dataset.py:173-183 (label -> one-hot -> 5 * (2x - 1)), :355-372 (ToTensor + Normalize(0.5, 0.5)),
// models/synthesizer.py:444 (input = cat([vid, lyt], dim=2)).
//   input[f, 0:3]   = ((rgb / 255) - 0.5) / 0.5          (or a copy of an fp32 frame that is already normalised)
//   input[f, 3 + c] = label == c ? on : off               (on = 5, off = -5; labels >= Nl give all-off)
// The output is written as channels-last records (include/waldo_b200.h): `input` (n, HW, Cp).
#pragma once
#include "wb_common.cuh"
#include "../../include/waldo_b200.h"

WB_DEV float wb_norm_u8(unsigned v) { return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), 0.5f), 0.5f); }

// One thread per pixel builds the whole channels-last record (Cp floats, a multiple of 4: 128-bit stores, padding zeroed).
__global__ void __launch_bounds__(256) k_pack_input(waldo_pack_input_t p) {
  const int C = 3 + p.Nl, Cp = p.Cp;
  const size_t HW = (size_t)p.HW;
  const int f = blockIdx.y;
  for (size_t q = (size_t)blockIdx.x * wb_nthr() + wb_tid(); q < HW; q += (size_t)gridDim.x * wb_nthr()) {
    const unsigned lab = p.label[(size_t)f * HW + q];
    float rgb[3];
    WB_UNROLL for (int c = 0; c < 3; ++c)
      rgb[c] = p.rgb_u8 ? wb_norm_u8(p.rgb_u8[((size_t)f * 3 + c) * HW + q]) : p.rgb_f32[((size_t)f * 3 + c) * HW + q];
    float* out = p.input + ((size_t)f * HW + q) * Cp;
    for (int j = 0; j < Cp / 4; ++j) {
      float4 v;
      WB_UNROLL for (int e = 0; e < 4; ++e) {
        const int ch = 4 * j + e;
        const float x = ch < 3 ? (ch == 0 ? rgb[0] : (ch == 1 ? rgb[1] : rgb[2])) : (ch < C ? (lab == (unsigned)(ch - 3) ? p.on : p.off) : 0.f);
        wb_set(v, e, x);
      }
      wb_st4(out + 4 * j, v);
    }
  }
}

// Generic relayout: planar (n, C, HW) fp32 -> channels-last records (n, HW, Cp) with zeroed padding, and back.  Used where a
// caller hands over (or wants) an NCHW-contiguous tensor; the path itself only ever touches records.
__global__ void __launch_bounds__(256) k_to_records(int n, int C, int Cp, long long HW, const float* __restrict__ src, float* __restrict__ dst) {
  const int f = blockIdx.y;
  for (long long q = (long long)blockIdx.x * wb_nthr() + wb_tid(); q < HW; q += (long long)gridDim.x * wb_nthr()) {
    float* out = dst + ((size_t)f * HW + q) * Cp;
    const float* in = src + (size_t)f * C * HW + q;
    for (int j = 0; j < Cp / 4; ++j) {
      float4 v;
      WB_UNROLL for (int e = 0; e < 4; ++e) wb_set(v, e, 4 * j + e < C ? __ldg(in + (size_t)(4 * j + e) * HW) : 0.f);
      wb_st4(out + 4 * j, v);
    }
  }
}
__global__ void __launch_bounds__(256) k_from_records(int n, int C, int Cp, long long HW, const float* __restrict__ src, float* __restrict__ dst) {
  const int f = blockIdx.y;
  for (long long q = (long long)blockIdx.x * wb_nthr() + wb_tid(); q < HW; q += (long long)gridDim.x * wb_nthr()) {
    const float* in = src + ((size_t)f * HW + q) * Cp;
    float* out = dst + (size_t)f * C * HW + q;
    for (int c = 0; c < C; ++c) out[(size_t)c * HW] = __ldg(in + c);
  }
}
