// WIF fuse tail (models/nets/wif.py:50-54), forward and backward, one thread per (b,tp,pixel).
//   w_tc   = softmax_tc(u[b,tp,tc,3]);   a_tc = sigmoid(raw[b,tc,tp,4] + 5)   (INPUT channel 4, as the reference)
//   frame  = sum_tc w_tc * (a_tc * raw[b,tc,tp,0:3] + u[b,tp,tc,0:3])
#pragma once
#include "wb_common.cuh"
#include "../../include/waldo_b200.h"

#define WB_WIF_MAX_TC 8

__global__ void __launch_bounds__(256) k_wif_fuse_fwd(waldo_wif_fuse_fwd_t p) {
  const int Cu = 4 + (p.ab ? 1 : 0);
  const size_t HW = (size_t)p.HW;
  const int btp = blockIdx.y, b = btp / p.Tp, tp = btp - b * p.Tp;
  for (size_t q = (size_t)blockIdx.x * wb_nthr() + wb_tid(); q < HW; q += (size_t)gridDim.x * wb_nthr()) {
    const float* u = p.unet_out + (((size_t)b * p.Tp + tp) * p.Tc) * Cu * HW + q;
    float mx = -INFINITY;
    for (int tc = 0; tc < p.Tc; ++tc) mx = fmaxf(mx, __ldg(u + ((size_t)tc * Cu + 3) * HW));
    float den = 0.f, acc[3] = {0.f, 0.f, 0.f};
    for (int tc = 0; tc < p.Tc; ++tc) {
      const float* r = p.raw_output + (((size_t)b * p.Tc + tc) * p.Tp + tp) * p.Cr * HW + q;
      const float* ut = u + (size_t)tc * Cu * HW;
      float e = expf(__ldg(ut + 3 * HW) - mx);
      float a = p.ab ? 1.f / (1.f + expf(-(__ldg(r + 4 * HW) + 5.f))) : 0.f;
      den += e;
      WB_UNROLL for (int c = 0; c < 3; ++c) acc[c] += e * (a * __ldg(r + c * HW) + __ldg(ut + c * HW));
    }
    float inv = 1.f / den;
    float* o = p.frame + ((size_t)b * p.Tp + tp) * 3 * HW + q;
    WB_UNROLL for (int c = 0; c < 3; ++c) o[c * HW] = acc[c] * inv;
  }
}

__global__ void __launch_bounds__(256) k_wif_fuse_bwd(waldo_wif_fuse_bwd_t pb) {
  const waldo_wif_fuse_fwd_t& p = pb.f;
  const int Cu = 4 + (p.ab ? 1 : 0);
  const size_t HW = (size_t)p.HW;
  const int btp = blockIdx.y, b = btp / p.Tp, tp = btp - b * p.Tp;
  for (size_t q = (size_t)blockIdx.x * wb_nthr() + wb_tid(); q < HW; q += (size_t)gridDim.x * wb_nthr()) {
    const float* u = p.unet_out + (((size_t)b * p.Tp + tp) * p.Tc) * Cu * HW + q;
    const float* df = pb.d_frame + ((size_t)b * p.Tp + tp) * 3 * HW + q;
    float g[3] = {__ldg(df), __ldg(df + HW), __ldg(df + 2 * HW)};
    float mx = -INFINITY;
    for (int tc = 0; tc < p.Tc; ++tc) mx = fmaxf(mx, __ldg(u + ((size_t)tc * Cu + 3) * HW));
    float den = 0.f, dotv[WB_WIF_MAX_TC], e[WB_WIF_MAX_TC], mean = 0.f;
    for (int tc = 0; tc < p.Tc; ++tc) {
      const float* r = p.raw_output + (((size_t)b * p.Tc + tc) * p.Tp + tp) * p.Cr * HW + q;
      const float* ut = u + (size_t)tc * Cu * HW;
      e[tc] = expf(__ldg(ut + 3 * HW) - mx);
      float a = p.ab ? 1.f / (1.f + expf(-(__ldg(r + 4 * HW) + 5.f))) : 0.f;
      den += e[tc];
      float dv = 0.f;
      WB_UNROLL for (int c = 0; c < 3; ++c) dv += g[c] * (a * __ldg(r + c * HW) + __ldg(ut + c * HW));
      dotv[tc] = dv;   // d frame / d w_tc
    }
    float inv = 1.f / den;
    for (int tc = 0; tc < p.Tc; ++tc) mean += dotv[tc] * e[tc] * inv;
    for (int tc = 0; tc < p.Tc; ++tc) {
      const float w = e[tc] * inv;
      const float* r = p.raw_output + (((size_t)b * p.Tc + tc) * p.Tp + tp) * p.Cr * HW + q;
      if (pb.d_unet_out) {
        float* du = pb.d_unet_out + ((((size_t)b * p.Tp + tp) * p.Tc) + tc) * Cu * HW + q;
        WB_UNROLL for (int c = 0; c < 3; ++c) du[c * HW] = w * g[c];
        du[3 * HW] = w * (dotv[tc] - mean);
        if (p.ab) du[4 * HW] = 0.f;          // the reference never reads UNet channel 4
      }
      if (pb.d_raw_output) {
        float* dr = pb.d_raw_output + (((size_t)b * p.Tc + tc) * p.Tp + tp) * p.Cr * HW + q;
        float a = p.ab ? 1.f / (1.f + expf(-(__ldg(r + 4 * HW) + 5.f))) : 0.f;
        float da = 0.f;
        WB_UNROLL for (int c = 0; c < 3; ++c) { dr[c * HW] = w * a * g[c]; da += w * g[c] * __ldg(r + c * HW); }
        dr[3 * HW] = 0.f;
        dr[4 * HW] = da * a * (1.f - a);
      }
    }
  }
}
