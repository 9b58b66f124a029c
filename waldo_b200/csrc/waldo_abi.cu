// C-ABI entry points of libwaldo_b200.so (see include/waldo_b200.h).  Argument validation + kernel launches;
// no allocation, no synchronisation, launches on the caller's stream.
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include "wb_common.cuh"
#include "wb_geom.cuh"
#include "wb_prep.cuh"
#include "wb_composite.cuh"
#define WB_DET 0
#include "wb_composite_bwd.cuh"   // namespace wb_plain: float reductions (default)
#undef WB_DET
#define WB_DET 1
#include "wb_composite_bwd.cuh"   // namespace wb_fixed: 64-bit fixed-point accumulation (deterministic gradients)
#undef WB_DET
#include "wb_wif.cuh"
#include "wb_pack.cuh"
#include "wb_field.cuh"
#include "wb_loss.cuh"
#include "wb_conv.cuh"

static thread_local char g_err[512] = "";
std::atomic<long long> g_wb_launches{0};

static int wb_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#ifndef WB_HOST_EMU
static int wb_check_launch(const char* file, int line) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return wb_fail(WALDO_ECUDA, "%s:%d: CUDA launch failed: %s", file, line, cudaGetErrorString(e));
  return 0;
}
#endif

#define WB_REQUIRE(cond, ...) do { if (!(cond)) return wb_fail(WALDO_EINVAL, __VA_ARGS__); } while (0)
#define WB_LAUNCHED() do { int rc_ = WB_CHECK_LAUNCH(); if (rc_) return rc_; } while (0)

// WALDO_INV_UNFUSED=1: objects take the phase-per-kernel inverse warp (A/B runs, and the tests of that path)
static bool wb_inv_unfused() {
  const char* e = getenv("WALDO_INV_UNFUSED");
  return e && e[0] == '1';
}

static inline unsigned wb_blocks(long long total, int threads, long long cap = 1 << 20) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (unsigned)b;
}

// ST = storage type of input / alpha / raw_output / out_full (float, or wb_bf16 for the forward-only bf16-storage variant)
template <typename ST>
static int wb_decode_fwd_launch(const waldo_decode_fwd_t& A, waldo_stream_t st) {
  const waldo_decode_fwd_t* a = &A;
  const waldo_geom_t& g = a->g;
  const bool filt = (g.flags & WALDO_F_FILTER) != 0;
  const bool from_cls = (g.flags & WALDO_F_HAS_CLS) && !(g.flags & WALDO_F_WEIGHT_CLS);
  const int L = g.No + 1, HW = g.H * g.W;
  const long long HWd = (long long)g.Hd * g.Wd;
  const bool st_prep = a->stages == 0 || (a->stages & 1), st_layers = a->stages == 0 || (a->stages & 2), st_gather = a->stages == 0 || (a->stages & 4);
  const bool st_aprep = a->stages == 0 || (a->stages & 8);
  if (st_prep) {
  // B1
  WB_LAUNCH(k_project_alpha, dim3(wb_blocks((long long)g.B * g.Tw * HW, 128)), dim3(128), 0, st, *a);
  WB_LAUNCHED();
  // B2
  if (filt) {
    if (!from_cls) {
      if (g.Nl == 20) WB_LAUNCH((k_class_profile<20, ST>), dim3(a->prof_ctas, g.B), dim3(256), 0, st, *a);
      else if (g.Nl == 19) WB_LAUNCH((k_class_profile<19, ST>), dim3(a->prof_ctas, g.B), dim3(256), 0, st, *a);
      else WB_LAUNCH((k_class_profile<0, ST>), dim3(a->prof_ctas, g.B), dim3(256), 0, st, *a);
      WB_LAUNCHED();
    }
    WB_LAUNCH(k_profile_final, dim3(g.B), dim3(352), 0, st, *a);
    WB_LAUNCHED();
  }
  }
  // B2b-B4
  if (st_aprep) {
    const dim3 pgrid(wb_blocks(HWd, WB_TILE_PX, 1024), g.B * g.Tw);
    if (g.Nl == 20) WB_LAUNCH((k_alpha_prep<20, ST>), pgrid, dim3(WB_TILE_PX), 0, st, *a);        // Cityscapes
    else if (g.Nl == 19) WB_LAUNCH((k_alpha_prep<19, ST>), pgrid, dim3(WB_TILE_PX), 0, st, *a);   // KITTI
    else WB_LAUNCH((k_alpha_prep<0, ST>), pgrid, dim3(WB_TILE_PX), 0, st, *a);
    WB_LAUNCHED();
  }
  if (st_prep) {
  // B5
  WB_LAUNCH(k_layer_flow_lo, dim3(wb_blocks((long long)g.B * g.Tp * HW, 128)), dim3(128), 0, st, *a);
  WB_LAUNCHED();
  }
  // B5(up)-B9: the layer kernel
  const dim3 grid(wb_blocks(HWd, WB_TILE_PX, 1024), g.B * g.Tp);
  if (st_layers) {
    WB_LAUNCH(k_layers_fwd<ST>, grid, dim3(WB_TILE_PX), 0, st, *a);
    WB_LAUNCHED();
  }
  // stage C: the gather kernel
  if (st_gather) {
    const bool self = (g.flags & WALDO_F_INCLUDE_SELF) && g.Tp == g.T;
    if (g.Tc == 4 && !self) WB_LAUNCH((k_gather_fwd<4, true, ST>), grid, dim3(WB_TILE_PX), 0, st, *a);
    else if (g.Tc <= 4) WB_LAUNCH((k_gather_fwd<4, false, ST>), grid, dim3(WB_TILE_PX), 0, st, *a);
    else WB_LAUNCH((k_gather_fwd<8, false, ST>), grid, dim3(WB_TILE_PX), 0, st, *a);
    WB_LAUNCHED();
  }
  return 0;
}

extern "C" {

const char* waldo_last_error(void) { return g_err; }
int waldo_abi_version(void) { return WALDO_ABI_VERSION; }
long long waldo_launch_count(void) { return g_wb_launches.load(); }
int waldo_has_device_code(void) {
#ifdef WB_HOST_EMU
  return 0;
#else
  return 1;
#endif
}

// ------------------------------------------------------------------------------------------ TPS
int waldo_tps_fwd(const waldo_tps_fwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->N > 0 && a->P > 0, "tps_fwd: bad sizes");
  WB_REQUIRE(a->N + 3 <= WB_MAX_K, "tps_fwd: %d control points exceed the compiled maximum %d", a->N, WB_MAX_K - 3);
  WB_REQUIRE(a->inverse_kernel && a->tgt_grid_repr && a->pts && a->mapping && a->grid, "tps_fwd: null pointer");
  if (a->n == 0) return 0;
  WB_LAUNCH(k_tps_mapping, dim3(a->n), dim3(128), 0, st, a->n, a->N, a->inverse_kernel, a->pts, a->mapping);
  WB_LAUNCHED();
  dim3 grid(wb_blocks(a->P, 128), (a->n + WB_TPS_NI - 1) / WB_TPS_NI);
  WB_LAUNCH(k_tps_eval, grid, dim3(128), 0, st, a->n, a->N, a->P, a->tgt_grid_repr, a->mapping, a->grid);
  WB_LAUNCHED();
  return 0;
}

int waldo_tps_bwd(const waldo_tps_bwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->N > 0 && a->P > 0 && a->chunks > 0, "tps_bwd: bad sizes");
  WB_REQUIRE(a->N + 3 <= WB_MAX_K, "tps_bwd: too many control points");
  WB_REQUIRE(a->inverse_kernel && a->tgt_grid_repr && a->dgrid && a->partial && a->dpts, "tps_bwd: null pointer");
  if (a->n == 0) return 0;
  WB_LAUNCH(k_tps_bwd_partial, dim3(a->n, a->chunks), dim3(((a->N + 3 + 31) / 32) * 32), 0, st, a->n, a->N, a->P, a->chunks,
            a->tgt_grid_repr, a->dgrid, a->partial);
  WB_LAUNCHED();
  WB_LAUNCH(k_tps_bwd_final, dim3(a->n), dim3(128), 0, st, a->n, a->N, a->chunks, a->inverse_kernel, a->partial, a->dpts);
  WB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ inverse warp
int waldo_invwarp_fwd(const waldo_invwarp_fwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->Hs > 0 && a->Ws > 0 && a->Ht > 0 && a->Wt > 0, "invwarp_fwd: bad sizes");
  WB_REQUIRE(a->niter >= 0 && a->niter < 120, "invwarp_fwd: niter out of range");
  WB_REQUIRE(a->n <= 65535, "invwarp_fwd: more than 65535 items in one call");
  WB_REQUIRE(a->fwd_grid && a->id_src && a->id_tgt && a->gauss && a->out && a->field && a->winner && a->level && a->eroded && a->val,
             "invwarp_fwd: null pointer");
  if (a->n == 0) return 0;
  WB_REQUIRE(a->bbox, "invwarp_fwd: null bbox scratch");
  WbInvArgs k = {a->n, a->Hs, a->Ws, a->Ht, a->Wt, a->niter, a->erode, a->fwd_grid, a->id_src, a->id_tgt, a->gauss,
                 a->out, a->field, a->winner, a->level, a->eroded, a->val, a->bbox};
  const int m = a->niter + 1, PP = (a->Ht + 2 * m) * (a->Wt + 2 * m), P = a->Ht * a->Wt;
  const dim3 gpp(wb_blocks(PP, 256, 512), a->n), gp(wb_blocks(P, 256, 512), a->n);
  const dim3 gband((a->Ht + 2 * m + WB_INV_ROWS - 1) / WB_INV_ROWS, a->n);
  // a forward map defined on a lattice much smaller than the target (an object canvas) lands in a small box of the target
  const bool small_box = (long long)a->Hs * a->Ws * 4 <= (long long)a->Ht * a->Wt;
  if (small_box && a->Hs * a->Ws <= WB_INVF_MAX_SRC && !wb_inv_unfused()) {
    // ... and all its phases run in one CTA with the working set in shared memory (k_inv_fused)
    const int NS = a->Hs * a->Ws;
    const size_t budget = 224 * 1024;
    int cap = (int)((budget - (size_t)NS * 8) / 14) & ~3;
    // test knobs: a smaller cell budget (boxes that do not fit take the global-memory work area) and a margin adjustment
    // (a negative one makes samples land outside the predicted box, which must send the item down the fallback)
    if (const char* e = getenv("WALDO_INV_CAP")) cap = max(4, min(cap, atoi(e))) & ~3;
    const char* em = getenv("WALDO_INV_MARGIN");
    const int mg_adj = em ? atoi(em) : 0;
    const size_t smem = (size_t)NS * 8 + (size_t)cap * 14;
#ifndef WB_HOST_EMU
    cudaFuncSetAttribute(k_inv_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);   // > 48 KB: opt in
#endif
    WB_LAUNCH(k_inv_fused, dim3(a->n), dim3(WB_INVF_THREADS), smem, st, k, cap, mg_adj); WB_LAUNCHED();
    return 0;
  }
  WB_LAUNCH(k_inv_clear, gpp, dim3(256), 0, st, k); WB_LAUNCHED();
  WB_LAUNCH(k_inv_claim, gp, dim3(256), 0, st, k); WB_LAUNCHED();
  WB_LAUNCH(k_inv_deposit, gp, dim3(256), 0, st, k); WB_LAUNCHED();
  // one CTA per item runs all growth iterations of a small box; otherwise (background) one flat launch per iteration
  if (small_box) { WB_LAUNCH(k_inv_grow_fused, dim3(a->n), dim3(512), 0, st, k); WB_LAUNCHED(); }
  else {
    for (int it = 1; it <= a->niter; ++it) { WB_LAUNCH(k_inv_dilate, gband, dim3(256), 0, st, k, it); WB_LAUNCHED(); }
    if (a->erode)
      for (int it = 1; it <= a->niter; ++it) { WB_LAUNCH(k_inv_erode, gband, dim3(256), 0, st, k, it); WB_LAUNCHED(); }
  }
  WB_LAUNCH(k_inv_final, gp, dim3(256), 0, st, k); WB_LAUNCHED();
  return 0;
}

int waldo_invwarp_bwd(const waldo_invwarp_bwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->Hs > 0 && a->Ws > 0 && a->Ht > 0 && a->Wt > 0, "invwarp_bwd: bad sizes");
  WB_REQUIRE(a->gauss && a->dout && a->field && a->winner && a->level && a->eroded && a->gval && a->inv_sw && a->gdisp && a->dfwd_grid,
             "invwarp_bwd: null pointer");
  if (a->n == 0) return 0;
  WB_REQUIRE(a->bbox, "invwarp_bwd: null bbox");
  WB_REQUIRE(a->Wt <= 8 * a->Ws, "invwarp_bwd: target lattice more than 8x wider than the source lattice");
  WbInvBwdArgs k = {a->n, a->Hs, a->Ws, a->Ht, a->Wt, a->niter, a->gauss, a->dout, a->field, a->winner, a->level, a->eroded,
                    a->bbox, a->gval, a->inv_sw, a->gdisp, a->dfwd_grid};
  const int m = a->niter + 1, PP = (a->Ht + 2 * m) * (a->Wt + 2 * m), P = a->Ht * a->Wt;
  const dim3 gpp(wb_blocks(PP, 256, 512), a->n), gp(wb_blocks(P, 256, 512), a->n), gs(wb_blocks(a->Hs * a->Ws, 256, 512), a->n);
  const dim3 gband((a->Ht + 2 * m + WB_INV_ROWS - 1) / WB_INV_ROWS, a->n);
  WB_LAUNCH(k_invb_init, gband, dim3(256), 0, st, k); WB_LAUNCHED();
  if ((long long)a->Hs * a->Ws * 4 <= (long long)a->Ht * a->Wt) { WB_LAUNCH(k_invb_levels_fused, dim3(a->n), dim3(512), 0, st, k); WB_LAUNCHED(); }
  else
    for (int lv = a->niter - 1; lv >= 0; --lv) { WB_LAUNCH(k_invb_level, gband, dim3(256), 0, st, k, lv); WB_LAUNCHED(); }
  WB_LAUNCH(k_invb_handoff, gp, dim3(256), 0, st, k); WB_LAUNCHED();
  WB_LAUNCH(k_invb_resize_t, gs, dim3(256), 0, st, k); WB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ occlusion matrix
int waldo_occ_fwd(int BT, int No, const float* score, float* occ, waldo_stream_t st) {
  WB_REQUIRE(BT >= 0 && No > 0 && score && occ, "occ_fwd: bad arguments");
  if (BT == 0) return 0;
  long long total = (long long)BT * (No + 1) * (No + 1);
  WB_LAUNCH(k_occ_fwd, dim3(wb_blocks(total, 256)), dim3(256), 0, st, BT, No, score, occ);
  WB_LAUNCHED();
  return 0;
}
int waldo_occ_bwd(int BT, int No, const float* score, const float* docc, float* dscore, waldo_stream_t st) {
  WB_REQUIRE(BT >= 0 && No > 0 && score && docc && dscore, "occ_bwd: bad arguments");
  if (BT == 0) return 0;
  WB_LAUNCH(k_occ_bwd, dim3(wb_blocks((long long)BT * No, 128)), dim3(128), 0, st, BT, No, score, docc, dscore);
  WB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ decode_output
static int wb_check_geom(const waldo_geom_t& g, const char* who) {
  WB_REQUIRE(g.B > 0 && g.T > 0 && g.Tw > 0 && g.Tc > 0 && g.Tp > 0, "%s: bad batch/time sizes", who);
  WB_REQUIRE(g.Tw <= g.T, "%s: Tw > T", who);
  WB_REQUIRE(g.Tc <= 8, "%s: %d contexts exceed the compiled maximum 8", who, g.Tc);
  WB_REQUIRE(g.No >= 1 && g.No + 1 <= WB_MAX_L, "%s: num_obj=%d unsupported (compiled max %d)", who, g.No, WB_MAX_L - 1);
  WB_REQUIRE(g.Nl >= 1 && g.Nl <= WB_MAX_NL, "%s: num_lyt=%d unsupported (compiled max %d)", who, g.Nl, WB_MAX_NL);
  WB_REQUIRE(g.C == 3 + g.Nl && g.C <= WB_MAX_C, "%s: C=%d must equal 3+num_lyt and be <= %d", who, g.C, WB_MAX_C);
  WB_REQUIRE(g.H > 1 && g.W > 1 && g.Hd >= g.H && g.Wd >= g.W && g.Ho > 1 && g.Wo > 1, "%s: bad spatial sizes", who);
  WB_REQUIRE((long long)g.Hd * g.W == (long long)g.H * g.Wd, "%s: HD and low-res aspect differ", who);
  WB_REQUIRE((long long)g.Hd * g.Wd < (1ll << 30), "%s: frame too large for 32-bit pixel indices", who);
  if (g.flags & WALDO_F_RESTRICT_CTX) WB_REQUIRE(g.flags & WALDO_F_FILTER, "%s: restrict_to_ctx implies the filter", who);
  if (g.flags & WALDO_F_WEIGHT_CLS) WB_REQUIRE(g.flags & WALDO_F_HAS_CLS, "%s: weight_cls needs cls", who);
  return 0;
}

int waldo_decode_fwd(const waldo_decode_fwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a, "decode_fwd: null argument");
  const waldo_geom_t& g = a->g;
  int rc = wb_check_geom(g, "decode_fwd");
  if (rc) return rc;
  WB_REQUIRE(a->input && a->tgt_grid_obj && a->src_grid_obj && a->tgt_grid_bg && a->src_grid_bg && a->occ && a->obj_alpha &&
             a->bg_alpha && a->ctx_ts && a->pred_ts && a->xs_hd && a->ys_hd, "decode_fwd: null input pointer");
  WB_REQUIRE(a->a_lo && a->f_lo && a->alpha && a->flow && a->raw_output && a->out_full && a->live_ctx && a->live_pred && a->norm && a->score,
             "decode_fwd: null output pointer");
  const bool filt = (g.flags & WALDO_F_FILTER) != 0;
  const bool from_cls = (g.flags & WALDO_F_HAS_CLS) && !(g.flags & WALDO_F_WEIGHT_CLS);
  if (g.flags & WALDO_F_HAS_CLS) WB_REQUIRE(a->cls, "decode_fwd: cls flagged but null");
  if (g.flags & WALDO_F_IS_OBJ) WB_REQUIRE(a->s_lo, "decode_fwd: s_lo needed for is_obj");
  if (filt) {
    WB_REQUIRE(a->prof_p, "decode_fwd: prof_p needed for the filter");
    if (!from_cls) WB_REQUIRE(a->prof_part && a->prof_sum && a->prof_ctas > 0, "decode_fwd: profile scratch needed");
  }
  WB_REQUIRE(a->storage == WALDO_ST_F32 || a->storage == WALDO_ST_BF16, "decode_fwd: unknown storage type %d", a->storage);
  return a->storage == WALDO_ST_BF16 ? wb_decode_fwd_launch<wb_bf16>(*a, st) : wb_decode_fwd_launch<float>(*a, st);
}

int waldo_decode_bwd(const waldo_decode_bwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a, "decode_bwd: null argument");
  return a->det_shadow ? wb_fixed::wb_decode_bwd_launch(*a, st) : wb_plain::wb_decode_bwd_launch(*a, st);
}

// ------------------------------------------------------------------------------------------ WIF fuse tail
int waldo_wif_fuse_fwd(const waldo_wif_fuse_fwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->B > 0 && a->Tc > 0 && a->Tp > 0 && a->HW > 0, "wif_fuse_fwd: bad sizes");
  WB_REQUIRE(a->Cr >= 5 || !a->ab, "wif_fuse_fwd: raw_output needs >= 5 channels for the gate");
  WB_REQUIRE(a->Cr >= 3 && a->raw_output && a->unet_out && a->frame, "wif_fuse_fwd: null pointer");
  WB_LAUNCH(k_wif_fuse_fwd, dim3(wb_blocks(a->HW, 256), a->B * a->Tp), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}
int waldo_wif_fuse_bwd(const waldo_wif_fuse_bwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->f.B > 0 && a->f.Tc > 0 && a->f.Tp > 0 && a->f.HW > 0, "wif_fuse_bwd: bad sizes");
  WB_REQUIRE(a->f.Tc <= WB_WIF_MAX_TC, "wif_fuse_bwd: Tc=%d exceeds compiled maximum %d", a->f.Tc, WB_WIF_MAX_TC);
  WB_REQUIRE(a->f.raw_output && a->f.unet_out && a->d_frame, "wif_fuse_bwd: null pointer");
  WB_LAUNCH(k_wif_fuse_bwd, dim3(wb_blocks(a->f.HW, 256), a->f.B * a->f.Tp), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ field warp / scale
int waldo_warp_field_fwd(const waldo_warp_field_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->c > 0 && a->h > 1 && a->w > 1 && a->H > 0 && a->W > 0, "warp_field: bad sizes");
  WB_REQUIRE(a->field && a->grid && a->out, "warp_field: null pointer");
  if (a->n == 0) return 0;
  WB_LAUNCH(k_warp_field, dim3(wb_blocks((long long)a->n * a->H * a->W, 256, 8192)), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}
int waldo_resize_bilinear_fwd(const waldo_resize_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->h > 0 && a->w > 0 && a->H >= a->h && a->W >= a->w, "resize_bilinear: bad sizes (up-sampling only)");
  WB_REQUIRE(a->in && a->out, "resize_bilinear: null pointer");
  if (a->n == 0) return 0;
  WB_LAUNCH(k_resize_bilinear, dim3(wb_blocks((long long)a->n * a->H * a->W, 256, 8192)), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ input packing
int waldo_pack_input(const waldo_pack_input_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->Nl >= 1 && a->HW > 0, "pack_input: bad sizes");
  WB_REQUIRE((a->rgb_u8 || a->rgb_f32) && a->label && a->input, "pack_input: null pointer");
  if (a->n == 0) return 0;
  WB_REQUIRE(a->storage == WALDO_ST_F32 || a->storage == WALDO_ST_BF16, "pack_input: unknown storage type %d", a->storage);
  const dim3 grid(wb_blocks(((long long)a->HW + 3) / 4, 256, 1024), a->n);
  if (a->storage == WALDO_ST_BF16) WB_LAUNCH(k_pack_input<wb_bf16>, grid, dim3(256), 0, st, *a);
  else WB_LAUNCH(k_pack_input<float>, grid, dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ output side
int waldo_frames_to_u8(const waldo_frames_u8_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->HW > 0 && a->hi > a->lo, "frames_to_u8: bad sizes / span");
  WB_REQUIRE(a->frames && a->out, "frames_to_u8: null pointer");
  WB_REQUIRE(((uintptr_t)a->frames & 15) == 0 && ((uintptr_t)a->out & 3) == 0, "frames_to_u8: frames must be 16-byte, out 4-byte aligned");
  if (a->n == 0) return 0;
  WB_LAUNCH(k_frames_to_u8, dim3(wb_blocks(((long long)a->HW + 3) / 4, 256, 1024), a->n), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ loss epilogues
static int wb_blur_check(const waldo_blur_t* a, const char* who) {
  WB_REQUIRE(a && a->n >= 0 && a->H > 0 && a->W > 0, "%s: bad sizes", who);
  WB_REQUIRE(a->ksize >= 1 && (a->ksize & 1) && a->ksize <= WB_BLUR_MAX_K, "%s: kernel_size must be odd and <= %d", who, WB_BLUR_MAX_K);
  WB_REQUIRE(a->H > a->ksize / 2 && a->W > a->ksize / 2, "%s: reflect padding needs H, W > kernel_size / 2", who);
  WB_REQUIRE(a->sigma > 0.f && a->in && a->out && a->in != a->out, "%s: bad sigma / pointers", who);
  WB_REQUIRE(a->n <= 65535, "%s: more than 65535 planes in one call", who);
  return 0;
}
int waldo_blur_fwd(const waldo_blur_t* a, waldo_stream_t st) {
  int rc = wb_blur_check(a, "blur_fwd");
  if (rc) return rc;
  if (a->n == 0) return 0;
  const int tiles = ((a->W + WB_BLUR_T - 1) / WB_BLUR_T) * ((a->H + WB_BLUR_T - 1) / WB_BLUR_T);
  WB_LAUNCH(k_blur<false>, dim3(tiles, a->n), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}
int waldo_blur_bwd(const waldo_blur_t* a, waldo_stream_t st) {
  int rc = wb_blur_check(a, "blur_bwd");
  if (rc) return rc;
  if (a->n == 0) return 0;
  const int tiles = ((a->W + WB_BLUR_T - 1) / WB_BLUR_T) * ((a->H + WB_BLUR_T - 1) / WB_BLUR_T);
  WB_LAUNCH(k_blur<true>, dim3(tiles, a->n), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}
int waldo_layer_entropy_fwd(const waldo_layer_entropy_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->L >= 1 && a->HW > 0 && a->n <= 65535, "layer_entropy_fwd: bad sizes");
  WB_REQUIRE(a->alpha && (a->entropy || a->fg), "layer_entropy_fwd: null pointer");
  if (a->n == 0) return 0;
  WB_LAUNCH(k_layer_entropy_fwd, dim3(wb_blocks(a->HW, 256, 1024), a->n), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}
int waldo_layer_entropy_bwd(const waldo_layer_entropy_bwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->f.n >= 0 && a->f.L >= 1 && a->f.HW > 0 && a->f.n <= 65535, "layer_entropy_bwd: bad sizes");
  WB_REQUIRE(a->f.alpha && a->d_alpha, "layer_entropy_bwd: null pointer");
  if (a->f.n == 0) return 0;
  WB_LAUNCH(k_layer_entropy_bwd, dim3(wb_blocks(a->f.HW, 256, 1024), a->f.n), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}

static int wb_pose_dis_check(const waldo_pose_dis_t* a, const char* who) {
  WB_REQUIRE(a && a->n >= 0 && a->n <= 65535 && a->HW > 0, "pose_dis: bad sizes");
  WB_REQUIRE(a->No >= 1 && a->No <= WB_PD_MAX_NO && a->ho >= 2 && a->wo >= 2 && (long long)a->No * (a->ho - 1) * (a->wo - 1) <= WB_PD_MAX_CELLS,
             "pose_dis: No <= 32, obj_shape >= (2, 2), No (ho-1)(wo-1) <= 1024");
  WB_REQUIRE(a->mov && a->fg && a->pose && a->grid && a->cell_arg && a->center_arg, "pose_dis: null pointer");
  (void)who;
  return 0;
}
int waldo_pose_dis_fwd(const waldo_pose_dis_t* a, waldo_stream_t st) {
  int rc = wb_pose_dis_check(a, "pose_dis_fwd");
  if (rc) return rc;
  WB_REQUIRE(a->cell_min && a->center_min, "pose_dis_fwd: null output");
  if (a->n == 0) return 0;
  WB_LAUNCH(k_pose_dis_fwd, dim3(wb_blocks(a->HW, WB_PD_THREADS, 1024), a->n), dim3(WB_PD_THREADS), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}
int waldo_pose_dis_bwd(const waldo_pose_dis_bwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a, "pose_dis_bwd: null");
  int rc = wb_pose_dis_check(&a->f, "pose_dis_bwd");
  if (rc) return rc;
  WB_REQUIRE(a->part && a->d_pose && a->ctas >= 1 && a->ctas <= 1024, "pose_dis_bwd: part / d_pose / ctas");
  if (a->f.n == 0) return 0;
  WB_LAUNCH(k_pose_dis_bwd, dim3(a->ctas, a->f.n), dim3(WB_PD_THREADS), 0, st, *a);
  WB_LAUNCHED();
  WB_LAUNCH(k_pose_dis_bwd_final, dim3(wb_blocks((long long)a->f.n * a->f.No, 128, 1024)), dim3(128), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}

static int wb_obj_flow_check(const waldo_obj_flow_t* a) {
  WB_REQUIRE(a && a->n >= 0 && a->n <= 65535 && a->HW > 0 && a->L >= 2 && a->L <= WB_OF_MAX_L, "obj_flow: bad sizes (2 <= L <= 33)");
  WB_REQUIRE(a->alpha && a->flow && a->part && a->mom && a->ctas >= 1 && a->ctas <= 1024, "obj_flow: null pointer / ctas");
  return 0;
}
int waldo_obj_flow_fwd(const waldo_obj_flow_t* a, waldo_stream_t st) {
  int rc = wb_obj_flow_check(a);
  if (rc) return rc;
  WB_REQUIRE(a->dev_map, "obj_flow_fwd: null output");
  if (a->n == 0) return 0;
  const int No = a->L - 1;
  WB_LAUNCH(k_of_reduce<0>, dim3(a->ctas, a->n), dim3(WB_OF_THREADS), 0, st, *a, (const float*)nullptr);
  WB_LAUNCHED();
  WB_LAUNCH(k_of_reduce_final, dim3(wb_blocks((long long)a->n * No * 3, 128, 1024)), dim3(128), 0, st, (const float*)a->part, a->mom, a->n, a->ctas, No * 3);
  WB_LAUNCHED();
  WB_LAUNCH(k_of_map, dim3(wb_blocks(a->HW, WB_OF_THREADS, 1024), a->n), dim3(WB_OF_THREADS), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}
int waldo_obj_flow_bwd(const waldo_obj_flow_bwd_t* a, waldo_stream_t st) {
  WB_REQUIRE(a, "obj_flow_bwd: null");
  int rc = wb_obj_flow_check(&a->f);
  if (rc) return rc;
  WB_REQUIRE(a->d_map && a->tsum && a->d_alpha, "obj_flow_bwd: null pointer");
  if (a->f.n == 0) return 0;
  const int No = a->f.L - 1;
  WB_LAUNCH(k_of_reduce<1>, dim3(a->f.ctas, a->f.n), dim3(WB_OF_THREADS), 0, st, a->f, a->d_map);
  WB_LAUNCHED();
  WB_LAUNCH(k_of_reduce_final, dim3(wb_blocks((long long)a->f.n * No * 2, 128, 1024)), dim3(128), 0, st, (const float*)a->f.part, a->tsum, a->f.n, a->f.ctas, No * 2);
  WB_LAUNCHED();
  WB_LAUNCH(k_of_dalpha, dim3(wb_blocks(a->f.HW, WB_OF_THREADS, 1024), a->f.n), dim3(WB_OF_THREADS), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ first UNet layer
int waldo_conv3x3_fwd(const waldo_conv3x3_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->n >= 0 && a->H > 0 && a->W > 0, "conv3x3_fwd: bad sizes");
  WB_REQUIRE(a->Cin >= 1 && a->Cin <= WB_CV_MAX_CIN, "conv3x3_fwd: Cin=%d unsupported (1..%d)", a->Cin, WB_CV_MAX_CIN);
  WB_REQUIRE(a->Cout >= 1 && a->Cout <= WB_CV_MAX_COUT, "conv3x3_fwd: Cout=%d unsupported (1..%d)", a->Cout, WB_CV_MAX_COUT);
  WB_REQUIRE((a->Tc > 0) == (a->Tp > 0), "conv3x3_fwd: Tc and Tp must both be set or both be 0");
  if (a->Tc > 0) WB_REQUIRE(a->n % (a->Tc * a->Tp) == 0, "conv3x3_fwd: n must be a multiple of Tc*Tp");
  WB_REQUIRE(a->in && a->weight && a->out && a->in != a->out, "conv3x3_fwd: null pointer");
  if (a->n == 0) return 0;
  const int Cp = (a->Cin + 7) & ~7, WP = a->Cout <= 24 ? 24 : (a->Cout <= 40 ? 40 : 56);
  // 16-byte staging needs whole 4-pixel chunks inside / outside the image and 16-byte aligned rows
  const bool vec = a->W % 4 == 0 && ((uintptr_t)a->in & 15) == 0;
  const size_t smem = ((size_t)Cp * WB_CV_ROWS * (vec ? WB_CV_PITCH_V : WB_CV_PITCH_S) + (size_t)9 * Cp * WP) * sizeof(float);
  const long long tiles = (long long)a->n * ((a->W + WB_CV_TW - 1) / WB_CV_TW) * ((a->H + WB_CV_TH - 1) / WB_CV_TH);
  const dim3 grid(wb_blocks(tiles, 1, 2 * 148));   // persistent: 2 CTAs per SM
#ifdef WB_HOST_EMU
#define WB_CV_GO(K) WB_LAUNCH(K, grid, dim3(256), smem, st, *a)
#else
#define WB_CV_GO(K) do { cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); WB_LAUNCH(K, grid, dim3(256), smem, st, *a); } while (0)
#endif
  if (vec && Cp == 40 && a->Cout == 16) WB_CV_GO((k_conv3x3_fwd<40, 2, true>));        // WIF's to_emb: 3 + 20 + 17 channels -> 16
  else if (vec && Cp == 48 && a->Cout == 16) WB_CV_GO((k_conv3x3_fwd<48, 2, true>));   // ... with the disocc channel (41 -> 48)
  else if (vec && Cp == 16 && a->Cout == 40) WB_CV_GO((k_conv3x3_fwd<16, 5, true>));   // its backward-data: 16 -> 40 (flipped, transposed weights)
  else if (vec && Cp == 32 && a->Cout <= 8) WB_CV_GO((k_conv3x3_fwd<32, 1, true>));    // UNet.from_emb: 2 x 16 -> 5 (4) channels
  else if (vec && Cp == 8 && a->Cout == 32) WB_CV_GO((k_conv3x3_fwd<8, 4, true>));     // ... and its backward-data
  else if (vec) WB_CV_GO((k_conv3x3_fwd<0, 0, true>));
  else WB_CV_GO((k_conv3x3_fwd<0, 0, false>));
#undef WB_CV_GO
  WB_LAUNCHED();
  return 0;
}

int waldo_conv3x3_wgrad(const waldo_conv3x3_wgrad_t* a, waldo_stream_t st) {
  WB_REQUIRE(a && a->c.n >= 0 && a->c.H > 0 && a->c.W > 0, "conv3x3_wgrad: bad sizes");
  WB_REQUIRE(a->c.Cin >= 1 && a->c.Cin <= 40, "conv3x3_wgrad: Cin=%d unsupported (1..40)", a->c.Cin);
  WB_REQUIRE(a->c.Cout >= 1 && a->c.Cout <= 16, "conv3x3_wgrad: Cout=%d unsupported (1..16)", a->c.Cout);
  WB_REQUIRE((a->c.Tc > 0) == (a->c.Tp > 0), "conv3x3_wgrad: Tc and Tp must both be set or both be 0");
  if (a->c.Tc > 0) WB_REQUIRE(a->c.n % (a->c.Tc * a->c.Tp) == 0, "conv3x3_wgrad: n must be a multiple of Tc*Tp");
  WB_REQUIRE(a->c.in && a->dout && a->part && a->dweight && a->ctas >= 1 && a->ctas <= 2 * 148, "conv3x3_wgrad: bad pointers / ctas");
  WB_REQUIRE(a->c.W % 4 == 0 && ((uintptr_t)a->c.in & 15) == 0 && ((uintptr_t)a->dout & 15) == 0,
             "conv3x3_wgrad: needs W %% 4 == 0 and 16-byte aligned in / dout");
  const int Cp = (a->c.Cin + 7) & ~7;
  const size_t smem = ((size_t)Cp * WB_CW_XS + (size_t)16 * WB_CW_YS) * sizeof(float);
#ifdef WB_HOST_EMU
#define WB_CW_GO(K) WB_LAUNCH(K, dim3(a->ctas), dim3(256), smem, st, *a)
#else
#define WB_CW_GO(K) do { cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); WB_LAUNCH(K, dim3(a->ctas), dim3(256), smem, st, *a); } while (0)
#endif
  if (Cp == 40) WB_CW_GO(k_conv3x3_wgrad<40>);
  else WB_CW_GO(k_conv3x3_wgrad<0>);
#undef WB_CW_GO
  WB_LAUNCHED();
  WB_LAUNCH(k_conv3x3_wgrad_final, dim3(wb_blocks((long long)a->c.Cout * a->c.Cin * 9, 256)), dim3(256), 0, st, *a);
  WB_LAUNCHED();
  return 0;
}

}  // extern "C"
