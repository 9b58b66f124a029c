/*
 * waldo_b200 -- C ABI of the B200-native (sm_100a) warp+composite hot path of WALDO.
 *
 * This is the drop-in boundary: plain pointers, sizes and a CUDA stream.  No torch types.
 * All pointers are DEVICE pointers to dense row-major fp32 tensors unless stated; the caller
 * owns every buffer (outputs, saved state, scratch); kernels never allocate, never keep a pointer
 * after return and never synchronise.  Every entry point returns 0 on success or a negative
 * WALDO_E* code; `waldo_last_error()` gives the message for the calling thread.
 *
 * The reference (16lemoing/waldo) has no native interface for this path: it is ~170 ATen calls per
 * decode inside three Python classes.  Each entry point therefore cites the reference Python
 * function it replaces (paths relative to the reference root).  The reference-side binding is the
 * ctypes shim shown in INTEGRATION.md (shipped as waldo_b200/_lib.py).
 *
 * Symbols (SURVEY.md symbol table): B batch, T frames, Tw frames that get a context alpha
 * (Tc with restrict_to_ctx, else T), Tc contexts, Tp predicted frames, No objects, L = No+1 layers
 * (layer 0 = background), C = 3+Nl input channels, H x W low-res, Hd x Wd full-res, Ho x Wo object canvas.
 */
#ifndef WALDO_B200_H
#define WALDO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WALDO_ABI_VERSION 2

/* compile-time capacity of the kernels (per-thread register arrays) */
#define WALDO_MAX_LAYERS 17   /* L  = num_obj + 1 */
#define WALDO_MAX_CH     24   /* C  = 3 + num_lyt */
#define WALDO_MAX_LYT    21   /* Nl */
#define WALDO_MAX_TPS_K  256  /* control points + 3 */

enum {
  WALDO_OK = 0,
  WALDO_EINVAL = -1,    /* bad argument / unsupported size */
  WALDO_ECUDA = -2,     /* CUDA launch error */
  WALDO_ENOGPU = -3     /* built without device code (never the case for the shipped .so) */
};

/* flags of waldo_geom_t.flags */
enum {
  WALDO_F_RESTRICT_CTX = 1 << 0, /* lvd.py:707 grid_to_flow_ctx (else :602 grid_to_flow)            */
  WALDO_F_FILTER       = 1 << 1, /* semantic filter on (always with RESTRICT_CTX; else !no_filter)   */
  WALDO_F_WEIGHT_CLS   = 1 << 2, /* lvd.py:736-739 weight the class histogram by cls                 */
  WALDO_F_HAS_CLS      = 1 << 3, /* cls != None                                                      */
  WALDO_F_IS_OBJ       = 1 << 4, /* lvd.py:788-791 (RESTRICT_CTX and not allow_ghost)                */
  WALDO_F_INCLUDE_SELF = 1 << 5, /* lvd.py:842-845 extra context = the target frame itself           */
  WALDO_F_USE_DISOCC   = 1 << 6, /* lvd.py:148-151 disocc appended to raw_output                     */
  WALDO_F_OCC_PAIRS    = 1 << 7  /* backward only: d_occ feeds waldo_occ_bwd and nothing else, so its row 0, column 0
                                    and diagonal (constants of lvd.py:63-66, never read there) are left at zero      */
};

/* storage type of the big HD streams of waldo_decode_fwd_t: input, raw_output, out_full (alpha, flow, norm, score and every low-res
 * tensor stay fp32: the sampling positions do not depend on the storage type) */
enum {
  WALDO_ST_F32  = 0,  /* the reference's precision: parity rules 1-3 apply                                                */
  WALDO_ST_BF16 = 1   /* bf16 storage, fp32 arithmetic, forward / inference only (half the HBM bytes of those streams);
                         the three pointers then address bf16 elements; tolerance stated in tests/parity.py (rule 4)     */
};

typedef void* waldo_stream_t; /* cudaStream_t */

const char* waldo_last_error(void);
int waldo_abi_version(void);
/* 1 when the library holds sm_100a device code (the shipped build), 0 for the CPU logic-emulation
 * build that only the unit tests compile (tests/emu). */
int waldo_has_device_code(void);
/* number of kernels this library has launched in this process (bench.py's `gpu_launches`) */
long long waldo_launch_count(void);

/* ------------------------------------------------------------------ a-1  TPSWarp.forward
 * models/modules/warp.py:49-55.  grid[n,p,:] = tgt_grid_repr[p,:] @ (inverse_kernel @ [pts[n];0_3x2]).
 * Accumulates in fp64 (the reference's two fp32 matmuls are ill-conditioned for the background). */
typedef struct {
  int n;                      /* items (B*T*No objects or B*T backgrounds) */
  int N;                      /* control points per item */
  int P;                      /* lattice points h*w */
  const float* inverse_kernel;/* (N+3, N+3)  buffer TPSWarp.inverse_kernel  warp.py:45 */
  const float* tgt_grid_repr; /* (P, N+3)    buffer TPSWarp.tgt_grid_repr   warp.py:47 */
  const float* pts;           /* (n, N, 2) */
  double* mapping;            /* scratch (n, N+3, 2) */
  float* grid;                /* out (n, P, 2) */
} waldo_tps_fwd_t;
int waldo_tps_fwd(const waldo_tps_fwd_t*, waldo_stream_t);

typedef struct {
  int n, N, P;
  const float* inverse_kernel;
  const float* tgt_grid_repr;
  const float* dgrid;         /* (n, P, 2) */
  int chunks;                 /* P is reduced in `chunks` ordered slices */
  double* partial;            /* scratch (n, chunks, N+3, 2) */
  float* dpts;                /* out (n, N, 2) */
} waldo_tps_bwd_t;
int waldo_tps_bwd(const waldo_tps_bwd_t*, waldo_stream_t);

/* ------------------------------------------------------------------ a-2  InverseWarp.forward
 * models/modules/warp.py:71-174 (num_perm == 1, pad=True).  Tie rule: lowest sample index wins. */
typedef struct {
  int n;                      /* items */
  int Hs, Ws;                 /* lattice of the forward map */
  int Ht, Wt;                 /* target (image) lattice */
  int niter;                  /* dilation (and erosion) iterations, reference default 5 */
  int erode;                  /* 1 for objects (lvd.py:861), 0 for background (:867) */
  const float* fwd_grid;      /* (n, Hs, Ws, 2) */
  const float* id_src;        /* (Hs, Ws, 2) buffer InverseWarp.src_grid  warp.py:65 */
  const float* id_tgt;        /* (Ht, Wt, 2) buffer InverseWarp.tgt_grid  warp.py:66 */
  const float* gauss;         /* (9) buffer InverseWarp.kernel            warp.py:64 */
  float* out;                 /* out (n, Ht, Wt, 2) */
  /* saved for backward / exposed for the index-map parity checks (int32, uint8) */
  int32_t* field;             /* (n, Ht*Wt) landing cell of every sample, -1 = outside   warp.py:84-88 */
  int32_t* winner;            /* (n, Ht*Wt) surviving sample per cell, INT32_MAX = none  warp.py:113-123 */
  uint8_t* level;             /* (n, Hp*Wp) 0 = hit, k = filled at dilation k, 255 = unknown; Hp = Ht+2(niter+1) */
  uint8_t* eroded;            /* (n, Hp*Wp) 0 = kept, k = removed at erosion k */
  float* val;                 /* scratch+saved (n, 2, Hp*Wp) inverse displacement in pixels */
  int32_t* bbox;              /* saved (n, 4) x0,y0,x1,y1 of the hit cells in padded coordinates (INT32_MAX,.,-1,. if none) */
} waldo_invwarp_fwd_t;
int waldo_invwarp_fwd(const waldo_invwarp_fwd_t*, waldo_stream_t);

typedef struct {
  int n, Hs, Ws, Ht, Wt, niter;
  const float* gauss;
  const float* dout;          /* (n, Ht, Wt, 2) */
  const int32_t* field;
  const int32_t* winner;
  const uint8_t* level;
  const uint8_t* eroded;
  const int32_t* bbox;
  float* gval;                /* scratch (n, 2, Hp*Wp) */
  float* inv_sw;              /* scratch (n, Hp*Wp) */
  float* gdisp;               /* scratch (n, Ht*Wt, 2) gradient of the resampled displacement */
  float* dfwd_grid;           /* out (n, Hs, Ws, 2) */
} waldo_invwarp_bwd_t;
int waldo_invwarp_bwd(const waldo_invwarp_bwd_t*, waldo_stream_t);

/* ------------------------------------------------------------------ a-4  LVD.compute_occ
 * models/nets/lvd.py:59-68. */
int waldo_occ_fwd(int BT, int No, const float* occ_score /* (BT,No) */, float* occ /* (BT,L,L) */, waldo_stream_t);
int waldo_occ_bwd(int BT, int No, const float* occ_score, const float* docc, float* dscore, waldo_stream_t);

/* ------------------------------------------------------------------ a-5..a-8  decode_output
 * LVD.forward(mode="decode_output") models/nets/lvd.py:141-153 =
 *   Warper.grid_to_flow_ctx :707-828 | Warper.grid_to_flow :602-705, then Warper.input_to_output :830-853. */
typedef struct {
  int B, T, Tw, Tc, Tp;
  int No, Nl, C;
  int H, W, Hd, Wd, Ho, Wo;
  int flags;
  float min_cls;              /* lvd.py:492 */
} waldo_geom_t;

typedef struct {
  waldo_geom_t g;
  /* inputs */
  const float* input;         /* (B, T, C, Hd, Wd) */
  const float* tgt_grid_obj;  /* (B, T, No, Ho, Wo, 2) */
  const float* src_grid_obj;  /* (B, T, No, H, W, 2) */
  const float* tgt_grid_bg;   /* (B, T, H, W, 2) */
  const float* src_grid_bg;   /* (B, T, H, W, 2) */
  const float* occ;           /* (B, T, L, L) */
  const float* obj_alpha;     /* (B, No, Ho, Wo) in [-1,1] */
  const float* bg_alpha;      /* (B, H, W) in [-1,1] */
  const float* cls;           /* (B, No, Nl) or NULL */
  const int64_t* ctx_ts;      /* (B, Tc, Tp) */
  const int64_t* pred_ts;     /* (Tp) */
  const float* xs_hd;         /* (Wd) x of buffer Warper.src_grid_hd  lvd.py:484 */
  const float* ys_hd;         /* (Hd) y of the same buffer */
  /* low-res intermediates (saved for backward) */
  float* a_lo;                /* (B, Tw, L, H, W) projected opacities, lvd.py:727 */
  float* prof_part;           /* scratch (B, prof_ctas, No*Nl + No) partial sums of the class profile */
  int prof_ctas;
  float* prof_sum;            /* (B, No*Nl + No) reduced sums (num | den), lvd.py:740-742 */
  float* prof_p;              /* (B, No, Nl) class probabilities used by the filter */
  float* lyt_lo;              /* (B, Tw, Nl, H, W) down-sampled layout logits (lvd.py:716), saved for backward; NULL = recompute */
  float* f_lo;                /* (B, Tc, Tp, L, H, W, 2) per-layer flow on the low-res lattice, lvd.py:792 */
  float* s_lo;                /* (B, Tp, No, H, W) object support, lvd.py:788 */
  uint32_t* live_ctx;         /* (B, Tw, H, W) bit k: layer k has an in-range tap at this low-res cell of context frame t */
  uint32_t* live_pred;        /* (B, Tp, H, W) same for the target frames (bit 0 always set) */
  /* outputs */
  float* alpha;               /* (B, Tw, L, Hd, Wd) in [-1,1]                         lvd.py:822 */
  float* flow;                /* (B, Tc, Tp, 2, Hd, Wd)                               lvd.py:818 */
  float* raw_output;          /* (B, Tc+self, Tp, C+L+disocc, Hd, Wd)                 lvd.py:846,151; alpha_ctx = channels C..C+L-1 */
  float* out_full;            /* (B, Tp, C+1, Hd, Wd): output | raw_alpha             lvd.py:851,147,152 */
  float* norm;                /* (B, Tp, Hd, Wd) sum_tc(score+eps), saved for backward */
  float* score;               /* (B, Tc, Tp, Hd, Wd) sum_k Actx_k per pair (lvd.py:841), glue between the two HD kernels */
  int stages;                 /* 0 = everything; else bit 0 = low-res kernels (B1, B2, B5), bit 3 = HD context-alpha kernel (B2b-B4),
                                 bit 1 = HD layer kernel (B5up-B9), bit 2 = HD gather kernel (stage C) */
  int storage;                /* WALDO_ST_F32 | WALDO_ST_BF16: element type of input, raw_output, out_full (everything else, incl.
                                 alpha / flow / norm / score and all low-res tensors, stays fp32).  waldo_decode_bwd needs WALDO_ST_F32 */
} waldo_decode_fwd_t;
int waldo_decode_fwd(const waldo_decode_fwd_t*, waldo_stream_t);

typedef struct {
  waldo_decode_fwd_t f;       /* the forward call's arguments (inputs, saved intermediates, outputs) */
  /* upstream gradients; any may be NULL (= zero) */
  const float* d_output;      /* (B, Tp, C, Hd, Wd)   gradient of out_full[:, :, :C]  (the returned `output`)    */
  const float* d_raw_alpha;   /* (B, Tp, 1, Hd, Wd)   gradient of out_full[:, :, C:]  (the returned `raw_alpha`) */
  const float* d_raw_output;  /* (B, Tc+self, Tp, C+L+disocc, Hd, Wd) */
  const float* d_flow;        /* (B, Tc, Tp, 2, Hd, Wd) */
  const float* d_alpha;       /* (B, Tw, L, Hd, Wd) */
  /* gradient outputs; NULL = not needed.  Buffers must be ZERO-FILLED by the caller. */
  float* d_input;             /* (B, T, C, Hd, Wd) */
  float* d_tgt_grid_obj;      /* (B, T, No, Ho, Wo, 2) */
  float* d_src_grid_obj;      /* (B, T, No, H, W, 2) */
  float* d_tgt_grid_bg;       /* (B, T, H, W, 2) */
  float* d_src_grid_bg;       /* (B, T, H, W, 2) */
  float* d_occ;               /* (B, T, L, L) */
  float* d_obj_alpha;         /* (B, No, Ho, Wo) */
  float* d_bg_alpha;          /* (B, H, W) */
  float* d_cls;               /* (B, No, Nl) */
  /* zero-filled scratch */
  float* d_alpha_acc;         /* (B, Tw, L, Hd, Wd) gradient w.r.t. the stored context opacity A */
  float* d_f_lo;              /* (B, Tc, Tp, L, H, W, 2) */
  float* d_a_lo;              /* (B, Tw, L, H, W) */
  float* d_prof_p;            /* (B, No, Nl) */
  float* d_prof_sum;          /* (B, No*Nl + No) */
  /* per-CTA partial sums of the small reductions (reduced in CTA order => deterministic); need no zero-fill */
  int red_ctas;               /* CTAs per (b,tp) / (b,t) group of the two HD backward kernels */
  float* occ_part;            /* (max(B*Tp, B*Tw), red_ctas, L*L) */
  float* prof_p_part;         /* (B*Tw, red_ctas, No*Nl) */
  float* cls_part;            /* (B, prof_ctas, No*Nl) */
  float* up_tab;              /* scratch (W, 9): per low-res column, first HD column touching it and its 8 x-weights */
  float* glue;                /* scratch (B, Tc, Tp, 3, Hd, Wd): d score, d flow x, d flow y handed from the gather to the layer kernel */
  int stages;                 /* 0 = everything; else bit 0 = HD gather backward, bit 1 = HD layer backward,
                                 bit 3 = HD context-alpha backward, bit 2 = the rest (low-res chain) */
  /* Deterministic accumulation (optional; det_shadow == NULL selects fire-and-forget float reductions, whose addition order
     -- like ATen's grid_sampler_2d_backward -- is not fixed).  With det_shadow set, every scatter target (d_input, the four
     d_*_grid_*, d_obj_alpha, d_bg_alpha, d_alpha_acc, d_f_lo, d_a_lo) is accumulated as 64-bit fixed point (integer
     additions commute: the result does not depend on the order) and converted to fp32 once all its addends are in.  All
     those targets must then be carved from ONE float arena [det_base, det_base + det_n); det_shadow is an int64 arena of
     det_n elements (element i shadows det_base[i]), zero-filled by the caller like the targets.  The fixed-point unit is
     2^-26 of the largest upstream gradient magnitude (rounded up to a power of two), found by the call itself; sums up to
     2^37 times that magnitude are representable. */
  const float* det_base;
  int64_t* det_shadow;
  int64_t det_n;
  float* det_scale;           /* device scratch, 4 floats: [0] scale, [1] 1/scale, [2] bits of max|upstream|, [3] != 0 after the
                                 call if an addend left the fixed-point range (the gradients are then invalid) */
} waldo_decode_bwd_t;
int waldo_decode_bwd(const waldo_decode_bwd_t*, waldo_stream_t);

/* ------------------------------------------------------------------ a-9  WIF.forward fuse tail
 * models/nets/wif.py:50-54: frame = sum_tc softmax_tc(u[3]) * (sigmoid(raw[4]+5) * raw[0:3] + u[0:3]). */
typedef struct {
  int B, Tc, Tp, Cr;          /* Cr = channels of raw_output */
  int HW;                     /* Hd*Wd */
  int ab;                     /* opt.ii_ab */
  const float* raw_output;    /* (B, Tc, Tp, Cr, HW) */
  const float* unet_out;      /* (B, Tp, Tc, 4+ab, HW) */
  float* frame;               /* out (B, Tp, 3, HW) */
} waldo_wif_fuse_fwd_t;
int waldo_wif_fuse_fwd(const waldo_wif_fuse_fwd_t*, waldo_stream_t);

typedef struct {
  waldo_wif_fuse_fwd_t f;
  const float* d_frame;       /* (B, Tp, 3, HW) */
  float* d_raw_output;        /* (B, Tc, Tp, Cr, HW) or NULL; channels 0..4 written, the rest must be pre-zeroed */
  float* d_unet_out;          /* (B, Tp, Tc, 4+ab, HW) or NULL */
} waldo_wif_fuse_bwd_t;
int waldo_wif_fuse_bwd(const waldo_wif_fuse_bwd_t*, waldo_stream_t);

/* ------------------------------------------------------------------ a-5 / a-11  stand-alone field warp and `scale`
 * Warper.obj_to_output / bg_to_output, models/nets/lvd.py:538-559: out = grid_sample(field + delta, grid) - delta
 * (bilinear, zeros, align_corners=False).  With `scale` (lvd.py:175-179) they make up the MAT propagation flows
 * grid_to_{bg,obj}_flow_from_{ref_to_pred,ctx_to_ref} (lvd.py:575-600).  Forward only (inference helpers; inside
 * decode_output the same steps are fused into waldo_decode_fwd/bwd). */
typedef struct {
  int n;                      /* items */
  int c;                      /* channels per item */
  int h, w;                   /* lattice of the field (canonical frame) */
  int H, W;                   /* lattice of the grid / output (image) */
  float delta;                /* lvd.py:548,559 */
  const float* field;         /* (n, c, h, w) */
  const float* grid;          /* (n, H, W, 2) */
  float* out;                 /* out (n, c, H, W) */
} waldo_warp_field_t;
int waldo_warp_field_fwd(const waldo_warp_field_t*, waldo_stream_t);

typedef struct {
  int n;                      /* planes */
  int h, w, H, W;             /* in / out sizes */
  const float* in;            /* (n, h, w) */
  float* out;                 /* out (n, H, W) */
} waldo_resize_t;
int waldo_resize_bilinear_fwd(const waldo_resize_t*, waldo_stream_t);

/* ------------------------------------------------------------------ f-3  input packing (caller side of the path)
 * data/base_dataset.py:173-183 (label map -> one-hot -> 5*(2x-1)), :355-372 (ToTensor, Normalize(0.5,0.5)) and
 * models/synthesizer.py:444 (input = cat([vid, lyt], dim=2)), done on the device so that only 8-bit planes cross PCIe. */
typedef struct {
  int n;                      /* frames B*T */
  int Nl;                     /* classes */
  int HW;                     /* Hd*Wd */
  float on, off;              /* 5, -5 in the reference */
  const uint8_t* rgb_u8;      /* (n, 3, HW) raw 8-bit RGB, or NULL */
  const float* rgb_f32;       /* (n, 3, HW) already normalised frames (used when rgb_u8 is NULL) */
  const uint8_t* label;       /* (n, HW) class ids; ids >= Nl light no channel */
  float* input;               /* out (n, 3+Nl, HW) */
  int storage;                /* element type of `input`: WALDO_ST_F32, or WALDO_ST_BF16 (the pointer then addresses bf16 elements;
                                 +-5 and the 8-bit colours k/127.5 - 1 round to bf16 once, here) */
} waldo_pack_input_t;
int waldo_pack_input(const waldo_pack_input_t*, waldo_stream_t);

/* ------------------------------------------------------------------ f-4  output side (caller side of the path)
 * tools/utils.py:246-249 (normalize: clamp to [lo, hi], rescale to [0, 1]) and :258-264 (dump_video: permute to
 * (T, H, W, 3), * 255, truncate to uint8), done on the device so that one byte per sample crosses PCIe and the layout
 * the video writer wants comes out of the kernel.  models/synthesizer.py:184-193 save_vid then only encodes. */
typedef struct {
  int n;                      /* frames */
  int HW;                     /* Hd*Wd */
  float lo, hi;               /* span, [-1, 1] in the reference */
  const float* frames;        /* (n, 3, HW) fp32, planar */
  uint8_t* out;               /* out (n, HW, 3) */
} waldo_frames_u8_t;
int waldo_frames_to_u8(const waldo_frames_u8_t*, waldo_stream_t);

/* ------------------------------------------------------------------ f-2  loss epilogues of LVD training (consumer side of the path)
 * models/synthesizer.py:1114-1118 `blur(vid, sigma, kernel_size)`: torchvision GaussianBlur = per-plane 2-D convolution
 * with the outer product of the normalised Gaussian taps, reflect padding (used at :893, :914, :946-947, :974-975).
 * waldo_blur_bwd is its adjoint: in = d out, out = d in. */
typedef struct {
  int n;                      /* planes (product of all leading dims) */
  int H, W;                   /* both > ksize / 2 (reflect padding) */
  int ksize;                  /* odd, <= 31; reference default 23 */
  float sigma;                /* reference default 3 */
  const float* in;            /* (n, H, W) */
  float* out;                 /* out (n, H, W) */
} waldo_blur_t;
int waldo_blur_fwd(const waldo_blur_t*, waldo_stream_t);
int waldo_blur_bwd(const waldo_blur_t*, waldo_stream_t);

/* models/synthesizer.py:886-889 (entropy of the normalised layer opacities, divided by 0.37) and :933 (fg_mask):
 *   x_k = (alpha_k + 1) / 2 + 1e-6;  p = x / max(sum_k |x_k|, 1e-12);  entropy = -sum_k p_k log(p_k + 1e-6) / 0.37
 *   fg = sum_{k >= 1} (alpha_k + 1) / 2 */
typedef struct {
  int n;                      /* B*T */
  int L;                      /* layers, layer 0 = background */
  int HW;
  const float* alpha;         /* (n, L, HW) in [-1, 1] */
  float* entropy;             /* out (n, HW) or NULL */
  float* fg;                  /* out (n, HW) or NULL */
} waldo_layer_entropy_t;
int waldo_layer_entropy_fwd(const waldo_layer_entropy_t*, waldo_stream_t);

typedef struct {
  waldo_layer_entropy_t f;    /* alpha as in the forward; the two outputs are not read */
  const float* d_entropy;     /* (n, HW) or NULL (= zero) */
  const float* d_fg;          /* (n, HW) or NULL (= zero) */
  float* d_alpha;             /* out (n, L, HW), written (not accumulated) */
} waldo_layer_entropy_bwd_t;
int waldo_layer_entropy_bwd(const waldo_layer_entropy_bwd_t*, waldo_stream_t);

/* models/synthesizer.py:965-979 (`cell_dis`, `center_dis` of LVD training): per low-res pixel g (warper.src_grid),
 *   cell_min   = min_o ( (mov + eps) (1 - fg) * sum_{cells k of object o} d(g, c_{o,k}) )     c = mean of the cell's 4 control points
 *   center_min = min_o ( mov * d(g, m_o) )                                                    m = mean of all control points of o
 *   d(g, c) = |g|^2 + |c|^2 - 2 c.g   (the reference's expanded form)
 * The reference's two scalars are the means of the two maps.  The (B, T, No, cells, H, W) distance tensor is never stored. */
typedef struct {
  int n;                      /* B*T, <= 65535 */
  int No, ho, wo;             /* objects (<= 32), control-point lattice of an object (opt.obj_shape), No (ho-1)(wo-1) <= 1024 */
  int HW;
  float eps;                  /* opt.cell_dis_eps */
  const float* mov;           /* (n, HW) mov_obj_mask (blurred) */
  const float* fg;            /* (n, HW) fg_mask (blurred) */
  const float* pose;          /* (n, No, ho*wo, 2) obj_pose */
  const float* grid;          /* (HW, 2) */
  float* cell_min;            /* out (n, HW) */
  float* center_min;          /* out (n, HW) */
  uint8_t* cell_arg;          /* out (n, HW) argmin object of cell_min (first index on ties), read by the backward */
  uint8_t* center_arg;       /* out (n, HW) */
} waldo_pose_dis_t;
int waldo_pose_dis_fwd(const waldo_pose_dis_t*, waldo_stream_t);

typedef struct {
  waldo_pose_dis_t f;         /* inputs and the two argmin maps as in the forward; cell_min / center_min are not read */
  const float* d_cell;        /* (n, HW) or NULL (= zero) */
  const float* d_center;      /* (n, HW) or NULL (= zero) */
  float* d_fg;                /* out (n, HW) or NULL */
  float* d_mov;               /* out (n, HW) or NULL (the reference's mask carries no gradient) */
  int ctas;                   /* CTAs per frame of the pixel pass, 1 .. 1024 */
  float* part;                /* scratch (n, ctas, No, 6): per-CTA sums, added in CTA order (deterministic) */
  float* d_pose;              /* out (n, No, ho*wo, 2), written (not accumulated) */
} waldo_pose_dis_bwd_t;
int waldo_pose_dis_bwd(const waldo_pose_dis_bwd_t*, waldo_stream_t);

/* models/synthesizer.py:864-868 (`obj_flow`, "same mean motion in layers"):
 *   a_o = (alpha_{o+1} + 1) / 2 + 1e-6 (object layers),  m_o = sum_p a_o f / sum_p a_o  (f = real_flow),
 *   dev_map(p) = sum_o a_o(p) (|fx(p) - mx_o| + |fy(p) - my_o|);   the reference's scalar is sum(dev_map) / (n (L-1) HW). */
typedef struct {
  int n;                      /* B*T, <= 65535 */
  int L;                      /* layers incl. the background layer 0 (skipped), 2 .. 33 */
  int HW;
  const float* alpha;         /* (n, L, HW) in [-1, 1] */
  const float* flow;          /* (n, 2, HW) */
  int ctas;                   /* CTAs per frame of the reduction passes, 1 .. 1024 */
  float* part;                /* scratch (n, ctas, L-1, 3): per-CTA sums, added in CTA order (deterministic) */
  float* mom;                 /* out (n, L-1, 3): sum a, sum a fx, sum a fy (read by the backward) */
  float* dev_map;             /* out (n, HW) */
} waldo_obj_flow_t;
int waldo_obj_flow_fwd(const waldo_obj_flow_t*, waldo_stream_t);

typedef struct {
  waldo_obj_flow_t f;         /* alpha, flow, mom as in / from the forward; part is scratch again; dev_map is not read */
  const float* d_map;         /* (n, HW) upstream gradient of dev_map */
  float* tsum;                /* scratch (n, L-1, 2) */
  float* d_alpha;             /* out (n, L, HW), written (layer 0 = 0) */
} waldo_obj_flow_bwd_t;
int waldo_obj_flow_bwd(const waldo_obj_flow_bwd_t*, waldo_stream_t);

/* ------------------------------------------------------------------ f-1 (the two full-resolution layers)  the consumer of raw_output
 * models/modules/conv.py:9-11, :36-37, :54, :63: UNet.to_emb = conv3x3(Cin -> 16) as WIF.forward applies it to raw_output
 * (models/nets/wif.py:33-38) and UNet.from_emb = conv3x3(2 x 16 -> 4 | 5); stride 1, padding 1, no bias.  TF32 tensor-core products, fp32 accumulation (what the reference's cuDNN convolution
 * does on this GPU with torch's default allow_tf32).  The backward-data of the same layer is this entry point again, called with the
 * flipped, transposed weights (Cout <-> Cin) and Tc <-> Tp (the image permute is its own inverse with the two swapped). */
typedef struct {
  int n;                      /* images */
  int Cin, Cout;              /* Cin <= 48, Cout <= 48 */
  int H, W;
  int Tc, Tp;                 /* both > 0: output image (b, tp, tc) reads input image (b, tc, tp) -- the permute of wif.py:33 folded into
                                 the addressing (n = B*Tc*Tp); 0, 0: output image i reads input image i */
  const float* in;            /* (n, Cin, H, W) */
  const float* weight;        /* (Cout, Cin, 3, 3) */
  float* out;                 /* out (n, Cout, H, W) */
} waldo_conv3x3_t;
int waldo_conv3x3_fwd(const waldo_conv3x3_t*, waldo_stream_t);

/* Weight gradient of the same layer: dweight[co][ci][ky][kx] = sum_{image, y, x} dout[image][co][y][x] * in[src(image)][ci][y+ky-1][x+kx-1]
 * (zero padding).  Per-CTA partial sums in registers over all tiles of the CTA, added in CTA order (deterministic).
 * Needs Cout <= 16, Cin <= 40, W % 4 == 0 and 16-byte aligned `in` / `dout`. */
typedef struct {
  waldo_conv3x3_t c;          /* the forward call's arguments (weight and out are not read) */
  const float* dout;          /* (n, Cout, H, W), in the forward's OUTPUT image order */
  int ctas;                   /* CTAs of the launch = number of partials, <= 296 */
  float* part;                /* scratch (ctas, Cout, Cin, 3, 3) */
  float* dweight;             /* out (Cout, Cin, 3, 3) */
} waldo_conv3x3_wgrad_t;
int waldo_conv3x3_wgrad(const waldo_conv3x3_wgrad_t*, waldo_stream_t);

#ifdef __cplusplus
}
#endif
#endif /* WALDO_B200_H */
