"""Drop-in replacements for the reference's hot-path modules, same constructor arguments, registered buffer
names (state-dict compatible), method names, positional signatures and return tuples:

    TPSWarp, InverseWarp      <- models/modules/warp.py:21-55, :58-174
    Warper                    <- models/nets/lvd.py:469-870
    compute_occ, decode_output, estimate_alpha_grid_occ   <- LVD.compute_occ :59-68, LVD.forward :126-153
    wif_fuse                  <- WIF.forward tail, models/nets/wif.py:50-54

Everything executes in the sm_100a library through waldo_b200.functional; there is no PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import functional as Fn


# ----------------------------------------------------------------------------- a-0 helpers (tools/utils.py:273-297)
def get_grid(height: int, width: int) -> torch.Tensor:
    """Pixel-centre normalised lattice (1,H,W,2), (x,y) last -- tools/utils.py:293-297."""
    xs = torch.linspace(-1.0 + 1.0 / width, 1.0 - 1.0 / width, width)
    ys = torch.linspace(-1.0 + 1.0 / height, 1.0 - 1.0 / height, height)
    return torch.stack([xs[None, :].expand(height, width), ys[:, None].expand(height, width)], dim=-1)[None]


def get_gaussian_kernel(k: int, sigma_div: float = 6) -> torch.Tensor:
    """k x k Gaussian with sigma = k / sigma_div, normalised to sum 1 -- tools/utils.py:273-291."""
    r = torch.arange(k, dtype=torch.float32) - (k - 1) / 2.0
    var = (k / sigma_div) ** 2.0
    ker = (1.0 / (2.0 * math.pi * var)) * torch.exp(-(r[None, :] ** 2 + r[:, None] ** 2) / (2 * var))
    return ker / ker.sum()


def kernel_distance(pts_1: torch.Tensor, pts_2: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """TPS radial basis 0.5 * d * log(d + eps) with d in expanded form -- models/modules/warp.py:15-18."""
    d = (pts_1 ** 2).sum(-1)[:, None] + (pts_2 ** 2).sum(-1)[None, :] - 2 * pts_1 @ pts_2.t()
    return 0.5 * d * torch.log(d + eps)


# ----------------------------------------------------------------------------- a-1
class TPSWarp(nn.Module):
    """models/modules/warp.py:21-55.  Buffers: inverse_kernel, pad, tgt_grid_repr."""

    def __init__(self, tgt_height, tgt_width, tgt_pts):
        super().__init__()
        self.tgt_shape = [tgt_height, tgt_width]
        pts = tgt_pts.float()
        n = pts.size(0)
        system = torch.zeros(n + 3, n + 3)
        system[:n, :n] = kernel_distance(pts, pts)
        system[:n, n] = 1
        system[n, :n] = 1
        system[:n, n + 1:] = pts
        system[n + 1:, :n] = pts.t()
        lattice = get_grid(tgt_height, tgt_width).view(-1, 2)
        repr_ = torch.cat([kernel_distance(lattice, pts), torch.ones(lattice.size(0), 1), lattice], dim=1)
        self.register_buffer("inverse_kernel", torch.inverse(system).contiguous())
        self.register_buffer("pad", torch.zeros(3, 2))
        self.register_buffer("tgt_grid_repr", repr_)

    def forward(self, src_pts):
        h, w = self.tgt_shape
        return Fn.tps_eval(src_pts, self.inverse_kernel, self.tgt_grid_repr, h, w)


# ----------------------------------------------------------------------------- a-2
class InverseWarp(nn.Module):
    """models/modules/warp.py:58-174.  Buffers: kernel, src_grid, tgt_grid, x_grid, y_grid, perm."""

    def __init__(self, src_height, src_width, tgt_height, tgt_width, kernel_size=3, num_perm=1):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("waldo_b200.InverseWarp: only the reference's 3x3 fill kernel is compiled")
        if num_perm != 1:
            raise NotImplementedError("waldo_b200.InverseWarp: num_perm > 1 (averaged permutations, warp.py:91-111) "
                                      "is used by no shipped script and is not built")
        self.kernel_size = kernel_size
        self.tgt_shape = [tgt_height, tgt_width]
        self.num_perm = num_perm
        self.register_buffer("kernel", get_gaussian_kernel(kernel_size).view(1, 1, kernel_size, kernel_size))
        self.register_buffer("src_grid", get_grid(src_height, src_width))
        self.register_buffer("tgt_grid", get_grid(tgt_height, tgt_width))
        self.register_buffer("x_grid", torch.arange(tgt_width).view(1, -1).repeat(tgt_height, 1).view(1, -1).float())
        self.register_buffer("y_grid", torch.arange(tgt_height).view(-1, 1).repeat(1, tgt_width).view(1, -1).float())
        self.register_buffer("perm", torch.stack([torch.randperm(tgt_height * tgt_width) for _ in range(num_perm)]))

    def forward(self, src_grid, niter=5, pad=True, erode=True, trace_box=None):
        if not pad:
            raise NotImplementedError("waldo_b200.InverseWarp: pad=False has no caller in the reference and is not built")
        h, w = self.tgt_shape
        return Fn.inverse_warp(src_grid, self.src_grid[0], self.tgt_grid[0], self.kernel.view(-1), h, w, niter, erode, trace_box)


OVERLAP_BG = True   # Warper.forward: background TPS + inverse warp on a side stream (see there)


# ----------------------------------------------------------------------------- a-3, a-5..a-7
class Warper(nn.Module):
    """models/nets/lvd.py:469-870.  Same option fields, buffers (src_pts, tgt_pts, src_grid, src_grid_hd, tgt_grid,
    tps_obj.*, invert_obj.*, tps_bg.*, invert_bg.*) and methods as the reference."""

    def __init__(self, opt, repeat_border=False):
        super().__init__()
        src_pts = get_grid(*opt.latent_shape).view(-1, 2)
        tgt_pts = get_grid(*opt.obj_shape).view(-1, 2)
        self.time_dropout = opt.time_dropout
        self.num_obj = opt.num_obj
        self.latent_obj_size = opt.obj_shape[0] * opt.obj_shape[1]
        self.latent_size = opt.latent_shape[0] * opt.latent_shape[1]
        self.tgt_shape = [int(opt.obj_shape[0] * opt.patch_size * opt.scale_factor),
                          int(opt.obj_shape[1] * opt.patch_size * opt.scale_factor)]
        self.src_shape = [opt.dim, int(opt.dim * opt.aspect_ratio)]
        self.src_shape_hd = [opt.load_dim, int(opt.load_dim * opt.aspect_ratio)] if opt.load_dim > 0 else self.src_shape
        self.register_buffer("src_pts", src_pts)
        self.register_buffer("tgt_pts", tgt_pts)
        self.register_buffer("src_grid", get_grid(*self.src_shape))
        self.register_buffer("src_grid_hd", get_grid(*self.src_shape_hd))
        self.register_buffer("tgt_grid", get_grid(*self.tgt_shape))
        self.tps_obj = TPSWarp(*self.tgt_shape, tgt_pts)
        self.invert_obj = InverseWarp(*self.tgt_shape, *self.src_shape, num_perm=opt.num_perm_grid)
        self.normalize_alpha = opt.normalize_alpha
        self.use_lyt_filtering = opt.use_lyt_filtering
        self.use_lyt_opacity = opt.use_lyt_opacity
        self.weight_cls = opt.weight_cls
        self.min_cls = opt.min_cls
        self.include_self = opt.include_self
        self.fast = opt.load_dim == 0
        self.scale_hd = opt.load_dim / opt.dim if opt.load_dim > 0 else 1
        self.tps_bg = TPSWarp(*self.src_shape, src_pts)
        self.invert_bg = InverseWarp(*self.src_shape, *self.src_shape, num_perm=opt.num_perm_grid)
        self.no_filter = opt.no_filter
        self.allow_ghost = opt.allow_ghost
        self.use_disocc = getattr(opt, "use_disocc", False)
        self._fused = None   # results of the last fused decode, handed out by input_to_output

    # -- a-3 ---------------------------------------------------------------------------------------------
    def forward(self, obj_pose, bg_pose, invert=True):
        """lvd.py:855-870 -> (tgt_grid_obj, src_grid_obj, tgt_grid_bg, src_grid_bg)."""
        B, T, No = obj_pose.shape[:3]
        Lo, Lb = self.latent_obj_size, self.latent_size
        # The background chain (40 items at the benchmark shape: small grids, one launch per dilation) and the object chain
        # (640 items) are independent: the background runs on a side stream and overlaps with the object kernels, forward
        # and -- autograd replays each node on its forward stream -- backward.  OVERLAP_BG = False serialises them.
        dev = obj_pose.device
        fork = OVERLAP_BG and obj_pose.is_cuda
        if fork:
            main, side = torch.cuda.current_stream(dev), Fn._side_stream(dev, 1)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                tgb = self.tps_bg(bg_pose.reshape(B * T, Lb, 2))
                sgb = self.invert_bg(tgb, erode=False) if invert else None
        tgo = self.tps_obj(obj_pose.reshape(B * T * No, Lo, 2))
        sgo = self.invert_obj(tgo) if invert else None
        if fork:
            # join.  No Tensor.record_stream on tgb / sgb: it would defer the release of their blocks until the host sees the
            # main stream pass this point, and the host runs many steps ahead of the GPU -- reserved memory then grows by a set
            # of buffers per step (measured: profiles/r2/r2_notes.md).  It is not needed either: the side stream touches memory
            # again only after its next `wait_stream(main)` (the next fork, or autograd's own synchronisation before a backward
            # node of this chain), i.e. after every main-stream consumer of these tensors has been enqueued before it.
            main.wait_stream(side)
        else:
            tgb = self.tps_bg(bg_pose.reshape(B * T, Lb, 2))
            sgb = self.invert_bg(tgb, erode=False) if invert else None
        tgo = tgo.view(B, T, No, *tgo.shape[1:])
        tgb = tgb.view(B, T, *tgb.shape[1:])
        if invert:
            sgo = sgo.view(B, T, No, *sgo.shape[1:])
            sgb = sgb.view(B, T, *sgb.shape[1:])
        return tgo, sgo, tgb, sgb

    # -- a-5 (stand-alone, forward only) ----------------------------------------------------------------
    def obj_to_output(self, obj, grid, delta_obj=1):
        """lvd.py:538-548: obj (B,No,c,Ho,Wo) or (B,T,No,c,Ho,Wo) warped by src_grid_obj -> (B,T,No,c,H,W)."""
        _, sgo, _, _ = grid
        B, T, No = sgo.shape[:3]
        Ho, Wo = self.tgt_shape
        H, W = self.src_shape
        c = obj.size(-3)
        obj = obj.view(B, 1, No, c, Ho, Wo).expand(-1, T, -1, -1, -1, -1) if obj.ndim == 5 else obj
        out = Fn.warp_field(obj.reshape(B * T * No, c, Ho, Wo), sgo.reshape(B * T * No, H, W, 2), delta_obj)
        return out.view(B, T, No, c, H, W)

    def bg_to_output(self, bg, grid, delta_bg=1, eps=1e-6):
        """lvd.py:550-559: bg (B,c,H,W) or (B,T,c,H,W) warped by src_grid_bg -> (B,T,1,c,H,W)."""
        _, _, _, sgb = grid
        B, T = sgb.shape[:2]
        H, W = self.src_shape
        c = bg.size(-3)
        bg = bg.view(B, 1, c, H, W).expand(-1, T, -1, -1, -1) if bg.ndim == 4 else bg
        out = Fn.warp_field(bg.reshape(B * T, c, H, W), sgb.reshape(B * T, H, W, 2), delta_bg)
        return out.view(B, T, 1, c, H, W)

    def layer_to_output(self, obj, bg, grid, delta_bg=1, delta_obj=1):
        """lvd.py:533-536."""
        return torch.cat([self.bg_to_output(bg, grid, delta_bg), self.obj_to_output(obj, grid, delta_obj)], dim=2)

    # -- a-11 MAT propagation flows (inference, --s_use_inpainter) -------------------------------------------
    def grid_to_bg_flow_from_ref_to_pred(self, grid, ctx_len, ref):
        """lvd.py:575-582 -> (B, T-ctx_len, Hd, Wd, 2)."""
        _, _, tgb, sgb = grid
        flow = (tgb[:, [ref]] - tgb[:, ctx_len:]).permute(0, 1, 4, 2, 3)
        flow = self.bg_to_output(flow, [None, None, None, sgb[:, ctx_len:]], delta_bg=0).squeeze(2)
        return Fn.resize_bilinear(flow, self.scale_hd).permute(0, 1, 3, 4, 2)

    def grid_to_obj_flow_from_ref_to_pred(self, grid, ctx_len, ref, obj_id):
        """lvd.py:584-591 -> (B, T-ctx_len, Hd, Wd, 2).  (The reference's `tgt_grid_obj[:, [ref], [obj_id]]` drops a dim and
        only broadcasts for B == 1, the batch size its inpainting path runs with; this is the B-general form.)"""
        tgo, sgo, _, _ = grid
        flow = (tgo[:, [ref]][:, :, [obj_id]] - tgo[:, ctx_len:][:, :, [obj_id]]).permute(0, 1, 2, 5, 3, 4)
        flow = self.obj_to_output(flow, [None, sgo[:, ctx_len:][:, :, [obj_id]], None, None], delta_obj=0).squeeze(2)
        return Fn.resize_bilinear(flow, self.scale_hd).permute(0, 1, 3, 4, 2)

    def grid_to_bg_flow_from_ctx_to_ref(self, grid, ctx_len, ref):
        """lvd.py:593-600 -> (B, ctx_len, Hd, Wd, 2)."""
        _, _, tgb, sgb = grid
        flow = (tgb[:, :ctx_len] - tgb[:, [ref]]).permute(0, 1, 4, 2, 3)
        flow = self.bg_to_output(flow, [None, None, None, sgb[:, [ref]].repeat(1, ctx_len, 1, 1, 1)], delta_bg=0).squeeze(2)
        return Fn.resize_bilinear(flow, self.scale_hd).permute(0, 1, 3, 4, 2)

    # -- a-6 + a-7 fused ---------------------------------------------------------------------------------
    def _spec(self, restrict_to_ctx: bool) -> Fn.DecodeSpec:
        H, W = self.src_shape
        Hd, Wd = self.src_shape_hd
        Ho, Wo = self.tgt_shape
        return Fn.DecodeSpec(H=H, W=W, Hd=Hd, Wd=Wd, Ho=Ho, Wo=Wo, num_obj=self.num_obj, restrict_to_ctx=restrict_to_ctx,
                             use_filter=True if restrict_to_ctx else not self.no_filter, weight_cls=bool(self.weight_cls),
                             allow_ghost=bool(self.allow_ghost), include_self=bool(self.include_self),
                             use_disocc=bool(self.use_disocc), min_cls=float(self.min_cls))

    def _decode(self, restrict, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts, want_disocc):
        """One fused decode.  Returns (output, raw_alpha, raw, flow, alpha, alpha_unflt, alpha_ctx, disocc); `raw` holds
        C + L (+1 with want_disocc) channels, alpha_ctx / disocc are channel-slice views of it."""
        self._fused = None   # never keep the HD tensors of an earlier call alive
        spec = self._spec(restrict)
        spec.use_disocc = bool(want_disocc)
        xs = self.src_grid_hd[0, 0, :, 0].contiguous()
        ys = self.src_grid_hd[0, :, 0, 1].contiguous()
        output, raw_alpha, raw, flow, alpha = Fn.decode(spec, ctx_ts, pred_ts, xs, ys, input, grid, occ, obj_alpha, bg_alpha, cls)
        C = input.size(2)
        Lr = self.num_obj + 1
        Tc = ctx_ts.size(1)
        alpha_ctx = raw[:, :Tc, :, C:C + Lr]                       # a channel-slice view of raw_output
        disocc = raw[:, :Tc, :, C + Lr:C + Lr + 1] if want_disocc else None
        alpha_unflt = alpha if self.fast else None                  # lvd.py:702-705 / :825-828
        return output, raw_alpha, raw, flow, alpha, alpha_unflt, alpha_ctx, disocc

    def _method_decode(self, restrict, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts):
        # Through the reference's METHOD interface the disocclusion map is always produced, as the reference does
        # (lvd.py:680 / :803): its caller decides whether to append it (lvd.py:148-151).
        output, raw_alpha, raw, flow, alpha, alpha_unflt, alpha_ctx, disocc = self._decode(
            restrict, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts, want_disocc=True)
        self._fused = (alpha_ctx, flow, output, raw_alpha, raw[:, :, :, :input.size(2) + self.num_obj + 1])
        return flow, alpha_unflt, alpha, alpha_ctx, disocc

    def grid_to_flow_ctx(self, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts):
        """lvd.py:707-828.  Returns (flow, alpha_unflt, alpha, alpha_ctx, disocc), `disocc` (B,Tc,Tp,1,Hd,Wd) always, as
        the reference."""
        return self._method_decode(True, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts)

    def grid_to_flow(self, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts):
        """lvd.py:602-705."""
        return self._method_decode(False, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts)

    def input_to_output(self, input, alpha, flow, ctx_ts, eps=1e-6):
        """lvd.py:830-853 -> (output (B,Tp,C+1,Hd,Wd), raw_output (B,Tc[+1],Tp,C+L,Hd,Wd)), both plain tensors exactly as
        the reference returns them (raw_output WITHOUT the disocc channel: the unmodified LVD.forward appends it itself,
        lvd.py:148-151).  The warp + context fusion was already done by the fused kernel of grid_to_flow[_ctx]; this hands
        its results out, which is why it must be called with the tensors that call returned, as LVD.forward does
        (lvd.py:143-146).  waldo_b200.decode_output() is the copy-free form of the same pair of calls."""
        f = self._fused
        self._fused = None
        if f is None or f[0] is not alpha or f[1] is not flow:
            raise NotImplementedError(
                "waldo_b200.Warper.input_to_output must be called with the (alpha_ctx, flow) tensors returned by the "
                "immediately preceding grid_to_flow[_ctx] call, as LVD.forward does (lvd.py:143-146): the warp of the "
                "context frames is fused into that kernel.")
        return torch.cat([f[2], f[3]], dim=2), f[4]


# ----------------------------------------------------------------------------- a-4, a-8
def compute_occ(occ_score, eps=1e-6):
    """LVD.compute_occ, lvd.py:59-68."""
    if eps != 1e-6:
        raise NotImplementedError("waldo_b200.compute_occ: eps is compiled as 1e-6 (the reference default)")
    return Fn.compute_occ(occ_score)


def decode_output(warper: Warper, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts,
                  restrict_to_ctx: bool, use_disocc: Optional[bool] = None, include_self: Optional[bool] = None):
    """LVD.forward(mode="decode_output"), lvd.py:141-153 -> the reference's 7-tuple
    (output, flow, alpha_unflt, alpha, raw_alpha, raw_output, alpha_ctx), copy-free: `output` / `raw_alpha` are the two
    channel slices of one buffer, `raw_output` already carries the disocc channel when opt.use_disocc, `alpha_ctx` is a
    view of it.  `use_disocc` / `include_self` override the warper's option fields (LVD.use_disocc / LVD.include_self in
    the reference, lvd.py:21-22) for this and later calls."""
    if use_disocc is not None:
        warper.use_disocc = bool(use_disocc)
    if include_self is not None:
        warper.include_self = bool(include_self)
    output, raw_alpha, raw, flow, alpha, alpha_unflt, alpha_ctx, _ = warper._decode(
        restrict_to_ctx, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts, want_disocc=warper.use_disocc)
    return output, flow, alpha_unflt, alpha, raw_alpha, raw, alpha_ctx


def alpha_masks(opt):
    """The constant buffers LVD.__init__ builds at lvd.py:25-44: (obj_alpha_mask (1,1,1,Ho,Wo) or 1, bg_alpha (1,1,H,W))."""
    Ho = int(opt.obj_shape[0] * opt.patch_size * opt.scale_factor)
    Wo = int(opt.obj_shape[1] * opt.patch_size * opt.scale_factor)
    mask = 1
    if opt.pad_obj_alpha > 0:
        p = int(opt.pad_obj_alpha * opt.scale_factor)
        mask = torch.ones(Ho, Wo)
        mask[:p] = 0
        mask[-p:] = 0
        mask[:, :p] = 0
        mask[:, -p:] = 0
        mask = mask.view(1, 1, 1, Ho, Wo)
    bg = torch.ones(1, 1, opt.dim, int(opt.dim * opt.aspect_ratio))
    if opt.pad_bg_alpha > 0:
        p = int(opt.pad_bg_alpha * opt.scale_factor)
        bg[:, :, :p] = -1
        bg[:, :, -p:] = -1
        bg[:, :, :, :p] = -1
        bg[:, :, :, -p:] = -1
    return mask, bg


def estimate_alpha_grid_occ(warper: Warper, obj_alpha, obj_alpha_mask, bg_alpha_buf, obj_pose, bg_pose, occ_score):
    """LVD.forward(mode="estimate_alpha_grid_occ") after the decoder, lvd.py:127-135
    -> (occ, obj_alpha, bg_alpha, grid)."""
    bg_alpha = bg_alpha_buf.expand(obj_alpha.size(0), -1, -1, -1)
    obj_alpha = obj_alpha_mask * obj_alpha + (1 - obj_alpha_mask) * (-1.0)
    grid = warper(obj_pose, bg_pose)
    occ = compute_occ(occ_score)
    return occ, obj_alpha, bg_alpha, grid


# ----------------------------------------------------------------------------- a-9
def wif_fuse(vid, unet_out, ab=True):
    """Tail of WIF.forward (wif.py:50-54).  vid = raw_output (B,Tc,Tp,Cin,H,W) exactly as WIF.forward receives it;
    unet_out = UNet output reshaped (B,Tp,Tc,5|4,H,W).  Returns (B,Tp,3,H,W)."""
    return Fn.wif_fuse(vid, unet_out, ab)


# ----------------------------------------------------------------------------- f-3 (caller side: the input pipeline)
def pack_input(rgb, label, num_lyt, out=None, dtype=torch.float32):
    """`input` = cat([vid, lyt], dim=2) (synthesizer.py:444) built on the device from 8-bit RGB (or normalised fp32
    frames) and the 8-bit label map, i.e. base_dataset.py:173-183 / :355-372 moved after the host->device copy.
    dtype=torch.bfloat16 builds the input of the bf16-storage inference variant of decode_output."""
    return Fn.pack_input(rgb, label, num_lyt, out=out, dtype=dtype)


# ----------------------------------------------------------------------------- f-4 (caller side: the output pipeline)
def frames_to_u8(vid, span=(-1.0, 1.0), out=None):
    """What dump_video does to a predicted video before encoding (tools/utils.py:246-249, :258-264), on the device:
    (..., 3, H, W) fp32 -> (..., H, W, 3) uint8.  save_vid (synthesizer.py:184-193) then copies a quarter of the bytes."""
    return Fn.frames_to_u8(vid, span, out)


# ----------------------------------------------------------------------------- f-2 (consumer side: LVD-training loss epilogues)
def blur(vid, sigma=3.0, kernel_size=23):
    """models/synthesizer.py:1114-1118: Gaussian blur (torchvision GaussianBlur, reflect padding) of every (H, W) plane of
    vid (..., C, H, W); differentiable (the adjoint folds the reflected margins back)."""
    return Fn.blur(vid, sigma, kernel_size)


def layer_entropy(alpha):
    """models/synthesizer.py:886-889 and :933 in one pass over the layer stack alpha (B, T, L, H, W) in [-1, 1]:
    (entropy (B, T, 1, H, W) = -sum_k p_k log(p_k + 1e-6) / 0.37 with p = normalize((alpha + 1) / 2 + 1e-6, p=1),
     fg_mask (B, T, 1, H, W) = sum_{k >= 1} (alpha_k + 1) / 2); differentiable."""
    return Fn.layer_entropy(alpha)


def pose_distance_losses(mov_obj_mask, fg_mask, obj_pose, grid, obj_shape, cell_dis_eps):
    """models/synthesizer.py:965-979: `cell_dis` and `center_dis` of LVD training -- the mean over pixels of the minimum over
    the objects of (mov_obj_mask + eps)(1 - fg_mask) * (summed squared distance of the pixel to the object's cell centres) and of
    mov_obj_mask * (squared distance to the object's mean control point) -- without the reference's (B, T, No, cells, H, W)
    distance tensor.  mov_obj_mask, fg_mask (B, T, 1, H, W) (already blurred if opt.blur_alpha); obj_pose (B, T, No, ho*wo, 2);
    grid = warper.src_grid (1, H, W, 2).  Differentiable in fg_mask and obj_pose (and mov_obj_mask)."""
    cell, center, _, _ = Fn.pose_distances(mov_obj_mask, fg_mask, obj_pose, grid, obj_shape, cell_dis_eps)
    return cell.mean(), center.mean()


def obj_flow_loss(rec_output_alpha, real_flow):
    """models/synthesizer.py:864-868 (`obj_flow`): mean over (B, T, No, H, W) of a_o |real_flow - mean flow of layer o|_1 with
    a_o = (alpha_{o+1} + 1) / 2 + 1e-6, without the reference's (B, T, No, 2, H, W) products.
    rec_output_alpha (B, T, No+1, H, W), real_flow (B, T, 2, H, W); differentiable in rec_output_alpha."""
    dev = Fn.obj_flow_map(rec_output_alpha, real_flow)
    return dev.sum() / (dev.numel() * (rec_output_alpha.shape[-3] - 1))


# ----------------------------------------------------------------------------- f-1, first layer (consumer side: WIF's UNet)
def wif_to_emb(raw_output, weight):
    """`UNet.to_emb` as WIF.forward applies it to raw_output (models/nets/wif.py:33-38 + models/modules/conv.py:54):
    raw_output (B, Tc, Tp, Cin, H, W) straight from decode_output, weight = unet.to_emb.weight (Cout, Cin, 3, 3)
    -> (B*Tp*Tc, Cout, H, W), the first feature map of the UNet.  The permute of wif.py:33 is folded into the kernel's
    addressing.  TF32 products / fp32 accumulation.  Differentiable: `d raw_output` (the upstream gradient of the warp
    backward) comes from the same kernel with the flipped, transposed weights; `d weight` from torch."""
    return Fn.conv3x3(raw_output, weight, wif_permute=True)


def conv3x3(x, weight):
    """models/modules/conv.py:9-11 conv3x3 (stride 1, padding 1, no bias) on (n, Cin <= 48, H, W), Cout <= 48 -- the UNet's two
    full-resolution layers, `to_emb` and `from_emb` (conv.py:36-37); differentiable."""
    return Fn.conv3x3(x, weight)


class Conv3x3(nn.Module):
    """Drop-in for `conv3x3(in_planes, out_planes)` at stride 1 (models/modules/conv.py:9-11: nn.Conv2d(kernel_size=3, padding=1,
    bias=False)): same parameter name and shape (`weight`, (out, in, 3, 3)) -- state dicts are interchangeable --, same default
    initialisation; forward and both gradients on waldo_conv3x3_fwd / waldo_conv3x3_wgrad."""

    def __init__(self, in_planes, out_planes):
        super().__init__()
        if in_planes > 48 or out_planes > 48:
            raise NotImplementedError("waldo_b200.Conv3x3: up to 48 input and 48 output channels are compiled")
        ref = nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=1, padding=1, bias=False)   # the reference's initialisation
        self.weight = nn.Parameter(ref.weight.detach().clone())
        self.in_channels, self.out_channels = in_planes, out_planes

    def forward(self, x):
        return Fn.conv3x3(x, self.weight)
