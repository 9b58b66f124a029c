"""torch.autograd bindings of the C ABI (include/waldo_b200.h).

PyTorch here is plumbing only: it owns device memory and the stream, and chains the backward entry points.
Every op allocates its outputs / saved state / scratch with torch and hands raw pointers to the library.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib as L


def _c(t: Optional[torch.Tensor], dtype=torch.float32):
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


# ===================================================================================== a-1 TPS
class _TpsEval(torch.autograd.Function):
    """TPSWarp.forward, models/modules/warp.py:49-55."""

    @staticmethod
    def forward(ctx, src_pts, inverse_kernel, tgt_grid_repr, h, w):
        lib = L.load()
        pts = _c(src_pts.detach())
        inverse_kernel, tgt_grid_repr = _c(inverse_kernel), _c(tgt_grid_repr)
        n, N = pts.shape[0], pts.shape[1]
        P = h * w
        grid = torch.empty(n, h, w, 2, device=pts.device, dtype=torch.float32)
        mapping = torch.empty(n, N + 3, 2, device=pts.device, dtype=torch.float64)
        a = L.TpsFwd(n, N, P, L.ptr(inverse_kernel, name="inverse_kernel"), L.ptr(tgt_grid_repr, name="tgt_grid_repr"),
                     L.ptr(pts, name="src_pts"), L.ptr(mapping, torch.float64), L.ptr(grid))
        L.call(lib.waldo_tps_fwd, a, pts, "tps_fwd")
        ctx.save_for_backward(inverse_kernel, tgt_grid_repr)
        ctx.dims = (n, N, P)
        return grid

    @staticmethod
    def backward(ctx, dgrid):
        lib = L.load()
        inverse_kernel, tgt_grid_repr = ctx.saved_tensors
        n, N, P = ctx.dims
        dgrid = _c(dgrid)
        chunks = max(1, min(256, P // 128))   # ordered slices of the pixel reduction: short serial loops, many CTAs
        partial = torch.empty(n, chunks, N + 3, 2, device=dgrid.device, dtype=torch.float64)
        dpts = torch.empty(n, N, 2, device=dgrid.device, dtype=torch.float32)
        a = L.TpsBwd(n, N, P, L.ptr(inverse_kernel), L.ptr(tgt_grid_repr), L.ptr(dgrid), chunks,
                     L.ptr(partial, torch.float64), L.ptr(dpts))
        L.call(lib.waldo_tps_bwd, a, dgrid, "tps_bwd")
        return dpts, None, None, None, None


def tps_eval(src_pts, inverse_kernel, tgt_grid_repr, h, w):
    return _TpsEval.apply(src_pts, inverse_kernel, tgt_grid_repr, h, w)


# ===================================================================================== a-2 inverse warp
@dataclass
class InverseWarpTrace:
    """Index maps of one inverse warp (parity rule 1: bit-exact)."""
    field: torch.Tensor    # (n, Ht*Wt) int32, -1 = outside                 warp.py:84-88
    winner: torch.Tensor   # (n, Ht*Wt) int32, INT32_MAX = empty cell       warp.py:113-123
    level: torch.Tensor    # (n, Hp, Wp) uint8
    eroded: torch.Tensor   # (n, Hp, Wp) uint8


class _InverseWarp(torch.autograd.Function):
    """InverseWarp.forward, models/modules/warp.py:71-174 (num_perm == 1, pad=True)."""

    @staticmethod
    def forward(ctx, fwd_grid, id_src, id_tgt, gauss, tgt_h, tgt_w, niter, erode, trace_box):
        lib = L.load()
        fg = _c(fwd_grid.detach())
        n, Hs, Ws, _ = fg.shape
        dev = fg.device
        m = niter + 1
        Hp, Wp = tgt_h + 2 * m, tgt_w + 2 * m
        out = torch.empty(n, tgt_h, tgt_w, 2, device=dev, dtype=torch.float32)
        field = torch.empty(n, tgt_h * tgt_w, device=dev, dtype=torch.int32)
        winner = torch.empty(n, tgt_h * tgt_w, device=dev, dtype=torch.int32)
        level = torch.empty(n, Hp, Wp, device=dev, dtype=torch.uint8)
        eroded = torch.empty(n, Hp, Wp, device=dev, dtype=torch.uint8)
        val = torch.empty(n, 2, Hp, Wp, device=dev, dtype=torch.float32)
        bbox = torch.empty(n, 4, device=dev, dtype=torch.int32)
        id_src_c, id_tgt_c, gauss_c = _c(id_src), _c(id_tgt), _c(gauss)
        a = L.InvWarpFwd(n, Hs, Ws, tgt_h, tgt_w, niter, 1 if erode else 0, L.ptr(fg, name="src_grid"), L.ptr(id_src_c),
                         L.ptr(id_tgt_c), L.ptr(gauss_c), L.ptr(out), L.ptr(field, torch.int32), L.ptr(winner, torch.int32),
                         L.ptr(level, torch.uint8), L.ptr(eroded, torch.uint8), L.ptr(val), L.ptr(bbox, torch.int32))
        L.call(lib.waldo_invwarp_fwd, a, fg, "invwarp_fwd")
        ctx.save_for_backward(gauss_c, field, winner, level, eroded, bbox)
        ctx.dims = (n, Hs, Ws, tgt_h, tgt_w, niter)
        if trace_box is not None:
            trace_box.append(InverseWarpTrace(field, winner, level, eroded))
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = L.load()
        gauss, field, winner, level, eroded, bbox = ctx.saved_tensors
        n, Hs, Ws, Ht, Wt, niter = ctx.dims
        dout = _c(dout)
        dev = dout.device
        m = niter + 1
        PP = (Ht + 2 * m) * (Wt + 2 * m)
        gval = torch.empty(n, 2, PP, device=dev, dtype=torch.float32)
        inv_sw = torch.empty(n, PP, device=dev, dtype=torch.float32)
        gdisp = torch.empty(n, Ht * Wt, 2, device=dev, dtype=torch.float32)
        dfwd = torch.empty(n, Hs, Ws, 2, device=dev, dtype=torch.float32)
        a = L.InvWarpBwd(n, Hs, Ws, Ht, Wt, niter, L.ptr(gauss), L.ptr(dout), L.ptr(field, torch.int32),
                         L.ptr(winner, torch.int32), L.ptr(level, torch.uint8), L.ptr(eroded, torch.uint8),
                         L.ptr(bbox, torch.int32), L.ptr(gval), L.ptr(inv_sw), L.ptr(gdisp), L.ptr(dfwd))
        L.call(lib.waldo_invwarp_bwd, a, dout, "invwarp_bwd")
        return dfwd, None, None, None, None, None, None, None, None


def inverse_warp(fwd_grid, id_src, id_tgt, gauss, tgt_h, tgt_w, niter=5, erode=True, trace_box=None):
    return _InverseWarp.apply(fwd_grid, id_src, id_tgt, gauss, tgt_h, tgt_w, niter, erode, trace_box)


# ===================================================================================== a-4 occlusion matrix
class _ComputeOcc(torch.autograd.Function):
    """LVD.compute_occ, models/nets/lvd.py:59-68."""

    @staticmethod
    def forward(ctx, occ_score):
        lib = L.load()
        s = _c(occ_score.detach())
        B, T, No = s.shape
        occ = torch.empty(B, T, No + 1, No + 1, device=s.device, dtype=torch.float32)
        with L.device_of(s.device):
            L.check(lib.waldo_occ_fwd(B * T, No, L.ptr(s, name="occ_score"), L.ptr(occ), L.stream_of(s)), "occ_fwd")
        ctx.save_for_backward(s)
        return occ

    @staticmethod
    def backward(ctx, docc):
        lib = L.load()
        (s,) = ctx.saved_tensors
        B, T, No = s.shape
        docc = _c(docc)
        ds = torch.empty_like(s)
        with L.device_of(s.device):
            L.check(lib.waldo_occ_bwd(B * T, No, L.ptr(s), L.ptr(docc), L.ptr(ds), L.stream_of(s)), "occ_bwd")
        return ds


def compute_occ(occ_score):
    return _ComputeOcc.apply(occ_score)


# ===================================================================================== a-5..a-8 decode_output
@dataclass
class DecodeSpec:
    """Static description of one decode call (everything that is not a tensor)."""
    H: int
    W: int
    Hd: int
    Wd: int
    Ho: int
    Wo: int
    num_obj: int
    restrict_to_ctx: bool
    use_filter: bool
    weight_cls: bool
    allow_ghost: bool
    include_self: bool
    use_disocc: bool
    min_cls: float
    occ_pairs_only: bool = False   # occ comes straight from compute_occ: its backward reads object-object pairs only


def _geom(spec: DecodeSpec, B, T, Tc, Tp, Nl, has_cls):
    Tw = Tc if spec.restrict_to_ctx else T
    flags = 0
    if spec.restrict_to_ctx:
        flags |= L.F_RESTRICT_CTX
    if spec.use_filter:
        flags |= L.F_FILTER
    if spec.weight_cls:
        flags |= L.F_WEIGHT_CLS
    if has_cls:
        flags |= L.F_HAS_CLS
    if spec.restrict_to_ctx and not spec.allow_ghost:
        flags |= L.F_IS_OBJ
    if spec.include_self:
        flags |= L.F_INCLUDE_SELF
    if spec.use_disocc:
        flags |= L.F_USE_DISOCC
    if spec.occ_pairs_only:
        flags |= L.F_OCC_PAIRS
    return L.Geom(B, T, Tw, Tc, Tp, spec.num_obj, Nl, 3 + Nl, spec.H, spec.W, spec.Hd, spec.Wd, spec.Ho, spec.Wo,
                  flags, float(spec.min_cls))


# Deterministic gradients (BASELINE north star: "deterministic gradient accumulation ... instead of global atomics").
# Default off: the scatter targets of decode_bwd are then accumulated with fire-and-forget fp32 reductions whose order is
# not fixed (reproducible to fp32 rounding only, like ATen's grid_sampler_2d_backward).  On -- set_deterministic(True), or
# torch.use_deterministic_algorithms(True) -- they are accumulated as 64-bit fixed point (order-independent), so every
# gradient of the path is bit-identical from run to run; costs an int64 shadow of the scatter targets and a few
# conversion passes (include/waldo_b200.h, det_* fields).
_DETERMINISTIC = False
LAST_DET_SCALE = None   # (4,) device tensor of the last deterministic backward: scale, 1/scale, max|upstream|, overflow flag


def set_deterministic(on: bool = True):
    global _DETERMINISTIC
    _DETERMINISTIC = bool(on)


def is_deterministic() -> bool:
    return _DETERMINISTIC or torch.are_deterministic_algorithms_enabled()


def _scatter_targets(refs, det, fill=True):
    """Zero-filled gradient / scratch buffers shaped like `refs` (None stays None).  In deterministic mode they are views
    of ONE float arena (16-byte aligned each) shadowed by an int64 arena: returns (targets, arena, shadow).  fill=False: the
    caller zero-fills (the targets, or the two arenas in deterministic mode)."""
    new = torch.zeros if fill else torch.empty
    if not det:
        return [(torch.zeros_like(r) if fill else torch.empty_like(r)) if r is not None else None for r in refs], None, None
    live = [r for r in refs if r is not None]
    if not live:   # nothing to accumulate: the call degenerates to the default path
        return [None] * len(refs), None, None
    offs, n = [], 0
    for r in live:
        offs.append(n)
        n += (r.numel() + 3) // 4 * 4
    dev = live[0].device
    arena = new(max(n, 4), device=dev, dtype=torch.float32)
    shadow = new(max(n, 4), device=dev, dtype=torch.int64)
    it = iter(offs)
    out = []
    for r in refs:
        if r is None:
            out.append(None)
        else:
            o = next(it)
            out.append(arena[o:o + r.numel()].view(r.shape))
    return out, arena, shadow


# The scatter targets of decode_bwd must be zero-filled (1.9 GB of `d input` + 1.1 GB of context-opacity gradient at the
# benchmark shape: 0.5 ms of pure HBM writes), by default at the start of the backward, on its stream.
# PREFILL = True (or WALDO_PREFILL=1) issues the fills at the end of the FORWARD on a side stream instead (see
# _Decode.forward), so that they overlap with the caller's work between forward and backward.  It is OFF by default: the
# buffers then carry a Tensor.record_stream mark, which defers the release of their blocks until the host observes the side
# stream's event -- with the host running many steps ahead of the GPU the caching allocator keeps allocating fresh 3 GB
# sets (reserved memory 13 -> 89 GiB over 40 steps on the B200, profiles/r2/r2_notes.md) -- and inside a step with nothing
# between forward and backward the overlap gains nothing (the fills compete with the forward kernels for SM slots).
PREFILL = os.environ.get("WALDO_PREFILL", "") == "1"
_side_streams = {}


def _side_stream(dev, which=0):
    """Side streams of a device: 0 = gradient-buffer fills during the forward, 1 = background chain of Warper.forward."""
    dev = torch.device(dev)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    s = _side_streams.get((dev, which))
    if s is None:
        s = _side_streams[(dev, which)] = torch.cuda.Stream(dev)
    return s


def _scatter_plan(need, g, cls_present):
    """Which gradients / scratch buffers a backward of this call needs (need = ctx.needs_input_grad[5:])."""
    n_inp, n_tgo, n_sgo, n_tgb, n_sgb, n_occ, n_oa, n_ba, n_cls = need
    n_cls = n_cls and cls_present
    filt = bool(g.flags & L.F_FILTER)
    geom = n_tgo or n_sgo or n_tgb or n_sgb
    chain = geom or n_occ or n_oa or n_ba or n_cls or (filt and n_inp)
    return dict(n_inp=n_inp, n_tgo=n_tgo, n_sgo=n_sgo, n_tgb=n_tgb, n_sgb=n_sgb, n_occ=n_occ, n_oa=n_oa, n_ba=n_ba, n_cls=n_cls,
                filt=filt, geom=geom, chain=chain)


def _alloc_scatter(plan, refs, det, fill=True):
    inp_c, alpha, f_lo, a_lo, tgo_c, sgo_c, tgb_c, sgb_c, oa_c, ba_c = refs
    on = lambda ref, cond: ref if cond else None
    p = plan
    return _scatter_targets(
        [on(inp_c, p["n_inp"]), on(alpha, p["chain"]), on(f_lo, p["geom"]), on(a_lo, p["chain"]), on(tgo_c, p["geom"]), on(sgo_c, p["geom"]),
         on(tgb_c, p["geom"]), on(sgb_c, p["geom"]), on(oa_c, p["n_oa"]), on(ba_c, p["n_ba"])], det, fill)


# When bench.py sets PROFILE = {"decode_fwd": [], "decode_bwd": []}, the two entry points are issued stage by stage
# (same kernels, same order, same stream) with a CUDA-event pair around the dominant fused HD kernel, so that its
# duration can be read inside the timed region (the roofline figure).  None = one call per entry point.
PROFILE = None


def _staged(fn, arg, stream, what, dev):
    if PROFILE is None:
        with L.device_of(dev):
            L.check(fn(C.byref(arg), stream), what)
        return
    # forward: low-res prep (1), context alpha (8), layers (2), gather (4); backward: gather (1), layers (2), context alpha (8),
    # rest (4) -- kernels in the same dependency order as with stages = 0
    order = (((1, "prep"), (8, "alpha_prep"), (2, "layers"), (4, "gather")) if what == "decode_fwd" else
             ((1, "gather"), (2, "layers"), (8, "alpha_prep"), (4, "rest")))
    for bit, key in order:
        arg.stages = bit
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        with L.device_of(dev):
            L.check(fn(C.byref(arg), stream), what)
        e1.record(torch.cuda.current_stream(dev))
        PROFILE.setdefault(what + ":" + key, []).append((e0, e1))
    arg.stages = 0


def _fwd_struct(g, prof_ctas, t):
    (inp_c, tgo_c, sgo_c, tgb_c, sgb_c, occ_c, oa_c, ba_c, cls_c, ts_c, ps_c, xs_hd, ys_hd, a_lo, prof_part, prof_sum, prof_p,
     f_lo, s_lo, live_ctx, live_pred, alpha, flow, raw, out_full, norm, score, lyt_lo) = t
    st = inp_c.dtype   # storage type of the HD activations: fp32, or bf16 (forward / inference only)
    return L.DecodeFwd(g, L.ptr(inp_c, st, name="input"), L.ptr(tgo_c), L.ptr(sgo_c), L.ptr(tgb_c), L.ptr(sgb_c), L.ptr(occ_c),
                       L.ptr(oa_c), L.ptr(ba_c), L.ptr(cls_c), L.ptr(ts_c, torch.int64), L.ptr(ps_c, torch.int64),
                       L.ptr(xs_hd), L.ptr(ys_hd), L.ptr(a_lo), L.ptr(prof_part), prof_ctas, L.ptr(prof_sum),
                       L.ptr(prof_p), L.ptr(lyt_lo), L.ptr(f_lo), L.ptr(s_lo), L.ptr(live_ctx, torch.int32), L.ptr(live_pred, torch.int32),
                       L.ptr(alpha), L.ptr(flow), L.ptr(raw, st), L.ptr(out_full, st), L.ptr(norm), L.ptr(score), 0,
                       L.ST_BF16 if st == torch.bfloat16 else L.ST_F32)


_ts_checked = {}


def _check_time_indices(ts_user, ps_user, ts, ps, Tw, T):
    """The kernels index frames with ctx_ts / pred_ts unchecked: validate them here.  Reading the extrema back is a
    device synchronisation, so the verdict is remembered for as long as the caller passes the very same (unmodified)
    tensor objects -- the time indices of a rollout or training loop are the same tensors step after step."""
    import weakref
    c = _ts_checked
    if (c.get("ts") is not None and c["ts"]() is ts_user and c["ps"]() is ps_user and
            c["ver"] == (ts_user._version, ps_user._version, Tw, T)):
        return
    ext = torch.stack([ts.max(), ts.min(), ps.max(), ps.min()]).tolist()   # one read-back
    if ext[0] >= Tw or ext[1] < 0 or ext[2] >= T or ext[3] < 0:
        raise RuntimeError("waldo_b200.decode: ctx_ts / pred_ts out of range")
    c["ts"], c["ps"], c["ver"] = weakref.ref(ts_user), weakref.ref(ps_user), (ts_user._version, ps_user._version, Tw, T)


class _Decode(torch.autograd.Function):
    """LVD.forward(mode="decode_output") = Warper.grid_to_flow[_ctx] + Warper.input_to_output, lvd.py:141-153.

    Returns (output (B,Tp,C,Hd,Wd), raw_alpha (B,Tp,1,Hd,Wd) -- two channel-slice views of one (B,Tp,C+1,Hd,Wd) buffer,
    lvd.py:147,152 --, raw_output (B,Tc+self,Tp,C+L+disocc,Hd,Wd), flow (B,Tc,Tp,2,Hd,Wd), alpha (B,Tw,L,Hd,Wd))."""

    @staticmethod
    def forward(ctx, spec, ctx_ts, pred_ts, xs_hd, ys_hd, inp, tgo, sgo, tgb, sgb, occ, obj_alpha, bg_alpha, cls):
        lib = L.load()
        ctx.set_materialize_grads(False)   # unused outputs must not cost a zero-filled HD tensor each
        # bf16 `input` selects the bf16-storage variant: input / raw_output / output are bf16 in HBM (alpha stays fp32), all
        # arithmetic is fp32 (include/waldo_b200.h WALDO_ST_BF16; tolerance: tests/parity.py TOL_BF16).  Forward / inference only.
        st = torch.bfloat16 if inp.dtype == torch.bfloat16 else torch.float32
        if st == torch.bfloat16 and any(ctx.needs_input_grad):
            raise RuntimeError("waldo_b200.decode: bf16 storage is forward / inference only (run under torch.no_grad(), or pass fp32 input)")
        inp_c = _c(inp.detach(), st)
        tgo_c, sgo_c, tgb_c, sgb_c = (_c(t.detach()) for t in (tgo, sgo, tgb, sgb))
        occ_c, oa_c, ba_c = _c(occ.detach()), _c(obj_alpha.detach()), _c(bg_alpha.detach())
        cls_c = _c(cls.detach()) if cls is not None else None
        B, T, Cc, Hd, Wd = inp_c.shape
        Tc, Tp = ctx_ts.shape[1], pred_ts.shape[0]
        Nl = Cc - 3
        if (Hd, Wd) != (spec.Hd, spec.Wd):
            raise RuntimeError(f"waldo_b200.decode: input is {Hd}x{Wd}, warper expects {spec.Hd}x{spec.Wd}")
        if spec.weight_cls and cls is None:
            raise RuntimeError("waldo_b200.decode: weight_cls needs cls (the reference fails here too, lvd.py:737)")
        g = _geom(spec, B, T, Tc, Tp, Nl, cls is not None)
        Lr = spec.num_obj + 1
        dev = inp_c.device
        # the kernels index every buffer from the geometry: a tensor of another shape would be read / written out of bounds
        No = spec.num_obj
        want = dict(tgt_grid_obj=(tgo_c, (B, T, No, spec.Ho, spec.Wo, 2)), src_grid_obj=(sgo_c, (B, T, No, spec.H, spec.W, 2)),
                    tgt_grid_bg=(tgb_c, (B, T, spec.H, spec.W, 2)), src_grid_bg=(sgb_c, (B, T, spec.H, spec.W, 2)),
                    occ=(occ_c, (B, T, Lr, Lr)), ctx_ts=(ctx_ts, (B, Tc, Tp)), xs_hd=(xs_hd, (Wd,)), ys_hd=(ys_hd, (Hd,)))
        if cls_c is not None:
            want["cls"] = (cls_c, (B, No, Nl))
        for name, (t, shape) in want.items():
            if tuple(t.shape) != shape:
                raise RuntimeError(f"waldo_b200.decode: {name} has shape {tuple(t.shape)}, expected {shape} "
                                   f"(B={B}, T={T}, num_obj={No}, low-res {spec.H}x{spec.W}, canvas {spec.Ho}x{spec.Wo})")
        if oa_c.numel() != B * No * spec.Ho * spec.Wo or ba_c.numel() != B * spec.H * spec.W:
            raise RuntimeError(f"waldo_b200.decode: obj_alpha {tuple(oa_c.shape)} / bg_alpha {tuple(ba_c.shape)} do not match "
                               f"(B={B}, num_obj={No}, canvas {spec.Ho}x{spec.Wo}, low-res {spec.H}x{spec.W})")
        for name, t in (("grid", tgo_c), ("occ", occ_c), ("obj_alpha", oa_c), ("bg_alpha", ba_c), ("cls", cls_c)):
            if t is not None and t.device != dev:
                raise RuntimeError(f"waldo_b200.decode: {name} is on {t.device}, input on {dev}")
        ts_c = _c(ctx_ts, torch.int64).to(dev)
        ps_c = _c(pred_ts, torch.int64).to(dev)
        _check_time_indices(ctx_ts, pred_ts, ts_c, ps_c, g.Tw, T)
        self_ctx = spec.include_self and Tp == T
        TcR, CR = Tc + (1 if self_ctx else 0), Cc + Lr + (1 if spec.use_disocc else 0)
        f32 = dict(device=dev, dtype=torch.float32)
        a_lo = torch.empty(B, g.Tw, Lr, spec.H, spec.W, **f32)
        nout = spec.num_obj * Nl + spec.num_obj
        prof_ctas = int(os.environ.get("WALDO_PROF_CTAS", 0)) or max(1, min(148, (g.Tw * spec.H * spec.W + 255) // 256))   # one per SM
        prof_part = torch.empty(B, prof_ctas, nout, **f32)
        prof_sum = torch.empty(B, nout, **f32)
        prof_p = torch.empty(B, spec.num_obj, Nl, **f32)
        f_lo = torch.empty(B, Tc, Tp, Lr, spec.H, spec.W, 2, **f32)
        s_lo = torch.empty(B, Tp, spec.num_obj, spec.H, spec.W, **f32)
        live_ctx = torch.empty(B, g.Tw, spec.H, spec.W, device=dev, dtype=torch.int32)
        live_pred = torch.empty(B, Tp, spec.H, spec.W, device=dev, dtype=torch.int32)
        alpha = torch.empty(B, g.Tw, Lr, Hd, Wd, **f32)   # fp32 in every storage variant: the flow is computed from it
        flow = torch.empty(B, Tc, Tp, 2, Hd, Wd, **f32)
        raw = torch.empty(B, TcR, Tp, CR, Hd, Wd, device=dev, dtype=st)
        out_full = torch.empty(B, Tp, Cc + 1, Hd, Wd, device=dev, dtype=st)
        norm = torch.empty(B, Tp, Hd, Wd, **f32)
        score = torch.empty(B, Tc, Tp, Hd, Wd, **f32)
        # low-res layout logits: kept only when a backward will follow (it saves re-reading the HD layout planes)
        lyt_lo = torch.empty(B, g.Tw, Nl, spec.H, spec.W, **f32) if (spec.use_filter and any(ctx.needs_input_grad)) else None
        tensors = (inp_c, tgo_c, sgo_c, tgb_c, sgb_c, occ_c, oa_c, ba_c, cls_c, ts_c, ps_c, xs_hd, ys_hd, a_lo, prof_part,
                   prof_sum, prof_p, f_lo, s_lo, live_ctx, live_pred, alpha, flow, raw, out_full, norm, score, lyt_lo)
        # (PREFILL only, off by default -- see the note at PREFILL.)
        # The fills are issued right after the forward kernels, on a side stream (`side.wait_stream(main)`: after the forward),
        # and the backward waits for them: they overlap with whatever the caller runs between this forward and its backward
        # (WIF's UNet, the losses).  Overlapping them with the forward kernels themselves was measured on the B200 and gains
        # nothing: fill CTAs and forward CTAs compete for the same SM slots (forward + 0.45 ms, exactly the fills' own time;
        # profiles/r2/r2_notes.md).  The buffers come from the MAIN stream's pool (allocation and release follow the main
        # stream like every other tensor of the step); only the fill kernels run on the side stream.
        ctx.prefill = None
        hook = None
        if PREFILL and inp_c.is_cuda and any(ctx.needs_input_grad[5:]):
            plan = _scatter_plan(ctx.needs_input_grad[5:], g, cls_c is not None)
            det = is_deterministic()

            def hook():
                main, side = torch.cuda.current_stream(dev), _side_stream(dev)
                bufs = _alloc_scatter(plan, (inp_c, alpha, f_lo, a_lo, tgo_c, sgo_c, tgb_c, sgb_c, oa_c, ba_c), det, fill=False)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    for t in ([bufs[1], bufs[2]] if det else bufs[0]):
                        if t is not None:
                            t.zero_()
                            t.record_stream(side)
                    ev = torch.cuda.Event()
                    ev.record(side)
                ctx.prefill = (plan, det, bufs, ev)
        a = _fwd_struct(g, prof_ctas, tensors)
        _staged(lib.waldo_decode_fwd, a, L.stream_of(inp_c), "decode_fwd", dev=dev)
        if hook is not None:
            hook()
        # saved through save_for_backward (NOT as ctx attributes): four of them are outputs of this very Function, and an
        # attribute would tie them into a reference cycle that only the cyclic GC frees (tens of GB per step)
        ctx.save_for_backward(*[t for t in tensors if t is not None])
        ctx.present = [t is not None for t in tensors]
        ctx.geom, ctx.prof_ctas = g, prof_ctas
        ctx.shapes = dict(obj_alpha=obj_alpha.shape, bg_alpha=bg_alpha.shape)
        return out_full[:, :, :Cc], out_full[:, :, Cc:], raw, flow, alpha

    @staticmethod
    def backward(ctx, d_output, d_raw_alpha, d_raw, d_flow, d_alpha):
        lib = L.load()
        it = iter(ctx.saved_tensors)
        tensors = tuple(next(it) if p else None for p in ctx.present)
        fwd = _fwd_struct(ctx.geom, ctx.prof_ctas, tensors)
        (inp_c, tgo_c, sgo_c, tgb_c, sgb_c, occ_c, oa_c, ba_c, cls_c) = tensors[:9]
        g = fwd.g
        dev = inp_c.device
        need = ctx.needs_input_grad[5:]   # inp, tgo, sgo, tgb, sgb, occ, obj_alpha, bg_alpha, cls
        plan = _scatter_plan(need, g, cls_c is not None)
        n_inp, n_tgo, n_sgo, n_tgb, n_sgb, n_occ, n_oa, n_ba, n_cls = (plan[k] for k in ("n_inp", "n_tgo", "n_sgo", "n_tgb", "n_sgb", "n_occ", "n_oa", "n_ba", "n_cls"))
        z = lambda ref, on: torch.zeros_like(ref) if on else None
        d_occ, d_cls = z(occ_c, n_occ), z(cls_c, n_cls)
        filt, geom, chain = plan["filt"], plan["geom"], plan["chain"]
        a_lo, prof_part, prof_sum, prof_p, f_lo, s_lo = tensors[13:19]
        alpha = tensors[21]
        Lr = g.No + 1
        f32 = dict(device=dev, dtype=torch.float32)
        # the scatter targets (geometry gradients flow through all four grids together: the ones not asked for are scratch)
        det = is_deterministic()
        on = lambda ref, cond: ref if cond else None
        pre, ctx.prefill = getattr(ctx, "prefill", None), None
        if pre is not None and pre[0] == plan and pre[1] == det:
            # zero-filled during the forward on the side stream: wait for those fills (long finished), use them once
            torch.cuda.current_stream(dev).wait_event(pre[3])
            (d_input, d_alpha_acc, d_f_lo, d_a_lo, d_tgo_s, d_sgo_s, d_tgb_s, d_sgb_s, d_oa, d_ba), det_arena, det_shadow = pre[2]
        else:
            (d_input, d_alpha_acc, d_f_lo, d_a_lo, d_tgo_s, d_sgo_s, d_tgb_s, d_sgb_s, d_oa, d_ba), det_arena, det_shadow = _alloc_scatter(
                plan, (inp_c, alpha, f_lo, a_lo, tgo_c, sgo_c, tgb_c, sgb_c, oa_c, ba_c), det)
        d_tgo, d_sgo, d_tgb, d_sgb = on(d_tgo_s, n_tgo), on(d_sgo_s, n_sgo), on(d_tgb_s, n_tgb), on(d_sgb_s, n_sgb)
        det_scale = torch.empty(4, **f32) if det else None
        d_prof_p = torch.zeros_like(prof_p) if (chain and filt) else None
        d_prof_sum = torch.zeros_like(prof_sum) if (chain and filt) else None
        tiles = ((g.Wd + 31) // 32) * ((g.Hd + 7) // 8)
        groups = max(g.B * g.Tp, g.B * g.Tw)
        # CTAs per frame of the two reducing HD kernels: whole waves of the 148 SMs at their 2 resident CTAs per SM
        red_ctas = int(os.environ.get("WALDO_RED_CTAS", 0)) or max(1, min(tiles, -(-148 * 8 // (g.B * g.Tp))))
        occ_part = torch.empty(groups, red_ctas, Lr * Lr, **f32) if n_occ else None
        prof_p_part = torch.empty(g.B * g.Tw, red_ctas, g.No * g.Nl, **f32) if d_prof_p is not None else None
        cls_part = torch.empty(g.B, fwd.prof_ctas, g.No * g.Nl, **f32) if (n_cls and (g.flags & L.F_WEIGHT_CLS)) else None
        up_tab = torch.empty(g.W, 9, **f32)
        glue = torch.empty(g.B, g.Tc, g.Tp, 3, g.Hd, g.Wd, **f32) if chain else None
        d_cls_s = d_cls
        if d_cls_s is None and cls_c is not None and chain and filt and not (g.flags & L.F_WEIGHT_CLS):
            d_cls_s = None   # P = cls path: nothing to propagate unless cls needs grad
        grads_in = [_c(t) if t is not None else None for t in (d_output, d_raw_alpha, d_raw, d_flow, d_alpha)]
        b = L.DecodeBwd(fwd, L.ptr(grads_in[0]), L.ptr(grads_in[1]), L.ptr(grads_in[2]), L.ptr(grads_in[3]), L.ptr(grads_in[4]),
                        L.ptr(d_input), L.ptr(d_tgo_s), L.ptr(d_sgo_s), L.ptr(d_tgb_s), L.ptr(d_sgb_s), L.ptr(d_occ),
                        L.ptr(d_oa), L.ptr(d_ba), L.ptr(d_cls_s), L.ptr(d_alpha_acc), L.ptr(d_f_lo), L.ptr(d_a_lo),
                        L.ptr(d_prof_p), L.ptr(d_prof_sum), red_ctas, L.ptr(occ_part), L.ptr(prof_p_part), L.ptr(cls_part), L.ptr(up_tab), L.ptr(glue), 0,
                        L.ptr(det_arena), L.ptr(det_shadow, torch.int64), det_arena.numel() if det_arena is not None else 0,
                        L.ptr(det_scale))
        _staged(lib.waldo_decode_bwd, b, L.stream_of(inp_c), "decode_bwd", dev=dev)
        if det:
            global LAST_DET_SCALE
            LAST_DET_SCALE = det_scale
        if det:
            # The two background-grid gradients are consumed on the side stream of Warper.forward's background chain, and autograd
            # marks a tensor that crosses streams with record_stream.  As views of the deterministic arena that mark would sit
            # on the whole 3.2 GB block: its release would wait until the host has seen the side stream pass, and with the host
            # many steps ahead of the GPU every step took a fresh arena (reserved memory 20 -> 72 GiB over 25 steps, the
            # allocations inside the timed steps; profiles/r2/r2_notes.md).  Two 10 MB copies keep the arena out of it.
            d_tgb = d_tgb.clone() if d_tgb is not None else None
            d_sgb = d_sgb.clone() if d_sgb is not None else None
        if d_oa is not None:
            d_oa = d_oa.view(ctx.shapes["obj_alpha"])
        if d_ba is not None:
            d_ba = d_ba.view(ctx.shapes["bg_alpha"])
        return (None, None, None, None, None, d_input, d_tgo, d_sgo, d_tgb, d_sgb, d_occ, d_oa, d_ba, d_cls)


def decode(spec: DecodeSpec, ctx_ts, pred_ts, xs_hd, ys_hd, inp, grid, occ, obj_alpha, bg_alpha, cls):
    tgo, sgo, tgb, sgb = grid
    spec.occ_pairs_only = type(occ.grad_fn).__name__ == "_ComputeOccBackward"
    return _Decode.apply(spec, ctx_ts, pred_ts, xs_hd, ys_hd, inp, tgo, sgo, tgb, sgb, occ, obj_alpha, bg_alpha, cls)


# ===================================================================================== a-9 WIF fuse tail
class _WifFuse(torch.autograd.Function):
    """WIF.forward tail, models/nets/wif.py:50-54."""

    @staticmethod
    def forward(ctx, raw_output, unet_out, ab):
        lib = L.load()
        r, u = _c(raw_output.detach()), _c(unet_out.detach())
        B, Tc, Tp, Cr, H, W = r.shape
        if u.shape[:3] != (B, Tp, Tc) or u.shape[3] != 4 + (1 if ab else 0):
            raise RuntimeError(f"waldo_b200.wif_fuse: unet_out shape {tuple(u.shape)} does not match raw_output {tuple(r.shape)}")
        if Cr < 5:   # wif.py:53 reads INPUT channel 4 as the gate; the backward writes d raw_output channels 0..4
            raise RuntimeError(f"waldo_b200.wif_fuse: raw_output needs at least 5 channels, got {Cr}")
        frame = torch.empty(B, Tp, 3, H, W, device=r.device, dtype=torch.float32)
        a = L.WifFuseFwd(B, Tc, Tp, Cr, H * W, 1 if ab else 0, L.ptr(r, name="raw_output"), L.ptr(u, name="unet_out"), L.ptr(frame))
        L.call(lib.waldo_wif_fuse_fwd, a, r, "wif_fuse_fwd")
        ctx.keep = (a, r, u)
        return frame

    @staticmethod
    def backward(ctx, d_frame):
        lib = L.load()
        a, r, u = ctx.keep
        d_frame = _c(d_frame)
        d_raw = torch.zeros_like(r) if ctx.needs_input_grad[0] else None
        d_u = torch.empty_like(u) if ctx.needs_input_grad[1] else None
        b = L.WifFuseBwd(a, L.ptr(d_frame), L.ptr(d_raw), L.ptr(d_u))
        L.call(lib.waldo_wif_fuse_bwd, b, r, "wif_fuse_bwd")
        return d_raw, d_u, None


def wif_fuse(raw_output, unet_out, ab=True):
    """raw_output (B,Tc,Tp,Cin,H,W) as produced by decode_output; unet_out (B,Tp,Tc,4|5,H,W)."""
    return _WifFuse.apply(raw_output, unet_out, ab)


# ===================================================================================== f-3 input packing
def pack_input(rgb, label, num_lyt, on=5.0, off=-5.0, out=None, dtype=torch.float32):
    """Build `input` (B,T,3+Nl,Hd,Wd) on the device from rgb (B,T,3,Hd,Wd; uint8 raw pixels or fp32 already in
    [-1,1]) and label (B,T,Hd,Wd) uint8 class ids -- data/base_dataset.py:173-183,:355-372 + synthesizer.py:444.
    dtype: torch.float32 (the reference's), or torch.bfloat16 for the bf16-storage inference variant of decode.
    Data preparation: no gradient."""
    lib = L.load()
    if rgb.dim() != 5 or rgb.size(2) != 3 or label.shape != rgb.shape[:2] + rgb.shape[3:]:
        raise RuntimeError(f"waldo_b200.pack_input: rgb {tuple(rgb.shape)} / label {tuple(label.shape)} shapes do not match")
    if label.dtype != torch.uint8:
        raise RuntimeError("waldo_b200.pack_input: label must be uint8 class ids")
    B, T, _, Hd, Wd = rgb.shape
    rgb_c, lab_c = rgb.detach().contiguous(), label.contiguous()
    if dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("waldo_b200.pack_input: dtype must be torch.float32 or torch.bfloat16")
    if out is None:
        out = torch.empty(B, T, 3 + num_lyt, Hd, Wd, device=rgb.device, dtype=dtype)
    elif out.shape != (B, T, 3 + num_lyt, Hd, Wd) or out.dtype != dtype:
        raise RuntimeError("waldo_b200.pack_input: out has the wrong shape / dtype")
    u8 = rgb_c.dtype == torch.uint8
    a = L.PackInput(B * T, num_lyt, Hd * Wd, float(on), float(off),
                    L.ptr(rgb_c, torch.uint8, "rgb") if u8 else None, None if u8 else L.ptr(rgb_c, name="rgb"),
                    L.ptr(lab_c, torch.uint8, "label"), L.ptr(out, dtype, name="out"),
                    L.ST_BF16 if dtype == torch.bfloat16 else L.ST_F32)
    L.call(lib.waldo_pack_input, a, out, "pack_input")
    return out


# ===================================================================================== f-4 output side
def frames_to_u8(vid, span=(-1.0, 1.0), out=None):
    """tools/utils.py:246-249 + :258-264 on the device: vid (..., 3, H, W) fp32 in `span` -> (..., H, W, 3) uint8, the
    tensor `torchvision.io.write_video` takes (dump_video), so that save_vid's device->host copy (synthesizer.py:184-193)
    moves one byte per sample.  Data preparation: no gradient."""
    lib = L.load()
    if vid.dim() < 3 or vid.size(-3) != 3:
        raise RuntimeError(f"waldo_b200.frames_to_u8: expected (..., 3, H, W), got {tuple(vid.shape)}")
    v = _c(vid.detach())
    *lead, _, H, W = v.shape
    n = 1
    for x in lead:
        n *= x
    if out is None:
        out = torch.empty(*lead, H, W, 3, device=v.device, dtype=torch.uint8)
    elif tuple(out.shape) != (*lead, H, W, 3) or out.dtype != torch.uint8:
        raise RuntimeError("waldo_b200.frames_to_u8: out has the wrong shape / dtype")
    a = L.FramesU8(n, H * W, float(span[0]), float(span[1]), L.ptr(v, name="vid"), L.ptr(out, torch.uint8, "out"))
    L.call(lib.waldo_frames_to_u8, a, v, "frames_to_u8")
    return out


# ===================================================================================== f-2 loss epilogues
class _Blur(torch.autograd.Function):
    """models/synthesizer.py:1114-1118 `blur`: torchvision GaussianBlur(kernel_size, sigma) over the last two dims."""

    @staticmethod
    def forward(ctx, vid, sigma, kernel_size):
        lib = L.load()
        x = _c(vid.detach())
        H, W = x.shape[-2:]
        n = x.numel() // (H * W)
        out = torch.empty_like(x)
        a = L.Blur(n, H, W, int(kernel_size), float(sigma), L.ptr(x, name="vid"), L.ptr(out))
        L.call(lib.waldo_blur_fwd, a, x, "blur_fwd")
        ctx.dims = (n, H, W, int(kernel_size), float(sigma))
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = L.load()
        n, H, W, K, sigma = ctx.dims
        g = _c(d_out)
        d_in = torch.empty_like(g)
        a = L.Blur(n, H, W, K, sigma, L.ptr(g), L.ptr(d_in))
        L.call(lib.waldo_blur_bwd, a, g, "blur_bwd")
        return d_in, None, None


def blur(vid, sigma=3.0, kernel_size=23):
    if vid.dim() < 3:
        raise RuntimeError(f"waldo_b200.blur: expected (..., C, H, W), got {tuple(vid.shape)}")
    return _Blur.apply(vid, sigma, kernel_size)


class _LayerEntropy(torch.autograd.Function):
    """models/synthesizer.py:886-889 (entropy of the normalised layer opacities / 0.37) and :933 (fg_mask), one pass."""

    @staticmethod
    def forward(ctx, alpha):
        lib = L.load()
        a_c = _c(alpha.detach())
        *lead, Lr, H, W = a_c.shape
        n = 1
        for v in lead:
            n *= v
        ent = torch.empty(*lead, 1, H, W, device=a_c.device, dtype=torch.float32)
        fg = torch.empty(*lead, 1, H, W, device=a_c.device, dtype=torch.float32)
        a = L.LayerEntropy(n, Lr, H * W, L.ptr(a_c, name="alpha"), L.ptr(ent), L.ptr(fg))
        L.call(lib.waldo_layer_entropy_fwd, a, a_c, "layer_entropy_fwd")
        ctx.save_for_backward(a_c)
        ctx.dims = (n, Lr, H * W)
        ctx.set_materialize_grads(False)
        return ent, fg

    @staticmethod
    def backward(ctx, d_ent, d_fg):
        lib = L.load()
        (a_c,) = ctx.saved_tensors
        n, Lr, HW = ctx.dims
        d_ent = _c(d_ent) if d_ent is not None else None
        d_fg = _c(d_fg) if d_fg is not None else None
        d_alpha = torch.empty_like(a_c)
        b = L.LayerEntropyBwd(L.LayerEntropy(n, Lr, HW, L.ptr(a_c), None, None), L.ptr(d_ent), L.ptr(d_fg), L.ptr(d_alpha))
        L.call(lib.waldo_layer_entropy_bwd, b, a_c, "layer_entropy_bwd")
        return d_alpha


def layer_entropy(alpha):
    """alpha (..., L, H, W) in [-1, 1] -> (entropy (..., 1, H, W), fg_mask (..., 1, H, W))."""
    if alpha.dim() < 3:
        raise RuntimeError(f"waldo_b200.layer_entropy: expected (..., L, H, W), got {tuple(alpha.shape)}")
    return _LayerEntropy.apply(alpha)


class _PoseDis(torch.autograd.Function):
    """models/synthesizer.py:965-979: the per-pixel minima over the objects behind `cell_dis` and `center_dis`."""

    @staticmethod
    def forward(ctx, mov, fg, pose, grid, eps, ho, wo):
        lib = L.load()
        mov_c, fg_c, pose_c, grid_c = _c(mov.detach()), _c(fg.detach()), _c(pose.detach()), _c(grid.detach())
        *lead, one, H, W = mov_c.shape
        n = 1
        for v in lead:
            n *= v
        No = pose_c.shape[-3]
        cell = torch.empty(*lead, H, W, device=mov_c.device, dtype=torch.float32)
        center = torch.empty_like(cell)
        carg = torch.empty(*lead, H, W, device=mov_c.device, dtype=torch.uint8)
        marg = torch.empty_like(carg)
        a = L.PoseDis(n, No, ho, wo, H * W, float(eps), L.ptr(mov_c, name="mov_obj_mask"), L.ptr(fg_c, name="fg_mask"), L.ptr(pose_c, name="obj_pose"),
                      L.ptr(grid_c, name="grid"), L.ptr(cell), L.ptr(center), L.ptr(carg, torch.uint8), L.ptr(marg, torch.uint8))
        L.call(lib.waldo_pose_dis_fwd, a, mov_c, "pose_dis_fwd")
        ctx.save_for_backward(mov_c, fg_c, pose_c, grid_c, carg, marg)
        ctx.dims = (n, No, ho, wo, H * W, float(eps))
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(carg, marg)
        return cell, center, carg, marg

    @staticmethod
    def backward(ctx, d_cell, d_center, _a, _b):
        lib = L.load()
        mov_c, fg_c, pose_c, grid_c, carg, marg = ctx.saved_tensors
        n, No, ho, wo, HW, eps = ctx.dims
        d_cell = _c(d_cell) if d_cell is not None else None
        d_center = _c(d_center) if d_center is not None else None
        d_fg = torch.empty_like(fg_c) if ctx.needs_input_grad[1] else None
        d_mov = torch.empty_like(mov_c) if ctx.needs_input_grad[0] else None
        d_pose = torch.empty_like(pose_c)
        ctas = max(1, min(1024, (HW + 255) // 256))
        part = torch.empty(max(n, 1), ctas, No, 6, device=mov_c.device, dtype=torch.float32)
        f = L.PoseDis(n, No, ho, wo, HW, eps, L.ptr(mov_c), L.ptr(fg_c), L.ptr(pose_c), L.ptr(grid_c), None, None, L.ptr(carg, torch.uint8), L.ptr(marg, torch.uint8))
        b = L.PoseDisBwd(f, L.ptr(d_cell), L.ptr(d_center), L.ptr(d_fg), L.ptr(d_mov), ctas, L.ptr(part), L.ptr(d_pose))
        L.call(lib.waldo_pose_dis_bwd, b, mov_c, "pose_dis_bwd")
        return d_mov, d_fg, (d_pose if ctx.needs_input_grad[2] else None), None, None, None, None


def pose_distances(mov_obj_mask, fg_mask, obj_pose, grid, obj_shape, eps):
    """(cell_min, center_min, cell_arg, center_arg), each (..., H, W): the maps whose means are the reference's `cell_dis` and
    `center_dis` (synthesizer.py:965-979) and the objects that attain the minima.
    mov_obj_mask, fg_mask (..., 1, H, W); obj_pose (..., No, ho*wo, 2); grid (1, H, W, 2) = warper.src_grid."""
    ho, wo = int(obj_shape[0]), int(obj_shape[1])
    if mov_obj_mask.dim() < 3 or mov_obj_mask.shape[-3] != 1 or mov_obj_mask.shape != fg_mask.shape:
        raise RuntimeError(f"waldo_b200.pose_distances: masks must both be (..., 1, H, W), got {tuple(mov_obj_mask.shape)} / {tuple(fg_mask.shape)}")
    H, W = mov_obj_mask.shape[-2:]
    lead = tuple(mov_obj_mask.shape[:-3])
    if obj_pose.dim() != len(lead) + 3 or tuple(obj_pose.shape[:len(lead)]) != lead or obj_pose.shape[-2] != ho * wo or obj_pose.shape[-1] != 2:
        raise RuntimeError(f"waldo_b200.pose_distances: obj_pose must be {lead + ('No', ho * wo, 2)}, got {tuple(obj_pose.shape)}")
    if grid.numel() != H * W * 2 or grid.shape[-1] != 2:
        raise RuntimeError(f"waldo_b200.pose_distances: grid must be (1, {H}, {W}, 2), got {tuple(grid.shape)}")
    return _PoseDis.apply(mov_obj_mask, fg_mask, obj_pose, grid, eps, ho, wo)


class _ObjFlow(torch.autograd.Function):
    """models/synthesizer.py:864-868: dev_map (..., H, W) = sum_o a_o (|fx - mx_o| + |fy - my_o|) of the object layers."""

    @staticmethod
    def forward(ctx, alpha, flow):
        lib = L.load()
        a_c, f_c = _c(alpha.detach()), _c(flow.detach())
        *lead, Lr, H, W = a_c.shape
        n = 1
        for v in lead:
            n *= v
        HW = H * W
        ctas = max(1, min(64, (HW + 1023) // 1024))
        part = torch.empty(max(n, 1), ctas, Lr - 1, 3, device=a_c.device, dtype=torch.float32)
        mom = torch.empty(max(n, 1), Lr - 1, 3, device=a_c.device, dtype=torch.float32)
        dev = torch.empty(*lead, H, W, device=a_c.device, dtype=torch.float32)
        a = L.ObjFlow(n, Lr, HW, L.ptr(a_c, name="alpha"), L.ptr(f_c, name="flow"), ctas, L.ptr(part), L.ptr(mom), L.ptr(dev))
        L.call(lib.waldo_obj_flow_fwd, a, a_c, "obj_flow_fwd")
        ctx.save_for_backward(a_c, f_c, mom)
        ctx.dims = (n, Lr, HW, ctas)
        return dev

    @staticmethod
    def backward(ctx, d_map):
        lib = L.load()
        a_c, f_c, mom = ctx.saved_tensors
        n, Lr, HW, ctas = ctx.dims
        d_map = _c(d_map)
        part = torch.empty(max(n, 1), ctas, Lr - 1, 3, device=a_c.device, dtype=torch.float32)
        tsum = torch.empty(max(n, 1), Lr - 1, 2, device=a_c.device, dtype=torch.float32)
        d_alpha = torch.empty_like(a_c)
        f = L.ObjFlow(n, Lr, HW, L.ptr(a_c), L.ptr(f_c), ctas, L.ptr(part), L.ptr(mom), None)
        b = L.ObjFlowBwd(f, L.ptr(d_map), L.ptr(tsum), L.ptr(d_alpha))
        L.call(lib.waldo_obj_flow_bwd, b, a_c, "obj_flow_bwd")
        return d_alpha, None


def obj_flow_map(alpha, flow):
    """alpha (..., L, H, W) in [-1, 1], flow (..., 2, H, W) -> dev_map (..., H, W); differentiable in alpha."""
    if alpha.dim() < 3 or flow.dim() != alpha.dim() or flow.shape[-3] != 2 or flow.shape[-2:] != alpha.shape[-2:] or flow.shape[:-3] != alpha.shape[:-3]:
        raise RuntimeError(f"waldo_b200.obj_flow: expected alpha (..., L, H, W) and flow (..., 2, H, W), got {tuple(alpha.shape)} / {tuple(flow.shape)}")
    if alpha.shape[-3] < 2:
        raise RuntimeError("waldo_b200.obj_flow: needs at least one object layer besides the background")
    if flow.requires_grad:
        raise NotImplementedError("waldo_b200.obj_flow: no gradient with respect to the flow (real_flow is data in the reference)")
    return _ObjFlow.apply(alpha, flow)


# ===================================================================================== f-1 first UNet layer
def _conv3x3_launch(x, weight, n, Cin, H, W, Tc, Tp):
    lib = L.load()
    out = torch.empty(n, weight.shape[0], H, W, device=x.device, dtype=torch.float32)
    a = L.Conv3x3(n, Cin, weight.shape[0], H, W, Tc, Tp, L.ptr(x, name="x"), L.ptr(weight, name="weight"), L.ptr(out))
    L.call(lib.waldo_conv3x3_fwd, a, x, "conv3x3_fwd")
    return out


class _Conv3x3(torch.autograd.Function):
    """conv3x3(stride 1, padding 1, no bias), models/modules/conv.py:9-11, TF32 products / fp32 accumulation.
    forward + backward-data on waldo_conv3x3_fwd (the backward-data of a stride-1 3x3 convolution is the same convolution with
    the weights flipped and transposed, and the image permute is undone by swapping Tc and Tp); the weight gradient -- a
    reduction over all pixels of all images -- on waldo_conv3x3_wgrad for the shapes it covers (Cout <= 16, Cin <= 40,
    W % 4 == 0; per-CTA partials added in CTA order: deterministic), else torch (cuDNN)."""

    @staticmethod
    def forward(ctx, x, weight, wif_permute):
        xc, wc = _c(x.detach()), _c(weight.detach())
        if wc.dim() != 4 or wc.shape[2:] != (3, 3):
            raise RuntimeError(f"waldo_b200.conv3x3: weight must be (Cout, Cin, 3, 3), got {tuple(wc.shape)}")
        if wif_permute:
            if xc.dim() != 6:
                raise RuntimeError(f"waldo_b200.conv3x3: wif_permute expects raw_output (B, Tc, Tp, C, H, W), got {tuple(xc.shape)}")
            B, Tc, Tp, Cin, H, W = xc.shape
            n = B * Tc * Tp
        else:
            if xc.dim() != 4:
                raise RuntimeError(f"waldo_b200.conv3x3: expected (n, Cin, H, W), got {tuple(xc.shape)}")
            n, Cin, H, W = xc.shape
            Tc = Tp = 0
        if wc.shape[1] != Cin:
            raise RuntimeError(f"waldo_b200.conv3x3: weight has {wc.shape[1]} input channels, x has {Cin}")
        ctx.save_for_backward(xc, wc)
        ctx.dims = (n, Cin, H, W, Tc, Tp, bool(wif_permute))
        return _conv3x3_launch(xc, wc, n, Cin, H, W, Tc, Tp)

    @staticmethod
    def backward(ctx, d_out):
        xc, wc = ctx.saved_tensors
        n, Cin, H, W, Tc, Tp, perm = ctx.dims
        g = _c(d_out)
        d_x = d_w = None
        if ctx.needs_input_grad[0]:
            wt = wc.flip(2, 3).transpose(0, 1).contiguous()                       # (Cin, Cout, 3, 3)
            d_x = _conv3x3_launch(g, wt, n, wc.shape[0], H, W, Tp, Tc)            # Tc <-> Tp: the inverse image permute
            d_x = d_x.view(xc.shape)
        if ctx.needs_input_grad[1]:
            Cout = wc.shape[0]
            if Cout <= 16 and Cin <= 40 and W % 4 == 0 and xc.data_ptr() % 16 == 0 and g.data_ptr() % 16 == 0:
                lib = L.load()
                ctas = 2 * 148
                part = torch.empty(ctas, Cout, Cin, 3, 3, device=xc.device, dtype=torch.float32)
                d_w = torch.empty_like(wc)
                a = L.Conv3x3Wgrad(L.Conv3x3(n, Cin, Cout, H, W, Tc, Tp, L.ptr(xc), None, None), L.ptr(g), ctas, L.ptr(part), L.ptr(d_w))
                L.call(lib.waldo_conv3x3_wgrad, a, xc, "conv3x3_wgrad")
            else:   # shapes the kernel does not cover: cuDNN
                x4 = xc.permute(0, 2, 1, 3, 4, 5).reshape(n, Cin, H, W) if perm else xc
                d_w = torch.nn.grad.conv2d_weight(x4, wc.shape, g, padding=1)
        return d_x, d_w, None


def conv3x3(x, weight, wif_permute=False):
    """UNet.to_emb (models/modules/conv.py:9-11, :54): conv3x3(stride 1, padding 1, no bias) with TF32 tensor-core products and
    fp32 accumulation.  x (n, Cin, H, W) -> (n, Cout, H, W); with wif_permute, x is raw_output (B, Tc, Tp, Cin, H, W) as
    decode_output returns it and the result is (B*Tp*Tc, Cout, H, W) in the image order of WIF.forward (wif.py:33-38), without the
    permuted copy.  Differentiable: d x (= d raw_output, the upstream gradient of the warp backward) by the same kernel, d weight
    by torch."""
    return _Conv3x3.apply(x, weight, wif_permute)


# ===================================================================================== a-5 / a-11 field warp, scale
def _no_grad_path(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError("waldo_b200: the stand-alone field warp / scale helpers are forward-only (inference); inside "
                                  "decode_output these steps are fused into kernels that carry the gradients")


def warp_field(field, grid, delta=0.0):
    """grid_sample(field + delta, grid) - delta, bilinear / zeros / align_corners=False (lvd.py:548,559).
    field (n, c, h, w), grid (n, H, W, 2) -> (n, c, H, W)."""
    _no_grad_path(field, grid)
    lib = L.load()
    f, g = _c(field.detach()), _c(grid.detach())
    n, c, h, w = f.shape
    if g.shape[0] != n or g.shape[-1] != 2:
        raise RuntimeError(f"waldo_b200.warp_field: field {tuple(f.shape)} and grid {tuple(g.shape)} do not match")
    H, W = g.shape[1], g.shape[2]
    out = torch.empty(n, c, H, W, device=f.device, dtype=torch.float32)
    a = L.WarpField(n, c, h, w, H, W, float(delta), L.ptr(f, name="field"), L.ptr(g, name="grid"), L.ptr(out))
    L.call(lib.waldo_warp_field_fwd, a, f, "warp_field_fwd")
    return out


def resize_bilinear(x, scale_factor):
    """lvd.py:175-179 `scale`: F.interpolate(bilinear, scale_factor) over the last two dims (up-sampling), any leading dims."""
    _no_grad_path(x)
    if scale_factor == 1:
        return x
    lib = L.load()
    xc = _c(x.detach())
    h, w = xc.shape[-2:]
    H, W = int(h * scale_factor), int(w * scale_factor)
    n = xc.numel() // (h * w)
    out = torch.empty(*xc.shape[:-2], H, W, device=xc.device, dtype=torch.float32)
    a = L.Resize(n, h, w, H, W, L.ptr(xc, name="x"), L.ptr(out))
    L.call(lib.waldo_resize_bilinear_fwd, a, xc, "resize_bilinear_fwd")
    return out
