"""Parity checks shared by the CPU (kernel-logic emulation) and GPU (the real sm_100a library) test files.

Each check feeds identical inputs to the product path (waldo_b200, through the C ABI) and to the oracle
(oracle/waldo_oracle.py, fp32 and its fp64 twin) / the committed reference outputs (tests/golden), using the tiers of
SURVEY.md §8(d):
  T0 stage-local, T1 chained from an identical `grid`, T2 end-to-end from control points (statistical).
Tolerances (stated once, used everywhere):
  * index / threshold maps: bit-exact;
  * fp32 forward: max-abs <= 1e-5 wherever the reference's own fp32-vs-fp64 floor is below that, else fp64-arbitrated:
        max|k - f64| <= 2 * max|ref32 - f64| + 1e-6;
  * fp32 gradients: ||g - g_ref||_inf / ||g_ref||_inf <= 1e-4, or fp64-arbitrated the same way.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import waldo_oracle as wo  # noqa: E402
import waldo_b200 as wb  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["city_x4", "kitti_x2", "train_lo", "cls_plain", "many_obj", "c23_x4", "c22_x2", "city_real", "kitti_real"]
# city_real / kitti_real: the REAL structure of the benchmarked configurations at reduced resolution -- 16 objects with 4x4
# control points, the K = 131 / 211 background TPS systems, 4 contexts, 20 / 19 classes, scale_hd 4 / 2 (oracle/make_golden.py)
PROJ_SEED, PROJ_ORDER = 1234, ("output", "flow", "alpha", "raw_alpha", "raw_output")
OUT_NAMES = ["output", "flow", "alpha_unflt", "alpha", "raw_alpha", "raw_output", "alpha_ctx"]

FWD_TOL = 1e-5
GRAD_TOL = 1e-4
# Rule (4), the bf16-STORAGE variant (input / raw_output / output held as bf16 in HBM; alpha, flow and all arithmetic fp32; forward
# only; include/waldo_b200.h WALDO_ST_BF16), stated against the fp32 path on the same inputs:
#   * flow, alpha (fp32 in both variants; they only see the bf16 rounding of the INPUT's layout logits): with the dataset's one-hot
#     +-5 logits (data/base_dataset.py:173-183: exactly representable in bf16) they are BIT-IDENTICAL to the fp32 path -- sampling
#     positions and every index-valued decision do not depend on the storage type; with arbitrary (smooth) logits: flow max-abs
#     <= 1e-3 (normalised units), alpha max-abs <= 1.6e-2;
#   * alpha_ctx, raw_alpha (values in [-1, 1], one bf16 rounding = 2^-9 relative): max-abs <= 1.6e-2 (= 2^-6);
#   * raw_output image / layout channels (values in [-5, 5]): one rounding at |v| >= 4 is 1.56e-2 (the whole deviation with one-hot
#     logits); with smooth logits the flow deviation above shifts the taps: mean-abs <= 5e-3, 99.9th percentile <= 0.1, max-abs <= 0.5;
#   * output (the score-weighted mean of the contexts, lvd.py:850-851: divides by the summed score, which is ~1e-6 where no
#     context sees the pixel -- the reference's own fp32-vs-fp64 error is O(1) there): where the summed
#     context score sum_tc sum_k A_k is >= 0.1: mean-abs <= 8e-3, max-abs <= 0.25; elsewhere finite.
TOL_BF16 = dict(flow_max=1e-3, alpha_max=1.6e-2, raw_mean=5e-3, raw_p999=0.1, raw_max=0.5, out_norm=0.1, out_mean=8e-3, out_max=0.25)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    B, T, Tc, smooth = meta.pop("B"), meta.pop("T"), meta.pop("Tc"), meta.pop("smooth")
    cfg = wo.PathConfig(**meta)
    d = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    if "proj_output" not in d:   # lean fixture: the projections / fake UNet output are re-drawn from the seed, in the
        gen = torch.Generator().manual_seed(PROJ_SEED)   # order oracle/make_golden.draw_projections drew them
        for n in PROJ_ORDER:
            d["proj_" + n] = torch.randn(tuple(d[n].shape), generator=gen)
        Bq, Tcq, Tpq, Cq, Hq, Wq = d["raw_output"].shape
        d["wif_unet_out"] = torch.randn(Bq * Tpq * Tcq, 5, Hq, Wq, generator=gen).view(Bq, Tpq, Tcq, 5, Hq, Wq)
        C, L = d["in_input"].shape[2], cfg.num_obj + 1
        d["alpha_ctx"] = d["raw_output"][:, :Tc, :, C:C + L]
    return cfg, (B, T, Tc), d


def make_opt(cfg):
    """The option namespace the reference's Warper reads (lvd.py:470-499), filled from a PathConfig."""
    from waldo_b200 import workloads
    return workloads.make_opt(cfg)


def arbitrated(k, ref32, f64, tol, what):
    """max|k-ref32| <= tol, or k is as close to the fp64 twin as the fp32 reference itself is."""
    k, ref32, f64 = k.detach().cpu().double(), ref32.detach().cpu().double(), f64.detach().cpu().double()
    direct = float((k - ref32).abs().max())
    if direct <= tol:
        return
    ek, er = float((k - f64).abs().max()), float((ref32 - f64).abs().max())
    assert ek <= 2 * er + 1e-6, f"{what}: |k-ref32|={direct:.3e} > {tol:.0e} and |k-f64|={ek:.3e} > 2*|ref32-f64|={er:.3e}"


def grad_close(g, g32, g64, what, tol=GRAD_TOL):
    g, g32, g64 = g.detach().cpu().double(), g32.detach().cpu().double(), g64.detach().cpu().double()
    scale = float(g64.abs().max())
    if scale == 0:
        assert float(g.abs().max()) == 0, f"{what}: reference gradient is zero, kernel's is not"
        return
    direct = float((g - g32).abs().max()) / scale
    if direct <= tol:
        return
    ek, er = float((g - g64).abs().max()) / scale, float((g32 - g64).abs().max()) / scale
    assert ek <= 2 * er + 1e-6, f"{what}: rel err vs ref32 {direct:.3e} > {tol:.0e}; vs f64 {ek:.3e} > 2*ref floor {er:.3e}"


# ------------------------------------------------------------------------------------------------ stage A
def check_tps(dev, case):
    """a-1 TPSWarp.forward + backward (warp.py:49-55): fp64-arbitrated (bg system cond ~1e4)."""
    cfg, (B, T, Tc), z = load_case(case)
    warper = wb.Warper(make_opt(cfg)).to(dev)
    st32, st64 = wo.make_state(cfg), wo.make_state(cfg, torch.float64)
    for which, tps, basis32, basis64, pts in (
            ("obj", warper.tps_obj, st32.tps_obj, st64.tps_obj, z["in_obj_pose"].reshape(-1, z["in_obj_pose"].shape[-2], 2)),
            ("bg", warper.tps_bg, st32.tps_bg, st64.tps_bg, z["in_bg_pose"].reshape(-1, z["in_bg_pose"].shape[-2], 2))):
        p = pts.clone().to(dev).requires_grad_(True)
        out = tps(p)
        p32 = pts.clone().requires_grad_(True)
        p64 = pts.double().clone().requires_grad_(True)
        o32, o64 = wo.tps_eval(basis32, p32), wo.tps_eval(basis64, p64)
        arbitrated(out, o32, o64, FWD_TOL, f"tps_{which}")
        # never worse than the reference's own fp32 error (+1e-6): SURVEY.md §8d
        assert float((out.detach().cpu().double() - o64).abs().max()) <= float((o32.double() - o64).abs().max()) + 1e-6
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
        (out * w.to(dev)).sum().backward()
        (o32 * w).sum().backward()
        (o64 * w.double()).sum().backward()
        grad_close(p.grad, p32.grad, p64.grad, f"tps_{which} d pts")
    # buffers are the reference's (state-dict compatible)
    assert torch.equal(warper.tps_bg.inverse_kernel.cpu(), st32.tps_bg.inverse_kernel)
    assert torch.equal(warper.tps_obj.tgt_grid_repr.cpu(), st32.tps_obj.tgt_grid_repr)


def check_inverse_warp(dev, case):
    """a-2 InverseWarp.forward + backward (warp.py:71-174) fed the REFERENCE's tgt_grid: index maps bit-exact."""
    cfg, (B, T, Tc), z = load_case(case)
    warper = wb.Warper(make_opt(cfg)).to(dev)
    H, W = cfg.lo_shape
    for which, mod, key, erode in (("obj", warper.invert_obj, "tgt_grid_obj", True), ("bg", warper.invert_bg, "tgt_grid_bg", False)):
        fwd = z[key].reshape(-1, *z[key].shape[-3:])
        x = fwd.clone().to(dev).requires_grad_(True)
        box = []
        out = mod(x, erode=erode, trace_box=box)
        tr = box[0]
        x32 = fwd.clone().requires_grad_(True)
        ref = wo.inverse_warp(x32, (H, W), erode=erode, trace=True)
        n, P = fwd.shape[0], H * W
        # rule (1): landing cells, hit mask, winners, known mask -- bit exact
        assert torch.equal(tr.field.cpu().long(), ref.field), f"{which}: field differs"
        win = tr.winner.cpu().long()
        win = torch.where(win == 2 ** 31 - 1, torch.full_like(win, P), win)
        assert torch.equal(win, ref.winner), f"{which}: winners differ"
        m = 6
        known = ((tr.level != 255) & (tr.eroded == 0))[:, m:-m, m:-m].cpu()
        assert torch.equal(known, ref.known), f"{which}: known mask differs"
        assert torch.equal((tr.level[:, m:-m, m:-m] == 0).cpu(), ref.hit), f"{which}: hit mask differs"
        # rule (2): values
        golden = z["src_grid_" + which].reshape(out.shape)
        assert float((out.detach().cpu() - ref.grid).abs().max()) <= FWD_TOL
        assert float((out.detach().cpu() - golden).abs().max()) <= FWD_TOL, f"{which}: differs from the reference's output"
        # the hit/known masks do not depend on the tie rule: compare with the UNPATCHED reference too
        unp = z["src_grid_" + which + "_unpatched"].reshape(out.shape)
        sentinel = unp[..., 0] > 1.5
        assert torch.equal(sentinel, ~known), f"{which}: known mask differs from the unpatched reference"
        # backward (values only)
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
        (out * w.to(dev)).sum().backward()
        (ref.grid * w).sum().backward()
        x64 = fwd.double().clone().requires_grad_(True)
        (wo.inverse_warp(x64, (H, W), erode=erode) * w.double()).sum().backward()
        grad_close(x.grad, x32.grad, x64.grad, f"inverse_warp_{which} d src_grid")


def check_occ(dev, case):
    """a-4 LVD.compute_occ (lvd.py:59-68)."""
    cfg, _, z = load_case(case)
    s = z["in_occ_score"].clone().to(dev).requires_grad_(True)
    occ = wb.compute_occ(s)
    assert float((occ.detach().cpu() - z["occ"]).abs().max()) <= 1e-6
    s32 = z["in_occ_score"].clone().requires_grad_(True)
    s64 = z["in_occ_score"].double().clone().requires_grad_(True)
    w = torch.randn(occ.shape, generator=torch.Generator().manual_seed(7))
    (occ * w.to(dev)).sum().backward()
    (wo.compute_occ(s32) * w).sum().backward()
    (wo.compute_occ(s64) * w.double()).sum().backward()
    grad_close(s.grad, s32.grad, s64.grad, "compute_occ d occ_score")


# ------------------------------------------------------------------------------------------------ decode (T1)
LEAF_KEYS = dict(input="in_input", tgo="tgt_grid_obj", sgo="src_grid_obj", tgb="tgt_grid_bg", sgb="src_grid_bg",
                 occ="occ", oar="in_obj_alpha_raw", cls="in_cls")


def oracle_decode(cfg, z, dtype, with_grad=True):
    st = wo.make_state(cfg, dtype)
    lv = {k: z[k2].to(dtype).clone().requires_grad_(with_grad) for k, k2 in LEAF_KEYS.items()}
    om, bg = wo.alpha_masks(cfg, dtype)
    oa = om * lv["oar"] + (1 - om) * (-1.0)
    ba = bg.expand(oa.shape[0], -1, -1, -1)
    out = wo.decode_output(st, lv["input"], (lv["tgo"], lv["sgo"], lv["tgb"], lv["sgb"]), lv["occ"], oa, ba, lv["cls"],
                           z["in_ctx_ts"], z["in_pred_ts"])
    if with_grad:
        loss = 0
        for n, o in zip(OUT_NAMES, out):
            if o is not None and "proj_" + n in z:
                loss = loss + (o * z["proj_" + n].to(dtype)).sum()
        loss.backward()
    return out, {k: v.grad for k, v in lv.items()}


def kernel_decode(dev, cfg, z, with_grad=True):
    opt = make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    lv = {k: z[k2].clone().to(dev).requires_grad_(with_grad) for k, k2 in LEAF_KEYS.items()}
    om, bg = wb.alpha_masks(opt)
    om = om.to(dev) if torch.is_tensor(om) else om
    oa = om * lv["oar"] + (1 - om) * (-1.0)
    ba = bg.to(dev).expand(oa.shape[0], -1, -1, -1)
    out = wb.decode_output(warper, lv["input"], (lv["tgo"], lv["sgo"], lv["tgb"], lv["sgb"]), lv["occ"], oa, ba, lv["cls"],
                           z["in_ctx_ts"].to(dev), z["in_pred_ts"].to(dev), cfg.restrict_to_ctx)
    if with_grad:
        loss = 0
        for n, o in zip(OUT_NAMES, out):
            if o is not None and "proj_" + n in z:
                loss = loss + (o * z["proj_" + n].to(dev)).sum()
        loss.backward()
    return out, {k: v.grad for k, v in lv.items()}


def check_decode(dev, case):
    """a-5..a-8 decode_output (lvd.py:141-153) from the REFERENCE's grid tuple (tier T1): forward vs the committed
    reference outputs and the oracle, gradients vs oracle autograd; fp64-arbitrated where the reference's own floor is
    above the tolerance; layer-assignment argmax maps bit-exact."""
    cfg, _, z = load_case(case)
    o32, g32 = oracle_decode(cfg, z, torch.float32)
    o64, g64 = oracle_decode(cfg, z, torch.float64)
    out, g = kernel_decode(dev, cfg, z)
    for n, o, a, b in zip(OUT_NAMES, out, o32, o64):
        if a is None:
            assert o is None, f"{n} must be None (lvd.py:825-828)"
            continue
        assert tuple(o.shape) == tuple(a.shape), f"{n}: shape {tuple(o.shape)} vs reference {tuple(a.shape)}"
        assert tuple(o.shape) == tuple(z[n].shape)
        arbitrated(o, z[n], b, FWD_TOL, f"{case}/{n} (vs reference fixture)")
        arbitrated(o, a, b, FWD_TOL, f"{case}/{n} (vs oracle)")
    # rule (1): layer assignment (logger.py:172 `alpha.max(dim=-3)[1]`) bit-exact wherever the reference's own top-2
    # margin exceeds its fp32 noise floor
    for n in ("alpha", "alpha_ctx"):
        i = OUT_NAMES.index(n)
        k, r = out[i].detach().cpu(), z[n]
        top2 = r.topk(2, dim=-3)[0]
        safe = (top2.select(-3, 0) - top2.select(-3, 1)) > 1e-4
        assert torch.equal(k.argmax(dim=-3)[safe], r.argmax(dim=-3)[safe]), f"{n}: layer argmax differs"
    for kname in LEAF_KEYS:
        grad_close(g[kname], g32[kname], g64[kname], f"{case}/d {kname}")


def _pct(e, q):
    e = e.flatten()
    return float(e.kthvalue(max(1, min(e.numel(), int(round(e.numel() * q)))))[0])


def bf16_close(out32, out16, what, report=None, exact_positions=False):
    """TOL_BF16 between a tuple of fp32 outputs (OUT_NAMES order) and the bf16-storage variant's.  exact_positions: the layout
    logits of the input are exactly representable in bf16 (one-hot +-5), so flow and alpha must be bit-identical."""
    t = TOL_BF16
    ac = out32[OUT_NAMES.index("alpha_ctx")].detach().float().cpu()
    seen = (((ac + 1) / 2).sum(3).sum(1) >= t["out_norm"]).unsqueeze(2)   # (B, Tp, 1, Hd, Wd): some context sees the pixel
    for n, x, y in zip(OUT_NAMES, out32, out16):
        assert (x is None) == (y is None), n
        if x is None:
            continue
        want = torch.float32 if n in ("flow", "alpha", "alpha_unflt") else torch.bfloat16   # alpha stays fp32: the flow is computed from it
        assert y.dtype == want and tuple(y.shape) == tuple(x.shape), f"{what}/{n}: {y.dtype} {tuple(y.shape)}"
        e = (x.detach().float().cpu() - y.detach().float().cpu()).abs()
        assert bool(torch.isfinite(e).all()), f"{what}/{n}: non-finite"
        mx, mean = float(e.max()), float(e.mean())
        if report is not None:
            report[n] = dict(max_abs=mx, mean_abs=mean, p99=_pct(e, 0.99), p999=_pct(e, 0.999))
        if exact_positions and n in ("flow", "alpha", "alpha_unflt"):
            assert mx == 0.0, f"{what}/{n}: not bit-identical with exactly representable layout logits ({mx:.3e})"
        if n == "flow":
            assert mx <= t["flow_max"], f"{what}/flow: {mx:.3e}"
        elif n in ("alpha", "alpha_unflt", "alpha_ctx", "raw_alpha"):
            assert mx <= t["alpha_max"], f"{what}/{n}: {mx:.3e}"
        elif n == "raw_output":
            assert mean <= t["raw_mean"] and _pct(e, 0.999) <= t["raw_p999"] and mx <= t["raw_max"], \
                f"{what}/raw_output: mean {mean:.3e} p99.9 {_pct(e, 0.999):.3e} max {mx:.3e}"
        elif n == "output":
            es = e[seen.expand_as(e)]
            if report is not None:
                report[n].update(seen_frac=float(seen.float().mean()), seen_max_abs=float(es.max()), seen_mean_abs=float(es.mean()))
            assert float(es.mean()) <= t["out_mean"] and float(es.max()) <= t["out_max"], \
                f"{what}/output: mean {float(es.mean()):.3e} max {float(es.max()):.3e} over the seen pixels"


def check_decode_bf16(dev, case):
    """Rule (4): the bf16-storage variant of decode_output (forward / inference) against the fp32 kernels AND the oracle on the
    same inputs, within TOL_BF16; index-valued decisions (live-layer masks, is_obj) do not depend on the storage type, which
    shows as `flow` agreeing to 1e-3.  Asking for gradients through it must raise."""
    cfg, _, z = load_case(case)
    with torch.no_grad():
        o32, _ = oracle_decode(cfg, z, torch.float32, with_grad=False)
        k32, _ = kernel_decode(dev, cfg, z, with_grad=False)
        zb = dict(z)
        zb["in_input"] = z["in_input"].to(torch.bfloat16)
        k16, _ = kernel_decode(dev, cfg, zb, with_grad=False)
    bf16_close(k32, k16, f"{case} (vs fp32 kernels)")
    bf16_close(o32, k16, f"{case} (vs oracle)")
    try:
        kernel_decode(dev, cfg, zb, with_grad=True)
    except RuntimeError as e:
        assert "forward / inference only" in str(e)
    else:
        raise AssertionError("bf16 storage with gradients must raise")


def check_decode_deterministic(dev, case):
    """Deterministic mode (64-bit fixed-point accumulation of every scatter target, include/waldo_b200.h det_* fields):
    the gradients meet the same bar as the default mode, are bit-identical from run to run, no addend left the
    fixed-point range, and the forward outputs are untouched."""
    from waldo_b200 import functional as F
    cfg, _, z = load_case(case)
    _, g32 = oracle_decode(cfg, z, torch.float32)
    _, g64 = oracle_decode(cfg, z, torch.float64)
    out0, g0 = kernel_decode(dev, cfg, z)
    assert not wb.is_deterministic()
    wb.set_deterministic(True)
    try:
        assert wb.is_deterministic()
        out1, g1 = kernel_decode(dev, cfg, z)
        sc = F.LAST_DET_SCALE.cpu()
        assert float(sc[3]) == 0.0, "fixed-point overflow flag set"
        assert float(sc[0]) * float(sc[1]) == 1.0 and float(sc[2]) > 0
        out2, g2 = kernel_decode(dev, cfg, z)
    finally:
        wb.set_deterministic(False)
    for a, b in zip(out0, out1):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.equal(a, b)
    for kname in LEAF_KEYS:
        assert torch.equal(g1[kname], g2[kname]), f"{case}/d {kname}: deterministic mode is not run-to-run identical"
        grad_close(g1[kname], g32[kname], g64[kname], f"{case}/d {kname} (deterministic)")
        # against the float-reduction path: both are sums of the same addends, differing only by rounding
        ref = g0[kname]
        assert float((g1[kname] - ref).abs().max()) <= 1e-5 * max(float(ref.abs().max()), 1e-30), f"{case}/d {kname}: det vs default"


def check_end_to_end(dev, case):
    """Tier T2: control points -> grids -> decode, plus gradients reaching every leaf (obj_pose, bg_pose, occ_score,
    obj_alpha, cls, input) through the whole chain.  The chain is discontinuous (round / dedupe / > 0.9 in the inverse
    warp) and the K = 131 / 211 background TPS system has cond ~1e4, so the reference's OWN fp32 run differs from exact
    arithmetic by O(1) on some pixels: agreement is statistical and fp64-arbitrated -- the kernel must
    be as close to the oracle's fp64 twin as the reference's fp32 outputs (the committed fixture) are, within 2x."""
    cfg, _, z = load_case(case)
    opt = make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    names = ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")
    lv = {k: z["in_" + k].clone().to(dev).requires_grad_(True) for k in names}
    om, bg = wb.alpha_masks(opt)
    om = om.to(dev) if torch.is_tensor(om) else om
    occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, lv["obj_alpha_raw"], om, bg.to(dev), lv["obj_pose"], lv["bg_pose"], lv["occ_score"])
    out = wb.decode_output(warper, lv["input"], grid, occ, oa, ba, lv["cls"], z["in_ctx_ts"].to(dev), z["in_pred_ts"].to(dev), cfg.restrict_to_ctx)
    # the fp64 twin of the whole chain
    st64 = wo.make_state(cfg, torch.float64)
    l64 = {k: z["in_" + k].double().clone().requires_grad_(True) for k in names}
    occ64, oa64, ba64, grid64 = wo.estimate_alpha_grid_occ(st64, l64["obj_alpha_raw"], l64["obj_pose"], l64["bg_pose"], l64["occ_score"])
    out64 = wo.decode_output(st64, l64["input"], grid64, occ64, oa64, ba64, l64["cls"], z["in_ctx_ts"], z["in_pred_ts"])
    loss, loss64 = 0, 0
    for n, o, o64 in zip(OUT_NAMES, out, out64):
        if o is None:
            continue
        ek = (o.detach().cpu().double() - o64.detach()).abs().mean()
        er = (z[n].double() - o64.detach()).abs().mean()
        # floor: the reference's own fp32-vs-fp64 mean discrepancy end to end (2.5e-3 on raw_output)
        assert float(ek) <= max(2 * float(er), 2.5e-3), f"{case}/{n}: mean |k - f64| {float(ek):.3e} vs reference's {float(er):.3e}"
        if "proj_" + n in z:
            loss = loss + (o * z["proj_" + n].to(dev)).sum()
            loss64 = loss64 + (o64 * z["proj_" + n].double()).sum()
    dk, dr = abs(float(loss) - float(loss64)), abs(float(z["loss"]) - float(loss64))
    assert dk <= 2 * dr + 1e-3 * max(1.0, abs(float(loss64))), f"{case}: loss {float(loss):.6f} vs f64 {float(loss64):.6f} (reference {float(z['loss']):.6f})"
    loss.backward()
    loss64.backward()
    for k, v in lv.items():
        gref, g64 = z["grad_" + k].double(), l64[k].grad
        scale = max(float(g64.abs().max()), 1e-30)
        ek = float((v.grad.cpu().double() - g64).abs().max()) / scale
        er = float((gref - g64).abs().max()) / scale
        assert ek <= 2 * er + 5e-2, f"{case}/d {k}: rel err vs f64 {ek:.3e}, reference's own {er:.3e}"
        assert float(v.grad.abs().max()) > 0 or float(gref.abs().max()) == 0


def _chain_grads(dev, cfg, d, warper, om, bg):
    lv = {k: d[k].clone().to(dev).requires_grad_(True) for k in ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")}
    occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, lv["obj_alpha_raw"], om, bg, lv["obj_pose"], lv["bg_pose"], lv["occ_score"])
    out = wb.decode_output(warper, lv["input"], grid, occ, oa, ba, lv["cls"], d["ctx_ts"].to(dev), d["pred_ts"].to(dev), cfg.restrict_to_ctx)
    output, flow, _, alpha, raw_alpha, raw, _ = out
    # a loss that reaches every output with non-uniform weights
    w = torch.linspace(0.5, 1.5, output.shape[-1], device=dev)
    loss = (output * w).sum() + (raw * raw).mean() * 1e3 + (flow * w).sum() * 0.1 + (raw_alpha * w).sum() + alpha.sum() * 1e-2
    loss.backward()
    return {k: v.grad for k, v in lv.items()}


def check_chain_deterministic(dev, cfg=None, B=1, T=5, Tc=4, seed=0):
    """Control points -> grids -> occlusion matrix -> decode -> loss, backward to every leaf, in deterministic mode: all
    gradients bit-identical from run to run (TPS / inverse-warp / occlusion backward use ordered reductions, the decode
    backward 64-bit fixed-point accumulation), and equal to the default mode up to rounding."""
    from waldo_b200 import functional as F
    cfg = cfg or wo.PathConfig()
    opt = make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    d = wo.synth_inputs(cfg, B, T, Tc, seed=seed)
    om, bg = wb.alpha_masks(opt)
    om, bg = (om.to(dev) if torch.is_tensor(om) else om), bg.to(dev)
    g0 = _chain_grads(dev, cfg, d, warper, om, bg)
    wb.set_deterministic(True)
    try:
        g1 = _chain_grads(dev, cfg, d, warper, om, bg)
        assert float(F.LAST_DET_SCALE[3]) == 0.0, "fixed-point overflow flag set"
        g2 = _chain_grads(dev, cfg, d, warper, om, bg)
    finally:
        wb.set_deterministic(False)
    for k in g1:
        assert bool(torch.isfinite(g1[k]).all()), k
        assert torch.equal(g1[k], g2[k]), f"d {k}: not bit-identical from run to run in deterministic mode"
        scale = max(float(g0[k].abs().max()), 1e-30)
        assert float((g1[k] - g0[k]).abs().max()) <= 1e-4 * scale, f"d {k}: deterministic vs default mode"
    return g0, g1


def check_wif(dev, case):
    """a-9 WIF.forward tail (wif.py:50-54) on the reference's raw_output and a seeded UNet output."""
    cfg, _, z = load_case(case)
    raw = z["raw_output"].clone().to(dev).requires_grad_(True)
    u = z["wif_unet_out"].clone().to(dev).requires_grad_(True)
    y = wb.wif_fuse(raw, u)
    assert float((y.detach().cpu() - z["wif_fused"]).abs().max()) <= FWD_TOL
    w = torch.randn(y.shape, generator=torch.Generator().manual_seed(11))
    (y * w.to(dev)).sum().backward()
    r32 = z["raw_output"].clone().requires_grad_(True)
    u32 = z["wif_unet_out"].clone().requires_grad_(True)
    (wo.wif_fuse(r32, u32) * w).sum().backward()
    r64 = z["raw_output"].double().clone().requires_grad_(True)
    u64 = z["wif_unet_out"].double().clone().requires_grad_(True)
    (wo.wif_fuse(r64, u64) * w.double()).sum().backward()
    grad_close(raw.grad, r32.grad, r64.grad, "wif d raw_output")
    grad_close(u.grad, u32.grad, u64.grad, "wif d unet_out")


def check_kats(dev):
    """Known-answer tests of SURVEY.md §4 through the product path."""
    # KAT1: TPS of the rest control points is the identity lattice
    g44 = wb.get_grid(4, 4).view(-1, 2)
    tps = wb.TPSWarp(64, 64, g44).to(dev)
    out = tps(g44[None].to(dev))
    assert float((out.cpu() - wb.get_grid(64, 64)).abs().max()) <= 2e-6
    # KAT2: TPS reproduces an affine map of the control points
    A = torch.tensor([[0.9, 0.2], [-0.1, 1.1]])
    tvec = torch.tensor([0.05, -0.02])
    out = tps((g44 @ A.t() + tvec)[None].to(dev))
    want = wb.get_grid(64, 64).view(-1, 2) @ A.t() + tvec
    assert float((out.cpu().view(-1, 2) - want).abs().max()) <= 2e-6
    # KAT3: inverse of the identity is the identity, bit-exactly; integer translation inverts exactly
    inv = wb.InverseWarp(16, 32, 16, 32).to(dev)
    ident = wb.get_grid(16, 32)
    out = inv(ident.to(dev), erode=False)
    assert torch.equal(out.cpu(), ident)
    shift = ident + torch.tensor([3 * 2 / 32, -2 * 2 / 16])
    out = inv(shift.to(dev), erode=False).cpu()
    want = ident - torch.tensor([3 * 2 / 32, -2 * 2 / 16])
    assert float((out - want).abs().max()) <= 1e-6
    # KAT5: with a zero UNet output WIF fuses sigma(v4+5)*rgb averaged over contexts
    v = torch.randn(1, 3, 2, 9, 8, 8, generator=torch.Generator().manual_seed(2))
    y = wb.wif_fuse(v.to(dev), torch.zeros(1, 2, 3, 5, 8, 8, device=dev)).cpu()
    want = (torch.sigmoid(v[:, :, :, 4:5] + 5) * v[:, :, :, :3]).mean(dim=1)
    assert float((y - want).abs().max()) <= 1e-6


def check_full_size(dev, B=1, T=5, Tc=4):
    """Full BASELINE shape through the product path only (the oracle would need minutes and tens of GB):
      * identity geometry (rest control points, zero flow): raw_output[:, :, :, :C] reproduces the context frames
        exactly and `flow` is ~0;
      * output channels are a convex combination of the warped contexts (min <= out <= max);
      * alpha_ctx is a view of raw_output; all alphas within [-1, 1]; run-to-run determinism of the forward;
      * the oracle agrees on a random crop-free sub-problem: checked separately at small sizes (check_decode)."""
    cfg = wo.PathConfig()
    opt = make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    d = wo.synth_inputs(cfg, B, T, Tc, seed=0)
    om, bg = wb.alpha_masks(opt)
    to = lambda t: t.to(dev)
    occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, to(d["obj_alpha_raw"]), to(om), to(bg), to(d["obj_pose"]), to(d["bg_pose"]), to(d["occ_score"]))
    args = (to(d["input"]), grid, occ, oa, ba, to(d["cls"]), to(d["ctx_ts"]), to(d["pred_ts"]), True)
    out1 = wb.decode_output(warper, *args)
    out2 = wb.decode_output(warper, *args)
    for a, b in zip(out1, out2):
        if a is not None:
            assert torch.equal(a, b), "forward is not run-to-run deterministic"
    output, flow, a_unflt, alpha, raw_alpha, raw, alpha_ctx = out1
    C, L = d["input"].shape[2], cfg.num_obj + 1
    assert a_unflt is None
    assert tuple(raw.shape) == (B, Tc, T - Tc, C + L, 512, 1024)
    assert alpha_ctx.data_ptr() == raw[:, :, :, C:].data_ptr()
    for t_ in (alpha, alpha_ctx, raw_alpha):
        assert float(t_.min()) >= -1 - 1e-5 and float(t_.max()) <= 1 + 1e-5
    warped = raw[:, :, :, :C]
    lo, hi = warped.min(dim=1)[0], warped.max(dim=1)[0]
    assert bool(((output >= lo - 1e-4) & (output <= hi + 1e-4)).all())
    assert bool(torch.isfinite(output).all()) and bool(torch.isfinite(flow).all())
    # identity geometry: every frame has the same control points -> zero flow -> exact copy of the context frames
    obj_pose = d["obj_pose"][:, :1].expand(-1, T, -1, -1, -1).contiguous()
    bg_pose = d["bg_pose"][:, :1].expand(-1, T, -1, -1, -1).contiguous()
    occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, to(d["obj_alpha_raw"]), to(om), to(bg), to(obj_pose), to(bg_pose), to(d["occ_score"]))
    output, flow, _, _, _, raw, _ = wb.decode_output(warper, to(d["input"]), grid, occ, oa, ba, to(d["cls"]), to(d["ctx_ts"]), to(d["pred_ts"]), True)
    assert float(flow.abs().max()) == 0.0
    want = to(d["input"])[:, :Tc].unsqueeze(2)
    assert torch.equal(raw[:, :, :, :C], want), "zero flow must copy the context frames bit-exactly"


# ------------------------------------------------------------------------------------------------ f-3 input packing
def reference_pack(rgb, label, num_lyt):
    """What the reference's dataset + Synthesizer do on the host (data/base_dataset.py:173-183, :355-372;
    models/synthesizer.py:444), restated with the same torch ops: ToTensor (x/255), Normalize(0.5, 0.5), one-hot via
    scatter_, 5 * (2x - 1), cat on the channel dim."""
    if rgb.dtype == torch.uint8:
        vid = (rgb.to(torch.float32).div(255) - 0.5) / 0.5
    else:
        vid = rgb
    B, T, Hd, Wd = label.shape
    onehot = torch.zeros(B, T, num_lyt, Hd, Wd).scatter_(2, label.long().unsqueeze(2), 1)
    return torch.cat([vid, 5 * (onehot * 2 - 1)], dim=2)


def check_pack_input(dev, B=2, T=3, Hd=12, Wd=20, num_lyt=20, seed=5):
    """Bit-exact against the reference's host-side formulas, uint8 and fp32 frames, sizes with and without a
    multiple-of-4 pixel count."""
    g = torch.Generator().manual_seed(seed)
    for (h, w) in ((Hd, Wd), (7, 9)):
        rgb8 = torch.randint(0, 256, (B, T, 3, h, w), generator=g, dtype=torch.uint8)
        lab = torch.randint(0, num_lyt, (B, T, h, w), generator=g, dtype=torch.uint8)
        want = reference_pack(rgb8, lab, num_lyt)
        got = wb.pack_input(rgb8.to(dev), lab.to(dev), num_lyt).cpu()
        assert torch.equal(got, want), "pack_input(uint8 rgb) differs from ToTensor/Normalize/one-hot"
        rgbf = torch.rand(B, T, 3, h, w, generator=g) * 2 - 1
        got = wb.pack_input(rgbf.to(dev), lab.to(dev), num_lyt).cpu()
        assert torch.equal(got, reference_pack(rgbf, lab, num_lyt))
        # bf16 storage: the same values rounded once to bf16 (round to nearest even, as Tensor.to(torch.bfloat16))
        got = wb.pack_input(rgb8.to(dev), lab.to(dev), num_lyt, dtype=torch.bfloat16).cpu()
        assert got.dtype == torch.bfloat16 and torch.equal(got, want.to(torch.bfloat16)), "pack_input(bf16) differs"
    # pinned by the reference's own dataset class: tests/golden/pack.npz holds what BaseDataset.load_rgb_path / load_layout_path
    # (data/base_dataset.py:167-183 with the transforms of :213-220) made of seeded PNG files (oracle/make_golden.pack_fixture)
    z = np.load(os.path.join(GOLDEN, "pack.npz"))
    i = 0
    while f"rgb{i}" in z.files:
        rgb8, lab, want = torch.from_numpy(z[f"rgb{i}"])[None, None], torch.from_numpy(z[f"lab{i}"])[None, None], torch.from_numpy(z[f"input{i}"])
        got = wb.pack_input(rgb8.to(dev), lab.to(dev), want.shape[0] - 3).cpu()[0, 0]
        assert torch.equal(got, want), f"pack_input differs from the reference's dataset class (fixture {i})"
        assert torch.equal(reference_pack(rgb8, lab, want.shape[0] - 3)[0, 0], want)   # ... and so does the restatement above
        i += 1
    assert i >= 2
    # every 8-bit value maps exactly as torchvision's ToTensor + Normalize
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 1, 1, 16, 16).expand(1, 1, 3, 16, 16).contiguous()
    lab = torch.zeros(1, 1, 16, 16, dtype=torch.uint8)
    assert torch.equal(wb.pack_input(ramp.to(dev), lab.to(dev), 2).cpu()[:, :, :3], reference_pack(ramp, lab, 2)[:, :, :3])


# ------------------------------------------------------------------------------------------------ f-2 loss epilogues
def check_blur(dev):
    """f-2 `blur` (synthesizer.py:1114-1118) against the outputs and autograd gradients of the REFERENCE's own function
    (tests/golden/blur.npz, oracle/make_golden.blur_fixture) and against the oracle restatement + its fp64 twin."""
    z = np.load(os.path.join(GOLDEN, "blur.npz"))
    i = 0
    while f"x{i}" in z.files:
        x, w, y_ref, g_ref = (torch.from_numpy(z[f"{n}{i}"]) for n in "xwyg")
        sigma, k = float(z[f"p{i}"][0]), int(z[f"p{i}"][1])
        xd = x.clone().to(dev).requires_grad_(True)
        y = wb.blur(xd, sigma=sigma, kernel_size=k)
        (y * w.to(dev)).sum().backward()
        x64 = x.double().requires_grad_(True)
        y64 = wo.blur(x64, sigma, k)
        (y64 * w.double()).sum().backward()
        arbitrated(y, y_ref, y64, FWD_TOL, f"blur[{i}] vs the reference")
        arbitrated(y, wo.blur(x, sigma, k), y64, FWD_TOL, f"blur[{i}] vs the oracle")
        grad_close(xd.grad, g_ref, x64.grad, f"blur[{i}] d vid")
        i += 1
    assert i >= 4


def check_layer_entropy(dev, seed=11):
    """f-2 layer entropy + fg_mask (synthesizer.py:886-889, :933) against the oracle (fp32 + fp64 twin), forward and
    gradients, on a random stack, a saturated one (alpha = +-1 exactly: the 1e-6 floors matter) and one with L = 2."""
    gen = torch.Generator().manual_seed(seed)
    cases = [torch.tanh(2 * torch.randn(2, 3, 17, 20, 36, generator=gen)),
             torch.where(torch.rand(1, 2, 5, 9, 11, generator=gen) > 0.5, torch.ones(()), -torch.ones(())),
             torch.tanh(torch.randn(1, 1, 2, 7, 5, generator=gen))]
    for i, a in enumerate(cases):
        we, wf = torch.randn(a.shape[0], a.shape[1], 1, *a.shape[3:], generator=gen), torch.randn(a.shape[0], a.shape[1], 1, *a.shape[3:], generator=gen)
        ad = a.clone().to(dev).requires_grad_(True)
        ent, fg = wb.layer_entropy(ad)
        ((ent * we.to(dev)).sum() + (fg * wf.to(dev)).sum()).backward()
        a32, a64 = a.clone().requires_grad_(True), a.double().requires_grad_(True)
        e32, f32 = wo.layer_entropy(a32)
        e64, f64 = wo.layer_entropy(a64)
        ((e32 * we).sum() + (f32 * wf).sum()).backward()
        ((e64 * we.double()).sum() + (f64 * wf.double()).sum()).backward()
        arbitrated(ent, e32, e64, FWD_TOL, f"layer_entropy[{i}]")
        arbitrated(fg, f32, f64, FWD_TOL, f"fg_mask[{i}]")
        grad_close(ad.grad, a32.grad, a64.grad, f"layer_entropy[{i}] d alpha")
        # each output alone (the other upstream gradient is None)
        ad2 = a.clone().to(dev).requires_grad_(True)
        (wb.layer_entropy(ad2)[1] * wf.to(dev)).sum().backward()
        a2 = a.clone().requires_grad_(True)
        (wo.layer_entropy(a2)[1] * wf).sum().backward()
        grad_close(ad2.grad, a2.grad, a2.grad.double(), f"fg_mask[{i}] d alpha")


def check_pose_distances(dev):
    """f-2 `cell_dis` / `center_dis` (synthesizer.py:965-979) against the scalars and autograd gradients of the REFERENCE's own
    source lines (tests/golden/pose_dis.npz, oracle/make_golden.pose_dis_fixture) and, map by map, against the oracle
    restatement + its fp64 twin; the pose gradient is a fixed-order reduction: bit-identical from run to run."""
    z = np.load(os.path.join(GOLDEN, "pose_dis.npz"))
    i = 0
    while f"pose{i}" in z.files:
        grid, pose, mov, fg = (torch.from_numpy(z[f"{n}{i}"]) for n in ("grid", "pose", "mov", "fg"))
        obj_shape, eps = (int(z[f"p{i}"][0]), int(z[f"p{i}"][1])), float(z[f"p{i}"][2])
        runs = []
        for _ in range(2):
            pd, fd = pose.clone().to(dev).requires_grad_(True), fg.clone().to(dev).requires_grad_(True)
            cell, center = wb.pose_distance_losses(mov.to(dev), fd, pd, grid.to(dev), obj_shape, eps)
            (cell * 1.5 + center * 0.75).backward()
            runs.append((cell.detach().cpu(), center.detach().cpu(), pd.grad.cpu(), fd.grad.cpu()))
        assert all(torch.equal(a, b) for a, b in zip(*runs)), f"pose_dis[{i}]: not bit-identical from run to run"
        cell, center, d_pose, d_fg = runs[0]
        p64, f64 = pose.double().requires_grad_(True), fg.double().requires_grad_(True)
        c64, m64 = wo.pose_distances(mov.double(), f64, p64, grid.double(), obj_shape, eps)
        (c64.mean() * 1.5 + m64.mean() * 0.75).backward()
        arbitrated(cell, torch.from_numpy(z[f"cell{i}"]), c64.mean(), FWD_TOL, f"cell_dis[{i}] vs the reference")
        arbitrated(center, torch.from_numpy(z[f"center{i}"]), m64.mean(), FWD_TOL, f"center_dis[{i}] vs the reference")
        grad_close(d_pose, torch.from_numpy(z[f"d_pose{i}"]), p64.grad, f"pose_dis[{i}] d obj_pose")
        grad_close(d_fg, torch.from_numpy(z[f"d_fg{i}"]), f64.grad, f"pose_dis[{i}] d fg_mask")
        # the two maps and the argmin objects, pixel by pixel (ties -- a zero weight -- go to the first object, as torch.min does on the CPU)
        with torch.no_grad():
            cm, mm, ca, ma = wb.functional.pose_distances(mov.to(dev), fg.to(dev), pose.to(dev), grid.to(dev), obj_shape, eps)
            c32, m32 = wo.pose_distances(mov, fg, pose, grid, obj_shape, eps)
            arbitrated(cm, c32, c64, FWD_TOL, f"cell_min[{i}]")
            arbitrated(mm, m32, m64, FWD_TOL, f"center_min[{i}]")
            assert int(ca.max()) < pose.shape[2] and int(ma.max()) < pose.shape[2]
        # d mov_obj_mask (no gradient in the reference: the mask comes from a threshold), where no two objects tie
        md = mov.clone().to(dev).requires_grad_(True)
        cell, center = wb.pose_distance_losses(md, fg.to(dev), pose.to(dev), grid.to(dev), obj_shape, eps)
        (cell * 1.5 + center * 0.75).backward()
        m64_ = mov.double().requires_grad_(True)
        c, m = wo.pose_distances(m64_, fg.double(), pose.double(), grid.double(), obj_shape, eps)
        (c.mean() * 1.5 + m.mean() * 0.75).backward()
        live = (mov > 0) & ((mov + eps) * (1 - fg) != 0)
        sc = float(m64_.grad.abs().max())
        assert float(((md.grad.cpu().double() - m64_.grad) * live).abs().max()) <= GRAD_TOL * sc, f"pose_dis[{i}] d mov_obj_mask"
        i += 1
    assert i >= 3
    with pytest.raises(RuntimeError, match="obj_pose must be"):
        wb.pose_distance_losses(mov.to(dev), fg.to(dev), pose[:, :, :, :-1].contiguous().to(dev), grid.to(dev), obj_shape, eps)


def check_obj_flow(dev):
    """f-2 `obj_flow` (synthesizer.py:865-868) against the scalar and the autograd gradient of the REFERENCE's own source lines
    (tests/golden/obj_flow.npz, oracle/make_golden.obj_flow_fixture) and against the oracle's fp64 twin; fixed-order reductions:
    bit-identical from run to run."""
    z = np.load(os.path.join(GOLDEN, "obj_flow.npz"))
    i = 0
    while f"alpha{i}" in z.files:
        alpha, flow = torch.from_numpy(z[f"alpha{i}"]), torch.from_numpy(z[f"flow{i}"])
        runs = []
        for _ in range(2):
            ad = alpha.clone().to(dev).requires_grad_(True)
            val = wb.obj_flow_loss(ad, flow.to(dev))
            (val * 3.0).backward()
            runs.append((val.detach().cpu(), ad.grad.cpu()))
        assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1]), f"obj_flow[{i}]: not bit-identical from run to run"
        a64 = alpha.double().requires_grad_(True)
        v64 = wo.obj_flow(a64, flow.double())
        (v64 * 3.0).backward()
        arbitrated(runs[0][0], torch.from_numpy(z[f"val{i}"]), v64, FWD_TOL, f"obj_flow[{i}] vs the reference")
        grad_close(runs[0][1], torch.from_numpy(z[f"d_alpha{i}"]), a64.grad, f"obj_flow[{i}] d rec_output_alpha")
        assert float(runs[0][1][:, :, 0].abs().max()) == 0.0, "the background layer takes no part"
        # a non-uniform upstream gradient of the per-pixel map (the T sums then depend on it)
        gen = torch.Generator().manual_seed(3 + i)
        w = torch.randn(alpha.shape[0], alpha.shape[1], *alpha.shape[3:], generator=gen)
        ad = alpha.clone().to(dev).requires_grad_(True)
        (wb.functional.obj_flow_map(ad, flow.to(dev)) * w.to(dev)).sum().backward()
        a64 = alpha.double().requires_grad_(True)
        a_ = (a64[:, :, 1:] + 1) / 2 + 1e-6
        f_ = flow.double().unsqueeze(2)
        m_ = (f_ * a_.unsqueeze(3)).sum(dim=(4, 5), keepdim=True) / a_.sum(dim=(3, 4), keepdim=True).unsqueeze(3)
        ((a_ * (f_ - m_).abs().sum(dim=3)).sum(dim=2) * w.double()).sum().backward()
        grad_close(ad.grad, a64.grad.float(), a64.grad, f"obj_flow[{i}] d alpha under a weighted map")
        i += 1
    assert i >= 3
    with pytest.raises(RuntimeError, match="expected alpha"):
        wb.obj_flow_loss(alpha.to(dev), flow[:, :, :1].contiguous().to(dev))


def check_loss_epilogue_shapes(dev, seed=23):
    """pose distances and obj_flow on shapes outside the fixtures (ragged sizes, one object, 2 .. 17 layers, lattices from 2x2 to
    4x5, more pixels than one CTA covers) against the oracle's fp64 twin."""
    gen = torch.Generator().manual_seed(seed)
    for (B, T, No, obj_shape, H, W, eps) in ((1, 1, 1, (2, 2), 5, 7, 0.01), (2, 1, 7, (4, 5), 33, 47, 0.0), (1, 2, 16, (4, 4), 40, 72, 0.1), (1, 1, 32, (2, 3), 8, 8, 0.0)):
        ys, xs = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
        grid = torch.stack([xs, ys], dim=-1)[None]
        pose = torch.rand(B, T, No, obj_shape[0] * obj_shape[1], 2, generator=gen) * 2 - 1
        mov, fg = torch.rand(B, T, 1, H, W, generator=gen), torch.rand(B, T, 1, H, W, generator=gen) * 1.1
        pd, fd = pose.clone().to(dev).requires_grad_(True), fg.clone().to(dev).requires_grad_(True)
        cell, center = wb.pose_distance_losses(mov.to(dev), fd, pd, grid.to(dev), obj_shape, eps)
        (cell - 2 * center).backward()
        p64, f64 = pose.double().requires_grad_(True), fg.double().requires_grad_(True)
        c64, m64 = wo.pose_distances(mov.double(), f64, p64, grid.double(), obj_shape, eps)
        (c64.mean() - 2 * m64.mean()).backward()
        c32, m32 = wo.pose_distances(mov, fg, pose, grid, obj_shape, eps)
        what = f"pose_dis {(B, T, No, obj_shape, H, W)}"
        arbitrated(cell, c32.mean(), c64.mean(), FWD_TOL, what + " cell_dis")
        arbitrated(center, m32.mean(), m64.mean(), FWD_TOL, what + " center_dis")
        grad_close(pd.grad, p64.grad.float(), p64.grad, what + " d obj_pose")
        grad_close(fd.grad, f64.grad.float(), f64.grad, what + " d fg_mask")
    for (B, T, L, H, W) in ((1, 1, 2, 5, 7), (2, 1, 8, 33, 47), (1, 2, 17, 40, 72), (1, 1, 33, 8, 8)):
        alpha = torch.tanh(2 * torch.randn(B, T, L, H, W, generator=gen))
        flow = torch.randn(B, T, 2, H, W, generator=gen) * 0.1
        ad = alpha.clone().to(dev).requires_grad_(True)
        val = wb.obj_flow_loss(ad, flow.to(dev))
        val.backward()
        a64 = alpha.double().requires_grad_(True)
        v64 = wo.obj_flow(a64, flow.double())
        v64.backward()
        arbitrated(val, wo.obj_flow(alpha, flow), v64, FWD_TOL, f"obj_flow {(B, T, L, H, W)}")
        grad_close(ad.grad, a64.grad.float(), a64.grad, f"obj_flow {(B, T, L, H, W)} d alpha")


def check_deterministic_large_gradients(dev, seed=1, B=8):
    """Deterministic mode on the benchmark inputs of rank 1 (seed 1), whose background-grid gradient is amplified 1e8 x (max
    |d bg_pose| 6e8 for upstream gradients of ~6): single addends exceed 2^52 fixed-point units there.  They must be
    accumulated (checked 64-bit add), not flagged: gradients finite, flag clear, equal to the float-reduction mode to rounding."""
    from waldo_b200 import functional as F, workloads as wl
    cfg, spec = wl.workload("city_train")
    T, Tc = spec["T"], spec["Tc"]
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in wl.synth_inputs(cfg, B, T, Tc, seed=seed).items()}
    opt = wl.make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    om, bg = wb.alpha_masks(opt)
    om, bg = om.to(dev), bg.to(dev)
    C, L = 3 + cfg.num_lyt, cfg.num_obj + 1
    Hd, Wd = cfg.hd_shape
    gen = torch.Generator(device=dev).manual_seed(1 + seed) if dev.type == "cuda" else torch.Generator().manual_seed(1 + seed)
    ups = [torch.randn(B, T - Tc, C, Hd, Wd, device=dev, generator=gen), torch.randn(B, Tc, T - Tc, 2, Hd, Wd, device=dev, generator=gen),
           torch.randn(B, Tc, T - Tc, C + L, Hd, Wd, device=dev, generator=gen)]
    names = ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")

    def run():
        lv = {k: d[k].detach().clone().requires_grad_(True) for k in names}
        occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, lv["obj_alpha_raw"], om, bg, lv["obj_pose"], lv["bg_pose"], lv["occ_score"])
        out = wb.decode_output(warper, lv["input"], grid, occ, oa, ba, lv["cls"], d["ctx_ts"].contiguous(), d["pred_ts"], cfg.restrict_to_ctx)
        torch.autograd.backward([out[0], out[1], out[5]], ups)
        return {k: v.grad for k, v in lv.items()}
    g0 = run()
    wb.set_deterministic(True)
    try:
        g1 = run()
        flag = float(F.LAST_DET_SCALE.cpu()[3])
    finally:
        wb.set_deterministic(False)
    assert flag == 0.0, "fixed-point overflow flag raised"
    big = max(float(g0[k].abs().max()) for k in names)
    assert big > 1e8, f"this input no longer produces the large gradients the test is about (max {big:.3e})"
    for k in names:
        assert bool(torch.isfinite(g1[k]).all()), k
        ref = g0[k]
        assert float((g1[k] - ref).abs().max()) <= 1e-4 * float(ref.abs().max()), f"d {k}: deterministic vs default"


# ------------------------------------------------------------------------------------------------ f-1 first UNet layer
TOL_TF32 = 3e-3   # conv3x3 with TF32 products (10-bit mantissa operands: activations truncated by the tensor core, weights
                  # rounded to nearest; fp32 accumulation): max|k - exact| <= 3e-3 * max|exact|


def _tf32(x, nearest=True):
    """TF32 on the host, 10 mantissa bits: cvt.rna.tf32.f32 (round to nearest, ties away; the staged weights) or what the tensor
    core makes of a raw fp32 pattern (low 13 bits ignored: toward zero; the activations)."""
    u = x.contiguous().view(torch.int32)
    return (((u + 0x1000) if nearest else u) & ~0x1fff).view(torch.float32)


def check_conv3x3(dev, seed=21):
    """f-1 first layer, `UNet.to_emb` (conv.py:9-11, :54) as WIF.forward feeds it (wif.py:33-38): against F.conv2d in fp64 on
    TF32-rounded operands (the kernel's exact arithmetic up to fp32 accumulation order: 2e-5) and on the original operands (the
    stated TF32 tolerance); ragged sizes, the WIF image permute, Cin that is not a multiple of 8, every supported Cout."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(seed)
    # (W % 4 == 0 takes the 16-byte staging; Cin 40 / 41 with Cout 16 the compile-time-unrolled WIF instantiations)
    for (B, Tc, Tp, Cin, H, W, Cout) in ((1, 2, 2, 40, 19, 45, 16), (2, 1, 3, 41, 8, 32, 8), (1, 2, 1, 24, 33, 70, 32), (1, 1, 1, 3, 9, 7, 24),
                                         (1, 2, 1, 40, 17, 64, 16), (1, 1, 2, 41, 9, 36, 16), (1, 1, 1, 16, 8, 4, 32)):
        raw = torch.randn(B, Tc, Tp, Cin, H, W, generator=gen)
        wgt = torch.randn(Cout, Cin, 3, 3, generator=gen) / (3 * Cin ** 0.5)
        want_order = raw.permute(0, 2, 1, 3, 4, 5).reshape(B * Tp * Tc, Cin, H, W)     # wif.py:33-38
        exact = F.conv2d(want_order.double(), wgt.double(), padding=1)
        exact_tf32 = F.conv2d(_tf32(want_order, nearest=False).double(), _tf32(wgt).double(), padding=1)
        got = wb.wif_to_emb(raw.to(dev), wgt.to(dev)).cpu().double()
        assert tuple(got.shape) == tuple(exact.shape)
        scale = float(exact.abs().max())
        assert float((got - exact_tf32).abs().max()) <= 2e-5 * scale, f"conv3x3 vs TF32-operand convolution: {float((got - exact_tf32).abs().max()) / scale:.3e}"
        assert float((got - exact).abs().max()) <= TOL_TF32 * scale, f"conv3x3 vs exact convolution: {float((got - exact).abs().max()) / scale:.3e}"
        plain = wb.conv3x3(want_order.contiguous().to(dev), wgt.to(dev)).cpu().double()
        assert torch.equal(plain, got), "conv3x3: the permuted addressing and the plain one disagree"
    # gradients: d raw_output through the same kernel (flipped, transposed weights; Tc <-> Tp), d weight through torch
    for (B, Tc, Tp, Cin, H, W, Cout) in ((1, 2, 2, 40, 12, 36, 16), (1, 3, 1, 16, 9, 20, 8), (2, 1, 2, 48, 8, 33, 16), (1, 1, 2, 41, 8, 12, 16)):
        raw = torch.randn(B, Tc, Tp, Cin, H, W, generator=gen)
        wgt = torch.randn(Cout, Cin, 3, 3, generator=gen) / (3 * Cin ** 0.5)
        proj = torch.randn(B * Tp * Tc, Cout, H, W, generator=gen)
        rd, wd = raw.clone().to(dev).requires_grad_(True), wgt.clone().to(dev).requires_grad_(True)
        (wb.wif_to_emb(rd, wd) * proj.to(dev)).sum().backward()
        r64, w64 = raw.double().requires_grad_(True), wgt.double().requires_grad_(True)
        (F.conv2d(r64.permute(0, 2, 1, 3, 4, 5).reshape(B * Tp * Tc, Cin, H, W), w64, padding=1) * proj.double()).sum().backward()
        for name, k, e in (("d raw_output", rd.grad, r64.grad), ("d weight", wd.grad, w64.grad)):
            err = float((k.detach().cpu().double() - e).abs().max()) / float(e.abs().max())
            assert err <= TOL_TF32, f"conv3x3 {name}: {err:.3e}"
    # UNet.from_emb (conv.py:37, :63): 2 x 16 -> 5 channels at full resolution, forward and both gradients
    x = torch.randn(3, 32, 10, 24, generator=gen)
    wgt = torch.randn(5, 32, 3, 3, generator=gen) / (3 * 32 ** 0.5)
    proj = torch.randn(3, 5, 10, 24, generator=gen)
    xd, wd = x.clone().to(dev).requires_grad_(True), wgt.clone().to(dev).requires_grad_(True)
    y = wb.conv3x3(xd, wd)
    (y * proj.to(dev)).sum().backward()
    x64, w64 = x.double().requires_grad_(True), wgt.double().requires_grad_(True)
    y64 = F.conv2d(x64, w64, padding=1)
    (y64 * proj.double()).sum().backward()
    for name, k, e in (("from_emb", y, y64), ("from_emb d x", xd.grad, x64.grad), ("from_emb d weight", wd.grad, w64.grad)):
        err = float((k.detach().cpu().double() - e.detach()).abs().max()) / float(e.abs().max())
        assert err <= TOL_TF32, f"conv3x3 {name}: {err:.3e}"


# ------------------------------------------------------------------------------------------------ f-4 output side
def check_frames_to_u8(dev, seed=9):
    """Bit-exact against the reference's own formulas (tools/utils.py:246-249 normalize, :258-264 dump_video), restated
    with the same torch ops, on values inside, outside and exactly on the span, multiple-of-4 and ragged pixel counts."""
    g = torch.Generator().manual_seed(seed)
    for shape in ((2, 3, 3, 16, 24), (5, 3, 7, 9)):
        vid = torch.rand(*shape, generator=g) * 2.6 - 1.3
        vid.view(-1)[:6] = torch.tensor([-1.0, 1.0, 0.0, -1.5, 1.5, 1.0 - 1e-7])
        t = vid.clamp(-1, 1)
        t = (t - (-1)) / (1 - (-1))
        want = (t.movedim(-3, -1) * 255).to(dtype=torch.uint8)
        got = wb.frames_to_u8(vid.to(dev)).cpu()
        assert got.shape == want.shape and torch.equal(got, want), "frames_to_u8 differs from normalize + dump_video"
    ramp = torch.linspace(-1, 1, 3 * 64 * 64).view(1, 3, 64, 64)
    want = (((ramp.clamp(-1, 1) + 1) / 2).movedim(-3, -1) * 255).to(torch.uint8)
    assert torch.equal(wb.frames_to_u8(ramp.to(dev)).cpu(), want)


# ------------------------------------------------------------------------------------------------ a-5 / a-11
def check_field_warps(dev, case):
    """Stand-alone obj/bg/layer_to_output (lvd.py:533-559) and the MAT propagation flows (lvd.py:575-600) against the same
    steps written with the oracle's grid_sample / interpolate wrappers, on the reference's own grids."""
    cfg, (B, T, Tc), z = load_case(case)
    opt = make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    grid = tuple(z[k] for k in ("tgt_grid_obj", "src_grid_obj", "tgt_grid_bg", "src_grid_bg"))
    gdev = tuple(g.to(dev) for g in grid)
    tgo, sgo, tgb, sgb = grid
    No = sgo.shape[2]
    Ho, Wo = cfg.obj_hw
    H, W = cfg.lo_shape
    gen = torch.Generator().manual_seed(21)
    obj = torch.randn(B, No, 3, Ho, Wo, generator=gen)
    bg = torch.randn(B, T, 2, H, W, generator=gen)
    # a-5, both deltas (delta = 1: out-of-range taps read as -1)
    for delta in (1, 0):
        want_o = wo.bil0(obj.unsqueeze(1).expand(-1, T, -1, -1, -1, -1).reshape(B * T * No, 3, Ho, Wo) + delta,
                         sgo.reshape(B * T * No, H, W, 2)).view(B, T, No, 3, H, W) - delta
        got_o = warper.obj_to_output(obj.to(dev), gdev, delta).cpu()
        assert float((got_o - want_o).abs().max()) <= FWD_TOL, f"obj_to_output delta={delta}"
        want_b = wo.bil0(bg.reshape(B * T, 2, H, W) + delta, sgb.reshape(B * T, H, W, 2)).view(B, T, 1, 2, H, W) - delta
        got_b = warper.bg_to_output(bg.to(dev), gdev, delta).cpu()
        assert float((got_b - want_b).abs().max()) <= FWD_TOL, f"bg_to_output delta={delta}"
    lay = warper.layer_to_output(obj[:, :, :2].to(dev), bg.to(dev), gdev, 0, 0).cpu()
    assert tuple(lay.shape) == (B, T, No + 1, 2, H, W)
    # a-11
    ctx_len, ref, obj_id = Tc, Tc - 1, min(1, No - 1)
    s = cfg.scale_hd

    def want_flow(diff, g):   # diff (B,t,h,w,2) in a canonical frame, g (B,t,H,W,2)
        b_, t_ = diff.shape[:2]
        f = wo.bil0(diff.permute(0, 1, 4, 2, 3).reshape(b_ * t_, 2, *diff.shape[2:4]), g.reshape(b_ * t_, H, W, 2))
        return wo.resize(f, s).view(b_, t_, 2, int(H * s), int(W * s)).permute(0, 1, 3, 4, 2)

    got = warper.grid_to_bg_flow_from_ref_to_pred(gdev, ctx_len, ref).cpu()
    assert float((got - want_flow(tgb[:, [ref]] - tgb[:, ctx_len:], sgb[:, ctx_len:])).abs().max()) <= FWD_TOL
    got = warper.grid_to_bg_flow_from_ctx_to_ref(gdev, ctx_len, ref).cpu()
    assert float((got - want_flow(tgb[:, :ctx_len] - tgb[:, [ref]], sgb[:, [ref]].repeat(1, ctx_len, 1, 1, 1))).abs().max()) <= FWD_TOL
    got = warper.grid_to_obj_flow_from_ref_to_pred(gdev, ctx_len, ref, obj_id).cpu()
    want = want_flow(tgo[:, [ref], obj_id] - tgo[:, ctx_len:, obj_id], sgo[:, ctx_len:, obj_id])
    assert float((got - want).abs().max()) <= FWD_TOL
    # forward-only helpers refuse to silently drop a gradient
    try:
        warper.bg_to_output(bg.to(dev).requires_grad_(True), gdev, 0)
        raise AssertionError("expected NotImplementedError")
    except NotImplementedError:
        pass


# ------------------------------------------------------------------------------------------------ oracle-direct cases
# Shapes the committed reference fixtures do not hold, checked against the (fixture-pinned) oracle directly: the
# benchmark's 4-context fast paths, several future frames per video, frame sizes that are not multiples of the 32x8
# pixel tile, an odd channel count.
SYNTH_CASES = {
    # name: (PathConfig kwargs, B, T, Tc)
    "tc4_ragged": (dict(dim=10, load_dim=20, aspect_ratio=2.2, num_obj=5, num_lyt=6, latent_shape=(2, 4)), 2, 6, 4),
    "tc4_x4": (dict(dim=8, load_dim=32, aspect_ratio=2.0, num_obj=4, num_lyt=20, latent_shape=(2, 4)), 1, 5, 4),
    "tc3_odd": (dict(dim=12, load_dim=24, aspect_ratio=1.5, num_obj=2, num_lyt=5, latent_shape=(3, 4)), 1, 5, 3),
}


# Checked on the host emulation only so far (added after the round's last GPU session): the square, scale-2 shape family
# of BASELINE configs[4] (bench.py --workload nonrigid_train), reduced.
SYNTH_CASES_EMU_ONLY = {
    "square_x2": (dict(dim=12, load_dim=24, aspect_ratio=1.0, num_obj=3, num_lyt=20, latent_shape=(2, 2)), 2, 5, 4),
}


def synth_case(name):
    kw, B, T, Tc = (SYNTH_CASES.get(name) or SYNTH_CASES_EMU_ONLY[name])
    cfg = wo.PathConfig(**kw)
    d = wo.synth_inputs(cfg, B, T, Tc, seed=17, radius=0.2)
    st = wo.make_state(cfg)
    with torch.no_grad():
        occ, _, _, grid = wo.estimate_alpha_grid_occ(st, d["obj_alpha_raw"], d["obj_pose"], d["bg_pose"], d["occ_score"])
    z = {"in_input": d["input"], "tgt_grid_obj": grid[0], "src_grid_obj": grid[1], "tgt_grid_bg": grid[2], "src_grid_bg": grid[3],
         "occ": occ, "in_obj_alpha_raw": d["obj_alpha_raw"], "in_cls": d["cls"],
         "in_ctx_ts": d["ctx_ts"].contiguous(), "in_pred_ts": d["pred_ts"]}
    with torch.no_grad():
        out, _ = oracle_decode(cfg, z, torch.float32, with_grad=False)
    gen = torch.Generator().manual_seed(5)
    for n, o in zip(OUT_NAMES, out):
        if o is not None:
            z["proj_" + n] = torch.randn(o.shape, generator=gen)
    return cfg, z


def check_decode_synth(dev, name):
    """decode_output against the oracle (fp32, fp64-arbitrated) on a shape outside the reference fixtures."""
    cfg, z = synth_case(name)
    o32, g32 = oracle_decode(cfg, z, torch.float32)
    o64, g64 = oracle_decode(cfg, z, torch.float64)
    out, g = kernel_decode(dev, cfg, z)
    for n, o, a, b in zip(OUT_NAMES, out, o32, o64):
        if a is None:
            assert o is None
            continue
        assert tuple(o.shape) == tuple(a.shape), f"{n}: shape {tuple(o.shape)} vs oracle {tuple(a.shape)}"
        arbitrated(o, a, b, FWD_TOL, f"{name}/{n} (vs oracle)")
    for kname in LEAF_KEYS:
        grad_close(g[kname], g32[kname], g64[kname], f"{name}/d {kname}")


# ------------------------------------------------------------------------------------------------ the benchmarked shapes
# the shapes that are shipped and benchmarked, compared with the oracle (fp32 + fp64 twin) and -- where
# the staged reference oracle/_ref (or /root/reference) is present -- with the reference's own code, at FULL size:
# 16 objects x 4x4 control points, K = 131 / 211 / 67 background TPS systems, 17 layers, scale_hd 4 / 2, the dense-layer
# branch, the FAST gather kernels at real tile counts.
FULL_SHAPES = {
    # name: (PathConfig kwargs, B, T, Tc)          BASELINE.json configs
    "city_512x1024": (dict(), 1, 5, 4),                                                                      # [0], [1], [3]
    "kitti_256x832": (dict(dim=128, load_dim=256, aspect_ratio=3.25, latent_shape=(8, 26), num_lyt=19), 1, 6, 4),   # [2]
    "nonrigid_256x256": (dict(dim=128, load_dim=256, aspect_ratio=1.0, latent_shape=(8, 8)), 2, 5, 4),       # [4]
}
REDUCED_SHAPES = {   # same structure at 1/4 of the resolution: the host-emulation (CPU suite) version of the same check
    "city_128x256": (dict(dim=32, load_dim=128), 1, 5, 4),
    "kitti_64x208": (dict(dim=32, load_dim=64, aspect_ratio=3.25, latent_shape=(8, 26), num_lyt=19), 1, 6, 4),
    "nonrigid_64x64": (dict(dim=32, load_dim=64, aspect_ratio=1.0, latent_shape=(8, 8)), 2, 5, 4),
}


def check_full_shape_bf16(dev, name, report=None):
    """Rule (4) at the benchmarked shapes: the bf16-storage forward against the fp32 kernels on the same inputs and grids (chain from
    the control points), TOL_BF16; `flow` -- and with it every index-valued decision behind it -- within 1e-3."""
    kw, B, T, Tc = (FULL_SHAPES.get(name) or REDUCED_SHAPES[name])
    cfg = wo.PathConfig(**kw)
    opt = make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in wo.synth_inputs(cfg, B, T, Tc, seed=0).items()}
    om, bg = wb.alpha_masks(opt)
    om = om.to(dev) if torch.is_tensor(om) else om
    with torch.no_grad():
        occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, d["obj_alpha_raw"], om, bg.to(dev), d["obj_pose"], d["bg_pose"], d["occ_score"])
        args = (grid, occ, oa, ba, d["cls"], d["ctx_ts"].contiguous(), d["pred_ts"], cfg.restrict_to_ctx)
        k32 = wb.decode_output(warper, d["input"], *args)
        k16 = wb.decode_output(warper, d["input"].to(torch.bfloat16), *args)
    rep = report if report is not None else {}
    bf16_close(k32, k16, f"{name} bf16 storage", rep, exact_positions=True)   # synth_inputs: one-hot +-5 layout logits
    return rep


def _argmax_report(k, r, what, tol=FWD_TOL):
    """Rule (1): the layer-assignment map (logger.py:172 `alpha.max(dim=-3)[1]`) is identical.  No margin filter: every
    mismatching pixel is counted, and each one must be a numerical tie of the reference itself (top-2 margin within twice
    the forward tolerance, i.e. the two layers are interchangeable within rule (2))."""
    ka, ra = k.argmax(dim=-3), r.argmax(dim=-3)
    bad = ka != ra
    n_bad = int(bad.sum())
    if n_bad:
        top2 = r.topk(2, dim=-3)[0]
        margin = (top2.select(-3, 0) - top2.select(-3, 1))[bad]
        assert float(margin.max()) <= 2 * tol, f"{what}: {n_bad} argmax mismatches, largest reference margin {float(margin.max()):.3e}"
    return n_bad, ka.numel()


def check_full_shape(dev, name, with_reference=True, report=None):
    kw, B, T, Tc = (FULL_SHAPES.get(name) or REDUCED_SHAPES[name])
    cfg = wo.PathConfig(**kw)
    opt = make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    d = wo.synth_inputs(cfg, B, T, Tc, seed=0)
    H, W = cfg.lo_shape
    rep = report if report is not None else {}
    st32, st64 = wo.make_state(cfg), wo.make_state(cfg, torch.float64)
    # ---------------- stage A, T0: TPS at the real K (fp64-arbitrated), inverse warp fed the ORACLE's tgt_grid (bit-exact maps)
    with torch.no_grad():
        for which, tps, b32, b64, pts in (("obj", warper.tps_obj, st32.tps_obj, st64.tps_obj, d["obj_pose"].reshape(-1, d["obj_pose"].shape[-2], 2)),
                                          ("bg", warper.tps_bg, st32.tps_bg, st64.tps_bg, d["bg_pose"].reshape(-1, d["bg_pose"].shape[-2], 2))):
            out = tps(pts.to(dev))
            o32, o64 = wo.tps_eval(b32, pts), wo.tps_eval(b64, pts.double())
            arbitrated(out, o32, o64, FWD_TOL, f"{name}/tps_{which}")
            ek, er = float((out.cpu().double() - o64).abs().max()), float((o32.double() - o64).abs().max())
            assert ek <= er + 1e-6, f"{name}/tps_{which}: kernel {ek:.3e} further from fp64 than the reference's fp32 {er:.3e}"
            rep[f"tps_{which}_err_vs_f64"] = (ek, er)
        occ32, oa32, ba32, grid32 = wo.estimate_alpha_grid_occ(st32, d["obj_alpha_raw"], d["obj_pose"], d["bg_pose"], d["occ_score"])
        for which, mod, fwd, erode in (("obj", warper.invert_obj, grid32[0], True), ("bg", warper.invert_bg, grid32[2], False)):
            fwd = fwd.reshape(-1, *fwd.shape[-3:])
            box = []
            out = mod(fwd.to(dev), erode=erode, trace_box=box)
            tr = box[0]
            ref = wo.inverse_warp(fwd, (H, W), erode=erode, trace=True)
            P = H * W
            assert torch.equal(tr.field.cpu().long(), ref.field), f"{name}/{which}: field differs"
            win = tr.winner.cpu().long()
            assert torch.equal(torch.where(win == 2 ** 31 - 1, torch.full_like(win, P), win), ref.winner), f"{name}/{which}: winners differ"
            m = 6
            assert torch.equal(((tr.level != 255) & (tr.eroded == 0))[:, m:-m, m:-m].cpu(), ref.known), f"{name}/{which}: known mask differs"
            assert torch.equal((tr.level[:, m:-m, m:-m] == 0).cpu(), ref.hit), f"{name}/{which}: hit mask differs"
            assert float((out.cpu() - ref.grid).abs().max()) <= FWD_TOL, f"{name}/{which}: inverse-warp values"
        occ_k = wb.compute_occ(d["occ_score"].to(dev))
        assert float((occ_k.cpu() - occ32).abs().max()) <= 1e-6
    # ---------------- decode, T1 (identical grid tuple on both sides): every output, all 8 leaf gradients
    z = {"in_input": d["input"], "tgt_grid_obj": grid32[0], "src_grid_obj": grid32[1], "tgt_grid_bg": grid32[2], "src_grid_bg": grid32[3],
         "occ": occ32, "in_obj_alpha_raw": d["obj_alpha_raw"], "in_cls": d["cls"],
         "in_ctx_ts": d["ctx_ts"].contiguous(), "in_pred_ts": d["pred_ts"]}
    trace = {}
    with torch.no_grad():
        om, bg = wo.alpha_masks(cfg)
        oa = om * d["obj_alpha_raw"] + (1 - om) * (-1.0)
        o_probe = wo.decode_output(st32, d["input"], grid32, occ32, oa, bg.expand(B, -1, -1, -1), d["cls"], z["in_ctx_ts"], z["in_pred_ts"], trace=trace)
    gen = torch.Generator().manual_seed(5)
    for n, o in zip(OUT_NAMES, o_probe):
        if o is not None:
            z["proj_" + n] = torch.randn(o.shape, generator=gen)
    del o_probe
    o32, g32 = oracle_decode(cfg, z, torch.float32)
    o32 = [o.detach() if o is not None else None for o in o32]
    out, g = kernel_decode(dev, cfg, z)
    out = [o.detach().cpu() if o is not None else None for o in out]
    g = {k: v.cpu() for k, v in g.items()}
    o64, g64 = oracle_decode(cfg, z, torch.float64)
    o64 = [o.detach() if o is not None else None for o in o64]
    for n, o, a, b in zip(OUT_NAMES, out, o32, o64):
        if a is None:
            assert o is None, f"{n} must be None (lvd.py:825-828)"
            continue
        assert tuple(o.shape) == tuple(a.shape), f"{n}: shape {tuple(o.shape)} vs oracle {tuple(a.shape)}"
        arbitrated(o, a, b, FWD_TOL, f"{name}/{n} (vs oracle)")
        rep[f"{n}_max_abs"] = [float((o - a).abs().max()), float((o.double() - b).abs().max()), float((a.double() - b).abs().max())]
    for kname in LEAF_KEYS:
        grad_close(g[kname], g32[kname], g64[kname], f"{name}/d {kname}")
        sc = max(float(g64[kname].abs().max()), 1e-30)
        # [kernel vs oracle fp32, kernel vs fp64 twin, oracle fp32 vs fp64 twin (= the reference arithmetic's own floor)]
        rep[f"d_{kname}_rel"] = [float((g[kname].double() - g32[kname].double()).abs().max()) / sc,
                                 float((g[kname].double() - g64[kname]).abs().max()) / sc,
                                 float((g32[kname].double() - g64[kname]).abs().max()) / sc]
    # ---------------- rule (1): index / threshold maps
    ia, ic = OUT_NAMES.index("alpha"), OUT_NAMES.index("alpha_ctx")
    rep["alpha_argmax_mismatch"] = _argmax_report(out[ia], o32[ia], f"{name}/alpha")
    rep["alpha_ctx_argmax_mismatch"] = _argmax_report(out[ic], o32[ic], f"{name}/alpha_ctx")
    if "is_obj" in trace:
        # is_obj (lvd.py:788-791) is not an output; it gates alpha_ctx: where the reference's mask is off the composited
        # opacity is EXACTLY 0 (alpha_ctx == -1), where it is on and the reference shows the layer, so does the kernel
        off = ~trace["is_obj"].unsqueeze(1).expand_as(out[ic])
        assert bool((out[ic][off] == -1).all()), f"{name}: a layer shows where the reference's is_obj mask is off"
        shown = (~off) & (o32[ic] > -1 + 1e-4)
        assert bool((out[ic][shown] > -1).all()), f"{name}: a layer is hidden where the reference's is_obj mask is on"
    # ---------------- the reference's own code on the same inputs (staged copy oracle/_ref on the GPU box)
    if with_reference:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_runner
        if ref_runner.available():
            rp = ref_runner.RefPath(cfg)
            lv = {k: z[k2].clone().requires_grad_(True) for k, k2 in LEAF_KEYS.items()}
            oa_r = rp.om * lv["oar"] + (1 - rp.om) * (-1.0)
            r_out = rp.decode(lv["input"], (lv["tgo"], lv["sgo"], lv["tgb"], lv["sgb"]), lv["occ"], oa_r, rp.bg.expand(B, -1, -1, -1),
                              lv["cls"], z["in_ctx_ts"], z["in_pred_ts"])
            sum((o * z["proj_" + n]).sum() for n, o in zip(OUT_NAMES, r_out) if o is not None).backward()
            for n, o, r, b in zip(OUT_NAMES, out, r_out, o64):
                assert (o is None) == (r is None), n
                if r is not None:
                    arbitrated(o, r.detach(), b, FWD_TOL, f"{name}/{n} (vs the reference's own code)")
                    rep[f"{n}_max_abs_vs_reference"] = float((o - r.detach()).abs().max())
            for kname in LEAF_KEYS:
                grad_close(g[kname], lv[kname].grad, g64[kname], f"{name}/d {kname} (vs the reference's autograd)")
            rep["reference"] = "staged copy oracle/_ref" if ref_runner.ref_loader.is_staged_copy() else ref_runner.ref_loader.REF_ROOT
        else:
            rep["reference"] = "absent"
    return rep
