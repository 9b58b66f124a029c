#!/usr/bin/env python
"""Diagnostic: histogram of the number of live layers per warp-row (32 HD pixels) in the context-alpha kernels
(live_ctx) and the layer kernels (live_pred & co) for the benchmark workload."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # run from profiles/ or scratch/
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import waldo_oracle as wo
import waldo_b200 as wb
from waldo_b200 import functional as F
from tests.parity import make_opt
dev = torch.device("cuda:0")
cfg = wo.PathConfig(); B, T, Tc = 2, 5, 4
opt = make_opt(cfg); warper = wb.Warper(opt).to(dev)
om, bg = wb.alpha_masks(opt); om, bg = om.to(dev), bg.to(dev)
d = wo.synth_inputs(cfg, B, T, Tc, seed=0)
cap = {}
orig = F._fwd_struct
def patched(g, pc, t):
    cap["t"] = t
    return orig(g, pc, t)
F._fwd_struct = patched
to = lambda t: t.to(dev)
with torch.no_grad():
    occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, to(d["obj_alpha_raw"]), om, bg, to(d["obj_pose"]), to(d["bg_pose"]), to(d["occ_score"]))
    wb.decode_output(warper, to(d["input"]), grid, occ, oa, ba, to(d["cls"]), to(d["ctx_ts"]), to(d["pred_ts"]), True)
t = cap["t"]
live_ctx, live_pred = t[19], t[20]
H, W = cfg.lo_shape; Hd, Wd = cfg.hd_shape
def axis(n_hd, n_lo):
    X = torch.arange(n_hd, device=dev, dtype=torch.float32)
    src = ((X + 0.5) * (n_lo / n_hd) - 0.5).clamp(min=0)
    i0 = src.floor().long().clamp(max=n_lo - 1); i1 = (i0 + 1).clamp(max=n_lo - 1)
    return i0, i1
def hist(live, name):
    y0, y1 = axis(Hd, H); x0, x1 = axis(Wd, W)
    rows = live[..., y0, :] | live[..., y1, :]          # (..., Hd, W)
    nw = Wd // 32
    out = torch.zeros(*rows.shape[:-1], nw, dtype=torch.int32, device=dev)
    for w in range(nw):
        lo, hi = int(x0[32 * w]), int(x1[32 * w + 31])
        acc = rows[..., lo]
        for c in range(lo + 1, hi + 1):
            acc = acc | rows[..., c]
        out[..., w] = acc
    n = torch.zeros_like(out)
    for b in range(17):
        n += (out >> b) & 1
    h = torch.bincount(n.flatten().long(), minlength=18).float()
    h = h / h.sum()
    print(name, "mean n %.2f" % float((h * torch.arange(len(h), device=dev)).sum()), " ".join(f"{i}:{100*float(v):.1f}%" for i, v in enumerate(h) if v > 0))
hist(live_ctx, "live_ctx (context-alpha kernels)")
hist(live_pred, "live_pred (target frames)")
