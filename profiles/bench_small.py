#!/usr/bin/env python
"""Roofline lines for the path's smaller entry points (not part of bench.py's step): the WIF fuse tail (a-9), input
packing (f-3) and Warper.forward (a-1..a-3) at the Cityscapes training shape.  One JSON line each; run on the B200."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import waldo_b200 as wb
import waldo_oracle as wo
from tests.parity import make_opt

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B, Tc, Tp, C, L, Hd, Wd = 8, 4, 1, 23, 17, 512, 1024
px = Hd * Wd
# ---- a-9 WIF fuse tail, forward and forward+backward
raw = torch.randn(B, Tc, Tp, C + L, Hd, Wd, device=dev)
u = torch.randn(B, Tp, Tc, 5, Hd, Wd, device=dev)
ms = timed(lambda: wb.wif_fuse(raw, u))
by = px * 4 * B * Tp * (Tc * (5 + 4) + 3)        # reads 5 raw + 4 used UNet channels per context, writes 3
print(json.dumps({"kernel": "k_wif_fuse_fwd", "ms": ms, "alg_bytes": by, "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak}))
rg, ug = raw.clone().requires_grad_(True), u.clone().requires_grad_(True)
gy = torch.randn(B, Tp, 3, Hd, Wd, device=dev)


def fb():
    rg.grad = None; ug.grad = None
    wb.wif_fuse(rg, ug).backward(gy)


ms2 = timed(fb, n=10)
print(json.dumps({"kernel": "wif_fuse fwd+bwd (incl. the zero-fill of d raw_output)", "ms": ms2}))
del raw, u, rg, ug, gy
# ---- f-3 input packing
rgb = torch.randint(0, 256, (B, 5, 3, Hd, Wd), device=dev, dtype=torch.uint8)
lab = torch.randint(0, 20, (B, 5, Hd, Wd), device=dev, dtype=torch.uint8)
out = torch.empty(B, 5, 23, Hd, Wd, device=dev)
ms = timed(lambda: wb.pack_input(rgb, lab, 20, out=out))
by = B * 5 * px * (4 + 23 * 4)
print(json.dumps({"kernel": "k_pack_input", "ms": ms, "alg_bytes": by, "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak}))
del rgb, lab, out
# ---- a-1..a-3 Warper.forward (TPS + inverse warps), T = 5 frames, 16 objects
cfg = wo.PathConfig()
warper = wb.Warper(make_opt(cfg)).to(dev)
d = wo.synth_inputs(cfg, B, 5, 4, seed=0)
op, bp = d["obj_pose"].to(dev), d["bg_pose"].to(dev)
with torch.no_grad():
    ms = timed(lambda: warper(op, bp), n=10)
print(json.dumps({"kernel": "Warper.forward (B=8, T=5, 16 objects: 640 + 40 inverse warps)", "ms": ms}))
