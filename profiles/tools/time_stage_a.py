#!/usr/bin/env python
"""Time stage A (Warper.forward: TPS + inverse warps, and its backward) alone, per variant, with CUDA events.
usage: python scratch/time_stage_a.py [workload]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import waldo_b200 as wb
from waldo_b200 import workloads as wl, modules as M

name = sys.argv[1] if len(sys.argv) > 1 else "city_train"
cfg, spec = wl.workload(name)
dev = torch.device("cuda:0")
warper = wb.Warper(wl.make_opt(cfg)).to(dev)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in wl.synth_inputs(cfg, spec["B"], spec["T"], spec["Tc"], seed=0).items()}


def run(backward, steps=20):
    def once():
        op = d["obj_pose"].detach().requires_grad_(backward)
        bp = d["bg_pose"].detach().requires_grad_(backward)
        with torch.set_grad_enabled(backward):
            g = warper(op, bp)
        if backward:
            torch.autograd.backward(list(g), [torch.ones_like(t) for t in g])
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        once()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for label, env, overlap in (("fused+overlap", {}, True), ("fused", {}, False), ("unfused+overlap", {"WALDO_INV_UNFUSED": "1"}, True), ("unfused", {"WALDO_INV_UNFUSED": "1"}, False)):
    for k in ("WALDO_INV_UNFUSED",):
        os.environ.pop(k, None)
    os.environ.update(env)
    M.OVERLAP_BG = overlap
    print(f"{name:14s} {label:16s} fwd {run(False):7.3f} ms   fwd+bwd {run(True):7.3f} ms", flush=True)
