#!/usr/bin/env python
"""Where does the step go?  Event timeline of the training step: durations of the instrumented stages and the gaps between them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import argparse, types
import torch
import bench
from waldo_b200 import functional as Fn, modules as M

args = types.SimpleNamespace(no_graph=True, no_input_grad=False)
cfg, spec = bench.workload_cfg(sys.argv[1] if len(sys.argv) > 1 else "city_train")
bench.load_peak()
r = bench.Runner(args, cfg, spec, 0, 0, 1)
for _ in range(5):
    r.step(r.resident)
torch.cuda.synchronize()
Fn.PROFILE = {}
N = 6
marks = []
for i in range(N):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
    r.step(r.resident)
e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
torch.cuda.synchronize()
prof = Fn.PROFILE; Fn.PROFILE = None
order = ["decode_fwd:prep", "decode_fwd:alpha_prep", "decode_fwd:layers", "decode_fwd:gather", "decode_bwd:gather", "decode_bwd:layers", "decode_bwd:alpha_prep", "decode_bwd:rest"]
for i in range(1, N):
    t0 = marks[i]
    line = [f"step {i}: total {marks[i].elapsed_time(marks[i+1]):.3f} |"]
    prev = t0
    for k in order:
        a, b = prof[k][i]
        line.append(f"gap {prev.elapsed_time(a):.3f} {k.split(':')[0][7:]}:{k.split(':')[1]} {a.elapsed_time(b):.3f}")
        prev = b
    line.append(f"tail {prev.elapsed_time(marks[i+1]):.3f}")
    print(" ".join(line), flush=True)
