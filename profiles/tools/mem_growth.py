#!/usr/bin/env python
"""Does reserved device memory keep growing over steps?"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
import waldo_b200 as wb
from waldo_b200 import modules as M
args = types.SimpleNamespace(no_graph=True, no_input_grad=False)
cfg, spec = bench.workload_cfg(sys.argv[1] if len(sys.argv) > 1 else "city_train")
bench.load_peak()
det = os.environ.get("DET", "0") == "1"
M.OVERLAP_BG = os.environ.get("NO_OVERLAP", "0") != "1"
r = bench.Runner(args, cfg, spec, 0, 0, 1, deterministic=det)
wb.set_deterministic(det)
out = []
for i in range(40):
    r.step(r.resident)
    if os.environ.get("GC_EACH_STEP") == "1":
        import gc
        gc.collect()
    ms = torch.cuda.memory_stats()
    out.append((round(torch.cuda.memory_reserved() / 2**30, 2), ms.get("num_device_alloc", 0), round(torch.cuda.memory_allocated() / 2**30, 2)))
torch.cuda.synchronize()
print("det", det, "overlap_bg", M.OVERLAP_BG, "prefill", wb.functional.PREFILL)
print(" reserved GiB:", [o[0] for o in out[::3]])
print(" device allocs:", [o[1] for o in out[::3]])
print(" allocated GiB (live tensors at the end of the step):", [o[2] for o in out[::3]])
