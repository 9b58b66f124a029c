#!/usr/bin/env python
"""Why is the deterministic step 3x slower under torchrun?  Per-step device time + allocator counters, variants via env."""
import os, sys, types, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
args = types.SimpleNamespace(no_graph=True, no_input_grad=False)
cfg, spec = bench.workload_cfg("city_train")
bench.load_peak()
det = os.environ.get("DET", "1") == "1"
r = bench.Runner(args, cfg, spec, rank, lr, world, deterministic=det)
if os.environ.get("NO_REDUCE") == "1":
    r.grad_buf = None
if os.environ.get("WARM_COMM") == "1" and world > 1:
    x = torch.ones(1 << 20, device=dev); dist.all_reduce(x); torch.cuda.synchronize()
import waldo_b200 as wb
wb.set_deterministic(det)
RED = []
if r.grad_buf is not None and os.environ.get("TIME_REDUCE") == "1":
    orig_reduce = r.grad_buf.reduce
    def timed_reduce():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = orig_reduce(); b.record(); RED.append((a, b)); return out
    r.grad_buf.reduce = timed_reduce
for _ in range(4):
    r.step(r.resident)
torch.cuda.synchronize()
marks = []
stats = []
for i in range(8):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
    t0 = time.perf_counter()
    r.step(r.resident)
    ms = torch.cuda.memory_stats(dev)
    stats.append((time.perf_counter() - t0, ms.get("num_device_alloc", 0), ms.get("num_device_free", 0), ms.get("num_alloc_retries", 0)))
e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
torch.cuda.synchronize()
if os.environ.get("STAGES") == "1":
    from waldo_b200 import functional as Fn
    Fn.PROFILE = {}
    m0 = torch.cuda.Event(enable_timing=True); m0.record()
    for _ in range(4):
        r.step(r.resident)
    m1 = torch.cuda.Event(enable_timing=True); m1.record()
    torch.cuda.synchronize()
    prof = Fn.PROFILE; Fn.PROFILE = None
    if rank == 0:
        print("   staged: step", round(m0.elapsed_time(m1) / 4, 2), {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 2) for k, v in prof.items()})
if rank == 0 and RED:
    print("   reduce ms:", " ".join(f"{a.elapsed_time(b):.2f}" for a, b in RED[-8:]))
if True:
    print("rank", rank, "det", det, "world", world, "no_reduce", os.environ.get("NO_REDUCE"), "gpu ms:", " ".join(f"{marks[i].elapsed_time(marks[i+1]):.1f}" for i in range(8)))
    print("   cpu ms:", " ".join(f"{s[0]*1e3:.1f}" for s in stats), "| device allocs:", [s[1] for s in stats], "frees:", [s[2] for s in stats], "retries:", stats[-1][3])
if world > 1:
    dist.barrier(); dist.destroy_process_group()
