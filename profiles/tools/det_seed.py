import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
import waldo_b200 as wb
from waldo_b200 import functional as Fn
args = types.SimpleNamespace(no_graph=True, no_input_grad=False)
cfg, spec = bench.workload_cfg("city_train")
bench.load_peak()
for seed in (1,):
  for DET in (False, True):
    r = bench.Runner(args, cfg, spec, seed, 0, 1, deterministic=DET)   # rank -> seed of the synthetic inputs
    wb.set_deterministic(DET)
    for _ in range(3):
        r.step(r.resident)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        r.step(r.resident)
    e1.record(); torch.cuda.synchronize()
    sc = Fn.LAST_DET_SCALE.cpu().tolist() if DET else None
    # one more step keeping the leaf gradients
    lv = {k: r.resident[k].detach().requires_grad_(True) for k in r.keys}
    occ, oa, ba, grid = wb.estimate_alpha_grid_occ(r.warper, lv["obj_alpha_raw"], r.om, r.bg, lv["obj_pose"], lv["bg_pose"], lv["occ_score"])
    out = wb.decode_output(r.warper, lv["input"], grid, occ, oa, ba, lv["cls"], r.ctx_ts, r.pred_ts, cfg.restrict_to_ctx)
    torch.autograd.backward([out[0], out[1], out[5]], [r.g_output, r.g_flow, r.g_raw])
    nan = {k: int((~torch.isfinite(v.grad)).sum()) for k, v in lv.items()}
    print("   outputs finite:", [bool(torch.isfinite(o).all()) if o is not None else None for o in out], "grid finite:", [bool(torch.isfinite(g).all()) for g in grid],
          "occ", bool(torch.isfinite(occ).all()), "upstream max", float(r.g_output.abs().max()), float(r.g_flow.abs().max()), float(r.g_raw.abs().max()))
    mx = {k: float(v.grad.abs().max()) for k, v in lv.items()}
    print(f"seed {seed} det={DET}: step {e0.elapsed_time(e1) / 4:.2f} ms, det_scale {sc}, NaNs {nan}, max|grad| {mx}", flush=True)
    wb.set_deterministic(False)
    del r, lv, out, grid
    torch.cuda.empty_cache()
