#!/usr/bin/env python
"""Benchmark of the WALDO warp+composite hot path (BASELINE.json metric: warped+composited frames/s, fwd+bwd, 512x1024).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl waldo|reference|reference-gpu]
                    [--workload city_train|city_rollout|kitti_rollout|nonrigid_train] [--deterministic]

One "step" = one pass of the hot path over one batch of synthetic input:
  control points -> TPS grids -> inverse warps -> occlusion matrix -> context alpha -> fused warp+composite (forward),
  then the whole backward down to input / obj_alpha / obj_pose / bg_pose / occ_score / cls gradients (city_train), or
  forward only (the *_rollout workloads).
Default workload = BASELINE.json configs[1]: Cityscapes shape 512x1024, batch 8 per GPU, 4 contexts -> 1 future frame,
fp32, fwd+bwd.  Under torchrun every rank runs the same per-GPU batch (weak scaling, no data-path collective; the
training workload adds the DDP-equivalent flat gradient all-reduce of a WIF-sized buffer, overlapped with the backward,
SURVEY.md section 8e).  The default line also carries, measured in the same process at every N:
  `deterministic`   the same step with order-independent (bit-reproducible) gradient accumulation -- the north star's contract;
  `other_workloads` BASELINE configs[3] / [2] / [4] (Cityscapes and KITTI rollouts, 256x256 training step);
and at N = 1: `gpu_stock_baseline` (the reference's own code on the same B200 with stock PyTorch kernels), `cpu_baseline`
(the reference's own code on the host cores, incl. the C1 forward split of BASELINE configs[0]).

`--impl reference` times the reference's own code (the byte-for-byte staged copy oracle/_ref; the oracle port only if that
is absent) on the host cores, one video per step; `--impl reference-gpu` the same code on CUDA.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WIF_GRAD_ELEMS = 14_160_000   # WIF parameters: the per-step DDP all-reduce payload of train_wif.sh


def workload_cfg(name):
    from waldo_b200 import workloads as wl
    return wl.workload(name)


def alg_bytes(cfg, B, Tc, Tp, backward):
    """Algorithmic HBM bytes of one step (SURVEY.md §8d / BASELINE.md §3), fp32."""
    from waldo_b200 import workloads as wl
    return wl.alg_bytes(cfg, B, Tc, Tp, backward)


def step_bytes(cfg, B, Tc, Tp, backward, storage="f32"):
    """alg_bytes for a storage type: with bf16 storage the big HD streams (input, raw_output, output) are 2-byte elements; alpha
    (written once per context frame, gathered by the layer kernel), flow and score stay fp32."""
    from waldo_b200 import workloads as wl
    if storage != "bf16":
        return wl.alg_bytes(cfg, B, Tc, Tp, backward)
    assert not backward
    fwd2, _ = wl.alg_bytes(cfg, B, Tc, Tp, False, elem=2)
    Hd, Wd = cfg.hd_shape
    L = cfg.num_obj + 1
    return fwd2 + Hd * Wd * 2 * (B * Tc * Tp * (2 + 2 + 1) + B * Tc * L + B * Tc * Tp * L), 0   # flow w + r, score r, alpha w + gathered: fp32


def kernel_bytes(cfg, B, Tc, Tp):
    """Algorithmic bytes per launch of the four HD kernels (DESIGN.md "kernels"), fp32.  pair = one (b, tc, tp)."""
    Hd, Wd = cfg.hd_shape
    px, s = Hd * Wd, 4
    C, L, Nl = 3 + cfg.num_lyt, cfg.num_obj + 1, cfg.num_lyt
    pairs, frames = B * Tc * Tp, B * Tp
    return {
        # read the HD layout logits, write the context alpha stack (low-res inputs are L2-resident)
        "k_alpha_prep": px * s * (B * Tc * (Nl + L)),
        # read d alpha (scatter-accumulated), re-read the layout logits, write d layout logits
        "k_alpha_prep_bwd": px * s * (B * Tc * (L + 2 * Nl)),
        # gather-read the context frame, read flow + score; write the warped channels, the fused output and its norm
        "k_gather_fwd": px * s * (pairs * (C + 3 + C) + frames * (C + 1 + 1)),
        # gather-read the context alpha stack; write alpha_ctx, flow, score
        "k_layers_fwd": px * s * (pairs * (L + L + 2 + 1)),
        # read d raw_output (image channels), re-read the context frame, flow, score; read d output, output, norm;
        # write d input once per context frame and the 3-float glue
        "k_gather_bwd": px * s * (pairs * (C + C + 3 + 3) + frames * (2 * (C + 1) + 1) + B * Tc * C),
        # read d raw_output (alpha channels), re-read the alpha stack, glue; write d alpha once per context frame
        "k_layers_bwd": px * s * (pairs * (L + L + 3) + B * Tc * L),
    }


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an in-process NVML poll every 50 ms on a daemon thread
    (pynvml; two light driver queries per poll).  A polling `nvidia-smi -lms` child process, used before, re-queries the
    whole device state on every poll and -- in the first process on a fresh box -- stalled kernel launches for 3 to 200 ms per
    poll (measured: profiles/r2/r2_notes.md); it remains the fallback when pynvml is missing.
    summary() keeps the samples taken inside [mark_begin(), mark_end()]."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None
        self.stop, self.max_mhz, self.source = False, None, None
        if os.environ.get("WALDO_NO_SAMPLER"):   # experiments only
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            threading.Thread(target=self._poll_nvml, daemon=True).start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop:
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((time.time(), [str(mhz), str(self.max_mhz)] + ["Active" if mask & bit else "Not Active" for bit, _ in self.REASONS]))
            except Exception:
                pass
            time.sleep(0.05)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_ready(self, timeout=5.0):
        t = time.time()
        while self.source and not self.rows and time.time() - t < timeout:
            time.sleep(0.05)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def close(self):
        self.stop = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = [r for t, r in self.rows if self.t0 is not None and self.t0 - 0.05 <= t <= (self.t1 or t) + 0.25 and len(r) >= 6]
        if not rows and self.rows:   # timed region shorter than the polling period: take the sample nearest to it
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - (self.t0 or tr[0])))[1]]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def make_inputs(cfg, B, T, Tc, seed):
    from waldo_b200 import workloads as wl
    return wl.synth_inputs(cfg, B, T, Tc, seed=seed)


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_loss(out):
    return out[0].abs().mean() + out[1].abs().mean()


def oracle_step(wo, st, d, backward):
    lv = {k: d[k].clone().requires_grad_(backward) for k in ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")}
    occ, oa, ba, grid = wo.estimate_alpha_grid_occ(st, lv["obj_alpha_raw"], lv["obj_pose"], lv["bg_pose"], lv["occ_score"])
    with torch.set_grad_enabled(backward):
        out = wo.decode_output(st, lv["input"], grid, occ, oa, ba, lv["cls"], d["ctx_ts"], d["pred_ts"])
        loss = _ref_loss(out)
    if backward:
        loss.backward()
    return float(loss)


def reference_runner(cfg, device):
    """(step(d, backward), kind, what): the UNMODIFIED reference's own code (oracle/ref_runner.py: /root/reference here, its
    staged byte-for-byte copy oracle/_ref on the GPU box) when present, else the oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import waldo_oracle as wo
    try:
        import ref_runner
        if ref_runner.available():
            rp = ref_runner.RefPath(cfg, device)

            def step(d, backward):
                # the stable tie rule only matters for parity; the timed arm runs the reference exactly as shipped
                out, _, _, _, loss = rp.chain(d, backward=backward, stable=False, loss_fn=_ref_loss)
                return float(loss) if loss is not None else float(_ref_loss(out))
            where = "staged copy oracle/_ref" if ref_runner.ref_loader.is_staged_copy() else ref_runner.ref_loader.REF_ROOT
            return step, "reference", f"the reference's own code ({where}: Warper.forward + compute_occ + LVD.forward decode_output)"
    except Exception as e:   # pragma: no cover - the port below always exists
        print(f"[bench] reference not loadable ({e}); timing the oracle port", file=sys.stderr)
    st = wo.make_state(cfg)
    return (lambda d, backward: oracle_step(wo, st, d, backward)), "port", "oracle/waldo_oracle.py (port of the reference)"


def cpu_baseline(cfg, spec, steps=1, warmup=0):
    """The reference on the host cores, on a bounded sample of the workload: ONE video (B=1) per step."""
    import warnings
    warnings.filterwarnings("ignore")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, what = reference_runner(cfg, "cpu")
    d = make_inputs(cfg, 1, spec["T"], spec["Tc"], seed=0)
    frames = spec["T"] - spec["Tc"]
    for _ in range(warmup):
        step(d, spec["backward"])
    t0 = time.perf_counter()
    for _ in range(steps):
        step(d, spec["backward"])
    dt = (time.perf_counter() - t0) / steps
    return {"value": frames / dt, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"B=1 of the workload ({frames} future frame(s), {'fwd+bwd' if spec['backward'] else 'fwd'}), "
                      f"{steps} step(s) of {dt:.1f} s, torch {torch.__version__} CPU, {what}"}, dt


def cpu_c1_split(cfg):
    """BASELINE configs[0] (C1): the reference's forward on the host cores, B=1, 4 contexts -> 1 frame, split per stage
    (SURVEY.md §8d): Warper.forward + compute_occ / decode_output / WIF.forward (UNet + fuse tail)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_runner
        if not ref_runner.available():
            return None
        rp = ref_runner.RefPath(cfg, "cpu")
        d = make_inputs(cfg, 1, 5, 4, seed=0)
        ns = ref_runner.ref_loader.load()
        wif = ns.wif.WIF(ref_runner.ref_opt(cfg)).eval()
        with torch.no_grad():
            t0 = time.perf_counter()
            occ, oa, ba, grid = rp.stage_a(d["obj_alpha_raw"], d["obj_pose"], d["bg_pose"], d["occ_score"], stable=False)
            t1 = time.perf_counter()
            out = rp.decode(d["input"], grid, occ, oa, ba, d["cls"], d["ctx_ts"], d["pred_ts"])
            t2 = time.perf_counter()
            wif(out[5])
            t3 = time.perf_counter()
        return {"config": "C1: Cityscapes-shape forward on CPU, B=1, 4 contexts -> 1 frame (BASELINE configs[0])",
                "warper_forward_s": t1 - t0, "decode_output_s": t2 - t1, "wif_forward_s": t3 - t2,
                "frames_per_s_hot_path": 1.0 / (t2 - t0), "frames_per_s_with_wif": 1.0 / (t3 - t0), "cores": os.cpu_count()}
    except Exception as e:
        return {"error": str(e)[:200]}


def gpu_stock_baseline(cfg, spec, dev, steps=3, batch=1):
    """The honest GPU comparator ('the bar is stock PyTorch-on-B200 running the reference code'): the
    reference's own modules moved to CUDA, ~170 ATen launches per decode, same workload at a bounded batch."""
    step, kind, what = reference_runner(cfg, dev)
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_inputs(cfg, batch, spec["T"], spec["Tc"], seed=0).items()}
    frames = batch * (spec["T"] - spec["Tc"])
    for _ in range(2):
        step(d, spec["backward"])
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(d, spec["backward"])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    peak = torch.cuda.max_memory_allocated(dev) / 1e9
    return {"value": frames / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "kind": kind, "device": "cuda (stock PyTorch ATen kernels)",
            "sample": f"B={batch} of the workload per step ({'fwd+bwd' if spec['backward'] else 'fwd'}), {steps} timed steps after 2 warm-ups, "
                      f"peak {peak:.1f} GB; {what}"}


def run_reference(args, cfg, spec, rank, world):
    if rank != 0:
        return
    if args.impl == "reference-gpu":
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        g = gpu_stock_baseline(cfg, spec, dev, steps=max(1, args.steps), batch=args.batch or 1)
        line = {"impl": "reference-gpu", "metric": "warped+composited frames/s", "value": g["value"], "unit": "frames/s", "n_gpus": 1,
                "steps": max(1, args.steps), "warmup": 2, "ms_per_step": g["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": spec["label"], "sample": g["sample"]},
                "gpu_stock_baseline": g, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    budget_s = 170.0
    base, dt = cpu_baseline(cfg, spec, steps=1, warmup=0)   # doubles as the first warm-up step
    steps = max(1, min(args.steps, int(budget_s / dt) - min(args.warmup, 1)))
    warm = 0 if steps + 1 > budget_s / dt else min(args.warmup, 1)
    base, dt = cpu_baseline(cfg, spec, steps=steps, warmup=warm)
    line = {"impl": "reference", "metric": "warped+composited frames/s", "value": base["value"], "unit": "frames/s",
            "n_gpus": world, "steps": steps, "steps_requested": args.steps, "warmup": warm + 1, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": spec["label"], "sample": "one video (B=1) per step on the host cores"},
            "cpu_baseline": base, "gpu_launches": 0,
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class Runner:
    """One workload set up on one GPU: resident inputs, pinned host copies, the step function and the two timers."""

    def __init__(self, args, cfg, spec, rank, local_rank, world, deterministic=False, storage="f32"):
        import torch.distributed as dist
        import waldo_b200 as wb
        from waldo_b200 import _lib, functional as Fn, sharding, workloads as wl
        self.wb, self.Fn, self.sharding, self.dist, self.lib = wb, Fn, sharding, dist, _lib.load()
        self.args, self.cfg, self.spec, self.rank, self.world = args, cfg, spec, rank, world
        self.dev = dev = torch.device("cuda", local_rank)
        self.deterministic = deterministic
        if storage == "bf16" and spec["backward"]:
            raise SystemExit("--storage bf16 is the forward / inference variant: use it with a *_rollout workload")
        self.storage = storage
        self.st_dtype = torch.bfloat16 if storage == "bf16" else torch.float32
        B, T, Tc, backward = spec["B"], spec["T"], spec["Tc"], spec["backward"]
        self.B, self.T, self.Tc, self.Tp, self.backward = B, T, Tc, T - Tc, backward
        opt = wl.make_opt(cfg)
        self.warper = wb.Warper(opt).to(dev)
        om, bg = wb.alpha_masks(opt)
        self.om, self.bg = om.to(dev), bg.to(dev)
        host = make_inputs(cfg, B, T, Tc, seed=rank)
        self.keys = ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")
        self.small = ("obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")
        # what the dataset really holds (data/base_dataset.py:173-183, :355-372): 8-bit RGB and an 8-bit label map; the
        # fp32 `input` (normalised RGB | +-5 one-hot logits) is expanded from them on the device by waldo_pack_input
        rgb8 = ((host["input"][:, :, :3] + 1) * 127.5).round().clamp(0, 255).to(torch.uint8)
        lab8 = host["input"][:, :, 3:].argmax(dim=2).to(torch.uint8)
        self.pinned = {k: host[k].pin_memory() for k in self.small}
        self.pinned["rgb"], self.pinned["label"] = rgb8.pin_memory(), lab8.pin_memory()
        self.ctx_ts, self.pred_ts = host["ctx_ts"].contiguous().to(dev), host["pred_ts"].to(dev)
        self.resident = {k: self.pinned[k].to(dev) for k in self.small}
        self.resident["input"] = wb.pack_input(self.pinned["rgb"].to(dev), self.pinned["label"].to(dev), cfg.num_lyt, dtype=self.st_dtype)
        del host
        self.grad_buf = None
        if world > 1 and backward and not args.no_exchange:   # the trainable net's gradients (WIF-sized), exchanged as DDP would: one flat all-reduce per step
            wif_like = torch.nn.Parameter(torch.zeros(WIF_GRAD_ELEMS, device=dev))
            wif_like.grad = torch.zeros_like(wif_like)
            # (if init_nccl capped the default group's CTAs: deterministic mode exchanges before the backward starts, on a
            #  communicator of NCCL's own width)
            self.grad_buf = sharding.FlatGradReducer([wif_like], group=sharding.full_width_group() if deterministic else None)
        self.loss_host = torch.zeros(1).pin_memory()
        # upstream gradients as a downstream consumer (WIF / losses) would supply them: fixed seeded tensors, so that no
        # loss kernel of torch sits inside the timed region and the backward reads d output, d flow and d raw_output
        C, L = 3 + cfg.num_lyt, cfg.num_obj + 1
        Hd, Wd = cfg.hd_shape
        if backward:
            gen = torch.Generator(device=dev).manual_seed(1 + rank)
            self.g_output = torch.randn(B, self.Tp, C, Hd, Wd, device=dev, generator=gen)
            self.g_flow = torch.randn(B, Tc, self.Tp, 2, Hd, Wd, device=dev, generator=gen)
            self.g_raw = torch.randn(B, Tc, self.Tp, C + L, Hd, Wd, device=dev, generator=gen)
        self.graphed = wb.GraphedDecode(self.warper, self.om, self.bg, cfg.restrict_to_ctx, max_graphs=8) if (not backward and not args.no_graph) else None
        self.feeder = None

    def step(self, src):
        wb, cfg = self.wb, self.cfg
        if self.graphed is not None:   # inference: the whole chain replayed from one CUDA graph per set of input buffers
            out = self.graphed(src["input"], src["obj_alpha_raw"], src["obj_pose"], src["bg_pose"], src["occ_score"], src["cls"], self.ctx_ts, self.pred_ts)
            return out[0][:, :, :3].mean()
        backward = self.backward
        lv = {k: src[k].detach().requires_grad_(backward and not (self.args.no_input_grad and k == "input")) for k in self.keys}
        with torch.set_grad_enabled(backward):
            occ, oa, ba, grid = wb.estimate_alpha_grid_occ(self.warper, lv["obj_alpha_raw"], self.om, self.bg, lv["obj_pose"], lv["bg_pose"], lv["occ_score"])
            out = wb.decode_output(self.warper, lv["input"], grid, occ, oa, ba, lv["cls"], self.ctx_ts, self.pred_ts, cfg.restrict_to_ctx)
        if backward:
            # the trainable net's gradient exchange rides NCCL's own stream while this path's backward runs (what DDP's
            # bucketed all-reduce does with the rest of a backward pass); the step ends when both are done
            # (deterministic mode: the exchange runs to completion first.  Concurrent with the 64-bit reductions of the
            #  deterministic backward kernels the step took 3x longer on the B200, profiles/r2/r2_notes.md)
            overlap = self.grad_buf is not None and not self.deterministic
            if self.grad_buf is not None and not overlap:
                self.grad_buf.reduce()
            work = self.grad_buf.reduce_async() if overlap else None
            torch.autograd.backward([out[0], out[1], out[5]], [self.g_output, self.g_flow, self.g_raw])
            if work is not None:
                self.grad_buf.finish(work)
        return out[0].detach()[:, :, :3].mean()   # the step's metric (mean predicted RGB), read back in the e2e leg

    def e2e_step(self):
        # the public feeding API: the next batch is copied from pinned host memory on a side stream while this one is
        # processed; every step's inputs cross PCIe inside the timed region and the step's metric is read back
        metric = self.step(next(self.feeder))
        self.loss_host.copy_(metric.reshape(1), non_blocking=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        import gc
        gc.collect()
        gc.disable()   # no collector pause inside the timed region (thousands of short-lived tensor wrappers per step)
        try:
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            marks = []
            for _ in range(steps):
                fn()
                if os.environ.get("WALDO_STEP_TIMES"):   # experiments only: per-step device times
                    m = torch.cuda.Event(enable_timing=True); m.record(); marks.append(m)
            e1.record()
            self.barrier()
            if marks:
                ts = [e0.elapsed_time(m) for m in marks]
                print("[step times]", " ".join(f"{b - a:.2f}" for a, b in zip([0.0] + ts[:-1], ts)), file=sys.stderr, flush=True)
        finally:
            gc.enable()
        mine = e0.elapsed_time(e1)
        self.last_per_rank_ms = [v / steps for v in self.sharding.gather_over_ranks(mine, device=self.dev)]
        return self.sharding.max_over_ranks([mine], device=self.dev)[0] / steps

    def measure(self, steps, warmup, profile=True):
        """(ms per step, launches of this library in the timed region, per-stage event times)."""
        wb, Fn = self.wb, self.Fn
        wb.set_deterministic(self.deterministic)
        try:
            for _ in range(max(warmup, 3)):
                self.step(self.resident)
            if self.graphed is None and profile:
                Fn.PROFILE = {}
            launches0 = self.lib.waldo_launch_count()
            ms_step = self.timed(lambda: self.step(self.resident), steps)
            launches = self.lib.waldo_launch_count() - launches0
            prof = {k: sum(a.elapsed_time(b) for a, b in v) / max(len(v), 1) for k, v in (Fn.PROFILE or {}).items()}
        finally:
            Fn.PROFILE = None
            wb.set_deterministic(False)
        return ms_step, launches, prof

    def frames(self):
        return self.world * self.B * self.Tp

    def summary(self, ms):
        fwd_b, bwd_b = step_bytes(self.cfg, self.B, self.Tc, self.Tp, self.backward, self.storage)
        return {"value": self.frames() / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                "hbm_frac_step": ((fwd_b + bwd_b) / (ms * 1e-3) / 1e9) / PEAK["gbs"]}


def loss_epilogues(dev, steps=20):
    """f-2 kernels at the LVD-training shape (B=8, T=5, 17 layers, 128x256 low-res lattice): layer entropy + fg_mask forward and
    backward (HBM streaming: (L + 2) + (2L + 2) floats per pixel) and the 23-tap Gaussian blur of a 3-plane map, fwd + bwd."""
    import waldo_b200 as wb
    B, T, L, H, W = 8, 5, 17, 128, 256
    gen = torch.Generator(device=dev).manual_seed(3)
    alpha = torch.tanh(torch.randn(B, T, L, H, W, device=dev, generator=gen)).requires_grad_(True)
    maps = torch.randn(B, T, 3, H, W, device=dev, generator=gen).requires_grad_(True)
    g1 = torch.randn(B, T, 1, H, W, device=dev, generator=gen)
    g3 = torch.randn(B, T, 3, H, W, device=dev, generator=gen)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    def ent():
        alpha.grad = None
        e, f = wb.layer_entropy(alpha)
        torch.autograd.backward([e, f], [g1, g1])

    def blr():
        maps.grad = None
        wb.blur(maps, 3.0, 23).backward(g3)
    ms_e, ms_b = timed(ent), timed(blr)
    px = B * T * H * W * 4
    be, bb = px * ((L + 2) + (2 * L + 2)), px * 3 * 4
    return {"shape": f"B={B} T={T} L={L} {H}x{W}", "timing": "autograd fwd+bwd, launch overhead of the two small kernels included",
            "layer_entropy_fwd_bwd": {"ms": ms_e, "alg_bytes": be, "frac": be / (ms_e * 1e-3) / 1e9 / PEAK["gbs"]},
            "blur23_fwd_bwd": {"ms": ms_b, "alg_bytes": bb, "frac": bb / (ms_b * 1e-3) / 1e9 / PEAK["gbs"]}}


def wif_to_emb_leg(dev, steps=10):
    """f-1 first layer at the WIF shape of the default workload: raw_output (B=8, Tc=4, Tp=1, 40 channels, 512x1024) -> 16 planes
    per image; HBM streaming: (40 + 16) floats per pixel and image."""
    import waldo_b200 as wb
    B, Tc, Tp, Cin, Cout, H, W = 8, 4, 1, 40, 16, 512, 1024
    gen = torch.Generator(device=dev).manual_seed(4)
    raw = torch.randn(B, Tc, Tp, Cin, H, W, device=dev, generator=gen)
    wgt = torch.randn(Cout, Cin, 3, 3, device=dev, generator=gen) * 0.05
    with torch.no_grad():
        for _ in range(3):
            wb.wif_to_emb(raw, wgt)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            wb.wif_to_emb(raw, wgt)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        # the same layer as the reference runs it: permute + reshape copy, then cuDNN (torch default: TF32 allowed)
        x = raw.permute(0, 2, 1, 3, 4, 5).reshape(B * Tp * Tc, Cin, H, W)
        for _ in range(3):
            torch.nn.functional.conv2d(raw.permute(0, 2, 1, 3, 4, 5).reshape(B * Tp * Tc, Cin, H, W), wgt, padding=1)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(steps):
            torch.nn.functional.conv2d(raw.permute(0, 2, 1, 3, 4, 5).reshape(B * Tp * Tc, Cin, H, W), wgt, padding=1)
        e1.record()
        torch.cuda.synchronize(dev)
        ms_ref = e0.elapsed_time(e1) / steps
        del x
    # training: forward + backward-data (d raw_output) + weight gradient, ours and stock torch
    raw_g, wgt_g = raw.detach().requires_grad_(True), wgt.detach().requires_grad_(True)
    gy = torch.randn(B * Tc * Tp, Cout, H, W, device=dev, generator=gen)

    def train_step(fn):
        raw_g.grad = None
        wgt_g.grad = None
        fn().backward(gy)

    def timed(fn, n=5):
        for _ in range(2):
            train_step(fn)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            train_step(fn)
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / n
    ms_train = timed(lambda: wb.wif_to_emb(raw_g, wgt_g))
    ms_train_ref = timed(lambda: torch.nn.functional.conv2d(raw_g.permute(0, 2, 1, 3, 4, 5).reshape(B * Tp * Tc, Cin, H, W), wgt_g, padding=1))
    del raw_g, gy
    nbytes = B * Tc * Tp * H * W * 4 * (Cin + Cout)
    return {"shape": f"raw_output (B={B}, Tc={Tc}, Tp={Tp}, {Cin}, {H}, {W}) -> ({B * Tc * Tp}, {Cout}, {H}, {W})", "ms": ms, "alg_bytes": nbytes,
            "fwd_bwd_ms": ms_train, "stock_torch_fwd_bwd_ms": ms_train_ref,
            "frac": nbytes / (ms * 1e-3) / 1e9 / PEAK["gbs"], "tflops": 2 * 9 * Cin * Cout * B * Tc * Tp * H * W / (ms * 1e-3) / 1e12,
            "stock_torch_ms": ms_ref, "stock_torch": "permute + reshape copy + F.conv2d (cuDNN, allow_tf32 default)"}


PEAK = {"gbs": 6650.0, "src": "B200_PROFILING.md fallback"}


def load_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        PEAK["gbs"], PEAK["src"] = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="waldo", choices=["waldo", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="city_train")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE workloads and the deterministic-mode leg")
    ap.add_argument("--no-input-grad", action="store_true", help="experiment: do not request d input")
    ap.add_argument("--no-exchange", action="store_true", help="experiment (N > 1): no gradient all-reduce, so that ms_per_step_per_rank shows the GPUs' own spread")
    ap.add_argument("--no-graph", action="store_true", help="inference workloads: eager launches instead of CUDA-graph replay")
    ap.add_argument("--storage", default="f32", choices=["f32", "bf16"],
                    help="bf16: input / alpha / raw_output / output stored as bf16, fp32 arithmetic (forward / inference workloads only)")
    ap.add_argument("--deterministic", action="store_true",
                    help="backward with 64-bit fixed-point accumulation of the scatter targets (bit-identical gradients run to run)")
    args = ap.parse_args()

    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    cfg, spec = workload_cfg(args.workload)
    if args.batch:
        spec["B"] = args.batch
    if args.impl in ("reference", "reference-gpu"):
        run_reference(args, cfg, spec, rank, world)
        return

    import torch.distributed as dist
    import waldo_b200 as wb
    from waldo_b200 import functional as Fn

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        from waldo_b200 import sharding
        sharding.init_nccl(dev)
    load_peak()
    r = Runner(args, cfg, spec, rank, local_rank, world, deterministic=args.deterministic, storage=args.storage)
    B, T, Tc, Tp, backward = r.B, r.T, r.Tc, r.Tp, r.backward
    Hd, Wd = cfg.hd_shape

    # The sampler's nvidia-smi process takes driver locks while it initialises (seconds on a fresh box: the first process to
    # open the driver pays for it) and would stall kernel launches in whatever runs meanwhile: wait until it delivers samples,
    # THEN warm up, then time.
    clocks = ClockSampler(local_rank)
    clocks.wait_ready(timeout=60.0)
    for _ in range(max(args.warmup, 3) + 3):
        r.step(r.resident)
    torch.cuda.synchronize()
    # ---- device-resident timing.  Two timed regions of K steps each, back to back: the plain one gives the headline (the calls
    # a user makes: one C-ABI call per entry point), the instrumented one issues the same kernels stage by stage with a CUDA-event
    # pair around each HD kernel (the roofline figures); both numbers are reported.
    clocks.mark_begin()
    ms_step, launches, _ = r.measure(args.steps, 0, profile=False)
    per_rank_ms = list(r.last_per_rank_ms)   # the headline region's device time on every rank (ms_step is their max)
    ms_instr, prof = ms_step, {}
    if r.graphed is None:
        ms_instr, _, prof = r.measure(args.steps, 0)
    clocks.mark_end()
    kernel_timing = "CUDA events around each kernel in a second timed region of the same K steps, right after the plain one (%.3f ms/step instrumented)" % ms_instr
    if r.graphed is not None:
        # a replayed graph cannot be bracketed kernel by kernel: time the same kernels in an extra eager pass
        with torch.no_grad():
            launches1 = r.lib.waldo_launch_count()
            occ_, oa_, ba_, grid_ = wb.estimate_alpha_grid_occ(r.warper, r.resident["obj_alpha_raw"], r.om, r.bg, r.resident["obj_pose"], r.resident["bg_pose"], r.resident["occ_score"])
            wb.decode_output(r.warper, r.resident["input"], grid_, occ_, oa_, ba_, r.resident["cls"], r.ctx_ts, r.pred_ts, cfg.restrict_to_ctx)
            per_step = r.lib.waldo_launch_count() - launches1   # kernels of this library in one chain = kernel nodes of the graph
            Fn.PROFILE = {}
            for _ in range(args.steps):
                wb.decode_output(r.warper, r.resident["input"], grid_, occ_, oa_, ba_, r.resident["cls"], r.ctx_ts, r.pred_ts, cfg.restrict_to_ctx)
            torch.cuda.synchronize()
            del occ_, oa_, ba_, grid_
        launches = args.steps * per_step
        kernel_timing = "extra eager pass after the timed region (the timed region replays one CUDA graph per step)"
        prof = {k: sum(a.elapsed_time(b) for a, b in v) / max(len(v), 1) for k, v in Fn.PROFILE.items()}
        Fn.PROFILE = None
    # ---- end to end: pinned host -> device copies and the loss read-back inside the timed region
    e2e = None
    if not args.no_e2e:
        nbytes = lambda d: sum(t.numel() * t.element_size() for t in d.values())

        def endless(batch):
            while True:
                yield batch
        # (a) 8-bit frames + labels over PCIe, expanded on the device (the shipped input pipeline)
        r.feeder = wb.DevicePrefetcher(endless(r.pinned), dev, num_lyt=cfg.num_lyt, input_dtype=r.st_dtype)
        for _ in range(3):   # every prefetch slot has been seen once (graph replay captures one graph per slot)
            r.e2e_step()
        ms_e2e = r.timed(r.e2e_step, args.steps)
        e2e = {"value": world * B * Tp / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": nbytes(r.pinned), "d2h_bytes_per_step": 4,
               "api": "waldo_b200.DevicePrefetcher (8-bit RGB + label map, pack_input on device, copy overlapped with the previous step)"}
        if args.storage == "f32":
            # (b) for comparison: the fp32 `input` tensor itself shipped every step (what the reference's to_cuda moves)
            r.feeder = None
            pinned32 = {k: r.pinned[k] for k in r.small}
            pinned32["input"] = r.resident["input"].float().cpu().pin_memory()
            r.feeder = wb.DevicePrefetcher(endless(pinned32), dev)
            for _ in range(3):
                r.e2e_step()
            ms32 = r.timed(r.e2e_step, args.steps)
            e2e["fp32_input"] = {"value": world * B * Tp / (ms32 * 1e-3), "ms_per_step": ms32, "h2d_bytes_per_step": nbytes(pinned32)}
            r.feeder = None
            del pinned32

    # ---- the north star's deterministic-gradient contract and the other BASELINE configurations, same process, every N
    det_leg, others = None, {}
    if not args.no_others and args.workload == "city_train" and not args.batch:
        if backward and not args.deterministic:
            rd = Runner(args, cfg, spec, rank, local_rank, world, deterministic=True)
            rd.resident, rd.pinned = r.resident, r.pinned      # same inputs, no second copy
            ms_d, _, _ = rd.measure(max(3, args.steps // 2), 3, profile=False)
            det_leg = dict(rd.summary(ms_d), mode="order-independent 64-bit fixed-point accumulation of every scatter target; bit-identical gradients run to run",
                           slowdown_vs_default=ms_d / ms_step)
            del rd
        del r.g_output, r.g_flow, r.g_raw
        keep_input = r.resident.pop("input")
        del keep_input
        torch.cuda.empty_cache()
        for name, sto in (("city_rollout", "f32"), ("kitti_rollout", "f32"), ("nonrigid_train", "f32"),
                          ("city_rollout", "bf16"), ("kitti_rollout", "bf16")):
            try:
                c2, s2 = workload_cfg(name)
                r2 = Runner(args, c2, s2, rank, local_rank, world, storage=sto)
                ms2, _, _ = r2.measure(max(5, args.steps), 3, profile=False)   # (steps of 1.4 - 3 ms: the full count, for a stable mean)
                if sto == "bf16":
                    name, s2["label"] = name + "_bf16", s2["label"] + ", bf16 storage / fp32 arithmetic (tolerance: tests/parity.py TOL_BF16)"
                others[name] = dict(r2.summary(ms2), workload=s2["label"],
                                    launch="one CUDA graph replay per step" if r2.graphed is not None else "eager kernel launches (autograd)")
                del r2
                torch.cuda.empty_cache()
            except Exception as e:   # never lose the headline line to a side measurement
                others[name] = {"error": str(e)[:200]}
        if world == 1:
            try:
                others["loss_epilogues"] = loss_epilogues(dev)
            except Exception as e:
                others["loss_epilogues"] = {"error": str(e)[:200]}
            try:
                torch.cuda.empty_cache()
                others["wif_to_emb"] = wif_to_emb_leg(dev)
            except Exception as e:
                others["wif_to_emb"] = {"error": str(e)[:200]}

    clocks.close()
    if rank == 0:
        frames = world * B * Tp
        fwd_b, bwd_b = step_bytes(cfg, B, Tc, Tp, backward, args.storage)
        kb = kernel_bytes(cfg, B, Tc, Tp)
        if args.storage == "bf16":   # per-kernel figures: 2-byte elements except flow / score (see step_bytes)
            px = Hd * Wd
            kb = {k: v // 2 for k, v in kb.items()}
            kb["k_gather_fwd"] += px * 2 * B * Tc * Tp * 3
            kb["k_layers_fwd"] += px * 2 * B * Tc * Tp * 3
        peak, peak_src = PEAK["gbs"], PEAK["src"]
        traffic = {}
        if args.workload == "city_train" and not args.batch:   # the ncu --set full capture in profiles/ is of this workload
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            except Exception:
                pass
        kern = {"k_gather_fwd": (prof.get("decode_fwd:gather", 0.0), kb["k_gather_fwd"]),
                "k_layers_fwd": (prof.get("decode_fwd:layers", 0.0), kb["k_layers_fwd"]),
                "k_alpha_prep": (prof.get("decode_fwd:alpha_prep", 0.0), kb["k_alpha_prep"])}
        if backward:
            kern["k_gather_bwd"] = (prof.get("decode_bwd:gather", 0.0), kb["k_gather_bwd"])
            kern["k_layers_bwd"] = (prof.get("decode_bwd:layers", 0.0), kb["k_layers_bwd"])
            kern["k_alpha_prep_bwd"] = (prof.get("decode_bwd:alpha_prep", 0.0), kb["k_alpha_prep_bwd"])
        dom = max(kern, key=lambda k: kern[k][0])
        dms, dbytes = kern[dom]
        achieved = dbytes / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
        line = {
            "metric": "warped+composited frames/s", "value": frames / (ms_step * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.storage == "f32" else "f32 arithmetic on bf16 storage", "data": "synthetic",
            "config": {"workload": spec["label"], "per_gpu_batch": B, "contexts": Tc, "future_frames": Tp,
                       "layers": cfg.num_obj + 1, "channels": 3 + cfg.num_lyt, "l2": "working set per step far larger than the 126 MB L2 (input %.2f GB), no flush needed" % (B * T * (3 + cfg.num_lyt) * Hd * Wd * 4 / 1e9),
                       "input": "8-bit RGB + label map expanded to %s on the device (pack_input)" % ("fp32" if args.storage == "f32" else "bf16"),
                       "launch": "one CUDA graph replay per step" if r.graphed is not None else "eager kernel launches (autograd)",
                       "gradient_accumulation": ("64-bit fixed point (deterministic)" if args.deterministic else "fp32 red.global (default)") if backward else None,
                       "parallelism": f"dp{world} batch-sharded" + (", flat fp32 all-reduce of 56.6 MB per step on NCCL's stream, overlapped with this path's backward" if r.grad_buf is not None else "")},
            "clocks": clocks.summary(), "gpu_launches": launches,
            "hbm_frac_step": ((fwd_b + bwd_b) / (ms_step * 1e-3) / 1e9) / peak,
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic.get(dom), "ms_per_launch": dms, "alg_bytes_per_launch": dbytes, "peak_source": peak_src,
                         "other": {k: {"ms_per_launch": v[0], "alg_bytes_per_launch": v[1],
                                       "frac": (v[1] / (v[0] * 1e-3) / 1e9 / peak) if v[0] > 0 else None} for k, v in kern.items() if k != dom},
                         "stages_ms": {k: round(v, 4) for k, v in prof.items()}, "kernel_timing": kernel_timing,
                         "ms_per_step_instrumented": ms_instr},
        }
        if world > 1:
            line["ms_per_step_per_rank"] = [round(v, 4) for v in per_rank_ms]
        if e2e:
            line["e2e"] = e2e
        if det_leg:
            line["deterministic"] = det_leg
        if others:
            line["other_workloads"] = others
        if world == 1 and not args.no_cpu_baseline:
            torch.cuda.empty_cache()
            try:
                line["gpu_stock_baseline"] = gpu_stock_baseline(cfg, spec, dev, steps=3, batch=1)
            except Exception as e:   # never lose the line to the comparator
                line["gpu_stock_baseline"] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()
            line["cpu_baseline"], _ = cpu_baseline(cfg, spec, steps=1, warmup=0)
            if args.workload == "city_train":
                line["cpu_baseline"]["c1_forward_split"] = cpu_c1_split(cfg)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
