"""Batch sharding of the hot path across the GPUs of one box: one process per GPU, videos are
independent units, no data-path collective.  The only exchange is what the reference's DDP wrapper does for the net
being trained -- a gradient all-reduce (SUM / world) per step, tools/engine.py:46-49 -- plus the 1-element NaN-flag
gather of models/synthesizer.py:619 and, for benchmarking, a max over ranks of the device time.

Backend-agnostic host logic (NCCL on the GPU box, gloo in the CPU tests); it never touches the kernels.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def world():
    """(rank, world_size); (0, 1) when no process group is initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_videos(num_videos: int, rank: int, world_size: int, drop_last: bool = True) -> List[int]:
    """Indices of the videos rank `rank` processes: rank, rank + world, ... (what DistributedSampler(shuffle=False)
    hands out, tools/engine.py:61-65).  With drop_last every rank gets floor(num_videos / world) videos (the reference's
    loader uses drop_last=True and batch_size // world_size); otherwise the tail is spread over the first ranks."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    per = num_videos // world_size
    n = per * world_size if drop_last else num_videos
    return list(range(rank, n, world_size))


def per_rank_batch(global_batch: int, world_size: int) -> int:
    """tools/engine.py:63: batch_size // world_size."""
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} is not divisible by the world size {world_size}")
    return global_batch // world_size


EXCHANGE_CTAS = 0   # 0 = NCCL's own choice
_CAPPED = False


def init_nccl(device: torch.device, exchange_ctas: int = EXCHANGE_CTAS) -> None:
    """init_process_group("nccl"), optionally with the collectives' footprint capped at `exchange_ctas` CTAs per GPU
    (ncclConfig maxCTAs; NCCL_MAX_CTAS in the environment wins).

    The idea: this path's backward kernels are bound by instruction issue, so every SM a concurrent all-reduce occupies is
    taken from them, while the exchange (56.6 MB per step) has the whole backward (7.7 ms) to finish in.  Measured on one box
    with 2 B200s, 20 steps each, back to back (profiles/r2/r2_notes.md): no exchange 13.346 / 13.358 ms per step, cap 2
    13.359, NCCL's own width 13.373 -- the overlapped exchange costs 0.02 ms either way, the cap is worth 0.01 ms at best.
    Hence off by default; the knob stays for larger payloads."""
    global _CAPPED
    import os
    opts = dist.ProcessGroupNCCL.Options()
    if "NCCL_MAX_CTAS" not in os.environ and exchange_ctas > 0:
        opts.config.max_ctas = exchange_ctas
        _CAPPED = True
    dist.init_process_group("nccl", device_id=device, pg_options=opts)


_FULL_WIDTH_GROUP = None


def full_width_group():
    """A second NCCL communicator with NCCL's own CTA count, for exchanges that are NOT overlapped with this path's kernels
    (deterministic mode: the exchange runs to completion before the backward, so it should be as short as it can be).
    None (= the default group) when init_nccl() applied no cap, on other backends or in a single process."""
    global _FULL_WIDTH_GROUP
    if not _CAPPED or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or dist.get_backend() != "nccl":
        return None
    if _FULL_WIDTH_GROUP is None:
        _FULL_WIDTH_GROUP = dist.new_group(backend="nccl", pg_options=dist.ProcessGroupNCCL.Options())
    return _FULL_WIDTH_GROUP


class FlatGradReducer:
    """One flat-buffer gradient all-reduce per step -- the DDP-equivalent exchange of a data-parallel training step.

    Gradients of `params` are packed into ONE contiguous buffer (fp32, or bf16 to halve the bytes on the wire),
    all-reduced once (SUM), divided by the world size and unpacked in place.  One launch-latency instead of one per
    bucket: the payload (56.7 MB fp32 for WIF) is far below what NVLink 5 / NVSwitch moves in the time of a step."""

    def __init__(self, params: Iterable[torch.Tensor], wire_dtype: torch.dtype = torch.float32, group=None):
        self.params = [p for p in params]
        self.group = group   # process group of the exchange (None = the default group)
        if not self.params:
            raise ValueError("FlatGradReducer: no parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=wire_dtype)
        self._divide = True
        self.offsets = []
        o = 0
        for p in self.params:
            self.offsets.append(o)
            o += p.numel()
        # Gradients that already exist and have the wire dtype become VIEWS of the flat buffer (DDP's
        # gradient_as_bucket_view): autograd accumulates into them in place, so packing and unpacking move no bytes.  A
        # .grad that is later replaced by another tensor (set_to_none, first backward of a fresh parameter) falls back to
        # the copies below for that parameter.
        for p, o in zip(self.params, self.offsets):
            if p.grad is not None and p.grad.dtype == wire_dtype and p.grad.device == dev:
                seg = self.flat[o:o + p.numel()].view_as(p)
                seg.copy_(p.grad)
                p.grad = seg

    def _is_view(self, p, o) -> bool:
        g = p.grad
        return (g is not None and g.dtype == self.flat.dtype and g.is_contiguous()
                and g.data_ptr() == self.flat.data_ptr() + o * self.flat.element_size())

    def _pack(self):
        for p, o in zip(self.params, self.offsets):
            if self._is_view(p, o):
                continue
            seg = self.flat[o:o + p.numel()]
            if p.grad is None:
                seg.zero_()
            else:
                seg.copy_(p.grad.reshape(-1))

    def _all_reduce_mean(self, async_op: bool):
        """SUM / world in one collective where the backend can (NCCL: ReduceOp.AVG); returns (work, still_to_divide)."""
        if dist.get_backend(self.group) == "nccl" and self.flat.is_cuda:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op), False
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op), True

    def reduce_async(self):
        """Pack and START the all-reduce on the backend's own stream (NCCL: its internal stream, so it overlaps with
        whatever the caller enqueues next -- e.g. the rest of a backward pass).  Returns the work handle for finish()."""
        _, ws = world()
        self._pack()
        if ws == 1:
            return None
        work, self._divide = self._all_reduce_mean(True)
        return work

    def finish(self, work) -> torch.Tensor:
        """Wait for reduce_async()'s exchange (the caller's stream waits, not the host), average and unpack into .grad."""
        _, ws = world()
        if work is not None:
            work.wait()
            if self._divide:
                self.flat.div_(ws)
        return self._unpack(ws)

    def reduce(self) -> torch.Tensor:
        """Pack -> all_reduce(SUM) -> / world -> unpack into .grad.  Returns the flat (averaged) buffer."""
        _, ws = world()
        self._pack()
        if ws > 1:
            _, divide = self._all_reduce_mean(False)
            if divide:
                self.flat.div_(ws)
        return self._unpack(ws)

    def _unpack(self, ws) -> torch.Tensor:
        for p, o in zip(self.params, self.offsets):
            if self._is_view(p, o):
                continue
            seg = self.flat[o:o + p.numel()]
            if p.grad is not None:
                p.grad.copy_(seg.view_as(p.grad))
            elif ws > 1:
                # unused on THIS rank but maybe not on another: every rank must apply the same averaged gradient (what DDP
                # does), or the replicas diverge
                p.grad = seg.to(p.dtype).view_as(p).clone()
        return self.flat


def any_nan(flag: torch.Tensor) -> bool:
    """models/synthesizer.py:619: all-gather a 1-element NaN flag; True if any rank saw a NaN."""
    _, ws = world()
    f = flag.reshape(1).to(torch.uint8)
    if ws == 1:
        return bool(f.item())
    got = [torch.zeros_like(f) for _ in range(ws)]
    dist.all_gather(got, f)
    return bool(torch.stack(got).any().item())


def max_over_ranks(values: Sequence[float], device=None) -> List[float]:
    """Element-wise max over ranks of a short list of floats (device times in ms)."""
    _, ws = world()
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if ws > 1:
        if t.is_cuda:
            t = t.float()   # NCCL has no fp64 restriction, but keep the payload small and uniform
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def gather_over_ranks(value: float, device=None) -> List[float]:
    """Every rank's value, in rank order, on every rank (per-rank device times beside their max: tells a slow GPU from
    a communication cost in the weak-scaling numbers)."""
    _, ws = world()
    if ws == 1:
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float32, device=device)
    got = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(got, t)
    return [float(g.item()) for g in got]


def whole_job_rate(units_per_rank: int, ms: float) -> float:
    """Whole-job throughput: units all ranks processed / (max-over-ranks time)."""
    _, ws = world()
    return ws * units_per_rank / (ms * 1e-3)
