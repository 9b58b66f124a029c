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


class FlatGradReducer:
    """One flat-buffer gradient all-reduce per step -- the DDP-equivalent exchange of a data-parallel training step.

    Gradients of `params` are packed into ONE contiguous buffer (fp32, or bf16 to halve the bytes on the wire),
    all-reduced once (SUM), divided by the world size and unpacked in place.  One launch-latency instead of one per
    bucket: the payload (56.7 MB fp32 for WIF) is far below what NVLink 5 / NVSwitch moves in the time of a step."""

    def __init__(self, params: Iterable[torch.Tensor], wire_dtype: torch.dtype = torch.float32):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatGradReducer: no parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=wire_dtype)
        self.offsets = []
        o = 0
        for p in self.params:
            self.offsets.append(o)
            o += p.numel()

    def _pack(self):
        for p, o in zip(self.params, self.offsets):
            seg = self.flat[o:o + p.numel()]
            if p.grad is None:
                seg.zero_()
            else:
                seg.copy_(p.grad.reshape(-1))

    def reduce_async(self):
        """Pack and START the all-reduce on the backend's own stream (NCCL: its internal stream, so it overlaps with
        whatever the caller enqueues next -- e.g. the rest of a backward pass).  Returns the work handle for finish()."""
        _, ws = world()
        self._pack()
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True) if ws > 1 else None

    def finish(self, work) -> torch.Tensor:
        """Wait for reduce_async()'s exchange (the caller's stream waits, not the host), average and unpack into .grad."""
        _, ws = world()
        if work is not None:
            work.wait()
            self.flat.div_(ws)
        return self._unpack(ws)

    def reduce(self) -> torch.Tensor:
        """Pack -> all_reduce(SUM) -> / world -> unpack into .grad.  Returns the flat (averaged) buffer."""
        _, ws = world()
        self._pack()
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(ws)
        return self._unpack(ws)

    def _unpack(self, ws) -> torch.Tensor:
        for p, o in zip(self.params, self.offsets):
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


def whole_job_rate(units_per_rank: int, ms: float) -> float:
    """Whole-job throughput: units all ranks processed / (max-over-ranks time)."""
    _, ws = world()
    return ws * units_per_rank / (ms * 1e-3)
