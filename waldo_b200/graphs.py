"""CUDA-graph replay of the inference path (control points -> grids -> occlusion matrix -> fused warp + composite).

One rollout step issues ~60 small-to-medium kernels plus the allocations behind them; at KITTI size (256x832, one
video per GPU) the device finishes them faster than Python can launch them.  GraphedDecode captures the whole chain
once per distinct set of input buffers and replays it with a single launch.  Inference only (no autograd); the
training path keeps the eager autograd Functions."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import modules as M


class GraphedDecode:
    """Callable with the arguments of the eager pair
        occ, oa, ba, grid = estimate_alpha_grid_occ(warper, obj_alpha, mask, bg_alpha_buf, obj_pose, bg_pose, occ_score)
        decode_output(warper, input, grid, occ, oa, ba, cls, ctx_ts, pred_ts, restrict_to_ctx)
    returning the same 7-tuple.  A graph is captured the first time a given set of input BUFFERS (data pointers and
    shapes) is seen and replayed afterwards: callers that reuse their device buffers (DevicePrefetcher does) pay the
    capture once.  The returned tensors live in the graph's memory pool and are overwritten by its next replay."""

    def __init__(self, warper: M.Warper, obj_alpha_mask, bg_alpha_buf, restrict_to_ctx: bool, max_graphs: int = 4):
        self.warper, self.mask, self.bg, self.restrict = warper, obj_alpha_mask, bg_alpha_buf, bool(restrict_to_ctx)
        self.max_graphs = max_graphs
        self.cache: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, tuple, tuple]] = {}

    def _eager(self, input, obj_alpha_raw, obj_pose, bg_pose, occ_score, cls, ctx_ts, pred_ts):
        occ, oa, ba, grid = M.estimate_alpha_grid_occ(self.warper, obj_alpha_raw, self.mask, self.bg, obj_pose, bg_pose, occ_score)
        return M.decode_output(self.warper, input, grid, occ, oa, ba, cls, ctx_ts, pred_ts, self.restrict)

    def __call__(self, input, obj_alpha_raw, obj_pose, bg_pose, occ_score, cls, ctx_ts, pred_ts):
        args = (input, obj_alpha_raw, obj_pose, bg_pose, occ_score, cls, ctx_ts, pred_ts)
        for t in args:
            if t is not None and (not t.is_cuda or not t.is_contiguous()):
                raise RuntimeError("waldo_b200.GraphedDecode: inputs must be contiguous CUDA tensors (static buffers)")
        key = tuple((t.data_ptr(), tuple(t.shape), t.dtype) if t is not None else None for t in args)
        hit = self.cache.get(key)
        if hit is None:
            if len(self.cache) >= self.max_graphs:
                self.cache.pop(next(iter(self.cache)))
            with torch.no_grad():
                side = torch.cuda.Stream(input.device)
                side.wait_stream(torch.cuda.current_stream(input.device))
                with torch.cuda.stream(side):          # warm-up off the capture: builds caches, validates the time indices
                    for _ in range(2):
                        self._eager(*args)
                torch.cuda.current_stream(input.device).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self._eager(*args)
            hit = (graph, out, args)   # keep the input buffers alive: the graph reads them by address
            self.cache[key] = hit
        hit[0].replay()
        return hit[1]
