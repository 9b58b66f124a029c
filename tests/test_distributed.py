"""The N > 1 host logic (waldo_b200/sharding.py) under a world_size-2 gloo process group on CPU: video sharding,
the flat DDP-equivalent gradient all-reduce, the NaN-flag gather and the max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from waldo_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.world() == (rank, world)
        # --- sharding by video: every video exactly once, equal per-rank counts
        mine = sharding.shard_videos(11, rank, world)
        got = [None] * world
        dist.all_gather_object(got, mine)
        flat = sorted(i for g in got for i in g)
        assert flat == list(range(10)) and all(len(g) == 5 for g in got)
        assert sharding.per_rank_batch(16, world) == 8
        # --- flat gradient all-reduce == mean of the per-rank gradients, written back into .grad
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(s)) for s in ((3, 5), (7,), (2, 2, 2))]
        per_rank = []
        for r in range(world):
            g = torch.Generator().manual_seed(100 + r)
            per_rank.append([torch.randn(p.shape, generator=g) for p in params])
        for p, g in zip(params, per_rank[rank]):
            p.grad = g.clone()
        for wire, tol in ((torch.float32, 1e-6), (torch.bfloat16, 2e-2)):
            for p, g in zip(params, per_rank[rank]):
                p.grad.copy_(g)
            red = sharding.FlatGradReducer(params, wire_dtype=wire)
            flatbuf = red.reduce()
            assert flatbuf.numel() == sum(p.numel() for p in params)
            for i, p in enumerate(params):
                want = sum(per_rank[r][i] for r in range(world)) / world
                assert torch.allclose(p.grad, want, atol=tol, rtol=tol), (wire, i)
        # --- the overlapped form (start on the backend's stream, finish later) gives the same averages
        for p, g in zip(params, per_rank[rank]):
            p.grad.copy_(g)
        red = sharding.FlatGradReducer(params)
        work = red.reduce_async()
        red.finish(work)
        for i, p in enumerate(params):
            want = sum(per_rank[r][i] for r in range(world)) / world
            assert torch.allclose(p.grad, want, atol=1e-6, rtol=1e-6), ("async", i)
        # --- gradients of the wire dtype are views of the flat buffer (nothing to pack / unpack); a replaced .grad falls back to copies
        assert all(red._is_view(p, o) for p, o in zip(red.params, red.offsets))
        for p, g in zip(params, per_rank[rank]):
            p.grad.copy_(g)                       # what autograd's in-place accumulation does
        params[1].grad = per_rank[rank][1].clone()   # ... and a gradient tensor the reducer has not seen
        assert not red._is_view(params[1], red.offsets[1]) and red._is_view(params[0], red.offsets[0])
        red.reduce()
        for i, p in enumerate(params):
            want = sum(per_rank[r][i] for r in range(world)) / world
            assert torch.allclose(p.grad, want, atol=1e-6, rtol=1e-6), ("views", i)
        # --- a parameter unused on THIS rank still receives the averaged gradient (replicas must not diverge)
        q = [torch.nn.Parameter(torch.zeros(4))]
        q[0].grad = torch.full((4,), 2.0) if rank == 0 else None
        sharding.FlatGradReducer(q).reduce()
        assert q[0].grad is not None and torch.allclose(q[0].grad, torch.full((4,), 1.0)), "unused-parameter gradient"
        # --- NaN flag and timing reductions
        assert sharding.any_nan(torch.tensor(rank == 1)) is True
        assert sharding.any_nan(torch.tensor(False)) is False
        assert sharding.max_over_ranks([1.0 + rank, 5.0 - rank]) == [2.0, 5.0]
        assert sharding.whole_job_rate(8, 1000.0) == 16.0
        assert sharding.gather_over_ranks(3.0 + rank) == [3.0, 4.0]
        out.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_single_process_defaults():
    assert sharding.world() == (0, 1)
    assert sharding.shard_videos(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.shard_videos(5, 1, 2, drop_last=False) == [1, 3]
    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.full((4,), 3.0)
    sharding.FlatGradReducer([p]).reduce()
    assert torch.equal(p.grad, torch.full((4,), 3.0))
    assert sharding.any_nan(torch.tensor(True)) is True
