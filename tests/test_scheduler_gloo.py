"""Multi-rank segment scheduler on CPU: world_size-2 gloo processes exercise the sharding, the block
delta broadcast and the gain gather (the N>1 host logic of bench.py) with an injected compute stub."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mav_active_3d_planning_b200 import scheduler


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [scheduler.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(s for s in sizes if s or n == 0 or True) <= -(-n // world)


def test_shard_views_keeps_segments_whole():
    seg = np.array([0, 0, 0, 1, 3, 3, 4, 4, 4, 4, 6], np.int32)   # segments 2 and 5 have no views
    assert scheduler.shard_views(seg, 0, 4) == (0, 6)
    assert scheduler.shard_views(seg, 4, 7) == (6, 11)
    assert scheduler.shard_views(seg, 2, 3) == (4, 4)


def _fake_gain(pos, quat, seg_local, n_local):
    # deterministic per-segment function of the poses: sum over the segment's views
    out = np.zeros(n_local)
    np.add.at(out, seg_local, pos[:, 0] * 3.0 + quat[:, 3])
    return out


def _worker(rank, world, port, n_segments, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)            # same global candidate list on every rank
        views_per = rng.integers(0, 4, n_segments)
        seg = np.repeat(np.arange(n_segments, dtype=np.int32), views_per)
        pos = rng.random((len(seg), 3))
        quat = rng.random((len(seg), 4))
        # block delta: rank 0 owns it, everybody must end up with the same bytes
        idx = torch.zeros((5, 3), dtype=torch.int32)
        payload = torch.zeros(5 * 64 * 5, dtype=torch.uint8)
        if rank == 0:
            idx = torch.arange(15, dtype=torch.int32).reshape(5, 3)
            payload = (torch.arange(payload.numel()) % 251).to(torch.uint8)
        scheduler.broadcast_delta(idx, payload)
        sched = scheduler.SegmentScheduler(_fake_gain)
        gains = sched.evaluate(pos, quat, seg, n_segments)
        want = _fake_gain(pos, quat, seg, n_segments)
        ok = (np.allclose(gains, want) and int(idx.sum()) == 105 and
              int(payload.to(torch.int64).sum()) == int((torch.arange(payload.numel()) % 251).sum()))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_segments", [1, 9, 64])
def test_two_rank_gloo_evaluate(n_segments):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_segments, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True
