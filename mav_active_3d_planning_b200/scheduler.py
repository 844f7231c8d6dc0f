"""Multi-GPU segment scheduler: one process per GPU, segments sharded, map replicated.

The reference evaluates one segment per call from a single thread
(active_3d_planning_core/src/planner/online_planner.cpp:414-416); a segment's gain depends only on
(map, its own poses, static parameters) (simulated_sensor_evaluator.cpp:45-82), so segments are
independent units.  This module is the plumbing around the C-ABI calls:

* ``shard_bounds``     contiguous block partition of the segment list over ranks (all views of a
                       segment stay on one rank: std::unique and the gain sum span them);
* ``broadcast_delta``  the planning step's updated-block list travels rank 0 → all ranks (NCCL
                       broadcast over NVLink when the tensors are on the GPU, gloo on CPU);
* ``gather_gains``     per-segment gains (8 B each) are all-gathered back in segment order.

``torch.distributed`` is only used for these two exchanges; every number is computed by the CUDA
kernels behind ``capi.Context``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_views", "broadcast_delta", "gather_gains", "SegmentScheduler"]


def shard_bounds(n_segments: int, world: int, rank: int):
    """[lo, hi) of the segments rank ``rank`` owns: ceil(n/world) contiguous segments per rank."""
    per = -(-n_segments // world) if world > 0 else n_segments
    lo = min(rank * per, n_segments)
    return lo, min(lo + per, n_segments)


def shard_views(view_segment: np.ndarray, lo: int, hi: int):
    """Views (index range) belonging to segments [lo, hi) of a non-decreasing view→segment map."""
    view_segment = np.asarray(view_segment)
    v0 = int(np.searchsorted(view_segment, lo, side="left"))
    v1 = int(np.searchsorted(view_segment, hi, side="left"))
    return v0, v1


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def broadcast_delta(block_idx: torch.Tensor, payload: torch.Tensor, src: int = 0, group=None):
    """Broadcast one packed block delta (``block_idx`` int32 n×3, ``payload`` uint8 = fp32 distances
    followed by the observed bytes) from ``src`` to every rank, in place."""
    world, _ = _world(group)
    if world > 1:
        dist.broadcast(block_idx, src=src, group=group)
        dist.broadcast(payload, src=src, group=group)
    return block_idx, payload


def gather_gains(local: torch.Tensor, n_segments: int, group=None) -> torch.Tensor:
    """All-gather the per-rank gain shards (equal-sized, padded by shard_bounds' ceil split) into the
    full per-segment vector, identical on every rank."""
    world, _ = _world(group)
    if world == 1:
        return local[:n_segments].clone()
    per = -(-n_segments // world)
    buf = torch.zeros(per, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(per * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[:n_segments]


class SegmentScheduler:
    """Evaluate a global list of segments with every rank's GPU and return all gains everywhere.

    ``compute_fn(pos, quat, view_segment_local, n_local) -> np.ndarray[n_local]`` is the per-rank
    evaluation (normally ``capi.Context.compute_gains(...)['gain']``); it is injected so that the
    sharding / gather logic can be exercised with the gloo backend on CPU-only machines."""

    def __init__(self, compute_fn, device="cpu", group=None):
        self.compute_fn = compute_fn
        self.device = torch.device(device)
        self.group = group

    def evaluate(self, pos: np.ndarray, quat: np.ndarray, view_segment: np.ndarray, n_segments: int):
        world, rank = _world(self.group)
        lo, hi = shard_bounds(n_segments, world, rank)
        v0, v1 = shard_views(view_segment, lo, hi)
        local_seg = (np.asarray(view_segment[v0:v1]) - lo).astype(np.int32)
        gains = np.zeros(0, np.float64)
        if hi > lo:
            gains = np.asarray(self.compute_fn(pos[v0:v1], quat[v0:v1], local_seg, hi - lo), np.float64)
        t = torch.from_numpy(gains).to(self.device)
        return gather_gains(t, n_segments, self.group).cpu().numpy()
