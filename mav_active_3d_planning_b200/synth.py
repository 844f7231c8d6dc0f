"""Deterministic synthetic voxblox-style maps and candidate views.

The reference has no datasets on this path: its maps come from a live voxblox ``esdf_server``
(``active_3d_planning_app_reconstruction/launch/run_experiment.launch:159-181``).  These generators
produce the same *data format* the map adapter consumes — allocated 3-D block indices plus, per
block, ``vps**3`` ESDF distances (fp32) and ``observed`` flags in voxblox's linear voxel order
``x + vps*(y + vps*z)`` — for an analytic scene (room shell + axis-aligned boxes) with a partially
observed region.  numpy only; no oracle, no GPU.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

__all__ = ["SynthMap", "room_map", "clutter_map", "candidate_views", "yaw_quat", "CONFIGS"]


@dataclasses.dataclass
class SynthMap:
    voxel_size: float          # fp32 value as voxblox stores it
    vps: int                   # voxels per block side
    block_idx: np.ndarray      # (n,3) int32
    dist: np.ndarray           # (n, vps^3) float32
    observed: np.ndarray       # (n, vps^3) uint8
    weight: np.ndarray         # (n, vps^3) float32 TSDF weights (0 where unobserved)
    extent: tuple              # scene AABB (Lx, Ly, Lz) metres
    boxes: np.ndarray          # (k,6) centre xyz, half-size xyz

    @property
    def n_blocks(self) -> int:
        return int(self.block_idx.shape[0])

    @property
    def block_size(self) -> float:
        return float(np.float32(self.voxel_size) * np.float32(self.vps))

    def nearest(self, pts: np.ndarray):
        """Nearest-voxel (distance, observed) for (n,3) points; unallocated → (nan, 0)."""
        vs = float(np.float32(self.voxel_size))
        g = np.floor(np.asarray(pts, np.float64) / vs).astype(np.int64)
        b = np.floor_divide(g, self.vps)
        v = g - b * self.vps
        lo = self.block_idx.min(axis=0).astype(np.int64)
        hi = self.block_idx.max(axis=0).astype(np.int64)
        grid = np.full(tuple(hi - lo + 1), -1, np.int64)
        rel = self.block_idx.astype(np.int64) - lo
        grid[rel[:, 0], rel[:, 1], rel[:, 2]] = np.arange(self.n_blocks)
        inb = np.all((b >= lo) & (b <= hi), axis=1)
        bc = np.clip(b, lo, hi) - lo
        slot = np.where(inb, grid[bc[:, 0], bc[:, 1], bc[:, 2]], -1)
        lin = v[:, 0] + self.vps * (v[:, 1] + self.vps * v[:, 2])
        ok = slot >= 0
        d = np.full(len(g), np.nan, np.float32)
        o = np.zeros(len(g), np.uint8)
        d[ok] = self.dist[slot[ok], lin[ok]]
        o[ok] = self.observed[slot[ok], lin[ok]]
        return d, o


def _splitmix64(state: int):
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


class _Rng:
    """Tiny portable generator so that scenes do not depend on the numpy version."""

    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def uniform(self, lo: float, hi: float) -> float:
        self.s, z = _splitmix64(self.s)
        return lo + (hi - lo) * ((z >> 11) / float(1 << 53))


def _make_boxes(n, seed, extent, edge_lo, edge_hi, z_hi):
    r = _Rng(seed)
    out = np.zeros((n, 6), np.float64)
    for k in range(n):
        out[k, 0] = r.uniform(3.0, extent[0] - 3.0)
        out[k, 1] = r.uniform(3.0, extent[1] - 3.0)
        out[k, 2] = r.uniform(0.0, z_hi)
        out[k, 3:6] = [0.5 * r.uniform(edge_lo, edge_hi) for _ in range(3)]
    return out


def _sdf_chunk(x, y, z, extent, shell, boxes, dmax):
    """Signed distance (float64) on the tensor grid x⊗y⊗z to room shell ∪ boxes, clipped to ±dmax."""
    X = x[:, None, None]
    Y = y[None, :, None]
    Z = z[None, None, :]
    d = np.minimum(X - shell, extent[0] - shell - X)
    d = np.minimum(d, np.minimum(Y - shell, extent[1] - shell - Y))
    d = np.minimum(d, np.minimum(Z - shell, extent[2] - shell - Z))
    d = np.broadcast_to(d, (len(x), len(y), len(z))).copy()
    lo = np.array([x[0], y[0], z[0]]) - dmax
    hi = np.array([x[-1], y[-1], z[-1]]) + dmax
    for c in boxes:
        if np.any(c[:3] + c[3:] < lo) or np.any(c[:3] - c[3:] > hi):
            continue
        qx = np.abs(X - c[0]) - c[3]
        qy = np.abs(Y - c[1]) - c[4]
        qz = np.abs(Z - c[2]) - c[5]
        outside = np.sqrt(np.maximum(qx, 0.0) ** 2 + np.maximum(qy, 0.0) ** 2 + np.maximum(qz, 0.0) ** 2)
        inside = np.minimum(np.maximum(qx, np.maximum(qy, qz)), 0.0)
        d = np.minimum(d, outside + inside)
    return np.clip(d, -dmax, dmax)


def _build_slab(args):
    """One x-slab of blocks (bx): (block_idx, dist, observed) of its allocated blocks, or None."""
    bx, vs, vps, nb, extent, boxes, wp, obs_radius, dmax = args
    nv = vps ** 3
    yv = (np.arange(nb[1] * vps) + 0.5) * vs
    zv = (np.arange(nb[2] * vps) + 0.5) * vs
    xv = (np.arange(bx * vps, (bx + 1) * vps) + 0.5) * vs
    # skip x-slabs that no waypoint sphere reaches
    if np.all(np.abs(np.clip(wp[:, 0], xv[0], xv[-1]) - wp[:, 0]) > obs_radius):
        return None
    d = _sdf_chunk(xv, yv, zv, extent, vs, boxes, dmax)
    obs = np.zeros(d.shape, bool)
    for w in wp:
        r2 = ((xv - w[0]) ** 2)[:, None, None] + ((yv - w[1]) ** 2)[None, :, None] \
            + ((zv - w[2]) ** 2)[None, None, :]
        obs |= r2 <= obs_radius * obs_radius
    # (vps, nby, vps, nbz, vps) → (nby, nbz, z, y, x)
    d5 = d.reshape(vps, nb[1], vps, nb[2], vps).transpose(1, 3, 4, 2, 0).reshape(nb[1], nb[2], nv)
    o5 = obs.reshape(vps, nb[1], vps, nb[2], vps).transpose(1, 3, 4, 2, 0).reshape(nb[1], nb[2], nv)
    keep = o5.any(axis=2)
    by, bz = np.nonzero(keep)
    if len(by) == 0:
        return None
    return (np.stack([np.full(len(by), bx), by, bz], 1).astype(np.int32), d5[by, bz].astype(np.float32),
            o5[by, bz].astype(np.uint8))


def _build(voxel_size, vps, extent, boxes, waypoints, obs_radius, dmax, weight_seed, workers=None):
    vs = float(np.float32(voxel_size))
    bs = vs * vps
    nb = [int(math.ceil(extent[k] / bs - 1e-9)) for k in range(3)]
    wp = np.asarray(waypoints, np.float64)
    nv = vps ** 3
    jobs = [(bx, vs, vps, nb, tuple(extent), boxes, wp, obs_radius, dmax) for bx in range(nb[0])]
    # big scenes (cfg3: 63 slabs of 6.4 M voxels) are built slab-parallel; the result does not depend
    # on the number of workers (slabs are independent and concatenated in order)
    if workers is None:
        import os
        workers = min(16, os.cpu_count() or 1) if nb[0] * nb[1] * nb[2] * nv > 64_000_000 else 1
    if workers > 1:
        import concurrent.futures as cf
        import multiprocessing as mp
        with cf.ProcessPoolExecutor(workers, mp_context=mp.get_context("fork")) as ex:
            slabs = list(ex.map(_build_slab, jobs))
    else:
        slabs = [_build_slab(j) for j in jobs]
    slabs = [s for s in slabs if s is not None]
    block_idx = np.concatenate([s[0] for s in slabs]) if slabs else np.zeros((0, 3), np.int32)
    dist = np.concatenate([s[1] for s in slabs]) if slabs else np.zeros((0, nv), np.float32)
    observed = np.concatenate([s[2] for s in slabs]) if slabs else np.zeros((0, nv), np.uint8)
    wrng = np.random.Generator(np.random.PCG64(weight_seed))
    weight = (wrng.random(dist.shape, dtype=np.float32) * 99.0 + 1.0) * observed
    return SynthMap(vs, vps, block_idx, dist, observed, weight.astype(np.float32), tuple(extent), boxes)


def room_map(voxel_size=0.1, vps=16, dmax=2.0) -> SynthMap:
    """M-room: 20×20×10 m room, 1-voxel shell + 6 boxes, observed within 6 m of 3 waypoints."""
    extent = (20.0, 20.0, 10.0)
    boxes = _make_boxes(6, 1, extent, 1.0, 3.0, 6.0)
    wps = [(4.0, 4.0, 2.0), (10.0, 10.0, 3.0), (16.0, 14.0, 4.0)]
    return _build(voxel_size, vps, extent, boxes, wps, 6.0, dmax, 2)


def clutter_map(voxel_size=0.05, vps=16, dmax=2.0, extent=(50.0, 50.0, 20.0), n_boxes=200,
                obs_radius=8.0) -> SynthMap:
    """M-clutter: shell + 200 boxes, observed within 8 m of a 12-waypoint lawn-mower path."""
    boxes = _make_boxes(n_boxes, 7, extent, 0.5, 3.0, min(12.0, extent[2] * 0.6))
    wps = []
    for row in range(3):
        ys = extent[1] * (row + 0.5) / 3.0
        xs = np.linspace(6.0, extent[0] - 6.0, 4)
        if row % 2:
            xs = xs[::-1]
        wps += [(float(x), ys, 3.0) for x in xs]
    return _build(voxel_size, vps, extent, boxes, wps, obs_radius, dmax, 2)


def yaw_quat(yaw: np.ndarray) -> np.ndarray:
    """Quaterniond(AngleAxisd(yaw, UnitZ)) as (n,4) w,x,y,z (rrt_star.cpp:113-114, trajectory.h:18-20)."""
    yaw = np.asarray(yaw, np.float64)
    q = np.zeros(yaw.shape + (4,), np.float64)
    q[..., 0] = np.cos(yaw * 0.5)
    q[..., 3] = np.sin(yaw * 0.5)
    return q


def candidate_views(m: SynthMap, n: int, seed: int = 42, min_clearance: float = 0.5):
    """n candidate poses: uniform positions in the shrunk AABB that are observed and
    ≥ min_clearance from obstacles; uniform yaw, roll = pitch = 0.  Returns (pos n×3, quat n×4)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = np.zeros((0, 3), np.float64)
    lo = np.array([0.5, 0.5, 0.5])
    hi = np.array(m.extent) - 0.5
    while len(pos) < n:
        cand = lo + (hi - lo) * rng.random((max(2 * n, 256), 3))
        d, o = m.nearest(cand)
        ok = (o != 0) & (d > min_clearance)
        pos = np.concatenate([pos, cand[ok]])
    pos = np.ascontiguousarray(pos[:n])
    yaw = rng.random(n) * (2.0 * math.pi)
    return pos, np.ascontiguousarray(yaw_quat(yaw))


# BASELINE.json configs made concrete.  caster kind: 0 Simple, 1 Iterative.
CONFIGS = {
    "cfg1": dict(map="room", voxel_size=0.2, views=64, caster=1, evaluator="VoxelType",
                 resolution=(320, 240), focal_length=160.0, ray_length=5.0),
    "cfg2": dict(map="room", voxel_size=0.1, views=4096, caster=1, evaluator="VoxelType",
                 resolution=(320, 240), focal_length=160.0, ray_length=5.0),
    "cfg3": dict(map="clutter", voxel_size=0.05, views=4096, caster=1, evaluator="Frontier",
                 resolution=(320, 240), focal_length=160.0, ray_length=5.0),
    "cfg4": dict(map="room", voxel_size=0.1, views=64, caster=0, evaluator="VoxelType",
                 resolution=(640, 480), focal_length=320.0, ray_length=8.0),
}
