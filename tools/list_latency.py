#!/usr/bin/env python
"""Latency of the per-segment list call (a3d_visible_voxels: what SimulatedSensorEvaluator::computeGain
costs) with the wavefront kernel's list emission versus the scan-order kernel, and of the batched
stored-list call for a few batch sizes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mav_active_3d_planning_b200 import capi, synth  # noqa: E402


def main():
    cfg = synth.CONFIGS["cfg2"]
    m = synth.room_map(cfg["voxel_size"])
    ctx = capi.Context(0)
    ctx.upload_map(m)
    caster = ctx.caster("IterativeRayCaster", resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0)
    ev = ctx.evaluator("VoxelTypeEvaluator")
    pos, quat = synth.candidate_views(m, 2048, seed=42)
    out = {}
    for name, opt in (("chain", 0), ("wave", 2), ("scan", 1)):  # default = kernel chain with list emission
        ctx.set_option("cast_kernel", opt)
        ks, ws = [], []
        for v in range(6):
            t0 = time.perf_counter()
            lst, _, _ = ctx.visible_voxels(caster, pos[v:v + 1], quat[v:v + 1], evaluator=ev)
            ws.append((time.perf_counter() - t0) * 1e3)
            ks.append(ctx.last_kernel_ms)
        out["single_view_" + name] = dict(kernel_ms=min(ks[1:]), wall_ms=min(ws[1:]), entries=len(lst))
        for n in (12, 148, 2048):
            store = capi.Store(ctx)
            keys = np.arange(1, n + 1, dtype=np.uint64)
            best = 1e9
            try:
                for r in range(3):
                    t0 = time.perf_counter()
                    store.compute_gains(caster, ev, keys, pos[:n], quat[:n])
                    best = min(best, (time.perf_counter() - t0) * 1e3)
                out[f"stored_{n}_{name}"] = dict(wall_ms=best)
            except capi.A3DError as e:  # (the wavefront kernel emits lists for up to 2 x SM count segments)
                out[f"stored_{n}_{name}"] = dict(error=str(e))
            store.close()
    ctx.set_option("cast_kernel", 0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
