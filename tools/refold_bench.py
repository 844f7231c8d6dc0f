#!/usr/bin/env python
"""Time the device-resident list path on the bench workload: evaluate + store
4096 segments, then re-fold all stored lists after a 1 % block delta (what UpdateAll/
SimulatedSensorUpdater do for the whole tree after every executed segment)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mav_active_3d_planning_b200 import capi  # noqa: E402
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    m, pos, quat, (d_idx, d_dist, d_obs) = bench.make_workload(n)
    ctx = capi.Context(0)
    ctx.upload_map(m)
    caster = ctx.caster("IterativeRayCaster", **bench.CAM)
    out = {"segments": n}
    for ekind in ("VoxelTypeEvaluator", "NaiveEvaluator"):
        ev = ctx.evaluator(ekind)
        store = capi.Store(ctx)
        keys = np.arange(1, n + 1, dtype=np.uint64)  # (key 0 means "no parent")
        t0 = time.perf_counter()
        r = store.compute_gains(caster, ev, keys, pos, quat)
        t_store = time.perf_counter() - t0
        ctx.upload_blocks(d_idx, d_dist, d_obs)
        times, kms = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            gains, left = store.refold(ev, keys)
            times.append(time.perf_counter() - t0)
            kms.append(ctx.last_kernel_ms)
        recast = ctx.compute_gains(caster, ev, pos, quat, digests=False, set_digests=False)
        ok = bool(np.allclose(gains, recast["gain"], rtol=1e-5))
        entries = int(r["n_visible"].sum())
        out[ekind] = {"store_s": t_store, "store_bytes": store.nbytes, "entries": entries,
                      "refold_wall_ms": 1e3 * min(times), "refold_kernel_ms": min(kms),
                      "recast_kernel_ms": ctx.last_kernel_ms, "refold_equals_recast": ok,
                      "refold_GBps_24B_per_entry": entries * 24 / (min(kms) * 1e-3) / 1e9}
        store.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
