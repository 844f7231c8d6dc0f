#!/usr/bin/env python
"""Scaling of the single-process multi-GPU group (a3d_group_*, ncclCommInitAll) on one box.

    python tools/group_bench.py --gpus 1,2,4,8 --workload cfg2 --views 4096 [--weak] [--steps 10]

strong scaling (default): the SAME global candidate list (--views segments) on N GPUs — the planner's
step latency; --weak: --views segments PER GPU.  Each step = one 1 % block delta (host → GPU 0 →
ncclBroadcast → K1 on every GPU) + a3d_group_compute_gains (host buffers in, gathered gains out).
Prints one JSON line per N with the step time (host wall clock around the synchronous calls), the
slowest GPU's cast kernel time, and a digest of the gathered gain vector (equal for every N on the
same list)."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mav_active_3d_planning_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--views", type=int, default=4096)
    ap.add_argument("--weak", action="store_true")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--evaluator", default=None)
    ap.add_argument("--deal-block", type=int, default=64)
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    cfg = dict(synth.CONFIGS[args.workload])
    if args.evaluator:
        cfg["evaluator"] = args.evaluator
    m = synth.room_map(cfg["voxel_size"]) if cfg["map"] == "room" else synth.clutter_map(cfg["voxel_size"])
    kind = "IterativeRayCaster" if cfg["caster"] == 1 else "SimpleRayCaster"
    cam = dict(resolution_x=cfg["resolution"][0], resolution_y=cfg["resolution"][1],
               focal_length=cfg["focal_length"], ray_length=cfg["ray_length"])
    rng = np.random.default_rng(3)
    n_delta = max(1, m.n_blocks // 100)
    sel = np.sort(rng.choice(m.n_blocks, n_delta, replace=False))
    d_idx = np.ascontiguousarray(m.block_idx[sel])
    d_dist = (m.dist[sel] + rng.normal(0, 0.002, m.dist[sel].shape)).astype(np.float32)
    d_obs = np.ascontiguousarray(m.observed[sel])
    t1 = None
    for n in [int(x) for x in args.gpus.split(",")]:
        if n > have:
            print(json.dumps({"n_gpus": n, "skipped": f"only {have} device(s)"}), flush=True)
            continue
        total = args.views * n if args.weak else args.views
        pos, quat = synth.candidate_views(m, total, seed=42)
        g = capi.Group(list(range(n)))
        g.upload_map(m)
        g.set_option("deal_block", args.deal_block)
        caster = g.caster(kind, **cam)
        ev = g.evaluator(cfg["evaluator"] + "Evaluator")
        for _ in range(2):
            g.upload_blocks(d_idx, d_dist, d_obs)
            r = g.compute_gains(caster, ev, pos, quat, digests=False, set_digests=False)
        t_up, t_cast, k_ms = [], [], []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            g.upload_blocks(d_idx, d_dist, d_obs)
            t_mid = time.perf_counter()
            r = g.compute_gains(caster, ev, pos, quat, digests=False, set_digests=False)
            t_end = time.perf_counter()
            t_up.append(t_mid - t0)
            t_cast.append(t_end - t_mid)
            k_ms.append(g.last_kernel_ms)
        step = float(np.median(t_up) + np.median(t_cast))
        if n == 1:
            t1 = step
        print(json.dumps({
            "workload": args.workload, "evaluator": cfg["evaluator"], "scaling": "weak" if args.weak else "strong",
            "n_gpus": n, "global_views": total, "step_ms": 1e3 * step, "delta_ms": 1e3 * float(np.median(t_up)),
            "gains_call_ms": 1e3 * float(np.median(t_cast)), "slowest_kernel_ms": float(np.median(k_ms)),
            "views_per_s": total / step,
            "speedup_vs_1": (t1 / step if (t1 and not args.weak) else None),
            "gain_digest": hashlib.sha256(r["gain"].tobytes()).hexdigest()[:16],
            "n_visible_sum": int(r["n_visible"].sum())}), flush=True)
        g.close()


if __name__ == "__main__":
    main()
