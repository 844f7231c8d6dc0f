#!/usr/bin/env python
"""Scaling sweep on one GPU (BASELINE config 5 shape): views/s, lookups/s and algorithmic GB/s of
a3d_compute_gains_dev for batch sizes 2^10 … 2^20, plus the other BASELINE configs as single points.
    python tools/sweep.py --workload cfg2 --views 1024,4096,16384,65536,262144,1048576
    python tools/sweep.py --workload cfg3 --views 4096        (50x50x20 m clutter map at 0.05 m, Frontier)
    python tools/sweep.py --workload cfg1|cfg4 --views 64
One JSON line per point."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mav_active_3d_planning_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--views", default="1024,4096,16384,65536")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--evaluator", default=None, help="override the workload's evaluator, e.g. VoxelWeight")
    ap.add_argument("--cast-kernel", type=int, default=0, help="0 auto, 1 scan-order, 2 wavefront, 3 kernel chain")
    ap.add_argument("--wave-cta", type=int, default=0, help="0 auto, 1 two-warp CTAs, 2 eight-warp CTAs")
    args = ap.parse_args()
    cfg = dict(synth.CONFIGS[args.workload])
    if args.evaluator:
        cfg["evaluator"] = args.evaluator
    t0 = time.time()
    if cfg["map"] == "room":
        m = synth.room_map(cfg["voxel_size"])
    else:
        m = synth.clutter_map(cfg["voxel_size"])
    t_map = time.time() - t0
    ctx = capi.Context(0)
    t0 = time.time()
    ctx.upload_map(m)
    t_up = time.time() - t0
    kind = "IterativeRayCaster" if cfg["caster"] == 1 else "SimpleRayCaster"
    caster = ctx.caster(kind, resolution_x=cfg["resolution"][0], resolution_y=cfg["resolution"][1],
                        focal_length=cfg["focal_length"], ray_length=cfg["ray_length"])
    ev = ctx.evaluator(cfg["evaluator"] + "Evaluator")
    ctx.set_option("cast_kernel", args.cast_kernel)
    ctx.set_option("wave_cta", args.wave_cta)
    dev = torch.device("cuda", 0)
    peak = 6555.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for n in [int(v) for v in args.views.split(",")]:
        pos, quat = synth.candidate_views(m, n, seed=42)
        t_pos, t_quat = torch.from_numpy(pos).to(dev), torch.from_numpy(quat).to(dev)
        t_first = torch.arange(n + 1, dtype=torch.int32, device=dev)
        t_gain = torch.zeros(n, dtype=torch.float64, device=dev)
        t_nvis = torch.zeros(n, dtype=torch.int64, device=dev)
        t_nsamp = torch.zeros(n, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        ms = []
        for r in range(args.reps + 1):
            ctx.compute_gains_dev(caster, ev, n, t_pos.data_ptr(), t_quat.data_ptr(), t_first.data_ptr(), n,
                                  t_gain.data_ptr(), t_nvis.data_ptr(), t_nsamp.data_ptr(), 0)
            k = ctx.last_kernel_ms
            if r:
                ms.append(k)
                stages = getattr(ctx, "last_stage_ms", [])
        S, V = int(t_nsamp.sum().item()), int(t_nvis.sum().item())
        # (lower bound for Frontier; VoxelWeight: centre state + one TSDF weight)
        fold = {"VoxelType": 8 * V, "Naive": V, "Frontier": V, "VoxelWeight": 9 * V}[cfg["evaluator"]]
        lookups = 8 * S + fold
        best = min(ms)
        alg = 32 * S + 4 * fold + 64 * n
        print(json.dumps({"workload": args.workload, "evaluator": cfg["evaluator"], "views": n, "map_blocks": m.n_blocks,
                          "map_gen_s": round(t_map, 1), "map_upload_s": round(t_up, 2),
                          "kernel_ms": best, "views_per_s": n / (best * 1e-3),
                          "lookups_per_s": lookups / (best * 1e-3),
                          "samples_per_view": S / n, "algorithmic_GBps": alg / (best * 1e-3) / 1e9,
                          "frac_of_hbm": alg / (best * 1e-3) / 1e9 / peak,
                          "kernel": ctx.last_cast_kernel, "stage_ms": [round(x, 3) for x in stages],
                          "gain_sum": float(t_gain.sum().item())}), flush=True)


if __name__ == "__main__":
    main()
