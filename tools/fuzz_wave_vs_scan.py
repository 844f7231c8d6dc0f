#!/usr/bin/env python
"""Randomised cross-check of the wavefront kernel (both CTA widths, gains and ordered lists) against the
scan-order kernel, which tests/test_gpu_parity.py pins to the oracle: random voxel sizes, voxels per
side, camera intrinsics / mounting, casters, evaluators, bounding volumes, map offsets and segment
layouts.  Integer outputs and lists must agree bit for bit, gains to 1e-9 relative.
    python tools/fuzz_wave_vs_scan.py --cases 60 --seed 1"""
import argparse
import dataclasses
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mav_active_3d_planning_b200 import capi, synth  # noqa: E402


def one_case(rng, case, oracle=None, max_views=None):
    vs = float(rng.choice([0.08, 0.1, 0.15, 0.2, 0.3]))
    vps = int(rng.choice([8, 16, 16, 32]))
    m = synth.room_map(vs, vps=vps)
    shift = np.zeros(3, np.int32)
    if rng.random() < 0.3:
        shift = rng.integers(-400, 400, 3).astype(np.int32)
        m = dataclasses.replace(m, block_idx=m.block_idx + shift)
    ctx = capi.Context(0)
    ctx.upload_map(m)
    om = oracle.OracleMap.from_synth(m) if oracle is not None else None
    edits = rng.random() < 0.4
    if edits:
        # a cast first (builds the dense volume), then overwrites, removals and new blocks outside the box
        warm = ctx.caster("SimpleRayCaster", resolution_x=160, resolution_y=120, focal_length=100.0, ray_length=3.0)
        p0, q0 = synth.candidate_views(synth.room_map(vs, vps=vps), 1, seed=3)
        ctx.compute_gains(warm, ctx.evaluator("NaiveEvaluator"),
                          p0 + shift.astype(np.float64) * float(np.float32(vs) * np.float32(vps)), q0, digests=False)
        sel = rng.choice(m.n_blocks, max(1, m.n_blocks // 10), replace=False)
        nd = (m.dist[sel] + rng.normal(0, 0.05, m.dist[sel].shape)).astype(np.float32)
        no = (rng.random(m.observed[sel].shape) < 0.7).astype(np.uint8)
        rem = rng.choice(m.n_blocks, max(1, m.n_blocks // 20), replace=False)
        far = (m.block_idx[rng.choice(m.n_blocks, 2, replace=False)] + np.array([[40, 0, 0], [0, -35, 3]])).astype(np.int32)
        fd = np.full((2, vps ** 3), 0.5 * vs, np.float32)
        for mm in (ctx, om):
            if mm is None:
                continue
            mm.upload_blocks(m.block_idx[sel], nd, no)
            mm.remove_blocks(m.block_idx[rem])
            mm.upload_blocks(far, fd, np.ones((2, vps ** 3), np.uint8))
    ckind = str(rng.choice(["IterativeRayCaster", "SimpleRayCaster", "IterativeRayCaster"]))
    cam = dict(resolution_x=int(rng.choice([160, 320, 400])), resolution_y=int(rng.choice([120, 240, 300])),
               focal_length=float(rng.choice([100.0, 160.0, 250.0])), ray_length=float(rng.choice([3.0, 5.0, 7.5])),
               downsampling_factor=float(rng.choice([1.0, 1.5, 2.5])))
    if rng.random() < 0.5:
        q = rng.normal(size=4)
        q[0] = abs(q[0]) + 2.0
        cam.update(mounting_translation=tuple(rng.normal(0, 0.1, 3)), mounting_rotation=tuple(q / np.linalg.norm(q)))
    caster = ctx.caster(ckind, **cam)
    ekind = str(rng.choice(["VoxelTypeEvaluator", "NaiveEvaluator", "FrontierEvaluator", "VoxelWeightEvaluator"]))
    kw = {}
    if rng.random() < 0.6:
        lo = rng.random(3) * 4.0
        hi = np.array(m.extent) - rng.random(3) * 4.0
        off = shift.astype(np.float64) * float(np.float32(vs) * np.float32(vps))
        kw["bounding_volume"] = dict(x_min=lo[0] + off[0], x_max=hi[0] + off[0], y_min=lo[1] + off[1],
                                     y_max=hi[1] + off[1], z_min=lo[2] + off[2], z_max=hi[2] + off[2],
                                     rotation=float(rng.choice([0.0, 0.0, 17.0])))
    if ekind in ("FrontierEvaluator", "VoxelWeightEvaluator"):
        kw.update(accurate_frontiers=bool(rng.random() < 0.4), surface_frontiers=bool(rng.random() < 0.6),
                  checking_distance=float(rng.choice([1.0, 1.0, 1.0, 2.0])))
    if ekind == "VoxelWeightEvaluator":
        kw.update(frontier_voxel_weight=float(rng.choice([0.0, 0.5, 1.0])), new_voxel_weight=0.05)
    ev = ctx.evaluator(ekind, **kw)
    n_views = int(rng.choice([1, 5, 24, 160, 320]))
    if max_views:
        n_views = min(n_views, max_views)
    pos, quat = synth.candidate_views(synth.room_map(vs, vps=vps), n_views, seed=int(rng.integers(1 << 30)))
    pos = pos + shift.astype(np.float64) * float(np.float32(vs) * np.float32(vps))
    if rng.random() < 0.4 and n_views >= 5:   # multi-view segments
        per = int(rng.choice([2, 3, 5]))
        n_seg = n_views // per
        seg = np.repeat(np.arange(n_seg, dtype=np.int32), per)
        pos, quat = pos[:n_seg * per], quat[:n_seg * per]
        args = (pos, quat, seg, n_seg)
    else:
        args = (pos, quat)
    sd = ekind != "VoxelWeightEvaluator"
    ctx.set_option("cast_kernel", 1)
    ref = ctx.compute_gains(caster, ev, *args, digests=False, set_digests=sd)
    ctx.set_option("cast_kernel", 2)
    bad = []
    if oracle is not None:
        # the scan-order kernel itself against the CPU oracle on this random configuration
        oc = om.caster(ckind, **cam)
        want = oc.compute_gains(oracle.eval_params(ekind, **kw), *args, n_threads=8)
        for k in ("n_samples", "n_visible") + (("set_digest",) if sd else ()):
            if not np.array_equal(ref[k], want[k]):
                bad.append(f"oracle {k}")
        if not np.allclose(ref["gain"], want["gain"], rtol=1e-5, atol=1e-12):
            bad.append("oracle gain")
    for width in (1, 2, 3):
        ctx.set_option("wave_cta", width)
        got = ctx.compute_gains(caster, ev, *args, digests=False, set_digests=sd)
        for k in ("n_samples", "n_visible") + (("set_digest",) if sd else ()):
            if not np.array_equal(got[k], ref[k]):
                bad.append(f"{k} width {width}")
        if not np.allclose(got["gain"], ref["gain"], rtol=1e-9, atol=1e-12):
            bad.append(f"gain width {width}")
    ctx.set_option("wave_cta", 0)
    # the kernel chain (a3d_split.cu): gains and counts only; VoxelWeight gains are summed in another
    # order than the scan-order kernel's
    ctx.set_option("cast_kernel", 0)
    got = ctx.compute_gains(caster, ev, *args, digests=False, set_digests=False)
    if ctx.last_cast_kernel == 3:
        for k in ("n_samples", "n_visible"):
            if not np.array_equal(got[k], ref[k]):
                bad.append(f"{k} chain")
        if not np.allclose(got["gain"], ref["gain"], rtol=1e-9, atol=1e-12):
            bad.append("gain chain")
    elif vps in (8, 16, 32) and not edits:
        bad.append(f"kernel chain not used (kernel {ctx.last_cast_kernel})")
    # ordered list of the first segment's views
    first = slice(0, 1) if len(args) == 2 else slice(0, int(np.sum(args[2] == 0)))
    ctx.set_option("cast_kernel", 1)
    l_ref, _, g_ref = ctx.visible_voxels(caster, pos[first], quat[first], evaluator=ev)
    for kernel in (2, 0):   # wavefront kernel's emission; default = the kernel chain's where it applies
        ctx.set_option("cast_kernel", kernel)
        l_got, _, g_got = ctx.visible_voxels(caster, pos[first], quat[first], evaluator=ev)
        if kernel == 2 and ctx.last_cast_kernel != 2:
            bad.append("list call did not use the wavefront kernel")
        if l_ref.shape != l_got.shape or not np.array_equal(l_ref, l_got):
            bad.append(f"ordered list (kernel {ctx.last_cast_kernel})")
        if not np.isclose(g_ref, g_got, rtol=1e-9, atol=1e-12):
            bad.append(f"list gain (kernel {ctx.last_cast_kernel})")
    ctx.set_option("cast_kernel", 0)
    # stored lists of the whole batch (two passes of the kernel chain, the second writing the keys straight
    # into the arena; the scan-order kernel where the chain does not apply): every segment's stored list
    # against the scan-order kernel's ordered list of its views (Naive / Frontier store the list after the
    # erasure of observed entries), counts and gains against the batch call
    n_seg = len(ref["gain"])
    if n_seg <= 64:
        seg_of = np.arange(len(pos), dtype=np.int32) if len(args) == 2 else args[2]
        keys = np.arange(1, n_seg + 1, dtype=np.uint64)
        for kernel in (0, 1):
            ctx.set_option("cast_kernel", kernel)
            store = capi.Store(ctx)
            st = store.compute_gains(caster, ev, keys, pos, quat, view_segment=seg_of)
            erases = ekind in ("NaiveEvaluator", "FrontierEvaluator")
            if not erases and not np.array_equal(st["n_visible"], ref["n_visible"]):
                bad.append(f"stored n_visible (kernel option {kernel})")
            if not np.array_equal(st["n_samples"], ref["n_samples"]):
                bad.append(f"stored n_samples (kernel option {kernel})")
            if not np.allclose(st["gain"], ref["gain"], rtol=1e-9, atol=1e-12):
                bad.append(f"stored gain (kernel option {kernel})")
            ctx.set_option("cast_kernel", 1)
            for sgm in range(min(n_seg, 4)):
                sel = seg_of == sgm
                want_l = ctx.visible_voxels(caster, pos[sel], quat[sel], evaluator=ev)[0] if sel.any() else np.zeros((0, 3))
                if erases and len(want_l):
                    want_l = want_l[ctx.is_observed(want_l) == 0]
                got_l = store.get(keys[sgm])
                if got_l.shape != want_l.shape or not np.array_equal(got_l, want_l):
                    bad.append(f"stored list of segment {sgm} (kernel option {kernel})")
            store.close()
        ctx.set_option("cast_kernel", 0)
    ctx.close()
    return dict(case=case, vs=vs, vps=vps, edits=bool(edits), caster=ckind, evaluator=ekind, kw={k: (v if not isinstance(v, dict) else "bv") for k, v in kw.items()},
                cam=cam["resolution_x"], shift=shift.tolist(), views=n_views, segments=len(ref["gain"]),
                samples=int(ref["n_samples"].sum()), list_len=int(len(l_ref)), bad=bad)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--oracle", action="store_true",
                    help="also check the scan-order kernel against the CPU oracle (at most 5 views per case)")
    args = ap.parse_args()
    oracle = None
    if args.oracle:
        from oracle import oracle_py as oracle  # noqa: F811  (test infrastructure: this tool is a checker)
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    n_bad = 0
    total = 0
    for c in range(args.cases):
        r = one_case(rng, c, oracle=oracle, max_views=5 if oracle else None)
        total += r["samples"]
        if r["bad"]:
            n_bad += 1
            print(json.dumps(r), flush=True)
    print(json.dumps(dict(cases=args.cases, seed=args.seed, mismatching_cases=n_bad, samples_checked=total,
                          seconds=round(time.time() - t0, 1))))
    sys.exit(1 if n_bad else 0)


if __name__ == "__main__":
    main()
