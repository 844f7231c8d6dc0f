#!/usr/bin/env python
"""bench.py — candidate views evaluated per second on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--views V]

Workload (config.workload): BASELINE config 2 — synthetic 20×20×10 m room at 0.1 m voxels (vps 16),
4096 candidate segments per GPU (one view each, RRT*-style poses), IterativeRayCaster
320×240 f=160 (79×65 rays, 6 sections) + VoxelTypeEvaluator defaults — plus, per planning step,
the 1 % map-block delta of BASELINE config 5 (re-uploaded through the C ABI; broadcast over NCCL
from rank 0 when N > 1) and the gather of the per-segment gains.

One "step" = apply the block delta + evaluate every candidate segment once.  Prints ONE JSON line.
`value` is measured with poses resident in HBM (a3d_compute_gains_dev); `e2e` goes through the
host-buffer C-ABI call (a3d_map_upload_blocks + a3d_compute_gains) from pinned host memory.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CAM = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0)
WORKLOAD = ("cfg2: room 20x20x10 m @0.1 m voxels (vps 16), IterativeRayCaster 320x240 f=160 L=5 "
            "(79x65 rays) + VoxelTypeEvaluator, one view per segment, + 1% block delta per step (cfg5)")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--views", type=int, default=4096, help="candidate segments per GPU")
    ap.add_argument("--global-views", type=int, default=0,
                    help="strong scaling: this many candidate segments in total, the SAME list for every N "
                         "(gain_digest in the JSON line is then equal for every N)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="views in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(n_views, seed=42):
    from mav_active_3d_planning_b200 import synth
    m = synth.room_map(0.1)
    pos, quat = synth.candidate_views(m, n_views, seed=seed)
    # cfg5 delta: 1 % of blocks re-sent each step (distances jittered, observed unchanged)
    rng = np.random.default_rng(3)
    n_delta = max(1, m.n_blocks // 100)
    sel = np.sort(rng.choice(m.n_blocks, n_delta, replace=False))
    d_idx = np.ascontiguousarray(m.block_idx[sel])
    d_dist = (m.dist[sel] + rng.normal(0, 0.002, m.dist[sel].shape)).astype(np.float32)
    d_obs = np.ascontiguousarray(m.observed[sel])
    return m, pos, quat, (d_idx, d_dist, d_obs)


def cpu_impl():
    """("ref", "reference") when oracle/_ref/liba3d_ref.so — the reference's own unmodified sources
    behind the orc_* API, built in the container that has /root/reference — travelled with the
    snapshot; else ("oracle", "port"), the restatement."""
    from oracle import oracle_py as O
    return ("ref", "reference") if O.have_ref() else ("oracle", "port")


def cpu_oracle_rate(m, delta, pos, quat, n_sample, n_threads):
    """views/s of the reference's CPU implementation of the path on the first n_sample views."""
    from oracle import oracle_py as O
    om = O.OracleMap.from_synth(m, impl=cpu_impl()[0])
    om.upload_blocks(*delta)
    oc = om.caster("IterativeRayCaster", **CAM)
    ep = O.eval_params("VoxelTypeEvaluator")
    oc.compute_gains(ep, pos[:2], quat[:2])  # warm
    t0 = time.perf_counter()
    r = oc.compute_gains(ep, pos[:n_sample], quat[:n_sample], n_threads=n_threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, r


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref: its
    unmodified sources, one planner + map copy per host thread; falls back to the restatement in
    oracle/ when that library is absent), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m, pos, quat, delta = make_workload(max(args.views, 256))
    cores = os.cpu_count() or 1
    n_sample = args.cpu_sample or min(args.views, max(256, 16 * cores))
    from oracle import oracle_py as O
    impl, kind = cpu_impl()
    om = O.OracleMap.from_synth(m, impl=impl)
    om.upload_blocks(*delta)
    oc = om.caster("IterativeRayCaster", **CAM)
    ep = O.eval_params("VoxelTypeEvaluator")
    for _ in range(args.warmup):
        oc.compute_gains(ep, pos[:n_sample], quat[:n_sample], n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oc.compute_gains(ep, pos[:n_sample], quat[:n_sample], n_threads=cores)
    dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    sample = f"first {n_sample} of {args.views} candidate views per step, {cores} threads over views"
    print(json.dumps({
        "impl": "reference", "metric": "candidate_views_per_sec", "value": value, "unit": "views/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 step / f32 lookup", "data": "synthetic",
        "config": {"workload": WORKLOAD, "views_per_gpu": args.views, "views_timed": n_sample},
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mav_active_3d_planning_b200 import capi, scheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # every rank evaluates its own shard of the global candidate list: --views per GPU (weak scaling) or
    # --global-views in total (strong scaling, the same list whatever N is)
    strong = args.global_views > 0
    n_global = args.global_views if strong else args.views * world
    m, pos_all, quat_all, (d_idx, d_dist, d_obs) = make_workload(n_global)
    lo, hi = scheduler.shard_bounds(n_global, world, rank)
    n_seg = hi - lo
    pos = np.ascontiguousarray(pos_all[lo:hi])
    quat = np.ascontiguousarray(quat_all[lo:hi])

    ctx = capi.Context(local)
    ctx.upload_map(m)
    caster = ctx.caster("IterativeRayCaster", **CAM)
    ev = ctx.evaluator("VoxelTypeEvaluator")
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- device-resident inputs/outputs -------------------------------------------------------
    t_pos = torch.from_numpy(pos).to(dev)
    t_quat = torch.from_numpy(quat).to(dev)
    t_first = torch.arange(n_seg + 1, dtype=torch.int32, device=dev)
    t_gain = torch.zeros(n_seg, dtype=torch.float64, device=dev)
    t_nvis = torch.zeros(n_seg, dtype=torch.int64, device=dev)
    t_nsamp = torch.zeros(n_seg, dtype=torch.int64, device=dev)
    nv = m.vps ** 3
    n_delta = d_idx.shape[0]
    dist_bytes = n_delta * nv * 4
    # two receive buffers: the delta of planning step k+1 is broadcast (NCCL, its own stream) while
    # step k casts; step k+1 waits for it before applying it
    payloads = [torch.zeros(dist_bytes + n_delta * nv, dtype=torch.uint8, device=dev) for _ in range(2)]
    if rank == 0:
        for pl in payloads:
            pl[:dist_bytes] = torch.from_numpy(d_dist.view(np.uint8).reshape(-1)).to(dev)
            pl[dist_bytes:] = torch.from_numpy(d_obs.reshape(-1)).to(dev)
    pending = {"work": None, "buf": 0}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.cuda.synchronize()

    def step_device():
        payload = payloads[pending["buf"]]
        if world > 1:
            if pending["work"] is None:          # first step: nothing was prefetched
                dist.broadcast(payload, src=0)
            else:
                pending["work"].wait()           # this step's delta, sent while the previous step cast
        ctx.upload_blocks_dev(d_idx, payload.data_ptr(), payload.data_ptr() + dist_bytes)
        if world > 1:                            # next step's delta travels during this step's cast
            pending["buf"] ^= 1
            pending["work"] = dist.broadcast(payloads[pending["buf"]], src=0, async_op=True)
        ctx.compute_gains_dev(caster, ev, n_seg, t_pos.data_ptr(), t_quat.data_ptr(),
                              t_first.data_ptr(), n_seg, t_gain.data_ptr(), t_nvis.data_ptr(),
                              t_nsamp.data_ptr(), 0)
        return scheduler.gather_gains(t_gain, n_global)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step_device()
        barrier()
        launches0 = ctx.kernel_launches
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        kernel_ms = []
        stage_ms = []
        barrier()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)  # evict L2 between timed iterations (outside the events)
            ev_a[i].record(stream)
            step_device()
            ev_b[i].record(stream)
            kernel_ms.append(None)
            ev_b[i].synchronize()
            kernel_ms[-1] = ctx.last_kernel_ms
            stage_ms.append(ctx.last_stage_ms)
        barrier()
        launches = ctx.kernel_launches - launches0
        total_ms = sum(a.elapsed_time(b) for a, b in zip(ev_a, ev_b))
        t_ms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        total_ms = float(t_ms.item())

        # ---- e2e: host buffers through the C ABI ---------------------------------------------
        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t, t.numpy()
        keep = []
        hp = [pinned(x) for x in (pos, quat, np.arange(n_seg, dtype=np.int32), d_dist, d_obs)]
        keep.extend(hp)
        h_pos, h_quat, h_seg, h_ddist, h_dobs = [x[1] for x in hp]

        def step_host():
            if world > 1:
                dist.broadcast(payloads[0], src=0)  # delta still travels rank0 → all on the device side
            ctx.upload_blocks(d_idx, h_ddist, h_dobs)
            r = ctx.compute_gains(caster, ev, h_pos, h_quat, h_seg, n_seg, digests=False,
                                  set_digests=False)
            if world > 1:   # the gains of every rank, gathered (host result → device → NCCL)
                t_gain_h.copy_(torch.from_numpy(r["gain"]), non_blocking=False)
                r["all_gains"] = scheduler.gather_gains(t_gain_h, n_global)
            return r
        t_gain_h = torch.zeros(n_seg, dtype=torch.float64, device=dev)
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res_host = step_host()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e_s = float(t_e.item())
        clocks = sampler.stop() if rank == 0 else None

    # ---- derived numbers ----------------------------------------------------------------------
    nsamp = t_nsamp.cpu().numpy().astype(np.int64)
    nvis = t_nvis.cpu().numpy().astype(np.int64)
    gains = t_gain.cpu().numpy()
    assert np.array_equal(res_host["n_samples"].astype(np.int64), nsamp), "host/dev paths disagree"
    S, V = int(nsamp.sum()), int(nvis.sum())
    alg_bytes = 32 * S + 32 * V + (56 + 8) * n_seg       # SURVEY.md §8(d), VoxelType fold
    lookups = 8 * S + 8 * V
    k_ms = float(np.mean(kernel_ms))
    chain = ctx.last_cast_kernel == 3
    stages = {}
    if chain and stage_ms and all(len(x) == 5 for x in stage_ms):
        names = ["k_split_views", "k_split_mark", "k_split_leaf<deferred marking parts>", "k_split_leaf", "k_split_fix"]
        stages = {n: float(np.mean([x[k] for x in stage_ms])) for k, n in enumerate(names)}
    peak, peak_src = measured_peak_hbm()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    steps_views = n_global * args.steps
    import hashlib
    all_gains = step_device().cpu().numpy()     # gathered vector, identical on every rank
    gain_digest = hashlib.sha256(np.ascontiguousarray(all_gains).tobytes()).hexdigest()[:16]
    if pending["work"] is not None:
        pending["work"].wait()
        torch.cuda.synchronize()
    value = steps_views / (total_ms * 1e-3)
    h2d = h_pos.nbytes + h_quat.nbytes + h_seg.nbytes + h_ddist.nbytes + h_dobs.nbytes + d_idx.nbytes
    d2h = n_seg * 8 * 3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    out = {
        "metric": "candidate_views_per_sec", "value": value, "unit": "views/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64 step / f32 lookup", "data": "synthetic",
        "config": {"workload": WORKLOAD, "views_per_gpu": n_seg, "global_views": n_global,
                   "map_blocks": m.n_blocks, "delta_blocks_per_step": int(n_delta),
                   "samples_per_view": S / n_seg, "visible_per_view": V / n_seg,
                   "l2": "512 MiB buffer written between timed iterations (L2 flush)",
                   "parallelism": f"segments sharded over {world} rank(s), map replicated"},
        "voxel_lookups_per_sec": lookups * world / (total_ms / args.steps * 1e-3),
        "kernel_ms": k_ms,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                     "traffic_source": "sum of dram__bytes_read+write over the kernels of one cast in an ncu --set full "
                                       "capture of this workload at 4096 views (profiles/traffic.json), not measured "
                                       "in this run",
                     # one "launch" = the cast of the batch: the kernel chain of a3d_split.cu (its kernels'
                     # shares below), or the single wavefront / scan-order kernel
                     "kernel": ("kernel chain k_split_views+mark+leaf<deferred>+leaf+fix" if chain else
                                "k_cast_wave" if ctx.last_cast_kernel == 2 else "k_cast_gain") +
                     "<Iterative,VoxelType>",
                     "kernel_ms_by_stage": stages,
                     "algorithmic_bytes_per_launch": alg_bytes},
        "e2e": {"value": steps_views / e2e_s, "unit": "views/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "gain_checksum": float(gains.sum()), "gain_digest": gain_digest,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sample = args.cpu_sample or min(n_seg, max(256, 16 * cores))
        rate_all, r = cpu_oracle_rate(m, (d_idx, d_dist, d_obs), pos, quat, n_sample, cores)
        n1 = min(n_sample, 24)
        rate_1, _ = cpu_oracle_rate(m, (d_idx, d_dist, d_obs), pos, quat, n1, 1)
        ok = bool(np.array_equal(r["n_samples"].astype(np.int64), nsamp[:n_sample]) and
                  np.allclose(r["gain"], gains[:n_sample], rtol=1e-5, atol=0))
        out["cpu_baseline"] = {
            "value": rate_all, "unit": "views/s", "cores": cores, "kind": cpu_impl()[1],
            "sample": f"first {n_sample} of {n_seg} candidate views, {cores} threads over views",
            "single_thread_value": rate_1, "single_thread_sample": f"first {n1} views",
            "matches_gpu": ok}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
