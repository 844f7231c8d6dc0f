"""C++ host mirror of the reference module API (mav_active_3d_planning_b200/host): YAML → ParamMap
exactly like ModuleFactoryROS, registry of type strings (CPU); modules built from YAML evaluating
segments through the C ABI, against the oracle (GPU)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "mav_active_3d_planning_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(HOST, "test_host")


def run_config(yaml, ns=None):
    out = subprocess.run([EXE, "config", yaml] + ([ns] if ns else []), capture_output=True, text=True,
                         check=True).stdout
    params, types, registered = {}, {}, []
    cur = None
    for line in out.splitlines():
        t = line.split()
        if t[0] == "namespace":
            cur = t[1]
            types[cur] = t[3]
            params[cur] = {}
        elif t[0] == "param":
            params[cur][t[1]] = line.split(" = ", 1)[1]
        elif t[0] == "registered":
            registered = [x.strip() for x in line[len("registered "):].split(",")]
    return params, types, registered


def test_yaml_factory_builds_param_maps_like_the_ros_factory(built):
    p, t, reg = run_config(os.path.join(GOLD, "planner_naive_iterative.yaml"))
    ev, sm = "/trajectory_evaluator", "/trajectory_evaluator/sensor_model"
    assert t[ev] == "NaiveEvaluator" and t[sm] == "IterativeRayCaster"
    # direct scalar children only; bools as 1/0; quotes stripped; param_namespace recorded
    assert p[ev] == {"type": "NaiveEvaluator", "clear_from_parents": "0", "visualize_sensor_view": "1",
                     "bounding_volume_args": "/target_bounding_volume", "param_namespace": ev}
    assert p[sm]["downsampling_factor"] == "1.0" and p[sm]["resolution_x"] == "320"
    assert p[sm]["param_namespace"] == sm and "cost_computer" not in p[ev]
    # every type string of the path resolves
    for name in ("SimpleRayCaster", "IterativeRayCaster", "VoxelTypeEvaluator", "NaiveEvaluator",
                 "FrontierEvaluator", "BoundingVolume", "VoxbloxMap"):
        assert name in reg
    p, t, _ = run_config(os.path.join(GOLD, "planner_naive_iterative.yaml"), "/target_bounding_volume")
    assert p["/target_bounding_volume"]["x_max"] == "18.0"


def test_yaw_planning_config_parses_like_the_reference_layout(built):
    """following_evaluator nests under the yaw-planning evaluator exactly like in
    app_reconstruction/cfg/planners/reconstruction_planner.yaml:59-100."""
    yml = os.path.join(GOLD, "planner_yaw_continuous.yaml")
    p, t, reg = run_config(yml)
    assert t["/trajectory_evaluator"] == "ContinuousYawPlanningEvaluator"
    assert p["/trajectory_evaluator"]["n_directions"] == "12" and p["/trajectory_evaluator"]["n_sections_fov"] == "3"
    for name in ("SimpleYawPlanningEvaluator", "ContinuousYawPlanningEvaluator"):
        assert name in reg
    p, t, _ = run_config(yml, "/trajectory_evaluator/following_evaluator")
    assert t["/trajectory_evaluator/following_evaluator"] == "NaiveEvaluator"
    assert t["/trajectory_evaluator/following_evaluator/sensor_model"] == "IterativeRayCaster"
    assert p["/trajectory_evaluator/following_evaluator/sensor_model"]["downsampling_factor"] == "2.0"


def test_reference_planner_configs_load_unchanged(built):
    cfg = "/root/reference/active_3d_planning_app_reconstruction/cfg/planners"
    if not os.path.isdir(cfg):
        pytest.skip("reference checkout not present on this machine")
    p, t, _ = run_config(os.path.join(cfg, "example_config.yaml"))
    assert t["/trajectory_evaluator"] == "NaiveEvaluator"
    assert t["/trajectory_evaluator/sensor_model"] == "IterativeRayCaster"
    assert p["/trajectory_evaluator/sensor_model"]["downsampling_factor"] == "3.0"
    assert p["/trajectory_evaluator"]["clear_from_parents"] == "0"
    p, t, _ = run_config(os.path.join(cfg, "reconstruction_planner.yaml"))
    assert t["/trajectory_evaluator"] == "RRTStarEvaluatorAdapter"   # out-of-scope wrapper, still parsed
    for f in os.listdir(cfg):
        subprocess.run([EXE, "config", os.path.join(cfg, f)], capture_output=True, check=True)


def test_updater_chain_and_update_tree_without_a_device(built):
    """evaluator_updater chain from YAML (ConstrainedUpdater -> UpdateAll -> PruneDirect, parameters and
    semantics of constrained_updater.cpp:17-30, update_all.cpp:16-27, prune_direct.cpp:33-39) through
    TrajectoryEvaluator::updateTree on a hand-made tree, no GPU involved: only the segments the
    constraint lets through are evaluated again, a dropped segment takes its subtree with it, the
    others keep their numbers; updateSegment is the one-element case."""
    out = subprocess.run([EXE, "tree", os.path.join(GOLD, "planner_update_tree_cpu.yaml")], capture_output=True,
                         text=True, check=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        t = line.split()
        if t[0] == "node":
            if t[1] == "root":
                cur = {}
                blocks.append(cur)
            cur[t[1]] = tuple(float(x) for x in t[2:5])
        elif t[0] == "removed":
            removed, calls = int(t[1]), int(t[3])
        elif t[0] == "single":
            single = int(t[1])
    before, after = blocks
    # gain = x of the end point, cost = 1 (TestSegmentLength at yaw 0), value = gain - 100 cost
    assert before == {"root": (0, 0, 0), "1": (5, 1, -95), "1.11": (6, 1, -94), "1.11.111": (1, 1, -99),
                      "2": (2, 1, -98), "2.21": (9, 1, -91), "3": (40, 1, -60), "3.31": (8, 1, -92)}
    # within 10 m of the robot and above gain 1.5: segments 1 and 2 -> evaluated again with the doubled
    # factor (10 and 4); 2 falls below PruneDirect's 5 and goes with its child; the rest is untouched
    assert calls == 2 and removed == 2
    assert after == {"root": (0, 0, 0), "1": (10, 1, -90), "1.11": (6, 1, -94), "1.11.111": (1, 1, -99),
                     "3": (40, 1, -60), "3.31": (8, 1, -92)}
    assert single == 1


def test_yaw_planning_updaters_without_a_device(built):
    """YawPlanningUpdater -> YawPlanningUpdateAdapter -> (an updater that knows nothing about yaw planning)
    from YAML, semantics of yaw_planning_updaters.cpp:21-41,56-119: the orientations of the segments in
    reach are evaluated again as one batch and the best one becomes the segment; segments out of reach keep
    their orientation; the adapter hands the ACTIVE copies on and copies gain / cost / value back (the
    copies never had a cost or value, so the segment's become 0 like in the reference)."""
    out = subprocess.run([EXE, "yawtree", os.path.join(GOLD, "planner_yaw_updaters_cpu.yaml")], capture_output=True,
                         text=True, check=True).stdout
    blocks, cur, info = [], None, {}
    for line in out.splitlines():
        t = line.split()
        if t[0] == "gain_calls":
            info["first_calls"] = int(t[1])
            cur = {}
            blocks.append(cur)
        elif t[0] == "removed":
            info.update(removed=int(t[1]), calls=int(t[3]), batches=int(t[5]))
            cur = {}
            blocks.append(cur)
        elif t[0] == "node":
            cur[int(t[1])] = tuple(float(x) for x in t[2:8])
    before, after = blocks
    # 4 segments x 4 orientations; gain 10 + 5 cos(yaw - 0): yaw 0 wins everywhere
    assert info["first_calls"] == 16
    for k in (1, 2, 21, 3):
        assert before[k] == (15.0, 1.0, -85.0, 0.0, 0.0, 15.0)
    # phase -> pi/2, factor 2: segments 1 and 2 end within 10 m of the robot -> 2 x 4 gains in the update,
    # orientation 1 (yaw pi/2) now scores 2 * 15 = 30; the adapter's follower triples the active copy's gain
    assert info == {"first_calls": 16, "removed": 0, "calls": 8, "batches": 1}
    for k in (1, 2):
        gain, cost, value, yaw, active, copy_gain = after[k]
        assert (gain, cost, value, active, copy_gain) == (90.0, 0.0, 0.0, 1.0, 90.0)
        assert abs(yaw - np.pi / 2) < 1e-9
    for k in (21, 3):   # out of reach: orientation kept, only the adapter's follower acts
        assert after[k] == (45.0, 0.0, 0.0, 0.0, 0.0, 45.0)


def write_inputs(tmp, m, pos, quat, pts_per_seg, dt_ns):
    mp, sp = os.path.join(tmp, "map.bin"), os.path.join(tmp, "segments.bin")
    with open(mp, "wb") as f:
        f.write(struct.pack("<ifi", m.vps, m.voxel_size, m.n_blocks))
        f.write(np.ascontiguousarray(m.block_idx, np.int32).tobytes())
        f.write(np.ascontiguousarray(m.dist, np.float32).tobytes())
        f.write(np.ascontiguousarray(m.observed, np.uint8).tobytes())
        f.write(np.ascontiguousarray(m.weight, np.float32).tobytes())
    n_seg = len(pos) // pts_per_seg
    with open(sp, "wb") as f:
        f.write(struct.pack("<ii", n_seg, pts_per_seg))
        for s in range(n_seg):
            for k in range(pts_per_seg):
                i = s * pts_per_seg + k
                f.write(struct.pack("<q", k * dt_ns))
                f.write(np.concatenate([pos[i], quat[i]]).astype(np.float64).tobytes())
    return mp, sp


def parse_out(path):
    res = {"single": {}, "refold": {}, "batch": {}, "stored": {}, "update": {}}
    for line in open(path):
        t = line.split()
        if not t or t[0] == "#":
            continue
        if t[0] == "grid":
            res["grid"] = tuple(int(x) for x in t[1:4])
        elif t[0] == "single":
            res["single"][int(t[1])] = (int(t[2]), float(t[3]), int(t[4]), int(t[5]))
        elif t[0] == "refold":
            res["refold"][int(t[1])] = int(t[2])
        elif t[0] == "batch":
            res["batch"][int(t[1])] = (int(t[2]), float(t[3]))
        elif t[0] == "stored":
            res["stored"][int(t[1])] = (int(t[2]), int(t[3]), int(t[4]))
        elif t[0] == "update":
            res["update"][int(t[1])] = (int(t[2]), float(t[3]))
        elif t[0] == "update_single":
            res["update_single"] = float(t[2])
        elif t[0] == "sensor":
            res["sensor"] = (int(t[2]), int(t[3]))
        elif t[0] == "center":
            res["center"] = [float(x) for x in t[1:4]]
            res["state"], res["observed"], res["voxel_size"] = int(t[5]), int(t[7]), float(t[9])
    return res


@pytest.mark.gpu
def test_modules_from_yaml_match_oracle_naive_iterative(built, room02, tmp_path):
    from mav_active_3d_planning_b200 import synth
    from oracle import oracle_py as O
    pos, quat = synth.candidate_views(room02, 6, seed=3)
    mp, sp = write_inputs(str(tmp_path), room02, pos, quat, 1, 0)
    out = os.path.join(str(tmp_path), "out.txt")
    subprocess.run([EXE, "gains", os.path.join(GOLD, "planner_naive_iterative.yaml"), mp, sp, out],
                   check=True, timeout=300)
    r = parse_out(out)
    om = O.OracleMap.from_synth(room02)
    cam = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0)
    oc = om.caster("IterativeRayCaster", **cam)
    bv = dict(x_min=2, x_max=18, y_min=2, y_max=18, z_min=0.5, z_max=7)
    ep = O.eval_params("NaiveEvaluator", bounding_volume=bv)
    want = oc.compute_gains(ep, pos, quat)
    assert r["grid"] == (oc.c_res_x, oc.c_res_y, oc.c_n_sections)
    for s in range(6):
        ok, gain, n_list, dig = r["single"][s]
        assert ok == 1 and gain == want["gain"][s]
        # Naive erases the observed entries from the stored list: what is left is what it counted
        assert n_list == int(want["gain"][s])
        assert r["refold"][s] == 1
        assert r["batch"][s] == (1, want["gain"][s])
        # device-resident lists: same gains, list size before the fold's erasure, re-fold unchanged map
        assert r["stored"][s] == (1, 1, int(want["n_visible"][s]))
        assert r["update"][s] == (1, want["gain"][s])
    assert r["update_single"] == want["gain"][0]
    lst, _ = oc.visible_voxels(pos[:1], quat[:1])
    assert r["sensor"] == (len(lst), O.digest(lst))
    assert r["center"] == list(om.voxel_center(pos[:1])[0])
    assert r["state"] == om.voxel_state(pos[:1])[0] and r["observed"] == om.is_observed(pos[:1])[0]
    assert r["voxel_size"] == om.voxel_size


@pytest.mark.gpu
def test_modules_from_yaml_match_oracle_voxeltype_parents(built, room02, tmp_path):
    """sampling_time > 0 (several views per segment), mounted camera, clear_from_parents along a chain."""
    from mav_active_3d_planning_b200 import synth
    from oracle import oracle_py as O
    n_seg, pts = 4, 5
    pos, quat = synth.candidate_views(room02, n_seg * pts, seed=29)
    # segments revisit nearby places so that parents actually clear voxels
    for s in range(1, n_seg):
        pos[s * pts:(s + 1) * pts] = pos[:pts] + 0.05 * s
        quat[s * pts:(s + 1) * pts] = quat[:pts]
    dt = 250_000_000   # 0.25 s between trajectory points → sampling_time 0.5 s picks points 2 and 4
    mp, sp = write_inputs(str(tmp_path), room02, pos, quat, pts, dt)
    out = os.path.join(str(tmp_path), "out.txt")
    subprocess.run([EXE, "gains", os.path.join(GOLD, "planner_voxeltype_simple_parents.yaml"), mp, sp, out],
                   check=True, timeout=300)
    r = parse_out(out)
    om = O.OracleMap.from_synth(room02)
    cam = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0,
               downsampling_factor=2.0, mounting_translation=(0.1, 0.0, -0.05),
               mounting_rotation=(0.99, 0.0, 0.1, 0.0))
    oc = om.caster("SimpleRayCaster", **cam)
    ep = O.eval_params("VoxelTypeEvaluator", clear_from_parents=True, gain_unknown=1.0, gain_free=0.5,
                       gain_occupied=0.25)
    times = np.arange(pts, dtype=np.int64) * dt
    sel = O.sample_viewpoints(0.5, times)
    assert list(sel) == [2, 4]
    vp = np.concatenate([pos[s * pts + sel] for s in range(n_seg)])
    vq = np.concatenate([quat[s * pts + sel] for s in range(n_seg)])
    seg = np.repeat(np.arange(n_seg, dtype=np.int32), len(sel))
    parent = np.arange(-1, n_seg - 1, dtype=np.int32)     # a chain, like the driver builds
    want = oc.compute_gains(ep, vp, vq, seg, n_seg, parent=parent)
    for s in range(n_seg):
        ok, gain, n_list, dig = r["single"][s]
        assert ok == 1 and n_list == want["n_visible"][s] and dig == want["digest"][s]
        assert np.isclose(gain, want["gain"][s], rtol=1e-5, atol=0)
        assert r["refold"][s] == 1
    assert want["n_visible"][1] < want["n_visible"][0]    # parents did clear something
    # the batch call subtracts the ancestors' lists on the device (a3d_compute_gains_tree)
    for s in range(n_seg):
        ok, gain = r["batch"][s]
        assert ok == 1 and np.isclose(gain, want["gain"][s], rtol=1e-5, atol=0)


def _yaw_of(q):
    w, x, y, z = q
    return np.arctan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))


def _quat_of_yaw(yaw):
    return np.array([np.cos(yaw / 2.0), 0.0, 0.0, np.sin(yaw / 2.0)])


def _angle_scaled(a):
    a = np.fmod(a, 2.0 * np.pi)
    return a + 2.0 * np.pi * (a < 0)


def parse_yaw(path):
    res = {"yaw": {}, "yupdate": {}, "ybatch": {}, "ok": [], "device_list": None}
    for line in open(path):
        t = line.split()
        if t[0] in ("yaw", "yupdate", "ybatch"):
            v = [float(x) for x in t[2:]]
            res[t[0]][int(t[1])] = dict(gain=v[0], cost=v[1], value=v[2], active=int(v[3]), yaw=v[4],
                                        front_yaw=v[5], gains=v[6:])
        elif t[0] == "ok":
            res["ok"].append(int(t[1]))
        elif t[0] == "device_list":
            res["device_list"] = int(t[1])
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,ekind,n_dir,n_fov,by_value", [
    ("planner_yaw_simple.yaml", "NaiveEvaluator", 12, 1, False),
    ("planner_yaw_simple_value.yaml", "VoxelTypeEvaluator", 5, 1, True),
    ("planner_yaw_continuous.yaml", "NaiveEvaluator", 12, 3, False),
])
def test_yaw_planning_evaluators_match_oracle(built, room02, tmp_path, cfg, ekind, n_dir, n_fov, by_value):
    """Simple/ContinuousYawPlanningEvaluator (yaw_planning_evaluator.cpp:48-110,
    continuous_yaw_planning_evaluator.cpp:44-92): per-orientation gains against the oracle, the
    reference's selection rule restated here, and the one-launch batch call against the
    per-segment call."""
    from mav_active_3d_planning_b200 import synth
    from oracle import oracle_py as O
    n_seg, pts = 5, 3
    pos, quat = synth.candidate_views(room02, n_seg * pts, seed=41)
    pos[pts * 2: pts * 3] = pos[pts * 2]          # a segment that does not move: "indifferent" needs atan2(0,0)
    mp, sp = write_inputs(str(tmp_path), room02, pos, quat, pts, 100_000_000)
    out = os.path.join(str(tmp_path), "yaw.txt")
    subprocess.run([EXE, "yaw", os.path.join(GOLD, cfg), mp, sp, out], check=True, timeout=300)
    r = parse_yaw(out)
    assert r["ok"] and all(r["ok"])
    om = O.OracleMap.from_synth(room02)
    cam = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0, downsampling_factor=2.0)
    oc = om.caster("IterativeRayCaster", **cam)
    bv = dict(x_min=2, x_max=18, y_min=2, y_max=18, z_min=0.5, z_max=7)
    ep = O.eval_params(ekind, bounding_volume=bv)
    for s in range(n_seg):
        last = s * pts + pts - 1
        original = _yaw_of(quat[last])
        yaws = [_angle_scaled(original + float(i) * 2.0 * np.pi / n_dir) for i in range(n_dir)]
        qs = np.array([_quat_of_yaw(y) for y in yaws])
        want = oc.compute_gains(ep, np.repeat(pos[last:last + 1], n_dir, axis=0), qs)["gain"]
        applied = [_yaw_of(q) for q in qs]
        if n_fov > 1:
            score = [sum(want[(i + j) % n_dir] for j in range(n_fov)) for i in range(n_dir)]
        elif by_value:
            score = [want[i] - 100.0 * (1.0 + applied[i]) for i in range(n_dir)]
        else:
            score = list(want)
        # the reference's scan (first strict maximum; "min" only updated in the else branch)
        if n_fov > 1:
            mx = mn = score[0]
            best = 0
            rng = range(1, n_dir)
        else:
            mx, mn, best, rng = -np.inf, np.inf, 0, range(n_dir)
            mx, mn = np.finfo(float).min, np.finfo(float).max
        for i in rng:
            if score[i] > mx:
                mx, best = score[i], i
            elif score[i] < mn:
                mn = score[i]
        if mx > mn:
            best_yaw = _angle_scaled(applied[best] + (float(n_fov - 1) / n_dir * np.pi if n_fov > 1 else 0.0))
        else:
            d = pos[last] - pos[s * pts]
            best_yaw = _angle_scaled(np.arctan2(d[1], d[0]))
        final = _yaw_of(_quat_of_yaw(best_yaw))
        gain = sum(want[(best + j) % n_dir] for j in range(n_fov))
        for tag in ("yaw", "ybatch", "yupdate"):
            got = r[tag][s]
            assert got["gains"] == list(want), (tag, s)
            assert got["active"] == best and got["gain"] == gain, (tag, s)
            assert got["yaw"] == final and got["front_yaw"] == final, (tag, s)
            if n_fov > 1:
                assert got["cost"] == 1.0 + final and got["value"] == gain - 100.0 * (1.0 + final)
            elif by_value:
                assert got["cost"] == 1.0 + applied[best] and got["value"] == score[best]
        assert max(want) > 0
    assert (r["device_list"] > 0) == (cfg == "planner_yaw_simple_value.yaml")


@pytest.mark.gpu
def test_voxblox_layer_adapter(built, room02, tmp_path):
    """blocks flagged Update::kMap in (mock) voxblox ESDF/TSDF layers reach the
    device mirror through map::syncUpdatedBlocks; gains afterwards equal the oracle's on the same map,
    the flag is cleared (other flags untouched), a second sync sends nothing, a changed block one."""
    from mav_active_3d_planning_b200 import synth
    from oracle import oracle_py as O
    pos, quat = synth.candidate_views(room02, 5, seed=11)
    mp, sp = write_inputs(str(tmp_path), room02, pos, quat, 1, 0)
    out = os.path.join(str(tmp_path), "adapter.txt")
    subprocess.run([EXE, "adapter", os.path.join(GOLD, "planner_naive_iterative.yaml"), mp, sp, out],
                   check=True, timeout=300)
    lines = [l.split() for l in open(out)]
    sent = next(l for l in lines if l[0] == "sent")
    assert int(sent[1]) == room02.n_blocks and int(sent[3]) == room02.n_blocks
    assert int(sent[5]) == 0 and int(sent[7]) == room02.n_blocks      # kMap cleared, kMesh untouched
    assert int(next(l for l in lines if l[0] == "again")[1]) == 0
    om = O.OracleMap.from_synth(room02)
    cam = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0)
    oc = om.caster("IterativeRayCaster", **cam)
    ep = O.eval_params("NaiveEvaluator", bounding_volume=dict(x_min=2, x_max=18, y_min=2, y_max=18, z_min=0.5, z_max=7))
    want = oc.compute_gains(ep, pos, quat)["gain"]
    got = {int(l[1]): float(l[3]) for l in lines if l[0] == "batch"}
    assert [got[s] for s in range(5)] == list(want)
    # TSDF weight at the first pose: the mock TSDF layer lacks every third block (weight 0 there)
    w = np.array(room02.weight, np.float32)
    w[0::3] = 0.0
    om.upload_blocks(room02.block_idx, room02.dist, room02.observed, w)
    assert float(next(l for l in lines if l[0] == "weight")[1]) == om.voxel_weight(pos[:1])[0]
    # the changed block
    d = next(l for l in lines if l[0] == "delta")
    assert int(d[1]) == 1
    bi = np.array([[int(d[2]), int(d[3]), int(d[4])]], np.int32)
    nv = room02.vps ** 3
    om.upload_blocks(bi, np.full((1, nv), 1.0, np.float32), np.ones((1, nv), np.uint8), np.zeros((1, nv), np.float32))
    want2 = oc.compute_gains(ep, pos, quat)["gain"]
    got2 = {int(l[1]): float(l[3]) for l in lines if l[0] == "batch2"}
    assert [got2[s] for s in range(5)] == list(want2)


def _changed_map(m, seed=5):
    """the same blocks with a third of them observed free space now (what an integrator update does)"""
    import dataclasses
    rng = np.random.default_rng(seed)
    sel = rng.choice(m.n_blocks, m.n_blocks // 3, replace=False)
    dist, obs = m.dist.copy(), m.observed.copy()
    dist[sel] = 1.0
    obs[sel] = 1
    return dataclasses.replace(m, dist=dist, observed=obs)


def parse_update(path):
    res = {"before": {}, "after": {}, "single": []}
    for line in open(path):
        t = line.split()
        if t[0] in ("before", "after"):
            res[t[0]][int(t[1])] = tuple(float(x) for x in t[2:5])
        elif t[0] == "removed":
            res["removed"], res["launches"], res["gain_ms"] = int(t[1]), int(t[3]), float(t[5])
        elif t[0] == "single":
            res["single"].append((int(t[1]), float(t[2])))
        elif t[0] == "casttime":
            res["casttime"] = (float(t[1]), float(t[2]), int(t[3]))
        elif t[0] == "ok":
            res["ok"] = int(t[1])
    return res


@pytest.mark.gpu
def test_simulated_sensor_updater_refolds_the_tree_in_one_launch(built, room02, tmp_path):
    """evaluator_updater: SimulatedSensorUpdater -> PruneDirect (simulated_sensor_updater.cpp:18-22,
    prune_direct.cpp:33-39) through TrajectoryEvaluator::updateTree: the gain of every stored list
    against the changed map equals the oracle's fold of the SAME lists, the chain is cut where the gain
    falls below the bound, and the whole tree costs one re-fold launch."""
    from mav_active_3d_planning_b200 import synth
    from oracle import oracle_py as O
    n = 7
    pos, quat = synth.candidate_views(room02, n, seed=61)
    m2 = _changed_map(room02)
    cam = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0, downsampling_factor=2.0)
    ep = O.eval_params("VoxelTypeEvaluator", gain_unknown=1.0, gain_occupied=0.5, gain_free=0.25)
    om1, om2 = O.OracleMap.from_synth(room02), O.OracleMap.from_synth(m2)
    oc = om1.caster("IterativeRayCaster", **cam)

    def refolded(i):
        lst, _ = oc.visible_voxels(pos[i:i + 1], quat[i:i + 1])
        return om2.gain_from_voxels(ep, lst)[0]
    # order the chain below the root by falling gain-after-the-change, so that a bound cuts it in the middle
    order = [0] + sorted(range(1, n), key=refolded, reverse=True)
    pos, quat = pos[order], quat[order]
    after = np.array([refolded(i) for i in range(1, n)])
    before = oc.compute_gains(ep, pos[1:], quat[1:])["gain"]
    assert np.any(after != before)
    cut = (n - 1) // 2
    assert after[cut - 1] > after[cut]
    bound = 0.5 * float(after[cut - 1] + after[cut])
    mp, sp = write_inputs(str(tmp_path), room02, pos, quat, 1, 0)
    d2 = os.path.join(str(tmp_path), "m2")
    os.makedirs(d2)
    mp2, _ = write_inputs(d2, m2, pos, quat, 1, 0)
    yml = os.path.join(str(tmp_path), "planner.yaml")
    open(yml, "w").write(open(os.path.join(GOLD, "planner_update_sensor.yaml")).read()
                         .replace("minimum_gain: 1200.0", f"minimum_gain: {bound!r}"))
    out = os.path.join(str(tmp_path), "update.txt")
    subprocess.run([EXE, "update", yml, mp, sp, out, mp2, "stored"], check=True, timeout=300)
    r = parse_update(out)
    assert r["ok"] == 1
    for s in range(n - 1):
        assert np.isclose(r["before"][s][0], before[s], rtol=1e-5, atol=0)
    assert r["removed"] == (n - 1) - cut and sorted(r["after"]) == list(range(cut))
    for s in range(cut):
        assert np.isclose(r["after"][s][0], after[s], rtol=1e-5, atol=0)
        assert r["after"][s][1:] == r["before"][s][1:]          # this chain does not touch cost / value
    assert r["launches"] <= 2                                   # the tree is one re-fold, not n
    assert r["gain_ms"] >= 0.0
    assert [x[0] for x in r["single"]] == [1] * cut
    assert np.allclose([x[1] for x in r["single"]], after[:cut], rtol=1e-5, atol=0)
    # `test: true` of the sensor model: one timing entry per view cast (camera_model.cpp:31-54)
    mean, std, count = r["casttime"]
    assert count == n - 1 and mean > 0.0


@pytest.mark.gpu
def test_update_all_recasts_the_tree_as_one_batch(built, room02, tmp_path):
    """evaluator_updater: ConstrainedUpdater -> UpdateAll (constrained_updater.cpp:17-30,
    update_all.cpp:16-27): segments within update_range of the robot get gain, cost and value again —
    one cast for all of them —, the others keep theirs; nothing is dropped."""
    from mav_active_3d_planning_b200 import synth
    from oracle import oracle_py as O
    n = 9
    pos, quat = synth.candidate_views(room02, n, seed=67)
    m2 = _changed_map(room02, seed=9)
    mp, sp = write_inputs(str(tmp_path), room02, pos, quat, 1, 0)
    d2 = os.path.join(str(tmp_path), "m2")
    os.makedirs(d2)
    mp2, _ = write_inputs(d2, m2, pos, quat, 1, 0)
    cam = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0, downsampling_factor=2.0)
    ep = O.eval_params("NaiveEvaluator")
    om1, om2 = O.OracleMap.from_synth(room02), O.OracleMap.from_synth(m2)
    before = om1.caster("SimpleRayCaster", **cam).compute_gains(ep, pos[1:], quat[1:])["gain"]
    recast = om2.caster("SimpleRayCaster", **cam).compute_gains(ep, pos[1:], quat[1:])["gain"]
    out = os.path.join(str(tmp_path), "update.txt")
    subprocess.run([EXE, "update", os.path.join(GOLD, "planner_update_all.yaml"), mp, sp, out, mp2, "host"],
                   check=True, timeout=300)
    r = parse_update(out)
    near = (np.linalg.norm(pos[1:] - pos[0], axis=1) <= 6.0) & (before > 0.0)
    assert near.any() and (~near).any() and np.any(recast[near] != before[near])
    assert r["removed"] == 0 and sorted(r["after"]) == list(range(n - 1))
    for s in range(n - 1):
        want = recast[s] if near[s] else before[s]
        gain, cost, value = r["after"][s]
        assert gain == want and cost == r["before"][s][1] and value == want - 100.0 * cost
    assert r["launches"] <= 8      # one batch cast (kernel chain + flagged-segment pass), not one per segment


@pytest.mark.gpu
def test_host_batch_gains_sharded_over_two_gpus(built, room02, tmp_path):
    """map: devices: "0,1" — VoxbloxMap holds an a3d_group (replicated mirror, NCCL broadcast of the
    upload), SimulatedSensorEvaluator::computeGains shards the segments over both GPUs and gathers the
    gains: same numbers as one GPU, as the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from mav_active_3d_planning_b200 import synth
    from oracle import oracle_py as O
    pos, quat = synth.candidate_views(room02, 6, seed=3)
    mp, sp = write_inputs(str(tmp_path), room02, pos, quat, 1, 0)
    yml = os.path.join(str(tmp_path), "planner.yaml")
    open(yml, "w").write(open(os.path.join(GOLD, "planner_naive_iterative.yaml")).read() +
                         '\nmap:\n  devices: "0,1"\n')
    out = os.path.join(str(tmp_path), "out.txt")
    subprocess.run([EXE, "gains", yml, mp, sp, out], check=True, timeout=300)
    r = parse_out(out)
    om = O.OracleMap.from_synth(room02)
    oc = om.caster("IterativeRayCaster", resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0)
    ep = O.eval_params("NaiveEvaluator", bounding_volume=dict(x_min=2, x_max=18, y_min=2, y_max=18, z_min=0.5, z_max=7))
    want = oc.compute_gains(ep, pos, quat)
    for s in range(6):
        assert r["batch"][s] == (1, want["gain"][s])          # sharded batch call
        assert r["single"][s][0] == 1 and r["single"][s][1] == want["gain"][s]   # per-segment call, GPU 0


@pytest.mark.gpu
def test_layer_messages(built, room02, tmp_path):
    """SURVEY.md §8(f)-3, the wire half: voxblox_msgs/Layer messages (ESDF: distance + observed, TSDF:
    weight) decoded by map::LayerReceiver into the device mirror, in an order where TSDF weights arrive
    before their ESDF blocks; VoxelWeightEvaluator gains and a TSDF weight equal the oracle's on the same
    map; mismatching messages are refused; ACTION_RESET replaces the mirror."""
    from mav_active_3d_planning_b200 import synth
    from oracle import oracle_py as O
    pos, quat = synth.candidate_views(room02, 5, seed=13)
    mp, sp = write_inputs(str(tmp_path), room02, pos, quat, 1, 0)
    out = os.path.join(str(tmp_path), "layer.txt")
    subprocess.run([EXE, "layer", os.path.join(GOLD, "planner_voxelweight.yaml"), mp, sp, out], check=True, timeout=300)
    lines = {l.split()[0]: l.split() for l in open(out) if not l.startswith("batch")}
    nb = room02.n_blocks
    assert lines["tsdf_first"][1] == "1" and int(lines["tsdf_first"][3]) == nb // 2     # all kept for later
    assert lines["esdf_a"][1] == "1" and lines["esdf_b"][1] == "1" and int(lines["esdf_b"][3]) == 0
    assert lines["tsdf_rest"][1] == "1" and int(lines["tsdf_rest"][3]) == 0
    assert int(lines["blocks"][1]) == nb and int(lines["blocks"][3]) == nb
    assert lines["refused"][1:] == ["1", "1"]
    om = O.OracleMap.from_synth(room02)
    oc = om.caster("IterativeRayCaster", resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0,
                   downsampling_factor=2.0)
    ep = O.eval_params("VoxelWeightEvaluator", accurate_frontiers=True, checking_distance=1.0,
                       frontier_voxel_weight=1.0, new_voxel_weight=0.0, min_impact_factor=0.01,
                       ray_angle_x=0.002454, ray_angle_y=0.002681)
    want = oc.compute_gains(ep, pos, quat)["gain"]
    got = {int(l.split()[1]): float(l.split()[3]) for l in open(out) if l.startswith("batch")}
    assert np.allclose([got[s] for s in range(5)], want, rtol=1e-5, atol=1e-12) and want.max() > 0
    assert float(lines["weight"][1]) == om.voxel_weight(pos[:1])[0]
    # getVoxelDistance (voxblox.cpp:84-98) reads the other member of the same TSDF voxel; the test's TSDF
    # messages carry the ESDF distance array there: the oracle's weight lookup on a map whose weight member
    # holds that array is the same arithmetic.  No block: 0.
    od = O.OracleMap(room02.voxel_size, room02.vps)
    od.upload_blocks(room02.block_idx, room02.dist, room02.observed, room02.dist)
    assert float(lines["tsdf_distance"][1]) == od.voxel_weight(pos[:1])[0]
    assert float(lines["tsdf_distance"][2]) == 0.0
    assert lines["reset"][1] == "1" and int(lines["reset"][3]) == nb // 3
