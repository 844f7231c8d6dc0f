"""Pins the CPU oracle (oracle/a3d_oracle.cpp, a restatement) to oracle/_ref — the reference's own
UNMODIFIED sources (camera_model.cpp, simple_ray_caster.cpp, iterative_ray_caster.cpp,
iterative_ray_caster_lidar.cpp, simulated_sensor_evaluator.cpp, voxel_type/naive/frontier/
voxel_weight_evaluator.cpp, simulated_sensor_updater.cpp, bounding_volume.cpp, sensor_model.cpp,
vbx/src/map/voxblox.cpp, module*.cpp) compiled against shim Eigen/glog/voxblox headers
(oracle/Makefile target `_ref`, oracle/ref_driver.cpp).  Every comparison is bit-for-bit: ordered
visible-voxel lists, counts, digests, gains, fold lookups.

`_ref` is built here from /root/reference; on the GPU box (no /root/reference) the prebuilt
oracle/_ref/liba3d_ref.so that travelled with the snapshot is used.  Neither → skip.
"""
import itertools

import numpy as np
import pytest

from conftest import CAM_BASE, CAM_CFG4
from mav_active_3d_planning_b200 import synth
from oracle import oracle_py as O


@pytest.fixture(scope="module")
def ref(built):
    if O.build_ref() is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    assert O.lib("ref").orc_impl() == b"reference" and O.lib().orc_impl() == b"port"
    return True


def both(m):
    return O.OracleMap.from_synth(m), O.OracleMap.from_synth(m, impl="ref")


def general_quats(n, seed):
    """Unit quaternions with roll and pitch, not only yaw (exercises Quaterniond*Quaterniond)."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


MOUNT = dict(mounting_translation=(0.1, 0.02, -0.05), mounting_rotation=(0.9, 0.1, -0.2, 0.3))
BV = dict(x_min=2.0, x_max=15.0, y_min=1.0, y_max=12.0, z_min=0.5, z_max=6.0, rotation=10.0)
EVALS = [
    ("VoxelTypeEvaluator", {}),
    ("VoxelTypeEvaluator", dict(gain_unknown=1.0, gain_occupied=0.5, gain_free=0.25, gain_unknown_outer=0.1,
                                gain_occupied_outer=0.05, gain_free_outer=0.025, bounding_volume=BV)),
    ("NaiveEvaluator", {}),
    ("NaiveEvaluator", dict(bounding_volume=BV)),
    ("FrontierEvaluator", {}),
    ("FrontierEvaluator", dict(accurate_frontiers=True)),
    ("FrontierEvaluator", dict(surface_frontiers=False, checking_distance=2.0)),
    ("VoxelWeightEvaluator", {}),
    ("VoxelWeightEvaluator", dict(frontier_voxel_weight=0.0, min_impact_factor=0.2, new_voxel_weight=0.5)),
]


def assert_same(ra, rb, what):
    for k in ra:
        assert np.array_equal(ra[k], rb[k]), (what, k, ra[k][:4], rb[k][:4])


def test_map_functions(ref, room02):
    """vbx/src/map/voxblox.cpp:32-115 through the reference's VoxbloxMap vs the restatement."""
    A, B = both(room02)
    rng = np.random.default_rng(0)
    pts = rng.uniform([-1, -1, -1], [21, 21, 11], (300000, 3))
    # points within a few ulp of voxel faces and block faces
    g = np.round(rng.uniform(0, 100, (20000, 3))) * float(np.float32(0.2))
    faces = np.concatenate([np.nextafter(g, -np.inf), g, np.nextafter(g, np.inf),
                            g.astype(np.float32).astype(np.float64)])
    pts = np.concatenate([pts, faces])
    for f in ("voxel_state", "voxel_center", "is_observed", "voxel_weight"):
        assert np.array_equal(getattr(A, f)(pts), getattr(B, f)(pts)), f
    da, oa = A.distance(pts)
    db, ob = B.distance(pts)
    assert np.array_equal(oa, ob) and np.array_equal(da[oa > 0], db[ob > 0])
    for r in (0.3, 1.0):
        assert np.array_equal(A.traversable(pts, r), B.traversable(pts, r))


@pytest.mark.parametrize("kind", ["SimpleRayCaster", "IterativeRayCaster", "IterativeRayCasterLidar"])
def test_cfg1_lists_and_gains(ref, room02, kind):
    """cfg1 map, 8 poses: derived constants, directions, ordered lists, all evaluators."""
    A, B = both(room02)
    pos, quat = synth.candidate_views(room02, 8)
    kw = dict(CAM_BASE)
    if "Lidar" in kind:
        kw.update(resolution_x=64, resolution_y=32)
    for mount in ({}, MOUNT):
        ca, cb = A.caster(kind, **kw, **mount), B.caster(kind, **kw, **mount)
        assert (ca.c_res_x, ca.c_res_y, ca.c_n_sections) == (cb.c_res_x, cb.c_res_y, cb.c_n_sections)
        assert np.array_equal(ca.split_distances, cb.split_distances)
        assert np.array_equal(ca.split_widths, cb.split_widths)
        assert (ca.fov_x, ca.fov_y, ca.ray_step) == (cb.fov_x, cb.fov_y, cb.ray_step)
        for i, j in itertools.product(range(0, ca.c_res_x, 3), range(0, ca.c_res_y, 3)):
            assert np.array_equal(ca.direction(i, j), cb.direction(i, j))
        qs = quat if not mount else general_quats(8, 5)
        for v in range(8):
            la, na = ca.visible_voxels(pos[v:v + 1], qs[v:v + 1])
            lb, nb = cb.visible_voxels(pos[v:v + 1], qs[v:v + 1])
            assert na == nb and np.array_equal(la, lb), (kind, v)
        # multi-view segment: std::unique runs over the concatenation (camera_model.cpp:58)
        la, _ = ca.visible_voxels(pos[:5], qs[:5])
        lb, _ = cb.visible_voxels(pos[:5], qs[:5])
        assert np.array_equal(la, lb)
        vseg = np.array([0, 0, 0, 1, 2, 2, 3, 3], np.int32)
        for ek, ekw in EVALS:
            e = O.eval_params(ek, **ekw)
            assert_same(ca.compute_gains(e, pos, qs), cb.compute_gains(e, pos, qs), (kind, ek, ekw))
            assert_same(ca.compute_gains(e, pos, qs, view_segment=vseg, n_segments=5),
                        cb.compute_gains(e, pos, qs, view_segment=vseg, n_segments=5), (kind, ek, "multi"))


@pytest.mark.parametrize("kind", ["SimpleRayCaster", "IterativeRayCaster"])
def test_clear_from_parents_chains(ref, room02, kind):
    """simulated_sensor_evaluator.cpp:60-76 with 3-deep parent chains, every evaluator."""
    A, B = both(room02)
    pos, quat = synth.candidate_views(room02, 8)
    pos[1] = pos[0] + [0.2, 0.0, 0.0]      # children that really overlap their parents
    pos[2] = pos[0] + [0.2, 0.2, 0.0]
    pos[3] = pos[0] + [0.4, 0.2, 0.1]
    quat[1:4] = quat[0]
    parent = np.array([-1, 0, 1, 2, -1, 4, 4, 6], np.int32)
    ca, cb = A.caster(kind, **CAM_BASE), B.caster(kind, **CAM_BASE)
    for ek, ekw in EVALS:
        e = O.eval_params(ek, clear_from_parents=True, **ekw)
        ra, rb = ca.compute_gains(e, pos, quat, parent=parent), cb.compute_gains(e, pos, quat, parent=parent)
        assert_same(ra, rb, (kind, ek, ekw))
        if "bounding_volume" not in ekw:  # the chain really removes entries
            lone = ca.compute_gains(O.eval_params(ek, **ekw), pos[3:4], quat[3:4])
            assert 0 < ra["n_visible"][3] < lone["n_visible"][0]


def test_refold_path(ref, room02):
    """SimulatedSensorUpdater::updateSegment (simulated_sensor_updater.cpp:18-22) →
    computeGainFromVisibleVoxels on a stored list, incl. the Naive/Frontier erasure."""
    A, B = both(room02)
    pos, quat = synth.candidate_views(room02, 3)
    lst, _ = A.caster("IterativeRayCaster", **CAM_BASE).visible_voxels(pos[:1], quat[:1])
    for ek, ekw in EVALS:
        e = O.eval_params(ek, **ekw)
        assert A.gain_from_voxels(e, lst, origin=pos[0]) == B.gain_from_voxels(e, lst, origin=pos[0]), (ek, ekw)


def test_bounding_volume_and_sampling(ref):
    rng = np.random.default_rng(3)
    pts = rng.uniform(-2, 18, (50000, 3))
    for bv in (BV, dict(BV, rotation=-37.5), dict(BV, rotation=0.0), dict(x_min=1, x_max=1), {}):
        e = O.eval_params("VoxelTypeEvaluator", bounding_volume=bv)
        assert np.array_equal(O.bv_contains(e, pts), O.bv_contains(e, pts, impl="ref")), bv
    t = np.cumsum(rng.integers(1, 400_000_000, 40)).astype(np.int64)
    for st in (0.0, -1.0, 0.05, 0.31, 1.0, 2.5, 1e3):
        for n in (1, 2, 7, 40):
            assert np.array_equal(O.sample_viewpoints(st, t[:n]), O.sample_viewpoints(st, t[:n], impl="ref")), (st, n)


def test_cfg2_and_cfg4_poses(ref, room01):
    """0.1 m room: cfg2 camera (Iterative) and cfg4 camera (Simple + Iterative, 8 m range)."""
    A, B = both(room01)
    pos, quat = synth.candidate_views(room01, 6)
    # (the reference's Naive/Frontier folds erase one entry at a time — O(V^2) on these 10^5-entry
    # lists — so the big lists are folded with VoxelType and VoxelWeight only; cfg1 covers the rest)
    evs = (O.eval_params("VoxelTypeEvaluator"), O.eval_params("VoxelWeightEvaluator"))
    for kind, cam, n in (("IterativeRayCaster", CAM_BASE, 4), ("SimpleRayCaster", CAM_BASE, 1),
                         ("IterativeRayCaster", CAM_CFG4, 1), ("SimpleRayCaster", CAM_CFG4, 1)):
        ca, cb = A.caster(kind, **cam), B.caster(kind, **cam)
        la, na = ca.visible_voxels(pos[:1], quat[:1])
        lb, nb = cb.visible_voxels(pos[:1], quat[:1])
        assert na == nb and np.array_equal(la, lb), (kind, cam)
        for e in evs:
            assert_same(ca.compute_gains(e, pos[:n], quat[:n]), cb.compute_gains(e, pos[:n], quat[:n]), (kind, cam))


def test_map_edits(ref, room02):
    """Block overwrite / removal between casts (the incremental-delta path of cfg5)."""
    A, B = both(room02)
    pos, quat = synth.candidate_views(room02, 4)
    rng = np.random.default_rng(11)
    ca, cb = A.caster("IterativeRayCaster", **CAM_BASE), B.caster("IterativeRayCaster", **CAM_BASE)
    e = O.eval_params("FrontierEvaluator")
    for step in range(3):
        pick = rng.choice(room02.n_blocks, 12, replace=False)
        d = room02.dist[pick] + rng.normal(0, 0.05, room02.dist[pick].shape).astype(np.float32)
        o = (rng.random(room02.observed[pick].shape) < 0.9).astype(np.uint8)
        gone = room02.block_idx[rng.choice(room02.n_blocks, 3, replace=False)][:0 if step == 0 else 3]
        for M in (A, B):
            M.upload_blocks(room02.block_idx[pick], d, o, room02.weight[pick])
            M.remove_blocks(gone)
        assert A.num_blocks == B.num_blocks
        assert_same(ca.compute_gains(e, pos, quat), cb.compute_gains(e, pos, quat), step)


def test_threads_match_sequential(ref, room02):
    """The multi-threaded reference arm (one planner + map copy per thread) returns what the
    single-threaded reference returns."""
    _, B = both(room02)
    pos, quat = synth.candidate_views(room02, 12)
    cb = B.caster("IterativeRayCaster", **CAM_BASE)
    e = O.eval_params("VoxelTypeEvaluator")
    assert_same(cb.compute_gains(e, pos, quat), cb.compute_gains(e, pos, quat, n_threads=4), "threads")


def _fuzz_case(rng):
    vs = float(rng.choice([0.05, 0.1, 0.2, 0.25]))
    vps = int(rng.choice([8, 16]))
    nb = int(rng.integers(2, 5))
    idx = np.array(list(itertools.product(range(-1, nb - 1), repeat=3)), np.int32)
    keep = rng.random(len(idx)) < 0.85
    idx = idx[keep]
    nv = vps ** 3
    dist = rng.uniform(-0.5 * vs, 6 * vs, (len(idx), nv)).astype(np.float32)
    obs = (rng.random((len(idx), nv)) < rng.uniform(0.6, 1.0)).astype(np.uint8)
    w = (rng.uniform(0, 50, (len(idx), nv)) * obs).astype(np.float32)
    ext = vs * vps * (nb - 1)
    n = int(rng.integers(1, 4))
    pos = rng.uniform(-0.5 * vs * vps, ext, (n, 3))
    quat = general_quats(n, int(rng.integers(1 << 30)))
    kind = str(rng.choice(["SimpleRayCaster", "IterativeRayCaster", "IterativeRayCasterLidar"]))
    cam = dict(resolution_x=int(rng.integers(8, 90)), resolution_y=int(rng.integers(8, 70)),
               focal_length=float(rng.uniform(20, 120)), ray_length=float(rng.uniform(4 * vs, 3.0)),
               downsampling_factor=float(rng.uniform(0.7, 3.0)),
               ray_step=float(rng.choice([0.0, 0.0, vs * 0.7, vs * 1.3])))
    if rng.random() < 0.5:
        cam.update(mounting_translation=tuple(rng.uniform(-0.2, 0.2, 3)),
                   mounting_rotation=tuple(rng.normal(size=4)))
    if "Lidar" in kind:
        cam.update(field_of_view_x=float(rng.uniform(60, 360)), field_of_view_y=float(rng.uniform(10, 60)))
    ek, ekw = EVALS[int(rng.integers(len(EVALS)))]
    ekw = dict(ekw)
    if "bounding_volume" in ekw:
        ekw["bounding_volume"] = dict(x_min=-ext, x_max=0.6 * ext, y_min=-ext, y_max=0.7 * ext, z_min=-ext,
                                      z_max=0.5 * ext, rotation=float(rng.uniform(-90, 90)))
    return vs, vps, idx, dist, obs, w, pos, quat, kind, cam, ek, ekw


def test_fuzz_200(ref):
    """200 random (map, camera, mount, pose, evaluator) cases, general orientations."""
    rng = np.random.default_rng(20261020)
    n_cases = 0
    for case in range(240):
        vs, vps, idx, dist, obs, w, pos, quat, kind, cam, ek, ekw = _fuzz_case(rng)
        A, B = O.OracleMap(vs, vps), O.OracleMap(vs, vps, impl="ref")
        for M in (A, B):
            M.upload_blocks(idx, dist, obs, w)
        ca, cb = A.caster(kind, **cam), B.caster(kind, **cam)
        assert (ca.c_res_x, ca.c_res_y, ca.c_n_sections) == (cb.c_res_x, cb.c_res_y, cb.c_n_sections), case
        if min(ca.c_res_x, ca.c_res_y) < 2:
            # undefined in the reference: relative_y = 0/0 (simple_ray_caster.cpp:44-45) and, for the
            # iterative casters, c_split_distances_[1] is read past the end (iterative_ray_caster.cpp:72)
            continue
        n_cases += 1
        la, na = ca.visible_voxels(pos, quat)
        lb, nb = cb.visible_voxels(pos, quat)
        assert na == nb and np.array_equal(la, lb), (case, kind, cam)
        e = O.eval_params(ek, clear_from_parents=bool(case % 3 == 0), **ekw)
        parent = np.arange(-1, len(pos) - 1, dtype=np.int32)
        assert_same(ca.compute_gains(e, pos, quat, parent=parent), cb.compute_gains(e, pos, quat, parent=parent),
                    (case, kind, ek))
    assert n_cases >= 200


def test_esdf_only_update_leaves_the_tsdf_layer_alone(ref, room02):
    """The map stand-ins of both checkers model voxblox's two layers: a block update without weights (an
    ESDF layer message) keeps the block's TSDF weights, one with weights replaces them, a removal drops
    both; VoxelWeightEvaluator gains agree bit for bit afterwards."""
    m = room02
    rng = np.random.default_rng(8)
    sel = rng.choice(m.n_blocks, m.n_blocks // 4, replace=False)
    nd = (m.dist[sel] + rng.normal(0, 0.03, m.dist[sel].shape)).astype(np.float32)
    pts = (np.asarray(m.block_idx[sel][:, None, :], np.float64) + rng.random((len(sel), 40, 3))).reshape(-1, 3) \
        * (m.voxel_size * m.vps)
    pos, quat = synth.candidate_views(m, 3, seed=12)
    res = []
    for om in both(m):
        w0 = om.voxel_weight(pts)
        assert np.count_nonzero(w0) > 100
        om.upload_blocks(m.block_idx[sel], nd, m.observed[sel])                 # ESDF only
        assert np.array_equal(om.voxel_weight(pts), w0)
        om.upload_blocks(m.block_idx[sel[:5]], nd[:5], m.observed[sel[:5]], 2.0 * m.weight[sel[:5]])
        om.remove_blocks(m.block_idx[sel[5:8]])
        w1 = om.voxel_weight(pts)
        per = 40
        assert np.array_equal(w1[:5 * per], 2.0 * w0[:5 * per]) and not w1[5 * per:8 * per].any()
        assert np.array_equal(w1[8 * per:], w0[8 * per:])
        res.append(om.caster("IterativeRayCaster", **CAM_BASE).compute_gains(
            O.eval_params("VoxelWeightEvaluator"), pos, quat))
    assert_same(res[0], res[1], "VoxelWeight after an ESDF-only update")
