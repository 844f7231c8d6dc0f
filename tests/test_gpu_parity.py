"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for voxel states, centres, ordered visible-voxel lists, counts and digests; gains within
1e-5 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import CAM_BASE, CAM_CFG4

pytestmark = pytest.mark.gpu

GAIN_RTOL = 1e-5


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle_py
    return oracle_py


@pytest.fixture(scope="module")
def capi(built):
    from mav_active_3d_planning_b200 import capi
    return capi


def make_pair(capi, O, m):
    ctx = capi.Context(0)
    ctx.upload_map(m)
    return ctx, O.OracleMap.from_synth(m)


def probe_points(m, n, seed):
    """Random points + points hugging voxel/block faces and voxel centres (fp edge cases)."""
    rng = np.random.default_rng(seed)
    ext = np.array(m.extent)
    pts = [rng.random((n, 3)) * (ext + 2.0) - 1.0]
    vs = float(np.float32(m.voxel_size))
    g = rng.integers(0, (ext / vs).astype(int), (n, 3))
    pts.append(g * vs)                                    # on voxel faces
    pts.append(g * vs + rng.normal(0, 2e-7, (n, 3)))      # ulps around faces
    pts.append((g + 0.5) * vs)                            # voxel centres (double)
    pts.append((g + 0.5) * vs + rng.normal(0, 2e-7, (n, 3)))
    bs = vs * m.vps
    gb = rng.integers(0, (ext / bs).astype(int) + 1, (n, 3))
    pts.append(gb * bs + rng.normal(0, 1e-6, (n, 3)))     # around block faces
    pts.append(np.nextafter(np.float32(gb * bs), np.float32(-1e9)).astype(np.float64))
    return np.concatenate(pts)


@pytest.mark.parametrize("vs", [0.2, 0.1])
def test_map_queries_bit_exact(capi, O, vs):
    from mav_active_3d_planning_b200 import synth
    m = synth.room_map(vs)
    ctx, om = make_pair(capi, O, m)
    pts = probe_points(m, 20000, 11)
    assert np.array_equal(ctx.voxel_state(pts), om.voxel_state(pts))
    assert np.array_equal(ctx.voxel_center(pts), om.voxel_center(pts))
    assert np.array_equal(ctx.is_observed(pts), om.is_observed(pts))
    assert np.array_equal(ctx.voxel_weight(pts), om.voxel_weight(pts))
    d, ok = ctx.distance(pts)
    od, ook = om.distance(pts)
    assert np.array_equal(ok, ook) and np.array_equal(d[ok == 1], od[ook == 1])
    for radius in (0.3, 1.0):   # VoxbloxMap::isTraversable (voxblox.cpp:32-41)
        assert np.array_equal(ctx.traversable(pts, radius), ((ook == 1) & (od > radius)).astype(np.uint8))
    # states at emitted voxel centres (the fold's lookups)
    c = om.voxel_center(pts)
    assert np.array_equal(ctx.voxel_state(c), om.voxel_state(c))
    ctx.close()


def test_fast_path_equals_literal_path_at_scale(capi, room01):
    """Apron + meta-byte fast path vs the literal 8-lookup path on 4M points (incl. fp edge cases)."""
    ctx = capi.Context(0)
    ctx.upload_map(room01)
    pts = probe_points(room01, 600000, 23)
    assert np.array_equal(ctx.voxel_state(pts), ctx.voxel_state_literal(pts))
    c = ctx.voxel_center(pts)
    assert np.array_equal(ctx.voxel_state(c), ctx.voxel_state_literal(c))
    ctx.close()


def test_map_vps8_and_delta_updates(capi, O):
    from mav_active_3d_planning_b200 import synth
    m = synth.room_map(0.2, vps=8)
    ctx, om = make_pair(capi, O, m)
    assert ctx.num_blocks == om.num_blocks == m.n_blocks
    pts = probe_points(m, 5000, 3)
    assert np.array_equal(ctx.voxel_state(pts), om.voxel_state(pts))
    # overwrite 10 % of blocks, remove 5 %, add new ones
    rng = np.random.default_rng(5)
    sel = rng.choice(m.n_blocks, m.n_blocks // 10, replace=False)
    nd = m.dist[sel] + rng.normal(0, 0.05, m.dist[sel].shape).astype(np.float32)
    no = np.ones_like(m.observed[sel])
    ctx.upload_blocks(m.block_idx[sel], nd, no)
    om.upload_blocks(m.block_idx[sel], nd, no)
    rem = rng.choice(m.n_blocks, m.n_blocks // 20, replace=False)
    ctx.remove_blocks(m.block_idx[rem])
    om.remove_blocks(m.block_idx[rem])
    new_idx = np.array([[-1, 0, 0], [-1, 1, 0], [40, 40, 40]], np.int32)
    nd = np.full((3, 512), 0.7, np.float32)
    ctx.upload_blocks(new_idx, nd, np.ones((3, 512), np.uint8))
    om.upload_blocks(new_idx, nd, np.ones((3, 512), np.uint8))
    assert ctx.num_blocks == om.num_blocks
    pts = np.concatenate([pts, probe_points(m, 2000, 4) - [1.6, 0, 0]])
    assert np.array_equal(ctx.voxel_state(pts), om.voxel_state(pts))
    assert np.array_equal(ctx.is_observed(pts), om.is_observed(pts))
    ctx.clear()
    assert ctx.num_blocks == 0
    assert np.all(ctx.voxel_state(pts[:100]) == 2)
    ctx.close()


def test_empty_map_and_empty_batches(capi, O):
    ctx = capi.Context(0)
    ctx.map_config(0.2, 16)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    ev = ctx.evaluator("VoxelTypeEvaluator")
    pos = np.array([[1.0, 2.0, 3.0]])
    quat = np.array([[1.0, 0, 0, 0]])
    r = ctx.compute_gains(caster, ev, pos, quat)
    # nothing allocated: every sample UNKNOWN, nothing occludes → free-space count, gain = count
    assert r["n_samples"][0] == 19529 and r["gain"][0] == r["n_visible"][0] > 0
    # a segment without views and a zero-length batch
    r = ctx.compute_gains(caster, ev, pos, quat, view_segment=np.array([1], np.int32), n_segments=3)
    assert list(r["n_visible"]) == [0, r["n_visible"][1], 0] and r["gain"][0] == 0
    r = ctx.compute_gains(caster, ev, np.zeros((0, 3)), np.zeros((0, 4)), n_segments=0)
    assert len(r["gain"]) == 0
    ctx.close()


@pytest.mark.parametrize("kind", ["IterativeRayCaster", "SimpleRayCaster"])
def test_cfg1_visible_voxel_lists_bit_exact(capi, O, room02, kind):
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    caster = ctx.caster(kind, **CAM_BASE)
    oc = om.caster(kind, **CAM_BASE)
    info = caster.info
    assert (info.c_res_x, info.c_res_y, info.c_n_sections) == (oc.c_res_x, oc.c_res_y,
                                                                oc.c_n_sections if kind[0] == "I" else 0)
    for i, j in [(0, 0), (3, 7), (info.c_res_x - 1, info.c_res_y - 1)]:
        assert np.array_equal(caster.direction(i, j), oc.direction(i, j))
    pos, quat = synth.candidate_views(room02, 8, seed=21)
    for v in range(8):
        got, ns, _ = ctx.visible_voxels(caster, pos[v:v + 1], quat[v:v + 1])
        want, ons = oc.visible_voxels(pos[v:v + 1], quat[v:v + 1])
        assert ns == ons
        assert got.shape == want.shape and np.array_equal(got, want)
    ctx.close()


def test_cfg1_gains_64_views(capi, O, room02):
    """BASELINE config 1: 0.2 m room, IterativeRayCaster + VoxelTypeEvaluator, 64 views."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    oc = om.caster("IterativeRayCaster", **CAM_BASE)
    pos, quat = synth.candidate_views(room02, 64)
    for kw in (dict(), dict(gain_unknown=1.0, gain_free=0.5, gain_occupied=0.25),
               dict(gain_unknown=1.0, gain_free=0.5, gain_occupied=0.25, gain_unknown_outer=0.1,
                    gain_free_outer=0.05, gain_occupied_outer=0.025,
                    bounding_volume=dict(x_min=4, x_max=16, y_min=4, y_max=16, z_min=0.5, z_max=6,
                                         rotation=20.0))):
        got = ctx.compute_gains(caster, ctx.evaluator("VoxelTypeEvaluator", **kw), pos, quat)
        want = oc.compute_gains(O.eval_params("VoxelTypeEvaluator", **kw), pos, quat)
        assert np.array_equal(got["n_samples"], want["n_samples"])
        assert np.array_equal(got["n_visible"], want["n_visible"])
        assert np.array_equal(got["digest"], want["digest"])
        assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=0)
        assert got["gain"].max() > 0
    ctx.close()


@pytest.mark.parametrize("ekind,kw", [
    ("NaiveEvaluator", {}),
    ("FrontierEvaluator", {}),
    ("FrontierEvaluator", dict(accurate_frontiers=True)),
    ("FrontierEvaluator", dict(surface_frontiers=False, checking_distance=2.0)),
])
def test_naive_and_frontier_gains(capi, O, room02, ekind, kw):
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    pos, quat = synth.candidate_views(room02, 16, seed=9)
    for ckind in ("IterativeRayCaster", "SimpleRayCaster"):
        caster = ctx.caster(ckind, **CAM_BASE)
        oc = om.caster(ckind, **CAM_BASE)
        got = ctx.compute_gains(caster, ctx.evaluator(ekind, **kw), pos, quat)
        want = oc.compute_gains(O.eval_params(ekind, **kw), pos, quat)
        assert np.array_equal(got["digest"], want["digest"])
        assert np.array_equal(got["n_visible"], want["n_visible"])
        assert np.array_equal(got["gain"], want["gain"])  # integer-valued gains: exact
    ctx.close()


def test_multi_view_segments_and_mounting(capi, O, room02):
    """sampling_time > 0: several views per segment share one list (std::unique and the gain sum
    span them, camera_model.cpp:21-58); non-default mounting transform (camera_model.cpp:24-28)."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    cam = dict(CAM_BASE, mounting_translation=(0.1, -0.05, 0.2), mounting_rotation=(0.9, 0.1, -0.2, 0.3))
    caster = ctx.caster("IterativeRayCaster", **cam)
    oc = om.caster("IterativeRayCaster", **cam)
    pos, quat = synth.candidate_views(room02, 20, seed=13)
    # make consecutive views of a segment nearly identical so cross-view adjacency matters
    pos[1::2] = pos[0::2]
    quat[1::2] = quat[0::2]
    seg = np.repeat(np.arange(4, dtype=np.int32), 5)
    got = ctx.compute_gains(caster, ctx.evaluator("VoxelTypeEvaluator"), pos, quat, seg, 4)
    want = oc.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos, quat, seg, 4)
    for k in ("n_samples", "n_visible", "digest"):
        assert np.array_equal(got[k], want[k]), k
    assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=0)
    lst, _, _ = ctx.visible_voxels(caster, pos[:5], quat[:5])
    olst, _ = oc.visible_voxels(pos[:5], quat[:5])
    assert np.array_equal(lst, olst)
    ctx.close()


def test_stored_list_and_gain_from_voxels(capi, O, room02):
    """SimulatedSensorUpdater path: fold a stored list again (simulated_sensor_updater.cpp:18-22)."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    bv = dict(x_min=2, x_max=18, y_min=2, y_max=18, z_min=0.5, z_max=7)
    pos, quat = synth.candidate_views(room02, 3, seed=17)
    caster = ctx.caster("SimpleRayCaster", **CAM_BASE)
    oc = om.caster("SimpleRayCaster", **CAM_BASE)
    for ekind in ("VoxelTypeEvaluator", "NaiveEvaluator", "FrontierEvaluator"):
        ev = ctx.evaluator(ekind, bounding_volume=bv, gain_free=0.5)
        oe = O.eval_params(ekind, bounding_volume=bv, gain_free=0.5)
        for v in range(3):
            lst, ns, gain = ctx.visible_voxels(caster, pos[v:v + 1], quat[v:v + 1], evaluator=ev)
            want = oc.compute_gains(oe, pos[v:v + 1], quat[v:v + 1])
            assert len(lst) == want["n_visible"][0] and O.digest(lst) == want["digest"][0]
            assert np.isclose(gain, want["gain"][0], rtol=GAIN_RTOL, atol=0)
            assert np.all(O.bv_contains(oe, lst) == 1)
            assert np.array_equal(ctx.bv_contains(ev, lst), O.bv_contains(oe, lst))
            g2, left, keep = ctx.gain_from_voxels(ev, lst, want_keep=True)
            og2, oleft = om.gain_from_voxels(oe, lst)
            assert np.isclose(g2, og2, rtol=GAIN_RTOL, atol=0) and left == oleft
            assert keep.sum() == left
    ctx.close()


def test_cfg2_subset_digests(capi, O, room01):
    """BASELINE config 2 (0.1 m room, Iterative + VoxelType): oracle-checked subset of 48 views."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room01)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    oc = om.caster("IterativeRayCaster", **CAM_BASE)
    pos, quat = synth.candidate_views(room01, 4096)
    sel = slice(0, 48)
    got = ctx.compute_gains(caster, ctx.evaluator("VoxelTypeEvaluator"), pos[sel], quat[sel])
    want = oc.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos[sel], quat[sel], n_threads=8)
    for k in ("n_samples", "n_visible", "digest"):
        assert np.array_equal(got[k], want[k]), k
    assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=0)
    ctx.close()


def test_cfg2_full_size_properties(capi, room01):
    """Full 4096-segment batch: size-independent properties (no oracle at this size):
    permutation equivariance, batch-split invariance, digest == digest of the materialised list,
    gain == count of UNKNOWN-state entries of that list, no adjacent duplicates."""
    from mav_active_3d_planning_b200 import synth
    ctx = capi.Context(0)
    ctx.upload_map(room01)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    ev = ctx.evaluator("VoxelTypeEvaluator")
    pos, quat = synth.candidate_views(room01, 4096)
    full = ctx.compute_gains(caster, ev, pos, quat)
    assert full["gain"].sum() > 0 and (full["n_visible"] <= full["n_samples"]).all()
    perm = np.random.default_rng(1).permutation(4096)
    p = ctx.compute_gains(caster, ev, pos[perm], quat[perm])
    for k in full:
        assert np.array_equal(p[k], full[k][perm]), k
    half = ctx.compute_gains(caster, ev, pos[1000:1300], quat[1000:1300])
    for k in full:
        assert np.array_equal(half[k], full[k][1000:1300]), k
    from oracle import oracle_py as O
    for v in (0, 777, 4095):
        lst, ns, gain = ctx.visible_voxels(caster, pos[v:v + 1], quat[v:v + 1], evaluator=ev)
        assert len(lst) == full["n_visible"][v] and ns == full["n_samples"][v]
        assert O.digest(lst) == full["digest"][v]
        assert gain == full["gain"][v] == float((ctx.voxel_state(lst) == 2).sum())
        assert not np.any(np.all(lst[1:] == lst[:-1], axis=1))
    ctx.close()


def test_cfg4_simple_and_iterative_640x480(capi, O, room01):
    """BASELINE config 4: 640×480 f=320, 8 m range, bit-exact ordered voxel sets (ds=1 → 126×103
    rays) for both casters, plus one literal 640×480-ray view (downsampling_factor 0.19)."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room01)
    pos, quat = synth.candidate_views(room01, 4, seed=33)
    for kind in ("SimpleRayCaster", "IterativeRayCaster"):
        caster = ctx.caster(kind, **CAM_CFG4)
        oc = om.caster(kind, **CAM_CFG4)
        assert (caster.info.c_res_x, caster.info.c_res_y) == (126, 103)
        for v in range(2):
            got, ns, _ = ctx.visible_voxels(caster, pos[v:v + 1], quat[v:v + 1])
            want, ons = oc.visible_voxels(pos[v:v + 1], quat[v:v + 1])
            assert ns == ons and np.array_equal(got, want)
    kw = dict(CAM_CFG4, downsampling_factor=0.19)
    caster = ctx.caster("IterativeRayCaster", **kw)
    oc = om.caster("IterativeRayCaster", **kw)
    assert (caster.info.c_res_x, caster.info.c_res_y, caster.info.c_n_sections) == (640, 480, 8)
    got = ctx.compute_gains(caster, ctx.evaluator("VoxelTypeEvaluator"), pos[:1], quat[:1])
    want = oc.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos[:1], quat[:1])
    for k in ("n_samples", "n_visible", "digest", "gain"):
        assert np.array_equal(got[k], want[k]), k
    ctx.close()


def test_error_codes(capi):
    ctx = capi.Context(0)
    with pytest.raises(capi.A3DError) as e:
        ctx.caster("SimpleRayCaster")          # map not configured
    assert e.value.code == -3
    ctx.map_config(0.1, 16)
    with pytest.raises(capi.A3DError) as e:
        ctx.caster("SimpleRayCaster", ray_length=-1.0)   # camera_model.cpp:185-188
    assert e.value.code == -1 and "ray_length" in str(e.value)
    caster = ctx.caster("SimpleRayCaster", **CAM_BASE)
    ev = ctx.evaluator("VoxelTypeEvaluator", clear_from_parents=True)
    with pytest.raises(capi.A3DError) as e:
        ctx.compute_gains(caster, ev, np.zeros((1, 3)), np.array([[1.0, 0, 0, 0]]))
    assert e.value.code == -5
    ev = ctx.evaluator("VoxelTypeEvaluator")
    with pytest.raises(capi.A3DError) as e:
        ctx.compute_gains(caster, ev, np.zeros((2, 3)), np.array([[1.0, 0, 0, 0]] * 2),
                          view_segment=np.array([1, 0], np.int32), n_segments=2)
    assert e.value.code == -1
    ctx.close()


# ---- wavefront kernel (k_cast_wave): same gains / counts / order-independent digest -----------------
def _wave(ctx, caster, ev, *a, **kw):
    """Wavefront kernel, all CTA widths (2 warps: throughput, 8 / 4 warps: small / mid-size batches);
    they must agree bit for bit on the integer outputs."""
    ctx.set_option("cast_kernel", 2)
    try:
        ctx.set_option("wave_cta", 1)
        r = ctx.compute_gains(caster, ev, *a, digests=False, **kw)
        assert ctx.last_cast_kernel == 2
        for width in (2, 3):
            ctx.set_option("wave_cta", width)
            w = ctx.compute_gains(caster, ev, *a, digests=False, **kw)
            assert ctx.last_cast_kernel == 2
            for k in ("n_samples", "n_visible", "set_digest"):
                assert np.array_equal(r[k], w[k]), (k, width)
            assert np.allclose(r["gain"], w["gain"], rtol=1e-12, atol=0)
    finally:
        ctx.set_option("cast_kernel", 0)
        ctx.set_option("wave_cta", 0)
    for k in ("n_samples", "n_visible", "set_digest"):
        assert np.array_equal(r[k], w[k]), k
    assert np.allclose(r["gain"], w["gain"], rtol=1e-12, atol=0)
    return r


def _check_wave(got, want):
    for k in ("n_samples", "n_visible", "set_digest"):
        assert np.array_equal(got[k], want[k]), k
    assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=0)


@pytest.mark.parametrize("ckind", ["IterativeRayCaster", "SimpleRayCaster"])
@pytest.mark.parametrize("ekind,kw", [
    ("VoxelTypeEvaluator", {}),
    ("VoxelTypeEvaluator", dict(gain_unknown=1.0, gain_free=0.5, gain_occupied=0.25,
                                bounding_volume=dict(x_min=4, x_max=16, y_min=4, y_max=16, z_min=0.5,
                                                     z_max=6, rotation=20.0))),
    ("NaiveEvaluator", {}),
    ("FrontierEvaluator", dict(accurate_frontiers=True)),
    ("FrontierEvaluator", dict(surface_frontiers=False)),
    # another checking distance: the precomputed frontier bits do not apply, literal neighbour test
    ("FrontierEvaluator", dict(surface_frontiers=False, checking_distance=2.0)),
])
def test_wave_kernel_cfg1(capi, O, room02, ckind, ekind, kw):
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    caster = ctx.caster(ckind, **CAM_BASE)
    oc = om.caster(ckind, **CAM_BASE)
    pos, quat = synth.candidate_views(room02, 64)
    got = _wave(ctx, caster, ctx.evaluator(ekind, **kw), pos, quat)
    want = oc.compute_gains(O.eval_params(ekind, **kw), pos, quat, n_threads=8)
    _check_wave(got, want)
    assert got["gain"].max() > 0
    # scan-order kernel agrees on the order-independent digest too
    ref = ctx.compute_gains(caster, ctx.evaluator(ekind, **kw), pos, quat)
    assert np.array_equal(ref["set_digest"], want["set_digest"])
    assert np.array_equal(ref["digest"], want["digest"])
    ctx.close()


def test_wave_kernel_multi_view_mounting_vps8(capi, O):
    from mav_active_3d_planning_b200 import synth
    m = synth.room_map(0.2, vps=8)
    ctx, om = make_pair(capi, O, m)
    cam = dict(CAM_BASE, mounting_translation=(0.1, -0.05, 0.2), mounting_rotation=(0.9, 0.1, -0.2, 0.3))
    pos, quat = synth.candidate_views(m, 20, seed=13)
    pos[1::2] = pos[0::2]
    quat[1::2] = quat[0::2]
    seg = np.repeat(np.arange(4, dtype=np.int32), 5)
    for ckind in ("IterativeRayCaster", "SimpleRayCaster"):
        caster = ctx.caster(ckind, **cam)
        oc = om.caster(ckind, **cam)
        got = _wave(ctx, caster, ctx.evaluator("VoxelTypeEvaluator"), pos, quat, seg, 5)
        want = oc.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos, quat, seg, 5)
        _check_wave(got, want)
        assert got["n_visible"][4] == 0
    ctx.close()


def test_wave_kernel_cfg2_and_cfg4(capi, O, room01):
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room01)
    pos, quat = synth.candidate_views(room01, 4096)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    oc = om.caster("IterativeRayCaster", **CAM_BASE)
    ev = ctx.evaluator("VoxelTypeEvaluator")
    got = _wave(ctx, caster, ev, pos[:48], quat[:48])
    want = oc.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos[:48], quat[:48], n_threads=8)
    _check_wave(got, want)
    # full 4096-segment batch: wavefront kernel == scan-order kernel on every segment
    full_w = _wave(ctx, caster, ev, pos, quat)
    full_s = ctx.compute_gains(caster, ev, pos, quat)
    _check_wave(full_w, full_s)
    # cfg4: 640x480 f=320, 8 m, both casters (126x103 rays) and one literal 640x480-ray view
    for kind in ("SimpleRayCaster", "IterativeRayCaster"):
        c4 = ctx.caster(kind, **CAM_CFG4)
        o4 = om.caster(kind, **CAM_CFG4)
        _check_wave(_wave(ctx, c4, ev, pos[:6], quat[:6]),
                    o4.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos[:6], quat[:6], n_threads=6))
    kw = dict(CAM_CFG4, downsampling_factor=0.19)
    c4 = ctx.caster("IterativeRayCaster", **kw)
    o4 = om.caster("IterativeRayCaster", **kw)
    _check_wave(_wave(ctx, c4, ev, pos[:1], quat[:1]),
                o4.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos[:1], quat[:1]))
    ctx.close()


def test_wave_axis_aligned_bounding_volume(capi, O, room02):
    """Axis-aligned volume whose faces coincide with voxel centres / voxel faces: the wavefront
    kernel tests it on integer voxel coordinates (BvGrid), the reference on the centre doubles."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    oc = om.caster("IterativeRayCaster", **CAM_BASE)
    pos, quat = synth.candidate_views(room02, 32, seed=3)
    vs = float(np.float32(0.2))
    for bv in (dict(x_min=20.5 * vs, x_max=70.5 * vs, y_min=4.0, y_max=16.0, z_min=0.5 * vs, z_max=12.5 * vs),
               dict(x_min=4.1, x_max=13.9, y_min=30 * vs, y_max=60 * vs, z_min=-100.0, z_max=1e9),
               dict(x_min=6.0, x_max=5.0, y_min=0.0, y_max=20.0, z_min=0.0, z_max=10.0)):  # not a volume: ignored
        for ekind in ("VoxelTypeEvaluator", "NaiveEvaluator"):
            kw = dict(bounding_volume=bv)
            got = _wave(ctx, caster, ctx.evaluator(ekind, **kw), pos, quat)
            want = oc.compute_gains(O.eval_params(ekind, **kw), pos, quat, n_threads=8)
            _check_wave(got, want)
    ctx.close()


def test_wave_far_from_origin(capi, O):
    """The same room 9.6 km / 6.4 km / 1.6 km away from the origin: fp32 coordinates carry ~1e-3 m
    of rounding there, voxblox's index arithmetic produces artefact indices and most samples fail
    the wavefront kernel's certification margin (literal path)."""
    import dataclasses
    from mav_active_3d_planning_b200 import synth
    m0 = synth.room_map(0.2)
    shift = np.array([3000, -2000, 500], np.int32)
    m = dataclasses.replace(m0, block_idx=m0.block_idx + shift)
    ctx, om = make_pair(capi, O, m)
    pos, quat = synth.candidate_views(m0, 24, seed=5)
    pos = pos + shift.astype(np.float64) * float(np.float32(0.2) * np.float32(16))
    for ckind in ("IterativeRayCaster", "SimpleRayCaster"):
        caster = ctx.caster(ckind, **CAM_BASE)
        oc = om.caster(ckind, **CAM_BASE)
        for ekind in ("VoxelTypeEvaluator", "FrontierEvaluator"):
            got = _wave(ctx, caster, ctx.evaluator(ekind), pos, quat)
            want = oc.compute_gains(O.eval_params(ekind), pos, quat, n_threads=8)
            _check_wave(got, want)
    assert got["n_visible"].max() > 0
    ctx.close()


def test_wave_dense_volume_tracks_map_edits(capi, O):
    """Uploads that grow the block box, overwrites and removals between casts: the dense meta
    volume the wavefront kernel reads must follow every edit."""
    from mav_active_3d_planning_b200 import synth
    m = synth.room_map(0.2)
    rng = np.random.default_rng(17)
    order = rng.permutation(m.n_blocks)
    ctx = capi.Context(0)
    ctx.map_config(m.voxel_size, m.vps)
    om = O.OracleMap(m.voxel_size, m.vps)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    oc = om.caster("IterativeRayCaster", **CAM_BASE)
    ev, ep = ctx.evaluator("VoxelTypeEvaluator"), O.eval_params("VoxelTypeEvaluator")
    # the frontier bits of the meta words (allocated and unallocated voxels) must follow the edits too
    fr = [(ctx.evaluator("FrontierEvaluator", **kw), O.eval_params("FrontierEvaluator", **kw))
          for kw in (dict(), dict(accurate_frontiers=True, surface_frontiers=False))]
    pos, quat = synth.candidate_views(m, 16, seed=9)

    def check():
        _check_wave(_wave(ctx, caster, ev, pos, quat), oc.compute_gains(ep, pos, quat, n_threads=8))
        for fe, fp in fr:
            _check_wave(_wave(ctx, caster, fe, pos, quat), oc.compute_gains(fp, pos, quat, n_threads=8))

    third = m.n_blocks // 3
    for part in (order[:third], order[third:2 * third], order[2 * third:]):   # box grows
        for mm in (ctx, om):
            mm.upload_blocks(m.block_idx[part], m.dist[part], m.observed[part])
        check()
    sel = order[:m.n_blocks // 8]                                             # overwrite in place
    nd = (m.dist[sel] * 0.5).astype(np.float32)
    for mm in (ctx, om):
        mm.upload_blocks(m.block_idx[sel], nd, np.ones_like(m.observed[sel]))
    check()
    rem = order[m.n_blocks // 8: m.n_blocks // 4]                             # remove
    for mm in (ctx, om):
        mm.remove_blocks(m.block_idx[rem])
    check()
    far = np.array([[60, 2, 1], [-9, 0, 0]], np.int32)                        # far blocks: new box
    fd = np.full((2, m.vps ** 3), 0.9, np.float32)
    for mm in (ctx, om):
        mm.upload_blocks(far, fd, np.ones((2, m.vps ** 3), np.uint8))
    check()
    ctx.clear()
    for mm in (ctx,):
        mm.upload_blocks(m.block_idx[sel], m.dist[sel], m.observed[sel])
    om2 = O.OracleMap(m.voxel_size, m.vps)
    om2.upload_blocks(m.block_idx[sel], m.dist[sel], m.observed[sel])
    _check_wave(_wave(ctx, caster, ev, pos, quat),
                om2.caster("IterativeRayCaster", **CAM_BASE).compute_gains(ep, pos, quat, n_threads=8))
    ctx.close()


@pytest.mark.parametrize("ckind", ["IterativeRayCaster", "SimpleRayCaster"])
def test_wave_kernel_emits_ordered_lists(capi, O, room02, ckind):
    """Ordered visible-voxel lists from the wavefront kernel (per-part slot ranges + scan-order
    offsets): bit-exact against the oracle's list, single view and multi-view segment, with a
    bounding volume, both CTA widths, and identical to what the scan-order kernel writes."""
    import dataclasses
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    cam = dict(CAM_BASE, mounting_translation=(0.05, 0.0, 0.1))
    caster = ctx.caster(ckind, **cam)
    oc = om.caster(ckind, **cam)
    pos, quat = synth.candidate_views(room02, 6, seed=23)
    pos[3:] = pos[:3] + 0.07          # overlapping views: std::unique across view boundaries matters
    bv = dict(x_min=2, x_max=18, y_min=2, y_max=18, z_min=0.5, z_max=7)
    for ekind, kw in (("VoxelTypeEvaluator", {}), ("NaiveEvaluator", dict(bounding_volume=bv)),
                      ("VoxelWeightEvaluator", {})):
        ev = ctx.evaluator(ekind, **kw)
        oe = O.eval_params(ekind, **kw)
        for sl in (slice(0, 1), slice(0, 6)):
            want, _ = oc.visible_voxels(pos[sl], quat[sl])
            if kw:
                want = want[O.bv_contains(oe, want) == 1]
            wg = oc.compute_gains(oe, pos[sl], quat[sl], np.zeros(sl.stop - sl.start, np.int32), 1)
            # the kernel chain with list emission (default) ...
            lst, ns, gain = ctx.visible_voxels(caster, pos[sl], quat[sl], evaluator=ev)
            assert ctx.last_cast_kernel == 3
            assert lst.shape == want.shape and np.array_equal(lst, want)
            assert ns == wg["n_samples"][0] and np.isclose(gain, wg["gain"][0], rtol=GAIN_RTOL, atol=1e-12)
            # ... the wavefront kernel's emission ...
            ctx.set_option("cast_kernel", 2)
            for width in (1, 2):
                ctx.set_option("wave_cta", width)
                lst, ns, gain = ctx.visible_voxels(caster, pos[sl], quat[sl], evaluator=ev)
                assert ctx.last_cast_kernel == 2
                assert lst.shape == want.shape and np.array_equal(lst, want)
                assert ns == wg["n_samples"][0] and np.isclose(gain, wg["gain"][0], rtol=GAIN_RTOL, atol=1e-12)
            ctx.set_option("wave_cta", 0)
            # ... and the scan-order kernel
            ctx.set_option("cast_kernel", 1)
            ref, _, _ = ctx.visible_voxels(caster, pos[sl], quat[sl], evaluator=ev)
            ctx.set_option("cast_kernel", 0)
            assert ctx.last_cast_kernel == 1 and np.array_equal(ref, want)
        # capacity too small: the count is still reported
        with pytest.raises(capi.A3DError):
            ctx.visible_voxels(caster, pos[:1], quat[:1], evaluator=ev, cap=10)
    ctx.close()
    # 9.6 km from the origin voxblox produces artefact indices: those segments take the scan-order path
    m0 = synth.room_map(0.2)
    shift = np.array([3000, -2000, 500], np.int32)
    m = dataclasses.replace(m0, block_idx=m0.block_idx + shift)
    ctx, om = make_pair(capi, O, m)
    caster = ctx.caster(ckind, **CAM_BASE)
    oc = om.caster(ckind, **CAM_BASE)
    pos, quat = synth.candidate_views(m0, 3, seed=5)
    pos = pos + shift.astype(np.float64) * float(np.float32(0.2) * np.float32(16))
    ev = ctx.evaluator("VoxelTypeEvaluator")
    for v in range(3):
        lst, _, _ = ctx.visible_voxels(caster, pos[v:v + 1], quat[v:v + 1], evaluator=ev)
        want, _ = oc.visible_voxels(pos[v:v + 1], quat[v:v + 1])
        assert np.array_equal(lst, want)
    ctx.close()


def test_wave_kernel_empty_map_and_far_poses(capi, O):
    """Unallocated space and poses beyond the 20-bit voxel-key range (scan-order re-evaluation)."""
    ctx = capi.Context(0)
    ctx.map_config(0.2, 16)
    om = O.OracleMap(0.2, 16)
    pos = np.array([[1.0, 2.0, 3.0], [2.0e5, -3.0e5, 10.0], [104857.0, 0.3, 0.2]])
    quat = np.array([[1.0, 0, 0, 0], [0.6, 0, 0, 0.8], [1.0, 0, 0, 0]])
    for ckind in ("IterativeRayCaster", "SimpleRayCaster"):
        caster = ctx.caster(ckind, **CAM_BASE)
        oc = om.caster(ckind, **CAM_BASE)
        got = _wave(ctx, caster, ctx.evaluator("VoxelTypeEvaluator"), pos, quat)
        want = oc.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos, quat)
        _check_wave(got, want)
    ctx.close()


def test_cfg3_small_clutter_frontier(capi, O):
    """BASELINE config 3 at reduced extent: cluttered map at 0.05 m voxels (158x129 rays, 7 sections),
    frontier / unknown-voxel gain, wavefront and scan-order kernels against the oracle."""
    from mav_active_3d_planning_b200 import synth
    m = synth.clutter_map(0.05, extent=(16.0, 16.0, 8.0), n_boxes=30, obs_radius=3.5)
    assert 0.2 < m.observed.mean() < 0.95
    ctx, om = make_pair(capi, O, m)
    pos, quat = synth.candidate_views(m, 6, seed=41)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    oc = om.caster("IterativeRayCaster", **CAM_BASE)
    assert (caster.info.c_res_x, caster.info.c_res_y, caster.info.c_n_sections) == (158, 129, 7)
    for ekind, kw in (("FrontierEvaluator", {}), ("FrontierEvaluator", dict(accurate_frontiers=True)),
                      ("VoxelTypeEvaluator", {})):
        want = oc.compute_gains(O.eval_params(ekind, **kw), pos, quat, n_threads=6)
        ev = ctx.evaluator(ekind, **kw)
        _check_wave(_wave(ctx, caster, ev, pos, quat), want)
        got = ctx.compute_gains(caster, ev, pos, quat)
        for k in ("n_samples", "n_visible", "digest", "set_digest"):
            assert np.array_equal(got[k], want[k]), k
        assert np.array_equal(got["gain"], want["gain"])
        assert want["gain"].max() > 0
    ctx.close()


@pytest.mark.parametrize("ckind", ["IterativeRayCaster", "SimpleRayCaster"])
def test_voxel_weight_evaluator(capi, O, room02, ckind):
    """VoxelWeightEvaluator (voxel_weight_evaluator.cpp:43-83): TSDF-weight impact of
    surface voxels + frontier / new-voxel weights; atan2/pow/sqrt in double → 1e-5 relative."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    caster = ctx.caster(ckind, **CAM_BASE)
    oc = om.caster(ckind, **CAM_BASE)
    pos, quat = synth.candidate_views(room02, 24, seed=51)
    for kw in (dict(), dict(frontier_voxel_weight=0.5, min_impact_factor=0.02, new_voxel_weight=0.05,
                            ray_angle_x=0.003, ray_angle_y=0.002, accurate_frontiers=True,
                            bounding_volume=dict(x_min=3, x_max=17, y_min=3, y_max=17, z_min=0.5, z_max=8)),
               # another checking distance: the precomputed frontier bits do not apply
               dict(frontier_voxel_weight=0.5, new_voxel_weight=0.05, checking_distance=1.5)):
        ev = ctx.evaluator("VoxelWeightEvaluator", **kw)
        oe = O.eval_params("VoxelWeightEvaluator", **kw)
        got = ctx.compute_gains(caster, ev, pos, quat)
        want = oc.compute_gains(oe, pos, quat, n_threads=8)
        assert ctx.last_cast_kernel == 1          # scan-order kernel
        for k in ("n_samples", "n_visible", "digest"):
            assert np.array_equal(got[k], want[k]), k
        assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=1e-12)
        assert want["gain"].max() > 1.0
        # without digests the kernel chain evaluates it (partial sums per ray part) ...
        got2 = ctx.compute_gains(caster, ev, pos, quat, digests=False, set_digests=False)
        assert ctx.last_cast_kernel == 3
        assert np.allclose(got2["gain"], want["gain"], rtol=GAIN_RTOL, atol=1e-12)
        # ... and bit-reproducibly so (fixed summation order)
        again = ctx.compute_gains(caster, ev, pos, quat, digests=False, set_digests=False)
        assert np.array_equal(got2["gain"], again["gain"])
        # ... or the wavefront kernel (partial sums per lane)
        ctx.set_option("cast_kernel", 2)
        got2 = ctx.compute_gains(caster, ev, pos, quat, digests=False, set_digests=False)
        ctx.set_option("cast_kernel", 0)
        assert ctx.last_cast_kernel == 2
        assert np.allclose(got2["gain"], want["gain"], rtol=GAIN_RTOL, atol=1e-12)
        for k in ("n_samples", "n_visible"):
            assert np.array_equal(got2[k], want[k]), k
        # multi-view segments with an explicit origin per segment, and the stored-list fold
        seg = np.repeat(np.arange(6, dtype=np.int32), 4)
        org = pos[::4] + 0.3
        got = ctx.compute_gains(caster, ev, pos, quat, seg, 6, seg_origin=org)
        want = oc.compute_gains(oe, pos, quat, seg, 6, seg_origin=org)
        assert np.array_equal(got["digest"], want["digest"])
        assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=1e-12)
        got2 = ctx.compute_gains(caster, ev, pos, quat, seg, 6, seg_origin=org, digests=False, set_digests=False)
        assert ctx.last_cast_kernel == 3
        assert np.allclose(got2["gain"], want["gain"], rtol=GAIN_RTOL, atol=1e-12)
        assert np.array_equal(got2["n_visible"], want["n_visible"])
        lst, _, g1 = ctx.visible_voxels(caster, pos[:4], quat[:4], evaluator=ev, origin=org[0])
        assert np.isclose(g1, want["gain"][0], rtol=GAIN_RTOL)
        g2, left, _ = ctx.gain_from_voxels(ev, lst, origin=org[0])
        og2, oleft = om.gain_from_voxels(oe, lst, origin=org[0])
        assert np.isclose(g2, og2, rtol=GAIN_RTOL) and left == oleft == len(lst)
    ctx.close()


def test_iterative_ray_caster_lidar(capi, O, room02):
    """IterativeRayCasterLidar (iterative_ray_caster_lidar.cpp, lidar_model.cpp:159-165; SURVEY.md
    §8(f)-4): the iterative scheme over the spherical direction table."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    lid = dict(ray_length=5.0, field_of_view_x=360.0, field_of_view_y=30.0, resolution_x=1024,
               resolution_y=128, downsampling_factor=2.0)
    caster = ctx.caster("IterativeRayCasterLidar", **lid)
    oc = om.caster("IterativeRayCasterLidar", **lid)
    assert (caster.info.c_res_x, caster.info.c_res_y, caster.info.c_n_sections) == (
        oc.c_res_x, oc.c_res_y, oc.c_n_sections)
    for i, j in [(0, 0), (5, 3), (oc.c_res_x - 1, oc.c_res_y - 1)]:
        assert np.array_equal(caster.direction(i, j), oc.direction(i, j))
    pos, quat = synth.candidate_views(room02, 12, seed=61)
    lst, ns, _ = ctx.visible_voxels(caster, pos[:1], quat[:1])
    olst, ons = oc.visible_voxels(pos[:1], quat[:1])
    assert ns == ons and np.array_equal(lst, olst)
    ev = ctx.evaluator("VoxelTypeEvaluator")
    want = oc.compute_gains(O.eval_params("VoxelTypeEvaluator"), pos, quat, n_threads=8)
    got = ctx.compute_gains(caster, ev, pos, quat)
    for k in ("n_samples", "n_visible", "digest", "gain"):
        assert np.array_equal(got[k], want[k]), k
    _check_wave(_wave(ctx, caster, ev, pos, quat), want)
    ctx.close()


def test_device_resident_lists_and_refold(capi, O, room02):
    """stored lists kept in HBM and re-folded after map deltas without re-casting
    (SimulatedSensorUpdater, simulated_sensor_updater.cpp:18-22)."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    oc = om.caster("IterativeRayCaster", **CAM_BASE)
    pos, quat = synth.candidate_views(room02, 12, seed=71)
    keys = np.arange(1000, 1012, dtype=np.uint64)
    rng = np.random.default_rng(9)
    for ekind in ("VoxelTypeEvaluator", "NaiveEvaluator", "FrontierEvaluator"):
        kw = dict(gain_free=0.5) if ekind[0] == "V" else {}
        ev = ctx.evaluator(ekind, **kw)
        oe = O.eval_params(ekind, **kw)
        store = capi.Store(ctx)
        got = store.compute_gains(caster, ev, keys, pos, quat, digests=True)
        want = oc.compute_gains(oe, pos, quat)
        for k in ("n_visible", "n_samples", "digest", "set_digest"):
            assert np.array_equal(got[k], want[k]), (ekind, k)
        assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=0)
        lists = [store.get(k) for k in keys]
        for s in range(12):
            # what the reference holds after computeGain: Naive / Frontier erased the observed entries
            ref_list, _ = oc.visible_voxels(pos[s:s + 1], quat[s:s + 1])
            if ekind[0] != "V":
                ref_list = ref_list[om.is_observed(ref_list) == 0]
            assert np.array_equal(lists[s], ref_list), (ekind, s)
        assert store.nbytes == 8 * int(want["n_visible"].sum())   # packed voxel keys
        # two rounds of map deltas: more of the map becomes observed, distances move
        for _ in range(2):
            sel = rng.choice(room02.n_blocks, room02.n_blocks // 5, replace=False)
            nd = (room02.dist[sel] + rng.normal(0, 0.05, room02.dist[sel].shape)).astype(np.float32)
            no = np.maximum(room02.observed[sel], (rng.random(room02.observed[sel].shape) < 0.5)).astype(np.uint8)
            ctx.upload_blocks(room02.block_idx[sel], nd, no)
            om.upload_blocks(room02.block_idx[sel], nd, no)
            gains, left = store.refold(ev, keys)
            for s in range(12):
                og, oleft = om.gain_from_voxels(oe, lists[s])
                assert np.isclose(gains[s], og, rtol=GAIN_RTOL, atol=0) and left[s] == oleft
                if ekind[0] != "V":   # the reference erases observed entries from the stored list
                    lists[s] = lists[s][om.is_observed(lists[s]) == 0] if len(lists[s]) else lists[s]
                    assert np.array_equal(store.get(keys[s]), lists[s])
        # restore the map for the next evaluator, replace half of the keys, erase the rest
        ctx.upload_blocks(room02.block_idx, room02.dist, room02.observed, room02.weight)
        om.upload_blocks(room02.block_idx, room02.dist, room02.observed, room02.weight)
        store.compute_gains(caster, ev, keys[:6], pos[:6], quat[:6])
        store.erase(keys[6:])
        assert store.list_size(keys[7]) == -1 and store.list_size(keys[0]) == len(store.get(keys[0]))
        assert store.nbytes == 8 * int(want["n_visible"][:6].sum())
        with pytest.raises(capi.A3DError):
            store.refold(ev, keys[6:8])
        with pytest.raises(capi.A3DError):   # repeated keys in one call
            store.refold(ev, keys[[0, 0]])
        with pytest.raises(capi.A3DError):
            store.compute_gains(caster, ev, keys[[1, 1]], pos[:2], quat[:2])
        store.close()
    ctx.close()


def test_store_clear_from_parents(capi, O, room02):
    """Batch clear_from_parents on the device (simulated_sensor_evaluator.cpp:60-76): ancestors given by
    store keys, in the same batch and from earlier calls; every evaluator; against the oracle's tree."""
    from mav_active_3d_planning_b200 import synth
    import golden_cases as G
    ctx, om = make_pair(capi, O, room02)
    pos, quat = G.chain_poses(*synth.candidate_views(room02, 8))
    parent = G.CHAIN_PARENT
    keys = np.arange(1, 9, dtype=np.uint64) * 1000
    pkeys = np.where(parent < 0, 0, keys[np.maximum(parent, 0)]).astype(np.uint64)
    for ckind in ("IterativeRayCaster", "SimpleRayCaster"):
        caster = ctx.caster(ckind, **CAM_BASE)
        oc = om.caster(ckind, **CAM_BASE)
        for ekind, kw in (("VoxelTypeEvaluator", dict(gain_free=0.5, gain_occupied=0.25)), ("NaiveEvaluator", {}),
                          ("FrontierEvaluator", {}), ("VoxelWeightEvaluator", {})):
            ev = ctx.evaluator(ekind, clear_from_parents=True, **kw)
            want = oc.compute_gains(O.eval_params(ekind, clear_from_parents=True, **kw), pos, quat, parent=parent)
            # (a) the whole tree in one call
            got = ctx.compute_gains(caster, ev, pos, quat, parent=parent)
            for k in ("n_visible", "n_samples", "digest", "set_digest"):
                assert np.array_equal(got[k], want[k]), (ckind, ekind, k)
            assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=0)
            # (b) like the planner: the tree grows call by call, parents are already in the store
            store = capi.Store(ctx)
            for lo, hi in ((0, 1), (1, 3), (3, 8)):
                r = store.compute_gains(caster, ev, keys[lo:hi], pos[lo:hi], quat[lo:hi], parent_keys=pkeys[lo:hi],
                                        digests=True)
                for k in ("n_visible", "digest"):
                    assert np.array_equal(r[k], want[k][lo:hi]), (ckind, ekind, k, lo)
                assert np.allclose(r["gain"], want["gain"][lo:hi], rtol=GAIN_RTOL, atol=0)
            with pytest.raises(capi.A3DError):
                store.compute_gains(caster, ev, keys[:1] + 7, pos[:1], quat[:1], parent_keys=np.array([424242], np.uint64))
            store.close()
    ctx.close()


def test_map_reconfigure_after_clear(capi, O):
    """a3d_map_config with a different voxels_per_side after a3d_map_clear: slabs are re-sized, objects
    built for the old geometry are refused."""
    from mav_active_3d_planning_b200 import synth
    m8, m16 = synth.room_map(0.2, 8), synth.room_map(0.2, 16)
    ctx = capi.Context(0)
    ctx.upload_map(m8)
    old_caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    old_ev = ctx.evaluator("VoxelTypeEvaluator")
    pos, quat = synth.candidate_views(m16, 6, seed=3)
    ctx.compute_gains(old_caster, old_ev, pos, quat)
    ctx.clear()
    ctx.upload_map(m16)     # map_config(0.2, 16) + upload
    with pytest.raises(capi.A3DError) as e:
        ctx.compute_gains(old_caster, old_ev, pos, quat)
    assert e.value.code == -3
    om = O.OracleMap.from_synth(m16)
    pts = probe_points(m16, 5000, 5)
    assert np.array_equal(ctx.voxel_state(pts), om.voxel_state(pts))
    assert np.array_equal(ctx.voxel_weight(pts), om.voxel_weight(pts))
    caster, ev = ctx.caster("IterativeRayCaster", **CAM_BASE), ctx.evaluator("VoxelTypeEvaluator")
    want = om.caster("IterativeRayCaster", **CAM_BASE).compute_gains(O.eval_params("VoxelTypeEvaluator"), pos, quat)
    got = ctx.compute_gains(caster, ev, pos, quat)
    for k in ("n_visible", "n_samples", "digest"):
        assert np.array_equal(got[k], want[k]), k
    ctx.close()


def test_esdf_only_delta_keeps_tsdf_weights(capi, O, room02):
    """A delta without weights (ESDF-only layer message, the multi-GPU broadcast payload) must not wipe
    the TSDF weights of blocks that already exist: VoxelWeightEvaluator parity afterwards."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    rng = np.random.default_rng(21)
    sel = rng.choice(room02.n_blocks, room02.n_blocks // 3, replace=False)
    nd = (room02.dist[sel] + rng.normal(0, 0.02, room02.dist[sel].shape)).astype(np.float32)
    ctx.upload_blocks(room02.block_idx[sel], nd, room02.observed[sel])   # no weights: the TSDF layer stays
    om.upload_blocks(room02.block_idx[sel], nd, room02.observed[sel])    # (same contract in the checker)
    assert np.count_nonzero(om.voxel_weight(probe_points(room02, 500, 6))) > 50
    pts = probe_points(room02, 5000, 6)
    assert np.array_equal(ctx.voxel_weight(pts), om.voxel_weight(pts))
    pos, quat = synth.candidate_views(room02, 8, seed=4)
    caster, ev = ctx.caster("IterativeRayCaster", **CAM_BASE), ctx.evaluator("VoxelWeightEvaluator")
    want = om.caster("IterativeRayCaster", **CAM_BASE).compute_gains(O.eval_params("VoxelWeightEvaluator"), pos, quat)
    got = ctx.compute_gains(caster, ev, pos, quat, digests=False, set_digests=False)
    assert np.allclose(got["gain"], want["gain"], rtol=GAIN_RTOL, atol=0)
    ctx.close()


def test_filter_not_in_is_std_find_on_the_three_doubles(capi):
    """a3d_filter_not_in (clear_from_parents, simulated_sensor_evaluator.cpp:60-76) as a hash-set probe:
    same answers as std::find with operator== on Eigen::Vector3d — duplicates on both sides, -0.0 == 0.0,
    NaN equal to nothing, entries that differ in one ulp of one coordinate, empty lists."""
    rng = np.random.default_rng(11)
    ctx = capi.Context(0)
    for n, m in [(0, 5), (5, 0), (1, 1), (1000, 37), (20000, 150000), (150000, 20000)]:
        grid = rng.integers(-40, 40, (max(m, 1), 3)) * 0.1 + 0.05
        other = grid[:m].copy()
        a = np.concatenate([other[rng.integers(0, max(m, 1), n // 2)] if m else np.zeros((0, 3)),
                            rng.integers(-40, 40, (n - (n // 2 if m else 0), 3)) * 0.1 + 0.05])
        if n >= 1000:
            a[0] = other[0]; a[0, 1] = np.nextafter(a[0, 1], 1e9)      # one ulp off: not equal
            a[1] = (0.0, -0.0, 0.0); other[1] = (-0.0, 0.0, -0.0)        # signed zeros are equal
            a[2] = (np.nan, 1.0, 2.0); other[2] = (np.nan, 1.0, 2.0)     # NaN equals nothing
            a[3] = a[4] = other[5]                                       # duplicates
            other[6] = other[5]
        keep = ctx.filter_not_in(a, other)
        with np.errstate(invalid="ignore"):
            if n and m:
                # exact reference by sorting on the bit patterns (zeros canonicalised, NaN rows never match)
                ca, co = a + 0.0, other + 0.0
                okey = {tuple(r) for r in co[~np.isnan(co).any(axis=1)].tolist()}
                want = np.array([0 if (not np.isnan(r).any() and tuple(r) in okey) else 1 for r in ca], np.uint8)
            else:
                want = np.ones(n, np.uint8)
        assert np.array_equal(keep, want), (n, m)
    ctx.close()


def test_voxel_tsdf_distance(capi, O, room02):
    """VoxbloxMap::getVoxelDistance (voxblox.cpp:84-98) = the getVoxelWeight lookup (:101-115) on the other
    member of the TSDF voxel: checked against the oracle's weight lookup on a map whose weight member holds
    the TSDF distances.  0 before any upload, for blocks never sent, for unknown blocks (reported through the
    applied mask) and for a block that was removed and allocated again."""
    m = room02
    rng = np.random.default_rng(3)
    tsdf = rng.normal(0.0, 0.3, m.dist.shape).astype(np.float32)
    ctx = capi.Context(0)
    ctx.upload_map(m)
    pts = probe_points(m, 4000, 17)
    assert not ctx.voxel_tsdf_distance(pts).any()
    half = m.n_blocks // 2
    unknown = np.array([[900, 900, 900]], np.int32)
    applied = ctx.upload_tsdf_distances(np.concatenate([m.block_idx[:half], unknown]),
                                        np.concatenate([tsdf[:half], tsdf[:1]]))
    assert applied[:half].all() and not applied[half]
    want_src = tsdf.copy()
    want_src[half:] = 0.0
    om = O.OracleMap(m.voxel_size, m.vps)
    om.upload_blocks(m.block_idx, m.dist, m.observed, want_src)
    want = om.voxel_weight(pts)
    got = ctx.voxel_tsdf_distance(pts)
    assert np.array_equal(got, want) and np.count_nonzero(got) > 1000
    # the weights are a separate member
    assert np.array_equal(ctx.voxel_weight(pts), O.OracleMap.from_synth(m).voxel_weight(pts))
    # a removed block's slot is reused by the next new block: it starts without TSDF distances
    ctx.remove_blocks(m.block_idx[:1])
    ctx.upload_blocks(m.block_idx[:1], m.dist[:1], m.observed[:1], m.weight[:1])
    want_src[0] = 0.0
    om2 = O.OracleMap(m.voxel_size, m.vps)
    om2.upload_blocks(m.block_idx, m.dist, m.observed, want_src)
    assert np.array_equal(ctx.voxel_tsdf_distance(pts), om2.voxel_weight(pts))
    ctx.close()


def test_refold_without_parents_voxelweight_and_bounding_volume(capi, O, room02):
    """The re-fold kernel for lists nothing can leave (k_list_fold_plain: VoxelType / VoxelWeight, no parents,
    no digests) on the cases the main re-fold test does not reach: VoxelWeightEvaluator with a reference
    point per segment (multi-view segments), VoxelTypeEvaluator with a rotated bounding volume — gains after
    a map delta against the oracle's fold of the same lists."""
    from mav_active_3d_planning_b200 import synth
    ctx, om = make_pair(capi, O, room02)
    caster = ctx.caster("IterativeRayCaster", **CAM_BASE)
    pos, quat = synth.candidate_views(room02, 12, seed=33)
    seg = np.repeat(np.arange(6, dtype=np.int32), 2)
    org = pos[1::2] + 0.2
    keys = np.arange(1, 7, dtype=np.uint64)
    bv = dict(x_min=3.0, x_max=16.0, y_min=2.0, y_max=15.0, z_min=0.5, z_max=6.0, rotation=25.0)
    rng = np.random.default_rng(5)
    sel = rng.choice(room02.n_blocks, room02.n_blocks // 4, replace=False)
    nd = (room02.dist[sel] + rng.normal(0, 0.05, room02.dist[sel].shape)).astype(np.float32)
    no = np.maximum(room02.observed[sel], rng.random(room02.observed[sel].shape) < 0.4).astype(np.uint8)
    for ekind, kw, origin in (("VoxelWeightEvaluator", dict(frontier_voxel_weight=0.5, new_voxel_weight=0.02), org),
                              ("VoxelWeightEvaluator", dict(frontier_voxel_weight=0.0, min_impact_factor=0.1), org),
                              ("VoxelTypeEvaluator", dict(gain_free=0.25, gain_occupied=2.0, gain_unknown_outer=0.5,
                                                          bounding_volume=bv), None)):
        ctx.upload_blocks(room02.block_idx, room02.dist, room02.observed, room02.weight)
        om.upload_blocks(room02.block_idx, room02.dist, room02.observed, room02.weight)
        ev, oe = ctx.evaluator(ekind, **kw), O.eval_params(ekind, **kw)
        store = capi.Store(ctx)
        store.compute_gains(caster, ev, keys, pos, quat, view_segment=seg, seg_origin=origin)
        lists = [store.get(k) for k in keys]
        ctx.upload_blocks(room02.block_idx[sel], nd, no)
        om.upload_blocks(room02.block_idx[sel], nd, no)
        gains, left = store.refold(ev, keys, seg_origin=origin)
        for s in range(6):
            og, oleft = om.gain_from_voxels(oe, lists[s], origin=None if origin is None else origin[s])
            assert np.isclose(gains[s], og, rtol=GAIN_RTOL, atol=1e-12) and left[s] == oleft == len(lists[s]), (ekind, s)
        assert np.array_equal(store.refold(ev, keys, seg_origin=origin)[0], gains)   # reproducible
        store.close()
    ctx.close()
