"""Known-answer tests that pin the CPU oracle to facts derivable by hand from the reference source
(SURVEY.md §4, Appendix B).  The reference ships no tests or golden vectors (parity unpinned)."""
import itertools

import numpy as np
import pytest

from conftest import CAM_BASE
from oracle import oracle_py as O


def free_map(vs, radius_m=5.6, dist=1.0):
    om = O.OracleMap(vs, 16)
    r = int(np.ceil(radius_m / (vs * 16))) + 1
    idx = np.array(list(itertools.product(range(-r, r + 1), repeat=3)), np.int32)
    om.upload_blocks(idx, np.full((len(idx), 4096), dist, np.float32),
                     np.ones((len(idx), 4096), np.uint8))
    return om


POSE = (np.array([[0.03, 0.04, 0.05]]), np.array([[1.0, 0.0, 0.0, 0.0]]))


@pytest.mark.parametrize("vs,res,nsec,spr,simple_s,iter_s", [
    (0.2, (40, 33), 5, 25, 33000, 19529),
    (0.1, (79, 65), 6, 50, 256750, 147904),
])
def test_derived_constants_and_free_space_counts(built, vs, res, nsec, spr, simple_s, iter_s):
    # iterative_ray_caster.cpp:24-45 and Appendix B
    om = free_map(vs)
    it = om.caster("IterativeRayCaster", **CAM_BASE)
    si = om.caster("SimpleRayCaster", **CAM_BASE)
    assert (it.c_res_x, it.c_res_y) == res and it.c_n_sections == nsec
    assert list(it.split_widths) == [2 ** (nsec - 1 - t) for t in range(nsec)] + [0]
    assert list(it.split_distances) == [0.0] + [5.0 / 2 ** (nsec - t) for t in range(1, nsec + 1)]
    assert it.ray_step == float(np.float32(vs))
    _, ns = si.visible_voxels(*POSE)
    assert ns == simple_s == res[0] * res[1] * spr
    _, ni = it.visible_voxels(*POSE)
    assert ni == iter_s


def test_shipped_config_resolutions(built):
    # app/cfg/planners/example_config.yaml:64-71 (ds=3, 640x480 f=320) → 27x22, 4 sections
    om = free_map(0.1, 1.0)
    c = om.caster("IterativeRayCaster", resolution_x=640, resolution_y=480, focal_length=320.0,
                  ray_length=5.0, downsampling_factor=3.0)
    assert (c.c_res_x, c.c_res_y, c.c_n_sections) == (27, 22, 4)
    # reconstruction_planner.yaml:83-90 (172x480, ds=5) → 6x13, 2 sections
    c = om.caster("IterativeRayCaster", resolution_x=172, resolution_y=480, focal_length=320.0,
                  ray_length=5.0, downsampling_factor=5.0)
    assert (c.c_res_x, c.c_res_y, c.c_n_sections) == (6, 13, 2)


def test_direction_vector(built):
    # camera_model.cpp:161-168: centre-ish ray points along +x; corners are symmetric
    om = free_map(0.2, 1.0)
    c = om.caster("SimpleRayCaster", **CAM_BASE)
    d00 = c.direction(0, 0)
    d11 = c.direction(c.c_res_x - 1, c.c_res_y - 1)
    assert np.allclose(np.linalg.norm(d00), 1.0)
    assert d00[0] == d11[0] and d00[1] == -d11[1] and d00[2] == -d11[2]
    assert d00[1] > 0 and d00[2] > 0  # y left, z up
    # fov 90° horizontally: corner ray's horizontal angle is 45°
    assert np.isclose(np.arctan2(d00[1], d00[0]), np.pi / 4)


def test_simple_excludes_iterative_includes_hit_voxel(built):
    # D11: a wall of occupied voxels 2 m ahead; a single central ray (1-ray-wide grid impossible,
    # so compare per-view totals): Iterative's list contains OCCUPIED centres, Simple's never does.
    vs = 0.2
    om = O.OracleMap(vs, 16)
    r = 3
    idx = np.array(list(itertools.product(range(-r, r + 1), repeat=3)), np.int32)
    nv = 4096
    dist = np.zeros((len(idx), nv), np.float32)
    lin = np.arange(nv)
    vx = lin % 16
    for k, b in enumerate(idx):
        x = (b[0] * 16 + vx + 0.5) * vs
        dist[k] = np.abs(2.0 - x) - 0.0  # distance to the plane x = 2
    om.upload_blocks(idx, dist, np.ones((len(idx), nv), np.uint8))
    si = om.caster("SimpleRayCaster", **CAM_BASE)
    it = om.caster("IterativeRayCaster", **CAM_BASE)
    vs_list, _ = si.visible_voxels(*POSE)
    vi_list, _ = it.visible_voxels(*POSE)
    assert len(vs_list) and len(vi_list)
    assert not np.any(om.voxel_state(vs_list) == 0)
    assert np.any(om.voxel_state(vi_list) == 0)
    # nothing behind the wall is seen by either
    assert vs_list[:, 0].max() < 2.0 + vs and vi_list[:, 0].max() < 2.0 + 2 * vs


def test_interpolation_needs_all_eight_observed(built):
    # Appendix A.3: UNKNOWN unless the base voxel and its +1 neighbours are allocated and observed,
    # including across a block face.
    vs = 0.2
    om = O.OracleMap(vs, 16)
    one = np.array([[0, 0, 0]], np.int32)
    om.upload_blocks(one, np.full((1, 4096), 1.0, np.float32), np.ones((1, 4096), np.uint8))
    c = lambda i: (i + 0.5) * vs
    inside = np.array([[c(3) + 0.01, c(4) + 0.01, c(5) + 0.01]])
    assert om.voxel_state(inside)[0] == 1
    # base voxel 15 along x needs block (1,0,0)
    edge = np.array([[c(15) + 0.01, c(4) + 0.01, c(5) + 0.01]])
    assert om.voxel_state(edge)[0] == 2
    om.upload_blocks(np.array([[1, 0, 0]], np.int32), np.full((1, 4096), 1.0, np.float32),
                     np.ones((1, 4096), np.uint8))
    assert om.voxel_state(edge)[0] == 1
    # below the first voxel centre the base shifts into block (-1,0,0)
    low = np.array([[c(0) - 0.01, c(4) + 0.01, c(5) + 0.01]])
    assert om.voxel_state(low)[0] == 2
    # one unobserved corner → UNKNOWN
    obs = np.ones((1, 4096), np.uint8)
    obs[0, 4 + 16 * (4 + 16 * 6)] = 0  # voxel (4,4,6)
    om.upload_blocks(one, np.full((1, 4096), 1.0, np.float32), obs)
    assert om.voxel_state(inside)[0] == 2
    # occupied iff interpolated distance < voxel size
    om.upload_blocks(one, np.full((1, 4096), 0.1, np.float32), np.ones((1, 4096), np.uint8))
    assert om.voxel_state(inside)[0] == 0


def test_interpolation_is_trilinear(built):
    vs = 0.1
    om = O.OracleMap(vs, 8)
    lin = np.arange(512)
    x, y, z = lin % 8, (lin // 8) % 8, lin // 64
    f = (0.3 * x + 0.2 * y - 0.1 * z + 0.05 * x * y).astype(np.float32)
    om.upload_blocks(np.array([[0, 0, 0]], np.int32), f[None], np.ones((1, 512), np.uint8))
    rng = np.random.default_rng(0)
    g = 1.0 + rng.random((200, 3)) * 5.0  # continuous voxel coords (centres at integer+0.5)
    pts = g * float(np.float32(vs))
    d, ok = om.distance(pts)
    u = g - 0.5
    want = 0.3 * u[:, 0] + 0.2 * u[:, 1] - 0.1 * u[:, 2] + 0.05 * u[:, 0] * u[:, 1]
    assert ok.all() and np.allclose(d, want, atol=2e-5)


def test_voxel_center_and_observed(built):
    vs = 0.2
    om = O.OracleMap(vs, 16)
    om.upload_blocks(np.array([[0, 0, 0]], np.int32), np.full((1, 4096), 1.0, np.float32),
                     np.ones((1, 4096), np.uint8))
    p = np.array([[0.31, 1.05, 2.99], [-0.01, 0.0, 3.3]])
    c = om.voxel_center(p)
    assert np.allclose(c, [[0.3, 1.1, 2.9], [-0.1, 0.1, 3.3]], atol=1e-6)
    assert list(om.is_observed(p)) == [1, 0]  # second point lies in unallocated block (-1,0,1)


def test_voxel_type_fallthrough_and_naive_duplicates(built):
    # D7 and D6
    vs = 0.2
    om = O.OracleMap(vs, 16)
    nv = 4096
    lin = np.arange(nv)
    x = lin % 16
    dist = np.where(x < 8, 1.0, 0.05).astype(np.float32)[None]  # x<8 free, x>=8 occupied
    obs = np.ones((1, nv), np.uint8)
    idx = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1],
                    [1, 1, 1]], np.int32)
    om.upload_blocks(idx, np.repeat(dist, 8, 0), np.repeat(obs, 8, 0))
    c = lambda i: (i + 0.5) * float(np.float32(vs))
    free_v = [c(2), c(2), c(2)]
    occ_v = [c(12), c(2), c(2)]
    unk_v = [c(2) - 20.0, c(2), c(2)]
    lst = np.array([free_v, occ_v, unk_v, unk_v, free_v])
    assert list(om.voxel_state(lst)) == [1, 0, 2, 2, 1]
    e = O.eval_params("VoxelTypeEvaluator", gain_unknown=1.0, gain_free=0.5, gain_occupied=0.25)
    g, _ = om.gain_from_voxels(e, lst)
    # UNKNOWN adds 1+.5+.25, FREE adds .5+.25, OCCUPIED adds .25
    assert g == 2 * 1.75 + 2 * 0.75 + 0.25
    g, left = om.gain_from_voxels(O.eval_params("NaiveEvaluator"), lst)
    assert g == 2.0 and left == 2  # both copies of the unknown voxel are counted
    # outer gains apply outside the bounding volume
    e = O.eval_params("VoxelTypeEvaluator", gain_unknown=1.0, gain_unknown_outer=0.125,
                      bounding_volume=dict(x_min=0, x_max=5, y_min=0, y_max=5, z_min=0, z_max=5))
    g, _ = om.gain_from_voxels(e, lst)
    assert g == 2 * 0.125


def test_bounding_volume(built):
    # bounding_volume.cpp:33-57
    e = O.eval_params("VoxelTypeEvaluator")
    assert e.bv_is_setup == 0
    assert O.bv_contains(e, np.array([[1e9, 0, 0]]))[0] == 1  # not set up → always true
    e = O.eval_params("VoxelTypeEvaluator",
                      bounding_volume=dict(x_min=-1, x_max=1, y_min=-2, y_max=2, z_min=0, z_max=1))
    pts = np.array([[0, 0, 0.5], [1.0, 2.0, 1.0], [1.01, 0, 0.5], [0, 0, -0.01]])
    assert list(O.bv_contains(e, pts)) == [1, 1, 0, 0]  # closed intervals
    e = O.eval_params("VoxelTypeEvaluator",
                      bounding_volume=dict(x_min=-1, x_max=1, y_min=-0.1, y_max=0.1, z_min=0,
                                           z_max=1, rotation=90.0))
    # rotated by -90° about z: world +y maps onto box +x
    assert list(O.bv_contains(e, np.array([[0, 0.9, 0.5], [0.9, 0, 0.5]]))) == [1, 0]


def test_sample_viewpoints(built):
    # camera_model.cpp:62-83
    t = np.array([0, 100, 200, 300, 400, 500], np.int64) * 1_000_000
    assert list(O.sample_viewpoints(0.0, t)) == [5]
    assert list(O.sample_viewpoints(1.0, t)) == [5]
    assert list(O.sample_viewpoints(0.2, t)) == [2, 4]
    assert list(O.sample_viewpoints(0.15, t)) == [2, 3, 5]


def test_adjacent_unique_only(built, room02):
    # D6: the per-segment list may contain non-adjacent duplicates but never adjacent ones
    from mav_active_3d_planning_b200 import synth
    om = O.OracleMap.from_synth(room02)
    c = om.caster("IterativeRayCaster", **CAM_BASE)
    pos, quat = synth.candidate_views(room02, 2, seed=3)
    v, ns = c.visible_voxels(pos, quat)
    assert len(v) <= ns
    assert not np.any(np.all(v[1:] == v[:-1], axis=1))
    assert len(np.unique(v, axis=0)) < len(v)
    assert O.digest(v) != O.digest(v[::-1])


def test_frontier_order(built):
    # frontier_evaluator.cpp:69-98: first non-UNKNOWN neighbour decides
    vs = 0.2
    om = O.OracleMap(vs, 16)
    nv = 4096
    lin = np.arange(nv)
    x, y = lin % 16, (lin // 16) % 16
    # x<8 unobserved; observed part: y<8 free, y>=8 occupied.  (With the 8-corner interpolation a
    # neighbour is only "known" if its +1 corners are observed too, so the frontier is detectable
    # from the unobserved voxel whose +x neighbour is observed, not the other way round.)
    obs = (x >= 8).astype(np.uint8)[None]
    dist = np.where(y < 8, 1.0, 0.05).astype(np.float32)[None]
    idx = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [0, -1, 0], [0, -1, 1]], np.int32)
    om.upload_blocks(idx, np.repeat(dist, 6, 0), np.repeat(obs, 6, 0))
    c = lambda i: (i + 0.5) * float(np.float32(vs))
    # unobserved voxel (7, 3, 3): +x neighbour is observed free → not a surface frontier
    v_free = [c(7), c(3), c(3)]
    # unobserved voxel (7, 12, 3): +x neighbour occupied → surface frontier
    v_occ = [c(7), c(12), c(3)]
    lst = np.array([v_free, v_occ, [c(11), c(3), c(3)]])  # last one is observed → erased
    g, left = om.gain_from_voxels(O.eval_params("FrontierEvaluator"), lst)
    assert (g, left) == (1.0, 2)
    g, left = om.gain_from_voxels(O.eval_params("FrontierEvaluator", surface_frontiers=False), lst)
    assert (g, left) == (2.0, 2)
