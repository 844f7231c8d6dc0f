"""The cases behind tests/golden/ref_goldens.npz: inputs are regenerated from seeds on both sides;
the expected outputs were produced by oracle/_ref (the reference's own sources, see
tests/golden/make_ref_goldens.py).  Shared by the generator, the CPU check and the GPU suite."""
import numpy as np

from conftest import CAM_BASE, CAM_CFG4

MOUNT = dict(mounting_translation=(0.1, 0.02, -0.05), mounting_rotation=(0.9, 0.1, -0.2, 0.3))
BV = dict(x_min=2.0, x_max=15.0, y_min=1.0, y_max=12.0, z_min=0.5, z_max=6.0, rotation=10.0)
GAINS = dict(gain_unknown=1.0, gain_occupied=0.5, gain_free=0.25, gain_unknown_outer=0.1,
             gain_occupied_outer=0.05, gain_free_outer=0.025)
LIDAR = dict(resolution_x=64, resolution_y=32, ray_length=5.0, field_of_view_x=360.0, field_of_view_y=30.0)

EVALS = {
    "vt": ("VoxelTypeEvaluator", {}),
    "vt_bv": ("VoxelTypeEvaluator", dict(GAINS, bounding_volume=BV)),
    "naive": ("NaiveEvaluator", {}),
    "naive_bv": ("NaiveEvaluator", dict(bounding_volume=BV)),
    "fr": ("FrontierEvaluator", {}),
    "fr26": ("FrontierEvaluator", dict(accurate_frontiers=True)),
    "fr_nosurf": ("FrontierEvaluator", dict(surface_frontiers=False, checking_distance=2.0)),
    "vw": ("VoxelWeightEvaluator", {}),
    "vw2": ("VoxelWeightEvaluator", dict(frontier_voxel_weight=0.0, min_impact_factor=0.2, new_voxel_weight=0.5)),
}


def general_quats(n, seed):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def chain_poses(pos, quat):
    """Children that overlap their parents (clear_from_parents)."""
    pos, quat = pos.copy(), quat.copy()
    pos[1] = pos[0] + [0.2, 0.0, 0.0]
    pos[2] = pos[0] + [0.2, 0.2, 0.0]
    pos[3] = pos[0] + [0.4, 0.2, 0.1]
    quat[1:4] = quat[0]
    return pos, quat


CHAIN_PARENT = np.array([-1, 0, 1, 2, -1, 4, 4, 6], np.int32)
MULTI_VSEG = np.array([0, 0, 0, 1, 2, 2, 3, 3], np.int32)   # 5 segments, segment 4 has no view


LIST_CASES = ("cfg1_iter", "cfg1_simple_mount", "cfg1_lidar")


def cases():
    """(name, map_key, caster kind, camera kwargs, pose spec, eval keys, variant)"""
    out = []
    for kind, tag in (("SimpleRayCaster", "simple"), ("IterativeRayCaster", "iter")):
        out.append((f"cfg1_{tag}", "room02", kind, CAM_BASE, ("views", 8, 42), list(EVALS), "plain"))
        out.append((f"cfg1_{tag}_mount", "room02", kind, dict(CAM_BASE, **MOUNT), ("general", 8, 5), ["vt", "naive", "vw"], "plain"))
        out.append((f"cfg1_{tag}_multi", "room02", kind, CAM_BASE, ("views", 8, 42), ["vt_bv", "naive", "fr", "vw"], "multi"))
        out.append((f"cfg1_{tag}_chain", "room02", kind, CAM_BASE, ("chain", 8, 42), ["vt", "vt_bv", "naive", "fr", "vw"], "chain"))
    out.append(("cfg1_lidar", "room02", "IterativeRayCasterLidar", LIDAR, ("general", 6, 9), ["vt", "naive", "fr"], "plain"))
    out.append(("cfg2_iter", "room01", "IterativeRayCaster", CAM_BASE, ("views", 6, 42), ["vt", "vw"], "plain"))
    out.append(("cfg2_simple", "room01", "SimpleRayCaster", CAM_BASE, ("views", 2, 42), ["vt"], "plain"))
    out.append(("cfg4_iter", "room01", "IterativeRayCaster", CAM_CFG4, ("views", 2, 7), ["vt"], "plain"))
    out.append(("cfg4_simple", "room01", "SimpleRayCaster", CAM_CFG4, ("views", 1, 7), ["vt"], "plain"))
    return out


def poses(synth, m, spec):
    mode, n, seed = spec
    pos, quat = synth.candidate_views(m, n, seed=seed)
    if mode == "general":
        quat = general_quats(n, seed)
    elif mode == "chain":
        pos, quat = chain_poses(pos, quat)
    return pos, quat


def call_kwargs(variant):
    if variant == "multi":
        return dict(view_segment=MULTI_VSEG, n_segments=5)
    if variant == "chain":
        return dict(parent=CHAIN_PARENT)
    return {}
