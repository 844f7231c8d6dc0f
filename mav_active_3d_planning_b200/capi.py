"""ctypes binding of the C ABI in ``include/a3d.h`` (``liba3d_b200.so``).

This is the host-side mirror used by the tests, the benchmark and Python callers.  It mirrors the
reference's interface for the path: a *map* (``VoxbloxMap``, active_3d_planning_voxblox/src/map/
voxblox.cpp), a *sensor model* (``SimpleRayCaster`` / ``IterativeRayCaster``) and a
``SimulatedSensorEvaluator`` subclass (``VoxelTypeEvaluator`` / ``NaiveEvaluator`` /
``FrontierEvaluator``), with the reference's parameter names and defaults.

There is no fallback of any kind: if the shared library is missing or no B200 is usable, the
constructors raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("A3D_LIB_PATH") or os.path.join(_HERE, "liba3d_b200.so")

OCCUPIED, FREE, UNKNOWN = 0, 1, 2
CASTER_SIMPLE, CASTER_ITERATIVE, CASTER_ITERATIVE_LIDAR = 0, 1, 2
EVAL_VOXEL_TYPE, EVAL_NAIVE, EVAL_FRONTIER, EVAL_VOXEL_WEIGHT = 0, 1, 2, 3
_CASTER_KINDS = {"SimpleRayCaster": CASTER_SIMPLE, "IterativeRayCaster": CASTER_ITERATIVE,
                 "IterativeRayCasterLidar": CASTER_ITERATIVE_LIDAR}
_EVAL_KINDS = {"VoxelTypeEvaluator": EVAL_VOXEL_TYPE, "NaiveEvaluator": EVAL_NAIVE,
               "FrontierEvaluator": EVAL_FRONTIER, "VoxelWeightEvaluator": EVAL_VOXEL_WEIGHT}

# every symbol include/a3d.h declares
EXPORTS = [
    "a3d_ctx_create", "a3d_ctx_destroy", "a3d_last_error", "a3d_ctx_stream", "a3d_last_kernel_ms", "a3d_last_stage_ms",
    "a3d_kernel_launches", "a3d_set_option", "a3d_last_cast_kernel", "a3d_map_config", "a3d_map_upload_blocks", "a3d_map_upload_blocks_dev", "a3d_map_upload_weights", "a3d_map_upload_tsdf_distances",
    "a3d_map_remove_blocks", "a3d_map_clear", "a3d_map_num_blocks", "a3d_map_voxel_size",
    "a3d_map_voxel_state", "a3d_map_voxel_state_literal", "a3d_map_voxel_center", "a3d_map_is_observed", "a3d_map_voxel_weight", "a3d_map_voxel_tsdf_distance",
    "a3d_map_distance", "a3d_map_traversable", "a3d_caster_create", "a3d_caster_destroy", "a3d_caster_get_info",
    "a3d_caster_direction", "a3d_eval_create", "a3d_eval_destroy", "a3d_compute_gains",
    "a3d_compute_gains_dev", "a3d_visible_voxels", "a3d_gain_from_voxels", "a3d_filter_not_in",
    "a3d_bv_contains", "a3d_store_create", "a3d_store_destroy", "a3d_compute_gains_stored",
    "a3d_compute_gains_tree", "a3d_group_create", "a3d_group_destroy", "a3d_group_last_error", "a3d_group_size",
    "a3d_group_ctx", "a3d_group_map_config", "a3d_group_map_upload_blocks", "a3d_group_map_remove_blocks",
    "a3d_group_map_clear", "a3d_group_caster_create", "a3d_group_caster_destroy", "a3d_group_caster_get_info",
    "a3d_group_eval_create", "a3d_group_eval_destroy", "a3d_group_set_option", "a3d_group_compute_gains",
    "a3d_group_last_kernel_ms", "a3d_group_kernel_launches",
    "a3d_store_refold", "a3d_store_erase", "a3d_store_list_size", "a3d_store_get", "a3d_store_bytes",
]


class CasterParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("resolution_x", C.c_int32), ("resolution_y", C.c_int32),
                ("_pad", C.c_int32), ("ray_length", C.c_double), ("focal_length", C.c_double),
                ("ray_step", C.c_double), ("downsampling_factor", C.c_double),
                ("mount_t", C.c_double * 3), ("mount_q", C.c_double * 4),
                ("field_of_view_x", C.c_double), ("field_of_view_y", C.c_double)]


class EvalParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("clear_from_parents", C.c_int32),
                ("accurate_frontiers", C.c_int32), ("surface_frontiers", C.c_int32),
                ("checking_distance", C.c_double), ("gains", C.c_double * 6),
                ("bv_is_setup", C.c_int32), ("_pad", C.c_int32), ("bv_min", C.c_double * 3),
                ("bv_max", C.c_double * 3), ("bv_quat", C.c_double * 4),
                ("frontier_voxel_weight", C.c_double), ("min_impact_factor", C.c_double),
                ("new_voxel_weight", C.c_double), ("ray_angle_x", C.c_double),
                ("ray_angle_y", C.c_double)]


class CasterInfo(C.Structure):
    _fields_ = [("c_res_x", C.c_int32), ("c_res_y", C.c_int32), ("c_n_sections", C.c_int32),
                ("samples_per_full_ray", C.c_int32), ("fov_x", C.c_double), ("fov_y", C.c_double),
                ("ray_step", C.c_double), ("split_distances", C.c_double * 33),
                ("split_widths", C.c_int32 * 33), ("_pad", C.c_int32)]


def caster_params(kind, *, ray_length=5.0, focal_length=320.0, resolution_x=640, resolution_y=480,
                  ray_step=0.0, downsampling_factor=1.0, mounting_translation=(0.0, 0.0, 0.0),
                  mounting_rotation=(1.0, 0.0, 0.0, 0.0), field_of_view_x=360.0, field_of_view_y=30.0,
                  cls=CasterParams):
    """Reference parameter names/defaults (camera_model.cpp:170-183, sensor_model.cpp:18-27).
    ``mounting_rotation`` is (w,x,y,z) and is normalised here like sensor_model.cpp:27."""
    if isinstance(kind, str):
        kind = _CASTER_KINDS[kind]
    p = cls()
    p.kind = kind
    p.field_of_view_x, p.field_of_view_y = float(field_of_view_x), float(field_of_view_y)
    p.resolution_x, p.resolution_y = int(resolution_x), int(resolution_y)
    p.ray_length, p.focal_length = float(ray_length), float(focal_length)
    p.ray_step, p.downsampling_factor = float(ray_step), float(downsampling_factor)
    p.mount_t[:] = [float(x) for x in mounting_translation]
    w, x, y, z = [float(v) for v in mounting_rotation]
    # Eigen's Quaterniond::normalized(): coeffs (x,y,z,w) squared and summed two lanes at a time
    n = math.sqrt((x * x + z * z) + (y * y + w * w))
    p.mount_q[:] = [w / n, x / n, y / n, z / n]
    return p


def eval_params(kind, *, clear_from_parents=False, accurate_frontiers=False, surface_frontiers=True,
                checking_distance=1.0, gain_unknown=1.0, gain_occupied=0.0, gain_free=0.0,
                gain_unknown_outer=0.0, gain_occupied_outer=0.0, gain_free_outer=0.0,
                bounding_volume=None, frontier_voxel_weight=1.0, min_impact_factor=0.0,
                new_voxel_weight=0.01, ray_angle_x=0.0025, ray_angle_y=0.0025, cls=EvalParams):
    """Reference parameter names/defaults (voxel_type_evaluator.cpp:24-31,
    frontier_evaluator.cpp:17-20, simulated_sensor_evaluator.cpp:17).  ``bounding_volume`` is a dict
    with x_min..z_max and optional rotation [deg] (bounding_volume.cpp:12-31) or None."""
    if isinstance(kind, str):
        kind = _EVAL_KINDS[kind]
    p = cls()
    p.kind = kind
    p.frontier_voxel_weight, p.min_impact_factor = float(frontier_voxel_weight), float(min_impact_factor)
    p.new_voxel_weight = float(new_voxel_weight)
    p.ray_angle_x, p.ray_angle_y = float(ray_angle_x), float(ray_angle_y)
    p.clear_from_parents = int(bool(clear_from_parents))
    p.accurate_frontiers = int(bool(accurate_frontiers))
    p.surface_frontiers = int(bool(surface_frontiers))
    p.checking_distance = float(checking_distance)
    p.gains[:] = [gain_unknown, gain_occupied, gain_free, gain_unknown_outer, gain_occupied_outer,
                  gain_free_outer]
    bv = bounding_volume or {}
    lo = [float(bv.get(k, 0.0)) for k in ("x_min", "y_min", "z_min")]
    hi = [float(bv.get(k, 0.0)) for k in ("x_max", "y_max", "z_max")]
    p.bv_min[:] = lo
    p.bv_max[:] = hi
    p.bv_is_setup = int(hi[0] > lo[0] and hi[1] > lo[1] and hi[2] > lo[2])
    ang = float(bv.get("rotation", 0.0)) * math.pi / -180.0  # bounding_volume.cpp:21
    p.bv_quat[:] = [math.cos(ang * 0.5), 0.0, 0.0, math.sin(ang * 0.5)]
    return p


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen liba3d_b200.so and declare prototypes.  Raises OSError if it has not been built
    (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise OSError(f"{path} not built: run __graft_entry__.build() (nvcc, sm_100a)")
    lib = C.CDLL(path)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER
    lib.a3d_ctx_create.argtypes = [C.c_int, P(vp)]
    lib.a3d_ctx_destroy.argtypes = [vp]
    lib.a3d_ctx_destroy.restype = None
    lib.a3d_last_error.argtypes = [vp]
    lib.a3d_last_error.restype = C.c_char_p
    lib.a3d_ctx_stream.argtypes = [vp]
    lib.a3d_ctx_stream.restype = vp
    lib.a3d_last_kernel_ms.argtypes = [vp]
    lib.a3d_last_kernel_ms.restype = C.c_float
    lib.a3d_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    lib.a3d_last_stage_ms.restype = C.c_int
    lib.a3d_kernel_launches.argtypes = [vp]
    lib.a3d_kernel_launches.restype = u64
    lib.a3d_map_config.argtypes = [vp, C.c_float, C.c_int]
    lib.a3d_map_upload_blocks.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.a3d_map_upload_blocks_dev.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.a3d_map_upload_weights.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.a3d_map_upload_tsdf_distances.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.a3d_map_remove_blocks.argtypes = [vp, C.c_int, vp]
    lib.a3d_map_clear.argtypes = [vp]
    lib.a3d_map_num_blocks.argtypes = [vp]
    lib.a3d_map_num_blocks.restype = i64
    lib.a3d_map_voxel_size.argtypes = [vp]
    lib.a3d_map_voxel_size.restype = dbl
    for name in ("a3d_map_voxel_state", "a3d_map_voxel_state_literal", "a3d_map_voxel_center", "a3d_map_is_observed",
                 "a3d_map_voxel_weight", "a3d_map_voxel_tsdf_distance"):
        getattr(lib, name).argtypes = [vp, i64, vp, vp]
    lib.a3d_map_distance.argtypes = [vp, i64, vp, vp, vp]
    lib.a3d_map_traversable.argtypes = [vp, i64, vp, dbl, vp]
    lib.a3d_caster_create.argtypes = [vp, P(CasterParams), P(vp)]
    lib.a3d_caster_destroy.argtypes = [vp]
    lib.a3d_caster_destroy.restype = None
    lib.a3d_caster_get_info.argtypes = [vp, P(CasterInfo)]
    lib.a3d_caster_direction.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.a3d_eval_create.argtypes = [vp, P(EvalParams), P(vp)]
    lib.a3d_eval_destroy.argtypes = [vp]
    lib.a3d_eval_destroy.restype = None
    lib.a3d_compute_gains.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.a3d_compute_gains_dev.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp,
                                          vp]
    lib.a3d_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.a3d_last_cast_kernel.argtypes = [vp]
    lib.a3d_visible_voxels.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, i64, P(i64), P(i64), P(dbl), vp]
    lib.a3d_gain_from_voxels.argtypes = [vp, vp, i64, vp, P(dbl), P(i64), vp, vp]
    lib.a3d_bv_contains.argtypes = [vp, vp, i64, vp, vp]
    lib.a3d_filter_not_in.argtypes = [vp, i64, vp, i64, vp, vp]
    lib.a3d_store_create.argtypes = [vp, P(vp)]
    lib.a3d_store_destroy.argtypes = [vp]
    lib.a3d_store_destroy.restype = None
    lib.a3d_compute_gains_stored.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp,
                                             vp, vp, vp]
    lib.a3d_group_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    lib.a3d_group_destroy.argtypes = [vp]
    lib.a3d_group_destroy.restype = None
    lib.a3d_group_last_error.argtypes = [vp]
    lib.a3d_group_last_error.restype = C.c_char_p
    lib.a3d_group_size.argtypes = [vp]
    lib.a3d_group_ctx.argtypes = [vp, C.c_int]
    lib.a3d_group_ctx.restype = vp
    lib.a3d_group_map_config.argtypes = [vp, C.c_float, C.c_int]
    lib.a3d_group_map_upload_blocks.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.a3d_group_map_remove_blocks.argtypes = [vp, C.c_int, vp]
    lib.a3d_group_map_clear.argtypes = [vp]
    lib.a3d_group_caster_create.argtypes = [vp, C.POINTER(CasterParams), C.POINTER(vp)]
    lib.a3d_group_caster_destroy.argtypes = [vp]
    lib.a3d_group_caster_destroy.restype = None
    lib.a3d_group_caster_get_info.argtypes = [vp, C.POINTER(CasterInfo)]
    lib.a3d_group_eval_create.argtypes = [vp, C.POINTER(EvalParams), C.POINTER(vp)]
    lib.a3d_group_eval_destroy.argtypes = [vp]
    lib.a3d_group_eval_destroy.restype = None
    lib.a3d_group_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.a3d_group_compute_gains.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.a3d_group_last_kernel_ms.argtypes = [vp]
    lib.a3d_group_last_kernel_ms.restype = C.c_float
    lib.a3d_group_kernel_launches.argtypes = [vp]
    lib.a3d_group_kernel_launches.restype = C.c_uint64
    lib.a3d_compute_gains_tree.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.a3d_store_refold.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp]
    lib.a3d_store_erase.argtypes = [vp, C.c_int, vp]
    lib.a3d_store_list_size.argtypes = [vp, u64]
    lib.a3d_store_list_size.restype = i64
    lib.a3d_store_get.argtypes = [vp, vp, u64, vp, i64, P(i64)]
    lib.a3d_store_bytes.argtypes = [vp]
    lib.a3d_store_bytes.restype = i64
    _lib = lib
    return lib


class A3DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"a3d error {code}: {msg}")
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pts(p):
    p = np.ascontiguousarray(p, np.float64)
    if p.ndim != 2 or p.shape[1] != 3:
        raise ValueError("points must be (n,3)")
    return p


class Context:
    """One GPU context + its map mirror (the ``VoxbloxMap`` role)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.a3d_ctx_create(device, C.byref(h))
        if rc:
            raise A3DError(rc, self.lib.a3d_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.a3d_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise A3DError(rc, self.lib.a3d_last_error(self.h).decode())

    # ---- map ----
    def map_config(self, voxel_size: float, voxels_per_side: int = 16):
        self._ck(self.lib.a3d_map_config(self.h, C.c_float(voxel_size), voxels_per_side))
        self.vps = voxels_per_side

    def upload_blocks(self, block_idx, dist, observed, weight=None):
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        dist = np.ascontiguousarray(dist, np.float32)
        observed = np.ascontiguousarray(observed, np.uint8)
        n = block_idx.shape[0]
        nv = self.vps ** 3
        if dist.size != n * nv or observed.size != n * nv:
            raise ValueError("payload shape mismatch")
        if weight is not None:
            weight = np.ascontiguousarray(weight, np.float32)
        self._ck(self.lib.a3d_map_upload_blocks(self.h, n, _ptr(block_idx), _ptr(dist),
                                                _ptr(observed), _ptr(weight)))

    def upload_blocks_dev(self, block_idx, dist_ptr: int, observed_ptr: int, weight_ptr: int = 0):
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        self._ck(self.lib.a3d_map_upload_blocks_dev(self.h, block_idx.shape[0], _ptr(block_idx),
                                                    dist_ptr, observed_ptr, weight_ptr or None))

    def upload_weights(self, block_idx, weight):
        """TSDF weights alone, for blocks the mirror holds; returns the mask of blocks written."""
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        weight = np.ascontiguousarray(weight, np.float32)
        applied = np.zeros(block_idx.shape[0], np.uint8)
        self._ck(self.lib.a3d_map_upload_weights(self.h, block_idx.shape[0], _ptr(block_idx), _ptr(weight),
                                                 _ptr(applied)))
        return applied.astype(bool)

    def upload_tsdf_distances(self, block_idx, tsdf_distance):
        """TSDF distances alone (VoxbloxMap::getVoxelDistance), for blocks the mirror holds; returns the
        mask of blocks written."""
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        tsdf_distance = np.ascontiguousarray(tsdf_distance, np.float32)
        applied = np.zeros(block_idx.shape[0], np.uint8)
        self._ck(self.lib.a3d_map_upload_tsdf_distances(self.h, block_idx.shape[0], _ptr(block_idx),
                                                        _ptr(tsdf_distance), _ptr(applied)))
        return applied.astype(bool)

    def upload_map(self, m):
        """Upload a synth.SynthMap."""
        self.map_config(m.voxel_size, m.vps)
        self.upload_blocks(m.block_idx, m.dist, m.observed, m.weight)

    def remove_blocks(self, block_idx):
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        self._ck(self.lib.a3d_map_remove_blocks(self.h, block_idx.shape[0], _ptr(block_idx)))

    def clear(self):
        self._ck(self.lib.a3d_map_clear(self.h))

    @property
    def num_blocks(self):
        return int(self.lib.a3d_map_num_blocks(self.h))

    @property
    def voxel_size(self):
        return float(self.lib.a3d_map_voxel_size(self.h))

    def voxel_state(self, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.uint8)
        self._ck(self.lib.a3d_map_voxel_state(self.h, len(pts), _ptr(pts), _ptr(out)))
        return out

    def voxel_state_literal(self, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.uint8)
        self._ck(self.lib.a3d_map_voxel_state_literal(self.h, len(pts), _ptr(pts), _ptr(out)))
        return out

    def voxel_center(self, pts):
        pts = _pts(pts)
        out = np.empty((len(pts), 3), np.float64)
        self._ck(self.lib.a3d_map_voxel_center(self.h, len(pts), _ptr(pts), _ptr(out)))
        return out

    def is_observed(self, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.uint8)
        self._ck(self.lib.a3d_map_is_observed(self.h, len(pts), _ptr(pts), _ptr(out)))
        return out

    def voxel_weight(self, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.float64)
        self._ck(self.lib.a3d_map_voxel_weight(self.h, len(pts), _ptr(pts), _ptr(out)))
        return out

    def voxel_tsdf_distance(self, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.float64)
        self._ck(self.lib.a3d_map_voxel_tsdf_distance(self.h, len(pts), _ptr(pts), _ptr(out)))
        return out

    def distance(self, pts):
        pts = _pts(pts)
        d = np.empty(len(pts), np.float64)
        ok = np.empty(len(pts), np.uint8)
        self._ck(self.lib.a3d_map_distance(self.h, len(pts), _ptr(pts), _ptr(d), _ptr(ok)))
        return d, ok

    def traversable(self, pts, collision_radius):
        pts = _pts(pts)
        out = np.empty(len(pts), np.uint8)
        self._ck(self.lib.a3d_map_traversable(self.h, len(pts), _ptr(pts), float(collision_radius),
                                              _ptr(out)))
        return out

    # ---- modules ----
    def caster(self, kind, **kw):
        return Caster(self, caster_params(kind, **kw))

    def evaluator(self, kind, **kw):
        return Evaluator(self, eval_params(kind, **kw))

    @property
    def stream(self) -> int:
        return int(self.lib.a3d_ctx_stream(self.h) or 0)

    @property
    def last_kernel_ms(self) -> float:
        return float(self.lib.a3d_last_kernel_ms(self.h))

    @property
    def last_stage_ms(self):
        """Device time of each kernel of the kernel chain in the last compute_gains call ([] if another
        kernel ran): views, mark, deferred marking parts, leaf parts, fix-up (iterative caster)."""
        buf = (C.c_float * 8)()
        n = int(self.lib.a3d_last_stage_ms(self.h, buf, 8))
        return [float(buf[k]) for k in range(max(n, 0))]

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.a3d_kernel_launches(self.h))

    # ---- hot call ----
    def set_option(self, name: str, value: int):
        self._ck(self.lib.a3d_set_option(self.h, name.encode(), int(value)))

    @property
    def last_cast_kernel(self) -> int:
        return int(self.lib.a3d_last_cast_kernel(self.h))

    def compute_gains(self, caster, evaluator, pos, quat, view_segment=None, n_segments=None,
                      digests=True, set_digests=True, seg_origin=None, parent=None):
        """Batched SimulatedSensorEvaluator::computeGain.  Returns dict(gain, n_visible, n_samples,
        digest, set_digest) of per-segment arrays.  Default: one view per segment.  ``digests=True``
        (ordered digest) selects the scan-order kernel; with ``digests=False`` the wavefront kernel
        runs.  ``parent`` (index of each segment's parent in the batch, -1 = none) routes through
        a3d_compute_gains_tree: the batch form of clear_from_parents."""
        pos = _pts(pos)
        quat = np.ascontiguousarray(quat, np.float64)
        n_views = len(pos)
        if view_segment is None:
            view_segment = np.arange(n_views, dtype=np.int32)
        view_segment = np.ascontiguousarray(view_segment, np.int32)
        if n_segments is None:
            n_segments = int(view_segment.max()) + 1 if n_views else 0
        gains = np.zeros(n_segments, np.float64)
        nvis = np.zeros(n_segments, np.uint64)
        nsamp = np.zeros(n_segments, np.uint64)
        dig = np.zeros(n_segments, np.uint64) if digests else None
        sdig = np.zeros(n_segments, np.uint64) if set_digests else None
        if parent is not None:
            parent = np.ascontiguousarray(parent, np.int32)
            assert len(parent) == n_segments
            self._ck(self.lib.a3d_compute_gains_tree(
                self.h, caster.h, evaluator.h, n_views, _ptr(pos), _ptr(quat), _ptr(view_segment), n_segments,
                _ptr(parent), _ptr(gains), _ptr(nvis), _ptr(nsamp), _ptr(dig), _ptr(sdig),
                _ptr(None if seg_origin is None else np.ascontiguousarray(seg_origin, np.float64))))
            return dict(gain=gains, n_visible=nvis, n_samples=nsamp, digest=dig, set_digest=sdig)
        self._ck(self.lib.a3d_compute_gains(self.h, caster.h, evaluator.h, n_views, _ptr(pos),
                                            _ptr(quat), _ptr(view_segment), n_segments, _ptr(gains),
                                            _ptr(nvis), _ptr(nsamp), _ptr(dig), _ptr(sdig),
                                            _ptr(None if seg_origin is None else
                                                 np.ascontiguousarray(seg_origin, np.float64))))
        return dict(gain=gains, n_visible=nvis, n_samples=nsamp, digest=dig, set_digest=sdig)

    def compute_gains_dev(self, caster, evaluator, n_views, pos_ptr, quat_ptr, first_ptr, n_segments,
                          gains_ptr, nvis_ptr=0, nsamp_ptr=0, dig_ptr=0, set_dig_ptr=0):
        """Device-pointer variant (enqueue only, no sync)."""
        self._ck(self.lib.a3d_compute_gains_dev(self.h, caster.h, evaluator.h, n_views, pos_ptr,
                                                quat_ptr, first_ptr, n_segments, gains_ptr,
                                                nvis_ptr or None, nsamp_ptr or None,
                                                dig_ptr or None, set_dig_ptr or None, None))

    def visible_voxels(self, caster, pos, quat, evaluator=None, cap=None, origin=None):
        """One segment: ordered voxel-centre list (+gain if an evaluator is given)."""
        pos = _pts(pos)
        quat = np.ascontiguousarray(quat, np.float64)
        info = caster.info
        if cap is None:
            cap = max(1, len(pos)) * info.c_res_x * info.c_res_y * max(1, info.samples_per_full_ray)
        out = np.empty((cap, 3), np.float64)
        n_out, n_samp, gain = C.c_int64(), C.c_int64(), C.c_double()
        self._ck(self.lib.a3d_visible_voxels(self.h, caster.h, evaluator.h if evaluator else None,
                                             len(pos), _ptr(pos), _ptr(quat), _ptr(out), cap,
                                             C.byref(n_out), C.byref(n_samp), C.byref(gain),
                                             _ptr(None if origin is None else
                                                  np.ascontiguousarray(origin, np.float64))))
        return out[:n_out.value].copy(), int(n_samp.value), float(gain.value)

    def gain_from_voxels(self, evaluator, centres, want_keep=False, origin=None):
        centres = _pts(centres) if len(centres) else np.zeros((0, 3))
        gain, left = C.c_double(), C.c_int64()
        keep = np.zeros(len(centres), np.uint8) if want_keep else None
        self._ck(self.lib.a3d_gain_from_voxels(self.h, evaluator.h, len(centres), _ptr(centres),
                                               C.byref(gain), C.byref(left), _ptr(keep),
                                               _ptr(None if origin is None else
                                                    np.ascontiguousarray(origin, np.float64))))
        return float(gain.value), int(left.value), keep

    def filter_not_in(self, centres, other):
        """keep mask of ``centres`` entries that do not occur in ``other`` (clear_from_parents)."""
        centres = _pts(centres) if len(centres) else np.zeros((0, 3))
        other = _pts(other) if len(other) else np.zeros((0, 3))
        keep = np.ones(len(centres), np.uint8)
        self._ck(self.lib.a3d_filter_not_in(self.h, len(centres), _ptr(centres), len(other),
                                            _ptr(other), _ptr(keep)))
        return keep

    def bv_contains(self, evaluator, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.uint8)
        self._ck(self.lib.a3d_bv_contains(self.h, evaluator.h, len(pts), _ptr(pts), _ptr(out)))
        return out


class Store:
    """Device-resident stored voxel lists (the SimulatedSensorInfo lists kept in HBM)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._ck(ctx.lib.a3d_store_create(ctx.h, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.a3d_store_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute_gains(self, caster, evaluator, keys, pos, quat, view_segment=None, seg_origin=None,
                      parent_keys=None, digests=False):
        """a3d_compute_gains_stored.  ``parent_keys``: store key of each segment's parent (0 = none),
        used when the evaluator has clear_from_parents."""
        pos = _pts(pos)
        quat = np.ascontiguousarray(quat, np.float64)
        keys = np.ascontiguousarray(keys, np.uint64)
        n_seg = len(keys)
        if view_segment is None:
            view_segment = np.arange(len(pos), dtype=np.int32)
        view_segment = np.ascontiguousarray(view_segment, np.int32)
        gains = np.zeros(n_seg, np.float64)
        nvis = np.zeros(n_seg, np.uint64)
        nsamp = np.zeros(n_seg, np.uint64)
        dig = np.zeros(n_seg, np.uint64) if digests else None
        sdig = np.zeros(n_seg, np.uint64) if digests else None
        org = None if seg_origin is None else np.ascontiguousarray(seg_origin, np.float64)
        pk = None if parent_keys is None else np.ascontiguousarray(parent_keys, np.uint64)
        self.ctx._ck(self.ctx.lib.a3d_compute_gains_stored(
            self.ctx.h, caster.h, evaluator.h, self.h, len(pos), _ptr(pos), _ptr(quat),
            _ptr(view_segment), n_seg, _ptr(keys), _ptr(pk), _ptr(gains), _ptr(nvis), _ptr(nsamp), _ptr(dig),
            _ptr(sdig), _ptr(org)))
        return dict(gain=gains, n_visible=nvis, n_samples=nsamp, digest=dig, set_digest=sdig)

    def refold(self, evaluator, keys, seg_origin=None):
        keys = np.ascontiguousarray(keys, np.uint64)
        gains = np.zeros(len(keys), np.float64)
        left = np.zeros(len(keys), np.uint64)
        org = None if seg_origin is None else np.ascontiguousarray(seg_origin, np.float64)
        self.ctx._ck(self.ctx.lib.a3d_store_refold(self.ctx.h, evaluator.h, self.h, len(keys),
                                                   _ptr(keys), _ptr(gains), _ptr(left), _ptr(org)))
        return gains, left

    def erase(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        self.ctx._ck(self.ctx.lib.a3d_store_erase(self.h, len(keys), _ptr(keys)))

    def list_size(self, key) -> int:
        return int(self.ctx.lib.a3d_store_list_size(self.h, int(key)))

    def get(self, key):
        n = self.list_size(key)
        if n < 0:
            raise KeyError(key)
        out = np.empty((max(n, 1), 3), np.float64)
        n_out = C.c_int64()
        self.ctx._ck(self.ctx.lib.a3d_store_get(self.ctx.h, self.h, int(key), _ptr(out), max(n, 1),
                                                C.byref(n_out)))
        return out[:n].copy()

    @property
    def nbytes(self) -> int:
        return int(self.ctx.lib.a3d_store_bytes(self.h))


class Caster:
    def __init__(self, ctx: Context, params: CasterParams):
        self.ctx, self.params = ctx, params
        h = C.c_void_p()
        ctx._ck(ctx.lib.a3d_caster_create(ctx.h, C.byref(params), C.byref(h)))
        self.h = h
        self.info = CasterInfo()
        ctx._ck(ctx.lib.a3d_caster_get_info(h, C.byref(self.info)))

    def direction(self, i, j):
        out = np.empty(3, np.float64)
        rc = self.ctx.lib.a3d_caster_direction(self.h, i, j, _ptr(out))
        if rc:
            raise A3DError(rc, "ray index out of range")
        return out

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.a3d_caster_destroy(self.h)
        except Exception:
            pass


class Evaluator:
    def __init__(self, ctx: Context, params: EvalParams):
        self.ctx, self.params = ctx, params
        h = C.c_void_p()
        ctx._ck(ctx.lib.a3d_eval_create(ctx.h, C.byref(params), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.a3d_eval_destroy(self.h)
        except Exception:
            pass


class Group:
    """N GPUs behind one handle, one host thread, one process (a3d_group_*): map replicated with an NCCL
    broadcast of every update, segments dealt to the GPUs, gains all-gathered."""

    def __init__(self, devices):
        self.lib = load_library()
        ids = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.a3d_group_create(ids, len(devices), C.byref(h))
        if rc:
            raise A3DError(rc, (self.lib.a3d_group_last_error(None) or b"").decode())
        self.h = h
        self.devices = list(devices)
        self._keep = []

    def _ck(self, rc):
        if rc:
            raise A3DError(rc, (self.lib.a3d_group_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            for kind, h in self._keep:
                (self.lib.a3d_group_caster_destroy if kind == "c" else self.lib.a3d_group_eval_destroy)(h)
            self._keep = []
            self.lib.a3d_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_map(self, m):
        self._ck(self.lib.a3d_group_map_config(self.h, C.c_float(m.voxel_size), m.vps))
        self.upload_blocks(m.block_idx, m.dist, m.observed, m.weight)

    def upload_blocks(self, block_idx, dist, observed, weight=None):
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        dist = np.ascontiguousarray(dist, np.float32)
        observed = np.ascontiguousarray(observed, np.uint8)
        if weight is not None:
            weight = np.ascontiguousarray(weight, np.float32)
        self._ck(self.lib.a3d_group_map_upload_blocks(self.h, block_idx.shape[0], _ptr(block_idx), _ptr(dist),
                                                      _ptr(observed), _ptr(weight)))

    def remove_blocks(self, block_idx):
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        self._ck(self.lib.a3d_group_map_remove_blocks(self.h, block_idx.shape[0], _ptr(block_idx)))

    def set_option(self, name, value):
        self._ck(self.lib.a3d_group_set_option(self.h, name.encode(), int(value)))

    def caster(self, kind, **kw):
        p = caster_params(kind, **kw)
        h = C.c_void_p()
        self._ck(self.lib.a3d_group_caster_create(self.h, C.byref(p), C.byref(h)))
        self._keep.append(("c", h))
        return h

    def evaluator(self, kind, **kw):
        p = eval_params(kind, **kw)
        h = C.c_void_p()
        self._ck(self.lib.a3d_group_eval_create(self.h, C.byref(p), C.byref(h)))
        self._keep.append(("e", h))
        return h

    @property
    def last_kernel_ms(self):
        return float(self.lib.a3d_group_last_kernel_ms(self.h))

    @property
    def kernel_launches(self):
        return int(self.lib.a3d_group_kernel_launches(self.h))

    def compute_gains(self, caster, evaluator, pos, quat, view_segment=None, n_segments=None, digests=True,
                      set_digests=True, seg_origin=None):
        pos = _pts(pos)
        quat = np.ascontiguousarray(quat, np.float64)
        n_views = len(pos)
        if view_segment is None:
            view_segment = np.arange(n_views, dtype=np.int32)
        view_segment = np.ascontiguousarray(view_segment, np.int32)
        if n_segments is None:
            n_segments = int(view_segment.max()) + 1 if n_views else 0
        gains = np.zeros(n_segments, np.float64)
        nvis = np.zeros(n_segments, np.uint64)
        nsamp = np.zeros(n_segments, np.uint64)
        dig = np.zeros(n_segments, np.uint64) if digests else None
        sdig = np.zeros(n_segments, np.uint64) if set_digests else None
        self._ck(self.lib.a3d_group_compute_gains(
            self.h, caster, evaluator, n_views, _ptr(pos), _ptr(quat), _ptr(view_segment), n_segments, _ptr(gains),
            _ptr(nvis), _ptr(nsamp), _ptr(dig), _ptr(sdig),
            _ptr(None if seg_origin is None else np.ascontiguousarray(seg_origin, np.float64))))
        return dict(gain=gains, n_visible=nvis, n_samples=nsamp, digest=dig, set_digest=sdig)
