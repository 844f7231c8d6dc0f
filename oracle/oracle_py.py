"""ctypes binding of the CPU oracle (oracle/liba3d_oracle.so).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
Parity status (pinned by oracle/_ref except for the voxblox arithmetic): see oracle/a3d_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liba3d_oracle.so")


class CasterParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("resolution_x", C.c_int32), ("resolution_y", C.c_int32),
                ("_pad", C.c_int32), ("ray_length", C.c_double), ("focal_length", C.c_double),
                ("ray_step", C.c_double), ("downsampling_factor", C.c_double),
                ("mount_t", C.c_double * 3), ("mount_q", C.c_double * 4),
                ("field_of_view_x", C.c_double), ("field_of_view_y", C.c_double),
                ("mount_q_raw", C.c_double * 4)]


class EvalParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("clear_from_parents", C.c_int32),
                ("accurate_frontiers", C.c_int32), ("surface_frontiers", C.c_int32),
                ("checking_distance", C.c_double), ("gains", C.c_double * 6),
                ("bv_is_setup", C.c_int32), ("_pad", C.c_int32), ("bv_min", C.c_double * 3),
                ("bv_max", C.c_double * 3), ("bv_quat", C.c_double * 4),
                ("frontier_voxel_weight", C.c_double), ("min_impact_factor", C.c_double),
                ("new_voxel_weight", C.c_double), ("ray_angle_x", C.c_double),
                ("ray_angle_y", C.c_double), ("bv_rotation_deg", C.c_double)]


_CASTER_KINDS = {"SimpleRayCaster": 0, "IterativeRayCaster": 1, "IterativeRayCasterLidar": 2}
_EVAL_KINDS = {"VoxelTypeEvaluator": 0, "NaiveEvaluator": 1, "FrontierEvaluator": 2,
               "VoxelWeightEvaluator": 3}


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++ -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "a3d_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


REF_LIB_PATH = os.path.join(_HERE, "_ref", "liba3d_ref.so")
REFERENCE_ROOT = "/root/reference"


def build_ref(force: bool = False) -> str | None:
    """Compile oracle/_ref (the reference's own sources behind the orc_* API) when /root/reference is
    present; returns the library path, or None when it can neither be built nor found prebuilt."""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.run(["make", "-C", _HERE, "-j8", "_ref"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return REF_LIB_PATH if os.path.exists(REF_LIB_PATH) else None


def have_ref() -> bool:
    return os.path.exists(REF_LIB_PATH)


_lib = None
_ref_lib = None


def ref_lib():
    """oracle/_ref/liba3d_ref.so — same entry points, reference code underneath."""
    global _ref_lib
    if _ref_lib is None:
        if not os.path.exists(REF_LIB_PATH):
            raise FileNotFoundError(REF_LIB_PATH + " (make -C oracle _ref needs /root/reference)")
        _ref_lib = _bind(C.CDLL(REF_LIB_PATH))
    return _ref_lib


def lib(impl: str = "oracle"):
    global _lib
    if impl == "ref":
        return ref_lib()
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    _lib = _bind(C.CDLL(LIB_PATH))
    return _lib


def _bind(L):
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    P = C.POINTER
    L.orc_map_create.argtypes = [C.c_float, C.c_int]
    L.orc_map_create.restype = vp
    L.orc_map_destroy.argtypes = [vp]
    L.orc_map_destroy.restype = None
    L.orc_map_upload_blocks.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.orc_map_remove_blocks.argtypes = [vp, C.c_int, vp]
    L.orc_map_num_blocks.argtypes = [vp]
    L.orc_map_num_blocks.restype = i64
    L.orc_map_voxel_size.argtypes = [vp]
    L.orc_map_voxel_size.restype = dbl
    for n in ("orc_voxel_state", "orc_voxel_center", "orc_is_observed", "orc_voxel_weight"):
        getattr(L, n).argtypes = [vp, i64, vp, vp]
        getattr(L, n).restype = None
    L.orc_distance.argtypes = [vp, i64, vp, vp, vp]
    L.orc_distance.restype = None
    L.orc_caster_create.argtypes = [vp, P(CasterParams)]
    L.orc_caster_create.restype = vp
    L.orc_caster_destroy.argtypes = [vp]
    L.orc_caster_destroy.restype = None
    L.orc_caster_info.argtypes = [vp, vp, vp, vp, vp, vp]
    L.orc_caster_info.restype = None
    L.orc_caster_direction.argtypes = [vp, C.c_int, C.c_int, vp]
    L.orc_caster_direction.restype = None
    L.orc_sample_viewpoints.argtypes = [dbl, C.c_int, vp, vp]
    L.orc_visible_voxels.argtypes = [vp, C.c_int, vp, vp, vp, i64, P(i64)]
    L.orc_visible_voxels.restype = i64
    L.orc_compute_gains.argtypes = [vp, P(EvalParams), C.c_int, vp, vp, vp, C.c_int, vp, vp, vp,
                                    vp, vp, vp, C.c_int, vp, vp]
    L.orc_digest_set.argtypes = [i64, vp]
    L.orc_digest_set.restype = C.c_uint64
    L.orc_gain_from_voxels.argtypes = [vp, P(EvalParams), i64, vp, P(i64), vp]
    L.orc_gain_from_voxels.restype = dbl
    L.orc_bv_contains.argtypes = [P(EvalParams), i64, vp, vp]
    L.orc_bv_contains.restype = None
    L.orc_digest.argtypes = [i64, vp]
    L.orc_digest.restype = C.c_uint64
    L.orc_traversable.argtypes = [vp, i64, vp, dbl, vp]
    L.orc_traversable.restype = None
    L.orc_impl.argtypes = []
    L.orc_impl.restype = C.c_char_p
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pts(p):
    p = np.ascontiguousarray(p, np.float64)
    assert p.ndim == 2 and p.shape[1] == 3
    return p


def caster_params(kind, *, ray_length=5.0, focal_length=320.0, resolution_x=640, resolution_y=480,
                  ray_step=0.0, downsampling_factor=1.0, mounting_translation=(0.0, 0.0, 0.0),
                  mounting_rotation=(1.0, 0.0, 0.0, 0.0), field_of_view_x=360.0, field_of_view_y=30.0):
    if isinstance(kind, str):
        kind = _CASTER_KINDS[kind]
    p = CasterParams()
    p.field_of_view_x, p.field_of_view_y = float(field_of_view_x), float(field_of_view_y)
    p.kind = kind
    p.resolution_x, p.resolution_y = int(resolution_x), int(resolution_y)
    p.ray_length, p.focal_length = float(ray_length), float(focal_length)
    p.ray_step, p.downsampling_factor = float(ray_step), float(downsampling_factor)
    p.mount_t[:] = [float(x) for x in mounting_translation]
    w, x, y, z = [float(v) for v in mounting_rotation]
    p.mount_q_raw[:] = [w, x, y, z]
    # Quaterniond::normalized(): coeffs (x,y,z,w) squared, summed two lanes at a time (SSE2)
    n = math.sqrt((x * x + z * z) + (y * y + w * w))
    p.mount_q[:] = [w / n, x / n, y / n, z / n]
    return p


def eval_params(kind, *, clear_from_parents=False, accurate_frontiers=False, surface_frontiers=True,
                checking_distance=1.0, gain_unknown=1.0, gain_occupied=0.0, gain_free=0.0,
                gain_unknown_outer=0.0, gain_occupied_outer=0.0, gain_free_outer=0.0,
                bounding_volume=None, frontier_voxel_weight=1.0, min_impact_factor=0.0,
                new_voxel_weight=0.01, ray_angle_x=0.0025, ray_angle_y=0.0025):
    if isinstance(kind, str):
        kind = _EVAL_KINDS[kind]
    p = EvalParams()
    p.frontier_voxel_weight, p.min_impact_factor = float(frontier_voxel_weight), float(min_impact_factor)
    p.new_voxel_weight = float(new_voxel_weight)
    p.ray_angle_x, p.ray_angle_y = float(ray_angle_x), float(ray_angle_y)
    p.kind = kind
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
    p.bv_rotation_deg = float(bv.get("rotation", 0.0))
    ang = float(bv.get("rotation", 0.0)) * math.pi / -180.0
    p.bv_quat[:] = [math.cos(ang * 0.5), 0.0, 0.0, math.sin(ang * 0.5)]
    return p


class OracleMap:
    def __init__(self, voxel_size: float, vps: int = 16, impl: str = "oracle"):
        self.L = lib(impl)
        self.impl = impl
        self.h = self.L.orc_map_create(C.c_float(voxel_size), vps)
        self.vps = vps

    @classmethod
    def from_synth(cls, m, impl: str = "oracle"):
        o = cls(m.voxel_size, m.vps, impl)
        o.upload_blocks(m.block_idx, m.dist, m.observed, m.weight)
        return o

    def __del__(self):
        try:
            if self.h:
                self.L.orc_map_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def upload_blocks(self, block_idx, dist, observed, weight=None):
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        dist = np.ascontiguousarray(dist, np.float32)
        observed = np.ascontiguousarray(observed, np.uint8)
        if weight is not None:
            weight = np.ascontiguousarray(weight, np.float32)
        self.L.orc_map_upload_blocks(self.h, block_idx.shape[0], _ptr(block_idx), _ptr(dist),
                                     _ptr(observed), _ptr(weight))

    def remove_blocks(self, block_idx):
        block_idx = np.ascontiguousarray(block_idx, np.int32)
        self.L.orc_map_remove_blocks(self.h, block_idx.shape[0], _ptr(block_idx))

    @property
    def num_blocks(self):
        return int(self.L.orc_map_num_blocks(self.h))

    @property
    def voxel_size(self):
        return float(self.L.orc_map_voxel_size(self.h))

    def voxel_state(self, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.uint8)
        self.L.orc_voxel_state(self.h, len(pts), _ptr(pts), _ptr(out))
        return out

    def voxel_center(self, pts):
        pts = _pts(pts)
        out = np.empty((len(pts), 3), np.float64)
        self.L.orc_voxel_center(self.h, len(pts), _ptr(pts), _ptr(out))
        return out

    def is_observed(self, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.uint8)
        self.L.orc_is_observed(self.h, len(pts), _ptr(pts), _ptr(out))
        return out

    def voxel_weight(self, pts):
        pts = _pts(pts)
        out = np.empty(len(pts), np.float64)
        self.L.orc_voxel_weight(self.h, len(pts), _ptr(pts), _ptr(out))
        return out

    def distance(self, pts):
        pts = _pts(pts)
        d = np.empty(len(pts), np.float64)
        ok = np.empty(len(pts), np.uint8)
        self.L.orc_distance(self.h, len(pts), _ptr(pts), _ptr(d), _ptr(ok))
        return d, ok

    def caster(self, kind, **kw):
        return OracleCaster(self, caster_params(kind, **kw))

    def traversable(self, pts, collision_radius):
        pts = _pts(pts)
        out = np.empty(len(pts), np.uint8)
        self.L.orc_traversable(self.h, len(pts), _ptr(pts), float(collision_radius), _ptr(out))
        return out

    def gain_from_voxels(self, eparams, centres, origin=None):
        centres = _pts(centres) if len(centres) else np.zeros((0, 3))
        left = C.c_int64()
        if origin is not None:
            origin = np.ascontiguousarray(origin, np.float64)
        g = self.L.orc_gain_from_voxels(self.h, C.byref(eparams), len(centres), _ptr(centres),
                                        C.byref(left), _ptr(origin))
        return float(g), int(left.value)


class OracleCaster:
    def __init__(self, omap: OracleMap, params: CasterParams):
        self.map, self.params, self.L = omap, params, omap.L
        self.h = self.L.orc_caster_create(omap.h, C.byref(params))
        info = np.zeros(3, np.int32)
        sd = np.zeros(33, np.float64)
        sw = np.zeros(33, np.int32)
        fov = np.zeros(2, np.float64)
        step = np.zeros(1, np.float64)
        self.L.orc_caster_info(self.h, _ptr(info), _ptr(sd), _ptr(sw), _ptr(fov), _ptr(step))
        self.c_res_x, self.c_res_y, self.c_n_sections = [int(v) for v in info]
        self.split_distances = sd[: self.c_n_sections + 1].copy()
        self.split_widths = sw[: self.c_n_sections + 1].copy()
        self.fov_x, self.fov_y = float(fov[0]), float(fov[1])
        self.ray_step = float(step[0])

    def __del__(self):
        try:
            if self.h:
                self.L.orc_caster_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def direction(self, i, j):
        out = np.empty(3, np.float64)
        self.L.orc_caster_direction(self.h, i, j, _ptr(out))
        return out

    def visible_voxels(self, pos, quat, cap=None):
        pos = _pts(pos)
        quat = np.ascontiguousarray(quat, np.float64)
        if cap is None:
            cap = len(pos) * self.c_res_x * self.c_res_y * 400
        out = np.empty((cap, 3), np.float64)
        ns = C.c_int64()
        n = self.L.orc_visible_voxels(self.h, len(pos), _ptr(pos), _ptr(quat), _ptr(out), cap,
                                      C.byref(ns))
        assert n <= cap
        return out[:n].copy(), int(ns.value)

    def compute_gains(self, eparams, pos, quat, view_segment=None, n_segments=None, parent=None,
                      n_threads=1, seg_origin=None):
        pos = _pts(pos)
        quat = np.ascontiguousarray(quat, np.float64)
        n_views = len(pos)
        if view_segment is None:
            view_segment = np.arange(n_views, dtype=np.int32)
        view_segment = np.ascontiguousarray(view_segment, np.int32)
        if n_segments is None:
            n_segments = int(view_segment.max()) + 1 if n_views else 0
        if parent is not None:
            parent = np.ascontiguousarray(parent, np.int32)
        gains = np.zeros(n_segments, np.float64)
        nvis = np.zeros(n_segments, np.uint64)
        nsamp = np.zeros(n_segments, np.uint64)
        dig = np.zeros(n_segments, np.uint64)
        fold = np.zeros(n_segments, np.uint64)
        sdig = np.zeros(n_segments, np.uint64)
        rc = self.L.orc_compute_gains(self.h, C.byref(eparams), n_views, _ptr(pos), _ptr(quat),
                                      _ptr(view_segment), n_segments, _ptr(parent), _ptr(gains),
                                      _ptr(nvis), _ptr(nsamp), _ptr(dig), _ptr(fold), n_threads,
                                      _ptr(sdig),
                                      _ptr(None if seg_origin is None
                                           else np.ascontiguousarray(seg_origin, np.float64)))
        if rc:
            raise ValueError("orc_compute_gains: bad view_segment")
        return dict(gain=gains, n_visible=nvis, n_samples=nsamp, digest=dig, set_digest=sdig,
                    n_fold_lookups=fold)


def sample_viewpoints(sampling_time, times_ns, impl="oracle"):
    t = np.ascontiguousarray(times_ns, np.int64)
    out = np.zeros(len(t) + 1, np.int32)
    n = lib(impl).orc_sample_viewpoints(float(sampling_time), len(t), _ptr(t), _ptr(out))
    return out[:n].copy()


def bv_contains(eparams, pts, impl="oracle"):
    pts = _pts(pts)
    out = np.empty(len(pts), np.uint8)
    lib(impl).orc_bv_contains(C.byref(eparams), len(pts), _ptr(pts), _ptr(out))
    return out


def digest_set(centres):
    centres = _pts(centres) if len(centres) else np.zeros((0, 3))
    return int(lib().orc_digest_set(len(centres), _ptr(centres)))


def digest(centres):
    centres = _pts(centres) if len(centres) else np.zeros((0, 3))
    return int(lib().orc_digest(len(centres), _ptr(centres)))
