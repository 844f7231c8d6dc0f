// a3d_device.cuh — device-side map view and the per-sample arithmetic of the view-gain path.
//
// Everything here is written for sm_100a and reproduces the reference's fp64-step / fp32-lookup
// arithmetic bit-for-bit (DESIGN.md "arithmetic contract"): every floating-point operation is an
// explicit round-to-nearest intrinsic (__dmul_rn, __fadd_rn, ...) so that nvcc can never contract
// a multiply-add into an FMA.
//
// Reference semantics implemented (paths relative to the reference checkout):
//   a3d::voxel_state   VoxbloxMap::getVoxelState   active_3d_planning_voxblox/src/map/voxblox.cpp:48-60
//                      (+ voxblox EsdfMap::getDistanceAtPosition, trilinear over 8 ESDF voxels)
//   a3d::voxel_center  VoxbloxMap::getVoxelCenter  voxblox.cpp:66-81
//   a3d::is_observed   VoxbloxMap::isObserved      voxblox.cpp:43-45
//   a3d::voxel_weight  VoxbloxMap::getVoxelWeight  voxblox.cpp:101-115
//   a3d::bv_contains   BoundingVolume::contains    active_3d_planning_core/src/data/bounding_volume.cpp:33-57
//   a3d::rotate        Eigen::Quaterniond * Vector3d as used at camera_model.cpp:27, simple_ray_caster.cpp:45
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace a3d {

// Unobserved voxels are stored as this exact bit pattern in the distance slab (observed flag
// folded into the 4-byte record).  Observed NaN distances are canonicalised to 0x7FC00000 on upload.
constexpr uint32_t kUnobservedBits = 0x7FD5A3D1u;
constexpr float kCoordEps = 1e-6f;

struct HashEntry {  // 16 B, one LDG.128 per probe
  int32_t x, y, z;
  int32_t slot;  // < 0: empty
};

struct MapView {
  const HashEntry* __restrict__ table;
  const float* __restrict__ dist;    // slab: [slot][vps^3], kUnobservedBits where !observed
  const float* __restrict__ weight;  // TSDF weights slab or nullptr
  const uint8_t* __restrict__ meta;  // per-voxel meta byte slab (see a3d_kernels.cu, k_build_meta)
  uint32_t mask;                     // table capacity - 1
  int32_t vps;
  int32_t nv;                        // vps^3
  float vs_f, vs_inv_f, bs_f, bs_inv_f;
  double c_vs;                       // double(vs_f), VoxbloxMap::c_voxel_size_
};

struct D3 {
  double x, y, z;
};
struct Q4 {
  double w, x, y, z;
};

__device__ __forceinline__ uint32_t hash3(int x, int y, int z) {
  uint32_t h = static_cast<uint32_t>(x) * 73856093u ^ static_cast<uint32_t>(y) * 19349663u ^
               static_cast<uint32_t>(z) * 83492791u;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

// Block index → slab slot, or -1.  Open addressing, linear probing, no tombstones (the host
// rebuilds the table on removal).
__device__ __forceinline__ int find_slot(const MapView& m, int bx, int by, int bz) {
  uint32_t h = hash3(bx, by, bz) & m.mask;
#pragma unroll 1
  for (;;) {
    const int4 e = __ldg(reinterpret_cast<const int4*>(m.table + h));
    if (e.w < 0) return -1;
    if (e.x == bx && e.y == by && e.z == bz) return e.w;
    h = (h + 1) & m.mask;
  }
}

// voxblox getGridIndexFromPoint, scalar: (int)floorf(p*inv + 1e-6f)
__device__ __forceinline__ int grid_index(float p, float inv) {
  return __float2int_rd(__fadd_rn(__fmul_rn(p, inv), kCoordEps));
}
// voxblox getCenterPointFromGridIndex: float((double(float(i)) + 0.5) * double(size))
__device__ __forceinline__ float center_point(int i, float size) {
  return __double2float_rn(
      __dmul_rn(__dadd_rn(static_cast<double>(__int2float_rn(i)), 0.5), static_cast<double>(size)));
}
__device__ __forceinline__ float origin_point(int i, float size) {
  return __fmul_rn(__int2float_rn(i), size);
}

// Eigen cross / quaternion-vector product, exact op order (see oracle/a3d_oracle.cpp rotate()).
__device__ __forceinline__ D3 cross(const D3& a, const D3& b) {
  D3 r;
  r.x = __dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y));
  r.y = __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z));
  r.z = __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x));
  return r;
}
__device__ __forceinline__ D3 rotate(const Q4& q, const D3& v) {
  const D3 qv{q.x, q.y, q.z};
  D3 uv = cross(qv, v);
  uv.x = __dadd_rn(uv.x, uv.x);
  uv.y = __dadd_rn(uv.y, uv.y);
  uv.z = __dadd_rn(uv.z, uv.z);
  const D3 c = cross(qv, uv);
  D3 r;
  r.x = __dadd_rn(__dadd_rn(v.x, __dmul_rn(q.w, uv.x)), c.x);
  r.y = __dadd_rn(__dadd_rn(v.y, __dmul_rn(q.w, uv.y)), c.y);
  r.z = __dadd_rn(__dadd_rn(v.z, __dmul_rn(q.w, uv.z)), c.z);
  return r;
}
__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
  Q4 r;
  r.w = __dsub_rn(__dsub_rn(__dsub_rn(__dmul_rn(a.w, b.w), __dmul_rn(a.x, b.x)), __dmul_rn(a.y, b.y)),
                  __dmul_rn(a.z, b.z));
  r.x = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.w, b.x), __dmul_rn(a.x, b.w)), __dmul_rn(a.y, b.z)),
                  __dmul_rn(a.z, b.y));
  r.y = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.w, b.y), __dmul_rn(a.y, b.w)), __dmul_rn(a.z, b.x)),
                  __dmul_rn(a.x, b.z));
  r.z = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.w, b.z), __dmul_rn(a.z, b.w)), __dmul_rn(a.x, b.y)),
                  __dmul_rn(a.y, b.x));
  return r;
}

// ---- interpolated ESDF distance (generic path) -------------------------------------------------
// Returns true and *dist when all 8 surrounding voxels exist and are observed.
__device__ inline bool interp_distance(const MapView& m, float px, float py, float pz,
                                             float* dist_out) {
  const float p[3] = {px, py, pz};
  int b[3], v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) b[k] = grid_index(p[k], m.bs_inv_f);
  int slot = find_slot(m, b[0], b[1], b[2]);
  if (slot < 0) return false;
  bool moved = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float origin = origin_point(b[k], m.bs_f);
    v[k] = grid_index(__fsub_rn(p[k], origin), m.vs_inv_f);
    const float centre = __fadd_rn(origin, center_point(v[k], m.vs_f));
    if (__fsub_rn(p[k], centre) < 0.0f) {
      v[k]--;
      if (v[k] < 0) {
        b[k]--;
        v[k] += m.vps;
        moved = true;
      }
    }
  }
  if (moved) {
    slot = find_slot(m, b[0], b[1], b[2]);
    if (slot < 0) return false;
  }
  // carry pattern of the +1 corners
  const int cx = (v[0] + 1 >= m.vps), cy = (v[1] + 1 >= m.vps), cz = (v[2] + 1 >= m.vps);
  const int c0x = (v[0] >= m.vps), c0y = (v[1] >= m.vps), c0z = (v[2] >= m.vps);  // fp artefact
  float d[8];
  float o[3];
  int cached_key = -1, cached_slot = -1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ox = (i >> 2) & 1, oy = (i >> 1) & 1, oz = i & 1;
    int vi[3] = {v[0] + ox, v[1] + oy, v[2] + oz};
    const int kx = ox ? cx : c0x, ky = oy ? cy : c0y, kz = oz ? cz : c0z;
    int s = slot;
    if (kx | ky | kz) {
      const int key = kx | (ky << 1) | (kz << 2);
      if (key != cached_key) {
        cached_slot = find_slot(m, b[0] + kx, b[1] + ky, b[2] + kz);
        cached_key = key;
      }
      s = cached_slot;
      if (s < 0) return false;
      vi[0] -= kx ? m.vps : 0;
      vi[1] -= ky ? m.vps : 0;
      vi[2] -= kz ? m.vps : 0;
    }
    if (i == 0) {
      const int nb[3] = {b[0] + kx, b[1] + ky, b[2] + kz};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float voxel_pos = __fadd_rn(origin_point(nb[k], m.bs_f), center_point(vi[k], m.vs_f));
        o[k] = __fmul_rn(__fsub_rn(p[k], voxel_pos), m.vs_inv_f);
      }
    }
    const int lin = vi[0] + m.vps * (vi[1] + m.vps * vi[2]);
    if (lin < 0 || lin >= m.nv) return false;
    const float dv = __ldg(m.dist + static_cast<size_t>(s) * m.nv + lin);
    if (__float_as_uint(dv) == kUnobservedBits) return false;
    d[i] = dv;
  }
  // q * (T * d) with the summation order fixed in SURVEY.md A.3 / oracle
  const float q4 = __fmul_rn(o[0], o[1]);
  const float q5 = __fmul_rn(o[1], o[2]);
  const float q6 = __fmul_rn(o[2], o[0]);
  const float q7 = __fmul_rn(q4, o[2]);
  const float w1 = __fsub_rn(d[4], d[0]);
  const float w2 = __fsub_rn(d[2], d[0]);
  const float w3 = __fsub_rn(d[1], d[0]);
  const float d01 = __fsub_rn(d[0], d[1]);
  const float w4 = __fadd_rn(__fsub_rn(d[0], d[2]), __fsub_rn(d[6], d[4]));
  const float w5 = __fadd_rn(d01, __fsub_rn(d[3], d[2]));
  const float w6 = __fadd_rn(d01, __fsub_rn(d[5], d[4]));
  const float w7 = __fadd_rn(__fadd_rn(w3, __fsub_rn(d[2], d[3])),
                             __fadd_rn(__fsub_rn(d[4], d[5]), __fsub_rn(d[7], d[6])));
  const float u0 = d[0];  // q0 == 1
  const float u1 = __fmul_rn(o[0], w1);
  const float u2 = __fmul_rn(o[1], w2);
  const float u3 = __fmul_rn(o[2], w3);
  const float u4 = __fmul_rn(q4, w4);
  const float u5 = __fmul_rn(q5, w5);
  const float u6 = __fmul_rn(q6, w6);
  const float u7 = __fmul_rn(q7, w7);
  *dist_out = __fadd_rn(__fadd_rn(__fadd_rn(u0, u4), __fadd_rn(u2, u6)),
                        __fadd_rn(__fadd_rn(u1, u5), __fadd_rn(u3, u7)));
  return true;
}

__device__ __forceinline__ int voxel_state_f(const MapView& m, float px, float py, float pz) {
  float dist;
  if (!interp_distance(m, px, py, pz, &dist)) return 2;         // UNKNOWN
  return (static_cast<double>(dist) < m.c_vs) ? 0 : 1;          // OCCUPIED : FREE
}
__device__ __forceinline__ int voxel_state(const MapView& m, const D3& p) {
  return voxel_state_f(m, __double2float_rn(p.x), __double2float_rn(p.y), __double2float_rn(p.z));
}

// getVoxelCenter; also returns the block index and (unclamped) voxel index it derived.
__device__ __forceinline__ D3 voxel_center(const MapView& m, const D3& p, int* b_out, int* v_out) {
  const double pp[3] = {p.x, p.y, p.z};
  double c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int b = grid_index(__double2float_rn(pp[k]), m.bs_inv_f);
    const double origin = static_cast<double>(origin_point(b, m.bs_f));
    const int v = grid_index(__double2float_rn(__dsub_rn(pp[k], origin)), m.vs_inv_f);
    c[k] = __dadd_rn(origin, static_cast<double>(center_point(v, m.vs_f)));
    b_out[k] = b;
    v_out[k] = v;
  }
  return D3{c[0], c[1], c[2]};
}

// Nearest-voxel linear address (block by coordinates, truncated voxel index); -1 if no block.
__device__ __forceinline__ int64_t nearest_voxel(const MapView& m, const D3& p) {
  const float pf[3] = {__double2float_rn(p.x), __double2float_rn(p.y), __double2float_rn(p.z)};
  int b[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) b[k] = grid_index(pf[k], m.bs_inv_f);
  const int slot = find_slot(m, b[0], b[1], b[2]);
  if (slot < 0) return -1;
  int v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[k] = grid_index(__fsub_rn(pf[k], origin_point(b[k], m.bs_f)), m.vs_inv_f);
    v[k] = max(min(v[k], m.vps - 1), 0);
  }
  return static_cast<int64_t>(slot) * m.nv + (v[0] + m.vps * (v[1] + m.vps * v[2]));
}
__device__ __forceinline__ bool is_observed(const MapView& m, const D3& p) {
  const int64_t a = nearest_voxel(m, p);
  if (a < 0) return false;
  return __float_as_uint(__ldg(m.dist + a)) != kUnobservedBits;
}
__device__ __forceinline__ double voxel_weight(const MapView& m, const D3& p) {
  if (!m.weight) return 0.0;
  const int64_t a = nearest_voxel(m, p);
  if (a < 0) return 0.0;
  return static_cast<double>(__ldg(m.weight + a));
}

struct EvalView {
  int32_t kind;
  int32_t accurate_frontiers, surface_frontiers;
  int32_t bv_is_setup;
  double nb_step;  // c_voxel_size_ * checking_distance (frontier_evaluator.cpp:31)
  double gains[6];
  double bv_min[3], bv_max[3];
  Q4 bv_q;
};

__device__ __forceinline__ bool bv_contains(const EvalView& e, const D3& p) {
  if (!e.bv_is_setup) return true;
  const D3 r = rotate(e.bv_q, p);
  if (r.x < e.bv_min[0]) return false;
  if (r.x > e.bv_max[0]) return false;
  if (r.y < e.bv_min[1]) return false;
  if (r.y > e.bv_max[1]) return false;
  if (r.z < e.bv_min[2]) return false;
  if (r.z > e.bv_max[2]) return false;
  return true;
}

// FrontierEvaluator neighbour order (frontier_evaluator.cpp:33-65), packed 2 bits per axis
// (0: 0, 1: +, 2: -), x | y<<2 | z<<4.
__device__ __forceinline__ void frontier_offset(bool accurate, int i, int* sx, int* sy, int* sz) {
  // 6-neighbourhood: +x,-x,+y,-y,+z,-z
  if (!accurate) {
    const int axis = i >> 1, sgn = (i & 1) ? -1 : 1;
    *sx = axis == 0 ? sgn : 0;
    *sy = axis == 1 ? sgn : 0;
    *sz = axis == 2 ? sgn : 0;
    return;
  }
  // 26-neighbourhood: x-major groups of 9 (x = +,0,-); within a group (y,z) runs
  // (0,0),(+,0),(-,0),(0,+),(+,+),(-,+),(0,-),(+,-),(-,-); the x==0 group omits (0,0).
  int g, r;
  if (i < 9) {
    g = 0;
    r = i;
  } else if (i < 17) {
    g = 1;
    r = i - 9 + 1;
  } else {
    g = 2;
    r = i - 17;
  }
  const int ys[3] = {0, 1, -1};
  *sx = g == 0 ? 1 : (g == 1 ? 0 : -1);
  *sy = ys[r % 3];
  *sz = ys[r / 3];
}

__device__ __forceinline__ bool is_frontier_voxel(const MapView& m, const EvalView& e, const D3& c) {
  const int n = e.accurate_frontiers ? 26 : 6;
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    int sx, sy, sz;
    frontier_offset(e.accurate_frontiers != 0, i, &sx, &sy, &sz);
    // Vector3d(±vs or 0) added component-wise; adding +0.0 is exact
    D3 q;
    q.x = __dadd_rn(c.x, sx > 0 ? e.nb_step : (sx < 0 ? -e.nb_step : 0.0));
    q.y = __dadd_rn(c.y, sy > 0 ? e.nb_step : (sy < 0 ? -e.nb_step : 0.0));
    q.z = __dadd_rn(c.z, sz > 0 ? e.nb_step : (sz < 0 ? -e.nb_step : 0.0));
    const int s = voxel_state(m, q);
    if (s == 2) continue;
    if (e.surface_frontiers) return s == 0;
    return true;
  }
  return false;
}

// order-sensitive digest (DESIGN.md "digest"): h = h * kDigestMul + digest_entry(c)
constexpr uint64_t kDigestMul = 0x100000001B3ull;
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t digest_entry(const D3& c) {
  uint64_t e = static_cast<uint64_t>(__double_as_longlong(c.x)) * 0x9E3779B97F4A7C15ull;
  e ^= rotl64(static_cast<uint64_t>(__double_as_longlong(c.y)), 23) * 0xC2B2AE3D27D4EB4Full;
  e ^= rotl64(static_cast<uint64_t>(__double_as_longlong(c.z)), 47) * 0x165667B19E3779F9ull;
  e ^= e >> 31;
  return e;
}

}  // namespace a3d
