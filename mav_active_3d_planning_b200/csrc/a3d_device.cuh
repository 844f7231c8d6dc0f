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
//
// HBM layout of the map mirror (DESIGN.md "data layout"):
//   table   open-addressing hash, 16 B entries {bx,by,bz,slot}
//   dist    per slot a PADDED cube of P^3 floats, P = vps+2: voxel (x,y,z), x,y,z in [-1,vps], lives at
//           ((z+1)*P + (y+1))*P + (x+1).  The interior is the block's ESDF distances; the one-voxel
//           apron mirrors the 26 neighbouring blocks (kUnobservedBits where the neighbour is not
//           allocated), so the 8 corners of any interpolation cell are addressed inside ONE block.
//   mword   per slot vps^3 32-bit meta words (voxel linear order x + vps*(y + vps*z)): bits 2c+1:2c =
//           class of the interpolation cell whose base corner is voxel - (c&1, c>>1&1, c>>2&1), c in
//           0..7, i.e. the 8 cells a point inside this voxel can interpolate in; bits 17:16 exact
//           voxel state at this voxel's centre; bit 18 observed.  One load per sample serves both
//           getVoxelState and the evaluator's lookup of the emitted centre.  Built on the device
//           after every delta (k_build_meta).
//   dir     dense directory of slots over the bounding box of allocated block indices, an
//           accelerator in front of the hash (1 load, no probe loop); absent for huge extents
//   weight  per slot vps^3 floats (TSDF weights, nearest-voxel reads only)
//   vol     the meta words a second time as ONE dense array over the directory's box, indexed by global
//           voxel coordinate (4x4x2 bricks = one 128-byte line each): the throughput kernel's
//           certified samples need a single load and no block lookup (a3d_wave.cu sample_prepare)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// -DA3D_BOUNDS_CHECK: every computed index into the map mirror / kernel scratch is checked on the
// device and a violation traps (compute-sanitizer is not available on the GPU pool this was
// developed on; the GPU test-suite is run once against such a build).
#ifdef A3D_BOUNDS_CHECK
#include <cstdio>
#define A3D_DEV_CHECK(cond)                                                                  \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      printf("A3D_DEV_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__);               \
      __trap();                                                                              \
    }                                                                                        \
  } while (0)
#else
#define A3D_DEV_CHECK(cond) \
  do {                      \
  } while (0)
#endif

namespace a3d {

// Unobserved voxels are stored as this exact bit pattern in the distance slab (observed flag
// folded into the 4-byte record).  Observed NaN distances are canonicalised to 0x7FC00000 on upload.
constexpr uint32_t kUnobservedBits = 0x7FD5A3D1u;
constexpr float kCoordEps = 1e-6f;

// cell classes (meta bits 1:0)
enum { kCellUnknown = 0, kCellFree = 1, kCellOccupied = 2, kCellMixed = 3 };
constexpr int kMaxVps = 64;

struct HashEntry {  // 16 B, one LDG.128 per probe
  int32_t x, y, z;
  int32_t slot;  // < 0: empty
};

struct MapView {
  const HashEntry* __restrict__ table;
  const float* __restrict__ dist;      // padded slab
  const uint32_t* __restrict__ mword;  // meta words, vps^3 per slot
  const float* __restrict__ weight;    // unpadded TSDF weights slab or nullptr
  const int32_t* __restrict__ dir;     // block directory or nullptr
  int32_t dir_lo[3];
  uint32_t dir_n[3];
  uint32_t mask;                     // table capacity - 1
  int32_t vps;
  int32_t nv;                        // vps^3
  int32_t P;                         // vps + 2
  int32_t P3;                        // P^3
  int32_t n_slots;                   // slots the slabs hold (bounds checks)
  int32_t brick_nx;                  // vps/4 if the meta words are stored in 4x4x2 bricks, else 0
  int32_t vps_shift;                 // log2(vps) when vps is a power of two, else -1
  // dense meta volume: the meta words again, indexed by GLOBAL voxel coordinate g = block*vps + voxel
  // over the directory's box, in 4x4x2 bricks (vol_offset); kWordNoBlock where no block is allocated.
  // nullptr when the box is too large or vps is not a power of two >= 4.
  const uint32_t* __restrict__ vol;
  int32_t vol_lo[3];                 // first voxel coordinate of the box
  uint32_t vol_n[3];                 // box size in voxels
  uint32_t vol_bx, vol_by;           // box size in bricks along x, y (vol_n >> 2)
  float fast_c0;                     // absolute part of the index-certification margin (a3d_wave.cu)
  float vs_f, vs_inv_f, bs_f, bs_inv_f;
  double c_vs;                       // double(vs_f), VoxbloxMap::c_voxel_size_
};

struct D3 {
  double x, y, z;
};
struct Q4 {
  double w, x, y, z;
};

// Per-CTA shared-memory tables of voxblox getCenterPointFromGridIndex(i, voxel_size) for
// i in [-1, vps]: removes 3 int→float→double→float conversion chains per lookup from the XU pipe.
struct CenterTables {
  float f[kMaxVps + 2];
  double d[kMaxVps + 2];
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
  if (m.dir) {
    const uint32_t ux = static_cast<uint32_t>(bx - m.dir_lo[0]);
    const uint32_t uy = static_cast<uint32_t>(by - m.dir_lo[1]);
    const uint32_t uz = static_cast<uint32_t>(bz - m.dir_lo[2]);
    if (ux >= m.dir_n[0] || uy >= m.dir_n[1] || uz >= m.dir_n[2]) return -1;
    return __ldg(m.dir + (static_cast<size_t>(uz) * m.dir_n[1] + uy) * m.dir_n[0] + ux);
  }
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

__device__ __forceinline__ void init_center_tables(const MapView& m, CenterTables* t) {
  for (int i = threadIdx.x; i < m.vps + 2; i += blockDim.x) {
    const float c = center_point(i - 1, m.vs_f);
    t->f[i] = c;
    t->d[i] = static_cast<double>(c);
  }
}

// Index of voxel (x,y,z) in a block's meta-word array.  When vps is a multiple of 4 the words are
// stored in 4x4x2 bricks (32 words = one 128-byte line) so that a step in any direction usually stays
// in the line the lane touched last; otherwise plain voxblox linear order.
__device__ __forceinline__ int mword_offset(const MapView& m, int x, int y, int z) {
  A3D_DEV_CHECK(static_cast<unsigned>(x) < static_cast<unsigned>(m.vps) &&
                static_cast<unsigned>(y) < static_cast<unsigned>(m.vps) &&
                static_cast<unsigned>(z) < static_cast<unsigned>(m.vps));
  if (m.brick_nx) {
    const int brick = ((z >> 1) * m.brick_nx + (y >> 2)) * m.brick_nx + (x >> 2);
    return (brick << 5) | ((z & 1) << 4) | ((y & 3) << 2) | (x & 3);
  }
  return (z * m.vps + y) * m.vps + x;
}

// Offset of box-relative voxel (u0,u1,u2) in the dense meta volume: 4x4x2 bricks of 32 words (one
// 128-byte line), bricks x fastest.  vol_n[] are multiples of vps, vps a power of two >= 4.
__device__ __forceinline__ uint32_t vol_offset(const MapView& m, uint32_t u0, uint32_t u1, uint32_t u2) {
  A3D_DEV_CHECK(u0 < m.vol_n[0] && u1 < m.vol_n[1] && u2 < m.vol_n[2]);
  const uint32_t brick = ((u2 >> 1) * m.vol_by + (u1 >> 2)) * m.vol_bx + (u0 >> 2);
  return (brick << 5) | ((u2 & 1u) << 4) | ((u1 & 3u) << 2) | (u0 & 3u);
}

__device__ __forceinline__ size_t padded_index(const MapView& m, int slot, int x, int y, int z) {
  A3D_DEV_CHECK(slot >= 0 && slot < m.n_slots && x >= -1 && x <= m.vps && y >= -1 && y <= m.vps && z >= -1 &&
                z <= m.vps);
  return static_cast<size_t>(slot) * m.P3 + ((z + 1) * m.P + (y + 1)) * m.P + (x + 1);
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

// voxblox Interpolator::interpMember: q * (T * d) with the summation order fixed in SURVEY.md A.3
// (zero-coefficient terms dropped).  d[i]: corner (i>>2&1, i>>1&1, i&1); o: voxel offsets.
__device__ __forceinline__ float interp8(const float (&d)[8], const float (&o)[3]) {
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
  return __fadd_rn(__fadd_rn(__fadd_rn(u0, u4), __fadd_rn(u2, u6)),
                   __fadd_rn(__fadd_rn(u1, u5), __fadd_rn(u3, u7)));
}

// ---- interpolated ESDF distance, literal generic path --------------------------------------------
// One hash lookup per distinct corner block, interiors only (never reads aprons or meta): this is
// the definition the apron/meta fast path must agree with, the builder of the meta bytes, and the
// fallback for the floating-point artefact cases (voxel index outside [0,vps) before the shift).
static __device__ __noinline__ bool interp_distance_generic(const MapView& m, float px, float py, float pz,
                                                     float* dist_out) {
  const float p[3] = {px, py, pz};
  int b[3], v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) b[k] = grid_index(p[k], m.bs_inv_f);
  if (find_slot(m, b[0], b[1], b[2]) < 0) return false;
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
      }
    }
  }
  float d[8];
  float o[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    int vi[3] = {v[0] + ((i >> 2) & 1), v[1] + ((i >> 1) & 1), v[2] + (i & 1)};
    int nb[3] = {b[0], b[1], b[2]};
    if (find_slot(m, nb[0], nb[1], nb[2]) < 0) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (vi[k] >= m.vps) {
        nb[k]++;
        vi[k] -= m.vps;
      }
    }
    const int s = find_slot(m, nb[0], nb[1], nb[2]);
    if (s < 0) return false;
    if (i == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float voxel_pos = __fadd_rn(origin_point(nb[k], m.bs_f), center_point(vi[k], m.vs_f));
        o[k] = __fmul_rn(__fsub_rn(p[k], voxel_pos), m.vs_inv_f);
      }
    }
    // voxblox computeLinearIndexFromVoxelIndex; out-of-range is UB upstream, defined as "no voxel"
    const int lin = vi[0] + m.vps * (vi[1] + m.vps * vi[2]);
    if (lin < 0 || lin >= m.nv) return false;
    const int z = lin / (m.vps * m.vps), rem = lin - z * m.vps * m.vps;
    const int y = rem / m.vps, x = rem - y * m.vps;
    const float dv = __ldg(m.dist + padded_index(m, s, x, y, z));
    if (__float_as_uint(dv) == kUnobservedBits) return false;
    d[i] = dv;
  }
  *dist_out = interp8(d, o);
  return true;
}

__device__ __forceinline__ int state_from_distance(const MapView& m, bool ok, float dist) {
  if (!ok) return 2;                                    // UNKNOWN
  return (static_cast<double>(dist) < m.c_vs) ? 0 : 1;  // OCCUPIED : FREE
}

static __device__ __noinline__ int voxel_state_generic(const MapView& m, float px, float py, float pz) {
  float dist = 0.f;
  const bool ok = interp_distance_generic(m, px, py, pz, &dist);
  return state_from_distance(m, ok, dist);
}

// ---- fast path ----------------------------------------------------------------------------------
constexpr uint32_t kWordNoBlock = 0x00020000u;  // centre state UNKNOWN, not observed, all cells UNKNOWN
// Meta-word bits 23:19 (k_build_frontier): FrontierEvaluator::isFrontierVoxel of this voxel for
// checking_distance 1, precomputed for every voxel whose centre is UNKNOWN or unobserved.
//   bit 19: 6-neighbourhood, surface_frontiers    bit 20: 6-neighbourhood, any known neighbour
//   bit 21: 26-neighbourhood, surface_frontiers   bit 22: 26-neighbourhood, any known neighbour
//   bit 23: bits 22:19 are valid
constexpr uint32_t kWordFrontierValid = 1u << 23;

// State of the float point p inside a MIXED interpolation cell: voxel v (in range) of allocated block
// b (slot), octant sc.  Reads the 8 corners from the padded cube and interpolates like the reference.
__device__ __forceinline__ int mixed_cell_state(const MapView& m, const CenterTables& ct,
                                                const float (&p)[3], const int (&b)[3],
                                                const float (&origin)[3], int slot, const int (&v)[3],
                                                int sc) {
  const int base[3] = {v[0] - (sc & 1), v[1] - ((sc >> 1) & 1), v[2] - (sc >> 2)};
  float d[8], o[3];
  const float* __restrict__ dp = m.dist + padded_index(m, slot, base[0], base[1], base[2]);
  const int P = m.P, PP = m.P * m.P;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    d[i] = __ldg(dp + ((i >> 2) & 1) + ((i >> 1) & 1) * P + (i & 1) * PP);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // corner 0 position as the reference forms it: in the shifted block's own frame
    const float voxel_pos = base[k] < 0
                                ? __fadd_rn(origin_point(b[k] - 1, m.bs_f), ct.f[m.vps])
                                : __fadd_rn(origin[k], ct.f[base[k] + 1]);
    o[k] = __fmul_rn(__fsub_rn(p[k], voxel_pos), m.vs_inv_f);
  }
  return state_from_distance(m, true, interp8(d, o));
}

// State of the float point p whose block index b, block origin and slot (= find_slot(b), may be
// < 0) the caller already has.  Uses the cell class when it is definite, else the 8 padded-cube
// corners; falls back to the generic path for artefact indices.  Also returns the voxel index v the
// point falls in and that voxel's meta word (valid when the return value of *v_ok is true).
__device__ __forceinline__ int voxel_state_in_block(const MapView& m, const CenterTables& ct,
                                                    const float (&p)[3], const int (&b)[3],
                                                    const float (&origin)[3], int slot, int (&v)[3],
                                                    uint32_t* word_out, bool* v_ok) {
  bool artefact = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[k] = grid_index(__fsub_rn(p[k], origin[k]), m.vs_inv_f);
    artefact |= static_cast<unsigned>(v[k]) >= static_cast<unsigned>(m.vps);
  }
  *v_ok = !artefact;
  *word_out = kWordNoBlock;
  if (slot < 0) return 2;
  if (artefact) return voxel_state_generic(m, p[0], p[1], p[2]);
  int sc = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float centre = __fadd_rn(origin[k], ct.f[v[k] + 1]);
    sc |= (__fsub_rn(p[k], centre) < 0.0f) ? (1 << k) : 0;
  }
  const uint32_t word =
      __ldg(m.mword + static_cast<size_t>(slot) * m.nv + mword_offset(m, v[0], v[1], v[2]));
  *word_out = word;
  const int cls = (word >> (2 * sc)) & 3;
  if (cls != kCellMixed) return cls == kCellUnknown ? 2 : (cls == kCellFree ? 1 : 0);
  return mixed_cell_state(m, ct, p, b, origin, slot, v, sc);
}

__device__ __forceinline__ int voxel_state(const MapView& m, const CenterTables& ct, const D3& pd) {
  const float p[3] = {__double2float_rn(pd.x), __double2float_rn(pd.y), __double2float_rn(pd.z)};
  int b[3];
  float origin[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    b[k] = grid_index(p[k], m.bs_inv_f);
    origin[k] = origin_point(b[k], m.bs_f);
  }
  int v[3];
  uint32_t word;
  bool v_ok;
  return voxel_state_in_block(m, ct, p, b, origin, find_slot(m, b[0], b[1], b[2]), v, &word, &v_ok);
}

// getVoxelCenter given the block index / origin of the point; v_out: unclamped voxel index.
__device__ __forceinline__ D3 voxel_center_in_block(const MapView& m, const CenterTables& ct,
                                                    const D3& p, const float (&origin)[3],
                                                    int (&v_out)[3]) {
  const double pp[3] = {p.x, p.y, p.z};
  double c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double od = static_cast<double>(origin[k]);
    const int v = grid_index(__double2float_rn(__dsub_rn(pp[k], od)), m.vs_inv_f);
    const double cp = (v >= -1 && v <= m.vps) ? ct.d[v + 1]
                                              : static_cast<double>(center_point(v, m.vs_f));
    c[k] = __dadd_rn(od, cp);
    v_out[k] = v;
  }
  return D3{c[0], c[1], c[2]};
}

__device__ __forceinline__ D3 voxel_center(const MapView& m, const CenterTables& ct, const D3& p) {
  float origin[3];
  origin[0] = origin_point(grid_index(__double2float_rn(p.x), m.bs_inv_f), m.bs_f);
  origin[1] = origin_point(grid_index(__double2float_rn(p.y), m.bs_inv_f), m.bs_f);
  origin[2] = origin_point(grid_index(__double2float_rn(p.z), m.bs_inv_f), m.bs_f);
  int v[3];
  return voxel_center_in_block(m, ct, p, origin, v);
}

// Nearest-voxel address parts (block by coordinates, truncated voxel index); slot -1 if no block.
__device__ __forceinline__ int nearest_voxel(const MapView& m, const D3& p, int (&v)[3]) {
  const float pf[3] = {__double2float_rn(p.x), __double2float_rn(p.y), __double2float_rn(p.z)};
  int b[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) b[k] = grid_index(pf[k], m.bs_inv_f);
  const int slot = find_slot(m, b[0], b[1], b[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[k] = grid_index(__fsub_rn(pf[k], origin_point(b[k], m.bs_f)), m.vs_inv_f);
    v[k] = max(min(v[k], m.vps - 1), 0);
  }
  return slot;
}
__device__ __forceinline__ bool is_observed(const MapView& m, const D3& p) {
  int v[3];
  const int slot = nearest_voxel(m, p, v);
  if (slot < 0) return false;
  return (__ldg(m.mword + static_cast<size_t>(slot) * m.nv + mword_offset(m, v[0], v[1], v[2])) >>
          18) & 1;
}
__device__ __forceinline__ double voxel_weight(const MapView& m, const D3& p) {
  if (!m.weight) return 0.0;
  int v[3];
  const int slot = nearest_voxel(m, p, v);
  if (slot < 0) return 0.0;
  return static_cast<double>(
      __ldg(m.weight + static_cast<size_t>(slot) * m.nv + v[0] + m.vps * (v[1] + m.vps * v[2])));
}

// State / observed flag at an emitted voxel CENTRE c = voxel_center(...) with block index b and
// unclamped voxel index v, slot = find_slot(b): one meta byte (DESIGN.md "centre state").
// Returns state | observed << 2.
__device__ __forceinline__ int centre_state(const MapView& m, const CenterTables& ct, const D3& c,
                                            int slot, const int (&v)[3]) {
  const bool in_range = (v[0] >= 0) & (v[0] < m.vps) & (v[1] >= 0) & (v[1] < m.vps) & (v[2] >= 0) &
                        (v[2] < m.vps);
  if (!in_range) {  // fp artefact: voxel index -1 / vps — evaluate literally
    return voxel_state(m, ct, c) | (is_observed(m, c) ? 4 : 0);
  }
  if (slot < 0) return 2;
  const uint32_t word =
      __ldg(m.mword + static_cast<size_t>(slot) * m.nv + mword_offset(m, v[0], v[1], v[2]));
  return static_cast<int>((word >> 16) & 7);
}
// the same from an already loaded meta word of that voxel
__device__ __forceinline__ int centre_state_from_word(uint32_t word) {
  return static_cast<int>((word >> 16) & 7);
}

struct EvalView {
  int32_t kind;
  int32_t accurate_frontiers, surface_frontiers;
  int32_t bv_is_setup;
  double nb_step;  // c_voxel_size_ * checking_distance (frontier_evaluator.cpp:31)
  double gains[6];
  double bv_min[3], bv_max[3];
  Q4 bv_q;
  // VoxelWeightEvaluator (voxel_weight_evaluator.cpp:17-22)
  double frontier_voxel_weight, min_impact_factor, new_voxel_weight, ray_angle_xy;  // rx*ry
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

// FrontierEvaluator neighbour order (frontier_evaluator.cpp:33-65).
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
  const int r3 = r / 3, rm = r - 3 * r3;
  *sx = g == 0 ? 1 : (g == 1 ? 0 : -1);
  *sy = rm == 0 ? 0 : (rm == 1 ? 1 : -1);
  *sz = r3 == 0 ? 0 : (r3 == 1 ? 1 : -1);
}

__device__ __forceinline__ bool is_frontier_voxel(const MapView& m, const CenterTables& ct,
                                                  const EvalView& e, const D3& c) {
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
    const int s = voxel_state(m, ct, q);
    if (s == 2) continue;
    if (e.surface_frontiers) return s == 0;
    return true;
  }
  return false;
}

// is_frontier_voxel for the voxel with global coordinate g and centre double c, from meta words where
// they suffice.  With checking_distance == 1 the neighbour points c ± voxel_size lie within a few float
// ulps of the neighbour voxels' centres, i.e. inside the neighbour voxel where its 8 interpolation
// cells meet; when all 8 cell classes of that voxel agree, the state there is that class whichever
// cell the rounding picks.  A neighbour whose cells disagree is evaluated literally.  Returns 1 / 0,
// or -1 when the shortcut does not apply at all (no dense volume, other checking distance).
__device__ __forceinline__ uint32_t neighbour_cells(const MapView& m, int g0, int g1, int g2) {
  const uint32_t u0 = static_cast<uint32_t>(g0 - m.vol_lo[0]), u1 = static_cast<uint32_t>(g1 - m.vol_lo[1]),
                 u2 = static_cast<uint32_t>(g2 - m.vol_lo[2]);
  uint32_t cells = 0;  // outside the box: no block, all cells UNKNOWN
  if ((u0 < m.vol_n[0]) & (u1 < m.vol_n[1]) & (u2 < m.vol_n[2]))
    cells = __ldg(m.vol + vol_offset(m, u0, u1, u2)) & 0xFFFFu;
  return cells;
}
static __device__ __noinline__ int neighbour_state_literal(const MapView& m, const CenterTables& ct,
                                                          const EvalView& e, const D3& c, int sx, int sy,
                                                          int sz) {
  D3 q;
  q.x = __dadd_rn(c.x, sx > 0 ? e.nb_step : (sx < 0 ? -e.nb_step : 0.0));
  q.y = __dadd_rn(c.y, sy > 0 ? e.nb_step : (sy < 0 ? -e.nb_step : 0.0));
  q.z = __dadd_rn(c.z, sz > 0 ? e.nb_step : (sz < 0 ? -e.nb_step : 0.0));
  return voxel_state(m, ct, q);
}
__device__ __forceinline__ int neighbour_state(const MapView& m, const CenterTables& ct, const EvalView& e,
                                               const D3& c, uint32_t cells, int sx, int sy, int sz) {
  if (cells == 0x0000u) return 2;  // UNKNOWN everywhere
  if (cells == 0x5555u) return 1;  // FREE everywhere
  if (cells == 0xAAAAu) return 0;  // OCCUPIED everywhere
  return neighbour_state_literal(m, ct, e, c, sx, sy, sz);
}
__device__ __forceinline__ int is_frontier_by_meta(const MapView& m, const CenterTables& ct,
                                                   const EvalView& e, const D3& c, int g0, int g1, int g2) {
  if (!m.vol || e.nb_step != m.c_vs) return -1;
  if (!e.accurate_frontiers) {
    // 6-neighbourhood (+x,-x,+y,-y,+z,-z): all six words requested before the first is looked at
    uint32_t cells[6];
    cells[0] = neighbour_cells(m, g0 + 1, g1, g2);
    cells[1] = neighbour_cells(m, g0 - 1, g1, g2);
    cells[2] = neighbour_cells(m, g0, g1 + 1, g2);
    cells[3] = neighbour_cells(m, g0, g1 - 1, g2);
    cells[4] = neighbour_cells(m, g0, g1, g2 + 1);
    cells[5] = neighbour_cells(m, g0, g1, g2 - 1);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int axis = i >> 1, sgn = (i & 1) ? -1 : 1;
      const int s = neighbour_state(m, ct, e, c, cells[i], axis == 0 ? sgn : 0, axis == 1 ? sgn : 0,
                                    axis == 2 ? sgn : 0);
      if (s == 2) continue;
      return e.surface_frontiers ? (s == 0) : 1;
    }
    return 0;
  }
#pragma unroll 1
  for (int i = 0; i < 26; ++i) {
    int sx, sy, sz;
    frontier_offset(true, i, &sx, &sy, &sz);
    const int s = neighbour_state(m, ct, e, c, neighbour_cells(m, g0 + sx, g1 + sy, g2 + sz), sx, sy, sz);
    if (s == 2) continue;
    return e.surface_frontiers ? (s == 0) : 1;
  }
  return 0;
}

// is_frontier_voxel from the precomputed bits of the voxel's own meta word: 1 / 0, or -1 when the
// word carries none (no block, other checking distance).
__device__ __forceinline__ int is_frontier_from_word(const MapView& m, const EvalView& e, uint32_t word) {
  if (!(word & kWordFrontierValid) || e.nb_step != m.c_vs) return -1;
  const int bit = (e.accurate_frontiers ? 21 : 19) + (e.surface_frontiers ? 0 : 1);
  return static_cast<int>((word >> bit) & 1u);
}

// Meta word of the voxel whose centre double is, bit for bit, the stored list entry c (true for every
// entry that came from a canonical voxel).  Such an entry's voxel state (= the precomputed centre
// state), observed flag and frontier bits are all in that word.  False: the caller evaluates the
// entry literally.
__device__ __forceinline__ bool centre_word(const MapView& m, const CenterTables& ct, const D3& c,
                                            uint32_t* word_out) {
  const double cd[3] = {c.x, c.y, c.z};
  int b[3], v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float p = __double2float_rn(cd[k]);
    b[k] = grid_index(p, m.bs_inv_f);
    const float origin = origin_point(b[k], m.bs_f);
    v[k] = grid_index(__fsub_rn(p, origin), m.vs_inv_f);
    if (static_cast<unsigned>(v[k]) >= static_cast<unsigned>(m.vps)) return false;
    if (__dadd_rn(static_cast<double>(origin), ct.d[v[k] + 1]) != cd[k]) return false;
  }
  uint32_t word = kWordNoBlock;
  if (m.vol) {
    const uint32_t u0 = static_cast<uint32_t>(b[0] * m.vps + v[0] - m.vol_lo[0]),
                   u1 = static_cast<uint32_t>(b[1] * m.vps + v[1] - m.vol_lo[1]),
                   u2 = static_cast<uint32_t>(b[2] * m.vps + v[2] - m.vol_lo[2]);
    if ((u0 < m.vol_n[0]) & (u1 < m.vol_n[1]) & (u2 < m.vol_n[2])) word = __ldg(m.vol + vol_offset(m, u0, u1, u2));
  } else {
    const int slot = find_slot(m, b[0], b[1], b[2]);
    if (slot >= 0) word = __ldg(m.mword + static_cast<size_t>(slot) * m.nv + mword_offset(m, v[0], v[1], v[2]));
  }
  *word_out = word;
  return true;
}
__device__ __forceinline__ int frontier_hint_for_centre(const MapView& m, const CenterTables& ct,
                                                        const EvalView& e, const D3& c) {
  uint32_t word;
  if (!centre_word(m, ct, c, &word)) return -1;
  return is_frontier_from_word(m, e, word);
}
// the frontier test of a stored list entry: precomputed bit where it applies, else literal
__device__ __forceinline__ bool is_frontier_entry(const MapView& m, const CenterTables& ct, const EvalView& e,
                                                  const D3& c) {
  const int hint = frontier_hint_for_centre(m, ct, e, c);
  return hint >= 0 ? hint != 0 : is_frontier_voxel(m, ct, e, c);
}

// VoxelWeightEvaluator::getVoxelValue (voxel_weight_evaluator.cpp:60-83) for a list entry with centre
// c, given its state cs (centre_state) and TSDF weight w.  frontier_hint: is_frontier_by_meta() of the
// voxel or -1.
__device__ __forceinline__ double voxel_weight_value(const MapView& m, const CenterTables& ct,
                                                     const EvalView& e, const D3& c, int cs, double w,
                                                     const D3& origin, int frontier_hint = -1) {
  const int state = cs & 3;
  if (state == 0) {  // surface voxel
    const double dx = c.x - origin.x, dy = c.y - origin.y, dz = c.z - origin.z;
    const double z = sqrt((dx * dx + dy * dy) + dz * dz);
    const double spanned_angle = 2.0 * atan2(m.c_vs, z * 2.0);
    const double new_weight = spanned_angle * spanned_angle / e.ray_angle_xy / (z * z);
    const double gain = new_weight / (new_weight + w);
    return gain > e.min_impact_factor ? gain : 0.0;
  }
  if (state == 2) {  // unobserved voxel
    if (e.frontier_voxel_weight > 0.0) {
      const bool frontier = frontier_hint >= 0 ? frontier_hint != 0 : is_frontier_voxel(m, ct, e, c);
      if (frontier) return e.frontier_voxel_weight;
    }
    return e.new_voxel_weight;
  }
  return 0.0;
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
