/*
 * a3d_oracle.cpp — CPU oracle for the view-gain hot path.  TEST INFRASTRUCTURE ONLY
 * (see a3d_oracle.h).
 *
 * PINNED against the reference's own code: oracle/_ref compiles the unmodified reference sources of
 * the path (ref_driver.cpp lists them) behind this same API, and tests/test_ref_pin.py demands
 * `_ref == oracle` bit for bit (ordered lists, counts, digests, gains; cfg1/cfg2/cfg4 poses, mounting,
 * multi-view segments, all five evaluators incl. clear_from_parents, 200 fuzz cases).  What reference
 * code does NOT pin — because it lives in un-vendored dependencies that are restated in both libraries —
 * is the voxblox arithmetic (grid-index helpers, nearest-voxel access, the 8-corner interpolation:
 * ref_shim/voxblox_shim.h) and the floating-point order of six Eigen expressions (ref_shim/eigen_shim.h).
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * "voxblox:" citations name files of the un-vendored dependency github.com/ethz-asl/voxblox
 * (unpinned default branch), restated from its published algorithm.
 *
 * Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC (oracle/Makefile).  No FMA
 * contraction, no fast-math: the op order written here IS the specification.
 */
#include "a3d_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct V3 {
  double x, y, z;
  bool operator==(const V3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct Q4 {
  double w, x, y, z;
};
struct I3 {
  int x, y, z;
  bool operator==(const I3& o) const { return x == o.x && y == o.y && z == o.z; }
};

// Eigen 3.3 Vector3d::cross (coefficient formula)
inline V3 cross(const V3& a, const V3& b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen 3.3 QuaternionBase::_transformVector: uv = q.vec x v; uv += uv; v + w*uv + q.vec x uv
inline V3 rotate(const Q4& q, const V3& v) {
  const V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = V3{uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  const V3 c = cross(qv, uv);
  return V3{(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}
// Hamilton product, textbook scalar order
inline Q4 qmul(const Q4& a, const Q4& b) {
  return Q4{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
            a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

// voxblox: core/block_hash.h AnyIndexHash
struct I3Hash {
  std::size_t operator()(const I3& i) const {
    constexpr std::size_t sl = 17191, sl2 = sl * sl;
    return static_cast<unsigned int>(i.x + i.y * sl + i.z * sl2);
  }
};

struct Block {
  float origin[3];
  std::vector<float> dist;
  std::vector<uint8_t> observed;
  std::vector<float> weight;  // TSDF layer weight (empty → block absent from the TSDF layer)
};

constexpr float kCoordinateEpsilon = 1e-6f;  // voxblox: core/common.h

// voxblox: core/common.h getGridIndexFromPoint (scalar component)
inline int gridIndex(float p, float inv) {
  return static_cast<int>(std::floor(p * inv + kCoordinateEpsilon));
}
// voxblox: core/common.h getCenterPointFromGridIndex (0.5 literal is double)
inline float centerPoint(int idx, float size) {
  return static_cast<float>((static_cast<float>(idx) + 0.5) * size);
}
// voxblox: core/common.h getOriginPointFromGridIndex
inline float originPoint(int idx, float size) { return static_cast<float>(idx) * size; }

}  // namespace

struct orc_map {
  float vs_f, vs_inv_f, bs_f, bs_inv_f;
  int vps;
  double c_voxel_size, c_block_size;  // vbx/src/map/voxblox.cpp:26-27
  std::unordered_map<I3, std::shared_ptr<Block>, I3Hash> blocks;

  // voxblox: core/layer.h getBlockPtrByIndex — returns a shared_ptr copy like upstream
  std::shared_ptr<Block> blockByIndex(const I3& b) const {
    auto it = blocks.find(b);
    if (it == blocks.end()) return std::shared_ptr<Block>();
    return it->second;
  }
  I3 blockIndexFromCoords(const float p[3]) const {
    return I3{gridIndex(p[0], bs_inv_f), gridIndex(p[1], bs_inv_f), gridIndex(p[2], bs_inv_f)};
  }
  // voxblox: core/block.h computeLinearIndexFromVoxelIndex; out-of-range is UB upstream, here
  // defined as "no voxel" (-1).
  int64_t linearIndex(const int v[3]) const {
    const int64_t lin = v[0] + static_cast<int64_t>(vps) * (v[1] + static_cast<int64_t>(vps) * v[2]);
    if (lin < 0 || lin >= static_cast<int64_t>(vps) * vps * vps) return -1;
    return lin;
  }
  // voxblox: core/block.h computeTruncatedVoxelIndexFromCoordinates
  int64_t truncatedLinearIndex(const Block& blk, const float p[3]) const {
    int v[3];
    for (int k = 0; k < 3; ++k) {
      v[k] = gridIndex(p[k] - blk.origin[k], vs_inv_f);
      v[k] = std::max(std::min(v[k], vps - 1), 0);
    }
    return linearIndex(v);
  }

  // voxblox: core/esdf_map.cc getDistanceAtPosition(position, interpolate=true) →
  // interpolator/interpolator_inl.h getInterpDistance / setIndexes / getVoxelsAndQVector /
  // getQVector / interpMember.  SURVEY.md A.3.
  bool getDistanceAtPosition(const V3& position, double* distance) const {
    const float p[3] = {static_cast<float>(position.x), static_cast<float>(position.y),
                        static_cast<float>(position.z)};
    // setIndexes
    I3 b = blockIndexFromCoords(p);
    std::shared_ptr<Block> block_ptr = blockByIndex(b);
    if (!block_ptr) return false;
    int v[3];
    int* bi[3] = {&b.x, &b.y, &b.z};
    for (int k = 0; k < 3; ++k) v[k] = gridIndex(p[k] - block_ptr->origin[k], vs_inv_f);
    for (int k = 0; k < 3; ++k) {
      const float centre = block_ptr->origin[k] + centerPoint(v[k], vs_f);
      const float center_offset = p[k] - centre;
      if (center_offset < 0) {
        v[k]--;
        if (v[k] < 0) {
          (*bi[k])--;
          v[k] += vps;
        }
      }
    }
    // getVoxelsAndQVector: 8 corners, column order (0,0,0),(0,0,1),(0,1,0),...,(1,1,1)
    float d[8];
    float q[8];
    for (int i = 0; i < 8; ++i) {
      std::shared_ptr<Block> bp = blockByIndex(b);
      if (!bp) return false;
      int vi[3] = {v[0] + ((i >> 2) & 1), v[1] + ((i >> 1) & 1), v[2] + (i & 1)};
      if (vi[0] >= vps || vi[1] >= vps || vi[2] >= vps) {
        I3 nb = b;
        int* nbi[3] = {&nb.x, &nb.y, &nb.z};
        for (int k = 0; k < 3; ++k) {
          if (vi[k] >= vps) {
            (*nbi[k])++;
            vi[k] -= vps;
          }
        }
        bp = blockByIndex(nb);
        if (!bp) return false;
      }
      if (i == 0) {
        // getQVector(voxel_pos, pos, voxel_size_inv)
        float o[3];
        for (int k = 0; k < 3; ++k) {
          const float voxel_pos = bp->origin[k] + centerPoint(vi[k], vs_f);
          o[k] = (p[k] - voxel_pos) * vs_inv_f;
        }
        q[0] = 1.0f;
        q[1] = o[0];
        q[2] = o[1];
        q[3] = o[2];
        q[4] = o[0] * o[1];
        q[5] = o[1] * o[2];
        q[6] = o[2] * o[0];
        q[7] = o[0] * o[1] * o[2];
      }
      const int64_t lin = linearIndex(vi);
      if (lin < 0) return false;          // UB upstream; defined here
      if (!bp->observed[lin]) return false;  // isVoxelValid<EsdfVoxel>
      d[i] = bp->dist[lin];
    }
    // interpMember: q * (T * d); summation order fixed per SURVEY.md A.3 step 7
    // (Eigen 3.3.7 col-major GEMV 4-columns-at-once, SSE2 redux for the dot product);
    // zero-coefficient terms dropped (x + 0 == x).
    float w[8];
    w[0] = d[0];
    w[1] = d[4] - d[0];
    w[2] = d[2] - d[0];
    w[3] = d[1] - d[0];
    w[4] = (d[0] - d[2]) + (d[6] - d[4]);
    w[5] = (d[0] - d[1]) + (d[3] - d[2]);
    w[6] = (d[0] - d[1]) + (d[5] - d[4]);
    w[7] = ((d[1] - d[0]) + (d[2] - d[3])) + ((d[4] - d[5]) + (d[7] - d[6]));
    float u[8];
    for (int i = 0; i < 8; ++i) u[i] = q[i] * w[i];
    const float dist = ((u[0] + u[4]) + (u[2] + u[6])) + ((u[1] + u[5]) + (u[3] + u[7]));
    *distance = static_cast<double>(dist);
    return true;
  }

  // vbx/src/map/voxblox.cpp:48-60
  unsigned char getVoxelState(const V3& point) const {
    double distance = 0.0;
    if (getDistanceAtPosition(point, &distance)) {
      if (distance < c_voxel_size) return ORC_OCCUPIED;
      return ORC_FREE;
    }
    return ORC_UNKNOWN;
  }

  // vbx/src/map/voxblox.cpp:66-81
  V3 getVoxelCenter(const V3& point) const {
    const float p[3] = {static_cast<float>(point.x), static_cast<float>(point.y),
                        static_cast<float>(point.z)};
    const I3 b = blockIndexFromCoords(p);
    const float bsz = static_cast<float>(c_block_size);
    V3 c{static_cast<double>(originPoint(b.x, bsz)), static_cast<double>(originPoint(b.y, bsz)),
         static_cast<double>(originPoint(b.z, bsz))};
    const float inv = static_cast<float>(1.0 / c_voxel_size);
    const float vsz = static_cast<float>(c_voxel_size);
    const int vx = gridIndex(static_cast<float>(point.x - c.x), inv);
    const int vy = gridIndex(static_cast<float>(point.y - c.y), inv);
    const int vz = gridIndex(static_cast<float>(point.z - c.z), inv);
    c.x += static_cast<double>(centerPoint(vx, vsz));
    c.y += static_cast<double>(centerPoint(vy, vsz));
    c.z += static_cast<double>(centerPoint(vz, vsz));
    return c;
  }

  // vbx/src/map/voxblox.cpp:43-45 → voxblox: core/esdf_map.cc isObserved
  bool isObserved(const V3& point) const {
    const float p[3] = {static_cast<float>(point.x), static_cast<float>(point.y),
                        static_cast<float>(point.z)};
    std::shared_ptr<Block> bp = blockByIndex(blockIndexFromCoords(p));
    if (!bp) return false;
    const int64_t lin = truncatedLinearIndex(*bp, p);
    return bp->observed[lin] != 0;
  }

  // vbx/src/map/voxblox.cpp:101-115
  double getVoxelWeight(const V3& point) const {
    const float p[3] = {static_cast<float>(point.x), static_cast<float>(point.y),
                        static_cast<float>(point.z)};
    std::shared_ptr<Block> bp = blockByIndex(blockIndexFromCoords(p));
    if (!bp || bp->weight.empty()) return 0.0;
    const int64_t lin = truncatedLinearIndex(*bp, p);
    return static_cast<double>(bp->weight[lin]);
  }
};

struct orc_caster {
  const orc_map* map;
  orc_caster_params p;
  double fov_x, fov_y;
  double ray_step;
  int c_res_x, c_res_y;
  int c_n_sections;
  std::vector<double> c_split_distances;
  std::vector<int> c_split_widths;

  // camera_model.cpp:161-168; lidar_model.cpp:159-165 for the lidar caster
  V3 directionVector(double relative_x, double relative_y) const {
    if (p.kind == ORC_CASTER_ITERATIVE_LIDAR) {
      const double phi = (0.5 - relative_x) * fov_x;
      const double theta = M_PI / 2.0 + (relative_y - 0.5) * fov_y;
      return V3{std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta)};
    }
    const V3 v{p.focal_length, (0.5 - relative_x) * static_cast<double>(p.resolution_x),
               (0.5 - relative_y) * static_cast<double>(p.resolution_y)};
    const double n = std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z);  // Eigen normalized()
    return V3{v.x / n, v.y / n, v.z / n};
  }
  V3 rayDirection(int i, int j) const {
    return directionVector(static_cast<double>(i) / (static_cast<double>(c_res_x) - 1.0),
                           static_cast<double>(j) / (static_cast<double>(c_res_y) - 1.0));
  }

  // simple_ray_caster.cpp:32-66
  void simpleVisible(std::vector<V3>* result, const V3& position, const Q4& orientation,
                     int64_t* n_samples) const {
    for (int i = 0; i < c_res_x; ++i) {
      for (int j = 0; j < c_res_y; ++j) {
        const V3 direction = rotate(orientation, rayDirection(i, j));
        double distance = 0.0;
        while (distance < p.ray_length) {
          const V3 cur{position.x + distance * direction.x, position.y + distance * direction.y,
                       position.z + distance * direction.z};
          distance += ray_step;
          ++*n_samples;
          if (map->getVoxelState(cur) == ORC_OCCUPIED) break;
          result->push_back(map->getVoxelCenter(cur));
        }
      }
    }
  }

  // iterative_ray_caster.cpp:110-119
  void markNeighboringRays(std::vector<int>* table, int x, int y, int segment, int value) const {
    for (int i = x; i < std::min(c_res_x, x + c_split_widths[segment]); ++i)
      for (int j = y; j < std::min(c_res_y, y + c_split_widths[segment]); ++j)
        (*table)[static_cast<size_t>(i) * c_res_y + j] = value;
  }

  // iterative_ray_caster.cpp:48-108
  void iterativeVisible(std::vector<V3>* result, const V3& position, const Q4& orientation,
                        int64_t* n_samples) const {
    std::vector<int> table(static_cast<size_t>(c_res_x) * c_res_y, 0);
    for (int i = 0; i < c_res_x; ++i) {
      for (int j = 0; j < c_res_y; ++j) {
        int current_segment = table[static_cast<size_t>(i) * c_res_y + j];
        if (current_segment < 0) continue;
        const V3 direction = rotate(orientation, rayDirection(i, j));
        double distance = c_split_distances[current_segment];
        bool cast_ray = true;
        while (cast_ray) {
          while (distance < c_split_distances[current_segment + 1]) {
            const V3 cur{position.x + distance * direction.x, position.y + distance * direction.y,
                         position.z + distance * direction.z};
            distance += ray_step;
            result->push_back(map->getVoxelCenter(cur));
            ++*n_samples;
            if (map->getVoxelState(cur) == ORC_OCCUPIED) {
              markNeighboringRays(&table, i, j, current_segment, -1);
              cast_ray = false;
              break;
            }
          }
          if (cast_ray) {
            current_segment++;
            if (current_segment >= c_n_sections) {
              cast_ray = false;
            } else {
              markNeighboringRays(&table, i, j, current_segment - 1, current_segment);
            }
          }
        }
      }
    }
  }

  // camera_model.cpp:15-60 (poses already sampled)
  void visibleFromPoses(std::vector<V3>* result, int n_views, const double* pos,
                        const double* quat, int64_t* n_samples) const {
    const V3 mt{p.mount_t[0], p.mount_t[1], p.mount_t[2]};
    const Q4 mq{p.mount_q[0], p.mount_q[1], p.mount_q[2], p.mount_q[3]};
    for (int v = 0; v < n_views; ++v) {
      V3 position{pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]};
      Q4 orientation{quat[4 * v], quat[4 * v + 1], quat[4 * v + 2], quat[4 * v + 3]};
      const V3 r = rotate(orientation, mt);
      position = V3{position.x + r.x, position.y + r.y, position.z + r.z};
      orientation = qmul(orientation, mq);
      if (p.kind == ORC_CASTER_SIMPLE)
        simpleVisible(result, position, orientation, n_samples);
      else  // IterativeRayCaster and IterativeRayCasterLidar share the loop (only directions differ)
        iterativeVisible(result, position, orientation, n_samples);
    }
    result->erase(std::unique(result->begin(), result->end()), result->end());
  }
};

namespace {

// bounding_volume.cpp:33-57
inline bool bvContains(const orc_eval_params& e, const V3& point) {
  if (!e.bv_is_setup) return true;
  const Q4 q{e.bv_quat[0], e.bv_quat[1], e.bv_quat[2], e.bv_quat[3]};
  const V3 r = rotate(q, point);
  if (r.x < e.bv_min[0]) return false;
  if (r.x > e.bv_max[0]) return false;
  if (r.y < e.bv_min[1]) return false;
  if (r.y > e.bv_max[1]) return false;
  if (r.z < e.bv_min[2]) return false;
  if (r.z > e.bv_max[2]) return false;
  return true;
}

// frontier_evaluator.cpp:30-66 neighbour table
std::vector<V3> frontierOffsets(const orc_map& m, const orc_eval_params& e) {
  const double vs = m.c_voxel_size * e.checking_distance;
  static const int k6[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  static const int k26[26][3] = {
      {1, 0, 0},   {1, 1, 0},   {1, -1, 0},  {1, 0, 1},   {1, 1, 1},    {1, -1, 1},  {1, 0, -1},
      {1, 1, -1},  {1, -1, -1}, {0, 1, 0},   {0, -1, 0},  {0, 0, 1},    {0, 1, 1},   {0, -1, 1},
      {0, 0, -1},  {0, 1, -1},  {0, -1, -1}, {-1, 0, 0},  {-1, 1, 0},   {-1, -1, 0}, {-1, 0, 1},
      {-1, 1, 1},  {-1, -1, 1}, {-1, 0, -1}, {-1, 1, -1}, {-1, -1, -1}};
  // the reference writes Vector3d(vs, 0, 0), Vector3d(-vs, ...): components are +vs, -vs or 0
  auto comp = [vs](int s) { return s > 0 ? vs : (s < 0 ? -vs : 0.0); };
  std::vector<V3> n;
  const int cnt = e.accurate_frontiers ? 26 : 6;
  for (int i = 0; i < cnt; ++i) {
    const int* o = e.accurate_frontiers ? k26[i] : k6[i];
    n.push_back(V3{comp(o[0]), comp(o[1]), comp(o[2])});
  }
  return n;
}

// frontier_evaluator.cpp:69-98
bool isFrontierVoxel(const orc_map& m, const orc_eval_params& e, const std::vector<V3>& nb,
                     const V3& voxel, uint64_t* lookups) {
  for (size_t i = 0; i < nb.size(); ++i) {
    *lookups += 8;
    const unsigned char s =
        m.getVoxelState(V3{voxel.x + nb[i].x, voxel.y + nb[i].y, voxel.z + nb[i].z});
    if (s == ORC_UNKNOWN) continue;
    if (e.surface_frontiers) return s == ORC_OCCUPIED;
    return true;
  }
  return false;
}

// computeGainFromVisibleVoxels: voxel_type_evaluator.cpp:57-89, naive_evaluator.cpp:20-38,
// frontier_evaluator.cpp:100-124.  Mutates the list like the reference does.
double foldGain(const orc_map& m, const orc_eval_params& e, std::vector<V3>* voxels,
                uint64_t* lookups, const V3& origin) {
  double gain = 0.0;
  if (e.kind == ORC_EVAL_VOXEL_WEIGHT) {
    // voxel_weight_evaluator.cpp:43-83 (the stored list is not modified)
    const std::vector<V3> nb = frontierOffsets(m, e);
    for (size_t i = 0; i < voxels->size(); ++i) {
      const V3& voxel = (*voxels)[i];
      const unsigned char s = m.getVoxelState(voxel);
      *lookups += 8;
      if (s == ORC_OCCUPIED) {
        const V3 dv{voxel.x - origin.x, voxel.y - origin.y, voxel.z - origin.z};
        const double z = std::sqrt((dv.x * dv.x + dv.y * dv.y) + dv.z * dv.z);
        const double spanned_angle = 2.0 * std::atan2(m.c_voxel_size, z * 2.0);
        const double new_weight =
            std::pow(spanned_angle, 2.0) / (e.ray_angle_x * e.ray_angle_y) / std::pow(z, 2.0);
        *lookups += 1;
        const double g = new_weight / (new_weight + m.getVoxelWeight(voxel));
        if (g > e.min_impact_factor) gain += g;
      } else if (s == ORC_UNKNOWN) {
        if (e.frontier_voxel_weight > 0.0 && isFrontierVoxel(m, e, nb, voxel, lookups)) {
          gain += e.frontier_voxel_weight;
        } else {
          gain += e.new_voxel_weight;
        }
      }
    }
  } else if (e.kind == ORC_EVAL_VOXEL_TYPE) {
    const double g_unknown = e.gains[0], g_occupied = e.gains[1], g_free = e.gains[2];
    const double g_unknown_o = e.gains[3], g_occupied_o = e.gains[4], g_free_o = e.gains[5];
    for (size_t i = 0; i < voxels->size(); ++i) {
      const unsigned char s = m.getVoxelState((*voxels)[i]);
      *lookups += 8;
      if (bvContains(e, (*voxels)[i])) {
        switch (s) {  // fall-through is the reference's behaviour (no breaks)
          case ORC_UNKNOWN:
            gain += g_unknown;  // fallthrough
          case ORC_FREE:
            gain += g_free;  // fallthrough
          case ORC_OCCUPIED:
            gain += g_occupied;
        }
      } else {
        switch (s) {
          case ORC_UNKNOWN:
            gain += g_unknown_o;  // fallthrough
          case ORC_FREE:
            gain += g_free_o;  // fallthrough
          case ORC_OCCUPIED:
            gain += g_occupied_o;
        }
      }
    }
  } else if (e.kind == ORC_EVAL_NAIVE) {
    // the reference erases one-by-one (O(n^2)); remove_if keeps the same survivors and order
    voxels->erase(std::remove_if(voxels->begin(), voxels->end(),
                                 [&](const V3& v) {
                                   *lookups += 1;
                                   return m.isObserved(v);
                                 }),
                  voxels->end());
    gain = static_cast<double>(voxels->size());
  } else {
    voxels->erase(std::remove_if(voxels->begin(), voxels->end(),
                                 [&](const V3& v) {
                                   *lookups += 1;
                                   return m.isObserved(v);
                                 }),
                  voxels->end());
    const std::vector<V3> nb = frontierOffsets(m, e);
    for (size_t i = 0; i < voxels->size(); ++i)
      if (isFrontierVoxel(m, e, nb, (*voxels)[i], lookups)) gain += 1.0;
  }
  return gain;
}

inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t bitsOf(double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return u;
}
inline uint64_t digestEntry(const V3& c) {
  uint64_t e = bitsOf(c.x) * 0x9E3779B97F4A7C15ull;
  e ^= rotl64(bitsOf(c.y), 23) * 0xC2B2AE3D27D4EB4Full;
  e ^= rotl64(bitsOf(c.z), 47) * 0x165667B19E3779F9ull;
  e ^= e >> 31;
  return e;
}
constexpr uint64_t kDigestMul = 0x100000001B3ull;
uint64_t digestList(const std::vector<V3>& v) {
  uint64_t h = 0;
  for (const V3& c : v) h = h * kDigestMul + digestEntry(c);
  return h;
}
// order-independent companion: sum of the per-entry hashes mod 2^64
uint64_t digestSet(const std::vector<V3>& v) {
  uint64_t h = 0;
  for (const V3& c : v) h += digestEntry(c);
  return h;
}

}  // namespace

extern "C" {

orc_map* orc_map_create(float voxel_size, int voxels_per_side) {
  orc_map* m = new orc_map();
  // voxblox: core/layer.h ctor, core/block.h ctor
  m->vs_f = voxel_size;
  m->vps = voxels_per_side;
  m->vs_inv_f = static_cast<float>(1.0 / m->vs_f);
  m->bs_f = m->vs_f * static_cast<float>(voxels_per_side);
  m->bs_inv_f = static_cast<float>(1.0 / m->bs_f);
  m->c_voxel_size = static_cast<double>(m->vs_f);
  m->c_block_size = static_cast<double>(m->bs_f);
  return m;
}
void orc_map_destroy(orc_map* m) { delete m; }

int orc_map_upload_blocks(orc_map* m, int n, const int32_t* idx, const float* dist,
                          const uint8_t* obs, const float* w) {
  const size_t nv = static_cast<size_t>(m->vps) * m->vps * m->vps;
  for (int i = 0; i < n; ++i) {
    const I3 b{idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]};
    auto blk = std::make_shared<Block>();
    blk->origin[0] = originPoint(b.x, m->bs_f);
    blk->origin[1] = originPoint(b.y, m->bs_f);
    blk->origin[2] = originPoint(b.z, m->bs_f);
    blk->dist.assign(dist + i * nv, dist + (i + 1) * nv);
    blk->observed.assign(obs + i * nv, obs + (i + 1) * nv);
    if (w) {
      blk->weight.assign(w + i * nv, w + (i + 1) * nv);
    } else {
      // an ESDF-only update leaves the TSDF layer alone (two voxblox layers): the block keeps its weights
      auto old = m->blocks.find(b);
      if (old != m->blocks.end()) blk->weight = old->second->weight;
    }
    m->blocks[b] = blk;
  }
  return 0;
}
int orc_map_remove_blocks(orc_map* m, int n, const int32_t* idx) {
  for (int i = 0; i < n; ++i) m->blocks.erase(I3{idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]});
  return 0;
}
int64_t orc_map_num_blocks(const orc_map* m) { return static_cast<int64_t>(m->blocks.size()); }
double orc_map_voxel_size(const orc_map* m) { return m->c_voxel_size; }

void orc_voxel_state(const orc_map* m, int64_t n, const double* p, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = m->getVoxelState(V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]});
}
void orc_voxel_center(const orc_map* m, int64_t n, const double* p, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    const V3 c = m->getVoxelCenter(V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]});
    out[3 * i] = c.x;
    out[3 * i + 1] = c.y;
    out[3 * i + 2] = c.z;
  }
}
void orc_is_observed(const orc_map* m, int64_t n, const double* p, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = m->isObserved(V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}) ? 1 : 0;
}
void orc_voxel_weight(const orc_map* m, int64_t n, const double* p, double* out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = m->getVoxelWeight(V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]});
}
void orc_distance(const orc_map* m, int64_t n, const double* p, double* dist, uint8_t* ok) {
  for (int64_t i = 0; i < n; ++i) {
    double d = 0.0;
    ok[i] = m->getDistanceAtPosition(V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}, &d) ? 1 : 0;
    dist[i] = d;
  }
}

const char* orc_impl(void) { return "port"; }

// vbx/src/map/voxblox.cpp:32-41
void orc_traversable(const orc_map* m, int64_t n, const double* p, double collision_radius, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    double d = 0.0;
    out[i] = (m->getDistanceAtPosition(V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}, &d) && d > collision_radius) ? 1 : 0;
  }
}

orc_caster* orc_caster_create(const orc_map* m, const orc_caster_params* p) {
  orc_caster* c = new orc_caster();
  c->map = m;
  c->p = *p;
  // camera_model.cpp:181-182
  c->fov_x = 2.0 * std::atan2(static_cast<double>(p->resolution_x), p->focal_length * 2.0);
  c->fov_y = 2.0 * std::atan2(static_cast<double>(p->resolution_y), p->focal_length * 2.0);
  if (p->kind == ORC_CASTER_ITERATIVE_LIDAR) {  // lidar_model.cpp:170-177
    c->fov_x = p->field_of_view_x;
    c->fov_y = p->field_of_view_y;
    c->fov_x *= M_PI / 180.0;
    c->fov_y *= M_PI / 180.0;
  }
  const double vs = m->c_voxel_size;
  c->ray_step = p->ray_step > 0.0 ? p->ray_step : vs;  // simple_ray_caster.cpp:17
  // simple_ray_caster.cpp:22-29 / iterative_ray_caster.cpp:25-32
  c->c_res_x = std::min(
      static_cast<int>(std::ceil(p->ray_length * c->fov_x / (vs * p->downsampling_factor))),
      p->resolution_x);
  c->c_res_y = std::min(
      static_cast<int>(std::ceil(p->ray_length * c->fov_y / (vs * p->downsampling_factor))),
      p->resolution_y);
  c->c_n_sections = 0;
  if (p->kind != ORC_CASTER_SIMPLE) {
    // iterative_ray_caster.cpp:35-45 == iterative_ray_caster_lidar.cpp:35-45
    c->c_n_sections = static_cast<int>(std::floor(std::log2(
        std::min(static_cast<double>(c->c_res_x), static_cast<double>(c->c_res_y)))));
    c->c_split_widths.push_back(0);
    for (int i = 0; i < c->c_n_sections; ++i) {
      c->c_split_widths.push_back(static_cast<int>(std::pow(2, i)));
      c->c_split_distances.push_back(p->ray_length / std::pow(2.0, static_cast<double>(i)));
    }
    c->c_split_distances.push_back(0.0);
    std::reverse(c->c_split_distances.begin(), c->c_split_distances.end());
    std::reverse(c->c_split_widths.begin(), c->c_split_widths.end());
  }
  return c;
}
void orc_caster_destroy(orc_caster* c) { delete c; }

void orc_caster_info(const orc_caster* c, int32_t* info, double* sd, int32_t* sw, double* fov,
                     double* ray_step) {
  info[0] = c->c_res_x;
  info[1] = c->c_res_y;
  info[2] = c->c_n_sections;
  for (size_t i = 0; i < c->c_split_distances.size() && i < 33; ++i) sd[i] = c->c_split_distances[i];
  for (size_t i = 0; i < c->c_split_widths.size() && i < 33; ++i) sw[i] = c->c_split_widths[i];
  if (fov) {
    fov[0] = c->fov_x;
    fov[1] = c->fov_y;
  }
  if (ray_step) *ray_step = c->ray_step;
}
void orc_caster_direction(const orc_caster* c, int i, int j, double* out3) {
  const V3 d = c->rayDirection(i, j);
  out3[0] = d.x;
  out3[1] = d.y;
  out3[2] = d.z;
}

// camera_model.cpp:62-83
int orc_sample_viewpoints(double sampling_time, int n_points, const int64_t* t_ns, int32_t* out) {
  int n = 0;
  if (sampling_time <= 0.0) {
    out[n++] = n_points - 1;
  } else {
    const int64_t sampling_time_ns = static_cast<int64_t>(sampling_time * 1.0e9);
    if (sampling_time_ns >= t_ns[n_points - 1]) {
      out[n++] = n_points - 1;
    } else {
      int64_t current_time = sampling_time_ns;
      for (int i = 0; i < n_points; ++i) {
        if (t_ns[i] >= current_time) {
          current_time += sampling_time_ns;
          out[n++] = i;
        }
      }
    }
  }
  return n;
}

int64_t orc_visible_voxels(const orc_caster* c, int n_views, const double* pos, const double* quat,
                           double* out, int64_t cap, int64_t* n_samples_out) {
  std::vector<V3> result;
  int64_t ns = 0;
  c->visibleFromPoses(&result, n_views, pos, quat, &ns);
  const int64_t n = static_cast<int64_t>(result.size());
  for (int64_t i = 0; i < std::min(n, cap); ++i) {
    out[3 * i] = result[i].x;
    out[3 * i + 1] = result[i].y;
    out[3 * i + 2] = result[i].z;
  }
  if (n_samples_out) *n_samples_out = ns;
  return n;
}

// simulated_sensor_evaluator.cpp:45-82 per segment
int orc_compute_gains(const orc_caster* c, const orc_eval_params* e, int n_views, const double* pos,
                      const double* quat, const int32_t* view_segment, int n_segments,
                      const int32_t* parent, double* gains_out, uint64_t* n_visible_out,
                      uint64_t* n_samples_out, uint64_t* digests_out, uint64_t* n_fold_out,
                      int n_threads, uint64_t* set_digests_out, const double* seg_origin) {
  // view ranges per segment
  std::vector<int> first(n_segments + 1, 0);
  for (int v = 0; v < n_views; ++v) {
    const int s = view_segment[v];
    if (s < 0 || s >= n_segments) return -1;
    if (v > 0 && s < view_segment[v - 1]) return -1;
    first[s + 1]++;
  }
  for (int s = 0; s < n_segments; ++s) first[s + 1] += first[s];

  const bool keep_lists = e->clear_from_parents && parent;
  std::vector<std::vector<V3>> stored(keep_lists ? n_segments : 0);

  auto do_segment = [&](int s) {
    std::vector<V3> new_voxels;
    int64_t ns = 0;
    const int v0 = first[s], nv = first[s + 1] - first[s];
    if (nv > 0) c->visibleFromPoses(&new_voxels, nv, pos + 3 * v0, quat + 4 * v0, &ns);
    if (e->bv_is_setup) {  // :50-57
      new_voxels.erase(std::remove_if(new_voxels.begin(), new_voxels.end(),
                                      [&](const V3& v) { return !bvContains(*e, v); }),
                       new_voxels.end());
    }
    if (keep_lists) {  // :60-76
      int prev = parent[s];
      while (prev >= 0) {
        const std::vector<V3>& old_voxels = stored[prev];
        new_voxels.erase(std::remove_if(new_voxels.begin(), new_voxels.end(),
                                        [&](const V3& v) {
                                          return std::find(old_voxels.begin(), old_voxels.end(),
                                                           v) != old_voxels.end();
                                        }),
                         new_voxels.end());
        prev = parent[prev];
      }
    }
    if (n_visible_out) n_visible_out[s] = new_voxels.size();
    if (n_samples_out) n_samples_out[s] = static_cast<uint64_t>(ns);
    if (digests_out) digests_out[s] = digestList(new_voxels);
    if (set_digests_out) set_digests_out[s] = digestSet(new_voxels);
    uint64_t lookups = 0;
    V3 origin{0, 0, 0};
    if (seg_origin) {
      origin = V3{seg_origin[3 * s], seg_origin[3 * s + 1], seg_origin[3 * s + 2]};
    } else if (nv > 0) {
      origin = V3{pos[3 * (v0 + nv - 1)], pos[3 * (v0 + nv - 1) + 1], pos[3 * (v0 + nv - 1) + 2]};
    }
    gains_out[s] = foldGain(*c->map, *e, &new_voxels, &lookups, origin);
    if (n_fold_out) n_fold_out[s] = lookups;
    if (keep_lists) stored[s] = std::move(new_voxels);
  };

  if (n_threads <= 1 || keep_lists) {
    for (int s = 0; s < n_segments; ++s) do_segment(s);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) {
      pool.emplace_back([&, t]() {
        for (int s = t; s < n_segments; s += n_threads) do_segment(s);
      });
    }
    for (auto& th : pool) th.join();
  }
  return 0;
}

double orc_gain_from_voxels(const orc_map* m, const orc_eval_params* e, int64_t n,
                            const double* centres, int64_t* n_left, const double* origin) {
  std::vector<V3> v(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) v[i] = V3{centres[3 * i], centres[3 * i + 1], centres[3 * i + 2]};
  uint64_t lookups = 0;
  const V3 o = origin ? V3{origin[0], origin[1], origin[2]} : V3{0, 0, 0};
  const double g = foldGain(*m, *e, &v, &lookups, o);
  if (n_left) *n_left = static_cast<int64_t>(v.size());
  return g;
}

void orc_bv_contains(const orc_eval_params* e, int64_t n, const double* p, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = bvContains(*e, V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}) ? 1 : 0;
}

uint64_t orc_digest_set(int64_t n, const double* c) {
  uint64_t h = 0;
  for (int64_t i = 0; i < n; ++i) h += digestEntry(V3{c[3 * i], c[3 * i + 1], c[3 * i + 2]});
  return h;
}

uint64_t orc_digest(int64_t n, const double* c) {
  uint64_t h = 0;
  for (int64_t i = 0; i < n; ++i)
    h = h * kDigestMul + digestEntry(V3{c[3 * i], c[3 * i + 1], c[3 * i + 2]});
  return h;
}

}  // extern "C"
