/*
 * ref_driver.cpp — oracle/_ref: the REFERENCE's own hot-path code behind the oracle's C API.
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product may link or load it.
 *
 * oracle/Makefile (target `_ref`) compiles this file together with the UNMODIFIED reference
 * sources where they lie under /root/reference (never copied into this repo):
 *   active_3d_planning_core/src/module/{module,module_factory,module_factory_registry}.cpp
 *   active_3d_planning_core/src/module/sensor_model/{sensor_model,camera_model,simple_ray_caster,
 *       iterative_ray_caster,lidar_model,iterative_ray_caster_lidar}.cpp
 *   active_3d_planning_core/src/module/trajectory_evaluator.cpp
 *   active_3d_planning_core/src/module/trajectory_evaluator/{simulated_sensor_evaluator,
 *       voxel_type_evaluator,naive_evaluator,frontier_evaluator,voxel_weight_evaluator}.cpp
 *   active_3d_planning_core/src/module/trajectory_evaluator/evaluator_updater/{
 *       simulated_sensor_updater,update_all,update_nothing}.cpp
 *   active_3d_planning_core/src/data/{bounding_volume,system_constraints,trajectory_segment,
 *       visualization_markers}.cpp, src/map/map.cpp
 *   active_3d_planning_voxblox/src/map/voxblox.cpp
 * against oracle/ref_shim/ (Eigen, glog, voxblox_ros stand-ins).  Everything the reference
 * decides — module registration and construction from string params, view sampling, mounting
 * transform, ray marching, the iterative ray table, std::unique, bounding-volume and
 * clear_from_parents filtering, the five gain folds, VoxbloxMap's state / centre / weight
 * functions — therefore runs as the reference wrote it.  What is NOT pinned by reference code is
 * listed in ref_shim/eigen_shim.h (floating-point order of six Eigen expressions) and
 * ref_shim/voxblox_shim.h (voxblox index helpers, nearest-voxel access, 8-corner interpolation).
 *
 * This file only (a) plays the planner that owns the map and the factory that feeds string
 * parameters, (b) observes — through subclasses that override nothing but
 * storeTrajectoryInformation — the list the reference stores in a segment, and (c) exports the
 * same orc_* entry points as a3d_oracle.h so tests can demand `_ref == oracle` bit for bit.
 */
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "active_3d_planning_core/data/bounding_volume.h"
#include "active_3d_planning_core/data/system_constraints.h"
#include "active_3d_planning_core/data/trajectory_segment.h"
#include "active_3d_planning_core/data/visualization_markers.h"
#include "active_3d_planning_core/module/module_factory.h"
#include "active_3d_planning_core/module/sensor_model/iterative_ray_caster.h"
#include "active_3d_planning_core/module/sensor_model/iterative_ray_caster_lidar.h"
#include "active_3d_planning_core/module/sensor_model/simple_ray_caster.h"
#include "active_3d_planning_core/module/trajectory_evaluator.h"
#include "active_3d_planning_core/module/trajectory_evaluator/evaluator_updater/simulated_sensor_updater.h"
#include "active_3d_planning_core/module/trajectory_evaluator/frontier_evaluator.h"
#include "active_3d_planning_core/module/trajectory_evaluator/naive_evaluator.h"
#include "active_3d_planning_core/module/trajectory_evaluator/simulated_sensor_evaluator.h"
#include "active_3d_planning_core/module/trajectory_evaluator/voxel_type_evaluator.h"
#include "active_3d_planning_core/module/trajectory_evaluator/voxel_weight_evaluator.h"
#include "active_3d_planning_core/planner/planner_I.h"
#include "active_3d_planning_voxblox/map/voxblox.h"

#include "a3d_oracle.h"

namespace a3d = active_3d_planning;
using a3d::Module;

namespace {

std::string num(double v) {  // round-trips through istringstream >> double (module.h:59-60)
  char b[64];
  std::snprintf(b, sizeof b, "%.17g", v);
  return b;
}
std::string num(int v) { return std::to_string(v); }

// ---- factory: namespaces of string params, like the ROS param server (module_factory_ros.cpp) --
class ParamFactory : public a3d::ModuleFactory {
 public:
  std::map<std::string, Module::ParamMap> ns;
  std::string last_error;
  bool getParamMapAndType(Module::ParamMap* map, std::string* type, std::string args) override {
    auto it = ns.find(args);
    if (it != ns.end()) {
      for (const auto& kv : it->second) (*map)[kv.first] = kv.second;
      auto t = it->second.find("type");
      if (t != it->second.end()) *type = t->second;
    }
    (*map)["param_namespace"] = args;
    return true;
  }

 protected:
  void printVerbose(const Module::ParamMap&) override {}
  void printError(const std::string& m) override { last_error = m; }
};

// ---- the planner the modules see -------------------------------------------------------------
class RefPlanner : public a3d::PlannerI {
 public:
  RefPlanner() : position_(0, 0, 0), orientation_(1, 0, 0, 0) {}
  const Eigen::Vector3d& getCurrentPosition() const override { return position_; }
  const Eigen::Quaterniond& getCurrentOrientation() const override { return orientation_; }
  a3d::BackTracker& getBackTracker() override { throw std::logic_error("no back tracker in oracle/_ref"); }
  a3d::TrajectoryGenerator& getTrajectoryGenerator() override { throw std::logic_error("no generator in oracle/_ref"); }
  a3d::TrajectoryEvaluator& getTrajectoryEvaluator() override { return *evaluator; }
  a3d::ModuleFactory& getFactory() override { return factory; }
  a3d::Map& getMap() override { return *map; }
  a3d::SystemConstraints& getSystemConstraints() override { return *constraints; }
  void publishVisualization(const a3d::VisualizationMarkers&) override {}
  void printInfo(const std::string&) override {}
  void printWarning(const std::string&) override {}
  void printError(const std::string& t) override { last_error = t; }

  ParamFactory factory;
  std::unique_ptr<a3d::SystemConstraints> constraints;
  std::unique_ptr<a3d::Map> map;
  std::unique_ptr<a3d::TrajectoryEvaluator> evaluator;
  std::string last_error;

 private:
  Eigen::Vector3d position_;
  Eigen::Quaterniond orientation_;
};

// ---- observers: reference classes with read access added, behaviour untouched -----------------
struct Captured {
  std::vector<Eigen::Vector3d> stored;  // the list handed to storeTrajectoryInformation
  voxblox::ShimCounters at_store;
  bool valid = false;
};
thread_local Captured* g_capture = nullptr;

template <class Base>
class TapEvaluator : public Base {
 public:
  explicit TapEvaluator(a3d::PlannerI& planner) : Base(planner) {}

 protected:
  bool storeTrajectoryInformation(a3d::TrajectorySegment* traj_in,
                                  const std::vector<Eigen::Vector3d>& new_voxels) override {
    if (g_capture) {
      g_capture->stored = new_voxels;
      g_capture->at_store = voxblox::ShimCounters::get();
      g_capture->valid = true;
    }
    return Base::storeTrajectoryInformation(traj_in, new_voxels);
  }
};

struct CasterInfo {
  int c_res_x = 0, c_res_y = 0, n_sections = 0;
  std::vector<double> split_distances;
  std::vector<int> split_widths;
  double fov_x = 0, fov_y = 0, ray_step = 0;
};
class CasterPeek {
 public:
  virtual ~CasterPeek() = default;
  virtual CasterInfo info() const = 0;
  virtual Eigen::Vector3d direction(int i, int j) = 0;
  virtual std::vector<int> sample(const a3d::TrajectorySegment& s) = 0;
};

class TapSimple : public a3d::sensor_model::SimpleRayCaster, public CasterPeek {
 public:
  explicit TapSimple(a3d::PlannerI& p) : SimpleRayCaster(p) {}
  CasterInfo info() const override {
    CasterInfo i;
    i.c_res_x = c_res_x_;
    i.c_res_y = c_res_y_;
    i.fov_x = c_field_of_view_x_;
    i.fov_y = c_field_of_view_y_;
    i.ray_step = p_ray_step_;
    return i;
  }
  Eigen::Vector3d direction(int i, int j) override {  // call site: simple_ray_caster.cpp:42-45
    Eigen::Vector3d d;
    CameraModel::getDirectionVector(&d, static_cast<double>(i) / (static_cast<double>(c_res_x_) - 1.0),
                                    static_cast<double>(j) / (static_cast<double>(c_res_y_) - 1.0));
    return d;
  }
  std::vector<int> sample(const a3d::TrajectorySegment& s) override {
    std::vector<int> r;
    sampleViewpoints(&r, s);
    return r;
  }
};
class TapIterative : public a3d::sensor_model::IterativeRayCaster, public CasterPeek {
 public:
  explicit TapIterative(a3d::PlannerI& p) : IterativeRayCaster(p) {}
  CasterInfo info() const override {
    CasterInfo i;
    i.c_res_x = c_res_x_;
    i.c_res_y = c_res_y_;
    i.n_sections = c_n_sections_;
    i.split_distances = c_split_distances_;
    i.split_widths = c_split_widths_;
    i.fov_x = c_field_of_view_x_;
    i.fov_y = c_field_of_view_y_;
    i.ray_step = p_ray_step_;
    return i;
  }
  Eigen::Vector3d direction(int i, int j) override {  // call site: iterative_ray_caster.cpp:62-66
    Eigen::Vector3d d;
    CameraModel::getDirectionVector(&d, static_cast<double>(i) / (static_cast<double>(c_res_x_) - 1.0),
                                    static_cast<double>(j) / (static_cast<double>(c_res_y_) - 1.0));
    return d;
  }
  std::vector<int> sample(const a3d::TrajectorySegment& s) override {
    std::vector<int> r;
    sampleViewpoints(&r, s);
    return r;
  }
};
class TapLidar : public a3d::sensor_model::IterativeRayCasterLidar, public CasterPeek {
 public:
  explicit TapLidar(a3d::PlannerI& p) : IterativeRayCasterLidar(p) {}
  CasterInfo info() const override {
    CasterInfo i;
    i.c_res_x = c_res_x_;
    i.c_res_y = c_res_y_;
    i.n_sections = c_n_sections_;
    i.split_distances = c_split_distances_;
    i.split_widths = c_split_widths_;
    i.fov_x = p_fov_x_;
    i.fov_y = p_fov_y_;
    i.ray_step = p_ray_step_;
    return i;
  }
  Eigen::Vector3d direction(int i, int j) override {
    Eigen::Vector3d d;
    LidarModel::getDirectionVector(&d, static_cast<double>(i) / (static_cast<double>(c_res_x_) - 1.0),
                                   static_cast<double>(j) / (static_cast<double>(c_res_y_) - 1.0));
    return d;
  }
  std::vector<int> sample(const a3d::TrajectorySegment& s) override {
    std::vector<int> r;
    sampleViewpoints(&r, s);
    return r;
  }
};

using a3d::ModuleFactoryRegistry;
namespace te = a3d::trajectory_evaluator;
ModuleFactoryRegistry::Registration<TapSimple> reg_tap_simple("Tap:SimpleRayCaster");
ModuleFactoryRegistry::Registration<TapIterative> reg_tap_iterative("Tap:IterativeRayCaster");
ModuleFactoryRegistry::Registration<TapLidar> reg_tap_lidar("Tap:IterativeRayCasterLidar");
ModuleFactoryRegistry::Registration<TapEvaluator<te::VoxelTypeEvaluator>> reg_tap_vt("Tap:VoxelTypeEvaluator");
ModuleFactoryRegistry::Registration<TapEvaluator<te::NaiveEvaluator>> reg_tap_naive("Tap:NaiveEvaluator");
ModuleFactoryRegistry::Registration<TapEvaluator<te::FrontierEvaluator>> reg_tap_frontier("Tap:FrontierEvaluator");
ModuleFactoryRegistry::Registration<TapEvaluator<te::VoxelWeightEvaluator>> reg_tap_vw("Tap:VoxelWeightEvaluator");

const char* casterType(int kind) {
  switch (kind) {
    case ORC_CASTER_SIMPLE: return "Tap:SimpleRayCaster";
    case ORC_CASTER_ITERATIVE: return "Tap:IterativeRayCaster";
    default: return "Tap:IterativeRayCasterLidar";
  }
}
const char* evalType(int kind) {
  switch (kind) {
    case ORC_EVAL_VOXEL_TYPE: return "Tap:VoxelTypeEvaluator";
    case ORC_EVAL_NAIVE: return "Tap:NaiveEvaluator";
    case ORC_EVAL_FRONTIER: return "Tap:FrontierEvaluator";
    default: return "Tap:VoxelWeightEvaluator";
  }
}

// The sampled body poses arrive one per view.  With sampling_time = 1 s and trajectory points at
// 1 s, 2 s, …, n s, CameraModel::sampleViewpoints (camera_model.cpp:62-83) selects every point;
// with one point it takes the "last point" branch.  A trailing point at (n + 0.5) s — never
// selected — carries the segment's end position when the caller gives one
// (voxel_weight_evaluator.cpp:53 reads trajectory.back()).
constexpr double kSamplingTime = 1.0;
constexpr int64_t kNs = 1000000000ll;

Module::ParamMap casterParamMap(const orc_caster_params& p) {
  Module::ParamMap m;
  m["type"] = casterType(p.kind);
  m["ray_length"] = num(p.ray_length);
  m["sampling_time"] = num(kSamplingTime);
  m["resolution_x"] = num(p.resolution_x);
  m["resolution_y"] = num(p.resolution_y);
  if (p.kind == ORC_CASTER_ITERATIVE_LIDAR) {
    m["field_of_view_x"] = num(p.field_of_view_x);
    m["field_of_view_y"] = num(p.field_of_view_y);
  } else {
    m["focal_length"] = num(p.focal_length);
  }
  if (p.ray_step > 0.0) m["ray_step"] = num(p.ray_step);  // else the reference's default (voxel size)
  m["downsampling_factor"] = num(p.downsampling_factor);
  m["mounting_translation_x"] = num(p.mount_t[0]);
  m["mounting_translation_y"] = num(p.mount_t[1]);
  m["mounting_translation_z"] = num(p.mount_t[2]);
  // sensor_model.cpp:27 normalises what it is given: hand it the caller's raw quaternion
  m["mounting_rotation_w"] = num(p.mount_q_raw[0]);
  m["mounting_rotation_x"] = num(p.mount_q_raw[1]);
  m["mounting_rotation_y"] = num(p.mount_q_raw[2]);
  m["mounting_rotation_z"] = num(p.mount_q_raw[3]);
  return m;
}

void evalParamMaps(const orc_eval_params& e, const std::string& ns, ParamFactory* f) {
  Module::ParamMap m;
  m["type"] = evalType(e.kind);
  m["clear_from_parents"] = e.clear_from_parents ? "1" : "0";
  m["accurate_frontiers"] = e.accurate_frontiers ? "1" : "0";
  m["surface_frontiers"] = e.surface_frontiers ? "1" : "0";
  m["checking_distance"] = num(e.checking_distance);
  m["gain_unknown"] = num(e.gains[0]);
  m["gain_occupied"] = num(e.gains[1]);
  m["gain_free"] = num(e.gains[2]);
  m["gain_unknown_outer"] = num(e.gains[3]);
  m["gain_occupied_outer"] = num(e.gains[4]);
  m["gain_free_outer"] = num(e.gains[5]);
  m["frontier_voxel_weight"] = num(e.frontier_voxel_weight);
  m["min_impact_factor"] = num(e.min_impact_factor);
  m["new_voxel_weight"] = num(e.new_voxel_weight);
  m["ray_angle_x"] = num(e.ray_angle_x);
  m["ray_angle_y"] = num(e.ray_angle_y);
  f->ns[ns] = m;
  Module::ParamMap bv;
  if (e.bv_is_setup) {
    bv["x_min"] = num(e.bv_min[0]);
    bv["y_min"] = num(e.bv_min[1]);
    bv["z_min"] = num(e.bv_min[2]);
    bv["x_max"] = num(e.bv_max[0]);
    bv["y_max"] = num(e.bv_max[1]);
    bv["z_max"] = num(e.bv_max[2]);
    bv["rotation"] = num(e.bv_rotation_deg);
  }
  f->ns[ns + "/bounding_volume"] = bv;
  f->ns[ns + "/outer_volume"] = Module::ParamMap();
}

std::mutex g_ros_params_mutex;  // voxblox::ShimRosParams stands in for the (global) ROS param server

struct World {  // one planner with its own VoxbloxMap; the reference is single-threaded per planner
  RefPlanner planner;
  a3d::map::VoxbloxMap* vmap = nullptr;

  World(float voxel_size, int vps) {
    planner.constraints = planner.factory.createModule<a3d::SystemConstraints>("/system_constraints", planner, false);
    std::lock_guard<std::mutex> lock(g_ros_params_mutex);
    voxblox::ShimRosParams::get().tsdf_voxel_size = voxel_size;
    voxblox::ShimRosParams::get().tsdf_voxels_per_side = vps;
    planner.factory.ns["/map"]["type"] = "VoxbloxMap";
    planner.map = planner.factory.createModule<a3d::Map>("/map", planner, false);
    vmap = dynamic_cast<a3d::map::VoxbloxMap*>(planner.map.get());
    if (!vmap) throw std::runtime_error("VoxbloxMap not registered");
  }
  voxblox::Layer<voxblox::EsdfVoxel>* esdf() { return vmap->getESDFServer().getEsdfMapPtr()->getEsdfLayerPtr(); }
  voxblox::Layer<voxblox::TsdfVoxel>* tsdf() { return vmap->getESDFServer().getTsdfMapPtr()->getTsdfLayerPtr(); }

  void upload(int n, const int32_t* idx, const float* dist, const uint8_t* obs, const float* w) {
    const size_t nv = esdf()->voxels_per_side() * esdf()->voxels_per_side() * esdf()->voxels_per_side();
    for (int i = 0; i < n; ++i) {
      const voxblox::BlockIndex b(idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]);
      // insert-or-overwrite, as a voxblox layer message with action "update" does; an ESDF-only update
      // (w == NULL) leaves the TSDF layer alone — they are two layers
      esdf()->removeBlock(b);
      if (w) tsdf()->removeBlock(b);
      auto eb = esdf()->allocateBlockPtrByIndex(b);
      for (size_t k = 0; k < nv; ++k) {
        voxblox::EsdfVoxel& v = eb->getVoxelByLinearIndex(k);
        v.distance = dist[i * nv + k];
        v.observed = obs[i * nv + k] != 0;
      }
      if (w) {
        auto tb = tsdf()->allocateBlockPtrByIndex(b);
        for (size_t k = 0; k < nv; ++k) tb->getVoxelByLinearIndex(k).weight = w[i * nv + k];
      }
    }
  }
  void remove(int n, const int32_t* idx) {
    for (int i = 0; i < n; ++i) {
      const voxblox::BlockIndex b(idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]);
      esdf()->removeBlock(b);
      tsdf()->removeBlock(b);
    }
  }
  void copyFrom(World& o, const std::set<std::array<int, 3>>& keys) {
    const size_t nv = esdf()->voxels_per_side() * esdf()->voxels_per_side() * esdf()->voxels_per_side();
    for (const auto& k : keys) {
      const voxblox::BlockIndex b(k[0], k[1], k[2]);
      auto src = o.esdf()->getBlockPtrByIndex(b);
      if (!src) continue;
      auto dst = esdf()->allocateBlockPtrByIndex(b);
      for (size_t i = 0; i < nv; ++i) dst->getVoxelByLinearIndex(i) = src->getVoxelByLinearIndex(i);
      auto ts = o.tsdf()->getBlockPtrByIndex(b);
      if (ts) {
        auto td = tsdf()->allocateBlockPtrByIndex(b);
        for (size_t i = 0; i < nv; ++i) td->getVoxelByLinearIndex(i) = ts->getVoxelByLinearIndex(i);
      }
    }
  }
};

void buildSegment(a3d::TrajectorySegment* seg, int nv, const double* pos, const double* quat, const double* end_pos) {
  seg->trajectory.clear();
  for (int v = 0; v < nv; ++v) {
    a3d::EigenTrajectoryPoint p;
    p.position_W = Eigen::Vector3d(pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]);
    p.orientation_W_B = Eigen::Quaterniond(quat[4 * v], quat[4 * v + 1], quat[4 * v + 2], quat[4 * v + 3]);
    p.time_from_start_ns = static_cast<int64_t>(v + 1) * kNs;
    seg->trajectory.push_back(p);
  }
  if (end_pos) {
    a3d::EigenTrajectoryPoint p;
    p.position_W = Eigen::Vector3d(end_pos[0], end_pos[1], end_pos[2]);
    p.time_from_start_ns = static_cast<int64_t>(nv) * kNs + kNs / 2;
    seg->trajectory.push_back(p);
  }
}

inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t bitsOf(double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return u;
}
// the repo's own list digest; same definition as a3d_oracle.cpp
inline uint64_t digestEntry(double x, double y, double z) {
  uint64_t e = bitsOf(x) * 0x9E3779B97F4A7C15ull;
  e ^= rotl64(bitsOf(y), 23) * 0xC2B2AE3D27D4EB4Full;
  e ^= rotl64(bitsOf(z), 47) * 0x165667B19E3779F9ull;
  e ^= e >> 31;
  return e;
}
constexpr uint64_t kDigestMul = 0x100000001B3ull;

}  // namespace

struct orc_map {
  float voxel_size;
  int vps;
  std::unique_ptr<World> world;
  std::set<std::array<int, 3>> keys;
  std::vector<std::unique_ptr<World>> replicas;  // for n_threads > 1; rebuilt after map edits
  bool replicas_stale = true;
};
struct orc_caster {
  orc_map* map;
  orc_caster_params p;
  std::unique_ptr<a3d::SensorModel> model;  // built by the reference factory in map->world
  CasterPeek* peek = nullptr;
};

extern "C" {

const char* orc_impl(void) { return "reference"; }

orc_map* orc_map_create(float voxel_size, int voxels_per_side) {
  orc_map* m = new orc_map();
  m->voxel_size = voxel_size;
  m->vps = voxels_per_side;
  m->world.reset(new World(voxel_size, voxels_per_side));
  return m;
}
void orc_map_destroy(orc_map* m) { delete m; }

int orc_map_upload_blocks(orc_map* m, int n, const int32_t* idx, const float* dist, const uint8_t* obs, const float* w) {
  m->world->upload(n, idx, dist, obs, w);
  for (int i = 0; i < n; ++i) m->keys.insert({idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]});
  m->replicas_stale = true;
  return 0;
}
int orc_map_remove_blocks(orc_map* m, int n, const int32_t* idx) {
  m->world->remove(n, idx);
  for (int i = 0; i < n; ++i) m->keys.erase({idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]});
  m->replicas_stale = true;
  return 0;
}
int64_t orc_map_num_blocks(const orc_map* m) {
  return static_cast<int64_t>(const_cast<orc_map*>(m)->world->esdf()->getNumberOfAllocatedBlocks());
}
double orc_map_voxel_size(const orc_map* m) { return m->world->vmap->getVoxelSize(); }

void orc_voxel_state(const orc_map* m, int64_t n, const double* p, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = m->world->vmap->getVoxelState(Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
void orc_voxel_center(const orc_map* m, int64_t n, const double* p, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector3d c;
    m->world->vmap->getVoxelCenter(&c, Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
    out[3 * i] = c.x();
    out[3 * i + 1] = c.y();
    out[3 * i + 2] = c.z();
  }
}
void orc_is_observed(const orc_map* m, int64_t n, const double* p, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = m->world->vmap->isObserved(Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2])) ? 1 : 0;
}
void orc_voxel_weight(const orc_map* m, int64_t n, const double* p, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = m->world->vmap->getVoxelWeight(Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
void orc_distance(const orc_map* m, int64_t n, const double* p, double* dist, uint8_t* ok) {
  for (int64_t i = 0; i < n; ++i) {
    double d = 0.0;
    ok[i] = m->world->vmap->getESDFServer().getEsdfMapPtr()->getDistanceAtPosition(Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]), &d) ? 1 : 0;
    dist[i] = d;
  }
}
/* VoxbloxMap::isTraversable (vbx/src/map/voxblox.cpp:32-41) with the given collision radius */
void orc_traversable(const orc_map* m, int64_t n, const double* p, double collision_radius, uint8_t* out) {
  m->world->planner.constraints->collision_radius = collision_radius;
  for (int64_t i = 0; i < n; ++i)
    out[i] = m->world->vmap->isTraversable(Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]), Eigen::Quaterniond(1, 0, 0, 0)) ? 1 : 0;
}

orc_caster* orc_caster_create(const orc_map* cm, const orc_caster_params* p) {
  orc_map* m = const_cast<orc_map*>(cm);
  orc_caster* c = new orc_caster();
  c->map = m;
  c->p = *p;
  m->world->planner.factory.ns["/caster"] = casterParamMap(*p);
  c->model = m->world->planner.factory.createModule<a3d::SensorModel>("/caster", m->world->planner, false);
  c->peek = dynamic_cast<CasterPeek*>(c->model.get());
  return c;
}
void orc_caster_destroy(orc_caster* c) { delete c; }

void orc_caster_info(const orc_caster* c, int32_t* info, double* sd, int32_t* sw, double* fov, double* ray_step) {
  const CasterInfo i = c->peek->info();
  info[0] = i.c_res_x;
  info[1] = i.c_res_y;
  info[2] = i.n_sections;
  for (size_t k = 0; k < i.split_distances.size() && k < 33; ++k) sd[k] = i.split_distances[k];
  for (size_t k = 0; k < i.split_widths.size() && k < 33; ++k) sw[k] = i.split_widths[k];
  if (fov) {
    fov[0] = i.fov_x;
    fov[1] = i.fov_y;
  }
  if (ray_step) *ray_step = i.ray_step;
}
void orc_caster_direction(const orc_caster* c, int i, int j, double* out3) {
  const Eigen::Vector3d d = c->peek->direction(i, j);
  out3[0] = d.x();
  out3[1] = d.y();
  out3[2] = d.z();
}

int orc_sample_viewpoints(double sampling_time, int n_points, const int64_t* t_ns, int32_t* out) {
  // a throw-away camera on a throw-away map: sampleViewpoints reads only p_sampling_time_
  World w(0.2f, 16);
  Module::ParamMap m;
  m["type"] = "Tap:SimpleRayCaster";
  m["sampling_time"] = num(sampling_time);
  w.planner.factory.ns["/caster"] = m;
  auto model = w.planner.factory.createModule<a3d::SensorModel>("/caster", w.planner, false);
  a3d::TrajectorySegment seg;
  for (int i = 0; i < n_points; ++i) {
    a3d::EigenTrajectoryPoint p;
    p.time_from_start_ns = t_ns[i];
    seg.trajectory.push_back(p);
  }
  const std::vector<int> r = dynamic_cast<CasterPeek*>(model.get())->sample(seg);
  for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
  return static_cast<int>(r.size());
}

int64_t orc_visible_voxels(const orc_caster* c, int n_views, const double* pos, const double* quat, double* out, int64_t cap,
                           int64_t* n_samples_out) {
  a3d::TrajectorySegment seg;
  buildSegment(&seg, n_views, pos, quat, nullptr);
  std::vector<Eigen::Vector3d> result;
  const voxblox::ShimCounters before = voxblox::ShimCounters::get();
  if (n_views > 0) c->model->getVisibleVoxelsFromTrajectory(&result, seg);
  if (n_samples_out) *n_samples_out = static_cast<int64_t>(voxblox::ShimCounters::get().interp - before.interp);
  const int64_t n = static_cast<int64_t>(result.size());
  for (int64_t i = 0; i < std::min(n, cap); ++i) {
    out[3 * i] = result[i].x();
    out[3 * i + 1] = result[i].y();
    out[3 * i + 2] = result[i].z();
  }
  return n;
}

int orc_compute_gains(const orc_caster* c, const orc_eval_params* e, int n_views, const double* pos, const double* quat,
                      const int32_t* view_segment, int n_segments, const int32_t* parent, double* gains_out,
                      uint64_t* n_visible_out, uint64_t* n_samples_out, uint64_t* digests_out, uint64_t* n_fold_out,
                      int n_threads, uint64_t* set_digests_out, const double* seg_origin) {
  std::vector<int> first(n_segments + 1, 0);
  for (int v = 0; v < n_views; ++v) {
    const int s = view_segment[v];
    if (s < 0 || s >= n_segments) return -1;
    if (v > 0 && s < view_segment[v - 1]) return -1;
    first[s + 1]++;
  }
  for (int s = 0; s < n_segments; ++s) first[s + 1] += first[s];
  const bool chained = e->clear_from_parents && parent;
  orc_map* m = c->map;

  // One evaluator (with its own sensor model and bounding volume) per world, built by the
  // reference's factory from string params: SimulatedSensorEvaluator::setupFromParamMap.
  auto make_eval = [&](World* w) {
    w->planner.factory.ns["/eval/sensor_model"] = casterParamMap(c->p);
    evalParamMaps(*e, "/eval", &w->planner.factory);
    return w->planner.factory.createModule<a3d::TrajectoryEvaluator>("/eval", w->planner, false);
  };
  // Segments live in a tree, as in the planner: children are spawned from their parent
  // (trajectory_segment.cpp spawnChild) so that `parent` pointers are what computeGain walks.
  std::vector<a3d::TrajectorySegment*> seg(n_segments, nullptr);
  std::vector<std::unique_ptr<a3d::TrajectorySegment>> roots;
  auto run_segment = [&](a3d::TrajectoryEvaluator* ev, int s) {
    const int v0 = first[s], nv = first[s + 1] - first[s];
    if (chained && parent[s] >= 0) {
      seg[s] = seg[parent[s]]->spawnChild();
    } else {
      roots.emplace_back(new a3d::TrajectorySegment());
      seg[s] = roots.back().get();
    }
    return std::make_pair(v0, nv);
  };
  auto eval_segment = [&](a3d::TrajectoryEvaluator* ev, a3d::TrajectorySegment* sg, int s, int v0, int nv) {
    const double* end = seg_origin ? seg_origin + 3 * s : nullptr;
    if (nv == 0) {  // no sampled pose: the oracle defines this as an empty list with gain 0
      gains_out[s] = 0.0;
      if (n_visible_out) n_visible_out[s] = 0;
      if (n_samples_out) n_samples_out[s] = 0;
      if (digests_out) digests_out[s] = 0;
      if (set_digests_out) set_digests_out[s] = 0;
      if (n_fold_out) n_fold_out[s] = 0;
      sg->info.reset(new te::SimulatedSensorInfo());
      return;
    }
    buildSegment(sg, nv, pos + 3 * v0, quat + 4 * v0, end);
    Captured cap;
    g_capture = &cap;
    const voxblox::ShimCounters before = voxblox::ShimCounters::get();
    ev->computeGain(sg);
    const voxblox::ShimCounters after = voxblox::ShimCounters::get();
    g_capture = nullptr;
    gains_out[s] = sg->gain;
    if (n_visible_out) n_visible_out[s] = cap.stored.size();
    if (n_samples_out) n_samples_out[s] = cap.at_store.interp - before.interp;
    if (digests_out || set_digests_out) {
      uint64_t h = 0, hs = 0;
      for (const auto& v : cap.stored) {
        const uint64_t d = digestEntry(v.x(), v.y(), v.z());
        h = h * kDigestMul + d;
        hs += d;
      }
      if (digests_out) digests_out[s] = h;
      if (set_digests_out) set_digests_out[s] = hs;
    }
    if (n_fold_out)
      n_fold_out[s] = 8 * (after.interp - cap.at_store.interp) + (after.observed - cap.at_store.observed) +
                      (after.weight - cap.at_store.weight);
  };

  if (n_threads <= 1 || chained) {
    auto ev = make_eval(m->world.get());
    for (int s = 0; s < n_segments; ++s) {
      auto r = run_segment(ev.get(), s);
      eval_segment(ev.get(), seg[s], s, r.first, r.second);
    }
    return 0;
  }
  // n_threads planners, each with a private copy of the map (the reference's modules are not
  // re-entrant: IterativeRayCaster::ray_table_ is a member)
  if (m->replicas_stale || static_cast<int>(m->replicas.size()) < n_threads - 1) {
    m->replicas.clear();
    for (int t = 1; t < n_threads; ++t) {
      m->replicas.emplace_back(new World(m->voxel_size, m->vps));
      m->replicas.back()->copyFrom(*m->world, m->keys);
    }
    m->replicas_stale = false;
  }
  for (int s = 0; s < n_segments; ++s) run_segment(nullptr, s);
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t) {
    pool.emplace_back([&, t]() {
      World* w = t == 0 ? m->world.get() : m->replicas[t - 1].get();
      auto ev = make_eval(w);
      for (int s = t; s < n_segments; s += n_threads) eval_segment(ev.get(), seg[s], s, first[s], first[s + 1] - first[s]);
    });
  }
  for (auto& th : pool) th.join();
  return 0;
}

double orc_gain_from_voxels(const orc_map* cm, const orc_eval_params* e, int64_t n, const double* centres, int64_t* n_left,
                            const double* origin) {
  // The reference's own re-fold path: SimulatedSensorUpdater::updateSegment
  // (simulated_sensor_updater.cpp:18-22) → evaluator->computeGainFromVisibleVoxels(segment).
  orc_map* m = const_cast<orc_map*>(cm);
  World* w = m->world.get();
  orc_caster_params cp;
  std::memset(&cp, 0, sizeof cp);
  cp.kind = ORC_CASTER_SIMPLE;
  cp.resolution_x = cp.resolution_y = 2;
  cp.ray_length = 1.0;
  cp.focal_length = 1.0;
  cp.downsampling_factor = 1.0;
  cp.mount_q_raw[0] = 1.0;
  w->planner.factory.ns["/eval/sensor_model"] = casterParamMap(cp);
  evalParamMaps(*e, "/eval", &w->planner.factory);
  auto ev = w->planner.factory.createModule<a3d::TrajectoryEvaluator>("/eval", w->planner, false);
  w->planner.factory.ns["/updater"]["type"] = "SimulatedSensorUpdater";
  auto up = w->planner.factory.createModule<a3d::EvaluatorUpdater>("/updater", w->planner, false);
  a3d::TrajectorySegment seg;
  a3d::EigenTrajectoryPoint p;
  if (origin) p.position_W = Eigen::Vector3d(origin[0], origin[1], origin[2]);
  seg.trajectory.push_back(p);
  te::SimulatedSensorInfo* info = new te::SimulatedSensorInfo();
  for (int64_t i = 0; i < n; ++i) info->visible_voxels.push_back(Eigen::Vector3d(centres[3 * i], centres[3 * i + 1], centres[3 * i + 2]));
  seg.info.reset(info);
  seg.gain = 0.0;
  up->updateSegment(&seg);
  if (n_left) *n_left = static_cast<int64_t>(info->visible_voxels.size());
  return seg.gain;
}

void orc_bv_contains(const orc_eval_params* e, int64_t n, const double* p, uint8_t* out) {
  RefPlanner planner;
  evalParamMaps(*e, "/eval", &planner.factory);
  auto bv = planner.factory.createModule<a3d::BoundingVolume>("/eval/bounding_volume", planner, false);
  for (int64_t i = 0; i < n; ++i) out[i] = bv->contains(Eigen::Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2])) ? 1 : 0;
}

uint64_t orc_digest_set(int64_t n, const double* c) {
  uint64_t h = 0;
  for (int64_t i = 0; i < n; ++i) h += digestEntry(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
  return h;
}
uint64_t orc_digest(int64_t n, const double* c) {
  uint64_t h = 0;
  for (int64_t i = 0; i < n; ++i) h = h * kDigestMul + digestEntry(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
  return h;
}

}  // extern "C"
