// a3d_host.cpp — C++ host side of the view-gain path: the reference's module classes as thin shells
// over the C ABI (include/a3d.h).  No map arithmetic, ray casting or gain folding happens here; the
// host code parses parameters, samples view points, moves buffers and keeps the reference's
// ownership / error conventions.  Reference citations are relative to the reference checkout.
#include <algorithm>
#include <map>
#include <cmath>
#include <limits>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <numeric>
#include <stdexcept>

#include "../../../include/a3d.h"
#include "a3d_host/trajectory_evaluator.h"
#include "a3d_host/yaml_factory.h"

namespace active_3d_planning {

[[noreturn]] void a3d_fatal(const std::string& message) {
  // the reference aborts through glog's LOG(FATAL); a throw lets embedding code decide
  throw std::runtime_error(message);
}

// ---- module system ------------------------------------------------------------------------------
void ModuleBase::assureParamsValid() {  // core/src/module/module.cpp:11-16
  std::string error_message;
  if (!checkParamsValid(&error_message)) {
    throw std::invalid_argument("Invalid parameters: " + error_message);
  }
}

ModuleFactoryRegistry& ModuleFactoryRegistry::Instance() {
  static ModuleFactoryRegistry instance;
  return instance;
}
ModuleFactoryRegistry::ModuleMap& ModuleFactoryRegistry::getRegistryMap() {
  static ModuleMap module_map;
  return module_map;
}
std::string ModuleFactoryRegistry::listTypes() {
  std::string out;
  for (const auto& kv : getRegistryMap()) out += (out.empty() ? "" : ", ") + kv.first;
  return out;
}
bool ModuleFactoryRegistry::isRegistered(const std::string& name) {
  return getRegistryMap().count(name) != 0;
}

void ModuleFactory::getDefaultType(const std::type_index& module_index, std::string* type) {
  // module_factory.cpp:14-25 — the reference's defaults for the module kinds of this path
  if (module_index == std::type_index(typeid(BoundingVolume))) *type = "BoundingVolume";
  if (module_index == std::type_index(typeid(EvaluatorUpdater))) *type = "UpdateNothing";
}
void ModuleFactory::registerLinkableModule(const std::string& name, Module* module) {
  linkable_module_list_[name] = module;  // module_factory.cpp:27-31
}
Module* ModuleFactory::readLinkableModule(const std::string& name) {
  auto it = linkable_module_list_.find(name);  // module_factory.cpp:33-41
  if (it == linkable_module_list_.end()) {
    printError("No module with name '" + name + "' registered to the linkable module list.");
    return nullptr;
  }
  return it->second;
}

// ---- data -----------------------------------------------------------------------------------------
TrajectorySegment* TrajectorySegment::spawnChild() {
  children.push_back(std::unique_ptr<TrajectorySegment>(new TrajectorySegment()));
  children.back()->parent = this;
  return children.back().get();
}
void TrajectorySegment::getChildren(std::vector<TrajectorySegment*>* result) {
  for (auto& c : children) result->push_back(c.get());
}
void TrajectorySegment::getLeaves(std::vector<TrajectorySegment*>* result) {
  if (children.empty()) {
    result->push_back(this);
    return;
  }
  for (auto& c : children) c->getLeaves(result);
}
void TrajectorySegment::getTree(std::vector<TrajectorySegment*>* result) {
  result->push_back(this);
  for (auto& c : children) c->getTree(result);
}
void TrajectorySegment::getTree(std::vector<TrajectorySegment*>* result, int maxdepth) {
  result->push_back(this);
  if (maxdepth > 0)
    for (auto& c : children) c->getTree(result, maxdepth - 1);
}
TrajectorySegment TrajectorySegment::shallowCopy() {
  TrajectorySegment copy;
  copy.trajectory = trajectory;
  copy.gain = gain;
  copy.cost = cost;
  copy.value = value;
  copy.tg_visited = tg_visited;
  copy.parent = parent;
  return copy;
}

ModuleFactoryRegistry::Registration<BoundingVolume> BoundingVolume::registration("BoundingVolume");

BoundingVolume::BoundingVolume(PlannerI& planner)
    : Module(planner), is_setup(false), x_min(0), x_max(0), y_min(0), y_max(0), z_min(0), z_max(0),
      rotation(0) {}

void BoundingVolume::setupFromParamMap(Module::ParamMap* param_map) {  // bounding_volume.cpp:12-31
  setParam<double>(param_map, "x_min", &x_min, 0.0);
  setParam<double>(param_map, "x_max", &x_max, 0.0);
  setParam<double>(param_map, "y_min", &y_min, 0.0);
  setParam<double>(param_map, "y_max", &y_max, 0.0);
  setParam<double>(param_map, "z_min", &z_min, 0.0);
  setParam<double>(param_map, "z_max", &z_max, 0.0);
  setParam<double>(param_map, "rotation", &rotation, 0.0);
  rotation_quat = Eigen::AngleAxisd(rotation * M_PI / -180.0, Eigen::Vector3d::UnitZ());
  is_setup = x_max > x_min && y_max > y_min && z_max > z_min;
}

bool BoundingVolume::contains(const Eigen::Vector3d& point) {  // bounding_volume.cpp:33-57
  // single points are host logic of the planner (RRT sampling etc.); the per-voxel tests of the
  // hot path run on the device (a3d_bv_contains / inside the cast kernels)
  if (!is_setup) return true;
  const Eigen::Vector3d r = rotation_quat * point;
  return !(r.x() < x_min || r.x() > x_max || r.y() < y_min || r.y() > y_max || r.z() < z_min ||
           r.z() > z_max);
}

// ---- map ------------------------------------------------------------------------------------------
namespace map {

ModuleFactoryRegistry::Registration<VoxbloxMap> VoxbloxMap::registration("VoxbloxMap");

VoxbloxMap::VoxbloxMap(PlannerI& planner)
    : TSDFMap(planner), ctx_(nullptr), group_(nullptr), p_voxel_size_(0.1), p_voxels_per_side_(16),
      c_voxel_size_(0.1), c_maximum_weight_(10000.0), p_device_(0) {}

VoxbloxMap::~VoxbloxMap() {
  if (group_)
    a3d_group_destroy(group_);  // owns the per-device contexts
  else if (ctx_)
    a3d_ctx_destroy(ctx_);
}
int VoxbloxMap::numDevices() const { return group_ ? a3d_group_size(group_) : (ctx_ ? 1 : 0); }

void VoxbloxMap::check(int rc, const char* what) {
  if (rc == A3D_OK) return;
  planner_.printError(std::string("VoxbloxMap: ") + what + ": " + a3d_last_error(ctx_));
  throw std::runtime_error(std::string(what) + ": " + a3d_last_error(ctx_));
}

void VoxbloxMap::setupFromParamMap(Module::ParamMap* param_map) {
  setParam<double>(param_map, "tsdf_voxel_size", &p_voxel_size_, 0.1);
  setParam<int>(param_map, "tsdf_voxels_per_side", &p_voxels_per_side_, 16);
  setParam<double>(param_map, "max_weight", &c_maximum_weight_, 10000.0);
  setParam<int>(param_map, "device", &p_device_, 0);
  setParam<std::string>(param_map, "devices", &p_devices_, "");
  if (!(p_voxel_size_ > 0.0) || p_voxels_per_side_ < 1) return;  // reported by checkParamsValid
  std::vector<int> ids;
  for (size_t at = 0; at < p_devices_.size();) {
    const size_t comma = p_devices_.find(',', at);
    const std::string tok = p_devices_.substr(at, comma == std::string::npos ? std::string::npos : comma - at);
    if (!tok.empty()) ids.push_back(std::stoi(tok));
    if (comma == std::string::npos) break;
    at = comma + 1;
  }
  if (ids.size() == 1) p_device_ = ids[0];
  // no CPU fallback: a map without a device cannot answer anything
  if (ids.size() > 1) {
    if (a3d_group_create(ids.data(), static_cast<int>(ids.size()), &group_) != A3D_OK)
      throw std::runtime_error(std::string("VoxbloxMap: no usable group of B200 devices: ") +
                               a3d_group_last_error(nullptr));
    ctx_ = a3d_group_ctx(group_, 0);
    if (a3d_group_map_config(group_, static_cast<float>(p_voxel_size_), p_voxels_per_side_) != A3D_OK)
      throw std::runtime_error(std::string("a3d_group_map_config: ") + a3d_group_last_error(group_));
  } else {
    if (a3d_ctx_create(p_device_, &ctx_) != A3D_OK)
      throw std::runtime_error(std::string("VoxbloxMap: no usable B200 device: ") + a3d_last_error(nullptr));
    check(a3d_map_config(ctx_, static_cast<float>(p_voxel_size_), p_voxels_per_side_), "a3d_map_config");
  }
  c_voxel_size_ = a3d_map_voxel_size(ctx_);  // double(float voxel size), voxblox.cpp:26
}

bool VoxbloxMap::checkParamsValid(std::string* error_message) {
  if (!(p_voxel_size_ > 0.0)) {
    *error_message = "tsdf_voxel_size expected > 0.0";
    return false;
  }
  if (p_voxels_per_side_ < 1 || p_voxels_per_side_ > 64) {
    *error_message = "tsdf_voxels_per_side expected in [1, 64]";
    return false;
  }
  return true;
}

bool VoxbloxMap::uploadBlocks(int n, const int32_t* block_idx, const float* esdf_distance,
                              const uint8_t* esdf_observed, const float* tsdf_weight) {
  if (group_) {  // host -> GPU 0 once, NCCL broadcast to the replicas, applied by every GPU
    if (a3d_group_map_upload_blocks(group_, n, block_idx, esdf_distance, esdf_observed, tsdf_weight) != A3D_OK)
      throw std::runtime_error(std::string("a3d_group_map_upload_blocks: ") + a3d_group_last_error(group_));
    return true;
  }
  check(a3d_map_upload_blocks(ctx_, n, block_idx, esdf_distance, esdf_observed, tsdf_weight),
        "a3d_map_upload_blocks");
  return true;
}
bool VoxbloxMap::uploadWeights(int n, const int32_t* block_idx, const float* tsdf_weight) {
  // every replica of a group holds the weights (the group has no weights-only broadcast: one plain
  // upload per device; TSDF updates are a few blocks per planner iteration)
  const int n_dev = numDevices();
  for (int i = 0; i < n_dev; ++i) {
    a3d_ctx* c = group_ ? a3d_group_ctx(group_, i) : ctx_;
    if (a3d_map_upload_weights(c, n, block_idx, tsdf_weight, nullptr) != A3D_OK)
      throw std::runtime_error(std::string("a3d_map_upload_weights: ") + a3d_last_error(c));
  }
  return true;
}
bool VoxbloxMap::uploadTsdfDistances(int n, const int32_t* block_idx, const float* tsdf_distance) {
  const int n_dev = numDevices();  // (queries go to the first device; the replicas stay interchangeable)
  for (int i = 0; i < n_dev; ++i) {
    a3d_ctx* c = group_ ? a3d_group_ctx(group_, i) : ctx_;
    if (a3d_map_upload_tsdf_distances(c, n, block_idx, tsdf_distance, nullptr) != A3D_OK)
      throw std::runtime_error(std::string("a3d_map_upload_tsdf_distances: ") + a3d_last_error(c));
  }
  return true;
}
bool VoxbloxMap::removeBlocks(int n, const int32_t* block_idx) {
  if (group_) {
    if (a3d_group_map_remove_blocks(group_, n, block_idx) != A3D_OK)
      throw std::runtime_error(std::string("a3d_group_map_remove_blocks: ") + a3d_group_last_error(group_));
    return true;
  }
  check(a3d_map_remove_blocks(ctx_, n, block_idx), "a3d_map_remove_blocks");
  return true;
}
int64_t VoxbloxMap::numBlocks() const { return a3d_map_num_blocks(ctx_); }

bool VoxbloxMap::isTraversable(const Eigen::Vector3d& position, const Eigen::Quaterniond&) {
  // voxblox.cpp:32-41: observed (interpolation succeeds) and farther than the collision radius
  double d = 0.0;
  uint8_t ok = 0;
  check(a3d_map_distance(ctx_, 1, position.data(), &d, &ok), "a3d_map_distance");
  return ok && d > planner_.getSystemConstraints().collision_radius;
}
bool VoxbloxMap::isObserved(const Eigen::Vector3d& point) {
  uint8_t o = 0;
  check(a3d_map_is_observed(ctx_, 1, point.data(), &o), "a3d_map_is_observed");
  return o != 0;
}
unsigned char VoxbloxMap::getVoxelState(const Eigen::Vector3d& point) {
  uint8_t s = UNKNOWN;
  check(a3d_map_voxel_state(ctx_, 1, point.data(), &s), "a3d_map_voxel_state");
  return s;
}
void VoxbloxMap::getVoxelStates(const std::vector<Eigen::Vector3d>& points,
                                std::vector<unsigned char>* states) {
  std::vector<double> flat(points.size() * 3);
  for (size_t i = 0; i < points.size(); ++i)
    for (int k = 0; k < 3; ++k) flat[3 * i + k] = points[i][k];
  states->assign(points.size(), UNKNOWN);
  check(a3d_map_voxel_state(ctx_, static_cast<int64_t>(points.size()), flat.data(), states->data()),
        "a3d_map_voxel_state");
}
void VoxbloxMap::areTraversable(const std::vector<Eigen::Vector3d>& points, std::vector<unsigned char>* out) {
  std::vector<double> flat(points.size() * 3);
  for (size_t i = 0; i < points.size(); ++i)
    for (int k = 0; k < 3; ++k) flat[3 * i + k] = points[i][k];
  out->assign(points.size(), 0);
  check(a3d_map_traversable(ctx_, static_cast<int64_t>(points.size()), flat.data(),
                            planner_.getSystemConstraints().collision_radius, out->data()),
        "a3d_map_traversable");
}
double VoxbloxMap::getVoxelSize() { return c_voxel_size_; }
bool VoxbloxMap::getVoxelCenter(Eigen::Vector3d* center, const Eigen::Vector3d& point) {
  double c[3];
  check(a3d_map_voxel_center(ctx_, 1, point.data(), c), "a3d_map_voxel_center");
  *center = Eigen::Vector3d(c[0], c[1], c[2]);
  return true;
}
double VoxbloxMap::getVoxelDistance(const Eigen::Vector3d& point) {
  // voxblox.cpp:84-98: the TSDF layer's stored distance of the voxel the point falls in (0 without one);
  // the mirror holds it for the blocks sent through uploadTsdfDistances
  double d = 0.0;
  check(a3d_map_voxel_tsdf_distance(ctx_, 1, point.data(), &d), "a3d_map_voxel_tsdf_distance");
  return d;
}
double VoxbloxMap::getVoxelWeight(const Eigen::Vector3d& point) {
  double w = 0.0;
  check(a3d_map_voxel_weight(ctx_, 1, point.data(), &w), "a3d_map_voxel_weight");
  return w;
}
double VoxbloxMap::getMaximumWeight() { return c_maximum_weight_; }

}  // namespace map

// ---- sensor models ----------------------------------------------------------------------------------
SensorModel::SensorModel(PlannerI& planner) : Module(planner), map_(nullptr) {}

void SensorModel::setupFromParamMap(Module::ParamMap* param_map) {  // sensor_model.cpp:9-27
  map_ = dynamic_cast<map::OccupancyMap*>(&(planner_.getMap()));
  if (!map_) planner_.printError("'SensorModel' requires a map of type 'OccupancyMap'!");
  double tx, ty, tz, rx, ry, rz, rw;
  setParam<double>(param_map, "mounting_translation_x", &tx, 0.0);
  setParam<double>(param_map, "mounting_translation_y", &ty, 0.0);
  setParam<double>(param_map, "mounting_translation_z", &tz, 0.0);
  setParam<double>(param_map, "mounting_rotation_x", &rx, 0.0);
  setParam<double>(param_map, "mounting_rotation_y", &ry, 0.0);
  setParam<double>(param_map, "mounting_rotation_z", &rz, 0.0);
  setParam<double>(param_map, "mounting_rotation_w", &rw, 1.0);
  mounting_translation_ = Eigen::Vector3d(tx, ty, tz);
  mounting_rotation_ = Eigen::Quaterniond(rw, rx, ry, rz).normalized();
}

namespace sensor_model {

CameraModel::CameraModel(PlannerI& planner)
    : SensorModel(planner), gpu_map_(nullptr), caster_(nullptr), caster_camera_(nullptr),
      group_caster_(nullptr), p_ray_length_(5.0), p_sampling_time_(0.0), p_focal_length_(320.0), p_ray_step_(0.0),
      p_downsampling_factor_(1.0), p_fov_x_deg_(360.0), p_fov_y_deg_(30.0), p_resolution_x_(640),
      p_resolution_y_(480), p_test_(false),
      c_field_of_view_x_(0.0), c_field_of_view_y_(0.0) {}

CameraModel::~CameraModel() {
  if (caster_) a3d_caster_destroy(caster_);
  if (caster_camera_) a3d_caster_destroy(caster_camera_);
  if (group_caster_) a3d_group_caster_destroy(group_caster_);
}

void CameraModel::setupFromParamMap(Module::ParamMap* param_map) {  // camera_model.cpp:170-183
  SensorModel::setupFromParamMap(param_map);
  setParam<double>(param_map, "ray_length", &p_ray_length_, 5.0);
  setParam<double>(param_map, "sampling_time", &p_sampling_time_, 0.0);
  setParam<double>(param_map, "focal_length", &p_focal_length_, 320.0);
  setParam<int>(param_map, "resolution_x", &p_resolution_x_, 640);
  setParam<int>(param_map, "resolution_y", &p_resolution_y_, 480);
  setParam<bool>(param_map, "test", &p_test_, false);
  c_field_of_view_x_ = 2.0 * atan2(p_resolution_x_, p_focal_length_ * 2.0);
  c_field_of_view_y_ = 2.0 * atan2(p_resolution_y_, p_focal_length_ * 2.0);
}

bool CameraModel::checkParamsValid(std::string* error_message) {  // camera_model.cpp:185-200
  if (p_ray_length_ <= 0.0) {
    *error_message = "ray_length expected > 0.0";
    return false;
  } else if (p_focal_length_ <= 0.0) {
    *error_message = "focal_length expected > 0.0";
    return false;
  } else if (p_resolution_x_ <= 0) {
    *error_message = "resolution_x expected > 0";
    return false;
  } else if (p_resolution_y_ <= 0) {
    *error_message = "resolution_y expected > 0";
    return false;
  } else if (!caster_) {
    *error_message = "ray caster could not be created on the device";
    return false;
  }
  return true;
}

void CameraModel::buildCasters(Module::ParamMap* param_map) {
  // ray_step / downsampling_factor and the derived ray grid (simple_ray_caster.cpp:15-30,
  // iterative_ray_caster.cpp:18-46); the tables are built inside a3d_caster_create
  setParam<double>(param_map, "ray_step", &p_ray_step_, map_ ? map_->getVoxelSize() : 0.0);
  setParam<double>(param_map, "downsampling_factor", &p_downsampling_factor_, 1.0);
  gpu_map_ = dynamic_cast<map::VoxbloxMap*>(map_);
  if (!gpu_map_) {
    planner_.printError("'CameraModel' (B200) requires the device map 'VoxbloxMap'!");
    return;
  }
  if (p_ray_length_ <= 0.0 || p_focal_length_ <= 0.0 || p_resolution_x_ <= 0 || p_resolution_y_ <= 0)
    return;  // reported by checkParamsValid
  a3d_caster_params p{};
  p.kind = casterKind();
  p.field_of_view_x = p_fov_x_deg_;
  p.field_of_view_y = p_fov_y_deg_;
  p.resolution_x = p_resolution_x_;
  p.resolution_y = p_resolution_y_;
  p.ray_length = p_ray_length_;
  p.focal_length = p_focal_length_;
  p.ray_step = p_ray_step_;
  p.downsampling_factor = p_downsampling_factor_;
  for (int k = 0; k < 3; ++k) p.mount_t[k] = mounting_translation_[k];
  p.mount_q[0] = mounting_rotation_.w();
  p.mount_q[1] = mounting_rotation_.x();
  p.mount_q[2] = mounting_rotation_.y();
  p.mount_q[3] = mounting_rotation_.z();
  if (a3d_caster_create(gpu_map_->context(), &p, &caster_) != A3D_OK) {
    planner_.printError(std::string("a3d_caster_create: ") + a3d_last_error(gpu_map_->context()));
    caster_ = nullptr;
    return;
  }
  if (gpu_map_->group() && a3d_group_caster_create(gpu_map_->group(), &p, &group_caster_) != A3D_OK) {
    planner_.printError(std::string("a3d_group_caster_create: ") + a3d_group_last_error(gpu_map_->group()));
    group_caster_ = nullptr;
  }
  for (int k = 0; k < 3; ++k) p.mount_t[k] = 0.0;
  p.mount_q[0] = 1.0;
  p.mount_q[1] = p.mount_q[2] = p.mount_q[3] = 0.0;
  if (a3d_caster_create(gpu_map_->context(), &p, &caster_camera_) != A3D_OK) caster_camera_ = nullptr;
}

int CameraModel::rayGridX() const {
  a3d_caster_info i{};
  return caster_ && a3d_caster_get_info(caster_, &i) == A3D_OK ? i.c_res_x : 0;
}
int CameraModel::rayGridY() const {
  a3d_caster_info i{};
  return caster_ && a3d_caster_get_info(caster_, &i) == A3D_OK ? i.c_res_y : 0;
}
int CameraModel::numSections() const {
  a3d_caster_info i{};
  return caster_ && a3d_caster_get_info(caster_, &i) == A3D_OK ? i.c_n_sections : 0;
}

void CameraModel::sampleViewpoints(std::vector<int>* result, const TrajectorySegment& traj_in) {
  // camera_model.cpp:62-83 (host logic on the trajectory's time stamps)
  const int n = static_cast<int>(traj_in.trajectory.size());
  if (p_sampling_time_ <= 0.0) {
    result->push_back(n - 1);
    return;
  }
  const int64_t sampling_time_ns = static_cast<int64_t>(p_sampling_time_ * 1.0e9);
  if (sampling_time_ns >= traj_in.trajectory.back().time_from_start_ns) {
    result->push_back(n - 1);
    return;
  }
  int64_t current_time = sampling_time_ns;
  for (int i = 0; i < n; ++i) {
    if (traj_in.trajectory[i].time_from_start_ns >= current_time) {
      current_time += sampling_time_ns;
      result->push_back(i);
    }
  }
}

void CameraModel::noteCastTime(double device_ms, int n_views) {
  if (!p_test_ || n_views <= 0 || !(device_ms >= 0.0)) return;
  const size_t before = time_count_.size();
  time_count_.insert(time_count_.end(), static_cast<size_t>(n_views), device_ms / n_views);
  if (time_count_.size() / 20 == before / 20) return;  // the reference logs at every 20th view
  double mean, stddev;
  size_t count;
  castTimeStats(&mean, &stddev, &count);
  planner_.printInfo("Raycasting: Mean: " + std::to_string(mean) + " ms, Stddev: " + std::to_string(stddev) +
                     " ms");
}
void CameraModel::castTimeStats(double* mean, double* stddev, size_t* count) const {
  const size_t n = time_count_.size();
  double m = 0.0, v = 0.0;
  for (double t : time_count_) m += t;
  if (n) m /= static_cast<double>(n);
  for (double t : time_count_) v += (t - m) * (t - m);
  *mean = m;
  *stddev = n ? std::sqrt(v / static_cast<double>(n)) : 0.0;
  *count = n;
}

void CameraModel::getDirectionVector(Eigen::Vector3d* result, double relative_x, double relative_y) {
  *result = Eigen::Vector3d(p_focal_length_, (0.5 - relative_x) * static_cast<double>(p_resolution_x_),
                            (0.5 - relative_y) * static_cast<double>(p_resolution_y_))
                .normalized();  // camera_model.cpp:161-168
}

namespace detail {
// poses of the sampled trajectory points, flat
void gatherPoses(const TrajectorySegment& traj, const std::vector<int>& indices, std::vector<double>* pos,
                 std::vector<double>* quat) {
  for (int idx : indices) {
    const EigenTrajectoryPoint& p = traj.trajectory[idx];
    for (int k = 0; k < 3; ++k) pos->push_back(p.position_W[k]);
    quat->push_back(p.orientation_W_B.w());
    quat->push_back(p.orientation_W_B.x());
    quat->push_back(p.orientation_W_B.y());
    quat->push_back(p.orientation_W_B.z());
  }
}

bool runVisible(a3d_ctx* ctx, a3d_caster* caster, a3d_eval* eval, int n_views, const double* pos,
                const double* quat, std::vector<Eigen::Vector3d>* out, double* gain, PlannerI& planner,
                const double* origin = nullptr) {
  a3d_caster_info info{};
  a3d_caster_get_info(caster, &info);
  int64_t cap = static_cast<int64_t>(std::max(1, n_views)) * info.c_res_x * info.c_res_y *
                std::max(1, info.samples_per_full_ray);
  // (kept between calls: a fresh vector of the worst-case size would be zero-filled — 6 MB for the
  // 320x240 camera — on every per-segment call)
  static thread_local std::vector<double> flat;
  if (flat.size() < static_cast<size_t>(cap) * 3) flat.resize(static_cast<size_t>(cap) * 3);
  int64_t n = 0, n_samples = 0;
  const int rc = a3d_visible_voxels(ctx, caster, eval, n_views, pos, quat, flat.data(), cap, &n,
                                    &n_samples, gain, origin);
  if (rc != A3D_OK) {
    planner.printError(std::string("a3d_visible_voxels: ") + a3d_last_error(ctx));
    return false;
  }
  static_assert(sizeof(Eigen::Vector3d) == 3 * sizeof(double), "Vector3d is three packed doubles");
  const Eigen::Vector3d* first = reinterpret_cast<const Eigen::Vector3d*>(flat.data());
  out->insert(out->end(), first, first + n);
  return true;
}
}  // namespace detail
using detail::gatherPoses;
using detail::runVisible;

bool CameraModel::getVisibleVoxelsFromTrajectory(std::vector<Eigen::Vector3d>* result,
                                                 const TrajectorySegment& traj_in) {
  // camera_model.cpp:15-60: sample poses, cast every pose, adjacent-unique over the whole list —
  // one device call (mounting transform and unique are applied inside)
  if (!caster_ || traj_in.trajectory.empty()) return false;
  std::vector<int> indices;
  sampleViewpoints(&indices, traj_in);
  std::vector<double> pos, quat;
  gatherPoses(traj_in, indices, &pos, &quat);
  const size_t before = result->size();
  if (!runVisible(gpu_map_->context(), caster_, nullptr, static_cast<int>(indices.size()), pos.data(),
                  quat.data(), result, nullptr, planner_))
    return false;
  noteCastTime(a3d_last_kernel_ms(gpu_map_->context()), static_cast<int>(indices.size()));
  // entries appended to a non-empty *result still get the reference's whole-vector unique
  if (before > 0) result->erase(std::unique(result->begin(), result->end()), result->end());
  return true;
}

bool CameraModel::getVisibleVoxels(std::vector<Eigen::Vector3d>* result, const Eigen::Vector3d& position,
                                   const Eigen::Quaterniond& orientation) {
  if (!caster_camera_) return false;
  const double quat[4] = {orientation.w(), orientation.x(), orientation.y(), orientation.z()};
  const bool ok = runVisible(gpu_map_->context(), caster_camera_, nullptr, 1, position.data(), quat, result,
                             nullptr, planner_);
  if (ok) noteCastTime(a3d_last_kernel_ms(gpu_map_->context()), 1);
  return ok;
}

ModuleFactoryRegistry::Registration<SimpleRayCaster> SimpleRayCaster::registration("SimpleRayCaster");
void SimpleRayCaster::setupFromParamMap(Module::ParamMap* param_map) {
  CameraModel::setupFromParamMap(param_map);
  buildCasters(param_map);
}
int SimpleRayCaster::casterKind() const { return A3D_CASTER_SIMPLE; }

ModuleFactoryRegistry::Registration<IterativeRayCaster> IterativeRayCaster::registration(
    "IterativeRayCaster");
void IterativeRayCaster::setupFromParamMap(Module::ParamMap* param_map) {
  CameraModel::setupFromParamMap(param_map);
  buildCasters(param_map);
}
int IterativeRayCaster::casterKind() const { return A3D_CASTER_ITERATIVE; }

ModuleFactoryRegistry::Registration<IterativeRayCasterLidar> IterativeRayCasterLidar::registration(
    "IterativeRayCasterLidar");
void IterativeRayCasterLidar::setupFromParamMap(Module::ParamMap* param_map) {  // lidar_model.cpp:167-178
  SensorModel::setupFromParamMap(param_map);
  setParam<double>(param_map, "ray_length", &p_ray_length_, 5.0);
  setParam<double>(param_map, "sampling_time", &p_sampling_time_, 0.0);
  setParam<double>(param_map, "field_of_view_x", &p_fov_x_deg_, 360.0);
  setParam<double>(param_map, "field_of_view_y", &p_fov_y_deg_, 30.0);
  setParam<int>(param_map, "resolution_x", &p_resolution_x_, 1024);
  setParam<int>(param_map, "resolution_y", &p_resolution_y_, 128);
  setParam<bool>(param_map, "test", &p_test_, false);
  c_field_of_view_x_ = p_fov_x_deg_ * M_PI / 180.0;
  c_field_of_view_y_ = p_fov_y_deg_ * M_PI / 180.0;
  p_focal_length_ = 1.0;  // unused by the lidar model
  buildCasters(param_map);
}
bool IterativeRayCasterLidar::checkParamsValid(std::string* error_message) {  // lidar_model.cpp:180-198
  if (p_fov_x_deg_ <= 0.0) {
    *error_message = "field_of_view_x expected > 0.0";
    return false;
  } else if (p_fov_y_deg_ <= 0.0) {
    *error_message = "field_of_view_y expected > 0.0";
    return false;
  }
  return CameraModel::checkParamsValid(error_message);
}
int IterativeRayCasterLidar::casterKind() const { return A3D_CASTER_ITERATIVE_LIDAR; }

}  // namespace sensor_model

// ---- evaluators ---------------------------------------------------------------------------------------
TrajectoryEvaluator::TrajectoryEvaluator(PlannerI& planner) : Module(planner) {}

void TrajectoryEvaluator::setupFromParamMap(Module::ParamMap* param_map) {
  // trajectory_evaluator.cpp:12-28: child modules default to namespace extensions
  const std::string ns = (*param_map)["param_namespace"];
  setParam<std::string>(param_map, "cost_computer_args", &p_cost_args_, ns + "/cost_computer");
  setParam<std::string>(param_map, "value_computer_args", &p_value_args_, ns + "/value_computer");
  setParam<std::string>(param_map, "next_selector_args", &p_next_args_, ns + "/next_selector");
  setParam<std::string>(param_map, "evaluator_updater_args", &p_updater_args_, ns + "/evaluator_updater");
  std::string bounding_volume_args;
  setParam<std::string>(param_map, "bounding_volume_args", &bounding_volume_args, ns + "/bounding_volume");
  bounding_volume_ =
      planner_.getFactory().createModule<BoundingVolume>(bounding_volume_args, planner_, verbose_modules_);
}
bool TrajectoryEvaluator::computeCost(TrajectorySegment* traj_in) {
  if (!cost_computer_)
    cost_computer_ = planner_.getFactory().createModule<CostComputer>(p_cost_args_, planner_, verbose_modules_);
  return cost_computer_->computeCost(traj_in);
}
bool TrajectoryEvaluator::computeValue(TrajectorySegment* traj_in) {
  if (!value_computer_)
    value_computer_ =
        planner_.getFactory().createModule<ValueComputer>(p_value_args_, planner_, verbose_modules_);
  return value_computer_->computeValue(traj_in);
}
int TrajectoryEvaluator::selectNextBest(TrajectorySegment* traj_in) {
  if (!next_selector_)
    next_selector_ = planner_.getFactory().createModule<NextSelector>(p_next_args_, planner_, verbose_modules_);
  return next_selector_->selectNextBest(traj_in);
}
EvaluatorUpdater* TrajectoryEvaluator::updater() {
  if (!evaluator_updater_)
    evaluator_updater_ =
        planner_.getFactory().createModule<EvaluatorUpdater>(p_updater_args_, planner_, verbose_modules_);
  return evaluator_updater_.get();
}
bool TrajectoryEvaluator::updateSegment(TrajectorySegment* segment) { return updater()->updateSegment(segment); }
void TrajectoryEvaluator::updateSegments(const std::vector<TrajectorySegment*>& segments,
                                         std::vector<uint8_t>* alive) {
  updater()->updateSegments(segments, alive);
}
int TrajectoryEvaluator::updateTree(TrajectorySegment* root) {
  // every segment below the root, parents before children
  std::vector<TrajectorySegment*> all;
  for (auto& c : root->children) all.push_back(c.get());
  for (size_t at = 0; at < all.size(); ++at)
    for (auto& c : all[at]->children) all.push_back(c.get());
  if (all.empty()) return 0;
  std::vector<uint8_t> alive(all.size(), 1);
  updateSegments(all, &alive);
  // a dropped segment takes its subtree with it (the reference never descends into it)
  std::map<const TrajectorySegment*, size_t> at;
  for (size_t i = 0; i < all.size(); ++i) at[all[i]] = i;
  int removed = 0;
  for (size_t i = 0; i < all.size(); ++i) {
    const TrajectorySegment* up = all[i]->parent;
    if (up != root && !alive[at[up]]) alive[i] = 0;
    removed += alive[i] ? 0 : 1;
  }
  // erase from the deepest level upwards so that no erased pointer is looked at again
  for (size_t i = all.size(); i-- > 0;) {
    if (alive[i]) continue;
    auto& siblings = all[i]->parent->children;
    siblings.erase(std::find_if(siblings.begin(), siblings.end(),
                                [&](const std::unique_ptr<TrajectorySegment>& c) { return c.get() == all[i]; }));
  }
  return removed;
}
bool EvaluatorUpdater::updateSegment(TrajectorySegment* segment) {
  std::vector<uint8_t> alive(1, 1);
  updateSegments({segment}, &alive);
  return alive[0] != 0;
}

namespace trajectory_evaluator {
using sensor_model::detail::gatherPoses;
using sensor_model::detail::runVisible;

SimulatedSensorEvaluator::SimulatedSensorEvaluator(PlannerI& planner)
    : TrajectoryEvaluator(planner), map_(nullptr), gpu_map_(nullptr), eval_(nullptr), group_eval_(nullptr),
      p_clear_from_parents_(false), p_visualize_sensor_view_(false) {}

DeviceListStore::~DeviceListStore() {
  if (store) a3d_store_destroy(store);
}
DeviceSensorInfo::~DeviceSensorInfo() {
  if (auto s = store.lock()) a3d_store_erase(s->store, 1, &key);
}

SimulatedSensorEvaluator::~SimulatedSensorEvaluator() {
  device_lists_.reset();  // before the evaluator object; the map (context) outlives the evaluator
  if (eval_) a3d_eval_destroy(eval_);
  if (group_eval_) a3d_group_eval_destroy(group_eval_);
}

double SimulatedSensorEvaluator::takeGainTimeMs() {
  const double t = perf_gain_ms_;
  perf_gain_ms_ = 0.0;
  return t;
}
void SimulatedSensorEvaluator::noteDeviceTime(int n_views, bool group) {
  const double ms = group ? a3d_group_last_kernel_ms(gpu_map_->group()) : a3d_last_kernel_ms(gpu_map_->context());
  if (!(ms >= 0.0)) return;
  perf_gain_ms_ += ms;
  if (auto* cam = dynamic_cast<sensor_model::CameraModel*>(sensor_model_.get())) cam->noteCastTime(ms, n_views);
}

bool SimulatedSensorEvaluator::computeGainsStored(const std::vector<TrajectorySegment*>& segments) {
  auto* cam = dynamic_cast<sensor_model::CameraModel*>(sensor_model_.get());
  if (!cam || !cam->handle() || !eval_) return false;
  if (!device_lists_) {
    a3d_store* st = nullptr;
    if (a3d_store_create(gpu_map_->context(), &st) != A3D_OK) return false;
    device_lists_ = std::make_shared<DeviceListStore>(st);
  }
  std::vector<double> pos, quat, origins(segments.size() * 3, 0.0);
  std::vector<int32_t> view_segment;
  std::vector<uint64_t> keys(segments.size()), n_vis(segments.size()), parent_keys(segments.size(), 0);
  std::map<const TrajectorySegment*, uint64_t> in_batch;
  for (size_t s = 0; s < segments.size(); ++s) {
    keys[s] = next_key_++;
    in_batch[segments[s]] = keys[s];
  }
  for (size_t s = 0; s < segments.size(); ++s) {
    if (p_clear_from_parents_) {
      // simulated_sensor_evaluator.cpp:60-76 walks traj_in->parent upwards; on the device the chain is
      // key → parent key.  The nearest ancestor that holds a list starts it (it links on to its own).
      for (const TrajectorySegment* a = segments[s]->parent; a; a = a->parent) {
        auto hit = in_batch.find(a);
        if (hit != in_batch.end()) {
          parent_keys[s] = hit->second;
          break;
        }
        if (auto* d = dynamic_cast<DeviceSensorInfo*>(a->info.get())) {
          if (d->store.lock() != device_lists_) return false;
          parent_keys[s] = d->key;
          break;
        }
        if (dynamic_cast<SimulatedSensorInfo*>(a->info.get())) return false;  // host-side list: per-segment path
      }
    }
    if (segments[s]->trajectory.empty()) continue;
    std::vector<int> indices;
    cam->sampleViewpoints(&indices, *segments[s]);
    gatherPoses(*segments[s], indices, &pos, &quat);
    view_segment.insert(view_segment.end(), indices.size(), static_cast<int32_t>(s));
    for (int k = 0; k < 3; ++k) origins[3 * s + k] = segments[s]->trajectory.back().position_W[k];
  }
  std::vector<double> gains(segments.size(), 0.0);
  if (a3d_compute_gains_stored(gpu_map_->context(), cam->handle(), eval_, device_lists_->store,
                               static_cast<int>(view_segment.size()), pos.data(), quat.data(),
                               view_segment.data(), static_cast<int>(segments.size()), keys.data(),
                               p_clear_from_parents_ ? parent_keys.data() : nullptr, gains.data(), n_vis.data(),
                               nullptr, nullptr, nullptr, origins.data()) != A3D_OK) {
    planner_.printError(std::string("a3d_compute_gains_stored: ") + a3d_last_error(gpu_map_->context()));
    return false;
  }
  noteDeviceTime(static_cast<int>(view_segment.size()), false);
  for (size_t s = 0; s < segments.size(); ++s) {
    auto* info = new DeviceSensorInfo();
    info->key = keys[s];
    info->n_voxels = n_vis[s];
    info->store = device_lists_;
    segments[s]->info.reset(info);  // releases whatever list the segment held before
    segments[s]->gain = gains[s];
  }
  return true;
}

bool SimulatedSensorEvaluator::updateGains(const std::vector<TrajectorySegment*>& segments) {
  if (!eval_) return false;
  std::vector<TrajectorySegment*> on_device;
  std::vector<uint64_t> keys;
  std::vector<double> origins;
  bool ok = true;
  for (TrajectorySegment* seg : segments) {
    auto* dinfo = dynamic_cast<DeviceSensorInfo*>(seg->info.get());
    if (dinfo && device_lists_ && dinfo->store.lock() == device_lists_) {
      on_device.push_back(seg);
      keys.push_back(dinfo->key);
      for (int k = 0; k < 3; ++k)
        origins.push_back(seg->trajectory.empty() ? 0.0 : seg->trajectory.back().position_W[k]);
    } else {
      ok = computeGainFromVisibleVoxels(seg) && ok;  // host-side list
    }
  }
  if (on_device.empty()) return ok;
  std::vector<double> gains(on_device.size());
  std::vector<uint64_t> left(on_device.size());
  if (a3d_store_refold(gpu_map_->context(), eval_, device_lists_->store, static_cast<int>(keys.size()),
                       keys.data(), gains.data(), left.data(), origins.data()) != A3D_OK) {
    planner_.printError(std::string("a3d_store_refold: ") + a3d_last_error(gpu_map_->context()));
    return false;
  }
  for (size_t i = 0; i < on_device.size(); ++i) {
    on_device[i]->gain = gains[i];
    static_cast<DeviceSensorInfo*>(on_device[i]->info.get())->n_voxels = left[i];
  }
  return ok;
}

void SimulatedSensorEvaluator::setupFromParamMap(Module::ParamMap* param_map) {
  // simulated_sensor_evaluator.cpp:16-43
  setParam<bool>(param_map, "clear_from_parents", &p_clear_from_parents_, false);
  setParam<bool>(param_map, "visualize_sensor_view", &p_visualize_sensor_view_, false);
  map_ = dynamic_cast<map::OccupancyMap*>(&(planner_.getMap()));
  if (!map_) planner_.printError("'SimulatedSensorEvaluator' requires a map of type 'OccupancyMap'!");
  gpu_map_ = dynamic_cast<map::VoxbloxMap*>(map_);
  planner_.getFactory().registerLinkableModule("SimulatedSensorEvaluator", this);
  std::string args;
  const std::string param_ns = (*param_map)["param_namespace"];
  setParam<std::string>(param_map, "sensor_model_args", &args, param_ns + "/sensor_model");
  sensor_model_ = planner_.getFactory().createModule<SensorModel>(args, planner_, verbose_modules_);
  TrajectoryEvaluator::setupFromParamMap(param_map);
}

void SimulatedSensorEvaluator::buildEval() {
  if (!gpu_map_) {
    planner_.printError("'SimulatedSensorEvaluator' (B200) requires the device map 'VoxbloxMap'!");
    return;
  }
  a3d_eval_params p{};
  p.kind = evalKind();
  // read by a3d_compute_gains_stored (computeGainsStored: ancestors' lists are subtracted on the
  // device); the per-segment computeGain below filters against host-side lists with a3d_filter_not_in
  p.clear_from_parents = p_clear_from_parents_ ? 1 : 0;
  p.surface_frontiers = 1;
  p.checking_distance = 1.0;
  p.gains[0] = 1.0;
  p.ray_angle_x = p.ray_angle_y = 0.0025;
  fillEvalParams(&p);
  p.bv_is_setup = bounding_volume_ && bounding_volume_->is_setup;
  if (bounding_volume_) {
    p.bv_min[0] = bounding_volume_->x_min;
    p.bv_min[1] = bounding_volume_->y_min;
    p.bv_min[2] = bounding_volume_->z_min;
    p.bv_max[0] = bounding_volume_->x_max;
    p.bv_max[1] = bounding_volume_->y_max;
    p.bv_max[2] = bounding_volume_->z_max;
    p.bv_quat[0] = bounding_volume_->rotation_quat.w();
    p.bv_quat[1] = bounding_volume_->rotation_quat.x();
    p.bv_quat[2] = bounding_volume_->rotation_quat.y();
    p.bv_quat[3] = bounding_volume_->rotation_quat.z();
  } else {
    p.bv_quat[0] = 1.0;
  }
  if (eval_) a3d_eval_destroy(eval_);
  eval_ = nullptr;
  if (a3d_eval_create(gpu_map_->context(), &p, &eval_) != A3D_OK)
    planner_.printError(std::string("a3d_eval_create: ") + a3d_last_error(gpu_map_->context()));
  if (group_eval_) a3d_group_eval_destroy(group_eval_);
  group_eval_ = nullptr;
  if (gpu_map_->group() && a3d_group_eval_create(gpu_map_->group(), &p, &group_eval_) != A3D_OK)
    planner_.printError(std::string("a3d_group_eval_create: ") + a3d_group_last_error(gpu_map_->group()));
}

bool SimulatedSensorEvaluator::computeGain(TrajectorySegment* traj_in) {
  // simulated_sensor_evaluator.cpp:45-82
  auto* cam = dynamic_cast<sensor_model::CameraModel*>(sensor_model_.get());
  if (!cam || !cam->handle() || !eval_ || traj_in->trajectory.empty()) return false;
  std::vector<int> indices;
  cam->sampleViewpoints(&indices, *traj_in);
  std::vector<double> pos, quat;
  gatherPoses(*traj_in, indices, &pos, &quat);
  // visible voxels, already restricted to the bounding volume (:50-57)
  // When nothing is taken out of the list afterwards (no clear_from_parents, an evaluator that does not
  // erase), the gain the cast computed on the way is the gain of the stored list: no second trip to the
  // device with the list (VoxelWeightEvaluator measures from the last trajectory point,
  // voxel_weight_evaluator.cpp:53).
  const bool gain_from_cast = !p_clear_from_parents_ && (evalKind() == A3D_EVAL_VOXEL_TYPE || evalKind() == A3D_EVAL_VOXEL_WEIGHT);
  double cast_gain = 0.0;
  std::vector<Eigen::Vector3d> new_voxels;
  if (!runVisible(gpu_map_->context(), cam->handle(), eval_, static_cast<int>(indices.size()), pos.data(),
                  quat.data(), &new_voxels, gain_from_cast ? &cast_gain : nullptr, planner_,
                  traj_in->trajectory.back().position_W.data()))
    return false;
  noteDeviceTime(static_cast<int>(indices.size()), false);
  if (p_clear_from_parents_) {  // :60-76, one device pass per ancestor
    std::vector<uint8_t> keep(new_voxels.size(), 1);
    std::vector<double> flat(new_voxels.size() * 3);
    for (size_t i = 0; i < new_voxels.size(); ++i)
      for (int k = 0; k < 3; ++k) flat[3 * i + k] = new_voxels[i][k];
    for (TrajectorySegment* previous = traj_in->parent; previous; previous = previous->parent) {
      auto* info = dynamic_cast<SimulatedSensorInfo*>(previous->info.get());
      if (!info || info->visible_voxels.empty()) continue;
      std::vector<double> old(info->visible_voxels.size() * 3);
      for (size_t i = 0; i < info->visible_voxels.size(); ++i)
        for (int k = 0; k < 3; ++k) old[3 * i + k] = info->visible_voxels[i][k];
      if (a3d_filter_not_in(gpu_map_->context(), static_cast<int64_t>(new_voxels.size()), flat.data(),
                            static_cast<int64_t>(info->visible_voxels.size()), old.data(),
                            keep.data()) != A3D_OK) {
        planner_.printError(std::string("a3d_filter_not_in: ") + a3d_last_error(gpu_map_->context()));
        return false;
      }
    }
    size_t w = 0;
    for (size_t i = 0; i < new_voxels.size(); ++i)
      if (keep[i]) new_voxels[w++] = new_voxels[i];
    new_voxels.resize(w);
  }
  storeTrajectoryInformation(traj_in, new_voxels);
  if (gain_from_cast)
    traj_in->gain = cast_gain;
  else
    computeGainFromVisibleVoxels(traj_in);
  return true;
}

bool SimulatedSensorEvaluator::storeTrajectoryInformation(TrajectorySegment* traj_in,
                                                          const std::vector<Eigen::Vector3d>& new_voxels) {
  SimulatedSensorInfo* new_info = new SimulatedSensorInfo();  // :84-91
  new_info->visible_voxels.assign(new_voxels.begin(), new_voxels.end());
  traj_in->info.reset(new_info);
  return true;
}

bool SimulatedSensorEvaluator::computeGainFromVisibleVoxels(TrajectorySegment* traj_in) {
  // voxel_type_evaluator.cpp:57-89 / naive_evaluator.cpp:20-38 / frontier_evaluator.cpp:100-124:
  // the fold runs on the device; Naive and Frontier also drop the observed entries from the list
  traj_in->gain = 0.0;
  if (!traj_in->info || !eval_) return false;
  if (dynamic_cast<DeviceSensorInfo*>(traj_in->info.get())) return updateGains({traj_in});
  auto* info = dynamic_cast<SimulatedSensorInfo*>(traj_in->info.get());
  if (!info) return false;
  const size_t n = info->visible_voxels.size();
  std::vector<double> flat(n * 3);
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) flat[3 * i + k] = info->visible_voxels[i][k];
  std::vector<uint8_t> keep(n, 1);
  double gain = 0.0;
  int64_t n_left = 0;
  const bool erases = evalKind() == A3D_EVAL_NAIVE || evalKind() == A3D_EVAL_FRONTIER;
  // VoxelWeightEvaluator measures from the segment's last trajectory point (voxel_weight_evaluator.cpp:53)
  const double* origin = traj_in->trajectory.empty() ? nullptr : traj_in->trajectory.back().position_W.data();
  if (a3d_gain_from_voxels(gpu_map_->context(), eval_, static_cast<int64_t>(n), flat.data(), &gain, &n_left,
                           erases ? keep.data() : nullptr, origin) != A3D_OK) {
    planner_.printError(std::string("a3d_gain_from_voxels: ") + a3d_last_error(gpu_map_->context()));
    return false;
  }
  if (erases) {
    size_t w = 0;
    for (size_t i = 0; i < n; ++i)
      if (keep[i]) info->visible_voxels[w++] = info->visible_voxels[i];
    info->visible_voxels.resize(w);
  }
  traj_in->gain = gain;
  return true;
}

bool SimulatedSensorEvaluator::computeGains(const std::vector<TrajectorySegment*>& segments) {
  auto* cam = dynamic_cast<sensor_model::CameraModel*>(sensor_model_.get());
  if (!cam || !cam->handle() || !eval_) return false;
  const size_t n = segments.size();
  // clear_from_parents (simulated_sensor_evaluator.cpp:60-76) as a batch: a segment's list loses what
  // its ancestors inside the batch see.  The device call wants parents before children, so the batch
  // is laid out by depth; `order[k]` = position in `segments` of the k-th segment handed over.
  std::vector<size_t> order(n);
  for (size_t s = 0; s < n; ++s) order[s] = s;
  std::vector<int32_t> parent_index;
  if (p_clear_from_parents_) {
    std::vector<int> depth(n, 0);
    for (size_t s = 0; s < n; ++s)
      for (TrajectorySegment* a = segments[s]->parent; a; a = a->parent) ++depth[s];
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return depth[a] < depth[b]; });
    std::map<const TrajectorySegment*, int32_t> slot;
    for (size_t k = 0; k < n; ++k) slot[segments[order[k]]] = static_cast<int32_t>(k);
    parent_index.assign(n, -1);
    for (size_t k = 0; k < n; ++k) {
      // nearest ancestor that is part of the batch (the ones in between contribute no list)
      for (TrajectorySegment* a = segments[order[k]]->parent; a; a = a->parent) {
        auto it = slot.find(a);
        if (it != slot.end()) {
          parent_index[k] = it->second;
          break;
        }
      }
    }
  }
  std::vector<double> pos, quat, origins(n * 3, 0.0);
  std::vector<int32_t> view_segment;
  for (size_t k = 0; k < n; ++k) {
    TrajectorySegment* seg = segments[order[k]];
    if (seg->trajectory.empty()) continue;
    std::vector<int> indices;
    cam->sampleViewpoints(&indices, *seg);
    gatherPoses(*seg, indices, &pos, &quat);
    view_segment.insert(view_segment.end(), indices.size(), static_cast<int32_t>(k));
    for (int c = 0; c < 3; ++c) origins[3 * k + c] = seg->trajectory.back().position_W[c];
  }
  std::vector<double> gains(n, 0.0);
  const int n_views = static_cast<int>(view_segment.size());
  // every GPU of the map's group takes a share of the segments (the tree form stays on one GPU: the
  // ancestors' lists would have to travel)
  const bool sharded = !p_clear_from_parents_ && gpu_map_->group() && cam->groupHandle() && group_eval_;
  int rc;
  if (sharded) {
    rc = a3d_group_compute_gains(gpu_map_->group(), cam->groupHandle(), group_eval_, n_views, pos.data(),
                                 quat.data(), view_segment.data(), static_cast<int>(n), gains.data(), nullptr,
                                 nullptr, nullptr, nullptr, origins.data());
  } else if (p_clear_from_parents_) {
    rc = a3d_compute_gains_tree(gpu_map_->context(), cam->handle(), eval_, n_views, pos.data(), quat.data(),
                                view_segment.data(), static_cast<int>(n), parent_index.data(), gains.data(),
                                nullptr, nullptr, nullptr, nullptr, origins.data());
  } else {
    rc = a3d_compute_gains(gpu_map_->context(), cam->handle(), eval_, n_views, pos.data(), quat.data(),
                           view_segment.data(), static_cast<int>(n), gains.data(), nullptr, nullptr, nullptr,
                           nullptr, origins.data());
  }
  if (rc != A3D_OK) {
    planner_.printError(std::string("computeGains: ") + (sharded ? a3d_group_last_error(gpu_map_->group())
                                                                 : a3d_last_error(gpu_map_->context())));
    return false;
  }
  noteDeviceTime(n_views, sharded);
  for (size_t k = 0; k < n; ++k) segments[order[k]]->gain = gains[k];
  return true;
}

ModuleFactoryRegistry::Registration<VoxelTypeEvaluator> VoxelTypeEvaluator::registration(
    "VoxelTypeEvaluator");
void VoxelTypeEvaluator::setupFromParamMap(Module::ParamMap* param_map) {  // voxel_type_evaluator.cpp:19-55
  SimulatedSensorEvaluator::setupFromParamMap(param_map);
  setParam<double>(param_map, "gain_unknown", &p_gain_unknown_, 1.0);
  setParam<double>(param_map, "gain_occupied", &p_gain_occupied_, 0.0);
  setParam<double>(param_map, "gain_free", &p_gain_free_, 0.0);
  setParam<double>(param_map, "gain_unknown_outer", &p_gain_unknown_outer_, 0.0);
  setParam<double>(param_map, "gain_occupied_outer", &p_gain_occupied_outer_, 0.0);
  setParam<double>(param_map, "gain_free_outer", &p_gain_free_outer_, 0.0);
  const std::string ns = (*param_map)["param_namespace"];
  std::string outer_volume_args;
  setParam<std::string>(param_map, "outer_volume_args", &outer_volume_args, ns + "/outer_volume");
  outer_volume_ =
      planner_.getFactory().createModule<BoundingVolume>(outer_volume_args, planner_, verbose_modules_);
  c_min_gain_ = std::min({p_gain_unknown_, p_gain_occupied_, p_gain_free_, p_gain_unknown_outer_,
                          p_gain_occupied_outer_, p_gain_free_outer_});
  c_max_gain_ = std::max({p_gain_unknown_, p_gain_occupied_, p_gain_free_, p_gain_unknown_outer_,
                          p_gain_occupied_outer_, p_gain_free_outer_});
  buildEval();
}
int VoxelTypeEvaluator::evalKind() const { return A3D_EVAL_VOXEL_TYPE; }
void VoxelTypeEvaluator::fillEvalParams(void* ptr) {
  auto* p = static_cast<a3d_eval_params*>(ptr);
  p->gains[0] = p_gain_unknown_;
  p->gains[1] = p_gain_occupied_;
  p->gains[2] = p_gain_free_;
  p->gains[3] = p_gain_unknown_outer_;
  p->gains[4] = p_gain_occupied_outer_;
  p->gains[5] = p_gain_free_outer_;
}

ModuleFactoryRegistry::Registration<NaiveEvaluator> NaiveEvaluator::registration("NaiveEvaluator");
void NaiveEvaluator::setupFromParamMap(Module::ParamMap* param_map) {  // naive_evaluator.cpp:15-18
  SimulatedSensorEvaluator::setupFromParamMap(param_map);
  buildEval();
}
int NaiveEvaluator::evalKind() const { return A3D_EVAL_NAIVE; }

ModuleFactoryRegistry::Registration<FrontierEvaluator> FrontierEvaluator::registration("FrontierEvaluator");
void FrontierEvaluator::setupFromParamMap(Module::ParamMap* param_map) {  // frontier_evaluator.cpp:15-28
  readFrontierParams(param_map);
  buildEval();
}
int FrontierEvaluator::evalKind() const { return A3D_EVAL_FRONTIER; }
void FrontierEvaluator::fillEvalParams(void* ptr) {
  auto* p = static_cast<a3d_eval_params*>(ptr);
  p->accurate_frontiers = p_accurate_frontiers_ ? 1 : 0;
  p->surface_frontiers = p_surface_frontiers_ ? 1 : 0;
  p->checking_distance = p_checking_distance_;
}
void FrontierEvaluator::readFrontierParams(Module::ParamMap* param_map) {
  SimulatedSensorEvaluator::setupFromParamMap(param_map);
  setParam<bool>(param_map, "accurate_frontiers", &p_accurate_frontiers_, false);
  setParam<bool>(param_map, "surface_frontiers", &p_surface_frontiers_, true);
  setParam<double>(param_map, "checking_distance", &p_checking_distance_, 1.0);
}

ModuleFactoryRegistry::Registration<VoxelWeightEvaluator> VoxelWeightEvaluator::registration(
    "VoxelWeightEvaluator");
void VoxelWeightEvaluator::setupFromParamMap(Module::ParamMap* param_map) {  // voxel_weight_evaluator.cpp:15-33
  readFrontierParams(param_map);
  setParam<double>(param_map, "frontier_voxel_weight", &p_frontier_voxel_weight_, 1.0);
  setParam<double>(param_map, "min_impact_factor", &p_min_impact_factor_, 0.0);
  setParam<double>(param_map, "new_voxel_weight", &p_new_voxel_weight_, 0.01);
  setParam<double>(param_map, "ray_angle_x", &p_ray_angle_x_, 0.0025);
  setParam<double>(param_map, "ray_angle_y", &p_ray_angle_y_, 0.0025);
  if (!dynamic_cast<map::TSDFMap*>(&(planner_.getMap())))
    planner_.printError("'VoxelWeightEvaluator' requires a map of type 'TSDFMap'!");
  buildEval();
}
int VoxelWeightEvaluator::evalKind() const { return A3D_EVAL_VOXEL_WEIGHT; }
void VoxelWeightEvaluator::fillEvalParams(void* ptr) {
  FrontierEvaluator::fillEvalParams(ptr);
  auto* p = static_cast<a3d_eval_params*>(ptr);
  p->frontier_voxel_weight = p_frontier_voxel_weight_;
  p->min_impact_factor = p_min_impact_factor_;
  p->new_voxel_weight = p_new_voxel_weight_;
  p->ray_angle_x = p_ray_angle_x_;
  p->ray_angle_y = p_ray_angle_y_;
}

// ---- yaw planning -------------------------------------------------------------------------------
// The reference evaluates the n_directions yaw copies of a segment one after the other and then
// picks one (yaw_planning_evaluator.cpp:48-110, continuous_yaw_planning_evaluator.cpp:35-92).  Here
// the copies of any number of segments form ONE batch for the device (evaluateOrientations) and the
// pick is a reduction over the copies' scores (pickOrientation), shared by both evaluators.
namespace {
double wrapAngle(double angle) {  // into [0, 2 pi), like tools/defaults.cpp:11-14
  angle = std::fmod(angle, 2.0 * M_PI);
  return angle < 0 ? angle + 2.0 * M_PI : angle;
}
// Index of the first maximum of `score`, and whether the scores discriminate between orientations at
// all.  The reference decides the latter with a running minimum that skips every element which raised
// the running maximum; `seeded` tells whether both start from score[0] (continuous evaluator) or from
// +-infinity (plain evaluator) — with the latter e.g. strictly increasing scores count as
// indifferent, which parity keeps.
struct Pick {
  int index;
  bool indifferent;
};
Pick pickOrientation(const std::vector<double>& score, bool seeded) {
  double hi = seeded ? score[0] : std::numeric_limits<double>::lowest();
  double lo = seeded ? score[0] : std::numeric_limits<double>::max();
  int best = 0;
  for (size_t i = seeded ? 1 : 0; i < score.size(); ++i) {
    const bool raises = score[i] > hi;
    if (raises) best = static_cast<int>(i);
    hi = raises ? score[i] : hi;
    lo = raises ? lo : std::min(lo, score[i]);
  }
  return Pick{best, !(hi > lo)};
}
double travelYaw(const TrajectorySegment& segment) {
  const Eigen::Vector3d d = segment.trajectory.back().position_W - segment.trajectory.front().position_W;
  return wrapAngle(std::atan2(d[1], d[0]));
}
}  // namespace

YawPlanningEvaluator::YawPlanningEvaluator(PlannerI& planner)
    : TrajectoryEvaluator(planner), p_n_directions_(4), p_select_by_value_(false), p_update_range_(-1.0),
      p_update_gain_(0.0), p_batch_orientations_(true), p_keep_visible_voxels_(false) {}

void YawPlanningEvaluator::setupFromParamMap(Module::ParamMap* param_map) {
  // parameter names and defaults of yaw_planning_evaluator.cpp:17-38 (update_range < 0: no updates)
  setParam<int>(param_map, "n_directions", &p_n_directions_, 4);
  setParam<bool>(param_map, "select_by_value", &p_select_by_value_, false);
  setParam<double>(param_map, "update_range", &p_update_range_, -1.0);
  setParam<double>(param_map, "update_gain", &p_update_gain_, 0.0);
  setParam<bool>(param_map, "batch_orientations", &p_batch_orientations_, true);
  setParam<bool>(param_map, "keep_visible_voxels", &p_keep_visible_voxels_, false);
  planner_.getFactory().registerLinkableModule("YawPlanningEvaluator", this);
  std::string args;
  setParam<std::string>(param_map, "following_evaluator_args", &args,
                        (*param_map)["param_namespace"] + "/following_evaluator");
  following_evaluator_ = planner_.getFactory().createModule<TrajectoryEvaluator>(args, planner_, verbose_modules_);
  TrajectoryEvaluator::setupFromParamMap(param_map);
}

bool YawPlanningEvaluator::checkParamsValid(std::string* error_message) {
  if (p_n_directions_ < 1) {
    *error_message = "n_directions expected > 0";
    return false;
  }
  return TrajectoryEvaluator::checkParamsValid(error_message);
}

void YawPlanningEvaluator::spawnOrientations(TrajectorySegment* traj_in) {
  auto info = std::make_unique<YawPlanningInfo>();
  const double from = traj_in->trajectory.front().getYaw(), to = traj_in->trajectory.back().getYaw();
  info->orientations.resize(p_n_directions_);
  for (int i = 0; i < p_n_directions_; ++i) {
    info->orientations[i] = traj_in->shallowCopy();
    setTrajectoryYaw(&info->orientations[i], from, sampleYaw(to, i));
  }
  traj_in->info = std::move(info);
}

bool YawPlanningEvaluator::evaluateOrientations(const std::vector<TrajectorySegment*>& copies) {
  if (copies.empty()) return true;
  auto* device = p_batch_orientations_ ? dynamic_cast<SimulatedSensorEvaluator*>(following_evaluator_.get()) : nullptr;
  bool ok = device && (p_keep_visible_voxels_ ? device->computeGainsStored(copies) : device->computeGains(copies));
  if (!ok) {  // another kind of following evaluator: copy by copy
    ok = true;
    for (TrajectorySegment* c : copies) ok &= following_evaluator_->computeGain(c);
  }
  for (TrajectorySegment* c : copies) {
    if (!p_select_by_value_) break;
    following_evaluator_->computeCost(c);
    following_evaluator_->computeValue(c);
  }
  return ok;
}

bool YawPlanningEvaluator::computeOrientationGains(const std::vector<TrajectorySegment*>& copies) {
  const bool by_value = p_select_by_value_;  // gains only: the caller decides about cost and value
  p_select_by_value_ = false;
  const bool ok = evaluateOrientations(copies);
  p_select_by_value_ = by_value;
  return ok;
}

bool YawPlanningEvaluator::computeGains(const std::vector<TrajectorySegment*>& segments) {
  std::vector<TrajectorySegment*> copies;
  copies.reserve(segments.size() * p_n_directions_);
  for (TrajectorySegment* seg : segments) {
    spawnOrientations(seg);
    for (auto& o : static_cast<YawPlanningInfo*>(seg->info.get())->orientations) copies.push_back(&o);
  }
  const bool ok = evaluateOrientations(copies);
  for (TrajectorySegment* seg : segments) setBestYaw(seg);
  return ok;
}
bool YawPlanningEvaluator::computeGain(TrajectorySegment* traj_in) {
  computeGains({traj_in});
  return true;
}

void YawPlanningEvaluator::setBestYaw(TrajectorySegment* segment) {
  auto* info = static_cast<YawPlanningInfo*>(segment->info.get());
  std::vector<double> score(info->orientations.size());
  for (size_t i = 0; i < score.size(); ++i)
    score[i] = p_select_by_value_ ? info->orientations[i].value : info->orientations[i].gain;
  const Pick pick = pickOrientation(score, false);
  const TrajectorySegment& chosen = info->orientations[pick.index];
  setTrajectoryYaw(segment, segment->trajectory.front().getYaw(),
                   pick.indifferent ? travelYaw(*segment) : wrapAngle(chosen.trajectory.back().getYaw()));
  info->active_orientation = pick.index;
  segment->gain = chosen.gain;
  if (p_select_by_value_) {
    segment->cost = chosen.cost;
    segment->value = chosen.value;
  }
}

bool YawPlanningEvaluator::computeCost(TrajectorySegment* traj_in) { return following_evaluator_->computeCost(traj_in); }
bool YawPlanningEvaluator::computeValue(TrajectorySegment* traj_in) {
  return following_evaluator_->computeValue(traj_in);
}
int YawPlanningEvaluator::selectNextBest(TrajectorySegment* traj_in) {
  return following_evaluator_->selectNextBest(traj_in);
}

// Which copies of `segment` an update re-evaluates (yaw_planning_evaluator.cpp:128-160): copies follow
// their segment when it was re-parented or moved; a non-root segment within update_range of the robot
// gets the gain of every copy above update_gain again.  Appends them to *again.
bool YawPlanningEvaluator::collectStaleOrientations(TrajectorySegment* segment, std::vector<TrajectorySegment*>* again) {
  auto* info = dynamic_cast<YawPlanningInfo*>(segment->info.get());
  if (!info) return false;
  for (auto& o : info->orientations) {
    const bool moved = o.parent != segment->parent ||
                       !(o.trajectory.front().position_W == segment->trajectory.front().position_W);
    if (!moved) continue;
    const double yaw = o.trajectory.back().getYaw();
    o.parent = segment->parent;
    o.trajectory = segment->trajectory;
    setTrajectoryYaw(&o, segment->trajectory.front().getYaw(), yaw);
  }
  if (!segment->parent) return false;
  const double dist = (planner_.getCurrentPosition() - segment->trajectory.back().position_W).norm();
  if (p_update_range_ == 0.0 || p_update_range_ > dist)
    for (auto& o : info->orientations)
      if (o.gain > p_update_gain_) again->push_back(&o);
  return true;
}

void YawPlanningEvaluator::updateSegments(const std::vector<TrajectorySegment*>& segments,
                                          std::vector<uint8_t>* alive) {
  // the stale copies of ALL segments are one batch; gains only, whatever select_by_value says
  std::vector<TrajectorySegment*> again, repick;
  for (size_t i = 0; i < segments.size(); ++i)
    if ((*alive)[i] && collectStaleOrientations(segments[i], &again)) repick.push_back(segments[i]);
  const bool by_value = p_select_by_value_;
  p_select_by_value_ = false;
  evaluateOrientations(again);
  p_select_by_value_ = by_value;
  for (TrajectorySegment* seg : repick) setBestYaw(seg);
  following_evaluator_->updateSegments(segments, alive);
}
bool YawPlanningEvaluator::updateSegment(TrajectorySegment* segment) {
  std::vector<uint8_t> alive(1, 1);
  updateSegments({segment}, &alive);
  return alive[0] != 0;
}

ModuleFactoryRegistry::Registration<SimpleYawPlanningEvaluator> SimpleYawPlanningEvaluator::registration(
    "SimpleYawPlanningEvaluator");
void SimpleYawPlanningEvaluator::setupFromParamMap(Module::ParamMap* param_map) {
  setParam<bool>(param_map, "visualize_followup", &p_visualize_followup_, true);
  YawPlanningEvaluator::setupFromParamMap(param_map);
}
double SimpleYawPlanningEvaluator::sampleYaw(double original_yaw, int sample_number) {
  return wrapAngle(original_yaw + static_cast<double>(sample_number) * 2.0 * M_PI / p_n_directions_);
}
void SimpleYawPlanningEvaluator::setTrajectoryYaw(TrajectorySegment* segment, double /*start_yaw*/,
                                                  double target_yaw) {
  for (auto& p : segment->trajectory) p.setFromYaw(target_yaw);
}

ModuleFactoryRegistry::Registration<ContinuousYawPlanningEvaluator>
    ContinuousYawPlanningEvaluator::registration("ContinuousYawPlanningEvaluator");
void ContinuousYawPlanningEvaluator::setupFromParamMap(Module::ParamMap* param_map) {
  setParam<bool>(param_map, "visualize_followup", &p_visualize_followup_, true);
  setParam<int>(param_map, "n_sections_fov", &p_n_sections_fov_, 1);
  YawPlanningEvaluator::setupFromParamMap(param_map);
  p_select_by_value_ = false;  // the sections of one view are not independent: gains only
}
bool ContinuousYawPlanningEvaluator::computeGain(TrajectorySegment* traj_in) {
  return YawPlanningEvaluator::computeGain(traj_in);
}
void ContinuousYawPlanningEvaluator::setBestYaw(TrajectorySegment* segment) {
  // A view covers n_sections_fov adjacent sections of the yaw circle: its score is the circular
  // window sum of the section gains (continuous_yaw_planning_evaluator.cpp:44-92), added in the
  // reference's order (section i first, then i+1 ...).
  auto* info = static_cast<YawPlanningInfo*>(segment->info.get());
  const int n = static_cast<int>(info->orientations.size());
  std::vector<double> window(n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < p_n_sections_fov_; ++j) window[i] += info->orientations[(i + j) % n].gain;
  const Pick pick = pickOrientation(window, true);
  const double centre = info->orientations[pick.index].trajectory.back().getYaw() +
                        static_cast<double>(p_n_sections_fov_ - 1) / p_n_directions_ * M_PI;
  setTrajectoryYaw(segment, segment->trajectory.front().getYaw(),
                   pick.indifferent ? travelYaw(*segment) : wrapAngle(centre));
  info->active_orientation = pick.index;
  segment->gain = window[pick.index];
  following_evaluator_->computeCost(segment);  // the trajectory may have changed
  following_evaluator_->computeValue(segment);
}
double ContinuousYawPlanningEvaluator::sampleYaw(double original_yaw, int sample_number) {
  return wrapAngle(original_yaw + static_cast<double>(sample_number) * 2.0 * M_PI / p_n_directions_);
}
void ContinuousYawPlanningEvaluator::setTrajectoryYaw(TrajectorySegment* segment, double /*start_yaw*/,
                                                      double target_yaw) {
  for (auto& p : segment->trajectory) p.setFromYaw(target_yaw);
}

}  // namespace trajectory_evaluator

// ---- YAML factory -----------------------------------------------------------------------------------------
namespace {
std::string trim(const std::string& s) {
  const size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  const size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}
// strip a trailing comment that is not inside quotes
std::string stripComment(const std::string& s) {
  char quote = 0;
  for (size_t i = 0; i < s.size(); ++i) {
    const char c = s[i];
    if (quote) {
      if (c == quote) quote = 0;
    } else if (c == '"' || c == '\'') {
      quote = c;
    } else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) {
      return s.substr(0, i);
    }
  }
  return s;
}
// YAML scalar → the string XmlRpc's operator<< would produce (module_factory_ros.cpp:37-39)
std::string scalarToParamString(std::string v) {
  v = trim(v);
  if (v.size() >= 2 && ((v.front() == '"' && v.back() == '"') || (v.front() == '\'' && v.back() == '\'')))
    return v.substr(1, v.size() - 2);
  std::string low = v;
  std::transform(low.begin(), low.end(), low.begin(), ::tolower);
  if (low == "true" || low == "yes" || low == "on") return "1";
  if (low == "false" || low == "no" || low == "off") return "0";
  return v;
}
}  // namespace

std::string YamlTree::normalize(const std::string& ns) const {
  std::string out;
  bool slash = false;
  for (char c : "/" + ns) {
    if (c == '/') {
      if (!slash) out += c;
      slash = true;
    } else {
      out += c;
      slash = false;
    }
  }
  while (out.size() > 1 && out.back() == '/') out.pop_back();
  return out;
}

bool YamlTree::loadFile(const std::string& path, std::string* error) {
  std::ifstream f(path);
  if (!f) {
    if (error) *error = "cannot open " + path;
    return false;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  return loadString(ss.str(), error);
}

bool YamlTree::loadString(const std::string& text, std::string* error) {
  std::vector<std::pair<int, std::string>> stack;  // (indent, key)
  std::istringstream in(text);
  std::string raw;
  int line_no = 0;
  while (std::getline(in, raw)) {
    ++line_no;
    const std::string line = stripComment(raw);
    if (trim(line).empty() || trim(line) == "---") continue;
    const int indent = static_cast<int>(line.find_first_not_of(' '));
    if (line[indent] == '\t' || line[indent] == '-') {
      if (error) *error = "line " + std::to_string(line_no) + ": tabs / sequences are not supported";
      return false;
    }
    const size_t colon = line.find(':', indent);
    if (colon == std::string::npos) {
      if (error) *error = "line " + std::to_string(line_no) + ": expected 'key: value'";
      return false;
    }
    std::string key = trim(line.substr(indent, colon - indent));
    if (key.size() >= 2 && (key.front() == '"' || key.front() == '\'')) key = key.substr(1, key.size() - 2);
    const std::string value = trim(line.substr(colon + 1));
    while (!stack.empty() && stack.back().first >= indent) stack.pop_back();
    std::string path = root_;
    for (const auto& e : stack) path += "/" + e.second;
    path += "/" + key;
    if (value.empty()) {
      stack.push_back({indent, key});
    } else {
      values_[normalize(path)] = scalarToParamString(value);
    }
  }
  return true;
}

bool YamlTree::hasNamespace(const std::string& ns) const {
  const std::string prefix = normalize(ns) + "/";
  auto it = values_.lower_bound(prefix);
  return it != values_.end() && it->first.compare(0, prefix.size(), prefix) == 0;
}

std::map<std::string, std::string> YamlTree::scalars(const std::string& ns) const {
  std::map<std::string, std::string> out;
  const std::string prefix = normalize(ns) == "/" ? "/" : normalize(ns) + "/";
  for (auto it = values_.lower_bound(prefix); it != values_.end(); ++it) {
    if (it->first.compare(0, prefix.size(), prefix) != 0) break;
    const std::string rest = it->first.substr(prefix.size());
    if (rest.find('/') == std::string::npos) out[rest] = it->second;
  }
  return out;
}

bool ModuleFactoryYaml::getParamMapAndType(Module::ParamMap* map, std::string* type, std::string args) {
  // module_factory_ros.cpp:17-56
  for (const auto& kv : tree_.scalars(args)) (*map)[kv.first] = kv.second;
  auto it = map->find("type");
  if (it != map->end()) *type = it->second;
  (*map)["param_namespace"] = args;
  if (map->find("verbose_text") == map->end())
    (*map)["verbose_text"] = "Creating Module '" + *type + "' from namespace '" + args + "'";
  return true;
}
void ModuleFactoryYaml::printVerbose(const Module::ParamMap& map) {
  auto it = map.find("verbose_text");
  if (it != map.end()) verboseLog.push_back(it->second);
}
void ModuleFactoryYaml::printError(const std::string& message) { std::cerr << "[a3d] " << message << "\n"; }

}  // namespace active_3d_planning
