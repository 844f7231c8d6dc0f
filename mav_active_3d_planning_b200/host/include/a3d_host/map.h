// map.h — map interface of the planner and its B200 implementation.
//   Map / OccupancyMap / TSDFMap   core/include/.../map/map.h:11-27, occupancy_map.h:11-30, tsdf_map.h:11-25
//   VoxbloxMap                     active_3d_planning_voxblox/include/.../map/voxblox.h, src/map/voxblox.cpp:17-115
//
// VoxbloxMap here owns the GPU mirror (a3d_ctx) instead of a voxblox::EsdfServer.  Map content arrives
// as block deltas — uploadBlocks() takes what voxblox's Layer::getAllUpdatedBlocks(Update::kMap) /
// a voxblox_msgs::Layer message holds: block indices plus per-block voxel arrays.  Every query is
// answered by the CUDA kernels behind the C ABI (include/a3d.h).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "a3d_host/data.h"

struct a3d_ctx;
struct a3d_group;

namespace active_3d_planning {

class Map : public Module {
 public:
  explicit Map(PlannerI& planner) : Module(planner) {}  // NOLINT
  virtual ~Map() = default;
  virtual bool isTraversable(const Eigen::Vector3d& position,
                             const Eigen::Quaterniond& orientation = Eigen::Quaterniond(1, 0, 0, 0)) = 0;
  virtual bool isObserved(const Eigen::Vector3d& point) = 0;
};

namespace map {

class OccupancyMap : public Map {
 public:
  explicit OccupancyMap(PlannerI& planner) : Map(planner) {}  // NOLINT
  static constexpr unsigned char OCCUPIED = 0, FREE = 1, UNKNOWN = 2;
  virtual unsigned char getVoxelState(const Eigen::Vector3d& point) = 0;
  virtual double getVoxelSize() = 0;
  virtual bool getVoxelCenter(Eigen::Vector3d* center, const Eigen::Vector3d& point) = 0;
};

class TSDFMap : public OccupancyMap {
 public:
  explicit TSDFMap(PlannerI& planner) : OccupancyMap(planner) {}  // NOLINT
  virtual double getVoxelDistance(const Eigen::Vector3d& point) = 0;
  virtual double getVoxelWeight(const Eigen::Vector3d& point) = 0;
  virtual double getMaximumWeight() = 0;
};

class VoxbloxMap : public TSDFMap {
 public:
  explicit VoxbloxMap(PlannerI& planner);  // NOLINT
  ~VoxbloxMap() override;

  // params: tsdf_voxel_size (0.1), tsdf_voxels_per_side (16), max_weight (10000) — the voxblox server
  // parameters of app_reconstruction/launch/example.launch:143-144 — and where the mirror lives:
  // device (0), or devices ("0,1,2,3": one replica per GPU behind an a3d_group; batch gain calls are
  // then sharded over all of them)
  void setupFromParamMap(Module::ParamMap* param_map) override;
  bool checkParamsValid(std::string* error_message) override;

  bool isTraversable(const Eigen::Vector3d& position, const Eigen::Quaterniond& orientation) override;
  bool isObserved(const Eigen::Vector3d& point) override;
  unsigned char getVoxelState(const Eigen::Vector3d& point) override;
  double getVoxelSize() override;
  bool getVoxelCenter(Eigen::Vector3d* center, const Eigen::Vector3d& point) override;
  double getVoxelDistance(const Eigen::Vector3d& point) override;
  double getVoxelWeight(const Eigen::Vector3d& point) override;
  double getMaximumWeight() override;

  // ---- delta interface (replaces the esdf_map_in / tsdf_map_in subscriptions) ----
  // insert-or-overwrite n blocks; voxel order x + vps*(y + vps*z); tsdf_weight may be null
  bool uploadBlocks(int n, const int32_t* block_idx, const float* esdf_distance,
                    const uint8_t* esdf_observed, const float* tsdf_weight);
  // TSDF weights alone, for blocks the mirror holds (a3d_map_upload_weights)
  bool uploadWeights(int n, const int32_t* block_idx, const float* tsdf_weight);
  // TSDF distances alone, for blocks the mirror holds (a3d_map_upload_tsdf_distances): what
  // getVoxelDistance reads
  bool uploadTsdfDistances(int n, const int32_t* block_idx, const float* tsdf_distance);
  bool removeBlocks(int n, const int32_t* block_idx);
  int64_t numBlocks() const;
  int voxelsPerSide() const { return p_voxels_per_side_; }

  // batched queries (one kernel launch each)
  void getVoxelStates(const std::vector<Eigen::Vector3d>& points, std::vector<unsigned char>* states);
  // isTraversable for many points (the collision checks of a whole candidate trajectory / RRT edge)
  void areTraversable(const std::vector<Eigen::Vector3d>& points, std::vector<unsigned char>* out);

  a3d_ctx* context() { return ctx_; }
  // non-null when `devices` names more than one GPU (context() is then the first GPU's)
  a3d_group* group() { return group_; }
  int numDevices() const;

 protected:
  static ModuleFactoryRegistry::Registration<VoxbloxMap> registration;
  void check(int rc, const char* what);
  a3d_ctx* ctx_;
  a3d_group* group_;
  std::string p_devices_;
  double p_voxel_size_;
  int p_voxels_per_side_;
  double c_voxel_size_;  // double(float voxel size), voxblox.cpp:26
  double c_maximum_weight_;
  int p_device_;
};

}  // namespace map
}  // namespace active_3d_planning
