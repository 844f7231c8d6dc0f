// sensor_model.h — simulated sensors of the path.
//   SensorModel          core/include/.../module/sensor_model.h:18-42, src/module/sensor_model/sensor_model.cpp:9-27
//   CameraModel          .../sensor_model/camera_model.h:16-66, camera_model.cpp:15-200
//   SimpleRayCaster      .../sensor_model/simple_ray_caster.h, simple_ray_caster.cpp:15-66
//   IterativeRayCaster   .../sensor_model/iterative_ray_caster.h, iterative_ray_caster.cpp:18-119
// The ray casting itself runs in k_cast_gain / k_cast_wave behind the C ABI.
#pragma once

#include <string>
#include <vector>

#include "a3d_host/map.h"

struct a3d_group_caster;

struct a3d_caster;

namespace active_3d_planning {

class SensorModel : public Module {
 public:
  explicit SensorModel(PlannerI& planner);  // NOLINT
  virtual ~SensorModel() = default;
  // appends the ordered voxel-centre list seen along the segment to *result
  virtual bool getVisibleVoxelsFromTrajectory(std::vector<Eigen::Vector3d>* result,
                                              const TrajectorySegment& traj_in) = 0;
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  map::OccupancyMap* map_;
  Eigen::Vector3d mounting_translation_;
  Eigen::Quaterniond mounting_rotation_;
};

namespace sensor_model {

class CameraModel : public SensorModel {
 public:
  explicit CameraModel(PlannerI& planner);  // NOLINT
  ~CameraModel() override;

  bool getVisibleVoxelsFromTrajectory(std::vector<Eigen::Vector3d>* result,
                                      const TrajectorySegment& traj_in) override;
  // One view from an explicit camera pose.  Appends the view's list AFTER the adjacent-unique that
  // every reference caller applies (camera_model.cpp:58); see INTEGRATION.md.
  virtual bool getVisibleVoxels(std::vector<Eigen::Vector3d>* result, const Eigen::Vector3d& position,
                                const Eigen::Quaterniond& orientation);
  void setupFromParamMap(Module::ParamMap* param_map) override;
  bool checkParamsValid(std::string* error_message) override;

  // indices of the trajectory points that are rendered (camera_model.cpp:62-83)
  void sampleViewpoints(std::vector<int>* result, const TrajectorySegment& traj_in);
  void getDirectionVector(Eigen::Vector3d* result, double relative_x, double relative_y);

  // derived constants (c_res_x_, c_res_y_, c_n_sections_ of the reference classes)
  int rayGridX() const;
  int rayGridY() const;
  int numSections() const;
  a3d_caster* handle() { return caster_; }
  a3d_group_caster* groupHandle() { return group_caster_; }  // non-null on a multi-GPU map

  // `test: true` (camera_model.cpp:31-54): the reference times every getVisibleVoxels call on the
  // wall clock and logs mean / stddev every 20 calls.  Here the time of a view is the CUDA-event
  // time of the device call that cast it, divided by the views of that call; same log line.
  void noteCastTime(double device_ms, int n_views);
  // (mean, stddev, count) over everything noted so far, in ms per view
  void castTimeStats(double* mean, double* stddev, size_t* count) const;

 protected:
  virtual int casterKind() const = 0;
  void buildCasters(Module::ParamMap* param_map);
  map::VoxbloxMap* gpu_map_;
  a3d_caster* caster_;          // with the mounting transform (trajectory-level calls)
  a3d_caster* caster_camera_;   // identity mount (getVisibleVoxels takes a camera pose)
  a3d_group_caster* group_caster_;  // the trajectory-level caster on every GPU of the map's group
  std::vector<double> time_count_;
  double p_ray_length_, p_sampling_time_, p_focal_length_, p_ray_step_, p_downsampling_factor_;
  double p_fov_x_deg_, p_fov_y_deg_;  // lidar casters only
  int p_resolution_x_, p_resolution_y_;
  bool p_test_;
  double c_field_of_view_x_, c_field_of_view_y_;
};

class SimpleRayCaster : public CameraModel {
 public:
  explicit SimpleRayCaster(PlannerI& planner) : CameraModel(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  int casterKind() const override;
  static ModuleFactoryRegistry::Registration<SimpleRayCaster> registration;
};

class IterativeRayCaster : public CameraModel {
 public:
  explicit IterativeRayCaster(PlannerI& planner) : CameraModel(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  int casterKind() const override;
  static ModuleFactoryRegistry::Registration<IterativeRayCaster> registration;
};

// IterativeRayCasterLidar (iterative_ray_caster_lidar.cpp, lidar_model.cpp:159-200): the iterative
// scheme over a spherical direction model.  The reference derives it from LidarModel, whose
// trajectory-level logic is identical to CameraModel's; here it shares that code by deriving from
// CameraModel and only parses the lidar parameter set.
class IterativeRayCasterLidar : public CameraModel {
 public:
  explicit IterativeRayCasterLidar(PlannerI& planner) : CameraModel(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;
  bool checkParamsValid(std::string* error_message) override;

 protected:
  int casterKind() const override;
  static ModuleFactoryRegistry::Registration<IterativeRayCasterLidar> registration;
};

}  // namespace sensor_model
}  // namespace active_3d_planning
