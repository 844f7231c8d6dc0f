// data.h — data carriers of the path.
//   EigenTrajectoryPoint, yaw helpers   core/include/.../data/trajectory.h:14-122
//   TrajectoryInfo, TrajectorySegment   core/include/.../data/trajectory_segment.h:12-68, src/data/trajectory_segment.cpp
//   BoundingVolume                      core/include/.../data/bounding_volume.h:14-36, src/data/bounding_volume.cpp:12-57
//   SystemConstraints                   core/include/.../data/system_constraints.h (fields only)
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "a3d_host/module.h"

namespace active_3d_planning {

inline double yawFromQuaternion(const Eigen::Quaterniond& q) {
  return atan2(2.0 * (q.w() * q.z() + q.x() * q.y()), 1.0 - 2.0 * (q.y() * q.y() + q.z() * q.z()));
}
inline Eigen::Quaterniond quaternionFromYaw(double yaw) {
  return Eigen::Quaterniond(Eigen::AngleAxisd(yaw, Eigen::Vector3d::UnitZ()));
}

struct EigenTrajectoryPoint {
  typedef std::vector<EigenTrajectoryPoint> Vector;
  int64_t timestamp_ns = -1;  // negative = invalid
  int64_t time_from_start_ns = 0;
  Eigen::Vector3d position_W, velocity_W, acceleration_W, jerk_W, snap_W;
  Eigen::Quaterniond orientation_W_B;
  Eigen::Vector3d angular_velocity_W, angular_acceleration_W;

  double getYaw() const { return yawFromQuaternion(orientation_W_B); }
  void setFromYaw(double yaw) { orientation_W_B = quaternionFromYaw(yaw); }
};
typedef EigenTrajectoryPoint::Vector EigenTrajectoryPointVector;

struct TrajectoryInfo {
  virtual ~TrajectoryInfo() = default;
};

struct TrajectorySegment {
  TrajectorySegment() : gain(0.0), cost(0.0), value(0.0), tg_visited(false), parent(nullptr) {}
  EigenTrajectoryPointVector trajectory;
  double gain, cost, value;
  bool tg_visited;
  TrajectorySegment* parent;
  std::vector<std::unique_ptr<TrajectorySegment>> children;
  std::unique_ptr<TrajectoryInfo> info;

  static bool comparePtr(TrajectorySegment* a, TrajectorySegment* b) { return a->value < b->value; }
  TrajectorySegment* spawnChild();
  void getChildren(std::vector<TrajectorySegment*>* result);
  void getLeaves(std::vector<TrajectorySegment*>* result);
  void getTree(std::vector<TrajectorySegment*>* result);
  void getTree(std::vector<TrajectorySegment*>* result, int maxdepth);
  // copies trajectory / gain / cost / value / parent, not info or children
  TrajectorySegment shallowCopy();
};

struct SystemConstraints {
  double v_max = 1.0, a_max = 1.0, yaw_rate_max = M_PI / 2.0, yaw_accel_max = M_PI / 2.0;
  double collision_radius = 1.0;
};

class BoundingVolume : public Module {
 public:
  explicit BoundingVolume(PlannerI& planner);  // NOLINT
  // true if not set up; else closed-interval box test after rotating by rotation_quat
  virtual bool contains(const Eigen::Vector3d& point);
  void setupFromParamMap(Module::ParamMap* param_map) override;

  bool is_setup;
  double x_min, x_max, y_min, y_max, z_min, z_max, rotation;
  Eigen::Quaterniond rotation_quat;

 protected:
  static ModuleFactoryRegistry::Registration<BoundingVolume> registration;
};

}  // namespace active_3d_planning
