// trajectory_evaluator.h — gain evaluators of the path.
//   TrajectoryEvaluator (+ Cost/Value/NextSelector/EvaluatorUpdater bases)
//                               core/include/.../module/trajectory_evaluator.h:24-90, src/module/trajectory_evaluator.cpp
//   SimulatedSensorEvaluator    .../trajectory_evaluator/simulated_sensor_evaluator.h:18-60, .cpp:16-91
//   VoxelTypeEvaluator          .../voxel_type_evaluator.h, .cpp:19-89
//   NaiveEvaluator              .../naive_evaluator.h, .cpp:15-38
//   FrontierEvaluator           .../frontier_evaluator.h, .cpp:15-124
//   VoxelWeightEvaluator        .../voxel_weight_evaluator.h, .cpp:15-83
//   YawPlanningEvaluator        .../yaw_planning_evaluator.h:18-72, .cpp:13-160
//   SimpleYawPlanningEvaluator  .../simple_yaw_planning_evaluator.h, .cpp:14-43
//   ContinuousYawPlanningEvaluator  .../continuous_yaw_planning_evaluator.h, .cpp:19-107
// computeGains() is the additive batch entry point: the reference evaluates one
// segment per call (online_planner.cpp:414-416); a B200 wants thousands per launch.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "a3d_host/sensor_model.h"

struct a3d_eval;
struct a3d_group_eval;
struct a3d_store;

namespace active_3d_planning {

class CostComputer : public Module {
 public:
  explicit CostComputer(PlannerI& planner) : Module(planner) {}  // NOLINT
  virtual bool computeCost(TrajectorySegment* traj_in) = 0;
};
class ValueComputer : public Module {
 public:
  explicit ValueComputer(PlannerI& planner) : Module(planner) {}  // NOLINT
  virtual bool computeValue(TrajectorySegment* traj_in) = 0;
};
class NextSelector : public Module {
 public:
  explicit NextSelector(PlannerI& planner) : Module(planner) {}  // NOLINT
  virtual int selectNextBest(TrajectorySegment* traj_in) = 0;
};
// Update step of the planner's evaluator (module/trajectory_evaluator.h:80-90): false = drop the
// segment and its subtree.  The reference walks the tree and updates one segment per call
// (online_planner.cpp:679-693); the device wants the tree at once, so the primary form here is the
// batch: updateSegments() looks at every segment whose `alive` flag is set, clears the flag of those it
// drops and may re-evaluate the others in ONE launch.  updateSegment() is the one-element batch.
class EvaluatorUpdater : public Module {
 public:
  explicit EvaluatorUpdater(PlannerI& planner) : Module(planner) {}  // NOLINT
  virtual void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) = 0;
  virtual bool updateSegment(TrajectorySegment* segment);
};

class TrajectoryEvaluator : public Module {
 public:
  explicit TrajectoryEvaluator(PlannerI& planner);  // NOLINT
  virtual ~TrajectoryEvaluator() = default;
  virtual bool computeGain(TrajectorySegment* traj_in) = 0;
  // these four delegate to modules created lazily from <ns>/cost_computer etc.; the concrete cost /
  // value / selector / updater types are outside this path — register your own
  virtual bool computeCost(TrajectorySegment* traj_in);
  virtual bool computeValue(TrajectorySegment* traj_in);
  virtual int selectNextBest(TrajectorySegment* traj_in);
  virtual bool updateSegment(TrajectorySegment* segment);
  // additive: OnlinePlanner::updateEvaluatorStep (online_planner.cpp:679-693) for the whole tree below
  // `root` in one pass of the updater chain — every surviving segment is updated, the subtrees of the
  // dropped ones are removed.  Returns how many segments were removed.
  virtual int updateTree(TrajectorySegment* root);
  // batch form of updateSegment (default: the updater chain's batch)
  virtual void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive);
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  EvaluatorUpdater* updater();
  std::unique_ptr<BoundingVolume> bounding_volume_;
  std::string p_cost_args_, p_value_args_, p_next_args_, p_updater_args_;
  std::unique_ptr<CostComputer> cost_computer_;
  std::unique_ptr<ValueComputer> value_computer_;
  std::unique_ptr<NextSelector> next_selector_;
  std::unique_ptr<EvaluatorUpdater> evaluator_updater_;
};

namespace trajectory_evaluator {

// what computeGain stores in TrajectorySegment::info
struct SimulatedSensorInfo : public TrajectoryInfo {
  virtual ~SimulatedSensorInfo() = default;
  std::vector<Eigen::Vector3d> visible_voxels;
};

// Handle of the evaluator's device-resident list store; segments only hold weak references, so
// they may outlive the evaluator.
struct DeviceListStore {
  explicit DeviceListStore(a3d_store* s) : store(s) {}
  ~DeviceListStore();
  a3d_store* store;
};
// TrajectorySegment::info of a segment whose voxel list lives in HBM instead of
// in SimulatedSensorInfo::visible_voxels.  Destroying the info releases the list.
struct DeviceSensorInfo : public TrajectoryInfo {
  ~DeviceSensorInfo() override;
  uint64_t key = 0;
  uint64_t n_voxels = 0;
  std::weak_ptr<DeviceListStore> store;
};

class SimulatedSensorEvaluator : public TrajectoryEvaluator {
 public:
  explicit SimulatedSensorEvaluator(PlannerI& planner);  // NOLINT
  ~SimulatedSensorEvaluator() override;
  void setupFromParamMap(Module::ParamMap* param_map) override;
  bool computeGain(TrajectorySegment* traj_in) override;

  // Batch: gains of many segments in one launch (no voxel lists are stored in the segments).
  // Returns false if clear_from_parents is set (needs the per-segment lists).
  virtual bool computeGains(const std::vector<TrajectorySegment*>& segments);
  // fold of the list already stored in the segment (SimulatedSensorUpdater path); works for
  // host-side SimulatedSensorInfo lists and for DeviceSensorInfo
  virtual bool computeGainFromVisibleVoxels(TrajectorySegment* traj_in);
  // Batch computeGain that keeps every segment's list on the device (info = DeviceSensorInfo) ...
  virtual bool computeGainsStored(const std::vector<TrajectorySegment*>& segments);
  // ... and the batch re-fold of such segments after the map changed: what UpdateAll /
  // SimulatedSensorUpdater do segment by segment (update_all.cpp:16-27,
  // simulated_sensor_updater.cpp:18-22), in one launch
  virtual bool updateGains(const std::vector<TrajectorySegment*>& segments);
  SensorModel* sensorModel() { return sensor_model_.get(); }
  // device time (CUDA events) spent in gain evaluation since the last call: the `Gain` column of the
  // planner's performance log (online_planner.cpp:97-101, 413-420)
  double takeGainTimeMs();

 protected:
  void noteDeviceTime(int n_views, bool group);
  virtual bool storeTrajectoryInformation(TrajectorySegment* traj_in,
                                          const std::vector<Eigen::Vector3d>& new_voxels);
  virtual int evalKind() const = 0;
  // evaluator-specific parameters → a3d_eval_params (called at the end of setupFromParamMap)
  void buildEval();
  virtual void fillEvalParams(void* a3d_eval_params_ptr) {}
  map::OccupancyMap* map_;
  map::VoxbloxMap* gpu_map_;
  std::unique_ptr<SensorModel> sensor_model_;
  a3d_eval* eval_;
  a3d_group_eval* group_eval_;
  double perf_gain_ms_ = 0.0;
  std::shared_ptr<DeviceListStore> device_lists_;
  uint64_t next_key_ = 1;
  bool p_clear_from_parents_, p_visualize_sensor_view_;
};

class VoxelTypeEvaluator : public SimulatedSensorEvaluator {
 public:
  explicit VoxelTypeEvaluator(PlannerI& planner) : SimulatedSensorEvaluator(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  int evalKind() const override;
  void fillEvalParams(void* p) override;
  static ModuleFactoryRegistry::Registration<VoxelTypeEvaluator> registration;
  double p_gain_unknown_, p_gain_occupied_, p_gain_free_, p_gain_unknown_outer_, p_gain_occupied_outer_,
      p_gain_free_outer_, c_min_gain_, c_max_gain_;
  std::unique_ptr<BoundingVolume> outer_volume_;  // constructed, never consulted (as in the reference)
};

class NaiveEvaluator : public SimulatedSensorEvaluator {
 public:
  explicit NaiveEvaluator(PlannerI& planner) : SimulatedSensorEvaluator(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  int evalKind() const override;
  static ModuleFactoryRegistry::Registration<NaiveEvaluator> registration;
};

class FrontierEvaluator : public SimulatedSensorEvaluator {
 public:
  explicit FrontierEvaluator(PlannerI& planner) : SimulatedSensorEvaluator(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  int evalKind() const override;
  void fillEvalParams(void* p) override;
  void readFrontierParams(Module::ParamMap* param_map);
  static ModuleFactoryRegistry::Registration<FrontierEvaluator> registration;
  bool p_accurate_frontiers_, p_surface_frontiers_;
  double p_checking_distance_;
};

// Reconstruction gain of the paper: TSDF-weight based impact of surface voxels, frontier and new
// voxels with fixed weights.  Evaluated by the scan-order kernel.
class VoxelWeightEvaluator : public FrontierEvaluator {
 public:
  explicit VoxelWeightEvaluator(PlannerI& planner) : FrontierEvaluator(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  int evalKind() const override;
  void fillEvalParams(void* p) override;
  static ModuleFactoryRegistry::Registration<VoxelWeightEvaluator> registration;
  double p_frontier_voxel_weight_, p_min_impact_factor_, p_new_voxel_weight_, p_ray_angle_x_,
      p_ray_angle_y_;
};

// What the yaw-planning evaluators store in TrajectorySegment::info: every oriented copy of the
// segment (yaw_planning_evaluator.h:64-71).
struct YawPlanningInfo : public TrajectoryInfo {
  virtual ~YawPlanningInfo() = default;
  std::vector<TrajectorySegment> orientations;
  int active_orientation = 0;
};

// Evaluates n_directions yaw copies of a segment with the following evaluator and keeps the best.
// The copies of one segment — or of a whole batch of segments (computeGains) — are evaluated in ONE
// launch when the following evaluator is a SimulatedSensorEvaluator that can batch (SURVEY.md
// §8(f)-1: "the 12 yaw copies as one mini-batch"); otherwise one by one like the reference.
class YawPlanningEvaluator : public TrajectoryEvaluator {
 public:
  explicit YawPlanningEvaluator(PlannerI& planner);  // NOLINT
  bool computeGain(TrajectorySegment* traj_in) override;
  bool computeCost(TrajectorySegment* traj_in) override;
  bool computeValue(TrajectorySegment* traj_in) override;
  int selectNextBest(TrajectorySegment* traj_in) override;
  bool updateSegment(TrajectorySegment* segment) override;
  // the stale orientations of all the segments are re-evaluated in one launch
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override;
  void setupFromParamMap(Module::ParamMap* param_map) override;
  bool checkParamsValid(std::string* error_message) override;
  // additive: computeGain for many segments, all their orientations in one launch
  virtual bool computeGains(const std::vector<TrajectorySegment*>& segments);
  TrajectoryEvaluator* followingEvaluator() { return following_evaluator_.get(); }
  // for evaluator_updater::YawPlanningUpdater: the gains of oriented copies (of any number of segments)
  // in one launch where the following evaluator can, and a copy turned to another yaw
  bool computeOrientationGains(const std::vector<TrajectorySegment*>& copies);
  void turnOrientation(TrajectorySegment* copy, double start_yaw, double target_yaw) {
    setTrajectoryYaw(copy, start_yaw, target_yaw);
  }

 protected:
  std::unique_ptr<TrajectoryEvaluator> following_evaluator_;
  int p_n_directions_;
  bool p_select_by_value_;
  double p_update_range_;  // update only gains within this distance (0.0: all, < 0: none)
  double p_update_gain_;   // update only gains above this
  bool p_batch_orientations_;     // additive parameter, default true
  bool p_keep_visible_voxels_;    // additive: batch path keeps each copy's voxel list on the device

  virtual double sampleYaw(double original_yaw, int sample) = 0;
  virtual void setTrajectoryYaw(TrajectorySegment* segment, double start_yaw, double target_yaw) = 0;
  virtual void setBestYaw(TrajectorySegment* segment);
  // creates info->orientations of the segment (yaw_planning_evaluator.cpp:49-60, without evaluating)
  void spawnOrientations(TrajectorySegment* traj_in);
  // gain (and cost/value if select_by_value) of the given oriented copies
  bool evaluateOrientations(const std::vector<TrajectorySegment*>& copies);
  bool collectStaleOrientations(TrajectorySegment* segment, std::vector<TrajectorySegment*>* again);
};

class SimpleYawPlanningEvaluator : public YawPlanningEvaluator {
 public:
  explicit SimpleYawPlanningEvaluator(PlannerI& planner) : YawPlanningEvaluator(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  double sampleYaw(double original_yaw, int sample_number) override;
  void setTrajectoryYaw(TrajectorySegment* segment, double start_yaw, double target_yaw) override;
  static ModuleFactoryRegistry::Registration<SimpleYawPlanningEvaluator> registration;
  bool p_visualize_followup_;
};

// Splits the full yaw circle into n_directions sections of which n_sections_fov adjacent ones make up
// the sensor's field of view; the gain of a yaw is the sum over its sections.
class ContinuousYawPlanningEvaluator : public YawPlanningEvaluator {
 public:
  explicit ContinuousYawPlanningEvaluator(PlannerI& planner) : YawPlanningEvaluator(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;
  bool computeGain(TrajectorySegment* traj_in) override;

 protected:
  double sampleYaw(double original_yaw, int sample_number) override;
  void setTrajectoryYaw(TrajectorySegment* segment, double start_yaw, double target_yaw) override;
  void setBestYaw(TrajectorySegment* segment) override;
  static ModuleFactoryRegistry::Registration<ContinuousYawPlanningEvaluator> registration;
  bool p_visualize_followup_;
  int p_n_sections_fov_;
};

}  // namespace trajectory_evaluator
}  // namespace active_3d_planning
