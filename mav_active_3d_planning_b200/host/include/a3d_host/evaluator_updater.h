// evaluator_updater.h — the update step of the evaluator (reference: core/include/.../module/
// trajectory_evaluator/evaluator_updater/*.h, src/.../evaluator_updater/*.cpp), batch first.
//
// Every updater implements updateSegments(): one pass over ALL segments that are still alive, so that
// the stages that touch the device do it once for the whole tree —
//   SimulatedSensorUpdater  re-folds every stored voxel list against the current map in one launch
//                           (a3d_store_refold) instead of one computeGainFromVisibleVoxels per segment
//                           (simulated_sensor_updater.cpp:18-22),
//   UpdateAll               re-casts every segment in one batch (update_all.cpp:16-27),
//   YawPlanningUpdater      re-evaluates the orientation copies of ALL segments in reach as one cast
//                           (yaw_planning_updaters.cpp:56-119), YawPlanningUpdateAdapter hands the active
//                           copies of all segments to its following updater as one batch (:21-41),
// and the host-only stages (PruneDirect, ConstrainedUpdater, ResetTree, UpdateNothing) are filters over
// the flags.  updateSegment(segment) — the reference's per-segment contract — is the one-element
// batch (EvaluatorUpdater::updateSegment).  Names, parameters and defaults are the reference's, so
// the shipped YAML files load unchanged; TrajectoryEvaluator::updateTree() is what a planner calls
// in place of its recursive updateEvaluatorStep (online_planner.cpp:679-693).
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "a3d_host/trajectory_evaluator.h"

namespace active_3d_planning {
namespace evaluator_updater {

// common part: "following_updater" chain (default: <ns>/following_updater)
class ChainedUpdater : public EvaluatorUpdater {
 public:
  explicit ChainedUpdater(PlannerI& planner) : EvaluatorUpdater(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  // hands the segments whose flag is set and that `select` accepts to the following updater
  void follow(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive,
              const std::vector<uint8_t>* select = nullptr);
  std::unique_ptr<EvaluatorUpdater> following_updater_;
};

class UpdateNothing : public EvaluatorUpdater {  // update_nothing.cpp
 public:
  explicit UpdateNothing(PlannerI& planner) : EvaluatorUpdater(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap*) override {}
  void updateSegments(const std::vector<TrajectorySegment*>&, std::vector<uint8_t>*) override {}

 protected:
  static ModuleFactoryRegistry::Registration<UpdateNothing> registration;
};

class ResetTree : public EvaluatorUpdater {  // reset_tree.cpp: drop everything
 public:
  explicit ResetTree(PlannerI& planner) : EvaluatorUpdater(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap*) override {}
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override;

 protected:
  static ModuleFactoryRegistry::Registration<ResetTree> registration;
};

class PruneDirect : public ChainedUpdater {  // prune_direct.cpp: minimum_value, minimum_gain, maximum_cost
 public:
  explicit PruneDirect(PlannerI& planner) : ChainedUpdater(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override;

 protected:
  static ModuleFactoryRegistry::Registration<PruneDirect> registration;
  double p_minimum_value_, p_minimum_gain_, p_maximum_cost_;
};

class ConstrainedUpdater : public ChainedUpdater {  // constrained_updater.cpp: minimum_gain, update_range
 public:
  explicit ConstrainedUpdater(PlannerI& planner) : ChainedUpdater(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override;

 protected:
  static ModuleFactoryRegistry::Registration<ConstrainedUpdater> registration;
  double p_minimum_gain_, p_update_range_;
};

class UpdateAll : public ChainedUpdater {  // update_all.cpp: update_gain, update_cost, update_value
 public:
  explicit UpdateAll(PlannerI& planner) : ChainedUpdater(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override;

 protected:
  static ModuleFactoryRegistry::Registration<UpdateAll> registration;
  bool update_gain_, update_cost_, update_value_;
};

class SimulatedSensorUpdater : public ChainedUpdater {  // simulated_sensor_updater.cpp
 public:
  explicit SimulatedSensorUpdater(PlannerI& planner) : ChainedUpdater(planner), evaluator_(nullptr) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override;

 protected:
  static ModuleFactoryRegistry::Registration<SimulatedSensorUpdater> registration;
  trajectory_evaluator::SimulatedSensorEvaluator* evaluator_;
};

// yaw_planning_updaters.cpp:14-52 — lets updaters that know nothing about YawPlanningInfo work on the
// active orientation of every segment: parameter dynamic_trajectories.
class YawPlanningUpdateAdapter : public ChainedUpdater {
 public:
  explicit YawPlanningUpdateAdapter(PlannerI& planner) : ChainedUpdater(planner) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override;

 protected:
  static ModuleFactoryRegistry::Registration<YawPlanningUpdateAdapter> registration;
  bool p_dynamic_trajectories_;
};

// yaw_planning_updaters.cpp:54-138 — every orientation of a segment in reach is evaluated again, the best
// one becomes the segment: parameters select_by_value, dynamic_trajectories, update_range (default -1:
// nothing is re-evaluated), update_gain.
class YawPlanningUpdater : public ChainedUpdater {
 public:
  explicit YawPlanningUpdater(PlannerI& planner) : ChainedUpdater(planner), evaluator_(nullptr) {}  // NOLINT
  void setupFromParamMap(Module::ParamMap* param_map) override;
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override;

 protected:
  static ModuleFactoryRegistry::Registration<YawPlanningUpdater> registration;
  trajectory_evaluator::YawPlanningEvaluator* evaluator_;
  bool p_select_by_value_, p_dynamic_trajectories_;
  double p_update_range_, p_update_gain_;
};

}  // namespace evaluator_updater
}  // namespace active_3d_planning
