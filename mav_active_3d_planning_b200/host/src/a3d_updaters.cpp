// a3d_updaters.cpp — evaluator updaters, batch first (see a3d_host/evaluator_updater.h).
#include "a3d_host/evaluator_updater.h"

namespace active_3d_planning {
namespace evaluator_updater {

namespace {
// the subset of `segments` with alive && (!select || select), and where each came from
struct Subset {
  std::vector<TrajectorySegment*> segments;
  std::vector<size_t> origin;
};
Subset gather(const std::vector<TrajectorySegment*>& segments, const std::vector<uint8_t>& alive,
              const std::vector<uint8_t>* select) {
  Subset out;
  for (size_t i = 0; i < segments.size(); ++i)
    if (alive[i] && (!select || (*select)[i])) {
      out.segments.push_back(segments[i]);
      out.origin.push_back(i);
    }
  return out;
}
}  // namespace

void ChainedUpdater::setupFromParamMap(Module::ParamMap* param_map) {
  std::string args;
  setParam<std::string>(param_map, "following_updater_args", &args,
                        (*param_map)["param_namespace"] + "/following_updater");
  following_updater_ = planner_.getFactory().createModule<EvaluatorUpdater>(args, planner_, verbose_modules_);
}
void ChainedUpdater::follow(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive,
                            const std::vector<uint8_t>* select) {
  Subset sub = gather(segments, *alive, select);
  if (sub.segments.empty()) return;
  std::vector<uint8_t> flags(sub.segments.size(), 1);
  following_updater_->updateSegments(sub.segments, &flags);
  for (size_t k = 0; k < flags.size(); ++k) (*alive)[sub.origin[k]] = flags[k];
}

ModuleFactoryRegistry::Registration<UpdateNothing> UpdateNothing::registration("UpdateNothing");

ModuleFactoryRegistry::Registration<ResetTree> ResetTree::registration("ResetTree");
void ResetTree::updateSegments(const std::vector<TrajectorySegment*>&, std::vector<uint8_t>* alive) {
  std::fill(alive->begin(), alive->end(), 0);
}

ModuleFactoryRegistry::Registration<PruneDirect> PruneDirect::registration("PruneDirect");
void PruneDirect::setupFromParamMap(Module::ParamMap* param_map) {
  setParam<double>(param_map, "minimum_value", &p_minimum_value_, -1e9);
  setParam<double>(param_map, "minimum_gain", &p_minimum_gain_, -1e9);
  setParam<double>(param_map, "maximum_cost", &p_maximum_cost_, 1e9);
  ChainedUpdater::setupFromParamMap(param_map);
}
void PruneDirect::updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) {
  for (size_t i = 0; i < segments.size(); ++i) {
    const TrajectorySegment& s = *segments[i];
    if (s.value < p_minimum_value_ || s.gain < p_minimum_gain_ || s.cost > p_maximum_cost_) (*alive)[i] = 0;
  }
  follow(segments, alive);
}

ModuleFactoryRegistry::Registration<ConstrainedUpdater> ConstrainedUpdater::registration("ConstrainedUpdater");
void ConstrainedUpdater::setupFromParamMap(Module::ParamMap* param_map) {
  setParam<double>(param_map, "minimum_gain", &p_minimum_gain_, -1e9);
  setParam<double>(param_map, "update_range", &p_update_range_, 1e9);
  ChainedUpdater::setupFromParamMap(param_map);
}
void ConstrainedUpdater::updateSegments(const std::vector<TrajectorySegment*>& segments,
                                        std::vector<uint8_t>* alive) {
  // only non-root segments above minimum_gain that end within update_range of the robot go on; the
  // others stay as they are
  const Eigen::Vector3d here = planner_.getCurrentPosition();
  std::vector<uint8_t> select(segments.size(), 0);
  for (size_t i = 0; i < segments.size(); ++i) {
    const TrajectorySegment& s = *segments[i];
    select[i] = s.parent && s.gain > p_minimum_gain_ &&
                (here - s.trajectory.back().position_W).norm() <= p_update_range_;
  }
  follow(segments, alive, &select);
}

ModuleFactoryRegistry::Registration<UpdateAll> UpdateAll::registration("UpdateAll");
void UpdateAll::setupFromParamMap(Module::ParamMap* param_map) {
  setParam<bool>(param_map, "update_gain", &update_gain_, true);
  setParam<bool>(param_map, "update_cost", &update_cost_, true);
  setParam<bool>(param_map, "update_value", &update_value_, true);
  ChainedUpdater::setupFromParamMap(param_map);
}
void UpdateAll::updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) {
  const Subset sub = gather(segments, *alive, nullptr);
  TrajectoryEvaluator& evaluator = planner_.getTrajectoryEvaluator();
  if (update_gain_ && !sub.segments.empty()) {
    // one cast of the whole batch where the evaluator can, else segment by segment
    bool done = false;
    if (auto* sim = dynamic_cast<trajectory_evaluator::SimulatedSensorEvaluator*>(&evaluator))
      done = sim->computeGains(sub.segments);
    else if (auto* yaw = dynamic_cast<trajectory_evaluator::YawPlanningEvaluator*>(&evaluator))
      done = yaw->computeGains(sub.segments);
    if (!done)
      for (TrajectorySegment* s : sub.segments) evaluator.computeGain(s);
  }
  for (TrajectorySegment* s : sub.segments) {
    if (update_cost_) evaluator.computeCost(s);
    if (update_value_) evaluator.computeValue(s);
  }
  follow(segments, alive);
}

ModuleFactoryRegistry::Registration<SimulatedSensorUpdater> SimulatedSensorUpdater::registration(
    "SimulatedSensorUpdater");
void SimulatedSensorUpdater::setupFromParamMap(Module::ParamMap* param_map) {
  ChainedUpdater::setupFromParamMap(param_map);
  evaluator_ = dynamic_cast<trajectory_evaluator::SimulatedSensorEvaluator*>(
      planner_.getFactory().readLinkableModule("SimulatedSensorEvaluator"));
}
void SimulatedSensorUpdater::updateSegments(const std::vector<TrajectorySegment*>& segments,
                                            std::vector<uint8_t>* alive) {
  // the gain of every stored voxel list against the current map, without casting a ray: lists that
  // live on the device in one launch, host-side lists one by one
  std::vector<TrajectorySegment*> on_device;
  for (size_t i = 0; i < segments.size(); ++i) {
    if (!(*alive)[i] || !evaluator_) continue;
    if (dynamic_cast<trajectory_evaluator::DeviceSensorInfo*>(segments[i]->info.get()))
      on_device.push_back(segments[i]);
    else
      evaluator_->computeGainFromVisibleVoxels(segments[i]);
  }
  if (!on_device.empty()) evaluator_->updateGains(on_device);
  follow(segments, alive);
}

}  // namespace evaluator_updater
}  // namespace active_3d_planning
