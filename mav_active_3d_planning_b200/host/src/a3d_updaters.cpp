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

ModuleFactoryRegistry::Registration<YawPlanningUpdateAdapter> YawPlanningUpdateAdapter::registration(
    "YawPlanningUpdateAdapter");
void YawPlanningUpdateAdapter::setupFromParamMap(Module::ParamMap* param_map) {
  setParam<bool>(param_map, "dynamic_trajectories", &p_dynamic_trajectories_, false);
  ChainedUpdater::setupFromParamMap(param_map);
}
void YawPlanningUpdateAdapter::updateSegments(const std::vector<TrajectorySegment*>& segments,
                                              std::vector<uint8_t>* alive) {
  // the active orientation of every segment, brought up to date with its segment, as ONE batch for the
  // following updater; what that one decides about a copy does not remove the segment (:40 returns true)
  std::vector<TrajectorySegment*> owners, active;
  for (size_t i = 0; i < segments.size(); ++i) {
    if (!(*alive)[i]) continue;
    auto* info = dynamic_cast<trajectory_evaluator::YawPlanningInfo*>(segments[i]->info.get());
    if (!info || info->orientations.empty()) continue;
    TrajectorySegment& copy = info->orientations[info->active_orientation];
    if (p_dynamic_trajectories_) copy.trajectory = segments[i]->trajectory;
    copy.parent = segments[i]->parent;
    owners.push_back(segments[i]);
    active.push_back(&copy);
  }
  if (active.empty()) return;
  std::vector<uint8_t> flags(active.size(), 1);
  following_updater_->updateSegments(active, &flags);
  for (size_t k = 0; k < active.size(); ++k) {
    owners[k]->gain = active[k]->gain;
    owners[k]->cost = active[k]->cost;
    owners[k]->value = active[k]->value;
  }
}

ModuleFactoryRegistry::Registration<YawPlanningUpdater> YawPlanningUpdater::registration("YawPlanningUpdater");
void YawPlanningUpdater::setupFromParamMap(Module::ParamMap* param_map) {
  setParam<bool>(param_map, "select_by_value", &p_select_by_value_, false);
  setParam<bool>(param_map, "dynamic_trajectories", &p_dynamic_trajectories_, false);
  setParam<double>(param_map, "update_range", &p_update_range_, -1.0);
  setParam<double>(param_map, "update_gain", &p_update_gain_, 0.0);
  ChainedUpdater::setupFromParamMap(param_map);
  evaluator_ = dynamic_cast<trajectory_evaluator::YawPlanningEvaluator*>(
      planner_.getFactory().readLinkableModule("YawPlanningEvaluator"));
}
void YawPlanningUpdater::updateSegments(const std::vector<TrajectorySegment*>& segments,
                                        std::vector<uint8_t>* alive) {
  using trajectory_evaluator::YawPlanningInfo;
  if (!evaluator_) {
    planner_.printError("'YawPlanningUpdater' requires a yaw planning evaluator!");
    follow(segments, alive);
    return;
  }
  // pass 1: copies follow a segment that changed; the copies to evaluate again, of all segments
  const Eigen::Vector3d here = planner_.getCurrentPosition();
  std::vector<TrajectorySegment*> owners, again;
  for (size_t i = 0; i < segments.size(); ++i) {
    TrajectorySegment* seg = segments[i];
    if (!(*alive)[i] || !seg->parent) continue;
    auto* info = dynamic_cast<YawPlanningInfo*>(seg->info.get());
    if (!info || info->orientations.empty()) continue;
    owners.push_back(seg);
    const TrajectorySegment& act = info->orientations[info->active_orientation];
    if (p_dynamic_trajectories_ && (act.gain != seg->gain || act.parent != seg->parent)) {
      for (TrajectorySegment& o : info->orientations) {
        const double yaw = o.trajectory.back().getYaw();
        o.parent = seg->parent;
        o.trajectory = seg->trajectory;
        evaluator_->turnOrientation(&o, seg->trajectory.front().getYaw(), yaw);
      }
    }
    if ((here - seg->trajectory.back().position_W).norm() <= p_update_range_)
      for (TrajectorySegment& o : info->orientations)
        if (o.gain > p_update_gain_) again.push_back(&o);
  }
  // one cast for all of them
  evaluator_->computeOrientationGains(again);
  // pass 2: the best orientation becomes the segment (first one wins ties; the first copy's cost and
  // value are taken as they are, :86-90)
  TrajectoryEvaluator* following = evaluator_->followingEvaluator();
  for (TrajectorySegment* seg : owners) {
    auto* info = static_cast<YawPlanningInfo*>(seg->info.get());
    int best = 0;
    double best_score = p_select_by_value_ ? info->orientations[0].value : info->orientations[0].gain;
    for (size_t k = 1; k < info->orientations.size(); ++k) {
      TrajectorySegment& o = info->orientations[k];
      if (p_select_by_value_) {
        following->computeCost(&o);
        following->computeValue(&o);
      }
      const double score = p_select_by_value_ ? o.value : o.gain;
      if (score > best_score) {
        best = static_cast<int>(k);
        best_score = score;
      }
    }
    info->active_orientation = best;
    const TrajectorySegment& chosen = info->orientations[best];
    seg->trajectory = chosen.trajectory;
    seg->gain = chosen.gain;
    if (p_select_by_value_) {
      seg->cost = chosen.cost;
      seg->value = chosen.value;
    }
  }
  follow(segments, alive);
}

}  // namespace evaluator_updater
}  // namespace active_3d_planning
