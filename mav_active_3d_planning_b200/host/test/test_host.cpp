// test_host.cpp — driver for the C++ host mirror.
//   test_host config <planner.yaml> [namespace]     CPU only: YAML → ParamMaps as the ROS factory would
//                                                    build them, registered module types
//   test_host gains <planner.yaml> <map.bin> <segments.bin> <out.txt>
//                                                    GPU: build VoxbloxMap + the configured evaluator
//                                                    from the YAML, upload the map dump, evaluate the
//                                                    segments one by one (reference API) and as a batch
#include <cstdio>
#include <bitset>
#include <map>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>

#include "../../../include/a3d.h"
#include "a3d_host/trajectory_evaluator.h"
#include "a3d_host/voxblox_adapter.h"
#include "a3d_host/evaluator_updater.h"
#include "a3d_host/layer_msg.h"
#include "a3d_host/yaml_factory.h"

using namespace active_3d_planning;  // NOLINT

namespace {

class TestPlanner : public PlannerI {
 public:
  explicit TestPlanner(const YamlTree& tree) : factory_(tree) {}
  void build(const std::string& map_ns, const std::string& eval_ns) {
    map_ = factory_.createModule<Map>(map_ns, *this, true);
    evaluator_ = factory_.createModule<TrajectoryEvaluator>(eval_ns, *this, true);
  }
  const Eigen::Vector3d& getCurrentPosition() const override { return position_; }
  const Eigen::Quaterniond& getCurrentOrientation() const override { return orientation_; }
  TrajectoryEvaluator& getTrajectoryEvaluator() override { return *evaluator_; }
  ModuleFactory& getFactory() override { return factory_; }
  Map& getMap() override { return *map_; }
  SystemConstraints& getSystemConstraints() override { return constraints_; }
  void printInfo(const std::string& t) override { std::cout << "[info] " << t << "\n"; }
  void printWarning(const std::string& t) override { std::cout << "[warn] " << t << "\n"; }
  void printError(const std::string& t) override { std::cout << "[error] " << t << "\n"; }
  ModuleFactoryYaml factory_;
  std::unique_ptr<Map> map_;
  std::unique_ptr<TrajectoryEvaluator> evaluator_;
  SystemConstraints constraints_;
  Eigen::Vector3d position_;
  Eigen::Quaterniond orientation_;
};

template <typename T>
bool readVec(std::ifstream& f, std::vector<T>* v, size_t n) {
  v->resize(n);
  f.read(reinterpret_cast<char*>(v->data()), static_cast<std::streamsize>(n * sizeof(T)));
  return static_cast<bool>(f);
}

uint64_t digestOf(const std::vector<Eigen::Vector3d>& v) {  // DESIGN.md "digest"
  auto rotl = [](uint64_t x, int r) { return (x << r) | (x >> (64 - r)); };
  uint64_t h = 0;
  for (const auto& c : v) {
    uint64_t b[3];
    for (int k = 0; k < 3; ++k) {
      const double d = c[k];
      std::memcpy(&b[k], &d, 8);
    }
    uint64_t e = b[0] * 0x9E3779B97F4A7C15ull;
    e ^= rotl(b[1], 23) * 0xC2B2AE3D27D4EB4Full;
    e ^= rotl(b[2], 47) * 0x165667B19E3779F9ull;
    e ^= e >> 31;
    h = h * 0x100000001B3ull + e;
  }
  return h;
}

int runConfig(int argc, char** argv) {
  YamlTree tree;
  std::string err;
  if (!tree.loadFile(argv[2], &err)) {
    std::cerr << err << "\n";
    return 2;
  }
  const std::string ns = argc > 3 ? argv[3] : "/trajectory_evaluator";
  ModuleFactoryYaml factory(tree);
  Module::ParamMap map;
  std::string type = "NoDefaultTypeAvailable";
  factory.getParamMapAndType(&map, &type, ns);
  std::cout << "namespace " << ns << " type " << type << "\n";
  for (const auto& kv : map)
    if (kv.first != "verbose_text") std::cout << "param " << kv.first << " = " << kv.second << "\n";
  Module::ParamMap smap;
  std::string stype = "NoDefaultTypeAvailable";
  factory.getParamMapAndType(&smap, &stype, ns + "/sensor_model");
  std::cout << "namespace " << ns << "/sensor_model type " << stype << "\n";
  for (const auto& kv : smap)
    if (kv.first != "verbose_text") std::cout << "param " << kv.first << " = " << kv.second << "\n";
  std::cout << "registered " << ModuleFactoryRegistry::listTypes() << "\n";
  return 0;
}

int runGains(char** argv) {
  YamlTree tree;
  std::string err;
  if (!tree.loadFile(argv[2], &err)) {
    std::cerr << err << "\n";
    return 2;
  }
  std::ifstream mf(argv[3], std::ios::binary), sf(argv[4], std::ios::binary);
  if (!mf || !sf) {
    std::cerr << "cannot open inputs\n";
    return 2;
  }
  int32_t vps = 0, n_blocks = 0;
  float vs = 0;
  mf.read(reinterpret_cast<char*>(&vps), 4);
  mf.read(reinterpret_cast<char*>(&vs), 4);
  mf.read(reinterpret_cast<char*>(&n_blocks), 4);
  const size_t nv = static_cast<size_t>(vps) * vps * vps;
  std::vector<int32_t> idx;
  std::vector<float> dist, weight;
  std::vector<uint8_t> obs;
  if (!readVec(mf, &idx, 3 * static_cast<size_t>(n_blocks)) || !readVec(mf, &dist, nv * n_blocks) ||
      !readVec(mf, &obs, nv * n_blocks) || !readVec(mf, &weight, nv * n_blocks)) {
    std::cerr << "short map file\n";
    return 2;
  }
  // map parameters live in the map namespace of the config store, like the voxblox server params
  tree.set("/map/type", "VoxbloxMap");
  tree.set("/map/tsdf_voxel_size", std::to_string(vs));
  tree.set("/map/tsdf_voxels_per_side", std::to_string(vps));

  TestPlanner planner(tree);
  planner.build("/map", "/trajectory_evaluator");
  auto* vmap = dynamic_cast<map::VoxbloxMap*>(planner.map_.get());
  auto* eval = dynamic_cast<trajectory_evaluator::SimulatedSensorEvaluator*>(planner.evaluator_.get());
  if (!vmap || !eval) {
    std::cerr << "config does not name a VoxbloxMap / SimulatedSensorEvaluator\n";
    return 2;
  }
  vmap->uploadBlocks(n_blocks, idx.data(), dist.data(), obs.data(), weight.data());

  int32_t n_seg = 0, pts = 0;
  sf.read(reinterpret_cast<char*>(&n_seg), 4);
  sf.read(reinterpret_cast<char*>(&pts), 4);
  TrajectorySegment root;
  std::vector<TrajectorySegment*> segments;
  for (int s = 0; s < n_seg; ++s) {
    TrajectorySegment* seg = (s == 0) ? &root : segments.back()->spawnChild();  // a chain: parents matter
    for (int p = 0; p < pts; ++p) {
      int64_t t;
      double v[7];
      sf.read(reinterpret_cast<char*>(&t), 8);
      sf.read(reinterpret_cast<char*>(v), 56);
      EigenTrajectoryPoint tp;
      tp.time_from_start_ns = t;
      tp.position_W = Eigen::Vector3d(v[0], v[1], v[2]);
      tp.orientation_W_B = Eigen::Quaterniond(v[3], v[4], v[5], v[6]);
      seg->trajectory.push_back(tp);
    }
    segments.push_back(seg);
  }
  std::ofstream out(argv[5]);
  out.precision(17);
  auto* cam = dynamic_cast<sensor_model::CameraModel*>(eval->sensorModel());
  out << "grid " << cam->rayGridX() << " " << cam->rayGridY() << " " << cam->numSections() << "\n";
  // reference API: one segment per call, voxel list stored in the segment
  for (int s = 0; s < n_seg; ++s) {
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = eval->computeGain(segments[s]);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    out << "# computeGain " << s << " " << ms << " ms\n";
    auto* info = dynamic_cast<trajectory_evaluator::SimulatedSensorInfo*>(segments[s]->info.get());
    out << "single " << s << " " << ok << " " << segments[s]->gain << " "
        << (info ? info->visible_voxels.size() : 0) << " " << (info ? digestOf(info->visible_voxels) : 0)
        << "\n";
  }
  // updater path: fold the stored lists again
  for (int s = 0; s < n_seg; ++s) {
    const double before = segments[s]->gain;
    eval->computeGainFromVisibleVoxels(segments[s]);
    out << "refold " << s << " " << (before == segments[s]->gain) << "\n";
  }
  // sensor model alone
  std::vector<Eigen::Vector3d> lst;
  eval->sensorModel()->getVisibleVoxelsFromTrajectory(&lst, *segments[0]);
  out << "sensor 0 " << lst.size() << " " << digestOf(lst) << "\n";
  // batch API
  std::vector<double> single(n_seg);
  for (int s = 0; s < n_seg; ++s) single[s] = segments[s]->gain;
  const bool bok = eval->computeGains(segments);
  for (int s = 0; s < n_seg; ++s) out << "batch " << s << " " << bok << " " << segments[s]->gain << "\n";
  // device-resident lists (batch) and their re-fold; then free them by dropping the infos
  if (bok) {
    std::vector<double> batch(n_seg);
    for (int s = 0; s < n_seg; ++s) batch[s] = segments[s]->gain;
    const bool sok = eval->computeGainsStored(segments);
    for (int s = 0; s < n_seg; ++s) {
      auto* di = dynamic_cast<trajectory_evaluator::DeviceSensorInfo*>(segments[s]->info.get());
      out << "stored " << s << " " << sok << " " << (segments[s]->gain == batch[s]) << " "
          << (di ? di->n_voxels : 0) << "\n";
    }
    const bool uok = eval->updateGains(segments);
    for (int s = 0; s < n_seg; ++s) out << "update " << s << " " << uok << " " << segments[s]->gain << "\n";
    eval->computeGainFromVisibleVoxels(segments[0]);  // single-segment updater path on a device list
    out << "update_single 0 " << segments[0]->gain << "\n";
    for (int s = 0; s < n_seg; ++s) segments[s]->info.reset();
  }
  // map virtuals
  Eigen::Vector3d c;
  vmap->getVoxelCenter(&c, segments[0]->trajectory.back().position_W);
  out << "center " << c.x() << " " << c.y() << " " << c.z() << " state "
      << static_cast<int>(vmap->getVoxelState(segments[0]->trajectory.back().position_W)) << " observed "
      << vmap->isObserved(segments[0]->trajectory.back().position_W) << " voxel_size " << vmap->getVoxelSize()
      << "\n";
  for (const auto& line : planner.factory_.verboseLog) out << "# " << line.substr(0, line.find('\n')) << "\n";
  return 0;
}

// Stand-ins for the cost / value / updater modules (outside this path) so that select_by_value,
// ContinuousYawPlanningEvaluator::setBestYaw and updateSegment can run.
class TestSegmentLength : public CostComputer {
 public:
  explicit TestSegmentLength(PlannerI& planner) : CostComputer(planner) {}
  void setupFromParamMap(Module::ParamMap*) override {}
  bool computeCost(TrajectorySegment* t) override {
    t->cost = 1.0 + t->trajectory.back().getYaw();  // depends on the applied yaw
    return true;
  }
  static ModuleFactoryRegistry::Registration<TestSegmentLength> registration;
};
ModuleFactoryRegistry::Registration<TestSegmentLength> TestSegmentLength::registration("TestSegmentLength");
class TestLinearValue : public ValueComputer {
 public:
  explicit TestLinearValue(PlannerI& planner) : ValueComputer(planner) {}
  void setupFromParamMap(Module::ParamMap*) override {}
  bool computeValue(TrajectorySegment* t) override {
    t->value = t->gain - 100.0 * t->cost;
    return true;
  }
  static ModuleFactoryRegistry::Registration<TestLinearValue> registration;
};
ModuleFactoryRegistry::Registration<TestLinearValue> TestLinearValue::registration("TestLinearValue");
class TestUpdater : public EvaluatorUpdater {
 public:
  explicit TestUpdater(PlannerI& planner) : EvaluatorUpdater(planner) {}
  void setupFromParamMap(Module::ParamMap*) override {}
  void updateSegments(const std::vector<TrajectorySegment*>&, std::vector<uint8_t>*) override {}
  static ModuleFactoryRegistry::Registration<TestUpdater> registration;
};
ModuleFactoryRegistry::Registration<TestUpdater> TestUpdater::registration("TestUpdater");

bool loadMap(const char* path, YamlTree* tree, int32_t* vps, int32_t* n_blocks, std::vector<int32_t>* idx,
             std::vector<float>* dist, std::vector<uint8_t>* obs, std::vector<float>* weight) {
  std::ifstream mf(path, std::ios::binary);
  if (!mf) return false;
  float vs = 0;
  mf.read(reinterpret_cast<char*>(vps), 4);
  mf.read(reinterpret_cast<char*>(&vs), 4);
  mf.read(reinterpret_cast<char*>(n_blocks), 4);
  const size_t nv = static_cast<size_t>(*vps) * *vps * *vps;
  if (!readVec(mf, idx, 3 * static_cast<size_t>(*n_blocks)) || !readVec(mf, dist, nv * *n_blocks) ||
      !readVec(mf, obs, nv * *n_blocks) || !readVec(mf, weight, nv * *n_blocks))
    return false;
  tree->set("/map/type", "VoxbloxMap");
  tree->set("/map/tsdf_voxel_size", std::to_string(vs));
  tree->set("/map/tsdf_voxels_per_side", std::to_string(*vps));
  return true;
}

// a chain of segments (parents matter for updateSegment) read from the segments file
void loadSegments(const char* path, TrajectorySegment* root, std::vector<TrajectorySegment*>* segments) {
  std::ifstream sf(path, std::ios::binary);
  int32_t n_seg = 0, pts = 0;
  sf.read(reinterpret_cast<char*>(&n_seg), 4);
  sf.read(reinterpret_cast<char*>(&pts), 4);
  for (int s = 0; s < n_seg; ++s) {
    TrajectorySegment* seg = (s == 0) ? root : segments->back()->spawnChild();
    for (int p = 0; p < pts; ++p) {
      int64_t t;
      double v[7];
      sf.read(reinterpret_cast<char*>(&t), 8);
      sf.read(reinterpret_cast<char*>(v), 56);
      EigenTrajectoryPoint tp;
      tp.time_from_start_ns = t;
      tp.position_W = Eigen::Vector3d(v[0], v[1], v[2]);
      tp.orientation_W_B = Eigen::Quaterniond(v[3], v[4], v[5], v[6]);
      seg->trajectory.push_back(tp);
    }
    segments->push_back(seg);
  }
}

void printYaw(std::ofstream& out, const char* tag, int s, const TrajectorySegment& seg) {
  auto* info = dynamic_cast<trajectory_evaluator::YawPlanningInfo*>(seg.info.get());
  out << tag << " " << s << " " << seg.gain << " " << seg.cost << " " << seg.value << " "
      << (info ? info->active_orientation : -1) << " " << seg.trajectory.back().getYaw() << " "
      << seg.trajectory.front().getYaw();
  if (info)
    for (const auto& o : info->orientations) out << " " << o.gain;
  out << "\n";
}

// yaw planning: reference call (one segment per computeGain), the batch call, updateSegment
int runYaw(char** argv) {
  YamlTree tree;
  std::string err;
  if (!tree.loadFile(argv[2], &err)) {
    std::cerr << err << "\n";
    return 2;
  }
  int32_t vps = 0, n_blocks = 0;
  std::vector<int32_t> idx;
  std::vector<float> dist, weight;
  std::vector<uint8_t> obs;
  if (!loadMap(argv[3], &tree, &vps, &n_blocks, &idx, &dist, &obs, &weight)) {
    std::cerr << "cannot read map\n";
    return 2;
  }
  TestPlanner planner(tree);
  planner.build("/map", "/trajectory_evaluator");
  auto* vmap = dynamic_cast<map::VoxbloxMap*>(planner.map_.get());
  auto* eval = dynamic_cast<trajectory_evaluator::YawPlanningEvaluator*>(planner.evaluator_.get());
  if (!vmap || !eval) {
    std::cerr << "config does not name a VoxbloxMap / YawPlanningEvaluator\n";
    return 2;
  }
  vmap->uploadBlocks(n_blocks, idx.data(), dist.data(), obs.data(), weight.data());
  std::ofstream out(argv[5]);
  out.precision(17);
  {
    TrajectorySegment root;
    std::vector<TrajectorySegment*> segments;
    loadSegments(argv[4], &root, &segments);
    for (size_t s = 0; s < segments.size(); ++s) {
      eval->computeGain(segments[s]);
      printYaw(out, "yaw", static_cast<int>(s), *segments[s]);
    }
    // updater path on an unchanged map: same orientation gains, same choice
    planner.position_ = segments.back()->trajectory.back().position_W;
    for (size_t s = 0; s < segments.size(); ++s) {
      const bool ok = eval->updateSegment(segments[s]);
      out << "ok " << ok << "\n";
      printYaw(out, "yupdate", static_cast<int>(s), *segments[s]);
    }
  }
  {
    TrajectorySegment root;
    std::vector<TrajectorySegment*> segments;
    loadSegments(argv[4], &root, &segments);
    const bool ok = eval->computeGains(segments);
    out << "ok " << ok << "\n";
    for (size_t s = 0; s < segments.size(); ++s) printYaw(out, "ybatch", static_cast<int>(s), *segments[s]);
    auto* info = dynamic_cast<trajectory_evaluator::YawPlanningInfo*>(segments[0]->info.get());
    auto* di = info ? dynamic_cast<trajectory_evaluator::DeviceSensorInfo*>(info->orientations[0].info.get())
                    : nullptr;
    out << "device_list " << (di ? di->n_voxels : 0) << "\n";
  }
  return 0;
}

// ---- a mock of the part of voxblox's layer API the adapter uses (same names and signatures) -------
namespace mockvbx {
struct Update {
  enum Status { kMap, kMesh, kEsdf, kCount };
};
struct BlockIndex {
  int v[3];
  int x() const { return v[0]; }
  int y() const { return v[1]; }
  int z() const { return v[2]; }
  bool operator<(const BlockIndex& o) const {
    return std::lexicographical_compare(v, v + 3, o.v, o.v + 3);
  }
};
typedef std::vector<BlockIndex> BlockIndexList;
struct EsdfVoxel {
  float distance = 0.0f;
  bool observed = false;
  bool hallucinated = false, in_queue = false, fixed = false;
};
struct TsdfVoxel {
  float distance = 0.0f;
  float weight = 0.0f;
};
template <class V>
struct Block {
  std::vector<V> voxels;
  std::bitset<Update::kCount> upd;
  size_t num_voxels() const { return voxels.size(); }
  const V& getVoxelByLinearIndex(size_t i) const { return voxels[i]; }
  V& getVoxelByLinearIndex(size_t i) { return voxels[i]; }
  std::bitset<Update::kCount>& updated() { return upd; }
};
template <class V>
struct Layer {
  float vs = 0.1f;
  size_t vps = 16;
  std::map<BlockIndex, std::shared_ptr<Block<V>>> blocks;
  size_t voxels_per_side() const { return vps; }
  float voxel_size() const { return vs; }
  void getAllUpdatedBlocks(Update::Status bit, BlockIndexList* out) const {
    for (const auto& kv : blocks)
      if (kv.second->upd[bit]) out->push_back(kv.first);
  }
  Block<V>& getBlockByIndex(const BlockIndex& i) { return *blocks.at(i); }
  std::shared_ptr<Block<V>> getBlockPtrByIndex(const BlockIndex& i) {
    auto it = blocks.find(i);
    return it == blocks.end() ? nullptr : it->second;
  }
};
}  // namespace mockvbx

// the voxblox layer adapter: fill mock layers from the map file, sync, evaluate, touch one block, sync
int runAdapter(char** argv) {
  YamlTree tree;
  std::string err;
  if (!tree.loadFile(argv[2], &err)) {
    std::cerr << err << "\n";
    return 2;
  }
  int32_t vps = 0, n_blocks = 0;
  std::vector<int32_t> idx;
  std::vector<float> dist, weight;
  std::vector<uint8_t> obs;
  if (!loadMap(argv[3], &tree, &vps, &n_blocks, &idx, &dist, &obs, &weight)) {
    std::cerr << "cannot read map\n";
    return 2;
  }
  const size_t nv = static_cast<size_t>(vps) * vps * vps;
  mockvbx::Layer<mockvbx::EsdfVoxel> esdf;
  mockvbx::Layer<mockvbx::TsdfVoxel> tsdf;
  esdf.vps = tsdf.vps = static_cast<size_t>(vps);
  for (int b = 0; b < n_blocks; ++b) {
    const mockvbx::BlockIndex bi{{idx[3 * b], idx[3 * b + 1], idx[3 * b + 2]}};
    auto eb = std::make_shared<mockvbx::Block<mockvbx::EsdfVoxel>>();
    eb->voxels.resize(nv);
    for (size_t i = 0; i < nv; ++i) {
      eb->voxels[i].distance = dist[b * nv + i];
      eb->voxels[i].observed = obs[b * nv + i] != 0;
    }
    eb->upd.set(mockvbx::Update::kMap);
    eb->upd.set(mockvbx::Update::kMesh);
    esdf.blocks[bi] = eb;
    if (b % 3 != 0) {  // the TSDF layer lacks some blocks: weight 0 there
      auto tb = std::make_shared<mockvbx::Block<mockvbx::TsdfVoxel>>();
      tb->voxels.resize(nv);
      for (size_t i = 0; i < nv; ++i) tb->voxels[i].weight = weight[b * nv + i];
      tsdf.blocks[bi] = tb;
    }
  }
  TestPlanner planner(tree);
  planner.build("/map", "/trajectory_evaluator");
  auto* vmap = dynamic_cast<map::VoxbloxMap*>(planner.map_.get());
  auto* eval = dynamic_cast<trajectory_evaluator::SimulatedSensorEvaluator*>(planner.evaluator_.get());
  if (!vmap || !eval) return 2;
  std::ofstream out(argv[5]);
  out.precision(17);
  const int sent = map::syncUpdatedBlocks<mockvbx::BlockIndexList>(vmap, &esdf, &tsdf, mockvbx::Update::kMap);
  size_t still_map = 0, still_mesh = 0;
  for (const auto& kv : esdf.blocks) {
    still_map += kv.second->upd[mockvbx::Update::kMap];
    still_mesh += kv.second->upd[mockvbx::Update::kMesh];
  }
  out << "sent " << sent << " blocks " << vmap->numBlocks() << " flags_map " << still_map << " flags_mesh "
      << still_mesh << "\n";
  out << "again " << map::syncUpdatedBlocks<mockvbx::BlockIndexList>(vmap, &esdf, &tsdf, mockvbx::Update::kMap) << "\n";
  TrajectorySegment root;
  std::vector<TrajectorySegment*> segments;
  loadSegments(argv[4], &root, &segments);
  const bool bok = eval->computeGains(segments);
  for (size_t s = 0; s < segments.size(); ++s) out << "batch " << s << " " << bok << " " << segments[s]->gain << "\n";
  // weights: present block vs block missing in the TSDF layer
  const Eigen::Vector3d p = segments[0]->trajectory.back().position_W;
  out << "weight " << vmap->getVoxelWeight(p) << "\n";
  // one block changes: everything becomes observed free space at 1 m
  auto& first = *esdf.blocks.begin();
  for (auto& v : first.second->voxels) {
    v.distance = 1.0f;
    v.observed = true;
  }
  first.second->upd.set(mockvbx::Update::kMap);
  out << "delta " << map::syncUpdatedBlocks<mockvbx::BlockIndexList>(vmap, &esdf, &tsdf, mockvbx::Update::kMap) << " "
      << first.first.x() << " " << first.first.y() << " " << first.first.z() << "\n";
  eval->computeGains(segments);
  for (size_t s = 0; s < segments.size(); ++s) out << "batch2 " << s << " 1 " << segments[s]->gain << "\n";
  return 0;
}

// ---- device-free: the updater chain and updateTree on a hand-made tree ----------------------------
// An evaluator without a sensor: the gain of a segment is the x coordinate of its end point times a
// factor that the test changes between evaluation and update ("the map changed").
class TestPlainEvaluator : public TrajectoryEvaluator {
 public:
  explicit TestPlainEvaluator(PlannerI& planner) : TrajectoryEvaluator(planner) {}
  bool computeGain(TrajectorySegment* t) override {
    ++n_gain_calls;
    t->gain = factor * t->trajectory.back().position_W.x();
    return true;
  }
  static double factor;
  static int n_gain_calls;
  static ModuleFactoryRegistry::Registration<TestPlainEvaluator> registration;
};
double TestPlainEvaluator::factor = 1.0;
int TestPlainEvaluator::n_gain_calls = 0;
ModuleFactoryRegistry::Registration<TestPlainEvaluator> TestPlainEvaluator::registration("TestPlainEvaluator");

void printTree(std::ostream& out, const TrajectorySegment& s, const std::string& path) {
  out << "node " << (path.empty() ? "root" : path) << " " << s.gain << " " << s.cost << " " << s.value << "\n";
  for (size_t i = 0; i < s.children.size(); ++i)
    printTree(out, *s.children[i], path + (path.empty() ? "" : ".") + std::to_string(static_cast<int>(
                                              s.children[i]->trajectory.back().position_W.y())));
}

// tree <yaml>: root at x = 0; children named by the y coordinate of their end point:
//   1 (x=5) -> 11 (x=6) -> 111 (x=1)       2 (x=2) -> 21 (x=9)       3 (x=40, far away) -> 31 (x=8)
int runTree(char** argv) {
  YamlTree tree;
  std::string err;
  if (!tree.loadFile(argv[2], &err)) {
    std::cerr << err << "\n";
    return 2;
  }
  TestPlanner planner(tree);
  planner.evaluator_ = planner.factory_.createModule<TrajectoryEvaluator>("/trajectory_evaluator", planner, true);
  TrajectoryEvaluator* eval = planner.evaluator_.get();
  TrajectorySegment root;
  auto point = [](double x, double y) {
    EigenTrajectoryPoint p;
    p.position_W = Eigen::Vector3d(x, y, 0.0);
    p.orientation_W_B = Eigen::Quaterniond(1, 0, 0, 0);
    return p;
  };
  root.trajectory.push_back(point(0, 0));
  auto add = [&](TrajectorySegment* parent, double x, int name) {
    TrajectorySegment* c = parent->spawnChild();
    c->trajectory.push_back(parent->trajectory.back());
    c->trajectory.push_back(point(x, name));
    return c;
  };
  TrajectorySegment* n1 = add(&root, 5, 1);
  TrajectorySegment* n11 = add(n1, 6, 11);
  add(n11, 1, 111);
  TrajectorySegment* n2 = add(&root, 2, 2);
  add(n2, 9, 21);
  TrajectorySegment* n3 = add(&root, 40, 3);
  add(n3, 8, 31);
  std::vector<TrajectorySegment*> all;
  root.getTree(&all);
  for (TrajectorySegment* s : all) {
    if (s == &root) continue;
    eval->computeGain(s);
    eval->computeCost(s);
    eval->computeValue(s);
  }
  printTree(std::cout, root, "");
  TestPlainEvaluator::factor = 2.0;
  TestPlainEvaluator::n_gain_calls = 0;
  planner.position_ = Eigen::Vector3d(0, 0, 0);
  const int removed = eval->updateTree(&root);
  std::cout << "removed " << removed << " gain_calls " << TestPlainEvaluator::n_gain_calls << "\n";
  printTree(std::cout, root, "");
  // the per-segment contract on a survivor and on a segment the chain drops
  std::cout << "single " << eval->updateSegment(root.children[0].get()) << "\n";
  return 0;
}

// ---- device-free: the yaw-planning updaters (yaw_planning_updaters.cpp) ------------------------------
// A following evaluator whose gain depends on where the segment looks at its end: 10 + 5 cos(yaw - phase),
// times a factor; the test turns the phase between evaluation and update ("what is worth seeing moved").
class TestYawEvaluator : public TrajectoryEvaluator {
 public:
  explicit TestYawEvaluator(PlannerI& planner) : TrajectoryEvaluator(planner) {}
  bool computeGain(TrajectorySegment* t) override {
    ++n_gain_calls;
    t->gain = factor * (10.0 + 5.0 * std::cos(t->trajectory.back().getYaw() - phase));
    return true;
  }
  static double factor, phase;
  static int n_gain_calls;
  static ModuleFactoryRegistry::Registration<TestYawEvaluator> registration;
};
double TestYawEvaluator::factor = 1.0;
double TestYawEvaluator::phase = 0.0;
int TestYawEvaluator::n_gain_calls = 0;
ModuleFactoryRegistry::Registration<TestYawEvaluator> TestYawEvaluator::registration("TestYawEvaluator");
// An updater that knows nothing about yaw planning: triples the gain of what it is given.
class TestTripleGain : public EvaluatorUpdater {
 public:
  explicit TestTripleGain(PlannerI& planner) : EvaluatorUpdater(planner) {}
  void setupFromParamMap(Module::ParamMap*) override {}
  void updateSegments(const std::vector<TrajectorySegment*>& segments, std::vector<uint8_t>* alive) override {
    ++n_batches;
    for (size_t i = 0; i < segments.size(); ++i)
      if ((*alive)[i]) segments[i]->gain *= 3.0;
  }
  static int n_batches;
  static ModuleFactoryRegistry::Registration<TestTripleGain> registration;
};
int TestTripleGain::n_batches = 0;
ModuleFactoryRegistry::Registration<TestTripleGain> TestTripleGain::registration("TestTripleGain");

// yawtree <yaml>: root at the origin, children 1 (ends at x=5), 2 (x=2, child 21 at x=3), 3 (x=40: out of reach)
int runYawTree(char** argv) {
  YamlTree tree;
  std::string err;
  if (!tree.loadFile(argv[2], &err)) {
    std::cerr << err << "\n";
    return 2;
  }
  TestPlanner planner(tree);
  planner.evaluator_ = planner.factory_.createModule<TrajectoryEvaluator>("/trajectory_evaluator", planner, true);
  TrajectoryEvaluator* eval = planner.evaluator_.get();
  TrajectorySegment root;
  auto point = [](double x, double y) {
    EigenTrajectoryPoint p;
    p.position_W = Eigen::Vector3d(x, y, 0.0);
    p.orientation_W_B = Eigen::Quaterniond(1, 0, 0, 0);
    return p;
  };
  root.trajectory.push_back(point(0, 0));
  auto add = [&](TrajectorySegment* parent, double x, int name) {
    TrajectorySegment* c = parent->spawnChild();
    c->trajectory.push_back(parent->trajectory.back());
    c->trajectory.push_back(point(x, name));
    return c;
  };
  add(&root, 5, 1);
  add(add(&root, 2, 2), 3, 21);
  add(&root, 40, 3);
  std::vector<TrajectorySegment*> all;
  root.getTree(&all);
  auto print = [&]() {
    for (TrajectorySegment* s : all) {
      if (s == &root) continue;
      auto* info = dynamic_cast<trajectory_evaluator::YawPlanningInfo*>(s->info.get());
      std::cout << "node " << static_cast<int>(s->trajectory.back().position_W.y()) << " " << s->gain << " " << s->cost
                << " " << s->value << " " << s->trajectory.back().getYaw() << " " << (info ? info->active_orientation : -1)
                << " " << (info ? info->orientations[info->active_orientation].gain : 0.0) << "\n";
    }
  };
  std::cout.precision(12);
  for (TrajectorySegment* s : all) {
    if (s == &root) continue;
    eval->computeGain(s);
    eval->computeCost(s);
    eval->computeValue(s);
  }
  std::cout << "gain_calls " << TestYawEvaluator::n_gain_calls << "\n";
  print();
  TestYawEvaluator::phase = M_PI / 2;
  TestYawEvaluator::factor = 2.0;
  TestYawEvaluator::n_gain_calls = 0;
  planner.position_ = Eigen::Vector3d(0, 0, 0);
  const int removed = eval->updateTree(&root);
  std::cout << "removed " << removed << " gain_calls " << TestYawEvaluator::n_gain_calls << " follower_batches "
            << TestTripleGain::n_batches << "\n";
  print();
  return 0;
}

// The planner's evaluator update step (online_planner.cpp:679-693) through the updater chain the YAML
// names: evaluate a chain of segments on the first map, overwrite blocks from the second map file,
// updateTree(), print what is left.  mode "stored": the segments' lists live on the device
// (computeGainsStored) — the SimulatedSensorUpdater then re-folds them in one launch.
int runUpdate(char** argv) {
  YamlTree tree;
  std::string err;
  if (!tree.loadFile(argv[2], &err)) {
    std::cerr << err << "\n";
    return 2;
  }
  int32_t vps = 0, n_blocks = 0;
  std::vector<int32_t> idx;
  std::vector<float> dist, weight;
  std::vector<uint8_t> obs;
  if (!loadMap(argv[3], &tree, &vps, &n_blocks, &idx, &dist, &obs, &weight)) return 2;
  TestPlanner planner(tree);
  planner.build("/map", "/trajectory_evaluator");
  auto* vmap = dynamic_cast<map::VoxbloxMap*>(planner.map_.get());
  auto* eval = planner.evaluator_.get();
  auto* sim = dynamic_cast<trajectory_evaluator::SimulatedSensorEvaluator*>(eval);
  if (!vmap || !sim) return 2;
  vmap->uploadBlocks(n_blocks, idx.data(), dist.data(), obs.data(), weight.data());
  std::ofstream out(argv[5]);
  out.precision(17);
  TrajectorySegment root;
  std::vector<TrajectorySegment*> segments;
  loadSegments(argv[4], &root, &segments);
  std::vector<TrajectorySegment*> below(segments.begin() + 1, segments.end());
  const bool stored = std::strcmp(argv[7], "stored") == 0;
  if (stored) {
    out << "ok " << sim->computeGainsStored(below) << "\n";
  } else {
    for (TrajectorySegment* s : below) sim->computeGain(s);
  }
  for (TrajectorySegment* s : below) {
    eval->computeCost(s);
    eval->computeValue(s);
  }
  for (size_t s = 0; s < below.size(); ++s)
    out << "before " << s << " " << below[s]->gain << " " << below[s]->cost << " " << below[s]->value << "\n";
  // the map changes
  YamlTree unused;
  int32_t vps2 = 0, n2 = 0;
  std::vector<int32_t> idx2;
  std::vector<float> dist2, weight2;
  std::vector<uint8_t> obs2;
  if (!loadMap(argv[6], &unused, &vps2, &n2, &idx2, &dist2, &obs2, &weight2)) return 2;
  vmap->uploadBlocks(n2, idx2.data(), dist2.data(), obs2.data(), weight2.data());
  planner.position_ = root.trajectory.back().position_W;
  sim->takeGainTimeMs();
  const uint64_t launches0 = a3d_kernel_launches(vmap->context());
  const int removed = eval->updateTree(&root);
  out << "removed " << removed << " launches " << (a3d_kernel_launches(vmap->context()) - launches0)
      << " gain_ms " << sim->takeGainTimeMs() << "\n";
  int depth = 0;
  for (TrajectorySegment* s = &root; !s->children.empty(); s = s->children[0].get(), ++depth) {
    const TrajectorySegment& c = *s->children[0];
    out << "after " << depth << " " << c.gain << " " << c.cost << " " << c.value << "\n";
  }
  // the per-segment contract on what is left (unchanged map: same numbers)
  for (TrajectorySegment* s = &root; !s->children.empty(); s = s->children[0].get())
    out << "single " << eval->updateSegment(s->children[0].get()) << " " << s->children[0]->gain << "\n";
  if (auto* cam = dynamic_cast<sensor_model::CameraModel*>(sim->sensorModel())) {
    double mean = 0, stddev = 0;
    size_t count = 0;
    cam->castTimeStats(&mean, &stddev, &count);
    out << "casttime " << mean << " " << stddev << " " << count << "\n";
  }
  return 0;
}

// voxblox_msgs/Layer messages → device mirror: the map file is encoded into ESDF and TSDF layer
// messages (encoder written from the message description in a3d_host/layer_msg.h), sent in the order a
// running system may deliver them (TSDF of some blocks first, ESDF in two messages, the remaining TSDF),
// then the evaluator of the YAML runs on the mirror.
int runLayer(char** argv) {
  YamlTree tree;
  std::string err;
  if (!tree.loadFile(argv[2], &err)) {
    std::cerr << err << "\n";
    return 2;
  }
  int32_t vps = 0, n_blocks = 0;
  std::vector<int32_t> idx;
  std::vector<float> dist, weight;
  std::vector<uint8_t> obs;
  if (!loadMap(argv[3], &tree, &vps, &n_blocks, &idx, &dist, &obs, &weight)) return 2;
  TestPlanner planner(tree);
  planner.build("/map", "/trajectory_evaluator");
  auto* vmap = dynamic_cast<map::VoxbloxMap*>(planner.map_.get());
  auto* sim = dynamic_cast<trajectory_evaluator::SimulatedSensorEvaluator*>(planner.evaluator_.get());
  if (!vmap || !sim) return 2;
  const size_t nv = static_cast<size_t>(vps) * vps * vps;
  auto encode = [&](const char* type, int b0, int b1, uint8_t action) {
    map::LayerMsg msg;
    msg.voxel_size = vmap->getVoxelSize();
    msg.voxels_per_side = static_cast<uint32_t>(vps);
    msg.layer_type = type;
    msg.action = action;
    const bool esdf = std::strcmp(type, "esdf") == 0;
    for (int b = b0; b < b1; ++b) {
      map::BlockMsg bm;
      bm.x_index = idx[3 * b];
      bm.y_index = idx[3 * b + 1];
      bm.z_index = idx[3 * b + 2];
      for (size_t i = 0; i < nv; ++i) {
        uint32_t w0, w1;
        std::memcpy(&w0, &dist[b * nv + i], 4);
        bm.data.push_back(w0);
        if (esdf) {
          // observed in the low byte, a parent direction in the upper three (ignored by the mirror)
          bm.data.push_back((obs[b * nv + i] ? 1u : 0u) | (static_cast<uint32_t>((i * 2654435761u) >> 8) << 8));
        } else {
          std::memcpy(&w1, &weight[b * nv + i], 4);
          bm.data.push_back(w1);
          bm.data.push_back(0x80FF40FFu);  // colour
        }
      }
      msg.blocks.push_back(std::move(bm));
    }
    return msg;
  };
  std::ofstream out(argv[5]);
  out.precision(17);
  map::LayerReceiver rx(vmap);
  rx.mirrorTsdfDistances(true);  // (the test's TSDF voxels carry the ESDF distance array as their distance)
  const int half = n_blocks / 2, third = n_blocks / 3;
  out << "tsdf_first " << rx.onTsdfLayer(encode("tsdf", 0, half, map::kLayerUpdate)) << " pending "
      << rx.pendingWeightBlocks() << "\n";
  out << "esdf_a " << rx.onEsdfLayer(encode("esdf", 0, third, map::kLayerUpdate)) << "\n";
  out << "esdf_b " << rx.onEsdfLayer(encode("esdf", third, n_blocks, map::kLayerUpdate)) << " pending "
      << rx.pendingWeightBlocks() << "\n";
  out << "tsdf_rest " << rx.onTsdfLayer(encode("tsdf", half, n_blocks, map::kLayerUpdate)) << " pending "
      << rx.pendingWeightBlocks() << "\n";
  out << "blocks " << vmap->numBlocks() << " received " << rx.blocksReceived() << "\n";
  // a message of another layer type / geometry is refused
  map::LayerMsg bad = encode("esdf", 0, 1, map::kLayerUpdate);
  bad.voxels_per_side += 1;
  out << "refused " << !rx.onEsdfLayer(bad) << " " << !rx.onTsdfLayer(encode("esdf", 0, 1, map::kLayerUpdate)) << "\n";
  TrajectorySegment root;
  std::vector<TrajectorySegment*> segments;
  loadSegments(argv[4], &root, &segments);
  sim->computeGains(segments);
  for (size_t s = 0; s < segments.size(); ++s) out << "batch " << s << " 1 " << segments[s]->gain << "\n";
  out << "weight " << vmap->getVoxelWeight(segments[0]->trajectory.back().position_W) << "\n";
  out << "tsdf_distance " << vmap->getVoxelDistance(segments[0]->trajectory.back().position_W) << " "
      << vmap->getVoxelDistance(Eigen::Vector3d(-50.0, -50.0, -50.0)) << "\n";
  // ACTION_RESET: the sender's layer replaces the mirror
  out << "reset " << rx.onEsdfLayer(encode("esdf", 0, third, map::kLayerReset)) << " blocks " << vmap->numBlocks() << "\n";
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    if (argc >= 3 && std::strcmp(argv[1], "config") == 0) return runConfig(argc, argv);
    if (argc >= 6 && std::strcmp(argv[1], "gains") == 0) return runGains(argv);
    if (argc >= 6 && std::strcmp(argv[1], "yaw") == 0) return runYaw(argv);
    if (argc >= 6 && std::strcmp(argv[1], "adapter") == 0) return runAdapter(argv);
    if (argc >= 8 && std::strcmp(argv[1], "update") == 0) return runUpdate(argv);
    if (argc >= 6 && std::strcmp(argv[1], "layer") == 0) return runLayer(argv);
    if (argc >= 3 && std::strcmp(argv[1], "tree") == 0) return runTree(argv);
    if (argc >= 3 && std::strcmp(argv[1], "yawtree") == 0) return runYawTree(argv);
  } catch (const std::exception& e) {
    std::cerr << "exception: " << e.what() << "\n";
    return 3;
  }
  std::cerr << "usage: test_host config <yaml> [ns] | tree <yaml> | yawtree <yaml> | gains|yaw|adapter|layer <yaml> <map.bin> <segments.bin> <out.txt> | "
               "update <yaml> <map.bin> <segments.bin> <out.txt> <map2.bin> stored|host\n";
  return 1;
}
