// layer_msg.h — voxblox_msgs/Layer → device mirror (the wire half).
//
// In the reference the planner node receives the map as voxblox_msgs/Layer messages on
// `esdf_map_in` / `tsdf_map_in` (app_reconstruction/launch/run_experiment.launch:175-181); voxblox_ros'
// EsdfServer deserialises them into its layers (conversions_inl.h deserializeMsgToLayer,
// core/block.cc Block<V>::deserializeFromIntegers) and the planner reads those.  A planner that holds
// only the device mirror can skip the host layers: LayerReceiver decodes the messages straight into
// block deltas.
//
// voxblox_msgs and voxblox are un-vendored dependencies (mav_active_3d_planning_https.rosinstall:5, no
// version pin); this restates the published message definition and voxel packing and is checked here
// against an encoder written from the same description (tests/test_host_cpp.py::test_layer_messages)
// — it is NOT pinned against voxblox itself:
//   voxblox_msgs/Layer: float64 voxel_size, uint32 voxels_per_side, string layer_type ("tsdf" | "esdf" |
//     "occupancy" | ...), uint8 action (ACTION_UPDATE = 0, ACTION_MERGE = 1, ACTION_RESET = 2),
//     voxblox_msgs/Block[] blocks;  Block: int32 x_index, y_index, z_index, uint32[] data.
//   data, per voxel in linear order x + vps*(y + vps*z):
//     EsdfVoxel  2 words: float distance (bit pattern) | word 1 = observed in bits 7:0, the parent
//                direction's three int8 in bits 31:8
//     TsdfVoxel  3 words: float distance | float weight | colour r<<24 | g<<16 | b<<8 | a
// The functions are templates over the message type, so that voxblox_msgs::Layer (and ConstPtr
// dereferenced) works as well as the plain structs below.
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "a3d_host/voxblox_adapter.h"

namespace active_3d_planning {
namespace map {

// Plain mirrors of the two message types (field names as in the .msg files).
struct BlockMsg {
  int32_t x_index = 0, y_index = 0, z_index = 0;
  std::vector<uint32_t> data;
};
struct LayerMsg {
  double voxel_size = 0.0;
  uint32_t voxels_per_side = 0;
  std::string layer_type;
  uint8_t action = 0;
  std::vector<BlockMsg> blocks;
};
enum LayerAction : uint8_t { kLayerUpdate = 0, kLayerMerge = 1, kLayerReset = 2 };

// Feeds a VoxbloxMap (device mirror) from layer messages.  ESDF messages carry distances and observed
// flags, TSDF messages the weights of the same blocks; the two arrive independently, so TSDF weights of
// blocks the mirror does not hold yet are kept until their ESDF block shows up.
class LayerReceiver {
 public:
  explicit LayerReceiver(VoxbloxMap* map) : map_(map) {}

  // false: the message does not fit the map (voxel size / voxels per side / layer type / block size)
  template <class Msg>
  bool onEsdfLayer(const Msg& msg) {
    if (!fits(msg, "esdf", 2)) return false;
    if (msg.action == kLayerReset) reset();
    const size_t nv = voxelsPerBlock();
    BlockDelta with_w, without_w;  // (the upload takes weights for all of its blocks or for none)
    std::vector<int32_t> late_idx;  // TSDF distances that waited for their ESDF block
    std::vector<float> late_dist;
    for (const auto& b : msg.blocks) {
      const Index idx{b.x_index, b.y_index, b.z_index};
      auto pending = pending_weights_.find(idx);
      BlockDelta& d = pending == pending_weights_.end() ? without_w : with_w;
      d.block_idx.insert(d.block_idx.end(), {b.x_index, b.y_index, b.z_index});
      for (size_t i = 0; i < nv; ++i) {
        float dist;
        std::memcpy(&dist, &b.data[2 * i], 4);
        d.esdf_distance.push_back(dist);
        d.esdf_observed.push_back((b.data[2 * i + 1] & 0xFFu) ? 1 : 0);
      }
      if (pending != pending_weights_.end()) {
        d.tsdf_weight.insert(d.tsdf_weight.end(), pending->second.weight.begin(), pending->second.weight.end());
        if (!pending->second.distance.empty()) {
          late_idx.insert(late_idx.end(), {b.x_index, b.y_index, b.z_index});
          late_dist.insert(late_dist.end(), pending->second.distance.begin(), pending->second.distance.end());
        }
        pending_weights_.erase(pending);
      }
      known_.insert({idx, true});
    }
    bool ok = true;
    if (with_w.n())
      ok &= map_->uploadBlocks(with_w.n(), with_w.block_idx.data(), with_w.esdf_distance.data(),
                               with_w.esdf_observed.data(), with_w.tsdf_weight.data());
    if (without_w.n())
      ok &= map_->uploadBlocks(without_w.n(), without_w.block_idx.data(), without_w.esdf_distance.data(),
                               without_w.esdf_observed.data(), nullptr);
    if (!late_idx.empty())
      ok &= map_->uploadTsdfDistances(static_cast<int>(late_idx.size() / 3), late_idx.data(), late_dist.data());
    blocks_received_ += with_w.n() + without_w.n();
    return ok;
  }

  template <class Msg>
  bool onTsdfLayer(const Msg& msg) {
    if (!fits(msg, "tsdf", 3)) return false;
    if (msg.action == kLayerReset) pending_weights_.clear();
    const size_t nv = voxelsPerBlock();
    std::vector<int32_t> idx;
    std::vector<float> weights, distances;
    for (const auto& b : msg.blocks) {
      TsdfPayload t;  // a TSDF voxel is {distance, weight, colour}
      t.weight.resize(nv);
      for (size_t i = 0; i < nv; ++i) std::memcpy(&t.weight[i], &b.data[3 * i + 1], 4);
      if (mirror_tsdf_distances_) {
        t.distance.resize(nv);
        for (size_t i = 0; i < nv; ++i) std::memcpy(&t.distance[i], &b.data[3 * i], 4);
      }
      const Index key{b.x_index, b.y_index, b.z_index};
      if (!known_.count(key)) {
        pending_weights_[key] = std::move(t);  // its ESDF block has not arrived yet
        continue;
      }
      idx.insert(idx.end(), {b.x_index, b.y_index, b.z_index});
      weights.insert(weights.end(), t.weight.begin(), t.weight.end());
      distances.insert(distances.end(), t.distance.begin(), t.distance.end());
    }
    if (idx.empty()) return true;
    bool ok = map_->uploadWeights(static_cast<int>(idx.size() / 3), idx.data(), weights.data());
    if (mirror_tsdf_distances_)
      ok &= map_->uploadTsdfDistances(static_cast<int>(idx.size() / 3), idx.data(), distances.data());
    return ok;
  }

  // also mirror the TSDF voxels' distances (VoxbloxMap::getVoxelDistance, voxblox.cpp:84-98); off by
  // default: nothing in the reference's core reads them
  void mirrorTsdfDistances(bool on) { mirror_tsdf_distances_ = on; }
  size_t blocksReceived() const { return blocks_received_; }
  size_t pendingWeightBlocks() const { return pending_weights_.size(); }

 private:
  struct Index {
    int32_t x, y, z;
    bool operator<(const Index& o) const { return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z); }
  };
  size_t voxelsPerBlock() const {
    const size_t v = static_cast<size_t>(map_->voxelsPerSide());
    return v * v * v;
  }
  template <class Msg>
  bool fits(const Msg& msg, const char* type, size_t words_per_voxel) const {
    if (msg.layer_type != type) return false;
    if (static_cast<int>(msg.voxels_per_side) != map_->voxelsPerSide()) return false;
    // voxblox compares the float voxel sizes it holds
    if (static_cast<float>(msg.voxel_size) != static_cast<float>(map_->getVoxelSize())) return false;
    for (const auto& b : msg.blocks)
      if (b.data.size() != words_per_voxel * voxelsPerBlock()) return false;
    return true;
  }
  void reset() {
    // ACTION_RESET: the sender's layer replaces ours
    std::vector<int32_t> all;
    for (const auto& kv : known_) all.insert(all.end(), {kv.first.x, kv.first.y, kv.first.z});
    if (!all.empty()) map_->removeBlocks(static_cast<int>(all.size() / 3), all.data());
    known_.clear();
  }
  VoxbloxMap* map_;
  std::map<Index, bool> known_;                          // blocks the mirror holds (sent through here)
  struct TsdfPayload {
    std::vector<float> weight, distance;  // (distance only with mirrorTsdfDistances)
  };
  std::map<Index, TsdfPayload> pending_weights_;  // TSDF voxels waiting for their ESDF block
  bool mirror_tsdf_distances_ = false;
  size_t blocks_received_ = 0;
};

}  // namespace map
}  // namespace active_3d_planning
