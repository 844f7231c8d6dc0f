// voxblox_adapter.h — voxblox layers → a3d block deltas.
//
// In the reference the map arrives through voxblox_ros: the planner's EsdfServer merges
// `esdf_map_in` / `tsdf_map_in` layer messages (app_reconstruction/launch/run_experiment.launch:175-181)
// and the VoxbloxMap adapter holds the server (vbx/src/map/voxblox.cpp:17-30).  This header is what a
// maintainer drops next to that server: once per planner iteration it walks the blocks voxblox
// flagged as updated and hands them to the device mirror as one delta.
//
// voxblox is not installed where this repository is built, so the functions are templates over the
// layer types and use only this part of voxblox's public API (names as in voxblox's
// core/layer.h, core/block.h, core/voxel.h; they are exercised here against a mock with the same
// signatures, tests/test_host_cpp.py::test_voxblox_layer_adapter):
//   Layer<V>::voxels_per_side(), ::voxel_size(), ::getAllUpdatedBlocks(Update::Status, BlockIndexList*),
//   ::getBlockByIndex(const BlockIndex&) → Block<V>&, ::getBlockPtrByIndex(idx) → shared_ptr (may be null)
//   Block<V>::num_voxels(), ::getVoxelByLinearIndex(size_t), ::updated() → std::bitset<Update::kCount>&
//   EsdfVoxel{float distance; bool observed; …}, TsdfVoxel{float distance; float weight; …}
// voxblox's linear voxel order, x + vps*(y + vps*z), is the order a3d_map_upload_blocks expects.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "a3d_host/map.h"

namespace active_3d_planning {
namespace map {

// One delta: n blocks, their indices and per-voxel payload in upload order.
struct BlockDelta {
  std::vector<int32_t> block_idx;     // n × 3
  std::vector<float> esdf_distance;   // n × vps³
  std::vector<uint8_t> esdf_observed; // n × vps³
  std::vector<float> tsdf_weight;     // n × vps³ (empty when no TSDF layer is given)
  int n() const { return static_cast<int>(block_idx.size() / 3); }
};

// Collects every ESDF block whose `status` bit is set (voxblox::Update::kMap for a planner that
// consumes the map), optionally with the TSDF weights of the same blocks (0 where the TSDF layer has
// no such block), and clears the bit.  `status` is passed in so that this header does not need
// voxblox's headers itself.
// Call as collectUpdatedBlocks<voxblox::BlockIndexList>(esdf_layer, tsdf_layer_or_null, voxblox::Update::kMap).
template <class BlockIndexList, class EsdfLayer, class TsdfLayer, class UpdateStatus>
BlockDelta collectUpdatedBlocks(EsdfLayer* esdf, TsdfLayer* tsdf, UpdateStatus status) {
  BlockDelta d;
  BlockIndexList updated;
  esdf->getAllUpdatedBlocks(status, &updated);
  const size_t nv = static_cast<size_t>(esdf->voxels_per_side()) * esdf->voxels_per_side() *
                    esdf->voxels_per_side();
  d.block_idx.reserve(updated.size() * 3);
  d.esdf_distance.reserve(updated.size() * nv);
  d.esdf_observed.reserve(updated.size() * nv);
  if (tsdf) d.tsdf_weight.reserve(updated.size() * nv);
  for (const auto& index : updated) {
    auto& block = esdf->getBlockByIndex(index);
    d.block_idx.push_back(static_cast<int32_t>(index.x()));
    d.block_idx.push_back(static_cast<int32_t>(index.y()));
    d.block_idx.push_back(static_cast<int32_t>(index.z()));
    for (size_t i = 0; i < nv; ++i) {
      const auto& v = block.getVoxelByLinearIndex(i);
      d.esdf_distance.push_back(v.distance);
      d.esdf_observed.push_back(v.observed ? 1 : 0);
    }
    if (tsdf) {
      auto tblock = tsdf->getBlockPtrByIndex(index);
      for (size_t i = 0; i < nv; ++i)
        d.tsdf_weight.push_back(tblock ? tblock->getVoxelByLinearIndex(i).weight : 0.0f);
    }
    block.updated().reset(static_cast<size_t>(status));
  }
  return d;
}

// collectUpdatedBlocks + VoxbloxMap::uploadBlocks.  Returns the number of blocks sent.
template <class BlockIndexList, class EsdfLayer, class TsdfLayer, class UpdateStatus>
int syncUpdatedBlocks(VoxbloxMap* map, EsdfLayer* esdf, TsdfLayer* tsdf, UpdateStatus status) {
  const BlockDelta d = collectUpdatedBlocks<BlockIndexList>(esdf, tsdf, status);
  if (d.n() == 0) return 0;
  const bool ok = map->uploadBlocks(d.n(), d.block_idx.data(), d.esdf_distance.data(), d.esdf_observed.data(),
                                    d.tsdf_weight.empty() ? nullptr : d.tsdf_weight.data());
  return ok ? d.n() : -1;
}

}  // namespace map
}  // namespace active_3d_planning
