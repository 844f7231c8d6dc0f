/*
 * voxblox_shim.h — stands in for <voxblox_ros/esdf_server.h> and <voxblox_ros/ros_params.h> so that
 * the reference's map adapter (/root/reference/active_3d_planning_voxblox/src/map/voxblox.cpp)
 * compiles UNMODIFIED into oracle/_ref.  TEST INFRASTRUCTURE ONLY; not used by the product.
 *
 * voxblox (github.com/ethz-asl/voxblox, un-vendored, unpinned in
 * mav_active_3d_planning_https.rosinstall:5) is not in this container.  The classes below carry
 * the upstream names and member functions that voxblox.cpp binds and restate the published
 * arithmetic of voxblox/core/common.h (grid index helpers), core/block.h, core/layer.h,
 * core/esdf_map.cc and interpolator/interpolator_inl.h — SURVEY.md Appendix A.1–A.3.  These
 * three groups of functions (index helpers, nearest-voxel access, 8-corner interpolation) are
 * the only arithmetic on the hot path that is NOT pinned by reference code in oracle/_ref.
 *
 * The summation order of Interpolator::interpMember (`q_vector * (interp_table * data^T)`,
 * Eigen 3.3.7, SSE2: column-major 8x8 GEMV processed four columns at a time, then a
 * vectorised dot) is fixed as SURVEY.md A.3 step 7 states it.
 */
#ifndef A3D_ORACLE_VOXBLOX_SHIM_H_
#define A3D_ORACLE_VOXBLOX_SHIM_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <Eigen/Core>

namespace ros {
// voxblox.cpp:19-20 constructs two handles and hands them to the server; nothing else.
class NodeHandle {
 public:
  explicit NodeHandle(const std::string& ns = std::string()) : ns_(ns) {}
  const std::string& getNamespace() const { return ns_; }

 private:
  std::string ns_;
};
}  // namespace ros

namespace voxblox {

// Call counters of this shim (not part of voxblox): how often the reference reached the three
// kinds of voxel access.  Per thread, monotone; ref_driver.cpp takes differences.
struct ShimCounters {
  uint64_t interp = 0;    // EsdfMap::getDistanceAtPosition (8 voxel records each)
  uint64_t observed = 0;  // EsdfMap::isObserved (1 record)
  uint64_t weight = 0;    // Block::getVoxelPtrByCoordinates (1 record: TSDF weight / distance)
  static ShimCounters& get() {
    static thread_local ShimCounters c;
    return c;
  }
};

// ---- core/common.h --------------------------------------------------------------------------
typedef float FloatingPoint;
typedef int IndexElement;
typedef Eigen::Matrix<FloatingPoint, 3, 1> Point;
typedef Eigen::Matrix<IndexElement, 3, 1> AnyIndex;
typedef AnyIndex VoxelIndex;
typedef AnyIndex BlockIndex;
constexpr FloatingPoint kCoordinateEpsilon = 1e-6;

template <typename IndexType>
inline IndexType getGridIndexFromPoint(const Point& point, const FloatingPoint grid_size_inv) {
  return IndexType(std::floor(point.x() * grid_size_inv + kCoordinateEpsilon),
                   std::floor(point.y() * grid_size_inv + kCoordinateEpsilon),
                   std::floor(point.z() * grid_size_inv + kCoordinateEpsilon));
}
template <typename IndexType>
inline Point getCenterPointFromGridIndex(const IndexType& idx, FloatingPoint grid_size) {
  // the 0.5 literal is a double: the sum and product are evaluated in double, then rounded
  return Point((static_cast<FloatingPoint>(idx.x()) + 0.5) * grid_size,
               (static_cast<FloatingPoint>(idx.y()) + 0.5) * grid_size,
               (static_cast<FloatingPoint>(idx.z()) + 0.5) * grid_size);
}
template <typename IndexType>
inline Point getOriginPointFromGridIndex(const IndexType& idx, FloatingPoint grid_size) {
  return Point(static_cast<FloatingPoint>(idx.x()) * grid_size, static_cast<FloatingPoint>(idx.y()) * grid_size,
               static_cast<FloatingPoint>(idx.z()) * grid_size);
}

// ---- core/voxel.h ---------------------------------------------------------------------------
struct TsdfVoxel {
  float distance = 0.0f;
  float weight = 0.0f;
  uint8_t color[4] = {0, 0, 0, 0};
};
struct EsdfVoxel {
  float distance = 0.0f;
  bool observed = false;
  bool hallucinated = false;
  bool in_queue = false;
  bool fixed = false;
  Eigen::Vector3i parent = Eigen::Vector3i::Zero();
};

// ---- core/block_hash.h ------------------------------------------------------------------------
struct AnyIndexHash {
  static constexpr size_t sl = 17191;
  static constexpr size_t sl2 = sl * sl;
  std::size_t operator()(const AnyIndex& index) const {
    return static_cast<unsigned int>(index.x() + index.y() * sl + index.z() * sl2);
  }
};

// ---- core/block.h -----------------------------------------------------------------------------
template <typename VoxelType>
class Block {
 public:
  typedef std::shared_ptr<Block<VoxelType>> Ptr;
  typedef std::shared_ptr<const Block<VoxelType>> ConstPtr;

  Block(size_t voxels_per_side, FloatingPoint voxel_size, const Point& origin)
      : voxels_per_side_(voxels_per_side), voxel_size_(voxel_size), origin_(origin) {
    num_voxels_ = voxels_per_side_ * voxels_per_side_ * voxels_per_side_;
    voxel_size_inv_ = 1.0 / voxel_size_;
    block_size_ = voxels_per_side_ * voxel_size_;
    block_size_inv_ = 1.0 / block_size_;
    voxels_.reset(new VoxelType[num_voxels_]);
  }

  inline size_t computeLinearIndexFromVoxelIndex(const VoxelIndex& index) const {
    size_t linear_index = static_cast<size_t>(index.x() + voxels_per_side_ * (index.y() + index.z() * voxels_per_side_));
    return linear_index;
  }
  inline VoxelIndex computeTruncatedVoxelIndexFromCoordinates(const Point& coords) const {
    const IndexElement max_value = voxels_per_side_ - 1;
    VoxelIndex voxel_index = getGridIndexFromPoint<VoxelIndex>(coords - origin_, voxel_size_inv_);
    return VoxelIndex(std::max(std::min(voxel_index.x(), max_value), 0), std::max(std::min(voxel_index.y(), max_value), 0),
                      std::max(std::min(voxel_index.z(), max_value), 0));
  }
  inline VoxelIndex computeVoxelIndexFromCoordinates(const Point& coords) const {
    return getGridIndexFromPoint<VoxelIndex>(coords - origin_, voxel_size_inv_);
  }
  inline size_t computeLinearIndexFromCoordinates(const Point& coords) const {
    return computeLinearIndexFromVoxelIndex(computeTruncatedVoxelIndexFromCoordinates(coords));
  }
  inline Point computeCoordinatesFromVoxelIndex(const VoxelIndex& index) const {
    return origin_ + getCenterPointFromGridIndex(index, voxel_size_);
  }
  // Upstream indexes the array unchecked (UB outside [0, num_voxels)); the oracle, the device code
  // and this shim all define an out-of-range linear index as "no voxel".
  inline bool isValidLinearIndex(size_t index) const { return index < num_voxels_; }
  inline const VoxelType& getVoxelByLinearIndex(size_t index) const { return voxels_[index]; }
  inline VoxelType& getVoxelByLinearIndex(size_t index) { return voxels_[index]; }
  inline const VoxelType& getVoxelByCoordinates(const Point& coords) const {
    return voxels_[computeLinearIndexFromCoordinates(coords)];
  }
  inline VoxelType* getVoxelPtrByCoordinates(const Point& coords) {
    ShimCounters::get().weight++;
    return &voxels_[computeLinearIndexFromCoordinates(coords)];
  }

  size_t voxels_per_side() const { return voxels_per_side_; }
  FloatingPoint voxel_size() const { return voxel_size_; }
  FloatingPoint voxel_size_inv() const { return voxel_size_inv_; }
  size_t num_voxels() const { return num_voxels_; }
  const Point& origin() const { return origin_; }
  FloatingPoint block_size() const { return block_size_; }

 private:
  std::unique_ptr<VoxelType[]> voxels_;
  size_t voxels_per_side_;
  FloatingPoint voxel_size_;
  Point origin_;
  size_t num_voxels_;
  FloatingPoint voxel_size_inv_;
  FloatingPoint block_size_;
  FloatingPoint block_size_inv_;
};

// ---- core/layer.h -----------------------------------------------------------------------------
template <typename VoxelType>
class Layer {
 public:
  typedef std::shared_ptr<Layer> Ptr;
  typedef Block<VoxelType> BlockType;
  typedef std::unordered_map<BlockIndex, typename BlockType::Ptr, AnyIndexHash> BlockHashMap;

  Layer(FloatingPoint voxel_size, size_t voxels_per_side) : voxel_size_(voxel_size), voxels_per_side_(voxels_per_side) {
    voxel_size_inv_ = 1.0 / voxel_size_;
    block_size_ = voxel_size_ * voxels_per_side_;
    block_size_inv_ = 1.0 / block_size_;
    voxels_per_side_inv_ = 1.0f / static_cast<FloatingPoint>(voxels_per_side_);
  }

  inline typename BlockType::ConstPtr getBlockPtrByIndex(const BlockIndex& index) const {
    typename BlockHashMap::const_iterator it = block_map_.find(index);
    if (it != block_map_.end()) return it->second;
    return typename BlockType::ConstPtr();
  }
  inline typename BlockType::Ptr getBlockPtrByIndex(const BlockIndex& index) {
    typename BlockHashMap::iterator it = block_map_.find(index);
    if (it != block_map_.end()) return it->second;
    return typename BlockType::Ptr();
  }
  inline BlockIndex computeBlockIndexFromCoordinates(const Point& coords) const {
    return getGridIndexFromPoint<BlockIndex>(coords, block_size_inv_);
  }
  inline typename BlockType::Ptr getBlockPtrByCoordinates(const Point& coords) {
    return getBlockPtrByIndex(computeBlockIndexFromCoordinates(coords));
  }
  inline typename BlockType::ConstPtr getBlockPtrByCoordinates(const Point& coords) const {
    return getBlockPtrByIndex(computeBlockIndexFromCoordinates(coords));
  }
  typename BlockType::Ptr allocateBlockPtrByIndex(const BlockIndex& index) {
    typename BlockHashMap::iterator it = block_map_.find(index);
    if (it != block_map_.end()) return it->second;
    typename BlockType::Ptr b(new BlockType(voxels_per_side_, voxel_size_, getOriginPointFromGridIndex(index, block_size_)));
    block_map_.emplace(index, b);
    return b;
  }
  bool removeBlock(const BlockIndex& index) { return block_map_.erase(index) > 0; }
  void removeAllBlocks() { block_map_.clear(); }
  size_t getNumberOfAllocatedBlocks() const { return block_map_.size(); }

  FloatingPoint block_size() const { return block_size_; }
  FloatingPoint block_size_inv() const { return block_size_inv_; }
  FloatingPoint voxel_size() const { return voxel_size_; }
  FloatingPoint voxel_size_inv() const { return voxel_size_inv_; }
  size_t voxels_per_side() const { return voxels_per_side_; }

 private:
  FloatingPoint voxel_size_;
  size_t voxels_per_side_;
  FloatingPoint voxel_size_inv_;
  FloatingPoint block_size_;
  FloatingPoint block_size_inv_;
  FloatingPoint voxels_per_side_inv_;
  BlockHashMap block_map_;
};

// ---- interpolator/interpolator{.h,_inl.h} (EsdfVoxel instantiation) -------------------------
template <typename VoxelType>
class Interpolator {
 public:
  explicit Interpolator(const Layer<VoxelType>* layer) : layer_(layer) {}

  bool getDistance(const Point& pos, FloatingPoint* distance, bool interpolate = false) const {
    if (interpolate) return getInterpDistance(pos, distance);
    return getNearestDistance(pos, distance);
  }

 private:
  // setIndexes: base voxel of the 2x2x2 cell that has `pos` at or above its centre on every axis
  bool setIndexes(const Point& pos, BlockIndex* block_index, VoxelIndex* voxel_index_out) const {
    *block_index = layer_->computeBlockIndexFromCoordinates(pos);
    typename Layer<VoxelType>::BlockType::ConstPtr block_ptr = layer_->getBlockPtrByIndex(*block_index);
    if (block_ptr == nullptr) return false;
    VoxelIndex voxel_index = block_ptr->computeVoxelIndexFromCoordinates(pos);
    Point center_offset = pos - block_ptr->computeCoordinatesFromVoxelIndex(voxel_index);
    for (int i = 0; i < 3; ++i) {
      if (center_offset(i) < 0) {
        voxel_index(i)--;
        if (voxel_index(i) < 0) {
          (*block_index)(i)--;
          voxel_index(i) += block_ptr->voxels_per_side();
        }
      }
    }
    *voxel_index_out = voxel_index;
    return true;
  }

  void getQVector(const Point& voxel_pos, const Point& pos, const FloatingPoint voxel_size_inv, FloatingPoint* q_vector) const {
    const Point voxel_offset = (pos - voxel_pos) * voxel_size_inv;
    q_vector[0] = 1;
    q_vector[1] = voxel_offset[0];
    q_vector[2] = voxel_offset[1];
    q_vector[3] = voxel_offset[2];
    q_vector[4] = voxel_offset[0] * voxel_offset[1];
    q_vector[5] = voxel_offset[1] * voxel_offset[2];
    q_vector[6] = voxel_offset[2] * voxel_offset[0];
    q_vector[7] = voxel_offset[0] * voxel_offset[1] * voxel_offset[2];
  }

  bool getVoxelsAndQVector(const Point& pos, const VoxelType** voxels, FloatingPoint* q_vector) const {
    BlockIndex block_index;
    VoxelIndex base;
    if (!setIndexes(pos, &block_index, &base)) return false;
    // InterpIndexes columns: (0,0,0) (0,0,1) (0,1,0) (0,1,1) (1,0,0) (1,0,1) (1,1,0) (1,1,1)
    for (int i = 0; i < 8; ++i) {
      typename Layer<VoxelType>::BlockType::ConstPtr block_ptr = layer_->getBlockPtrByIndex(block_index);
      if (block_ptr == nullptr) return false;
      VoxelIndex voxel_index(base.x() + ((i >> 2) & 1), base.y() + ((i >> 1) & 1), base.z() + (i & 1));
      const IndexElement vps = static_cast<IndexElement>(block_ptr->voxels_per_side());
      if (voxel_index.x() >= vps || voxel_index.y() >= vps || voxel_index.z() >= vps) {
        BlockIndex new_block_index = block_index;
        for (int j = 0; j < 3; ++j) {
          if (voxel_index(j) >= vps) {
            new_block_index(j)++;
            voxel_index(j) -= vps;
          }
        }
        block_ptr = layer_->getBlockPtrByIndex(new_block_index);
        if (block_ptr == nullptr) return false;
      }
      if (i == 0) {
        getQVector(block_ptr->computeCoordinatesFromVoxelIndex(voxel_index), pos, block_ptr->voxel_size_inv(), q_vector);
      }
      const size_t lin = block_ptr->computeLinearIndexFromVoxelIndex(voxel_index);
      if (!block_ptr->isValidLinearIndex(lin)) return false;  // UB upstream; defined as "no voxel"
      const VoxelType& voxel = block_ptr->getVoxelByLinearIndex(lin);
      voxels[i] = &voxel;
      if (!voxel.observed) return false;  // utils::isObservedVoxel<EsdfVoxel>
    }
    return true;
  }

  bool getInterpDistance(const Point& pos, FloatingPoint* distance) const {
    const VoxelType* voxels[8];
    FloatingPoint q[8];
    if (!getVoxelsAndQVector(pos, voxels, q)) return false;
    // interpMember: q_vector * (interp_table * data^T), order per SURVEY.md A.3 step 7
    FloatingPoint d[8];
    for (int i = 0; i < 8; ++i) d[i] = voxels[i]->distance;
    static const int T[8][8] = {{1, 0, 0, 0, 0, 0, 0, 0},    {-1, 0, 0, 0, 1, 0, 0, 0},  {-1, 0, 1, 0, 0, 0, 0, 0},
                                {-1, 1, 0, 0, 0, 0, 0, 0},   {1, 0, -1, 0, -1, 0, 1, 0}, {1, -1, -1, 1, 0, 0, 0, 0},
                                {1, -1, 0, 0, -1, 1, 0, 0},  {-1, 1, 1, -1, 1, -1, -1, 1}};
    FloatingPoint u[8];
    for (int r = 0; r < 8; ++r) {
      FloatingPoint t[8];
      for (int c = 0; c < 8; ++c) t[c] = static_cast<FloatingPoint>(T[r][c]) * d[c];
      const FloatingPoint w = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
      u[r] = q[r] * w;
    }
    *distance = ((u[0] + u[4]) + (u[2] + u[6])) + ((u[1] + u[5]) + (u[3] + u[7]));
    return true;
  }

  bool getNearestDistance(const Point& pos, FloatingPoint* distance) const {
    typename Layer<VoxelType>::BlockType::ConstPtr block_ptr = layer_->getBlockPtrByCoordinates(pos);
    if (block_ptr == nullptr) return false;
    const VoxelType& voxel = block_ptr->getVoxelByCoordinates(pos);
    *distance = voxel.distance;
    return voxel.observed;
  }

  const Layer<VoxelType>* layer_;
};

// ---- core/esdf_map.{h,cc}, core/tsdf_map.h ----------------------------------------------------
class EsdfMap {
 public:
  struct Config {
    FloatingPoint esdf_voxel_size = 0.2;
    size_t esdf_voxels_per_side = 16u;
  };
  explicit EsdfMap(const Config& config)
      : esdf_layer_(new Layer<EsdfVoxel>(config.esdf_voxel_size, config.esdf_voxels_per_side)), interpolator_(esdf_layer_.get()) {
    block_size_ = config.esdf_voxel_size * config.esdf_voxels_per_side;
  }
  Layer<EsdfVoxel>* getEsdfLayerPtr() { return esdf_layer_.get(); }
  const Layer<EsdfVoxel>& getEsdfLayer() const { return *esdf_layer_; }
  FloatingPoint block_size() const { return block_size_; }
  FloatingPoint voxel_size() const { return esdf_layer_->voxel_size(); }

  bool getDistanceAtPosition(const Eigen::Vector3d& position, double* distance) const {
    constexpr bool interpolate = true;
    return getDistanceAtPosition(position, interpolate, distance);
  }
  bool getDistanceAtPosition(const Eigen::Vector3d& position, bool interpolate, double* distance) const {
    FloatingPoint distance_fp;
    ShimCounters::get().interp++;
    bool success = interpolator_.getDistance(position.cast<FloatingPoint>(), &distance_fp, interpolate);
    if (success) *distance = static_cast<double>(distance_fp);
    return success;
  }
  bool isObserved(const Eigen::Vector3d& position) const {
    ShimCounters::get().observed++;
    Block<EsdfVoxel>::ConstPtr block_ptr = esdf_layer_->getBlockPtrByCoordinates(position.cast<FloatingPoint>());
    if (block_ptr) {
      const EsdfVoxel& voxel = block_ptr->getVoxelByCoordinates(position.cast<FloatingPoint>());
      return voxel.observed;
    }
    return false;
  }

 private:
  FloatingPoint block_size_;
  Layer<EsdfVoxel>::Ptr esdf_layer_;
  Interpolator<EsdfVoxel> interpolator_;
};

class TsdfMap {
 public:
  struct Config {
    FloatingPoint tsdf_voxel_size = 0.2;
    size_t tsdf_voxels_per_side = 16u;
  };
  explicit TsdfMap(const Config& config) : tsdf_layer_(new Layer<TsdfVoxel>(config.tsdf_voxel_size, config.tsdf_voxels_per_side)) {}
  Layer<TsdfVoxel>* getTsdfLayerPtr() { return tsdf_layer_.get(); }

 private:
  Layer<TsdfVoxel>::Ptr tsdf_layer_;
};

// ---- voxblox_ros/{esdf_server,tsdf_server,ros_params}.h -------------------------------------
// The real server reads tsdf_voxel_size / tsdf_voxels_per_side / max_weight from the private ROS
// namespace; the test driver sets them here before the reference's VoxbloxMap is created.
struct ShimRosParams {
  FloatingPoint tsdf_voxel_size = 0.2;
  int tsdf_voxels_per_side = 16;
  float max_weight = 10000.0f;
  static ShimRosParams& get() {
    static ShimRosParams p;
    return p;
  }
};

struct TsdfIntegratorConfig {
  float max_weight = 10000.0f;
};
inline TsdfIntegratorConfig getTsdfIntegratorConfigFromRosParam(const ros::NodeHandle& /*nh_private*/) {
  TsdfIntegratorConfig c;
  c.max_weight = ShimRosParams::get().max_weight;
  return c;
}

class EsdfServer {
 public:
  EsdfServer(const ros::NodeHandle& /*nh*/, const ros::NodeHandle& /*nh_private*/) {
    const ShimRosParams& p = ShimRosParams::get();
    TsdfMap::Config tc;
    tc.tsdf_voxel_size = p.tsdf_voxel_size;
    tc.tsdf_voxels_per_side = static_cast<size_t>(p.tsdf_voxels_per_side);
    tsdf_map_.reset(new TsdfMap(tc));
    EsdfMap::Config ec;  // esdf_server.cc: the ESDF layer takes the TSDF layer's geometry
    ec.esdf_voxel_size = p.tsdf_voxel_size;
    ec.esdf_voxels_per_side = static_cast<size_t>(p.tsdf_voxels_per_side);
    esdf_map_.reset(new EsdfMap(ec));
  }
  void setTraversabilityRadius(float r) { traversability_radius_ = r; }
  float getTraversabilityRadius() const { return traversability_radius_; }
  std::shared_ptr<EsdfMap> getEsdfMapPtr() { return esdf_map_; }
  std::shared_ptr<TsdfMap> getTsdfMapPtr() { return tsdf_map_; }

 private:
  std::shared_ptr<TsdfMap> tsdf_map_;
  std::shared_ptr<EsdfMap> esdf_map_;
  float traversability_radius_ = 0.0f;
};

}  // namespace voxblox
#endif  // A3D_ORACLE_VOXBLOX_SHIM_H_
