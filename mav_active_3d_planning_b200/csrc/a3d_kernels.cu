// a3d_kernels.cu — sm_100a kernels of the view-gain path.
//
//   k_apply_blocks       K1a: merge an uploaded block delta (distance + observed flag) into the
//                        padded slab interiors
//   k_refresh_aprons     K1b: re-mirror the one-voxel aprons of every block touched by the delta
//   k_build_meta         K1c: per-voxel meta byte (cell class, exact centre state, observed)
//   k_point_query        batched OccupancyMap/TSDFMap virtuals (voxblox.cpp:43-115)
//   k_cast_gain          K2/K3 + fold: SimpleRayCaster / IterativeRayCaster march
//                        (simple_ray_caster.cpp:32-66, iterative_ray_caster.cpp:48-119), adjacent
//                        unique (camera_model.cpp:58), bounding-volume filter
//                        (simulated_sensor_evaluator.cpp:50-57) and the evaluator fold
//                        (voxel_type_evaluator.cpp:57-89, naive_evaluator.cpp:20-38,
//                        frontier_evaluator.cpp:100-124) fused; one warp per trajectory segment
//   k_gain_from_voxels   K4: the fold alone on a stored list (SimulatedSensorUpdater path)
//
// Mapping of k_cast_gain (DESIGN.md "kernel K3"): the IterativeRayCaster's ray table makes rays of
// one view order-dependent, and std::unique + the gain sum make views of one segment
// order-dependent, so a warp owns a whole segment and walks views → rays in the reference's scan
// order, while its 32 lanes take 32 consecutive ray samples.  The first OCCUPIED sample is found
// with a ballot; the previous raw list entry for the adjacent-unique test comes from the
// neighbouring lane (shuffle) or from a warp-uniform carry across chunks, rays and views.
#include "a3d_kernels.cuh"
#include "a3d_keys.cuh"

#include <algorithm>

#include "../../include/a3d.h"

namespace a3d {

constexpr unsigned kFull = 0xFFFFFFFFu;

static inline int grid_for(long long total, int block, int per_sm = 16) {
  return static_cast<int>(std::max<long long>(
      1, std::min<long long>((total + block - 1) / block, 148LL * per_sm)));
}

// ---------------------------------------------------------------------------------------------
__global__ void k_apply_blocks(MapView m, int n, const int32_t* __restrict__ slots,
                               const float* __restrict__ dist_src, const uint8_t* __restrict__ obs_src,
                               const float* __restrict__ w_src, float* __restrict__ dist_slab,
                               float* __restrict__ weight_slab, float* __restrict__ tsdf_slab) {
  const long long total = static_cast<long long>(n) * m.nv;
  const int vv = m.vps * m.vps;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int blk = static_cast<int>(i / m.nv);
    const int lin = static_cast<int>(i - static_cast<long long>(blk) * m.nv);
    const int z = lin / vv, rem = lin - z * vv, y = rem / m.vps, x = rem - y * m.vps;
    // bit 30: the slot was allocated by this upload.  Without a weight payload (an ESDF-only delta) the
    // TSDF weights of blocks that already exist stay as they are; only fresh slots start at 0.
    const bool fresh = (slots[blk] >> 30) & 1;
    const int slot = slots[blk] & ~(1 << 30);
    const float d = dist_src[i];
    uint32_t bits = __float_as_uint(d);
    if (!obs_src[i]) {
      bits = kUnobservedBits;
    } else if (d != d) {
      bits = 0x7FC00000u;  // canonical NaN for observed voxels
    }
    dist_slab[padded_index(m, slot, x, y, z)] = __uint_as_float(bits);
    if (weight_slab && (w_src || fresh)) weight_slab[static_cast<size_t>(slot) * m.nv + lin] = w_src ? w_src[i] : 0.0f;
    // (TSDF distances arrive on their own, a3d_map_upload_tsdf_distances: a fresh slot has none yet)
    if (tsdf_slab && fresh) tsdf_slab[static_cast<size_t>(slot) * m.nv + lin] = 0.0f;
  }
}

void launch_apply_blocks(cudaStream_t st, const MapView& m, int n, const int32_t* slots_dev,
                         const float* dist_src, const uint8_t* obs_src, const float* w_src,
                         float* dist_slab, float* weight_slab, float* tsdf_slab) {
  if (n <= 0) return;
  const long long total = static_cast<long long>(n) * m.nv;
  k_apply_blocks<<<grid_for(total, 256), 256, 0, st>>>(m, n, slots_dev, dist_src, obs_src, w_src,
                                                       dist_slab, weight_slab, tsdf_slab);
}

// For every affected block (slot, 27 neighbour slots in (dz,dy,dx) order, -1 = not allocated) copy
// the facing layer of each neighbour's interior into the apron; kUnobservedBits if it is missing.
__global__ void k_refresh_aprons(MapView m, int n, const int32_t* __restrict__ slots,
                                 const int32_t* __restrict__ nbr, float* __restrict__ dist_slab) {
  const int P = m.P;
  for (int a = blockIdx.x; a < n; a += gridDim.x) {
    const int slot = slots[a];
    const int32_t* nb = nbr + static_cast<size_t>(a) * 27;
    for (int i = threadIdx.x; i < m.P3; i += blockDim.x) {
      const int z = i / (P * P) - 1, rem = i % (P * P), y = rem / P - 1, x = rem % P - 1;
      const int dx = x < 0 ? -1 : (x >= m.vps ? 1 : 0);
      const int dy = y < 0 ? -1 : (y >= m.vps ? 1 : 0);
      const int dz = z < 0 ? -1 : (z >= m.vps ? 1 : 0);
      if ((dx | dy | dz) == 0) continue;  // interior
      const int src = nb[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)];
      float val = __uint_as_float(kUnobservedBits);
      if (src >= 0)
        val = dist_slab[padded_index(m, src, x - dx * m.vps, y - dy * m.vps, z - dz * m.vps)];
      dist_slab[static_cast<size_t>(slot) * m.P3 + i] = val;
    }
  }
}

void launch_refresh_aprons(cudaStream_t st, const MapView& m, int n, const int32_t* slots_dev,
                           const int32_t* nbr_dev, float* dist_slab) {
  if (n <= 0) return;
  k_refresh_aprons<<<std::min(n, 148 * 8), 256, 0, st>>>(m, n, slots_dev, nbr_dev, dist_slab);
}

// Class of the interpolation cell whose base corner is padded voxel i of the block at dp:
// UNKNOWN if any of the 8 corners is unobserved/missing; FREE / OCCUPIED only when every corner is
// on the same side of the voxel size by more than the worst-case rounding error of interp8
// (DESIGN.md "cell class margin"), so the class equals the interpolated test for every point of
// the cell; MIXED otherwise (the kernels interpolate).
__device__ __forceinline__ int cell_class(const MapView& m, const float* __restrict__ dp, int i) {
  const int P = m.P, PP = m.P * m.P;
  bool all_obs = true, any_nan = false;
  float mn = 3.0e38f, mx = -3.0e38f, amax = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float d = dp[i + ((c >> 2) & 1) + ((c >> 1) & 1) * P + (c & 1) * PP];
    all_obs &= (__float_as_uint(d) != kUnobservedBits);
    any_nan |= (d != d);
    mn = fminf(mn, d);
    mx = fmaxf(mx, d);
    amax = fmaxf(amax, fabsf(d));
  }
  if (!all_obs) return kCellUnknown;
  const double margin = 1.0e-4 * fmax(static_cast<double>(amax), m.c_vs);
  if (any_nan || !(amax < 1.0e30f)) return kCellMixed;
  if (static_cast<double>(mn) - margin >= m.c_vs) return kCellFree;
  if (static_cast<double>(mx) + margin < m.c_vs) return kCellOccupied;
  return kCellMixed;
}

// Meta word per voxel of every affected block (layout: a3d_device.cuh header).  One CTA per block;
// the (vps+1)^3 cell classes are staged in shared memory when they fit.
// Index of voxel (x,y,z) of block bi in the dense meta volume (the block lies inside the box).
__device__ __forceinline__ size_t vol_index(const MapView& m, const int32_t* bi, int x, int y, int z) {
  return vol_offset(m, static_cast<uint32_t>(bi[0] * m.vps + x - m.vol_lo[0]),
                    static_cast<uint32_t>(bi[1] * m.vps + y - m.vol_lo[1]),
                    static_cast<uint32_t>(bi[2] * m.vps + z - m.vol_lo[2]));
}

// `split` CTAs share a block, each taking a slab of z layers (a delta of a few blocks is a latency
// problem: one CTA per block leaves most of the machine idle).
__global__ void k_build_meta(MapView m, int n, const int32_t* __restrict__ slots,
                             const int32_t* __restrict__ block_idx, uint32_t* __restrict__ mword_slab,
                             uint32_t* __restrict__ vol, int use_smem, int split) {
  extern __shared__ uint8_t s_cls[];
  const int P = m.P, PP = m.P * m.P, vps = m.vps;
  const int layers = (vps + split - 1) / split;  // z layers per CTA
  for (int unit = blockIdx.x; unit < n * split; unit += gridDim.x) {
    const int a = unit / split, z_lo = (unit - a * split) * layers, z_hi = min(vps, z_lo + layers);
    if (z_lo >= z_hi) continue;
    const int slot = slots[a];
    const float* __restrict__ dp = m.dist + static_cast<size_t>(slot) * m.P3;
    if (use_smem) {
      // the cells the slab's voxels look at: padded z in [z_lo, z_hi]
      __syncthreads();
      for (int i = z_lo * PP + threadIdx.x; i < (z_hi + 1) * PP; i += blockDim.x) {
        const int z = i / PP, rem = i % PP, y = rem / P, x = rem % P;  // padded coords (base + 1)
        s_cls[i] = (x <= vps && y <= vps && z <= vps) ? cell_class(m, dp, i) : kCellUnknown;
      }
      __syncthreads();
    }
    const int32_t* bi = block_idx + 3 * a;
    for (int lin = z_lo * vps * vps + threadIdx.x; lin < z_hi * vps * vps; lin += blockDim.x) {
      const int z = lin / (vps * vps), rem = lin % (vps * vps), y = rem / vps, x = rem % vps;
      uint32_t word = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        // cell with base = voxel - (c&1, c>>1&1, c>>2&1); padded index of the base = base + 1
        const int i = ((z + 1 - (c >> 2)) * P + (y + 1 - ((c >> 1) & 1))) * P + (x + 1 - (c & 1));
        const int cls = use_smem ? s_cls[i] : cell_class(m, dp, i);
        word |= static_cast<uint32_t>(cls) << (2 * c);
      }
      const int obs = __float_as_uint(dp[((z + 1) * P + (y + 1)) * P + (x + 1)]) != kUnobservedBits;
      // centre exactly as getVoxelCenter forms it (voxblox.cpp:72-79), cast like getVoxelState
      const double cx = __dadd_rn(static_cast<double>(origin_point(bi[0], m.bs_f)),
                                  static_cast<double>(center_point(x, m.vs_f)));
      const double cy = __dadd_rn(static_cast<double>(origin_point(bi[1], m.bs_f)),
                                  static_cast<double>(center_point(y, m.vs_f)));
      const double cz = __dadd_rn(static_cast<double>(origin_point(bi[2], m.bs_f)),
                                  static_cast<double>(center_point(z, m.vs_f)));
      // The centre lies within a few float ulps of where the voxel's 8 cells meet: when their classes
      // agree, the state there is that class whichever cell the rounding picks; else literally.
      const uint32_t cells = word & 0xFFFFu;
      const int centre = cells == 0x0000u ? 2
                         : cells == 0x5555u ? 1
                         : cells == 0xAAAAu ? 0
                                            : voxel_state_generic(m, __double2float_rn(cx), __double2float_rn(cy),
                                                                  __double2float_rn(cz));
      word |= static_cast<uint32_t>(centre) << 16;
      word |= static_cast<uint32_t>(obs) << 18;
      mword_slab[static_cast<size_t>(slot) * m.nv + mword_offset(m, x, y, z)] = word;
      if (vol) vol[vol_index(m, bi, x, y, z)] = word;
    }
  }
}

void launch_build_meta(cudaStream_t st, const MapView& m, int n, const int32_t* slots_dev,
                       const int32_t* block_idx_dev, uint32_t* mword_slab, uint32_t* vol) {
  if (n <= 0) return;
  const int use_smem = m.P3 <= 40 * 1024;
  int split = 1;
  while (split < 8 && n * split * 2 <= 148 * 8 && split * 2 <= m.vps) split *= 2;
  k_build_meta<<<std::min(n * split, 148 * 8), 256, use_smem ? m.P3 : 0, st>>>(m, n, slots_dev, block_idx_dev,
                                                                               mword_slab, vol, use_smem, split);
}

// Bits 23:19 of a meta word for the voxel with global coordinate g and centre c: first neighbour that
// is not UNKNOWN in the 6- and in the 26-neighbourhood (frontier_evaluator.cpp:69-98 with
// checking_distance 1).  A neighbour whose 8 cell classes agree (read from the dense volume, whose
// class bits the preceding k_build_meta pass wrote) has that state; the others are evaluated
// literally at the point the reference would form (neighbour_cells / is_frontier_by_meta explain why).
__device__ __forceinline__ uint32_t frontier_bits(const MapView& m, const CenterTables& ct, const D3& c,
                                                 int g0, int g1, int g2) {
  uint32_t bits = kWordFrontierValid;
#pragma unroll 1
  for (int nb = 0; nb < 2; ++nb) {
    const int count = nb ? 26 : 6;
#pragma unroll 1
    for (int i = 0; i < count; ++i) {
      int sx, sy, sz;
      frontier_offset(nb != 0, i, &sx, &sy, &sz);
      int st = -1;
      if (m.vol) {
        const uint32_t cells = neighbour_cells(m, g0 + sx, g1 + sy, g2 + sz);
        st = cells == 0x0000u ? 2 : (cells == 0x5555u ? 1 : (cells == 0xAAAAu ? 0 : -1));
      }
      if (st < 0) {
        D3 q;
        q.x = __dadd_rn(c.x, sx > 0 ? m.c_vs : (sx < 0 ? -m.c_vs : 0.0));
        q.y = __dadd_rn(c.y, sy > 0 ? m.c_vs : (sy < 0 ? -m.c_vs : 0.0));
        q.z = __dadd_rn(c.z, sz > 0 ? m.c_vs : (sz < 0 ? -m.c_vs : 0.0));
        st = voxel_state(m, ct, q);
      }
      if (st == 2) continue;
      bits |= (st == 0 ? 1u : 0u) << (19 + 2 * nb);
      bits |= 1u << (20 + 2 * nb);
      break;
    }
  }
  return bits;
}

// K1d: frontier bits (meta-word bits 23:19) of every voxel of the affected blocks whose centre is
// UNKNOWN or unobserved.  Runs after k_build_meta of ALL affected blocks: the neighbour states come
// from voxel_state(), i.e. from the cell classes / padded cubes that pass just wrote.  The neighbour
// points are formed exactly like frontier_evaluator.cpp:69-98 does for checking_distance 1.
__global__ void k_build_frontier(MapView m, int n, const int32_t* __restrict__ slots,
                                 const int32_t* __restrict__ block_idx, uint32_t* __restrict__ mword_slab,
                                 uint32_t* __restrict__ vol) {
  __shared__ CenterTables ct;
  init_center_tables(m, &ct);
  __syncthreads();
  const int vps = m.vps;
  // flat over (block, voxel): a delta of a few blocks still fills the machine
  const long long total = static_cast<long long>(n) * m.nv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int a = static_cast<int>(i / m.nv), lin = static_cast<int>(i - static_cast<long long>(a) * m.nv);
    const int slot = slots[a];
    const int32_t* bi = block_idx + 3 * a;
    const int z = lin / (vps * vps), rem = lin % (vps * vps), y = rem / vps, x = rem % vps;
    uint32_t* wp = mword_slab + static_cast<size_t>(slot) * m.nv + mword_offset(m, x, y, z);
    uint32_t word = *wp & ~(0x1Fu << 19);
    const bool candidate = ((word >> 16) & 3u) == 2u || !((word >> 18) & 1u);
    if (candidate) {
      const D3 c{__dadd_rn(static_cast<double>(origin_point(bi[0], m.bs_f)), ct.d[x + 1]),
                 __dadd_rn(static_cast<double>(origin_point(bi[1], m.bs_f)), ct.d[y + 1]),
                 __dadd_rn(static_cast<double>(origin_point(bi[2], m.bs_f)), ct.d[z + 1])};
      word |= frontier_bits(m, ct, c, bi[0] * vps + x, bi[1] * vps + y, bi[2] * vps + z);
    }
    *wp = word;
    if (vol) vol[vol_index(m, bi, x, y, z)] = word;
  }
}

void launch_build_frontier(cudaStream_t st, const MapView& m, int n, const int32_t* slots_dev,
                           const int32_t* block_idx_dev, uint32_t* mword_slab, uint32_t* vol) {
  if (n > 0)
    k_build_frontier<<<grid_for(static_cast<long long>(n) * m.nv, 256), 256, 0, st>>>(m, n, slots_dev, block_idx_dev,
                                                                                       mword_slab, vol);
}

// Dense-volume words of UNALLOCATED blocks next to allocated ones: "no block" plus the frontier bits
// of every voxel (a voxel without a block can still be a frontier voxel when its neighbour lies in an
// allocated block).  Unallocated voxels elsewhere hold kWordNoBlock | kWordFrontierValid: all their
// neighbours are UNKNOWN.
__global__ void k_vol_frontier_unallocated(MapView m, uint32_t* __restrict__ vol, int n,
                                           const int32_t* __restrict__ block_idx) {
  __shared__ CenterTables ct;
  init_center_tables(m, &ct);
  __syncthreads();
  const int vps = m.vps;
  const long long total = static_cast<long long>(n) * m.nv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int a = static_cast<int>(i / m.nv), lin = static_cast<int>(i - static_cast<long long>(a) * m.nv);
    const int32_t* bi = block_idx + 3 * a;
    const int z = lin / (vps * vps), rem = lin % (vps * vps), y = rem / vps, x = rem % vps;
    uint32_t bits = kWordFrontierValid;
    // only the outer voxel layer of the block has a neighbour in another block
    if (x == 0 || y == 0 || z == 0 || x == vps - 1 || y == vps - 1 || z == vps - 1) {
      const D3 c{__dadd_rn(static_cast<double>(origin_point(bi[0], m.bs_f)), ct.d[x + 1]),
                 __dadd_rn(static_cast<double>(origin_point(bi[1], m.bs_f)), ct.d[y + 1]),
                 __dadd_rn(static_cast<double>(origin_point(bi[2], m.bs_f)), ct.d[z + 1])};
      bits = frontier_bits(m, ct, c, bi[0] * vps + x, bi[1] * vps + y, bi[2] * vps + z);
    }
    vol[vol_index(m, bi, x, y, z)] = kWordNoBlock | bits;
  }
}

void launch_vol_frontier_unallocated(cudaStream_t st, const MapView& m, uint32_t* vol, int n,
                                     const int32_t* block_idx_dev) {
  if (n > 0)
    k_vol_frontier_unallocated<<<grid_for(static_cast<long long>(n) * m.nv, 256), 256, 0, st>>>(m, vol, n,
                                                                                                block_idx_dev);
}

__global__ void k_vol_fill_all(uint32_t* __restrict__ vol, size_t n_words) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_words;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    vol[i] = kWordNoBlock | kWordFrontierValid;  // no block around: not a frontier voxel
}

// one CTA per listed block: slots == nullptr writes "no block", else copies the block's meta words
__global__ void k_vol_blocks(MapView m, uint32_t* __restrict__ vol, int n,
                             const int32_t* __restrict__ slots, const int32_t* __restrict__ block_idx) {
  const int vps = m.vps;
  for (int a = blockIdx.x; a < n; a += gridDim.x) {
    const int32_t* bi = block_idx + 3 * a;
    const uint32_t* src = slots ? m.mword + static_cast<size_t>(slots[a]) * m.nv : nullptr;
    for (int lin = threadIdx.x; lin < m.nv; lin += blockDim.x) {
      const int z = lin / (vps * vps), rem = lin % (vps * vps), y = rem / vps, x = rem % vps;
      vol[vol_index(m, bi, x, y, z)] = src ? src[mword_offset(m, x, y, z)] : (kWordNoBlock | kWordFrontierValid);
    }
  }
}

void launch_vol_fill(cudaStream_t st, const MapView& m, uint32_t* vol, int n_blocks,
                     const int32_t* block_idx_dev) {
  if (!block_idx_dev) {
    const size_t n_words = static_cast<size_t>(m.vol_n[0]) * m.vol_n[1] * m.vol_n[2];
    if (n_words) k_vol_fill_all<<<148 * 8, 256, 0, st>>>(vol, n_words);
  } else if (n_blocks > 0) {
    k_vol_blocks<<<std::min(n_blocks, 148 * 8), 256, 0, st>>>(m, vol, n_blocks, nullptr, block_idx_dev);
  }
}

void launch_vol_scatter(cudaStream_t st, const MapView& m, uint32_t* vol, int n, const int32_t* slots_dev,
                        const int32_t* block_idx_dev) {
  if (n > 0) k_vol_blocks<<<std::min(n, 148 * 8), 256, 0, st>>>(m, vol, n, slots_dev, block_idx_dev);
}

// ---------------------------------------------------------------------------------------------
__global__ void k_point_query(MapView m, int what, long long n, const double* __restrict__ pts,
                              void* out0, void* out1, double arg) {
  __shared__ CenterTables ct;
  init_center_tables(m, &ct);
  __syncthreads();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const D3 p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    switch (what) {
      case kQueryState:
        static_cast<uint8_t*>(out0)[i] = static_cast<uint8_t>(voxel_state(m, ct, p));
        break;
      case kQueryStateGeneric:
        static_cast<uint8_t*>(out0)[i] = static_cast<uint8_t>(voxel_state_generic(
            m, __double2float_rn(p.x), __double2float_rn(p.y), __double2float_rn(p.z)));
        break;
      case kQueryCenter: {
        const D3 c = voxel_center(m, ct, p);
        double* o = static_cast<double*>(out0);
        o[3 * i] = c.x;
        o[3 * i + 1] = c.y;
        o[3 * i + 2] = c.z;
        break;
      }
      case kQueryObserved:
        static_cast<uint8_t*>(out0)[i] = is_observed(m, p) ? 1 : 0;
        break;
      case kQueryWeight:
        static_cast<double*>(out0)[i] = voxel_weight(m, p);
        break;
      case kQueryTsdfDistance: {  // VoxbloxMap::getVoxelDistance (voxblox.cpp:84-98): same voxel as the weight
        const float* tsdf = static_cast<const float*>(out1);
        int v[3];
        const int slot = tsdf ? nearest_voxel(m, p, v) : -1;
        static_cast<double*>(out0)[i] =
            slot < 0 ? 0.0
                     : static_cast<double>(__ldg(tsdf + static_cast<size_t>(slot) * m.nv + v[0] + m.vps * (v[1] + m.vps * v[2])));
        break;
      }
      case kQueryTraversable: {  // VoxbloxMap::isTraversable (voxblox.cpp:32-41)
        float d = 0.0f;
        const bool ok = interp_distance_generic(m, __double2float_rn(p.x), __double2float_rn(p.y),
                                                __double2float_rn(p.z), &d);
        static_cast<uint8_t*>(out0)[i] = (ok && static_cast<double>(d) > arg) ? 1 : 0;
        break;
      }
      case kQueryDistance: {
        float d = 0.0f;
        const bool ok = interp_distance_generic(m, __double2float_rn(p.x), __double2float_rn(p.y),
                                                __double2float_rn(p.z), &d);
        static_cast<double*>(out0)[i] = ok ? static_cast<double>(d) : 0.0;
        static_cast<uint8_t*>(out1)[i] = ok ? 1 : 0;
        break;
      }
    }
  }
}

void launch_point_query(cudaStream_t st, const MapView& m, int what, long long n, const double* pts,
                        void* out0, void* out1, double arg) {
  if (n <= 0) return;
  k_point_query<<<grid_for(n, 128), 128, 0, st>>>(m, what, n, pts, out0, out1, arg);
}

__global__ void k_bv_contains(EvalView e, long long n, const double* __restrict__ pts,
                              uint8_t* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = bv_contains(e, D3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}) ? 1 : 0;
}
void launch_bv_contains(cudaStream_t st, const EvalView& e, long long n, const double* pts,
                        uint8_t* out) {
  if (n <= 0) return;
  k_bv_contains<<<grid_for(n, 128), 128, 0, st>>>(e, n, pts, out);
}

// ---------------------------------------------------------------------------------------------
constexpr int kCastBlock = 32;  // one warp per CTA: a warp owns a whole segment
#ifndef A3D_CAST_MIN_BLOCKS
#define A3D_CAST_MIN_BLOCKS 28  // resident CTAs (= warps) per SM the register allocation must allow
#endif

template <int CASTER, int EVAL, bool DIGEST, bool EMIT>
__global__ void __launch_bounds__(kCastBlock, A3D_CAST_MIN_BLOCKS)
    k_cast_gain(const MapView m, const CasterView c, const EvalView e, const Batch b) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const size_t gwarp = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  int8_t* const table = ITER ? b.tables + gwarp * b.table_stride : nullptr;

  __shared__ CenterTables ct;
  __shared__ unsigned long long pw[33];  // kDigestMul^i
  init_center_tables(m, &ct);
  if (DIGEST && threadIdx.x == 0) {
    unsigned long long p = 1;
    for (int i = 0; i < 33; ++i) {
      pw[i] = p;
      p *= kDigestMul;
    }
  }
  __syncthreads();
  const double kNaN = __longlong_as_double(0x7FF8000000000000ll);

  for (;;) {
    int seg = 0;
    if (lane == 0) seg = static_cast<int>(atomicAdd(b.work_counter, 1u));
    seg = __shfl_sync(kFull, seg, 0);
    if (seg >= b.n_segments) break;
    if (b.seg_mask && !((b.seg_mask[seg >> 5] >> (seg & 31)) & 1u)) continue;
    const int v_begin = b.view_first[seg], v_end = b.view_first[seg + 1];

    D3 carry{kNaN, kNaN, kNaN};  // previous raw list entry (NaN: none yet)
    // per-lane counters: VoxelType → [state + 3*outside]; Naive/Frontier → [0]
    unsigned cnt[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long n_vis = 0, n_samp = 0, digest = 0, set_digest = 0;
    const long long emit_base = (EMIT && b.emit_offset) ? b.emit_offset[seg] : 0;
    const long long emit_cap = (EMIT && b.emit_offset) ? b.emit_offset[seg + 1] - emit_base : b.cap;
    double wsum = 0.0;  // VoxelWeightEvaluator: per-lane partial gain
    D3 seg_org{0.0, 0.0, 0.0};
    if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
      if (b.seg_origin) {
        seg_org = D3{b.seg_origin[3 * seg], b.seg_origin[3 * seg + 1], b.seg_origin[3 * seg + 2]};
      } else if (v_end > v_begin) {
        seg_org = D3{b.pos[3 * (v_end - 1)], b.pos[3 * (v_end - 1) + 1], b.pos[3 * (v_end - 1) + 2]};
      }
    }

    for (int view = v_begin; view < v_end; ++view) {
      // camera pose = body pose ∘ mounting transform (camera_model.cpp:24-28)
      D3 P0{b.pos[3 * view], b.pos[3 * view + 1], b.pos[3 * view + 2]};
      Q4 q{b.quat[4 * view], b.quat[4 * view + 1], b.quat[4 * view + 2], b.quat[4 * view + 3]};
      {
        const D3 mt = rotate(q, D3{c.mount_t[0], c.mount_t[1], c.mount_t[2]});
        P0.x = __dadd_rn(P0.x, mt.x);
        P0.y = __dadd_rn(P0.y, mt.y);
        P0.z = __dadd_rn(P0.z, mt.z);
        q = qmul(q, c.mount_q);
      }
      if (ITER) {  // ray_table_ = Zero (iterative_ray_caster.cpp:52)
        int4* t4 = reinterpret_cast<int4*>(table);
        for (size_t x = lane; x < b.table_stride / 16; x += 32) t4[x] = make_int4(0, 0, 0, 0);
        __syncwarp();
      }

      // rays in scan order (i-major), 32 at a time: lane L prepares ray g+L (direction, start
      // section), then the warp walks them one by one
      for (int g = 0; g < c.n_rays; g += 32) {
        const int my_r = g + static_cast<int>(lane);
        const bool my_valid = my_r < c.n_rays;
        const int my_ri = my_valid ? my_r / c.res_y : 0;
        const int my_rj = my_valid ? my_r - my_ri * c.res_y : 0;
        D3 my_dir{0.0, 0.0, 0.0};
        int my_s = -1;
        if (my_valid) {
          my_dir = rotate(q, D3{c.cam_dirs[3 * my_r], c.cam_dirs[3 * my_r + 1],
                                c.cam_dirs[3 * my_r + 2]});
          my_s = ITER ? static_cast<int>(table[my_r]) : 0;
        }
        const int n_in_group = min(32, c.n_rays - g);
        for (int l = 0; l < n_in_group; ++l) {
          const int s = __shfl_sync(kFull, my_s, l);
          if (s < 0) continue;  // already occluded ray
          const D3 dir{__shfl_sync(kFull, my_dir.x, l), __shfl_sync(kFull, my_dir.y, l),
                       __shfl_sync(kFull, my_dir.z, l)};
          const double* __restrict__ ds = c.dseq + static_cast<size_t>(s) * c.kmax;
          const int K = c.kcount[s];
          int hit_k = -1;

          for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + static_cast<int>(lane);
            const bool valid = k < K;
            const double d = valid ? ds[k] : 0.0;
            const D3 Pd{__dadd_rn(P0.x, __dmul_rn(d, dir.x)), __dadd_rn(P0.y, __dmul_rn(d, dir.y)),
                        __dadd_rn(P0.z, __dmul_rn(d, dir.z))};
            // shared by getVoxelState and getVoxelCenter: float point, block index, block origin
            const float pf[3] = {__double2float_rn(Pd.x), __double2float_rn(Pd.y),
                                 __double2float_rn(Pd.z)};
            int bi[3];
            float origin[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              bi[a] = grid_index(pf[a], m.bs_inv_f);
              origin[a] = origin_point(bi[a], m.bs_f);
            }
            const int slot = find_slot(m, bi[0], bi[1], bi[2]);
            int st = 1, vs_i[3];
            uint32_t word;
            bool v_ok;
            if (valid) st = voxel_state_in_block(m, ct, pf, bi, origin, slot, vs_i, &word, &v_ok);
            const unsigned occ = __ballot_sync(kFull, valid && st == 0);
            int n_emit = min(32, K - k0);
            if (occ) {
              const int h = __ffs(occ) - 1;
              hit_k = k0 + h;
              // Simple: state tested before the push (hit excluded, simple_ray_caster.cpp:52-61);
              // Iterative: pushed, then tested (hit included, iterative_ray_caster.cpp:81-93)
              n_emit = ITER ? h + 1 : h;
            }
            int vi[3];
            const D3 cc = voxel_center_in_block(m, ct, Pd, origin, vi);
            double px = __shfl_up_sync(kFull, cc.x, 1);
            double py = __shfl_up_sync(kFull, cc.y, 1);
            double pz = __shfl_up_sync(kFull, cc.z, 1);
            if (lane == 0) {
              px = carry.x;
              py = carry.y;
              pz = carry.z;
            }
            bool keep =
                (static_cast<int>(lane) < n_emit) && !(cc.x == px && cc.y == py && cc.z == pz);
            if (n_emit > 0) {
              carry.x = __shfl_sync(kFull, cc.x, n_emit - 1);
              carry.y = __shfl_sync(kFull, cc.y, n_emit - 1);
              carry.z = __shfl_sync(kFull, cc.z, n_emit - 1);
            }
            bool inside = true;
            if (EVAL != kEvalNone) {
              if (e.bv_is_setup) {
                inside = bv_contains(e, cc);
                keep = keep && inside;  // simulated_sensor_evaluator.cpp:50-57
              }
            }
            const unsigned km = __ballot_sync(kFull, keep);
            const int nk = __popc(km);
            const int rank = __popc(km & lt_mask);
            if (EMIT) {
              const long long idx = static_cast<long long>(n_vis) + rank;
              if (keep && idx < emit_cap) {
                if (b.keys_out) {
                  unsigned bad = 0;
                  b.keys_out[emit_base + idx] = key_from_block_voxel(m, ct, bi, vi, &bad);
                  if (bad) *b.key_error = 1u;
                } else {
                  double* dst = b.centres_out + 3 * (emit_base + idx);
                  dst[0] = cc.x;
                  dst[1] = cc.y;
                  dst[2] = cc.z;
                }
              }
            }
            if (DIGEST) {
              const unsigned long long eh = keep ? digest_entry(cc) : 0ull;
              unsigned long long contrib = keep ? eh * pw[nk - 1 - rank] : 0ull;
              unsigned long long plain = eh;
#pragma unroll
              for (int off = 16; off > 0; off >>= 1) {
                contrib += __shfl_xor_sync(kFull, contrib, off);
                plain += __shfl_xor_sync(kFull, plain, off);
              }
              digest = digest * pw[nk] + contrib;
              set_digest += plain;
            }
            n_vis += nk;
            if (EVAL != kEvalNone) {
              if (keep) {
                // the entry's centre lies in block bi (same block index as the sample point)
                const int cs = centre_state(m, ct, cc, slot, vi);
                if (EVAL == A3D_EVAL_VOXEL_TYPE) {
                  cnt[cs & 3]++;  // retained entries are always inside the bounding volume
                } else if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
                  const double w = (cs & 3) == 0 ? voxel_weight(m, cc) : 0.0;
                  wsum += voxel_weight_value(m, ct, e, cc, cs, w, seg_org);
                } else if (!(cs & 4)) {
                  if (EVAL == A3D_EVAL_NAIVE) {
                    cnt[0]++;
                  } else if (is_frontier_voxel(m, ct, e, cc)) {
                    cnt[0]++;
                  }
                }
              }
            }
            if (occ) break;
          }
          n_samp += (hit_k >= 0) ? (hit_k + 1) : K;

          if (ITER) {
            // markNeighboringRays calls of this ray, applied in one pass (DESIGN.md "ray table"):
            // completed sections t = s..t_last each mark the w[t]-square with t+1; a hit in section
            // seg_h marks its square with -1 last.
            const bool hit = hit_k >= 0;
            int seg_h = s;
            if (hit) {
              while (hit_k >= c.seg_end[s * c.n_sections + seg_h]) ++seg_h;
            }
            const int t_last = hit ? seg_h : c.n_sections - 2;
            const int w0 = 1 << (c.n_sections - 1 - s);
            if (t_last >= s && w0 > 1) {
              const int r = g + l;
              const int ri = r / c.res_y, rj = r - ri * c.res_y;
              const int nx = min(w0, c.res_x - ri), ny = min(w0, c.res_y - rj);
              for (int cell = lane; cell < nx * ny; cell += 32) {
                const int dx = cell / ny, dy = cell - dx * ny;
                const int mx = max(dx, dy);
                const int tmax = mx == 0 ? c.n_sections - 1 : c.n_sections - 2 - (31 - __clz(mx));
                const int t = min(tmax, t_last);
                const int val = (hit && t == seg_h) ? -1 : t + 1;
                table[(ri + dx) * c.res_y + rj + dy] = static_cast<int8_t>(val);
              }
              // the same marks applied to the start sections already held in registers
              if (my_valid) {
                const int dx = my_ri - ri, dy = my_rj - rj;
                if (dx >= 0 && dy >= 0 && dx < nx && dy < ny) {
                  const int mx = max(dx, dy);
                  const int tmax = mx == 0 ? c.n_sections - 1 : c.n_sections - 2 - (31 - __clz(mx));
                  const int t = min(tmax, t_last);
                  my_s = (hit && t == seg_h) ? -1 : t + 1;
                }
              }
              __syncwarp();
            }
          }
        }  // rays of the group
      }    // ray groups
    }      // views

    // per-segment reduction and write-out
    unsigned long long tot[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) tot[i] = __reduce_add_sync(kFull, cnt[i]);
    if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(kFull, wsum, off);
    }
    if (lane == 0) {
      if (EVAL == A3D_EVAL_VOXEL_TYPE) {
        b.gains[seg] = voxel_type_gain(e, tot);
      } else if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
        b.gains[seg] = wsum;
      } else if (EVAL != kEvalNone) {
        b.gains[seg] = static_cast<double>(tot[0]);
      } else if (b.gains) {
        b.gains[seg] = 0.0;
      }
      if (b.n_visible) b.n_visible[seg] = n_vis;
      if (b.n_samples) b.n_samples[seg] = n_samp;
      if (DIGEST && b.digests) b.digests[seg] = digest;
      if (DIGEST && b.set_digests) b.set_digests[seg] = set_digest;
    }
  }
}

template <int CASTER, int EVAL>
static void launch_cast_gain_t(cudaStream_t st, const MapView& m, const CasterView& c,
                               const EvalView& e, const Batch& b, bool digest, bool emit, int grid,
                               int block) {
  if (digest && emit)
    k_cast_gain<CASTER, EVAL, true, true><<<grid, block, 0, st>>>(m, c, e, b);
  else if (digest)
    k_cast_gain<CASTER, EVAL, true, false><<<grid, block, 0, st>>>(m, c, e, b);
  else if (emit)
    k_cast_gain<CASTER, EVAL, false, true><<<grid, block, 0, st>>>(m, c, e, b);
  else
    k_cast_gain<CASTER, EVAL, false, false><<<grid, block, 0, st>>>(m, c, e, b);
}

template <int CASTER>
static void launch_cast_gain_c(cudaStream_t st, const MapView& m, const CasterView& c,
                               const EvalView* e, const Batch& b, bool digest, bool emit, int grid,
                               int block) {
  EvalView none{};
  if (!e) {
    launch_cast_gain_t<CASTER, kEvalNone>(st, m, c, none, b, digest, emit, grid, block);
    return;
  }
  switch (e->kind) {
    case A3D_EVAL_VOXEL_TYPE:
      launch_cast_gain_t<CASTER, A3D_EVAL_VOXEL_TYPE>(st, m, c, *e, b, digest, emit, grid, block);
      break;
    case A3D_EVAL_NAIVE:
      launch_cast_gain_t<CASTER, A3D_EVAL_NAIVE>(st, m, c, *e, b, digest, emit, grid, block);
      break;
    case A3D_EVAL_VOXEL_WEIGHT:
      launch_cast_gain_t<CASTER, A3D_EVAL_VOXEL_WEIGHT>(st, m, c, *e, b, digest, emit, grid, block);
      break;
    default:
      launch_cast_gain_t<CASTER, A3D_EVAL_FRONTIER>(st, m, c, *e, b, digest, emit, grid, block);
      break;
  }
}

void launch_cast_gain(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView* e,
                      const Batch& b, bool digest, bool emit, int grid, int block) {
  if (c.kind == A3D_CASTER_ITERATIVE)
    launch_cast_gain_c<A3D_CASTER_ITERATIVE>(st, m, c, e, b, digest, emit, grid, block);
  else
    launch_cast_gain_c<A3D_CASTER_SIMPLE>(st, m, c, e, b, digest, emit, grid, block);
}

int cast_gain_block_threads() { return kCastBlock; }

// ---------------------------------------------------------------------------------------------
// clear_from_parents (simulated_sensor_evaluator.cpp:60-76): keep[i] = 0 if centre i occurs anywhere
// in the ancestor's stored list (exact double comparison, like std::find on Eigen::Vector3d).
// The ancestor's list goes into an open-addressing hash set of entry indices (linear probing, load
// <= 1/2), the new list probes it: O(n + m) instead of the n x m comparisons of the literal loop.
// Equality is operator== on the three doubles: -0.0 equals 0.0 (hashed as 0.0), a NaN equals nothing
// (never inserted, never found).
__device__ __forceinline__ bool centre_hash(const double* __restrict__ c, uint32_t* h_out) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double v = c[k];
    ok &= (v == v);
    unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v == 0.0 ? 0.0 : v));
    h = (h ^ bits) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 29;
  }
  *h_out = static_cast<uint32_t>(h ^ (h >> 32));
  return ok;
}
__device__ __forceinline__ bool centre_equal(const double* __restrict__ a, const double* __restrict__ b) {
  return (a[0] == b[0]) & (a[1] == b[1]) & (a[2] == b[2]);
}

__global__ void k_centre_set_build(long long m, const double* __restrict__ other, int* __restrict__ table,
                                   uint32_t mask) {
  const long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (j >= m) return;
  const double* c = other + 3 * j;
  uint32_t h;
  if (!centre_hash(c, &h)) return;
  for (uint32_t s = h & mask;; s = (s + 1) & mask) {
    const int old = atomicCAS(&table[s], -1, static_cast<int>(j));
    if (old < 0) return;                                     // claimed an empty slot
    if (centre_equal(other + 3 * static_cast<long long>(old), c)) return;  // already present
  }
}

__global__ void k_filter_not_in(long long n, const double* __restrict__ a, const double* __restrict__ other,
                                const int* __restrict__ table, uint32_t mask, uint8_t* __restrict__ keep) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double* c = a + 3 * i;
  uint32_t h;
  if (!centre_hash(c, &h)) return;
  for (uint32_t s = h & mask;; s = (s + 1) & mask) {
    const int j = table[s];
    if (j < 0) return;
    if (centre_equal(other + 3 * static_cast<long long>(j), c)) {
      keep[i] = 0;
      return;
    }
  }
}

// table: filter_table_slots(m) ints of scratch
size_t filter_table_slots(long long m) {
  size_t cap = 64;
  while (cap < 2 * static_cast<size_t>(m)) cap <<= 1;
  return cap;
}

void launch_filter_not_in(cudaStream_t st, long long n, const double* a, long long m,
                          const double* other, uint8_t* keep, int* table) {
  if (n <= 0 || m <= 0) return;
  const size_t cap = filter_table_slots(m);
  cudaMemsetAsync(table, 0xFF, cap * sizeof(int), st);
  k_centre_set_build<<<static_cast<unsigned>((m + 255) / 256), 256, 0, st>>>(m, other, table,
                                                                             static_cast<uint32_t>(cap - 1));
  k_filter_not_in<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(n, a, other, table,
                                                                          static_cast<uint32_t>(cap - 1), keep);
}

// ---------------------------------------------------------------------------------------------
// K4: fold of a stored list.  counters[0..5] VoxelType state counts / [0] gain count, [6] survivors.
__global__ void k_gain_from_voxels(MapView m, EvalView e, long long n,
                                   const double* __restrict__ centres,
                                   unsigned long long* __restrict__ counters,
                                   uint8_t* __restrict__ keep, D3 origin, double* __restrict__ gain_acc) {
  __shared__ CenterTables ct;
  init_center_tables(m, &ct);
  __syncthreads();
  unsigned cnt[6] = {0, 0, 0, 0, 0, 0};
  unsigned left = 0;
  double wsum = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const D3 c{centres[3 * i], centres[3 * i + 1], centres[3 * i + 2]};
    bool kp = true;
    uint32_t word = 0;
    const bool have = centre_word(m, ct, c, &word);
    if (e.kind == A3D_EVAL_VOXEL_TYPE) {
      const int s = have ? static_cast<int>((word >> 16) & 3u) : voxel_state(m, ct, c);
      cnt[s + (bv_contains(e, c) ? 0 : 3)]++;
    } else if (e.kind == A3D_EVAL_VOXEL_WEIGHT) {
      const int s = have ? static_cast<int>((word >> 16) & 3u) : voxel_state(m, ct, c);
      const int hint = (have && s == 2 && e.frontier_voxel_weight > 0.0) ? is_frontier_from_word(m, e, word) : -1;
      wsum += voxel_weight_value(m, ct, e, c, s, s == 0 ? voxel_weight(m, c) : 0.0, origin, hint);
    } else {
      kp = have ? !((word >> 18) & 1u) : !is_observed(m, c);
      if (kp && e.kind == A3D_EVAL_NAIVE) cnt[0]++;
      if (kp && e.kind == A3D_EVAL_FRONTIER) {
        const int hint = have ? is_frontier_from_word(m, e, word) : -1;
        if (hint >= 0 ? hint != 0 : is_frontier_voxel(m, ct, e, c)) cnt[0]++;
      }
    }
    if (kp) left++;
    if (keep) keep[i] = kp ? 1 : 0;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const unsigned t = __reduce_add_sync(kFull, cnt[i]);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(counters + i, static_cast<unsigned long long>(t));
  }
  const unsigned tl = __reduce_add_sync(kFull, left);
  if ((threadIdx.x & 31) == 0 && tl) atomicAdd(counters + 6, static_cast<unsigned long long>(tl));
  if (e.kind == A3D_EVAL_VOXEL_WEIGHT) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(kFull, wsum, off);
    if ((threadIdx.x & 31) == 0 && wsum != 0.0) atomicAdd(gain_acc, wsum);
  }
}

void launch_gain_from_voxels(cudaStream_t st, const MapView& m, const EvalView& e, long long n,
                             const double* centres, unsigned long long* counters, uint8_t* keep,
                             D3 origin, double* gain_acc) {
  if (n <= 0) return;
  k_gain_from_voxels<<<grid_for(n, 128), 128, 0, st>>>(m, e, n, centres, counters, keep, origin,
                                                       gain_acc);
}

}  // namespace a3d
