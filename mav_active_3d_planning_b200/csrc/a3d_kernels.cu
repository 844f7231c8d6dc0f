// a3d_kernels.cu — sm_100a kernels of the view-gain path.
//
//   k_apply_blocks       K1: merge an uploaded block delta (distance + observed flag) into the slab
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

#include "../../include/a3d.h"

namespace a3d {

constexpr unsigned kFull = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
__global__ void k_apply_blocks(int n, int nv, const int32_t* __restrict__ slots,
                               const float* __restrict__ dist_src, const uint8_t* __restrict__ obs_src,
                               const float* __restrict__ w_src, float* __restrict__ dist_slab,
                               float* __restrict__ weight_slab) {
  const long long total = static_cast<long long>(n) * nv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int blk = static_cast<int>(i / nv);
    const int lin = static_cast<int>(i - static_cast<long long>(blk) * nv);
    const size_t dst = static_cast<size_t>(slots[blk]) * nv + lin;
    float d = dist_src[i];
    uint32_t bits = __float_as_uint(d);
    if (!obs_src[i]) {
      bits = kUnobservedBits;
    } else if (d != d) {
      bits = 0x7FC00000u;  // canonical NaN for observed voxels
    }
    dist_slab[dst] = __uint_as_float(bits);
    if (weight_slab) weight_slab[dst] = w_src ? w_src[i] : 0.0f;
  }
}

void launch_apply_blocks(cudaStream_t st, int n, int nv, const int32_t* slots_dev,
                         const float* dist_src, const uint8_t* obs_src, const float* w_src,
                         float* dist_slab, float* weight_slab) {
  if (n <= 0) return;
  const long long total = static_cast<long long>(n) * nv;
  const int block = 256;
  const int grid = static_cast<int>(std::min<long long>((total + block - 1) / block, 148LL * 16));
  k_apply_blocks<<<grid, block, 0, st>>>(n, nv, slots_dev, dist_src, obs_src, w_src, dist_slab,
                                         weight_slab);
}

// ---------------------------------------------------------------------------------------------
__global__ void k_point_query(MapView m, int what, long long n, const double* __restrict__ pts,
                              void* out0, void* out1) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const D3 p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    switch (what) {
      case kQueryState:
        static_cast<uint8_t*>(out0)[i] = static_cast<uint8_t>(voxel_state(m, p));
        break;
      case kQueryCenter: {
        int b[3], v[3];
        const D3 c = voxel_center(m, p, b, v);
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
      case kQueryDistance: {
        float d = 0.0f;
        const bool ok = interp_distance(m, __double2float_rn(p.x), __double2float_rn(p.y),
                                        __double2float_rn(p.z), &d);
        static_cast<double*>(out0)[i] = ok ? static_cast<double>(d) : 0.0;
        static_cast<uint8_t*>(out1)[i] = ok ? 1 : 0;
        break;
      }
    }
  }
}

void launch_point_query(cudaStream_t st, const MapView& m, int what, long long n, const double* pts,
                        void* out0, void* out1) {
  if (n <= 0) return;
  const int block = 128;
  const int grid = static_cast<int>(std::min<long long>((n + block - 1) / block, 148LL * 16));
  k_point_query<<<grid, block, 0, st>>>(m, what, n, pts, out0, out1);
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
  const int block = 128;
  const int grid = static_cast<int>(std::min<long long>((n + block - 1) / block, 148LL * 16));
  k_bv_contains<<<grid, block, 0, st>>>(e, n, pts, out);
}

// ---------------------------------------------------------------------------------------------
// Evaluator fold of one retained entry.  cnt: VoxelType → [state + 3*outside]; Naive/Frontier → [0].
template <int EVAL>
__device__ __forceinline__ void fold_entry(const MapView& m, const EvalView& e, const D3& c,
                                           unsigned (&cnt)[6]) {
  if (EVAL == A3D_EVAL_VOXEL_TYPE) {
    const int s = voxel_state(m, c);
    const int outside = bv_contains(e, c) ? 0 : 3;
    cnt[s + outside]++;
  } else if (EVAL == A3D_EVAL_NAIVE) {
    if (!is_observed(m, c)) cnt[0]++;
  } else if (EVAL == A3D_EVAL_FRONTIER) {
    if (!is_observed(m, c)) {
      cnt[1]++;
      if (is_frontier_voxel(m, e, c)) cnt[0]++;
    }
  }
}

// VoxelTypeEvaluator's switch without breaks (voxel_type_evaluator.cpp:69-85): UNKNOWN adds
// unknown+free+occupied, FREE adds free+occupied, OCCUPIED adds occupied.
__device__ __forceinline__ double voxel_type_gain(const EvalView& e, const unsigned long long* n) {
  // n[state + 3*outside]; state 0 OCC, 1 FREE, 2 UNKNOWN; gains: unk, occ, free, unk_o, occ_o, free_o
  double g = 0.0;
  for (int o = 0; o < 2; ++o) {
    const double n_unk = static_cast<double>(n[2 + 3 * o]);
    const double n_free = static_cast<double>(n[1 + 3 * o] + n[2 + 3 * o]);
    const double n_occ = static_cast<double>(n[0 + 3 * o] + n[1 + 3 * o] + n[2 + 3 * o]);
    g += n_unk * e.gains[0 + 3 * o];
    g += n_free * e.gains[2 + 3 * o];
    g += n_occ * e.gains[1 + 3 * o];
  }
  return g;
}

constexpr int kCastBlock = 32;  // one warp per CTA: a warp owns a whole segment

template <int CASTER, int EVAL, bool DIGEST, bool EMIT>
__global__ void __launch_bounds__(kCastBlock)
    k_cast_gain(const MapView m, const CasterView c, const EvalView e, const Batch b) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const size_t gwarp = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  int8_t* const table = ITER ? b.tables + gwarp * b.table_stride : nullptr;

  __shared__ unsigned long long pw[33];  // kDigestMul^i
  if (DIGEST) {
    if (threadIdx.x == 0) {
      unsigned long long p = 1;
      for (int i = 0; i < 33; ++i) {
        pw[i] = p;
        p *= kDigestMul;
      }
    }
    __syncthreads();
  }
  const double kNaN = __longlong_as_double(0x7FF8000000000000ll);

  for (;;) {
    int seg = 0;
    if (lane == 0) seg = static_cast<int>(atomicAdd(b.work_counter, 1u));
    seg = __shfl_sync(kFull, seg, 0);
    if (seg >= b.n_segments) break;
    const int v_begin = b.view_first[seg], v_end = b.view_first[seg + 1];

    D3 carry{kNaN, kNaN, kNaN};  // previous raw list entry (NaN: none yet)
    unsigned cnt[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long n_vis = 0, n_samp = 0, digest = 0;

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

      int ri = 0, rj = 0;  // ray r = ri*res_y + rj, scan order i-major
      for (int r = 0; r < c.n_rays; ++r, ++rj) {
        if (rj == c.res_y) {
          rj = 0;
          ++ri;
        }
        int s = 0;
        if (ITER) {
          s = table[r];
          if (s < 0) continue;  // already occluded ray
        }
        const D3 dir = rotate(q, D3{c.cam_dirs[3 * r], c.cam_dirs[3 * r + 1], c.cam_dirs[3 * r + 2]});
        const double* __restrict__ ds = c.dseq + static_cast<size_t>(s) * c.kmax;
        const int K = c.kcount[s];
        int hit_k = -1;

        for (int k0 = 0; k0 < K; k0 += 32) {
          const int k = k0 + static_cast<int>(lane);
          const bool valid = k < K;
          const double d = valid ? ds[k] : 0.0;
          const D3 P{__dadd_rn(P0.x, __dmul_rn(d, dir.x)), __dadd_rn(P0.y, __dmul_rn(d, dir.y)),
                     __dadd_rn(P0.z, __dmul_rn(d, dir.z))};
          int st = 1;
          if (valid) st = voxel_state(m, P);
          const unsigned occ = __ballot_sync(kFull, valid && st == 0);
          int n_emit = min(32, K - k0);
          if (occ) {
            const int h = __ffs(occ) - 1;
            hit_k = k0 + h;
            // Simple: state tested before the push (hit excluded, simple_ray_caster.cpp:52-61);
            // Iterative: pushed, then tested (hit included, iterative_ray_caster.cpp:81-93)
            n_emit = ITER ? h + 1 : h;
          }
          int bi[3], vi[3];
          const D3 cc = voxel_center(m, P, bi, vi);
          double px = __shfl_up_sync(kFull, cc.x, 1);
          double py = __shfl_up_sync(kFull, cc.y, 1);
          double pz = __shfl_up_sync(kFull, cc.z, 1);
          if (lane == 0) {
            px = carry.x;
            py = carry.y;
            pz = carry.z;
          }
          bool keep = (static_cast<int>(lane) < n_emit) && !(cc.x == px && cc.y == py && cc.z == pz);
          if (n_emit > 0) {
            carry.x = __shfl_sync(kFull, cc.x, n_emit - 1);
            carry.y = __shfl_sync(kFull, cc.y, n_emit - 1);
            carry.z = __shfl_sync(kFull, cc.z, n_emit - 1);
          }
          if (EVAL != kEvalNone) {
            if (keep && e.bv_is_setup) keep = bv_contains(e, cc);
          }
          const unsigned km = __ballot_sync(kFull, keep);
          const int nk = __popc(km);
          const int rank = __popc(km & lt_mask);
          if (EMIT) {
            const long long idx = static_cast<long long>(n_vis) + rank;
            if (keep && idx < b.cap) {
              b.centres_out[3 * idx] = cc.x;
              b.centres_out[3 * idx + 1] = cc.y;
              b.centres_out[3 * idx + 2] = cc.z;
            }
          }
          if (DIGEST) {
            unsigned long long contrib = keep ? digest_entry(cc) * pw[nk - 1 - rank] : 0ull;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) contrib += __shfl_xor_sync(kFull, contrib, off);
            digest = digest * pw[nk] + contrib;
          }
          n_vis += nk;
          if (EVAL != kEvalNone) {
            if (keep) fold_entry<EVAL>(m, e, cc, cnt);
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
            const int nx = min(w0, c.res_x - ri), ny = min(w0, c.res_y - rj);
            for (int cell = lane; cell < nx * ny; cell += 32) {
              const int dx = cell / ny, dy = cell - dx * ny;
              const int mx = max(dx, dy);
              const int tmax = mx == 0 ? c.n_sections - 1 : c.n_sections - 2 - (31 - __clz(mx));
              const int t = min(tmax, t_last);
              const int val = (hit && t == seg_h) ? -1 : t + 1;
              table[(ri + dx) * c.res_y + rj + dy] = static_cast<int8_t>(val);
            }
            __syncwarp();
          }
        }
      }  // rays
    }    // views

    // per-segment reduction and write-out
    unsigned long long tot[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) tot[i] = __reduce_add_sync(kFull, cnt[i]);
    if (lane == 0) {
      if (EVAL == A3D_EVAL_VOXEL_TYPE) {
        b.gains[seg] = voxel_type_gain(e, tot);
      } else if (EVAL != kEvalNone) {
        b.gains[seg] = static_cast<double>(tot[0]);
      } else if (b.gains) {
        b.gains[seg] = 0.0;
      }
      if (b.n_visible) b.n_visible[seg] = n_vis;
      if (b.n_samples) b.n_samples[seg] = n_samp;
      if (DIGEST && b.digests) b.digests[seg] = digest;
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
// K4: fold of a stored list.  counters[0..5] VoxelType state counts / [0] gain count, [6] survivors.
__global__ void k_gain_from_voxels(MapView m, EvalView e, long long n,
                                   const double* __restrict__ centres,
                                   unsigned long long* __restrict__ counters,
                                   uint8_t* __restrict__ keep) {
  unsigned cnt[6] = {0, 0, 0, 0, 0, 0};
  unsigned left = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const D3 c{centres[3 * i], centres[3 * i + 1], centres[3 * i + 2]};
    bool kp = true;
    if (e.kind == A3D_EVAL_VOXEL_TYPE) {
      fold_entry<A3D_EVAL_VOXEL_TYPE>(m, e, c, cnt);
    } else {
      kp = !is_observed(m, c);
      if (kp && e.kind == A3D_EVAL_NAIVE) cnt[0]++;
      if (kp && e.kind == A3D_EVAL_FRONTIER && is_frontier_voxel(m, e, c)) cnt[0]++;
    }
    if (kp) left++;
    if (keep) keep[i] = kp ? 1 : 0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const unsigned t = __reduce_add_sync(kFull, cnt[i]);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(counters + i, static_cast<unsigned long long>(t));
  }
  const unsigned tl = __reduce_add_sync(kFull, left);
  if ((threadIdx.x & 31) == 0 && tl) atomicAdd(counters + 6, static_cast<unsigned long long>(tl));
}

void launch_gain_from_voxels(cudaStream_t st, const MapView& m, const EvalView& e, long long n,
                             const double* centres, unsigned long long* counters, uint8_t* keep) {
  if (n <= 0) return;
  const int block = 128;
  const int grid = static_cast<int>(std::min<long long>((n + block - 1) / block, 148LL * 16));
  k_gain_from_voxels<<<grid, block, 0, st>>>(m, e, n, centres, counters, keep);
}

}  // namespace a3d
