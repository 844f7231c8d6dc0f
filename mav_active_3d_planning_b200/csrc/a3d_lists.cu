// a3d_lists.cu — device-resident visible-voxel lists as packed 64-bit voxel keys (a3d_keys.cuh).
//
// What SimulatedSensorEvaluator does to a list after the caster returned it
// (simulated_sensor_evaluator.cpp:60-91) and what SimulatedSensorUpdater does to it later
// (simulated_sensor_updater.cpp:18-22), on lists that never leave HBM:
//
//   k_list_pipeline   one CTA per list, chunk by chunk, in place:
//       1. clear_from_parents (:60-76): drop entries found in any ancestor's stored list — a probe of
//          that ancestor's hash set per entry instead of std::find over its vector (O(V·ΣV) → O(V·depth));
//       2. the size / ordered digest / set digest of the list the reference stores (:79, :84-91);
//       3. computeGainFromVisibleVoxels: VoxelType / Naive / Frontier / VoxelWeight fold straight from
//          the key (meta word of voxel g, no index arithmetic on doubles); Naive and Frontier erase the
//          observed entries from the stored list like the reference does (naive_evaluator.cpp:28-35,
//          frontier_evaluator.cpp:110-116) — stable compaction.
//   k_build_sets      open-addressing hash set (linear probing, load <= 0.5) of a stored list's keys
//   k_keys_to_centres the centre doubles the reference holds, for a3d_store_get
#include <algorithm>

#include "../../include/a3d.h"
#include "a3d_kernels.cuh"
#include "a3d_keys.cuh"

namespace a3d {

namespace {
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kListThreads = 256;

__device__ __forceinline__ uint32_t key_hash(unsigned long long key) {
  return static_cast<uint32_t>((key * 0x9E3779B97F4A7C15ull) >> 32);
}

__device__ __forceinline__ bool want_digest_any(const ListBatch& lb) { return lb.digests || lb.set_digests; }

// exclusive rank of this thread among the threads of the CTA with flag set + the CTA total
// (two __syncthreads; s_warp: 8 counters)
__device__ __forceinline__ unsigned block_rank(bool flag, unsigned* s_warp, unsigned* total) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(kFull, flag);
  __syncthreads();  // previous use of s_warp is over
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  unsigned before = 0, tot = 0;
#pragma unroll
  for (unsigned w = 0; w < kListThreads / 32; ++w) {
    before += w < warp ? s_warp[w] : 0;
    tot += s_warp[w];
  }
  *total = tot;
  return before + __popc(bal & ((1u << lane) - 1u));
}
}  // namespace

__global__ void __launch_bounds__(kListThreads, 4)
    k_list_pipeline(const MapView m, const EvalView e, const ListBatch lb) {
  __shared__ CenterTables ct;
  __shared__ unsigned s_warp[kListThreads / 32];
  __shared__ unsigned long long s_pw[kListThreads + 1];  // kDigestMul^i
  __shared__ unsigned long long s_acc[2];                // per-chunk digest contributions
  __shared__ double s_wpart[kListThreads / 32];
  __shared__ unsigned long long s_cnt[6];
  init_center_tables(m, &ct);
  if (want_digest_any(lb) && threadIdx.x == 0) {
    unsigned long long p = 1;
    for (int i = 0; i <= kListThreads; ++i) {
      s_pw[i] = p;
      p *= kDigestMul;
    }
  }
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const bool fold = lb.gains != nullptr;
  const bool erases = fold && (e.kind == A3D_EVAL_NAIVE || e.kind == A3D_EVAL_FRONTIER);
  const bool want_digest = lb.digests != nullptr || lb.set_digests != nullptr;

  for (int li = blockIdx.x; li < lb.n; li += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long* const list = lb.pool + lb.offsets[li];
    const long long cnt = lb.counts[li];
    const int a_lo = lb.anc_first ? lb.anc_first[li] : 0, a_hi = lb.anc_first ? lb.anc_first[li + 1] : 0;
    const D3 origin = lb.origins ? D3{lb.origins[3 * li], lb.origins[3 * li + 1], lb.origins[3 * li + 2]}
                                 : D3{0.0, 0.0, 0.0};
    unsigned c[6] = {0, 0, 0, 0, 0, 0};
    double wsum = 0.0;
    unsigned long long digest = 0, set_digest = 0;  // (uniform across the CTA)
    long long n_stored = 0, n_left = 0;             // after the parent filter / after the fold's erasure

    for (long long c0 = 0; c0 < cnt; c0 += kListThreads) {
      const long long i = c0 + threadIdx.x;
      unsigned long long key = kNoKey;
      bool keep = false;
      if (i < cnt) {
        key = list[i];
        keep = true;
        // 1. clear_from_parents: any ancestor's stored list holds this voxel?
        for (int a = a_lo; a < a_hi && keep; ++a) {
          const unsigned long long* __restrict__ tab = lb.anc_table[a];
          const uint32_t mask = lb.anc_mask[a];
          uint32_t h = key_hash(key) & mask;
          for (;;) {
            const unsigned long long t = __ldg(tab + h);
            if (t == key) {
              keep = false;
              break;
            }
            if (t == kNoKey) break;
            h = (h + 1) & mask;
          }
        }
      }
      // 2. the stored list (what storeTrajectoryInformation receives): size and digests
      unsigned nk = 0;
      const unsigned rank = block_rank(keep, s_warp, &nk);  // (syncs: every key of the chunk is in a register)
      D3 v{0.0, 0.0, 0.0};
      bool have_v = false;
      if (want_digest) {
        if (threadIdx.x < 2) s_acc[threadIdx.x] = 0;
        __syncthreads();
        if (keep) {
          v = key_centre(m, ct, key);
          have_v = true;
        }
        const unsigned long long eh = keep ? digest_entry(v) : 0ull;
        unsigned long long contrib = keep ? eh * s_pw[nk - 1 - rank] : 0ull, plain = eh;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          contrib += __shfl_xor_sync(kFull, contrib, off);
          plain += __shfl_xor_sync(kFull, plain, off);
        }
        if (lane == 0) {
          atomicAdd(&s_acc[0], contrib);
          atomicAdd(&s_acc[1], plain);
        }
        __syncthreads();
        digest = digest * s_pw[nk] + s_acc[0];
        set_digest += s_acc[1];
      }
      n_stored += nk;
      // 3. the fold
      bool kp = keep;
      if (fold && keep) {
        uint32_t word = 0;
        const bool have = key_word(m, key, &word);
        if (!have_v && (!have || e.kind == A3D_EVAL_VOXEL_WEIGHT || (e.kind == A3D_EVAL_VOXEL_TYPE && e.bv_is_setup))) {
          v = key_centre(m, ct, key);
          have_v = true;
        }
        if (e.kind == A3D_EVAL_VOXEL_TYPE) {
          const int st = have ? static_cast<int>((word >> 16) & 3u) : voxel_state(m, ct, v);
          c[st + ((!e.bv_is_setup || bv_contains(e, v)) ? 0 : 3)]++;
        } else if (e.kind == A3D_EVAL_VOXEL_WEIGHT) {
          const int st = have ? static_cast<int>((word >> 16) & 3u) : voxel_state(m, ct, v);
          const int hint = (have && st == 2 && e.frontier_voxel_weight > 0.0) ? is_frontier_from_word(m, e, word) : -1;
          wsum += voxel_weight_value(m, ct, e, v, st, st == 0 ? voxel_weight(m, v) : 0.0, origin, hint);
        } else {
          kp = have ? !((word >> 18) & 1u) : !is_observed(m, v);
          if (kp && e.kind == A3D_EVAL_FRONTIER) {
            const int hint = have ? is_frontier_from_word(m, e, word) : -1;
            if (hint < 0 && !have_v) v = key_centre(m, ct, key);
            if (hint >= 0 ? hint != 0 : is_frontier_voxel(m, ct, e, v)) c[0]++;
          } else if (kp) {
            c[0]++;
          }
        }
      }
      // stable compaction of what stays stored: parent-filtered, and fold-erased for Naive/Frontier
      unsigned n_stay = nk, pos = rank;
      if (erases) pos = block_rank(kp, s_warp, &n_stay);
      if (erases ? kp : keep) list[n_left + pos] = key;
      n_left += n_stay;
    }
    // reductions in a fixed order: the VoxelWeight sum is reproducible run to run
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const unsigned t = __reduce_add_sync(kFull, c[k]);
      if (lane == 0 && t) atomicAdd(&s_cnt[k], static_cast<unsigned long long>(t));
    }
    if (fold && e.kind == A3D_EVAL_VOXEL_WEIGHT) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) wsum += __shfl_down_sync(kFull, wsum, off);
      if (lane == 0) s_wpart[warp] = wsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (fold) {
        if (e.kind == A3D_EVAL_VOXEL_TYPE) {
          lb.gains[li] = voxel_type_gain(e, s_cnt);
        } else if (e.kind == A3D_EVAL_VOXEL_WEIGHT) {
          double g = 0.0;
          for (int w = 0; w < kListThreads / 32; ++w) g += s_wpart[w];
          lb.gains[li] = g;
        } else {
          lb.gains[li] = static_cast<double>(s_cnt[0]);
        }
      }
      lb.counts[li] = n_left;
      if (lb.n_stored) lb.n_stored[li] = static_cast<unsigned long long>(n_stored);
      if (lb.digests) lb.digests[li] = digest;
      if (lb.set_digests) lb.set_digests[li] = set_digest;
    }
  }
}

// The re-fold of lists nothing can leave and whose order nobody asks about (VoxelType / VoxelWeight, no
// parent filter, no digests: SimulatedSensorUpdater over a tree): no ranks, no barriers in the loop, no
// write-back — every thread folds its entries with several loads in flight.  One CTA per list; sums in a
// fixed order (the VoxelWeight gain is reproducible run to run).
template <int KIND>
__global__ void __launch_bounds__(kListThreads, 4)
    k_list_fold_plain(const MapView m, const EvalView e, const ListBatch lb) {
  __shared__ CenterTables ct;
  __shared__ double s_wpart[kListThreads / 32];
  __shared__ unsigned long long s_cnt[6];
  init_center_tables(m, &ct);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (int li = blockIdx.x; li < lb.n; li += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long* const list = lb.pool + lb.offsets[li];
    const long long cnt = lb.counts[li];
    const D3 origin = lb.origins ? D3{lb.origins[3 * li], lb.origins[3 * li + 1], lb.origins[3 * li + 2]}
                                 : D3{0.0, 0.0, 0.0};
    unsigned c[6] = {0, 0, 0, 0, 0, 0};
    double wsum = 0.0;
#pragma unroll 4
    for (long long i = threadIdx.x; i < cnt; i += kListThreads) {
      const unsigned long long key = __ldcs(list + i);
      uint32_t word = 0;
      const bool have = key_word(m, key, &word);
      D3 v{0.0, 0.0, 0.0};
      if (!have || KIND == A3D_EVAL_VOXEL_WEIGHT || e.bv_is_setup) v = key_centre(m, ct, key);
      const int st = have ? static_cast<int>((word >> 16) & 3u) : voxel_state(m, ct, v);
      if (KIND == A3D_EVAL_VOXEL_TYPE) {
        c[st + ((!e.bv_is_setup || bv_contains(e, v)) ? 0 : 3)]++;
      } else {
        const int hint = (have && st == 2 && e.frontier_voxel_weight > 0.0) ? is_frontier_from_word(m, e, word) : -1;
        wsum += voxel_weight_value(m, ct, e, v, st, st == 0 ? voxel_weight(m, v) : 0.0, origin, hint);
      }
    }
    if (KIND == A3D_EVAL_VOXEL_TYPE) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const unsigned t = __reduce_add_sync(kFull, c[k]);
        if (lane == 0 && t) atomicAdd(&s_cnt[k], static_cast<unsigned long long>(t));
      }
    } else {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) wsum += __shfl_down_sync(kFull, wsum, off);
      if (lane == 0) s_wpart[warp] = wsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (KIND == A3D_EVAL_VOXEL_TYPE) {
        lb.gains[li] = voxel_type_gain(e, s_cnt);
      } else {
        double g = 0.0;
        for (int w = 0; w < kListThreads / 32; ++w) g += s_wpart[w];
        lb.gains[li] = g;
      }
      if (lb.n_stored) lb.n_stored[li] = static_cast<unsigned long long>(cnt);
    }
  }
}

void launch_list_pipeline(cudaStream_t st, const MapView& m, const EvalView& e, const ListBatch& lb) {
  if (lb.n <= 0) return;
  const bool plain = lb.gains && !lb.anc_first && !lb.digests && !lb.set_digests;
  if (plain && e.kind == A3D_EVAL_VOXEL_TYPE)
    k_list_fold_plain<A3D_EVAL_VOXEL_TYPE><<<std::min(lb.n, 148 * 4), kListThreads, 0, st>>>(m, e, lb);
  else if (plain && e.kind == A3D_EVAL_VOXEL_WEIGHT)
    k_list_fold_plain<A3D_EVAL_VOXEL_WEIGHT><<<std::min(lb.n, 148 * 4), kListThreads, 0, st>>>(m, e, lb);
  else
    k_list_pipeline<<<std::min(lb.n, 148 * 4), kListThreads, 0, st>>>(m, e, lb);
}

// Hash sets of stored lists: tables are pre-filled with kNoKey (0xFF bytes); duplicates collapse.
__global__ void k_build_sets(int n, const unsigned long long* const* __restrict__ lists,
                             const long long* __restrict__ counts, unsigned long long* const* __restrict__ tables,
                             const uint32_t* __restrict__ masks) {
  for (int li = blockIdx.y; li < n; li += gridDim.y) {
    const unsigned long long* list = lists[li];
    unsigned long long* tab = tables[li];
    const uint32_t mask = masks[li];
    const long long cnt = counts[li];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < cnt;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const unsigned long long key = list[i];
      uint32_t h = key_hash(key) & mask;
      for (;;) {
        const unsigned long long old = atomicCAS(tab + h, kNoKey, key);
        if (old == kNoKey || old == key) break;
        h = (h + 1) & mask;
      }
    }
  }
}

void launch_build_sets(cudaStream_t st, int n, const unsigned long long* const* lists, const long long* counts,
                       unsigned long long* const* tables, const uint32_t* masks) {
  if (n <= 0) return;
  const dim3 grid(8, static_cast<unsigned>(std::min(n, 148 * 4)));
  k_build_sets<<<grid, 256, 0, st>>>(n, lists, counts, tables, masks);
}

__global__ void k_keys_to_centres(const MapView m, long long n, const unsigned long long* __restrict__ keys,
                                  double* __restrict__ centres) {
  __shared__ CenterTables ct;
  init_center_tables(m, &ct);
  __syncthreads();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const D3 c = key_centre(m, ct, keys[i]);
    centres[3 * i] = c.x;
    centres[3 * i + 1] = c.y;
    centres[3 * i + 2] = c.z;
  }
}

void launch_keys_to_centres(cudaStream_t st, const MapView& m, long long n, const unsigned long long* keys,
                            double* centres) {
  if (n <= 0) return;
  const long long blocks = std::min<long long>((n + 255) / 256, 148 * 16);
  k_keys_to_centres<<<static_cast<unsigned>(blocks), 256, 0, st>>>(m, n, keys, centres);
}

}  // namespace a3d
