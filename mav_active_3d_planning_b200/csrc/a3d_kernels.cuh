// a3d_kernels.cuh — kernel-side views of caster / batch state + launch prototypes.
#pragma once

#include "a3d_device.cuh"

namespace a3d {

// Host-precomputed, view-independent tables of one CameraModel subclass.
struct CasterView {
  int32_t kind;        // A3D_CASTER_*
  int32_t res_x, res_y, n_rays;
  int32_t n_sections;  // iterative only (0 for simple)
  int32_t n_starts;    // number of distance sequences (1 for simple, n_sections for iterative)
  int32_t kmax;        // row stride of dseq
  int32_t _pad;
  const double* __restrict__ cam_dirs;  // n_rays×3 camera-frame unit directions, ray r = i*res_y + j
  const double* __restrict__ dseq;      // n_starts×kmax accumulated sample distances
  const int32_t* __restrict__ kcount;   // n_starts: samples of a full (unoccluded) ray
  const int32_t* __restrict__ seg_end;  // n_starts×n_sections: #samples lying in sections <= t
  double mount_t[3];
  Q4 mount_q;
  double ray_step, ray_length;
  double split[33];    // c_split_distances_ (iterative)
  double leaf_d0[33];  // per start section: first accumulated sample distance >= split[n-1] (or ray_length)
  // leaf-supply units of k_cast_wave: iterative → unit u covers rays (i, diag-i), i in
  // [unit_i0[u], unit_i0[u]+32) of anti-diagonal unit_diag[u]; simple → rays [32u, 32u+32)
  int32_t n_units;
  int32_t diag_max;  // longest anti-diagonal = min(res_x, res_y)
  const int32_t* __restrict__ unit_diag;
  const int32_t* __restrict__ unit_i0;
};

struct Batch {
  const double* __restrict__ pos;          // n_views×3
  const double* __restrict__ quat;         // n_views×4 (w,x,y,z)
  const int32_t* __restrict__ view_first;  // n_segments+1
  int32_t n_segments;
  int32_t _pad;
  double* gains;
  unsigned long long* n_visible;
  unsigned long long* n_samples;
  unsigned long long* digests;      // order-sensitive digest (k_cast_gain only)
  unsigned long long* set_digests;  // order-independent digest (both kernels)
  const unsigned int* seg_mask;     // k_cast_gain: if set, only segments whose bit is 1 are evaluated
  const double* seg_origin;         // n_segments×3 or null (VoxelWeightEvaluator's origin)
  double* centres_out;  // emit mode: cap×3 (single segment) or the pool the offsets index
  unsigned long long* keys_out;  // emit mode, instead of centres_out: packed voxel keys (a3d_keys.cuh)
  unsigned int* key_error;       // set to 1 when an entry does not fit a key (keys_out only)
  long long cap;
  const long long* emit_offset;  // emit mode, many segments: n_segments+1 entry offsets into centres_out
  int8_t* tables;       // per-warp ray tables (iterative), table_stride bytes each
  size_t table_stride;
  unsigned int* work_counter;
};

enum { kEvalNone = -1 };

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

// One batch of device-resident lists for k_list_pipeline (a3d_lists.cu); every array has n entries
// unless noted.
struct ListBatch {
  int n;
  unsigned long long* pool;        // packed voxel keys of the arena
  const long long* offsets;        // where each list starts in pool
  long long* counts;               // in: list sizes; out: sizes of what stays stored
  unsigned long long* n_stored;    // out (nullable): size after the parent filter, before the fold's erasure
  unsigned long long* digests;     // out (nullable): ordered digest of that list
  unsigned long long* set_digests; // out (nullable)
  double* gains;                   // out (nullable: no fold)
  const double* origins;           // n×3 (VoxelWeightEvaluator) or null
  const int* anc_first;            // n+1 offsets into anc_table / anc_mask, or null: no parent filter
  const unsigned long long* const* anc_table;  // hash set of each ancestor's stored list
  const uint32_t* anc_mask;        // its size - 1
};

struct LaunchCfg {
  int grid, block;
};

// launchers (a3d_kernels.cu)
void launch_apply_blocks(cudaStream_t st, const MapView& m, int n, const int32_t* slots_dev,
                         const float* dist_src, const uint8_t* obs_src, const float* w_src,
                         float* dist_slab, float* weight_slab, float* tsdf_slab /* nullable */);
void launch_refresh_aprons(cudaStream_t st, const MapView& m, int n, const int32_t* slots_dev,
                           const int32_t* nbr_dev /*n×27*/, float* dist_slab);
void launch_build_meta(cudaStream_t st, const MapView& m, int n, const int32_t* slots_dev,
                       const int32_t* block_idx_dev /*n×3*/, uint32_t* mword_slab,
                       uint32_t* vol /*dense meta volume described by m.vol_lo / m.vol_n, or null*/);
// frontier bits of the meta words (after launch_build_meta of all affected blocks)
void launch_build_frontier(cudaStream_t st, const MapView& m, int n, const int32_t* slots_dev,
                           const int32_t* block_idx_dev /*n×3*/, uint32_t* mword_slab, uint32_t* vol);
void launch_vol_frontier_unallocated(cudaStream_t st, const MapView& m, uint32_t* vol, int n,
                                     const int32_t* block_idx_dev /*n×3, unallocated blocks inside the box*/);
// dense meta volume maintenance: fill everything / the listed blocks with "no block", copy the
// listed blocks' meta words in
void launch_vol_fill(cudaStream_t st, const MapView& m, uint32_t* vol, int n_blocks,
                     const int32_t* block_idx_dev /*n×3, or null: whole volume*/);
void launch_vol_scatter(cudaStream_t st, const MapView& m, uint32_t* vol, int n, const int32_t* slots_dev,
                        const int32_t* block_idx_dev /*n×3*/);
void launch_point_query(cudaStream_t st, const MapView& m, int what, long long n, const double* pts,
                        void* out0, void* out1, double arg = 0.0);
void launch_bv_contains(cudaStream_t st, const EvalView& e, long long n, const double* pts,
                        uint8_t* out);
void launch_cast_gain(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView* e,
                      const Batch& b, bool digest, bool emit, int grid, int block);
void launch_gain_from_voxels(cudaStream_t st, const MapView& m, const EvalView& e, long long n,
                             const double* centres, unsigned long long* counters /*8*/,
                             uint8_t* keep, D3 origin, double* gain_acc /*1, VoxelWeight*/);
size_t filter_table_slots(long long m);  // ints of scratch launch_filter_not_in needs for a list of m entries
void launch_filter_not_in(cudaStream_t st, long long n, const double* a, long long m,
                          const double* other, uint8_t* keep, int* table);
void launch_list_pipeline(cudaStream_t st, const MapView& m, const EvalView& e, const ListBatch& lb);
void launch_build_sets(cudaStream_t st, int n, const unsigned long long* const* lists, const long long* counts,
                       unsigned long long* const* tables, const uint32_t* masks);
void launch_keys_to_centres(cudaStream_t st, const MapView& m, long long n, const unsigned long long* keys,
                            double* centres);
int cast_gain_block_threads();
// throughput kernel (a3d_wave.cu)
size_t wave_scratch_bytes(int n_rays, int grid, bool digest, size_t* rays_stride,
                          int emit_kmax /*> 0: list emission, slots per ray part*/);
void launch_cast_wave(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e,
                      const Batch& b, bool digest, int grid, void* scratch, unsigned int* inexact,
                      int warps /*CTA width: 2 (throughput), 4 or 8 (mid-size / small batches)*/);
// throughput path as a chain of kernels (a3d_split.cu): gains / counts and, when Batch::centres_out or
// keys_out is set, the ordered lists; no digests; dense meta volume required
bool cast_split_supported(const MapView& m, const CasterView& c);
// mode 0: gains / counts; 1: list emission (b.centres_out / b.keys_out); 2: gains / counts + every ray part's
// offset within its segment's list, for launch_cast_split_direct on the same scratch afterwards
size_t cast_split_bytes_per_view(const CasterView& c, int eval_kind, size_t* stride_out, int mode);
int cast_split_launches(const CasterView& c, int mode);
// stage_events (6, or null): recorded before the first and after each kernel; returns how many were recorded
int launch_cast_split(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                      int seg0, int n_seg, int view0, int n_views, void* scratch, unsigned int* inexact,
                      int sm_count, cudaEvent_t* stage_events, int mode);
// second pass of a stored-list call (a3d_split_direct.cu): casts the whole batch again and writes the keys of
// the retained entries to b.keys_out[b.emit_offset[segment] + place in the list]
int launch_cast_split_direct(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                             int n_seg, int n_views, void* scratch, int sm_count);
int cast_wave_threads();
int cast_wave_max_diag();
int cast_wave_blocks_per_sm();

enum { kQueryState = 0, kQueryCenter = 1, kQueryObserved = 2, kQueryWeight = 3, kQueryDistance = 4,
       kQueryStateGeneric = 5, kQueryTraversable = 6,
       kQueryTsdfDistance = 7 /* out1 = the TSDF distance slab (may be null) */ };

}  // namespace a3d
