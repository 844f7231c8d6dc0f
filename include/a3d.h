/*
 * a3d.h — C ABI of the B200 view-gain library (liba3d_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of ethz-asl/mav_active_3d_planning:
 * SimulatedSensorEvaluator::computeGain → CameraModel ray casters → voxblox map lookups →
 * VoxelType/Naive/Frontier gain fold.  Each entry point names the reference interface it replaces
 * (paths relative to the reference checkout).  Plain pointers and sizes only; all pointers are HOST
 * memory unless the function name ends in _dev.  The caller owns every buffer.  One thread per
 * context (thread-compatible, not thread-safe) — like the reference, which is single-threaded
 * (active_3d_planning_ros/src/planner/ros_planner.cpp:93-105).
 *
 * There is no CPU fallback: every function that computes does so in sm_100a CUDA kernels and
 * returns A3D_ERR_CUDA if no device is usable.
 *
 * All functions return 0 (A3D_OK) or a negative error code; a3d_last_error() gives the text.
 */
#ifndef A3D_H_
#define A3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A3D_OK 0
#define A3D_ERR_INVALID (-1)     /* bad argument (the reference throws std::invalid_argument,
                                    active_3d_planning_core/src/module/module.cpp:11-16) */
#define A3D_ERR_CUDA (-2)        /* CUDA runtime error / no device */
#define A3D_ERR_STATE (-3)       /* call order (e.g. upload before a3d_map_config) */
#define A3D_ERR_CAPACITY (-4)    /* caller buffer too small (a3d_visible_voxels) */
#define A3D_ERR_UNSUPPORTED (-5)

/* OccupancyMap states, active_3d_planning_core/include/active_3d_planning_core/map/occupancy_map.h:17-19 */
#define A3D_OCCUPIED 0
#define A3D_FREE 1
#define A3D_UNKNOWN 2

/* sensor model kinds = registered type strings "SimpleRayCaster" / "IterativeRayCaster"
   (core/src/module/sensor_model/simple_ray_caster.cpp:10-11, iterative_ray_caster.cpp:12-13) */
#define A3D_CASTER_SIMPLE 0
#define A3D_CASTER_ITERATIVE 1
/* "IterativeRayCasterLidar" (iterative_ray_caster_lidar.cpp:12-13): the iterative scheme over the
   spherical direction model of LidarModel (lidar_model.cpp:159-165) */
#define A3D_CASTER_ITERATIVE_LIDAR 2

/* evaluator kinds = "VoxelTypeEvaluator" / "NaiveEvaluator" / "FrontierEvaluator"
   (core/src/module/trajectory_evaluator/voxel_type_evaluator.cpp:13-14, naive_evaluator.cpp:9-10,
   frontier_evaluator.cpp:9-10) */
#define A3D_EVAL_VOXEL_TYPE 0
#define A3D_EVAL_NAIVE 1
#define A3D_EVAL_FRONTIER 2
/* "VoxelWeightEvaluator" (voxel_weight_evaluator.cpp:12-13) */
#define A3D_EVAL_VOXEL_WEIGHT 3

typedef struct a3d_ctx a3d_ctx;
typedef struct a3d_caster a3d_caster;
typedef struct a3d_eval a3d_eval;
typedef struct a3d_store a3d_store;
typedef struct a3d_group a3d_group;
typedef struct a3d_group_caster a3d_group_caster;
typedef struct a3d_group_eval a3d_group_eval;

/* Parameters of a CameraModel subclass exactly as setupFromParamMap parses them
   (core/src/module/sensor_model/sensor_model.cpp:9-27, camera_model.cpp:170-183,
   simple_ray_caster.cpp:15-30, iterative_ray_caster.cpp:18-46). */
typedef struct {
  int32_t kind;               /* A3D_CASTER_* */
  int32_t resolution_x;       /* "resolution_x" (640) */
  int32_t resolution_y;       /* "resolution_y" (480) */
  int32_t _pad;
  double ray_length;          /* "ray_length" (5.0) */
  double focal_length;        /* "focal_length" (320.0) */
  double ray_step;            /* "ray_step"; <= 0 selects the default = map voxel size */
  double downsampling_factor; /* "downsampling_factor" (1.0) */
  double mount_t[3];          /* "mounting_translation_{x,y,z}" */
  double mount_q[4];          /* "mounting_rotation_{w,x,y,z}", normalised by the caller like
                                 sensor_model.cpp:27 does */
  double field_of_view_x;     /* lidar only: "field_of_view_x" [deg] (360.0), lidar_model.cpp:170 */
  double field_of_view_y;     /* lidar only: "field_of_view_y" [deg] (30.0) */
} a3d_caster_params;

/* Parameters of a SimulatedSensorEvaluator subclass + its BoundingVolume
   (core/src/module/trajectory_evaluator/simulated_sensor_evaluator.cpp:16-43,
   voxel_type_evaluator.cpp:19-55, frontier_evaluator.cpp:15-67, core/src/data/bounding_volume.cpp:12-31). */
typedef struct {
  int32_t kind;               /* A3D_EVAL_* */
  int32_t clear_from_parents; /* "clear_from_parents" (false) */
  int32_t accurate_frontiers; /* "accurate_frontiers" (false) */
  int32_t surface_frontiers;  /* "surface_frontiers" (true) */
  double checking_distance;   /* "checking_distance" (1.0) */
  double gains[6];            /* gain_unknown, gain_occupied, gain_free, gain_unknown_outer,
                                 gain_occupied_outer, gain_free_outer */
  int32_t bv_is_setup;        /* BoundingVolume::is_setup */
  int32_t _pad;
  double bv_min[3];           /* x_min, y_min, z_min */
  double bv_max[3];           /* x_max, y_max, z_max */
  double bv_quat[4];          /* w,x,y,z of AngleAxis(rotation*pi/-180, UnitZ) (bounding_volume.cpp:21) */
  /* VoxelWeightEvaluator (voxel_weight_evaluator.cpp:17-22): "frontier_voxel_weight" (1.0),
     "min_impact_factor" (0.0), "new_voxel_weight" (0.01), "ray_angle_x" / "ray_angle_y" (0.0025) */
  double frontier_voxel_weight, min_impact_factor, new_voxel_weight, ray_angle_x, ray_angle_y;
} a3d_eval_params;

/* Derived constants of a caster (members c_res_x_, c_res_y_, c_n_sections_, c_split_distances_,
   c_split_widths_, c_field_of_view_{x,y}_ of the reference classes). */
typedef struct {
  int32_t c_res_x, c_res_y, c_n_sections, samples_per_full_ray;
  double fov_x, fov_y, ray_step;
  double split_distances[33];
  int32_t split_widths[33];
  int32_t _pad;
} a3d_caster_info;

/* ---- context ------------------------------------------------------------------------------ */
/* One context = one GPU (one process per GPU; multi-GPU runs shard segments across processes). */
int a3d_ctx_create(int device_id, a3d_ctx** out);
void a3d_ctx_destroy(a3d_ctx* ctx);
/* Text of the last error on this context (or of the last failed a3d_ctx_create if ctx == NULL). */
const char* a3d_last_error(const a3d_ctx* ctx);
/* The cudaStream_t all work of this context is issued on (for event timing by the caller). */
void* a3d_ctx_stream(a3d_ctx* ctx);
/* Device time in ms of the ray-cast/gain kernel of the most recent a3d_compute_gains*() call
   (CUDA events on the context stream); < 0 if none. */
float a3d_last_kernel_ms(a3d_ctx* ctx);
/* When that call ran the kernel chain (a3d_last_cast_kernel() == 3): device time of each of its kernels
   (last chunk of the batch), in launch order — IterativeRayCaster: views, mark, section n-2 marking
   parts, leaf parts, fix-up; SimpleRayCaster: views, rays, fix-up.  Returns how many were written
   (<= cap; 0: not available), < 0 on a bad argument. */
int a3d_last_stage_ms(a3d_ctx* ctx, float* out_ms, int cap);
/* Number of kernels this context has launched so far. */
uint64_t a3d_kernel_launches(a3d_ctx* ctx);
/* Tuning / test knobs.  "cast_kernel": 0 = automatic (default), 1 = always the scan-order kernel
   (k_cast_gain), 2 = require the wavefront kernel (k_cast_wave; errors if ordered digests are asked),
   3 = require the kernel chain (a3d_split.cu: gains, counts and ordered lists; errors if digests are
   asked or the dense meta volume does not exist).
   "wave_cta": CTA width of the wavefront kernel, 0 = automatic (8 or 4 warps for small and mid-size
   batches, else 2), 1 = always 2 warps, 2 = always 8 warps, 3 = always 4 warps. */
int a3d_set_option(a3d_ctx* ctx, const char* name, int64_t value);
/* Which kernel the last a3d_compute_gains*() used: 1 scan-order, 2 wavefront, 3 kernel chain (2 and 3:
   + scan-order re-evaluation of flagged segments). */
int a3d_last_cast_kernel(const a3d_ctx* ctx);

/* ---- map mirror: replaces VoxbloxMap (active_3d_planning_voxblox/src/map/voxblox.cpp) ------- */
/* voxblox::EsdfServer layer parameters (tsdf_voxel_size, tsdf_voxels_per_side,
   app_reconstruction/launch/example.launch:143-144); voxblox.cpp:26-27 caches the derived sizes. */
int a3d_map_config(a3d_ctx* ctx, float voxel_size, int voxels_per_side);
/* Insert-or-overwrite n blocks = the integrator's updated-block list for this planning step.
   block_idx n×3; esdf_dist / esdf_observed / tsdf_weight n×vps³ in voxblox linear voxel order
   x + vps*(y + vps*z); tsdf_weight may be NULL.  Replaces the esdf_map_in/tsdf_map_in layer
   merge of voxblox::EsdfServer (run_experiment.launch:180-181). */
int a3d_map_upload_blocks(a3d_ctx* ctx, int n, const int32_t* block_idx, const float* esdf_dist,
                          const uint8_t* esdf_observed, const float* tsdf_weight);
/* TSDF weights alone (what a `tsdf_map_in` layer message carries, run_experiment.launch:181) for blocks
   the mirror already holds: n blocks, vps^3 floats each, voxblox's linear voxel order.  Blocks
   without an ESDF block on the device are skipped; applied_out (nullable, n bytes) tells which ones
   were written.  Distances, observed flags and the meta words are untouched. */
int a3d_map_upload_weights(a3d_ctx* ctx, int n, const int32_t* block_idx, const float* tsdf_weight,
                           uint8_t* applied_out);
/* TSDF distances alone (the other member of the TSDF voxels a `tsdf_map_in` layer message carries), for
   VoxbloxMap::getVoxelDistance (active_3d_planning_voxblox/src/map/voxblox.cpp:84-98).  Same contract as
   a3d_map_upload_weights; the slab is allocated by the first call, blocks never sent read 0. */
int a3d_map_upload_tsdf_distances(a3d_ctx* ctx, int n, const int32_t* block_idx, const float* tsdf_distance,
                                  uint8_t* applied_out);
/* Same with payload pointers in device memory (e.g. the receive buffer of an NCCL broadcast);
   block_idx stays on the host. */
int a3d_map_upload_blocks_dev(a3d_ctx* ctx, int n, const int32_t* block_idx,
                              const float* esdf_dist_dev, const uint8_t* esdf_observed_dev,
                              const float* tsdf_weight_dev);
int a3d_map_remove_blocks(a3d_ctx* ctx, int n, const int32_t* block_idx);
int a3d_map_clear(a3d_ctx* ctx);
int64_t a3d_map_num_blocks(const a3d_ctx* ctx);
/* OccupancyMap::getVoxelSize (voxblox.cpp:63) */
double a3d_map_voxel_size(const a3d_ctx* ctx);
/* Batched map virtuals, n points (n×3 doubles):
   getVoxelState voxblox.cpp:48-60 | getVoxelCenter :66-81 | isObserved :43-45 |
   getVoxelWeight :101-115 | getVoxelDistance :84-98 | EsdfMap::getDistanceAtPosition as used at :35-36,:50 */
int a3d_map_voxel_state(a3d_ctx* ctx, int64_t n, const double* pts, uint8_t* out);
/* Same result as a3d_map_voxel_state, computed by the literal 8-lookup path that never reads the
   apron / meta acceleration structures (self-check of the device mirror). */
int a3d_map_voxel_state_literal(a3d_ctx* ctx, int64_t n, const double* pts, uint8_t* out);
int a3d_map_voxel_center(a3d_ctx* ctx, int64_t n, const double* pts, double* out);
int a3d_map_is_observed(a3d_ctx* ctx, int64_t n, const double* pts, uint8_t* out);
int a3d_map_voxel_weight(a3d_ctx* ctx, int64_t n, const double* pts, double* out);
/* VoxbloxMap::getVoxelDistance (voxblox.cpp:84-98): the stored TSDF distance of the voxel a point falls in
   (same voxel as getVoxelWeight); 0 without a block or before a3d_map_upload_tsdf_distances. */
int a3d_map_voxel_tsdf_distance(a3d_ctx* ctx, int64_t n, const double* pts, double* out);
int a3d_map_distance(a3d_ctx* ctx, int64_t n, const double* pts, double* dist, uint8_t* ok);
/* Batched VoxbloxMap::isTraversable (voxblox.cpp:32-41): observed and interpolated ESDF distance >
   collision_radius — the collision checks of TrajectoryGenerator::checkTraversable
   (core/src/module/trajectory_generator.cpp:30-46) and RRT::connectPoses (rrt.cpp:251-258). */
int a3d_map_traversable(a3d_ctx* ctx, int64_t n, const double* pts, double collision_radius,
                        uint8_t* out);

/* ---- sensor model / evaluator objects ------------------------------------------------------- */
/* Needs a configured map (ray_step default and c_res_* depend on the voxel size,
   simple_ray_caster.cpp:17-29). */
int a3d_caster_create(a3d_ctx* ctx, const a3d_caster_params* p, a3d_caster** out);
void a3d_caster_destroy(a3d_caster* c);
int a3d_caster_get_info(const a3d_caster* c, a3d_caster_info* out);
/* CameraModel::getDirectionVector for ray (i,j) as the casters call it (camera_model.cpp:161-168) */
int a3d_caster_direction(const a3d_caster* c, int i, int j, double* out3);
int a3d_eval_create(a3d_ctx* ctx, const a3d_eval_params* p, a3d_eval** out);
void a3d_eval_destroy(a3d_eval* e);

/* ---- the hot call ---------------------------------------------------------------------------
 * Batched SimulatedSensorEvaluator::computeGain (simulated_sensor_evaluator.cpp:45-82): for every
 * segment, ray-cast its sampled views (CameraModel::getVisibleVoxelsFromTrajectory,
 * camera_model.cpp:15-60), adjacent-unique the voxel-centre list, drop entries outside the
 * bounding volume, and fold the list into a gain (computeGainFromVisibleVoxels).
 *   pos n_views×3, quat n_views×4 (w,x,y,z): body poses trajectory[idx].position_W /
 *   orientation_W_B of the sampled trajectory points (the mounting transform is applied inside);
 *   view_segment n_views, non-decreasing, values in [0,n_segments): which segment a view belongs to.
 * Outputs, one per segment (each nullable except gains_out):
 *   gains_out      TrajectorySegment::gain
 *   n_visible_out  size of the voxel list the reference hands to storeTrajectoryInformation (after
 *                  unique and the bounding-volume filter; Naive and Frontier erase the observed entries
 *                  from the stored list afterwards, naive_evaluator.cpp:28-35 — that is not subtracted)
 *   n_samples_out  ray samples taken (= getVoxelState calls made by the caster)
 *   digests_out    order-sensitive 64-bit digest of that stored list (DESIGN.md "digest");
 *                  asking for it selects the scan-order kernel
 *   set_digests_out  order-independent digest of the same list (sum of per-entry hashes mod 2^64)
 *   seg_origin (input, nullable, n_segments×3): position of each segment's last trajectory point,
 *                  the `origin` of VoxelWeightEvaluator (voxel_weight_evaluator.cpp:53); default =
 *                  body position of the segment's last view
 * Synchronous on return.  An evaluator with clear_from_parents needs the tree: use
 * a3d_compute_gains_tree or a3d_compute_gains_stored (this call returns A3D_ERR_UNSUPPORTED). */
int a3d_compute_gains(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval, int n_views,
                      const double* pos, const double* quat, const int32_t* view_segment,
                      int n_segments, double* gains_out, uint64_t* n_visible_out,
                      uint64_t* n_samples_out, uint64_t* digests_out, uint64_t* set_digests_out,
                      const double* seg_origin);
/* Same with every array in device memory; enqueued on the context stream, NOT synchronised.
   view_first_dev: n_segments+1 int32 prefix offsets of each segment's views. */
int a3d_compute_gains_dev(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval,
                          int n_views, const double* pos_dev, const double* quat_dev,
                          const int32_t* view_first_dev, int n_segments, double* gains_dev,
                          uint64_t* n_visible_dev, uint64_t* n_samples_dev, uint64_t* digests_dev,
                          uint64_t* set_digests_dev, const double* seg_origin_dev);

/* One segment with its voxel list materialised (parity / visualisation / the per-segment
 * SensorModel::getVisibleVoxelsFromTrajectory + storeTrajectoryInformation API):
 *   eval == NULL: the sensor model's list (camera_model.cpp:15-60), no gain;
 *   eval != NULL: the list after the evaluator's bounding-volume filter and the gain of that list
 *                 (clear_from_parents is not applied here: the caller owns the ancestors' lists).
 * centres_out cap×3 doubles; *n_out = full list length even if > cap (then A3D_ERR_CAPACITY). */
int a3d_visible_voxels(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval, int n_views,
                       const double* pos, const double* quat, double* centres_out, int64_t cap,
                       int64_t* n_out, int64_t* n_samples_out, double* gain_out,
                       const double* origin /* 3 doubles, nullable: see seg_origin */);

/* computeGainFromVisibleVoxels on an explicit stored list (the SimulatedSensorUpdater path,
 * core/src/module/trajectory_evaluator/evaluator_updater/simulated_sensor_updater.cpp:18-22).
 * n_left_out (nullable): list size after the fold's own erasures (Naive/Frontier drop observed
 * voxels, naive_evaluator.cpp:28-35, frontier_evaluator.cpp:110-116); keep_out (nullable, n bytes):
 * which entries survive. */
int a3d_gain_from_voxels(a3d_ctx* ctx, const a3d_eval* eval, int64_t n, const double* centres,
                         double* gain_out, int64_t* n_left_out, uint8_t* keep_out,
                         const double* origin /* 3 doubles, VoxelWeightEvaluator only, else nullable */);

/* ---- device-resident stored lists -------------------------------------------
 * The reference keeps each segment's voxel list in TrajectorySegment::info (SimulatedSensorInfo,
 * simulated_sensor_evaluator.h:54-58) so that SimulatedSensorUpdater can re-fold the gain after the
 * map changed without casting rays again (evaluator_updater/simulated_sensor_updater.cpp:18-22), and
 * so that clear_from_parents can subtract what the ancestors already see
 * (simulated_sensor_evaluator.cpp:60-76).  A store keeps those lists in HBM as packed 64-bit voxel
 * keys (8 B per entry instead of the 24-B centre triple; two entries have equal centre doubles iff
 * their keys are equal), keyed by a caller-chosen non-zero 64-bit id (e.g. the segment's address). */
int a3d_store_create(a3d_ctx* ctx, a3d_store** out);
void a3d_store_destroy(a3d_store* store);
/* Batched SimulatedSensorEvaluator::computeGain that also keeps every segment's stored list under
 * seg_keys[s] (replacing older lists of the same key; keys must be distinct and non-zero).
 *   parent_keys (nullable, n_segments): store key of each segment's parent (TrajectorySegment::parent),
 *     0 for none.  Read only when the evaluator has clear_from_parents: entries that occur in the
 *     stored list of any ancestor are dropped (:60-76; a hash-set probe per ancestor instead of
 *     std::find).  A parent is either stored already or part of this batch (any order).
 *   The stored list is what the reference holds after computeGain returned: for Naive and Frontier
 *   evaluators the observed entries are erased (naive_evaluator.cpp:28-35, frontier_evaluator.cpp:110-116);
 *   n_visible_out is the size handed to storeTrajectoryInformation (before that erasure), digests /
 *   set digests are of that list.  Outputs other than gains_out are nullable.
 * Passes: gains / sizes / every ray part's place in its list (kernel chain) → ordered key lists into an
 * exactly sized arena (the chain's second pass writes each key straight to its place; list emission through
 * the chain's slots or the scan-order kernel where that does not apply) → where needed, k_list_pipeline
 * (parent filter, digests, fold + erasure), level by level of the tree. */
int a3d_compute_gains_stored(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval,
                             a3d_store* store, int n_views, const double* pos, const double* quat,
                             const int32_t* view_segment, int n_segments, const uint64_t* seg_keys,
                             const uint64_t* parent_keys, double* gains_out, uint64_t* n_visible_out,
                             uint64_t* n_samples_out, uint64_t* digests_out, uint64_t* set_digests_out,
                             const double* seg_origin);
/* a3d_compute_gains for a batch that carries its own tree: parent_index[s] (nullable) = index of
 * segment s's parent inside the batch (must be < s) or -1.  This is the batch form of computeGain
 * with clear_from_parents (simulated_sensor_evaluator.cpp:60-76); lists live in a temporary store. */
int a3d_compute_gains_tree(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval, int n_views,
                           const double* pos, const double* quat, const int32_t* view_segment,
                           int n_segments, const int32_t* parent_index, double* gains_out,
                           uint64_t* n_visible_out, uint64_t* n_samples_out, uint64_t* digests_out,
                           uint64_t* set_digests_out, const double* seg_origin);
/* computeGainFromVisibleVoxels for n stored lists against the CURRENT map, one launch per arena
 * (k_list_pipeline: the fold reads each entry's meta word straight from its key).  Naive and
 * Frontier evaluators also erase the observed entries from the stored lists, like the reference.
 * n_left_out (nullable): list sizes afterwards.  Unknown or repeated key → A3D_ERR_INVALID. */
int a3d_store_refold(a3d_ctx* ctx, const a3d_eval* eval, a3d_store* store, int n,
                     const uint64_t* seg_keys, double* gains_out, uint64_t* n_left_out,
                     const double* seg_origin);
int a3d_store_erase(a3d_store* store, int n, const uint64_t* seg_keys);
/* number of entries stored under key, -1 if absent */
int64_t a3d_store_list_size(const a3d_store* store, uint64_t key);
/* copy a stored list back as centre doubles (cap×3); *n_out = its length */
int a3d_store_get(a3d_ctx* ctx, const a3d_store* store, uint64_t key, double* centres_out, int64_t cap,
                  int64_t* n_out);
/* bytes of HBM currently held by the store (lists and hash sets) */
int64_t a3d_store_bytes(const a3d_store* store);

/* clear_from_parents step of SimulatedSensorEvaluator::computeGain
 * (simulated_sensor_evaluator.cpp:60-76) for ONE ancestor: keep_inout[i] is cleared if centres[i]
 * occurs anywhere in other[0..m) (exact comparison of the three doubles, like std::find). */
int a3d_filter_not_in(a3d_ctx* ctx, int64_t n, const double* centres, int64_t m,
                      const double* other, uint8_t* keep_inout);

/* BoundingVolume::contains (core/src/data/bounding_volume.cpp:33-57), batched. */
int a3d_bv_contains(a3d_ctx* ctx, const a3d_eval* eval, int64_t n, const double* pts, uint8_t* out);

/* ---- multi-GPU group (SURVEY.md §8(b),(e)) ----------------------------------------------------------
 * The N-GPU form of the calls above for ONE host thread of ONE process — what the reference's planner
 * (single-threaded, online_planner.cpp:414-416) would hold instead of a context.  A group owns one
 * a3d_ctx per device and one NCCL communicator per device (ncclCommInitAll; NCCL is dlopen-ed here, a
 * single-GPU group never loads it).  The map mirror is REPLICATED: an update goes host → GPU 0 once,
 * is ncclBroadcast over NVLink to the other GPUs and applied by every GPU to its own mirror.
 * Segments are SHARDED: dealt to the GPUs in blocks of `deal_block` segments (option, default 64; all
 * views of a segment stay on one GPU), every GPU casts its share, the per-segment records
 * (gain, n_visible, n_samples, digests) are ncclAllGather-ed and returned in the caller's order.
 * Same argument meaning, error codes and outputs as the single-context calls they mirror. */
int a3d_group_create(const int* device_ids, int n_dev, a3d_group** out);
void a3d_group_destroy(a3d_group* g);
/* text of the last error (g == NULL: of the last failed a3d_group_create) */
const char* a3d_group_last_error(const a3d_group* g);
int a3d_group_size(const a3d_group* g);
/* the i-th device's context, e.g. for point queries or event timing (owned by the group) */
a3d_ctx* a3d_group_ctx(a3d_group* g, int i);
int a3d_group_map_config(a3d_group* g, float voxel_size, int voxels_per_side);
int a3d_group_map_upload_blocks(a3d_group* g, int n, const int32_t* block_idx, const float* esdf_dist,
                                const uint8_t* esdf_observed, const float* tsdf_weight);
int a3d_group_map_remove_blocks(a3d_group* g, int n, const int32_t* block_idx);
int a3d_group_map_clear(a3d_group* g);
int a3d_group_caster_create(a3d_group* g, const a3d_caster_params* p, a3d_group_caster** out);
void a3d_group_caster_destroy(a3d_group_caster* c);
int a3d_group_caster_get_info(const a3d_group_caster* c, a3d_caster_info* out);
int a3d_group_eval_create(a3d_group* g, const a3d_eval_params* p, a3d_group_eval** out);
void a3d_group_eval_destroy(a3d_group_eval* e);
/* "deal_block" (see above) or any a3d_set_option name, applied to every device */
int a3d_group_set_option(a3d_group* g, const char* name, int64_t value);
/* a3d_compute_gains over all GPUs of the group; synchronous on return */
int a3d_group_compute_gains(a3d_group* g, const a3d_group_caster* caster, const a3d_group_eval* eval, int n_views,
                            const double* pos, const double* quat, const int32_t* view_segment, int n_segments,
                            double* gains_out, uint64_t* n_visible_out, uint64_t* n_samples_out,
                            uint64_t* digests_out, uint64_t* set_digests_out, const double* seg_origin);
/* max over the devices of a3d_last_kernel_ms / sum of a3d_kernel_launches */
float a3d_group_last_kernel_ms(a3d_group* g);
uint64_t a3d_group_kernel_launches(a3d_group* g);

#ifdef __cplusplus
}
#endif
#endif /* A3D_H_ */
