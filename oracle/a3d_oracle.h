/*
 * a3d_oracle.h — CPU oracle for the view-gain hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a literal single-purpose restatement of the reference algorithm
 * (ethz-asl/mav_active_3d_planning) used as the parity checker and as the timed CPU
 * baseline.  Nothing in the product (mav_active_3d_planning_b200/, include/) may call it;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PINNED by the reference's own code, except for the voxblox arithmetic: the reference ships no
 * tests or golden vectors, but its sources for this path compile against Eigen / glog / voxblox
 * stand-ins (oracle/Makefile target `_ref`, ref_driver.cpp, ref_shim/), and tests/test_ref_pin.py
 * demands that library == this one bit for bit; the .npz files under tests/golden hold its outputs for the GPU box.
 * The voxblox arithmetic (un-vendored dependency, unpinned in
 * mav_active_3d_planning_https.rosinstall:5) is restated from its published sources (SURVEY.md
 * Appendix A) in both libraries — VoxbloxLite::getDistanceAtPosition here, ref_shim/voxblox_shim.h
 * there — and is the part that stays "parity unpinned".
 */
#ifndef A3D_ORACLE_H_
#define A3D_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_map orc_map;
typedef struct orc_caster orc_caster;

enum { ORC_OCCUPIED = 0, ORC_FREE = 1, ORC_UNKNOWN = 2 }; /* occupancy_map.h:17-19 */
enum { ORC_CASTER_SIMPLE = 0, ORC_CASTER_ITERATIVE = 1, ORC_CASTER_ITERATIVE_LIDAR = 2 };
enum { ORC_EVAL_VOXEL_TYPE = 0, ORC_EVAL_NAIVE = 1, ORC_EVAL_FRONTIER = 2, ORC_EVAL_VOXEL_WEIGHT = 3 };

typedef struct {
  int32_t kind;                /* ORC_CASTER_* */
  int32_t resolution_x;        /* camera_model.cpp:174 */
  int32_t resolution_y;        /* camera_model.cpp:175 */
  int32_t _pad;
  double ray_length;           /* camera_model.cpp:171 */
  double focal_length;         /* camera_model.cpp:173 */
  double ray_step;             /* <=0 → map voxel size (simple_ray_caster.cpp:17) */
  double downsampling_factor;  /* simple_ray_caster.cpp:18 */
  double mount_t[3];           /* sensor_model.cpp:18-20 */
  double mount_q[4];           /* w,x,y,z, already normalised (sensor_model.cpp:21-27) */
  double field_of_view_x;      /* lidar only, degrees (lidar_model.cpp:170-171) */
  double field_of_view_y;
  double mount_q_raw[4];       /* w,x,y,z as configured, before sensor_model.cpp:27 normalises it
                                  (read by oracle/_ref only, which lets the reference normalise) */
} orc_caster_params;

typedef struct {
  int32_t kind;                 /* ORC_EVAL_* */
  int32_t clear_from_parents;   /* simulated_sensor_evaluator.cpp:17 */
  int32_t accurate_frontiers;   /* frontier_evaluator.cpp:17 */
  int32_t surface_frontiers;    /* frontier_evaluator.cpp:19 */
  double checking_distance;     /* frontier_evaluator.cpp:20 */
  /* voxel_type_evaluator.cpp:24-31: unknown, occupied, free, unknown_outer, occupied_outer,
     free_outer */
  double gains[6];
  int32_t bv_is_setup;          /* bounding_volume.cpp:27-31 */
  int32_t _pad;
  double bv_min[3];
  double bv_max[3];
  double bv_quat[4];            /* w,x,y,z of AngleAxis(-rotation*pi/180, Z) (bounding_volume.cpp:21) */
  /* voxel_weight_evaluator.cpp:17-22 */
  double frontier_voxel_weight, min_impact_factor, new_voxel_weight, ray_angle_x, ray_angle_y;
  double bv_rotation_deg;       /* bounding_volume.cpp:19 `rotation` as configured (read by oracle/_ref
                                   only; the restatement uses bv_quat) */
} orc_eval_params;

/* ---- map (VoxbloxLite: restated voxblox Layer<EsdfVoxel>/<TsdfVoxel> + Interpolator) ---- */
orc_map* orc_map_create(float voxel_size, int voxels_per_side);
void orc_map_destroy(orc_map* m);
/* insert-or-overwrite; voxel linear order x + vps*(y + vps*z); tsdf_weight nullable */
int orc_map_upload_blocks(orc_map* m, int n, const int32_t* block_idx, const float* esdf_dist,
                          const uint8_t* esdf_observed, const float* tsdf_weight);
int orc_map_remove_blocks(orc_map* m, int n, const int32_t* block_idx);
int64_t orc_map_num_blocks(const orc_map* m);
double orc_map_voxel_size(const orc_map* m);

/* vbx/src/map/voxblox.cpp:48-60, 66-81, 43-45, 101-115 — batched over n points (n×3) */
void orc_voxel_state(const orc_map* m, int64_t n, const double* pts, uint8_t* out);
void orc_voxel_center(const orc_map* m, int64_t n, const double* pts, double* out);
void orc_is_observed(const orc_map* m, int64_t n, const double* pts, uint8_t* out);
void orc_voxel_weight(const orc_map* m, int64_t n, const double* pts, double* out);
/* interpolated distance + success flag (EsdfMap::getDistanceAtPosition) */
void orc_distance(const orc_map* m, int64_t n, const double* pts, double* dist, uint8_t* ok);

/* VoxbloxMap::isTraversable (vbx/src/map/voxblox.cpp:32-41) for a given collision radius */
void orc_traversable(const orc_map* m, int64_t n, const double* pts, double collision_radius, uint8_t* out);
/* "port" for liba3d_oracle.so (the restatement), "reference" for _ref/liba3d_ref.so */
const char* orc_impl(void);

/* ---- caster (camera_model.cpp, simple_ray_caster.cpp, iterative_ray_caster.cpp) ---- */
orc_caster* orc_caster_create(const orc_map* m, const orc_caster_params* p);
void orc_caster_destroy(orc_caster* c);
/* info: [c_res_x, c_res_y, n_sections]; split_distances/widths need room for 33 entries */
void orc_caster_info(const orc_caster* c, int32_t* info, double* split_distances,
                     int32_t* split_widths, double* fov_xy, double* ray_step);
/* camera_model.cpp:161-168 direction for ray (i,j) as the casters call it */
void orc_caster_direction(const orc_caster* c, int i, int j, double* out3);
/* camera_model.cpp:62-83 */
int orc_sample_viewpoints(double sampling_time, int n_points, const int64_t* time_from_start_ns,
                          int32_t* out_indices);
/* CameraModel::getVisibleVoxelsFromTrajectory for one segment whose sampled body poses are given
   (pos n×3, quat n×4 wxyz).  Returns the number of entries; writes min(n, cap) centres.
   n_samples_out (nullable) = ray samples taken (= getVoxelState calls in the caster). */
int64_t orc_visible_voxels(const orc_caster* c, int n_views, const double* pos, const double* quat,
                           double* centres_out, int64_t cap, int64_t* n_samples_out);

/* ---- SimulatedSensorEvaluator::computeGain over a batch of segments ----
 * view_segment (n_views, non-decreasing) maps views to segments.  parent (nullable, n_segments)
 * gives the index of each segment's parent inside this batch (-1 = root); parents must precede
 * children (only read when clear_from_parents).  Outputs per segment: gain; n_visible = size of
 * the list stored in the segment (after unique + bounding volume + parents, before the fold
 * mutates it); n_samples; digest of that stored list (nullable); n_fold_lookups = voxel-record
 * fetches of the fold (nullable).  n_threads<=1 → sequential like the reference. */
int orc_compute_gains(const orc_caster* c, const orc_eval_params* e, int n_views, const double* pos,
                      const double* quat, const int32_t* view_segment, int n_segments,
                      const int32_t* parent, double* gains_out, uint64_t* n_visible_out,
                      uint64_t* n_samples_out, uint64_t* digests_out,
                      uint64_t* n_fold_lookups_out, int n_threads, uint64_t* set_digests_out,
                      const double* seg_origin /* n_segments x 3, nullable: position of the segment's last
                                                  trajectory point (voxel_weight_evaluator.cpp:53);
                                                  default = body position of its last view */);
/* computeGainFromVisibleVoxels on an explicit list (n×3 centres) — SimulatedSensorUpdater path.
   Returns gain; n_left (nullable) = list size after the fold's own erasures. */
double orc_gain_from_voxels(const orc_map* m, const orc_eval_params* e, int64_t n,
                            const double* centres, int64_t* n_left, const double* origin /*3, nullable*/);
/* BoundingVolume::contains (bounding_volume.cpp:33-57) */
void orc_bv_contains(const orc_eval_params* e, int64_t n, const double* pts, uint8_t* out);
/* order-sensitive digest of a centre list (shared definition with the product) */
uint64_t orc_digest(int64_t n, const double* centres);
/* order-independent companion: sum of the per-entry hashes mod 2^64 */
uint64_t orc_digest_set(int64_t n, const double* centres);

#ifdef __cplusplus
}
#endif
#endif /* A3D_ORACLE_H_ */
