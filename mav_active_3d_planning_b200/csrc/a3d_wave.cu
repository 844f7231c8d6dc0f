// a3d_wave.cu — k_cast_wave: the throughput kernel of the view-gain path (DESIGN.md "kernel K3w").
//
// Same results as k_cast_gain (a3d_kernels.cu) for gain / n_visible / n_samples and an
// order-independent digest of the stored list, but mapped lane-per-ray instead of warp-per-ray:
//
//  * One CTA of 2 warps owns a trajectory segment at a time (atomic work counter).
//  * Every lane marches its OWN ray part: the previous raw list entry it needs for the
//    adjacent-unique test (camera_model.cpp:58) is in its own registers, the sample distance is
//    accumulated in a register exactly like the reference does (`distance += p_ray_step_`), and no
//    shuffle is needed per sample.  Voxel identity is the integer triple
//    g = block*vps + voxel (+ artefact flags), equal iff the reference's double centre triples are
//    equal (fp artefact indices are canonicalised or flagged, see make_ident_slow).  Lanes are
//    persistent: when a ray ends (range or OCCUPIED hit) the lane pulls the next ready ray, so
//    occlusion does not idle lanes.
//  * A sample whose voxel coordinate floor(pf / voxel_size) is provably what voxblox's own index
//    arithmetic yields ("certified", see sample_prepare) costs ONE load of the voxel's meta word
//    from the dense meta volume: cell class, centre state, observed flag and frontier bits all
//    come out of that word.  Everything else is evaluated literally.
//  * What couples rays — `std::unique` across the boundary between two consecutive rays / views —
//    is repaired afterwards: each ray part records its first and last key and what its first entry
//    contributed; a scan-order pass (fix-up) un-counts first entries that equal the previous
//    emitted entry.
//  * IterativeRayCaster (iterative_ray_caster.cpp:48-119): the ray table makes ray (i,j) depend on
//    rays (i'<=i, j'<=j) only, so rays on one anti-diagonal are independent.  Warp 0 walks the
//    anti-diagonals: it casts the rays that start before the last section only up to
//    c_split_distances_[n-1] — the part that decides every markNeighboringRays call ("marking
//    parts") — and applies their marks: an entry keeps the value of its highest writer (= the
//    last writer in scan order), so the marks commute and are applied with atomicMax.  Every ray whose table entry then reads n-1 (fresh last-section starts AND
//    survivors of the marking parts, whose own cell gets marked n-1) has its last section cast as a
//    "leaf part".  Leaf parts come from a shared supply (one unit = up to 32 rays of a finished
//    anti-diagonal); warp 1 only casts leaf parts, warp 0 fills the lanes its marking parts leave
//    idle with leaf parts too.
//  * SimpleRayCaster (simple_ray_caster.cpp:32-66): all rays independent; both warps cast leaf
//    parts (whole rays).
#include <algorithm>

#include "../../include/a3d.h"
#include "a3d_kernels.cuh"
#include "a3d_keys.cuh"
#include "a3d_sample.cuh"

// The kernel has ~40 instantiations; to halve the build time this file is compiled as two translation
// units: as itself (part 1: IterativeRayCaster instantiations + the public entry points) and through
// a3d_wave_simple.cu (part 2: SimpleRayCaster instantiations only).
#ifndef A3D_WAVE_PART
#define A3D_WAVE_PART 1
#endif

namespace a3d {

void launch_wave_iterative(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                           bool digest, int grid, const void* ws, unsigned int* inexact, int warps);
void launch_wave_simple(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                        bool digest, int grid, const void* ws, unsigned int* inexact, int warps);

namespace {

#ifndef A3D_WAVE_WARPS
#define A3D_WAVE_WARPS 2  // warps per CTA: warp 0 schedules (iterative caster), all cast leaf parts
#endif
constexpr int kWaveWarps = A3D_WAVE_WARPS;
constexpr int kWaveThreads = 32 * kWaveWarps;
constexpr int kWaveWideWarps = 8;  // CTA width for small batches (2 such CTAs per SM)
constexpr int kWaveMidWarps = 4;   // ... for mid-size batches (5 per SM)
#ifndef A3D_WAVE_MIN_BLOCKS
#define A3D_WAVE_MIN_BLOCKS 10  // resident 2-warp CTAs per SM the register allocation must allow
#endif
constexpr int kPendCap = 96;    // per-warp ring of ready rays (top-up adds <= 32 while count <= 64)
constexpr int kMarkCap = 1024;  // max rays of one anti-diagonal (min(c_res_x, c_res_y))

// info byte of a ray-part record: bits 2:0 category counted for the first entry (kCatNone: not
// counted), bit 3: part emitted at least one entry, bit 4: its first entry was kept (counted in
// n_visible)
struct Records {
  unsigned long long* first_key;   // [2][n_rays]
  unsigned long long* last_key;    // [2][n_rays]
  unsigned long long* first_hash;  // [2][n_rays] (digest mode only)
  uint8_t* info;                   // [2][n_rays]
};

struct WaveScratch {
  uint32_t* table;     // per CTA: n_rays entries (writer+1) << 8 | uint8(value)
  uint8_t* start_s;    // per CTA: original start section of rays cast as marking parts, 0xFF none
  Records rec;         // per CTA
  size_t rays_stride;  // n_rays rounded up (elements) per CTA
  // list emission (null / 0 when only gains are asked): the kept entries of every ray part as packed
  // keys, kmax slots per part; how many; where the part's entries go in the segment's ordered list
  unsigned long long* ent;  // [2][n_rays][kmax]
  uint16_t* cnt;            // [2][n_rays]
  uint32_t* off;            // [2][n_rays]: offset within the segment's list | (first entry dropped) << 31
  int kmax;
};

// Per-lane accumulators of a segment.  `packed`: three 21-bit fields, field c counts list entries of
// fold category c (VoxelType: voxel state 0/1/2; Naive/Frontier: field 0 = entries that score).
// field 0..2: VoxelType → entries of voxel state 0/1/2; Naive/Frontier → field 0 = entries that score,
// field 1 = other retained entries.  n_visible = sum of the fields.
struct LaneCounters {
  unsigned long long packed;
  unsigned long long digest;  // DIGEST instantiations only
  double wsum;                // VoxelWeightEvaluator instantiations only: partial gain
  unsigned n_samp;
};

// Per-lane state of the ray part a lane is marching (kept small: it lives in registers across the
// whole engine loop).
// Index part of a sample: voxel coordinate, cell octant and the meta word (requested).
struct Prepared {
  int g0, g1, g2;  // floor(pf / voxel_size) per axis
  int flags;       // bits 2:0 cell octant, bit 3: certified (see ray_step)
  uint32_t word;   // meta word of voxel g (certified samples only)
};

// Per-lane state that is read once per sample or less lives in shared memory, one column per
// thread: the kernel is register-bound (measured: every spilled value costs more than these loads).
template <int THREADS>
struct LaneTable {
  double dir[3][THREADS];  // world direction of the lane's ray
  double d_end[THREADS];   // the part ends before this distance
  int r[THREADS];          // ray id
  int mark[THREADS];       // index of the lane's marking part in the diagonal's list
  int kept[THREADS];       // list emission: entries of the lane's part stored so far
  unsigned inexact;             // any lane met a case the scan-order kernel must decide
};

struct RayState {
  double d;           // distance of the next sample (accumulated like `distance += p_ray_step_`)
  Ident prev;         // previous raw entry (g0 == kNoIdent: none yet)
};

struct ViewShared {  // per-view constants, one copy per CTA in shared memory
  double P0[3];
  double q[4];
  // index-certification margin of this view (NaN: no sample of the view may take the fast path)
  float delta, delta_hi;  // delta_hi = 0.5f - delta
  double org[3];          // VoxelWeightEvaluator: the segment's reference point (last trajectory point)
};

template <class LT>
__device__ __forceinline__ void ray_begin(RayState& rs, LT& lt, const CasterView& c,
                                          const ViewShared& vw, int r, double d0) {
  const Q4 q{vw.q[0], vw.q[1], vw.q[2], vw.q[3]};
  const D3 dir = rotate(q, D3{c.cam_dirs[3 * r], c.cam_dirs[3 * r + 1], c.cam_dirs[3 * r + 2]});
  lt.dir[0][threadIdx.x] = dir.x;
  lt.dir[1][threadIdx.x] = dir.y;
  lt.dir[2][threadIdx.x] = dir.z;
  rs.d = d0;
  rs.prev = Ident{kNoIdent, 0, 0};
  lt.r[threadIdx.x] = r;
  lt.kept[threadIdx.x] = 0;
}

// Position of the sample at distance d of the lane's ray, double and float (the reference casts the
// double point to voxblox::Point).
template <class LT>
__device__ __forceinline__ void sample_point(const ViewShared& vw, const LT& lt, double d,
                                             double (&pp)[3], float (&pf)[3]) {
  const double dx = lt.dir[0][threadIdx.x], dy = lt.dir[1][threadIdx.x], dz = lt.dir[2][threadIdx.x];
  pp[0] = __dadd_rn(vw.P0[0], __dmul_rn(d, dx));
  pp[1] = __dadd_rn(vw.P0[1], __dmul_rn(d, dy));
  pp[2] = __dadd_rn(vw.P0[2], __dmul_rn(d, dz));
#pragma unroll
  for (int a = 0; a < 3; ++a) pf[a] = __double2float_rn(pp[a]);
}

// Index arithmetic of the sample at distance d and the request for its meta word.
//
// With T = pf * (1/voxel_size) per axis, g = floor(T) is the global voxel coordinate
// block*vps + voxel and frac(T) tells the half of the voxel (= interpolation cell octant).  The
// reference reaches the same numbers through ~8 float roundings (block index, block origin,
// in-block offset, voxel index, each with voxblox's +1e-6; once more from the double point for
// getVoxelCenter), whose combined error in voxel units is below
//     delta = 5e-7 * max|T| + (vps+1) * (16 * 2^-24 + 1.1e-6).
// When frac(T) keeps a distance > delta from 0, 1/2 and 1 on all three axes, every one of those
// floors and the centre comparison provably land where floor(T) / frac(T) say: block and voxel
// indices are in range, getVoxelState's and getVoxelCenter's voxels coincide and no artefact index
// can occur ("certified").  Such a sample costs one load from the dense meta volume (or a directory
// load and a meta-word load when the volume is not available).  Everything else (≈ 12*delta of the
// samples, plus cells whose class is MIXED) is evaluated literally by ray_step().
template <class LT>
__device__ __forceinline__ Prepared sample_prepare(const MapView& m, const ViewShared& vw, const LT& lt,
                                                   double d) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: (T + kMagic) - kMagic = rint(T) for |T| < 2^22
  double pp[3];
  float pf[3];
  sample_point(vw, lt, d, pp, pf);
  int g[3];
  int sc = 0;
  bool safe = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float T = __fmul_rn(pf[a], m.vs_inv_f);
    const float s = __fadd_rn(T, kMagic);
    const float fr = __fsub_rn(T, __fsub_rn(s, kMagic));  // T - rint(T), exact, in [-1/2, 1/2]
    const float af = fabsf(fr);
    safe &= (af > vw.delta) & (af < vw.delta_hi);
    const int below = __float_as_int(fr) >> 31;  // -1 when T is below its nearest integer
    g[a] = (__float_as_int(s) - 0x4B400000) + below;
    sc |= (below + 1) << a;  // bit set: the point is below the voxel centre on this axis
  }
  if (!safe) {
    // Margin not met: the reference's own index arithmetic, inline.  If it stays in range and
    // getVoxelState's voxel equals getVoxelCenter's, the sample is an ordinary one after all.
    int sc_l = 0;
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int bi = grid_index(pf[a], m.bs_inv_f);
      const float origin = origin_point(bi, m.bs_f);
      const int vs = grid_index(__fsub_rn(pf[a], origin), m.vs_inv_f);
      const int vi = grid_index(__double2float_rn(__dsub_rn(pp[a], static_cast<double>(origin))), m.vs_inv_f);
      ok &= (static_cast<unsigned>(vs) < static_cast<unsigned>(m.vps)) & (vs == vi);
      const float centre = __fadd_rn(origin, center_point(min(max(vs, 0), m.vps - 1), m.vs_f));
      sc_l |= (__fsub_rn(pf[a], centre) < 0.0f) ? (1 << a) : 0;
      g[a] = bi * m.vps + vs;
    }
    sc = sc_l;
    safe = ok && m.vps_shift >= 0 && vw.delta == vw.delta;  // (NaN margin: certification is off)
  }
  Prepared nx;
  nx.g0 = g[0];
  nx.g1 = g[1];
  nx.g2 = g[2];
  nx.flags = sc | (safe ? 8 : 0);
  nx.word = kWordNoBlock;
  if (safe) {
    if (m.vol) {
      // dense meta volume: one load, no block lookup
      const unsigned u0 = static_cast<unsigned>(g[0] - m.vol_lo[0]),
                     u1 = static_cast<unsigned>(g[1] - m.vol_lo[1]),
                     u2 = static_cast<unsigned>(g[2] - m.vol_lo[2]);
      if ((u0 < m.vol_n[0]) & (u1 < m.vol_n[1]) & (u2 < m.vol_n[2]))
        nx.word = __ldg(m.vol + vol_offset(m, u0, u1, u2));
    } else {
      const int vmask = m.vps - 1;
      const int slot = find_slot(m, g[0] >> m.vps_shift, g[1] >> m.vps_shift, g[2] >> m.vps_shift);
      if (slot >= 0)
        nx.word = __ldg(m.mword + static_cast<size_t>(slot) * m.nv +
                        mword_offset(m, g[0] & vmask, g[1] & vmask, g[2] & vmask));
    }
  }
  return nx;
}

// One sample of the lane's ray (distance rs.d): index arithmetic + meta word, then the fold.
// Returns true when the part is finished (hit or end reached) and then leaves in rs.d the distance
// of the OCCUPIED sample that ended it, or -1.
// rec_base + (marking part ? 0 : n_rays) + ray: index of this part's record.  Precondition:
// rs.d < lt.d_end.
template <int CASTER, int EVAL, bool DIGEST, bool EMIT, class LT>
__device__ __forceinline__ bool ray_step(const MapView& m, const CenterTables& ct, const EvalView& e,
                                         const ViewShared& vw, const BvGrid& bvg, const int bv_mode,
                                         const double step, RayState& rs, LT& lt, LaneCounters& lc,
                                         const Records& rec, size_t rec_base, int role, int n_rays,
                                         unsigned long long* ent, int kmax) {
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  const double d_this = rs.d;
  const Prepared cur = sample_prepare(m, vw, lt, d_this);
  rs.d = __dadd_rn(d_this, step);
  const bool more = rs.d < lt.d_end[threadIdx.x];
  lc.n_samp++;
  const int g[3] = {cur.g0, cur.g1, cur.g2};
  const int sc = cur.flags & 7;
  const bool safe = (cur.flags & 8) != 0;
  SampleInfo si;
  D3 cc{0.0, 0.0, 0.0};  // centre of the entry's voxel, formed only where something needs it
  bool have_cc = false;
  bool in_bv = true;
  auto centre = [&]() {
    if (have_cc) return;
    const int vmask = m.vps - 1;  // (certified samples only: the literal path delivers its centre)
    cc.x = __dadd_rn(static_cast<double>(origin_point(g[0] >> m.vps_shift, m.bs_f)), ct.d[(g[0] & vmask) + 1]);
    cc.y = __dadd_rn(static_cast<double>(origin_point(g[1] >> m.vps_shift, m.bs_f)), ct.d[(g[1] & vmask) + 1]);
    cc.z = __dadd_rn(static_cast<double>(origin_point(g[2] >> m.vps_shift, m.bs_f)), ct.d[(g[2] & vmask) + 1]);
    have_cc = true;
  };
  if (safe) {
    const uint32_t word = cur.word;
    const int cls = (word >> (2 * sc)) & 3;
    si.st = cls == kCellUnknown ? 2 : (cls == kCellFree ? 1 : 0);
    si.cs = centre_state_from_word(word);
    si.id = Ident{g[0], g[1], g[2]};
    if (cls == kCellMixed) {
      double pp[3];
      float pf[3];
      sample_point(vw, lt, d_this, pp, pf);
      si.st = mixed_state(m, ct, pf[0], pf[1], pf[2], g[0], g[1], g[2], sc);
    }
    if (bv_mode == 2) {
      centre();
      in_bv = bv_contains(e, cc);  // simulated_sensor_evaluator.cpp:50-57
    } else if (bv_mode) {
      in_bv = (g[0] >= bvg.lo[0]) & (g[0] <= bvg.hi[0]) & (g[1] >= bvg.lo[1]) & (g[1] <= bvg.hi[1]) &
              (g[2] >= bvg.lo[2]) & (g[2] <= bvg.hi[2]);
    }
  } else {
    ExactOut out;
    unsigned bad = 0;
    double pp[3];
    float pf[3];
    sample_point(vw, lt, d_this, pp, pf);
    sample_exact(m, ct, pp[0], pp[1], pp[2], &out, &bad);
    if (bad) lt.inexact = 1;
    si = out.si;
    cc = out.cc;
    have_cc = true;
    if (bv_mode) in_bv = bv_contains(e, cc);
  }
  const bool emit = ITER || si.st != 0;
  const bool keep = emit & in_bv &
                    ((si.id.g0 != rs.prev.g0) | (si.id.g1 != rs.prev.g1) | (si.id.g2 != rs.prev.g2));
  const bool first = emit & (rs.prev.g0 == kNoIdent);
  if (emit) rs.prev = si.id;
  int cat;
  double wval = 0.0;
  if (EVAL == A3D_EVAL_VOXEL_TYPE) {
    cat = si.cs & 3;  // retained entries are always inside the bounding volume
  } else if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
    cat = 1;
    if (keep && (si.cs & 3) != 1) {  // voxel_weight_evaluator.cpp:43-83 (free voxels score nothing)
      int hint = -1;
      if (safe && (si.cs & 3) == 2 && e.frontier_voxel_weight > 0.0) hint = is_frontier_from_word(m, e, cur.word);
      double w = 0.0;
      if ((si.cs & 3) == 0 || (hint < 0 && e.frontier_voxel_weight > 0.0)) centre();
      if ((si.cs & 3) == 0) w = voxel_weight(m, cc);
      if (safe && hint < 0 && (si.cs & 3) == 2 && e.frontier_voxel_weight > 0.0)
        hint = is_frontier_by_meta(m, ct, e, cc, g[0], g[1], g[2]);
      wval = voxel_weight_value(m, ct, e, cc, si.cs, w, D3{vw.org[0], vw.org[1], vw.org[2]}, hint);
      lc.wsum += wval;
    }
  } else {
    cat = 1;
    if (keep && !(si.cs & 4)) {
      if (EVAL == A3D_EVAL_NAIVE) {
        cat = 0;
      } else {
        int hint = safe ? is_frontier_from_word(m, e, cur.word) : -1;
        if (hint < 0) {
          centre();
          if (safe) hint = is_frontier_by_meta(m, ct, e, cc, g[0], g[1], g[2]);
        }
        if (hint >= 0 ? hint != 0 : is_frontier_voxel(m, ct, e, cc)) cat = 0;
      }
    }
  }
  lc.packed += bvg.inc[keep ? cat : 3];
  unsigned long long h = 0;
  if (DIGEST) {
    if (keep) {
      centre();
      h = digest_entry(cc);
    }
    lc.digest += h;
  }
  if (EMIT && keep) {
    // list emission: the kept entry, in the part's own slot range (ordered within the part)
    unsigned bad = 0;
    const unsigned long long key = pack_key(si.id, &bad);
    if (bad) lt.inexact = 1;
    const int k = lt.kept[threadIdx.x];
    A3D_DEV_CHECK(lt.r[threadIdx.x] >= 0 && lt.r[threadIdx.x] < n_rays && k < kmax);
    if (k < kmax)
      __stcs(&ent[(rec_base + (role == 2 ? 0 : n_rays) + lt.r[threadIdx.x]) * static_cast<size_t>(kmax) + k], key);
    lt.kept[threadIdx.x] = k + 1;
  }
  if (first) {
    // record of the part's first entry, written once
    unsigned bad = 0;
    const size_t rec_at = rec_base + (role == 2 ? 0 : n_rays) + lt.r[threadIdx.x];
    A3D_DEV_CHECK(lt.r[threadIdx.x] >= 0 && lt.r[threadIdx.x] < n_rays);
    __stcs(&rec.first_key[rec_at], pack_key(si.id, &bad));
    if (bad) lt.inexact = 1;
    __stcs(&rec.info[rec_at], static_cast<uint8_t>((keep ? cat : kCatNone) | 8 | (keep ? 16 : 0)));
    if (DIGEST) rec.first_hash[rec_at] = h;
    if (EVAL == A3D_EVAL_VOXEL_WEIGHT) rec.first_hash[rec_at] = __double_as_longlong(wval);
  }
  if (si.st == 0) {
    rs.d = d_this;
    return true;
  }
  if (!more) rs.d = -1.0;
  return !more;
}

#ifndef A3D_WAVE_STEPS
#define A3D_WAVE_STEPS 64  // most samples marched between two rounds of lane bookkeeping
#endif
#ifndef A3D_WAVE_DRY_STEPS
#define A3D_WAVE_DRY_STEPS 8  // round length while idle lanes cannot be refilled
#endif
#ifndef A3D_WAVE_SCHED_LEAF_MAX
#define A3D_WAVE_SCHED_LEAF_MAX 32  // leaf parts the scheduler warp may cast while anti-diagonals are open
#endif
#ifndef A3D_WAVE_IDLE_BREAK
#define A3D_WAVE_IDLE_BREAK 12  // idle lanes that end a round early (32: never)
#endif
#ifndef A3D_WAVE_MARK_STEPS
#define A3D_WAVE_MARK_STEPS 64  // ... while marking parts are in flight (scheduler warp)
#endif

}  // namespace

// WARPS: 2 for throughput (as many CTAs as the registers allow), 4 and 8 for mid-size and small
// batches (5 / 2 such CTAs fit an SM), where a view's leaf parts are better spread over more lanes of the one CTA that owns it.
// EMIT: also write the ordered visible-voxel lists (8-warp CTAs only).
template <int CASTER, int EVAL, bool DIGEST, int WARPS, bool EMIT>
__global__ void __launch_bounds__(32 * WARPS, WARPS == kWaveWarps      ? A3D_WAVE_MIN_BLOCKS
                                             : WARPS == kWaveMidWarps ? 5
                                                                      : 2)
    k_cast_wave(const MapView m, const CasterView c, const EvalView e, const Batch b,
                const WaveScratch ws, unsigned int* __restrict__ inexact_flags) {
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int n_rays = c.n_rays;
  const int n = c.n_sections;
  const int n_diag = c.res_x + c.res_y - 1;
  const int n_units = c.n_units;

  __shared__ CenterTables ct;
  __shared__ ViewShared vw;
  __shared__ int s_seg;
  __shared__ volatile int s_done_diag;         // anti-diagonals whose table entries are final
  __shared__ unsigned s_unit_cursor;           // next leaf-supply unit
  constexpr int kThreads = 32 * WARPS;
  __shared__ unsigned s_pending[WARPS][kPendCap];  // per-warp ring of ready rays waiting for a lane
  // marking rays of the current diagonal (dynamic shared memory, sized by the longest diagonal):
  // their hit distance (< 0: none) and r << 8 | s
  extern __shared__ double s_dyn[];
  double* const s_mark_hit = s_dyn;
  unsigned* const s_mark_list = reinterpret_cast<unsigned*>(s_dyn + c.diag_max);
  __shared__ unsigned long long s_packed, s_digest;
  __shared__ double s_wsum;
  __shared__ unsigned s_nsamp;
  __shared__ BvGrid bvg;
  __shared__ LaneTable<kThreads> lt;
  init_center_tables(m, &ct);
  __syncthreads();
  if (threadIdx.x < 32) init_bv_grid(m, ct, e, &bvg);
  __syncthreads();
  const int bv_mode = bvg.mode;

  const size_t cta_base = static_cast<size_t>(blockIdx.x) * ws.rays_stride;
  uint32_t* const table = ws.table + cta_base;  // written by atomics, read with __ldcg
  volatile uint8_t* const start_s = ws.start_s + cta_base;
  const double step = c.ray_step;
  const double d_leaf_end = c.ray_length;
  const double d_mark_end = (ITER && n > 1) ? c.split[n - 1] : 0.0;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      s_seg = static_cast<int>(atomicAdd(b.work_counter, 1u));
      lt.inexact = 0;
    }
    __syncthreads();
    const int seg = s_seg;
    if (seg >= b.n_segments) break;
    const int v_begin = b.view_first[seg], v_end = b.view_first[seg + 1];

    LaneCounters lc{};
    uint64_t carry_key = kNoKey;  // last emitted key so far in this segment (warp 0, uniform)
    unsigned emit_pos = 0;        // list emission: entries of the segment's list placed so far (warp 0)

    for (int view = v_begin; view < v_end; ++view) {
      // per-view init (after the previous view's fix-up): pose, ray table, start sections, records
      __syncthreads();
      if (threadIdx.x == 0) {
        // camera pose = body pose ∘ mounting transform (camera_model.cpp:24-28)
        D3 P0{b.pos[3 * view], b.pos[3 * view + 1], b.pos[3 * view + 2]};
        Q4 q{b.quat[4 * view], b.quat[4 * view + 1], b.quat[4 * view + 2], b.quat[4 * view + 3]};
        const D3 mt = rotate(q, D3{c.mount_t[0], c.mount_t[1], c.mount_t[2]});
        vw.P0[0] = __dadd_rn(P0.x, mt.x);
        vw.P0[1] = __dadd_rn(P0.y, mt.y);
        vw.P0[2] = __dadd_rn(P0.z, mt.z);
        q = qmul(q, c.mount_q);
        vw.q[0] = q.w;
        vw.q[1] = q.x;
        vw.q[2] = q.y;
        vw.q[3] = q.z;
        // margin of the certified index path for every sample this view can produce
        const double reach = fmax(fmax(fabs(vw.P0[0]), fabs(vw.P0[1])), fabs(vw.P0[2])) + c.ray_length;
        const double t_max = reach * static_cast<double>(m.vs_inv_f) * 1.001 + 2.0;
        float delta = __int_as_float(0x7FC00000);
        if (m.vps_shift >= 0 && t_max < 1048576.0) delta = static_cast<float>(t_max * 5e-7) + m.fast_c0;
        vw.delta = delta;
        vw.delta_hi = 0.5f - delta;
        if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
          const double* o = b.seg_origin ? b.seg_origin + 3 * seg : b.pos + 3 * (v_end - 1);
          vw.org[0] = o[0];
          vw.org[1] = o[1];
          vw.org[2] = o[2];
        }
        s_done_diag = (ITER && n > 1) ? 0 : n_diag;
        s_unit_cursor = 0;
      }
      for (int r = threadIdx.x; r < n_rays; r += kThreads) {
        if (ITER) {
          table[r] = 0u;
          start_s[r] = 0xFF;
        }
        __stcs(&ws.rec.info[cta_base * 2 + r], static_cast<uint8_t>(kCatNone));
        __stcs(&ws.rec.info[cta_base * 2 + n_rays + r], static_cast<uint8_t>(kCatNone));
      }
      __syncthreads();

      // ---------------- persistent-lane engine ----------------------------------------------------
      {
        unsigned* const pending = s_pending[warp];
        int pend_head = 0, pend_count = 0;  // ring state (warp-uniform)
        bool supply_done = false;
        // scheduler state (warp 0 of the iterative caster)
        const bool scheduler = ITER && warp == 0 && n > 1;
        bool sched_done = !scheduler;
        int cur_diag = -1, n_mark = 0, mark_head = 0, n_run = 0;
        // lane state
        RayState rs;
        int role = 0;  // 0 idle, 1 leaf part, 2 marking part

        for (;;) {
          // ---- 1. scheduler: finish the current diagonal, open the next ones ----
          while (!sched_done && mark_head == n_mark && n_run == 0) {
            if (cur_diag >= 0) {
              // markNeighboringRays of every marking ray of the diagonal.  An entry keeps the value
              // of its highest writer (= last writer in scan order), so the marks commute: they are
              // applied with fire-and-forget atomicMax on (writer << 8 | value), one ray per lane.
              for (int base = 0; base < n_mark; base += 32) {
                const int qi = base + static_cast<int>(lane);
                const bool valid = qi < n_mark;
                const unsigned it = valid ? s_mark_list[qi] : 0u;
                const int rr = static_cast<int>(it >> 8), ss = static_cast<int>(it & 0xFF);
                const double hd = valid ? s_mark_hit[qi] : -1.0;
                const bool hit = hd >= 0.0;
                int seg_h = ss;
                if (hit) {
                  while (!(hd < c.split[seg_h + 1])) ++seg_h;
                }
                const int t_last = hit ? seg_h : n - 2;
                const int w0 = 1 << (n - 1 - ss);
                const int ri = rr / c.res_y, rj = rr - ri * c.res_y;
                const unsigned wr = (static_cast<unsigned>(rr) + 1u) << 8;
                if (valid && w0 == 2) {
                  // start section n-2 (most marking rays): the 2x2 square gets one value
                  const unsigned ent = wr | static_cast<uint8_t>(hit ? -1 : n - 1);
                  const bool in_x = ri + 1 < c.res_x, in_y = rj + 1 < c.res_y;
                  atomicMax(&table[rr], ent);
                  if (in_y) atomicMax(&table[rr + 1], ent);
                  if (in_x) atomicMax(&table[rr + c.res_y], ent);
                  if (in_x && in_y) atomicMax(&table[rr + c.res_y + 1], ent);
                }
                // larger squares: the warp shares the cells of one ray at a time
                unsigned big = __ballot_sync(kFull, valid && w0 > 2);
                while (big) {
                  const int src = __ffs(big) - 1;
                  big &= big - 1;
                  const int b_ri = __shfl_sync(kFull, ri, src), b_rj = __shfl_sync(kFull, rj, src);
                  const int b_w0 = __shfl_sync(kFull, w0, src), b_tl = __shfl_sync(kFull, t_last, src);
                  const int b_sh = __shfl_sync(kFull, hit ? seg_h : -1, src);
                  const unsigned b_wr = __shfl_sync(kFull, wr, src);
                  const int nx = min(b_w0, c.res_x - b_ri), ny = min(b_w0, c.res_y - b_rj);
                  for (int cell = lane; cell < nx * ny; cell += 32) {
                    const int dx = cell / ny, dy = cell - dx * ny;
                    const int mx = max(dx, dy);
                    const int tmax = mx == 0 ? n - 1 : n - 2 - (31 - __clz(mx));
                    const int t = min(tmax, b_tl);
                    const int val = (t == b_sh) ? -1 : t + 1;
                    A3D_DEV_CHECK((b_ri + dx) * c.res_y + b_rj + dy < n_rays);
                    atomicMax(&table[(b_ri + dx) * c.res_y + b_rj + dy], b_wr | static_cast<uint8_t>(val));
                  }
                }
              }
              __threadfence_block();
              if (lane == 0) s_done_diag = cur_diag + 1;
            }
            ++cur_diag;
            n_mark = 0;
            mark_head = 0;
            if (cur_diag >= n_diag) {
              sched_done = true;
              break;
            }
            // gather the rays of this diagonal that start before the last section (scan order)
            const int i_lo = max(0, cur_diag - (c.res_y - 1)), i_hi = min(c.res_x - 1, cur_diag);
            for (int base = i_lo; base <= i_hi; base += 32) {
              const int i = base + static_cast<int>(lane);
              int s = -1, r = 0;
              if (i <= i_hi) {
                r = i * c.res_y + (cur_diag - i);
                s = static_cast<int>(static_cast<int8_t>(__ldcg(&table[r]) & 0xFFu));
              }
              const bool marking = s >= 0 && s < n - 1;
              const unsigned mm = __ballot_sync(kFull, marking);
              A3D_DEV_CHECK(n_mark + __popc(mm) <= c.diag_max);
              if (marking)
                s_mark_list[n_mark + __popc(mm & lt_mask)] = (static_cast<unsigned>(r) << 8) | s;
              n_mark += __popc(mm);
            }
            __syncwarp();
          }

          // ---- 2. give work to idle lanes: marking parts first, then leaf parts ----
          const unsigned idle = __ballot_sync(kFull, role == 0);
          bool starved = false;  // lanes stayed idle: nothing to hand out right now
          if (idle) {
            int n_idle = __popc(idle);
            const int rank = __popc(idle & lt_mask);
            int take_mark = 0;
            if (scheduler) {
              take_mark = min(n_idle, n_mark - mark_head);
              if (role == 0 && rank < take_mark) {
                const unsigned it = s_mark_list[mark_head + rank];
                const int mark_idx = mark_head + rank;
                lt.mark[threadIdx.x] = mark_idx;
                const int ms = static_cast<int>(it & 0xFF);
                ray_begin(rs, lt, c, vw, static_cast<int>(it >> 8), c.split[ms]);
                start_s[it >> 8] = static_cast<uint8_t>(ms);  // read by leaf lanes after the diagonal
                lt.d_end[threadIdx.x] = d_mark_end;
                role = 2;  // split[s] < split[n-1] for s < n-1: at least one sample
              }
              mark_head += take_mark;
              n_run += take_mark;
              n_idle -= take_mark;
#if A3D_WAVE_SCHED_LEAF_MAX < 32
              // keep lanes of the scheduler warp free for the marking parts of the next anti-diagonals
              if (!sched_done)
                n_idle = min(n_idle, max(0, A3D_WAVE_SCHED_LEAF_MAX - __popc(__ballot_sync(kFull, role == 1))));
#endif
            }
            if (n_idle > 0) {
              // top up the ring from the shared supply
              while (!supply_done && pend_count < n_idle && pend_count <= kPendCap - 32) {
                int u = -1;
                if (lane == 0) {
                  const unsigned cur = *reinterpret_cast<volatile unsigned*>(&s_unit_cursor);
                  if (cur >= static_cast<unsigned>(n_units)) {
                    u = -2;
                  } else if (!ITER || c.unit_diag[cur] < s_done_diag) {
                    if (atomicCAS(&s_unit_cursor, cur, cur + 1) == cur) u = static_cast<int>(cur);
                  }
                }
                u = __shfl_sync(kFull, u, 0);
                if (u == -2) supply_done = true;
                if (u < 0) break;
                bool ready = false;
                int rr = 0;
                if (ITER) {
                  const int dg = c.unit_diag[u];
                  const int i = c.unit_i0[u] + static_cast<int>(lane);
                  if (i <= min(c.res_x - 1, dg)) {
                    rr = i * c.res_y + (dg - i);
                    ready = static_cast<int>(static_cast<int8_t>(__ldcg(&table[rr]) & 0xFFu)) == n - 1;
                  }
                } else {
                  rr = u * 32 + static_cast<int>(lane);
                  ready = rr < n_rays;
                }
                const unsigned mm = __ballot_sync(kFull, ready);
                if (ready) pending[(pend_head + pend_count + __popc(mm & lt_mask)) % kPendCap] = rr;
                pend_count += __popc(mm);
                __syncwarp();
              }
              const int lrank = rank - take_mark;
              if (role == 0 && lrank >= 0 && lrank < min(pend_count, n_idle)) {
                const int r = static_cast<int>(pending[(pend_head + lrank) % kPendCap]);
                double d0 = 0.0;
                if (ITER) {
                  const int s0 = start_s[r] == 0xFF ? n - 1 : start_s[r];
                  d0 = c.leaf_d0[s0];  // accumulated distance at which a ray started in s0 enters section n-1
                }
                ray_begin(rs, lt, c, vw, r, d0);
                lt.d_end[threadIdx.x] = d_leaf_end;
                role = (rs.d < d_leaf_end) ? 1 : 0;
              }
              const int taken = min(n_idle, pend_count);
              pend_head = (pend_head + taken) % kPendCap;
              pend_count -= taken;
              starved = taken < n_idle;
              __syncwarp();
            }
          }

          // ---- 3. anything left to do? ----
          const unsigned busy = __ballot_sync(kFull, role != 0);
          if (busy == 0) {
            if (sched_done && supply_done && pend_count == 0) break;
            if (sched_done && pend_count == 0) __nanosleep(64);  // wait for the scheduler warp
            continue;
          }

          // ---- 4. samples: every busy lane marches its own part ----
          // A round ends when (a) the marking parts in flight are all done, so that the scheduler can
          // open the next anti-diagonal without delay, (b) enough lanes ran out of ray and there is
          // work to hand out, or (c) after max_it samples.  While the supply is dry (the scheduler
          // has not released more rays yet) rounds are short instead.
          const bool marks_running = scheduler && n_run > 0;
          const int max_it = marks_running ? A3D_WAVE_MARK_STEPS
                                           : ((starved && !supply_done) ? A3D_WAVE_DRY_STEPS : A3D_WAVE_STEPS);
          const int idle_limit = (starved || supply_done) ? 33 : A3D_WAVE_IDLE_BREAK;
#pragma unroll 1
          for (int it = 0; it < max_it; ++it) {
            if (role > 0) {
              if (ray_step<CASTER, EVAL, DIGEST, EMIT>(m, ct, e, vw, bvg, bv_mode, step, rs, lt, lc, ws.rec, cta_base * 2,
                                                 role, n_rays, ws.ent, ws.kmax))
                role = -role;  // finished: bookkeeping after the round
            }
            bool stop = __popc(__ballot_sync(kFull, role <= 0)) >= idle_limit;
            if (marks_running) stop |= __popc(__ballot_sync(kFull, role == -2)) == n_run;
            if (stop) break;
          }
          // parts that ended in this round: last key for the fix-up, hit distance for the marks
          n_run -= __popc(__ballot_sync(kFull, role == -2));
          if (role < 0) {
            if (rs.prev.g0 != kNoIdent) {
              unsigned bad = 0;
              const size_t rec_at =
                  cta_base * 2 + (role == -2 ? 0 : static_cast<size_t>(n_rays)) + lt.r[threadIdx.x];
              __stcs(&ws.rec.last_key[rec_at], pack_key(rs.prev, &bad));
              if (bad) lt.inexact = 1;
              if (EMIT) {
                if (lt.kept[threadIdx.x] > ws.kmax) lt.inexact = 1;  // cannot happen (kmax bounds a part)
                ws.cnt[rec_at] = static_cast<uint16_t>(min(lt.kept[threadIdx.x], ws.kmax));
              }
            }
            if (role == -2) s_mark_hit[lt.mark[threadIdx.x]] = rs.d;
            role = 0;
          }
          __syncwarp();
        }
      }
      __syncthreads();

      // ---------------- fix-up: adjacent-unique across ray-part boundaries, scan order ----------
      if (warp == 0) {
        // lane L looks at record (ray = base + L/2, part = L&1): scan order == lane order
        for (int base = 0; base < n_rays; base += 16) {
          const int rr = base + static_cast<int>(lane >> 1);
          const int part = lane & 1;
          int info = kCatNone;
          uint64_t fk = kNoKey, lk = kNoKey;
          size_t at = 0;
          if (rr < n_rays) {
            at = cta_base * 2 + static_cast<size_t>(part) * n_rays + rr;
            info = __ldcs(&ws.rec.info[at]);
            if (info & 8) {
              fk = __ldcs(&ws.rec.first_key[at]);
              lk = __ldcs(&ws.rec.last_key[at]);
            }
          }
          const bool emitted = (info & 8) != 0;
          const unsigned em = __ballot_sync(kFull, emitted);
          const unsigned below = em & lt_mask;
          const int src = below ? 31 - __clz(below) : 0;
          uint64_t prev = __shfl_sync(kFull, lk, src);
          if (!below) prev = carry_key;
          const bool drop = emitted && fk == prev && (info & 16);
          if (EMIT) {
            // where this part's kept entries go in the segment's ordered list
            const unsigned n_out = emitted ? static_cast<unsigned>(ws.cnt[at]) - (drop ? 1u : 0u) : 0u;
            unsigned incl = n_out;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const unsigned t = __shfl_up_sync(kFull, incl, o);
              if (static_cast<int>(lane) >= o) incl += t;
            }
            if (emitted) ws.off[at] = (emit_pos + incl - n_out) | (drop ? 0x80000000u : 0u);
            emit_pos += __shfl_sync(kFull, incl, 31);
          }
          if (drop) {
            // the first entry of this part repeats the previous raw entry: std::unique drops it
            lc.packed -= 1ull << (kFieldBits * (info & 7));
            if (DIGEST) lc.digest -= ws.rec.first_hash[at];
            if (EVAL == A3D_EVAL_VOXEL_WEIGHT)
              lc.wsum -= __longlong_as_double(static_cast<long long>(ws.rec.first_hash[at]));
          }
          if (em) carry_key = __shfl_sync(kFull, lk, 31 - __clz(em));
        }
      }

      // ---------------- list emission: every warp copies ray parts into the ordered list ----------
      if (EMIT) {
        __syncthreads();
        const long long emit_base = b.emit_offset ? b.emit_offset[seg] : 0;
        const long long emit_cap = b.emit_offset ? b.emit_offset[seg + 1] - emit_base : b.cap;
        for (int q = static_cast<int>(warp); q < 2 * n_rays; q += WARPS) {
          const size_t at = cta_base * 2 + static_cast<size_t>(q & 1) * n_rays + (q >> 1);
          if (!(ws.rec.info[at] & 8)) continue;
          const unsigned o = ws.off[at];
          const int skip = static_cast<int>(o >> 31);
          const int n_out = static_cast<int>(ws.cnt[at]) - skip;
          const long long first_idx = static_cast<long long>(o & 0x7FFFFFFFu);
          for (int k = static_cast<int>(lane); k < n_out; k += 32) {
            const long long idx = first_idx + k;
            if (idx >= emit_cap) break;
            const unsigned long long key = __ldcs(&ws.ent[at * static_cast<size_t>(ws.kmax) + k + skip]);
            A3D_DEV_CHECK(k + skip < static_cast<int>(ws.cnt[at]) && ws.cnt[at] <= ws.kmax);
            if (b.keys_out) {
              b.keys_out[emit_base + idx] = key;
            } else {
              const D3 cc = key_centre(m, ct, key);
              double* dst = b.centres_out + 3 * (emit_base + idx);
              dst[0] = cc.x;
              dst[1] = cc.y;
              dst[2] = cc.z;
            }
          }
        }
      }
    }  // views

    // ---------------- per-segment reduction ---------------------------------------------------------
    // Lane accumulators are modular (the fix-up may decrement a lane that never incremented); the
    // field sums are exact as long as every per-segment count stays below 2^21 per field — larger
    // segments are flagged for the scan-order kernel.
    if (threadIdx.x == 0) {
      s_packed = 0;
      s_digest = 0;
      s_wsum = 0.0;
      s_nsamp = 0;
    }
    __syncthreads();
    {
      unsigned long long pk = lc.packed, dg = lc.digest;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        pk += __shfl_xor_sync(kFull, pk, off);
        if (DIGEST) dg += __shfl_xor_sync(kFull, dg, off);
      }
      double wv = lc.wsum;
      if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) wv += __shfl_xor_sync(kFull, wv, off);
        if (lane == 0) atomicAdd(&s_wsum, wv);
      }
      const unsigned tsmp = __reduce_add_sync(kFull, lc.n_samp);
      if (lane == 0) {
        atomicAdd(&s_packed, pk);
        atomicAdd(&s_nsamp, tsmp);
        if (DIGEST) atomicAdd(&s_digest, dg);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long fmask = (1ull << kFieldBits) - 1;
      const unsigned long long c0 = s_packed & fmask, c1 = (s_packed >> kFieldBits) & fmask,
                               c2 = (s_packed >> (2 * kFieldBits)) & fmask;
      unsigned inexact = lt.inexact;
      // a field may have wrapped if the segment has 2^21 or more samples
      if (s_nsamp >= (1u << kFieldBits)) inexact = 1;
      const unsigned long long n_vis_total = c0 + c1 + c2;
      // VoxelTypeEvaluator's switch without breaks (voxel_type_evaluator.cpp:69-85)
      double g = 0.0;
      if (EVAL == A3D_EVAL_VOXEL_TYPE) {
        const double n_unk = static_cast<double>(c2);
        const double n_free = static_cast<double>(c1) + n_unk;
        const double n_occ = static_cast<double>(c0) + n_free;
        g += n_unk * e.gains[0];
        g += n_free * e.gains[2];
        g += n_occ * e.gains[1];
      } else if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
        g = s_wsum;
      } else {
        g = static_cast<double>(c0);
      }
      b.gains[seg] = g;
      if (b.n_visible) b.n_visible[seg] = n_vis_total;
      if (b.n_samples) b.n_samples[seg] = s_nsamp;
      if (DIGEST && b.set_digests) b.set_digests[seg] = s_digest;
      if (inexact) atomicOr(&inexact_flags[seg >> 5], 1u << (seg & 31));
    }
  }
}

#if A3D_WAVE_PART == 1
size_t wave_scratch_bytes(int n_rays, int grid, bool digest, size_t* rays_stride, int emit_kmax) {
  const size_t stride = (static_cast<size_t>(n_rays) + 31) / 32 * 32;
  *rays_stride = stride;
  size_t per_cta = stride * 4 + stride + 2 * stride * (8 + 8 + 1) + (digest ? 2 * stride * 8 : 0);
  if (emit_kmax > 0) per_cta += 2 * stride * (static_cast<size_t>(emit_kmax) * 8 + 2 + 4);
  return per_cta * static_cast<size_t>(grid) + 256;
}
#endif

template <int CASTER, int EVAL>
static void launch_wave_t(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e,
                          const Batch& b, bool digest, int grid, const WaveScratch& ws,
                          unsigned int* inexact, int warps) {
  const size_t smem = static_cast<size_t>(c.diag_max) * 12;
  if (ws.ent)  // list emission: always the wide CTA, never together with a digest
    k_cast_wave<CASTER, EVAL, false, kWaveWideWarps, true>
        <<<grid, 32 * kWaveWideWarps, smem, st>>>(m, c, e, b, ws, inexact);
  else if (digest && warps >= kWaveWideWarps)
    k_cast_wave<CASTER, EVAL, true, kWaveWideWarps, false>
        <<<grid, 32 * kWaveWideWarps, smem, st>>>(m, c, e, b, ws, inexact);
  else if (digest && warps >= kWaveMidWarps)
    k_cast_wave<CASTER, EVAL, true, kWaveMidWarps, false>
        <<<grid, 32 * kWaveMidWarps, smem, st>>>(m, c, e, b, ws, inexact);
  else if (digest)
    k_cast_wave<CASTER, EVAL, true, kWaveWarps, false><<<grid, kWaveThreads, smem, st>>>(m, c, e, b, ws, inexact);
  else if (warps >= kWaveWideWarps)
    k_cast_wave<CASTER, EVAL, false, kWaveWideWarps, false>
        <<<grid, 32 * kWaveWideWarps, smem, st>>>(m, c, e, b, ws, inexact);
  else if (warps >= kWaveMidWarps)
    k_cast_wave<CASTER, EVAL, false, kWaveMidWarps, false>
        <<<grid, 32 * kWaveMidWarps, smem, st>>>(m, c, e, b, ws, inexact);
  else
    k_cast_wave<CASTER, EVAL, false, kWaveWarps, false><<<grid, kWaveThreads, smem, st>>>(m, c, e, b, ws, inexact);
}

template <int CASTER>
static void launch_wave_c(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e,
                          const Batch& b, bool digest, int grid, const WaveScratch& ws,
                          unsigned int* inexact, int warps) {
  switch (e.kind) {
    case A3D_EVAL_VOXEL_TYPE:
      launch_wave_t<CASTER, A3D_EVAL_VOXEL_TYPE>(st, m, c, e, b, digest, grid, ws, inexact, warps);
      break;
    case A3D_EVAL_NAIVE:
      launch_wave_t<CASTER, A3D_EVAL_NAIVE>(st, m, c, e, b, digest, grid, ws, inexact, warps);
      break;
    case A3D_EVAL_VOXEL_WEIGHT:  // (never with a digest: the value record shares the hash slot)
      launch_wave_t<CASTER, A3D_EVAL_VOXEL_WEIGHT>(st, m, c, e, b, false, grid, ws, inexact, warps);
      break;
    default:
      launch_wave_t<CASTER, A3D_EVAL_FRONTIER>(st, m, c, e, b, digest, grid, ws, inexact, warps);
      break;
  }
}

#if A3D_WAVE_PART == 1
void launch_wave_iterative(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                           bool digest, int grid, const void* ws, unsigned int* inexact, int warps) {
  launch_wave_c<A3D_CASTER_ITERATIVE>(st, m, c, e, b, digest, grid, *static_cast<const WaveScratch*>(ws), inexact,
                                      warps);
}
#else
void launch_wave_simple(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                        bool digest, int grid, const void* ws, unsigned int* inexact, int warps) {
  launch_wave_c<A3D_CASTER_SIMPLE>(st, m, c, e, b, digest, grid, *static_cast<const WaveScratch*>(ws), inexact,
                                   warps);
}
#endif

#if A3D_WAVE_PART == 1
// scratch: wave_scratch_bytes() bytes, carved here.  b.set_digests receives the ORDER-INDEPENDENT
// digest (sum of digest_entry over the stored list, mod 2^64).  inexact: one bit per segment, set
// when a voxel key overflowed the 20-bit packing (the host re-evaluates those with k_cast_gain).
void launch_cast_wave(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e,
                      const Batch& b, bool digest, int grid, void* scratch, unsigned int* inexact,
                      int warps) {
  size_t stride = 0;
  const bool extra = digest || e.kind == A3D_EVAL_VOXEL_WEIGHT;  // per-part first_hash / first value
  const int emit_kmax = (b.centres_out || b.keys_out) ? c.kmax : 0;
  wave_scratch_bytes(c.n_rays, grid, extra, &stride, emit_kmax);
  WaveScratch ws{};
  ws.rays_stride = stride;
  uint8_t* p = static_cast<uint8_t*>(scratch);
  const size_t g = static_cast<size_t>(grid);
  ws.rec.first_key = reinterpret_cast<unsigned long long*>(p);
  p += g * stride * 2 * 8;
  ws.rec.last_key = reinterpret_cast<unsigned long long*>(p);
  p += g * stride * 2 * 8;
  if (extra) {
    ws.rec.first_hash = reinterpret_cast<unsigned long long*>(p);
    p += g * stride * 2 * 8;
  }
  if (emit_kmax > 0) {
    ws.ent = reinterpret_cast<unsigned long long*>(p);
    p += g * stride * 2 * 8 * static_cast<size_t>(emit_kmax);
    ws.kmax = emit_kmax;
  }
  ws.table = reinterpret_cast<uint32_t*>(p);
  p += g * stride * 4;
  if (emit_kmax > 0) {
    ws.off = reinterpret_cast<uint32_t*>(p);
    p += g * stride * 2 * 4;
    ws.cnt = reinterpret_cast<uint16_t*>(p);
    p += g * stride * 2 * 2;
  }
  ws.rec.info = p;
  p += g * stride * 2;
  ws.start_s = p;
  if (c.kind == A3D_CASTER_ITERATIVE)
    launch_wave_iterative(st, m, c, e, b, digest, grid, &ws, inexact, warps);
  else
    launch_wave_simple(st, m, c, e, b, digest, grid, &ws, inexact, warps);
}

int cast_wave_threads() { return kWaveThreads; }
int cast_wave_max_diag() { return kMarkCap; }
int cast_wave_blocks_per_sm() { return A3D_WAVE_MIN_BLOCKS; }
#endif  // A3D_WAVE_PART == 1

}  // namespace a3d
