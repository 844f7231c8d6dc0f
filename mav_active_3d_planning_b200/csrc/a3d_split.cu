// a3d_split.cu — the throughput path of the view-gain evaluation as a chain of small kernels
// (DESIGN.md "kernel K3s").  Same results as k_cast_wave / k_cast_gain for gain, n_visible and
// n_samples; used when only those are asked (no digests, no list emission) and the dense meta volume
// exists.
//
// k_cast_wave does everything for a segment inside one 2-warp CTA — ray table, marking parts, leaf
// parts, fix-up — and needs 96 registers for it: 20 warps per SM, bound by the latency of the meta
// word load.  Here the work is cut along its dependencies instead, so that the bulk of it (the last
// section of every ray, 80 % of the samples) runs in a kernel that does nothing else:
//
//   k_split_views  camera pose (camera_model.cpp:24-28) and the constants of the certified index path
//                  of every view of the batch.
//   k_split_mark   IterativeRayCaster only (iterative_ray_caster.cpp:48-119): one warp per view walks
//                  the anti-diagonals of the ray table, casts the rays that start before the last
//                  section up to c_split_distances_[n-1] ("marking parts", one per lane) and applies
//                  markNeighboringRays with atomicMax on (writer << 8 | value): an entry keeps the
//                  value of its highest writer = last writer in scan order.  Leaves the final table.
//   k_split_leaf   every ray whose table entry reads n-1 (SimpleRayCaster: every ray) is cast through
//                  its last section (whole ray), one ray per lane, 32 consecutive rays of a view per
//                  warp, any warp takes any unit: no per-segment state, no tail.
//   k_split_fix    per segment, scan order: std::unique across ray-part boundaries
//                  (camera_model.cpp:58) from the parts' first / last keys, sums the parts' counters,
//                  folds the gain (voxel_type_evaluator.cpp:57-89 | naive_evaluator.cpp:20-38 |
//                  frontier_evaluator.cpp:100-124 | voxel_weight_evaluator.cpp:43-83).
//
// A ray part leaves one record: key of its first and last raw list entry, a 64-bit word of four
// 15-bit counters (entries per fold category, samples) + what the first entry was counted as, and for
// VoxelWeightEvaluator the sum of the values of the entries that have one of their own (surface voxels;
// unobserved voxels are counted per constant), without the first entry's (the fix-up decides whether the
// first entry stays, so nothing is added and subtracted again, and the sums are formed in a fixed order:
// gains are reproducible run to run).
//
// Stored lists (a3d_compute_gains_stored) run the chain twice: the first pass also leaves every ray
// part's offset within its segment's list (SplitScratch::off, from the fix-up); once the arena has been
// sized from the segments' counts, the DIRECT instantiations of the mark / leaf kernels cast again and
// write every retained entry's key straight to its final place (launch_cast_split_direct; compiled as a
// second translation unit, a3d_split_direct.cu, in parallel with this one).
#include <algorithm>

#include "../../include/a3d.h"
#include "a3d_kernels.cuh"
#include "a3d_keys.cuh"
#include "a3d_sample.cuh"

#ifndef A3D_SPLIT_PART
#define A3D_SPLIT_PART 1  // 1: this file as such; 2: the DIRECT instantiations (a3d_split_direct.cu)
#endif

namespace a3d {

namespace {

constexpr int kSplitWarps = 4;  // warps per CTA (independent: no CTA-wide barrier after the prologue)
constexpr int kSplitThreads = 32 * kSplitWarps;
#ifndef A3D_SPLIT_LEAF_BLOCKS
#define A3D_SPLIT_LEAF_BLOCKS 9  // resident CTAs per SM the leaf kernel's register allocation must allow
#endif
#ifndef A3D_SPLIT_MARK_BLOCKS
#define A3D_SPLIT_MARK_BLOCKS 7
#endif
// (measured on cfg2, 4096 views: VoxelType 9 CTAs 3.35 ms / 8: 3.45 / 6: 4.1; Frontier 8: 4.38 / 7: 4.20 / 6: 4.40;
// VoxelWeight 9: 5.52 / 8: 5.41 / 7: 5.20 / 6: 5.27 — 72 registers spill no more than 80 do)
#ifndef A3D_SPLIT_LEAF_BLOCKS_VW
#define A3D_SPLIT_LEAF_BLOCKS_VW 7
#endif
#ifndef A3D_SPLIT_LEAF_BLOCKS_FR
#define A3D_SPLIT_LEAF_BLOCKS_FR 7
#endif
constexpr int leaf_blocks(int caster, int eval) {
  return caster != A3D_CASTER_ITERATIVE ? 5
         : eval == A3D_EVAL_VOXEL_WEIGHT ? A3D_SPLIT_LEAF_BLOCKS_VW
         : eval == A3D_EVAL_FRONTIER     ? A3D_SPLIT_LEAF_BLOCKS_FR
                                         : A3D_SPLIT_LEAF_BLOCKS;
}
constexpr int kSplitMaxDiag = 256;  // longest anti-diagonal k_split_mark keeps in shared memory

// record word: 4 x 15-bit counters, then what the first entry was counted as
constexpr int kCntBits = 15;
constexpr unsigned long long kCntMask = (1ull << kCntBits) - 1;
constexpr int kCntSampShift = 3 * kCntBits;
constexpr int kRecCatShift = 60;                 // bits 61:60 fold category of the first entry
constexpr unsigned long long kRecEmitted = 1ull << 62;  // the part emitted at least one raw entry
constexpr unsigned long long kRecKept = 1ull << 63;     // its first entry was retained (counted)

struct ViewRec {
  double P0[3];   // camera position
  double q[4];    // camera orientation (w, x, y, z)
  double org[3];  // VoxelWeightEvaluator: reference point of the view's segment
  // certified index path: half-voxel coordinate z = fma(float(d), dz, z0), certified when
  // |z - rint(z)| < lim on all axes (lim NaN: never)
  float z0[3], lim;
};

struct SplitScratch {
  ViewRec* views;                 // [n_views]
  unsigned int* view_flags;       // [n_views] != 0: a sample of the view needs the scan-order kernel
  uint32_t* table;                // [n_views][stride] (writer + 1) << 8 | uint8(value)   (iterative)
  uint8_t* start_s;               // [n_views][stride] start section of marking rays, 0xFF (iterative)
  uint32_t* dlist;                // [n_views][stride] rays that start in section n-2 (k_split_mark2)
  unsigned int* dcount;           // [n_views] how many
  unsigned long long* first_key;  // [n_views][parts][stride]
  unsigned long long* last_key;
  unsigned long long* rec;        // counters + first-entry info
  double* wsum;                   // VoxelWeightEvaluator: sum of the part's category-1 entry values, first entry excluded
  double* first_val;              // ... value of the first entry (0 unless it is a category-1 entry with a value)
  size_t stride;                  // n_rays rounded up to 32
  int parts;                      // 2 (iterative: marking part, leaf part) or 1
  // list emission (null / 0 when only gains are asked): the retained entries of every ray part as
  // packed keys, kmax slots per part; where the part's entries go in its segment's ordered list
  unsigned long long* ent;        // [n_views][parts][stride][kmax]
  uint32_t* off;                  // [n_views][parts][stride]: offset in the segment's list | (first entry dropped) << 31
  int32_t* view_seg;              // [n_views] segment of each view (batch-wide numbering)
  int kmax;
  // second pass of a stored-list call (DIRECT kernels): keys go straight to direct[seg_base[segment] + offset of
  // the part (off, first pass) + index within the part]
  unsigned long long* direct;
  const long long* seg_base;      // [n_segments + 1]
};

// Where a ray part's retained entries go in the second pass of a stored-list call.
struct DirectOut {
  unsigned long long* dst;  // place of the part's entry number `skip`
  int skip;                 // 1: std::unique dropped the part's first entry (fix-up of the first pass)
  int room;                 // entries left in the segment's range from dst on
};
__device__ __forceinline__ DirectOut direct_begin(const SplitScratch& sc, int view, size_t at) {
  const uint32_t o = __ldg(&sc.off[at]);
  const int seg = __ldg(&sc.view_seg[view]);
  const long long base = __ldg(&sc.seg_base[seg]), len = __ldg(&sc.seg_base[seg + 1]) - base;
  const long long first = static_cast<long long>(o & 0x7FFFFFFFu);
  DirectOut d;
  d.dst = sc.direct + base + first;
  d.skip = static_cast<int>(o >> 31);
  d.room = static_cast<int>(max(0ll, min(len - first, 2147483647ll)));
  return d;
}
__device__ __forceinline__ void direct_put(const DirectOut& d, int k, unsigned long long key) {
  const int i = k - d.skip;
  if (i >= 0 && i < d.room) __stcs(d.dst + i, key);
}

// entries a part has retained so far, from its record word
__device__ __forceinline__ int rec_kept(unsigned long long rec) {
  return static_cast<int>((rec & kCntMask) + ((rec >> kCntBits) & kCntMask) + ((rec >> (2 * kCntBits)) & kCntMask));
}

__host__ __device__ inline uint32_t vol_tab_entries(const MapView& m) {
  const uint32_t n = m.vol_n[0] + m.vol_n[1] + m.vol_n[2];
  return n <= 3072u ? n : 0u;
}

// per-axis terms of vol_offset() (the brick offset is their sum)
__device__ __forceinline__ void fill_vol_tab(const MapView& m, uint32_t* tab, uint32_t n) {
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    uint32_t u = i, v;
    if (u < m.vol_n[0]) {
      v = ((u >> 2) << 5) | (u & 3u);
    } else if ((u -= m.vol_n[0]) < m.vol_n[1]) {
      v = (((u >> 2) * m.vol_bx) << 5) | ((u & 3u) << 2);
    } else {
      u -= m.vol_n[1];
      v = (((u >> 1) * m.vol_by * m.vol_bx) << 5) | ((u & 1u) << 4);
    }
    tab[i] = v;
  }
}

// meta word of voxel g from the dense volume (kWordNoBlock outside of it)
__device__ __forceinline__ uint32_t vol_word(const MapView& m, const int (&g)[3],
                                             const uint32_t* __restrict__ tab) {
  const unsigned u0 = static_cast<unsigned>(g[0] - m.vol_lo[0]), u1 = static_cast<unsigned>(g[1] - m.vol_lo[1]),
                 u2 = static_cast<unsigned>(g[2] - m.vol_lo[2]);
  uint32_t word = kWordNoBlock;
  if ((u0 < m.vol_n[0]) & (u1 < m.vol_n[1]) & (u2 < m.vol_n[2])) {
    const uint32_t off = tab ? tab[u0] + tab[m.vol_n[0] + u1] + tab[m.vol_n[0] + m.vol_n[1] + u2]
                             : vol_offset(m, u0, u1, u2);
    A3D_DEV_CHECK(off == vol_offset(m, u0, u1, u2));
    word = __ldg(m.vol + off);
  }
  return word;
}

// What every kernel holds per CTA in shared memory.
struct CtaShared {
  CenterTables ct;
  BvGrid bvg;
  ViewRec vw[kSplitWarps];          // the view each warp is working on
  double dir[3][kSplitThreads];     // world direction of the lane's ray (literal paths only)
};

// State of the ray part a lane is marching.
struct Part {
  double d;      // distance of the next sample (accumulated like `distance += p_ray_step_`)
  double d_end;  // the part ends before this distance
  Ident prev;    // previous raw entry (g0 == kNoIdent: none yet)
  unsigned long long rec;  // counters + first-entry info
  float dz[3];   // 2 * dir / voxel_size
  size_t at;     // index of the part's record
  DirectOut out; // DIRECT kernels only
};

template <class S>
__device__ __forceinline__ void part_begin(Part& p, S& sh, const MapView& m, const CasterView& c, const ViewRec& vw,
                                           int r, double d0, double d_end, size_t at) {
  const Q4 q{vw.q[0], vw.q[1], vw.q[2], vw.q[3]};
  const D3 dir = rotate(q, D3{c.cam_dirs[3 * r], c.cam_dirs[3 * r + 1], c.cam_dirs[3 * r + 2]});
  sh.dir[0][threadIdx.x] = dir.x;
  sh.dir[1][threadIdx.x] = dir.y;
  sh.dir[2][threadIdx.x] = dir.z;
  const double k = 2.0 * static_cast<double>(m.vs_inv_f);
  p.dz[0] = __double2float_rn(__dmul_rn(dir.x, k));
  p.dz[1] = __double2float_rn(__dmul_rn(dir.y, k));
  p.dz[2] = __double2float_rn(__dmul_rn(dir.z, k));
  p.d = d0;
  p.d_end = d_end;
  p.prev = Ident{kNoIdent, 0, 0};
  p.rec = 0;
  p.at = at;
}

// One sample, up to the point where the list logic takes over: voxel state at the sample point,
// identity / centre state of the voxel getVoxelCenter names, bounding-volume test.
//
// Index arithmetic: with T = pf / voxel_size per axis, floor(2T) names the voxel (g = floor(2T) >> 1)
// and the half of it the point lies in (interpolation cell octant).  voxblox reaches the same through
// ~8 float roundings whose combined error in voxel units is below
// delta = 5e-7 max|T| + (vps+1)(16 * 2^-24 + 1.1e-6) (a3d_wave.cu, tests/test_certified_index.py).
// The test does not need the reference's point, only a number close to it: z = fma(float(d), dz, z0)
// = 2T - 1/2 up to 4e-7 max|T| (three fp32 roundings on terms <= 2 max|T|, plus the two of the
// reference's own fp32 T).  When |z - rint(z)| < 1/2 - 2 (delta + 4e-7 max|T|) on all axes the sample
// is "certified": rint(z) = floor(2T), block and voxel indices are in range, getVoxelState's and
// getVoxelCenter's voxels coincide, and it costs one load from the dense meta volume.  The others
// (~0.3 %) and the MIXED cells form the exact fp64 point and are evaluated literally.
struct Eval {
  SampleInfo si;
  int g[3];       // voxel coordinate (certified samples)
  uint32_t word;  // meta word of that voxel (certified samples)
  bool safe;      // certified
  bool in_bv;
  bool have_cc;
  D3 cc;          // centre of the entry's voxel, formed only where something needs it
};

template <class S>
__device__ __forceinline__ Eval sample_eval(const MapView& m, const EvalView& e, const S& sh, const ViewRec& vw,
                                            const int bv_mode, const uint32_t* __restrict__ tab, const double d_this,
                                            const float (&dz)[3], unsigned& inexact) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: (z + kMagic) - kMagic = rint(z) for |z| < 2^22
  const CenterTables& ct = sh.ct;
  const BvGrid& bvg = sh.bvg;
  const float df = __double2float_rn(d_this);
  Eval ev;
  int oct = 0;
  bool safe = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float z = __fmaf_rn(df, dz[a], vw.z0[a]);
    const float s = __fadd_rn(z, kMagic);
    const float fr = __fsub_rn(z, __fsub_rn(s, kMagic));  // z - rint(z), exact, in [-1/2, 1/2]
    safe &= fabsf(fr) < vw.lim;
    const int h = __float_as_int(s) - 0x4B400000;  // rint(z) = floor(2T)
    ev.g[a] = h >> 1;
    oct |= (~h & 1) << a;  // bit set: the point is below the voxel centre on this axis
  }
  ev.cc = D3{0.0, 0.0, 0.0};
  ev.have_cc = false;
  ev.in_bv = true;
  ev.word = kWordNoBlock;
  double pp[3];
  float pf[3];
  auto point = [&]() {  // the reference's sample point, double and float
    pp[0] = __dadd_rn(vw.P0[0], __dmul_rn(d_this, sh.dir[0][threadIdx.x]));
    pp[1] = __dadd_rn(vw.P0[1], __dmul_rn(d_this, sh.dir[1][threadIdx.x]));
    pp[2] = __dadd_rn(vw.P0[2], __dmul_rn(d_this, sh.dir[2][threadIdx.x]));
#pragma unroll
    for (int a = 0; a < 3; ++a) pf[a] = __double2float_rn(pp[a]);
  };
  if (!safe) {
    // Margin not met: the reference's own index arithmetic, inline.  If it stays in range and
    // getVoxelState's voxel equals getVoxelCenter's, the sample is an ordinary one after all.
    point();
    int oct_l = 0;
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int bi = grid_index(pf[a], m.bs_inv_f);
      const float origin = origin_point(bi, m.bs_f);
      const int vs = grid_index(__fsub_rn(pf[a], origin), m.vs_inv_f);
      const int vi = grid_index(__double2float_rn(__dsub_rn(pp[a], static_cast<double>(origin))), m.vs_inv_f);
      ok &= (static_cast<unsigned>(vs) < static_cast<unsigned>(m.vps)) & (vs == vi);
      const float centre = __fadd_rn(origin, center_point(min(max(vs, 0), m.vps - 1), m.vs_f));
      oct_l |= (__fsub_rn(pf[a], centre) < 0.0f) ? (1 << a) : 0;
      ev.g[a] = bi * m.vps + vs;
    }
    oct = oct_l;
    safe = ok && vw.lim == vw.lim;  // (NaN margin: certification is off)
  }
  ev.safe = safe;
  if (safe) {
    const uint32_t word = vol_word(m, ev.g, tab);
    ev.word = word;
    const int cls = (word >> (2 * oct)) & 3;
    static_assert(kCellUnknown == 0 && kCellFree == 1 && kCellOccupied == 2, "state = 2 - class");
    ev.si.st = 2 - cls;
    ev.si.cs = centre_state_from_word(word);
    ev.si.id = Ident{ev.g[0], ev.g[1], ev.g[2]};
    if (cls == kCellMixed) {
      point();
      ev.si.st = mixed_state(m, ct, pf[0], pf[1], pf[2], ev.g[0], ev.g[1], ev.g[2], oct);
    }
    if (bv_mode == 2) {
      const int vmask = m.vps - 1;
      ev.cc.x = __dadd_rn(static_cast<double>(origin_point(ev.g[0] >> m.vps_shift, m.bs_f)), ct.d[(ev.g[0] & vmask) + 1]);
      ev.cc.y = __dadd_rn(static_cast<double>(origin_point(ev.g[1] >> m.vps_shift, m.bs_f)), ct.d[(ev.g[1] & vmask) + 1]);
      ev.cc.z = __dadd_rn(static_cast<double>(origin_point(ev.g[2] >> m.vps_shift, m.bs_f)), ct.d[(ev.g[2] & vmask) + 1]);
      ev.have_cc = true;
      ev.in_bv = bv_contains(e, ev.cc);  // simulated_sensor_evaluator.cpp:50-57
    } else if (bv_mode) {
      ev.in_bv = (ev.g[0] >= bvg.lo[0]) & (ev.g[0] <= bvg.hi[0]) & (ev.g[1] >= bvg.lo[1]) & (ev.g[1] <= bvg.hi[1]) &
                 (ev.g[2] >= bvg.lo[2]) & (ev.g[2] <= bvg.hi[2]);
    }
  } else {
    ExactOut out;
    unsigned bad = 0;
    point();
    sample_exact(m, ct, pp[0], pp[1], pp[2], &out, &bad);
    inexact |= bad;
    ev.si = out.si;
    ev.cc = out.cc;
    ev.have_cc = true;
    if (bv_mode) ev.in_bv = bv_contains(e, ev.cc);
  }
  return ev;
}

// The evaluators' per-entry work beyond a look at the meta word lives in functions of its own
// (not inlined): their registers would otherwise be charged to the sample loop, which runs at 64.
__device__ __forceinline__ D3 centre_of_voxel(const MapView& m, const CenterTables& ct, int g0, int g1, int g2) {
  const int vmask = m.vps - 1;
  return D3{__dadd_rn(static_cast<double>(origin_point(g0 >> m.vps_shift, m.bs_f)), ct.d[(g0 & vmask) + 1]),
            __dadd_rn(static_cast<double>(origin_point(g1 >> m.vps_shift, m.bs_f)), ct.d[(g1 & vmask) + 1]),
            __dadd_rn(static_cast<double>(origin_point(g2 >> m.vps_shift, m.bs_f)), ct.d[(g2 & vmask) + 1])};
}
// FrontierEvaluator::isFrontierVoxel (frontier_evaluator.cpp:69-98) where the precomputed bit of the
// voxel's meta word does not apply: neighbours' cell classes, else literally.
__device__ __noinline__ bool frontier_test_slow(const MapView& m, const CenterTables& ct, const EvalView& e,
                                                int g0, int g1, int g2, double cx, double cy, double cz,
                                                bool have_cc, bool safe) {
  const D3 cc = have_cc ? D3{cx, cy, cz} : centre_of_voxel(m, ct, g0, g1, g2);
  const int hint = safe ? is_frontier_by_meta(m, ct, e, cc, g0, g1, g2) : -1;
  return hint >= 0 ? hint != 0 : is_frontier_voxel(m, ct, e, cc);
}
// VoxelWeightEvaluator::getVoxelValue (voxel_weight_evaluator.cpp:60-83) of a retained entry that is
// not FREE.
__device__ __noinline__ double voxel_weight_entry(const MapView& m, const CenterTables& ct, const EvalView& e,
                                                  int g0, int g1, int g2, double cx, double cy, double cz,
                                                  bool have_cc, bool safe, uint32_t word, int cs, double ox,
                                                  double oy, double oz) {
  int hint = -1;
  if (safe && (cs & 3) == 2 && e.frontier_voxel_weight > 0.0) hint = is_frontier_from_word(m, e, word);
  if ((cs & 3) == 2 && (hint >= 0 || !(e.frontier_voxel_weight > 0.0)))
    return (hint > 0) ? e.frontier_voxel_weight : e.new_voxel_weight;  // nothing else to look at
  const D3 cc = have_cc ? D3{cx, cy, cz} : centre_of_voxel(m, ct, g0, g1, g2);
  double w = 0.0;
  if ((cs & 3) == 0) w = voxel_weight(m, cc);
  if (safe && hint < 0 && (cs & 3) == 2 && e.frontier_voxel_weight > 0.0)
    hint = is_frontier_by_meta(m, ct, e, cc, g0, g1, g2);
  return voxel_weight_value(m, ct, e, cc, cs, w, D3{ox, oy, oz}, hint);
}

// Fold category of a retained (or not) list entry and, for VoxelWeightEvaluator, its value.
template <int EVAL, class S>
__device__ __forceinline__ int entry_category(const MapView& m, const EvalView& e, const S& sh, const ViewRec& vw,
                                              Eval& ev, const bool keep, double& wval) {
  const CenterTables& ct = sh.ct;
  const SampleInfo& si = ev.si;
  int cat;
  wval = 0.0;
  if (EVAL == A3D_EVAL_VOXEL_TYPE) {
    cat = si.cs & 3;  // retained entries are always inside the bounding volume
  } else if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
    // Unobserved voxels score one of two constants (voxel_weight_evaluator.cpp:72-80): they are counted —
    // category 0: frontier_voxel_weight, 2: new_voxel_weight — and priced by the fix-up.  Category 1:
    // everything else; only there an entry has a value of its own (surface voxels, and the unobserved
    // voxels whose frontier test is not in the meta word).
    cat = 1;
    if (keep && (si.cs & 3) != 1) {  // (free voxels score nothing)
      const bool fw = e.frontier_voxel_weight > 0.0;
      const int hint = ((si.cs & 3) == 2 && ev.safe && fw) ? is_frontier_from_word(m, e, ev.word) : -1;
      if ((si.cs & 3) == 2 && (hint >= 0 || !fw))
        cat = hint > 0 ? 0 : 2;
      else
        wval = voxel_weight_entry(m, ct, e, ev.g[0], ev.g[1], ev.g[2], ev.cc.x, ev.cc.y, ev.cc.z, ev.have_cc, ev.safe,
                                  ev.word, si.cs, vw.org[0], vw.org[1], vw.org[2]);
    }
  } else {
    cat = 1;
    if (keep && !(si.cs & 4)) {
      if (EVAL == A3D_EVAL_NAIVE) {
        cat = 0;
      } else {
        const int hint = ev.safe ? is_frontier_from_word(m, e, ev.word) : -1;
        if (hint >= 0 ? hint != 0
                      : frontier_test_slow(m, ct, e, ev.g[0], ev.g[1], ev.g[2], ev.cc.x, ev.cc.y, ev.cc.z, ev.have_cc,
                                           ev.safe))
          cat = 0;
      }
    }
  }
  return cat;
}

// One sample of the lane's part (camera_model / ray caster inner loop + the evaluator's per-voxel
// test).  Returns true when the part is finished (hit or end reached) and then leaves in p.d the
// distance of the OCCUPIED sample that ended it, or -1.  Precondition: p.d < p.d_end.
template <int CASTER, int EVAL, bool DIRECT, class S>
__device__ __forceinline__ bool part_step(const MapView& m, const EvalView& e, const S& sh, const ViewRec& vw,
                                          const int bv_mode, const double step, const uint32_t* __restrict__ tab,
                                          const SplitScratch& sc, Part& p, unsigned& inexact) {
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  const double d_this = p.d;
  Eval ev = sample_eval(m, e, sh, vw, bv_mode, tab, d_this, p.dz, inexact);
  p.d = __dadd_rn(d_this, step);
  const bool more = p.d < p.d_end;
  p.rec += 1ull << kCntSampShift;
  const SampleInfo& si = ev.si;
  const bool emit = ITER || si.st != 0;
  const bool keep = emit & ev.in_bv &
                    ((si.id.g0 != p.prev.g0) | (si.id.g1 != p.prev.g1) | (si.id.g2 != p.prev.g2));
  const bool first = emit & (p.prev.g0 == kNoIdent);
  if (emit) p.prev = si.id;
  double wval;
  const int cat = entry_category<EVAL>(m, e, sh, vw, ev, keep, wval);
  if (EVAL == A3D_EVAL_VOXEL_WEIGHT && wval != 0.0) {
    // rare (a surface voxel at the end of a ray): straight to the part's record, no register held for it
    if (first) __stcs(&sc.first_val[p.at], wval); else sc.wsum[p.at] += wval;
  }
  if (DIRECT) {
    if (keep) {  // second pass of a stored-list call: the retained entry at its final place
      unsigned bad = 0;
      direct_put(p.out, rec_kept(p.rec), pack_key(si.id, &bad));
    }
  } else if (sc.ent && keep) {  // list emission: the retained entry, in the part's own slot range
    unsigned bad = 0;
    A3D_DEV_CHECK(rec_kept(p.rec) < sc.kmax);
    __stcs(&sc.ent[p.at * static_cast<size_t>(sc.kmax) + rec_kept(p.rec)], pack_key(si.id, &bad));
    inexact |= bad;
  }
  p.rec += static_cast<unsigned long long>(keep) << (kCntBits * cat);
  if (first) {
    // the part's first raw entry: its key now, what it was counted as in the record word
    unsigned bad = 0;
    __stcs(&sc.first_key[p.at], pack_key(si.id, &bad));
    inexact |= bad;
    p.rec |= kRecEmitted | (keep ? kRecKept : 0ull) | (static_cast<unsigned long long>(keep ? cat : 3) << kRecCatShift);
  }
  if (si.st == 0) {
    p.d = d_this;
    return true;
  }
  if (!more) p.d = -1.0;
  return !more;
}

// A whole marking part by the warp: lane k evaluates sample k of the ray (accumulated distances from
// CasterView::dseq, as k_cast_gain does), the first OCCUPIED sample ends it.  Used on the
// anti-diagonal chain, where the latency of a part — not lane occupancy — is what counts.  Returns the
// hit distance (-1: none).  All lanes must call; `at` = the part's record.
template <int EVAL, bool DIRECT, class S>
__device__ __forceinline__ double part_by_warp(const MapView& m, const CasterView& c, const EvalView& e, S& sh,
                                               const ViewRec& vw, const int bv_mode, const uint32_t* __restrict__ tab,
                                               const SplitScratch& sc, const int view, const int r, const int s0,
                                               const size_t at, unsigned& inexact) {
  const unsigned lane = threadIdx.x & 31u;
  const int n = c.n_sections;
  const int n_s = c.seg_end[s0 * n + (n - 2)];  // samples before c_split_distances_[n-1]
  const double* const dseq = c.dseq + static_cast<size_t>(s0) * c.kmax;
  // every lane holds the ray's direction (the literal paths read it per thread)
  const Q4 q{vw.q[0], vw.q[1], vw.q[2], vw.q[3]};
  const D3 dir = rotate(q, D3{c.cam_dirs[3 * r], c.cam_dirs[3 * r + 1], c.cam_dirs[3 * r + 2]});
  sh.dir[0][threadIdx.x] = dir.x;
  sh.dir[1][threadIdx.x] = dir.y;
  sh.dir[2][threadIdx.x] = dir.z;
  const double k2 = 2.0 * static_cast<double>(m.vs_inv_f);
  const float dz[3] = {__double2float_rn(__dmul_rn(dir.x, k2)), __double2float_rn(__dmul_rn(dir.y, k2)),
                       __double2float_rn(__dmul_rn(dir.z, k2))};
  unsigned long long rec = 0;
  DirectOut out{};
  if (DIRECT) out = direct_begin(sc, view, at);
  double wsum = 0.0, first_val = 0.0, hit_d = -1.0;
  Ident carry{kNoIdent, 0, 0};  // last raw entry of the previous chunk of 32 samples
  Ident last{kNoIdent, 0, 0};
  for (int k0 = 0; k0 < n_s; k0 += 32) {
    const int k = k0 + static_cast<int>(lane);
    const bool in = k < n_s;
    const double d = in ? __ldg(dseq + k) : 0.0;
    Eval ev;
    ev.si.st = 1;
    ev.si.cs = 0;
    ev.si.id = Ident{0, 0, 0};
    ev.in_bv = false;
    if (in) ev = sample_eval(m, e, sh, vw, bv_mode, tab, d, dz, inexact);
    const unsigned occ = __ballot_sync(kFull, in && ev.si.st == 0);
    const int n_valid = occ ? __ffs(occ) : min(32, n_s - k0);  // the hit sample is part of the list
    const bool valid = static_cast<int>(lane) < n_valid;
    // previous raw entry: the lane below (every sample of an iterative ray is emitted)
    Ident prev;
    prev.g0 = __shfl_up_sync(kFull, ev.si.id.g0, 1);
    prev.g1 = __shfl_up_sync(kFull, ev.si.id.g1, 1);
    prev.g2 = __shfl_up_sync(kFull, ev.si.id.g2, 1);
    if (lane == 0) prev = carry;
    const bool keep = valid & ev.in_bv &
                      ((ev.si.id.g0 != prev.g0) | (ev.si.id.g1 != prev.g1) | (ev.si.id.g2 != prev.g2));
    double wval = 0.0;
    int cat = 3;
    if (valid) cat = entry_category<EVAL>(m, e, sh, vw, ev, keep, wval);
    const bool first = k == 0;
    if (DIRECT) {
      const unsigned km = __ballot_sync(kFull, keep);
      if (keep) {
        unsigned bad = 0;
        direct_put(out, rec_kept(rec) + __popc(km & ((1u << lane) - 1u)), pack_key(ev.si.id, &bad));
      }
    } else if (sc.ent) {  // list emission: retained entries in sample order
      const unsigned km = __ballot_sync(kFull, keep);
      if (keep) {
        unsigned bad = 0;
        A3D_DEV_CHECK(rec_kept(rec) + __popc(km & ((1u << lane) - 1u)) < sc.kmax);
        __stcs(&sc.ent[at * static_cast<size_t>(sc.kmax) + rec_kept(rec) + __popc(km & ((1u << lane) - 1u))],
               pack_key(ev.si.id, &bad));
        inexact |= bad;
      }
    }
    rec += static_cast<unsigned long long>(n_valid) << kCntSampShift;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
      rec += static_cast<unsigned long long>(__popc(__ballot_sync(kFull, keep && cat == cc))) << (kCntBits * cc);
    if (first) {
      unsigned bad = 0;
      __stcs(&sc.first_key[at], pack_key(ev.si.id, &bad));
      inexact |= bad;
    }
    // lane 0 of the first chunk describes the first entry to everybody
    if (k0 == 0) {
      const int keep0 = __shfl_sync(kFull, static_cast<int>(keep), 0), cat0 = __shfl_sync(kFull, cat, 0);
      rec |= kRecEmitted | (keep0 ? kRecKept : 0ull) | (static_cast<unsigned long long>(keep0 ? cat0 : 3) << kRecCatShift);
    }
    if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
      double w = (keep && !first) ? wval : 0.0;
      if (__any_sync(kFull, w != 0.0)) {  // (entries with a value of their own are rare)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(kFull, w, off);
        wsum += w;
      }
      if (k0 == 0) first_val = __shfl_sync(kFull, keep ? wval : 0.0, 0);
    }
    const int tail = n_valid - 1;
    last.g0 = __shfl_sync(kFull, ev.si.id.g0, tail);
    last.g1 = __shfl_sync(kFull, ev.si.id.g1, tail);
    last.g2 = __shfl_sync(kFull, ev.si.id.g2, tail);
    carry = last;
    if (occ) {
      hit_d = __shfl_sync(kFull, d, tail);
      break;
    }
  }
  if (lane == 0) {
    if (last.g0 != kNoIdent) {
      unsigned bad = 0;
      __stcs(&sc.last_key[at], pack_key(last, &bad));
      inexact |= bad;
    }
    __stcs(&sc.rec[at], rec);
    if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
      __stcs(&sc.wsum[at], wsum);
      __stcs(&sc.first_val[at], first_val);
    }
  }
  return hit_d;
}

template <int EVAL>
__device__ __forceinline__ void part_end(const SplitScratch& sc, const Part& p, unsigned& inexact) {
  if (p.prev.g0 != kNoIdent) {
    unsigned bad = 0;
    __stcs(&sc.last_key[p.at], pack_key(p.prev, &bad));
    inexact |= bad;
  }
  __stcs(&sc.rec[p.at], p.rec);  // (VoxelWeightEvaluator: wsum / first_val were written as they came)
}

// ---- k_split_views -----------------------------------------------------------------------------------
__global__ void k_split_views(const MapView m, const CasterView c, const Batch b, int seg0, int n_seg,
                              int view0, const SplitScratch sc) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const int seg = seg0 + s;
  const int v_begin = b.view_first[seg], v_end = b.view_first[seg + 1];
  for (int view = v_begin; view < v_end; ++view) {
    ViewRec vw;
    // camera pose = body pose ∘ mounting transform (camera_model.cpp:24-28)
    const D3 P0{b.pos[3 * view], b.pos[3 * view + 1], b.pos[3 * view + 2]};
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
    const double* o = b.seg_origin ? b.seg_origin + 3 * seg : b.pos + 3 * (v_end - 1);
    vw.org[0] = o[0];
    vw.org[1] = o[1];
    vw.org[2] = o[2];
    // margin of the certified index path for every sample this view can produce
    const double reach = fmax(fmax(fabs(vw.P0[0]), fabs(vw.P0[1])), fabs(vw.P0[2])) + c.ray_length;
    const double t_max = reach * static_cast<double>(m.vs_inv_f) * 1.001 + 2.0;
    float lim = __int_as_float(0x7FC00000);
    if (m.vps_shift >= 0 && t_max < 1048576.0) lim = 0.5f - 2.0f * (static_cast<float>(t_max * 9e-7) + m.fast_c0);
    vw.lim = lim;
    const double k = 2.0 * static_cast<double>(m.vs_inv_f);
#pragma unroll
    for (int a = 0; a < 3; ++a) vw.z0[a] = static_cast<float>(vw.P0[a] * k - 0.5);
    sc.views[view - view0] = vw;
    sc.view_flags[view - view0] = 0u;
    if (sc.view_seg) sc.view_seg[view - view0] = seg;
  }
}

template <class S>
__device__ __forceinline__ void cta_prologue(const MapView& m, const EvalView& e, S& sh, uint32_t* tab, uint32_t tab_n) {
  init_center_tables(m, &sh.ct);
  fill_vol_tab(m, tab, tab_n);
  __syncthreads();
  if (threadIdx.x < 32) init_bv_grid(m, sh.ct, e, &sh.bvg);
  __syncthreads();
}

template <class S>
__device__ __forceinline__ void load_view(S& sh, const SplitScratch& sc, int view_local, unsigned warp, unsigned lane) {
  __syncwarp();
  static_assert(sizeof(ViewRec) % 8 == 0, "ViewRec is copied in 8-byte words");
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(sc.views + view_local);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sh.vw[warp]);
  if (lane < sizeof(ViewRec) / 8) dst[lane] = __ldg(src + lane);
  __syncwarp();
}

// ---- k_split_mark ------------------------------------------------------------------------------------
// One warp per view, ray table in global memory (any camera size).  Dynamic shared memory: per warp
// diag_max hit distances + diag_max list items, then the volume offset tables.
template <int EVAL, bool DIRECT>
__global__ void __launch_bounds__(kSplitThreads, A3D_SPLIT_MARK_BLOCKS)
    k_split_mark(const MapView m, const CasterView c, const EvalView e, int n_views, const SplitScratch sc) {
  constexpr bool SMEM_TABLE = false;
  constexpr int kWarps = kSplitWarps;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int n_rays = c.n_rays, n = c.n_sections, n_diag = c.res_x + c.res_y - 1;
  __shared__ CtaShared sh;
  extern __shared__ double s_dyn[];
  double* const s_hit = s_dyn + static_cast<size_t>(warp) * c.diag_max;
  unsigned* const s_list_all = reinterpret_cast<unsigned*>(s_dyn + static_cast<size_t>(kWarps) * c.diag_max);
  unsigned* const s_list = s_list_all + static_cast<size_t>(warp) * c.diag_max;
  uint32_t* const s_tab = s_list_all + static_cast<size_t>(kWarps) * c.diag_max;
  const uint32_t tab_n = vol_tab_entries(m);
  uint32_t* const s_table = nullptr;
  const uint32_t* const tab = tab_n ? s_tab : nullptr;
  cta_prologue(m, e, sh, s_tab, tab_n);
  const int bv_mode = sh.bvg.mode;

  for (int view = blockIdx.x * kWarps + warp; view < n_views; view += gridDim.x * kWarps) {
    load_view(sh, sc, view, warp, lane);
    const ViewRec& vw = sh.vw[warp];
    uint32_t* const g_table = sc.table + static_cast<size_t>(view) * sc.stride;
    uint8_t* const start_s = sc.start_s + static_cast<size_t>(view) * sc.stride;
    const size_t rec0 = static_cast<size_t>(view) * 2 * sc.stride;  // marking parts: part 0
    for (int r = lane; r < n_rays; r += 32) {
      if (SMEM_TABLE) s_table[r] = 0u; else g_table[r] = 0u;
      start_s[r] = 0xFF;
      __stcs(&sc.rec[rec0 + r], 0ull);
    }
    if (!SMEM_TABLE) __threadfence_block();
    __syncwarp();
    // table access: shared memory, or L2 (atomics are performed there, reads bypass L1)
    auto t_read = [&](int r) -> uint32_t { return SMEM_TABLE ? s_table[r] : __ldcg(&g_table[r]); };
    auto t_max = [&](int r, unsigned v) {
      A3D_DEV_CHECK(r >= 0 && r < n_rays);
      if (SMEM_TABLE) atomicMax(&s_table[r], v); else atomicMax(&g_table[r], v);
    };
    auto t_fence = [&]() {
      if (!SMEM_TABLE) __threadfence_block();
      __syncwarp();
    };
    unsigned inexact = 0;
    uint32_t* const dlist = sc.dlist + static_cast<size_t>(view) * sc.stride;
    int n_def = 0;
    unsigned any_def = 0;

    for (int diag = 0; diag < n_diag; ++diag) {
      // The rays of this diagonal that start before the last section, in scan order.  Those that start
      // in section n-2 (most of them) mark only their own 2x2 square, with n-1 or — if they hit
      // something — -1: neither value makes a marking ray, so nothing on the chain depends on their
      // outcome.  They get the provisional mark n-1 now and are cast later, all at once, by
      // k_split_mark2, which raises the marks of the rays that hit to -1 (same writer, larger value).
      int n_mark = 0;
      const int i_lo = max(0, diag - (c.res_y - 1)), i_hi = min(c.res_x - 1, diag);
      for (int base = i_lo; base <= i_hi; base += 32) {
        const int i = base + static_cast<int>(lane);
        int s = -1, r = 0;
        if (i <= i_hi) {
          r = i * c.res_y + (diag - i);
          s = static_cast<int>(static_cast<int8_t>(t_read(r) & 0xFFu));
        }
        const bool deferred = s == n - 2;
        const bool marking = s >= 0 && s < n - 2;
        const unsigned mm = __ballot_sync(kFull, marking), dm = __ballot_sync(kFull, deferred);
        A3D_DEV_CHECK(n_mark + __popc(mm) <= c.diag_max && n_def + __popc(dm) <= n_rays);
        if (marking) s_list[n_mark + __popc(mm & lt_mask)] = (static_cast<unsigned>(r) << 8) | s;
        n_mark += __popc(mm);
        if (deferred) {
          dlist[n_def + __popc(dm & lt_mask)] = static_cast<uint32_t>(r);
          start_s[r] = static_cast<uint8_t>(n - 2);
          const int ri = i, rj = diag - i;
          const unsigned ent = ((static_cast<unsigned>(r) + 1u) << 8) | static_cast<uint8_t>(n - 1);
          const bool in_x = ri + 1 < c.res_x, in_y = rj + 1 < c.res_y;
          t_max(r, ent);
          if (in_y) t_max(r + 1, ent);
          if (in_x) t_max(r + c.res_y, ent);
          if (in_x && in_y) t_max(r + c.res_y + 1, ent);
        }
        n_def += __popc(dm);
        any_def |= dm;
      }
      __syncwarp();
      if (n_mark == 0) {
        if (any_def) t_fence();
        any_def = 0;
        continue;
      }
      any_def = 0;
      // cast them, one after the other, each by the whole warp
      for (int qi = 0; qi < n_mark; ++qi) {
        const unsigned it = s_list[qi];
        const int r = static_cast<int>(it >> 8), ms = static_cast<int>(it & 0xFF);
        const double hd = part_by_warp<EVAL, DIRECT>(m, c, e, sh, vw, bv_mode, tab, sc, view, r, ms, rec0 + r, inexact);
        if (lane == 0) {
          s_hit[qi] = hd;
          start_s[r] = static_cast<uint8_t>(ms);  // read by k_split_leaf
        }
      }
      __syncwarp();
      // markNeighboringRays of every marking ray of the diagonal.  An entry keeps the value of its
      // highest writer (= last writer in scan order), so the marks commute: fire-and-forget
      // atomicMax on (writer << 8 | value), one ray per lane.
      for (int base = 0; base < n_mark; base += 32) {
        const int qi = base + static_cast<int>(lane);
        const bool valid = qi < n_mark;
        const unsigned it = valid ? s_list[qi] : 0u;
        const int rr = static_cast<int>(it >> 8), ss = static_cast<int>(it & 0xFF);
        const double hd = valid ? s_hit[qi] : -1.0;
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
          t_max(rr, ent);
          if (in_y) t_max(rr + 1, ent);
          if (in_x) t_max(rr + c.res_y, ent);
          if (in_x && in_y) t_max(rr + c.res_y + 1, ent);
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
            t_max((b_ri + dx) * c.res_y + b_rj + dy, b_wr | static_cast<uint8_t>(val));
          }
        }
      }
      t_fence();
    }
    if (SMEM_TABLE) {
      __syncwarp();
      for (int r = lane; r < n_rays; r += 32) g_table[r] = s_table[r];
    }
    if (lane == 0) sc.dcount[view] = static_cast<unsigned>(n_def);
    if (__any_sync(kFull, inexact != 0) && lane == 0) sc.view_flags[view] = 1u;
  }
}

// ---- k_split_mark_cta --------------------------------------------------------------------------------
// The same walk with one CTA (4 warps) per view: the marking parts of one anti-diagonal are cast by
// different warps at the same time, so that what remains per anti-diagonal is one part's latency.
// SMEM_TABLE: the ray table lives in shared memory (cameras whose table fits) — every step of the chain
// is a shared-memory access instead of an L2 round trip — and is copied out at the end; otherwise it
// stays in global memory (L2 atomics, reads that bypass L1).  Dynamic shared memory: diag_max hit
// distances, diag_max list items, the volume offset tables, (SMEM_TABLE) n_rays table entries.
template <int EVAL, bool SMEM_TABLE, bool DIRECT>
__global__ void __launch_bounds__(kSplitThreads, SMEM_TABLE ? 6 : 7)
    k_split_mark_cta(const MapView m, const CasterView c, const EvalView e, int n_views, const SplitScratch sc) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int n_rays = c.n_rays, n = c.n_sections, n_diag = c.res_x + c.res_y - 1;
  __shared__ CtaShared sh;
  __shared__ int s_n_mark, s_n_def;
  __shared__ unsigned s_inexact;
  extern __shared__ double s_dyn[];
  double* const s_hit = s_dyn;
  unsigned* const s_list = reinterpret_cast<unsigned*>(s_dyn + c.diag_max);
  uint32_t* const s_tab = s_list + c.diag_max;
  const uint32_t tab_n = vol_tab_entries(m);
  const uint32_t* const tab = tab_n ? s_tab : nullptr;
  uint32_t* const s_table = s_tab + tab_n;
  cta_prologue(m, e, sh, s_tab, tab_n);
  const int bv_mode = sh.bvg.mode;

  for (int view = blockIdx.x; view < n_views; view += gridDim.x) {
    __syncthreads();
    load_view(sh, sc, view, warp, lane);  // (every warp keeps its own copy)
    const ViewRec& vw = sh.vw[warp];
    uint8_t* const start_s = sc.start_s + static_cast<size_t>(view) * sc.stride;
    uint32_t* const dlist = sc.dlist + static_cast<size_t>(view) * sc.stride;
    const size_t rec0 = static_cast<size_t>(view) * 2 * sc.stride;  // marking parts: part 0
    uint32_t* const g_table = sc.table + static_cast<size_t>(view) * sc.stride;
    auto t_read = [&](int r) -> uint32_t { return SMEM_TABLE ? s_table[r] : __ldcg(&g_table[r]); };
    auto t_max = [&](int r, unsigned v) {
      A3D_DEV_CHECK(r >= 0 && r < n_rays);
      if (SMEM_TABLE) atomicMax(&s_table[r], v); else atomicMax(&g_table[r], v);
    };
    auto t_sync = [&]() {  // marks of every warp visible to every warp
      if (!SMEM_TABLE) __threadfence_block();
      __syncthreads();
    };
    for (int r = threadIdx.x; r < n_rays; r += kSplitThreads) {
      if (SMEM_TABLE) s_table[r] = 0u; else g_table[r] = 0u;
      start_s[r] = 0xFF;
      __stcs(&sc.rec[rec0 + r], 0ull);
    }
    if (threadIdx.x == 0) {
      s_n_def = 0;
      s_inexact = 0;
    }
    unsigned inexact = 0;

    for (int diag = 0; diag < n_diag; ++diag) {
      if (threadIdx.x == 0) s_n_mark = 0;
      t_sync();  // table final up to this diagonal; counters reset
      // The rays of this diagonal that start before the last section (see k_split_mark): section n-2
      // starts get their provisional mark and go to the list of k_split_mark2, earlier starts are cast
      // here.  Neither list needs an order (marks commute, records are per ray).
      const int i_lo = max(0, diag - (c.res_y - 1)), i_hi = min(c.res_x - 1, diag);
      for (int base = i_lo + 32 * static_cast<int>(warp); base <= i_hi; base += kSplitThreads) {
        const int i = base + static_cast<int>(lane);
        int s = -1, r = 0;
        if (i <= i_hi) {
          r = i * c.res_y + (diag - i);
          s = static_cast<int>(static_cast<int8_t>(t_read(r) & 0xFFu));
        }
        const bool deferred = s == n - 2;
        const bool marking = s >= 0 && s < n - 2;
        const unsigned mm = __ballot_sync(kFull, marking), dm = __ballot_sync(kFull, deferred);
        int m_at = 0, d_at = 0;
        if (lane == 0) {
          if (mm) m_at = atomicAdd(&s_n_mark, __popc(mm));
          if (dm) d_at = atomicAdd(&s_n_def, __popc(dm));
        }
        m_at = __shfl_sync(kFull, m_at, 0);
        d_at = __shfl_sync(kFull, d_at, 0);
        A3D_DEV_CHECK(m_at + __popc(mm) <= c.diag_max && d_at + __popc(dm) <= n_rays);
        if (marking) s_list[m_at + __popc(mm & lt_mask)] = (static_cast<unsigned>(r) << 8) | s;
        if (deferred) {
          dlist[d_at + __popc(dm & lt_mask)] = static_cast<uint32_t>(r);
          start_s[r] = static_cast<uint8_t>(n - 2);
          const int ri = i, rj = diag - i;
          const unsigned ent = ((static_cast<unsigned>(r) + 1u) << 8) | static_cast<uint8_t>(n - 1);
          const bool in_x = ri + 1 < c.res_x, in_y = rj + 1 < c.res_y;
          t_max(r, ent);
          if (in_y) t_max(r + 1, ent);
          if (in_x) t_max(r + c.res_y, ent);
          if (in_x && in_y) t_max(r + c.res_y + 1, ent);
        }
      }
      __syncthreads();
      const int n_mark = s_n_mark;
      if (n_mark == 0) continue;
      // cast them: each warp takes every fourth one
      for (int qi = static_cast<int>(warp); qi < n_mark; qi += kSplitWarps) {
        const unsigned it = s_list[qi];
        const int r = static_cast<int>(it >> 8), ms = static_cast<int>(it & 0xFF);
        const double hd = part_by_warp<EVAL, DIRECT>(m, c, e, sh, vw, bv_mode, tab, sc, view, r, ms, rec0 + r, inexact);
        if (lane == 0) {
          s_hit[qi] = hd;
          start_s[r] = static_cast<uint8_t>(ms);  // read by k_split_leaf
        }
      }
      __syncthreads();
      // markNeighboringRays of every marking ray of the diagonal (see k_split_mark)
      for (int base = 32 * static_cast<int>(warp); base < n_mark; base += kSplitThreads) {
        const int qi = base + static_cast<int>(lane);
        const bool valid = qi < n_mark;
        const unsigned it = valid ? s_list[qi] : 0u;
        const int rr = static_cast<int>(it >> 8), ss = static_cast<int>(it & 0xFF);
        const double hd = valid ? s_hit[qi] : -1.0;
        const bool hit = hd >= 0.0;
        int seg_h = ss;
        if (hit) {
          while (!(hd < c.split[seg_h + 1])) ++seg_h;
        }
        const int t_last = hit ? seg_h : n - 2;
        const int w0 = 1 << (n - 1 - ss);
        const int ri = rr / c.res_y, rj = rr - ri * c.res_y;
        const unsigned wr = (static_cast<unsigned>(rr) + 1u) << 8;
        unsigned big = __ballot_sync(kFull, valid);
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
            t_max((b_ri + dx) * c.res_y + b_rj + dy, b_wr | static_cast<uint8_t>(val));
          }
        }
      }
    }
    __syncthreads();
    if (SMEM_TABLE)
      for (int r = threadIdx.x; r < n_rays; r += kSplitThreads) g_table[r] = s_table[r];
    if (__any_sync(kFull, inexact != 0) && lane == 0) atomicOr(&s_inexact, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
      sc.dcount[view] = static_cast<unsigned>(s_n_def);
      sc.view_flags[view] = s_inexact;
    }
  }
}

// ---- k_split_leaf / k_split_mark2 --------------------------------------------------------------------
// DEFERRED = false: unit = 32 consecutive rays of one view, cast through the last section (whole ray).
// DEFERRED = true (k_split_mark2): unit = 32 entries of the view's list of rays that start in section
// n-2; cast through that section; a ray that hits raises its provisional marks to -1.
template <int CASTER, int EVAL, bool DEFERRED, bool DIRECT>
__global__ void __launch_bounds__(kSplitThreads, leaf_blocks(CASTER, EVAL))
    k_split_leaf(const MapView m, const CasterView c, const EvalView e, int n_views, const SplitScratch sc) {
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int n_rays = c.n_rays, n = c.n_sections;
  __shared__ CtaShared sh;
  extern __shared__ double s_dyn[];
  uint32_t* const s_tab = reinterpret_cast<uint32_t*>(s_dyn);
  const uint32_t tab_n = vol_tab_entries(m);
  const uint32_t* const tab = tab_n ? s_tab : nullptr;
  cta_prologue(m, e, sh, s_tab, tab_n);
  const int bv_mode = sh.bvg.mode;
  const double step = c.ray_step;
  const int upv = (n_rays + 31) / 32;
  const long long n_units = static_cast<long long>(n_views) * upv;
  const size_t part_off = (ITER && !DEFERRED) ? sc.stride : 0;  // leaf parts: part 1 of 2 (iterative), part 0 of 1

  int cur_view = -1;
  for (long long unit = static_cast<long long>(blockIdx.x) * kSplitWarps + warp; unit < n_units;
       unit += static_cast<long long>(gridDim.x) * kSplitWarps) {
    const int view = static_cast<int>(unit / upv);
    int r = static_cast<int>(unit - static_cast<long long>(view) * upv) * 32 + static_cast<int>(lane);
    bool active;
    if (DEFERRED) {
      const int n_def = static_cast<int>(__ldg(&sc.dcount[view]));
      if (r - static_cast<int>(lane) >= n_def) continue;  // (warp-uniform)
      active = r < n_def;
      r = active ? static_cast<int>(sc.dlist[static_cast<size_t>(view) * sc.stride + r]) : 0;
    } else {
      active = r < n_rays;
    }
    if (view != cur_view) {
      load_view(sh, sc, view, warp, lane);
      cur_view = view;
    }
    const ViewRec& vw = sh.vw[warp];
    const size_t at = static_cast<size_t>(view) * sc.parts * sc.stride + part_off + r;
    double d0 = 0.0;
    if (ITER && !DEFERRED && active) {
      const size_t tr = static_cast<size_t>(view) * sc.stride + r;
      active = static_cast<int>(static_cast<int8_t>(__ldcs(&sc.table[tr]) & 0xFFu)) == n - 1;
      const int s0 = sc.start_s[tr];
      d0 = c.leaf_d0[s0 == 0xFF ? n - 1 : s0];  // accumulated distance at which a ray started in s0 enters section n-1
      active &= d0 < c.ray_length;
      if (!active) __stcs(&sc.rec[at], 0ull);
    }
    unsigned inexact = 0;
    Part p;
    if (active) {
      part_begin(p, sh, m, c, vw, r, DEFERRED ? c.split[n - 2] : d0, DEFERRED ? c.split[n - 1] : c.ray_length, at);
      if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
        __stcs(&sc.wsum[at], 0.0);
        __stcs(&sc.first_val[at], 0.0);
      }
      if (DIRECT) p.out = direct_begin(sc, view, at);
    }
    while (__any_sync(kFull, active)) {
      if (active && part_step<CASTER, EVAL, DIRECT>(m, e, sh, vw, bv_mode, step, tab, sc, p, inexact)) {
        part_end<EVAL>(sc, p, inexact);
        if (DEFERRED && p.d >= 0.0) {
          // hit: the 2x2 square's marks become -1 where this ray is still the highest writer
          uint32_t* const table = sc.table + static_cast<size_t>(view) * sc.stride;
          const int ri = r / c.res_y, rj = r - ri * c.res_y;
          const unsigned ent = ((static_cast<unsigned>(r) + 1u) << 8) | 0xFFu;
          const bool in_x = ri + 1 < c.res_x, in_y = rj + 1 < c.res_y;
          atomicMax(&table[r], ent);
          if (in_y) atomicMax(&table[r + 1], ent);
          if (in_x) atomicMax(&table[r + c.res_y], ent);
          if (in_x && in_y) atomicMax(&table[r + c.res_y + 1], ent);
        }
        active = false;
      }
    }
    if (__any_sync(kFull, inexact != 0) && lane == 0) atomicOr(&sc.view_flags[view], 1u);
  }
}

// ---- k_split_fix -------------------------------------------------------------------------------------
// One warp per segment: adjacent-unique across ray-part boundaries in scan order (ray by ray, the
// marking part before the leaf part), counters summed, gain folded.
template <int CASTER, int EVAL>
__global__ void __launch_bounds__(kSplitThreads)
    k_split_fix(const EvalView e, const Batch b, int seg0, int n_seg, int view0, int n_rays, const SplitScratch sc,
                unsigned int* __restrict__ inexact_flags) {
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int s = blockIdx.x * kSplitWarps + static_cast<int>(threadIdx.x >> 5);
  if (s >= n_seg) return;
  const int seg = seg0 + s;
  const int v_begin = b.view_first[seg], v_end = b.view_first[seg + 1];
  unsigned c0 = 0, c1 = 0, c2 = 0, ns = 0;  // per-lane sums (a segment with 2^32 samples is flagged below)
  unsigned long long ns_wide = 0;
  double wsum = 0.0;
  unsigned inexact = 0;
  uint64_t carry = kNoKey;  // last raw entry so far in this segment (warp-uniform)
  unsigned emit_pos = 0;    // list emission: entries of the segment's list placed so far (warp-uniform)
  const int per_it = ITER ? 16 : 32;
  for (int view = v_begin; view < v_end; ++view) {
    const int vl = view - view0;
    if (lane == 0 && sc.view_flags[vl]) inexact = 1;
    const size_t rec0 = static_cast<size_t>(vl) * sc.parts * sc.stride;
    constexpr int kAhead = 8;  // chunks whose records are requested before the first is looked at
    for (int base0 = 0; base0 < n_rays; base0 += kAhead * per_it) {
      unsigned long long recs[kAhead];
      uint64_t fks[kAhead], lks[kAhead];
      size_t ats[kAhead];
#pragma unroll
      for (int k = 0; k < kAhead; ++k) {
        // iterative: lane L looks at (ray base + L/2, part L&1); scan order == lane order
        const int base = base0 + k * per_it;
        const int rr = ITER ? base + static_cast<int>(lane >> 1) : base + static_cast<int>(lane);
        ats[k] = rec0 + (ITER ? static_cast<size_t>(lane & 1) * sc.stride : 0) + rr;
        recs[k] = 0;
        fks[k] = kNoKey;
        lks[k] = kNoKey;
        if (rr < n_rays) {
          // (the keys of a part that emitted nothing are never written: read, not used)
          recs[k] = __ldcs(&sc.rec[ats[k]]);
          fks[k] = __ldcs(&sc.first_key[ats[k]]);
          lks[k] = __ldcs(&sc.last_key[ats[k]]);
        }
      }
#pragma unroll
      for (int k = 0; k < kAhead; ++k) {
        const unsigned long long rec = recs[k];
        const uint64_t fk = fks[k], lk = lks[k];
        const bool emitted = (rec & kRecEmitted) != 0;
        const unsigned em = __ballot_sync(kFull, emitted);
        const unsigned below = em & lt_mask;
        uint64_t prev = __shfl_sync(kFull, lk, below ? 31 - __clz(below) : 0);
        if (!below) prev = carry;
        const bool kept = (rec & kRecKept) != 0;
        const bool drop = emitted && kept && fk == prev;  // std::unique drops the part's first entry
        const int cat = static_cast<int>((rec >> kRecCatShift) & 3);
        c0 += static_cast<unsigned>(rec & kCntMask) - ((drop && cat == 0) ? 1u : 0u);
        c1 += static_cast<unsigned>((rec >> kCntBits) & kCntMask) - ((drop && cat == 1) ? 1u : 0u);
        c2 += static_cast<unsigned>((rec >> (2 * kCntBits)) & kCntMask) - ((drop && cat == 2) ? 1u : 0u);
        ns += static_cast<unsigned>((rec >> kCntSampShift) & kCntMask);
        if (EVAL == A3D_EVAL_VOXEL_WEIGHT && ((rec >> kCntSampShift) & kCntMask) != 0) {
          wsum += sc.wsum[ats[k]];
          if (emitted && kept && !drop) wsum += sc.first_val[ats[k]];
        }
        if (sc.off) {
          // where this part's retained entries go in the segment's ordered list
          const unsigned n_out = static_cast<unsigned>(rec_kept(rec)) - (drop ? 1u : 0u);
          unsigned incl = n_out;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, incl, o);
            if (static_cast<int>(lane) >= o) incl += t;
          }
          if (((rec >> kCntSampShift) & kCntMask) != 0 || emitted)
            sc.off[ats[k]] = (emit_pos + incl - n_out) | (drop ? 0x80000000u : 0u);
          emit_pos += __shfl_sync(kFull, incl, 31);
        }
        if (em) carry = __shfl_sync(kFull, lk, 31 - __clz(em));
      }
    }
    ns_wide += ns;
    ns = 0;
  }
  const unsigned long long t0 = __reduce_add_sync(kFull, c0), t1 = __reduce_add_sync(kFull, c1),
                           t2 = __reduce_add_sync(kFull, c2);
  unsigned long long tn = ns_wide;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) tn += __shfl_xor_sync(kFull, tn, off);
  if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(kFull, wsum, off);
  }
  inexact = __any_sync(kFull, inexact != 0);
  if (lane == 0) {
    // (per-lane 32-bit counters: a segment whose views hold 2^32 entries per lane is out of reach of
    // the record format anyway; __reduce_add_sync wraps at 2^32 in total, guard it)
    if (tn >= (1ull << 32)) inexact = 1;
    double g = 0.0;
    if (EVAL == A3D_EVAL_VOXEL_TYPE) {
      // VoxelTypeEvaluator's switch without breaks (voxel_type_evaluator.cpp:69-85)
      const double n_unk = static_cast<double>(t2);
      const double n_free = static_cast<double>(t1) + n_unk;
      const double n_occ = static_cast<double>(t0) + n_free;
      g += n_unk * e.gains[0];
      g += n_free * e.gains[2];
      g += n_occ * e.gains[1];
    } else if (EVAL == A3D_EVAL_VOXEL_WEIGHT) {
      g = wsum;  // entries with a value of their own, summed in a fixed order; then the counted ones
      g += static_cast<double>(t0) * e.frontier_voxel_weight;
      g += static_cast<double>(t2) * e.new_voxel_weight;
    } else {
      g = static_cast<double>(t0);
    }
    b.gains[seg] = g;
    if (b.n_visible) b.n_visible[seg] = t0 + t1 + t2;
    if (b.n_samples) b.n_samples[seg] = tn;
    if (inexact) atomicOr(&inexact_flags[seg >> 5], 1u << (seg & 31));
  }
}

// ---- k_split_emit ------------------------------------------------------------------------------------
// List emission, last step: every ray part's retained entries move from its slot range to their place
// in the segment's ordered list (offsets from k_split_fix), as packed keys or as the centre doubles
// the reference holds.  One lane per part, 32 consecutive parts in scan order per warp: neighbouring
// lanes write neighbouring ranges.
__global__ void __launch_bounds__(kSplitThreads)
    k_split_emit(const MapView m, const Batch b, int n_views, int n_rays, const SplitScratch sc) {
  __shared__ CenterTables ct;
  init_center_tables(m, &ct);
  __syncthreads();
  const long long n_parts = static_cast<long long>(n_views) * sc.parts * n_rays;
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < n_parts;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    // scan order within a view: ray by ray, marking part before leaf part
    const int view = static_cast<int>(q / (static_cast<long long>(sc.parts) * n_rays));
    const int rest = static_cast<int>(q - static_cast<long long>(view) * sc.parts * n_rays);
    const int ray = rest / sc.parts, part = rest - ray * sc.parts;
    const size_t at = (static_cast<size_t>(view) * sc.parts + part) * sc.stride + ray;
    const unsigned long long rec = sc.rec[at];
    const int kept = rec_kept(rec);
    if (kept == 0) continue;
    const uint32_t o = sc.off[at];
    const int skip = static_cast<int>(o >> 31);
    const long long first_idx = static_cast<long long>(o & 0x7FFFFFFFu);
    const int seg = sc.view_seg[view];
    const long long base = b.emit_offset ? b.emit_offset[seg] : 0;
    const long long cap = b.emit_offset ? b.emit_offset[seg + 1] - base : b.cap;
    const unsigned long long* src = sc.ent + at * static_cast<size_t>(sc.kmax);
    for (int k = skip; k < kept; ++k) {
      const long long idx = first_idx + (k - skip);
      if (idx >= cap) break;
      A3D_DEV_CHECK(k < sc.kmax && idx >= 0);
      const unsigned long long key = __ldcs(src + k);
      if (b.keys_out) {
        b.keys_out[base + idx] = key;
      } else {
        const D3 cc = key_centre(m, ct, key);
        double* dst = b.centres_out + 3 * (base + idx);
        dst[0] = cc.x;
        dst[1] = cc.y;
        dst[2] = cc.z;
      }
    }
  }
}

template <int CASTER, int EVAL, bool DIRECT>
int launch_split_t(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                   int seg0, int n_seg, int view0, int n_views, const SplitScratch& sc, unsigned int* inexact,
                   int sm_count, cudaEvent_t* evs) {
  int n_ev = 0;
  auto stamp = [&]() {
    if (evs) cudaEventRecord(evs[n_ev++], st);
  };
  stamp();
  k_split_views<<<(n_seg + 127) / 128, 128, 0, st>>>(m, c, b, seg0, n_seg, view0, sc);
  stamp();
  const size_t tab_bytes = static_cast<size_t>(vol_tab_entries(m)) * 4;
  if (CASTER == A3D_CASTER_ITERATIVE && c.n_sections > 1) {
    // the ray table in shared memory when it fits next to the other tables (at most 8 views per SM)
    const size_t smem_1 = static_cast<size_t>(c.diag_max) * 12 + tab_bytes + static_cast<size_t>(c.n_rays) * 4;
    // ... and the batch is small enough for latency to matter more than the number of resident views
    // (cfg2: one view 0.89 -> 0.37 ms, 148 views 1.23 -> 0.69 ms; beyond ~2000 views the one-warp-per-view
    // kernel wins: 4096 views 1.7 vs 2.6 ms)
    if (smem_1 <= 40 * 1024 && n_views <= sm_count * 12) {
      k_split_mark_cta<EVAL, true, DIRECT><<<std::max(1, std::min(n_views, sm_count * 8)), kSplitThreads, smem_1, st>>>(m, c, e, n_views, sc);
    } else if (smem_1 > 40 * 1024 && n_views <= sm_count * 7) {
      // table too large for shared memory, batch small enough for one resident CTA per view: the
      // CTA-per-view walk on the global table (cfg3, 1024 views: 5.8 ms with one warp per view)
      const size_t smem_g = static_cast<size_t>(c.diag_max) * 12 + tab_bytes;
      k_split_mark_cta<EVAL, false, DIRECT><<<std::max(1, n_views), kSplitThreads, smem_g, st>>>(m, c, e, n_views, sc);
    } else {
      const size_t smem = static_cast<size_t>(kSplitWarps) * c.diag_max * 12 + tab_bytes;
      const int grid =
          std::max(1, std::min((n_views + kSplitWarps - 1) / kSplitWarps, sm_count * A3D_SPLIT_MARK_BLOCKS));
      k_split_mark<EVAL, DIRECT><<<grid, kSplitThreads, smem, st>>>(m, c, e, n_views, sc);
    }
  } else if (CASTER == A3D_CASTER_ITERATIVE) {
    // a single section: no marking parts; every ray is cast whole by k_split_leaf
    cudaMemsetAsync(sc.table, 0, static_cast<size_t>(n_views) * sc.stride * 4, st);
    cudaMemsetAsync(sc.start_s, 0xFF, static_cast<size_t>(n_views) * sc.stride, st);
    cudaMemsetAsync(sc.rec, 0, static_cast<size_t>(n_views) * 2 * sc.stride * 8, st);
    cudaMemsetAsync(sc.dcount, 0, static_cast<size_t>(n_views) * 4, st);
  }
  if (CASTER == A3D_CASTER_ITERATIVE) stamp();
  {
    const long long units = static_cast<long long>(n_views) * ((c.n_rays + 31) / 32);
    const int per_sm = leaf_blocks(CASTER, EVAL);
    const int grid = static_cast<int>(std::max(1ll, std::min((units + kSplitWarps - 1) / kSplitWarps,
                                                             static_cast<long long>(sm_count) * per_sm * 4)));
    if (CASTER == A3D_CASTER_ITERATIVE && c.n_sections > 1)
      k_split_leaf<CASTER, EVAL, true, DIRECT><<<grid, kSplitThreads, tab_bytes, st>>>(m, c, e, n_views, sc);
    if (CASTER == A3D_CASTER_ITERATIVE) stamp();
    k_split_leaf<CASTER, EVAL, false, DIRECT><<<grid, kSplitThreads, tab_bytes, st>>>(m, c, e, n_views, sc);
    stamp();
  }
  if constexpr (!DIRECT) {  // (DIRECT: counters, gains and the parts' offsets are the first pass's)
    k_split_fix<CASTER, EVAL><<<(n_seg + kSplitWarps - 1) / kSplitWarps, kSplitThreads, 0, st>>>(
        e, b, seg0, n_seg, view0, c.n_rays, sc, inexact);
    stamp();
    if (sc.ent) {
      const long long n_parts = static_cast<long long>(n_views) * sc.parts * c.n_rays;
      const int grid = static_cast<int>(std::max(1ll, std::min((n_parts + kSplitThreads - 1) / kSplitThreads,
                                                               static_cast<long long>(sm_count) * 16)));
      k_split_emit<<<grid, kSplitThreads, 0, st>>>(m, b, n_views, c.n_rays, sc);
    }
  }
  return n_ev;
}

template <int CASTER, bool DIRECT>
int launch_split_c(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                   int seg0, int n_seg, int view0, int n_views, const SplitScratch& sc, unsigned int* inexact,
                   int sm_count, cudaEvent_t* evs) {
  switch (e.kind) {
    case A3D_EVAL_VOXEL_TYPE:
      return launch_split_t<CASTER, A3D_EVAL_VOXEL_TYPE, DIRECT>(st, m, c, e, b, seg0, n_seg, view0, n_views, sc, inexact, sm_count, evs);
    case A3D_EVAL_NAIVE:
      return launch_split_t<CASTER, A3D_EVAL_NAIVE, DIRECT>(st, m, c, e, b, seg0, n_seg, view0, n_views, sc, inexact, sm_count, evs);
    case A3D_EVAL_VOXEL_WEIGHT:
      return launch_split_t<CASTER, A3D_EVAL_VOXEL_WEIGHT, DIRECT>(st, m, c, e, b, seg0, n_seg, view0, n_views, sc, inexact, sm_count, evs);
    default:
      return launch_split_t<CASTER, A3D_EVAL_FRONTIER, DIRECT>(st, m, c, e, b, seg0, n_seg, view0, n_views, sc, inexact, sm_count, evs);
  }
}

// bytes of scratch per view and where everything lies.  mode 0: gains / counts; 1: list emission through
// per-part key slots; 2: gains / counts + the parts' offsets within their segment's list (first pass of a
// stored-list call; the second pass — DIRECT — uses the same layout)
size_t split_bytes_per_view(const CasterView& c, int eval_kind, size_t* stride_out, int mode) {
  const size_t stride = (static_cast<size_t>(c.n_rays) + 31) / 32 * 32;
  if (stride_out) *stride_out = stride;
  const size_t parts = c.kind == A3D_CASTER_ITERATIVE ? 2 : 1;
  size_t per_view = parts * stride * (8 + 8 + 8) + sizeof(ViewRec) + 8;
  if (mode == 1) per_view += parts * stride * static_cast<size_t>(c.kmax) * 8;
  if (mode != 0) per_view += parts * stride * 4 + 8;
  if (eval_kind == A3D_EVAL_VOXEL_WEIGHT) per_view += parts * stride * 16;
  if (c.kind == A3D_CASTER_ITERATIVE) per_view += stride * 9 + 8;
  return per_view;
}

SplitScratch split_layout(const CasterView& c, int eval_kind, int mode, int n_views, void* scratch) {
  SplitScratch sc{};
  size_t stride = 0;
  split_bytes_per_view(c, eval_kind, &stride, mode);
  sc.stride = stride;
  sc.parts = c.kind == A3D_CASTER_ITERATIVE ? 2 : 1;
  uint8_t* p = static_cast<uint8_t*>(scratch);
  const size_t nv = static_cast<size_t>(n_views), slots = nv * sc.parts * stride;
  sc.views = reinterpret_cast<ViewRec*>(p);
  p += nv * sizeof(ViewRec);
  sc.first_key = reinterpret_cast<unsigned long long*>(p);
  p += slots * 8;
  sc.last_key = reinterpret_cast<unsigned long long*>(p);
  p += slots * 8;
  sc.rec = reinterpret_cast<unsigned long long*>(p);
  p += slots * 8;
  if (eval_kind == A3D_EVAL_VOXEL_WEIGHT) {
    sc.wsum = reinterpret_cast<double*>(p);
    p += slots * 8;
    sc.first_val = reinterpret_cast<double*>(p);
    p += slots * 8;
  }
  if (mode == 1) {
    sc.kmax = c.kmax;
    sc.ent = reinterpret_cast<unsigned long long*>(p);
    p += slots * static_cast<size_t>(c.kmax) * 8;
  }
  if (mode != 0) {
    sc.off = reinterpret_cast<uint32_t*>(p);
    p += (slots * 4 + 7) / 8 * 8;
    sc.view_seg = reinterpret_cast<int32_t*>(p);
    p += (nv * 4 + 7) / 8 * 8;
  }
  sc.view_flags = reinterpret_cast<unsigned int*>(p);
  p += (nv * 4 + 7) / 8 * 8;
  if (c.kind == A3D_CASTER_ITERATIVE) {
    sc.table = reinterpret_cast<uint32_t*>(p);
    p += nv * stride * 4;
    sc.dlist = reinterpret_cast<uint32_t*>(p);
    p += nv * stride * 4;
    sc.dcount = reinterpret_cast<unsigned int*>(p);
    p += (nv * 4 + 7) / 8 * 8;
    sc.start_s = p;
  }
  return sc;
}

}  // namespace

#if A3D_SPLIT_PART == 1
bool cast_split_supported(const MapView& m, const CasterView& c) {
  if (!m.vol || m.vps_shift < 0) return false;
  if (c.kmax >= (1 << kCntBits) - 1) return false;  // a part's counters are 15 bits wide
  if (c.kind == A3D_CASTER_ITERATIVE && (c.diag_max > kSplitMaxDiag || c.n_sections > 31)) return false;
  return c.n_rays < (1 << 23);
}

size_t cast_split_bytes_per_view(const CasterView& c, int eval_kind, size_t* stride_out, int mode) {
  return split_bytes_per_view(c, eval_kind, stride_out, mode);
}

int cast_split_launches(const CasterView& c, int mode) {
  return (c.kind == A3D_CASTER_ITERATIVE ? 5 : 3) + (mode == 1 ? 1 : 0);
}

// Segments [seg0, seg0 + n_seg) = views [view0, view0 + n_views) of the batch; scratch holds
// n_views * cast_split_bytes_per_view() + 256 bytes.  inexact: one bit per segment (batch-wide
// numbering), set when a sample needs the scan-order kernel's decision.  mode: see split_bytes_per_view
// (1 needs b.centres_out or b.keys_out).
int launch_cast_split(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                      int seg0, int n_seg, int view0, int n_views, void* scratch, unsigned int* inexact,
                      int sm_count, cudaEvent_t* stage_events, int mode) {
  const SplitScratch sc = split_layout(c, e.kind, mode, n_views, scratch);
  if (c.kind == A3D_CASTER_ITERATIVE)
    return launch_split_c<A3D_CASTER_ITERATIVE, false>(st, m, c, e, b, seg0, n_seg, view0, n_views, sc, inexact,
                                                       sm_count, stage_events);
  return launch_split_c<A3D_CASTER_SIMPLE, false>(st, m, c, e, b, seg0, n_seg, view0, n_views, sc, inexact, sm_count,
                                                  stage_events);
}
#else
// Second pass of a stored-list call over the whole batch (segments [0, n_seg), views [0, n_views)): the
// scratch is the one launch_cast_split(mode 2) left behind for the same batch; every retained entry goes
// to b.keys_out[b.emit_offset[segment] + its place in the segment's list] as a packed key.  Launches:
// cast_split_launches(c, 0) - 1 (no fix-up).
int launch_cast_split_direct(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e, const Batch& b,
                             int n_seg, int n_views, void* scratch, int sm_count) {
  SplitScratch sc = split_layout(c, e.kind, 2, n_views, scratch);
  sc.direct = b.keys_out;
  sc.seg_base = b.emit_offset;
  if (c.kind == A3D_CASTER_ITERATIVE)
    return launch_split_c<A3D_CASTER_ITERATIVE, true>(st, m, c, e, b, 0, n_seg, 0, n_views, sc, nullptr, sm_count,
                                                      nullptr);
  return launch_split_c<A3D_CASTER_SIMPLE, true>(st, m, c, e, b, 0, n_seg, 0, n_views, sc, nullptr, sm_count, nullptr);
}
#endif

}  // namespace a3d
