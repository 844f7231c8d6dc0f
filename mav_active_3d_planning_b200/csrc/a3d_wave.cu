// a3d_wave.cu — k_cast_wave: the throughput kernel of the view-gain path (DESIGN.md "kernel K3w").
//
// Same results as k_cast_gain (a3d_kernels.cu) for gain / n_visible / n_samples and an
// order-independent digest of the stored list, but mapped lane-per-ray instead of warp-per-ray:
//
//  * One CTA of 2 warps owns a trajectory segment at a time (atomic work counter).
//  * Every lane marches its OWN ray part: the previous raw list entry it needs for the
//    adjacent-unique test (camera_model.cpp:58) is in its own registers, the sample distance is
//    accumulated in a register exactly like the reference does (`distance += p_ray_step_`), and no
//    shuffle or ballot is needed per sample.  Voxel identity is the integer triple
//    g = block*vps + voxel (+ artefact flags), equal iff the reference's double centre triples are
//    equal (fp artefact indices are canonicalised or flagged, see make_ident_slow).  Lanes are
//    persistent: when a ray ends (range or OCCUPIED hit) the lane pulls the next ready ray, so
//    occlusion does not idle lanes.
//  * What couples rays — `std::unique` across the boundary between two consecutive rays / views —
//    is repaired afterwards: each ray part records its first and last key and what its first entry
//    contributed; a scan-order pass (fix-up) un-counts first entries that equal the previous
//    emitted entry.
//  * IterativeRayCaster (iterative_ray_caster.cpp:48-119): the ray table makes ray (i,j) depend on
//    rays (i'<=i, j'<=j) only, so rays on one anti-diagonal are independent.  Warp 0 walks the
//    anti-diagonals: it casts the rays that start before the last section only up to
//    c_split_distances_[n-1] — the part that decides every markNeighboringRays call ("marking
//    parts") — and applies their marks in scan order, last-writer-wins resolved by a per-entry
//    writer index.  Every ray whose table entry then reads n-1 (fresh last-section starts AND
//    survivors of the marking parts, whose own cell gets marked n-1) has its last section cast as a
//    "leaf part".  Leaf parts come from a shared supply (one unit = up to 32 rays of a finished
//    anti-diagonal); warp 1 only casts leaf parts, warp 0 fills the lanes its marking parts leave
//    idle with leaf parts too.
//  * SimpleRayCaster (simple_ray_caster.cpp:32-66): all rays independent; both warps cast leaf
//    parts (whole rays).
#include <algorithm>

#include "../../include/a3d.h"
#include "a3d_kernels.cuh"

namespace a3d {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
#ifndef A3D_WAVE_WARPS
#define A3D_WAVE_WARPS 2  // warps per CTA: warp 0 schedules (iterative caster), all cast leaf parts
#endif
constexpr int kWaveWarps = A3D_WAVE_WARPS;
constexpr int kWaveThreads = 32 * kWaveWarps;
constexpr uint64_t kNoKey = ~0ull;  // "no entry"
constexpr uint64_t kKeyOverflow = 1ull << 63;
constexpr int kKeyBits = 20, kKeyBias = 1 << 19;
constexpr int kCatNone = 7;
constexpr int kNoIdent = INT32_MIN;
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
};

// Per-lane accumulators of a segment.  `packed`: three 21-bit fields, field c counts list entries of
// fold category c (VoxelType: voxel state 0/1/2; Naive/Frontier: field 0 = entries that score).
// field 0..2: VoxelType → entries of voxel state 0/1/2; Naive/Frontier → field 0 = entries that score,
// field 1 = other retained entries.  n_visible = sum of the fields.
struct LaneCounters {
  unsigned long long packed;
  unsigned long long digest;  // DIGEST instantiations only
  unsigned n_samp;
  unsigned inexact;
};
constexpr int kFieldBits = 21;

// Identity of an emitted list entry: g[k] = b[k]*vps + v[k]; bit 30 of g[k] set: axis k keeps a raw
// artefact voxel index (v = -1 or vps) whose centre double differs from the canonical voxel's
// (valid g are < 2^19 in magnitude, larger ones flag the segment for the scan-order kernel).
struct Ident {
  int g0, g1, g2;
};
constexpr int kRawAxisBit = 1 << 30;

// Artefact voxel indices (getGridIndexFromPoint lands one voxel outside the block because of the
// float block-index rounding): per axis, name the voxel canonically when its centre double
// (voxblox.cpp:72-79) equals the canonical voxel's centre double, else keep it raw and flag it.
__device__ __noinline__ Ident make_ident_slow(const MapView& m, const CenterTables& ct, int b0, int b1,
                                              int b2, int v0, int v1, int v2, unsigned* inexact) {
  const int b[3] = {b0, b1, b2}, v[3] = {v0, v1, v2};
  int g[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g[k] = b[k] * m.vps + v[k];
    if (g[k] <= -kKeyBias || g[k] >= kKeyBias) *inexact = 1;
    if (v[k] >= 0 && v[k] < m.vps) continue;
    if (v[k] < -1 || v[k] > m.vps) {
      *inexact = 1;  // not an artefact the reference arithmetic can produce; scan-order kernel decides
      continue;
    }
    const int bq = v[k] < 0 ? b[k] - 1 : b[k] + 1;
    const int vq = v[k] < 0 ? m.vps - 1 : 0;
    const double a = __dadd_rn(static_cast<double>(origin_point(b[k], m.bs_f)), ct.d[v[k] + 1]);
    const double c = __dadd_rn(static_cast<double>(origin_point(bq, m.bs_f)), ct.d[vq + 1]);
    if (a != c) g[k] ^= kRawAxisBit;
  }
  return Ident{g[0], g[1], g[2]};
}

__device__ __forceinline__ uint64_t pack_key(const Ident& id, unsigned* inexact) {
  // bit30 != bit31 ⇒ raw-axis flag
  const int flags = static_cast<int>(static_cast<unsigned>(id.g0 ^ (id.g0 << 1)) >> 31) |
                    static_cast<int>(static_cast<unsigned>(id.g1 ^ (id.g1 << 1)) >> 31) << 1 |
                    static_cast<int>(static_cast<unsigned>(id.g2 ^ (id.g2 << 1)) >> 31) << 2;
  const int h0 = (flags & 1) ? id.g0 ^ kRawAxisBit : id.g0, h1 = (flags & 2) ? id.g1 ^ kRawAxisBit : id.g1,
            h2 = (flags & 4) ? id.g2 ^ kRawAxisBit : id.g2;
  const unsigned u0 = static_cast<unsigned>(h0 + kKeyBias), u1 = static_cast<unsigned>(h1 + kKeyBias),
                 u2 = static_cast<unsigned>(h2 + kKeyBias);
  if ((u0 | u1 | u2) >> kKeyBits) {
    *inexact = 1;
    return kKeyOverflow;
  }
  return static_cast<uint64_t>(u0) | (static_cast<uint64_t>(u1) << kKeyBits) |
         (static_cast<uint64_t>(u2) << (2 * kKeyBits)) | (static_cast<uint64_t>(flags) << 60);
}

// Per-lane state of the ray part a lane is marching (kept small: it lives in registers across the
// whole engine loop).
struct RayState {
  double dx, dy, dz;  // world direction
  double d;           // distance of the next sample (accumulated like `distance += p_ray_step_`)
  Ident prev;         // previous raw entry (g0 == kNoIdent: none yet)
  int r;              // ray id
};

struct ViewShared {  // per-view constants, one copy per CTA in shared memory
  double P0[3];
  double q[4];
};

__device__ __forceinline__ void ray_begin(RayState& rs, const CasterView& c, const ViewShared& vw,
                                          int r, double d0) {
  const Q4 q{vw.q[0], vw.q[1], vw.q[2], vw.q[3]};
  const D3 dir = rotate(q, D3{c.cam_dirs[3 * r], c.cam_dirs[3 * r + 1], c.cam_dirs[3 * r + 2]});
  rs.dx = dir.x;
  rs.dy = dir.y;
  rs.dz = dir.z;
  rs.d = d0;
  rs.prev = Ident{kNoIdent, 0, 0};
  rs.r = r;
}

// Everything about one sample that the list logic needs.
struct SampleInfo {
  int st;    // voxel state at the sample point
  int cs;    // state | observed << 2 at the centre of the voxel getVoxelCenter names
  Ident id;  // identity of that voxel
};

// Rare cases of a sample, evaluated literally: artefact voxel indices, interpolation cells that
// straddle the occupied threshold, centre voxel != state voxel.
__device__ __noinline__ SampleInfo sample_slow(const MapView& m, const CenterTables& ct, float pf0,
                                               float pf1, float pf2, int b0, int b1, int b2, int slot,
                                               int vi0, int vi1, int vi2, unsigned* inexact) {
  const float pf[3] = {pf0, pf1, pf2};
  const int bi[3] = {b0, b1, b2}, vi[3] = {vi0, vi1, vi2};
  float origin[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) origin[a] = origin_point(bi[a], m.bs_f);
  SampleInfo si;
  int vs[3];
  uint32_t word;
  bool vs_ok;
  si.st = voxel_state_in_block(m, ct, pf, bi, origin, slot, vs, &word, &vs_ok);
  const unsigned vmax = static_cast<unsigned>(m.vps);
  const bool canonical = (static_cast<unsigned>(vi0) < vmax) & (static_cast<unsigned>(vi1) < vmax) &
                         (static_cast<unsigned>(vi2) < vmax);
  if (canonical) {
    si.id = Ident{b0 * m.vps + vi0, b1 * m.vps + vi1, b2 * m.vps + vi2};
  } else {
    si.id = make_ident_slow(m, ct, b0, b1, b2, vi0, vi1, vi2, inexact);
  }
  if (vs_ok & (vi0 == vs[0]) & (vi1 == vs[1]) & (vi2 == vs[2])) {
    si.cs = centre_state_from_word(word);
  } else {
    double cp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      cp[a] = (vi[a] >= -1 && vi[a] <= m.vps) ? ct.d[vi[a] + 1]
                                              : static_cast<double>(center_point(vi[a], m.vs_f));
    const D3 cc{__dadd_rn(static_cast<double>(origin[0]), cp[0]),
                __dadd_rn(static_cast<double>(origin[1]), cp[1]),
                __dadd_rn(static_cast<double>(origin[2]), cp[2])};
    si.cs = centre_state(m, ct, cc, slot, vi);
  }
  return si;
}

// One sample of the lane's ray.  Returns true when the part is finished (hit or end reached);
// *hit_d = distance of the OCCUPIED sample that ended it (unchanged otherwise).
// rec_at: index of this part's record.  Precondition: rs.d < d_end.
//
// The common case — allocated or unallocated block, in-range voxel indices that agree between
// getVoxelState's and getVoxelCenter's index arithmetic, definite cell class — is straight-line
// code around one meta-word load; everything else goes through sample_slow().
template <int CASTER, int EVAL, bool DIGEST>
__device__ __forceinline__ bool ray_step(const MapView& m, const CenterTables& ct, const EvalView& e,
                                         const ViewShared& vw, const double step, const double d_end,
                                         RayState& rs, LaneCounters& lc, const Records& rec,
                                         size_t rec_at, double* hit_d) {
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  const double d_this = rs.d;
  const double pp[3] = {__dadd_rn(vw.P0[0], __dmul_rn(d_this, rs.dx)),
                        __dadd_rn(vw.P0[1], __dmul_rn(d_this, rs.dy)),
                        __dadd_rn(vw.P0[2], __dmul_rn(d_this, rs.dz))};
  rs.d = __dadd_rn(d_this, step);
  lc.n_samp++;
  float pf[3], origin[3];
  int bi[3], vs[3], vi[3];
  int sc = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    pf[a] = __double2float_rn(pp[a]);
    bi[a] = grid_index(pf[a], m.bs_inv_f);
    origin[a] = origin_point(bi[a], m.bs_f);
    // voxel index as getVoxelState derives it (float subtraction) ...
    vs[a] = grid_index(__fsub_rn(pf[a], origin[a]), m.vs_inv_f);
    // ... and as getVoxelCenter derives it (double subtraction, voxblox.cpp:74-77)
    vi[a] = grid_index(__double2float_rn(__dsub_rn(pp[a], static_cast<double>(origin[a]))),
                       m.vs_inv_f);
  }
  const unsigned vmax = static_cast<unsigned>(m.vps);
  const bool in_range = (static_cast<unsigned>(vs[0]) < vmax) & (static_cast<unsigned>(vs[1]) < vmax) &
                        (static_cast<unsigned>(vs[2]) < vmax);
  const bool same = (vs[0] == vi[0]) & (vs[1] == vi[1]) & (vs[2] == vi[2]);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float centre = __fadd_rn(origin[a], ct.f[min(static_cast<unsigned>(vs[a]), vmax) + 1]);
    sc |= (__fsub_rn(pf[a], centre) < 0.0f) ? (1 << a) : 0;
  }
  const int slot = find_slot(m, bi[0], bi[1], bi[2]);
  uint32_t word = kWordNoBlock;
  if (slot >= 0 && in_range)
    word = __ldg(m.mword + static_cast<size_t>(slot) * m.nv + mword_offset(m, vs[0], vs[1], vs[2]));
  const int cls = (word >> (2 * sc)) & 3;
  SampleInfo si;
  si.st = cls == kCellUnknown ? 2 : (cls == kCellFree ? 1 : 0);
  si.cs = centre_state_from_word(word);
  si.id = Ident{bi[0] * m.vps + vi[0], bi[1] * m.vps + vi[1], bi[2] * m.vps + vi[2]};
  if (!(in_range & same) | (cls == kCellMixed)) {
    unsigned bad = 0;
    si = sample_slow(m, ct, pf[0], pf[1], pf[2], bi[0], bi[1], bi[2], slot, vi[0], vi[1], vi[2], &bad);
    lc.inexact |= bad;
  }
  const bool emit = ITER || si.st != 0;
  bool keep = emit & ((si.id.g0 != rs.prev.g0) | (si.id.g1 != rs.prev.g1) | (si.id.g2 != rs.prev.g2));
  const bool first = emit & (rs.prev.g0 == kNoIdent);
  if (emit) rs.prev = si.id;
  D3 cc{0.0, 0.0, 0.0};
  if (DIGEST || e.bv_is_setup || EVAL == A3D_EVAL_FRONTIER) {
    double cp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      cp[a] = (vi[a] >= -1 && vi[a] <= m.vps) ? ct.d[vi[a] + 1]
                                              : static_cast<double>(center_point(vi[a], m.vs_f));
    cc.x = __dadd_rn(static_cast<double>(origin[0]), cp[0]);
    cc.y = __dadd_rn(static_cast<double>(origin[1]), cp[1]);
    cc.z = __dadd_rn(static_cast<double>(origin[2]), cp[2]);
    if (e.bv_is_setup) keep = keep && bv_contains(e, cc);  // simulated_sensor_evaluator.cpp:50-57
  }
  int cat;
  if (EVAL == A3D_EVAL_VOXEL_TYPE) {
    cat = si.cs & 3;  // retained entries are always inside the bounding volume
  } else {
    cat = 1;
    if (keep && !(si.cs & 4)) {
      if (EVAL == A3D_EVAL_NAIVE) {
        cat = 0;
      } else if (is_frontier_voxel(m, ct, e, cc)) {
        cat = 0;
      }
    }
  }
  lc.packed += static_cast<unsigned long long>(keep) << (kFieldBits * cat);
  unsigned long long h = 0;
  if (DIGEST) {
    h = keep ? digest_entry(cc) : 0ull;
    lc.digest += h;
  }
  if (first) {
    // record of the part's first entry, written once
    unsigned bad = 0;
    rec.first_key[rec_at] = pack_key(si.id, &bad);
    lc.inexact |= bad;
    rec.info[rec_at] = static_cast<uint8_t>((keep ? cat : kCatNone) | 8 | (keep ? 16 : 0));
    if (DIGEST) rec.first_hash[rec_at] = h;
  }
  if (si.st == 0) {
    *hit_d = d_this;
    return true;
  }
  return !(rs.d < d_end);
}

#ifndef A3D_WAVE_STEPS
#define A3D_WAVE_STEPS 8  // samples marched between two rounds of lane bookkeeping
#endif

}  // namespace

template <int CASTER, int EVAL, bool DIGEST>
__global__ void __launch_bounds__(kWaveThreads, A3D_WAVE_MIN_BLOCKS)
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
  __shared__ unsigned s_pending[kWaveWarps][kPendCap];  // per-warp ring of ready rays waiting for a lane
  // marking rays of the current diagonal (dynamic shared memory, sized by the longest diagonal):
  // their hit distance (< 0: none) and r << 8 | s
  extern __shared__ double s_dyn[];
  double* const s_mark_hit = s_dyn;
  unsigned* const s_mark_list = reinterpret_cast<unsigned*>(s_dyn + c.diag_max);
  __shared__ unsigned long long s_packed, s_digest;
  __shared__ unsigned s_nsamp, s_inexact;
  init_center_tables(m, &ct);

  const size_t cta_base = static_cast<size_t>(blockIdx.x) * ws.rays_stride;
  volatile uint32_t* const table = ws.table + cta_base;
  volatile uint8_t* const start_s = ws.start_s + cta_base;
  const double step = c.ray_step;
  const double d_leaf_end = c.ray_length;
  const double d_mark_end = (ITER && n > 1) ? c.split[n - 1] : 0.0;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_seg = static_cast<int>(atomicAdd(b.work_counter, 1u));
    __syncthreads();
    const int seg = s_seg;
    if (seg >= b.n_segments) break;
    const int v_begin = b.view_first[seg], v_end = b.view_first[seg + 1];

    LaneCounters lc{};
    uint64_t carry_key = kNoKey;  // last emitted key so far in this segment (warp 0, uniform)

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
        s_done_diag = (ITER && n > 1) ? 0 : n_diag;
        s_unit_cursor = 0;
      }
      for (int r = threadIdx.x; r < n_rays; r += kWaveThreads) {
        if (ITER) {
          table[r] = 0u;
          start_s[r] = 0xFF;
        }
        ws.rec.info[cta_base * 2 + r] = kCatNone;
        ws.rec.info[cta_base * 2 + n_rays + r] = kCatNone;
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
        rs.r = -1;
        int role = 0;  // 0 idle, 1 leaf part, 2 marking part
        int mark_idx = 0;

        for (;;) {
          // ---- 1. scheduler: finish the current diagonal, open the next ones ----
          while (!sched_done && mark_head == n_mark && n_run == 0) {
            if (cur_diag >= 0) {
              // markNeighboringRays of every marking ray of the diagonal, in scan order
              for (int qi = 0; qi < n_mark; ++qi) {
                const unsigned it = s_mark_list[qi];
                const int rr = static_cast<int>(it >> 8), ss = static_cast<int>(it & 0xFF);
                const double hd = s_mark_hit[qi];
                const bool hit = hd >= 0.0;
                int seg_h = ss;
                if (hit) {
                  while (!(hd < c.split[seg_h + 1])) ++seg_h;
                }
                const int t_last = hit ? seg_h : n - 2;
                const int w0 = 1 << (n - 1 - ss);
                const int ri = rr / c.res_y, rj = rr - ri * c.res_y;
                if (w0 == 2) {
                  // start section n-2 (most marking rays): the 2x2 square gets one value
                  const int val = hit ? -1 : n - 1;
                  const int dx = lane >> 1, dy = lane & 1;
                  if (lane < 4 && ri + dx < c.res_x && rj + dy < c.res_y) {
                    const int at = (ri + dx) * c.res_y + rj + dy;
                    const unsigned writer = static_cast<unsigned>(rr) + 1u;
                    if ((table[at] >> 8) < writer) table[at] = (writer << 8) | static_cast<uint8_t>(val);
                  }
                  __syncwarp();
                  continue;
                }
                const int nx = min(w0, c.res_x - ri), ny = min(w0, c.res_y - rj);
                const unsigned writer = static_cast<unsigned>(rr) + 1u;
                for (int cell = lane; cell < nx * ny; cell += 32) {
                  const int dx = cell / ny, dy = cell - dx * ny;
                  const int mx = max(dx, dy);
                  const int tmax = mx == 0 ? n - 1 : n - 2 - (31 - __clz(mx));
                  const int t = min(tmax, t_last);
                  const int val = (hit && t == seg_h) ? -1 : t + 1;
                  const int at = (ri + dx) * c.res_y + rj + dy;
                  if ((table[at] >> 8) < writer) table[at] = (writer << 8) | static_cast<uint8_t>(val);
                }
                __syncwarp();
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
                s = static_cast<int>(static_cast<int8_t>(table[r] & 0xFFu));
              }
              const bool marking = s >= 0 && s < n - 1;
              const unsigned mm = __ballot_sync(kFull, marking);
              if (marking)
                s_mark_list[n_mark + __popc(mm & lt_mask)] = (static_cast<unsigned>(r) << 8) | s;
              n_mark += __popc(mm);
            }
            __syncwarp();
          }

          // ---- 2. give work to idle lanes: marking parts first, then leaf parts ----
          const unsigned idle = __ballot_sync(kFull, role == 0);
          if (idle) {
            int n_idle = __popc(idle);
            const int rank = __popc(idle & lt_mask);
            int take_mark = 0;
            if (scheduler) {
              take_mark = min(n_idle, n_mark - mark_head);
              if (role == 0 && rank < take_mark) {
                const unsigned it = s_mark_list[mark_head + rank];
                mark_idx = mark_head + rank;
                const int ms = static_cast<int>(it & 0xFF);
                ray_begin(rs, c, vw, static_cast<int>(it >> 8), c.split[ms]);
                start_s[rs.r] = static_cast<uint8_t>(ms);  // read by leaf lanes after the diagonal
                s_mark_hit[mark_idx] = -1.0;
                role = 2;  // split[s] < split[n-1] for s < n-1: at least one sample
              }
              mark_head += take_mark;
              n_run += take_mark;
              n_idle -= take_mark;
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
                    ready = static_cast<int>(static_cast<int8_t>(table[rr] & 0xFFu)) == n - 1;
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
              if (role == 0 && lrank >= 0 && lrank < pend_count) {
                const int r = static_cast<int>(pending[(pend_head + lrank) % kPendCap]);
                double d0 = 0.0;
                if (ITER) {
                  const int s0 = start_s[r] == 0xFF ? n - 1 : start_s[r];
                  d0 = c.leaf_d0[s0];  // accumulated distance at which a ray started in s0 enters section n-1
                }
                ray_begin(rs, c, vw, r, d0);
                role = (rs.d < d_leaf_end) ? 1 : 0;
              }
              const int taken = min(n_idle, pend_count);
              pend_head = (pend_head + taken) % kPendCap;
              pend_count -= taken;
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

          // ---- 4. a few samples per busy lane ----
          bool fin_mark = false;
#pragma unroll 1
          for (int it = 0; it < A3D_WAVE_STEPS; ++it) {
            if (role != 0) {
              const size_t rec_at = cta_base * 2 + (role == 2 ? 0 : static_cast<size_t>(n_rays)) + rs.r;
              double hit_d = -1.0;
              if (ray_step<CASTER, EVAL, DIGEST>(m, ct, e, vw, step, role == 2 ? d_mark_end : d_leaf_end,
                                                 rs, lc, ws.rec, rec_at, &hit_d)) {
                if (rs.prev.g0 != kNoIdent) {
                  unsigned bad = 0;
                  ws.rec.last_key[rec_at] = pack_key(rs.prev, &bad);
                  lc.inexact |= bad;
                }
                if (role == 2) {
                  s_mark_hit[mark_idx] = hit_d;
                  fin_mark = true;
                }
                role = 0;
              }
            }
          }
          if (scheduler) n_run -= __popc(__ballot_sync(kFull, fin_mark));
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
            info = ws.rec.info[at];
            if (info & 8) {
              fk = ws.rec.first_key[at];
              lk = ws.rec.last_key[at];
            }
          }
          const bool emitted = (info & 8) != 0;
          const unsigned em = __ballot_sync(kFull, emitted);
          const unsigned below = em & lt_mask;
          const int src = below ? 31 - __clz(below) : 0;
          uint64_t prev = __shfl_sync(kFull, lk, src);
          if (!below) prev = carry_key;
          if (emitted && fk == prev && (info & 16)) {
            // the first entry of this part repeats the previous raw entry: std::unique drops it
            lc.packed -= 1ull << (kFieldBits * (info & 7));
            if (DIGEST) lc.digest -= ws.rec.first_hash[at];
          }
          if (em) carry_key = __shfl_sync(kFull, lk, 31 - __clz(em));
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
      s_nsamp = 0;
      s_inexact = 0;
    }
    __syncthreads();
    {
      unsigned long long pk = lc.packed, dg = lc.digest;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        pk += __shfl_xor_sync(kFull, pk, off);
        if (DIGEST) dg += __shfl_xor_sync(kFull, dg, off);
      }
      const unsigned tsmp = __reduce_add_sync(kFull, lc.n_samp);
      const unsigned tin = __reduce_or_sync(kFull, lc.inexact);
      if (lane == 0) {
        atomicAdd(&s_packed, pk);
        atomicAdd(&s_nsamp, tsmp);
        atomicOr(&s_inexact, tin);
        if (DIGEST) atomicAdd(&s_digest, dg);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long fmask = (1ull << kFieldBits) - 1;
      const unsigned long long c0 = s_packed & fmask, c1 = (s_packed >> kFieldBits) & fmask,
                               c2 = (s_packed >> (2 * kFieldBits)) & fmask;
      unsigned inexact = s_inexact;
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

size_t wave_scratch_bytes(int n_rays, int grid, bool digest, size_t* rays_stride) {
  const size_t stride = (static_cast<size_t>(n_rays) + 31) / 32 * 32;
  *rays_stride = stride;
  size_t per_cta = stride * 4 + stride + 2 * stride * (8 + 8 + 1) + (digest ? 2 * stride * 8 : 0);
  return per_cta * static_cast<size_t>(grid) + 256;
}

template <int CASTER, int EVAL>
static void launch_wave_t(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e,
                          const Batch& b, bool digest, int grid, const WaveScratch& ws,
                          unsigned int* inexact) {
  if (digest)
    k_cast_wave<CASTER, EVAL, true>
        <<<grid, kWaveThreads, static_cast<size_t>(c.diag_max) * 12, st>>>(m, c, e, b, ws, inexact);
  else
    k_cast_wave<CASTER, EVAL, false>
        <<<grid, kWaveThreads, static_cast<size_t>(c.diag_max) * 12, st>>>(m, c, e, b, ws, inexact);
}

template <int CASTER>
static void launch_wave_c(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e,
                          const Batch& b, bool digest, int grid, const WaveScratch& ws,
                          unsigned int* inexact) {
  switch (e.kind) {
    case A3D_EVAL_VOXEL_TYPE:
      launch_wave_t<CASTER, A3D_EVAL_VOXEL_TYPE>(st, m, c, e, b, digest, grid, ws, inexact);
      break;
    case A3D_EVAL_NAIVE:
      launch_wave_t<CASTER, A3D_EVAL_NAIVE>(st, m, c, e, b, digest, grid, ws, inexact);
      break;
    default:
      launch_wave_t<CASTER, A3D_EVAL_FRONTIER>(st, m, c, e, b, digest, grid, ws, inexact);
      break;
  }
}

// scratch: wave_scratch_bytes() bytes, carved here.  b.set_digests receives the ORDER-INDEPENDENT
// digest (sum of digest_entry over the stored list, mod 2^64).  inexact: one bit per segment, set
// when a voxel key overflowed the 20-bit packing (the host re-evaluates those with k_cast_gain).
void launch_cast_wave(cudaStream_t st, const MapView& m, const CasterView& c, const EvalView& e,
                      const Batch& b, bool digest, int grid, void* scratch, unsigned int* inexact) {
  size_t stride = 0;
  wave_scratch_bytes(c.n_rays, grid, digest, &stride);
  WaveScratch ws{};
  ws.rays_stride = stride;
  uint8_t* p = static_cast<uint8_t*>(scratch);
  const size_t g = static_cast<size_t>(grid);
  ws.rec.first_key = reinterpret_cast<unsigned long long*>(p);
  p += g * stride * 2 * 8;
  ws.rec.last_key = reinterpret_cast<unsigned long long*>(p);
  p += g * stride * 2 * 8;
  if (digest) {
    ws.rec.first_hash = reinterpret_cast<unsigned long long*>(p);
    p += g * stride * 2 * 8;
  }
  ws.table = reinterpret_cast<uint32_t*>(p);
  p += g * stride * 4;
  ws.rec.info = p;
  p += g * stride * 2;
  ws.start_s = p;
  if (c.kind == A3D_CASTER_ITERATIVE)
    launch_wave_c<A3D_CASTER_ITERATIVE>(st, m, c, e, b, digest, grid, ws, inexact);
  else
    launch_wave_c<A3D_CASTER_SIMPLE>(st, m, c, e, b, digest, grid, ws, inexact);
}

int cast_wave_threads() { return kWaveThreads; }
int cast_wave_max_diag() { return kMarkCap; }
int cast_wave_blocks_per_sm() { return A3D_WAVE_MIN_BLOCKS; }

}  // namespace a3d
