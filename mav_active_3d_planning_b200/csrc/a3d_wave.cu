// a3d_wave.cu — k_cast_wave: the throughput kernel of the view-gain path (DESIGN.md "kernel K3w").
//
// Same results as k_cast_gain (a3d_kernels.cu) for gain / n_visible / n_samples and an
// order-independent digest of the stored list, but mapped lane-per-ray instead of warp-per-ray:
//
//  * One CTA of 2 warps owns a trajectory segment at a time (atomic work counter).
//  * Every lane marches its OWN ray: the previous raw list entry it needs for the adjacent-unique
//    test (camera_model.cpp:58) is in its own registers, the sample distance is accumulated in a
//    register exactly like the reference does (`distance += p_ray_step_`), and no shuffle or ballot
//    is needed per sample.  Voxel identity is a packed integer key (block*vps+voxel per axis, 20
//    bits each) that is equal iff the reference's double centre triples are equal (fp artefact
//    indices are canonicalised or flagged, see make_key).
//  * What couples rays — `std::unique` across the boundary between two consecutive rays / views —
//    is repaired afterwards: each ray part records its first and last key and what its first entry
//    contributed; a scan-order pass (fix_up) un-counts first entries that equal the previous
//    emitted entry.
//  * IterativeRayCaster (iterative_ray_caster.cpp:48-119): the ray table makes ray (i,j) depend on
//    rays (i'<=i, j'<=j) only, so rays on one anti-diagonal are independent.  Warp 0 ("scheduler")
//    walks the anti-diagonals, casts the rays that start before the last section only up to
//    c_split_distances_[n-1] — the part that decides every markNeighboringRays call — and applies
//    their marks in scan order with last-writer-wins resolved by a per-entry writer index.  Rays
//    whose entry reads n-1 (fresh last-section starts AND survivors of warp 0, whose own cell gets
//    marked n-1) are cast through the last section by warp 1 ("leaf") with persistent lanes that
//    pull a new ray as soon as theirs ends, so occlusion does not idle lanes.
//  * SimpleRayCaster (simple_ray_caster.cpp:32-66): all rays independent; both warps are leaf warps.
#include <algorithm>

#include "../../include/a3d.h"
#include "a3d_kernels.cuh"

namespace a3d {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kWaveThreads = 64;
constexpr uint64_t kNoKey = ~0ull;          // "no previous entry"
constexpr uint64_t kKeyOverflow = 1ull << 63;
constexpr int kKeyBits = 20, kKeyBias = 1 << 19;
constexpr int kCatNone = 7;
#ifndef A3D_WAVE_MIN_BLOCKS
#define A3D_WAVE_MIN_BLOCKS 10  // resident 2-warp CTAs per SM the register allocation must allow
#endif

// info byte of a ray-part record: bits 2:0 category counted for the first entry (kCatNone: not
// counted), bit 3: part emitted at least one entry
struct Records {
  unsigned long long* first_key;   // [2][n_rays]
  unsigned long long* last_key;    // [2][n_rays]
  unsigned long long* first_hash;  // [2][n_rays] (digest mode only)
  uint8_t* info;                   // [2][n_rays]
};

struct LaneCounters {
  unsigned cnt[6];
  unsigned n_vis, n_samp;
  unsigned long long digest;
  unsigned inexact;
};

// Packed voxel identity of getVoxelCenter's result for block index b, unclamped voxel index v.
// Two entries have equal keys iff their centre doubles (voxblox.cpp:72-79) are equal:
//  - v in [0,vps) on every axis (canonical): the centre is a function of g = b*vps+v alone;
//  - an artefact index (v = -1, vps, ...) names the same voxel as (b', v mod vps); its centre double
//    is compared with that voxel's canonical centre and the axis is flagged if they differ, so a
//    flagged key only equals keys built from the very same (b, v).
__device__ __noinline__ uint64_t make_key_slow(const MapView& m, const CenterTables& ct,
                                               const int (&b)[3], const int (&v)[3]) {
  uint64_t key = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const long long g = static_cast<long long>(b[k]) * m.vps + v[k];
    if (g < -kKeyBias || g >= kKeyBias) return kKeyOverflow;
    key |= static_cast<uint64_t>(g + kKeyBias) << (kKeyBits * k);
    if (v[k] < 0 || v[k] >= m.vps) {
      const int bq = b[k] + (v[k] < 0 ? -((-v[k] + m.vps - 1) / m.vps) : v[k] / m.vps);
      const int vq = v[k] - (bq - b[k]) * m.vps;
      const double cp_a = (v[k] >= -1 && v[k] <= m.vps)
                              ? ct.d[v[k] + 1]
                              : static_cast<double>(center_point(v[k], m.vs_f));
      const double a = __dadd_rn(static_cast<double>(origin_point(b[k], m.bs_f)), cp_a);
      const double c = __dadd_rn(static_cast<double>(origin_point(bq, m.bs_f)), ct.d[vq + 1]);
      if (a != c) key |= 1ull << (60 + k);
    }
  }
  return key;
}

__device__ __forceinline__ uint64_t make_key(const MapView& m, const CenterTables& ct,
                                             const int (&b)[3], const int (&v)[3], bool* canonical) {
  const unsigned vmax = static_cast<unsigned>(m.vps);
  const bool canon = (static_cast<unsigned>(v[0]) < vmax) & (static_cast<unsigned>(v[1]) < vmax) &
                     (static_cast<unsigned>(v[2]) < vmax);
  const int g0 = b[0] * m.vps + v[0] + kKeyBias, g1 = b[1] * m.vps + v[1] + kKeyBias,
            g2 = b[2] * m.vps + v[2] + kKeyBias;
  const bool in_range = (static_cast<unsigned>(g0 | g1 | g2) >> kKeyBits) == 0 &&
                        (abs(b[0]) | abs(b[1]) | abs(b[2])) < (1 << 24);
  *canonical = canon;
  if (canon & in_range)
    return static_cast<uint64_t>(g0) | (static_cast<uint64_t>(g1) << kKeyBits) |
           (static_cast<uint64_t>(g2) << (2 * kKeyBits));
  return make_key_slow(m, ct, b, v);
}

// Per-lane state of the ray part a lane is marching.
struct RayState {
  D3 dir;
  double d;      // distance of the next sample (accumulated like `distance += p_ray_step_`)
  double d_end;  // exclusive end of this part
  uint64_t prev_key, first_key;
  unsigned long long first_hash;
  double hit_d;  // distance of the OCCUPIED sample that ended the part, < 0 if none
  int info;
  int r;
};

__device__ __forceinline__ void ray_begin(RayState& rs, const CasterView& c, const Q4& q, int r,
                                          double d0, double d_end) {
  rs.dir = rotate(q, D3{c.cam_dirs[3 * r], c.cam_dirs[3 * r + 1], c.cam_dirs[3 * r + 2]});
  rs.d = d0;
  rs.d_end = d_end;
  rs.prev_key = kNoKey;
  rs.first_key = kNoKey;
  rs.first_hash = 0;
  rs.hit_d = -1.0;
  rs.info = kCatNone;
  rs.r = r;
}

// One sample of the lane's ray.  Returns true when the part is finished (hit or end reached).
// Precondition: rs.d < rs.d_end.
template <int CASTER, int EVAL, bool DIGEST>
__device__ __forceinline__ bool ray_step(const MapView& m, const CenterTables& ct, const EvalView& e,
                                         const D3& P0, const double step, RayState& rs,
                                         LaneCounters& lc) {
  constexpr bool ITER = (CASTER == A3D_CASTER_ITERATIVE);
  const double d_this = rs.d;
  const D3 Pd{__dadd_rn(P0.x, __dmul_rn(d_this, rs.dir.x)), __dadd_rn(P0.y, __dmul_rn(d_this, rs.dir.y)),
              __dadd_rn(P0.z, __dmul_rn(d_this, rs.dir.z))};
  rs.d = __dadd_rn(d_this, step);
  lc.n_samp++;
  const float pf[3] = {__double2float_rn(Pd.x), __double2float_rn(Pd.y), __double2float_rn(Pd.z)};
  int bi[3];
  float origin[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    bi[a] = grid_index(pf[a], m.bs_inv_f);
    origin[a] = origin_point(bi[a], m.bs_f);
  }
  const int slot = find_slot(m, bi[0], bi[1], bi[2]);
  const int st = voxel_state_in_block(m, ct, pf, bi, origin, slot);
  if (ITER || st != 0) {
    // getVoxelCenter identity (voxblox.cpp:66-81)
    const double pp[3] = {Pd.x, Pd.y, Pd.z};
    int vi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      vi[a] = grid_index(__double2float_rn(__dsub_rn(pp[a], static_cast<double>(origin[a]))),
                         m.vs_inv_f);
    bool canonical;
    const uint64_t key = make_key(m, ct, bi, vi, &canonical);
    if (key == kKeyOverflow) lc.inexact = 1;
    bool keep = key != rs.prev_key;
    rs.prev_key = key;
    D3 cc{0.0, 0.0, 0.0};
    const bool need_cc = DIGEST || e.bv_is_setup || EVAL == A3D_EVAL_FRONTIER || !canonical;
    if (need_cc) {
      double cp[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        cp[a] = (vi[a] >= -1 && vi[a] <= m.vps) ? ct.d[vi[a] + 1]
                                                : static_cast<double>(center_point(vi[a], m.vs_f));
      cc.x = __dadd_rn(static_cast<double>(origin[0]), cp[0]);
      cc.y = __dadd_rn(static_cast<double>(origin[1]), cp[1]);
      cc.z = __dadd_rn(static_cast<double>(origin[2]), cp[2]);
    }
    if (e.bv_is_setup) keep = keep && bv_contains(e, cc);  // simulated_sensor_evaluator.cpp:50-57
    int cat = kCatNone;
    unsigned long long h = 0;
    if (keep) {
      lc.n_vis++;
      const int cs = centre_state(m, ct, cc, slot, vi);
      if (EVAL == A3D_EVAL_VOXEL_TYPE) {
        cat = cs & 3;  // retained entries are always inside the bounding volume
      } else if (!(cs & 4)) {
        if (EVAL == A3D_EVAL_NAIVE) {
          cat = 0;
        } else if (is_frontier_voxel(m, ct, e, cc)) {
          cat = 0;
        }
      }
      if (cat != kCatNone) lc.cnt[cat]++;
      if (DIGEST) {
        h = digest_entry(cc);
        lc.digest += h;
      }
    }
    if (rs.first_key == kNoKey) {
      rs.first_key = key;
      rs.info = (keep ? cat : kCatNone) | 8 | (keep ? 16 : 0);
      rs.first_hash = h;
    }
  }
  if (st == 0) {
    rs.hit_d = d_this;
    return true;
  }
  return !(rs.d < rs.d_end);
}

struct WaveScratch {
  uint32_t* table;     // per CTA: n_rays entries (writer+1) << 8 | uint8(value)
  uint8_t* start_s;    // per CTA: original start section of rays cast by the scheduler, 0xFF none
  Records rec;         // per CTA
  size_t rays_stride;  // n_rays rounded up (elements) per CTA
};

__device__ __forceinline__ void store_record(const Records& rec, size_t base, int part, int n_rays,
                                             const RayState& rs, bool digest) {
  const size_t at = base * 2 + static_cast<size_t>(part) * n_rays + rs.r;
  rec.first_key[at] = rs.first_key;
  rec.last_key[at] = rs.prev_key;
  rec.info[at] = static_cast<uint8_t>(rs.info);
  if (digest) rec.first_hash[at] = rs.first_hash;
}

constexpr int kPendCap = 96;     // leaf ring capacity (top-up adds <= 32 while count <= 64)
constexpr int kMarkCap = 1024;   // max rays of one anti-diagonal (min(c_res_x, c_res_y))

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

  __shared__ CenterTables ct;
  __shared__ int s_seg;
  __shared__ volatile int s_done_diag;        // anti-diagonals fully processed by the scheduler
  __shared__ unsigned s_next_ray;             // simple caster: next unassigned ray
  __shared__ unsigned s_pending[2][kPendCap];  // per-warp ring of rays waiting for a lane
  __shared__ unsigned s_mark_list[kMarkCap];  // scheduler: marking rays of the current diagonal
  __shared__ unsigned s_cnt[6], s_nvis, s_nsamp, s_inexact;
  __shared__ unsigned long long s_digest;
  init_center_tables(m, &ct);

  const size_t cta_base = static_cast<size_t>(blockIdx.x) * ws.rays_stride;
  volatile uint32_t* const table = ws.table + cta_base;
  volatile uint8_t* const start_s = ws.start_s + cta_base;
  const double step = c.ray_step;
  const int n_diag = c.res_x + c.res_y - 1;

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
      // per-view init: ray table, start sections, record flags (after the previous view's fix-up)
      __syncthreads();
      for (int r = threadIdx.x; r < n_rays; r += kWaveThreads) {
        if (ITER) {
          table[r] = 0u;
          start_s[r] = 0xFF;
        }
        ws.rec.info[cta_base * 2 + r] = kCatNone;
        ws.rec.info[cta_base * 2 + n_rays + r] = kCatNone;
      }
      if (threadIdx.x == 0) {
        s_done_diag = 0;
        s_next_ray = 0;
      }
      __syncthreads();

      if (ITER && warp == 0) {
        // ---------------- scheduler: sections before the last one + all marks -----------------
        const double d_end = n > 1 ? c.split[n - 1] : 0.0;
        for (int dg = 0; n > 1 && dg < n_diag; ++dg) {
          const int i_lo = max(0, dg - (c.res_y - 1)), i_hi = min(c.res_x - 1, dg);
          // gather the rays of this diagonal that start before the last section (scan order)
          int n_mark = 0;
          for (int base = i_lo; base <= i_hi; base += 32) {
            const int i = base + static_cast<int>(lane);
            int s = -1, r = 0;
            if (i <= i_hi) {
              r = i * c.res_y + (dg - i);
              s = static_cast<int>(static_cast<int8_t>(table[r] & 0xFFu));
            }
            const bool marking = s >= 0 && s < n - 1;
            const unsigned mm = __ballot_sync(kFull, marking);
            if (marking)
              s_mark_list[n_mark + __popc(mm & lt_mask)] = (static_cast<unsigned>(r) << 8) | s;
            n_mark += __popc(mm);
          }
          __syncwarp();
          for (int q0 = 0; q0 < n_mark; q0 += 32) {
            const int nb = min(32, n_mark - q0);
            RayState rs;
            rs.r = -1;
            rs.hit_d = -1.0;
            int s = 0;
            if (static_cast<int>(lane) < nb) {
              const unsigned it = s_mark_list[q0 + lane];
              s = static_cast<int>(it & 0xFF);
              ray_begin(rs, c, q, static_cast<int>(it >> 8), c.split[s], d_end);
              bool done = !(rs.d < rs.d_end);
              while (!done) done = ray_step<CASTER, EVAL, DIGEST>(m, ct, e, P0, step, rs, lc);
              store_record(ws.rec, cta_base, 0, n_rays, rs, DIGEST);
              start_s[rs.r] = static_cast<uint8_t>(s);
            }
            __syncwarp();
            // markNeighboringRays of each ray of the batch, in scan order (DESIGN.md "ray table")
            for (int qi = 0; qi < nb; ++qi) {
              const int rr = __shfl_sync(kFull, rs.r, qi);
              const int ss = __shfl_sync(kFull, s, qi);
              const double hd = __shfl_sync(kFull, rs.hit_d, qi);
              const bool hit = hd >= 0.0;
              int seg_h = ss;
              if (hit) {
                while (!(hd < c.split[seg_h + 1])) ++seg_h;
              }
              const int t_last = hit ? seg_h : n - 2;
              const int w0 = 1 << (n - 1 - ss);
              const int ri = rr / c.res_y, rj = rr - ri * c.res_y;
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
          }
          __threadfence_block();
          if (lane == 0) s_done_diag = dg + 1;
        }
      } else {
        // ---------------- leaf: persistent lanes over the rays that are ready -------------------
        // Iterative: last section of every ray whose table entry reads n-1, discovered diagonal by
        // diagonal behind the scheduler.  Simple: every ray, full length.
        unsigned* const pending = s_pending[warp];
        int pend_head = 0, pend_count = 0;  // ring state (warp-uniform)
        int cur_diag = 0, cur_i = 0;        // iterative: scan position
        bool source_done = false;
        const double d_end = c.ray_length;
        RayState rs;
        rs.r = -1;
        bool active = false;
        for (;;) {
          const unsigned idle = __ballot_sync(kFull, !active);
          if (idle) {
            // ---- top up the pending ring ----
            while (!source_done && pend_count <= kPendCap - 32 && pend_count < __popc(idle)) {
              if (ITER) {
                if (cur_diag >= n_diag) {
                  source_done = true;
                  break;
                }
                if (n > 1 && s_done_diag <= cur_diag) break;  // entries of this diagonal not final
                const int i_lo = max(0, cur_diag - (c.res_y - 1)), i_hi = min(c.res_x - 1, cur_diag);
                const int i = i_lo + cur_i + static_cast<int>(lane);
                bool ready = false;
                int rr = 0;
                if (i <= i_hi) {
                  rr = i * c.res_y + (cur_diag - i);
                  ready = static_cast<int>(static_cast<int8_t>(table[rr] & 0xFFu)) == n - 1;
                }
                const unsigned mm = __ballot_sync(kFull, ready);
                if (ready) pending[(pend_head + pend_count + __popc(mm & lt_mask)) % kPendCap] = rr;
                pend_count += __popc(mm);
                cur_i += 32;
                if (i_lo + cur_i > i_hi) {
                  cur_i = 0;
                  ++cur_diag;
                }
              } else {
                unsigned first = 0;
                if (lane == 0) first = atomicAdd(&s_next_ray, 32u);
                first = __shfl_sync(kFull, first, 0);
                if (first >= static_cast<unsigned>(n_rays)) {
                  source_done = true;
                  break;
                }
                const int cnt = min(32, n_rays - static_cast<int>(first));
                if (static_cast<int>(lane) < cnt)
                  pending[(pend_head + pend_count + lane) % kPendCap] = first + lane;
                pend_count += cnt;
              }
              __syncwarp();
            }
            // ---- hand rays to idle lanes ----
            const int rank = __popc(idle & lt_mask);
            if (!active && rank < pend_count) {
              const int r = static_cast<int>(pending[(pend_head + rank) % kPendCap]);
              double d0 = 0.0;
              if (ITER) {
                const int s0 = start_s[r] == 0xFF ? n - 1 : start_s[r];
                const int k0 = s0 == n - 1 ? 0 : c.seg_end[s0 * n + (n - 2)];
                d0 = k0 < c.kcount[s0] ? c.dseq[static_cast<size_t>(s0) * c.kmax + k0] : d_end;
              }
              ray_begin(rs, c, q, r, d0, d_end);
              active = rs.d < rs.d_end;
              if (!active) store_record(ws.rec, cta_base, 1, n_rays, rs, DIGEST);
            }
            const int taken = min(__popc(idle), pend_count);
            pend_head = (pend_head + taken) % kPendCap;
            pend_count -= taken;
            __syncwarp();
            if (__ballot_sync(kFull, active) == 0) {
              if (source_done && pend_count == 0) break;
              if (pend_count == 0) __nanosleep(64);
              continue;
            }
          }
          if (active) {
            if (ray_step<CASTER, EVAL, DIGEST>(m, ct, e, P0, step, rs, lc)) {
              store_record(ws.rec, cta_base, 1, n_rays, rs, DIGEST);
              active = false;
            }
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
            lc.n_vis--;
            if ((info & 7) != kCatNone) lc.cnt[info & 7]--;
            if (DIGEST) lc.digest -= ws.rec.first_hash[at];
          }
          if (em) carry_key = __shfl_sync(kFull, lk, 31 - __clz(em));
        }
      }
    }  // views

    // ---------------- per-segment reduction (lane counters are modular: fix-up may decrement) ----
    if (threadIdx.x == 0) {
      for (int i = 0; i < 6; ++i) s_cnt[i] = 0;
      s_nvis = 0;
      s_nsamp = 0;
      s_inexact = 0;
      s_digest = 0;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const unsigned t = __reduce_add_sync(kFull, lc.cnt[i]);
      if (lane == 0 && t) atomicAdd(&s_cnt[i], t);
    }
    {
      const unsigned tv = __reduce_add_sync(kFull, lc.n_vis);
      const unsigned tsmp = __reduce_add_sync(kFull, lc.n_samp);
      const unsigned tin = __reduce_or_sync(kFull, lc.inexact);
      unsigned long long dg = lc.digest;
      if (DIGEST) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dg += __shfl_xor_sync(kFull, dg, off);
      }
      if (lane == 0) {
        atomicAdd(&s_nvis, tv);
        atomicAdd(&s_nsamp, tsmp);
        atomicOr(&s_inexact, tin);
        if (DIGEST) atomicAdd(&s_digest, dg);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // VoxelTypeEvaluator's switch without breaks (voxel_type_evaluator.cpp:69-85)
      double g = 0.0;
      if (EVAL == A3D_EVAL_VOXEL_TYPE) {
        const double n_unk = static_cast<double>(s_cnt[2]);
        const double n_free = static_cast<double>(s_cnt[1]) + n_unk;
        const double n_occ = static_cast<double>(s_cnt[0]) + n_free;
        g += n_unk * e.gains[0];
        g += n_free * e.gains[2];
        g += n_occ * e.gains[1];
      } else {
        g = static_cast<double>(s_cnt[0]);
      }
      b.gains[seg] = g;
      if (b.n_visible) b.n_visible[seg] = s_nvis;
      if (b.n_samples) b.n_samples[seg] = s_nsamp;
      if (DIGEST && b.set_digests) b.set_digests[seg] = s_digest;
      if (s_inexact) atomicOr(&inexact_flags[seg >> 5], 1u << (seg & 31));
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
    k_cast_wave<CASTER, EVAL, true><<<grid, kWaveThreads, 0, st>>>(m, c, e, b, ws, inexact);
  else
    k_cast_wave<CASTER, EVAL, false><<<grid, kWaveThreads, 0, st>>>(m, c, e, b, ws, inexact);
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

// scratch: wave_scratch_bytes() bytes, carved here.  b.digests receives the ORDER-INDEPENDENT digest
// (sum of digest_entry over the stored list, mod 2^64).  inexact: one bit per segment, set when a
// voxel key overflowed the 20-bit packing (the host re-evaluates those with k_cast_gain).
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

}  // namespace a3d
