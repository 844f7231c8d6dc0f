// a3d_sample.cuh — per-sample device code shared by the throughput kernels (a3d_wave.cu: one kernel
// per batch; a3d_split.cu: the same work as three kernels): the literal evaluation of the samples
// the certified index path cannot decide, and the bounding-volume test on voxel coordinates.
#pragma once

#include "a3d_kernels.cuh"
#include "a3d_keys.cuh"

namespace a3d {
namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kCatNone = 7;
constexpr int kNoIdent = INT32_MIN;
constexpr int kFieldBits = 21;  // a3d_wave.cu: three 21-bit counters per lane (BvGrid::inc)

// Bounding-volume test on integer voxel coordinates: for an axis-aligned volume, centre(g) inside
// [bv_min, bv_max] <=> lo <= g <= hi per axis (the centre double is monotone in g).  Built once per
// CTA from the literal centre arithmetic; fast == 0: use bv_contains() on the centre itself.
struct BvGrid {
  int lo[3], hi[3];
  int fast;
  int mode;  // 0: no volume, 1: test on voxel coordinates (lo/hi), 2: bv_contains() on the centre
  // what one retained entry of fold category 0..2 adds to LaneCounters::packed; [3]: nothing
  unsigned long long inc[4];
};

__device__ __forceinline__ double centre_of(const MapView& m, const CenterTables& ct, int g) {
  const int b = g >> m.vps_shift, v = g & (m.vps - 1);
  return __dadd_rn(static_cast<double>(origin_point(b, m.bs_f)), ct.d[v + 1]);
}

// threads 0..2, one axis each
__device__ void init_bv_grid(const MapView& m, const CenterTables& ct, const EvalView& e, BvGrid* g) {
  const int a = threadIdx.x;
  if (a == 0) g->fast = 1;
  __syncwarp();
  if (a >= 3) return;
  const bool identity = e.bv_q.w == 1.0 && e.bv_q.x == 0.0 && e.bv_q.y == 0.0 && e.bv_q.z == 0.0;
  bool ok = e.bv_is_setup && identity && m.vps_shift >= 0;
  int lo = INT32_MIN, hi = INT32_MAX;
  if (ok) {
    const double lim = 1048576.0;  // certified samples have |g| < 2^21
    const double tl = e.bv_min[a] / m.c_vs, th = e.bv_max[a] / m.c_vs;
    if (tl != tl || th != th) ok = false;
    if (ok && tl > -lim) {
      // smallest g whose centre is not below bv_min
      int q = static_cast<int>(floor(fmin(tl, lim))) - 3;
      ok = centre_of(m, ct, q) < e.bv_min[a];
      int it = 0;
      while (ok && centre_of(m, ct, q) < e.bv_min[a]) {
        ++q;
        if (++it > 8) ok = false;
      }
      lo = q;
    }
    if (ok && th < lim) {
      // largest g whose centre is not above bv_max
      int q = static_cast<int>(floor(fmax(th, -lim))) + 3;
      ok = centre_of(m, ct, q) > e.bv_max[a];
      int it = 0;
      while (ok && centre_of(m, ct, q) > e.bv_max[a]) {
        --q;
        if (++it > 8) ok = false;
      }
      hi = q;
    }
  }
  g->lo[a] = lo;
  g->hi[a] = hi;
  if (!ok) g->fast = 0;
  g->inc[a] = 1ull << (kFieldBits * a);
  if (a == 0) g->inc[3] = 0;
  __syncwarp(0x7u);  // threads 0..2 (the others left above)
  if (a == 0) g->mode = !e.bv_is_setup ? 0 : (g->fast ? 1 : 2);
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

// State of a certified sample whose interpolation cell straddles the occupied threshold: the
// literal trilinear interpolation over the cell's 8 corners.
__device__ __noinline__ int mixed_state(const MapView& m, const CenterTables& ct, float pf0, float pf1,
                                        float pf2, int g0, int g1, int g2, int sc) {
  const float pf[3] = {pf0, pf1, pf2};
  const int vmask = m.vps - 1;
  const int b[3] = {g0 >> m.vps_shift, g1 >> m.vps_shift, g2 >> m.vps_shift};
  const int v[3] = {g0 & vmask, g1 & vmask, g2 & vmask};
  float origin[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) origin[a] = origin_point(b[a], m.bs_f);
  return mixed_cell_state(m, ct, pf, b, origin, find_slot(m, b[0], b[1], b[2]), v, sc);
}

// One sample, literally as the reference computes it (voxblox index arithmetic on the float point
// for the state, on the double point for the voxel centre).  Used for the samples the margin test
// of ray_step() cannot certify.
struct ExactOut {
  SampleInfo si;
  D3 cc;  // voxel centre (voxblox.cpp:72-79)
};
__device__ __noinline__ void sample_exact(const MapView& m, const CenterTables& ct, double p0, double p1,
                                          double p2, ExactOut* out, unsigned* inexact) {
  const double pp[3] = {p0, p1, p2};
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
  if (!(in_range & same) | (cls == kCellMixed))
    si = sample_slow(m, ct, pf[0], pf[1], pf[2], bi[0], bi[1], bi[2], slot, vi[0], vi[1], vi[2], inexact);
  out->si = si;
  double cp[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    cp[a] = (vi[a] >= -1 && vi[a] <= m.vps) ? ct.d[vi[a] + 1]
                                            : static_cast<double>(center_point(vi[a], m.vs_f));
  out->cc.x = __dadd_rn(static_cast<double>(origin[0]), cp[0]);
  out->cc.y = __dadd_rn(static_cast<double>(origin[1]), cp[1]);
  out->cc.z = __dadd_rn(static_cast<double>(origin[2]), cp[2]);
}

}  // namespace
}  // namespace a3d
