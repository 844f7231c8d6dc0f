// a3d_keys.cuh — 64-bit identity of a visible-voxel list entry.
//
// The reference stores a list entry as the voxel-centre double triple getVoxelCenter builds
// (voxblox.cpp:66-81) and compares entries with == (std::unique camera_model.cpp:58, std::find
// simulated_sensor_evaluator.cpp:66-70).  On the device an entry is the integer triple
// g[k] = block[k]*vps + voxel[k]: two entries have equal centre doubles iff their keys are equal.
// voxblox's float block-index rounding can name a voxel one step outside its block (voxel index -1 or
// vps, "artefact index"); such an entry is renamed to the canonical voxel when the two centre doubles
// coincide, else it keeps its raw index and carries a per-axis flag, so the rule above still holds.
// Packed: 3 x 20 bit (biased) + 3 flag bits at 60..62; keys do not fit for |g| >= 2^19 (poses beyond
// +-52 km at 0.1 m), which callers detect through kKeyOverflow / *inexact.
#pragma once

#include "a3d_device.cuh"

namespace a3d {

constexpr uint64_t kNoKey = ~0ull;  // "no entry"
constexpr uint64_t kKeyOverflow = 1ull << 63;
constexpr int kKeyBits = 20, kKeyBias = 1 << 19;
constexpr int kRawAxisBit = 1 << 30;

// Identity of an emitted list entry: g[k] = b[k]*vps + v[k]; bit 30 of g[k] set: axis k keeps a raw
// artefact voxel index (v = -1 or vps) whose centre double differs from the canonical voxel's
// (valid g are < 2^19 in magnitude, larger ones flag the segment for the scan-order kernel).
struct Ident {
  int g0, g1, g2;
};

// Artefact voxel indices (getGridIndexFromPoint lands one voxel outside the block because of the
// float block-index rounding): per axis, name the voxel canonically when its centre double
// (voxblox.cpp:72-79) equals the canonical voxel's centre double, else keep it raw and flag it.
static __device__ __noinline__ Ident make_ident_slow(const MapView& m, const CenterTables& ct, int b0, int b1,
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

// Key of the entry getVoxelCenter produced from block b / unclamped voxel index v (any vps).
__device__ __forceinline__ uint64_t key_from_block_voxel(const MapView& m, const CenterTables& ct, const int (&b)[3],
                                                         const int (&v)[3], unsigned* inexact) {
  const unsigned vmax = static_cast<unsigned>(m.vps);
  const bool canonical = (static_cast<unsigned>(v[0]) < vmax) & (static_cast<unsigned>(v[1]) < vmax) &
                         (static_cast<unsigned>(v[2]) < vmax);
  const Ident id = canonical ? Ident{b[0] * m.vps + v[0], b[1] * m.vps + v[1], b[2] * m.vps + v[2]}
                             : make_ident_slow(m, ct, b[0], b[1], b[2], v[0], v[1], v[2], inexact);
  return pack_key(id, inexact);
}

__device__ __forceinline__ int floor_div(int a, int d) {
  const int q = a / d;
  return (a % d != 0 && ((a < 0) != (d < 0))) ? q - 1 : q;
}

// Block / voxel index of axis a of a key (voxel may be -1 or vps for a raw artefact axis).
__device__ __forceinline__ void key_axis(const MapView& m, uint64_t key, int a, int* b, int* v) {
  const int g = static_cast<int>((key >> (kKeyBits * a)) & ((1u << kKeyBits) - 1u)) - kKeyBias;
  int bb = m.vps_shift >= 0 ? (g >> m.vps_shift) : floor_div(g, m.vps);
  int vv = g - bb * m.vps;
  if ((key >> (60 + a)) & 1u) {
    // raw artefact index on this axis (make_ident_slow): voxel -1 of block b+1, or voxel vps of block b-1
    if (vv == m.vps - 1) {
      bb += 1;
      vv = -1;
    } else {
      bb -= 1;
      vv = m.vps;
    }
  }
  *b = bb;
  *v = vv;
}

// The centre double triple the reference holds for this entry (voxblox.cpp:72-79).
__device__ __forceinline__ D3 key_centre(const MapView& m, const CenterTables& ct, uint64_t key) {
  double c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int b, v;
    key_axis(m, key, a, &b, &v);
    c[a] = __dadd_rn(static_cast<double>(origin_point(b, m.bs_f)), ct.d[v + 1]);
  }
  return D3{c[0], c[1], c[2]};
}

// Meta word of the entry's voxel (centre state, observed flag, frontier bits) straight from the key:
// no index arithmetic on doubles.  False for raw artefact entries (evaluate those literally from the
// centre).
__device__ __forceinline__ bool key_word(const MapView& m, uint64_t key, uint32_t* word_out) {
  if ((key >> 60) & 7u) return false;
  const int g0 = static_cast<int>(key & ((1u << kKeyBits) - 1u)) - kKeyBias;
  const int g1 = static_cast<int>((key >> kKeyBits) & ((1u << kKeyBits) - 1u)) - kKeyBias;
  const int g2 = static_cast<int>((key >> (2 * kKeyBits)) & ((1u << kKeyBits) - 1u)) - kKeyBias;
  uint32_t word = kWordNoBlock;
  if (m.vol) {
    const uint32_t u0 = static_cast<uint32_t>(g0 - m.vol_lo[0]), u1 = static_cast<uint32_t>(g1 - m.vol_lo[1]),
                   u2 = static_cast<uint32_t>(g2 - m.vol_lo[2]);
    if ((u0 < m.vol_n[0]) & (u1 < m.vol_n[1]) & (u2 < m.vol_n[2])) word = __ldg(m.vol + vol_offset(m, u0, u1, u2));
  } else {
    const int b0 = m.vps_shift >= 0 ? (g0 >> m.vps_shift) : floor_div(g0, m.vps);
    const int b1 = m.vps_shift >= 0 ? (g1 >> m.vps_shift) : floor_div(g1, m.vps);
    const int b2 = m.vps_shift >= 0 ? (g2 >> m.vps_shift) : floor_div(g2, m.vps);
    const int slot = find_slot(m, b0, b1, b2);
    if (slot >= 0)
      word = __ldg(m.mword + static_cast<size_t>(slot) * m.nv +
                   mword_offset(m, g0 - b0 * m.vps, g1 - b1 * m.vps, g2 - b2 * m.vps));
  }
  *word_out = word;
  return true;
}

}  // namespace a3d
