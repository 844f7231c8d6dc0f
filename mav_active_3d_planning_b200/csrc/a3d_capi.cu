// a3d_capi.cu — host side of the C ABI declared in include/a3d.h: context, GPU block-hash mirror of
// the voxblox layers (slot allocator + open-addressing table + voxel slab), caster/evaluator
// objects with their host-precomputed tables, and the launch plumbing of the hot call.
//
// No CPU compute path exists here: every query and every gain goes through the kernels in
// a3d_kernels.cu; a missing/unusable device is an error (A3D_ERR_CUDA).
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/a3d.h"
#include "a3d_kernels.cuh"

using namespace a3d;

namespace {

thread_local std::string g_create_error;

struct Key3 {
  int32_t x, y, z;
  bool operator==(const Key3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct Key3Hash {
  size_t operator()(const Key3& k) const {
    return (static_cast<size_t>(static_cast<uint32_t>(k.x)) * 73856093u) ^
           (static_cast<size_t>(static_cast<uint32_t>(k.y)) * 19349663u) ^
           (static_cast<size_t>(static_cast<uint32_t>(k.z)) * 83492791u);
  }
};

inline uint32_t host_hash3(int x, int y, int z) {  // must equal a3d::hash3
  uint32_t h = static_cast<uint32_t>(x) * 73856093u ^ static_cast<uint32_t>(y) * 19349663u ^
               static_cast<uint32_t>(z) * 83492791u;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

template <typename T>
struct DevBuf {  // grow-only device scratch
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    const size_t want = std::max(n, cap + cap / 2);
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// Grow-only pinned host staging for the small index lists of a map update: copies from it are truly
// asynchronous (a pageable source would make every cudaMemcpyAsync synchronise the stream first).
// Reuse is fenced by an event recorded after the last copy that reads the buffer.
struct PinBuf {
  int32_t* p = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  bool pending = false;
  cudaError_t reserve(size_t n) {
    if (pending) {
      cudaError_t e = cudaEventSynchronize(ev);
      if (e != cudaSuccess) return e;
      pending = false;
    }
    if (n <= cap) return cudaSuccess;
    const size_t want = std::max<size_t>(std::max(n, cap + cap / 2), 1024);
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), want * sizeof(int32_t), cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  cudaError_t mark(cudaStream_t st) {
    if (!ev) {
      cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    pending = true;
    return cudaEventRecord(ev, st);
  }
  void release() {
    if (pending && ev) cudaEventSynchronize(ev);
    if (p) cudaFreeHost(p);
    if (ev) cudaEventDestroy(ev);
    p = nullptr;
    ev = nullptr;
    cap = 0;
    pending = false;
  }
};

}  // namespace

struct a3d_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_stage[6] = {};  // kernel chain (a3d_split.cu): after each of its kernels, last chunk
  int n_stage = 0;
  bool timed = false;
  uint64_t launches = 0;
  mutable std::string err;

  // map
  bool configured = false;
  float vs_f = 0, vs_inv_f = 0, bs_f = 0, bs_inv_f = 0;
  int vps = 0, nv = 0;
  std::unordered_map<Key3, int32_t, Key3Hash> slots;
  std::vector<int32_t> free_slots;
  int32_t next_slot = 0;
  std::vector<HashEntry> h_table;
  DevBuf<HashEntry> d_table;
  std::vector<Key3> slot_key;  // slot → block index (valid for allocated slots)
  float* d_dist = nullptr;     // padded cubes, P^3 floats per slot
  uint32_t* d_mword = nullptr; // meta words, vps^3 per slot
  std::vector<int32_t> h_dir;  // dense directory over [dir_lo, dir_lo + dir_n)
  DevBuf<int32_t> d_dir;
  int32_t dir_lo[3] = {0, 0, 0};
  uint32_t dir_n[3] = {0, 0, 0};
  bool dir_enabled = false;
  uint32_t* d_vol = nullptr;   // dense meta volume over the directory's box (a3d_device.cuh)
  size_t vol_words = 0;        // allocated size
  bool vol_stale = true;       // box changed (or never built): rebuilt lazily by vol_ensure()
  bool vol_on = false;         // d_vol holds the current map
  float* d_weight = nullptr;   // vps^3 floats per slot
  float* d_tsdf = nullptr;     // vps^3 floats per slot: TSDF distances (getVoxelDistance), allocated on first upload
  size_t slab_slots = 0;
  bool has_weight = false;
  int P = 0, P3 = 0;

  // staging / scratch
  DevBuf<float> st_dist, st_w;
  DevBuf<uint8_t> st_obs;
  DevBuf<int32_t> st_slots, st_aff_slots, st_aff_nbr, st_aff_idx, st_unalloc_idx;
  PinBuf pin_slots, pin_aff, pin_unalloc;
  DevBuf<double> sc_pts, sc_out;
  DevBuf<uint8_t> sc_u8;
  DevBuf<double> sc_pos, sc_quat, sc_gains, sc_centres, sc_origin;
  DevBuf<int32_t> sc_first;
  DevBuf<unsigned long long> sc_nvis, sc_nsamp, sc_dig, sc_cnt;
  DevBuf<int8_t> sc_tables;
  DevBuf<unsigned int> sc_counter, sc_inexact;
  DevBuf<uint8_t> sc_wave;
  uint64_t config_gen = 0;  // bumped when a3d_map_config changes the geometry: older casters/evals are stale
  // 0 auto, 1 scan-order kernel (k_cast_gain), 2 wavefront kernel (k_cast_wave), 3 kernel chain (a3d_split.cu)
  // stored-list calls: the next chain cast also leaves the ray parts' list offsets in the scratch
  // (want_part_offsets), and whether the last cast did for its whole batch (launch_cast_split_direct may follow)
  bool want_part_offsets = false, part_offsets_valid = false;
  int opt_cast_kernel = 0;
  int last_cast_kernel = 0;
  int opt_wave_cta = 0;  // 0 auto (wave_warps_auto), 1 always 2 warps, 2 always 8, 3 always 4

  int fail(int code, const std::string& msg) const {
    err = msg;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) const {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return A3D_ERR_CUDA;
  }
  MapView view() const {
    MapView m{};
    m.table = d_table.p;
    m.dist = d_dist;
    m.weight = has_weight ? d_weight : nullptr;
    m.mword = d_mword;
    m.dir = dir_enabled ? d_dir.p : nullptr;
    for (int k = 0; k < 3; ++k) {
      m.dir_lo[k] = dir_lo[k];
      m.dir_n[k] = dir_n[k];
    }
    m.P = P;
    m.P3 = P3;
    m.n_slots = static_cast<int32_t>(slab_slots);
    m.brick_nx = (vps % 4 == 0) ? vps / 4 : 0;
    m.vps_shift = -1;
    for (int k = 0; k < 7; ++k)
      if ((1 << k) == vps) m.vps_shift = k;
    // 16 float roundings of magnitude <= vps+1 voxels plus voxblox's 1e-6 index epsilons (block
    // epsilon counts vps times in voxel units); see ray_step() in a3d_wave.cu
    m.fast_c0 = static_cast<float>((vps + 1) * (16.0 * 5.97e-8 + 1.1e-6));
    m.vol = (vol_on && !vol_stale) ? d_vol : nullptr;
    for (int k = 0; k < 3; ++k) {
      m.vol_lo[k] = dir_lo[k] * vps;
      m.vol_n[k] = dir_n[k] * static_cast<uint32_t>(vps);
    }
    m.vol_bx = m.vol_n[0] >> 2;
    m.vol_by = m.vol_n[1] >> 2;
    m.mask = static_cast<uint32_t>(h_table.size() - 1);
    m.vps = vps;
    m.nv = nv;
    m.vs_f = vs_f;
    m.vs_inv_f = vs_inv_f;
    m.bs_f = bs_f;
    m.bs_inv_f = bs_inv_f;
    m.c_vs = static_cast<double>(vs_f);
    return m;
  }
};

struct a3d_caster {
  a3d_ctx* ctx;
  uint64_t config_gen = 0;
  a3d_caster_params p;
  a3d_caster_info info;
  std::vector<double> h_dirs;
  double* d_dirs = nullptr;
  double* d_dseq = nullptr;
  int32_t* d_kcount = nullptr;
  int32_t* d_seg_end = nullptr;
  int32_t* d_units = nullptr;  // unit_diag[n_units] followed by unit_i0[n_units]
  CasterView view{};
};

struct a3d_eval {
  a3d_ctx* ctx;
  uint64_t config_gen = 0;
  a3d_eval_params p;
  EvalView view{};
};

// Device-resident stored lists (packed voxel keys, a3d_keys.cuh): every a3d_compute_gains_stored call
// owns one arena; a key maps to a slice of an arena; an arena is freed when its last user is replaced
// or erased.  Hash sets of the lists (clear_from_parents membership tests) live in arenas of their own.
struct StoreArena {
  unsigned long long* p = nullptr;  // list keys, or hash-set tables
  size_t entries = 0;               // 8-byte words
  int live = 0;
  bool pooled = false;              // from the stream-ordered pool (else plain cudaMalloc)
};
struct StoreEntry {
  int arena;
  long long offset, count;
  uint64_t parent = 0;     // store key of the segment's parent (0: none)
  int set_arena = -1;      // hash set of the stored list: arena, offset (words), size - 1
  long long set_off = 0;
  uint32_t set_mask = 0;
  bool set_stale = true;   // the list changed since the set was built (refold erased entries)
};
struct a3d_store {
  a3d_ctx* ctx;
  std::vector<StoreArena> arenas;
  std::unordered_map<uint64_t, StoreEntry> lists;
  DevBuf<long long> d_off, d_cnt;
  DevBuf<double> d_gain, d_origin;
  DevBuf<unsigned long long> d_u64;   // n_stored | digests | set digests of a pipeline launch
  DevBuf<int> d_anc_first;
  DevBuf<unsigned long long*> d_tab;  // ancestor tables / tables to build
  DevBuf<uint32_t> d_mask;
  // Arenas come from the device's stream-ordered pool (cudaMallocAsync on the context's stream, the only
  // stream that touches them): a released arena is handed out again without a trip to the driver and
  // without the device-wide synchronisation of cudaFree — stored lists are replaced every time a planner
  // re-evaluates its tree.
  void free_arena(StoreArena& a);
  void unref(int arena) {
    if (arena < 0) return;
    StoreArena& a = arenas[arena];
    if (--a.live == 0 && a.p) free_arena(a);
  }
  void drop(uint64_t key) {
    auto it = lists.find(key);
    if (it == lists.end()) return;
    unref(it->second.arena);
    unref(it->second.set_arena);
    lists.erase(it);
  }
  int new_arena(size_t words, cudaError_t* err);
};

void a3d_store::free_arena(StoreArena& a) {
  if (!a.pooled || cudaFreeAsync(a.p, ctx->stream) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(a.p);
  }
  a.p = nullptr;
  a.entries = 0;
}
int a3d_store::new_arena(size_t words, cudaError_t* err) {
  StoreArena a;
  a.entries = words;
  void* p = nullptr;
  *err = cudaMallocAsync(&p, std::max<size_t>(words, 1) * 8, ctx->stream);
  a.pooled = *err == cudaSuccess;
  if (!a.pooled) {  // (a device or driver mode without memory pools)
    cudaGetLastError();
    *err = cudaMalloc(&p, std::max<size_t>(words, 1) * 8);
  }
  if (*err != cudaSuccess) return -1;
  a.p = static_cast<unsigned long long*>(p);
  arenas.push_back(a);
  return static_cast<int>(arenas.size()) - 1;
}

#define CK(call)                                          \
  do {                                                    \
    cudaError_t _e = (call);                              \
    if (_e != cudaSuccess) return ctx->cuda_fail(_e, #call); \
  } while (0)

namespace {

void table_insert(std::vector<HashEntry>& t, int x, int y, int z, int slot) {
  const uint32_t mask = static_cast<uint32_t>(t.size() - 1);
  uint32_t h = host_hash3(x, y, z) & mask;
  while (t[h].slot >= 0) {
    if (t[h].x == x && t[h].y == y && t[h].z == z) {
      t[h].slot = slot;
      return;
    }
    h = (h + 1) & mask;
  }
  t[h] = HashEntry{x, y, z, slot};
}

void table_rebuild(a3d_ctx* ctx, size_t min_entries) {
  size_t cap = 1024;
  while (cap < 4 * min_entries) cap <<= 1;  // load factor <= 0.25
  ctx->h_table.assign(cap, HashEntry{0, 0, 0, -1});
  for (const auto& kv : ctx->slots)
    table_insert(ctx->h_table, kv.first.x, kv.first.y, kv.first.z, kv.second);
}

// Dense directory over the bounding box of the allocated block indices (grown with slack so that a
// moving frontier does not rebuild it every step).  Disabled when the box exceeds 64 Mi entries.
void dir_rebuild(a3d_ctx* ctx) {
  ctx->dir_enabled = false;
  if (ctx->slots.empty()) {
    ctx->h_dir.clear();
    ctx->vol_stale = true;
    return;
  }
  int32_t lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  for (const auto& kv : ctx->slots) {
    const int32_t c[3] = {kv.first.x, kv.first.y, kv.first.z};
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(lo[k], c[k]);
      hi[k] = std::max(hi[k], c[k]);
    }
  }
  bool covered = !ctx->h_dir.empty();
  for (int k = 0; k < 3 && covered; ++k)
    covered = lo[k] >= ctx->dir_lo[k] &&
              static_cast<int64_t>(hi[k]) < static_cast<int64_t>(ctx->dir_lo[k]) + ctx->dir_n[k];
  if (!covered) {
    ctx->vol_stale = true;
    double total = 1.0;
    for (int k = 0; k < 3; ++k) {
      const int64_t ext = static_cast<int64_t>(hi[k]) - lo[k] + 1;
      const int64_t slack = std::max<int64_t>(2, ext / 8);
      ctx->dir_lo[k] = static_cast<int32_t>(std::max<int64_t>(INT32_MIN / 2, lo[k] - slack));
      ctx->dir_n[k] = static_cast<uint32_t>(ext + 2 * slack);
      total *= static_cast<double>(ctx->dir_n[k]);
    }
    if (total > 64.0 * 1024 * 1024) {
      ctx->h_dir.clear();
      return;
    }
  }
  const size_t n = static_cast<size_t>(ctx->dir_n[0]) * ctx->dir_n[1] * ctx->dir_n[2];
  ctx->h_dir.assign(n, -1);
  for (const auto& kv : ctx->slots) {
    const size_t ux = static_cast<size_t>(kv.first.x - ctx->dir_lo[0]);
    const size_t uy = static_cast<size_t>(kv.first.y - ctx->dir_lo[1]);
    const size_t uz = static_cast<size_t>(kv.first.z - ctx->dir_lo[2]);
    ctx->h_dir[(uz * ctx->dir_n[1] + uy) * ctx->dir_n[0] + ux] = kv.second;
  }
  ctx->dir_enabled = true;
}

// Uploads the hash table and (re)builds + uploads its directory.
int table_upload(a3d_ctx* ctx) {
  CK(ctx->d_table.reserve(ctx->h_table.size()));
  CK(cudaMemcpyAsync(ctx->d_table.p, ctx->h_table.data(), ctx->h_table.size() * sizeof(HashEntry),
                     cudaMemcpyHostToDevice, ctx->stream));
  dir_rebuild(ctx);
  if (ctx->dir_enabled) {
    CK(ctx->d_dir.reserve(ctx->h_dir.size()));
    CK(cudaMemcpyAsync(ctx->d_dir.p, ctx->h_dir.data(), ctx->h_dir.size() * sizeof(int32_t),
                       cudaMemcpyHostToDevice, ctx->stream));
  }
  return A3D_OK;
}

template <typename T>
int slab_grow(a3d_ctx* ctx, T** slab, size_t old_slots, size_t new_slots, size_t per_slot) {
  T* nd = nullptr;
  CK(cudaMalloc(&nd, new_slots * per_slot * sizeof(T)));
  if (*slab) {
    CK(cudaMemcpyAsync(nd, *slab, old_slots * per_slot * sizeof(T), cudaMemcpyDeviceToDevice,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(*slab);
  }
  *slab = nd;
  return A3D_OK;
}

int slab_reserve(a3d_ctx* ctx, size_t n_slots, bool want_weight) {
  const size_t nv = static_cast<size_t>(ctx->nv), p3 = static_cast<size_t>(ctx->P3);
  if (n_slots > ctx->slab_slots) {
    const size_t cap = std::max<size_t>(std::max<size_t>(n_slots, ctx->slab_slots * 2), 64);
    int rc = slab_grow(ctx, &ctx->d_dist, ctx->slab_slots, cap, p3);
    if (rc) return rc;
    rc = slab_grow(ctx, &ctx->d_mword, ctx->slab_slots, cap, nv);
    if (rc) return rc;
    if (ctx->d_weight) {
      rc = slab_grow(ctx, &ctx->d_weight, ctx->slab_slots, cap, nv);
      if (rc) return rc;
    }
    if (ctx->d_tsdf) {
      rc = slab_grow(ctx, &ctx->d_tsdf, ctx->slab_slots, cap, nv);
      if (rc) return rc;
    }
    ctx->slab_slots = cap;
  }
  if (want_weight && !ctx->d_weight) {
    CK(cudaMalloc(&ctx->d_weight, ctx->slab_slots * nv * sizeof(float)));
    CK(cudaMemsetAsync(ctx->d_weight, 0, ctx->slab_slots * nv * sizeof(float), ctx->stream));
    ctx->has_weight = true;
  }
  return A3D_OK;
}

// Frontier bits of the dense volume's words for the unallocated blocks (inside the volume's box) in the
// 3x3x3 neighbourhood of the changed blocks: exactly the unallocated blocks whose outer voxel layer can
// see a changed voxel through its neighbour test.
int vol_refresh_unallocated(a3d_ctx* ctx, const MapView& mv, int n_changed, const int32_t* changed_idx) {
  if (!mv.vol || n_changed <= 0) return A3D_OK;
  std::unordered_map<Key3, int32_t, Key3Hash> seen;
  std::vector<int32_t> idx;
  for (int i = 0; i < n_changed; ++i) {
    const Key3 k{changed_idx[3 * i], changed_idx[3 * i + 1], changed_idx[3 * i + 2]};
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const Key3 q{k.x + dx, k.y + dy, k.z + dz};
          if (ctx->slots.count(q) || seen.count(q)) continue;
          bool inside = true;
          const int32_t c[3] = {q.x, q.y, q.z};
          for (int a = 0; a < 3; ++a)
            inside = inside && c[a] >= ctx->dir_lo[a] &&
                     static_cast<int64_t>(c[a]) < static_cast<int64_t>(ctx->dir_lo[a]) + ctx->dir_n[a];
          if (!inside) continue;
          seen.emplace(q, 1);
          idx.insert(idx.end(), {q.x, q.y, q.z});
        }
  }
  if (idx.empty()) return A3D_OK;
  CK(ctx->pin_unalloc.reserve(idx.size()));
  std::copy(idx.begin(), idx.end(), ctx->pin_unalloc.p);
  CK(ctx->st_unalloc_idx.reserve(idx.size()));
  CK(cudaMemcpyAsync(ctx->st_unalloc_idx.p, ctx->pin_unalloc.p, idx.size() * sizeof(int32_t),
                     cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx->pin_unalloc.mark(ctx->stream));
  launch_vol_frontier_unallocated(ctx->stream, mv, ctx->d_vol, static_cast<int>(idx.size() / 3),
                                  ctx->st_unalloc_idx.p);
  ctx->launches++;
  CK(cudaGetLastError());
  return A3D_OK;
}

// K1b + K1c for every allocated block in `affected` (slots): refresh aprons, rebuild meta bytes.
int refresh_blocks(a3d_ctx* ctx, const std::vector<int32_t>& affected, int n_changed,
                   const int32_t* changed_idx) {
  const int n = static_cast<int>(affected.size());
  const MapView mv = ctx->view();
  if (n > 0) {
    // pinned staging: [slots n | neighbours 27n | block indices 3n]
    CK(ctx->pin_aff.reserve(static_cast<size_t>(n) * 31));
    int32_t* h_slots = ctx->pin_aff.p;
    int32_t* nbr = h_slots + n;
    int32_t* idx = nbr + static_cast<size_t>(n) * 27;
    for (int a = 0; a < n; ++a) {
      h_slots[a] = affected[a];
      const Key3 k = ctx->slot_key[affected[a]];
      idx[3 * a] = k.x;
      idx[3 * a + 1] = k.y;
      idx[3 * a + 2] = k.z;
      for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            auto it = ctx->slots.find(Key3{k.x + dx, k.y + dy, k.z + dz});
            nbr[static_cast<size_t>(a) * 27 + (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)] =
                it == ctx->slots.end() ? -1 : it->second;
          }
    }
    CK(ctx->st_aff_slots.reserve(n));
    CK(ctx->st_aff_nbr.reserve(static_cast<size_t>(n) * 27));
    CK(ctx->st_aff_idx.reserve(static_cast<size_t>(n) * 3));
    CK(cudaMemcpyAsync(ctx->st_aff_slots.p, h_slots, n * sizeof(int32_t), cudaMemcpyHostToDevice,
                       ctx->stream));
    CK(cudaMemcpyAsync(ctx->st_aff_nbr.p, nbr, static_cast<size_t>(n) * 27 * sizeof(int32_t),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->st_aff_idx.p, idx, static_cast<size_t>(n) * 3 * sizeof(int32_t),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->pin_aff.mark(ctx->stream));
    launch_refresh_aprons(ctx->stream, mv, n, ctx->st_aff_slots.p, ctx->st_aff_nbr.p, ctx->d_dist);
    launch_build_meta(ctx->stream, mv, n, ctx->st_aff_slots.p, ctx->st_aff_idx.p, ctx->d_mword,
                      mv.vol ? ctx->d_vol : nullptr);
    launch_build_frontier(ctx->stream, mv, n, ctx->st_aff_slots.p, ctx->st_aff_idx.p, ctx->d_mword,
                          mv.vol ? ctx->d_vol : nullptr);
    ctx->launches += 3;
    CK(cudaGetLastError());
  }
  return vol_refresh_unallocated(ctx, mv, n_changed, changed_idx);
}

// Slots of the allocated blocks in the 3x3x3 neighbourhood of the given block indices.
void collect_affected(a3d_ctx* ctx, int n, const int32_t* block_idx, std::vector<int32_t>* out) {
  std::vector<uint8_t> seen(static_cast<size_t>(ctx->next_slot), 0);
  for (int i = 0; i < n; ++i)
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          auto it = ctx->slots.find(
              Key3{block_idx[3 * i] + dx, block_idx[3 * i + 1] + dy, block_idx[3 * i + 2] + dz});
          if (it == ctx->slots.end() || seen[it->second]) continue;
          seen[it->second] = 1;
          out->push_back(it->second);
        }
}

// Brings the dense meta volume in line with the map before a cast: a no-op unless the directory's box
// changed since the last build (block updates inside the box are written by k_build_meta directly).
int vol_ensure(a3d_ctx* ctx) {
  if (!ctx->vol_stale) return A3D_OK;
  ctx->vol_on = false;
  const MapView m0 = ctx->view();
  const double words = static_cast<double>(m0.vol_n[0]) * m0.vol_n[1] * m0.vol_n[2];
  const double allocated = static_cast<double>(ctx->slots.size()) * ctx->nv;
  const bool eligible = ctx->dir_enabled && m0.vps_shift >= 2 && !ctx->slots.empty() &&
                        words < 2147483648.0 && words <= 64.0 * allocated + 16.0 * 1024 * 1024;
  if (!eligible) {
    ctx->vol_stale = false;
    return A3D_OK;
  }
  const size_t n_words = static_cast<size_t>(words);
  if (n_words > ctx->vol_words) {
    if (ctx->d_vol) cudaFree(ctx->d_vol);
    ctx->d_vol = nullptr;
    ctx->vol_words = 0;
    CK(cudaMalloc(&ctx->d_vol, n_words * sizeof(uint32_t)));
    ctx->vol_words = n_words;
  }
  const int n = static_cast<int>(ctx->slots.size());
  std::vector<int32_t> sl(n), idx(static_cast<size_t>(n) * 3);
  int a = 0;
  for (const auto& kv : ctx->slots) {
    sl[a] = kv.second;
    idx[3 * a] = kv.first.x;
    idx[3 * a + 1] = kv.first.y;
    idx[3 * a + 2] = kv.first.z;
    ++a;
  }
  CK(ctx->st_aff_slots.reserve(n));
  CK(ctx->st_aff_idx.reserve(idx.size()));
  CK(cudaMemcpyAsync(ctx->st_aff_slots.p, sl.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaMemcpyAsync(ctx->st_aff_idx.p, idx.data(), idx.size() * sizeof(int32_t),
                     cudaMemcpyHostToDevice, ctx->stream));
  launch_vol_fill(ctx->stream, m0, ctx->d_vol, 0, nullptr);
  launch_vol_scatter(ctx->stream, m0, ctx->d_vol, n, ctx->st_aff_slots.p, ctx->st_aff_idx.p);
  ctx->launches += 2;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));  // host vectors above are pageable
  ctx->vol_on = true;
  ctx->vol_stale = false;
  int rc = vol_refresh_unallocated(ctx, ctx->view(), n, idx.data());
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return A3D_OK;
}

// CTA width of the wavefront kernel for a batch (wave_cta = 0): 2, 4 or 8 warps, of which 10 / 5 / 2 fit
// an SM.  A W-warp CTA finishes a view up to W/2 times faster than a 2-warp CTA, bounded by the ratio
// of the view's leaf work on 64 lanes to the scheduler warp's anti-diagonal chain; the batch needs
// ceil(n / resident CTAs) passes.  Picks the width with the smallest estimate (measured on cfg2: 8 warps
// up to 296 segments, 4 up to 740, 2 beyond; SimpleRayCaster workload: 8 warps up to ~890).
int wave_warps_auto(const CasterView& c, int n_segments, int sm_count) {
  const bool iterative = c.kind == A3D_CASTER_ITERATIVE;
  const double kmax = std::max(1, c.kmax);
  const double leaf = c.n_rays * (iterative ? 0.5 * kmax : kmax) / 64.0;
  const double chain = iterative ? 0.5 * (c.res_x + c.res_y) * (0.25 * kmax + 4.0) : 0.0;
  const double ratio = std::min(4.0, std::max(1.0, leaf / std::max(chain, 0.25 * leaf)));
  int best = 2;
  double best_t = 1e300;
  const int widths[3] = {8, 4, 2}, per_sm[3] = {2, 5, cast_wave_blocks_per_sm()};
  for (int k = 0; k < 3; ++k) {
    const int resident = per_sm[k] * sm_count;
    const double passes = (n_segments + resident - 1) / resident;
    // (a wider CTA is worth >= 1.25x even for chain-bound views: no spills, fewer warps contending)
    const double t = widths[k] == 2 ? 0.95 * passes : passes / std::max(1.25, std::min(0.5 * widths[k], ratio));
    if (t < best_t) {
      best_t = t;
      best = widths[k];
    }
  }
  return best;
}

// Shared by the host- and device-pointer upload entry points.
int upload_blocks_impl(a3d_ctx* ctx, int n, const int32_t* block_idx, const float* dist,
                       const uint8_t* obs, const float* w, bool payload_on_device) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!ctx->configured) return ctx->fail(A3D_ERR_STATE, "a3d_map_config must be called first");
  if (n < 0 || (n > 0 && (!block_idx || !dist || !obs)))
    return ctx->fail(A3D_ERR_INVALID, "null block payload");
  if (n == 0) return A3D_OK;
  CK(cudaSetDevice(ctx->device));
  // slot assignment
  // slot | 1<<30 for blocks allocated by this call (their TSDF weights start at 0 when none are sent)
  std::vector<int32_t> slot_of(n);
  constexpr int32_t kFreshSlot = 1 << 30;
  bool table_dirty = false;
  for (int i = 0; i < n; ++i) {
    const Key3 k{block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]};
    auto it = ctx->slots.find(k);
    if (it != ctx->slots.end()) {
      slot_of[i] = it->second;
      continue;
    }
    int32_t s;
    if (!ctx->free_slots.empty()) {
      s = ctx->free_slots.back();
      ctx->free_slots.pop_back();
    } else {
      s = ctx->next_slot++;
    }
    ctx->slots.emplace(k, s);
    if (ctx->slot_key.size() <= static_cast<size_t>(s)) ctx->slot_key.resize(s + 1);
    ctx->slot_key[s] = k;
    slot_of[i] = s | kFreshSlot;
    table_dirty = true;
  }
  int rc = slab_reserve(ctx, static_cast<size_t>(ctx->next_slot), w != nullptr);
  if (rc) return rc;
  if (table_dirty) {
    if (ctx->h_table.empty() || 4 * ctx->slots.size() > ctx->h_table.size()) {
      table_rebuild(ctx, ctx->slots.size());
    } else {
      for (int i = 0; i < n; ++i)
        table_insert(ctx->h_table, block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2],
                     slot_of[i] & ~kFreshSlot);
    }
    rc = table_upload(ctx);
    if (rc) return rc;
  }
  // payload, in chunks bounded by the staging size
  const size_t nv = static_cast<size_t>(ctx->nv);
  const int chunk = static_cast<int>(std::max<size_t>(1, (size_t(256) << 20) / (nv * 4)));
  CK(ctx->pin_slots.reserve(n));
  std::copy(slot_of.begin(), slot_of.end(), ctx->pin_slots.p);
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int m = std::min(chunk, n - i0);
    CK(ctx->st_slots.reserve(m));
    CK(cudaMemcpyAsync(ctx->st_slots.p, ctx->pin_slots.p + i0, m * sizeof(int32_t),
                       cudaMemcpyHostToDevice, ctx->stream));
    const float* dsrc = dist + static_cast<size_t>(i0) * nv;
    const uint8_t* osrc = obs + static_cast<size_t>(i0) * nv;
    const float* wsrc = w ? w + static_cast<size_t>(i0) * nv : nullptr;
    if (!payload_on_device) {
      CK(ctx->st_dist.reserve(m * nv));
      CK(ctx->st_obs.reserve(m * nv));
      CK(cudaMemcpyAsync(ctx->st_dist.p, dsrc, m * nv * sizeof(float), cudaMemcpyHostToDevice,
                         ctx->stream));
      CK(cudaMemcpyAsync(ctx->st_obs.p, osrc, m * nv, cudaMemcpyHostToDevice, ctx->stream));
      dsrc = ctx->st_dist.p;
      osrc = ctx->st_obs.p;
      if (w) {
        CK(ctx->st_w.reserve(m * nv));
        CK(cudaMemcpyAsync(ctx->st_w.p, wsrc, m * nv * sizeof(float), cudaMemcpyHostToDevice,
                           ctx->stream));
        wsrc = ctx->st_w.p;
      }
    }
    launch_apply_blocks(ctx->stream, ctx->view(), m, ctx->st_slots.p, dsrc, osrc, wsrc, ctx->d_dist,
                        ctx->has_weight ? ctx->d_weight : nullptr, ctx->d_tsdf);
    ctx->launches++;
    CK(cudaGetLastError());  // (staging buffers are reused by the next chunk in stream order)
  }
  CK(ctx->pin_slots.mark(ctx->stream));
  // aprons + meta of the uploaded blocks and of their allocated neighbours
  std::vector<int32_t> affected;
  collect_affected(ctx, n, block_idx, &affected);
  rc = refresh_blocks(ctx, affected, n, block_idx);
  if (rc) return rc;
  // one synchronisation per update: on return the caller's payload has been consumed
  CK(cudaStreamSynchronize(ctx->stream));
  return A3D_OK;
}

int point_query(a3d_ctx* ctx, int what, int64_t n, const double* pts, void* out0, size_t out0_elt,
                void* out1, double arg = 0.0) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!ctx->configured) return ctx->fail(A3D_ERR_STATE, "a3d_map_config must be called first");
  if (n < 0 || (n > 0 && (!pts || !out0))) return ctx->fail(A3D_ERR_INVALID, "null argument");
  if (n == 0) return A3D_OK;
  CK(cudaSetDevice(ctx->device));
  if (ctx->h_table.empty()) {
    table_rebuild(ctx, 0);
    int rc = table_upload(ctx);
    if (rc) return rc;
  }
  CK(ctx->sc_pts.reserve(3 * n));
  CK(cudaMemcpyAsync(ctx->sc_pts.p, pts, 3 * n * sizeof(double), cudaMemcpyHostToDevice,
                     ctx->stream));
  const size_t out0_bytes = static_cast<size_t>(n) * out0_elt;
  CK(ctx->sc_out.reserve((out0_bytes + 7) / 8));
  CK(ctx->sc_u8.reserve(n));
  launch_point_query(ctx->stream, ctx->view(), what, n, ctx->sc_pts.p, ctx->sc_out.p,
                     what == kQueryTsdfDistance ? static_cast<void*>(ctx->d_tsdf) : static_cast<void*>(ctx->sc_u8.p),
                     arg);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out0, ctx->sc_out.p, out0_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (out1) CK(cudaMemcpyAsync(out1, ctx->sc_u8.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return A3D_OK;
}

EvalView make_eval_view(const a3d_ctx* ctx, const a3d_eval_params& p) {
  EvalView v{};
  v.kind = p.kind;
  v.accurate_frontiers = p.accurate_frontiers;
  v.surface_frontiers = p.surface_frontiers;
  v.bv_is_setup = p.bv_is_setup;
  v.nb_step = static_cast<double>(ctx->vs_f) * p.checking_distance;  // frontier_evaluator.cpp:30-31
  for (int i = 0; i < 6; ++i) v.gains[i] = p.gains[i];
  for (int i = 0; i < 3; ++i) {
    v.bv_min[i] = p.bv_min[i];
    v.bv_max[i] = p.bv_max[i];
  }
  v.bv_q = Q4{p.bv_quat[0], p.bv_quat[1], p.bv_quat[2], p.bv_quat[3]};
  v.frontier_voxel_weight = p.frontier_voxel_weight;
  v.min_impact_factor = p.min_impact_factor;
  v.new_voxel_weight = p.new_voxel_weight;
  v.ray_angle_xy = p.ray_angle_x * p.ray_angle_y;
  return v;
}

// The scan-order kernel on the segments whose bit is set in seg_mask (null: all of them).
int enqueue_scan(a3d_ctx* ctx, const a3d_caster* c, const a3d_eval* e, Batch* b, bool digest, bool emit,
                 const unsigned int* seg_mask) {
  const int block = cast_gain_block_threads();
  const int warps_per_block = block / 32;
  int grid = (b->n_segments + warps_per_block - 1) / warps_per_block;
  grid = std::max(1, std::min(grid, ctx->sm_count * 32));
  if (c->view.kind == A3D_CASTER_ITERATIVE) {
    b->table_stride = (static_cast<size_t>(c->view.n_rays) + 15) / 16 * 16;
    CK(ctx->sc_tables.reserve(b->table_stride * grid * warps_per_block));
    b->tables = ctx->sc_tables.p;
  }
  b->work_counter = ctx->sc_counter.p + 1;
  b->seg_mask = seg_mask;
  launch_cast_gain(ctx->stream, ctx->view(), c->view, e ? &e->view : nullptr, *b, digest, emit, grid, block);
  ctx->launches++;
  CK(cudaGetLastError());
  return A3D_OK;
}

// Enqueue the cast+gain work for a batch whose inputs/outputs are already on the device.
// Kernel choice: the wavefront kernel (k_cast_wave) when only order-free outputs are wanted, else the
// scan-order kernel (k_cast_gain), which also re-evaluates the rare segments k_cast_wave flags.
int enqueue_cast(a3d_ctx* ctx, const a3d_caster* c, const a3d_eval* e, const double* pos_dev,
                 const double* quat_dev, const int32_t* first_dev, int n_segments, double* gains_dev,
                 unsigned long long* nvis_dev, unsigned long long* nsamp_dev,
                 unsigned long long* dig_dev, unsigned long long* set_dig_dev, double* centres_dev,
                 long long cap, bool emit, const double* seg_origin_dev,
                 const long long* emit_offset_dev = nullptr, unsigned long long* keys_dev = nullptr,
                 int n_views = -1 /* total views of the batch when the caller knows it */) {
  if (ctx->h_table.empty()) {
    table_rebuild(ctx, 0);
    int rc = table_upload(ctx);
    if (rc) return rc;
  }
  ctx->part_offsets_valid = false;
  Batch b{};
  b.pos = pos_dev;
  b.quat = quat_dev;
  b.view_first = first_dev;
  b.n_segments = n_segments;
  b.gains = gains_dev;
  b.n_visible = nvis_dev;
  b.n_samples = nsamp_dev;
  b.digests = dig_dev;
  b.set_digests = set_dig_dev;
  b.centres_out = centres_dev;
  b.cap = cap;
  b.emit_offset = emit_offset_dev;
  b.seg_origin = seg_origin_dev;
  b.keys_out = keys_dev;
  CK(ctx->sc_counter.reserve(4));
  CK(cudaMemsetAsync(ctx->sc_counter.p, 0, 4 * sizeof(unsigned int), ctx->stream));
  b.key_error = ctx->sc_counter.p + 2;

  // The wavefront kernel emits lists through per-part slot ranges (4 MB of scratch per resident view
  // on the bench workload) that decode voxel keys with a shift: a latency win for the per-segment call
  // and small batches (2.2 vs 6.6 ms for one view), a loss against the scan-order kernel's streaming
  // writes for thousands of segments (36 vs 27 ms for 2048).
  const bool emit_by_wave = emit && ctx->view().vps_shift >= 0 && n_segments <= 2 * ctx->sm_count;
  // A list alone (no evaluator: the sensor model's own getVisibleVoxels) is the Naive fold with the
  // gain ignored and no bounding volume.
  EvalView no_eval{};
  no_eval.kind = A3D_EVAL_NAIVE;
  const EvalView& ev_view = e ? e->view : no_eval;
  const bool wave_ok = (e || emit_by_wave) && !dig_dev && (!emit || emit_by_wave) &&
                       !(ev_view.kind == A3D_EVAL_VOXEL_WEIGHT && set_dig_dev) && !(emit && set_dig_dev) &&
                       std::min(c->view.res_x, c->view.res_y) <= cast_wave_max_diag() &&
                       c->view.n_rays < (1 << 24);
  if (ctx->opt_cast_kernel == 2 && !wave_ok)
    return ctx->fail(A3D_ERR_UNSUPPORTED, "cast_kernel=2 (wavefront) cannot produce ordered digests");
  bool use_wave = wave_ok && ctx->opt_cast_kernel != 1 && ctx->opt_cast_kernel != 3;
  // the kernel chain of a3d_split.cu: gains / counts / ordered lists (no digests), dense meta volume required
  bool use_split = (e || emit) && !dig_dev && !set_dig_dev && n_views >= 0 && (!emit || ctx->view().vps_shift >= 0) &&
                   (ctx->opt_cast_kernel == 0 || ctx->opt_cast_kernel == 3);
  if (use_wave || use_split) {
    int rc = vol_ensure(ctx);
    if (rc) return rc;
  }
  use_split = use_split && cast_split_supported(ctx->view(), c->view);
  if (ctx->opt_cast_kernel == 3 && !use_split)
    return ctx->fail(A3D_ERR_UNSUPPORTED,
                     "cast_kernel=3 (kernel chain) needs the dense meta volume and cannot produce digests");
  if (use_split) use_wave = false;
  ctx->last_cast_kernel = use_split ? 3 : use_wave ? 2 : 1;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  const unsigned int* seg_mask = nullptr;
  if (use_split) {
    const int mode = emit ? 1 : (ctx->want_part_offsets ? 2 : 0);
    const size_t per_view = cast_split_bytes_per_view(c->view, ev_view.kind, nullptr, mode);
    const long long max_views = static_cast<long long>(std::max<size_t>(1, (size_t(4) << 30) / per_view));
    std::vector<int32_t> h_first;  // needed only to cut the batch into chunks that fit the scratch
    if (n_views > max_views) {
      h_first.resize(static_cast<size_t>(n_segments) + 1);
      CK(cudaMemcpyAsync(h_first.data(), first_dev, h_first.size() * sizeof(int32_t), cudaMemcpyDeviceToHost,
                         ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    const size_t n_words = static_cast<size_t>(n_segments) / 32 + 1;
    CK(ctx->sc_inexact.reserve(n_words));
    CK(cudaMemsetAsync(ctx->sc_inexact.p, 0, n_words * sizeof(unsigned int), ctx->stream));
    for (int seg = 0; seg < n_segments;) {
      int seg_end = n_segments, v0 = 0, v1 = n_views;
      if (!h_first.empty()) {
        v0 = h_first[seg];
        seg_end = seg + 1;  // (a single segment larger than the budget gets a larger scratch)
        while (seg_end < n_segments && h_first[seg_end + 1] - v0 <= max_views) ++seg_end;
        v1 = h_first[seg_end];
      }
      CK(ctx->sc_wave.reserve(static_cast<size_t>(v1 - v0) * per_view + 256));
      for (auto& evs : ctx->ev_stage)
        if (!evs) CK(cudaEventCreate(&evs));
      ctx->n_stage = launch_cast_split(ctx->stream, ctx->view(), c->view, ev_view, b, seg, seg_end - seg, v0, v1 - v0,
                                       ctx->sc_wave.p, ctx->sc_inexact.p, ctx->sm_count, ctx->ev_stage, mode);
      ctx->launches += cast_split_launches(c->view, mode);
      CK(cudaGetLastError());
      seg = seg_end;
    }
    seg_mask = ctx->sc_inexact.p;
    ctx->part_offsets_valid = mode == 2 && h_first.empty();  // (one chunk: the scratch holds the whole batch)
  }
  if (use_wave) {
    // as many CTAs as can be resident (measured: trimming the grid so that every CTA gets the same
    // number of segments is slower — throughput follows the number of resident warps)
    int grid = std::max(1, std::min(n_segments, ctx->sm_count * cast_wave_blocks_per_sm()));
    if (emit) grid = std::max(1, std::min(n_segments, 2 * ctx->sm_count));  // resident 8-warp CTAs
    size_t stride = 0;
    const bool extra = set_dig_dev != nullptr || ev_view.kind == A3D_EVAL_VOXEL_WEIGHT;
    const int emit_kmax = emit ? c->view.kmax : 0;
    while (grid > 1 &&
           wave_scratch_bytes(c->view.n_rays, grid, extra, &stride, emit_kmax) > (size_t(emit ? 16 : 6) << 30))
      grid /= 2;
    CK(ctx->sc_wave.reserve(wave_scratch_bytes(c->view.n_rays, grid, extra, &stride, emit_kmax)));
    const size_t n_words = static_cast<size_t>(n_segments) / 32 + 1;
    CK(ctx->sc_inexact.reserve(n_words));
    CK(cudaMemsetAsync(ctx->sc_inexact.p, 0, n_words * sizeof(unsigned int), ctx->stream));
    b.work_counter = ctx->sc_counter.p;
    launch_cast_wave(ctx->stream, ctx->view(), c->view, ev_view, b, set_dig_dev != nullptr, grid,
                     ctx->sc_wave.p, ctx->sc_inexact.p, ctx->opt_wave_cta == 1   ? 2
                     : ctx->opt_wave_cta == 2 ? 8
                     : ctx->opt_wave_cta == 3 ? 4
                                              : wave_warps_auto(c->view, n_segments, ctx->sm_count));
    ctx->launches++;
    CK(cudaGetLastError());
    seg_mask = ctx->sc_inexact.p;
  }
  {
    int rc = enqueue_scan(ctx, c, e, &b, dig_dev != nullptr || set_dig_dev != nullptr, emit, seg_mask);
    if (rc) return rc;
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  return A3D_OK;
}

// Second pass of a stored-list call, when the first pass (enqueue_cast with want_part_offsets) went through
// the kernel chain in one chunk: the chain casts again and writes the lists' keys straight into the arena;
// the segments the first pass flagged are emitted by the scan-order kernel, as in enqueue_cast.
int enqueue_direct_emit(a3d_ctx* ctx, const a3d_caster* c, const a3d_eval* e, int n_segments, int n_views,
                        const double* seg_origin_dev, const long long* emit_offset_dev,
                        unsigned long long* keys_dev) {
  Batch b{};
  b.pos = ctx->sc_pos.p;
  b.quat = ctx->sc_quat.p;
  b.view_first = ctx->sc_first.p;
  b.n_segments = n_segments;
  b.gains = ctx->sc_gains.p;
  b.n_visible = ctx->sc_nvis.p;
  b.n_samples = ctx->sc_nsamp.p;
  b.seg_origin = seg_origin_dev;
  b.emit_offset = emit_offset_dev;
  b.keys_out = keys_dev;
  CK(cudaMemsetAsync(ctx->sc_counter.p, 0, 4 * sizeof(unsigned int), ctx->stream));
  b.key_error = ctx->sc_counter.p + 2;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  launch_cast_split_direct(ctx->stream, ctx->view(), c->view, e->view, b, n_segments, n_views, ctx->sc_wave.p,
                           ctx->sm_count);
  ctx->launches += cast_split_launches(c->view, 0) - 1;
  CK(cudaGetLastError());
  int rc = enqueue_scan(ctx, c, e, &b, false, true, ctx->sc_inexact.p);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->part_offsets_valid = false;
  return A3D_OK;
}

}  // namespace

#define CHECK_FRESH(obj)                                                                              \
  if ((obj) && (obj)->config_gen != ctx->config_gen)                                                  \
  return ctx->fail(A3D_ERR_STATE, "caster/evaluator was created before a3d_map_config changed the map geometry")

extern "C" {

int a3d_ctx_create(int device_id, a3d_ctx** out) {
  if (!out) return A3D_ERR_INVALID;
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return A3D_ERR_CUDA;
  }
  if (device_id < 0 || device_id >= n_dev) {
    g_create_error = "device_id out of range";
    return A3D_ERR_INVALID;
  }
  e = cudaSetDevice(device_id);
  if (e != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return A3D_ERR_CUDA;
  }
  a3d_ctx* ctx = new a3d_ctx();
  ctx->device = device_id;
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, device_id);
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    g_create_error = "liba3d_b200 is built for sm_100a only; device is sm_" +
                     std::to_string(prop.major) + std::to_string(prop.minor);
    delete ctx;
    return A3D_ERR_CUDA;
  }
  {
    // list arenas of the stores (a3d_store::new_arena) are recycled through the device's default pool:
    // up to 8 GB of released arenas stay with it instead of going back to the driver at the next sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
      uint64_t keep = 8ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    g_create_error = "stream/event creation failed";
    delete ctx;
    return A3D_ERR_CUDA;
  }
  *out = ctx;
  return A3D_OK;
}

void a3d_ctx_destroy(a3d_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& evs : ctx->ev_stage)
    if (evs) cudaEventDestroy(evs);
  if (ctx->d_dist) cudaFree(ctx->d_dist);
  if (ctx->d_weight) cudaFree(ctx->d_weight);
  if (ctx->d_tsdf) cudaFree(ctx->d_tsdf);
  if (ctx->d_mword) cudaFree(ctx->d_mword);
  ctx->d_dir.release();
  if (ctx->d_vol) cudaFree(ctx->d_vol);
  ctx->d_vol = nullptr;
  ctx->vol_words = 0;
  ctx->vol_stale = true;
  ctx->vol_on = false;
  ctx->d_table.release();
  ctx->st_aff_slots.release();
  ctx->st_aff_nbr.release();
  ctx->st_aff_idx.release();
  ctx->st_unalloc_idx.release();
  ctx->pin_slots.release();
  ctx->pin_aff.release();
  ctx->pin_unalloc.release();
  ctx->st_dist.release();
  ctx->st_w.release();
  ctx->st_obs.release();
  ctx->st_slots.release();
  ctx->sc_pts.release();
  ctx->sc_out.release();
  ctx->sc_u8.release();
  ctx->sc_pos.release();
  ctx->sc_quat.release();
  ctx->sc_gains.release();
  ctx->sc_centres.release();
  ctx->sc_first.release();
  ctx->sc_nvis.release();
  ctx->sc_nsamp.release();
  ctx->sc_dig.release();
  ctx->sc_cnt.release();
  ctx->sc_tables.release();
  ctx->sc_counter.release();
  ctx->sc_origin.release();
  ctx->sc_inexact.release();
  ctx->sc_wave.release();
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* a3d_last_error(const a3d_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}
void* a3d_ctx_stream(a3d_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
float a3d_last_kernel_ms(a3d_ctx* ctx) {
  if (!ctx || !ctx->timed) return -1.0f;
  if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.0f;
  return ms;
}
int a3d_last_stage_ms(a3d_ctx* ctx, float* out_ms, int cap) {
  if (!ctx || !out_ms || cap < 0) return A3D_ERR_INVALID;
  if (!ctx->timed || ctx->last_cast_kernel != 3 || ctx->n_stage < 2) return 0;
  if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return 0;
  int n = 0;
  for (int k = 0; k + 1 < ctx->n_stage && n < cap; ++k, ++n)
    if (cudaEventElapsedTime(&out_ms[n], ctx->ev_stage[k], ctx->ev_stage[k + 1]) != cudaSuccess) return n;
  return n;
}
uint64_t a3d_kernel_launches(a3d_ctx* ctx) { return ctx ? ctx->launches : 0; }

int a3d_map_config(a3d_ctx* ctx, float voxel_size, int vps) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!(voxel_size > 0.0f) || vps < 1 || vps > 64)
    return ctx->fail(A3D_ERR_INVALID, "voxel_size must be > 0 and 1 <= voxels_per_side <= 64");
  if (ctx->configured && !ctx->slots.empty())
    return ctx->fail(A3D_ERR_STATE, "map already holds blocks; a3d_map_clear first");
  if (ctx->configured && (vps != ctx->vps || voxel_size != ctx->vs_f)) {
    // a new geometry: the slabs are sized per slot by vps, casters and evaluators cache the voxel size
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_dist) cudaFree(ctx->d_dist);
    if (ctx->d_mword) cudaFree(ctx->d_mword);
    if (ctx->d_weight) cudaFree(ctx->d_weight);
    if (ctx->d_tsdf) cudaFree(ctx->d_tsdf);
    if (ctx->d_vol) cudaFree(ctx->d_vol);
    ctx->d_tsdf = nullptr;
    ctx->d_dist = nullptr;
    ctx->d_mword = nullptr;
    ctx->d_weight = nullptr;
    ctx->d_vol = nullptr;
    ctx->vol_words = 0;
    ctx->vol_on = false;
    ctx->vol_stale = true;
    ctx->slab_slots = 0;
    ctx->has_weight = false;
    ctx->slot_key.clear();
    ctx->h_dir.clear();
    ctx->dir_enabled = false;
    ctx->config_gen++;
  }
  // voxblox Layer / Block constructors
  ctx->vs_f = voxel_size;
  ctx->vps = vps;
  ctx->nv = vps * vps * vps;
  ctx->P = vps + 2;
  ctx->P3 = ctx->P * ctx->P * ctx->P;
  ctx->vs_inv_f = static_cast<float>(1.0 / static_cast<double>(voxel_size));
  ctx->bs_f = voxel_size * static_cast<float>(vps);
  ctx->bs_inv_f = static_cast<float>(1.0 / static_cast<double>(ctx->bs_f));
  ctx->configured = true;
  return A3D_OK;
}

int a3d_map_upload_blocks(a3d_ctx* ctx, int n, const int32_t* block_idx, const float* esdf_dist,
                          const uint8_t* esdf_observed, const float* tsdf_weight) {
  return upload_blocks_impl(ctx, n, block_idx, esdf_dist, esdf_observed, tsdf_weight, false);
}
int a3d_map_upload_blocks_dev(a3d_ctx* ctx, int n, const int32_t* block_idx, const float* dist_dev,
                              const uint8_t* obs_dev, const float* w_dev) {
  return upload_blocks_impl(ctx, n, block_idx, dist_dev, obs_dev, w_dev, true);
}

int a3d_map_upload_weights(a3d_ctx* ctx, int n, const int32_t* block_idx, const float* tsdf_weight,
                           uint8_t* applied_out) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!ctx->configured) return ctx->fail(A3D_ERR_STATE, "a3d_map_config must be called first");
  if (n < 0 || (n > 0 && (!block_idx || !tsdf_weight))) return ctx->fail(A3D_ERR_INVALID, "null block payload");
  if (n == 0) return A3D_OK;
  CK(cudaSetDevice(ctx->device));
  int rc = slab_reserve(ctx, static_cast<size_t>(std::max(ctx->next_slot, 1)), true);
  if (rc) return rc;
  const size_t nv = static_cast<size_t>(ctx->nv);
  for (int i = 0; i < n; ++i) {
    const Key3 k{block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]};
    auto it = ctx->slots.find(k);
    if (applied_out) applied_out[i] = it != ctx->slots.end();
    if (it == ctx->slots.end()) continue;  // no ESDF block there yet: the caller keeps it
    CK(cudaMemcpyAsync(ctx->d_weight + static_cast<size_t>(it->second) * nv, tsdf_weight + static_cast<size_t>(i) * nv,
                       nv * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return A3D_OK;
}

int a3d_map_upload_tsdf_distances(a3d_ctx* ctx, int n, const int32_t* block_idx, const float* tsdf_distance,
                                  uint8_t* applied_out) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!ctx->configured) return ctx->fail(A3D_ERR_STATE, "a3d_map_config must be called first");
  if (n < 0 || (n > 0 && (!block_idx || !tsdf_distance))) return ctx->fail(A3D_ERR_INVALID, "null block payload");
  if (n == 0) return A3D_OK;
  CK(cudaSetDevice(ctx->device));
  int rc = slab_reserve(ctx, static_cast<size_t>(std::max(ctx->next_slot, 1)), false);
  if (rc) return rc;
  const size_t nv = static_cast<size_t>(ctx->nv);
  if (!ctx->d_tsdf) {
    CK(cudaMalloc(&ctx->d_tsdf, ctx->slab_slots * nv * sizeof(float)));
    CK(cudaMemsetAsync(ctx->d_tsdf, 0, ctx->slab_slots * nv * sizeof(float), ctx->stream));
  }
  for (int i = 0; i < n; ++i) {
    const Key3 k{block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]};
    auto it = ctx->slots.find(k);
    if (applied_out) applied_out[i] = it != ctx->slots.end();
    if (it == ctx->slots.end()) continue;  // no ESDF block there yet: the caller keeps it
    CK(cudaMemcpyAsync(ctx->d_tsdf + static_cast<size_t>(it->second) * nv, tsdf_distance + static_cast<size_t>(i) * nv,
                       nv * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return A3D_OK;
}

int a3d_map_remove_blocks(a3d_ctx* ctx, int n, const int32_t* block_idx) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!ctx->configured) return ctx->fail(A3D_ERR_STATE, "a3d_map_config must be called first");
  if (n < 0 || (n > 0 && !block_idx)) return ctx->fail(A3D_ERR_INVALID, "null block_idx");
  CK(cudaSetDevice(ctx->device));
  std::vector<int32_t> removed;  // indices of the blocks that were allocated (they lie inside the box)
  for (int i = 0; i < n; ++i) {
    auto it = ctx->slots.find(Key3{block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]});
    if (it == ctx->slots.end()) continue;
    ctx->free_slots.push_back(it->second);
    ctx->slots.erase(it);
    removed.insert(removed.end(), {block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]});
  }
  const bool dirty = !removed.empty();
  const int n_removed = static_cast<int>(removed.size() / 3);
  if (dirty) {
    table_rebuild(ctx, ctx->slots.size());
    int rc = table_upload(ctx);
    if (rc) return rc;
    const MapView mv = ctx->view();
    if (mv.vol) {
      // removed blocks inside the (unchanged) box read "no block" again; blocks that were not
      // allocated hold that value already
      CK(ctx->st_aff_idx.reserve(removed.size()));
      CK(cudaMemcpyAsync(ctx->st_aff_idx.p, removed.data(), removed.size() * sizeof(int32_t),
                         cudaMemcpyHostToDevice, ctx->stream));
      launch_vol_fill(ctx->stream, mv, ctx->d_vol, n_removed, ctx->st_aff_idx.p);
      ctx->launches++;
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(ctx->stream));
    }
    // the removed blocks' neighbours must stop mirroring them
    std::vector<int32_t> affected;
    collect_affected(ctx, n_removed, removed.data(), &affected);
    rc = refresh_blocks(ctx, affected, n_removed, removed.data());
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return A3D_OK;
}

int a3d_map_clear(a3d_ctx* ctx) {
  if (!ctx) return A3D_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  ctx->slots.clear();
  ctx->free_slots.clear();
  ctx->next_slot = 0;
  ctx->vol_stale = true;
  if (ctx->configured) {
    table_rebuild(ctx, 0);
    int rc = table_upload(ctx);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return A3D_OK;
}

int64_t a3d_map_num_blocks(const a3d_ctx* ctx) {
  return ctx ? static_cast<int64_t>(ctx->slots.size()) : 0;
}
double a3d_map_voxel_size(const a3d_ctx* ctx) {
  return ctx ? static_cast<double>(ctx->vs_f) : 0.0;
}

int a3d_map_voxel_state(a3d_ctx* ctx, int64_t n, const double* pts, uint8_t* out) {
  return point_query(ctx, kQueryState, n, pts, out, 1, nullptr);
}
int a3d_map_voxel_state_literal(a3d_ctx* ctx, int64_t n, const double* pts, uint8_t* out) {
  return point_query(ctx, kQueryStateGeneric, n, pts, out, 1, nullptr);
}
int a3d_map_voxel_center(a3d_ctx* ctx, int64_t n, const double* pts, double* out) {
  return point_query(ctx, kQueryCenter, n, pts, out, 24, nullptr);
}
int a3d_map_is_observed(a3d_ctx* ctx, int64_t n, const double* pts, uint8_t* out) {
  return point_query(ctx, kQueryObserved, n, pts, out, 1, nullptr);
}
int a3d_map_voxel_weight(a3d_ctx* ctx, int64_t n, const double* pts, double* out) {
  return point_query(ctx, kQueryWeight, n, pts, out, 8, nullptr);
}
int a3d_map_voxel_tsdf_distance(a3d_ctx* ctx, int64_t n, const double* pts, double* out) {
  return point_query(ctx, kQueryTsdfDistance, n, pts, out, 8, nullptr);
}
int a3d_map_traversable(a3d_ctx* ctx, int64_t n, const double* pts, double collision_radius,
                        uint8_t* out) {
  return point_query(ctx, kQueryTraversable, n, pts, out, 1, nullptr, collision_radius);
}
int a3d_map_distance(a3d_ctx* ctx, int64_t n, const double* pts, double* dist, uint8_t* ok) {
  if (ctx && n > 0 && !ok) return ctx->fail(A3D_ERR_INVALID, "null argument");
  return point_query(ctx, kQueryDistance, n, pts, dist, 8, ok);
}

// ---- caster ---------------------------------------------------------------------------------
int a3d_caster_create(a3d_ctx* ctx, const a3d_caster_params* p, a3d_caster** out) {
  if (!ctx || !out) return A3D_ERR_INVALID;
  *out = nullptr;
  if (!p) return ctx->fail(A3D_ERR_INVALID, "null params");
  if (!ctx->configured) return ctx->fail(A3D_ERR_STATE, "a3d_map_config must be called first");
  // CameraModel::checkParamsValid (camera_model.cpp:185-200)
  if (!(p->ray_length > 0.0)) return ctx->fail(A3D_ERR_INVALID, "ray_length expected > 0.0");
  const bool lidar = p->kind == A3D_CASTER_ITERATIVE_LIDAR;
  if (!lidar && !(p->focal_length > 0.0))
    return ctx->fail(A3D_ERR_INVALID, "focal_length expected > 0.0");
  if (lidar && !(p->field_of_view_x > 0.0))
    return ctx->fail(A3D_ERR_INVALID, "field_of_view_x expected > 0.0");  // lidar_model.cpp:185-187
  if (lidar && !(p->field_of_view_y > 0.0))
    return ctx->fail(A3D_ERR_INVALID, "field_of_view_y expected > 0.0");
  if (p->resolution_x <= 0) return ctx->fail(A3D_ERR_INVALID, "resolution_x expected > 0");
  if (p->resolution_y <= 0) return ctx->fail(A3D_ERR_INVALID, "resolution_y expected > 0");
  if (p->kind != A3D_CASTER_SIMPLE && p->kind != A3D_CASTER_ITERATIVE && !lidar)
    return ctx->fail(A3D_ERR_INVALID, "unknown caster kind");
  const bool iterative = p->kind != A3D_CASTER_SIMPLE;
  if (!(p->downsampling_factor > 0.0))
    return ctx->fail(A3D_ERR_INVALID, "downsampling_factor expected > 0.0");
  CK(cudaSetDevice(ctx->device));

  a3d_caster_info info{};
  const double vs = static_cast<double>(ctx->vs_f);  // OccupancyMap::getVoxelSize
  // camera_model.cpp:181-182
  info.fov_x = 2.0 * std::atan2(static_cast<double>(p->resolution_x), p->focal_length * 2.0);
  info.fov_y = 2.0 * std::atan2(static_cast<double>(p->resolution_y), p->focal_length * 2.0);
  if (lidar) {  // lidar_model.cpp:170-177
    info.fov_x = p->field_of_view_x;
    info.fov_y = p->field_of_view_y;
    info.fov_x *= M_PI / 180.0;
    info.fov_y *= M_PI / 180.0;
  }
  info.ray_step = p->ray_step > 0.0 ? p->ray_step : vs;
  // simple_ray_caster.cpp:22-29 == iterative_ray_caster.cpp:25-32
  info.c_res_x = std::min(
      static_cast<int>(std::ceil(p->ray_length * info.fov_x / (vs * p->downsampling_factor))),
      p->resolution_x);
  info.c_res_y = std::min(
      static_cast<int>(std::ceil(p->ray_length * info.fov_y / (vs * p->downsampling_factor))),
      p->resolution_y);
  if (info.c_res_x < 2 || info.c_res_y < 2)
    return ctx->fail(A3D_ERR_INVALID,
                     "derived ray grid smaller than 2x2 (the reference divides by c_res-1)");
  const int n_rays = info.c_res_x * info.c_res_y;
  int n_sections = 0;
  std::vector<double> split;  // c_split_distances_
  if (iterative) {
    // iterative_ray_caster.cpp:35-45
    n_sections = static_cast<int>(std::floor(std::log2(
        std::min(static_cast<double>(info.c_res_x), static_cast<double>(info.c_res_y)))));
    if (n_sections < 1 || n_sections > 31)
      return ctx->fail(A3D_ERR_INVALID, "c_n_sections out of range");
    split.assign(n_sections + 1, 0.0);
    for (int t = 1; t <= n_sections; ++t)
      split[t] = p->ray_length / std::pow(2.0, static_cast<double>(n_sections - t));
    for (int t = 0; t <= n_sections; ++t) {
      info.split_distances[t] = split[t];
      info.split_widths[t] = t < n_sections ? (1 << (n_sections - 1 - t)) : 0;
    }
  }
  info.c_n_sections = n_sections;

  // Accumulated sample distances (simple_ray_caster.cpp:46-50, iterative_ray_caster.cpp:75-79):
  // d_0 = start, d_{k+1} = d_k + ray_step while d_k < ray_length.  One sequence per start section.
  const int n_starts = iterative ? n_sections : 1;
  std::vector<std::vector<double>> seqs(n_starts);
  size_t kmax = 1;
  for (int s = 0; s < n_starts; ++s) {
    double d = iterative ? split[s] : 0.0;
    while (d < p->ray_length) {
      seqs[s].push_back(d);
      d += info.ray_step;
      if (seqs[s].size() > (size_t(1) << 24))
        return ctx->fail(A3D_ERR_INVALID, "ray_step too small for ray_length");
    }
    kmax = std::max(kmax, seqs[s].size());
  }
  info.samples_per_full_ray = static_cast<int32_t>(seqs[0].size());
  kmax = (kmax + 31) / 32 * 32;
  std::vector<double> dseq(static_cast<size_t>(n_starts) * kmax, 0.0);
  std::vector<int32_t> kcount(n_starts), seg_end(static_cast<size_t>(n_starts) * std::max(1, n_sections), 0);
  for (int s = 0; s < n_starts; ++s) {
    std::copy(seqs[s].begin(), seqs[s].end(), dseq.begin() + static_cast<size_t>(s) * kmax);
    kcount[s] = static_cast<int32_t>(seqs[s].size());
    for (int t = 0; t < n_sections; ++t) {
      int cnt = 0;
      for (double d : seqs[s])
        if (d < split[t + 1]) ++cnt;
      seg_end[static_cast<size_t>(s) * n_sections + t] = cnt;
    }
  }

  // Camera-frame directions (camera_model.cpp:161-168 as called at simple_ray_caster.cpp:41-44)
  a3d_caster* c = new a3d_caster();
  c->config_gen = ctx->config_gen;
  c->ctx = ctx;
  c->p = *p;
  c->info = info;
  c->h_dirs.resize(static_cast<size_t>(n_rays) * 3);
  for (int i = 0; i < info.c_res_x; ++i) {
    for (int j = 0; j < info.c_res_y; ++j) {
      const double rx = static_cast<double>(i) / (static_cast<double>(info.c_res_x) - 1.0);
      const double ry = static_cast<double>(j) / (static_cast<double>(info.c_res_y) - 1.0);
      double* o = &c->h_dirs[(static_cast<size_t>(i) * info.c_res_y + j) * 3];
      if (lidar) {  // LidarModel::getDirectionVector (lidar_model.cpp:159-165)
        const double phi = (0.5 - rx) * info.fov_x;
        const double theta = M_PI / 2.0 + (ry - 0.5) * info.fov_y;
        o[0] = std::sin(theta) * std::cos(phi);
        o[1] = std::sin(theta) * std::sin(phi);
        o[2] = std::cos(theta);
        continue;
      }
      const double vx = p->focal_length;
      const double vy = (0.5 - rx) * static_cast<double>(p->resolution_x);
      const double vz = (0.5 - ry) * static_cast<double>(p->resolution_y);
      const double nrm = std::sqrt((vx * vx + vy * vy) + vz * vz);
      o[0] = vx / nrm;
      o[1] = vy / nrm;
      o[2] = vz / nrm;
    }
  }
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  };
  cudaError_t e = up(reinterpret_cast<void**>(&c->d_dirs), c->h_dirs.data(), c->h_dirs.size() * 8);
  if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&c->d_dseq), dseq.data(), dseq.size() * 8);
  if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&c->d_kcount), kcount.data(), kcount.size() * 4);
  if (e == cudaSuccess) e = up(reinterpret_cast<void**>(&c->d_seg_end), seg_end.data(), seg_end.size() * 4);
  // leaf-supply units of the wavefront kernel
  std::vector<int32_t> units;
  int n_units = (n_rays + 31) / 32;
  if (iterative) {
    std::vector<int32_t> ud, ui;
    for (int dg = 0; dg < info.c_res_x + info.c_res_y - 1; ++dg) {
      const int i_lo = std::max(0, dg - (info.c_res_y - 1)), i_hi = std::min(info.c_res_x - 1, dg);
      for (int i0 = i_lo; i0 <= i_hi; i0 += 32) {
        ud.push_back(dg);
        ui.push_back(i0);
      }
    }
    n_units = static_cast<int>(ud.size());
    units = ud;
    units.insert(units.end(), ui.begin(), ui.end());
    if (e == cudaSuccess)
      e = up(reinterpret_cast<void**>(&c->d_units), units.data(), units.size() * 4);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    a3d_caster_destroy(c);
    return ctx->cuda_fail(e, "caster table upload");
  }
  CasterView& v = c->view;
  v.kind = iterative ? A3D_CASTER_ITERATIVE : A3D_CASTER_SIMPLE;  // the kernels know two schemes
  v.res_x = info.c_res_x;
  v.res_y = info.c_res_y;
  v.n_rays = n_rays;
  v.n_sections = n_sections;
  v.n_starts = n_starts;
  v.kmax = static_cast<int32_t>(kmax);
  v.cam_dirs = c->d_dirs;
  v.dseq = c->d_dseq;
  v.kcount = c->d_kcount;
  v.seg_end = c->d_seg_end;
  for (int i = 0; i < 3; ++i) v.mount_t[i] = p->mount_t[i];
  v.mount_q = Q4{p->mount_q[0], p->mount_q[1], p->mount_q[2], p->mount_q[3]};
  v.n_units = n_units;
  v.diag_max = std::min(info.c_res_x, info.c_res_y);
  v.unit_diag = c->d_units;
  v.unit_i0 = c->d_units ? c->d_units + n_units : nullptr;
  v.ray_step = info.ray_step;
  v.ray_length = p->ray_length;
  for (size_t t = 0; t < split.size(); ++t) v.split[t] = split[t];
  for (int s0 = 0; s0 < n_sections; ++s0) {
    v.leaf_d0[s0] = p->ray_length;
    for (double d : seqs[s0]) {
      if (!(d < split[n_sections - 1])) {
        v.leaf_d0[s0] = d;
        break;
      }
    }
  }
  *out = c;
  return A3D_OK;
}

void a3d_caster_destroy(a3d_caster* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  if (c->d_dirs) cudaFree(c->d_dirs);
  if (c->d_dseq) cudaFree(c->d_dseq);
  if (c->d_kcount) cudaFree(c->d_kcount);
  if (c->d_seg_end) cudaFree(c->d_seg_end);
  if (c->d_units) cudaFree(c->d_units);
  delete c;
}

int a3d_caster_get_info(const a3d_caster* c, a3d_caster_info* out) {
  if (!c || !out) return A3D_ERR_INVALID;
  *out = c->info;
  return A3D_OK;
}
int a3d_caster_direction(const a3d_caster* c, int i, int j, double* out3) {
  if (!c || !out3 || i < 0 || j < 0 || i >= c->info.c_res_x || j >= c->info.c_res_y)
    return A3D_ERR_INVALID;
  const double* d = &c->h_dirs[(static_cast<size_t>(i) * c->info.c_res_y + j) * 3];
  out3[0] = d[0];
  out3[1] = d[1];
  out3[2] = d[2];
  return A3D_OK;
}

int a3d_eval_create(a3d_ctx* ctx, const a3d_eval_params* p, a3d_eval** out) {
  if (!ctx || !out) return A3D_ERR_INVALID;
  *out = nullptr;
  if (!p) return ctx->fail(A3D_ERR_INVALID, "null params");
  if (!ctx->configured) return ctx->fail(A3D_ERR_STATE, "a3d_map_config must be called first");
  if (p->kind < A3D_EVAL_VOXEL_TYPE || p->kind > A3D_EVAL_VOXEL_WEIGHT)
    return ctx->fail(A3D_ERR_INVALID, "unknown evaluator kind");
  if (p->kind == A3D_EVAL_VOXEL_WEIGHT && !(p->ray_angle_x * p->ray_angle_y > 0.0))
    return ctx->fail(A3D_ERR_INVALID, "ray_angle_x * ray_angle_y expected > 0.0");
  a3d_eval* e = new a3d_eval();
  e->config_gen = ctx->config_gen;
  e->ctx = ctx;
  e->p = *p;
  e->view = make_eval_view(ctx, *p);
  *out = e;
  return A3D_OK;
}
void a3d_eval_destroy(a3d_eval* e) { delete e; }

// ---- hot call --------------------------------------------------------------------------------
int a3d_compute_gains_dev(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval, int n_views,
                          const double* pos_dev, const double* quat_dev,
                          const int32_t* view_first_dev, int n_segments, double* gains_dev,
                          uint64_t* n_visible_dev, uint64_t* n_samples_dev, uint64_t* digests_dev,
                          uint64_t* set_digests_dev, const double* seg_origin_dev) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!caster || !eval || caster->ctx != ctx || eval->ctx != ctx)
    return ctx->fail(A3D_ERR_INVALID, "caster/eval missing or from another context");
  CHECK_FRESH(caster);
  CHECK_FRESH(eval);
  if (eval->p.clear_from_parents)
    return ctx->fail(A3D_ERR_UNSUPPORTED,
                     "clear_from_parents needs the tree: use a3d_compute_gains_tree or a3d_compute_gains_stored");
  if (n_views < 0 || n_segments < 0) return ctx->fail(A3D_ERR_INVALID, "negative count");
  if (n_segments == 0) return A3D_OK;
  if (!pos_dev || !quat_dev || !view_first_dev || !gains_dev)
    return ctx->fail(A3D_ERR_INVALID, "null argument");
  CK(cudaSetDevice(ctx->device));
  return enqueue_cast(ctx, caster, eval, pos_dev, quat_dev, view_first_dev, n_segments, gains_dev,
                      reinterpret_cast<unsigned long long*>(n_visible_dev),
                      reinterpret_cast<unsigned long long*>(n_samples_dev),
                      reinterpret_cast<unsigned long long*>(digests_dev),
                      reinterpret_cast<unsigned long long*>(set_digests_dev), nullptr, 0, false,
                      seg_origin_dev, nullptr, nullptr, n_views);
}

int a3d_set_option(a3d_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return A3D_ERR_INVALID;
  if (std::strcmp(name, "cast_kernel") == 0) {
    if (value < 0 || value > 3) return ctx->fail(A3D_ERR_INVALID, "cast_kernel must be 0, 1, 2 or 3");
    ctx->opt_cast_kernel = static_cast<int>(value);
    return A3D_OK;
  }
  if (std::strcmp(name, "wave_cta") == 0) {
    if (value < 0 || value > 3) return ctx->fail(A3D_ERR_INVALID, "wave_cta must be 0, 1, 2 or 3");
    ctx->opt_wave_cta = static_cast<int>(value);
    return A3D_OK;
  }
  return ctx->fail(A3D_ERR_INVALID, std::string("unknown option ") + name);
}

int a3d_last_cast_kernel(const a3d_ctx* ctx) { return ctx ? ctx->last_cast_kernel : 0; }

int a3d_compute_gains(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval, int n_views,
                      const double* pos, const double* quat, const int32_t* view_segment,
                      int n_segments, double* gains_out, uint64_t* n_visible_out,
                      uint64_t* n_samples_out, uint64_t* digests_out, uint64_t* set_digests_out,
                      const double* seg_origin) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!caster || !eval || caster->ctx != ctx || eval->ctx != ctx)
    return ctx->fail(A3D_ERR_INVALID, "caster/eval missing or from another context");
  CHECK_FRESH(caster);
  CHECK_FRESH(eval);
  if (n_views < 0 || n_segments < 0) return ctx->fail(A3D_ERR_INVALID, "negative count");
  if (n_segments == 0) return A3D_OK;
  if ((n_views > 0 && (!pos || !quat || !view_segment)) || !gains_out)
    return ctx->fail(A3D_ERR_INVALID, "null argument");
  std::vector<int32_t> first(static_cast<size_t>(n_segments) + 1, 0);
  for (int v = 0; v < n_views; ++v) {
    const int s = view_segment[v];
    if (s < 0 || s >= n_segments) return ctx->fail(A3D_ERR_INVALID, "view_segment out of range");
    if (v > 0 && s < view_segment[v - 1])
      return ctx->fail(A3D_ERR_INVALID, "view_segment must be non-decreasing");
    first[s + 1]++;
  }
  for (int s = 0; s < n_segments; ++s) first[s + 1] += first[s];
  CK(cudaSetDevice(ctx->device));
  CK(ctx->sc_pos.reserve(3 * static_cast<size_t>(std::max(n_views, 1))));
  CK(ctx->sc_quat.reserve(4 * static_cast<size_t>(std::max(n_views, 1))));
  CK(ctx->sc_first.reserve(first.size()));
  CK(ctx->sc_gains.reserve(n_segments));
  CK(ctx->sc_nvis.reserve(n_segments));
  CK(ctx->sc_nsamp.reserve(n_segments));
  if (digests_out || set_digests_out) CK(ctx->sc_dig.reserve(2 * static_cast<size_t>(n_segments)));
  if (n_views > 0) {
    CK(cudaMemcpyAsync(ctx->sc_pos.p, pos, 3 * sizeof(double) * n_views, cudaMemcpyHostToDevice,
                       ctx->stream));
    CK(cudaMemcpyAsync(ctx->sc_quat.p, quat, 4 * sizeof(double) * n_views, cudaMemcpyHostToDevice,
                       ctx->stream));
  }
  CK(cudaMemcpyAsync(ctx->sc_first.p, first.data(), first.size() * sizeof(int32_t),
                     cudaMemcpyHostToDevice, ctx->stream));
  const double* origin_dev = nullptr;
  if (seg_origin) {
    CK(ctx->sc_origin.reserve(3 * static_cast<size_t>(n_segments)));
    CK(cudaMemcpyAsync(ctx->sc_origin.p, seg_origin, 24 * static_cast<size_t>(n_segments),
                       cudaMemcpyHostToDevice, ctx->stream));
    origin_dev = ctx->sc_origin.p;
  }
  int rc = a3d_compute_gains_dev(ctx, caster, eval, n_views, ctx->sc_pos.p, ctx->sc_quat.p,
                                 ctx->sc_first.p, n_segments, ctx->sc_gains.p,
                                 reinterpret_cast<uint64_t*>(ctx->sc_nvis.p),
                                 reinterpret_cast<uint64_t*>(ctx->sc_nsamp.p),
                                 digests_out ? reinterpret_cast<uint64_t*>(ctx->sc_dig.p) : nullptr,
                                 set_digests_out
                                     ? reinterpret_cast<uint64_t*>(ctx->sc_dig.p + n_segments)
                                     : nullptr,
                                 origin_dev);
  if (rc) return rc;
  CK(cudaMemcpyAsync(gains_out, ctx->sc_gains.p, sizeof(double) * n_segments,
                     cudaMemcpyDeviceToHost, ctx->stream));
  if (n_visible_out)
    CK(cudaMemcpyAsync(n_visible_out, ctx->sc_nvis.p, 8 * static_cast<size_t>(n_segments),
                       cudaMemcpyDeviceToHost, ctx->stream));
  if (n_samples_out)
    CK(cudaMemcpyAsync(n_samples_out, ctx->sc_nsamp.p, 8 * static_cast<size_t>(n_segments),
                       cudaMemcpyDeviceToHost, ctx->stream));
  if (digests_out)
    CK(cudaMemcpyAsync(digests_out, ctx->sc_dig.p, 8 * static_cast<size_t>(n_segments),
                       cudaMemcpyDeviceToHost, ctx->stream));
  if (set_digests_out)
    CK(cudaMemcpyAsync(set_digests_out, ctx->sc_dig.p + n_segments,
                       8 * static_cast<size_t>(n_segments), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return A3D_OK;
}

int a3d_visible_voxels(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval, int n_views,
                       const double* pos, const double* quat, double* centres_out, int64_t cap,
                       int64_t* n_out, int64_t* n_samples_out, double* gain_out, const double* origin) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!caster || caster->ctx != ctx || (eval && eval->ctx != ctx))
    return ctx->fail(A3D_ERR_INVALID, "caster/eval missing or from another context");
  CHECK_FRESH(caster);
  CHECK_FRESH(eval);
  if (n_views < 0 || cap < 0 || (n_views > 0 && (!pos || !quat)) || (cap > 0 && !centres_out) ||
      !n_out)
    return ctx->fail(A3D_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(ctx->device));
  const int32_t first[2] = {0, n_views};
  CK(ctx->sc_pos.reserve(3 * static_cast<size_t>(std::max(n_views, 1))));
  CK(ctx->sc_quat.reserve(4 * static_cast<size_t>(std::max(n_views, 1))));
  CK(ctx->sc_first.reserve(2));
  CK(ctx->sc_gains.reserve(1));
  CK(ctx->sc_nvis.reserve(1));
  CK(ctx->sc_nsamp.reserve(1));
  CK(ctx->sc_centres.reserve(3 * static_cast<size_t>(std::max<int64_t>(cap, 1))));
  if (n_views > 0) {
    CK(cudaMemcpyAsync(ctx->sc_pos.p, pos, 3 * sizeof(double) * n_views, cudaMemcpyHostToDevice,
                       ctx->stream));
    CK(cudaMemcpyAsync(ctx->sc_quat.p, quat, 4 * sizeof(double) * n_views, cudaMemcpyHostToDevice,
                       ctx->stream));
  }
  CK(cudaMemcpyAsync(ctx->sc_first.p, first, sizeof(first), cudaMemcpyHostToDevice, ctx->stream));
  const double* origin_dev = nullptr;
  if (origin) {
    CK(ctx->sc_origin.reserve(3));
    CK(cudaMemcpyAsync(ctx->sc_origin.p, origin, 24, cudaMemcpyHostToDevice, ctx->stream));
    origin_dev = ctx->sc_origin.p;
  }
  int rc = enqueue_cast(ctx, caster, eval, ctx->sc_pos.p, ctx->sc_quat.p, ctx->sc_first.p, 1,
                        ctx->sc_gains.p, ctx->sc_nvis.p, ctx->sc_nsamp.p, nullptr, nullptr,
                        ctx->sc_centres.p, cap, true, origin_dev, nullptr, nullptr, n_views);
  if (rc) return rc;
  unsigned long long nvis = 0, nsamp = 0;
  double gain = 0.0;
  CK(cudaMemcpyAsync(&nvis, ctx->sc_nvis.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&nsamp, ctx->sc_nsamp.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&gain, ctx->sc_gains.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const int64_t n_copy = std::min<int64_t>(static_cast<int64_t>(nvis), cap);
  if (n_copy > 0) {
    CK(cudaMemcpyAsync(centres_out, ctx->sc_centres.p, 24 * static_cast<size_t>(n_copy),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  *n_out = static_cast<int64_t>(nvis);
  if (n_samples_out) *n_samples_out = static_cast<int64_t>(nsamp);
  if (gain_out) *gain_out = gain;
  if (static_cast<int64_t>(nvis) > cap)
    return ctx->fail(A3D_ERR_CAPACITY, "centres_out too small for the voxel list");
  return A3D_OK;
}

int a3d_gain_from_voxels(a3d_ctx* ctx, const a3d_eval* eval, int64_t n, const double* centres,
                         double* gain_out, int64_t* n_left_out, uint8_t* keep_out, const double* origin) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!eval || eval->ctx != ctx) return ctx->fail(A3D_ERR_INVALID, "eval missing");
  CHECK_FRESH(eval);
  if (n < 0 || (n > 0 && !centres) || !gain_out) return ctx->fail(A3D_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(ctx->device));
  if (ctx->h_table.empty()) {
    table_rebuild(ctx, 0);
    int rc = table_upload(ctx);
    if (rc) return rc;
  }
  unsigned long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (eval->p.kind == A3D_EVAL_VOXEL_WEIGHT && !origin)
    return ctx->fail(A3D_ERR_INVALID, "VoxelWeightEvaluator needs the segment origin");
  const D3 o = origin ? D3{origin[0], origin[1], origin[2]} : D3{0.0, 0.0, 0.0};
  if (n > 0) {
    CK(ctx->sc_pts.reserve(3 * n));
    CK(ctx->sc_cnt.reserve(8));
    CK(ctx->sc_u8.reserve(n));
    CK(cudaMemcpyAsync(ctx->sc_pts.p, centres, 24 * static_cast<size_t>(n), cudaMemcpyHostToDevice,
                       ctx->stream));
    CK(cudaMemsetAsync(ctx->sc_cnt.p, 0, 64, ctx->stream));
    // counters[7] doubles as the double accumulator of the VoxelWeight gain
    launch_gain_from_voxels(ctx->stream, ctx->view(), eval->view, n, ctx->sc_pts.p, ctx->sc_cnt.p,
                            keep_out ? ctx->sc_u8.p : nullptr, o,
                            reinterpret_cast<double*>(ctx->sc_cnt.p + 7));
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cnt, ctx->sc_cnt.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
    if (keep_out)
      CK(cudaMemcpyAsync(keep_out, ctx->sc_u8.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if (eval->p.kind == A3D_EVAL_VOXEL_WEIGHT) {
    double g = 0.0;
    std::memcpy(&g, &cnt[7], sizeof(double));
    *gain_out = g;
    if (n_left_out) *n_left_out = n;  // the list is not modified (voxel_weight_evaluator.cpp:43-58)
    return A3D_OK;
  }
  if (eval->p.kind == A3D_EVAL_VOXEL_TYPE) {
    // fall-through sums of voxel_type_evaluator.cpp:69-85
    double g = 0.0;
    for (int o = 0; o < 2; ++o) {
      const double n_unk = static_cast<double>(cnt[2 + 3 * o]);
      const double n_free = static_cast<double>(cnt[1 + 3 * o] + cnt[2 + 3 * o]);
      const double n_occ = static_cast<double>(cnt[0 + 3 * o] + cnt[1 + 3 * o] + cnt[2 + 3 * o]);
      g += n_unk * eval->p.gains[0 + 3 * o];
      g += n_free * eval->p.gains[2 + 3 * o];
      g += n_occ * eval->p.gains[1 + 3 * o];
    }
    *gain_out = g;
  } else {
    *gain_out = static_cast<double>(cnt[0]);
  }
  if (n_left_out) *n_left_out = static_cast<int64_t>(cnt[6]);
  return A3D_OK;
}

int a3d_store_create(a3d_ctx* ctx, a3d_store** out) {
  if (!ctx || !out) return A3D_ERR_INVALID;
  *out = new a3d_store();
  (*out)->ctx = ctx;
  return A3D_OK;
}
void a3d_store_destroy(a3d_store* store) {
  if (!store) return;
  cudaSetDevice(store->ctx->device);
  cudaStreamSynchronize(store->ctx->stream);
  for (auto& a : store->arenas)
    if (a.p) store->free_arena(a);
  cudaStreamSynchronize(store->ctx->stream);
  store->d_off.release();
  store->d_cnt.release();
  store->d_gain.release();
  store->d_origin.release();
  store->d_u64.release();
  store->d_anc_first.release();
  store->d_tab.release();
  store->d_mask.release();
  delete store;
}
int a3d_store_erase(a3d_store* store, int n, const uint64_t* seg_keys) {
  if (!store || n < 0 || (n > 0 && !seg_keys)) return A3D_ERR_INVALID;
  cudaSetDevice(store->ctx->device);
  cudaStreamSynchronize(store->ctx->stream);
  for (int i = 0; i < n; ++i) store->drop(seg_keys[i]);
  return A3D_OK;
}
int64_t a3d_store_list_size(const a3d_store* store, uint64_t key) {
  if (!store) return -1;
  auto it = store->lists.find(key);
  return it == store->lists.end() ? -1 : it->second.count;
}
int64_t a3d_store_bytes(const a3d_store* store) {
  int64_t b = 0;
  if (store)
    for (const auto& a : store->arenas) b += static_cast<int64_t>(a.entries) * 8;
  return b;
}
int a3d_store_get(a3d_ctx* ctx, const a3d_store* store, uint64_t key, double* centres_out, int64_t cap,
                  int64_t* n_out) {
  if (!ctx || !store || store->ctx != ctx || !n_out || cap < 0 || (cap > 0 && !centres_out))
    return ctx ? ctx->fail(A3D_ERR_INVALID, "bad argument") : A3D_ERR_INVALID;
  auto it = store->lists.find(key);
  if (it == store->lists.end()) return ctx->fail(A3D_ERR_INVALID, "key not in store");
  *n_out = it->second.count;
  const int64_t n = std::min<int64_t>(cap, it->second.count);
  CK(cudaSetDevice(ctx->device));
  if (n > 0) {
    // keys → the centre doubles the reference holds in SimulatedSensorInfo::visible_voxels
    CK(ctx->sc_centres.reserve(3 * static_cast<size_t>(n)));
    launch_keys_to_centres(ctx->stream, ctx->view(), n, store->arenas[it->second.arena].p + it->second.offset,
                           ctx->sc_centres.p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(centres_out, ctx->sc_centres.p, 24 * static_cast<size_t>(n), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return it->second.count > cap ? ctx->fail(A3D_ERR_CAPACITY, "centres_out too small") : A3D_OK;
}

namespace {

uint32_t set_size_for(long long count) {
  uint32_t sz = 16;
  while (sz < 2ull * static_cast<unsigned long long>(std::max<long long>(count, 1))) sz <<= 1;
  return sz;
}

// (Re)builds the hash sets of the given stored lists where missing or stale, one arena for the new ones.
int ensure_sets(a3d_ctx* ctx, a3d_store* store, const std::vector<uint64_t>& keys) {
  std::vector<StoreEntry*> todo;
  size_t new_words = 0;
  for (uint64_t k : keys) {
    auto it = store->lists.find(k);
    if (it == store->lists.end()) return ctx->fail(A3D_ERR_INVALID, "parent key not in store");
    StoreEntry& e = it->second;
    if (e.set_arena >= 0 && !e.set_stale) continue;
    if (std::find(todo.begin(), todo.end(), &e) != todo.end()) continue;
    todo.push_back(&e);
    if (e.set_arena < 0) new_words += set_size_for(e.count);
  }
  if (todo.empty()) return A3D_OK;
  int arena = -1;
  if (new_words > 0) {
    cudaError_t err;
    arena = store->new_arena(new_words, &err);
    if (arena < 0) return ctx->cuda_fail(err, "cudaMalloc(hash sets)");
  }
  long long cursor = 0;
  const int n = static_cast<int>(todo.size());
  std::vector<long long> cnt(n);
  std::vector<unsigned long long*> tabs(2 * static_cast<size_t>(n));  // [tables | lists] (lists may sit in different arenas)
  std::vector<uint32_t> masks(n);
  for (int i = 0; i < n; ++i) {
    StoreEntry& e = *todo[i];
    if (e.set_arena < 0) {
      e.set_arena = arena;
      e.set_off = cursor;
      e.set_mask = set_size_for(e.count) - 1;
      cursor += static_cast<long long>(e.set_mask) + 1;
      store->arenas[arena].live++;
    }
    tabs[i] = store->arenas[e.set_arena].p + e.set_off;
    masks[i] = e.set_mask;
    CK(cudaMemsetAsync(tabs[i], 0xFF, (static_cast<size_t>(e.set_mask) + 1) * 8, ctx->stream));
    tabs[n + i] = store->arenas[e.arena].p + e.offset;
    cnt[i] = e.count;
    e.set_stale = false;
  }
  CK(store->d_cnt.reserve(n));
  CK(store->d_tab.reserve(tabs.size()));
  CK(store->d_mask.reserve(n));
  CK(cudaMemcpyAsync(store->d_cnt.p, cnt.data(), n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(store->d_tab.p, tabs.data(), tabs.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(store->d_mask.p, masks.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  launch_build_sets(ctx->stream, n, store->d_tab.p + n, store->d_cnt.p, store->d_tab.p, store->d_mask.p);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));  // host vectors above are pageable
  return A3D_OK;
}

// k_list_pipeline over the given stored lists (all in one arena): parent filter (when `parents`),
// digests, fold.  Results come back in the order of `keys`.
int run_list_pipeline(a3d_ctx* ctx, const a3d_eval* eval, a3d_store* store, const std::vector<uint64_t>& keys,
                      bool parents, bool fold, const double* seg_origin /*per key, nullable*/,
                      double* gains_out, uint64_t* n_stored_out, uint64_t* digests_out, uint64_t* set_digests_out) {
  const int n = static_cast<int>(keys.size());
  if (n == 0) return A3D_OK;
  std::vector<StoreEntry*> ent(n);
  for (int i = 0; i < n; ++i) ent[i] = &store->lists.at(keys[i]);
  const int arena = ent[0]->arena;
  std::vector<long long> off(n), cnt(n);
  std::vector<int> anc_first(n + 1, 0);
  std::vector<unsigned long long*> tabs;
  std::vector<uint32_t> masks;
  if (parents) {
    std::vector<uint64_t> anc_keys;
    for (int i = 0; i < n; ++i)
      for (uint64_t k = ent[i]->parent; k != 0;) {
        auto it = store->lists.find(k);
        if (it == store->lists.end()) return ctx->fail(A3D_ERR_INVALID, "parent key not in store");
        anc_keys.push_back(k);
        k = it->second.parent;
        if (anc_keys.size() > (size_t(1) << 26)) return ctx->fail(A3D_ERR_INVALID, "parent chain does not end");
      }
    int rc = ensure_sets(ctx, store, anc_keys);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
      for (uint64_t k = ent[i]->parent; k != 0; k = store->lists.at(k).parent) {
        const StoreEntry& a = store->lists.at(k);
        tabs.push_back(store->arenas[a.set_arena].p + a.set_off);
        masks.push_back(a.set_mask);
      }
      anc_first[i + 1] = static_cast<int>(tabs.size());
    }
  }
  for (int i = 0; i < n; ++i) {
    if (ent[i]->arena != arena) return ctx->fail(A3D_ERR_STATE, "lists of one pipeline launch span arenas");
    off[i] = ent[i]->offset;
    cnt[i] = ent[i]->count;
  }
  CK(store->d_off.reserve(n));
  CK(store->d_cnt.reserve(n));
  CK(store->d_gain.reserve(n));
  CK(store->d_u64.reserve(3 * static_cast<size_t>(n)));
  CK(cudaMemcpyAsync(store->d_off.p, off.data(), n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(store->d_cnt.p, cnt.data(), n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  ListBatch lb{};
  lb.n = n;
  lb.pool = store->arenas[arena].p;
  lb.offsets = store->d_off.p;
  lb.counts = store->d_cnt.p;
  lb.n_stored = store->d_u64.p;
  lb.digests = digests_out ? store->d_u64.p + n : nullptr;
  lb.set_digests = set_digests_out ? store->d_u64.p + 2 * static_cast<size_t>(n) : nullptr;
  lb.gains = fold ? store->d_gain.p : nullptr;
  if (seg_origin) {
    CK(store->d_origin.reserve(3 * static_cast<size_t>(n)));
    CK(cudaMemcpyAsync(store->d_origin.p, seg_origin, 24 * static_cast<size_t>(n), cudaMemcpyHostToDevice,
                       ctx->stream));
    lb.origins = store->d_origin.p;
  }
  if (parents && !tabs.empty()) {
    CK(store->d_anc_first.reserve(n + 1));
    CK(store->d_tab.reserve(tabs.size()));
    CK(store->d_mask.reserve(masks.size()));
    CK(cudaMemcpyAsync(store->d_anc_first.p, anc_first.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice,
                       ctx->stream));
    CK(cudaMemcpyAsync(store->d_tab.p, tabs.data(), tabs.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(store->d_mask.p, masks.data(), masks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                       ctx->stream));
    lb.anc_first = store->d_anc_first.p;
    lb.anc_table = store->d_tab.p;
    lb.anc_mask = store->d_mask.p;
  }
  if (ctx->h_table.empty()) {
    table_rebuild(ctx, 0);
    int rc = table_upload(ctx);
    if (rc) return rc;
  }
  int rc = vol_ensure(ctx);
  if (rc) return rc;
  launch_list_pipeline(ctx->stream, ctx->view(), eval->view, lb);
  ctx->launches++;
  CK(cudaGetLastError());
  std::vector<unsigned long long> u(3 * static_cast<size_t>(n));
  std::vector<double> g(n);
  CK(cudaMemcpyAsync(cnt.data(), store->d_cnt.p, n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(u.data(), store->d_u64.p, u.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (fold) CK(cudaMemcpyAsync(g.data(), store->d_gain.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) {
    if (cnt[i] != ent[i]->count) ent[i]->set_stale = true;
    ent[i]->count = cnt[i];
    if (fold && gains_out) gains_out[i] = g[i];
    if (n_stored_out) n_stored_out[i] = u[i];
    if (digests_out) digests_out[i] = u[n + i];
    if (set_digests_out) set_digests_out[i] = u[2 * static_cast<size_t>(n) + i];
  }
  return A3D_OK;
}

}  // namespace

int a3d_compute_gains_stored(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval,
                             a3d_store* store, int n_views, const double* pos, const double* quat,
                             const int32_t* view_segment, int n_segments, const uint64_t* seg_keys,
                             const uint64_t* parent_keys, double* gains_out, uint64_t* n_visible_out,
                             uint64_t* n_samples_out, uint64_t* digests_out, uint64_t* set_digests_out,
                             const double* seg_origin) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!store || store->ctx != ctx || (n_segments > 0 && !seg_keys))
    return ctx->fail(A3D_ERR_INVALID, "store/keys missing");
  if (!eval) return ctx->fail(A3D_ERR_INVALID, "eval missing");
  const bool chained = eval->p.clear_from_parents && parent_keys;
  {
    std::unordered_map<uint64_t, int> seen;
    for (int s = 0; s < n_segments; ++s) {
      if (seg_keys[s] == 0) return ctx->fail(A3D_ERR_INVALID, "segment key 0 is reserved (no parent)");
      if (!seen.emplace(seg_keys[s], s).second) return ctx->fail(A3D_ERR_INVALID, "duplicate segment key");
    }
  }
  // pass 1: gains (final unless the list pipeline folds below), raw list sizes, sample counts
  std::vector<uint64_t> nvis(static_cast<size_t>(std::max(n_segments, 0)));
  a3d_eval plain = *eval;
  plain.p.clear_from_parents = 0;  // (the batch cast itself never looks at parents)
  ctx->want_part_offsets = true;
  int rc = a3d_compute_gains(ctx, caster, &plain, n_views, pos, quat, view_segment, n_segments, gains_out,
                             nvis.data(), n_samples_out, nullptr, nullptr, seg_origin);
  ctx->want_part_offsets = false;
  const bool direct = ctx->part_offsets_valid;  // the chain ran in one chunk: its scratch knows where every part goes
  ctx->part_offsets_valid = false;
  if (rc) return rc;
  if (n_segments == 0) return A3D_OK;
  std::vector<long long> off(static_cast<size_t>(n_segments) + 1, 0);
  for (int s = 0; s < n_segments; ++s) off[s + 1] = off[s] + static_cast<long long>(nvis[s]);
  const size_t total = static_cast<size_t>(off[n_segments]);
  // pass 2: ordered lists as packed voxel keys into one exactly sized arena (inputs are still in the
  // context scratch)
  cudaError_t merr;
  const int arena_id = store->new_arena(total, &merr);
  if (arena_id < 0) return ctx->cuda_fail(merr, "cudaMalloc(list arena)");
  auto fail_arena = [&](int code) {
    store->free_arena(store->arenas[arena_id]);
    return code;
  };
  if (cudaSuccess != store->d_off.reserve(off.size())) return fail_arena(ctx->fail(A3D_ERR_CUDA, "cudaMalloc"));
  CK(cudaMemcpyAsync(store->d_off.p, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice,
                     ctx->stream));
  if (direct)
    rc = enqueue_direct_emit(ctx, caster, &plain, n_segments, n_views, seg_origin ? ctx->sc_origin.p : nullptr,
                             store->d_off.p, store->arenas[arena_id].p);
  else
    rc = enqueue_cast(ctx, caster, &plain, ctx->sc_pos.p, ctx->sc_quat.p, ctx->sc_first.p, n_segments,
                      ctx->sc_gains.p, ctx->sc_nvis.p, ctx->sc_nsamp.p, nullptr, nullptr, nullptr, 0, true,
                      seg_origin ? ctx->sc_origin.p : nullptr, store->d_off.p, store->arenas[arena_id].p, n_views);
  if (rc) return fail_arena(rc);
  unsigned int key_error = 0;
  CK(cudaMemcpyAsync(&key_error, ctx->sc_counter.p + 2, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail_arena(ctx->fail(A3D_ERR_CUDA, "list emission failed"));
  if (key_error)
    return fail_arena(ctx->fail(A3D_ERR_UNSUPPORTED,
                                "a voxel coordinate does not fit the 20-bit list keys (pose beyond 2^19 voxels from the origin)"));
  // register the lists (old lists of the same keys go first, so the new arena is never released here)
  for (int s = 0; s < n_segments; ++s) store->drop(seg_keys[s]);
  for (int s = 0; s < n_segments; ++s) {
    StoreEntry e{};
    e.arena = arena_id;
    e.offset = off[s];
    e.count = static_cast<long long>(nvis[s]);
    e.parent = chained ? parent_keys[s] : 0;
    store->lists[seg_keys[s]] = e;
    store->arenas[arena_id].live++;
  }
  // list pipeline: needed when parents filter the lists, when the fold erases entries (Naive and
  // Frontier store the post-erasure list, naive_evaluator.cpp:28-35) or when digests are asked
  const bool erases = eval->p.kind == A3D_EVAL_NAIVE || eval->p.kind == A3D_EVAL_FRONTIER;
  if (chained || erases || digests_out || set_digests_out) {
    // levels: a segment is ready when its parent is not in this batch or was processed already
    std::unordered_map<uint64_t, int> in_batch;
    for (int s = 0; s < n_segments; ++s) in_batch.emplace(seg_keys[s], s);
    std::vector<int> level(n_segments, 0);
    int max_level = 0;
    if (chained) {
      for (int s = 0; s < n_segments; ++s) {
        if (parent_keys[s] != 0 && !in_batch.count(parent_keys[s]) && !store->lists.count(parent_keys[s]))
          return ctx->fail(A3D_ERR_INVALID, "parent key neither stored nor in this batch");
      }
      // (depth by following parents; chains are short, cycles are rejected)
      for (int s = 0; s < n_segments; ++s) {
        int d = 0;
        uint64_t k = parent_keys[s];
        while (k != 0) {
          auto it = in_batch.find(k);
          if (it == in_batch.end()) break;
          k = parent_keys[it->second];
          if (++d > n_segments) return ctx->fail(A3D_ERR_INVALID, "parent keys form a cycle");
        }
        level[s] = d;
        max_level = std::max(max_level, d);
      }
    }
    for (int lv = 0; lv <= max_level; ++lv) {
      std::vector<uint64_t> keys;
      std::vector<int> idx;
      std::vector<double> org;
      for (int s = 0; s < n_segments; ++s)
        if (level[s] == lv) {
          keys.push_back(seg_keys[s]);
          idx.push_back(s);
          if (eval->p.kind == A3D_EVAL_VOXEL_WEIGHT) {
            // the segment's reference point: given, or the body position of its last view
            const double* o = nullptr;
            if (seg_origin) {
              o = seg_origin + 3 * s;
            } else {
              int last = -1;
              for (int v = 0; v < n_views; ++v)
                if (view_segment[v] == s) last = v;
              o = last >= 0 ? pos + 3 * last : nullptr;
            }
            for (int c = 0; c < 3; ++c) org.push_back(o ? o[c] : 0.0);
          }
        }
      const int m = static_cast<int>(keys.size());
      std::vector<double> g(m);
      std::vector<uint64_t> ns(m), dg(m), sd(m);
      rc = run_list_pipeline(ctx, eval, store, keys, chained, true, org.empty() ? nullptr : org.data(), g.data(),
                             ns.data(), digests_out ? dg.data() : nullptr, set_digests_out ? sd.data() : nullptr);
      if (rc) return rc;
      for (int i = 0; i < m; ++i) {
        gains_out[idx[i]] = g[i];
        nvis[idx[i]] = ns[i];
        if (digests_out) digests_out[idx[i]] = dg[i];
        if (set_digests_out) set_digests_out[idx[i]] = sd[i];
      }
    }
  }
  if (n_visible_out) std::copy(nvis.begin(), nvis.end(), n_visible_out);
  return A3D_OK;
}

int a3d_compute_gains_tree(a3d_ctx* ctx, const a3d_caster* caster, const a3d_eval* eval, int n_views,
                           const double* pos, const double* quat, const int32_t* view_segment,
                           int n_segments, const int32_t* parent_index, double* gains_out,
                           uint64_t* n_visible_out, uint64_t* n_samples_out, uint64_t* digests_out,
                           uint64_t* set_digests_out, const double* seg_origin) {
  if (!ctx) return A3D_ERR_INVALID;
  if (n_segments < 0) return ctx->fail(A3D_ERR_INVALID, "negative count");
  std::vector<uint64_t> keys(static_cast<size_t>(n_segments)), parents(static_cast<size_t>(n_segments), 0);
  for (int s = 0; s < n_segments; ++s) {
    keys[s] = static_cast<uint64_t>(s) + 1;
    if (parent_index) {
      if (parent_index[s] >= s || parent_index[s] < -1)
        return ctx->fail(A3D_ERR_INVALID, "parent_index must be -1 or the index of an earlier segment");
      parents[s] = parent_index[s] < 0 ? 0 : static_cast<uint64_t>(parent_index[s]) + 1;
    }
  }
  a3d_store* tmp = nullptr;
  int rc = a3d_store_create(ctx, &tmp);
  if (rc) return rc;
  rc = a3d_compute_gains_stored(ctx, caster, eval, tmp, n_views, pos, quat, view_segment, n_segments, keys.data(),
                                parent_index ? parents.data() : nullptr, gains_out, n_visible_out, n_samples_out,
                                digests_out, set_digests_out, seg_origin);
  a3d_store_destroy(tmp);
  return rc;
}

int a3d_store_refold(a3d_ctx* ctx, const a3d_eval* eval, a3d_store* store, int n,
                     const uint64_t* seg_keys, double* gains_out, uint64_t* n_left_out,
                     const double* seg_origin) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!store || store->ctx != ctx || !eval || eval->ctx != ctx || n < 0 ||
      (n > 0 && (!seg_keys || !gains_out)))
    return ctx->fail(A3D_ERR_INVALID, "bad argument");
  CHECK_FRESH(eval);
  if (eval->p.kind == A3D_EVAL_VOXEL_WEIGHT && !seg_origin)
    return ctx->fail(A3D_ERR_INVALID, "VoxelWeightEvaluator needs the segment origins");
  if (n == 0) return A3D_OK;
  CK(cudaSetDevice(ctx->device));
  // group the requested lists by arena (one launch per arena; typically one)
  std::vector<int> order(n);
  {
    std::unordered_map<uint64_t, int> seen;
    for (int i = 0; i < n; ++i) {
      if (!store->lists.count(seg_keys[i])) return ctx->fail(A3D_ERR_INVALID, "key not in store");
      if (!seen.emplace(seg_keys[i], i).second) return ctx->fail(A3D_ERR_INVALID, "duplicate segment key");
      order[i] = i;
    }
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return store->lists.at(seg_keys[a]).arena < store->lists.at(seg_keys[b]).arena;
  });
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int k0 = 0; k0 < n;) {
    int k1 = k0;
    const int arena = store->lists.at(seg_keys[order[k0]]).arena;
    while (k1 < n && store->lists.at(seg_keys[order[k1]]).arena == arena) ++k1;
    const int m = k1 - k0;
    std::vector<uint64_t> keys(m);
    std::vector<double> org(seg_origin ? 3 * static_cast<size_t>(m) : 0), g(m);
    for (int i = 0; i < m; ++i) {
      keys[i] = seg_keys[order[k0 + i]];
      if (seg_origin)
        for (int c = 0; c < 3; ++c) org[3 * i + c] = seg_origin[3 * order[k0 + i] + c];
    }
    int rc = run_list_pipeline(ctx, eval, store, keys, false, true, seg_origin ? org.data() : nullptr, g.data(), nullptr,
                               nullptr, nullptr);
    if (rc) return rc;
    for (int i = 0; i < m; ++i) {
      gains_out[order[k0 + i]] = g[i];
      if (n_left_out) n_left_out[order[k0 + i]] = static_cast<uint64_t>(store->lists.at(keys[i]).count);
    }
    k0 = k1;
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  return A3D_OK;
}

int a3d_filter_not_in(a3d_ctx* ctx, int64_t n, const double* centres, int64_t m,
                      const double* other, uint8_t* keep_inout) {
  if (!ctx) return A3D_ERR_INVALID;
  if (n < 0 || m < 0 || (n > 0 && (!centres || !keep_inout)) || (m > 0 && !other))
    return ctx->fail(A3D_ERR_INVALID, "bad argument");
  if (n == 0 || m == 0) return A3D_OK;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->sc_pts.reserve(3 * n));
  CK(ctx->sc_out.reserve(3 * m));
  CK(ctx->sc_u8.reserve(n));
  if (m >= (1ll << 30)) return ctx->fail(A3D_ERR_INVALID, "list too long");
  CK(ctx->sc_first.reserve(filter_table_slots(m)));  // (hash set of the ancestor's entries)
  CK(cudaMemcpyAsync(ctx->sc_pts.p, centres, 24 * static_cast<size_t>(n), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaMemcpyAsync(ctx->sc_out.p, other, 24 * static_cast<size_t>(m), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaMemcpyAsync(ctx->sc_u8.p, keep_inout, n, cudaMemcpyHostToDevice, ctx->stream));
  launch_filter_not_in(ctx->stream, n, ctx->sc_pts.p, m, ctx->sc_out.p, ctx->sc_u8.p, ctx->sc_first.p);
  ctx->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(keep_inout, ctx->sc_u8.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return A3D_OK;
}

int a3d_bv_contains(a3d_ctx* ctx, const a3d_eval* eval, int64_t n, const double* pts, uint8_t* out) {
  if (!ctx) return A3D_ERR_INVALID;
  if (!eval || eval->ctx != ctx) return ctx->fail(A3D_ERR_INVALID, "eval missing");
  CHECK_FRESH(eval);
  if (n < 0 || (n > 0 && (!pts || !out))) return ctx->fail(A3D_ERR_INVALID, "bad argument");
  if (n == 0) return A3D_OK;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->sc_pts.reserve(3 * n));
  CK(ctx->sc_u8.reserve(n));
  CK(cudaMemcpyAsync(ctx->sc_pts.p, pts, 24 * static_cast<size_t>(n), cudaMemcpyHostToDevice,
                     ctx->stream));
  launch_bv_contains(ctx->stream, eval->view, n, ctx->sc_pts.p, ctx->sc_u8.p);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->sc_u8.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return A3D_OK;
}

}  // extern "C"
