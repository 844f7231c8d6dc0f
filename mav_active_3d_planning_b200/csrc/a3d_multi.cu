// a3d_multi.cu — the N-GPU form of the boundary (include/a3d.h "multi-GPU group", SURVEY.md §8(b),(e)).
//
// One host thread of one process drives N GPUs: a group owns one a3d_ctx per device (map mirror
// replicated on every GPU) and one NCCL communicator per device created with ncclCommInitAll.
//   * map updates   host → GPU 0 staging once (PCIe), ncclBroadcast over NVLink to the other GPUs'
//                   staging, then every GPU applies the delta to its own mirror (K1 kernels) — the
//                   per-GPU applies run concurrently from short-lived host threads;
//   * gains         segments are dealt to the GPUs in blocks (all views of a segment stay together:
//                   std::unique and the gain sum span them, camera_model.cpp:58); every GPU casts its
//                   share, the per-segment result records are ncclAllGather-ed and GPU 0 returns them
//                   in the caller's segment order.
// NCCL is loaded with dlopen at group creation, so a single-GPU process never needs it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/a3d.h"

namespace {

// ---- the slice of the NCCL API this file uses (nccl.h 2.x; types restated to avoid a build-time
//      dependency on the header's location) ------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;  // ncclSuccess == 0
enum { kNcclChar = 0, kNcclUint64 = 5 };  // ncclInt8 / ncclUint64 in ncclDataType_t

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string load() {
    if (lib) return "";
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return std::string("cannot load NCCL: ") + dlerror();
#define A3D_NCCL_SYM(field, name)                                      \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name));         \
  if (!field) return std::string("NCCL symbol missing: ") + name;
    A3D_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    A3D_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    A3D_NCCL_SYM(GroupStart, "ncclGroupStart")
    A3D_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    A3D_NCCL_SYM(Broadcast, "ncclBroadcast")
    A3D_NCCL_SYM(AllGather, "ncclAllGather")
    A3D_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef A3D_NCCL_SYM
    return "";
  }
};
Nccl g_nccl;
std::string g_group_create_error;

template <typename T>
struct Buf {  // grow-only device buffer on a fixed device
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct Pinned {  // grow-only page-locked host buffer: copies from / to it are truly asynchronous
  uint8_t* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = std::max<size_t>(n + n / 2, 4096);
    const cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), want, cudaHostAllocPortable);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

constexpr int kRecWords = 5;  // per segment: gain bits | n_visible | n_samples | digest | set digest

struct Device {
  int id = 0;
  a3d_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;  // the context's stream: NCCL and kernels are ordered on it
  Buf<uint8_t> st_bytes;          // delta staging: distances | weights | observed
  Buf<double> pos, quat, origin;
  Buf<int32_t> first;
  Buf<unsigned long long> rec, all;  // this GPU's records (per×5 words) / everyone's
  Pinned stage;                      // this GPU's share of a batch: pos | quat | origin | first
};

}  // namespace

struct a3d_group {
  std::vector<Device> dev;
  Pinned result;  // gathered records on their way back to the caller
  std::string err;
  int block = 64;  // segments are dealt to the GPUs in blocks of this many
  float voxel_size = 0.0f;
  int vps = 0, nv = 0;
  int fail(int code, const std::string& m) {
    err = m;
    return code;
  }
};
struct a3d_group_caster {
  a3d_group* g;
  std::vector<a3d_caster*> c;
};
struct a3d_group_eval {
  a3d_group* g;
  std::vector<a3d_eval*> e;
};

#define GCK(call)                                                                     \
  do {                                                                                \
    cudaError_t _e = (call);                                                          \
    if (_e != cudaSuccess) return g->fail(A3D_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)
#define NCK(call)                                                                     \
  do {                                                                                \
    ncclResult_t _r = (call);                                                         \
    if (_r != 0) return g->fail(A3D_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

namespace {
int sub_fail(a3d_group* g, int i, int rc) {
  g->err = "GPU " + std::to_string(g->dev[i].id) + ": " + a3d_last_error(g->dev[i].ctx);
  return rc;
}
// runs fn(i) for every device concurrently (host threads: the per-context calls block on their own stream)
template <class F>
int for_each_device(a3d_group* g, F fn) {
  const int n = static_cast<int>(g->dev.size());
  std::vector<int> rc(n, A3D_OK);
  std::vector<std::thread> th;
  for (int i = 1; i < n; ++i) th.emplace_back([&, i]() { rc[i] = fn(i); });
  rc[0] = fn(0);
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rc[i]) return sub_fail(g, i, rc[i]);
  return A3D_OK;
}
}  // namespace

extern "C" {

int a3d_group_create(const int* device_ids, int n_dev, a3d_group** out) {
  if (!out) return A3D_ERR_INVALID;
  *out = nullptr;
  if (!device_ids || n_dev < 1) {
    g_group_create_error = "device list missing";
    return A3D_ERR_INVALID;
  }
  a3d_group* g = new a3d_group();
  g->dev.resize(n_dev);
  for (int i = 0; i < n_dev; ++i) {
    g->dev[i].id = device_ids[i];
    const int rc = a3d_ctx_create(device_ids[i], &g->dev[i].ctx);
    if (rc) {
      g_group_create_error = std::string("device ") + std::to_string(device_ids[i]) + ": " + a3d_last_error(nullptr);
      for (int k = 0; k < i; ++k) a3d_ctx_destroy(g->dev[k].ctx);
      delete g;
      return rc;
    }
    g->dev[i].stream = static_cast<cudaStream_t>(a3d_ctx_stream(g->dev[i].ctx));
  }
  if (n_dev > 1) {
    const std::string e = g_nccl.load();
    std::vector<ncclComm_t> comms(n_dev, nullptr);
    ncclResult_t r = 0;
    if (e.empty()) r = g_nccl.CommInitAll(comms.data(), n_dev, device_ids);
    if (!e.empty() || r != 0) {
      g_group_create_error = !e.empty() ? e : std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r);
      for (auto& d : g->dev) a3d_ctx_destroy(d.ctx);
      delete g;
      return A3D_ERR_CUDA;
    }
    for (int i = 0; i < n_dev; ++i) g->dev[i].comm = comms[i];
  }
  *out = g;
  return A3D_OK;
}

void a3d_group_destroy(a3d_group* g) {
  if (!g) return;
  for (auto& d : g->dev) {
    cudaSetDevice(d.id);
    cudaStreamSynchronize(d.stream);
    if (d.comm) g_nccl.CommDestroy(d.comm);
    d.st_bytes.release();
    d.pos.release();
    d.quat.release();
    d.origin.release();
    d.first.release();
    d.rec.release();
    d.all.release();
    d.stage.release();
    a3d_ctx_destroy(d.ctx);
  }
  g->result.release();
  delete g;
}

const char* a3d_group_last_error(const a3d_group* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }
int a3d_group_size(const a3d_group* g) { return g ? static_cast<int>(g->dev.size()) : 0; }
a3d_ctx* a3d_group_ctx(a3d_group* g, int i) {
  return (g && i >= 0 && i < static_cast<int>(g->dev.size())) ? g->dev[i].ctx : nullptr;
}

int a3d_group_map_config(a3d_group* g, float voxel_size, int voxels_per_side) {
  if (!g) return A3D_ERR_INVALID;
  for (size_t i = 0; i < g->dev.size(); ++i) {
    const int rc = a3d_map_config(g->dev[i].ctx, voxel_size, voxels_per_side);
    if (rc) return sub_fail(g, static_cast<int>(i), rc);
  }
  g->voxel_size = voxel_size;
  g->vps = voxels_per_side;
  g->nv = voxels_per_side * voxels_per_side * voxels_per_side;
  return A3D_OK;
}

int a3d_group_map_upload_blocks(a3d_group* g, int n, const int32_t* block_idx, const float* esdf_dist,
                                const uint8_t* esdf_observed, const float* tsdf_weight) {
  if (!g) return A3D_ERR_INVALID;
  if (g->nv == 0) return g->fail(A3D_ERR_STATE, "a3d_group_map_config must be called first");
  if (n < 0 || (n > 0 && (!block_idx || !esdf_dist || !esdf_observed))) return g->fail(A3D_ERR_INVALID, "null block payload");
  if (n == 0) return A3D_OK;
  const int N = static_cast<int>(g->dev.size());
  const size_t nv = static_cast<size_t>(g->nv);
  // chunks bound the staging buffers (the per-context upload is chunked the same way)
  const int chunk = static_cast<int>(std::max<size_t>(1, (size_t(256) << 20) / (nv * 9)));
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int m = std::min(chunk, n - i0);
    const size_t d_bytes = m * nv * 4, w_bytes = tsdf_weight ? m * nv * 4 : 0, o_bytes = m * nv;
    const size_t total = d_bytes + w_bytes + o_bytes;
    for (auto& d : g->dev) {
      GCK(cudaSetDevice(d.id));
      GCK(d.st_bytes.reserve(total));
    }
    // host → GPU 0, once
    Device& d0 = g->dev[0];
    GCK(cudaSetDevice(d0.id));
    GCK(cudaMemcpyAsync(d0.st_bytes.p, esdf_dist + static_cast<size_t>(i0) * nv, d_bytes, cudaMemcpyHostToDevice, d0.stream));
    if (tsdf_weight)
      GCK(cudaMemcpyAsync(d0.st_bytes.p + d_bytes, tsdf_weight + static_cast<size_t>(i0) * nv, w_bytes,
                          cudaMemcpyHostToDevice, d0.stream));
    GCK(cudaMemcpyAsync(d0.st_bytes.p + d_bytes + w_bytes, esdf_observed + static_cast<size_t>(i0) * nv, o_bytes,
                        cudaMemcpyHostToDevice, d0.stream));
    // GPU 0 → every GPU over NVLink
    if (N > 1) {
      NCK(g_nccl.GroupStart());
      for (auto& d : g->dev) {
        const ncclResult_t r = g_nccl.Broadcast(d0.st_bytes.p, d.st_bytes.p, total, kNcclChar, 0, d.comm, d.stream);
        if (r != 0) {
          g_nccl.GroupEnd();
          return g->fail(A3D_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(r));
        }
      }
      NCK(g_nccl.GroupEnd());
    }
    // every GPU applies the delta to its own mirror (stream order: after its broadcast)
    const int32_t* idx = block_idx + 3 * static_cast<size_t>(i0);
    const int rc = for_each_device(g, [&](int i) {
      Device& d = g->dev[i];
      const float* dist_dev = reinterpret_cast<const float*>(d.st_bytes.p);
      const float* w_dev = tsdf_weight ? reinterpret_cast<const float*>(d.st_bytes.p + d_bytes) : nullptr;
      return a3d_map_upload_blocks_dev(d.ctx, m, idx, dist_dev, d.st_bytes.p + d_bytes + w_bytes, w_dev);
    });
    if (rc) return rc;
  }
  return A3D_OK;
}

int a3d_group_map_remove_blocks(a3d_group* g, int n, const int32_t* block_idx) {
  if (!g) return A3D_ERR_INVALID;
  return for_each_device(g, [&](int i) { return a3d_map_remove_blocks(g->dev[i].ctx, n, block_idx); });
}
int a3d_group_map_clear(a3d_group* g) {
  if (!g) return A3D_ERR_INVALID;
  return for_each_device(g, [&](int i) { return a3d_map_clear(g->dev[i].ctx); });
}

int a3d_group_caster_create(a3d_group* g, const a3d_caster_params* p, a3d_group_caster** out) {
  if (!g || !out) return A3D_ERR_INVALID;
  *out = nullptr;
  a3d_group_caster* c = new a3d_group_caster();
  c->g = g;
  c->c.resize(g->dev.size(), nullptr);
  for (size_t i = 0; i < g->dev.size(); ++i) {
    const int rc = a3d_caster_create(g->dev[i].ctx, p, &c->c[i]);
    if (rc) {
      sub_fail(g, static_cast<int>(i), rc);
      for (size_t k = 0; k < i; ++k) a3d_caster_destroy(c->c[k]);
      delete c;
      return rc;
    }
  }
  *out = c;
  return A3D_OK;
}
void a3d_group_caster_destroy(a3d_group_caster* c) {
  if (!c) return;
  for (a3d_caster* x : c->c) a3d_caster_destroy(x);
  delete c;
}
int a3d_group_caster_get_info(const a3d_group_caster* c, a3d_caster_info* out) {
  return (c && !c->c.empty()) ? a3d_caster_get_info(c->c[0], out) : A3D_ERR_INVALID;
}
int a3d_group_eval_create(a3d_group* g, const a3d_eval_params* p, a3d_group_eval** out) {
  if (!g || !out) return A3D_ERR_INVALID;
  *out = nullptr;
  a3d_group_eval* e = new a3d_group_eval();
  e->g = g;
  e->e.resize(g->dev.size(), nullptr);
  for (size_t i = 0; i < g->dev.size(); ++i) {
    const int rc = a3d_eval_create(g->dev[i].ctx, p, &e->e[i]);
    if (rc) {
      sub_fail(g, static_cast<int>(i), rc);
      for (size_t k = 0; k < i; ++k) a3d_eval_destroy(e->e[k]);
      delete e;
      return rc;
    }
  }
  *out = e;
  return A3D_OK;
}
void a3d_group_eval_destroy(a3d_group_eval* e) {
  if (!e) return;
  for (a3d_eval* x : e->e) a3d_eval_destroy(x);
  delete e;
}

int a3d_group_set_option(a3d_group* g, const char* name, int64_t value) {
  if (!g || !name) return A3D_ERR_INVALID;
  if (std::strcmp(name, "deal_block") == 0) {
    if (value < 1 || value > (1 << 20)) return g->fail(A3D_ERR_INVALID, "deal_block must be in [1, 2^20]");
    g->block = static_cast<int>(value);
    return A3D_OK;
  }
  for (size_t i = 0; i < g->dev.size(); ++i) {
    const int rc = a3d_set_option(g->dev[i].ctx, name, value);
    if (rc) return sub_fail(g, static_cast<int>(i), rc);
  }
  return A3D_OK;
}

float a3d_group_last_kernel_ms(a3d_group* g) {
  float ms = -1.0f;
  if (g)
    for (auto& d : g->dev) ms = std::max(ms, a3d_last_kernel_ms(d.ctx));
  return ms;
}
uint64_t a3d_group_kernel_launches(a3d_group* g) {
  uint64_t n = 0;
  if (g)
    for (auto& d : g->dev) n += a3d_kernel_launches(d.ctx);
  return n;
}

int a3d_group_compute_gains(a3d_group* g, const a3d_group_caster* caster, const a3d_group_eval* eval, int n_views,
                            const double* pos, const double* quat, const int32_t* view_segment, int n_segments,
                            double* gains_out, uint64_t* n_visible_out, uint64_t* n_samples_out,
                            uint64_t* digests_out, uint64_t* set_digests_out, const double* seg_origin) {
  if (!g) return A3D_ERR_INVALID;
  if (!caster || !eval || caster->g != g || eval->g != g) return g->fail(A3D_ERR_INVALID, "caster/eval missing or from another group");
  if (n_views < 0 || n_segments < 0) return g->fail(A3D_ERR_INVALID, "negative count");
  if (n_segments == 0) return A3D_OK;
  if ((n_views > 0 && (!pos || !quat || !view_segment)) || !gains_out) return g->fail(A3D_ERR_INVALID, "null argument");
  const int N = static_cast<int>(g->dev.size());
  // views of each segment
  std::vector<int32_t> first(static_cast<size_t>(n_segments) + 1, 0);
  for (int v = 0; v < n_views; ++v) {
    const int s = view_segment[v];
    if (s < 0 || s >= n_segments) return g->fail(A3D_ERR_INVALID, "view_segment out of range");
    if (v > 0 && s < view_segment[v - 1]) return g->fail(A3D_ERR_INVALID, "view_segment must be non-decreasing");
    first[s + 1]++;
  }
  for (int s = 0; s < n_segments; ++s) first[s + 1] += first[s];
  // deal blocks of segments round-robin: neighbouring segments (similar cost) spread over the GPUs
  const int blk = std::max(1, std::min(g->block, (n_segments + N - 1) / N));
  std::vector<std::vector<int>> mine(N);
  for (int s0 = 0, k = 0; s0 < n_segments; s0 += blk, ++k)
    for (int s = s0; s < std::min(n_segments, s0 + blk); ++s) mine[k % N].push_back(s);
  size_t per = 0;
  for (int i = 0; i < N; ++i) per = std::max(per, mine[i].size());
  const bool want_dig = digests_out != nullptr, want_sdig = set_digests_out != nullptr;

  // Every GPU's share is laid out in page-locked memory first (whole blocks of segments at a time:
  // their views are contiguous in the caller's arrays), then all GPUs are started back to back with
  // asynchronous copies — no GPU waits for the host to prepare another one's share.
  std::vector<int> n_views_of(N, 0);
  for (int i = 0; i < N; ++i) {
    Device& d = g->dev[i];
    const int ns = static_cast<int>(mine[i].size());
    int nvw = 0;
    for (int s : mine[i]) nvw += first[s + 1] - first[s];
    n_views_of[i] = nvw;
    const size_t b_pos = 24 * static_cast<size_t>(nvw), b_quat = 32 * static_cast<size_t>(nvw),
                 b_org = seg_origin ? 24 * static_cast<size_t>(ns) : 0, b_first = 4 * (static_cast<size_t>(ns) + 1);
    GCK(cudaSetDevice(d.id));
    GCK(d.stream ? cudaStreamSynchronize(d.stream) : cudaSuccess);  // (the staging of the previous call is free)
    GCK(d.stage.reserve(b_pos + b_quat + b_org + b_first));
    double* hp = reinterpret_cast<double*>(d.stage.p);
    double* hq = reinterpret_cast<double*>(d.stage.p + b_pos);
    double* ho = reinterpret_cast<double*>(d.stage.p + b_pos + b_quat);
    int32_t* hf = reinterpret_cast<int32_t*>(d.stage.p + b_pos + b_quat + b_org);
    hf[0] = 0;
    size_t vw = 0;
    for (int k = 0; k < ns;) {
      // a run of consecutive segments = one run of consecutive views
      int k1 = k + 1;
      while (k1 < ns && mine[i][k1] == mine[i][k1 - 1] + 1) ++k1;
      const int s0 = mine[i][k], s1 = mine[i][k1 - 1] + 1;
      const size_t nv_run = static_cast<size_t>(first[s1] - first[s0]);
      std::memcpy(hp + 3 * vw, pos + 3 * static_cast<size_t>(first[s0]), 24 * nv_run);
      std::memcpy(hq + 4 * vw, quat + 4 * static_cast<size_t>(first[s0]), 32 * nv_run);
      if (seg_origin) std::memcpy(ho + 3 * static_cast<size_t>(k), seg_origin + 3 * static_cast<size_t>(s0), 24 * static_cast<size_t>(s1 - s0));
      for (int s = s0; s < s1; ++s, ++k) hf[k + 1] = hf[k] + (first[s + 1] - first[s]);
      vw += nv_run;
    }
  }
  // one host thread per GPU issues that GPU's copies and launches (a cast is ~30 runtime calls; issued
  // from one thread for 8 GPUs they cost more than the 512-view cast itself)
  std::vector<int> rcs(N, A3D_OK);
  std::vector<std::string> errs(N);
  auto issue = [&](int i) {
    Device& d = g->dev[i];
    const int ns = static_cast<int>(mine[i].size()), nvw = n_views_of[i];
    const size_t b_pos = 24 * static_cast<size_t>(nvw), b_quat = 32 * static_cast<size_t>(nvw),
                 b_org = seg_origin ? 24 * static_cast<size_t>(ns) : 0, b_first = 4 * (static_cast<size_t>(ns) + 1);
    auto ck = [&](cudaError_t e, const char* what) {
      if (e != cudaSuccess && rcs[i] == A3D_OK) {
        rcs[i] = A3D_ERR_CUDA;
        errs[i] = std::string(what) + ": " + cudaGetErrorString(e);
      }
      return e == cudaSuccess;
    };
    if (!ck(cudaSetDevice(d.id), "cudaSetDevice")) return;
    if (!ck(d.pos.reserve(3 * static_cast<size_t>(std::max(nvw, 1))), "cudaMalloc") ||
        !ck(d.quat.reserve(4 * static_cast<size_t>(std::max(nvw, 1))), "cudaMalloc") ||
        !ck(d.first.reserve(static_cast<size_t>(ns) + 1), "cudaMalloc") ||
        !ck(d.rec.reserve(per * kRecWords), "cudaMalloc") || !ck(d.all.reserve(per * kRecWords * N), "cudaMalloc"))
      return;
    ck(cudaMemsetAsync(d.rec.p, 0, per * kRecWords * 8, d.stream), "cudaMemsetAsync");
    if (nvw > 0) {
      ck(cudaMemcpyAsync(d.pos.p, d.stage.p, b_pos, cudaMemcpyHostToDevice, d.stream), "cudaMemcpyAsync");
      ck(cudaMemcpyAsync(d.quat.p, d.stage.p + b_pos, b_quat, cudaMemcpyHostToDevice, d.stream), "cudaMemcpyAsync");
    }
    ck(cudaMemcpyAsync(d.first.p, d.stage.p + b_pos + b_quat + b_org, b_first, cudaMemcpyHostToDevice, d.stream),
       "cudaMemcpyAsync");
    const double* org_dev = nullptr;
    if (seg_origin && ns > 0) {
      if (!ck(d.origin.reserve(3 * static_cast<size_t>(ns)), "cudaMalloc")) return;
      ck(cudaMemcpyAsync(d.origin.p, d.stage.p + b_pos + b_quat, b_org, cudaMemcpyHostToDevice, d.stream),
         "cudaMemcpyAsync");
      org_dev = d.origin.p;
    }
    if (ns > 0 && rcs[i] == A3D_OK) {
      // record layout on the device: five arrays of `per` words (struct of arrays)
      unsigned long long* r = d.rec.p;
      const int rc = a3d_compute_gains_dev(d.ctx, caster->c[i], eval->e[i], nvw, d.pos.p, d.quat.p, d.first.p, ns,
                                           reinterpret_cast<double*>(r), reinterpret_cast<uint64_t*>(r + per),
                                           reinterpret_cast<uint64_t*>(r + 2 * per),
                                           want_dig ? reinterpret_cast<uint64_t*>(r + 3 * per) : nullptr,
                                           want_sdig ? reinterpret_cast<uint64_t*>(r + 4 * per) : nullptr, org_dev);
      if (rc) {
        rcs[i] = rc;
        errs[i] = std::string("GPU ") + std::to_string(d.id) + ": " + a3d_last_error(d.ctx);
      }
    }
  };
  {
    std::vector<std::thread> th;
    for (int i = 1; i < N; ++i) th.emplace_back(issue, i);
    issue(0);
    for (auto& t : th) t.join();
  }
  for (int i = 0; i < N; ++i)
    if (rcs[i] != A3D_OK) return g->fail(rcs[i], errs[i]);
  const unsigned long long* gathered = g->dev[0].rec.p;
  if (N > 1) {
    NCK(g_nccl.GroupStart());
    for (auto& d : g->dev) {
      const ncclResult_t r = g_nccl.AllGather(d.rec.p, d.all.p, per * kRecWords, kNcclUint64, d.comm, d.stream);
      if (r != 0) {
        g_nccl.GroupEnd();
        return g->fail(A3D_ERR_CUDA, std::string("ncclAllGather: ") + g_nccl.GetErrorString(r));
      }
    }
    NCK(g_nccl.GroupEnd());
    gathered = g->dev[0].all.p;
  }
  const size_t n_words = per * kRecWords * static_cast<size_t>(N);
  GCK(cudaSetDevice(g->dev[0].id));
  GCK(g->result.reserve(n_words * 8));
  const unsigned long long* const host = reinterpret_cast<const unsigned long long*>(g->result.p);
  GCK(cudaMemcpyAsync(g->result.p, gathered, n_words * 8, cudaMemcpyDeviceToHost, g->dev[0].stream));
  for (auto& d : g->dev) {
    GCK(cudaSetDevice(d.id));
    GCK(cudaStreamSynchronize(d.stream));
  }
  for (int i = 0; i < N; ++i) {
    const unsigned long long* r = host + static_cast<size_t>(i) * per * kRecWords;
    for (size_t k = 0; k < mine[i].size(); ++k) {
      const int s = mine[i][k];
      std::memcpy(&gains_out[s], &r[k], 8);
      if (n_visible_out) n_visible_out[s] = r[per + k];
      if (n_samples_out) n_samples_out[s] = r[2 * per + k];
      if (digests_out) digests_out[s] = r[3 * per + k];
      if (set_digests_out) set_digests_out[s] = r[4 * per + k];
    }
  }
  return A3D_OK;
}

}  // extern "C"
