// group.cu — row-band groups: the frame strong-scaled over the GPUs of one node (BASELINE config 4).
//
// Every GPU holds the whole scene and renders the rows [row0, row1) of the image. ReSTIR couples rows in two places, and
// both are served by exchanging reservoir rows between band neighbours instead of re-rendering halo rows:
//   kind 0 (mid frame)  the TEMPORAL reservoirs (buffer half 1) of the `radius` rows next to a band edge, after the temporal
//                       stage: spatial reuse reads them up to `radius` rows away (RayGen.rgen:470-506, neighbour radius 30)
//   kind 1 (end frame)  the FINAL reservoirs (half 0) of the `margin` rows next to a band edge, after the spatial stage: the
//                       next frame's temporal reuse reads them one motion vector away (RayGen.rgen:424-461)
// Only the G-buffer keeps a halo (max(radius, margin) + 1 rows of primary rays: neighbour surfaces and previous-frame
// normals / depths). Three 16-byte SoA planes per reservoir half, contiguous per row range.
// Back ends: NCCL send / recv on the frame's stream (one process per GPU; libnccl.so.2 is resolved at run time, the 128-byte
// unique id travels through the host's own plumbing), or direct device-to-device copies between the contexts of ONE process
// (several GPUs, or one GPU in the tests), each context driven by its own host thread.
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <vector>

#include "kernels.h"

namespace snk {

namespace {
// the part of NCCL's C API the exchange needs (stable ABI since NCCL 2.7: point-to-point calls, 128-byte unique id)
struct NcclUniqueId { char internal[128]; };
using ncclComm_t = void*;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi* nccl_api(std::string& err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("SNKPT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
      api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
      api.Send = (decltype(api.Send))dlsym(api.lib, "ncclSend");
      api.Recv = (decltype(api.Recv))dlsym(api.lib, "ncclRecv");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd) {
    err = "row-band group: libnccl.so.2 not found (set SNKPT_NCCL_LIB)";
    return nullptr;
  }
  return &api;
}
constexpr int NCCL_CHAR = 0;  // ncclInt8 / ncclChar
}  // namespace

struct GroupState {
  int rank = 0, world = 1;
  uint32_t radius = 30, margin = 16;
  ncclComm_t comm = nullptr;            // NCCL back end
  std::vector<SnkPt*> peers;            // local back end: every context of the group by rank
  cudaEvent_t ev[2][2] = {};            // [kind][parity]: this band's rows of that kind are written
  std::atomic<uint64_t> posted[2];      // records issued per kind
  uint64_t seq[2] = {0, 0};
  uint64_t bytes_sent = 0;
  std::vector<uint32_t> starts;         // world + 1 row boundaries (equal bands, or the table of snkpt_group_set_bands)
};

static void equal_band(uint32_t H, int world, int r, uint32_t& b0, uint32_t& b1) {
  const uint32_t q = H / (uint32_t)world, rem = H % (uint32_t)world;
  b0 = (uint32_t)r * q + ((uint32_t)r < rem ? (uint32_t)r : rem);
  b1 = b0 + q + ((uint32_t)r < rem ? 1u : 0u);
}
static void band_of(const GroupState* g, int r, uint32_t& b0, uint32_t& b1) { b0 = g->starts[r]; b1 = g->starts[r + 1]; }
static void window_of(uint32_t H, uint32_t b0, uint32_t b1, uint32_t n, uint32_t& w0, uint32_t& w1) {
  w0 = b0 > n ? b0 - n : 0u;
  w1 = b1 + n < H ? b1 + n : H;
}

uint32_t group_gbuffer_halo(const SnkPt* ctx) {
  const GroupState* g = ctx->group;
  return (g->radius > g->margin ? g->radius : g->margin) + 1u;
}

// kind 0: temporal reservoirs (half 1), `radius` rows; kind 1: final reservoirs (half 0), `margin` rows
int group_exchange(SnkPt* ctx, int kind, cudaStream_t st) {
  GroupState* g = ctx->group;
  if (!g || g->world <= 1) return 0;
  const uint32_t H = ctx->H, W = ctx->W, n = kind == 0 ? g->radius : g->margin;
  const int half = kind == 0 ? 1 : 0;
  const size_t row_bytes = (size_t)W * sizeof(float4);
  uint32_t b0, b1, w0, w1;
  band_of(g, g->rank, b0, b1);
  window_of(H, b0, b1, n, w0, w1);
  struct Xfer { int peer; uint32_t s0, s1, r0, r1; };
  std::vector<Xfer> plan;
  for (int p = 0; p < g->world; ++p) {
    if (p == g->rank) continue;
    uint32_t p0, p1, pw0, pw1;
    band_of(g, p, p0, p1);
    window_of(H, p0, p1, n, pw0, pw1);
    Xfer x{p, b0 > pw0 ? b0 : pw0, b1 < pw1 ? b1 : pw1, p0 > w0 ? p0 : w0, p1 < w1 ? p1 : w1};  // my rows the peer needs | its rows I need
    if (x.s0 < x.s1 || x.r0 < x.r1) plan.push_back(x);
  }
  if (plan.empty()) return 0;
  if (g->comm) {
    std::string err;
    NcclApi* api = nccl_api(err);
    if (!api) { ctx->err = err; return SNKPT_E_STATE; }
    int rc = api->GroupStart();
    for (const Xfer& x : plan)
      for (int k = 0; k < 3 && rc == 0; ++k) {
        char* base = reinterpret_cast<char*>(ctx->planes.res[half][k]);
        if (x.s0 < x.s1) { rc = api->Send(base + x.s0 * row_bytes, (x.s1 - x.s0) * row_bytes, NCCL_CHAR, x.peer, g->comm, st); g->bytes_sent += (x.s1 - x.s0) * row_bytes; }
        if (rc == 0 && x.r0 < x.r1) rc = api->Recv(base + x.r0 * row_bytes, (x.r1 - x.r0) * row_bytes, NCCL_CHAR, x.peer, g->comm, st);
      }
    const int rc2 = api->GroupEnd();
    if (rc || rc2) { ctx->err = std::string("row-band group: NCCL ") + (api->GetErrorString ? api->GetErrorString(rc ? rc : rc2) : "error"); return SNKPT_E_CUDA; }
    return 1;
  }
  // local back end: publish my rows, then pull the peers' rows once they published theirs (lock step: one exchange of each
  // kind per frame on every band; the read-after-write order is the event, the write-after-read order follows from the next
  // exchange of the other kind, which waits for this band's later stage)
  const uint64_t k = g->seq[kind]++;
  cudaEventRecord(g->ev[kind][k & 1], st);
  g->posted[kind].store(k + 1, std::memory_order_release);
  for (const Xfer& x : plan) {
    if (x.r0 >= x.r1) continue;
    SnkPt* peer = g->peers[x.peer];
    GroupState* pg = peer->group;
    const auto t0 = std::chrono::steady_clock::now();
    while (pg->posted[kind].load(std::memory_order_acquire) < k + 1) {
      std::this_thread::yield();
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) { ctx->err = "row-band group: a band of the local group did not reach the exchange (every band needs its own host thread)"; return SNKPT_E_STATE; }
    }
    cudaStreamWaitEvent(st, pg->ev[kind][k & 1], 0);
    for (int pl = 0; pl < 3; ++pl) {
      char* dst = reinterpret_cast<char*>(ctx->planes.res[half][pl]) + x.r0 * row_bytes;
      const char* src = reinterpret_cast<const char*>(peer->planes.res[half][pl]) + x.r0 * row_bytes;
      if (peer->device == ctx->device) cudaMemcpyAsync(dst, src, (x.r1 - x.r0) * row_bytes, cudaMemcpyDeviceToDevice, st);
      else cudaMemcpyPeerAsync(dst, ctx->device, src, peer->device, (x.r1 - x.r0) * row_bytes, st);
    }
  }
  return 1;
}

static int attach_common(SnkPt* ctx, int rank, int world, uint32_t radius_rows, uint32_t margin_rows, GroupState** out) {
  if (!ctx || world < 1 || rank < 0 || rank >= world) return SNKPT_E_INVALID;
  if (ctx->group) { ctx->err = "row-band group: the context already belongs to a group"; return SNKPT_E_STATE; }
  std::vector<uint32_t> starts;
  if (!ctx->band_starts.empty()) {
    if ((int)ctx->band_starts.size() != world + 1) { ctx->err = "row-band group: the band table of snkpt_group_set_bands is for another group size"; return SNKPT_E_INVALID; }
    starts = ctx->band_starts;
  } else {
    starts.resize((size_t)world + 1);
    for (int r = 0; r < world; ++r) { uint32_t a, b; equal_band(ctx->H, world, r, a, b); starts[r] = a; starts[r + 1] = b; }
  }
  if (ctx->row0 != starts[rank] || ctx->row1 != starts[rank + 1]) {
    ctx->err = "row-band group: the context's rows are not band `rank` of the group (snkpt_group_band, or the table of snkpt_group_set_bands)";
    return SNKPT_E_INVALID;
  }
  GroupState* g = new GroupState();
  g->starts = starts;
  g->rank = rank; g->world = world; g->radius = radius_rows; g->margin = margin_rows;
  g->posted[0].store(0); g->posted[1].store(0);
  for (auto& a : g->ev) for (auto& e : a) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  ctx->group = g;
  ctx->halo_pass1 = 0;
  ctx->halo_gbuffer = group_gbuffer_halo(ctx);
  ctx->history_valid = false;
  *out = g;
  return SNKPT_OK;
}

void group_free(SnkPt* ctx) {
  GroupState* g = ctx->group;
  if (!g) return;
  if (g->comm) { std::string err; if (NcclApi* api = nccl_api(err)) if (api->CommDestroy) api->CommDestroy(g->comm); }
  for (auto& a : g->ev) for (auto& e : a) if (e) cudaEventDestroy(e);
  delete g;
  ctx->group = nullptr;
}

}  // namespace snk

using namespace snk;

extern "C" {

SNKPT_API void snkpt_group_band(uint32_t height, int world, int rank, uint32_t* row_begin, uint32_t* row_end) {
  uint32_t b0 = 0, b1 = 0;
  if (world >= 1 && rank >= 0 && rank < world) equal_band(height, world, rank, b0, b1);
  if (row_begin) *row_begin = b0;
  if (row_end) *row_end = b1;
}

SNKPT_API int snkpt_group_set_bands(SnkPt* ctx, int world, const uint32_t* row_starts) {
  if (!ctx) return SNKPT_E_INVALID;
  if (ctx->group) { ctx->err = "snkpt_group_set_bands: the context already belongs to a group (detach first)"; return SNKPT_E_STATE; }
  if (world == 0 || !row_starts) { ctx->band_starts.clear(); return SNKPT_OK; }  // back to equal bands
  if (world < 1) return SNKPT_E_INVALID;
  if (row_starts[0] != 0u || row_starts[world] != ctx->H) { ctx->err = "snkpt_group_set_bands: the table must start at row 0 and end at the image height"; return SNKPT_E_INVALID; }
  for (int r = 0; r < world; ++r)
    if (row_starts[r] >= row_starts[r + 1]) { ctx->err = "snkpt_group_set_bands: every band needs at least one row"; return SNKPT_E_INVALID; }
  ctx->band_starts.assign(row_starts, row_starts + world + 1);
  return SNKPT_OK;
}

SNKPT_API int snkpt_group_unique_id(void* out128) {
  if (!out128) return SNKPT_E_INVALID;
  std::string err;
  NcclApi* api = nccl_api(err);
  if (!api) return SNKPT_E_STATE;
  NcclUniqueId id;
  if (api->GetUniqueId(&id) != 0) return SNKPT_E_CUDA;
  memcpy(out128, id.internal, 128);
  return SNKPT_OK;
}

SNKPT_API int snkpt_group_attach_nccl(SnkPt* ctx, int rank, int world, const void* id128, uint32_t radius_rows, uint32_t margin_rows) {
  if (!ctx || !id128) return SNKPT_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return SNKPT_E_CUDA;
  std::string err;
  NcclApi* api = nccl_api(err);
  if (!api) { ctx->err = err; return SNKPT_E_STATE; }
  GroupState* g = nullptr;
  const int rc = attach_common(ctx, rank, world, radius_rows, margin_rows, &g);
  if (rc) return rc;
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  const int nrc = api->CommInitRank(&g->comm, world, id, rank);
  if (nrc != 0) { ctx->err = std::string("row-band group: ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(nrc) : "error"); g->comm = nullptr; group_free(ctx); return SNKPT_E_CUDA; }
  return SNKPT_OK;
}

SNKPT_API int snkpt_group_attach_local(SnkPt** ctxs, int n, uint32_t radius_rows, uint32_t margin_rows) {
  if (!ctxs || n < 1) return SNKPT_E_INVALID;
  for (int r = 0; r < n; ++r) if (!ctxs[r]) return SNKPT_E_INVALID;
  for (int r = 0; r < n; ++r) {
    GroupState* g = nullptr;
    cudaSetDevice(ctxs[r]->device);
    const int rc = attach_common(ctxs[r], r, n, radius_rows, margin_rows, &g);
    if (rc) { for (int k = 0; k < r; ++k) group_free(ctxs[k]); return rc; }
  }
  for (int r = 0; r < n; ++r) {
    ctxs[r]->group->peers.assign(ctxs, ctxs + n);
    for (int p = 0; p < n; ++p)
      if (ctxs[p]->device != ctxs[r]->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[p]->device);
        if (can) { cudaSetDevice(ctxs[r]->device); cudaDeviceEnablePeerAccess(ctxs[p]->device, 0); cudaGetLastError(); }
      }
  }
  return SNKPT_OK;
}

SNKPT_API int snkpt_group_detach(SnkPt* ctx) {
  if (!ctx) return SNKPT_E_INVALID;
  if (!ctx->group) return SNKPT_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  group_free(ctx);
  ctx->halo_pass1 = 0; ctx->halo_gbuffer = 0;
  ctx->history_valid = false;
  return SNKPT_OK;
}

SNKPT_API int snkpt_group_info(SnkPt* ctx, int* rank, int* world, uint64_t* bytes_sent) {
  if (!ctx || !ctx->group) return SNKPT_E_STATE;
  if (rank) *rank = ctx->group->rank;
  if (world) *world = ctx->group->world;
  if (bytes_sent) *bytes_sent = ctx->group->bytes_sent;
  return SNKPT_OK;
}

}  // extern "C"
