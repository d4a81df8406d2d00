// trace.cu — wavefront traversal kernels over ray queues in HBM.
//
// The generic ray query of the boundary (snkpt_trace_rays) and the secondary-ray
// stages of the frame go through these kernels: a persistent grid (a multiple of
// the 148 SMs) pulls rays from a 32-B/ray queue, traverses the two-level BVH8 and
// writes 16-B hit records + a 4-B instance plane. Replaces the driver side of
// vkCmdTraceRaysKHR (Core/src/rendering/Raytracing.cpp:253, :260).
#include "kernels.h"

#include <cstdlib>

namespace snk {

namespace {

constexpr int TRACE_THREADS = 128;

// ---- persistent wavefront kernel with dynamic ray fetch ----------------------------------------------------
// One warp keeps 32 traversal state machines; lanes whose ray finished pull the next ray index from a global
// ticket (one atomic per warp and refill) instead of idling until the slowest lane of the warp is done
// (the first, static one-ray-per-thread kernel ran at 7.4 of 32 active lanes on the bounce queues,
// profiles/r01a_prof_trace.txt).
// A refill is triggered once the warp has accumulated PT_LOST_MAX idle lane-iterations beyond PT_IDLE_OK per
// iteration (Aila & Laine 2009 / Ylitie et al. 2017 dynamic fetch). Results are written per ray index, so the
// output is independent of the fetch order.
#ifndef SNK_PT_IDLE_OK
#define SNK_PT_IDLE_OK 4
#endif
#ifndef SNK_PT_LOST_MAX
#define SNK_PT_LOST_MAX 16
#endif
#ifndef SNK_PT_CTAS
#define SNK_PT_CTAS 8
#endif
constexpr int PT_IDLE_OK = SNK_PT_IDLE_OK;
constexpr int PT_LOST_MAX = SNK_PT_LOST_MAX;
constexpr int PT_CTAS_PER_SM = SNK_PT_CTAS;

template <bool ANY>
__global__ void __launch_bounds__(TRACE_THREADS, PT_CTAS_PER_SM)
k_trace_pt(TraceScene S, const float4* __restrict__ rays, unsigned long long n, unsigned long long* __restrict__ ticket,
           SnkHit* __restrict__ hits_aos, float4* __restrict__ hits_soa, uint32_t* __restrict__ inst_soa, float* __restrict__ vis_out) {
  __shared__ uint2 sm_stack[SHORT_STACK * TRACE_THREADS];
  const uint32_t lane = threadIdx.x & 31u;
  // world-space reciprocal direction / origin / direction and the closest hit's u, v, ids: the thread's cold state (traverse.cuh, COLD)
  __shared__ float4 sm_world[4 * TRACE_THREADS];
  StackRef stk{sm_stack + threadIdx.x,
               S.spill + ((size_t)blockIdx.x * (TRACE_THREADS / 32) + (threadIdx.x >> 5)) * (SPILL_STACK * 32) + lane};
  stk.world = sm_world + threadIdx.x;
  Trav t;
  bool active = false, exhausted = false;
  unsigned long long idx = 0;
  auto write_result = [&](unsigned long long i, const HitRec& h) {
    if (hits_aos) { SnkHit o; o.t = h.t; o.u = h.u; o.v = h.v; o.prim = h.prim; o.inst = h.inst; hits_aos[i] = o; }
    if (hits_soa) hits_soa[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
    if (inst_soa) inst_soa[i] = h.inst;
    if (vis_out) vis_out[i] = h.t >= 0.0f ? 0.0f : 1.0f;  // Shadow.rmiss:12 → 1 - float(hit)
  };
  while (true) {
    // ---- refill: every lane without a ray takes the next index; dead slots (tmax <= tmin) resolve at once
    for (int round = 0; round < 4; ++round) {
      const uint32_t need = __ballot_sync(0xFFFFFFFFu, !active);
      if (need == 0u || exhausted) break;
      const uint32_t cnt = __popc(need);
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(ticket, (unsigned long long)cnt);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (!active) {
        const unsigned long long i = base + __popc(need & ((1u << lane) - 1u));
        if (i < n) {
          const float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
          trav_begin(t, a, b);
          if (b.w > a.w && S.tlas_node_count) {
            active = true; idx = i;
            sm_world[threadIdx.x] = make_float4(t.r.idx, t.r.idy, t.r.idz, __uint_as_float(t.r.octinv));
            sm_world[TRACE_THREADS + threadIdx.x] = make_float4(a.x, a.y, a.z, 0.f);
            sm_world[2 * TRACE_THREADS + threadIdx.x] = make_float4(b.x, b.y, b.z, 0.f);
          }
          else write_result(i, t.best);
        }
      }
      if (base + cnt >= n) exhausted = true;
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) {
      if (exhausted) break;
      continue;
    }
    // ---- traverse; the three phases are warp-synchronous (the ballots reconverge all 32 lanes per iteration)
    int lost = 0;
    while (true) {
      if (trav_step_impl<TRACE_THREADS, true, true>(S, t, stk, active, 0xFFFFFFFFu, ANY)) {
        HitRec h;
        h.t = t.best.t; h.u = 0.f; h.v = 0.f; h.prim = 0xFFFFFFFFu; h.inst = 0xFFFFFFFFu;
        if (h.t >= 0.0f) {
          const float4 b = sm_world[3 * TRACE_THREADS + threadIdx.x];
          h.u = b.x; h.v = b.y; h.prim = __float_as_uint(b.z); h.inst = __float_as_uint(b.w);
        }
        write_result(idx, h);
        active = false;
      }
      const uint32_t live = __ballot_sync(0xFFFFFFFFu, active);
      if (live == 0u) break;
      if (!exhausted) {
        lost += (32 - __popc(live)) - PT_IDLE_OK;
        if (lost < 0) lost = 0;
        if (lost >= PT_LOST_MAX) break;
      }
    }
  }
}

// Diagnostic (not on the hot path): traversal-order independent work of one ray = every wide node / leaf
// record / instance whose (quantised, padded) box overlaps the segment [tmin, t_end], with
// t_end = the reported hit distance, or tmax for a miss. Feeds the roofline's algorithmic bytes.
__device__ void count_subtree(const RaySpace& r0, const uint4* __restrict__ nodes, uint32_t root, float tmin, float tend,
                              uint32_t& n_nodes, uint32_t& leaf_bits_total, uint32_t* leaf_out, int& leaf_n, int leaf_cap) {
  RaySpace r = r0;
  r.octinv = 0;  // inner hit bits become 24 + slot
  uint32_t st[64];
  int sp = 0;
  st[sp++] = root;
  while (sp) {
    const uint32_t ni = st[--sp];
    uint4 n0, n1, n2, n3, n4;
    load_node(nodes, ni, n0, n1, n2, n3, n4);
    ++n_nodes;
    const uint32_t hm = intersect_node(r, n0, n1, n2, n3, n4, tmin, tend);
    const uint32_t imask = n0.w >> 24;
    for (uint32_t s = 0; s < 8; ++s)
      if (hm & (1u << (24 + s))) { if (sp < 64) st[sp++] = n1.x + __popc(imask & ((1u << s) - 1u)); }
    uint32_t tb = hm & 0x00FFFFFFu;
    leaf_bits_total += __popc(tb);
    while (leaf_out && tb) {
      const uint32_t bit = __ffs(tb) - 1;
      tb &= tb - 1;
      if (leaf_n < leaf_cap) leaf_out[leaf_n++] = n1.y + bit;
    }
  }
}

__global__ void __launch_bounds__(TRACE_THREADS) k_count_work(TraceScene S, const float4* __restrict__ rays,
                                                              const SnkHit* __restrict__ hits, unsigned long long n,
                                                              uint32_t* __restrict__ out3) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
  const float tend = hits[i].t >= 0.f ? hits[i].t : b.w;
  uint32_t nodes = 0, recs = 0, insts = 0, entered = 0;
  if (S.tlas_node_count) {
    RaySpace r;
    setup_ray_space(r, a.x, a.y, a.z, b.x, b.y, b.z);
    // pass 1: TLAS, collecting the leaf records (instances) the segment overlaps, 64 at a time
    uint32_t leaf[64];
    int ln = 0;
    uint32_t dummy = 0;
    count_subtree(r, S.tlas_nodes, 0u, a.w, tend, nodes, dummy, leaf, ln, 64);
    insts = dummy;
    (void)insts;
    for (int k = 0; k < ln; ++k) {
      const uint32_t slot = S.tlas_leaf_ids[leaf[k]];
      if (S.inst_sphere && !ray_may_hit_sphere(S.inst_sphere[slot], a.x, a.y, a.z, b.x, b.y, b.z, tend)) continue;  // culled before the entry (traverse.cuh)
      ++entered;
      const float4* M = reinterpret_cast<const float4*>(S.inst + slot);
      const float4 r0 = M[0], r1 = M[1], r2 = M[2];
      RaySpace o;
      setup_ray_space(o, r0.x * a.x + r0.y * a.y + r0.z * a.z + r0.w, r1.x * a.x + r1.y * a.y + r1.z * a.z + r1.w,
                      r2.x * a.x + r2.y * a.y + r2.z * a.z + r2.w, r0.x * b.x + r0.y * b.y + r0.z * b.z,
                      r1.x * b.x + r1.y * b.y + r1.z * b.z, r2.x * b.x + r2.y * b.y + r2.z * b.z);
      int unused = 0;
      count_subtree(o, S.blas_nodes, S.inst[slot].blas_root, a.w, tend, nodes, recs, nullptr, unused, 0);
    }
  }
  out3[3 * i] = nodes; out3[3 * i + 1] = recs; out3[3 * i + 2] = entered;  // instances entered after the bounding-sphere cull
}

// whole-queue sums of the same quantities (profiling level 2)
__global__ void __launch_bounds__(TRACE_THREADS) k_count_work_sum(TraceScene S, const float4* __restrict__ rays,
                                                                  const float4* __restrict__ hits, unsigned long long n,
                                                                  unsigned long long* __restrict__ sums4) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v[4] = {0, 0, 0, 0};
  if (i < n) {
    const float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
    if (b.w >= a.w && S.tlas_node_count) {
      const float ht = hits[i].x;
      const float tend = ht >= 0.f ? ht : b.w;
      uint32_t nodes = 0, recs = 0, insts = 0, entered = 0;
      RaySpace r;
      setup_ray_space(r, a.x, a.y, a.z, b.x, b.y, b.z);
      uint32_t leaf[64];
      int ln = 0;
      count_subtree(r, S.tlas_nodes, 0u, a.w, tend, nodes, insts, leaf, ln, 64);
      for (int k = 0; k < ln; ++k) {
        const uint32_t slot = S.tlas_leaf_ids[leaf[k]];
        if (S.inst_sphere && !ray_may_hit_sphere(S.inst_sphere[slot], a.x, a.y, a.z, b.x, b.y, b.z, tend)) continue;  // culled before the entry (traverse.cuh)
        ++entered;
        const float4* M = reinterpret_cast<const float4*>(S.inst + slot);
        const float4 r0 = M[0], r1 = M[1], r2 = M[2];
        RaySpace o;
        setup_ray_space(o, r0.x * a.x + r0.y * a.y + r0.z * a.z + r0.w, r1.x * a.x + r1.y * a.y + r1.z * a.z + r1.w,
                        r2.x * a.x + r2.y * a.y + r2.z * a.z + r2.w, r0.x * b.x + r0.y * b.y + r0.z * b.z,
                        r1.x * b.x + r1.y * b.y + r1.z * b.z, r2.x * b.x + r2.y * b.y + r2.z * b.z);
        int unused = 0;
        count_subtree(o, S.blas_nodes, S.inst[slot].blas_root, a.w, tend, nodes, recs, nullptr, unused, 0);
      }
      v[0] = 1; v[1] = nodes; v[2] = recs; v[3] = entered;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
    if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(sums4 + k, v[k]);
  }
}

}  // namespace

static int trace_grid(uint64_t n) {
  const uint64_t blocks = (n + TRACE_THREADS - 1) / TRACE_THREADS;
  const uint64_t persistent = (uint64_t)device_sm_count() * PT_CTAS_PER_SM;
  return (int)(blocks < persistent ? (blocks ? blocks : 1) : persistent);
}

size_t trace_spill_bytes() {
  const int sms = device_sm_count();
  const size_t one_ray = (size_t)sms * 8 * /* the largest persistent grid: 8 CTAs per SM */ (TRACE_THREADS / 32) * SPILL_STACK * 32 * sizeof(uint2);
  const size_t mx = trace_mx_spill_bytes(sms);
  return one_ray > mx ? one_ray : mx;
}

int launch_trace(const TraceScene& S, const float4* rays, uint64_t n, bool any, SnkHit* hits_aos, float4* hits_soa,
                 uint32_t* inst_soa, float* vis_out, unsigned long long* ticket, cudaStream_t st) {
  if (n == 0) return 0;
  if (!ticket || !S.spill) return 0;
  cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), st);
  const int grid = trace_grid(n);
  if (any) k_trace_pt<true><<<grid, TRACE_THREADS, 0, st>>>(S, rays, n, ticket, hits_aos, hits_soa, inst_soa, vis_out);
  else k_trace_pt<false><<<grid, TRACE_THREADS, 0, st>>>(S, rays, n, ticket, hits_aos, hits_soa, inst_soa, vis_out);
  return 1;
}

int launch_count_work(const TraceScene& S, const float4* rays, const SnkHit* hits, uint64_t n, bool any, uint32_t* out3,
                      cudaStream_t st) {
  (void)any;
  if (n == 0) return 0;
  k_count_work<<<(unsigned)((n + TRACE_THREADS - 1) / TRACE_THREADS), TRACE_THREADS, 0, st>>>(S, rays, hits, n, out3);
  return 1;
}

int launch_count_work_sum(const TraceScene& S, const float4* rays, const float4* hits_soa, uint64_t n, unsigned long long* sums4,
                          cudaStream_t st) {
  if (n == 0) return 0;
  k_count_work_sum<<<(unsigned)((n + TRACE_THREADS - 1) / TRACE_THREADS), TRACE_THREADS, 0, st>>>(S, rays, hits_soa, n, sums4);
  return 1;
}

}  // namespace snk
