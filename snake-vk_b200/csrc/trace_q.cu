// trace_q.cu — queue traversal: ONE persistent launch serves every ray the frame has ready at that point.
//
// The north-star frame used to launch the persistent traversal kernel once per ray kind and bounce
// (11 launches on C2: rs_any.init_pl … ns_any.4, profiles/r01n_quick_northstar.txt); the late ones carried
// 50-300 k live rays in 2 M slots and ran at 0.2-0.8 Grays/s because a ray's latency, not throughput, set their
// duration. Here the shading stages append 32-bit ray references (segment << 28 | slot) of their LIVE rays to a
// compact queue (warp-aggregated atomics, northstar.cuh:q_append3), and one launch of k_trace_q drains the queue:
// closest-hit and any-hit rays of different stages mixed in one warp (the kind is a per-lane flag), results written
// to the slot of the ray's own segment — so the output does not depend on queue order.
// Replaces the driver side of vkCmdTraceRaysKHR (Core/src/rendering/Raytracing.cpp:253, :260) like trace.cu.
#include "kernels.h"

#include <cstdlib>

namespace snk {

namespace {

constexpr int TQ_THREADS = 128;
#ifndef SNK_TQ_CTAS
#define SNK_TQ_CTAS 7  // 7 resident CTAs = 72 registers (8 B of spills left) instead of 8 = 64 registers: C2 frame 4.63 -> 4.58 ms, 1 M flat 6.23 -> 6.13 ms;
                       // 6 CTAs (80 registers): 4.78 / 6.49 ms; 5 or 6 short-stack entries at 8 CTAs: 4.57 / 6.17 ms (profiles/r07h_variants_tq_occupancy.txt)
#endif
#ifndef SNK_TQ_GRID_CTAS
#define SNK_TQ_GRID_CTAS SNK_TQ_CTAS  // (r04s, when 8 were resident:) CTAs per SM of the persistent grid. r04s (after instance culling): 5 / 6 / 7 / 8 CTAs give a C2 frame of
                            // 5.68 / 5.56 / 5.56 / 5.53 ms with the two chains overlapped and a kernel sum of 6.28 / 5.98 / 5.80 / 5.66 ms when
                            // every launch runs alone: a smaller grid leaves slots for the other chain's kernels, but a full grid is
                            // faster by about as much, so the full grid stays (profiles/r04s_variants_grid.txt)
#endif
#ifndef SNK_TQ_VOTE
#define SNK_TQ_VOTE 0  // 0: node + triangle + pop every iteration (trav_step); 1: one voted phase per iteration (trav_vote_step) — measured 20 % slower (r01r)
#endif
#ifndef SNK_TQ_CACHE_WORLD
#define SNK_TQ_CACHE_WORLD 1  // world-space 1/d and octant kept in shared memory for leaving an instance: +1-2 % (profiles/r02r_variants_tri_reps_world_cache.txt)
#endif
#ifndef SNK_TQ_COLD
#define SNK_TQ_COLD 1  // world-space ray + closest-hit ids in shared memory instead of registers (traverse.cuh, COLD): no spills around the node test
#endif
#ifndef SNK_TQ_BATCH
#define SNK_TQ_BATCH 0  // 1: queue entries reserved 32 at a time, their references prefetched into registers (one memory round trip per refill instead
                        // of three) — measured slower: C2 4.41 -> 4.53 ms, 1 M flat 5.92 -> 6.15 ms (three more live registers, 12 B of spills; profiles/r07o_variants_refill_batch.txt)
#endif
#ifndef SNK_TQ_REFILL_MIN
#define SNK_TQ_REFILL_MIN 8  // idle lanes that trigger a refill
#endif

__global__ void __launch_bounds__(TQ_THREADS, SNK_TQ_CTAS)
k_trace_q(TraceScene S, TraceSegs segs, const uint32_t* __restrict__ refs, const uint32_t* __restrict__ count_ptr,
          uint32_t* __restrict__ ticket, float* __restrict__ prof_t) {
  __shared__ uint2 sm_stack[SHORT_STACK * TQ_THREADS];
#if SNK_TQ_CACHE_WORLD
  __shared__ float4 sm_world[(SNK_TQ_COLD ? 4 : 1) * TQ_THREADS];
#endif
  const uint32_t lane = threadIdx.x & 31u;
  StackRef stk{sm_stack + threadIdx.x,
               S.spill + ((size_t)blockIdx.x * (TQ_THREADS / 32) + (threadIdx.x >> 5)) * (SPILL_STACK * 32) + lane};
#if SNK_TQ_CACHE_WORLD
  stk.world = sm_world + threadIdx.x;
#endif
  const uint32_t n = __ldcg(count_ptr);
  if (n == 0u || S.tlas_node_count == 0u) {
    // nothing to hit: every queued ray misses
    for (uint32_t q = blockIdx.x * TQ_THREADS + threadIdx.x; q < n; q += gridDim.x * TQ_THREADS) {
      const uint32_t ref = __ldg(refs + q);
      const TraceSeg& sg = segs.s[ref >> REF_SLOT_BITS];
      const uint32_t slot = ref & REF_SLOT_MASK;
      if (sg.vis) sg.vis[slot] = 1.0f;
      if (sg.hit) sg.hit[slot] = make_float4(-1.0f, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu));
      if (sg.inst) sg.inst[slot] = 0xFFFFFFFFu;
      if (prof_t) prof_t[q] = -1.0f;
    }
    return;
  }
  Trav t;
  bool active = false, exhausted = false, any = false;
  uint32_t qpos = 0;
#if SNK_TQ_BATCH
  uint32_t res_next = 0, res_end = 0, ref_cache = 0;
#endif
  while (true) {
    // ---- refill: lanes without a ray take the next queue entries (one atomic per warp)
    const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
    if (!exhausted && __popc(idle) >= SNK_TQ_REFILL_MIN) {
#if SNK_TQ_BATCH
      // the warp reserves 32 queue entries at a time and keeps their references in registers (lane i: entry res_end - 32 + i): a refill
      // then waits for ONE memory round trip (the ray) instead of three in a row (ticket atomic, reference, ray)
      uint32_t need = idle;
#pragma unroll 1
      for (int round = 0; round < 2 && need != 0u; ++round) {
        if (res_next == res_end) {
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(ticket, 32u);
          base = __shfl_sync(0xFFFFFFFFu, base, 0);
          if (base >= n) { exhausted = true; break; }
          res_next = base; res_end = base + 32u;
          ref_cache = base + lane < n ? __ldg(refs + base + lane) : 0u;
        }
        const uint32_t avail = res_end - res_next, rank = __popc(need & ((1u << lane) - 1u));
        const bool take = ((need >> lane) & 1u) != 0u && rank < avail;
        const uint32_t q = res_next + rank;
        const uint32_t ref = __shfl_sync(0xFFFFFFFFu, ref_cache, (q - res_end) & 31u);  // entry q sits in lane q - (res_end - 32)
        if (take && q < n) {
          const TraceSeg& sg = segs.s[ref >> REF_SLOT_BITS];
          const float4* rp = sg.rays + 2ull * (ref & REF_SLOT_MASK);
          const float4 a = __ldcg(rp), b = __ldcg(rp + 1);
          trav_begin(t, a, b);
#if SNK_TQ_CACHE_WORLD
          sm_world[threadIdx.x] = make_float4(t.r.idx, t.r.idy, t.r.idz, __uint_as_float(t.r.octinv));
#if SNK_TQ_COLD
          sm_world[TQ_THREADS + threadIdx.x] = make_float4(a.x, a.y, a.z, 0.f);
          sm_world[2 * TQ_THREADS + threadIdx.x] = make_float4(b.x, b.y, b.z, 0.f);
#endif
#endif
          any = sg.any != 0u;
          qpos = q;
          active = true;  // the producers only queue live rays (tmax > tmin)
        }
        const uint32_t taken = __ballot_sync(0xFFFFFFFFu, take);
        res_next += __popc(taken);
        need &= ~taken;
      }
#else
      const uint32_t cnt = __popc(idle);
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(ticket, cnt);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (!active) {
        const uint32_t q = base + __popc(idle & ((1u << lane) - 1u));
        if (q < n) {
          const uint32_t ref = __ldg(refs + q);
          const TraceSeg& sg = segs.s[ref >> REF_SLOT_BITS];
          const float4* rp = sg.rays + 2ull * (ref & REF_SLOT_MASK);
          const float4 a = __ldcg(rp), b = __ldcg(rp + 1);
          trav_begin(t, a, b);
#if SNK_TQ_CACHE_WORLD
          sm_world[threadIdx.x] = make_float4(t.r.idx, t.r.idy, t.r.idz, __uint_as_float(t.r.octinv));
#if SNK_TQ_COLD
          sm_world[TQ_THREADS + threadIdx.x] = make_float4(a.x, a.y, a.z, 0.f);
          sm_world[2 * TQ_THREADS + threadIdx.x] = make_float4(b.x, b.y, b.z, 0.f);
#endif
#endif
          any = sg.any != 0u;
          qpos = q;
          active = true;  // the producers only queue live rays (tmax > tmin)
        }
      }
      if (base + cnt >= n) exhausted = true;
#endif
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) {
      if (exhausted) break;
      continue;
    }
#if SNK_TQ_VOTE
    const bool finished = trav_vote_step<TQ_THREADS>(S, t, stk, active, any);
#else
    const bool finished = trav_step_impl<TQ_THREADS, SNK_TQ_CACHE_WORLD != 0, SNK_TQ_CACHE_WORLD != 0 && SNK_TQ_COLD != 0>(S, t, stk, active, 0xFFFFFFFFu, any);
#endif
    if (finished) {
      const uint32_t ref = __ldg(refs + qpos);
      const TraceSeg& sg = segs.s[ref >> REF_SLOT_BITS];
      const uint32_t slot = ref & REF_SLOT_MASK;
      if (any) {
        sg.vis[slot] = t.best.t >= 0.0f ? 0.0f : 1.0f;  // Shadow.rmiss:12 → 1 - float(hit)
      } else {
#if SNK_TQ_CACHE_WORLD && SNK_TQ_COLD
        float4 h = make_float4(-1.0f, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu));
        uint32_t hi = 0xFFFFFFFFu;
        if (t.best.t >= 0.0f) {
          const float4 b = sm_world[3 * TQ_THREADS + threadIdx.x];
          h = make_float4(t.best.t, b.x, b.y, b.z);
          hi = __float_as_uint(b.w);
        }
        sg.hit[slot] = h;
        sg.inst[slot] = hi;
#else
        sg.hit[slot] = make_float4(t.best.t, t.best.u, t.best.v, __uint_as_float(t.best.prim));
        sg.inst[slot] = t.best.inst;
#endif
      }
      if (prof_t) prof_t[qpos] = t.best.t;
      active = false;
    }
  }
}

// Diagnostic replay (profiling level 2): order-independent work of every queued ray, as k_count_work_sum (trace.cu);
// sums8 = {live rays, nodes, leaf records, instances, any-hit rays among the live ones}.
__device__ void count_subtree_q(const RaySpace& r0, const uint4* __restrict__ nodes, uint32_t root, float tmin, float tend, uint32_t& n_nodes,
                                uint32_t& leaf_bits_total, uint32_t* leaf_out, int& leaf_n, int leaf_cap) {
  RaySpace r = r0;
  r.octinv = 0;
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

__global__ void __launch_bounds__(TQ_THREADS) k_count_work_q(TraceScene S, TraceSegs segs, const uint32_t* __restrict__ refs,
                                                             const uint32_t* __restrict__ count_ptr, const float* __restrict__ prof_t,
                                                             unsigned long long* __restrict__ sums8) {
  const uint32_t n = __ldcg(count_ptr);
  unsigned long long v[5] = {0, 0, 0, 0, 0};
  for (uint32_t q = blockIdx.x * TQ_THREADS + threadIdx.x; q < n; q += gridDim.x * TQ_THREADS) {
    const uint32_t ref = refs[q];
    const TraceSeg& sg = segs.s[ref >> REF_SLOT_BITS];
    const float4* rp = sg.rays + 2ull * (ref & REF_SLOT_MASK);
    const float4 a = rp[0], b = rp[1];
    if (!(b.w >= a.w) || !S.tlas_node_count) continue;
    const float ht = prof_t[q];
    const float tend = ht >= 0.f ? ht : b.w;
    uint32_t nodes = 0, recs = 0, insts = 0, entered = 0;
    RaySpace r;
    setup_ray_space(r, a.x, a.y, a.z, b.x, b.y, b.z);
    uint32_t leaf[64];
    int ln = 0;
    count_subtree_q(r, S.tlas_nodes, 0u, a.w, tend, nodes, insts, leaf, ln, 64);
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
      count_subtree_q(o, S.blas_nodes, S.inst[slot].blas_root, a.w, tend, nodes, recs, nullptr, unused, 0);
    }
    insts = entered;  // instances actually entered (after the bounding-sphere cull)
    v[0] += 1; v[1] += nodes; v[2] += recs; v[3] += insts; v[4] += sg.any ? 1 : 0;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
    if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(sums8 + k, v[k]);
  }
}

}  // namespace

int device_sm_count() {
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

// SNKPT_TRACE_KERNEL=mx selects the multiplexed kernel of trace_mx.cu (A/B measurements: bit-identical results, 21 instead of
// 15 active threads per warp and 18 % fewer instructions, but 4-35 % slower because its shared-memory state caps the
// occupancy — profiles/r04a_*, r04b_*); default: the one-ray-per-lane kernel of this file
static bool use_mx_kernel() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SNKPT_TRACE_KERNEL"); v = (e && e[0] == 'm') ? 1 : 0; }
  return v == 1;
}

int launch_trace_q(const TraceScene& S, const TraceSegs& segs, const RefQueue& q, float* prof_t, cudaStream_t st) {
  if (!q.refs || !q.count || !q.ticket || !S.spill || q.cap == 0) return 0;
  if (use_mx_kernel()) return launch_trace_mx(S, segs, q, prof_t, device_sm_count(), st);
  // persistent grid; a queue that cannot fill it gets one CTA per 128 possible entries
  const uint64_t blocks = ((uint64_t)q.cap + TQ_THREADS - 1) / TQ_THREADS;
  const uint64_t persistent = (uint64_t)device_sm_count() * SNK_TQ_GRID_CTAS;
  const int grid = (int)(blocks < persistent ? blocks : persistent);
  k_trace_q<<<grid, TQ_THREADS, 0, st>>>(S, segs, q.refs, q.count, q.ticket, prof_t);
  return 1;
}

int launch_count_work_q(const TraceScene& S, const TraceSegs& segs, const RefQueue& q, const float* prof_t, unsigned long long* sums8,
                        cudaStream_t st) {
  if (!q.refs || !q.count || q.cap == 0 || !prof_t) return 0;
  const uint64_t blocks = ((uint64_t)q.cap + TQ_THREADS - 1) / TQ_THREADS;
  k_count_work_q<<<(unsigned)(blocks < (uint64_t)device_sm_count() * 16 ? blocks : (uint64_t)device_sm_count() * 16), TQ_THREADS, 0, st>>>(S, segs, q.refs, q.count, prof_t, sums8);
  return 1;
}

}  // namespace snk
