// trace_mx.cu — queue traversal with K rays multiplexed per lane ("mx"): the traversal kernel of the north-star frame.
//
// k_trace_q (trace_q.cu) keeps ONE ray per lane and runs instance entry, the node test, the triangle test and the pop of
// every lane in every iteration; ncu showed that only the node test is reasonably filled (23.5 of 32 lanes) while entry,
// Woop set-up, triangle test, pop and result write run at 2-8 lanes in ~96 % of the iterations: 46 % of the issue slots
// for ~5 lanes of work (profiles/r01u_sass_blocks_fused.txt, r03h_prof_trace.txt: 15.4 active threads per warp).
// Voting for ONE phase per iteration did not help with one ray per lane, because the lanes outside the winning phase
// idle (profiles/r01r_variants_vote_vs_fused.txt).
//
// Here every lane owns MX_K rays whose whole traversal state lives in SHARED memory (80 B + a short stack per ray,
// laid out in 16-byte chunks interleaved by thread, so every access is a conflict-free LDS.128 / STS.128) and keeps in
// registers only one byte per ray that names the phase this ray needs next. Each iteration the warp votes for the
// phase most lanes can serve, every lane picks (round robin) one of ITS rays that wants this phase, and the phase runs
// once at (nearly) full width. No state ever moves between lanes; rays that want another phase simply wait a turn.
//   FETCH  retire finished rays (result write) and pull new references from the queue ticket
//   NODE   open the nearest pending child: 5 x 128-bit node load + 8 quantised box tests (+ pop when nothing was hit)
//   TRI    one watertight triangle test of the pending leaf group (+ pop)
//   SPACE  change of space: enter the next instance of the pending TLAS leaf group (ray transform, box + Woop constants)
//          or return to world space after the last BLAS entry was popped
// Results are written to the slot of the ray's own segment and the closest hit is the one with the lowest
// (t, instance, primitive), so the output does not depend on the schedule: same bits as k_trace_q and the CPU oracle.
// Replaces the driver side of vkCmdTraceRaysKHR (Core/src/rendering/Raytracing.cpp:253, :260).
#include "kernels.h"

#include <cstdlib>

namespace snk {

namespace {

#ifndef SNK_MX_K
#define SNK_MX_K 3  // rays per lane (1..4)
#endif
#ifndef SNK_MX_THREADS
#define SNK_MX_THREADS 128
#endif
#ifndef SNK_MX_CTAS
#define SNK_MX_CTAS 3  // resident CTAs per SM the launch bounds ask for
#endif
#ifndef SNK_MX_SHORT
#define SNK_MX_SHORT 4  // stack entries per ray in shared memory
#endif
#ifndef SNK_MX_FETCH_MIN
#define SNK_MX_FETCH_MIN 12  // lanes with a finished / empty slot that trigger a FETCH phase
#endif
#ifndef SNK_MX_TRI_BIAS
#define SNK_MX_TRI_BIAS 0  // the TRI phase wins the vote against NODE with this many fewer lanes (earlier hits shrink tcur earlier)
#endif

constexpr int MX_K = SNK_MX_K, MX_T = SNK_MX_THREADS, MX_SHORT = SNK_MX_SHORT, MX_SPILL = 44;
static_assert(MX_K >= 1 && MX_K <= 4, "one phase byte per ray in a 32-bit word");
constexpr int MX_CHUNKS = 5;
constexpr size_t MX_SMEM = (size_t)MX_K * MX_T * (16 * MX_CHUNKS + 8 * MX_SHORT);

// phase byte of a ray slot (one hot)
enum : uint32_t { W_EMPTY = 1u, W_NODE = 2u, W_TRI = 4u, W_SPACE = 8u, W_DONE = 16u };
__host__ __device__ constexpr uint32_t rep4(uint32_t b) { return b * 0x01010101u; }
// packed word of the D chunk: stack height | stack height at instance entry | Woop axes | inverse octant | flags
constexpr uint32_t P_SP = 0xFFu, P_BSP_SH = 8u, P_K_SH = 16u, P_OCT_SH = 22u, P_INBLAS = 1u << 25, P_ANY = 1u << 26, P_HIT = 1u << 27;

// chunk c of ray slot k of this thread
#define MX_AT(c, k) sm[((c) * MX_K + (k)) * MX_T + tid]
enum { CH_A = 0 /* o.xyz, queue position */, CH_B = 1 /* 1/d.xyz, instance */, CH_C = 2 /* G, T */, CH_D = 3 /* tmin, tcur, packed, ref */,
       CH_S = 4 /* Woop shear, - */ };

struct MxStack {
  uint2* sm;   // entry e of slot k: sm[(e * MX_K + k) * MX_T]
  uint2* gl;   // entry e of slot k: gl[(k * MX_SPILL + e) * MX_T]
  uint32_t* overflow;
};
__device__ __forceinline__ void mx_push(const MxStack& s, uint32_t k, uint32_t& sp, uint2 v) {
  if (sp < (uint32_t)MX_SHORT) s.sm[(sp * MX_K + k) * MX_T] = v;
  else if (sp < (uint32_t)(MX_SHORT + MX_SPILL)) s.gl[(k * MX_SPILL + (sp - MX_SHORT)) * MX_T] = v;
  else atomicAdd(s.overflow, 1u);
  sp += sp < (uint32_t)(MX_SHORT + MX_SPILL) ? 1u : 0u;
}
__device__ __forceinline__ uint2 mx_pop_entry(const MxStack& s, uint32_t k, uint32_t& sp) {
  --sp;
  return sp < (uint32_t)MX_SHORT ? s.sm[(sp * MX_K + k) * MX_T] : s.gl[(k * MX_SPILL + (sp - MX_SHORT)) * MX_T];
}

// nothing pending: pop the next group; returns the phase the ray needs next
__device__ __forceinline__ uint32_t mx_pop(const MxStack& s, uint32_t k, uint32_t& packed, uint2& G, uint2& T) {
  uint32_t sp = packed & P_SP;
  bool leaving = false;
  if ((packed & P_INBLAS) && sp == ((packed >> P_BSP_SH) & 0xFFu)) { packed &= ~P_INBLAS; leaving = true; }
  if (sp == 0u) return W_DONE;
  const uint2 e = mx_pop_entry(s, k, sp);
  packed = (packed & ~P_SP) | sp;
  const bool inner = (e.y & 0xFF000000u) != 0u;
  if (inner) { G = e; T = make_uint2(0u, 0u); }
  else { T = e; G = make_uint2(0u, 0u); }
  if (packed & P_INBLAS) return inner ? W_NODE : W_TRI;
  // TLAS: a leaf group holds instances (SPACE enters the next one); after leaving an instance the world-space box
  // constants have to be restored before a TLAS node can be opened (SPACE does that when no instance is pending)
  return (inner && !leaving) ? W_NODE : W_SPACE;
}

__global__ void __launch_bounds__(MX_T, SNK_MX_CTAS)
k_trace_mx(TraceScene S, TraceSegs segs, const uint32_t* __restrict__ refs, const uint32_t* __restrict__ count_ptr,
           uint32_t* __restrict__ ticket, float* __restrict__ prof_t) {
  extern __shared__ uint4 sm[];
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint32_t n = __ldcg(count_ptr);
  if (n == 0u || S.tlas_node_count == 0u) {
    // nothing to hit: every queued ray misses
    for (uint32_t q = blockIdx.x * MX_T + tid; q < n; q += gridDim.x * MX_T) {
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
  MxStack stk;
  stk.sm = reinterpret_cast<uint2*>(sm + MX_CHUNKS * MX_K * MX_T) + tid;
  stk.gl = S.spill + (size_t)blockIdx.x * (MX_K * MX_SPILL * MX_T) + tid;
  stk.overflow = S.overflow;

  uint32_t w = 0;  // one phase byte per ray slot
#pragma unroll
  for (int k = 0; k < MX_K; ++k) w |= W_EMPTY << (8 * k);
  bool exhausted = false;
  uint32_t rr = 0;
  while (true) {
    // ---- vote: the phase most lanes can serve
    const uint32_t want_fetch = w & (rep4(W_DONE) | (exhausted ? 0u : rep4(W_EMPTY)));
    const int n_fetch = __popc(__ballot_sync(0xFFFFFFFFu, want_fetch != 0u));
    const int n_node = __popc(__ballot_sync(0xFFFFFFFFu, (w & rep4(W_NODE)) != 0u));
    const int n_tri = __popc(__ballot_sync(0xFFFFFFFFu, (w & rep4(W_TRI)) != 0u));
    const int n_space = __popc(__ballot_sync(0xFFFFFFFFu, (w & rep4(W_SPACE)) != 0u));
    uint32_t phase = W_NODE;
    int best = n_node;
    if (n_tri + (n_tri ? SNK_MX_TRI_BIAS : 0) > best) { phase = W_TRI; best = n_tri + SNK_MX_TRI_BIAS; }
    if (n_space > best) { phase = W_SPACE; best = n_space; }
    if (n_fetch >= SNK_MX_FETCH_MIN || best == 0) {
      if (n_fetch == 0) break;  // queue drained, every slot empty
      phase = W_EMPTY;
    }
    // ---- my ray for this phase (round robin over my slots)
    const uint32_t sel_mask = phase == W_EMPTY ? want_fetch : (w & rep4(phase));
    const uint32_t rot = __funnelshift_r(sel_mask, sel_mask, 8u * rr);
    const bool mine = rot != 0u;
    const uint32_t k = mine ? ((((uint32_t)__ffs(rot) - 1u) >> 3) + rr) & 3u : 0u;
    rr = (rr + 1u) & 3u;
    uint32_t next = 0;  // new phase byte of slot k (mine only)

    if (phase == W_NODE) {
      if (mine) {
        const uint4 A = MX_AT(CH_A, k), B = MX_AT(CH_B, k), C = MX_AT(CH_C, k), D = MX_AT(CH_D, k);
        uint32_t packed = D.z, sp = packed & P_SP;
        uint2 G = make_uint2(C.x, C.y), T;
        RaySpace r;
        r.ox = __uint_as_float(A.x); r.oy = __uint_as_float(A.y); r.oz = __uint_as_float(A.z);
        r.idx = __uint_as_float(B.x); r.idy = __uint_as_float(B.y); r.idz = __uint_as_float(B.z);
        r.octinv = (packed >> P_OCT_SH) & 7u;
        const uint32_t hits_imask = G.y;
        const uint32_t bit = 31u - __clz(hits_imask);
        G.y &= ~(1u << bit);
        if (G.y & 0xFF000000u) mx_push(stk, k, sp, G);
        const uint32_t cslot = (bit - 24u) ^ r.octinv;
        const uint32_t node_idx = G.x + __popc(hits_imask & ~(0xFFFFFFFFu << cslot) & 0xFFu);
        const uint4* nodes = (packed & P_INBLAS) ? S.blas_nodes : S.tlas_nodes;
        uint4 n0, n1, n2, n3, n4;
        load_node(nodes, node_idx, n0, n1, n2, n3, n4);
        const uint32_t hm = intersect_node(r, n0, n1, n2, n3, n4, __uint_as_float(D.x), __uint_as_float(D.y));
        G = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
        T = make_uint2(n1.y, hm & 0x00FFFFFFu);
        packed = (packed & ~P_SP) | sp;
        if (T.y) next = (packed & P_INBLAS) ? W_TRI : W_SPACE;
        else if (G.y & 0xFF000000u) next = W_NODE;
        else next = mx_pop(stk, k, packed, G, T);
        MX_AT(CH_C, k) = make_uint4(G.x, G.y, T.x, T.y);
        reinterpret_cast<uint32_t*>(&MX_AT(CH_D, k))[2] = packed;
      }
    } else if (phase == W_TRI) {
      if (mine) {
        const uint4 A = MX_AT(CH_A, k), C = MX_AT(CH_C, k), D = MX_AT(CH_D, k), Sh = MX_AT(CH_S, k);
        uint32_t packed = D.z;
        uint2 G = make_uint2(C.x, C.y), T = make_uint2(C.z, C.w);
        const float tmin = __uint_as_float(D.x);
        float tcur = __uint_as_float(D.y);
        RaySpace r;
        r.ox = __uint_as_float(A.x); r.oy = __uint_as_float(A.y); r.oz = __uint_as_float(A.z);
        r.Sx = __uint_as_float(Sh.x); r.Sy = __uint_as_float(Sh.y); r.Sz = __uint_as_float(Sh.z);
        r.kx = (packed >> P_K_SH) & 3u; r.ky = (packed >> (P_K_SH + 2)) & 3u; r.kz = (packed >> (P_K_SH + 4)) & 3u;
        const uint32_t bit = __ffs(T.y) - 1;
        T.y &= ~(1u << bit);
        const float4* tp = S.tris + 3ull * (T.x + bit);
        const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
        float th, u, v;
        bool finished = false;
        if (tri_test(r, p0, p1, p2, th, u, v) && th > tmin) {
          const uint32_t prim = __float_as_uint(p0.w);
          const bool any = (packed & P_ANY) != 0u;
          bool take = th < tcur;
          const TraceSeg& sg = segs.s[D.w >> REF_SLOT_BITS];
          const uint32_t slot = D.w & REF_SLOT_MASK;
          const uint32_t cur_inst = reinterpret_cast<const uint32_t*>(&MX_AT(CH_B, k))[3];
          if (!any && th == tcur && (packed & P_HIT)) {
            // at the current closest distance the lower (instance, primitive) wins; the best hit so far is in the result slot
            const uint32_t best_inst = sg.inst[slot], best_prim = __float_as_uint(sg.hit[slot].w);
            take = cur_inst < best_inst || (cur_inst == best_inst && prim < best_prim);
          }
          if (take) {
            tcur = th;
            packed |= P_HIT;
            if (any) finished = true;
            else { sg.hit[slot] = make_float4(th, u, v, __uint_as_float(prim)); sg.inst[slot] = cur_inst; }
          }
        }
        if (finished) next = W_DONE;
        else if (T.y) next = W_TRI;
        else if (G.y & 0xFF000000u) next = W_NODE;
        else next = mx_pop(stk, k, packed, G, T);
        MX_AT(CH_C, k) = make_uint4(G.x, G.y, T.x, T.y);
        uint2* dp = reinterpret_cast<uint2*>(&MX_AT(CH_D, k));
        dp[0].y = __float_as_uint(tcur);
        dp[1].x = packed;
      }
    } else if (phase == W_SPACE) {
      if (mine) {
        const uint4 C = MX_AT(CH_C, k), D = MX_AT(CH_D, k);
        uint32_t packed = D.z, sp = packed & P_SP;
        uint2 G = make_uint2(C.x, C.y), T = make_uint2(C.z, C.w);
        const TraceSeg& sg = segs.s[D.w >> REF_SLOT_BITS];
        const float4* rp = sg.rays + 2ull * (D.w & REF_SLOT_MASK);
        const float4 a = __ldcg(rp), b = __ldcg(rp + 1);
        RaySpace r;
        bool entered = false;
        if (T.y) {
          // enter the next instance of the pending TLAS leaf group; the rest of the group and the pending inner group wait
          const uint32_t bit = __ffs(T.y) - 1;
          T.y &= ~(1u << bit);
          const uint32_t islot = __ldg(S.tlas_leaf_ids + T.x + bit);
          if (!S.inst_sphere || ray_may_hit_sphere(__ldg(S.inst_sphere + islot), a.x, a.y, a.z, b.x, b.y, b.z, __uint_as_float(D.y))) {
            entered = true;
            if (T.y) mx_push(stk, k, sp, T);
            if (G.y & 0xFF000000u) mx_push(stk, k, sp, G);
            const float4* M = reinterpret_cast<const float4*>(S.inst + islot);
            const float4 r0 = __ldg(M), r1 = __ldg(M + 1), r2 = __ldg(M + 2);
            const uint4 meta = __ldg(reinterpret_cast<const uint4*>(M + 3));
            // o' = ((m0*x + m1*y) + m2*z) + m3 ; d' = (m0*x + m1*y) + m2*z   (unfused, as the oracle)
            const float oox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, a.x), __fmul_rn(r0.y, a.y)), __fmul_rn(r0.z, a.z)), r0.w);
            const float ooy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, a.x), __fmul_rn(r1.y, a.y)), __fmul_rn(r1.z, a.z)), r1.w);
            const float ooz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, a.x), __fmul_rn(r2.y, a.y)), __fmul_rn(r2.z, a.z)), r2.w);
            const float odx = __fadd_rn(__fadd_rn(__fmul_rn(r0.x, b.x), __fmul_rn(r0.y, b.y)), __fmul_rn(r0.z, b.z));
            const float ody = __fadd_rn(__fadd_rn(__fmul_rn(r1.x, b.x), __fmul_rn(r1.y, b.y)), __fmul_rn(r1.z, b.z));
            const float odz = __fadd_rn(__fadd_rn(__fmul_rn(r2.x, b.x), __fmul_rn(r2.y, b.y)), __fmul_rn(r2.z, b.z));
            setup_box_space(r, oox, ooy, ooz, odx, ody, odz);
            setup_tri_space(r);
            packed = (packed & (P_ANY | P_HIT)) | sp | (sp << P_BSP_SH) | ((uint32_t)r.kx << P_K_SH) | ((uint32_t)r.ky << (P_K_SH + 2)) |
                     ((uint32_t)r.kz << (P_K_SH + 4)) | (r.octinv << P_OCT_SH) | P_INBLAS;
            MX_AT(CH_S, k) = make_uint4(__float_as_uint(r.Sx), __float_as_uint(r.Sy), __float_as_uint(r.Sz), 0u);
            reinterpret_cast<uint32_t*>(&MX_AT(CH_B, k))[3] = meta.y;
            G = make_uint2(meta.x, 0x01000000u);  // one inner bit, empty child mask: the BLAS root
            T = make_uint2(0u, 0u);
            next = W_NODE;
          }
        }
        if (!entered) {
          // back in (or still in) world space: the box constants of the world-space ray; an instance whose bounding sphere the
          // ray misses was skipped above and the rest of its leaf group is tried in the next SPACE phase
          setup_box_space(r, a.x, a.y, a.z, b.x, b.y, b.z);
          packed = (packed & ~(7u << P_OCT_SH)) | (r.octinv << P_OCT_SH) | sp;
          packed = (packed & ~P_SP) | sp;
          if (T.y) next = W_SPACE;
          else if (G.y & 0xFF000000u) next = W_NODE;
          else next = mx_pop(stk, k, packed, G, T);
        }
        uint4* ap = &MX_AT(CH_A, k);
        reinterpret_cast<float*>(ap)[0] = r.ox; reinterpret_cast<float*>(ap)[1] = r.oy; reinterpret_cast<float*>(ap)[2] = r.oz;
        uint4* bp = &MX_AT(CH_B, k);
        reinterpret_cast<float*>(bp)[0] = r.idx; reinterpret_cast<float*>(bp)[1] = r.idy; reinterpret_cast<float*>(bp)[2] = r.idz;
        MX_AT(CH_C, k) = make_uint4(G.x, G.y, T.x, T.y);
        reinterpret_cast<uint32_t*>(&MX_AT(CH_D, k))[2] = packed;
      }
    } else {
      // ---- FETCH: retire my finished ray, then take the next queue entry (one atomic per warp)
      const bool retire = mine && ((w >> (8u * k)) & W_DONE) != 0u;
      if (retire) {
        const uint4 D = MX_AT(CH_D, k);
        const TraceSeg& sg = segs.s[D.w >> REF_SLOT_BITS];
        const uint32_t slot = D.w & REF_SLOT_MASK;
        const bool hit = (D.z & P_HIT) != 0u;
        if (D.z & P_ANY) sg.vis[slot] = hit ? 0.0f : 1.0f;  // Shadow.rmiss:12 → 1 - float(hit)
        else if (!hit) { sg.hit[slot] = make_float4(-1.0f, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu)); sg.inst[slot] = 0xFFFFFFFFu; }
        if (prof_t) prof_t[reinterpret_cast<const uint32_t*>(&MX_AT(CH_A, k))[3]] = hit ? __uint_as_float(D.y) : -1.0f;
      }
      next = W_EMPTY;
      if (!exhausted) {
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, mine);
        const uint32_t cnt = __popc(need);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(ticket, cnt);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (mine) {
          const uint32_t q = base + __popc(need & ((1u << lane) - 1u));
          if (q < n) {
            const uint32_t ref = __ldg(refs + q);
            const TraceSeg& sg = segs.s[ref >> REF_SLOT_BITS];
            const float4* rp = sg.rays + 2ull * (ref & REF_SLOT_MASK);
            const float4 a = __ldcg(rp), b = __ldcg(rp + 1);
            RaySpace r;
            setup_box_space(r, a.x, a.y, a.z, b.x, b.y, b.z);
            MX_AT(CH_A, k) = make_uint4(__float_as_uint(r.ox), __float_as_uint(r.oy), __float_as_uint(r.oz), q);
            MX_AT(CH_B, k) = make_uint4(__float_as_uint(r.idx), __float_as_uint(r.idy), __float_as_uint(r.idz), 0u);
            MX_AT(CH_C, k) = make_uint4(0u, 0x80000000u, 0u, 0u);  // the TLAS root as a one-child inner group
            MX_AT(CH_D, k) = make_uint4(__float_as_uint(a.w), __float_as_uint(b.w), (r.octinv << P_OCT_SH) | (sg.any ? P_ANY : 0u), ref);
            next = W_NODE;  // the producers only queue live rays (tmax > tmin)
          }
        }
        if (base + cnt >= n) exhausted = true;
      }
    }
    if (mine) w = (w & ~(0xFFu << (8u * k))) | (next << (8u * k));
    __syncwarp();
  }
}

int g_mx_sms = 0;

}  // namespace

size_t trace_mx_spill_bytes(int sms) { return (size_t)sms * SNK_MX_CTAS * MX_K * MX_SPILL * MX_T * sizeof(uint2); }

int launch_trace_mx(const TraceScene& S, const TraceSegs& segs, const RefQueue& q, float* prof_t, int sms, cudaStream_t st) {
  if (!q.refs || !q.count || !q.ticket || !S.spill || q.cap == 0) return 0;
  if (g_mx_sms == 0) {
    if (cudaFuncSetAttribute(k_trace_mx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MX_SMEM) != cudaSuccess) return 0;
    // ask for exactly the shared memory the resident CTAs need: the rest of the 256 KB stays L1 for the BVH nodes (with the
    // default carve-out the driver picked 233 KB and the L1 hit rate fell from 74 % to 57 %, profiles/r04a)
    const size_t need = (size_t)SNK_MX_CTAS * (MX_SMEM + 1024);
    int carve = (int)((need * 100 + 233472 - 1) / 233472);
    if (carve > 100) carve = 100;
    cudaFuncSetAttribute(k_trace_mx, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    g_mx_sms = sms;
  }
  // persistent grid; a queue that cannot fill it gets one CTA per MX_T * MX_K possible entries
  const uint64_t blocks = ((uint64_t)q.cap + MX_T * MX_K - 1) / (MX_T * MX_K);
  const uint64_t persistent = (uint64_t)sms * SNK_MX_CTAS;
  const int grid = (int)(blocks < persistent ? blocks : persistent);
  k_trace_mx<<<grid, MX_T, MX_SMEM, st>>>(S, segs, q.refs, q.count, q.ticket, prof_t);
  return 1;
}

}  // namespace snk
