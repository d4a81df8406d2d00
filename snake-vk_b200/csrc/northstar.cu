// northstar.cu — the north-star estimator as wavefront stages (BASELINE.json north_star items 3 + 4):
// 1 spp, up to max_bounces path vertices, GGX / Lambert BSDF sampling, MIS next-event estimation, Russian roulette.
//
// No counterpart in the reference (its estimator is the fixed 8x(1+2) tree of RayGen.rgen:139-213);
// shared with it are the material model (LightBuffers.glsl:65-144), the light records, the emission convention
// (RayClosestHit.rchit:119-123: emitted radiance = material albedo) and the direct term at the primary hit, which is
// produced by the same ReSTIR DI kernels as in reference mode (restir.cu). The CPU statement this file must match bit
// for bit is oracle/snk_northstar.h (same expression order, -fmad=false, float64 sin / cos rounded once).
//
//   k_ns_gen      pixel -> vertex x0 from the primary-hit ids: BSDF sample -> ray queue, throughput
//   repeat b = 1..max_bounces:
//     trace closest (persistent wavefront kernel, trace.cu)
//     k_ns_vertex pixel -> adds the previous vertex's NEE (C x visibility); hit -> surface; emissive hit -> MIS-weighted
//                 emission; else NEE sample (C + shadow ray), Russian roulette, BSDF sample -> next ray
//     trace any   shadow rays -> visibility
//   k_ns_resolve  pixel -> last NEE term, out = direct + indirect (or emission / sky)
// One path per pixel, so every queue is a per-pixel plane (slot = pixel of the band); dead paths carry tmax < 0.
// All surfaces share one BSDF model, so there is nothing to sort the shading queue by material for.
#include "kernels.h"
#include "northstar.cuh"

namespace snk {

// SNK_NS_INLINE: the path-vertex stages inline the surface / BSDF / pdf helpers (a handful of call sites per kernel; results are
// the same bits, -fmad=false). 0 keeps real calls (the round-1 rule against instruction-cache pressure).
#ifndef SNK_NS_INLINE
#define SNK_NS_INLINE 1
#endif
#if SNK_NS_INLINE
#define NS_SURFACE ns_surface_inl
#define NS_SURFACE_EM ns_surface_em_inl
#define NS_BSDF_COS ns_bsdf_cos_inl
#define NS_PDF ns_pdf_inl
#else
#define NS_SURFACE ns_surface
#define NS_SURFACE_EM ns_surface_em
#define NS_BSDF_COS ns_bsdf_cos
#define NS_PDF ns_pdf
#endif

namespace {

struct NsBuffers {
  float4* rays;     // 2 x float4 per pixel: continuation ray
  float4* hit;      // t,u,v,prim
  uint32_t* inst;
  float4* T;        // throughput rgb, pdf of the BSDF sample that produced the ray
  float4* L;        // indirect radiance so far
  float4* srays;    // 2 x float4 per pixel: NEE shadow ray
  float4* C;        // pending NEE contribution (unshadowed)
  float* vis;       // visibility of the pending NEE ray
};

// work of path vertex x_b (b = 0: the primary hit): NEE (b >= 1), Russian roulette (b >= 2), BSDF sample.
// Writes the shadow ray + pending C, the continuation ray, throughput and pdf for slot i.
__device__ void ns_vertex_work(const SceneDev& D, const NsBuffers& B, size_t i, uint32_t pixel, uint32_t frame, uint32_t b, uint32_t max_bounces,
                               uint32_t nee_from, const NsSurf& s, vec3 wo, vec3 T, bool& has_shadow, bool& has_closest) {
  const NsLightCounts lc = ns_light_counts(D);
  vec3 n = s.n;
  if (dot(n, wo) < 0.0f) n = -n;
  const vec3 o = s.p + n * 0.01f;
  const uint32_t dim = 16u * b;
  vec3 C = v3(0.0f);
  bool shadow = false;
  if (b >= nee_from && lc.n_types > 0) {
    const uint32_t t = ns_pick(ns_rand(pixel, frame, dim + 3), lc.n_types);
    uint32_t type = 0;
    {
      uint32_t k = 0;
      bool found = false;
      if (lc.has_dir) { if (k == t) { type = 0; found = true; } ++k; }
      if (!found && lc.n_pl) { if (k == t) { type = 1; found = true; } ++k; }
      if (!found && lc.n_em) { if (k == t) { type = 2; found = true; } ++k; }
    }
    const float p_type = 1.0f / (float)lc.n_types;
    const float u1 = ns_rand(pixel, frame, dim + 4), u2 = ns_rand(pixel, frame, dim + 5);
    const float u3 = ns_rand(pixel, frame, dim + 6), u4 = ns_rand(pixel, frame, dim + 7);
    // The three light types first only prepare their sample (direction, range, emitted colour, the scalars of their estimator);
    // the BSDF is then evaluated ONCE, by all lanes with a valid sample together. With one ns_bsdf_cos call per branch the stage
    // ran at 13-15 of 32 lanes (ncu r04r: the light type is a per-lane random choice). Same expressions in the same order as
    // before (and as oracle/snk_northstar.h), so the values do not change.
    vec3 wi = v3(0.0f, 1.0f, 0.0f), emitted = v3(0.0f);
    float tmax = -1.0f, s0 = 1.0f, s1 = 1.0f;  // type 1: s0 = attenuation, s1 = p_type / n_pl; type 2: s0 = MIS weight (set below), s1 = pdf_w
    bool valid = false;
    if (type == 0) {
      const float4 dl = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.dir));
      wi = v3(dl.x, dl.y, dl.z);
      if (dot(n, wi) > 0.0f) {
        const float4 dc = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.colour));
        emitted = v3(dc.x, dc.y, dc.z);
        tmax = 10000.0f;
        valid = true;
      }
    } else if (type == 1) {
      const uint32_t idx = ns_pick(u1, lc.n_pl);
      const PointLightRec pl = load_pointlight(D.lights, idx);
      const vec3 d = pl.pos - o;
      const float dist = length(d);
      wi = d / dist;
      if (dot(n, wi) > 0.0f && dist > 0.002f) {
        s0 = (pl.atten_constant + pl.atten_linear * dist) + pl.atten_exp * (dist * dist);
        s1 = p_type * (1.0f / (float)lc.n_pl);
        emitted = pl.colour;
        tmax = dist - 0.001f;
        valid = true;
      }
    } else {
      const uint32_t obj = __ldg(D.emissive_idx + ns_pick(u1, lc.n_em));
      const uint32_t ntri = load_instance(D, obj).num_mesh_indices / 3u;
      const uint32_t tri = ns_pick(u2, ntri);
      float bu = u3, bv = u4;
      if (bu + bv >= 1.0f) { bu = 1.0f - bu; bv = 1.0f - bv; }
      const NsSurf q = NS_SURFACE_EM(D, obj, tri, bu, bv);
      const vec3 d = q.p - o;
      const float dist = length(d);
      wi = d / dist;
      const float pdf_w = ns_nee_pdf_emissive(lc, ntri, q.area, dist, q.n, wi);
      if (dot(n, wi) > 0.0f && pdf_w > 0.0f && dist > 0.01f) {
        s1 = pdf_w;
        emitted = q.albedo;
        tmax = dist - 0.005f;
        valid = true;
      }
    }
    if (valid) {
      if (type == 2) {
        const float p_bsdf = NS_PDF(n, wo, wi, s.roughness, s.metallic);
        s0 = (s1 * s1) / (s1 * s1 + p_bsdf * p_bsdf);
      }
      const vec3 tb = (T * NS_BSDF_COS(n, wo, wi, s.albedo, s.roughness, s.metallic)) * emitted;
      if (type == 0) C = tb / p_type;
      else if (type == 1) C = (tb / s0) / s1;
      else C = (tb * s0) / s1;
    }
    if (tmax > 0.0f) {
      put_ray(B.srays, i, o, 0.001f, wi, tmax);
      shadow = true;
    } else {
      C = v3(0.0f);
    }
  }
  if (!shadow) B.vis[i] = 1.0f;  // no shadow ray: the pending contribution is zero and counts as unoccluded
  has_shadow = shadow;
  B.C[i] = make_float4(C.x, C.y, C.z, 0.0f);
  // ---- continuation
  bool alive = b < max_bounces;
  if (alive && b >= 2) {
    const float q = gmin(gmax(gmax(T.x, T.y), T.z), 0.95f);
    if (!(ns_rand(pixel, frame, dim + 8) < q)) alive = false;
    else T = T / q;
  }
  vec3 wi = v3(0.0f, 1.0f, 0.0f);
  float pdf_b = 0.0f;
  if (alive) alive = ns_sample(n, wo, s.roughness, s.metallic, ns_rand(pixel, frame, dim), ns_rand(pixel, frame, dim + 1), ns_rand(pixel, frame, dim + 2), wi);
  if (alive) {
    pdf_b = NS_PDF(n, wo, wi, s.roughness, s.metallic);
    alive = pdf_b > 0.0f;
  }
  if (alive) {
    T = (T * NS_BSDF_COS(n, wo, wi, s.albedo, s.roughness, s.metallic)) / pdf_b;
    alive = gmax(gmax(T.x, T.y), T.z) > 0.0f;
  }
  if (alive) put_ray(B.rays, i, o, 0.001f, wi, 1000.0f);
  else put_dead_dir(B.rays, i);
  has_closest = alive;
  B.T[i] = make_float4(T.x, T.y, T.z, pdf_b);
}

__global__ void __launch_bounds__(128, 8) k_ns_gen(SceneDev D, FrameConsts fc, PixelPlanes P, NsBuffers B, uint32_t npix, uint32_t max_bounces,
                                                uint32_t nee_from, unsigned long long* counters, uint32_t* q_refs, uint32_t* q_count,
                                                uint32_t* paths_out, uint32_t* paths_out_count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  bool has_shadow = false, has_closest = false;
  uint32_t x, y, i = 0;
  if (stage_pixel(fc, t, x, y)) {
    i = (y - fc.row0) * fc.W + x;  // slot of the per-pixel path planes (band-relative pixel index)
    const size_t px = (size_t)y * fc.W + x;
    B.L[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.vis[i] = 1.0f;
    bool alive = P.prim[px] != INVALID_IDX;
    NsSurf s;
    if (alive) {
      const float2 bc = P.bary[px];
      s = NS_SURFACE(D, P.inst[px], P.prim[px], bc.x, bc.y);
      alive = !(s.flags & SNKPT_MAT_FLAG_EMISSIVE);
    }
    if (alive) {
      const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
      const vec3 wo = normalize(cam - s.p);
      ns_vertex_work(D, B, i, (uint32_t)px, fc.frame_idx, 0u, max_bounces, nee_from, s, wo, v3(1.0f), has_shadow, has_closest);
    } else {
      put_dead_dir(B.rays, i);
      B.C[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      B.T[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  q_append3(q_refs, q_count, has_closest, make_ref(SEG_NS_CLOSEST, i), has_shadow, make_ref(SEG_NS_SHADOW, i), false, 0u);
  q_append1(paths_out, paths_out_count, has_closest, i);
  flush_counts(has_shadow ? 1ull : 0ull, has_closest ? 1ull : 0ull, counters);
}

// the hit of the ray that left vertex x_{b-1} becomes vertex x_b. One thread per path that is still alive (compact list
// written by the previous stage): a path that ended keeps its last pending next-event term in C / vis, which k_ns_resolve
// adds — the same sum, in the same order, as visiting the dead pixel here would give
#ifndef SNK_NS_VERTEX_CTAS
#define SNK_NS_VERTEX_CTAS 8  // 5 / 6 / 8 resident CTAs (96 / 80 / 64 registers, 52 B of spills at 64): vertex 1 0.215 / 0.189 / 0.170 ms on C2, 0.259 / 0.231 / 0.204 ms on C3
                              // (profiles/r07l_variants_shading_occupancy.txt): the stage waits on dependent gathers (issue slots 38 % busy), more warps hide more
#endif
__global__ void __launch_bounds__(128, SNK_NS_VERTEX_CTAS) k_ns_vertex(SceneDev D, FrameConsts fc, NsBuffers B, uint32_t npix, uint32_t b, uint32_t max_bounces,
                                                   uint32_t nee_from, unsigned long long* counters, uint32_t* q_refs, uint32_t* q_count,
                                                   const uint32_t* paths_in, const uint32_t* paths_in_count, uint32_t* paths_out,
                                                   uint32_t* paths_out_count) {
  // grid-stride over the compact list: the host sizes the grid for the paths it EXPECTS to be alive at bounce b (it never reads
  // the count back); 16 200 CTAs for the 100-300 k paths of the late bounces spent a third of the stage launching empty CTAs
  uint32_t count = __ldcg(paths_in_count);
  if (count > npix) count = npix;
  unsigned long long n_shadow = 0, n_closest = 0;
  // Two phases per batch of 128 paths. Phase 1, one thread per path: add the previous vertex's resolved next-event term and look
  // at the hit record; rays that left the scene or were dead end here. Phase 2 runs the surface work (vertex fetch, emission MIS,
  // next-event sample, BSDF sample) on the paths that DID hit, compacted through shared memory so that whole warps work: in an open
  // scene 40-50 % of the rays of a bounce miss, and with one thread per listed path the stage ran at 13-17 of 32 lanes (ncu r04r).
  // Per path the same operations in the same order as before; only the order of the queue entries changes (results do not depend on it).
  __shared__ uint32_t sm_list[128];
  __shared__ uint32_t sm_warp_n[4];
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  for (uint32_t base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x) {  // CTA-uniform trip count
  const uint32_t j = base + threadIdx.x;
  bool hit = false;
  uint32_t i = 0;
  if (j < count) {
    i = paths_in[j];
    // the previous vertex's next-event estimate, now that its shadow ray is resolved
    float4 L4 = B.L[i];
    vec3 L = v3(L4.x, L4.y, L4.z);
    if (b >= 1u + nee_from) {
      const float4 c = B.C[i];
      L += v3(c.x, c.y, c.z) * B.vis[i];
    }
    const float r1w = B.rays[2 * i + 1].w;
    hit = r1w >= 0.0f && !(B.hit[i].x < 0.0f);
    B.L[i] = make_float4(L.x, L.y, L.z, 0.0f);
    if (!hit) {
      if (r1w >= 0.0f) put_dead_dir(B.rays, i);
      B.C[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
  if (lane == 0) sm_warp_n[wid] = __popc(m);
  __syncthreads();
  uint32_t before = 0, n_hit = 0;
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) { const uint32_t c = sm_warp_n[k]; before += k < wid ? c : 0u; n_hit += c; }
  if (hit) sm_list[before + __popc(m & ((1u << lane) - 1u))] = i;
  __syncthreads();
  bool has_shadow = false, has_closest = false;
  i = 0;
  if (threadIdx.x < n_hit) {
    i = sm_list[threadIdx.x];
    const uint32_t x = i % fc.W, y = fc.row0 + i / fc.W;
    const uint32_t pixel = y * fc.W + x;
    const float4 r1 = B.rays[2 * i + 1];
    const float4 h = B.hit[i];
    const vec3 wi = v3(r1.x, r1.y, r1.z);
    const float4 T4 = B.T[i];
    const vec3 T = v3(T4.x, T4.y, T4.z);
    const NsSurf s = NS_SURFACE(D, B.inst[i], __float_as_uint(h.w), h.y, h.z);
    if (s.flags & SNKPT_MAT_FLAG_EMISSIVE) {
      if (b >= 1u + nee_from) {  // the ray left x_{b-1}, a vertex that also did next-event estimation
        const NsLightCounts lc = ns_light_counts(D);
        const float pdf_b = T4.w;
        const float p_nee = ns_nee_pdf_emissive(lc, s.ntri, s.area, h.x, s.n, wi);
        const float w = (pdf_b * pdf_b) / (pdf_b * pdf_b + p_nee * p_nee);
        const float4 L4 = B.L[i];
        const vec3 L = v3(L4.x, L4.y, L4.z) + (T * s.albedo) * w;
        B.L[i] = make_float4(L.x, L.y, L.z, 0.0f);
      }
      put_dead_dir(B.rays, i);
      B.C[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      ns_vertex_work(D, B, i, pixel, fc.frame_idx, b, max_bounces, nee_from, s, -wi, T, has_shadow, has_closest);
    }
  }
  __syncthreads();  // sm_list / sm_warp_n are rewritten by the next batch
  q_append3(q_refs, q_count, has_closest, make_ref(SEG_NS_CLOSEST, i), has_shadow, make_ref(SEG_NS_SHADOW, i), false, 0u);
  q_append1(paths_out, paths_out_count, has_closest, i);
  n_shadow += has_shadow ? 1ull : 0ull; n_closest += has_closest ? 1ull : 0ull;
  }
  flush_counts(n_shadow, n_closest, counters);
}

__global__ void __launch_bounds__(256) k_ns_resolve(SceneDev D, FrameConsts fc, PixelPlanes P, NsBuffers B, uint32_t npix, uint32_t max_bounces,
                                                    uint32_t nee_from) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const uint32_t x = i % fc.W, y = fc.row0 + i / fc.W;
  const size_t px = (size_t)y * fc.W + x;
  vec3 out = v3(0.0f);
  if (P.prim[px] != INVALID_IDX) {
    // an emissive primary hit shows its emission = the material albedo (what ns_surface returns for it)
    const SnkMaterial* m = D.materials + __ldg(&D.instances[P.inst[px]].material_idx);
    if (__ldg(&m->flags) & SNKPT_MAT_FLAG_EMISSIVE) {
      const float4 alb = __ldg(reinterpret_cast<const float4*>(m->albedo));
      out = v3(alb.x, alb.y, alb.z);
    } else {
      const float4 L4 = B.L[i];
      vec3 L = v3(L4.x, L4.y, L4.z);
      if (max_bounces >= nee_from) {
        const float4 c = B.C[i];
        L += v3(c.x, c.y, c.z) * B.vis[i];
      }
      const float4 d = P.direct[px];
      out = v3(d.x, d.y, d.z) + L;
    }
  }
  P.radiance[px] = make_float4(out.x, out.y, out.z, 1.0f);
  P.out16[px] = make_ushort4(f2h(out.x), f2h(out.y), f2h(out.z), f2h(1.0f));
}

__global__ void k_fold_counters_ns(unsigned long long* c) {
  c[4] += c[1];
  c[5] += c[2];
}

}  // namespace

int launch_northstar_frame(SnkPt* ctx, const SnkRtSettings& s, cudaStream_t st) {
  const FrameConsts& fc = ctx->fc;
  const uint32_t npix = fc.W * (fc.row1 - fc.row0), npix_frame = fc.W * fc.H;
  if (s.max_bounces > 16) { ctx->err = "snkpt_render: max_bounces > 16"; return SNKPT_E_INVALID; }
  if (npix_frame > REF_SLOT_MASK) { ctx->err = "snkpt_render: more than 2^28 pixels"; return SNKPT_E_INVALID; }
  int launches = 0;
  cudaMemsetAsync(ctx->dev_counters.as<uint32_t>() + 2, 0, 16, st);
  // direct term at the primary hit: the ReSTIR DI stages (restir_ns.cu); switched off, the path does its own
  // next-event estimation at x0 instead (plain path tracer, the unbiased yardstick of the statistical tests)
  const uint32_t nee_from = (s.flags & SNKPT_RT_NO_RESTIR) ? 0u : 1u;
  const size_t sizes[8] = {32ull * npix, 16ull * npix, 4ull * npix, 16ull * npix, 16ull * npix, 32ull * npix, 16ull * npix, 4ull * npix};
  for (int k = 0; k < 8; ++k) {
    cudaError_t e = ctx->wave[k].reserve(sizes[k], false, st);
    if (e != cudaSuccess) { ctx->err = std::string("north-star queues: ") + cudaGetErrorString(e); return SNKPT_E_NOMEM; }
  }
  // one compact reference queue per traversal launch; the largest (launch 2) holds 3 ReSTIR + 2 path rays per pixel
  const uint32_t qcap = 5u * npix_frame;
  int rc = reserve_ref_queues(ctx, qcap, st);
  if (rc < 0) return rc;
  // compact lists of the paths alive after each vertex stage (ping-pong) + one counter per stage
  if (ctx->wave_paths.reserve(8ull * npix, false, st) != cudaSuccess) { ctx->err = "north-star path lists: out of memory"; return SNKPT_E_NOMEM; }
  uint32_t* paths[2] = {ctx->wave_paths.as<uint32_t>(), ctx->wave_paths.as<uint32_t>() + npix};
  uint32_t* path_count = ctx->dev_counters.as<uint32_t>() + 112;
  reset_ref_queues(ctx, st);
  NsBuffers B;
  B.rays = ctx->wave[0].as<float4>(); B.hit = ctx->wave[1].as<float4>(); B.inst = ctx->wave[2].as<uint32_t>(); B.T = ctx->wave[3].as<float4>();
  B.L = ctx->wave[4].as<float4>(); B.srays = ctx->wave[5].as<float4>(); B.C = ctx->wave[6].as<float4>(); B.vis = ctx->wave[7].as<float>();
  TraceSegs segs{};
  if (nee_from) {
    rc = restir_ns_prepare(ctx, segs, st);
    if (rc < 0) return rc;
  }
  segs.s[SEG_NS_CLOSEST] = TraceSeg{B.rays, B.hit, B.inst, nullptr, 0u, 0u};
  segs.s[SEG_NS_SHADOW] = TraceSeg{B.srays, nullptr, nullptr, B.vis, 1u, 0u};
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(ctx->dev_counters.as<uint32_t>() + 2);
  const SceneDev D = ctx->scene_dev();
  const uint32_t g128 = (npix + 127) / 128;
  const bool bounce = !(s.flags & SNKPT_RT_NO_BOUNCE);
  const uint32_t max_bounces = bounce ? s.max_bounces : 0u;
  // Two independent chains share the frame after the G-buffer: the ReSTIR DI stages (direct light at x0) and the path
  // (vertices 1..max_bounces). Each chain appends its rays to its own queues and every traversal launch drains ONE queue:
  //   chain A (ReSTIR): rs_initial -> [queue A0: initial-candidate shadow rays] -> temporal, spatial, final -> [queue A1: final
  //                     shadow rays] -> rs_resolve
  //   chain B (path):   ns_gen -> [queue B0: first continuation ray (+ its NEE ray without ReSTIR)] -> vertex b ->
  //                     [queue Bb: NEE ray of vertex b + continuation ray b+1] ...
  // With overlap on (default outside profiling) chain A runs on the context's low-priority auxiliary stream: the late
  // bounces of chain B are latency-bound tails (0.05-0.65 M live rays, 18-42 % achieved occupancy, profiles/r03h) and the
  // traversal kernels leave two CTA slots per SM free (launch grid 6 of 8 resident CTAs), so the bulk kernels of chain A
  // fill the SMs the tails leave idle. Results do not depend on which queue or launch a ray went through.
  const bool overlap = nee_from && ns_overlap_enabled(ctx);
  cudaStream_t sa = st;
  if (overlap) {
    if (!ctx->aux_stream) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if (cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, lo) != cudaSuccess) { ctx->err = "north-star frame: cannot create the auxiliary stream"; return SNKPT_E_CUDA; }
      if (cudaStreamCreateWithPriority(&ctx->hi_stream, cudaStreamNonBlocking, hi) != cudaSuccess) { ctx->err = "north-star frame: cannot create the path stream"; return SNKPT_E_CUDA; }
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ctx->ev_join_b, cudaEventDisableTiming);
    }
    sa = ctx->aux_stream;
  }
  // the path chain is the frame's critical path: with overlap on it runs on a high-priority stream of the context, so its
  // (small, latency-bound) launches get the CTA slots the bulk kernels of the ReSTIR chain free first
  cudaStream_t sb = overlap ? ctx->hi_stream : st;
  cudaEventRecord(ctx->ev[2], st);
  launches += build_emissive_table(ctx, st);
  char name[32];
  int rec;
  // queue capacities: chain A up to 3 rays per pixel (final point / emissive / directional), chain B 2 (NEE + continuation)
  const uint32_t capA = 3u * npix_frame, capB = 2u * npix_frame;
  // Staggered start (SNKPT_NS_STAGGER = "a,b", path events: 0 = after ns_gen, k = after the traversal launch of bounce k - 1):
  // the ReSTIR chain's first half (initial candidates + their shadow rays) starts when path event a has passed, its second half
  // (temporal / spatial / final + shadow rays + resolve) when event b has passed; -1 = at the fork. The path chain is enqueued first then.
  static int stagger[2] = {-2, -2};
  if (stagger[0] == -2) {
    int a0 = -1, a1 = -1;
    const char* e = getenv("SNKPT_NS_STAGGER");
    if (e) sscanf(e, "%d,%d", &a0, &a1);
    stagger[1] = a1 < 18 ? a1 : 18; stagger[0] = a0 < 18 ? a0 : 18;
  }
  const bool staggered = overlap && (stagger[0] >= 0 || stagger[1] >= 0);
  auto stage_event = [&](int k) {
    if (!staggered || k >= 19) return;
    if (!ctx->ev_path[k]) cudaEventCreateWithFlags(&ctx->ev_path[k], cudaEventDisableTiming);
    cudaEventRecord(ctx->ev_path[k], sb);
  };
  int rcA = 0;
  auto chain_a = [&]() {
    if (!nee_from) {
      cudaEventRecord(ctx->ev[3], st);
      cudaMemsetAsync(ctx->planes.direct + (size_t)fc.row0 * fc.W, 0, 16ull * npix, st);
      return;
    }
    const int last = (int)max_bounces;  // events 0 .. max_bounces exist
    if (staggered && stagger[0] >= 0) cudaStreamWaitEvent(sa, ctx->ev_path[stagger[0] < last ? stagger[0] : last], 0);
    RefQueue QA = ref_queue(ctx, 0, capA, 0);
    launches += restir_ns_initial(ctx, s, QA, sa);
    launches += traced_launch_q(ctx, "trace.qA0", segs, QA, sa, overlap ? 1 : 0);
    if (staggered && stagger[1] >= 0) cudaStreamWaitEvent(sa, ctx->ev_path[stagger[1] < last ? stagger[1] : last], 0);
    QA = ref_queue(ctx, 1, capA, 0);
    const int r = restir_ns_reuse(ctx, s, QA, sa, overlap ? 1 : 0);
    if (r < 0) { rcA = r; return; }
    launches += r;
    launches += traced_launch_q(ctx, "trace.qA1", segs, QA, sa, overlap ? 1 : 0);
    launches += restir_ns_resolve(ctx, sa);
    if (overlap) cudaEventRecord(ctx->ev_join, sa);
  };
  if (!nee_from) ctx->prev_free = 1;  // without the ReSTIR stages nothing of this frame reads the previous normals / depths
  if (nee_from && overlap) { cudaEventRecord(ctx->ev_fork, st); cudaStreamWaitEvent(sa, ctx->ev_fork, 0); cudaStreamWaitEvent(sb, ctx->ev_fork, 0); }
  if (!staggered) { chain_a(); if (rcA < 0) return rcA; }
  int q = 2;
  RefQueue Q = ref_queue(ctx, q, capB, capA);
  rec = prof_begin(ctx, "ns_gen", npix, false, false, sb);
  k_ns_gen<<<(stage_threads(fc.W, fc.row1 - fc.row0) + 127) / 128, 128, 0, sb>>>(D, fc, ctx->planes, B, npix, max_bounces, nee_from, counters, Q.refs, Q.count, paths[0], path_count);
  prof_end(ctx, rec, sb);
  ++launches;
  stage_event(0);
  launches += traced_launch_q(ctx, "trace.qB0", segs, Q, sb, 0);
  stage_event(1);
  const uint32_t sms = (uint32_t)device_sm_count();
  for (uint32_t b = 1; b <= max_bounces; ++b) {
    Q = ref_queue(ctx, ++q, capB, capA);
    snprintf(name, sizeof(name), "ns_vertex.%u", b);
    rec = prof_begin(ctx, name, npix, false, false, sb);
    // expected survivors: everything at bounce 1, then a shrinking multiple of the resident grid (6 CTAs per SM)
    const uint32_t g_cap = sms * 6u * (b == 1 ? 32u : b == 2 ? 8u : b == 3 ? 4u : 2u);
    k_ns_vertex<<<g128 < g_cap ? g128 : g_cap, 128, 0, sb>>>(D, fc, B, npix, b, max_bounces, nee_from, counters, Q.refs, Q.count, paths[(b - 1) & 1], path_count + (b - 1),
                                      paths[b & 1], path_count + b);
    prof_end(ctx, rec, sb);
    snprintf(name, sizeof(name), "trace.qB%u", b);
    launches += 1 + traced_launch_q(ctx, name, segs, Q, sb, 0);
    if (b < max_bounces) stage_event((int)b + 1);
    if (b == 1) cudaEventRecord(ctx->ev[4], sb);
  }
  if (staggered) { chain_a(); if (rcA < 0) return rcA; }
  if (max_bounces == 0) cudaEventRecord(ctx->ev[4], sb);
  if (overlap) { cudaEventRecord(ctx->ev_join_b, sb); cudaStreamWaitEvent(st, ctx->ev_join_b, 0); }
  if (overlap) cudaStreamWaitEvent(st, ctx->ev_join, 0);
  wait_readback(ctx, st);  // the resolve overwrites the output image
  rec = prof_begin(ctx, "ns_resolve", npix, false, false, st);
  k_ns_resolve<<<(npix + 255) / 256, 256, 0, st>>>(D, fc, ctx->planes, B, npix, max_bounces, nee_from);
  prof_end(ctx, rec, st);
  k_fold_counters_ns<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(ctx->dev_counters.as<uint32_t>()));
  launches += 2;
  cudaEventRecord(ctx->ev[5], st);
  ctx->stats.kernel_launches += launches;
  ctx->render_pending = true;
  return launches;
}

}  // namespace snk
