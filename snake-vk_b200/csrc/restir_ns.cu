// restir_ns.cu — north-star ReSTIR DI (BASELINE.json north_star item 4) as separate wavefront stages:
//
//   k_rs_initial   candidate generation (RIS over point lights / emissive triangles) -> reservoir (half 1) + shadow ray
//   [trace any]    persistent wavefront traversal kernel (trace.cu), one launch per light type present
//   k_rs_temporal  visibility of the initial sample, temporal reuse at the motion-vector-reprojected pixel -> half 1
//   k_rs_spatial   spatial reuse with the bias-correcting MIS weight -> half 0 (next frame's history);
//   k_rs_final     final contributions of that reservoir on the exact primary-hit surface + 3 shadow rays (point, emissive,
//                  directional)
//   [trace any]
//   k_rs_resolve   direct = sum of contribution x visibility -> planes.direct
//
// Data contract kept from the reference: the 48-B PixelReservoir (two independent reservoirs per pixel) staged as
// three 16-B SoA planes per buffer half (RayGen.rgen:53-76, EditorLayer.cpp:299-302), RT::RtSettings and the shader
// constants (Raytracing.h:13-16, RayGen.rgen:95-97, :435, :488, :491). The estimator itself is restated in
// oracle/snk_restir_ns.h (which lists how it differs from the reference's biased passes); this file must match that
// statement bit for bit: same expression order, -fmad=false, PCG-hash random numbers, no trigonometry.
// No shadow ray is traced inside a shading kernel: every ray goes through the wavefront traversal kernel.
#include <cstdlib>

#include "kernels.h"
#include "northstar.cuh"
#include "packet.cuh"

namespace snk {

namespace {

constexpr uint32_t RS_DIM_INITIAL = 0x1000u, RS_DIM_TEMPORAL = 0x2000u, RS_DIM_SPATIAL = 0x3000u;
__device__ __forceinline__ uint32_t rs_seed(uint32_t pixel, uint32_t frame) { return pcg_hash(pcg_hash(pixel) + frame); }
__device__ __forceinline__ float rs_rand(uint32_t h0, uint32_t dim) { return (float)(pcg_hash(h0 + dim) >> 8) * (1.0f / 16777216.0f); }

struct RsParams {
  int32_t num_temporal, num_spatial;
  uint32_t ris_iters, m_cap;
  float radius, normal_threshold;
  uint32_t history_ok;
  uint32_t npix_frame;  // W * H: plane stride of the scratch queues
};
// scratch queues, indexed by full-frame pixel (plane k of pixel px at k * npix_frame + px)
struct RsBuffers {
  float4* rays;   // 3 planes x 2 float4
  float* vis;     // 3 planes
  float2* phat;   // p^ of the initial samples
  float4* C;      // 3 planes: final unshadowed contributions
};

struct RsSurf {
  vec3 n, wo, o, albedo;
  float roughness, metallic;
};
struct RsEval { vec3 C, wi; float tmax; };
__device__ __forceinline__ float rs_phat(const RsEval& e) { return e.tmax > 0.0f ? (e.C.x + e.C.y) + e.C.z : 0.0f; }

__device__ __forceinline__ bool px_shaded(const PixelPlanes& P, size_t px) {
  return P.prim[px] != INVALID_IDX && !(P.flags[px] & SNKPT_MAT_FLAG_EMISSIVE);
}
__device__ __forceinline__ vec3 px_normal(const ushort4* plane, size_t px) {
  const ushort4 n = plane[px];
  return v3(h2f(n.x), h2f(n.y), h2f(n.z));
}
// G-buffer surface of a shaded pixel (oracle: rs_gsurf)
__device__ __forceinline__ void rs_gsurf(const FrameConsts& fc, const PixelPlanes& P, uint32_t x, uint32_t y, RsSurf& s, vec3& pos, vec3& n_raw) {
  const size_t px = (size_t)y * fc.W + x;
  n_raw = px_normal(P.normal, px);
  const float u = ((float)x + 0.5f) / (float)fc.W, v = ((float)y + 0.5f) / (float)fc.H;
  pos = world_pos_from_depth(fc.ivp, P.depth[px], u, v);
  const char4 a = P.albedo[px];
  const ushort4 r = P.rma[px];
  s.albedo = v3(snorm8_2f(a.x), snorm8_2f(a.y), snorm8_2f(a.z));
  s.roughness = gmax(h2f(r.x), 0.05f);
  s.metallic = h2f(r.y);
  const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
  s.wo = normalize(cam - pos);
  s.n = n_raw;
  if (dot(s.n, s.wo) < 0.0f) s.n = -s.n;
  s.o = pos + s.n * 0.01f;
}
// the same surface from a packed record (a neighbour's pixel: one sector instead of five planes)
__device__ __forceinline__ void rs_gsurf_rec(const FrameConsts& fc, const SurfRec& r, uint32_t x, uint32_t y, RsSurf& s, vec3& pos, vec3& n_raw) {
  n_raw = surf_normal(r);
  const float u = ((float)x + 0.5f) / (float)fc.W, v = ((float)y + 0.5f) / (float)fc.H;
  pos = world_pos_from_depth(fc.ivp, __uint_as_float(r.a.w), u, v);
  char4 a;
  memcpy(&a, &r.b.x, 4);
  s.albedo = v3(snorm8_2f(a.x), snorm8_2f(a.y), snorm8_2f(a.z));
  s.roughness = gmax(h2f((uint16_t)(r.b.y & 0xFFFFu)), 0.05f);
  s.metallic = h2f((uint16_t)(r.b.y >> 16));
  const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
  s.wo = normalize(cam - pos);
  s.n = n_raw;
  if (dot(s.n, s.wo) < 0.0f) s.n = -s.n;
  s.o = pos + s.n * 0.01f;
}
__device__ __forceinline__ RsSurf rs_xsurf(const NsSurf& h, vec3 cam) {
  RsSurf s;
  s.albedo = h.albedo; s.roughness = h.roughness; s.metallic = h.metallic;
  s.wo = normalize(cam - h.p);
  s.n = h.n;
  if (dot(s.n, s.wo) < 0.0f) s.n = -s.n;
  s.o = h.p + s.n * 0.01f;
  return s;
}
// `INL`: the candidate loops of k_rs_initial inline the evaluation (SNK_RS_INLINE_EVAL; 16-32 calls per pixel, arguments and result
// through local memory otherwise); the reuse stages keep real calls (many call sites: instruction-cache pressure, r01c)
template <bool INL>
__device__ __forceinline__ RsEval rs_eval_pl_body(const RsSurf& s, const PointLightRec& pl) {
  RsEval e; e.C = v3(0.0f); e.wi = v3(0.0f, 1.0f, 0.0f); e.tmax = -1.0f;
  const vec3 d = pl.pos - s.o;
  const float dist = length(d);
  if (!(dist > 0.002f)) return e;
  const vec3 wi = inv_scale(d, dist);
  if (!(dot(s.n, wi) > 0.0f)) return e;
  const float atten = (pl.atten_constant + pl.atten_linear * dist) + pl.atten_exp * (dist * dist);
  e.C = inv_scale((INL ? ns_bsdf_cos_inl(s.n, s.wo, wi, s.albedo, s.roughness, s.metallic) : ns_bsdf_cos(s.n, s.wo, wi, s.albedo, s.roughness, s.metallic)) * pl.colour, atten);
  e.wi = wi; e.tmax = dist - 0.001f;
  return e;
}
template <bool INL>
__device__ __forceinline__ RsEval rs_eval_em_body(const RsSurf& s, const NsSurf& q) {
  RsEval e; e.C = v3(0.0f); e.wi = v3(0.0f, 1.0f, 0.0f); e.tmax = -1.0f;
  const vec3 d = q.p - s.o;
  const float dist = length(d);
  if (!(dist > 0.01f)) return e;
  const vec3 wi = inv_scale(d, dist);
  const float cos_l = fabsf(dot(q.n, wi));
  if (!(dot(s.n, wi) > 0.0f) || !(cos_l > 0.000001f)) return e;
  e.C = inv_scale(((INL ? ns_bsdf_cos_inl(s.n, s.wo, wi, s.albedo, s.roughness, s.metallic) : ns_bsdf_cos(s.n, s.wo, wi, s.albedo, s.roughness, s.metallic)) * q.albedo) * cos_l, dist * dist);
  e.wi = wi; e.tmax = dist - 0.005f;
  return e;
}
static __device__ __noinline__ RsEval rs_eval_pl(const RsSurf& s, const PointLightRec& pl) { return rs_eval_pl_body<false>(s, pl); }
static __device__ __noinline__ RsEval rs_eval_em(const RsSurf& s, const NsSurf& q) { return rs_eval_em_body<false>(s, q); }
#ifndef SNK_RS_INLINE_EVAL
#define SNK_RS_INLINE_EVAL 1
#endif
__device__ __forceinline__ RsEval rs_eval_pl_initial(const RsSurf& s, const PointLightRec& pl) {
#if SNK_RS_INLINE_EVAL
  return rs_eval_pl_body<true>(s, pl);
#else
  return rs_eval_pl(s, pl);
#endif
}
#ifndef SNK_RS_INLINE_REUSE
#define SNK_RS_INLINE_REUSE 1  // r06k: rs_spatial 0.384 -> 0.306 ms (C2), 0.672 -> 0.515 ms (C3)
#endif
__device__ __forceinline__ RsEval rs_eval_pl_reuse(const RsSurf& s, const PointLightRec& pl) {
#if SNK_RS_INLINE_REUSE
  return rs_eval_pl_body<true>(s, pl);
#else
  return rs_eval_pl(s, pl);
#endif
}
__device__ __forceinline__ RsEval rs_eval_em_reuse(const RsSurf& s, const NsSurf& q) {
#if SNK_RS_INLINE_REUSE
  return rs_eval_em_body<true>(s, q);
#else
  return rs_eval_em(s, q);
#endif
}
__device__ __forceinline__ RsEval rs_eval_em_initial(const RsSurf& s, const NsSurf& q) {
#if SNK_RS_INLINE_EVAL
  return rs_eval_em_body<true>(s, q);
#else
  return rs_eval_em(s, q);
#endif
}
// a stored emissive sample is usable when its instance is (still) in the emissive list — snkpt_set_instances drops the
// reservoir history when that list changes — and the triangle exists
__device__ __forceinline__ bool rs_em_valid(const SceneDev& D, uint32_t obj, uint32_t tri) {
  return obj < D.num_instances && __ldg(D.em_base_of_obj + obj) != INVALID_IDX && tri < __ldg(&D.instances[obj].num_mesh_indices) / 3u;
}
__device__ __forceinline__ void rs_empty(Reservoir& r) { init_reservoir(r); }

// thread -> pixel (northstar.cuh:stage_pixel); false outside the band
__device__ __forceinline__ bool pixel_of(const FrameConsts& fc, uint32_t t, uint32_t& x, uint32_t& y, size_t& px) {
  const bool ok = stage_pixel(fc, t, x, y);
  px = ok ? (size_t)y * fc.W + x : 0;
  return ok;
}

// ---- emissive-triangle table (northstar.cuh): one block per emissive-list entry ------------------------------------------
__global__ void __launch_bounds__(128) k_build_em_table(SceneDev D, float4* __restrict__ table) {
  const uint32_t k = blockIdx.x;
  const uint32_t obj = __ldg(D.emissive_idx + k);
  const uint32_t base = __ldg(D.em_base + k);
  const SnkInstanceData in = load_instance(D, obj);
  const uint32_t ntri = in.num_mesh_indices / 3u;
  const float* M = D.transforms + 16ull * in.transform_idx;
  const float* A = D.normal_mats + 9ull * in.transform_idx;
  const SnkMaterial* m = D.materials + in.material_idx;
  const float4 alb = __ldg(reinterpret_cast<const float4*>(m->albedo));
  const uint32_t flags = __ldg(&m->flags);
  for (uint32_t tri = threadIdx.x; tri < ntri; tri += blockDim.x) {
    uint32_t ia, ib, ic;
    tri_indices(D, in, tri, ia, ib, ic);
    // world-space vertices and vertex normals (oracle: ns_light_point)
    const vec3 wa = mul_point(M, ld3(D.positions, ia)), wb = mul_point(M, ld3(D.positions, ib)), wc = mul_point(M, ld3(D.positions, ic));
    const vec3 na = mul_mat3r(A, ld3(D.normals, ia)), nb = mul_mat3r(A, ld3(D.normals, ib)), nc = mul_mat3r(A, ld3(D.normals, ic));
    const float area = 0.5f * length(cross(wb - wa, wc - wa));
    float4* r = table + (size_t)EM_REC_F4 * (base + tri);
    r[0] = make_float4(wa.x, wa.y, wa.z, area);
    r[1] = make_float4(wb.x, wb.y, wb.z, alb.x);
    r[2] = make_float4(wc.x, wc.y, wc.z, alb.y);
    r[3] = make_float4(na.x, na.y, na.z, alb.z);
    r[4] = make_float4(nb.x, nb.y, nb.z, __uint_as_float(ntri));
    r[5] = make_float4(nc.x, nc.y, nc.z, __uint_as_float(flags));
  }
}

// ---- stage 1 ---------------------------------------------------------------------------------------------------------
#ifndef SNK_RS_INITIAL_CTAS
#define SNK_RS_INITIAL_CTAS 6  // 5 / 6 / 8: 0.279 / 0.254 / 0.236 ms on C2 but 1.335 / 1.272 / 1.410 ms on C3 (10 k emissive triangles; r07l): 6 stays
#endif
// HAS_EM = false: the build for scenes without emissive triangles — the emissive candidate loop is compiled out and the kernel runs at
// SNK_RS_INITIAL_CTAS_PL resident CTAs (the point-light loop alone gains from more warps, the emissive loop loses: r07l)
#ifndef SNK_RS_INITIAL_CTAS_PL
#define SNK_RS_INITIAL_CTAS_PL 8
#endif
template <bool HAS_EM>
__global__ void __launch_bounds__(128, HAS_EM ? SNK_RS_INITIAL_CTAS : SNK_RS_INITIAL_CTAS_PL) k_rs_initial(SceneDev D, FrameConsts fc, PixelPlanes P, RsParams rp, RsBuffers B, uint32_t npix,
                                                    unsigned long long* counters, uint32_t* q_refs, uint32_t* q_count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long n_shadow = 0;
  bool ray_pl = false, ray_em = false;
  size_t px = 0;
  uint32_t x, y;
  if (pixel_of(fc, i, x, y, px)) {
    const size_t np = rp.npix_frame;
    Reservoir r; rs_empty(r);
    float p_sel_pl = 0.0f, p_sel_em = 0.0f;
    const uint32_t n_pl = D.lights->num_pointlights, n_em = HAS_EM ? D.num_emissive : 0u;
    if (px_shaded(P, px)) {
      RsSurf s; vec3 pos, n_raw;
      rs_gsurf(fc, P, x, y, s, pos, n_raw);
      const uint32_t h0 = rs_seed((uint32_t)px, fc.frame_idx);
      const uint32_t iters = rp.ris_iters;
      if (n_pl > 0) {
        const bool all = n_pl <= iters;
        const uint32_t m = all ? n_pl : iters;
        RsEval e_sel; e_sel.C = v3(0.0f); e_sel.wi = v3(0.0f, 1.0f, 0.0f); e_sel.tmax = -1.0f;
        for (uint32_t c = 0; c < m; ++c) {
          const uint32_t idx = all ? c : ns_pick(rs_rand(h0, RS_DIM_INITIAL + 2u * c), n_pl);
          const RsEval e = rs_eval_pl_initial(s, load_pointlight(D.lights, idx));
          const float p_hat = rs_phat(e);
          const float w = p_hat * (float)n_pl;
          r.w_sum_pl += w;
          if (w > 0.0f && rs_rand(h0, RS_DIM_INITIAL + 2u * c + 1u) * r.w_sum_pl < w) { r.pl_idx = idx; p_sel_pl = p_hat; e_sel = e; }
        }
        r.pl_m = m;
        if (r.pl_idx != INVALID_IDX) {
          r.w_pl = (r.w_sum_pl / (float)r.pl_m) / p_sel_pl;
          put_ray(B.rays, px, s.o, 0.001f, e_sel.wi, e_sel.tmax);
          ray_pl = true;
          ++n_shadow;
        }
      }
      if (n_em > 0) {
        RsEval e_sel; e_sel.C = v3(0.0f); e_sel.wi = v3(0.0f, 1.0f, 0.0f); e_sel.tmax = -1.0f;
        for (uint32_t c = 0; c < iters; ++c) {
          const uint32_t d = RS_DIM_INITIAL + 0x800u + 5u * c;
          const uint32_t obj = __ldg(D.emissive_idx + ns_pick(rs_rand(h0, d), n_em));
          const uint32_t ntri = __ldg(&D.instances[obj].num_mesh_indices) / 3u;
          const uint32_t tri = ns_pick(rs_rand(h0, d + 1u), ntri);
          float bu = rs_rand(h0, d + 2u), bv = rs_rand(h0, d + 3u);
          if (bu + bv >= 1.0f) { bu = 1.0f - bu; bv = 1.0f - bv; }
          const NsSurf q = SNK_RS_INLINE_EVAL ? ns_surface_em_inl(D, obj, tri, bu, bv) : ns_surface_em(D, obj, tri, bu, bv);
          const RsEval e = rs_eval_em_initial(s, q);
          const float p_hat = rs_phat(e);
          const float w = (p_hat * ((float)n_em * (float)ntri)) * q.area;
          r.w_sum_em += w;
          if (w > 0.0f && rs_rand(h0, d + 4u) * r.w_sum_em < w) { r.obj = obj; r.tri = tri; r.bx = bu; r.by = bv; p_sel_em = p_hat; e_sel = e; }
        }
        r.m_em = iters;
        if (r.obj != INVALID_IDX) {
          r.w_em = (r.w_sum_em / (float)r.m_em) / p_sel_em;
          put_ray(B.rays, np + px, s.o, 0.001f, e_sel.wi, e_sel.tmax);
          ray_em = true;
          ++n_shadow;
        }
      }
    }
    // no shadow ray: the (empty) sample counts as unoccluded
    if (n_pl > 0 && !ray_pl) B.vis[px] = 1.0f;
    if (n_em > 0 && !ray_em) B.vis[np + px] = 1.0f;
    B.phat[px] = make_float2(p_sel_pl, p_sel_em);
    const ResPlanes half1{P.res[1][0], P.res[1][1], P.res[1][2]};
    store_reservoir(half1, px, r);
  }
  q_append3(q_refs, q_count, ray_pl, make_ref(SEG_RS_PL, px), ray_em, make_ref(SEG_RS_EM, px), false, 0u);
  flush_counts(n_shadow, 0ull, counters);
}

// ---- stage 2 ---------------------------------------------------------------------------------------------------------
#ifndef SNK_RS_TEMPORAL_CTAS
#define SNK_RS_TEMPORAL_CTAS 12  // 8 / 10 / 12 resident CTAs (64 / 48 / 40 registers): 0.108 / 0.097 / 0.095 ms on C2, 0.136 / 0.128 / 0.125 ms on C3 (r07m): DRAM-latency bound
#endif
// HAS_EM = false (here and in the spatial / final stages): build for scenes without emissive triangles — the emissive-reservoir work is
// compiled out (the M bookkeeping stays: the reservoirs are bit-identical), which frees registers in stages that wait on gathers
template <bool HAS_EM>
__global__ void __launch_bounds__(128, SNK_RS_TEMPORAL_CTAS) k_rs_temporal(SceneDev D, FrameConsts fc, PixelPlanes P, RsParams rp, RsBuffers B, uint32_t npix) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t x, y; size_t px;
  if (!pixel_of(fc, i, x, y, px)) return;
  if (!px_shaded(P, px)) return;  // stage 1 wrote the empty reservoir
  const size_t np = rp.npix_frame;
  const ResPlanes half0{P.res[0][0], P.res[0][1], P.res[0][2]}, half1{P.res[1][0], P.res[1][1], P.res[1][2]};
  RsSurf s; vec3 pos, n_raw;
  rs_gsurf(fc, P, x, y, s, pos, n_raw);
  Reservoir r = load_reservoir(half1, px);
  const float2 ph0 = B.phat[px];
  float p_pl = ph0.x, p_em = ph0.y;
  const uint32_t n_pl = D.lights->num_pointlights;
  if (n_pl > 0 && B.vis[px] == 0.0f) { r.w_sum_pl = 0.0f; r.w_pl = 0.0f; }
  if (HAS_EM && D.num_emissive > 0 && B.vis[np + px] == 0.0f) { r.w_sum_em = 0.0f; r.w_em = 0.0f; }
  if (rp.history_ok && rp.num_temporal != 0) {
    const ushort2 mv = P.motion[px];
    const float hx = ((float)x + 0.5f) / (float)fc.W + h2f(mv.x), hy = ((float)y + 0.5f) / (float)fc.H + h2f(mv.y);
    if ((hx > 0.0f && hy > 0.0f) && (hx < 1.0f && hy < 1.0f)) {
      uint32_t qx = to_uint(hx * (float)fc.W), qy = to_uint(hy * (float)fc.H);
      if (qx > fc.W - 1) qx = fc.W - 1;
      if (qy > fc.H - 1) qy = fc.H - 1;
      const size_t qp = (size_t)qy * fc.W + qx;
      const vec3 pn = px_normal(P.prev_normal, qp);
      const vec3 ppos = world_pos_from_depth(fc.prev_ivp, P.prev_depth[qp], ((float)qx + 0.5f) / (float)fc.W, ((float)qy + 0.5f) / (float)fc.H);
      const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
      const vec3 dp = ppos - pos;
      const float tol = 0.02f * length(cam - pos) + 0.02f;
      if (dot(pn, n_raw) > rp.normal_threshold && length(dp) < tol) {
        Reservoir h = load_reservoir(half0, qp);
        const uint32_t h0 = rs_seed((uint32_t)px, fc.frame_idx);
        if (h.pl_m > rp.m_cap * r.pl_m) h.pl_m = rp.m_cap * r.pl_m;
        if (h.m_em > rp.m_cap * r.m_em) h.m_em = rp.m_cap * r.m_em;
        if (h.pl_idx < n_pl) {
          const float ph = rs_phat(rs_eval_pl_reuse(s, load_pointlight(D.lights, h.pl_idx)));
          const float w = (ph * h.w_pl) * (float)h.pl_m;
          r.w_sum_pl += w;
          if (w > 0.0f && rs_rand(h0, RS_DIM_TEMPORAL) * r.w_sum_pl < w) { r.pl_idx = h.pl_idx; p_pl = ph; }
        }
        r.pl_m += h.pl_m;
        if (HAS_EM && rs_em_valid(D, h.obj, h.tri)) {
          const NsSurf q = ns_surface_em(D, h.obj, h.tri, h.bx, h.by);
          const float ph = rs_phat(rs_eval_em_reuse(s, q));
          const float w = (ph * h.w_em) * (float)h.m_em;
          r.w_sum_em += w;
          if (w > 0.0f && rs_rand(h0, RS_DIM_TEMPORAL + 1u) * r.w_sum_em < w) { r.obj = h.obj; r.tri = h.tri; r.bx = h.bx; r.by = h.by; p_em = ph; }
        }
        r.m_em += h.m_em;
      }
    }
  }
  r.w_pl = (r.pl_idx != INVALID_IDX && p_pl > 0.0f && r.pl_m > 0) ? (r.w_sum_pl / (float)r.pl_m) / p_pl : 0.0f;
  r.w_em = (r.obj != INVALID_IDX && p_em > 0.0f && r.m_em > 0) ? (r.w_sum_em / (float)r.m_em) / p_em : 0.0f;
  store_reservoir(half1, px, r);
}

// ---- stage 3 ---------------------------------------------------------------------------------------------------------
constexpr int RS_MAXN = 9;  // this pixel + up to 8 neighbours

#ifndef SNK_RS_PREFETCH
#define SNK_RS_PREFETCH 0  // measured (profiles/r02z_variants_rs_prefetch.txt): rs_spatial 0.690 -> 0.811 ms with the prefetches: not adopted
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifndef SNK_RS_SPATIAL_CTAS
#define SNK_RS_SPATIAL_CTAS 8  // 5 / 6 / 8 resident CTAs (96 / 80 / 64 registers): 0.492 / 0.443 / 0.449 ms on C2 (profiles/r04t_variants_occupancy.txt);
                               // after the helpers were inlined (r07l): 5 / 6 / 8 = 0.339 / 0.309 / 0.291 ms on C2, 0.524 / 0.517 / 0.507 ms on C3
#endif
// SNKPT_RT_UNBIASED_REUSE: the shadow rays of the MIS weights are the one place where a north-star shading stage traces inline
// (warp-synchronous tracer of traverse.cuh, as the reference-mode passes do): up to 2 rays per accepted neighbour, each needed
// by the very next addition of the same thread
struct RsTracer {
  const TraceScene* S;
  uint2 *sm_col, *gl_col;
  // true when nothing lies between the surface point and the light sample (oracle: rs_mis_visibility)
  __device__ __forceinline__ bool visible(vec3 o, const RsEval& e) const {
    return trace_warp_any<128>(S, o.x, o.y, o.z, 0.001f, e.wi.x, e.wi.y, e.wi.z, e.tmax, sm_col, gl_col) == 0.0f;
  }
};

template <bool UNBIASED, bool HAS_EM>
__device__ __forceinline__ void rs_spatial_thread(const SceneDev& D, const FrameConsts& fc, const PixelPlanes& P, const RsParams& rp, uint32_t i,
                                                  const RsTracer& T, unsigned long long& n_mis) {
  size_t px = 0;
  uint32_t x, y;
  if (pixel_of(fc, i, x, y, px)) {
    const ResPlanes half0{P.res[0][0], P.res[0][1], P.res[0][2]}, half1{P.res[1][0], P.res[1][1], P.res[1][2]};
    Reservoir out; rs_empty(out);
    const uint32_t n_pl = D.lights->num_pointlights;
    if (px_shaded(P, px)) {
      RsSurf s; vec3 pos, n_raw;
      rs_gsurf(fc, P, x, y, s, pos, n_raw);
      const uint32_t h0 = rs_seed((uint32_t)px, fc.frame_idx);
      uint32_t cpx[RS_MAXN];  // contributing pixels (x | y << 16): slot 0 = this pixel, then the accepted neighbours
      int nc = 0;
      cpx[nc++] = x | (y << 16);
      int ns = rp.num_spatial;
      if (ns > RS_MAXN - 1) ns = RS_MAXN - 1;
#if SNK_RS_PREFETCH
      // the stage is a chain of dependent reads of the neighbours' planes (ncu r02s: issue slots 36 % busy, L1 hit 33 %); the
      // neighbour positions depend only on the counter-based RNG, so their cache lines are requested before anything is consumed
      for (int k = 0; k < ns; ++k) {
        const float ua = rs_rand(h0, RS_DIM_SPATIAL + 2u * (uint32_t)k), ub = rs_rand(h0, RS_DIM_SPATIAL + 2u * (uint32_t)k + 1u);
        const int qx = (int)x + (int)((ua * 2.0f - 1.0f) * rp.radius), qy = (int)y + (int)((ub * 2.0f - 1.0f) * rp.radius);
        if (qx < 0 || qy < 0 || qx >= (int)fc.W || qy >= (int)fc.H) continue;
        const size_t qp = (size_t)qy * fc.W + qx;
        prefetch_l1(P.prim + qp); prefetch_l1(P.flags + qp); prefetch_l1(P.normal + qp); prefetch_l1(P.hit_t + qp);
        prefetch_l1(half1.a + qp); prefetch_l1(half1.b + qp); prefetch_l1(half1.c + qp);
        prefetch_l1(P.depth + qp); prefetch_l1(P.albedo + qp); prefetch_l1(P.rma + qp);
      }
#endif
      const float t_q = P.hit_t[px];
      for (int k = 0; k < ns; ++k) {
        const float ua = rs_rand(h0, RS_DIM_SPATIAL + 2u * (uint32_t)k), ub = rs_rand(h0, RS_DIM_SPATIAL + 2u * (uint32_t)k + 1u);
        const int ox = (int)((ua * 2.0f - 1.0f) * rp.radius), oy = (int)((ub * 2.0f - 1.0f) * rp.radius);
        const int qx = (int)x + ox, qy = (int)y + oy;
        if ((ox == 0 && oy == 0) || qx < 0 || qy < 0 || qx >= (int)fc.W || qy >= (int)fc.H) continue;
        const size_t qp = (size_t)qy * fc.W + qx;
        const SurfRec q = load_surf(P.surf, qp);  // prim / flags / normal / hit distance of the neighbour in one sector
        if (!surf_shaded(q)) continue;
        if (!(dot(surf_normal(q), n_raw) >= rp.normal_threshold)) continue;
        if (!(fabsf(__uint_as_float(q.a.z) - t_q) <= 0.1f * t_q)) continue;
        cpx[nc++] = (uint32_t)qx | ((uint32_t)qy << 16);
      }
      int src_pl = -1, src_em = -1;
      float p_pl = 0.0f, p_em = 0.0f;
      uint32_t m_pl[RS_MAXN], m_em[RS_MAXN];
      for (int j = 0; j < nc; ++j) {
        const Reservoir c = load_reservoir(half1, (size_t)(cpx[j] >> 16) * fc.W + (cpx[j] & 0xFFFFu));
        m_pl[j] = c.pl_m; m_em[j] = c.m_em;
        if (c.pl_idx < n_pl) {
          const float ph = rs_phat(rs_eval_pl_reuse(s, load_pointlight(D.lights, c.pl_idx)));
          const float w = (ph * c.w_pl) * (float)c.pl_m;
          out.w_sum_pl += w;
          if (w > 0.0f && rs_rand(h0, RS_DIM_SPATIAL + 0x100u + 2u * (uint32_t)j) * out.w_sum_pl < w) { out.pl_idx = c.pl_idx; p_pl = ph; src_pl = j; }
        }
        out.pl_m += c.pl_m;
        if (HAS_EM && rs_em_valid(D, c.obj, c.tri)) {
          const NsSurf q = ns_surface_em(D, c.obj, c.tri, c.bx, c.by);
          const float ph = rs_phat(rs_eval_em_reuse(s, q));
          const float w = (ph * c.w_em) * (float)c.m_em;
          out.w_sum_em += w;
          if (w > 0.0f && rs_rand(h0, RS_DIM_SPATIAL + 0x100u + 2u * (uint32_t)j + 1u) * out.w_sum_em < w) {
            out.obj = c.obj; out.tri = c.tri; out.bx = c.bx; out.by = c.by; p_em = ph; src_em = j;
          }
        }
        out.m_em += c.m_em;
      }
      // MIS weight of the selected sample: m_s(y) = p^_s(y) / sum_j p^_j(y) M_j (p^_j on pixel j's G-buffer surface)
      if (src_pl >= 0 || src_em >= 0) {
        float num_pl = 0.0f, den_pl = 0.0f, num_em = 0.0f, den_em = 0.0f;
        NsSurf q;
        if (src_em >= 0) q = ns_surface_em(D, out.obj, out.tri, out.bx, out.by);
        const PointLightRec pl = load_pointlight(D.lights, src_pl >= 0 ? out.pl_idx : 0u);
        for (int j = 0; j < nc; ++j) {
          float ph_pl = p_pl, ph_em = p_em;
          bool vis_pl = true, vis_em = true;
          if (j > 0) {
            RsSurf sj; vec3 pj, nj;
            const uint32_t jx = cpx[j] & 0xFFFFu, jy = cpx[j] >> 16;
            rs_gsurf_rec(fc, load_surf(P.surf, (size_t)jy * fc.W + jx), jx, jy, sj, pj, nj);
            if (src_pl >= 0) {
              const RsEval e = rs_eval_pl_reuse(sj, pl);
              ph_pl = rs_phat(e);
              if (UNBIASED && ph_pl > 0.0f) { vis_pl = T.visible(sj.o, e); ++n_mis; }
            }
            if (src_em >= 0) {
              const RsEval e = rs_eval_em_reuse(sj, q);
              ph_em = rs_phat(e);
              if (UNBIASED && ph_em > 0.0f) { vis_em = T.visible(sj.o, e); ++n_mis; }
            }
          }
          if (src_pl >= 0) {
            const float term = ph_pl * (float)m_pl[j];
            den_pl += vis_pl ? term : 0.0f;
            if (j == src_pl) num_pl = vis_pl ? ph_pl : 0.0f;
          }
          if (src_em >= 0) {
            const float term = ph_em * (float)m_em[j];
            den_em += vis_em ? term : 0.0f;
            if (j == src_em) num_em = vis_em ? ph_em : 0.0f;
          }
        }
        if (src_pl >= 0) out.w_pl = (den_pl > 0.0f && p_pl > 0.0f) ? (out.w_sum_pl * (num_pl / den_pl)) / p_pl : 0.0f;
        if (src_em >= 0) out.w_em = (den_em > 0.0f && p_em > 0.0f) ? (out.w_sum_em * (num_em / den_em)) / p_em : 0.0f;
      }
    }
    store_reservoir(half0, px, out);
  }
}

template <bool HAS_EM>
__global__ void __launch_bounds__(128, SNK_RS_SPATIAL_CTAS) k_rs_spatial(SceneDev D, FrameConsts fc, PixelPlanes P, RsParams rp, uint32_t npix) {
  unsigned long long n_mis = 0;
  rs_spatial_thread<false, HAS_EM>(D, fc, P, rp, blockIdx.x * blockDim.x + threadIdx.x, RsTracer{nullptr, nullptr, nullptr}, n_mis);
}

// unbiased reuse: persistent CTAs (the global spill area of the inline tracer is sized for resident grids, trace.cu:trace_spill_bytes)
__global__ void __launch_bounds__(128, 4) k_rs_spatial_unbiased(TraceScene S, SceneDev D, FrameConsts fc, PixelPlanes P, RsParams rp, uint32_t nthreads,
                                                                unsigned long long* counters) {
  __shared__ uint2 sm_stack[SHORT_STACK * 128];
  __shared__ TraceScene sS;
  if (threadIdx.x == 0) sS = S;
  __syncthreads();
  const RsTracer T{&sS, sm_stack + threadIdx.x, S.spill + ((size_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * (SPILL_STACK * 32) + (threadIdx.x & 31u)};
  unsigned long long n_mis = 0;
  for (uint32_t base = blockIdx.x * 128u; base < nthreads; base += gridDim.x * 128u) rs_spatial_thread<true, true>(D, fc, P, rp, base + threadIdx.x, T, n_mis);
  flush_counts(n_mis, 0ull, counters);
}

#ifndef SNK_RS_FINAL_CTAS
#define SNK_RS_FINAL_CTAS 8  // 6 / 8: 0.112 / 0.105 ms on C2, 0.145 / 0.137 ms on C3 (r07l)
#endif
// final contributions of the reservoir the spatial stage just stored, on the exact primary-hit surface. Its own kernel: together
// with the reuse loops the stage needed 96 registers AND 190 local-memory spill instructions (cuobjdump, r02z).
template <bool HAS_EM>
__global__ void __launch_bounds__(128, SNK_RS_FINAL_CTAS) k_rs_final(SceneDev D, FrameConsts fc, PixelPlanes P, RsParams rp, RsBuffers B, uint32_t npix,
                                                                     unsigned long long* counters, uint32_t* q_refs, uint32_t* q_count, uint32_t dir_packet) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long n_shadow = 0;
  bool ray0 = false, ray1 = false, ray2 = false;
  size_t px = 0;
  uint32_t x, y;
  if (pixel_of(fc, i, x, y, px)) {
    const size_t np = rp.npix_frame;
    const ResPlanes half0{P.res[0][0], P.res[0][1], P.res[0][2]};
    vec3 C0 = v3(0.0f), C1 = v3(0.0f), C2 = v3(0.0f);
    const uint32_t n_pl = D.lights->num_pointlights, n_em = D.num_emissive;
    if (px_shaded(P, px)) {
      const Reservoir out = load_reservoir(half0, px);
      // final contributions on the exact primary-hit surface
      const float2 bc = P.bary[px];
      const NsSurf hs = ns_surface(D, P.inst[px], P.prim[px], bc.x, bc.y);
      const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
      const RsSurf xs = rs_xsurf(hs, cam);
      if (out.pl_idx != INVALID_IDX && out.w_pl > 0.0f) {
        const RsEval e = rs_eval_pl_reuse(xs, load_pointlight(D.lights, out.pl_idx));
        if (e.tmax > 0.0f) {
          C0 = e.C * out.w_pl;
          put_ray(B.rays, px, xs.o, 0.001f, e.wi, e.tmax);
          ray0 = true; ++n_shadow;
        }
      }
      if (HAS_EM && out.obj != INVALID_IDX && out.w_em > 0.0f) {
        const NsSurf q = ns_surface_em(D, out.obj, out.tri, out.bx, out.by);
        const RsEval e = rs_eval_em_reuse(xs, q);
        if (e.tmax > 0.0f) {
          C1 = e.C * out.w_em;
          put_ray(B.rays, np + px, xs.o, 0.001f, e.wi, e.tmax);
          ray1 = true; ++n_shadow;
        }
      }
      {
        const float4 dl = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.dir));
        const float4 dc = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.colour));
        const vec3 wi = v3(dl.x, dl.y, dl.z), col = v3(dc.x, dc.y, dc.z);
        if ((col.x > 0.0f || col.y > 0.0f || col.z > 0.0f) && dot(xs.n, wi) > 0.0f) {
          C2 = ns_bsdf_cos(xs.n, xs.wo, wi, xs.albedo, xs.roughness, xs.metallic) * col;
          put_ray(B.rays, 2 * np + px, xs.o, 0.001f, wi, 10000.0f);
          ray2 = true; ++n_shadow;
        }
      }
    }
    if (n_pl > 0 && !ray0) B.vis[px] = 1.0f;
    if (n_em > 0 && !ray1) B.vis[np + px] = 1.0f;
    if (!ray2) {
      B.vis[2 * np + px] = 1.0f;
      if (dir_packet) put_dead(B.rays, 2 * np + px);  // k_rs_shadow_dir walks every pixel of the band
    }
    B.C[px] = make_float4(C0.x, C0.y, C0.z, 0.0f);
    B.C[np + px] = make_float4(C1.x, C1.y, C1.z, 0.0f);
    B.C[2 * np + px] = make_float4(C2.x, C2.y, C2.z, 0.0f);
  }
  // with dir_packet the directional shadow rays are not queued: k_rs_shadow_dir traces them tile by tile as packets
  q_append3(q_refs, q_count, ray0, make_ref(SEG_RS_PL, px), ray1, make_ref(SEG_RS_EM, px), ray2 && !dir_packet, make_ref(SEG_RS_DIR, px));
  flush_counts(n_shadow, 0ull, counters);
}

// shadow rays towards the directional light: the 32 rays of an 8x4 pixel tile are parallel and start from neighbouring surface
// points — one packet (packet.cuh) instead of 32 entries of the traversal queue. Same visibility bits: the triangle test is the
// queue kernel's, the any-hit answer does not depend on the order or the number of triangles looked at.
#ifndef SNK_SHADOW_PACKET_CTAS
#define SNK_SHADOW_PACKET_CTAS 8
#endif
__global__ void __launch_bounds__(128, SNK_SHADOW_PACKET_CTAS) k_rs_shadow_dir(TraceScene S, FrameConsts fc, const float4* __restrict__ rays, float* __restrict__ vis,
                                                                               uint32_t* ticket, uint32_t* next_ticket) {
  __shared__ uint32_t sm_stack[4 * PK_STACK];
  const uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const PkStack stk{sm_stack + w * PK_STACK, reinterpret_cast<uint32_t*>(S.spill + ((size_t)blockIdx.x * 4 + w) * (SPILL_STACK * 32))};
  const uint32_t tiles_x = (fc.W + 15u) / 16u, n_tiles = 4u * tiles_x * ((fc.row1 - fc.row0 + 7u) / 8u);
  if (blockIdx.x == 0 && threadIdx.x == 0) *next_ticket = 0u;
  for (;;) {
    uint32_t tile = 0;
    if (l == 0) tile = atomicAdd(ticket, 1u);
    tile = __shfl_sync(PK_FULL, tile, 0);
    if (tile >= n_tiles) break;
    const uint32_t q = tile & 3u, blk = tile >> 2;
    const uint32_t x = (blk % tiles_x) * 16 + (q & 1) * 8 + (l & 7);
    const uint32_t y = fc.row0 + (blk / tiles_x) * 8 + (q >> 1) * 4 + (l >> 3);
    const bool in_range = x < fc.W && y < fc.row1;
    const size_t px = in_range ? (size_t)y * fc.W + x : 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 1.f, 0.f, -1.0f);
    if (in_range) { a = rays[2 * px]; b = rays[2 * px + 1]; }
    const bool hit = pk_trace_parallel_any(S, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, stk, l);
    if (in_range && b.w > a.w) vis[px] = hit ? 0.0f : 1.0f;
  }
}

__global__ void __launch_bounds__(256) k_rs_resolve(FrameConsts fc, PixelPlanes P, RsBuffers B, uint32_t npix, uint32_t npix_frame, uint32_t has_pl,
                                                    uint32_t has_em) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const size_t px = (size_t)fc.row0 * fc.W + i, np = npix_frame;
  const float4 c0 = B.C[px], c1 = B.C[np + px], c2 = B.C[2 * np + px];
  const float v0 = has_pl ? B.vis[px] : 1.0f, v1 = has_em ? B.vis[np + px] : 1.0f, v2 = B.vis[2 * np + px];
  P.direct[px] = make_float4((c0.x * v0 + c1.x * v1) + c2.x * v2, (c0.y * v0 + c1.y * v1) + c2.y * v2, (c0.z * v0 + c1.z * v1) + c2.z * v2, 1.0f);
}

}  // namespace

namespace {
struct RsLaunch {
  RsBuffers B;
  FrameConsts f1;  // stages 1-2 also cover the halo rows whose temporal reservoirs this band's spatial reuse reads (snkpt_set_halo)
  uint32_t npix1, npix, npix_frame;
  RsParams rp;
};
RsLaunch rs_launch(SnkPt* ctx, const SnkRtSettings& s) {
  const FrameConsts& fc = ctx->fc;
  RsLaunch L;
  L.npix_frame = fc.W * fc.H;
  L.B.rays = ctx->wave[8].as<float4>(); L.B.vis = ctx->wave[9].as<float>(); L.B.phat = ctx->wave[10].as<float2>(); L.B.C = ctx->wave[11].as<float4>();
  L.f1 = fc;
  L.f1.row0 = fc.row0 > ctx->halo_pass1 ? fc.row0 - ctx->halo_pass1 : 0u;
  L.f1.row1 = fc.row1 + ctx->halo_pass1 < fc.H ? fc.row1 + ctx->halo_pass1 : fc.H;
  L.npix1 = fc.W * (L.f1.row1 - L.f1.row0);
  L.npix = fc.W * (fc.row1 - fc.row0);
  L.rp = RsParams{s.num_temporal_resamples, s.num_spatial_resamples, s.ris_iters, s.m_cap, s.neighbour_radius, s.normal_threshold,
                  (fc.prev_frame_invalid || !ctx->history_valid) ? 0u : 1u, L.npix_frame};
  return L;
}
}  // namespace

// scratch planes + the three shadow-ray segments (slot = full-frame pixel)
int restir_ns_prepare(SnkPt* ctx, TraceSegs& segs, cudaStream_t st) {
  const uint32_t npix_frame = ctx->fc.W * ctx->fc.H;
  const size_t sizes[4] = {96ull * npix_frame, 12ull * npix_frame, 8ull * npix_frame, 48ull * npix_frame};
  for (int k = 0; k < 4; ++k) {
    cudaError_t e = ctx->wave[8 + k].reserve(sizes[k], false, st);
    if (e != cudaSuccess) { ctx->err = std::string("north-star ReSTIR queues: ") + cudaGetErrorString(e); return SNKPT_E_NOMEM; }
  }
  float4* rays = ctx->wave[8].as<float4>();
  float* vis = ctx->wave[9].as<float>();
  for (uint32_t k = 0; k < 3; ++k) segs.s[SEG_RS_PL + k] = TraceSeg{rays + 2ull * k * npix_frame, nullptr, nullptr, vis + (size_t)k * npix_frame, 1u, 0u};
  return SNKPT_OK;
}

// flattens the emissive triangles of this frame (transforms and materials as currently set) into the lookup table
int build_emissive_table(SnkPt* ctx, cudaStream_t st) {
  if (ctx->num_emissive == 0 || ctx->num_emissive_tris == 0) return 0;
  const int rec = prof_begin(ctx, "em_table", ctx->num_emissive_tris, false, false, st);
  k_build_em_table<<<ctx->num_emissive, 128, 0, st>>>(ctx->scene_dev(), ctx->em_table.as<float4>());
  prof_end(ctx, rec, st);
  return 1;
}

// stage 1: candidate generation; appends the shadow rays of the initial samples to q
int restir_ns_initial(SnkPt* ctx, const SnkRtSettings& s, const RefQueue& q, cudaStream_t st) {
  const RsLaunch L = rs_launch(ctx, s);
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(ctx->dev_counters.as<uint32_t>() + 2);
  const int rec = prof_begin(ctx, "rs_initial", L.npix1, false, false, st);
  const SceneDev D = ctx->scene_dev();
  const uint32_t grid = (stage_threads(L.f1.W, L.f1.row1 - L.f1.row0) + 127) / 128;
  if (D.num_emissive) k_rs_initial<true><<<grid, 128, 0, st>>>(D, L.f1, ctx->planes, L.rp, L.B, L.npix1, counters, q.refs, q.count);
  else k_rs_initial<false><<<grid, 128, 0, st>>>(D, L.f1, ctx->planes, L.rp, L.B, L.npix1, counters, q.refs, q.count);
  prof_end(ctx, rec, st);
  return 1;
}

// stages 2-3 (after the initial shadow rays are resolved): temporal reuse, spatial reuse + final evaluation; appends
// the final shadow rays (point, emissive, directional) to q
int restir_ns_reuse(SnkPt* ctx, const SnkRtSettings& s, const RefQueue& q, cudaStream_t st, int spill_region) {
  const RsLaunch L = rs_launch(ctx, s);
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(ctx->dev_counters.as<uint32_t>() + 2);
  const SceneDev D = ctx->scene_dev();
  int rec = prof_begin(ctx, "rs_temporal", L.npix1, false, false, st);
  if (D.num_emissive) k_rs_temporal<true><<<(stage_threads(L.f1.W, L.f1.row1 - L.f1.row0) + 127) / 128, 128, 0, st>>>(D, L.f1, ctx->planes, L.rp, L.B, L.npix1);
  else k_rs_temporal<false><<<(stage_threads(L.f1.W, L.f1.row1 - L.f1.row0) + 127) / 128, 128, 0, st>>>(D, L.f1, ctx->planes, L.rp, L.B, L.npix1);
  // the last reader of the previous frame's normals / depths (the other copy of the G-buffer planes): the next frame's camera rays may overwrite them
  if (ctx->ev_prev_free) { cudaEventRecord(ctx->ev_prev_free, st); ctx->prev_free = 1; ctx->prev_free_event = true; }
  prof_end(ctx, rec, st);
  cudaEventRecord(ctx->ev[3], st);
  // row-band group: the neighbours' temporal reservoirs of the rows spatial reuse can reach arrive here ...
  if (ctx->group) { const int xr = group_exchange(ctx, 0, st); if (xr < 0) return xr; }
  rec = prof_begin(ctx, "rs_spatial", L.npix, false, false, st);
  const uint32_t sp_threads = stage_threads(ctx->fc.W, ctx->fc.row1 - ctx->fc.row0);
  if (s.flags & SNKPT_RT_UNBIASED_REUSE) {
    TraceScene S = ctx->trace_scene();
    S.spill += (size_t)spill_region * (trace_spill_bytes() / sizeof(uint2));
    const uint32_t resident = (uint32_t)device_sm_count() * 4u, want = (sp_threads + 127u) / 128u;
    k_rs_spatial_unbiased<<<want < resident ? want : resident, 128, 0, st>>>(S, D, ctx->fc, ctx->planes, L.rp, sp_threads, counters);
  } else {
    if (D.num_emissive) k_rs_spatial<true><<<(sp_threads + 127) / 128, 128, 0, st>>>(D, ctx->fc, ctx->planes, L.rp, L.npix);
    else k_rs_spatial<false><<<(sp_threads + 127) / 128, 128, 0, st>>>(D, ctx->fc, ctx->planes, L.rp, L.npix);
  }
  prof_end(ctx, rec, st);
  // ... and the final reservoirs the next frame's temporal reuse may reproject to leave as soon as they exist
  if (ctx->group) { const int xr = group_exchange(ctx, 1, st); if (xr < 0) return xr; }
  rec = prof_begin(ctx, "rs_final", L.npix, false, false, st);
  // SNKPT_SHADOW_DIR=packet traces the directional shadow rays as parallel-ray packets (k_rs_shadow_dir) instead of queueing them.
  // Measured and NOT adopted (profiles/r05g_variants_shadow_packet.txt): C2 0.81 M rays cost 0.39 ms as packets against 0.22 ms
  // inside trace.qA1 — the rays of a tile leave from surface points that lie metres apart at grazing view angles, the shared
  // volume is wide, and an any-hit ray is cheap on its own (8 nodes); C4 at 4K: 0.68 vs 0.79 ms, within noise of the frame.
  static const bool dir_packet_env = getenv("SNKPT_SHADOW_DIR") && getenv("SNKPT_SHADOW_DIR")[0] == 'p';
  const bool dir_packet = ctx->has_dir_light && dir_packet_env;
  const uint32_t fin_grid = (stage_threads(ctx->fc.W, ctx->fc.row1 - ctx->fc.row0) + 127) / 128;
  if (D.num_emissive) k_rs_final<true><<<fin_grid, 128, 0, st>>>(D, ctx->fc, ctx->planes, L.rp, L.B, L.npix, counters, q.refs, q.count, dir_packet ? 1u : 0u);
  else k_rs_final<false><<<fin_grid, 128, 0, st>>>(D, ctx->fc, ctx->planes, L.rp, L.B, L.npix, counters, q.refs, q.count, dir_packet ? 1u : 0u);
  prof_end(ctx, rec, st);
  if (dir_packet) {
    rec = prof_begin(ctx, "shadow.dir", L.npix, false, false, st);
    const uint32_t n_tiles = ((ctx->fc.W + 15u) / 16u) * ((ctx->fc.row1 - ctx->fc.row0 + 7u) / 8u);
    const uint32_t resident = (uint32_t)device_sm_count() * SNK_SHADOW_PACKET_CTAS;
    uint32_t* tickets2 = ctx->dev_counters.as<uint32_t>() + 202;  // second self-resetting ticket pair (the first: gbuffer.cu)
    const uint32_t par = ctx->shadow_parity++ & 1u;
    TraceScene Sd = ctx->trace_scene();
    Sd.spill += (size_t)spill_region * (trace_spill_bytes() / sizeof(uint2));  // the chain's own spill region (packet stacks spill there)
    k_rs_shadow_dir<<<n_tiles < resident ? n_tiles : resident, 128, 0, st>>>(Sd, ctx->fc, L.B.rays + 4ull * L.npix_frame, L.B.vis + 2ull * L.npix_frame,
                                                                             tickets2 + par, tickets2 + (par ^ 1u));
    prof_end(ctx, rec, st);
    return 4;
  }
  return 3;
}

// direct = sum of final contribution x visibility
int restir_ns_resolve(SnkPt* ctx, cudaStream_t st) {
  const FrameConsts& fc = ctx->fc;
  const uint32_t npix = fc.W * (fc.row1 - fc.row0), npix_frame = fc.W * fc.H;
  RsBuffers B;
  B.rays = ctx->wave[8].as<float4>(); B.vis = ctx->wave[9].as<float>(); B.phat = ctx->wave[10].as<float2>(); B.C = ctx->wave[11].as<float4>();
  // light counts as of the last snkpt_set_lights / snkpt_set_instances (host mirrors)
  const bool has_pl = ctx->num_pointlights > 0, has_em = ctx->num_emissive > 0;
  const int rec = prof_begin(ctx, "rs_resolve", npix, false, false, st);
  k_rs_resolve<<<(npix + 255) / 256, 256, 0, st>>>(fc, ctx->planes, B, npix, npix_frame, has_pl ? 1u : 0u, has_em ? 1u : 0u);
  prof_end(ctx, rec, st);
  return 1;
}

}  // namespace snk
