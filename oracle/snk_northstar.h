// oracle/snk_northstar.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// CPU statement of the north-star estimator (BASELINE.json north_star: 1 spp, up to 4 bounces, GGX / Lambert
// BSDF sampling, multiple-importance-sampled next-event estimation, Russian roulette) that the CUDA wavefront
// stages of snake-vk_b200/csrc/northstar.cu must reproduce bit for bit. It has NO counterpart in the reference
// (the reference's estimator is the fixed 8x(1+2) tree of RayGen.rgen:139-213); what it shares
// with the reference is the material model (LightBuffers.glsl:65-144 Cook-Torrance terms), the light records, the
// emission convention (emitted radiance = material albedo, RayClosestHit.rchit:119-123) and the direct-light term at the
// primary hit, which is taken from the reference's own ReSTIR DI passes (RayGen.rgen:411-528, restated in snk_raygen.h).
//
//   out(pixel) = Le                                   primary hit emissive
//              = 0                                    sky
//              = direct(pixel) + indirect(pixel)      otherwise
//   direct     = pass-1 / pass-2 result of the reference-mode passes (directional light + ReSTIR reservoirs)
//   indirect   = sum over path vertices x1..xN (N = max_bounces) of throughput x next-event estimate, plus emissive
//                surfaces hit by BSDF sampling from x1.. (MIS, power heuristic); Russian roulette from x2.
//
// Random numbers: PCG hash of (pixel, frame_idx, dimension) — stateless, so every wavefront stage can draw its own.
// Numerics contract as in snk_math.h (IEEE binary32 in the order written, float64 sin/cos rounded once).
#pragma once
#include "snk_raygen.h"

namespace snkor {

inline uint32_t pcg_hash(uint32_t v) {
  const uint32_t state = v * 747796405u + 2891336453u;
  const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}
inline float ns_rand(uint32_t pixel, uint32_t frame, uint32_t dim) {
  const uint32_t h = pcg_hash(pcg_hash(pcg_hash(pixel) + frame) + dim);
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

struct NsSurf {
  vec3 p, n, albedo;
  float roughness, metallic;
  uint32_t flags, ntri;
  float area;  // world-space area of the hit triangle
};

inline NsSurf ns_surface(const Scene& sc, uint32_t inst, uint32_t prim, float u, float v) {
  NsSurf s;
  const float b0 = (1.0f - u) - v;
  const InstanceData& in = sc.instances[inst];
  uint32_t ia, ib, ic; sc.tri_indices(in, prim, ia, ib, ic);
  const vec3 pa = sc.pos(ia), pb = sc.pos(ib), pc = sc.pos(ic);
  const vec3 p = (pa * b0 + pb * u) + pc * v;
  const vec3 n = (sc.nrm(ia) * b0 + sc.nrm(ib) * u) + sc.nrm(ic) * v;
  const mat4& M = sc.transforms[in.transform_idx];
  s.p = mul_point(M, p);
  s.n = normalize(mul(sc.normal_matrix[in.transform_idx], n));
  const Material& m = sc.materials[in.material_idx];
  // emitted radiance is the material albedo alone (as RayGen.rgen:270 and the emissive-triangle sampling); the albedo
  // texture modulates reflecting surfaces only
  s.albedo = (m.flags & MAT_FLAG_EMISSIVE) ? v3(m.albedo[0], m.albedo[1], m.albedo[2]) : sc.albedo_at(m, ia, ib, ic, b0, u, v);
  s.roughness = gmax(m.roughness, 0.05f);
  s.metallic = m.metallic;
  s.flags = m.flags;
  s.ntri = in.num_mesh_indices / 3u;
  const vec3 wa = mul_point(M, pa), wb = mul_point(M, pb), wc = mul_point(M, pc);
  s.area = 0.5f * length(cross(wb - wa, wc - wa));
  return s;
}

// A point on an emissive triangle as a LIGHT SAMPLE (next-event estimation, ReSTIR candidates and stored samples): the triangle's
// vertices and vertex normals are taken to world space first (mul_point / normal matrix per vertex) and interpolated there — what
// the CUDA path reads from its per-frame emissive-triangle table (northstar.cuh: 96 bytes per triangle instead of the instance ->
// indices -> vertices -> matrices -> material chain). Differs from ns_surface (interpolate, then transform) by rounding only.
// Emitted radiance = the material albedo (RayGen.rgen:270); roughness / metallic of an emitter are not used by any caller.
inline NsSurf ns_light_point(const Scene& sc, uint32_t inst, uint32_t prim, float u, float v) {
  NsSurf s;
  const float b0 = (1.0f - u) - v;
  const InstanceData& in = sc.instances[inst];
  uint32_t ia, ib, ic; sc.tri_indices(in, prim, ia, ib, ic);
  const mat4& M = sc.transforms[in.transform_idx];
  const vec3 wa = mul_point(M, sc.pos(ia)), wb = mul_point(M, sc.pos(ib)), wc = mul_point(M, sc.pos(ic));
  const vec3 na = mul(sc.normal_matrix[in.transform_idx], sc.nrm(ia)), nb = mul(sc.normal_matrix[in.transform_idx], sc.nrm(ib)),
             nc = mul(sc.normal_matrix[in.transform_idx], sc.nrm(ic));
  s.p = (wa * b0 + wb * u) + wc * v;
  s.n = normalize((na * b0 + nb * u) + nc * v);
  const Material& m = sc.materials[in.material_idx];
  s.albedo = v3(m.albedo[0], m.albedo[1], m.albedo[2]);
  s.roughness = 0.0f; s.metallic = 0.0f;
  s.flags = m.flags;
  s.ntri = in.num_mesh_indices / 3u;
  s.area = 0.5f * length(cross(wb - wa, wc - wa));
  return s;
}

// f(wi, wo) * cos(theta_i): the Cook-Torrance terms of LightBuffers.glsl:126-144 without the light colour
inline vec3 ns_bsdf_cos(vec3 n, vec3 wo, vec3 wi, vec3 albedo, float roughness, float metallic) {
  const vec3 f0 = mix(v3(0.04f), albedo, metallic);
  const vec3 h = normalize(wo + wi);
  const vec3 f = fresnel_schlick(gmax(dot(h, wo), 0.0f), f0);
  const float ndf = distribution_ggx(h, n, roughness);
  const float g = geometry_smith(wo, wi, n, roughness);
  const vec3 num = (ndf * g) * f;
  const float denom = 4.0f * gmax(dot(n, wo), 0.0f) * gmax(dot(n, wi), 0.0f) + 0.0001f;
  const vec3 specular = inv_scale(num, denom);
  vec3 kd = v3(1.0f) - f;
  kd *= 1.0f - metallic;
  const float n_dot_l = gmax(dot(n, wi), 0.0f);
  return ((kd * albedo) * (1.0f / PI_GLSL) + specular) * n_dot_l;
}
inline float ns_spec_prob(float metallic) { return 0.25f + 0.5f * metallic; }
inline float ns_pdf(vec3 n, vec3 wo, vec3 wi, float roughness, float metallic) {
  const float ps = ns_spec_prob(metallic);
  const float n_dot_l = gmax(dot(n, wi), 0.0f);
  const vec3 h = normalize(wo + wi);
  const float n_dot_h = gmax(dot(n, h), 0.0f);
  const float v_dot_h = gmax(dot(wo, h), 0.000001f);
  const float d = distribution_ggx(h, n, roughness);
  return (1.0f - ps) * (n_dot_l / PI_GLSL) + ps * ((d * n_dot_h) / (4.0f * v_dot_h));
}
// returns false when the sampled direction is below the surface
inline bool ns_sample(vec3 n, vec3 wo, float roughness, float metallic, float u0, float u1, float u2, vec3& wi) {
  vec3 nt, nb; Invocation::create_coordinate_system(n, nt, nb);
  const float phi = (2.0f * PI_GLSL) * u2;
  const float cp = cos_c(phi), sp = sin_c(phi);
  if (u0 < ns_spec_prob(metallic)) {
    const float a = roughness * roughness;
    const float a2 = a * a;
    const float cos_t = sqrtf((1.0f - u1) / (1.0f + (a2 - 1.0f) * u1));
    const float sin_t = sqrtf(gmax(1.0f - cos_t * cos_t, 0.0f));
    const vec3 l = v3(sin_t * cp, cos_t, sin_t * sp);
    const vec3 h = v3((l.x * nb.x + l.y * n.x) + l.z * nt.x, (l.x * nb.y + l.y * n.y) + l.z * nt.y, (l.x * nb.z + l.y * n.z) + l.z * nt.z);
    wi = h * (2.0f * dot(wo, h)) - wo;
  } else {
    const float r = sqrtf(u1);
    const vec3 l = v3(r * cp, sqrtf(gmax(1.0f - u1, 0.0f)), r * sp);
    wi = v3((l.x * nb.x + l.y * n.x) + l.z * nt.x, (l.x * nb.y + l.y * n.y) + l.z * nt.y, (l.x * nb.z + l.y * n.z) + l.z * nt.z);
  }
  wi = normalize(wi);
  return dot(n, wi) > 0.0f;
}

struct NsLightCounts { uint32_t has_dir, n_pl, n_em, n_types; };
inline NsLightCounts ns_light_counts(const Scene& sc) {
  NsLightCounts c;
  const float* col = sc.lights.dir_light.colour;
  c.has_dir = (col[0] > 0.0f || col[1] > 0.0f || col[2] > 0.0f) ? 1u : 0u;
  c.n_pl = sc.lights.num_pointlights;
  c.n_em = (uint32_t)sc.emissive_idx.size();
  c.n_types = c.has_dir + (c.n_pl ? 1u : 0u) + (c.n_em ? 1u : 0u);
  return c;
}
inline uint32_t ns_pick(float u, uint32_t n) { uint32_t i = to_uint(u * (float)n); return i >= n ? n - 1 : i; }

// solid-angle pdf with which next-event estimation at a vertex would have produced the point at distance `dist`
// along `wi` on an emissive triangle of world area `area`, light-side normal nl (MIS partner of BSDF sampling)
inline float ns_nee_pdf_emissive(const NsLightCounts& lc, uint32_t ntri, float area, float dist, vec3 nl, vec3 wi) {
  const float cos_l = fabsf(dot(nl, wi));
  if (!(cos_l > 0.000001f) || !(area > 0.0f)) return 0.0f;
  const float p_pick = (1.0f / (float)lc.n_types) * (1.0f / ((float)lc.n_em * (float)ntri));
  return p_pick * ((dist * dist) / (cos_l * area));
}

// one path of one pixel; `frame.g` holds the primary hit. Returns the indirect radiance; counts rays in inv.
inline vec3 ns_indirect(Frame& F, Invocation& inv, uint32_t x, uint32_t y) {
  const Scene& sc = *F.sc;
  const size_t px = (size_t)y * F.g.w + x;
  const uint32_t pixel = (uint32_t)px, frame = F.ubo.frame_idx;
  const NsLightCounts lc = ns_light_counts(sc);
  const uint32_t max_bounces = F.settings.max_bounces;
  // with the reservoir passes switched off the path does its own next-event estimation at x0 (plain path tracer)
  const uint32_t nee_from = (F.settings.flags & RT_NO_RESTIR) ? 0u : 1u;
  vec3 L = v3(0.0f);
  if (F.g.prim[px] == INVALID) return L;
  NsSurf s = ns_surface(sc, F.g.inst[px], F.g.prim[px], F.g.bary[2 * px], F.g.bary[2 * px + 1]);
  if (s.flags & MAT_FLAG_EMISSIVE) return L;
  const vec3 cam = v3(F.ubo.cam_pos[0], F.ubo.cam_pos[1], F.ubo.cam_pos[2]);
  vec3 wo = normalize(cam - s.p);
  vec3 T = v3(1.0f);
  float pdf_b = 0.0f;
  for (uint32_t b = 0; b <= max_bounces; ++b) {
    // vertex x_b: shading frame faces the incoming path
    vec3 n = s.n;
    if (dot(n, wo) < 0.0f) n = -n;
    const vec3 o = s.p + n * 0.01f;
    const uint32_t dim = 16u * b;
    // ---- next-event estimation (x0's direct light comes from the ReSTIR passes unless they are off)
    if (b >= nee_from && lc.n_types > 0) {
      uint32_t t = ns_pick(ns_rand(pixel, frame, dim + 3), lc.n_types);
      // map the t-th present type to 0 = directional, 1 = point, 2 = emissive
      uint32_t type = 0;
      { uint32_t k = 0; bool found = false;
        if (lc.has_dir) { if (k == t) { type = 0; found = true; } ++k; }
        if (!found && lc.n_pl) { if (k == t) { type = 1; found = true; } ++k; }
        if (!found && lc.n_em) { if (k == t) { type = 2; found = true; } ++k; } }
      const float p_type = 1.0f / (float)lc.n_types;
      const float u1 = ns_rand(pixel, frame, dim + 4), u2 = ns_rand(pixel, frame, dim + 5);
      const float u3 = ns_rand(pixel, frame, dim + 6), u4 = ns_rand(pixel, frame, dim + 7);
      vec3 C = v3(0.0f), wi = v3(0.0f, 1.0f, 0.0f);
      float tmax = -1.0f;
      if (type == 0) {
        wi = v3(sc.lights.dir_light.dir[0], sc.lights.dir_light.dir[1], sc.lights.dir_light.dir[2]);
        if (dot(n, wi) > 0.0f) {
          const vec3 col = v3(sc.lights.dir_light.colour[0], sc.lights.dir_light.colour[1], sc.lights.dir_light.colour[2]);
          C = ((T * ns_bsdf_cos(n, wo, wi, s.albedo, s.roughness, s.metallic)) * col) / p_type;
          tmax = 10000.0f;
        }
      } else if (type == 1) {
        const uint32_t idx = ns_pick(u1, lc.n_pl);
        const Pointlight& pl = sc.lights.pointlights[idx];
        const vec3 d = v3(pl.pos[0], pl.pos[1], pl.pos[2]) - o;
        const float dist = length(d);
        wi = d / dist;
        if (dot(n, wi) > 0.0f && dist > 0.002f) {
          const float atten = (pl.atten_constant + pl.atten_linear * dist) + pl.atten_exp * (dist * dist);
          const vec3 col = v3(pl.colour[0], pl.colour[1], pl.colour[2]);
          C = (((T * ns_bsdf_cos(n, wo, wi, s.albedo, s.roughness, s.metallic)) * col) / atten) / (p_type * (1.0f / (float)lc.n_pl));
          tmax = dist - 0.001f;
        }
      } else {
        const uint32_t obj = sc.emissive_idx[ns_pick(u1, lc.n_em)];
        const uint32_t ntri = sc.instances[obj].num_mesh_indices / 3u;
        const uint32_t tri = ns_pick(u2, ntri);
        float bu = u3, bv = u4;
        if (bu + bv >= 1.0f) { bu = 1.0f - bu; bv = 1.0f - bv; }
        const NsSurf q = ns_light_point(sc, obj, tri, bu, bv);
        const vec3 d = q.p - o;
        const float dist = length(d);
        wi = d / dist;
        const float pdf_w = ns_nee_pdf_emissive(lc, ntri, q.area, dist, q.n, wi);
        if (dot(n, wi) > 0.0f && pdf_w > 0.0f && dist > 0.01f) {
          const float p_bsdf = ns_pdf(n, wo, wi, s.roughness, s.metallic);
          const float w = (pdf_w * pdf_w) / (pdf_w * pdf_w + p_bsdf * p_bsdf);
          C = (((T * ns_bsdf_cos(n, wo, wi, s.albedo, s.roughness, s.metallic)) * q.albedo) * w) / pdf_w;
          tmax = dist - 0.005f;
        }
      }
      if (tmax > 0.0f) {
        ++inv.n_any;
        const Hit h = sc.trace(Ray{{o.x, o.y, o.z}, 0.001f, {wi.x, wi.y, wi.z}, tmax}, true);
        const float vis = h.t >= 0.0f ? 0.0f : 1.0f;
        L += C * vis;
      }
    }
    if (b == max_bounces) break;
    // ---- Russian roulette from x2
    if (b >= 2) {
      const float q = gmin(gmax(gmax(T.x, T.y), T.z), 0.95f);
      if (!(ns_rand(pixel, frame, dim + 8) < q)) break;
      T = T / q;
    }
    // ---- continue the path
    vec3 wi;
    if (!ns_sample(n, wo, s.roughness, s.metallic, ns_rand(pixel, frame, dim), ns_rand(pixel, frame, dim + 1), ns_rand(pixel, frame, dim + 2), wi)) break;
    pdf_b = ns_pdf(n, wo, wi, s.roughness, s.metallic);
    if (!(pdf_b > 0.0f)) break;
    T = (T * ns_bsdf_cos(n, wo, wi, s.albedo, s.roughness, s.metallic)) / pdf_b;
    if (!(gmax(gmax(T.x, T.y), T.z) > 0.0f)) break;
    ++inv.n_closest;
    const Hit h = sc.trace(Ray{{o.x, o.y, o.z}, 0.001f, {wi.x, wi.y, wi.z}, 1000.0f}, false);
    if (h.t < 0.0f) break;
    s = ns_surface(sc, h.inst, h.prim, h.u, h.v);
    if (s.flags & MAT_FLAG_EMISSIVE) {
      // BSDF sampling found a light: counted from the first NEE vertex on (x0's direct light is the ReSTIR term), MIS against NEE
      if (b >= nee_from) {
        const float p_nee = ns_nee_pdf_emissive(lc, s.ntri, s.area, h.t, s.n, wi);
        const float w = (pdf_b * pdf_b) / (pdf_b * pdf_b + p_nee * p_nee);
        L += (T * s.albedo) * w;
      }
      break;
    }
    wo = -wi;
  }
  return L;
}

// path stage of rows [y0,y1): out = direct + indirect
inline void render_northstar_paths(Frame& F, uint32_t y0, uint32_t y1, int threads) {
  parallel_rows(y0, y1, threads, [&](uint32_t y) {
    uint64_t nc = 0, na = 0;
    for (uint32_t x = 0; x < F.g.w; ++x) {
      const size_t px = (size_t)y * F.g.w + x;
      Invocation inv(F, x, y);
      vec3 out = v3(0.0f);
      if (F.g.prim[px] != INVALID) {
        const NsSurf s = ns_surface(*F.sc, F.g.inst[px], F.g.prim[px], F.g.bary[2 * px], F.g.bary[2 * px + 1]);
        if (s.flags & MAT_FLAG_EMISSIVE) out = s.albedo;
        else out = v3(F.direct_rgba[4 * px], F.direct_rgba[4 * px + 1], F.direct_rgba[4 * px + 2]) + ns_indirect(F, inv, x, y);
      }
      F.out_rgba[4 * px] = out.x; F.out_rgba[4 * px + 1] = out.y; F.out_rgba[4 * px + 2] = out.z; F.out_rgba[4 * px + 3] = 1.0f;
      nc += inv.n_closest; na += inv.n_any;
    }
    F.rays_closest += nc; F.rays_any += na;
  });
}

}  // namespace snkor

#include "snk_restir_ns.h"

namespace snkor {

// north-star frame over rows [y0,y1): ReSTIR DI stages (snk_restir_ns.h) for the direct term, then one path per pixel
inline void render_northstar(Frame& F, uint32_t y0, uint32_t y1, int threads) {
  F.ivp = inv_view_proj(F.ubo);
  if (F.settings.flags & RT_NO_RESTIR) {
    parallel_rows(y0, y1, threads, [&](uint32_t y) {
      for (uint32_t x = 0; x < F.g.w; ++x) { float* d = &F.direct_rgba[4 * ((size_t)y * F.g.w + x)]; d[0] = d[1] = d[2] = 0.0f; d[3] = 1.0f; }
    });
  } else {
    render_restir_ns(F, F.rs, !F.ubo.prev_frame_invalid, y0, y1, F.ns_halo, threads);
  }
  render_northstar_paths(F, y0, y1, threads);
}

}  // namespace snkor
