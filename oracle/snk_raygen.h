// oracle/snk_raygen.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Function-for-function CPU restatement of the reference's ray-tracing shaders:
//   Core/res/shaders/RayGen.rgen        (all functions, cited per function)
//   Core/res/shaders/RayClosestHit.rchit:87-142
//   Core/res/shaders/RayMiss.rmiss:15-26, Shadow.rmiss:11-13
//   Core/res/shaders/LightBuffers.glsl:65-157
//   Core/res/shaders/Util.glsl:1-7, 23-57, 77-82
// Expression order follows the GLSL text token for token (left-to-right float32,
// see snk_math.h for the numerics contract). Quirks of the reference
//  are reproduced; the four places where the reference
// would read out of bounds are guarded and marked DEVIATION.
// PARITY: the reference ships no tests or golden vectors and cannot run here, so whole frames are UNPINNED. Pinned against
// the reference's own text: the BRDF / light functions and the integer RNG (executed from the GLSL by tools/glsl_exec.py ->
// tests/golden/reference_shader_vectors.npz), 36 constants and the block layouts (tests/golden/reference_*.json).
#pragma once
#include <atomic>

#include "snk_gbuffer.h"

namespace snkor {

constexpr float PI_GLSL = 3.1415926f;                         // LightBuffers.glsl:29
constexpr float PHI_GLSL = 1.61803398874989484820459f;        // Util.glsl:77

struct RtSettings {
  int32_t num_temporal_resamples = 1, num_spatial_resamples = 3;  // Raytracing.h:13-16
  uint32_t mode = 0;
  uint32_t ris_iters = 32;    // RayGen.rgen:95
  uint32_t m_cap = 20;        // RayGen.rgen:97
  float neighbour_radius = 30.0f;   // RayGen.rgen:488
  float normal_threshold = 0.95f;   // RayGen.rgen:435, :491
  uint32_t max_bounces = 4, spp = 1, flags = 0;
  uint32_t reserved[6] = {0, 0, 0, 0, 0, 0};
};
static_assert(sizeof(RtSettings) == 64, "RtSettings mirrors SnkRtSettings");
constexpr uint32_t RT_NO_RESTIR = 1u, RT_NO_BOUNCE = 2u, RT_UNBIASED_REUSE = 4u;

// optional scripted random numbers for known-answer tests: when set, every
// gold_noise()/random() call returns the next value of the stream instead
struct RngScript { const float* values = nullptr; size_t count = 0; size_t next = 0; };

// ---- Util.glsl ---------------------------------------------------------------------
inline uint32_t hash1(uint32_t x) {  // Util.glsl:23-30
  x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u);
  return x;
}
inline float float_construct(uint32_t m) {  // Util.glsl:43-52
  m &= 0x007FFFFFu; m |= 0x3F800000u;
  return u2f(m) - 1.0f;
}
inline float random_f(float x) { return float_construct(hash1(f2u(x))); }  // Util.glsl:57
inline float gold_noise_f(float px, float py, float seed) {  // Util.glsl:79-82
  const float ax = px * PHI_GLSL, ay = py * PHI_GLSL;
  const float dx = ax - px, dy = ay - py;
  const float dist = sqrtf(dx * dx + dy * dy);
  const float t = tan_c(dist * seed) * px;
  return fract(t);
}
inline uint32_t to_uint(float f) { return (f >= 0.0f && f < 4294967296.0f) ? (uint32_t)f : 0u; }

// ---- LightBuffers.glsl ----------------------------------------------------------------
inline vec3 fresnel_schlick(float cos_theta, vec3 f0) {  // :65-67
  const float p = pow5(gclamp(1.0f - cos_theta, 0.0f, 1.0f));
  return f0 + (v3(1.0f) - f0) * p;
}
inline float distribution_ggx(vec3 h, vec3 n, float roughness) {  // :69-80
  const float a = roughness * roughness;
  const float a2 = a * a;
  const float n_dot_h = gmax(dot(n, h), 0.0f);
  const float n_dot_h2 = n_dot_h * n_dot_h;
  float denom = (n_dot_h2 * (a2 - 1.0f) + 1.0f);
  denom = PI_GLSL * denom * denom;
  return a2 / denom;
}
inline float geometry_schlick_ggx(float n_dot_v, float roughness) {  // :82-87
  const float r = (roughness + 1.0f);
  const float k = (r * r) / 8.0f;
  return n_dot_v / (n_dot_v * (1.0f - k) + k);
}
inline float geometry_smith(vec3 v, vec3 l, vec3 n, float roughness) {  // :89-96
  const float n_dot_v = gmax(dot(n, v), 0.0f);
  const float n_dot_l = gmax(dot(n, l), 0.0f);
  const float ggx2 = geometry_schlick_ggx(n_dot_v, roughness);
  const float ggx1 = geometry_schlick_ggx(n_dot_l, roughness);
  return ggx2 * ggx1;
}
inline vec3 cook_torrance_brdf(vec3 light_pos, vec3 surface_pos, vec3 v, vec3 f0, vec3 n, float roughness, float metallic) {  // :104-124
  const vec3 frag_to_light = light_pos - surface_pos;
  const vec3 l = normalize(frag_to_light);
  const vec3 h = normalize(v + l);
  const vec3 f = fresnel_schlick(gmax(dot(h, v), 0.0f), f0);
  const float ndf = distribution_ggx(h, n, roughness);
  const float g = geometry_smith(v, l, n, roughness);
  const float n_dot_l = gmax(dot(n, l), 0.0f);
  const vec3 num = (ndf * g) * f;
  const float denom = 4.0f * gmax(dot(n, v), 0.0f) * gmax(n_dot_l, 0.0f) + 0.0001f;
  const vec3 specular = num / denom;
  vec3 kd = v3(1.0f) - f;
  kd *= 1.0f - metallic;
  return (kd / PI_GLSL + specular) * n_dot_l;
}
inline vec3 calc_directional_light(const LightData& L, vec3 v, vec3 f0, vec3 n, float roughness, float metallic, vec3 albedo) {  // :126-144
  const vec3 l = v3(L.dir_light.dir[0], L.dir_light.dir[1], L.dir_light.dir[2]);
  const vec3 h = normalize(v + l);
  const vec3 f = fresnel_schlick(gmax(dot(h, v), 0.0f), f0);
  const float ndf = distribution_ggx(h, n, roughness);
  const float g = geometry_smith(v, l, n, roughness);
  const vec3 num = (ndf * g) * f;
  const float denom = 4.0f * gmax(dot(n, v), 0.0f) * gmax(dot(n, l), 0.0f) + 0.0001f;
  const vec3 specular = num / denom;
  vec3 kd = v3(1.0f) - f;
  kd *= 1.0f - metallic;
  const float n_dot_l = gmax(dot(n, l), 0.0f);
  const vec3 c = v3(L.dir_light.colour[0], L.dir_light.colour[1], L.dir_light.colour[2]);
  return ((kd * albedo / PI_GLSL + specular) * n_dot_l) * c;
}
inline vec3 calc_pointlight(const Pointlight& light, vec3 v, vec3 f0, vec3 world_pos, vec3 n, float roughness, float metallic, vec3 albedo) {  // :147-157
  const vec3 lp = v3(light.pos[0], light.pos[1], light.pos[2]);
  const vec3 frag_to_light = lp - world_pos;
  const float distance = length(frag_to_light);
  const float attenuation = (light.atten_constant + light.atten_linear * distance) + light.atten_exp * (distance * distance);
  const vec3 brdf = cook_torrance_brdf(lp, world_pos, v, f0, n, roughness, metallic);
  return ((brdf * albedo) * v3(light.colour[0], light.colour[1], light.colour[2])) / attenuation;
}

// ---- frame state -----------------------------------------------------------------------
struct Payload { vec3 colour{0, 0, 0}, normal{0, 0, 0}; float distance = -1.0f, roughness = 0.0f; uint32_t mat_flags = 0; };

// scratch planes of the north-star ReSTIR stages (snk_restir_ns.h)
struct RsScratch {
  std::vector<float> phat;  // 2 / px: p^ of the initial samples
  std::vector<float> vis;   // 3 / px
  std::vector<Ray> rays;    // 3 / px
  std::vector<float> C;     // 9 / px: final unshadowed contributions (pl, em, dir)
  void resize(size_t n) {
    phat.assign(2 * n, 0.0f); vis.assign(3 * n, 1.0f); rays.assign(3 * n, Ray{{0, 0, 0}, 0.0f, {0, 1, 0}, -1.0f}); C.assign(9 * n, 0.0f);
  }
};

struct Frame {
  const Scene* sc = nullptr;
  RsScratch rs;
  uint32_t ns_halo = 0;  // rows around the rendered band that also run the initial + temporal stages
  CommonUBO ubo{}, prev_ubo{};
  RtSettings settings{};
  GBuffer g;
  std::vector<PixelReservoir> reservoirs;  // 2·W·H: half 0 final, half 1 temporal (EditorLayer.cpp:299-302)
  std::vector<float> out_rgba;             // W·H·4 float32 (imageStore value before the fp16 store)
  std::vector<float> direct_rgba;          // dir light + reservoir radiance (pass-2 `light`)
  std::atomic<uint64_t> rays_closest{0}, rays_any{0};
  mat4 ivp{};
  bool force_visible = false;  // tests only: every shadow ray reports 'visible' (function-level checks against the executed shader text)
  float forced_rays[4][7] = {};  // ... and the last shadow rays asked for are kept: origin, direction, max distance
  uint32_t n_forced = 0;
  void resize(uint32_t w, uint32_t h) {
    g.resize(w, h);
    PixelReservoir e{}; e.pl_idx = INVALID; e.object_instance_idx = INVALID; e.triangle_idx = INVALID;
    reservoirs.assign((size_t)2 * w * h, e);
    out_rgba.assign((size_t)4 * w * h, 0.0f);
    direct_rgba.assign((size_t)4 * w * h, 0.0f);
  }
};

// one shader invocation (gl_LaunchIDEXT = (x, y))
struct Invocation {
  Frame& F;
  const Scene& sc;
  const LightData& L;
  uint32_t x, y;
  float fx, fy;  // vec2(gl_LaunchIDEXT.xy)
  Payload payload;
  RngScript* script;
  uint64_t n_closest = 0, n_any = 0;

  Invocation(Frame& f, uint32_t px, uint32_t py, RngScript* s = nullptr)
      : F(f), sc(*f.sc), L(f.sc->lights), x(px), y(py), fx((float)px), fy((float)py), script(s) {}

  float gold_noise(float px, float py, float seed) {
    if (script && script->values) return script->values[script->next++ % script->count];
    return gold_noise_f(px, py, seed);
  }
  float random(float s) {
    if (script && script->values) return script->values[script->next++ % script->count];
    return random_f(s);
  }
  float time() const { return F.ubo.time_elapsed; }

  // RayGen.rgen:101-121 / RayClosestHit.rchit:64-84
  float shadow_test(vec3 ro, vec3 rd, float max_dist) {
    ++n_any;
    if (F.force_visible) {
      float* q = F.forced_rays[F.n_forced++ & 3u];
      q[0] = ro.x; q[1] = ro.y; q[2] = ro.z; q[3] = rd.x; q[4] = rd.y; q[5] = rd.z; q[6] = max_dist;
      return 1.0f;
    }
    Ray r{{ro.x, ro.y, ro.z}, 0.001f, {rd.x, rd.y, rd.z}, max_dist};
    Hit h = sc.trace(r, true);
    return 1.0f - (h.t >= 0.0f ? 1.0f : 0.0f);
  }

  // traceRayEXT(closest) + RayClosestHit.rchit:87-142 + RayMiss.rmiss:15-26
  void trace_closest(vec3 ro, float tmin, vec3 rd, float tmax) {
    ++n_closest;
    Ray r{{ro.x, ro.y, ro.z}, tmin, {rd.x, rd.y, rd.z}, tmax};
    Hit h = sc.trace(r, false);
    if (h.t < 0.0f) { payload.colour = v3(0.0f); payload.distance = -1.0f; return; }
    shade_closest_hit(h, rd);
  }
  // RayClosestHit.rchit:87-142 for a given hit
  void shade_closest_hit(const Hit& h, vec3 rd) {
    const float b0 = (1.0f - h.u) - h.v, b1 = h.u, b2 = h.v;                     // rchit:88
    const InstanceData& in = sc.instances[h.inst];                              // rchit:90
    uint32_t ia, ib, ic; sc.tri_indices(in, h.prim, ia, ib, ic);                // rchit:91
    const vec3 p = (sc.pos(ia) * b0 + sc.pos(ib) * b1) + sc.pos(ic) * b2;       // rchit:105
    vec3 n = (sc.nrm(ia) * b0 + sc.nrm(ib) * b1) + sc.nrm(ic) * b2;             // rchit:106
    const vec3 pos = mul_point(sc.transforms[in.transform_idx], p);             // rchit:110
    n = normalize(mul(sc.normal_matrix[in.transform_idx], n));                  // rchit:111
    const Material& m = sc.materials[in.material_idx];                          // rchit:113
    const vec3 albedo = sc.albedo_at(m, ia, ib, ic, b0, b1, b2);                // rchit:100-116
    payload.mat_flags = m.flags;                                                // rchit:118
    if (m.flags & MAT_FLAG_EMISSIVE) { payload.colour = albedo; payload.distance = h.t; return; }  // rchit:119-123
    const vec3 v = normalize(-rd);                                              // rchit:127
    const vec3 f0 = mix(v3(0.04f), albedo, m.metallic);                         // rchit:131-132
    const vec3 biased_pos = pos + n * 0.01f;                                    // rchit:134
    const vec3 ld = v3(L.dir_light.dir[0], L.dir_light.dir[1], L.dir_light.dir[2]);
    const vec3 light = calc_directional_light(L, v, f0, n, m.roughness, m.metallic, albedo) * shadow_test(biased_pos, ld, 10000.0f);  // rchit:136
    payload.colour = light; payload.normal = n; payload.distance = h.t; payload.roughness = m.roughness;  // rchit:138-141
  }

  // RayGen.rgen:123-130
  static void create_coordinate_system(vec3 n, vec3& nt, vec3& nb) {
    if (fabsf(n.x) > fabsf(n.y)) nt = v3(n.z, 0.0f, -n.x) / length2(n.x, n.z);
    else nt = v3(0.0f, -n.z, n.y) / length2(n.y, n.z);
    nb = cross(n, nt);
  }
  // RayGen.rgen:132-136
  static vec3 uniform_sample_hemisphere(float r1, float r2) {
    const float sin_theta = sqrtf(1.0f - r1 * r1);
    const float phi = (2.0f * PI_GLSL) * r2;
    return v3(sin_theta * cos_c(phi), r1, sin_theta * sin_c(phi));
  }

  // RayGen.rgen:139-213
  vec3 bounce_diffuse(vec3 ro, vec3 normal, vec3 albedo) {
    vec3 colour_sum = v3(0.0f);
    vec3 nt, nb; create_coordinate_system(normal, nt, nb);
    vec3 nt_inner, nb_inner, n_inner;
    const float l2_ro = dot(ro, ro);
    float seed = fract((((((l2_ro + ro.x) + ro.y) + ro.z) + F.ubo.time_elapsed) + F.ubo.delta_time));
    float xx = fx, xy = fy;
    xx += fract(sin_c(xx * 1038.283726f + xy * F.ubo.delta_time));
    xy += fract(cos_c(fx * F.ubo.time_elapsed + fy * 9278.33337f));
    for (uint32_t ray_idx = 0; ray_idx < 8; ++ray_idx) {
      seed = fract(seed + (float)ray_idx / 8.0f);
      const float s = fract(seed + (float)ray_idx / 8.0f);
      const vec3 rnd = uniform_sample_hemisphere(gold_noise(xx, xy, s), gold_noise(xx, xy, s + 0.1f));
      vec3 rd = v3((rnd.x * nb.x + rnd.y * normal.x) + rnd.z * nt.x, (rnd.x * nb.y + rnd.y * normal.y) + rnd.z * nt.y,
                   (rnd.x * nb.z + rnd.y * normal.z) + rnd.z * nt.z);
      trace_closest(ro, 0.01f, rd, 1000.0f);
      colour_sum += ((albedo * payload.colour) * gmax(cos_c(dot(rd, normal)), 0.0f)) / 8.0f;  // :178 cos OF the dot (quirk 1)
      if (payload.distance < 0.0f || (payload.mat_flags & MAT_FLAG_EMISSIVE)) continue;
      const vec3 new_pos = ro + rd * payload.distance;
      n_inner = payload.normal;
      create_coordinate_system(n_inner, nt_inner, nb_inner);
      for (uint32_t bounce_idx = 0; bounce_idx < 2; ++bounce_idx) {
        const vec3 h = uniform_sample_hemisphere(gold_noise(xx, xy, fract((seed + (float)bounce_idx * 0.5f) + 0.2f)),
                                                 gold_noise(xx, xy, fract((seed + (float)bounce_idx * 0.5f) + 0.3f)));
        rd = v3((h.x * nb_inner.x + h.y * n_inner.x) + h.z * nt_inner.x, (h.x * nb_inner.y + h.y * n_inner.y) + h.z * nt_inner.y,
                (h.x * nb_inner.z + h.y * n_inner.z) + h.z * nt_inner.z);
        trace_closest(new_pos, 0.01f, rd, 1000.0f);
        colour_sum += (((albedo * payload.colour) * 0.25f) * gmax(cos_c(dot(rd, n_inner)), 0.0f)) / 8.0f;  // :208
      }
    }
    return colour_sum * (2.0f * PI_GLSL);
  }

  // RayGen.rgen:215-226
  bool update_pointlight_reservoir(PixelReservoir& r, uint32_t pl_idx, float weight, float rnd_seed) {
    r.pl_m++;
    const float rnd = gold_noise(fx, fy, fract((weight + time() * 0.0009783f) + rnd_seed));
    r.w_sum_pl += weight;
    if (r.w_sum_pl > 0.0f && rnd < weight / r.w_sum_pl) { r.pl_idx = pl_idx; return true; }
    return false;
  }
  // RayGen.rgen:228-239
  bool update_emissive_reservoir(PixelReservoir& r, uint32_t obj, uint32_t tri, float bx, float by, float weight, float rnd_seed) {
    r.m_em++;
    const float rnd = gold_noise(fx, fy, fract((weight + time() * 0.000789f) + rnd_seed));
    r.w_sum_em += weight;
    if (r.w_sum_em > 0.0f && rnd < weight / r.w_sum_em) {
      r.object_instance_idx = obj; r.triangle_idx = tri; r.bary[0] = bx; r.bary[1] = by;
      return true;
    }
    return false;
  }

  // the point on an emissive triangle (shared head of RayGen.rgen:241-251 and :258-268)
  void triangle_point(uint32_t obj, uint32_t tri, float bx, float by, vec3& p, vec3& point_normal, const InstanceData*& in) {
    in = &sc.instances[obj];
    uint32_t ia, ib, ic; sc.tri_indices(*in, tri, ia, ib, ic);
    const float bz = (1.0f - bx) - by;  // vec3(bary.xy, 1 - x - y): weights of (ti.x, ti.y, ti.z)
    p = (bx * sc.pos(ia) + by * sc.pos(ib)) + bz * sc.pos(ic);
    point_normal = (bx * sc.nrm(ia) + by * sc.nrm(ib)) + bz * sc.nrm(ic);
    p = mul_point(sc.transforms[in->transform_idx], p);
    point_normal = normalize(mul(sc.normal_matrix[in->transform_idx], point_normal));
  }
  // RayGen.rgen:241-256
  float triangle_shadow_test(uint32_t obj, uint32_t tri, float bx, float by, vec3 surface_pos) {
    vec3 p, pn; const InstanceData* in; triangle_point(obj, tri, bx, by, p, pn, in);
    const vec3 target = p + pn * 0.01f;
    const float point_dist = length(surface_pos - target);
    const vec3 dir = (target - surface_pos) / point_dist;
    return shadow_test(surface_pos, dir, point_dist);
  }
  // RayGen.rgen:258-282
  vec3 compute_triangle_radiance(uint32_t obj, uint32_t tri, float bx, float by, vec3 v, vec3 f0, vec3 world_pos, vec3 n,
                                 float roughness, float metallic, vec3 albedo, vec3 biased_pos, bool do_shadow) {
    vec3 p, pn; const InstanceData* in; triangle_point(obj, tri, bx, by, p, pn, in);
    const Material& m = sc.materials[in->material_idx];
    const vec3 emissive = v3(m.albedo[0], m.albedo[1], m.albedo[2]);
    const vec3 brdf = cook_torrance_brdf(p, world_pos, v, f0, n, roughness, metallic);
    const vec3 dlt = world_pos - p;
    vec3 radiance = ((emissive * albedo) * brdf) / fabsf(dot(dlt, dlt));
    if (do_shadow) {
      const vec3 target = p + pn * 0.01f;
      const float point_dist = length(biased_pos - target);
      const vec3 dir = (target - biased_pos) / point_dist;
      radiance *= shadow_test(biased_pos, dir, point_dist);
    }
    return radiance;
  }

  // RayGen.rgen:284-294
  static void init_pixel_reservoir(PixelReservoir& r) {
    r.pl_idx = INVALID; r.w_sum_pl = 0.0f; r.pl_m = 0;
    r.object_instance_idx = INVALID; r.triangle_idx = INVALID; r.w_sum_em = 0.0f; r.m_em = 0;
    // w_pl / w_em / bary are left uninitialised by the reference; zero here (DEVIATION: defined value)
    r.w_pl = 0.0f; r.w_em = 0.0f; r.bary[0] = r.bary[1] = 0.0f; r.pad = 0;
  }

  // RayGen.rgen:296-357
  // RayGen.rgen:300-309: the point-light half of DoInitialWRS; returns p_hat of the selected light
  float ris_pointlights(PixelReservoir& r, vec3 v, vec3 f0, vec3 world_pos, vec3 n, float roughness, float metallic, vec3 albedo) {
    float p_hat_pl = 0.0f;
    const uint32_t iters = F.settings.ris_iters;
    const uint32_t n_pl = L.num_pointlights;
    for (uint32_t i = 0; i < iters; ++i) {
      if (n_pl == 0) { r.pl_m++; continue; }  // DEVIATION: reference reads light slot 0 of an empty array (weight 0/NaN, never selected)
      const float seed = fract(((time() * 0.001197f + world_pos.x) + world_pos.y) + world_pos.z) + (float)i / (float)iters;
      uint32_t rnd_idx = to_uint(gold_noise(fx, fy, seed) * (float)n_pl);
      if (rnd_idx >= n_pl) rnd_idx = n_pl - 1;  // DEVIATION: guard
      const float p_hat = length(calc_pointlight(L.pointlights[rnd_idx], v, f0, world_pos, n, roughness, metallic, albedo));
      const float weight = p_hat / (1.0f / (float)n_pl);
      if (update_pointlight_reservoir(r, rnd_idx, weight, sin_c((float)i * 20382997.34741f))) p_hat_pl = p_hat;
    }
    return p_hat_pl;
  }
  // RayGen.rgen:311-342: the emissive-triangle half of DoInitialWRS; returns p_hat of the selected triangle point
  float ris_emissives(PixelReservoir& r, vec3 v, vec3 f0, vec3 world_pos, vec3 n, float roughness, float metallic, vec3 albedo, vec3 biased_pos) {
    float p_hat_em = 0.0f;
    const uint32_t iters = F.settings.ris_iters;
    const uint32_t n_em = (uint32_t)sc.emissive_idx.size();
    for (uint32_t i = 0; i < iters; ++i) {
      if (n_em == 0) { r.m_em++; continue; }  // DEVIATION: reference reads idx[0] of an empty list
      const float wsum = (world_pos.x + world_pos.y) + world_pos.z;
      uint32_t rnd_emissive_idx = to_uint(gold_noise(fx, fy, fabsf(fract((wsum + time() * 0.001534f) + (float)i / (float)iters))) * (float)n_em);
      if (rnd_emissive_idx >= n_em) rnd_emissive_idx = n_em - 1;  // DEVIATION: guard
      const uint32_t obj = sc.emissive_idx[rnd_emissive_idx];
      const InstanceData& in = sc.instances[obj];
      const uint32_t ntri = in.num_mesh_indices / 3;
      uint32_t tri = to_uint(gold_noise(fx, fy, fabsf(fract((wsum + time() * 0.00119f) + (float)i / (float)iters))) * (float)ntri);
      if (tri >= ntri) tri = ntri - 1;  // DEVIATION: guard
      uint32_t ia, ib, ic; sc.tri_indices(in, tri, ia, ib, ic);
      float bx = random(((sin_c((float)i * 27934.3221f + fx * 293.28f)) + time() * 0.001f) + (float)i);
      float by = random((((sin_c((float)i * 12391.8583f + fy * 983.3399f)) + (float)(int32_t)ib) + time() * 0.0091f) + (float)i);
      if (bx + by >= 1.0f) { bx = 1.0f - bx; by = 1.0f - by; }
      const float triangle_area = (sc.det3[in.transform_idx] * 0.5f) * length(cross(sc.pos(ib) - sc.pos(ia), sc.pos(ib) - sc.pos(ic)));
      const vec3 radiance = compute_triangle_radiance(obj, tri, bx, by, v, f0, world_pos, n, roughness, metallic, albedo, biased_pos, false);
      const float p_hat = length(radiance);
      const float pdf = 1.0f / ((float)(n_em * ntri) * triangle_area);
      const float weight = p_hat / pdf;
      if (update_emissive_reservoir(r, obj, tri, bx, by, weight, sin_c((float)i * 38293.28347f))) p_hat_em = p_hat;
    }
    return p_hat_em;
  }
  void do_initial_wrs(PixelReservoir& r, vec3 v, vec3 f0, vec3 world_pos, vec3 n, float roughness, float metallic, vec3 albedo, vec3 biased_pos) {
    const float p_hat_pl = ris_pointlights(r, v, f0, world_pos, n, roughness, metallic, albedo);
    const float p_hat_em = ris_emissives(r, v, f0, world_pos, n, roughness, metallic, albedo, biased_pos);
    finish_initial_wrs(r, biased_pos, p_hat_pl, p_hat_em);
  }
  // RayGen.rgen:344-356: visibility-tested W of the two selected samples
  void finish_initial_wrs(PixelReservoir& r, vec3 biased_pos, float p_hat_pl, float p_hat_em) {
    if (r.pl_idx != INVALID) {
      const Pointlight& pl = L.pointlights[r.pl_idx];
      const vec3 lp = v3(pl.pos[0], pl.pos[1], pl.pos[2]);
      const float pl_dist = length(lp - biased_pos);
      // :346 operator-precedence bug reproduced: pos - (biased_pos / pl_dist)  (quirk 3)
      r.w_pl = shadow_test(biased_pos, lp - biased_pos / pl_dist, pl_dist) * ((r.w_sum_pl / (float)r.pl_m) / p_hat_pl);
    } else {
      r.w_pl = 0.0f;
    }
    if (r.object_instance_idx != INVALID) {
      r.w_em = triangle_shadow_test(r.object_instance_idx, r.triangle_idx, r.bary[0], r.bary[1], biased_pos) * ((r.w_sum_em / (float)r.m_em) / p_hat_em);
    } else {
      r.w_em = 0.0f;
    }
  }

  // RayGen.rgen:361-386
  void combine_pixel_reservoir(PixelReservoir& cur, const PixelReservoir& nb, float& p_hat_pl, float& p_hat_em, vec3 v, vec3 f0,
                               vec3 world_pos, vec3 n, float roughness, float metallic, vec3 albedo, vec3 biased_pos, float rnd_seed) {
    if (nb.pl_idx != INVALID) {
      const Pointlight& pl = L.pointlights[nb.pl_idx];
      const vec3 lp = v3(pl.pos[0], pl.pos[1], pl.pos[2]);
      const float light_dist = length(lp - biased_pos);
      const float p_hat = length(calc_pointlight(pl, v, f0, world_pos, n, roughness, metallic, albedo)) *
                          shadow_test(biased_pos, (lp - biased_pos) / light_dist, light_dist - 0.001f);
      const float weight = (p_hat * (float)nb.pl_m) * nb.w_pl;
      if (update_pointlight_reservoir(cur, nb.pl_idx, weight, rnd_seed)) p_hat_pl = p_hat;
    }
    if (nb.object_instance_idx != INVALID) {
      const vec3 radiance = compute_triangle_radiance(nb.object_instance_idx, nb.triangle_idx, nb.bary[0], nb.bary[1], v, f0, world_pos, n,
                                                      roughness, metallic, albedo, biased_pos, true);
      const float p_hat = length(radiance);
      const float weight = (p_hat * nb.w_em) * (float)nb.m_em;
      if (update_emissive_reservoir(cur, nb.object_instance_idx, nb.triangle_idx, nb.bary[0], nb.bary[1], weight, rnd_seed + 0.283673699f))
        p_hat_em = p_hat;
    }
  }

  // RayGen.rgen:390-409
  void find_neighbour_pixel_coord(float sx, float sy, float radius, float seed_x, float seed_y, float& ox, float& oy) {
    const float W = (float)F.g.w, H = (float)F.g.h;
    float offx = (((gold_noise(fx, fy, fract(seed_x + time() * 0.0011f)) - 0.5f) * 2.0f) * radius) / W;
    float offy = (((gold_noise(fx, fy, fract(seed_y + time() * 0.0006f)) - 0.5f) * 2.0f) * radius) / H;
    int i = 0;
    while ((sx + offx < 0.0f || sy + offy < 0.0f) || (sx + offx > 1.0f || sy + offy > 1.0f)) {
      i++;
      if (i > 64) { offx = 0.0f; offy = 0.0f; break; }  // DEVIATION: the reference loop is unbounded
      const float s = sin_c((float)i * 293.28937f);
      const float qx = (float)(x + (uint32_t)i), qy = (float)(y + (uint32_t)i);
      offx = (((gold_noise(qx, qy, fract((seed_x + s) + time() * 0.0011f)) - 0.5f) * 2.0f) * radius) / W;
      offy = (((gold_noise(qx, qy, fract((seed_y + s) + time() * 0.0006f)) - 0.5f) * 2.0f) * radius) / H;
    }
    ox = sx + offx; oy = sy + offy;
  }

  vec3 prev_normal_texel(int px, int py) const {
    const size_t i = ((size_t)py * F.g.w + px) * 4;
    return v3(f16_to_f32(F.g.prev_normal[i]), f16_to_f32(F.g.prev_normal[i + 1]), f16_to_f32(F.g.prev_normal[i + 2]));
  }
  // texture(sampler_prev_frame_normals, uv): linear filter, repeat addressing (Images.cpp:16-37), full-precision weights
  vec3 prev_normal_bilinear(float u, float v) const {
    const int W = (int)F.g.w, H = (int)F.g.h;
    const float su = u * (float)W - 0.5f, sv = v * (float)H - 0.5f;
    const float fu = floorf(su), fv = floorf(sv);
    const float a = su - fu, b = sv - fv;
    int x0 = (int)fu, y0 = (int)fv;
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = ((x0 % W) + W) % W; x1 = ((x1 % W) + W) % W; y0 = ((y0 % H) + H) % H; y1 = ((y1 % H) + H) % H;
    const vec3 t00 = prev_normal_texel(x0, y0), t10 = prev_normal_texel(x1, y0), t01 = prev_normal_texel(x0, y1), t11 = prev_normal_texel(x1, y1);
    const vec3 top = t00 * (1.0f - a) + t10 * a, bot = t01 * (1.0f - a) + t11 * a;
    return top * (1.0f - b) + bot * b;
  }

  // RayGen.rgen:411-468
  void write_initial_reservoirs(vec3 v, vec3 f0, vec3 ro, vec3 normal, vec3 rma, vec3 albedo, vec3 biased_pos) {
    PixelReservoir pixel; init_pixel_reservoir(pixel);
    do_initial_wrs(pixel, v, f0, ro, normal, rma.x, rma.y, albedo, biased_pos);
    PixelReservoir temporal; init_pixel_reservoir(temporal);
    const uint32_t W = F.g.w, H = F.g.h;
    if (!F.ubo.prev_frame_invalid && F.settings.num_temporal_resamples != 0) {
      const size_t px = (size_t)y * W + x;
      const float vx = f16_to_f32(F.g.motion[2 * px]), vy = f16_to_f32(F.g.motion[2 * px + 1]);
      const float hx = (fx + 0.5f) / (float)W + vx, hy = (fy + 0.5f) / (float)H + vy;
      if ((hx > 0.0f && hy > 0.0f) && (hx < 1.0f && hy < 1.0f)) {
        const vec3 neighbour_normal = prev_normal_bilinear(hx, hy);
        if (dot(neighbour_normal, normal) > F.settings.normal_threshold) {
          int nx = (int)(hx * (float)W), ny = (int)(hy * (float)H);
          if (nx > (int)W - 1) nx = (int)W - 1;  // DEVIATION: guard (cannot trigger for hx < 1 except by rounding)
          if (ny > (int)H - 1) ny = (int)H - 1;
          PixelReservoir history = F.reservoirs[(size_t)ny * W + nx];
          const float cap = (float)F.settings.m_cap;
          if ((float)history.pl_m > cap * (float)pixel.pl_m) {
            history.w_sum_pl *= (cap * (float)pixel.pl_m) / (float)history.pl_m;
            history.pl_m = F.settings.m_cap * pixel.pl_m;
          }
          if ((float)history.m_em > cap * (float)pixel.m_em) {
            history.w_sum_em *= (cap * (float)pixel.m_em) / (float)history.m_em;
            history.m_em = F.settings.m_cap * pixel.m_em;
          }
          float p_hat_em = 0.0f, p_hat_pl = 0.0f;
          combine_pixel_reservoir(temporal, pixel, p_hat_pl, p_hat_em, v, f0, ro, normal, rma.x, rma.y, albedo, biased_pos, 0.846746997f);
          combine_pixel_reservoir(temporal, history, p_hat_pl, p_hat_em, v, f0, ro, normal, rma.x, rma.y, albedo, biased_pos, 0.0f);
          temporal.m_em = history.m_em + pixel.m_em;
          temporal.pl_m = history.pl_m + pixel.pl_m;
          temporal.w_em = p_hat_em > 0.0f ? (temporal.w_sum_em / (float)temporal.m_em) / p_hat_em : 0.0f;
          temporal.w_pl = p_hat_pl > 0.0f ? (temporal.w_sum_pl / (float)temporal.pl_m) / p_hat_pl : 0.0f;
        }
      }
    }
    // quirk 4: without usable history the EMPTY temporal reservoir is written
    F.reservoirs[(size_t)W * H + (size_t)y * W + x] = temporal;
  }

  // RayGen.rgen:470-506
  PixelReservoir compute_and_write_spatial(vec3 v, vec3 f0, vec3 ro, vec3 normal, vec3 rma, vec3 albedo, vec3 biased_pos) {
    const uint32_t W = F.g.w, H = F.g.h;
    const float pcx = fx / (float)W, pcy = fy / (float)H;
    PixelReservoir r; init_pixel_reservoir(r);
    const size_t read_offset = (size_t)W * H;
    const PixelReservoir pixel = F.reservoirs[read_offset + (size_t)y * W + x];
    float p_hat_pl = 0.0f, p_hat_em = 0.0f;
    combine_pixel_reservoir(r, pixel, p_hat_pl, p_hat_em, v, f0, ro, normal, rma.x, rma.y, albedo, biased_pos, 0.0f);
    r.m_em += pixel.m_em;
    r.pl_m += pixel.pl_m;
    const float seed = ((biased_pos.x + biased_pos.y) + biased_pos.z) + F.ubo.delta_time;
    for (int i = 0; i < F.settings.num_spatial_resamples; ++i) {
      const float s = sin_c((float)i * 13477.3793f);
      float cx, cy;
      find_neighbour_pixel_coord(pcx, pcy, F.settings.neighbour_radius, seed + s, -seed + s, cx, cy);
      int nx = (int)((float)W * cx), ny = (int)((float)H * cy);
      if (nx > (int)W - 1) nx = (int)W - 1;  // DEVIATION: quirk 10, the reference can index one past the edge
      if (ny > (int)H - 1) ny = (int)H - 1;
      if (nx < 0) nx = 0;
      if (ny < 0) ny = 0;
      const vec3 neighbour_normal = prev_normal_texel(nx, ny);  // quirk 5: PREVIOUS-frame normals
      if (dot(neighbour_normal, normal) < F.settings.normal_threshold) continue;
      const PixelReservoir nb = F.reservoirs[read_offset + (size_t)ny * W + nx];
      combine_pixel_reservoir(r, nb, p_hat_pl, p_hat_em, v, f0, ro, normal, rma.x, rma.y, albedo, biased_pos, sin_c((float)i * 234473.344f));
      r.m_em += nb.m_em;
      r.pl_m += nb.pl_m;
    }
    r.w_em = p_hat_em > 0.0f ? (r.w_sum_em / p_hat_em) / (float)r.m_em : 0.0f;
    r.w_pl = p_hat_pl > 0.0f ? (r.w_sum_pl / p_hat_pl) / (float)r.pl_m : 0.0f;
    F.reservoirs[(size_t)y * W + x] = r;
    return r;
  }

  // RayGen.rgen:508-528
  vec3 radiance_from_reservoirs(const PixelReservoir& r, vec3 v, vec3 f0, vec3 ro, vec3 normal, vec3 rma, vec3 albedo, vec3 biased_pos) {
    vec3 radiance = v3(0.0f);
    if (L.num_pointlights > 0 && r.pl_idx != INVALID) {
      const Pointlight& pl = L.pointlights[r.pl_idx];
      const vec3 lp = v3(pl.pos[0], pl.pos[1], pl.pos[2]);
      const float light_dist = length(lp - biased_pos);
      vec3 plr = calc_pointlight(pl, v, f0, ro, normal, rma.x, rma.y, albedo) * shadow_test(biased_pos, (lp - biased_pos) / light_dist, light_dist);
      plr *= r.w_pl;
      radiance += plr;
    }
    if (!sc.emissive_idx.empty() && r.object_instance_idx != INVALID) {
      const vec3 er = compute_triangle_radiance(r.object_instance_idx, r.triangle_idx, r.bary[0], r.bary[1], v, f0, ro, normal, rma.x, rma.y,
                                                albedo, biased_pos, true);
      radiance += er * r.w_em;
    }
    return radiance;
  }

  // RayGen.rgen:530-570; pass = push.initial_reservoir_pass
  void main(uint32_t pass) {
    const uint32_t W = F.g.w, H = F.g.h;
    const size_t px = (size_t)y * W + x;
    const float u = (fx + 0.5f) / (float)W, vv = (fy + 0.5f) / (float)H;
    const vec3 albedo = v3(snorm8_to_f32(F.g.albedo[4 * px]), snorm8_to_f32(F.g.albedo[4 * px + 1]), snorm8_to_f32(F.g.albedo[4 * px + 2]));
    const vec3 rma = v3(f16_to_f32(F.g.rma[4 * px]), f16_to_f32(F.g.rma[4 * px + 1]), f16_to_f32(F.g.rma[4 * px + 2]));
    const uint32_t mat_flags = F.g.flags[px];
    float* out = &F.out_rgba[4 * px];
    if (pass == 0) { float* dz = &F.direct_rgba[4 * px]; dz[0] = dz[1] = dz[2] = 0.0f; dz[3] = 1.0f; }
    if (mat_flags & MAT_FLAG_EMISSIVE) {
      if (pass == 0) { out[0] = albedo.x; out[1] = albedo.y; out[2] = albedo.z; out[3] = 1.0f; }
      return;
    }
    if (F.g.prim[px] == INVALID) {
      // DEVIATION (quirk 12): sky pixels (depth 1, normal 0) are masked instead of shaded with NaN frames;
      // the reference result for them is 0 (albedo 0 multiplies every term). Reservoirs are written empty.
      PixelReservoir e; init_pixel_reservoir(e);
      if (pass == 1) F.reservoirs[(size_t)W * H + px] = e;
      else { F.reservoirs[px] = e; out[0] = out[1] = out[2] = 0.0f; out[3] = 1.0f; }
      return;
    }
    const vec3 normal = v3(f16_to_f32(F.g.normal[4 * px]), f16_to_f32(F.g.normal[4 * px + 1]), f16_to_f32(F.g.normal[4 * px + 2]));
    const vec3 ro = world_pos_from_depth(F.ivp, F.g.depth[px], u, vv);
    const vec3 cam = v3(F.ubo.cam_pos[0], F.ubo.cam_pos[1], F.ubo.cam_pos[2]);
    const vec3 v = normalize(cam - ro);
    const vec3 f0 = mix(v3(0.04f), albedo, rma.y);
    const vec3 biased_pos = ro + normal * (0.01f + gmax(length(cam - ro) * 0.001f, 0.01f));
    const bool restir = !(F.settings.flags & RT_NO_RESTIR);
    if (pass == 1) {
      if (restir) write_initial_reservoirs(v, f0, ro, normal, rma, albedo, biased_pos);
      return;
    }
    const vec3 ld = v3(L.dir_light.dir[0], L.dir_light.dir[1], L.dir_light.dir[2]);
    vec3 light = calc_directional_light(L, v, f0, normal, rma.x, rma.y, albedo) * shadow_test(biased_pos, ld, 1000.0f);
    if (restir) {
      const PixelReservoir r = compute_and_write_spatial(v, f0, ro, normal, rma, albedo, biased_pos);
      light += radiance_from_reservoirs(r, v, f0, ro, normal, rma, albedo, biased_pos);
    }
    float* dir = &F.direct_rgba[4 * px];
    dir[0] = light.x; dir[1] = light.y; dir[2] = light.z; dir[3] = 1.0f;
    vec3 colour_sum = light;
    if (!(F.settings.flags & RT_NO_BOUNCE)) colour_sum += bounce_diffuse(biased_pos, normal, albedo);
    out[0] = colour_sum.x; out[1] = colour_sum.y; out[2] = colour_sum.z; out[3] = 1.0f;
  }
};

// RT::RecordRenderCmdBuf (Raytracing.cpp:227-267): pass 1 over all pixels, barrier, pass 2
// one pass of RT::RecordRenderCmdBuf over rows [y0,y1): pass 1 = initial + temporal reservoirs,
// pass 0 = spatial reuse, shading and bounce rays (Raytracing.cpp:250-260)
inline void render_reference_pass(Frame& F, uint32_t pass, uint32_t y0, uint32_t y1, int threads) {
  F.ivp = inv_view_proj(F.ubo);
  parallel_rows(y0, y1, threads, [&](uint32_t y) {
    uint64_t nc = 0, na = 0;
    for (uint32_t x = 0; x < F.g.w; ++x) { Invocation inv(F, x, y); inv.main(pass); nc += inv.n_closest; na += inv.n_any; }
    F.rays_closest += nc; F.rays_any += na;
  });
}

inline void render_reference(Frame& F, uint32_t y0, uint32_t y1, int threads) {
  F.ivp = inv_view_proj(F.ubo);
  for (uint32_t pass = 1;; pass = 0) {
    parallel_rows(y0, y1, threads, [&](uint32_t y) {
      uint64_t nc = 0, na = 0;
      for (uint32_t x = 0; x < F.g.w; ++x) { Invocation inv(F, x, y); inv.main(pass); nc += inv.n_closest; na += inv.n_any; }
      F.rays_closest += nc; F.rays_any += na;
    });
    if (pass == 0) break;
  }
}

}  // namespace snkor
