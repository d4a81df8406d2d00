// shading.cuh — device restatement of the reference's shading / sampling GLSL.
//
//   LightBuffers.glsl:65-157   FresnelSchlick, DistributionGGX, GeometrySmith, CookTorranceBRDF,
//                              CalcDirectionalLight, CalcPointlight
//   Util.glsl:23-57, 77-82     hash / floatConstruct / random, gold_noise
//   RayGen.rgen:123-136        CreateCoordinateSystem, UniformSampleHemisphere
//   RayGen.rgen:241-282        TriangleShadowTest / ComputeTriangleRadiance (point on an emissive triangle)
//   RayClosestHit.rchit:87-142 hit shading
// Expression order follows the GLSL text (left-to-right float32); the translation unit that includes
// this header must be compiled with -fmad=false (vecmath.cuh).
#pragma once
#include <cstring>
#include "context.h"
#include "vecmath.cuh"

namespace snk {

constexpr float PI_GLSL = 3.1415926f;                    // LightBuffers.glsl:29
constexpr float PHI_GLSL = 1.61803398874989484820459f;   // Util.glsl:77
constexpr uint32_t INVALID_IDX = 0xFFFFFFFFu;

// ---- Util.glsl ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash1(uint32_t x) {  // Util.glsl:23-30
  x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u);
  return x;
}
__device__ __forceinline__ float random_f(float x) {  // Util.glsl:43-57
  const uint32_t m = (hash1(__float_as_uint(x)) & 0x007FFFFFu) | 0x3F800000u;
  return __uint_as_float(m) - 1.0f;
}
__device__ __forceinline__ float gold_noise(float px, float py, float seed) {  // Util.glsl:79-82
  const float ax = px * PHI_GLSL, ay = py * PHI_GLSL;
  const float dx = ax - px, dy = ay - py;
  const float dist = sqrtf(dx * dx + dy * dy);
  const float t = tan_c(dist * seed) * px;
  return fract(t);
}
__device__ __forceinline__ uint32_t to_uint(float f) { return (f >= 0.0f && f < 4294967296.0f) ? (uint32_t)f : 0u; }

// ---- LightBuffers.glsl ---------------------------------------------------------------------
__device__ __forceinline__ vec3 fresnel_schlick(float cos_theta, vec3 f0) {
  const float p = pow5(gclamp(1.0f - cos_theta, 0.0f, 1.0f));
  return f0 + (v3(1.0f) - f0) * p;
}
__device__ __forceinline__ float distribution_ggx(vec3 h, vec3 n, float roughness) {
  const float a = roughness * roughness;
  const float a2 = a * a;
  const float n_dot_h = gmax(dot(n, h), 0.0f);
  const float n_dot_h2 = n_dot_h * n_dot_h;
  float denom = (n_dot_h2 * (a2 - 1.0f) + 1.0f);
  denom = PI_GLSL * denom * denom;
  return a2 / denom;
}
__device__ __forceinline__ float geometry_schlick_ggx(float n_dot_v, float roughness) {
  const float r = (roughness + 1.0f);
  const float k = (r * r) / 8.0f;
  return n_dot_v / (n_dot_v * (1.0f - k) + k);
}
__device__ __forceinline__ float geometry_smith(vec3 v, vec3 l, vec3 n, float roughness) {
  const float n_dot_v = gmax(dot(n, v), 0.0f);
  const float n_dot_l = gmax(dot(n, l), 0.0f);
  const float ggx2 = geometry_schlick_ggx(n_dot_v, roughness);
  const float ggx1 = geometry_schlick_ggx(n_dot_l, roughness);
  return ggx2 * ggx1;
}
#ifndef SNK_REF_INLINE
#define SNK_REF_INLINE 0  // 1: the reference-mode shading helpers are inlined (A/B: profiles/r06n)
#endif
#if SNK_REF_INLINE
#define SNK_REF_FN __device__ __forceinline__
#else
#define SNK_REF_FN static __device__ __noinline__
#endif
SNK_REF_FN vec3 cook_torrance_brdf(vec3 light_pos, vec3 surface_pos, vec3 v, vec3 f0, vec3 n, float roughness, float metallic) {
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
SNK_REF_FN vec3 calc_directional_light(vec3 ldir, vec3 lcol, vec3 v, vec3 f0, vec3 n, float roughness, float metallic, vec3 albedo) {
  const vec3 l = ldir;
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
  return ((kd * albedo / PI_GLSL + specular) * n_dot_l) * lcol;
}
struct PointLightRec { vec3 colour, pos; float atten_exp, atten_linear, atten_constant; };
__device__ __forceinline__ PointLightRec load_pointlight(const SnkLightData* __restrict__ L, uint32_t idx) {
  const float4* p = reinterpret_cast<const float4*>(&L->pointlights[idx]);
  const float4 c = __ldg(p), q = __ldg(p + 1), a = __ldg(p + 2);
  PointLightRec r;
  r.colour = v3(c.x, c.y, c.z); r.pos = v3(q.x, q.y, q.z);
  r.atten_exp = a.x; r.atten_linear = a.y; r.atten_constant = a.z;
  return r;
}
__device__ __forceinline__ vec3 calc_pointlight(const PointLightRec& light, vec3 v, vec3 f0, vec3 world_pos, vec3 n, float roughness, float metallic, vec3 albedo) {
  const vec3 frag_to_light = light.pos - world_pos;
  const float distance = length(frag_to_light);
  const float attenuation = (light.atten_constant + light.atten_linear * distance) + light.atten_exp * (distance * distance);
  const vec3 brdf = cook_torrance_brdf(light.pos, world_pos, v, f0, n, roughness, metallic);
  return ((brdf * albedo) * light.colour) / attenuation;
}

// ---- RayGen.rgen:123-136 -------------------------------------------------------------------------
__device__ __forceinline__ void create_coordinate_system(vec3 n, vec3& nt, vec3& nb) {
  if (fabsf(n.x) > fabsf(n.y)) nt = v3(n.z, 0.0f, -n.x) / length2(n.x, n.z);
  else nt = v3(0.0f, -n.z, n.y) / length2(n.y, n.z);
  nb = cross(n, nt);
}
__device__ __forceinline__ vec3 uniform_sample_hemisphere(float r1, float r2) {
  const float sin_theta = sqrtf(1.0f - r1 * r1);
  const float phi = (2.0f * PI_GLSL) * r2;
  return v3(sin_theta * cos_c(phi), r1, sin_theta * sin_c(phi));
}

// ---- scene fetch helpers ---------------------------------------------------------------------------
__device__ __forceinline__ vec3 ld3(const float* __restrict__ p, uint32_t i) { return v3(__ldg(p + 3ull * i), __ldg(p + 3ull * i + 1), __ldg(p + 3ull * i + 2)); }
__device__ __forceinline__ void tri_indices(const SceneDev& D, const SnkInstanceData& in, uint32_t prim, uint32_t& a, uint32_t& b, uint32_t& c) {
  const uint32_t base = 3u * (prim + in.mesh_buffer_index_offset / 3u);  // RayClosestHit.rchit:91
  a = __ldg(D.indices + base) + in.mesh_buffer_vertex_offset;
  b = __ldg(D.indices + base + 1) + in.mesh_buffer_vertex_offset;
  c = __ldg(D.indices + base + 2) + in.mesh_buffer_vertex_offset;
}
__device__ __forceinline__ SnkInstanceData load_instance(const SceneDev& D, uint32_t i) {
  const uint32_t* p = reinterpret_cast<const uint32_t*>(D.instances + i);
  SnkInstanceData r;
  r.transform_idx = __ldg(p); r.mesh_buffer_index_offset = __ldg(p + 1); r.mesh_buffer_vertex_offset = __ldg(p + 2);
  r.material_idx = __ldg(p + 3); r.num_mesh_indices = __ldg(p + 4);
  return r;
}

// ---- albedo texture (RayClosestHit.rchit:100-116, GBuffer.frag:56-58) ---------------------------------------------------
// bilinear, repeat addressing, base level, sRGB decode through a 256-entry table (CPU statement: oracle/snk_scene.h
// sample_texture_rgb — same float32 expressions in the same order)
__device__ __forceinline__ uint32_t tex_wrap(float f, uint32_t n) {
  long long i = (long long)f % (long long)n;
  return (uint32_t)(i < 0 ? i + (long long)n : i);
}
static __device__ __noinline__ vec3 sample_texture_rgb(const SceneDev& D, uint32_t tex, float s, float t) {
  const TexDev td = D.textures[tex];
  const float u = s * (float)td.w - 0.5f, v = t * (float)td.h - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  const uint32_t x0 = tex_wrap(fu, td.w), x1 = tex_wrap(fu + 1.0f, td.w), y0 = tex_wrap(fv, td.h), y1 = tex_wrap(fv + 1.0f, td.h);
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  const uchar4 t00 = __ldg(td.texels + (size_t)y0 * td.w + x0), t10 = __ldg(td.texels + (size_t)y0 * td.w + x1);
  const uchar4 t01 = __ldg(td.texels + (size_t)y1 * td.w + x0), t11 = __ldg(td.texels + (size_t)y1 * td.w + x1);
  vec3 c;
  if (td.srgb) {
    const float* L = D.srgb_lut;
    c.x = ((__ldg(L + t00.x) * w00 + __ldg(L + t10.x) * w10) + __ldg(L + t01.x) * w01) + __ldg(L + t11.x) * w11;
    c.y = ((__ldg(L + t00.y) * w00 + __ldg(L + t10.y) * w10) + __ldg(L + t01.y) * w01) + __ldg(L + t11.y) * w11;
    c.z = ((__ldg(L + t00.z) * w00 + __ldg(L + t10.z) * w10) + __ldg(L + t01.z) * w01) + __ldg(L + t11.z) * w11;
  } else {
    c.x = (((float)t00.x / 255.0f * w00 + (float)t10.x / 255.0f * w10) + (float)t01.x / 255.0f * w01) + (float)t11.x / 255.0f * w11;
    c.y = (((float)t00.y / 255.0f * w00 + (float)t10.y / 255.0f * w10) + (float)t01.y / 255.0f * w01) + (float)t11.y / 255.0f * w11;
    c.z = (((float)t00.z / 255.0f * w00 + (float)t10.z / 255.0f * w10) + (float)t01.z / 255.0f * w01) + (float)t11.z / 255.0f * w11;
  }
  return c;
}
// material albedo times the albedo texture at the interpolated texcoord (weights b0, b1, b2 of vertices ia, ib, ic)
__device__ __forceinline__ vec3 albedo_at(const SceneDev& D, const SnkMaterial* m, uint32_t ia, uint32_t ib, uint32_t ic, float b0, float b1, float b2) {
  const float4 alb = __ldg(reinterpret_cast<const float4*>(m->albedo));
  vec3 albedo = v3(alb.x, alb.y, alb.z);
  const uint32_t tex = __ldg(&m->albedo_tex_idx);
  if (tex != INVALID_IDX && tex < D.num_textures && D.uvs != nullptr && D.textures[tex].w != 0u) {
    const float tu = (__ldg(D.uvs + 2ull * ia) * b0 + __ldg(D.uvs + 2ull * ib) * b1) + __ldg(D.uvs + 2ull * ic) * b2;
    const float tv = (__ldg(D.uvs + 2ull * ia + 1) * b0 + __ldg(D.uvs + 2ull * ib + 1) * b1) + __ldg(D.uvs + 2ull * ic + 1) * b2;
    albedo = albedo * sample_texture_rgb(D, tex, tu, tv);
  }
  return albedo;
}

// GBuffer.vert:49-64 + GBuffer.frag:48-52: normal map through the interpolated per-vertex TBN (CPU statement: Scene::mapped_normal)
static __device__ __noinline__ bool mapped_normal(const SceneDev& D, const SnkMaterial* m, const float* __restrict__ M, uint32_t ia, uint32_t ib, uint32_t ic,
                                                  float b0, float b1, float b2, vec3& out) {
  const uint32_t tex = __ldg(&m->normal_tex_idx);
  if (tex == INVALID_IDX || tex >= D.num_textures || D.textures[tex].w == 0u || D.tangents == nullptr || D.uvs == nullptr) return false;
  const uint32_t vi[3] = {ia, ib, ic};
  const float w[3] = {b0, b1, b2};
  vec3 T = v3(0.0f), B = v3(0.0f), N = v3(0.0f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const vec3 tv = ld3(D.tangents, vi[k]), nv = ld3(D.normals, vi[k]);
    vec3 t = normalize(v3((M[0] * tv.x + M[4] * tv.y) + M[8] * tv.z, (M[1] * tv.x + M[5] * tv.y) + M[9] * tv.z, (M[2] * tv.x + M[6] * tv.y) + M[10] * tv.z));
    const vec3 n = normalize(v3((M[0] * nv.x + M[4] * nv.y) + M[8] * nv.z, (M[1] * nv.x + M[5] * nv.y) + M[9] * nv.z, (M[2] * nv.x + M[6] * nv.y) + M[10] * nv.z));
    t = normalize(t - n * dot(t, n));
    const vec3 b = cross(n, t);
    T = k == 0 ? t * w[k] : T + t * w[k]; B = k == 0 ? b * w[k] : B + b * w[k]; N = k == 0 ? n * w[k] : N + n * w[k];
  }
  const float tu = (__ldg(D.uvs + 2ull * ia) * b0 + __ldg(D.uvs + 2ull * ib) * b1) + __ldg(D.uvs + 2ull * ic) * b2;
  const float tv2 = (__ldg(D.uvs + 2ull * ia + 1) * b0 + __ldg(D.uvs + 2ull * ib + 1) * b1) + __ldg(D.uvs + 2ull * ic + 1) * b2;
  const vec3 c = sample_texture_rgb(D, tex, tu, tv2);
  const vec3 tn = normalize(c * 2.0f - v3(1.0f));
  out = normalize((T * tn.x + B * tn.y) + N * tn.z);
  return true;
}

// point on an emissive triangle: shared head of RayGen.rgen:241-251 and :258-268.
// bary = (bx, by, 1-bx-by) weights of (ti.x, ti.y, ti.z) — NOT the hit-attribute convention.
struct TriPoint { vec3 p, n, emissive; };
SNK_REF_FN TriPoint triangle_point(const SceneDev& D, uint32_t obj, uint32_t tri, float bx, float by) {
  const SnkInstanceData in = load_instance(D, obj);
  uint32_t ia, ib, ic;
  tri_indices(D, in, tri, ia, ib, ic);
  const float bz = (1.0f - bx) - by;
  vec3 p = (bx * ld3(D.positions, ia) + by * ld3(D.positions, ib)) + bz * ld3(D.positions, ic);
  vec3 n = (bx * ld3(D.normals, ia) + by * ld3(D.normals, ib)) + bz * ld3(D.normals, ic);
  TriPoint r;
  r.p = mul_point(D.transforms + 16ull * in.transform_idx, p);
  r.n = normalize(mul_mat3r(D.normal_mats + 9ull * in.transform_idx, n));
  const float4 alb = __ldg(reinterpret_cast<const float4*>(D.materials[in.material_idx].albedo));
  r.emissive = v3(alb.x, alb.y, alb.z);
  return r;
}
// unshadowed part of ComputeTriangleRadiance (RayGen.rgen:270-273)
__device__ __forceinline__ vec3 triangle_radiance_unshadowed(const TriPoint& tp, vec3 v, vec3 f0, vec3 world_pos, vec3 n, float roughness,
                                                             float metallic, vec3 albedo) {
  const vec3 brdf = cook_torrance_brdf(tp.p, world_pos, v, f0, n, roughness, metallic);
  const vec3 dlt = world_pos - tp.p;
  return ((tp.emissive * albedo) * brdf) / fabsf(dot(dlt, dlt));
}
// the shadow ray both TriangleShadowTest (:253-255) and ComputeTriangleRadiance (:276-278) shoot
__device__ __forceinline__ void triangle_shadow_ray(const TriPoint& tp, vec3 from, vec3& dir, float& dist) {
  const vec3 target = tp.p + tp.n * 0.01f;
  dist = length(from - target);
  dir = (target - from) / dist;
}

// RayClosestHit.rchit:87-142 up to (not including) the dir-light shadow ray.
// Returns the unshadowed colour; the caller multiplies by the visibility of `shadow_*` when needs_shadow.
struct HitShade {
  vec3 colour, normal, shadow_o;
  float distance;
  uint32_t mat_flags;
  bool needs_shadow;
};
__device__ __forceinline__ HitShade shade_hit(const SceneDev& D, vec3 ldir, vec3 lcol, vec3 rd, float t, float u, float v, uint32_t prim, uint32_t inst) {
  HitShade o;
  o.needs_shadow = false; o.normal = v3(0.0f); o.shadow_o = v3(0.0f); o.mat_flags = 0;
  if (t < 0.0f) { o.colour = v3(0.0f); o.distance = -1.0f; return o; }  // RayMiss.rmiss:22-24
  const float b0 = (1.0f - u) - v, b1 = u, b2 = v;
  const SnkInstanceData in = load_instance(D, inst);
  uint32_t ia, ib, ic;
  tri_indices(D, in, prim, ia, ib, ic);
  const vec3 p = (ld3(D.positions, ia) * b0 + ld3(D.positions, ib) * b1) + ld3(D.positions, ic) * b2;
  vec3 n = (ld3(D.normals, ia) * b0 + ld3(D.normals, ib) * b1) + ld3(D.normals, ic) * b2;
  const vec3 pos = mul_point(D.transforms + 16ull * in.transform_idx, p);
  n = normalize(mul_mat3r(D.normal_mats + 9ull * in.transform_idx, n));
  const SnkMaterial* m = D.materials + in.material_idx;
  const vec3 albedo = albedo_at(D, m, ia, ib, ic, b0, b1, b2);
  const float roughness = __ldg(&m->roughness), metallic = __ldg(&m->metallic);
  o.mat_flags = __ldg(&m->flags);
  o.distance = t;
  if (o.mat_flags & SNKPT_MAT_FLAG_EMISSIVE) { o.colour = albedo; return o; }
  const vec3 vv = normalize(-rd);
  const vec3 f0 = mix(v3(0.04f), albedo, metallic);
  o.shadow_o = pos + n * 0.01f;
  o.colour = calc_directional_light(ldir, lcol, vv, f0, n, roughness, metallic, albedo);
  o.normal = n;
  o.needs_shadow = true;
  return o;
}

// ---- reservoirs (SoA planes of 16 B) ---------------------------------------------------------------------
struct Reservoir {
  float w_sum_pl, w_pl; uint32_t pl_idx, pl_m;
  uint32_t obj, tri; float bx, by;
  float w_sum_em, w_em; uint32_t m_em;
};
__device__ __forceinline__ void init_reservoir(Reservoir& r) {  // RayGen.rgen:284-294 (+ defined values for the fields it leaves unset)
  r.pl_idx = INVALID_IDX; r.w_sum_pl = 0.0f; r.pl_m = 0; r.w_pl = 0.0f;
  r.obj = INVALID_IDX; r.tri = INVALID_IDX; r.w_sum_em = 0.0f; r.m_em = 0; r.w_em = 0.0f; r.bx = 0.0f; r.by = 0.0f;
}
struct ResPlanes { float4 *a, *b, *c; };
__device__ __forceinline__ Reservoir load_reservoir(const ResPlanes& pl, size_t i) {
  const float4 a = pl.a[i], b = pl.b[i], c = pl.c[i];
  Reservoir r;
  r.w_sum_pl = a.x; r.w_pl = a.y; r.pl_idx = __float_as_uint(a.z); r.pl_m = __float_as_uint(a.w);
  r.obj = __float_as_uint(b.x); r.tri = __float_as_uint(b.y); r.bx = b.z; r.by = b.w;
  r.w_sum_em = c.x; r.w_em = c.y; r.m_em = __float_as_uint(c.z);
  return r;
}
__device__ __forceinline__ void store_reservoir(const ResPlanes& pl, size_t i, const Reservoir& r) {
  pl.a[i] = make_float4(r.w_sum_pl, r.w_pl, __uint_as_float(r.pl_idx), __uint_as_float(r.pl_m));
  pl.b[i] = make_float4(__uint_as_float(r.obj), __uint_as_float(r.tri), r.bx, r.by);
  pl.c[i] = make_float4(r.w_sum_em, r.w_em, __uint_as_float(r.m_em), 0.0f);
}

// ---- packed surface record (PixelPlanes::surf) -----------------------------------------------------------------------------
// Spatial reuse looks at 3-8 RANDOM pixels within 30 px; through the G-buffer planes that is one 32-byte sector each for prim,
// flags, normal, hit distance, depth, albedo and rma — the stage was bound by those gathers (ncu r04r: 35 % of issue slots busy,
// L1 hit rate 33 %, 2 GB of L2 -> L1 traffic per 1080p frame). The record repeats exactly those bits in ONE sector:
//   a.x normal.x | normal.y << 16 (f16)   a.y normal.z | material flags << 16   a.z hit distance (f32)   a.w depth (f32)
//   b.x albedo (4 x snorm8)               b.y roughness | metallic << 16 (f16)   b.z 1 = a surface was hit   b.w 0
struct SurfRec { uint4 a, b; };
__device__ __forceinline__ void pack_surf(uint4* surf, size_t px, ushort4 normal, uint32_t flags, float hit_t, float depth, char4 albedo, ushort4 rma, bool hit) {
  uint32_t alb;
  memcpy(&alb, &albedo, 4);
  surf[2 * px] = make_uint4((uint32_t)normal.x | ((uint32_t)normal.y << 16), (uint32_t)normal.z | (flags << 16), __float_as_uint(hit_t), __float_as_uint(depth));
  surf[2 * px + 1] = make_uint4(alb, (uint32_t)rma.x | ((uint32_t)rma.y << 16), hit ? 1u : 0u, 0u);
}
__device__ __forceinline__ SurfRec load_surf(const uint4* __restrict__ surf, size_t px) { return SurfRec{__ldg(surf + 2 * px), __ldg(surf + 2 * px + 1)}; }
__device__ __forceinline__ bool surf_shaded(const SurfRec& r) { return r.b.z != 0u && !((r.a.y >> 16) & SNKPT_MAT_FLAG_EMISSIVE); }
__device__ __forceinline__ vec3 surf_normal(const SurfRec& r) {
  return v3(h2f((uint16_t)(r.a.x & 0xFFFFu)), h2f((uint16_t)(r.a.x >> 16)), h2f((uint16_t)(r.a.y & 0xFFFFu)));
}

}  // namespace snk
