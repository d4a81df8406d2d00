// northstar.cuh — device helpers shared by the north-star stages (northstar.cu: path vertices; restir_ns.cu: ReSTIR DI).
// CPU statement: oracle/snk_northstar.h (same expression order, -fmad=false).
#pragma once
#include "shading.cuh"

namespace snk {

__device__ __forceinline__ uint32_t pcg_hash(uint32_t v) {
  const uint32_t state = v * 747796405u + 2891336453u;
  const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}
__device__ __forceinline__ float ns_rand(uint32_t pixel, uint32_t frame, uint32_t dim) {
  const uint32_t h = pcg_hash(pcg_hash(pcg_hash(pixel) + frame) + dim);
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

struct NsSurf {
  vec3 p, n, albedo;
  float roughness, metallic;
  uint32_t flags, ntri;
  float area;
};

__device__ __forceinline__ NsSurf ns_surface_inl(const SceneDev& D, uint32_t inst, uint32_t prim, float u, float v) {
  NsSurf s;
  const float b0 = (1.0f - u) - v;
  const SnkInstanceData in = load_instance(D, inst);
  uint32_t ia, ib, ic;
  tri_indices(D, in, prim, ia, ib, ic);
  const vec3 pa = ld3(D.positions, ia), pb = ld3(D.positions, ib), pc = ld3(D.positions, ic);
  const vec3 p = (pa * b0 + pb * u) + pc * v;
  const vec3 n = (ld3(D.normals, ia) * b0 + ld3(D.normals, ib) * u) + ld3(D.normals, ic) * v;
  const float* M = D.transforms + 16ull * in.transform_idx;
  s.p = mul_point(M, p);
  s.n = normalize(mul_mat3r(D.normal_mats + 9ull * in.transform_idx, n));
  const SnkMaterial* m = D.materials + in.material_idx;
  s.flags = __ldg(&m->flags);
  // emitted radiance is the material albedo alone (RayGen.rgen:270, emissive table); the texture modulates reflectors only
  if (s.flags & SNKPT_MAT_FLAG_EMISSIVE) {
    const float4 alb = __ldg(reinterpret_cast<const float4*>(m->albedo));
    s.albedo = v3(alb.x, alb.y, alb.z);
  } else {
    s.albedo = albedo_at(D, m, ia, ib, ic, b0, u, v);
  }
  s.roughness = gmax(__ldg(&m->roughness), 0.05f);
  s.metallic = __ldg(&m->metallic);
  s.ntri = in.num_mesh_indices / 3u;
  const vec3 wa = mul_point(M, pa), wb = mul_point(M, pb), wc = mul_point(M, pc);
  s.area = 0.5f * length(cross(wb - wa, wc - wa));
  return s;
}
static __device__ __noinline__ NsSurf ns_surface(const SceneDev& D, uint32_t inst, uint32_t prim, float u, float v) { return ns_surface_inl(D, inst, prim, u, v); }

// ---- emissive-triangle table ------------------------------------------------------------------------------------
// Light sampling evaluates a point on 32 random emissive triangles per pixel; through the scene buffers that is ~60
// dependent 4-byte gathers per candidate (instance record, 3 indices, 3 positions, 3 normals, mat4, normal matrix,
// material) — ncu: 81 % excessive sectors, IPC 1.4 on C3 (profiles/r01t_prof_shade.txt). k_build_em_table (restir_ns.cu)
// flattens everything a LIGHT SAMPLE needs about one emissive triangle into ONE record per frame. Round 1 kept the record in
// object space (192 B: vertices, normals, mat4, normal matrix, material) and the lookup re-did the transforms; at 66 M candidates
// per frame on C3 that is 12.7 GB of L2 -> L1 traffic, the stage ran at the L2's bandwidth. Now the record is in WORLD space
// (oracle: ns_light_point — vertices and vertex normals transformed per vertex when the table is built, interpolated at the
// lookup): 96 B = 6 x 128-bit loads, no matrix work per candidate.
constexpr int EM_REC_F4 = 6;
//  r0 wa.xyz area | r1 wb.xyz albedo.r | r2 wc.xyz albedo.g | r3 wna.xyz albedo.b | r4 wnb.xyz ntri | r5 wnc.xyz flags
__device__ __forceinline__ NsSurf ns_surface_em_inl(const SceneDev& D, uint32_t obj, uint32_t tri, float u, float v) {
  const uint32_t base = __ldg(D.em_base_of_obj + obj);  // callers make sure obj is in the emissive list (rs_em_valid)
  const float4* r = D.em_table + (size_t)EM_REC_F4 * (base + tri);
  NsSurf s;
  const float b0 = (1.0f - u) - v;
  const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3), r4 = __ldg(r + 4), r5 = __ldg(r + 5);
  s.p = (v3(r0.x, r0.y, r0.z) * b0 + v3(r1.x, r1.y, r1.z) * u) + v3(r2.x, r2.y, r2.z) * v;
  s.n = normalize((v3(r3.x, r3.y, r3.z) * b0 + v3(r4.x, r4.y, r4.z) * u) + v3(r5.x, r5.y, r5.z) * v);
  s.area = r0.w;
  s.albedo = v3(r1.w, r2.w, r3.w);
  s.roughness = 0.0f; s.metallic = 0.0f;
  s.ntri = __float_as_uint(r4.w);
  s.flags = __float_as_uint(r5.w);
  return s;
}
// a real call for the kernels with many call sites (instruction-cache pressure, profiles/r01c); candidate loops use the _inl form
static __device__ __noinline__ NsSurf ns_surface_em(const SceneDev& D, uint32_t obj, uint32_t tri, float u, float v) { return ns_surface_em_inl(D, obj, tri, u, v); }

__device__ __forceinline__ vec3 ns_bsdf_cos_inl(vec3 n, vec3 wo, vec3 wi, vec3 albedo, float roughness, float metallic) {
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
static __device__ __noinline__ vec3 ns_bsdf_cos(vec3 n, vec3 wo, vec3 wi, vec3 albedo, float roughness, float metallic) {
  return ns_bsdf_cos_inl(n, wo, wi, albedo, roughness, metallic);
}
__device__ __forceinline__ float ns_spec_prob(float metallic) { return 0.25f + 0.5f * metallic; }
__device__ __forceinline__ float ns_pdf_inl(vec3 n, vec3 wo, vec3 wi, float roughness, float metallic) {
  const float ps = ns_spec_prob(metallic);
  const float n_dot_l = gmax(dot(n, wi), 0.0f);
  const vec3 h = normalize(wo + wi);
  const float n_dot_h = gmax(dot(n, h), 0.0f);
  const float v_dot_h = gmax(dot(wo, h), 0.000001f);
  const float d = distribution_ggx(h, n, roughness);
  return (1.0f - ps) * (n_dot_l / PI_GLSL) + ps * ((d * n_dot_h) / (4.0f * v_dot_h));
}
static __device__ __noinline__ float ns_pdf(vec3 n, vec3 wo, vec3 wi, float roughness, float metallic) { return ns_pdf_inl(n, wo, wi, roughness, metallic); }
__device__ __forceinline__ bool ns_sample(vec3 n, vec3 wo, float roughness, float metallic, float u0, float u1, float u2, vec3& wi) {
  vec3 nt, nb;
  create_coordinate_system(n, nt, nb);
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
__device__ __forceinline__ NsLightCounts ns_light_counts(const SceneDev& D) {
  NsLightCounts c;
  const float4 col = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.colour));
  c.has_dir = (col.x > 0.0f || col.y > 0.0f || col.z > 0.0f) ? 1u : 0u;
  c.n_pl = D.lights->num_pointlights;
  c.n_em = D.num_emissive;
  c.n_types = c.has_dir + (c.n_pl ? 1u : 0u) + (c.n_em ? 1u : 0u);
  return c;
}
__device__ __forceinline__ uint32_t ns_pick(float u, uint32_t n) { const uint32_t i = to_uint(u * (float)n); return i >= n ? n - 1 : i; }
__device__ __forceinline__ float ns_nee_pdf_emissive(const NsLightCounts& lc, uint32_t ntri, float area, float dist, vec3 nl, vec3 wi) {
  const float cos_l = fabsf(dot(nl, wi));
  if (!(cos_l > 0.000001f) || !(area > 0.0f)) return 0.0f;
  const float p_pick = (1.0f / (float)lc.n_types) * (1.0f / ((float)lc.n_em * (float)ntri));
  return p_pick * ((dist * dist) / (cos_l * area));
}

__device__ __forceinline__ void put_ray(float4* q, size_t i, vec3 o, float tmin, vec3 d, float tmax) {
  q[2 * i] = make_float4(o.x, o.y, o.z, tmin);
  q[2 * i + 1] = make_float4(d.x, d.y, d.z, tmax);
}
__device__ __forceinline__ void put_dead(float4* q, size_t i) {
  q[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
  q[2 * i + 1] = make_float4(0.f, 1.f, 0.f, -1.0f);
}

// ---- thread -> pixel mapping of the per-pixel north-star stages: a warp covers an 8x4 pixel tile of the band instead of 32
// pixels of one row. The rays a warp appends to the queue (and the paths it lists) then start from a compact patch, and the
// neighbour gathers of the spatial-reuse stage get 2-D locality. Storage stays indexed by pixel, so results do not change.
#ifndef SNK_TILE_PIXELS
#define SNK_TILE_PIXELS 0  // measured (profiles/r02w_variants_tile_pixels.txt): C2 7.19 -> 7.11 ms, C3 14.05 -> 13.96 ms, but C4 39.7 -> 40.2 ms (trace.q2 +20 % at 4K): a wash, rows kept
#endif
__device__ __forceinline__ bool stage_pixel(const FrameConsts& fc, uint32_t t, uint32_t& x, uint32_t& y) {
#if SNK_TILE_PIXELS
  const uint32_t tiles_x = (fc.W + 7u) >> 3;
  const uint32_t tile = t >> 5, lane = t & 31u;
  x = (tile % tiles_x) * 8u + (lane & 7u);
  y = fc.row0 + (tile / tiles_x) * 4u + (lane >> 3);
  return x < fc.W && y < fc.row1;
#else
  x = t % fc.W; y = fc.row0 + t / fc.W;
  return y < fc.row1;
#endif
}
inline uint32_t stage_threads(uint32_t W, uint32_t rows) {
#if SNK_TILE_PIXELS
  return ((W + 7u) / 8u) * ((rows + 3u) / 4u) * 32u;
#else
  return W * rows;
#endif
}

// ---- compact ray-reference queue (trace_q.cu): segments of the north-star frame
constexpr uint32_t SEG_RS_PL = 0u, SEG_RS_EM = 1u, SEG_RS_DIR = 2u, SEG_NS_CLOSEST = 3u, SEG_NS_SHADOW = 4u;
__device__ __forceinline__ uint32_t make_ref(uint32_t seg, size_t slot) { return (seg << REF_SLOT_BITS) | (uint32_t)slot; }
// warp-aggregated append of up to three references per lane (one atomic per warp); every lane of the warp must call it.
// The warp's references land segment by segment, so a traversal warp that takes 32 consecutive entries mostly gets
// rays of one kind from neighbouring pixels.
__device__ __forceinline__ void q_append3(uint32_t* refs, uint32_t* count, bool h0, uint32_t r0, bool h1, uint32_t r1, bool h2, uint32_t r2) {
  const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, h0), m1 = __ballot_sync(0xFFFFFFFFu, h1), m2 = __ballot_sync(0xFFFFFFFFu, h2);
  const uint32_t n0 = __popc(m0), n1 = __popc(m1), tot = n0 + n1 + __popc(m2);
  if (tot == 0u) return;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(count, tot);
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  const uint32_t below = (1u << lane) - 1u;
  if (h0) refs[base + __popc(m0 & below)] = r0;
  if (h1) refs[base + n0 + __popc(m1 & below)] = r1;
  if (h2) refs[base + n0 + n1 + __popc(m2 & below)] = r2;
}

__device__ __forceinline__ void q_append1(uint32_t* list, uint32_t* count, bool h, uint32_t v) {
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, h);
  if (m == 0u) return;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  if (h) list[base + __popc(m & ((1u << lane) - 1u))] = v;
}

// marks a path as ended: only the tmax word of the slot is read back by the next vertex stage
__device__ __forceinline__ void put_dead_dir(float4* q, size_t i) { q[2 * i + 1] = make_float4(0.f, 1.f, 0.f, -1.0f); }

__device__ __forceinline__ void flush_counts(unsigned long long n_shadow, unsigned long long n_closest, unsigned long long* counters) {
  for (int o = 16; o > 0; o >>= 1) {
    n_shadow += __shfl_xor_sync(0xFFFFFFFFu, n_shadow, o);
    n_closest += __shfl_xor_sync(0xFFFFFFFFu, n_closest, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_shadow) atomicAdd(counters, n_shadow);
    if (n_closest) atomicAdd(counters + 1, n_closest);
  }
}

}  // namespace snk
