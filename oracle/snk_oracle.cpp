// oracle/snk_oracle.cpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// C entry points over the CPU restatement (snk_scene.h, snk_gbuffer.h,
// snk_raygen.h) so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs can drive it through ctypes. Nothing in
// snake-vk_b200/ may link, import or execute this library.
// PARITY UNPINNED: the reference (orangefrog45/Snake-VK) has no tests, golden
// vectors or runnable headless path for this pass; the only
// external anchors are glm (vendored under /root/reference/Core/extern/glm), used by
// oracle/_ref to check the matrix helpers, and the reference's shader TEXT: constants,
// block layouts and the pure shading functions executed by tools/glsl_exec.py
// (tests/golden/reference_*.json, reference_shader_vectors.npz).
#include <chrono>
#include <cstdio>
#include <cstring>

#include "snk_northstar.h"
#include "snk_particles.h"
#include "snk_taa.h"

using namespace snkor;

extern "C" {

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API void* snkor_scene_create(const float* positions, const float* normals, const float* uvs, uint32_t num_vertices,
                                    const uint32_t* indices, uint32_t num_indices, const InstanceData* instances,
                                    uint32_t num_instances, const float* transforms, const float* prev_transforms,
                                    uint32_t num_transforms, const Material* materials, uint32_t num_materials,
                                    const void* light_data, size_t light_bytes, const uint32_t* emissive_idx,
                                    uint32_t num_emissive) {
  auto* sc = new Scene();
  sc->positions.assign(positions, positions + 3 * (size_t)num_vertices);
  if (normals) sc->normals.assign(normals, normals + 3 * (size_t)num_vertices);
  else sc->normals.assign(3 * (size_t)num_vertices, 0.0f);
  if (uvs) sc->uvs.assign(uvs, uvs + 2 * (size_t)num_vertices);
  sc->indices.assign(indices, indices + num_indices);
  sc->instances.assign(instances, instances + num_instances);
  sc->transforms.resize(num_transforms);
  sc->prev_transforms.resize(num_transforms);
  memcpy(sc->transforms.data(), transforms, sizeof(mat4) * num_transforms);
  memcpy(sc->prev_transforms.data(), prev_transforms ? prev_transforms : transforms, sizeof(mat4) * num_transforms);
  sc->materials.assign(materials, materials + num_materials);
  memset(&sc->lights, 0, sizeof(LightData));
  memcpy(&sc->lights, light_data, light_bytes < sizeof(LightData) ? light_bytes : sizeof(LightData));
  sc->emissive_idx.assign(emissive_idx, emissive_idx + num_emissive);
  sc->finalize();
  return sc;
}
ORACLE_API void snkor_scene_destroy(void* s) { delete (Scene*)s; }
// albedo texture `idx` (RGBA8, `srgb` != 0: R8G8B8A8_SRGB) — snk_scene.h:Texture
ORACLE_API void snkor_scene_set_texture(void* s, uint32_t idx, uint32_t w, uint32_t h, const uint8_t* rgba, uint32_t srgb) {
  Scene& sc = *(Scene*)s;
  if (sc.textures.size() <= idx) sc.textures.resize(idx + 1);
  Texture& t = sc.textures[idx];
  t.w = w; t.h = h; t.srgb = srgb;
  t.rgba.assign(rgba, rgba + 4 * (size_t)w * h);
}
ORACLE_API void snkor_scene_set_tangents(void* s, const float* tangents, uint32_t num_vertices) {
  ((Scene*)s)->tangents.assign(tangents, tangents + 3 * (size_t)num_vertices);
}
ORACLE_API void snkor_sample_texture(void* s, uint32_t idx, float u, float v, float* out3) {
  const vec3 c = sample_texture_rgb(((Scene*)s)->textures[idx], u, v);
  out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}

// flags: bit0 = any-hit, bit1 = brute force (no BVH)
ORACLE_API void snkor_trace(void* s, const Ray* rays, uint64_t n, uint32_t flags, Hit* hits, int threads) {
  const Scene& sc = *(Scene*)s;
  const bool any = flags & 1u, brute = flags & 2u;
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t]() {
      for (uint64_t i = t; i < n; i += threads) hits[i] = brute ? sc.trace_brute(rays[i], any) : sc.trace(rays[i], any);
    });
  for (auto& th : pool) th.join();
}

// ---- frame object -------------------------------------------------------------------
ORACLE_API void* snkor_frame_create(void* scene, uint32_t w, uint32_t h) {
  auto* f = new Frame();
  f->sc = (Scene*)scene;
  f->resize(w, h);
  return f;
}
ORACLE_API void snkor_frame_destroy(void* f) { delete (Frame*)f; }
// animated scenes: the frame state (reservoirs, G-buffer history) lives on while the scene buffers of the next frame replace the old ones
ORACLE_API void snkor_frame_set_scene(void* f, void* scene) { ((Frame*)f)->sc = (Scene*)scene; }
ORACLE_API void snkor_frame_set(void* fp, const CommonUBO* cur, const CommonUBO* prev, const RtSettings* st) {
  Frame& f = *(Frame*)fp;
  f.ubo = *cur; f.prev_ubo = *prev;
  if (st) f.settings = *st;
}
ORACLE_API void snkor_frame_gbuffer(void* fp, float jx, float jy, uint32_t y0, uint32_t y1, int threads) {
  Frame& f = *(Frame*)fp;
  if (y1 == 0) y1 = f.g.h;
  fill_gbuffer(*f.sc, f.ubo, f.prev_ubo, jx, jy, f.g, y0, y1, threads);
}
ORACLE_API void snkor_frame_render(void* fp, uint32_t y0, uint32_t y1, int threads) {
  Frame& f = *(Frame*)fp;
  if (y1 == 0) y1 = f.g.h;
  if (f.settings.mode == 1) render_northstar(f, y0, y1, threads);
  else render_reference(f, y0, y1, threads);
}
ORACLE_API void snkor_frame_set_halo(void* fp, uint32_t halo) { ((Frame*)fp)->ns_halo = halo; }
ORACLE_API void snkor_frame_render_pass(void* fp, uint32_t pass, uint32_t y0, uint32_t y1, int threads) {
  Frame& f = *(Frame*)fp;
  if (y1 == 0) y1 = f.g.h;
  render_reference_pass(f, pass, y0, y1, threads);
}
// what: same numbering as SNKPT_AUX_* (include/snkpt.h)
ORACLE_API int snkor_frame_read(void* fp, uint32_t what, void* dst, size_t bytes) {
  Frame& f = *(Frame*)fp;
  const size_t n = (size_t)f.g.w * f.g.h;
  const void* src = nullptr; size_t sz = 0;
  switch (what) {
    case 0: src = f.g.prim.data(); sz = 4 * n; break;
    case 1: src = f.g.inst.data(); sz = 4 * n; break;
    case 2: src = f.g.depth.data(); sz = 4 * n; break;
    case 3: src = f.g.normal.data(); sz = 8 * n; break;
    case 4: src = f.g.albedo.data(); sz = 4 * n; break;
    case 5: src = f.g.rma.data(); sz = 8 * n; break;
    case 6: src = f.g.motion.data(); sz = 4 * n; break;
    case 7: src = f.g.flags.data(); sz = 2 * n; break;
    case 8: src = f.g.prev_normal.data(); sz = 8 * n; break;
    case 9: src = f.g.prev_depth.data(); sz = 4 * n; break;
    case 10: src = f.reservoirs.data(); sz = 48 * n; break;
    case 11: src = f.reservoirs.data() + n; sz = 48 * n; break;
    case 14: src = f.g.hit_t.data(); sz = 4 * n; break;
    case 15: src = f.g.bary.data(); sz = 8 * n; break;
    case 16: src = f.out_rgba.data(); sz = 16 * n; break;
    case 18: src = f.direct_rgba.data(); sz = 16 * n; break;
    default: return -1;
  }
  if (bytes < sz) return -1;
  memcpy(dst, src, sz);
  return 0;
}
ORACLE_API int snkor_frame_write_reservoirs(void* fp, uint32_t half, const PixelReservoir* src, uint32_t first, uint32_t count) {
  Frame& f = *(Frame*)fp;
  const size_t n = (size_t)f.g.w * f.g.h;
  if (half > 1 || (size_t)first + count > n) return -1;
  memcpy(f.reservoirs.data() + half * n + first, src, sizeof(PixelReservoir) * count);
  return 0;
}
ORACLE_API void snkor_frame_ray_counts(void* fp, uint64_t* closest, uint64_t* any, int reset) {
  Frame& f = *(Frame*)fp;
  *closest = f.rays_closest.load(); *any = f.rays_any.load();
  if (reset) { f.rays_closest = 0; f.rays_any = 0; }
}
ORACLE_API void snkor_inv_view_proj(const CommonUBO* u, float* out16) {
  mat4 m = inv_view_proj(*u);
  memcpy(out16, m.m, 64);
}

// ---- known-answer helpers ---------------------------------------------------------------
ORACLE_API uint32_t snkor_hash(uint32_t x) { return hash1(x); }
ORACLE_API float snkor_random(float x) { return random_f(x); }
ORACLE_API float snkor_gold_noise(float px, float py, float seed) { return gold_noise_f(px, py, seed); }
ORACLE_API uint16_t snkor_f32_to_f16(float f) { return f32_to_f16(f); }
ORACLE_API float snkor_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
ORACLE_API int snkor_f32_to_snorm8(float f) { return f32_to_snorm8(f); }
// out: fresnel(3) ggx(1) smith(1) cook_torrance(3)
ORACLE_API void snkor_brdf_eval(const float* n, const float* v, const float* l, const float* f0, float roughness, float metallic,
                                float* out8) {
  const vec3 N = v3(n[0], n[1], n[2]), V = v3(v[0], v[1], v[2]), Lv = v3(l[0], l[1], l[2]), F0 = v3(f0[0], f0[1], f0[2]);
  const vec3 H = normalize(V + Lv);
  const vec3 f = fresnel_schlick(gmax(dot(H, V), 0.0f), F0);
  out8[0] = f.x; out8[1] = f.y; out8[2] = f.z;
  out8[3] = distribution_ggx(H, N, roughness);
  out8[4] = geometry_smith(V, Lv, N, roughness);
  // light position = surface (origin) + l
  const vec3 ct = cook_torrance_brdf(Lv, v3(0.0f), V, F0, N, roughness, metallic);
  out8[5] = ct.x; out8[6] = ct.y; out8[7] = ct.z;
}
ORACLE_API void snkor_calc_pointlight(const Pointlight* pl, const float* v, const float* f0, const float* wp, const float* n, float roughness,
                                      float metallic, const float* albedo, float* out3) {
  vec3 r = calc_pointlight(*pl, v3(v[0], v[1], v[2]), v3(f0[0], f0[1], f0[2]), v3(wp[0], wp[1], wp[2]), v3(n[0], n[1], n[2]), roughness,
                           metallic, v3(albedo[0], albedo[1], albedo[2]));
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
ORACLE_API void snkor_calc_dirlight(const void* light_data, const float* v, const float* f0, const float* n, float roughness, float metallic,
                                    const float* albedo, float* out3) {
  LightData L; memset(&L, 0, sizeof(L)); memcpy(&L, light_data, 112);
  vec3 r = calc_directional_light(L, v3(v[0], v[1], v[2]), v3(f0[0], f0[1], f0[2]), v3(n[0], n[1], n[2]), roughness, metallic,
                                  v3(albedo[0], albedo[1], albedo[2]));
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
// scripted WRS: runs update_pointlight_reservoir over `n` (idx, weight) candidates with the
// uniform numbers in `u` (RayGen.rgen:215-226 with the RNG injected), returns final state
ORACLE_API void snkor_wrs_pointlight_script(void* fp, const uint32_t* idx, const float* weight, const float* u, uint32_t n,
                                            PixelReservoir* out) {
  Frame& f = *(Frame*)fp;
  RngScript s{u, n, 0};
  Invocation inv(f, 1, 1, &s);
  PixelReservoir r; Invocation::init_pixel_reservoir(r);
  for (uint32_t i = 0; i < n; ++i) inv.update_pointlight_reservoir(r, idx[i], weight[i], 0.0f);
  *out = r;
}
// the same streaming updates with the shader's OWN random numbers (gold_noise of the pixel, the frame time and the weight):
// pinned against the reference functions executed from the shader text (tests/golden/reference_shader_vectors.npz)
ORACLE_API void snkor_wrs_native(void* fp, uint32_t x, uint32_t y, const uint32_t* idx, const uint32_t* tri, const float* bary2,
                                 const float* weight, const float* seed, uint32_t n, PixelReservoir* out_pl, PixelReservoir* out_em,
                                 uint8_t* selected_pl, uint8_t* selected_em) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, x, y, nullptr);
  PixelReservoir a; Invocation::init_pixel_reservoir(a);
  PixelReservoir b; Invocation::init_pixel_reservoir(b);
  for (uint32_t i = 0; i < n; ++i) {
    selected_pl[i] = inv.update_pointlight_reservoir(a, idx[i], weight[i], seed[i]) ? 1 : 0;
    selected_em[i] = inv.update_emissive_reservoir(b, idx[i], tri[i], bary2[2 * i], bary2[2 * i + 1], weight[i], seed[i]) ? 1 : 0;
  }
  *out_pl = a; *out_em = b;
}
// the point-light loop of DoInitialWRS on an explicit surface (not a G-buffer pixel)
ORACLE_API float snkor_ris_pointlights(void* fp, uint32_t x, uint32_t y, const float* v, const float* f0, const float* world_pos, const float* n,
                                       float roughness, float metallic, const float* albedo, PixelReservoir* out) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, x, y, nullptr);
  PixelReservoir r; Invocation::init_pixel_reservoir(r);
  const float p_hat = inv.ris_pointlights(r, v3(v[0], v[1], v[2]), v3(f0[0], f0[1], f0[2]), v3(world_pos[0], world_pos[1], world_pos[2]),
                                          v3(n[0], n[1], n[2]), roughness, metallic, v3(albedo[0], albedo[1], albedo[2]));
  *out = r;
  return p_hat;
}
// the emissive-triangle loop of DoInitialWRS on an explicit surface
ORACLE_API float snkor_ris_emissives(void* fp, uint32_t x, uint32_t y, const float* v, const float* f0, const float* world_pos, const float* n,
                                     float roughness, float metallic, const float* albedo, PixelReservoir* out) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, x, y, nullptr);
  PixelReservoir r; Invocation::init_pixel_reservoir(r);
  const vec3 wp = v3(world_pos[0], world_pos[1], world_pos[2]);
  const float p_hat = inv.ris_emissives(r, v3(v[0], v[1], v[2]), v3(f0[0], f0[1], f0[2]), wp, v3(n[0], n[1], n[2]), roughness, metallic,
                                        v3(albedo[0], albedo[1], albedo[2]), wp);
  *out = r;
  return p_hat;
}
// CombinePixelReservoir with every shadow ray visible: the merge arithmetic alone
ORACLE_API void snkor_combine_visible(void* fp, uint32_t x, uint32_t y, PixelReservoir* cur, const PixelReservoir* nb, float* p_hat2, const float* v,
                                      const float* f0, const float* world_pos, const float* n, float roughness, float metallic, const float* albedo,
                                      const float* biased_pos, float rnd_seed) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, x, y, nullptr);
  f.force_visible = true;
  inv.combine_pixel_reservoir(*cur, *nb, p_hat2[0], p_hat2[1], v3(v[0], v[1], v[2]), v3(f0[0], f0[1], f0[2]), v3(world_pos[0], world_pos[1], world_pos[2]),
                              v3(n[0], n[1], n[2]), roughness, metallic, v3(albedo[0], albedo[1], albedo[2]),
                              v3(biased_pos[0], biased_pos[1], biased_pos[2]), rnd_seed);
  f.force_visible = false;
}
// CalculateRadianceFromReservoirs with every shadow ray visible
ORACLE_API void snkor_reservoir_radiance_visible(void* fp, const PixelReservoir* r, const float* v, const float* f0, const float* ro, const float* n,
                                                 const float* rma, const float* albedo, const float* biased_pos, float* out3) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, 0, 0, nullptr);
  f.force_visible = true;
  const vec3 c = inv.radiance_from_reservoirs(*r, v3(v[0], v[1], v[2]), v3(f0[0], f0[1], f0[2]), v3(ro[0], ro[1], ro[2]), v3(n[0], n[1], n[2]),
                                              v3(rma[0], rma[1], rma[2]), v3(albedo[0], albedo[1], albedo[2]),
                                              v3(biased_pos[0], biased_pos[1], biased_pos[2]));
  f.force_visible = false;
  out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
ORACLE_API void snkor_find_neighbour(void* fp, uint32_t x, uint32_t y, float sx, float sy, float radius, float seed_x, float seed_y, float* out2) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, x, y, nullptr);
  inv.find_neighbour_pixel_coord(sx, sy, radius, seed_x, seed_y, out2[0], out2[1]);
}
// the tail of DoInitialWRS with visibility forced: W of both samples and the two shadow rays it asks for (rays7: origin, dir, max)
ORACLE_API uint32_t snkor_finish_initial_wrs(void* fp, PixelReservoir* r, const float* biased_pos, float p_hat_pl, float p_hat_em, float* rays7x2) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, 0, 0, nullptr);
  f.force_visible = true; f.n_forced = 0;
  inv.finish_initial_wrs(*r, v3(biased_pos[0], biased_pos[1], biased_pos[2]), p_hat_pl, p_hat_em);
  f.force_visible = false;
  memcpy(rays7x2, f.forced_rays, sizeof(float) * 14);
  return f.n_forced;
}
// RayClosestHit.rchit main() for a given hit, visibility forced; out9 = colour3 normal3 distance roughness mat_flags, ray7 = the shadow ray
ORACLE_API uint32_t snkor_shade_hit_visible(void* fp, uint32_t inst, uint32_t prim, float u, float v, float t, const float* rd, float* out9, float* ray7) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, 0, 0, nullptr);
  Hit h{}; h.t = t; h.u = u; h.v = v; h.prim = prim; h.inst = inst;
  f.force_visible = true; f.n_forced = 0;
  inv.payload = Payload{};
  inv.shade_closest_hit(h, v3(rd[0], rd[1], rd[2]));
  f.force_visible = false;
  const Payload& p = inv.payload;
  out9[0] = p.colour.x; out9[1] = p.colour.y; out9[2] = p.colour.z; out9[3] = p.normal.x; out9[4] = p.normal.y; out9[5] = p.normal.z;
  out9[6] = p.distance; out9[7] = p.roughness; out9[8] = (float)p.mat_flags;
  memcpy(ray7, f.forced_rays[0], sizeof(float) * 7);
  return f.n_forced;
}
// ComputeTriangleRadiance(…, shadow_test = false) for a point on an emissive triangle seen from an explicit surface
ORACLE_API void snkor_compute_triangle_radiance(void* fp, uint32_t obj, uint32_t tri, float bx, float by, const float* v, const float* f0,
                                                const float* world_pos, const float* n, float roughness, float metallic, const float* albedo,
                                                float* out3) {
  Frame& f = *(Frame*)fp;
  Invocation inv(f, 0, 0, nullptr);
  const vec3 wp = v3(world_pos[0], world_pos[1], world_pos[2]);
  const vec3 r = inv.compute_triangle_radiance(obj, tri, bx, by, v3(v[0], v[1], v[2]), v3(f0[0], f0[1], f0[2]), wp, v3(n[0], n[1], n[2]), roughness,
                                               metallic, v3(albedo[0], albedo[1], albedo[2]), wp, false);
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
// run one pixel of one pass with scripted random numbers (KAT for the temporal / spatial state machines)
ORACLE_API void snkor_pixel_scripted(void* fp, uint32_t x, uint32_t y, uint32_t pass, const float* u, uint32_t n) {
  Frame& f = *(Frame*)fp;
  f.ivp = inv_view_proj(f.ubo);
  RngScript s{u, n, 0};
  Invocation inv(f, x, y, n ? &s : nullptr);
  inv.main(pass);
  f.rays_closest += inv.n_closest; f.rays_any += inv.n_any;
}
ORACLE_API void snkor_sizes(uint32_t* out8) {
  out8[0] = sizeof(InstanceData); out8[1] = sizeof(Material); out8[2] = sizeof(LightData); out8[3] = sizeof(CommonUBO);
  out8[4] = sizeof(PixelReservoir); out8[5] = sizeof(Ray); out8[6] = sizeof(Hit); out8[7] = sizeof(RtSettings);
}
// particle collision pass (snk_particles.h): Particle.comp Init, `substeps` sub-steps of ParticleSystem::UpdateParticles
ORACLE_API void snkor_particles_init(Particle* p, uint32_t n) {
  std::vector<Particle> v(n);
  ptcl_init(v);
  memcpy(p, v.data(), sizeof(Particle) * n);
}
ORACLE_API uint64_t snkor_particles_update(void* scene, Particle* p, uint32_t n, const PhysicsParams* prm, int substeps, int threads) {
  std::vector<Particle> v(p, p + n);
  const uint64_t rays = ptcl_update(*(Scene*)scene, v, prm ? *prm : PhysicsParams(), substeps, threads);
  memcpy(p, v.data(), sizeof(Particle) * n);
  return rays;
}
ORACLE_API void snkor_particles_grid(const Particle* p, uint32_t n, CellKeyEntry* keys, uint32_t* start) {
  std::vector<Particle> v(p, p + n);
  std::vector<CellKeyEntry> k;
  std::vector<uint32_t> s;
  ptcl_build_grid(v, k, s);
  memcpy(keys, k.data(), sizeof(CellKeyEntry) * n);
  memcpy(start, s.data(), 4 * (size_t)n);
}
// TAA_Resolve.comp on raw RGBA16F / RG16F planes (snk_taa.h)
ORACLE_API void snkor_hemisphere_helpers(const float* n3, float r1, float r2, float* nt3, float* nb3, float* sample3) {
  vec3 nt, nb;
  Invocation::create_coordinate_system(v3(n3[0], n3[1], n3[2]), nt, nb);
  const vec3 s = Invocation::uniform_sample_hemisphere(r1, r2);
  nt3[0] = nt.x; nt3[1] = nt.y; nt3[2] = nt.z; nb3[0] = nb.x; nb3[1] = nb.y; nb3[2] = nb.z;
  sample3[0] = s.x; sample3[1] = s.y; sample3[2] = s.z;
}
ORACLE_API void snkor_calc_velocity(const float* cur4, const float* old4, float* out2) {
  calc_velocity(vec4{cur4[0], cur4[1], cur4[2], cur4[3]}, vec4{old4[0], old4[1], old4[2], old4[3]}, out2[0], out2[1]);
}
ORACLE_API uint32_t snkor_ptcl_cell_key(int cx, int cy, int cz, uint32_t num_particles) { return ptcl_cell_key(cx, cy, cz, num_particles); }
ORACLE_API void snkor_ptcl_cell_pos(const float* pos3, int* out3) {
  Particle p{};
  p.position_radius[0] = pos3[0]; p.position_radius[1] = pos3[1]; p.position_radius[2] = pos3[2];
  ptcl_cell_pos(p, out3[0], out3[1], out3[2]);
}
ORACLE_API void snkor_taa_clip_aabb(const float* aabb_min, const float* aabb_max, const float* p4, const float* q4, float* out4) {
  const vec4 r = taa_clip_aabb(v3(aabb_min[0], aabb_min[1], aabb_min[2]), v3(aabb_max[0], aabb_max[1], aabb_max[2]),
                               vec4{p4[0], p4[1], p4[2], p4[3]}, vec4{q4[0], q4[1], q4[2], q4[3]});
  out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
ORACLE_API void snkor_taa_resolve(uint32_t w, uint32_t h, const uint16_t* cur, const uint16_t* velocity, const uint16_t* history,
                                  const uint16_t* prev_vel, uint16_t* out) {
  taa_resolve(TaaImages{w, h, cur, velocity, history, prev_vel, out});
}
ORACLE_API int snkor_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
