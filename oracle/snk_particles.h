// oracle/snk_particles.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// CPU restatement of the reference's particle collision pass, the "next" row N4 of SURVEY §8(f): the second consumer of the
// acceleration structures of the path-tracing pass.
//   Core/res/shaders/Particle.glsl:1-36          Particle (3 x vec4), CellKeyEntry, GetCellKey, GetParticleCellPos, DT
//   Core/res/shaders/Particle.comp:27-105        Init, GenerateCellKey, (bitonic sort), TryGenerateCellKeyStartIdx
//   Core/res/shaders/ParticleUpdate.rgen:37-229  particle / particle collisions, 3 collision rays, gravity ray, offset constraint
//   Core/res/shaders/ParticleUpdate.rchit:44-72  contact normal / position / velocity of a hit
//   Core/src/scene/ParticleSystem.cpp:224-299    three sub-steps per frame: keys, sort, start indices, update, copy back
//   Core/headers/scene/ParticleSystem.h:16-24    PhysicsParamsUBO defaults
// PARITY UNPINNED (no reference run exists). Decisions where the reference is implementation-defined:
//   * the rays use gl_RayFlagsTerminateOnFirstHitEXT with a closest-hit shader: WHICH hit the driver reports is unspecified.
//     Here (and in the CUDA kernel) it is the closest hit — one of the allowed outcomes and the only order-independent one;
//   * the bitonic network leaves the order of equal keys unspecified; here the entries of one cell key are in ascending
//     particle index (a stable sort of the key list, which is generated in index order). The order only changes the
//     float summation order of the impulses;
//   * TryGenerateCellKeyStartIdx reads k[-1] for invocation 0 (out of bounds, result unused because of the `== 0` test): guarded;
//   * velocity / speed with speed == 0 is NaN in the reference (the three collision rays then have NaN directions and report no
//     hit on any driver that tests `t < tmax`); the rays are skipped instead. DEVIATION, same result.
// Float32 arithmetic in the order of the GLSL text (-ffp-contract=off); sin / cos in Init in float64 rounded once.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "snk_scene.h"

namespace snkor {

struct Particle { float position_radius[4], velocity[4], colour[4]; };  // Particle.glsl:3-7 (velocity.w: 0 asleep, 1 awake)
static_assert(sizeof(Particle) == 48, "Particle");
struct CellKeyEntry { uint32_t key, ptcl_idx; };
struct PhysicsParams {  // ParticleSystem.h:16-24
  float particle_restitution = 0.2f, timestep = 1.f / 120.f, friction_coefficient = 0.5f, repulsion_factor = 0.08f, penetration_slop = 0.005f,
        restitution_slop = 0.05f, sleep_threshold = 0.05f;
};
constexpr float PTCL_DT = (float)(1.0 / 120.0);  // Particle.glsl:1

inline uint32_t ptcl_cell_key(int cx, int cy, int cz, uint32_t num_particles) {  // Particle.glsl:14-16 (int32 wrap-around, unsigned %)
  const uint32_t h = ((uint32_t)cx * 73856093u) ^ ((uint32_t)cy * 10357u) ^ ((uint32_t)cz * 83492791u);
  return h % num_particles;
}
inline void ptcl_cell_pos(const Particle& p, int& x, int& y, int& z) {  // ivec3(pos * 5): truncation
  x = (int)(p.position_radius[0] * 5.0f); y = (int)(p.position_radius[1] * 5.0f); z = (int)(p.position_radius[2] * 5.0f);
}
inline bool ptcl_asleep(const Particle& p) { return p.velocity[3] < 0.5f; }

inline void ptcl_init(std::vector<Particle>& P) {  // Particle.comp:27-49
  const uint32_t n = (uint32_t)P.size();
  for (uint32_t i = 0; i < n; ++i) {
    Particle& p = P[i];
    const int bx = 100, by = 100;
    const vec3 grid = v3((float)(i % bx), (float)((i % (uint32_t)(bx * by)) / bx), (float)(i / (uint32_t)(bx * by)));
    const vec3 q = (grid - v3((float)bx, 0.0f, 100.0f) * 0.5f) * 0.5f;
    p.position_radius[0] = q.x + -100.0f; p.position_radius[1] = q.y + 20.0f; p.position_radius[2] = q.z + 190.0f; p.position_radius[3] = 0.15f;
    const float fi = (float)i;
    p.velocity[0] = 0.1f * sin_c(fi * 8.34534f); p.velocity[1] = 0.1f * cos_c(fi * 5.34342f); p.velocity[2] = 0.1f * sin_c(fi * 2.3435f);
    p.velocity[3] = 1.0f;
    const float third = (float)n / 3.0f;
    vec3 col = v3(0.0f, 0.0f, 1.0f);
    if ((float)i > third * 2.0f) col = v3(1.0f, 0.0f, 0.0f);
    else if ((float)i > third) col = v3(0.0f, 1.0f, 0.0f);
    p.colour[0] = col.x; p.colour[1] = col.y; p.colour[2] = col.z; p.colour[3] = 0.0f;
  }
}

struct PtclPayload { vec3 contact_normal, contact_velocity, contact_position; float distance; };

// traceRayEXT + ParticleUpdate.rchit / .rmiss; closest hit (see header)
inline PtclPayload ptcl_trace(const Scene& sc, vec3 o, vec3 d, float tmax, float timestep) {
  PtclPayload pl{v3(0.0f), v3(0.0f), v3(0.0f), -1.0f};
  if (!(tmax > 0.0f)) return pl;
  Ray r{{o.x, o.y, o.z}, 0.0f, {d.x, d.y, d.z}, tmax};
  const Hit h = sc.trace(r, false);
  if (h.t < 0.0f) return pl;
  const float b0 = (1.0f - h.u) - h.v, b1 = h.u, b2 = h.v;
  const InstanceData& in = sc.instances[h.inst];
  uint32_t ia, ib, ic; sc.tri_indices(in, h.prim, ia, ib, ic);
  vec3 n = (sc.nrm(ia) * b0 + sc.nrm(ib) * b1) + sc.nrm(ic) * b2;
  const vec3 p = (sc.pos(ia) * b0 + sc.pos(ib) * b1) + sc.pos(ic) * b2;
  n = normalize(mul(sc.normal_matrix[in.transform_idx], n));
  const vec3 contact = o + d * h.t;
  pl.distance = h.t;
  pl.contact_normal = n;
  pl.contact_position = contact;
  pl.contact_velocity = (contact - mul_point(sc.prev_transforms[in.transform_idx], p)) / timestep;
  return pl;
}

inline void ptcl_build_grid(const std::vector<Particle>& P, std::vector<CellKeyEntry>& keys, std::vector<uint32_t>& start) {
  const uint32_t n = (uint32_t)P.size();
  keys.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    int x, y, z; ptcl_cell_pos(P[i], x, y, z);
    keys[i] = CellKeyEntry{ptcl_cell_key(x, y, z, n), i};
  }
  std::stable_sort(keys.begin(), keys.end(), [](const CellKeyEntry& a, const CellKeyEntry& b) { return a.key < b.key; });
  start.assign(n, 0xFFFFFFFFu);
  for (uint32_t i = 0; i < n; ++i)
    if (i == 0 || keys[i].key != keys[i - 1].key) start[keys[i].key] = i;
}

inline Particle ptcl_update_one(const Scene& sc, const std::vector<Particle>& P, const std::vector<CellKeyEntry>& keys, const std::vector<uint32_t>& start,
                                const PhysicsParams& prm, uint32_t idx, uint64_t& rays) {
  const Particle& p = P[idx];
  const uint32_t n = (uint32_t)P.size();
  const vec3 pos = v3(p.position_radius[0], p.position_radius[1], p.position_radius[2]);
  const float radius = p.position_radius[3];
  const vec3 vel = v3(p.velocity[0], p.velocity[1], p.velocity[2]);
  const bool asleep = ptcl_asleep(p);
  // ---- main(): three collision rays (rgen:180-197)
  PtclPayload cand{v3(0.0f), v3(0.0f), v3(0.0f), -1.0f};
  float highest = 0.0f;
  const float speed = length(vel);
  if (!asleep && speed > 0.0f) {
    const vec3 vdir = vel / speed;
    const vec3 vright = fabsf(dot(vdir, v3(0.0f, 1.0f, 0.0f))) > 0.99f ? cross(vdir, v3(0.0f, 0.0f, 1.0f)) : cross(vdir, v3(0.0f, 1.0f, 0.0f));
    const vec3 vup = cross(vright, vdir);
    const vec3 dirs[3] = {vdir, vright, vup};
    const float lens[3] = {speed * PTCL_DT + radius * 2.0f, radius * 2.0f, radius * 2.0f};
    for (int k = 0; k < 3; ++k) {  // TraceCollisionRay (rgen:96-117): from the "back" of the particle
      const PtclPayload pl = ptcl_trace(sc, pos - dirs[k] * radius, dirs[k], lens[k], prm.timestep);
      ++rays;
      const float w = length(pl.contact_velocity - vel);
      if (pl.distance > 0.0f && w > highest) { highest = w; cand = pl; }
    }
  }
  // ---- HandleParticleOnParticleCollision (rgen:37-94)
  vec3 total_offset = v3(0.0f), total_impulse = v3(0.0f), colour_mix = v3(p.colour[0], p.colour[1], p.colour[2]);
  {
    int cx, cy, cz; ptcl_cell_pos(p, cx, cy, cz);
    vec3 local_impulse = v3(0.0f), local_offset = v3(0.0f);
    for (int xo = -1; xo <= 1; ++xo)
      for (int yo = -1; yo <= 1; ++yo)
        for (int zo = -1; zo <= 1; ++zo) {
          const uint32_t key = ptcl_cell_key(cx + xo, cy + yo, cz + zo, n);
          const uint32_t s0 = start[key];
          if (s0 == 0xFFFFFFFFu) continue;
          for (uint32_t i = s0; i < n; ++i) {
            if (keys[i].ptcl_idx == idx) continue;
            if (keys[i].key != key) break;
            const Particle& o = P[keys[i].ptcl_idx];
            const float min_dist = o.position_radius[3] + radius;
            const vec3 between = pos - v3(o.position_radius[0], o.position_radius[1], o.position_radius[2]);
            const float dist = length(between);
            if (dist < (min_dist - 0.005f)) {
              const vec3 cn = between / dist;
              const vec3 rel = vel - v3(o.velocity[0], o.velocity[1], o.velocity[2]);
              const float rs = length(rel);
              vec3 impulse = v3(0.0f);
              if (rs > prm.restitution_slop) impulse += (((cn * rs) * dot(cn, (-rel) / rs)) * 2.0f) * prm.particle_restitution;
              local_offset += (cn * (min_dist - dist)) * 0.5f;
              impulse += ((cn / PTCL_DT) * ((min_dist - dist) - 0.005f)) * prm.repulsion_factor;
              local_impulse += impulse;
              colour_mix = mix(colour_mix, v3(o.colour[0], o.colour[1], o.colour[2]), prm.timestep);
            }
          }
        }
    const float i2 = fabsf(dot(local_impulse, local_impulse));
    if ((asleep && i2 > prm.sleep_threshold * prm.sleep_threshold) || !asleep) { total_offset += local_offset; total_impulse += local_impulse; }
  }
  // ---- CalculateGravity (rgen:147-165)
  {
    const PtclPayload g = ptcl_trace(sc, pos, v3(0.0f, -1.0f, 0.0f), radius + 0.001f, prm.timestep);
    ++rays;
    total_impulse.y += (-9.8f * PTCL_DT) * (g.distance < 0.0f ? 1.0f : 0.0f);
  }
  if (cand.distance > 0.0f) {  // rgen:204-212
    const float restitution = 0.2f;
    const vec3 rel = vel - cand.contact_velocity;
    const float rs = length(rel);
    const float overlap = gmax(radius - length(cand.contact_position - pos), 0.0f);
    total_offset += cand.contact_normal * overlap;
    total_impulse += ((cand.contact_normal * rs) * dot((-rel) / rs, cand.contact_normal)) * 2.0f * (0.5f + restitution * 0.5f);
  }
  vec3 new_vel = vel + total_impulse;
  total_offset += new_vel * PTCL_DT;
  // ---- ConstrainOffsetWithScene (rgen:119-145)
  {
    const float ol = length(total_offset);
    if (ol < 0.00001f) total_offset = v3(0.0f);
    else {
      const vec3 od = total_offset / ol;
      const PtclPayload c = ptcl_trace(sc, pos, od, ol + radius, prm.timestep);
      ++rays;
      if (c.distance > 0.0f) total_offset = od * gmax(c.distance - radius * 0.9f, 0.0f) + (c.contact_normal * radius) * 0.05f;
    }
  }
  // ---- friction + sleep (rgen:219-241)
  const float mass = 1.0f;
  float ke = (0.5f * mass) * dot(new_vel, new_vel);
  const float loss = (mass * length(total_offset)) * prm.friction_coefficient;
  ke -= gmin(loss, ke);
  const float v2 = fabsf((2.0f * ke) / mass);
  if (fabsf(dot(new_vel, new_vel)) > 0.0001f) new_vel = normalize(new_vel) * sqrtf(v2);
  Particle out = p;
  out.velocity[3] = v2 > prm.sleep_threshold * prm.sleep_threshold ? 1.0f : 0.0f;
  const float awake = asleep ? 0.0f : 1.0f;
  out.position_radius[0] = pos.x + total_offset.x * awake; out.position_radius[1] = pos.y + total_offset.y * awake;
  out.position_radius[2] = pos.z + total_offset.z * awake; out.position_radius[3] = radius;
  out.velocity[0] = new_vel.x * awake; out.velocity[1] = new_vel.y * awake; out.velocity[2] = new_vel.z * awake;
  out.colour[0] = colour_mix.x; out.colour[1] = colour_mix.y; out.colour[2] = colour_mix.z;
  return out;
}

// ParticleSystem::UpdateParticles: `substeps` x (keys, sort, start indices, update, copy back); returns the rays traced
inline uint64_t ptcl_update(const Scene& sc, std::vector<Particle>& P, const PhysicsParams& prm, int substeps, int threads) {
  uint64_t total = 0;
  std::vector<CellKeyEntry> keys;
  std::vector<uint32_t> start;
  std::vector<Particle> out(P.size());
  for (int s = 0; s < substeps; ++s) {
    ptcl_build_grid(P, keys, start);
    const uint32_t n = (uint32_t)P.size();
    if (threads < 1) threads = 1;
    std::vector<uint64_t> rays(threads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
      pool.emplace_back([&, t]() { for (uint32_t i = t; i < n; i += threads) out[i] = ptcl_update_one(sc, P, keys, start, prm, i, rays[t]); });
    for (auto& th : pool) th.join();
    for (uint64_t r : rays) total += r;
    P.swap(out);
  }
  return total;
}

}  // namespace snkor
