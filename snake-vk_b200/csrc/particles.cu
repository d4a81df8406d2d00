// particles.cu — the reference's particle collision pass on the path tracer's acceleration structures.
//
// Replaces ParticleSystem::UpdateParticles (Core/src/scene/ParticleSystem.cpp:224-299): per sub-step the cell-key generation
// (Particle.comp:51-55), the key sort, the start-index table (:77-85) and the ray-traced collision update
// (ParticleUpdate.rgen:37-229 with ParticleUpdate.rchit:44-72), then the copy of the result buffer back.
//   * the O(log^2 N) bitonic dispatch chain of the reference (136 dispatches per sub-step for 100 k particles,
//     ParticleSystem.cpp:252-268) becomes 3 stable LSD radix passes over the 24 key bits (the sort kernels of bvh_build.cu):
//     entries of one cell end up in ascending particle index, the order the CPU statement defines;
//   * the <= 5 rays of a particle are traced inline by the warp-synchronous closest-hit tracer (traverse.cuh) — they depend
//     on each other (the constraint ray is built from the impulses of the others), so they cannot be queued;
//   * "terminate on first hit" is resolved to the closest hit (oracle/snk_particles.h explains why).
// CPU statement: oracle/snk_particles.h; this file must match it bit for bit (-fmad=false, same expression order).
#include "kernels.h"
#include "shading.cuh"

namespace snk {

namespace {

constexpr int PT_THREADS = 128;
constexpr int PT_CTAS_PER_SM = 6;
constexpr float PTCL_DT = (float)(1.0 / 120.0);  // Particle.glsl:1

__device__ __forceinline__ uint32_t cell_key(int cx, int cy, int cz, uint32_t n) {  // Particle.glsl:14-16
  const uint32_t h = ((uint32_t)cx * 73856093u) ^ ((uint32_t)cy * 10357u) ^ ((uint32_t)cz * 83492791u);
  return h % n;
}

__global__ void k_ptcl_init(SnkParticle* __restrict__ P, uint32_t n) {  // Particle.comp:27-49
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const vec3 grid = v3((float)(i % 100u), (float)((i % 10000u) / 100u), (float)(i / 10000u));
  const vec3 q = (grid - v3(100.0f, 0.0f, 100.0f) * 0.5f) * 0.5f;
  const float fi = (float)i;
  SnkParticle p;
  p.position_radius[0] = q.x + -100.0f; p.position_radius[1] = q.y + 20.0f; p.position_radius[2] = q.z + 190.0f; p.position_radius[3] = 0.15f;
  p.velocity[0] = 0.1f * sin_c(fi * 8.34534f); p.velocity[1] = 0.1f * cos_c(fi * 5.34342f); p.velocity[2] = 0.1f * sin_c(fi * 2.3435f);
  p.velocity[3] = 1.0f;
  const float third = (float)n / 3.0f;
  vec3 col = v3(0.0f, 0.0f, 1.0f);
  if ((float)i > third * 2.0f) col = v3(1.0f, 0.0f, 0.0f);
  else if ((float)i > third) col = v3(0.0f, 1.0f, 0.0f);
  p.colour[0] = col.x; p.colour[1] = col.y; p.colour[2] = col.z; p.colour[3] = 0.0f;
  P[i] = p;
}

__global__ void k_ptcl_keys(const SnkParticle* __restrict__ P, uint32_t n, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 pr = *reinterpret_cast<const float4*>(P[i].position_radius);
  keys[i] = cell_key((int)(pr.x * 5.0f), (int)(pr.y * 5.0f), (int)(pr.z * 5.0f), n);
  vals[i] = i;
}

__global__ void k_ptcl_cell_start(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* __restrict__ start) {  // Particle.comp:77-85
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = (uint32_t)keys[i];
  if (i == 0 || k != (uint32_t)keys[i - 1]) start[k] = i;
}

struct Payload { vec3 contact_normal, contact_velocity, contact_position; float distance; };

// all lanes of the warp call this together; lanes without a ray pass tmax <= 0
template <int STRIDE>
static __device__ __noinline__ HitRec trace_warp_closest(const TraceScene* S, float ox, float oy, float oz, float dx, float dy, float dz, float tmax, uint2* sm_col,
                                                         uint2* gl_col) {
  HitRec h;
  trace_warp_body<false, STRIDE>(*S, make_float4(ox, oy, oz, 0.0f), make_float4(dx, dy, dz, tmax), StackRef{sm_col, gl_col}, h);
  return h;
}

// ParticleUpdate.rchit:44-72 / .rmiss
__device__ __forceinline__ Payload make_payload(const SceneDev& D, const HitRec& h, vec3 o, vec3 d, float timestep) {
  Payload pl;
  pl.contact_normal = v3(0.0f); pl.contact_velocity = v3(0.0f); pl.contact_position = v3(0.0f); pl.distance = -1.0f;
  if (h.t < 0.0f) return pl;
  const float b0 = (1.0f - h.u) - h.v, b1 = h.u, b2 = h.v;
  const SnkInstanceData in = load_instance(D, h.inst);
  uint32_t ia, ib, ic;
  tri_indices(D, in, h.prim, ia, ib, ic);
  vec3 n = (ld3(D.normals, ia) * b0 + ld3(D.normals, ib) * b1) + ld3(D.normals, ic) * b2;
  const vec3 p = (ld3(D.positions, ia) * b0 + ld3(D.positions, ib) * b1) + ld3(D.positions, ic) * b2;
  n = normalize(mul_mat3r(D.normal_mats + 9ull * in.transform_idx, n));
  const vec3 contact = o + d * h.t;
  pl.distance = h.t;
  pl.contact_normal = n;
  pl.contact_position = contact;
  pl.contact_velocity = (contact - mul_point(D.prev_transforms + 16ull * in.transform_idx, p)) / timestep;
  return pl;
}

__global__ void __launch_bounds__(PT_THREADS, PT_CTAS_PER_SM)
k_ptcl_update(TraceScene S, SceneDev D, const SnkParticle* __restrict__ P, SnkParticle* __restrict__ out, const unsigned long long* __restrict__ keys,
              const uint32_t* __restrict__ order, const uint32_t* __restrict__ start, uint32_t n, SnkPhysicsParams prm, unsigned long long* ray_counter) {
  __shared__ uint2 sm_stack[SHORT_STACK * PT_THREADS];
  __shared__ TraceScene sS;
  if (threadIdx.x == 0) sS = S;
  __syncthreads();
  uint2* sm_col = sm_stack + threadIdx.x;
  uint2* gl_col = S.spill + ((size_t)blockIdx.x * (PT_THREADS / 32) + (threadIdx.x >> 5)) * (SPILL_STACK * 32) + (threadIdx.x & 31u);
  unsigned long long rays = 0;
  const uint32_t n32 = (n + 31u) & ~31u;
  for (uint32_t idx = blockIdx.x * PT_THREADS + threadIdx.x; idx < n32; idx += gridDim.x * PT_THREADS) {
    const bool live = idx < n;
    SnkParticle p;
    if (live) p = P[idx];
    else { p = SnkParticle{}; p.position_radius[3] = 0.0f; }
    const vec3 pos = v3(p.position_radius[0], p.position_radius[1], p.position_radius[2]);
    const float radius = p.position_radius[3];
    const vec3 vel = v3(p.velocity[0], p.velocity[1], p.velocity[2]);
    const bool asleep = p.velocity[3] < 0.5f;
    // ---- three collision rays (rgen:180-197)
    Payload cand;
    cand.contact_normal = v3(0.0f); cand.contact_velocity = v3(0.0f); cand.contact_position = v3(0.0f); cand.distance = -1.0f;
    float highest = 0.0f;
    const float speed = length(vel);
    const bool shoot = live && !asleep && speed > 0.0f;
    vec3 dirs[3];
    dirs[0] = shoot ? vel / speed : v3(0.0f, 1.0f, 0.0f);
    dirs[1] = fabsf(dot(dirs[0], v3(0.0f, 1.0f, 0.0f))) > 0.99f ? cross(dirs[0], v3(0.0f, 0.0f, 1.0f)) : cross(dirs[0], v3(0.0f, 1.0f, 0.0f));
    dirs[2] = cross(dirs[1], dirs[0]);
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
      const float len = k == 0 ? speed * PTCL_DT + radius * 2.0f : radius * 2.0f;
      const vec3 o = pos - dirs[k] * radius;
      const HitRec h = trace_warp_closest<PT_THREADS>(&sS, o.x, o.y, o.z, dirs[k].x, dirs[k].y, dirs[k].z, shoot ? len : -1.0f, sm_col, gl_col);
      if (shoot) {
        ++rays;
        const Payload pl = make_payload(D, h, o, dirs[k], prm.timestep);
        const float w = length(pl.contact_velocity - vel);
        if (pl.distance > 0.0f && w > highest) { highest = w; cand = pl; }
      }
    }
    // ---- particle / particle collisions over the 27 neighbouring cells (rgen:37-94)
    vec3 total_offset = v3(0.0f), total_impulse = v3(0.0f), colour_mix = v3(p.colour[0], p.colour[1], p.colour[2]);
    if (live) {
      const int cx = (int)(pos.x * 5.0f), cy = (int)(pos.y * 5.0f), cz = (int)(pos.z * 5.0f);
      vec3 local_impulse = v3(0.0f), local_offset = v3(0.0f);
      for (int xo = -1; xo <= 1; ++xo)
        for (int yo = -1; yo <= 1; ++yo)
          for (int zo = -1; zo <= 1; ++zo) {
            const uint32_t key = cell_key(cx + xo, cy + yo, cz + zo, n);
            const uint32_t s0 = __ldg(start + key);
            if (s0 == 0xFFFFFFFFu) continue;
            for (uint32_t i = s0; i < n; ++i) {
              const uint32_t other = __ldg(order + i);
              if (other == idx) continue;
              if ((uint32_t)__ldg(keys + i) != key) break;
              const float4 opr = __ldg(reinterpret_cast<const float4*>(P[other].position_radius));
              const float min_dist = opr.w + radius;
              const vec3 between = pos - v3(opr.x, opr.y, opr.z);
              const float dist = length(between);
              if (dist < (min_dist - 0.005f)) {
                const float4 ov = __ldg(reinterpret_cast<const float4*>(P[other].velocity));
                const float4 oc = __ldg(reinterpret_cast<const float4*>(P[other].colour));
                const vec3 cn = between / dist;
                const vec3 rel = vel - v3(ov.x, ov.y, ov.z);
                const float rs = length(rel);
                vec3 impulse = v3(0.0f);
                if (rs > prm.restitution_slop) impulse += (((cn * rs) * dot(cn, (-rel) / rs)) * 2.0f) * prm.particle_restitution;
                local_offset += (cn * (min_dist - dist)) * 0.5f;
                impulse += ((cn / PTCL_DT) * ((min_dist - dist) - 0.005f)) * prm.repulsion_factor;
                local_impulse += impulse;
                colour_mix = mix(colour_mix, v3(oc.x, oc.y, oc.z), prm.timestep);
              }
            }
          }
      const float i2 = fabsf(dot(local_impulse, local_impulse));
      if ((asleep && i2 > prm.sleep_threshold * prm.sleep_threshold) || !asleep) { total_offset += local_offset; total_impulse += local_impulse; }
    }
    // ---- gravity ray (rgen:147-165)
    {
      const HitRec h = trace_warp_closest<PT_THREADS>(&sS, pos.x, pos.y, pos.z, 0.0f, -1.0f, 0.0f, live ? radius + 0.001f : -1.0f, sm_col, gl_col);
      if (live) {
        ++rays;
        total_impulse.y += (-9.8f * PTCL_DT) * (h.t < 0.0f ? 1.0f : 0.0f);
      }
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
      const bool go = live && !(ol < 0.00001f);
      const vec3 od = go ? total_offset / ol : v3(0.0f, 1.0f, 0.0f);
      const HitRec h = trace_warp_closest<PT_THREADS>(&sS, pos.x, pos.y, pos.z, od.x, od.y, od.z, go ? ol + radius : -1.0f, sm_col, gl_col);
      if (live) {
        if (!go) total_offset = v3(0.0f);
        else {
          ++rays;
          const Payload c = make_payload(D, h, pos, od, prm.timestep);
          if (c.distance > 0.0f) total_offset = od * gmax(c.distance - radius * 0.9f, 0.0f) + (c.contact_normal * radius) * 0.05f;
        }
      }
    }
    if (!live) continue;
    // ---- friction + sleep (rgen:219-241)
    const float mass = 1.0f;
    float ke = (0.5f * mass) * dot(new_vel, new_vel);
    const float loss = (mass * length(total_offset)) * prm.friction_coefficient;
    ke -= gmin(loss, ke);
    const float v2 = fabsf((2.0f * ke) / mass);
    if (fabsf(dot(new_vel, new_vel)) > 0.0001f) new_vel = normalize(new_vel) * sqrtf(v2);
    SnkParticle o = p;
    o.velocity[3] = v2 > prm.sleep_threshold * prm.sleep_threshold ? 1.0f : 0.0f;
    const float awake = asleep ? 0.0f : 1.0f;
    o.position_radius[0] = pos.x + total_offset.x * awake; o.position_radius[1] = pos.y + total_offset.y * awake;
    o.position_radius[2] = pos.z + total_offset.z * awake; o.position_radius[3] = radius;
    o.velocity[0] = new_vel.x * awake; o.velocity[1] = new_vel.y * awake; o.velocity[2] = new_vel.z * awake;
    o.colour[0] = colour_mix.x; o.colour[1] = colour_mix.y; o.colour[2] = colour_mix.z;
    out[idx] = o;
  }
  for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(0xFFFFFFFFu, rays, o);
  if ((threadIdx.x & 31) == 0 && rays) atomicAdd(ray_counter, rays);
}

}  // namespace

int launch_particles_init(SnkParticle* P, uint32_t n, cudaStream_t st) {
  if (n == 0) return 0;
  k_ptcl_init<<<(n + 255) / 256, 256, 0, st>>>(P, n);
  return 1;
}

// one sub-step: P -> out (the caller swaps). scratch.keys / vals hold the key list; returns the number of launches
int launch_particles_substep(SnkPt* ctx, const SnkParticle* P, SnkParticle* out, uint32_t n, uint32_t* cell_start, const SnkPhysicsParams& prm,
                             unsigned long long* ray_counter, cudaStream_t st) {
  BuildScratch& s = ctx->scratch_ptcl;
  int launches = 0;
  k_ptcl_keys<<<(n + 255) / 256, 256, 0, st>>>(P, n, s.keys[0], s.vals[0]);
  ++launches;
  int bits = 1;
  while (bits < 32 && (1ull << bits) < (unsigned long long)n) ++bits;  // keys are < n
  int cur = 0;
  launches += launch_sort_pairs(s, n, bits, cur, st);
  cudaMemsetAsync(cell_start, 0xFF, 4ull * n, st);
  k_ptcl_cell_start<<<(n + 255) / 256, 256, 0, st>>>(s.keys[cur], n, cell_start);
  ++launches;
  const uint32_t blocks = (n + PT_THREADS - 1) / PT_THREADS, persistent = (uint32_t)device_sm_count() * PT_CTAS_PER_SM;
  const int rec = prof_begin(ctx, "ptcl_update", n, false, false, st);
  k_ptcl_update<<<blocks < persistent ? blocks : persistent, PT_THREADS, 0, st>>>(ctx->trace_scene(), ctx->scene_dev(), P, out, s.keys[cur], s.vals[cur],
                                                                                   cell_start, n, prm, ray_counter);
  prof_end(ctx, rec, st);
  return launches + 1;
}

}  // namespace snk
