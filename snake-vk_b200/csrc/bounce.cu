// bounce.cu — BounceDiffuse (RayGen.rgen:139-213) + RayClosestHit.rchit as wavefront stages, and the
// orchestration of one reference-mode frame (RT::RecordRenderCmdBuf, Raytracing.cpp:227-267).
//
// The reference traces, per pixel, 8 hemisphere rays from the primary hit and 2 more from every
// non-emissive first hit, each followed by a directional-light shadow ray in the hit shader:
// up to 24 closest-hit + 24 any-hit rays. Here that recursion is unrolled into queues in HBM:
//
//   k_bounce_gen     pixel  -> 8 first-bounce rays            (queue A, 32 B/ray, index = k*npix + pixel)
//   trace closest    A      -> hits A (16 B + 4 B)
//   k_bounce_shade   A      -> unshadowed hit colour (16 B), shadow ray (overwrites A), 2 inner rays (queue B)
//   trace any        A      -> visibility A (4 B);     trace closest  B -> hits B
//   k_bounce_shade   B      -> colour B, shadow rays (overwrite B)
//   trace any        B      -> visibility B
//   k_bounce_resolve pixel  -> replays the reference's accumulation order exactly, adds the direct part, writes the image
//
// Queue slots are fixed (no compaction yet): dead rays carry tmax < tmin and fall out of the traversal
// after one node fetch. Compiled with -fmad=false (shading numerics contract).
#include "kernels.h"
#include "shading.cuh"

namespace snk {

namespace {

struct BounceBuffers {
  float4* rays_a;  float4* hit_a;  uint32_t* inst_a;  float4* col_a;  float* cos_a;  float* vis_a;   // 8 per pixel
  float4* rays_b;  float4* hit_b;  uint32_t* inst_b;  float4* col_b;  float* cos_b;  float* vis_b;   // 16 per pixel
};

__device__ __forceinline__ void store_ray(float4* q, size_t i, vec3 o, float tmin, vec3 d, float tmax) {
  q[2 * i] = make_float4(o.x, o.y, o.z, tmin);
  q[2 * i + 1] = make_float4(d.x, d.y, d.z, tmax);
}
__device__ __forceinline__ void store_dead_ray(float4* q, size_t i) {
  q[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
  q[2 * i + 1] = make_float4(0.f, 1.f, 0.f, -1.0f);
}

// the two per-pixel constants of BounceDiffuse's RNG (RayGen.rgen:150-154)
__device__ __forceinline__ void bounce_rng_setup(const FrameConsts& fc, float fx, float fy, vec3 ro, float& seed, float& xx, float& xy) {
  const float l2_ro = dot(ro, ro);
  seed = fract((((((l2_ro + ro.x) + ro.y) + ro.z) + fc.time_elapsed) + fc.delta_time));
  xx = fx + fract(sin_c(fx * 1038.283726f + fy * fc.delta_time));
  xy = fy + fract(cos_c(fx * fc.time_elapsed + fy * 9278.33337f));
}

__device__ __forceinline__ bool pixel_shades(const PixelPlanes& P, size_t px) {
  return !(P.flags[px] & SNKPT_MAT_FLAG_EMISSIVE) && P.prim[px] != INVALID_IDX;
}
__device__ __forceinline__ void pixel_surface(const FrameConsts& fc, const PixelPlanes& P, uint32_t x, uint32_t y, vec3& normal, vec3& biased_pos) {
  const size_t px = (size_t)y * fc.W + x;
  const ushort4 n = P.normal[px];
  normal = v3(h2f(n.x), h2f(n.y), h2f(n.z));
  const float u = ((float)x + 0.5f) / (float)fc.W, vv = ((float)y + 0.5f) / (float)fc.H;
  const vec3 ro = world_pos_from_depth(fc.ivp, P.depth[px], u, vv);
  const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
  biased_pos = ro + normal * (0.01f + gmax(length(cam - ro) * 0.001f, 0.01f));  // RayGen.rgen:555
}

// RayGen.rgen:155-176: the 8 first-bounce rays of one pixel
__global__ void __launch_bounds__(256) k_bounce_gen(FrameConsts fc, PixelPlanes P, BounceBuffers B, uint32_t npix, unsigned long long* counter) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // pixel index inside the band
  unsigned long long emitted = 0;
  if (i < npix) {
    const uint32_t x = i % fc.W, y = fc.row0 + i / fc.W;
    const size_t px = (size_t)y * fc.W + x;
    if (!pixel_shades(P, px)) {
      for (uint32_t k = 0; k < 8; ++k) { store_dead_ray(B.rays_a, (size_t)k * npix + i); B.cos_a[(size_t)k * npix + i] = 0.0f; }
    } else {
      vec3 normal, ro;
      pixel_surface(fc, P, x, y, normal, ro);
      vec3 nt, nb;
      create_coordinate_system(normal, nt, nb);
      float seed, xx, xy;
      bounce_rng_setup(fc, (float)x, (float)y, ro, seed, xx, xy);
      for (uint32_t k = 0; k < 8; ++k) {
        seed = fract(seed + (float)k / 8.0f);
        const float s = fract(seed + (float)k / 8.0f);
        const vec3 rnd = uniform_sample_hemisphere(gold_noise(xx, xy, s), gold_noise(xx, xy, s + 0.1f));
        const vec3 rd = v3((rnd.x * nb.x + rnd.y * normal.x) + rnd.z * nt.x, (rnd.x * nb.y + rnd.y * normal.y) + rnd.z * nt.y,
                           (rnd.x * nb.z + rnd.y * normal.z) + rnd.z * nt.z);
        store_ray(B.rays_a, (size_t)k * npix + i, ro, 0.01f, rd, 1000.0f);
        B.cos_a[(size_t)k * npix + i] = gmax(cos_c(dot(rd, normal)), 0.0f);  // :178 cos OF the dot product (quirk 1)
        ++emitted;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) emitted += __shfl_xor_sync(0xFFFFFFFFu, emitted, o);
  if ((threadIdx.x & 31) == 0 && emitted) atomicAdd(counter, emitted);
}

// RayClosestHit.rchit for every ray of a queue. LEVEL 0: first-bounce rays (also emits the 2 inner rays,
// RayGen.rgen:183-207); LEVEL 1: inner rays.
template <int LEVEL>
__global__ void __launch_bounds__(256) k_bounce_shade(SceneDev D, FrameConsts fc, PixelPlanes P, BounceBuffers B, uint32_t npix,
                                                      unsigned long long* counters) {
  const size_t n = (size_t)(LEVEL == 0 ? 8 : 16) * npix;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long n_shadow = 0, n_closest = 0;
  if (i < n) {
    float4* rays = LEVEL == 0 ? B.rays_a : B.rays_b;
    const float4 h = (LEVEL == 0 ? B.hit_a : B.hit_b)[i];
    const uint32_t inst = (LEVEL == 0 ? B.inst_a : B.inst_b)[i];
    const float4 ro4 = rays[2 * i], rd4 = rays[2 * i + 1];
    const bool live = rd4.w >= 0.0f;
    const vec3 ro = v3(ro4.x, ro4.y, ro4.z), rd = v3(rd4.x, rd4.y, rd4.z);
    const float4 dl = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.dir));
    const float4 dc = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.colour));
    const vec3 ldir = v3(dl.x, dl.y, dl.z), lcol = v3(dc.x, dc.y, dc.z);
    HitShade s;
    s.colour = v3(0.0f); s.normal = v3(0.0f); s.shadow_o = v3(0.0f); s.distance = -1.0f; s.mat_flags = 0; s.needs_shadow = false;
    if (live) s = shade_hit(D, ldir, lcol, rd, h.x, h.y, h.z, __float_as_uint(h.w), inst);
    // continues == the reference does NOT `continue` at RayGen.rgen:179
    const bool continues = live && !(s.distance < 0.0f || (s.mat_flags & SNKPT_MAT_FLAG_EMISSIVE));
    (LEVEL == 0 ? B.col_a : B.col_b)[i] = make_float4(s.colour.x, s.colour.y, s.colour.z, continues ? 1.0f : 0.0f);
    // the hit shader's shadow ray (rchit:136) replaces the consumed ray in the same slot
    if (s.needs_shadow) { store_ray(rays, i, s.shadow_o, 0.001f, ldir, 10000.0f); ++n_shadow; }
    else store_dead_ray(rays, i);
    if (LEVEL == 0) {
      const uint32_t k = (uint32_t)(i / npix), pix = (uint32_t)(i % npix);
      const size_t j0 = (size_t)(2 * k) * npix + pix, j1 = (size_t)(2 * k + 1) * npix + pix;
      if (!continues) {
        store_dead_ray(B.rays_b, j0); store_dead_ray(B.rays_b, j1);
        B.cos_b[j0] = 0.0f; B.cos_b[j1] = 0.0f;
      } else {
        const uint32_t x = pix % fc.W, y = fc.row0 + pix / fc.W;
        float seed, xx, xy;
        bounce_rng_setup(fc, (float)x, (float)y, ro, seed, xx, xy);
        for (uint32_t kk = 0; kk <= k; ++kk) seed = fract(seed + (float)kk / 8.0f);  // seed after iteration k (:156)
        const vec3 new_pos = ro + rd * s.distance;
        const vec3 n_inner = s.normal;
        vec3 nt, nb;
        create_coordinate_system(n_inner, nt, nb);
        for (uint32_t b = 0; b < 2; ++b) {
          const vec3 hs = uniform_sample_hemisphere(gold_noise(xx, xy, fract((seed + (float)b * 0.5f) + 0.2f)),
                                                    gold_noise(xx, xy, fract((seed + (float)b * 0.5f) + 0.3f)));
          const vec3 rd2 = v3((hs.x * nb.x + hs.y * n_inner.x) + hs.z * nt.x, (hs.x * nb.y + hs.y * n_inner.y) + hs.z * nt.y,
                              (hs.x * nb.z + hs.y * n_inner.z) + hs.z * nt.z);
          const size_t j = b == 0 ? j0 : j1;
          store_ray(B.rays_b, j, new_pos, 0.01f, rd2, 1000.0f);
          B.cos_b[j] = gmax(cos_c(dot(rd2, n_inner)), 0.0f);  // :208
          ++n_closest;
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    n_shadow += __shfl_xor_sync(0xFFFFFFFFu, n_shadow, o);
    n_closest += __shfl_xor_sync(0xFFFFFFFFu, n_closest, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_shadow) atomicAdd(counters, n_shadow);
    if (n_closest) atomicAdd(counters + 1, n_closest);
  }
}

// RayGen.rgen:178, :208, :212, :566-569 — the accumulation replayed in the reference's order
__global__ void __launch_bounds__(256) k_bounce_resolve(FrameConsts fc, PixelPlanes P, BounceBuffers B, uint32_t npix, int with_bounce) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const uint32_t x = i % fc.W, y = fc.row0 + i / fc.W;
  const size_t px = (size_t)y * fc.W + x;
  const char4 a = P.albedo[px];
  const vec3 albedo = v3(snorm8_2f(a.x), snorm8_2f(a.y), snorm8_2f(a.z));
  vec3 out;
  if (P.flags[px] & SNKPT_MAT_FLAG_EMISSIVE) out = albedo;               // RayGen.rgen:540-543
  else if (P.prim[px] == INVALID_IDX) out = v3(0.0f);                     // sky (DEVIATION, quirk 12: masked)
  else {
    const float4 d = P.direct[px];
    out = v3(d.x, d.y, d.z);
    if (with_bounce) {
      vec3 colour_sum = v3(0.0f);
      for (uint32_t k = 0; k < 8; ++k) {
        const size_t ia = (size_t)k * npix + i;
        const float4 c1 = B.col_a[ia];
        const vec3 payload = v3(c1.x, c1.y, c1.z) * B.vis_a[ia];
        colour_sum += ((albedo * payload) * B.cos_a[ia]) / 8.0f;
        if (c1.w == 0.0f) continue;
        for (uint32_t b = 0; b < 2; ++b) {
          const size_t ib = (size_t)(2 * k + b) * npix + i;
          const float4 c2 = B.col_b[ib];
          const vec3 payload2 = v3(c2.x, c2.y, c2.z) * B.vis_b[ib];
          colour_sum += (((albedo * payload2) * 0.25f) * B.cos_b[ib]) / 8.0f;
        }
      }
      out += colour_sum * (2.0f * PI_GLSL);
    }
  }
  P.radiance[px] = make_float4(out.x, out.y, out.z, 1.0f);
  P.out16[px] = make_ushort4(f2h(out.x), f2h(out.y), f2h(out.z), f2h(1.0f));  // internal_render_image is RGBA16F
}

// frame counters [2..5] (shadow, closest as u64) are folded into the since-create totals [8..11]
__global__ void k_fold_counters(unsigned long long* c) {
  c[4] += c[1];
  c[5] += c[2];
}

}  // namespace

int launch_reference_frame(SnkPt* ctx, const SnkRtSettings& s, cudaStream_t st) {
  const FrameConsts& fc = ctx->fc;
  const uint32_t npix = fc.W * (fc.row1 - fc.row0);
  int launches = 0;
  cudaMemsetAsync(ctx->dev_counters.as<uint32_t>() + 2, 0, 16, st);
  launches += launch_restir_passes(ctx, s, st);
  const bool bounce = !(s.flags & SNKPT_RT_NO_BOUNCE);
  BounceBuffers B{};
  if (bounce) {
    const size_t na = (size_t)8 * npix, nb = (size_t)16 * npix;
    const size_t sizes[12] = {32 * na, 16 * na, 4 * na, 16 * na, 4 * na, 4 * na, 32 * nb, 16 * nb, 4 * nb, 16 * nb, 4 * nb, 4 * nb};
    for (int k = 0; k < 12; ++k) {
      cudaError_t e = ctx->wave[k].reserve(sizes[k], false, st);
      if (e != cudaSuccess) { ctx->err = std::string("bounce queues: ") + cudaGetErrorString(e); return SNKPT_E_NOMEM; }
    }
    B.rays_a = ctx->wave[0].as<float4>(); B.hit_a = ctx->wave[1].as<float4>(); B.inst_a = ctx->wave[2].as<uint32_t>();
    B.col_a = ctx->wave[3].as<float4>(); B.cos_a = ctx->wave[4].as<float>(); B.vis_a = ctx->wave[5].as<float>();
    B.rays_b = ctx->wave[6].as<float4>(); B.hit_b = ctx->wave[7].as<float4>(); B.inst_b = ctx->wave[8].as<uint32_t>();
    B.col_b = ctx->wave[9].as<float4>(); B.cos_b = ctx->wave[10].as<float>(); B.vis_b = ctx->wave[11].as<float>();
    unsigned long long* c_shadow = reinterpret_cast<unsigned long long*>(ctx->dev_counters.as<uint32_t>() + 2);
    unsigned long long* c_closest = c_shadow + 1;
    const SceneDev D = ctx->scene_dev();
    int rec = prof_begin(ctx, "bounce_gen", npix, false, false, st);
    k_bounce_gen<<<(npix + 255) / 256, 256, 0, st>>>(fc, ctx->planes, B, npix, c_closest);
    prof_end(ctx, rec, st);
    launches += 1 + traced_launch(ctx, "trace_closest.A", B.rays_a, na, false, nullptr, B.hit_a, B.inst_a, nullptr, st);
    rec = prof_begin(ctx, "bounce_shade.A", na, false, false, st);
    k_bounce_shade<0><<<(unsigned)((na + 255) / 256), 256, 0, st>>>(D, fc, ctx->planes, B, npix, c_shadow);
    prof_end(ctx, rec, st);
    launches += 1 + traced_launch(ctx, "trace_any.A", B.rays_a, na, true, nullptr, nullptr, nullptr, B.vis_a, st);
    launches += traced_launch(ctx, "trace_closest.B", B.rays_b, nb, false, nullptr, B.hit_b, B.inst_b, nullptr, st);
    rec = prof_begin(ctx, "bounce_shade.B", nb, false, false, st);
    k_bounce_shade<1><<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(D, fc, ctx->planes, B, npix, c_shadow);
    prof_end(ctx, rec, st);
    launches += 1 + traced_launch(ctx, "trace_any.B", B.rays_b, nb, true, nullptr, nullptr, nullptr, B.vis_b, st);
  }
  {
    wait_readback(ctx, st);  // the resolve overwrites the output image
    const int rec = prof_begin(ctx, "bounce_resolve", npix, false, false, st);
    k_bounce_resolve<<<(npix + 255) / 256, 256, 0, st>>>(fc, ctx->planes, B, npix, bounce ? 1 : 0);
    prof_end(ctx, rec, st);
  }
  k_fold_counters<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(ctx->dev_counters.as<uint32_t>()));
  launches += 2;
  cudaEventRecord(ctx->ev[5], st);
  ctx->stats.kernel_launches += launches;
  ctx->render_pending = true;
  return launches;
}

}  // namespace snk
