// restir.cu — ReSTIR DI passes of the reference frame (reference mode).
//
// k_restir_pass1  = RayGen.rgen main() with push.initial_reservoir_pass == 1:
//                   WriteInitialReservoirs (:411-468) = DoInitialWRS (:296-357) + temporal reuse
// k_restir_pass2  = pass 0 up to (not including) BounceDiffuse: directional light (:562),
//                   ComputeAndWriteSpatialReservoirReuse (:470-506), CalculateRadianceFromReservoirs (:508-528)
// (RT::RecordRenderCmdBuf, Core/src/rendering/Raytracing.cpp:250-260, launches both as traceRaysKHR(W,H,1).)
//
// One thread per pixel in 16x8 tiles; reservoirs are read / written as three 16-B SoA planes per half.
// Shadow rays are traced inline with the shared-memory stack; every lane of a warp calls the tracer at
// the same program points (inactive lanes pass tmax < tmin), so warps stay convergent through the
// traversal loops. Reference quirks are reproduced; guarded out-of-bounds reads are
// marked DEVIATION exactly as in the CPU oracle.
#include "kernels.h"
#include "shading.cuh"

namespace snk {

namespace {

constexpr int RT_THREADS = 128;

struct RestirParams {
  int32_t num_temporal, num_spatial;
  uint32_t ris_iters, m_cap;
  float radius, normal_threshold;
  uint32_t flags;
};

// per-pixel shading context (what main() computes at RayGen.rgen:531-555)
struct Px {
  uint32_t x, y, W, H;
  float fx, fy, time, delta;
  vec3 albedo, normal, ro, v, f0, biased_pos;
  float roughness, metallic;
  bool shade;  // false: sky / emissive / out of range → no reservoir work
};

struct Tracer {
  const TraceScene* S;  // CTA-shared copy of the kernel parameter (the tracer is a real call, see traverse.cuh)
  uint2 *sm_col, *gl_col;
  unsigned long long n_any;
  __device__ __forceinline__ float shadow(bool valid, vec3 ro, vec3 rd, float max_dist) {  // RayGen.rgen:101-121
    const float hit = trace_warp_any<RT_THREADS>(S, ro.x, ro.y, ro.z, 0.001f, rd.x, rd.y, rd.z, valid ? max_dist : -1.0f, sm_col, gl_col);
    if (valid) ++n_any;
    return 1.0f - hit;
  }
};

// persistent CTAs: RT_CTAS_PER_SM x 148 CTAs walk the 16x8 pixel tiles of the band
#ifndef SNK_RT_CTAS
#define SNK_RT_CTAS 8
#endif
constexpr int RT_CTAS_PER_SM = SNK_RT_CTAS;
#define SNK_RT_PROLOGUE()                                                                                              \
  __shared__ uint2 sm_stack[SHORT_STACK * RT_THREADS];                                                                 \
  __shared__ TraceScene sS;                                                                                            \
  if (threadIdx.x == 0) sS = S;                                                                                        \
  __syncthreads();                                                                                                     \
  Tracer T{&sS, sm_stack + threadIdx.x,                                                                                \
           S.spill + ((size_t)blockIdx.x * (RT_THREADS / 32) + (threadIdx.x >> 5)) * (SPILL_STACK * 32) + (threadIdx.x & 31u), 0ull}; \
  const uint32_t tiles_x = (fc.W + 15u) / 16u, n_tiles = tiles_x * ((fc.row1 - fc.row0 + 7u) / 8u)

__device__ __forceinline__ Px load_px(const FrameConsts& fc, const PixelPlanes& P, uint32_t x, uint32_t y, bool in_range) {
  Px q;
  q.x = x; q.y = y; q.W = fc.W; q.H = fc.H; q.fx = (float)x; q.fy = (float)y; q.time = fc.time_elapsed; q.delta = fc.delta_time;
  q.shade = false;
  q.albedo = q.normal = q.ro = q.v = q.f0 = q.biased_pos = v3(0.0f);
  q.roughness = q.metallic = 0.0f;
  if (!in_range) return q;
  const size_t px = (size_t)y * fc.W + x;
  const char4 a = P.albedo[px];
  const ushort4 r = P.rma[px];
  q.albedo = v3(snorm8_2f(a.x), snorm8_2f(a.y), snorm8_2f(a.z));
  q.roughness = h2f(r.x); q.metallic = h2f(r.y);
  const uint32_t mat_flags = P.flags[px];
  if ((mat_flags & SNKPT_MAT_FLAG_EMISSIVE) || P.prim[px] == INVALID_IDX) return q;
  const ushort4 n = P.normal[px];
  q.normal = v3(h2f(n.x), h2f(n.y), h2f(n.z));
  const float u = (q.fx + 0.5f) / (float)fc.W, vv = (q.fy + 0.5f) / (float)fc.H;
  q.ro = world_pos_from_depth(fc.ivp, P.depth[px], u, vv);
  const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
  q.v = normalize(cam - q.ro);
  q.f0 = mix(v3(0.04f), q.albedo, q.metallic);
  q.biased_pos = q.ro + q.normal * (0.01f + gmax(length(cam - q.ro) * 0.001f, 0.01f));
  q.shade = true;
  return q;
}

// RayGen.rgen:215-226
__device__ __forceinline__ bool update_pl(const Px& q, Reservoir& r, uint32_t pl_idx, float weight, float rnd_seed) {
  r.pl_m++;
  const float rnd = gold_noise(q.fx, q.fy, fract((weight + q.time * 0.0009783f) + rnd_seed));
  r.w_sum_pl += weight;
  if (r.w_sum_pl > 0.0f && rnd < weight / r.w_sum_pl) { r.pl_idx = pl_idx; return true; }
  return false;
}
// RayGen.rgen:228-239
__device__ __forceinline__ bool update_em(const Px& q, Reservoir& r, uint32_t obj, uint32_t tri, float bx, float by, float weight, float rnd_seed) {
  r.m_em++;
  const float rnd = gold_noise(q.fx, q.fy, fract((weight + q.time * 0.000789f) + rnd_seed));
  r.w_sum_em += weight;
  if (r.w_sum_em > 0.0f && rnd < weight / r.w_sum_em) { r.obj = obj; r.tri = tri; r.bx = bx; r.by = by; return true; }
  return false;
}

// RayGen.rgen:296-357. `on` = this lane really shades; all lanes run the two shadow tests together.
__device__ void do_initial_wrs(const SceneDev& D, const RestirParams& rp, Tracer& T, const Px& q, bool on, Reservoir& r) {
  float p_hat_pl = 0.0f, p_hat_em = 0.0f;
  const uint32_t iters = rp.ris_iters;
  const uint32_t n_pl = D.lights->num_pointlights, n_em = D.num_emissive;
  if (on) {
    const vec3 wp = q.ro;
    for (uint32_t i = 0; i < iters; ++i) {
      if (n_pl == 0) { r.pl_m++; continue; }  // DEVIATION (as oracle): the reference reads slot 0 of an empty light array
      const float seed = fract(((q.time * 0.001197f + wp.x) + wp.y) + wp.z) + (float)i / (float)iters;
      uint32_t rnd_idx = to_uint(gold_noise(q.fx, q.fy, seed) * (float)n_pl);
      if (rnd_idx >= n_pl) rnd_idx = n_pl - 1;  // DEVIATION: guard
      const PointLightRec pl = load_pointlight(D.lights, rnd_idx);
      const float p_hat = length(calc_pointlight(pl, q.v, q.f0, wp, q.normal, q.roughness, q.metallic, q.albedo));
      const float weight = p_hat / (1.0f / (float)n_pl);
      if (update_pl(q, r, rnd_idx, weight, sin_c((float)i * 20382997.34741f))) p_hat_pl = p_hat;
    }
    for (uint32_t i = 0; i < iters; ++i) {
      if (n_em == 0) { r.m_em++; continue; }  // DEVIATION (as oracle): the reference reads idx[0] of an empty list
      const float wsum = (wp.x + wp.y) + wp.z;
      uint32_t e = to_uint(gold_noise(q.fx, q.fy, fabsf(fract((wsum + q.time * 0.001534f) + (float)i / (float)iters))) * (float)n_em);
      if (e >= n_em) e = n_em - 1;  // DEVIATION: guard
      const uint32_t obj = __ldg(D.emissive_idx + e);
      const SnkInstanceData in = load_instance(D, obj);
      const uint32_t ntri = in.num_mesh_indices / 3u;
      uint32_t tri = to_uint(gold_noise(q.fx, q.fy, fabsf(fract((wsum + q.time * 0.00119f) + (float)i / (float)iters))) * (float)ntri);
      if (tri >= ntri) tri = ntri - 1;  // DEVIATION: guard
      uint32_t ia, ib, ic;
      tri_indices(D, in, tri, ia, ib, ic);
      float bx = random_f(((sin_c((float)i * 27934.3221f + q.fx * 293.28f)) + q.time * 0.001f) + (float)i);
      float by = random_f((((sin_c((float)i * 12391.8583f + q.fy * 983.3399f)) + (float)(int32_t)ib) + q.time * 0.0091f) + (float)i);
      if (bx + by >= 1.0f) { bx = 1.0f - bx; by = 1.0f - by; }
      const vec3 pa = ld3(D.positions, ia), pb = ld3(D.positions, ib), pc = ld3(D.positions, ic);
      const float triangle_area = (__ldg(D.det3 + in.transform_idx) * 0.5f) * length(cross(pb - pa, pb - pc));
      const TriPoint tp = triangle_point(D, obj, tri, bx, by);
      const vec3 radiance = triangle_radiance_unshadowed(tp, q.v, q.f0, wp, q.normal, q.roughness, q.metallic, q.albedo);
      const float p_hat = length(radiance);
      const float pdf = 1.0f / ((float)(n_em * ntri) * triangle_area);
      const float weight = p_hat / pdf;
      if (update_em(q, r, obj, tri, bx, by, weight, sin_c((float)i * 38293.28347f))) p_hat_em = p_hat;
    }
  }
  // :344-349 — visibility of the selected point light; operator-precedence bug reproduced (quirk 3)
  {
    const bool valid = on && r.pl_idx != INVALID_IDX;
    vec3 dir = v3(0.0f, 1.0f, 0.0f);
    float pl_dist = 1.0f;
    if (valid) {
      const PointLightRec pl = load_pointlight(D.lights, r.pl_idx);
      pl_dist = length(pl.pos - q.biased_pos);
      dir = pl.pos - q.biased_pos / pl_dist;
    }
    const float vis = T.shadow(valid, q.biased_pos, dir, pl_dist);
    r.w_pl = valid ? vis * ((r.w_sum_pl / (float)r.pl_m) / p_hat_pl) : 0.0f;
  }
  // :351-355 — TriangleShadowTest of the selected emissive sample
  {
    const bool valid = on && r.obj != INVALID_IDX;
    vec3 dir = v3(0.0f, 1.0f, 0.0f);
    float dist = 1.0f;
    if (valid) {
      const TriPoint tp = triangle_point(D, r.obj, r.tri, r.bx, r.by);
      triangle_shadow_ray(tp, q.biased_pos, dir, dist);
    }
    const float vis = T.shadow(valid, q.biased_pos, dir, dist);
    r.w_em = valid ? vis * ((r.w_sum_em / (float)r.m_em) / p_hat_em) : 0.0f;
  }
}

// RayGen.rgen:361-386. `on` gates the whole call (lanes with on == false only keep the warp convergent).
__device__ void combine(const SceneDev& D, Tracer& T, const Px& q, bool on, Reservoir& cur, const Reservoir& nb, float& p_hat_pl,
                        float& p_hat_em, float rnd_seed) {
  {
    const bool valid = on && nb.pl_idx != INVALID_IDX;
    vec3 dir = v3(0.0f, 1.0f, 0.0f);
    float light_dist = 1.0f, p_hat_unshadowed = 0.0f;
    if (valid) {
      const PointLightRec pl = load_pointlight(D.lights, nb.pl_idx);
      light_dist = length(pl.pos - q.biased_pos);
      p_hat_unshadowed = length(calc_pointlight(pl, q.v, q.f0, q.ro, q.normal, q.roughness, q.metallic, q.albedo));
      dir = (pl.pos - q.biased_pos) / light_dist;
    }
    const float vis = T.shadow(valid, q.biased_pos, dir, light_dist - 0.001f);
    if (valid) {
      const float p_hat = p_hat_unshadowed * vis;
      const float weight = (p_hat * (float)nb.pl_m) * nb.w_pl;
      if (update_pl(q, cur, nb.pl_idx, weight, rnd_seed)) p_hat_pl = p_hat;
    }
  }
  {
    const bool valid = on && nb.obj != INVALID_IDX;
    vec3 dir = v3(0.0f, 1.0f, 0.0f), radiance = v3(0.0f);
    float dist = 1.0f;
    if (valid) {
      const TriPoint tp = triangle_point(D, nb.obj, nb.tri, nb.bx, nb.by);
      radiance = triangle_radiance_unshadowed(tp, q.v, q.f0, q.ro, q.normal, q.roughness, q.metallic, q.albedo);
      triangle_shadow_ray(tp, q.biased_pos, dir, dist);
    }
    const float vis = T.shadow(valid, q.biased_pos, dir, dist);
    if (valid) {
      radiance *= vis;
      const float p_hat = length(radiance);
      const float weight = (p_hat * nb.w_em) * (float)nb.m_em;
      if (update_em(q, cur, nb.obj, nb.tri, nb.bx, nb.by, weight, rnd_seed + 0.283673699f)) p_hat_em = p_hat;
    }
  }
}

__device__ __forceinline__ vec3 prev_normal_texel(const PixelPlanes& P, uint32_t W, int px, int py) {
  const ushort4 n = P.prev_normal[(size_t)py * W + px];
  return v3(h2f(n.x), h2f(n.y), h2f(n.z));
}
// texture(sampler_prev_frame_normals, uv): linear filter, repeat addressing (Images.cpp:16-37)
__device__ vec3 prev_normal_bilinear(const PixelPlanes& P, int W, int H, float u, float v) {
  const float su = u * (float)W - 0.5f, sv = v * (float)H - 0.5f;
  const float fu = floorf(su), fv = floorf(sv);
  const float a = su - fu, b = sv - fv;
  int x0 = (int)fu, y0 = (int)fv;
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = ((x0 % W) + W) % W; x1 = ((x1 % W) + W) % W; y0 = ((y0 % H) + H) % H; y1 = ((y1 % H) + H) % H;
  const vec3 t00 = prev_normal_texel(P, W, x0, y0), t10 = prev_normal_texel(P, W, x1, y0), t01 = prev_normal_texel(P, W, x0, y1),
             t11 = prev_normal_texel(P, W, x1, y1);
  const vec3 top = t00 * (1.0f - a) + t10 * a, bot = t01 * (1.0f - a) + t11 * a;
  return top * (1.0f - b) + bot * b;
}

__device__ __forceinline__ void tile_xy(const FrameConsts& fc, uint32_t tile, uint32_t tiles_x, uint32_t& x, uint32_t& y, bool& in_range) {
  const uint32_t t = threadIdx.x, w = t >> 5, l = t & 31;
  x = (tile % tiles_x) * 16 + (w & 1) * 8 + (l & 7);
  y = fc.row0 + (tile / tiles_x) * 8 + (w >> 1) * 4 + (l >> 3);
  in_range = x < fc.W && y < fc.row1;
}

// ---- pass 1: WriteInitialReservoirs (RayGen.rgen:411-468) --------------------------------------------------
__global__ void __launch_bounds__(RT_THREADS, RT_CTAS_PER_SM) k_restir_pass1(TraceScene S, SceneDev D, FrameConsts fc, PixelPlanes P,
                                                                             RestirParams rp, unsigned long long* ray_counter) {
  SNK_RT_PROLOGUE();
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  uint32_t x, y; bool in_range;
  tile_xy(fc, tile, tiles_x, x, y, in_range);
  const Px q = load_px(fc, P, x, y, in_range);
  const ResPlanes half0{P.res[0][0], P.res[0][1], P.res[0][2]}, half1{P.res[1][0], P.res[1][1], P.res[1][2]};
  Reservoir pixel, temporal;
  init_reservoir(pixel);
  init_reservoir(temporal);
  do_initial_wrs(D, rp, T, q, q.shade, pixel);
  // temporal reuse
  bool reuse = false;
  Reservoir history;
  init_reservoir(history);
  if (q.shade && !fc.prev_frame_invalid && rp.num_temporal != 0) {
    const size_t px = (size_t)y * fc.W + x;
    const ushort2 mv = P.motion[px];
    const float hx = (q.fx + 0.5f) / (float)fc.W + h2f(mv.x), hy = (q.fy + 0.5f) / (float)fc.H + h2f(mv.y);
    if ((hx > 0.0f && hy > 0.0f) && (hx < 1.0f && hy < 1.0f)) {
      const vec3 nn = prev_normal_bilinear(P, (int)fc.W, (int)fc.H, hx, hy);
      if (dot(nn, q.normal) > rp.normal_threshold) {
        int nx = (int)(hx * (float)fc.W), ny = (int)(hy * (float)fc.H);
        if (nx > (int)fc.W - 1) nx = (int)fc.W - 1;  // DEVIATION: guard
        if (ny > (int)fc.H - 1) ny = (int)fc.H - 1;
        history = load_reservoir(half0, (size_t)ny * fc.W + nx);
        const float cap = (float)rp.m_cap;
        if ((float)history.pl_m > cap * (float)pixel.pl_m) {
          history.w_sum_pl *= (cap * (float)pixel.pl_m) / (float)history.pl_m;
          history.pl_m = rp.m_cap * pixel.pl_m;
        }
        if ((float)history.m_em > cap * (float)pixel.m_em) {
          history.w_sum_em *= (cap * (float)pixel.m_em) / (float)history.m_em;
          history.m_em = rp.m_cap * pixel.m_em;
        }
        reuse = true;
      }
    }
  }
  // the warp runs the two combines together if any of its lanes reuses history
  if (__any_sync(0xFFFFFFFFu, reuse)) {
    float p_hat_em = 0.0f, p_hat_pl = 0.0f;
    combine(D, T, q, reuse, temporal, pixel, p_hat_pl, p_hat_em, 0.846746997f);
    combine(D, T, q, reuse, temporal, history, p_hat_pl, p_hat_em, 0.0f);
    if (reuse) {
      temporal.m_em = history.m_em + pixel.m_em;
      temporal.pl_m = history.pl_m + pixel.pl_m;
      temporal.w_em = p_hat_em > 0.0f ? (temporal.w_sum_em / (float)temporal.m_em) / p_hat_em : 0.0f;
      temporal.w_pl = p_hat_pl > 0.0f ? (temporal.w_sum_pl / (float)temporal.pl_m) / p_hat_pl : 0.0f;
    }
  }
  // quirk 4: without usable history the EMPTY temporal reservoir is written (initial candidates are dropped)
  if (in_range && !(P.flags[(size_t)y * fc.W + x] & SNKPT_MAT_FLAG_EMISSIVE)) store_reservoir(half1, (size_t)y * fc.W + x, temporal);
  }
  // ray statistics: one atomic per warp
  unsigned long long n = T.n_any;
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(ray_counter, n);
}

// RayGen.rgen:390-409
__device__ void find_neighbour(const Px& q, float sx, float sy, float radius, float seed_x, float seed_y, float& ox, float& oy) {
  const float W = (float)q.W, H = (float)q.H;
  float offx = (((gold_noise(q.fx, q.fy, fract(seed_x + q.time * 0.0011f)) - 0.5f) * 2.0f) * radius) / W;
  float offy = (((gold_noise(q.fx, q.fy, fract(seed_y + q.time * 0.0006f)) - 0.5f) * 2.0f) * radius) / H;
  int i = 0;
  while ((sx + offx < 0.0f || sy + offy < 0.0f) || (sx + offx > 1.0f || sy + offy > 1.0f)) {
    i++;
    if (i > 64) { offx = 0.0f; offy = 0.0f; break; }  // DEVIATION (as oracle): the reference loop is unbounded
    const float s = sin_c((float)i * 293.28937f);
    const float qx = (float)(q.x + (uint32_t)i), qy = (float)(q.y + (uint32_t)i);
    offx = (((gold_noise(qx, qy, fract((seed_x + s) + q.time * 0.0011f)) - 0.5f) * 2.0f) * radius) / W;
    offy = (((gold_noise(qx, qy, fract((seed_y + s) + q.time * 0.0006f)) - 0.5f) * 2.0f) * radius) / H;
  }
  ox = sx + offx; oy = sy + offy;
}

// ---- pass 2 (direct part) ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(RT_THREADS, RT_CTAS_PER_SM) k_restir_pass2(TraceScene S, SceneDev D, FrameConsts fc, PixelPlanes P,
                                                                             RestirParams rp, unsigned long long* ray_counter) {
  SNK_RT_PROLOGUE();
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  uint32_t x, y; bool in_range;
  tile_xy(fc, tile, tiles_x, x, y, in_range);
  const Px q = load_px(fc, P, x, y, in_range);
  const ResPlanes half0{P.res[0][0], P.res[0][1], P.res[0][2]}, half1{P.res[1][0], P.res[1][1], P.res[1][2]};
  const size_t px = in_range ? (size_t)y * fc.W + x : 0;
  const bool restir = !(rp.flags & SNKPT_RT_NO_RESTIR);
  const float4 dl = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.dir));
  const float4 dc = __ldg(reinterpret_cast<const float4*>(D.lights->dir_light.colour));
  const vec3 ldir = v3(dl.x, dl.y, dl.z), lcol = v3(dc.x, dc.y, dc.z);
  // RayGen.rgen:562
  vec3 light = v3(0.0f);
  {
    const float vis = T.shadow(q.shade, q.biased_pos, ldir, 1000.0f);
    if (q.shade) light = calc_directional_light(ldir, lcol, q.v, q.f0, q.normal, q.roughness, q.metallic, q.albedo) * vis;
  }
  if (restir) {
    // ComputeAndWriteSpatialReservoirReuse (:470-506)
    const bool on = q.shade;
    Reservoir r;
    init_reservoir(r);
    Reservoir own;
    init_reservoir(own);
    if (on) own = load_reservoir(half1, px);
    float p_hat_pl = 0.0f, p_hat_em = 0.0f;
    combine(D, T, q, on, r, own, p_hat_pl, p_hat_em, 0.0f);
    if (on) { r.m_em += own.m_em; r.pl_m += own.pl_m; }
    const float pcx = q.fx / (float)fc.W, pcy = q.fy / (float)fc.H;
    const float seed = ((q.biased_pos.x + q.biased_pos.y) + q.biased_pos.z) + fc.delta_time;
    for (int i = 0; i < rp.num_spatial; ++i) {
      bool take = false;
      Reservoir nb;
      init_reservoir(nb);
      if (on) {
        const float s = sin_c((float)i * 13477.3793f);
        float cx, cy;
        find_neighbour(q, pcx, pcy, rp.radius, seed + s, -seed + s, cx, cy);
        int nx = (int)((float)fc.W * cx), ny = (int)((float)fc.H * cy);
        if (nx > (int)fc.W - 1) nx = (int)fc.W - 1;  // DEVIATION: quirk 10, the reference can index one past the edge
        if (ny > (int)fc.H - 1) ny = (int)fc.H - 1;
        if (nx < 0) nx = 0;
        if (ny < 0) ny = 0;
        const vec3 nn = prev_normal_texel(P, fc.W, nx, ny);  // quirk 5: PREVIOUS-frame normals
        if (!(dot(nn, q.normal) < rp.normal_threshold)) {
          take = true;
          nb = load_reservoir(half1, (size_t)ny * fc.W + nx);
        }
      }
      if (__any_sync(0xFFFFFFFFu, take)) {
        combine(D, T, q, take, r, nb, p_hat_pl, p_hat_em, sin_c((float)i * 234473.344f));
        if (take) { r.m_em += nb.m_em; r.pl_m += nb.pl_m; }
      }
    }
    if (on) {
      r.w_em = p_hat_em > 0.0f ? (r.w_sum_em / p_hat_em) / (float)r.m_em : 0.0f;
      r.w_pl = p_hat_pl > 0.0f ? (r.w_sum_pl / p_hat_pl) / (float)r.pl_m : 0.0f;
    }
    if (in_range && !(P.flags[px] & SNKPT_MAT_FLAG_EMISSIVE)) store_reservoir(half0, px, r);
    // CalculateRadianceFromReservoirs (:508-528)
    vec3 radiance = v3(0.0f);
    {
      const bool valid = on && D.lights->num_pointlights > 0 && r.pl_idx != INVALID_IDX;
      vec3 dir = v3(0.0f, 1.0f, 0.0f), plr = v3(0.0f);
      float light_dist = 1.0f;
      if (valid) {
        const PointLightRec pl = load_pointlight(D.lights, r.pl_idx);
        light_dist = length(pl.pos - q.biased_pos);
        plr = calc_pointlight(pl, q.v, q.f0, q.ro, q.normal, q.roughness, q.metallic, q.albedo);
        dir = (pl.pos - q.biased_pos) / light_dist;
      }
      const float vis = T.shadow(valid, q.biased_pos, dir, light_dist);
      if (valid) { plr = plr * vis; plr *= r.w_pl; radiance += plr; }
    }
    {
      const bool valid = on && D.num_emissive > 0 && r.obj != INVALID_IDX;
      vec3 dir = v3(0.0f, 1.0f, 0.0f), er = v3(0.0f);
      float dist = 1.0f;
      if (valid) {
        const TriPoint tp = triangle_point(D, r.obj, r.tri, r.bx, r.by);
        er = triangle_radiance_unshadowed(tp, q.v, q.f0, q.ro, q.normal, q.roughness, q.metallic, q.albedo);
        triangle_shadow_ray(tp, q.biased_pos, dir, dist);
      }
      const float vis = T.shadow(valid, q.biased_pos, dir, dist);
      if (valid) { er *= vis; radiance += er * r.w_em; }
    }
    if (on) light += radiance;
  }
  if (in_range) P.direct[px] = make_float4(light.x, light.y, light.z, 1.0f);
  }
  unsigned long long n = T.n_any;
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(ray_counter, n);
}

}  // namespace

int launch_restir_passes(SnkPt* ctx, const SnkRtSettings& s, cudaStream_t st) {
  const FrameConsts& fc = ctx->fc;
  const uint32_t n_tiles = ((fc.W + 15) / 16) * ((fc.row1 - fc.row0 + 7) / 8);
  const uint32_t grid = n_tiles < (uint32_t)device_sm_count() * RT_CTAS_PER_SM ? n_tiles : (uint32_t)device_sm_count() * RT_CTAS_PER_SM;
  // pass 1 also covers the halo rows whose temporal reservoirs this band's spatial reuse reads (snkpt_set_halo)
  FrameConsts f1 = fc;
  f1.row0 = fc.row0 > ctx->halo_pass1 ? fc.row0 - ctx->halo_pass1 : 0u;
  f1.row1 = fc.row1 + ctx->halo_pass1 < fc.H ? fc.row1 + ctx->halo_pass1 : fc.H;
  const uint32_t n_tiles1 = ((fc.W + 15) / 16) * ((f1.row1 - f1.row0 + 7) / 8);
  const uint32_t grid1 = n_tiles1 < (uint32_t)device_sm_count() * RT_CTAS_PER_SM ? n_tiles1 : (uint32_t)device_sm_count() * RT_CTAS_PER_SM;
  RestirParams rp{s.num_temporal_resamples, s.num_spatial_resamples, s.ris_iters, s.m_cap, s.neighbour_radius, s.normal_threshold, s.flags};
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(ctx->dev_counters.as<uint32_t>() + 2);
  FrameConsts f = fc;
  f.prev_frame_invalid = (fc.prev_frame_invalid || !ctx->history_valid) ? 1u : 0u;
  f1.prev_frame_invalid = f.prev_frame_invalid;
  int launches = 0;
  cudaEventRecord(ctx->ev[2], st);
  if (!(s.flags & SNKPT_RT_NO_RESTIR)) {
    const int rec = prof_begin(ctx, "restir_pass1", (uint64_t)fc.W * (f1.row1 - f1.row0), false, false, st);
    k_restir_pass1<<<grid1, RT_THREADS, 0, st>>>(ctx->trace_scene(), ctx->scene_dev(), f1, ctx->planes, rp, counter);
    prof_end(ctx, rec, st);
    ++launches;
  }
  cudaEventRecord(ctx->ev[3], st);
  const int rec2 = prof_begin(ctx, "restir_pass2", (uint64_t)fc.W * (fc.row1 - fc.row0), false, false, st);
  k_restir_pass2<<<grid, RT_THREADS, 0, st>>>(ctx->trace_scene(), ctx->scene_dev(), f, ctx->planes, rp, counter);
  prof_end(ctx, rec2, st);
  ++launches;
  cudaEventRecord(ctx->ev[4], st);
  return launches;
}

}  // namespace snk
