// taa.cu — the reference's TAA resolve as one CUDA kernel.
//
// Replaces TAA_ResolvePass::RecordCommandBuffer (Core/src/renderpasses/TAA_ResolvePass.cpp:39-56): the compute dispatch of
// Core/res/shaders/TAA_Resolve.comp:1-85 over the RGBA16F output image (Editor/src/EditorLayer.cpp:389-395, :636-640)
// and the two blits that follow it (output -> history, velocity -> previous velocity). CPU statement: oracle/snk_taa.h
// (which lists the reference quirks kept); this file must match it bit for bit (-fmad=false, same expression order).
// HBM-bound: compulsory traffic per pixel = 8 B frame + 8 B history + 4 B velocity + 4 B previous velocity read, 8 B
// resolved + 8 B output + 4 B velocity copy written = 44 B; the 3x3 neighbourhood and the bilinear taps come from L1/L2.
#include "kernels.h"
#include "shading.cuh"

namespace snk {

namespace {

struct vec4f { float x, y, z, w; };

__device__ __forceinline__ void taa_sample4(const ushort4* __restrict__ img, uint32_t W, uint32_t H, float s, float t, float* out) {
  const float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  int i0 = (int)fu, j0 = (int)fv;
  int x0 = i0 % (int)W; if (x0 < 0) x0 += (int)W;
  int x1 = (i0 + 1) % (int)W; if (x1 < 0) x1 += (int)W;
  int y0 = j0 % (int)H; if (y0 < 0) y0 += (int)H;
  int y1 = (j0 + 1) % (int)H; if (y1 < 0) y1 += (int)H;
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  const ushort4 t00 = __ldg(img + (size_t)y0 * W + x0), t10 = __ldg(img + (size_t)y0 * W + x1);
  const ushort4 t01 = __ldg(img + (size_t)y1 * W + x0), t11 = __ldg(img + (size_t)y1 * W + x1);
  out[0] = ((h2f(t00.x) * w00 + h2f(t10.x) * w10) + h2f(t01.x) * w01) + h2f(t11.x) * w11;
  out[1] = ((h2f(t00.y) * w00 + h2f(t10.y) * w10) + h2f(t01.y) * w01) + h2f(t11.y) * w11;
  out[2] = ((h2f(t00.z) * w00 + h2f(t10.z) * w10) + h2f(t01.z) * w01) + h2f(t11.z) * w11;
  out[3] = ((h2f(t00.w) * w00 + h2f(t10.w) * w10) + h2f(t01.w) * w01) + h2f(t11.w) * w11;
}
__device__ __forceinline__ void taa_sample2(const ushort2* __restrict__ img, uint32_t W, uint32_t H, float s, float t, float* out) {
  const float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  int i0 = (int)fu, j0 = (int)fv;
  int x0 = i0 % (int)W; if (x0 < 0) x0 += (int)W;
  int x1 = (i0 + 1) % (int)W; if (x1 < 0) x1 += (int)W;
  int y0 = j0 % (int)H; if (y0 < 0) y0 += (int)H;
  int y1 = (j0 + 1) % (int)H; if (y1 < 0) y1 += (int)H;
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  const ushort2 t00 = __ldg(img + (size_t)y0 * W + x0), t10 = __ldg(img + (size_t)y0 * W + x1);
  const ushort2 t01 = __ldg(img + (size_t)y1 * W + x0), t11 = __ldg(img + (size_t)y1 * W + x1);
  out[0] = ((h2f(t00.x) * w00 + h2f(t10.x) * w10) + h2f(t01.x) * w01) + h2f(t11.x) * w11;
  out[1] = ((h2f(t00.y) * w00 + h2f(t10.y) * w10) + h2f(t01.y) * w01) + h2f(t11.y) * w11;
}

// TAA_Resolve.comp:12-35
__device__ __forceinline__ vec4f clip_aabb(vec3 aabb_min, vec3 aabb_max, vec4f p, vec4f q) {
  vec4f r{q.x - p.x, q.y - p.y, q.z - p.z, q.w - p.w};
  const vec3 rmax = aabb_max - v3(p.x, p.y, p.z);
  const vec3 rmin = aabb_min - v3(p.x, p.y, p.z);
  const float eps = 0.00000001f;
#define SNK_SCALE(S) { const float s_ = (S); r.x *= s_; r.y *= s_; r.z *= s_; r.w *= s_; }
  if (r.x > rmax.x + eps) SNK_SCALE(rmax.x / r.x)
  if (r.y > rmax.y + eps) SNK_SCALE(rmax.y / r.y)
  if (r.z > rmax.z + eps) SNK_SCALE(rmax.z / r.z)
  if (r.x < rmin.x - eps) SNK_SCALE(rmin.x / r.x)
  if (r.y < rmin.y - eps) SNK_SCALE(rmin.y / r.y)
  if (r.z < rmin.z - eps) SNK_SCALE(rmin.z / r.z)
#undef SNK_SCALE
  return vec4f{p.x + r.x, p.y + r.y, p.z + r.z, p.w + r.w};
}

// one thread per pixel of the frame; pixels outside the reference's (W/8) x (H/8) dispatch keep the rendered value.
// Writes the resolved frame to `out` (= the next history image) and copies this frame's velocity to `vel_out`
// (= the next "previous velocity"); the caller then copies `out` over the output image (the in-place store of the shader).
__global__ void __launch_bounds__(256) k_taa_resolve(const ushort4* __restrict__ cur, const ushort2* __restrict__ velocity,
                                                     const ushort4* __restrict__ history, const ushort2* __restrict__ prev_vel,
                                                     ushort4* __restrict__ out, ushort2* __restrict__ vel_out, uint32_t W, uint32_t H) {
  const uint32_t gx = blockIdx.x * 32u + (threadIdx.x & 31u), gy = blockIdx.y * 8u + (threadIdx.x >> 5);
  if (gx >= W || gy >= H) return;
  const size_t px = (size_t)gy * W + gx;
  const ushort2 v16 = velocity[px];
  vel_out[px] = v16;
  const ushort4 c16 = cur[px];
  if (gx >= (W / 8u) * 8u || gy >= (H / 8u) * 8u) { out[px] = c16; return; }  // TAA_ResolvePass.cpp:48-49
  const float vx = h2f(v16.x), vy = h2f(v16.y);
  float hx = ((float)gx + 0.5f) / (float)W + vx, hy = ((float)gy + 0.5f) / (float)H + vy;
  hx = gclamp(hx, 0.0f, 1.0f); hy = gclamp(hy, 0.0f, 1.0f);
  const vec3 current = v3(h2f(c16.x), h2f(c16.y), h2f(c16.z));
  vec3 min_col = v3(9999.0f), max_col = v3(-9999.0f), m = v3(0.0f), m2 = v3(0.0f);
#pragma unroll
  for (int x = -1; x <= 1; ++x) {
#pragma unroll
    for (int y = -1; y <= 1; ++y) {
      int ox = (int)gx + x, oy = (int)gy + y;
      ox = ox < 0 ? 0 : (ox > (int)W - 1 ? (int)W - 1 : ox);
      oy = oy < 0 ? 0 : (oy > (int)H - 1 ? (int)H - 1 : oy);
      const ushort4 t = __ldg(cur + (size_t)oy * W + (size_t)ox);
      const vec3 col = v3(gmax(h2f(t.x), 0.0f), gmax(h2f(t.y), 0.0f), gmax(h2f(t.z), 0.0f));
      min_col = v3(gmin(min_col.x, col.x), gmin(min_col.y, col.y), gmin(min_col.z, col.z));
      max_col = v3(gmax(max_col.x, col.x), gmax(max_col.y, col.y), gmax(max_col.z, col.z));
      m += col;
      m2 += col * col;
    }
  }
  const float inv_n = 1.0f / 9.0f, gamma = 0.55f;
  const vec3 mu = m * inv_n;
  const vec3 var = (m2 * inv_n) - (mu * mu);
  const vec3 sigma = v3(sqrtf(fabsf(var.x)), sqrtf(fabsf(var.y)), sqrtf(fabsf(var.z)));
  const vec3 minc = mu - gamma * sigma, maxc = mu + gamma * sigma;
  float h[4];
  taa_sample4(history, W, H, hx, hy, h);
  const vec3 hist = v3(gmax(h[0], 0.0f), gmax(h[1], 0.0f), gmax(h[2], 0.0f));
  const vec3 mu_c = v3(gclamp(mu.x, min_col.x, max_col.x), gclamp(mu.y, min_col.y, max_col.y), gclamp(mu.z, min_col.z, max_col.z));
  const vec4f clipped = clip_aabb(minc, maxc, vec4f{mu_c.x, mu_c.y, mu_c.z, 1.0f}, vec4f{hist.x, hist.y, hist.z, 1.0f});
  float pv[2];
  taa_sample2(prev_vel, W, H, hx, hy, pv);
  const float sx = pv[0] + vx, sy = pv[1] + vy;
  const float velocity_length = sqrtf(sx * sx + sy * sy);
  const float disocclusion = gclamp((velocity_length - 0.001f) * 10.0f, 0.0f, 1.0f);
  vec3 fin = mix(current, v3(clipped.x, clipped.y, clipped.z), 0.9f);
  fin = mix(fin, mu, disocclusion);
  out[px] = make_ushort4(f2h(fin.x), f2h(fin.y), f2h(fin.z), f2h(1.0f));
}

}  // namespace

int launch_taa_resolve(const ushort4* cur, const ushort2* velocity, const ushort4* history, const ushort2* prev_vel, ushort4* out,
                       ushort2* vel_out, uint32_t W, uint32_t H, cudaStream_t st) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8);
  k_taa_resolve<<<grid, 256, 0, st>>>(cur, velocity, history, prev_vel, out, vel_out, W, H);
  return 1;
}

}  // namespace snk
