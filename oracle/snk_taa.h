// oracle/snk_taa.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// CPU restatement of the reference's TAA resolve, the "next" row N1 of SURVEY §8(f):
//   Core/res/shaders/TAA_Resolve.comp:1-85                    the compute shader
//   Core/src/renderpasses/TAA_ResolvePass.cpp:39-56           dispatch (W/8, H/8) + the two blits (output -> history,
//                                                             velocity -> previous velocity)
//   Core/src/resources/Images.cpp:18-22                       sampler: linear filter, repeat addressing
//   Editor/src/EditorLayer.cpp:389-395, :636-640              the image it runs on: internal_render_image, RGBA16F
// PARITY UNPINNED against a Vulkan driver: `texture()` filtering precision is implementation-defined (the Vulkan
// spec lets the weights be fixed point); here the bilinear weights are float32 and the four taps are summed left to
// right. Everything else is the shader's float32 arithmetic in the order of the GLSL text (-ffp-contract=off).
// Reference quirks reproduced (appendix A of SURVEY.md style):
//   * the shader reads the 3x3 neighbourhood from the image it writes (img_current_frame is both): a data race in the
//     reference; restated as "all reads see the frame as rendered" (the result goes to a separate image). DEVIATION.
//   * dispatch is (uint)(W/8) x (uint)(H/8) groups of 8x8: the right / bottom remainder of a size that is not a
//     multiple of 8 is left as rendered;
//   * history_tex_coord is clamped to [0,1] but the sampler repeats: a coordinate of exactly 0 or 1 blends the
//     opposite border texel with weight 0.5;
//   * the history image is never initialised by the reference; here the first resolve after a reset uses the current
//     frame (and its velocity) as history.
#pragma once
#include <cstdint>
#include <vector>

#include "snk_gbuffer.h"
#include "snk_math.h"

namespace snkor {

struct TaaImages {
  uint32_t W, H;
  const uint16_t* cur;       // RGBA16F current frame (img_current_frame as rendered)
  const uint16_t* velocity;  // RG16F (G-buffer motion plane)
  const uint16_t* history;   // RGBA16F
  const uint16_t* prev_vel;  // RG16F
  uint16_t* out;             // RGBA16F resolved frame (also the next history)
};

inline vec4 taa_clip_aabb(vec3 aabb_min, vec3 aabb_max, vec4 p, vec4 q) {  // TAA_Resolve.comp:12-35
  vec4 r{q.x - p.x, q.y - p.y, q.z - p.z, q.w - p.w};
  const vec3 rmax = aabb_max - v3(p.x, p.y, p.z);
  const vec3 rmin = aabb_min - v3(p.x, p.y, p.z);
  const float eps = 0.00000001f;
  auto scale = [&](float s) { r.x *= s; r.y *= s; r.z *= s; r.w *= s; };
  if (r.x > rmax.x + eps) scale(rmax.x / r.x);
  if (r.y > rmax.y + eps) scale(rmax.y / r.y);
  if (r.z > rmax.z + eps) scale(rmax.z / r.z);
  if (r.x < rmin.x - eps) scale(rmin.x / r.x);
  if (r.y < rmin.y - eps) scale(rmin.y / r.y);
  if (r.z < rmin.z - eps) scale(rmin.z / r.z);
  return vec4{p.x + r.x, p.y + r.y, p.z + r.z, p.w + r.w};
}

// bilinear fetch, repeat addressing, `ch` channels of f16 (Vulkan 1.3 §16.8: i0 = floor(u - 0.5), weight = frac)
inline void taa_sample(const uint16_t* img, int ch, uint32_t W, uint32_t H, float s, float t, float* out) {
  const float u = s * (float)W - 0.5f, v = t * (float)H - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  const int i0 = (int)fu, j0 = (int)fv;
  auto wrap = [](int i, int n) { int m = i % n; return m < 0 ? m + n : m; };
  const uint32_t x0 = (uint32_t)wrap(i0, (int)W), x1 = (uint32_t)wrap(i0 + 1, (int)W);
  const uint32_t y0 = (uint32_t)wrap(j0, (int)H), y1 = (uint32_t)wrap(j0 + 1, (int)H);
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  for (int c = 0; c < ch; ++c) {
    const float t00 = f16_to_f32(img[((size_t)y0 * W + x0) * ch + c]), t10 = f16_to_f32(img[((size_t)y0 * W + x1) * ch + c]);
    const float t01 = f16_to_f32(img[((size_t)y1 * W + x0) * ch + c]), t11 = f16_to_f32(img[((size_t)y1 * W + x1) * ch + c]);
    out[c] = ((t00 * w00 + t10 * w10) + t01 * w01) + t11 * w11;
  }
}

inline void taa_resolve_pixel(const TaaImages& I, uint32_t gx, uint32_t gy) {  // TAA_Resolve.comp:37-85
  const uint32_t W = I.W, H = I.H;
  const size_t px = (size_t)gy * W + gx;
  const float vx = f16_to_f32(I.velocity[2 * px]), vy = f16_to_f32(I.velocity[2 * px + 1]);
  float hx = ((float)gx + 0.5f) / (float)W + vx, hy = ((float)gy + 0.5f) / (float)H + vy;
  hx = gclamp(hx, 0.0f, 1.0f); hy = gclamp(hy, 0.0f, 1.0f);
  const vec3 current = v3(f16_to_f32(I.cur[4 * px]), f16_to_f32(I.cur[4 * px + 1]), f16_to_f32(I.cur[4 * px + 2]));
  vec3 min_col = v3(9999.0f), max_col = v3(-9999.0f), m = v3(0.0f), m2 = v3(0.0f);
  for (int x = -1; x <= 1; ++x) {
    for (int y = -1; y <= 1; ++y) {
      int ox = (int)gx + x, oy = (int)gy + y;
      ox = ox < 0 ? 0 : (ox > (int)W - 1 ? (int)W - 1 : ox);
      oy = oy < 0 ? 0 : (oy > (int)H - 1 ? (int)H - 1 : oy);
      const size_t q = (size_t)oy * W + (size_t)ox;
      const vec3 col = v3(gmax(f16_to_f32(I.cur[4 * q]), 0.0f), gmax(f16_to_f32(I.cur[4 * q + 1]), 0.0f), gmax(f16_to_f32(I.cur[4 * q + 2]), 0.0f));
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
  taa_sample(I.history, 4, W, H, hx, hy, h);
  const vec3 hist = v3(gmax(h[0], 0.0f), gmax(h[1], 0.0f), gmax(h[2], 0.0f));
  const vec3 mu_c = v3(gclamp(mu.x, min_col.x, max_col.x), gclamp(mu.y, min_col.y, max_col.y), gclamp(mu.z, min_col.z, max_col.z));
  const vec4 clipped = taa_clip_aabb(minc, maxc, vec4{mu_c.x, mu_c.y, mu_c.z, 1.0f}, vec4{hist.x, hist.y, hist.z, 1.0f});
  float pv[2];
  taa_sample(I.prev_vel, 2, W, H, hx, hy, pv);
  const float sx = pv[0] + vx, sy = pv[1] + vy;
  const float velocity_length = sqrtf(sx * sx + sy * sy);
  const float disocclusion = gclamp((velocity_length - 0.001f) * 10.0f, 0.0f, 1.0f);
  vec3 fin = mix(current, v3(clipped.x, clipped.y, clipped.z), 0.9f);
  fin = mix(fin, mu, disocclusion);
  I.out[4 * px] = f32_to_f16(fin.x); I.out[4 * px + 1] = f32_to_f16(fin.y); I.out[4 * px + 2] = f32_to_f16(fin.z);
  I.out[4 * px + 3] = f32_to_f16(1.0f);
}

// one resolve: out = cur everywhere, then the dispatched 8x8 groups overwrite their pixels
inline void taa_resolve(const TaaImages& I) {
  for (size_t k = 0; k < (size_t)I.W * I.H * 4; ++k) I.out[k] = I.cur[k];
  const uint32_t gw = (I.W / 8u) * 8u, gh = (I.H / 8u) * 8u;  // TAA_ResolvePass.cpp:48-49
  for (uint32_t y = 0; y < gh; ++y)
    for (uint32_t x = 0; x < gw; ++x) taa_resolve_pixel(I, x, y);
}

}  // namespace snkor
