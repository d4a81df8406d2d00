// oracle/snk_gbuffer.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Primary visibility restated with rays. The reference rasterises its G-buffer
// (GBufferPass.cpp:70-196, GBuffer.vert:66-96, GBuffer.frag:41-72); CUDA has no
// rasteriser, so both the product and this oracle shoot one camera ray per pixel
// (Util.glsl:9-13 CalcRayDir) and write the same quantities in the same storage
// formats (EditorLayer.cpp:229-256): albedo RGBA8_SNORM, normal RGBA16F, depth
// D32F, rma RGBA16F, motion RG16F, mat_flags R16UI (+ prev normal / prev depth).
#pragma once
#include <thread>

#include "snk_scene.h"

namespace snkor {

// GBuffer.frag:41-43 CalcVelocity: UV-space delta from the current to the previous clip position (x 0.5: NDC -> UV)
inline void calc_velocity(const vec4& cur_clip, const vec4& old_clip, float& mx, float& my) {
  mx = (old_clip.x / old_clip.w - cur_clip.x / cur_clip.w) * 0.5f;
  my = (old_clip.y / old_clip.w - cur_clip.y / cur_clip.w) * 0.5f;
}


// ---- storage formats -------------------------------------------------------------
inline uint16_t f32_to_f16(float f) {  // round to nearest even, IEEE binary16
  uint32_t x = f2u(f);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t e = (x >> 23) & 0xFFu, m = x & 0x7FFFFFu;
  if (e == 0xFF) return (uint16_t)(sign | 0x7C00u | (m ? 0x200u | (m >> 13) : 0u));
  int32_t he = (int32_t)e - 127 + 15;
  if (he >= 31) return (uint16_t)(sign | 0x7C00u);
  if (he <= 0) {
    if (he < -10) return (uint16_t)sign;
    m |= 0x800000u;
    uint32_t shift = (uint32_t)(14 - he);
    uint32_t hm = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (hm & 1u))) hm++;
    return (uint16_t)(sign | hm);
  }
  uint32_t hm = m >> 13, rem = m & 0x1FFFu;
  uint32_t h = sign | ((uint32_t)he << 10) | hm;
  if (rem > 0x1000u || (rem == 0x1000u && (hm & 1u))) h++;
  return (uint16_t)h;
}
inline float f16_to_f32(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
  if (e == 0) {
    if (m == 0) return u2f(sign);
    float v = (float)m * (1.0f / 16777216.0f);  // m * 2^-24
    return sign ? -v : v;
  }
  if (e == 31) return u2f(sign | 0x7F800000u | (m << 13));
  return u2f(sign | ((e + 112u) << 23) | (m << 13));
}
inline int8_t f32_to_snorm8(float f) {  // Vulkan: clamp to [-1,1], *127, round to nearest
  if (!(f == f)) return 0;
  float c = f < -1.0f ? -1.0f : (f > 1.0f ? 1.0f : f);
  return (int8_t)lrintf(c * 127.0f);
}
inline float snorm8_to_f32(int8_t v) { float f = (float)v / 127.0f; return f < -1.0f ? -1.0f : f; }

struct GBuffer {
  uint32_t w = 0, h = 0;
  std::vector<int8_t> albedo;      // 4/px
  std::vector<uint16_t> normal;    // 4/px f16
  std::vector<float> depth;        // 1/px
  std::vector<uint16_t> rma;       // 4/px f16
  std::vector<uint16_t> motion;    // 2/px f16
  std::vector<uint16_t> flags;     // 1/px
  std::vector<uint16_t> prev_normal;
  std::vector<float> prev_depth;
  std::vector<uint32_t> prim, inst;  // aux
  std::vector<float> hit_t, bary;    // aux (1, 2 per px)
  void resize(uint32_t W, uint32_t H) {
    w = W; h = H; size_t n = (size_t)W * H;
    albedo.assign(4 * n, 0); normal.assign(4 * n, 0); depth.assign(n, 1.0f); rma.assign(4 * n, 0);
    motion.assign(2 * n, 0); flags.assign(n, 0); prev_normal.assign(4 * n, 0); prev_depth.assign(n, 1.0f);
    prim.assign(n, INVALID); inst.assign(n, INVALID); hit_t.assign(n, -1.0f); bary.assign(2 * n, 0.0f);
  }
};

// hoisted inverse(view) * inverse(proj) (Util.glsl:3): float64, rounded once
inline mat4 inv_view_proj(const CommonUBO& u) {
  double v[16], p[16], pv[16], inv[16];
  for (int i = 0; i < 16; ++i) { v[i] = u.view.m[i]; p[i] = u.proj.m[i]; }
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += p[k * 4 + r] * v[c * 4 + k];
      pv[c * 4 + r] = s;
    }
  mat4 R;
  if (!invert4_d(pv, inv)) { for (int i = 0; i < 16; ++i) R.m[i] = NAN; return R; }
  for (int i = 0; i < 16; ++i) R.m[i] = (float)inv[i];
  return R;
}

// Util.glsl:1-7
inline vec3 world_pos_from_depth(const mat4& inv_vp, float depth, float u, float v) {
  vec4 clip = {u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f};
  vec4 ws = mul_vec4(inv_vp, clip);
  float w = gmax(ws.w, 1e-6f);
  return v3(ws.x / w, ws.y / w, ws.z / w);
}

constexpr float PRIMARY_TMIN = 0.1f;     // CameraComponent.h:11 z_near
constexpr float PRIMARY_TMAX = 10000.0f;

template <class F>
inline void parallel_rows(uint32_t y0, uint32_t y1, int threads, F f) {
  if (threads <= 1) { for (uint32_t y = y0; y < y1; ++y) f(y); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([=]() { for (uint32_t y = y0 + t; y < y1; y += threads) f(y); });
  for (auto& th : pool) th.join();
}

inline void fill_gbuffer(const Scene& sc, const CommonUBO& cur, const CommonUBO& prev, float jx, float jy, GBuffer& g,
                         uint32_t y0, uint32_t y1, int threads) {
  const mat4 ivp = inv_view_proj(cur);
  const vec3 cam = v3(cur.cam_pos[0], cur.cam_pos[1], cur.cam_pos[2]);
  // GBufferPass.cpp:73-77: blit normal/depth to the prev planes before drawing
  g.prev_normal = g.normal;
  g.prev_depth = g.depth;
  parallel_rows(y0, y1, threads, [&](uint32_t y) {
    for (uint32_t x = 0; x < g.w; ++x) {
      const size_t px = (size_t)y * g.w + x;
      const float u = (((float)x + 0.5f) + jx) / (float)g.w, v = (((float)y + 0.5f) + jy) / (float)g.h;
      const vec3 far = world_pos_from_depth(ivp, 1.0f, u, v);
      const vec3 d = normalize(far - cam);
      Ray r{{cam.x, cam.y, cam.z}, PRIMARY_TMIN, {d.x, d.y, d.z}, PRIMARY_TMAX};
      Hit h = sc.trace(r, false);
      g.prim[px] = h.prim; g.inst[px] = h.inst; g.hit_t[px] = h.t;
      g.bary[2 * px] = h.t >= 0 ? h.u : 0.0f; g.bary[2 * px + 1] = h.t >= 0 ? h.v : 0.0f;
      if (h.t < 0.0f) {
        for (int k = 0; k < 4; ++k) { g.albedo[4 * px + k] = 0; g.normal[4 * px + k] = 0; g.rma[4 * px + k] = 0; }
        g.depth[px] = 1.0f; g.motion[2 * px] = g.motion[2 * px + 1] = 0; g.flags[px] = 0;
        continue;
      }
      const InstanceData& in = sc.instances[h.inst];
      uint32_t ia, ib, ic; sc.tri_indices(in, h.prim, ia, ib, ic);
      const float b0 = (1.0f - h.u) - h.v, b1 = h.u, b2 = h.v;
      const vec3 p = (sc.pos(ia) * b0 + sc.pos(ib) * b1) + sc.pos(ic) * b2;
      const vec3 n0 = (sc.nrm(ia) * b0 + sc.nrm(ib) * b1) + sc.nrm(ic) * b2;
      const vec3 world = mul_point(sc.transforms[in.transform_idx], p);
      const Material& mat_n = sc.materials[in.material_idx];
      vec3 n = normalize(mul(sc.normal_matrix[in.transform_idx], n0));
      { vec3 nm; if (sc.mapped_normal(mat_n, sc.transforms[in.transform_idx], ia, ib, ic, b0, b1, b2, nm)) n = nm; }  // GBuffer.frag:48-52
      const vec4 clip = mul_vec4(cur.proj_view, vec4{world.x, world.y, world.z, 1.0f});
      const vec3 pworld = mul_point(sc.prev_transforms[in.transform_idx], p);
      const vec4 pclip = mul_vec4(prev.proj_view, vec4{pworld.x, pworld.y, pworld.z, 1.0f});
      g.depth[px] = clip.z / clip.w;
      float mx, my;
      calc_velocity(clip, pclip, mx, my);
      g.motion[2 * px] = f32_to_f16(mx); g.motion[2 * px + 1] = f32_to_f16(my);
      const Material& m = sc.materials[in.material_idx];
      const vec3 alb = sc.albedo_at(m, ia, ib, ic, b0, b1, b2);  // GBuffer.frag:56-58
      g.albedo[4 * px] = f32_to_snorm8(alb.x); g.albedo[4 * px + 1] = f32_to_snorm8(alb.y);
      g.albedo[4 * px + 2] = f32_to_snorm8(alb.z); g.albedo[4 * px + 3] = 127;
      g.normal[4 * px] = f32_to_f16(n.x); g.normal[4 * px + 1] = f32_to_f16(n.y);
      g.normal[4 * px + 2] = f32_to_f16(n.z); g.normal[4 * px + 3] = f32_to_f16(1.0f);
      g.rma[4 * px] = f32_to_f16(m.roughness); g.rma[4 * px + 1] = f32_to_f16(m.metallic);
      g.rma[4 * px + 2] = f32_to_f16(m.ao); g.rma[4 * px + 3] = f32_to_f16(1.0f);
      g.flags[px] = (uint16_t)m.flags;
    }
  });
}

}  // namespace snkor
