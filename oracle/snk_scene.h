// oracle/snk_scene.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Scene container in the reference's wire formats + the ray/triangle oracle:
//   * brute-force watertight closest / any hit over instanced triangles
//     (the contract of the opaque Vulkan driver, SURVEY §8a a16: closest hit in
//     (tmin,tmax), watertight shared edges, barycentrics of v1/v2, no culling);
//   * an independent median-split binary BVH (per submesh + over instances) that
//     returns bit-identical results to brute force and makes full frames feasible.
// PARITY UNPINNED: the reference has no tests / golden vectors and its
// traversal lives in the closed Vulkan ICD (TLAS.cpp:74,178; Raytracing.cpp:253,260).
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <tuple>
#include <vector>

#include "snk_math.h"

namespace snkor {

// ---- wire formats (asserted) ------------------------------------------------
struct InstanceData {  // RaytracingInstances.glsl:7-13
  uint32_t transform_idx, mesh_buffer_index_offset, mesh_buffer_vertex_offset, material_idx, num_mesh_indices;
};
static_assert(sizeof(InstanceData) == 20, "InstanceData is 20 B (std430 stride)");

struct Material {  // TexMatBuffers.glsl:5-20
  uint32_t albedo_tex_idx, normal_tex_idx, roughness_tex_idx, metallic_tex_idx, ao_tex_idx;
  float roughness, metallic, ao;
  float albedo[4];
  uint32_t flags, pad0, pad1, pad2;
};
static_assert(sizeof(Material) == 64 && offsetof(Material, albedo) == 32 && offsetof(Material, flags) == 48, "Material");

struct DirectionalLight { float colour[4], dir[4], light_transform[16]; };
struct Pointlight { float colour[4], pos[4], atten_exp, atten_linear, atten_constant, padding; };
struct Spotlight { float colour[4], pos[4], dir[4], atten_exp, atten_linear, atten_constant, aperture; };
struct LightData {  // LightBuffers.glsl:41-49
  DirectionalLight dir_light;
  uint32_t num_pointlights, num_spotlights;
  float pad0, pad1;
  Pointlight pointlights[1024];
  Spotlight spotlights[1024];
};
static_assert(sizeof(DirectionalLight) == 96 && sizeof(Pointlight) == 48 && sizeof(Spotlight) == 64, "lights");
static_assert(offsetof(LightData, num_pointlights) == 96 && offsetof(LightData, pointlights) == 112 &&
                  offsetof(LightData, spotlights) == 49264, "LightData offsets");

struct CommonUBO {  // CommonUBO.glsl:1-13 (std140)
  mat4 view, proj, proj_view;
  float cam_pos[4], cam_forward[4];
  uint32_t frame_idx;
  float time_elapsed, delta_time;
  uint32_t prev_frame_invalid;
};
static_assert(sizeof(CommonUBO) == 240 && offsetof(CommonUBO, frame_idx) == 224, "CommonUBO");

struct PixelReservoir {  // RayGen.rgen:53-76 (std430)
  float w_sum_pl, w_pl;
  uint32_t pl_idx, pl_m;
  uint32_t object_instance_idx, triangle_idx;
  float bary[2];
  float w_sum_em, w_em;
  uint32_t m_em, pad;
};
static_assert(sizeof(PixelReservoir) == 48 && offsetof(PixelReservoir, w_sum_em) == 32, "PixelReservoir");

struct Ray { float o[3], tmin, d[3], tmax; };
struct Hit { float t, u, v; uint32_t prim, inst; };
static_assert(sizeof(Ray) == 32 && sizeof(Hit) == 20, "ray/hit");

constexpr uint32_t INVALID = 0xFFFFFFFFu;
constexpr uint32_t MAT_FLAG_EMISSIVE = 1u;

// ---- watertight ray/triangle test (Woop, Benthin, Wald 2013) ------------------
struct RayPre {
  vec3 o;
  int kx, ky, kz;
  float Sx, Sy, Sz;
};
inline float comp(vec3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
inline RayPre ray_pre(vec3 o, vec3 d) {
  RayPre r;
  r.o = o;
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int kz = 0;
  if (ay > ax) kz = 1;
  if (az > (kz == 0 ? ax : ay)) kz = 2;
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  if (comp(d, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
  r.kx = kx; r.ky = ky; r.kz = kz;
  r.Sx = comp(d, kx) / comp(d, kz);
  r.Sy = comp(d, ky) / comp(d, kz);
  r.Sz = 1.0f / comp(d, kz);
  return r;
}
// returns true and (t,u,v) when the line hits; range test is left to the caller.
// u,v = barycentric weights of v1, v2 (hitAttributeEXT, RayClosestHit.rchit:88).
inline bool tri_test(const RayPre& r, vec3 p0, vec3 p1, vec3 p2, float& t, float& u, float& v) {
  const vec3 A = p0 - r.o, B = p1 - r.o, C = p2 - r.o;
  const float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
  const float Ax = comp(A, r.kx) - r.Sx * Akz, Ay = comp(A, r.ky) - r.Sy * Akz;
  const float Bx = comp(B, r.kx) - r.Sx * Bkz, By = comp(B, r.ky) - r.Sy * Bkz;
  const float Cx = comp(C, r.kx) - r.Sx * Ckz, Cy = comp(C, r.ky) - r.Sy * Ckz;
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
    U = (float)(CxBy - CyBx);
    double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
    V = (float)(AxCy - AyCx);
    double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
    W = (float)(BxAy - ByAx);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  const float det = (U + V) + W;
  if (det == 0.0f) return false;
  const float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
  const float T = (U * Az + V * Bz) + W * Cz;
  const float rcp = 1.0f / det;
  t = T * rcp;
  u = V * rcp;
  v = W * rcp;
  return true;
}

// ---- binary BVH (oracle-private) --------------------------------------------
struct Box {
  float lo[3], hi[3];
  void reset() { for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; } }
  void grow(vec3 p) {
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
  }
  void grow(const Box& b) { for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], b.lo[k]); hi[k] = fmaxf(hi[k], b.hi[k]); } }
  void pad() {  // make the slab test conservative w.r.t. the rounding of the triangle test
    for (int k = 0; k < 3; ++k) {
      float m = fmaxf(fabsf(lo[k]), fabsf(hi[k]));
      float e = 1e-5f * fmaxf(m, 1e-3f);
      lo[k] -= e; hi[k] += e;
    }
  }
};
struct BNode { Box box; int left, right, first, count; };

struct BinaryBvh {
  std::vector<BNode> nodes;
  std::vector<uint32_t> order;  // item ids in leaf order
  void build(const std::vector<Box>& item_boxes) {
    const int n = (int)item_boxes.size();
    order.resize(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    nodes.clear();
    nodes.reserve(2 * n + 1);
    if (n == 0) return;
    std::vector<vec3> cent(n);
    for (int i = 0; i < n; ++i)
      cent[i] = v3(0.5f * (item_boxes[i].lo[0] + item_boxes[i].hi[0]), 0.5f * (item_boxes[i].lo[1] + item_boxes[i].hi[1]),
                   0.5f * (item_boxes[i].lo[2] + item_boxes[i].hi[2]));
    build_rec(item_boxes, cent, 0, n);
  }
  int build_rec(const std::vector<Box>& ib, const std::vector<vec3>& cent, int first, int count) {
    int id = (int)nodes.size();
    nodes.push_back(BNode{});
    Box b; b.reset();
    Box cb; cb.reset();
    for (int i = first; i < first + count; ++i) { b.grow(ib[order[i]]); cb.grow(cent[order[i]]); }
    nodes[id].box = b; nodes[id].first = first; nodes[id].count = count; nodes[id].left = nodes[id].right = -1;
    if (count <= 4) return id;
    int axis = 0;
    float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
    if (ext[1] > ext[axis]) axis = 1;
    if (ext[2] > ext[axis]) axis = 2;
    int mid = first + count / 2;
    std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                     [&](uint32_t a, uint32_t b2) { return comp(cent[a], axis) < comp(cent[b2], axis); });
    nodes[id].count = 0;
    int l = build_rec(ib, cent, first, mid - first);
    int r = build_rec(ib, cent, mid, first + count - mid);
    nodes[id].left = l; nodes[id].right = r;
    return id;
  }
};

inline bool slab(const Box& b, vec3 o, vec3 inv_d, float tmin, float tmax) {
  float t0 = tmin, t1 = tmax;
  const float oo[3] = {o.x, o.y, o.z}, id[3] = {inv_d.x, inv_d.y, inv_d.z};
  for (int k = 0; k < 3; ++k) {
    float a = (b.lo[k] - oo[k]) * id[k], c = (b.hi[k] - oo[k]) * id[k];
    if (a != a || c != c) continue;  // 0 * inf: origin on the slab plane → do not cull
    float n = fminf(a, c), f = fmaxf(a, c);
    n = n - 2e-6f * fabsf(n);  // conservative w.r.t. the rounding of (v - o) in the triangle test
    f = f + 2e-6f * fabsf(f);
    t0 = fmaxf(t0, n); t1 = fminf(t1, f);
  }
  return t0 <= t1;
}

// ---- scene -------------------------------------------------------------------
struct Blas {  // one per distinct (index_offset, vertex_offset, num_indices)
  uint32_t index_offset, vertex_offset, num_tris;
  BinaryBvh bvh;
  Box bounds;
};

// Albedo textures (the `textures[]` descriptor array of TexMatBuffers.glsl, sampled by RayClosestHit.rchit:115-116 and
// GBuffer.frag:57-58). RGBA8; base-colour textures are created R8G8B8A8_SRGB (AssetLoader.cpp:570, AssetManager.cpp:37),
// sampler = linear filter + repeat addressing (Images.cpp:18-22). A ray-tracing stage has no derivatives, so `texture()`
// reads the base level; GBuffer.frag's implicit-LOD fetch is restated at the base level too (DEVIATION: the mip the
// rasteriser would pick is implementation-defined). Filter arithmetic: float32, taps summed left to right (unpinned).
struct Texture {
  uint32_t w = 0, h = 0, srgb = 0;
  std::vector<uint8_t> rgba;
};
inline float srgb8_to_linear(uint32_t c) {  // sRGB EOTF evaluated in float64, rounded once
  const double x = (double)c / 255.0;
  return (float)(x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4));
}
inline vec3 sample_texture_rgb(const Texture& t, float s, float tt) {
  const float u = s * (float)t.w - 0.5f, v = tt * (float)t.h - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  auto wrap = [](float f, uint32_t n) { long long i = (long long)f % (long long)n; return (uint32_t)(i < 0 ? i + (long long)n : i); };
  const uint32_t x0 = wrap(fu, t.w), x1 = wrap(fu + 1.0f, t.w), y0 = wrap(fv, t.h), y1 = wrap(fv + 1.0f, t.h);
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  float out[3];
  for (int c = 0; c < 3; ++c) {
    auto tex = [&](uint32_t x, uint32_t y) {
      const uint32_t q = t.rgba[((size_t)y * t.w + x) * 4 + c];
      return t.srgb ? srgb8_to_linear(q) : (float)q / 255.0f;
    };
    out[c] = ((tex(x0, y0) * w00 + tex(x1, y0) * w10) + tex(x0, y1) * w01) + tex(x1, y1) * w11;
  }
  return v3(out[0], out[1], out[2]);
}

struct Scene {
  std::vector<float> positions, normals, uvs, tangents;  // global SoA buffers (MeshBufferManager)
  std::vector<Texture> textures;
  std::vector<uint32_t> indices;
  std::vector<InstanceData> instances;
  std::vector<mat4> transforms, prev_transforms;
  std::vector<Material> materials;
  LightData lights;
  std::vector<uint32_t> emissive_idx;

  // derived
  std::vector<mat4> world_to_object;        // per transform, float64 inverse rounded once
  std::vector<mat3r> normal_matrix;         // transpose(inverse(mat3(M))), float64 rounded once
  std::vector<float> det3;                  // determinant(mat3(M)), float64 rounded once
  std::vector<std::unique_ptr<Blas>> blas;
  std::vector<int> blas_of_instance;
  BinaryBvh tlas;
  std::vector<uint32_t> tlas_items;         // instance slot of each tlas item

  vec3 pos(uint32_t i) const { return v3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]); }
  vec3 nrm(uint32_t i) const { return v3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]); }
  // GBuffer.vert:49-64 + GBuffer.frag:48-52: per-vertex TBN from the model matrix, interpolated by the hit's barycentrics, applied to
  // the normal texture (UNORM, AssetLoader.cpp:570). Returns false when the material has no usable normal map.
  bool mapped_normal(const Material& m, const mat4& M, uint32_t ia, uint32_t ib, uint32_t ic, float b0, float b1, float b2, vec3& out) const {
    if (m.normal_tex_idx == 0xFFFFFFFFu || m.normal_tex_idx >= textures.size() || textures[m.normal_tex_idx].w == 0) return false;
    if (tangents.size() < positions.size() || uvs.size() < 2 * positions.size() / 3) return false;
    const uint32_t vi[3] = {ia, ib, ic};
    const float w[3] = {b0, b1, b2};
    vec3 T = v3(0.0f), B = v3(0.0f), N = v3(0.0f);
    for (int k = 0; k < 3; ++k) {
      const vec3 tv = v3(tangents[3 * vi[k]], tangents[3 * vi[k] + 1], tangents[3 * vi[k] + 2]), nv = nrm(vi[k]);
      auto m3 = [&](vec3 v) { return v3((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z, (M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z, (M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z); };
      vec3 t = normalize(m3(tv));
      const vec3 n = normalize(m3(nv));
      t = normalize(t - n * dot(t, n));
      const vec3 b = cross(n, t);
      T = k == 0 ? t * w[k] : T + t * w[k]; B = k == 0 ? b * w[k] : B + b * w[k]; N = k == 0 ? n * w[k] : N + n * w[k];
    }
    const float tu = (uvs[2 * ia] * b0 + uvs[2 * ib] * b1) + uvs[2 * ic] * b2;
    const float tv2 = (uvs[2 * ia + 1] * b0 + uvs[2 * ib + 1] * b1) + uvs[2 * ic + 1] * b2;
    const vec3 c = sample_texture_rgb(textures[m.normal_tex_idx], tu, tv2);
    const vec3 tn = normalize(c * 2.0f - v3(1.0f));
    out = normalize((T * tn.x + B * tn.y) + N * tn.z);
    return true;
  }
  // material albedo times the albedo texture at the interpolated texcoord (rchit:100-116; weights b0,b1,b2 of the 3 vertices)
  vec3 albedo_at(const Material& m, uint32_t ia, uint32_t ib, uint32_t ic, float b0, float b1, float b2) const {
    vec3 albedo = v3(m.albedo[0], m.albedo[1], m.albedo[2]);
    if (m.albedo_tex_idx != 0xFFFFFFFFu && m.albedo_tex_idx < textures.size() && textures[m.albedo_tex_idx].w && uvs.size() >= 2 * (size_t)positions.size() / 3) {
      const float tu = (uvs[2 * ia] * b0 + uvs[2 * ib] * b1) + uvs[2 * ic] * b2;
      const float tv = (uvs[2 * ia + 1] * b0 + uvs[2 * ib + 1] * b1) + uvs[2 * ic + 1] * b2;
      albedo = albedo * sample_texture_rgb(textures[m.albedo_tex_idx], tu, tv);
    }
    return albedo;
  }

  void tri_indices(const InstanceData& in, uint32_t prim, uint32_t& a, uint32_t& b, uint32_t& c) const {
    // indices.i[prim + offset/3] + ivec3(vertex_offset)  (RayClosestHit.rchit:91)
    const uint32_t base = 3 * (prim + in.mesh_buffer_index_offset / 3);
    a = indices[base] + in.mesh_buffer_vertex_offset;
    b = indices[base + 1] + in.mesh_buffer_vertex_offset;
    c = indices[base + 2] + in.mesh_buffer_vertex_offset;
  }

  void finalize() {
    const size_t nt = transforms.size();
    world_to_object.resize(nt); normal_matrix.resize(nt); det3.resize(nt);
    for (size_t i = 0; i < nt; ++i) {
      const mat4& M = transforms[i];
      world_to_object[i] = inverse_f(M);
      // inverse of the upper 3x3 in float64 (cofactors), transposed
      double a = M.m[0], b = M.m[4], c = M.m[8], d = M.m[1], e = M.m[5], f = M.m[9], g = M.m[2], h = M.m[6], k = M.m[10];
      double det = a * (e * k - f * h) - b * (d * k - f * g) + c * (d * h - e * g);
      det3[i] = (float)det;
      double id = 1.0 / det;
      double inv[3][3] = {{(e * k - f * h) * id, (c * h - b * k) * id, (b * f - c * e) * id},
                          {(f * g - d * k) * id, (a * k - c * g) * id, (c * d - a * f) * id},
                          {(d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id}};
      for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) normal_matrix[i].r[r][cc] = (float)inv[cc][r];
    }
    std::map<std::tuple<uint32_t, uint32_t, uint32_t>, int> known;
    blas_of_instance.assign(instances.size(), -1);
    std::vector<Box> ibox;
    tlas_items.clear();
    for (size_t i = 0; i < instances.size(); ++i) {
      const InstanceData& in = instances[i];
      if (in.num_mesh_indices < 3 || in.transform_idx >= nt) continue;
      auto key = std::make_tuple(in.mesh_buffer_index_offset, in.mesh_buffer_vertex_offset, in.num_mesh_indices);
      auto it = known.find(key);
      if (it == known.end()) {
        auto bl = std::make_unique<Blas>();
        bl->index_offset = in.mesh_buffer_index_offset; bl->vertex_offset = in.mesh_buffer_vertex_offset;
        bl->num_tris = in.num_mesh_indices / 3;
        std::vector<Box> tb(bl->num_tris);
        bl->bounds.reset();
        for (uint32_t p = 0; p < bl->num_tris; ++p) {
          uint32_t a, b, c; tri_indices(in, p, a, b, c);
          tb[p].reset(); tb[p].grow(pos(a)); tb[p].grow(pos(b)); tb[p].grow(pos(c)); tb[p].pad();
          bl->bounds.grow(tb[p]);
        }
        bl->bvh.build(tb);
        blas.push_back(std::move(bl));
        it = known.emplace(key, (int)blas.size() - 1).first;
      }
      blas_of_instance[i] = it->second;
      // world box = transformed corners of the BLAS bounds
      const Box& ob = blas[it->second]->bounds;
      Box wb; wb.reset();
      for (int cx = 0; cx < 8; ++cx) {
        vec3 p = v3((cx & 1) ? ob.hi[0] : ob.lo[0], (cx & 2) ? ob.hi[1] : ob.lo[1], (cx & 4) ? ob.hi[2] : ob.lo[2]);
        wb.grow(mul_point(transforms[in.transform_idx], p));
      }
      wb.pad();
      ibox.push_back(wb);
      tlas_items.push_back((uint32_t)i);
    }
    tlas.build(ibox);
  }

  // candidate acceptance shared by brute force and BVH paths
  static inline bool better(float t, uint32_t inst, uint32_t prim, const Hit& best) {
    if (best.t < 0.0f) return true;
    if (t < best.t) return true;
    if (t > best.t) return false;
    if (inst != best.inst) return inst < best.inst;
    return prim < best.prim;
  }

  inline void ray_to_object(uint32_t inst, vec3 o, vec3 d, vec3& oo, vec3& od) const {
    const mat4& W = world_to_object[instances[inst].transform_idx];
    oo = mul_point(W, o);
    od = v3((W.m[0] * d.x + W.m[4] * d.y) + W.m[8] * d.z, (W.m[1] * d.x + W.m[5] * d.y) + W.m[9] * d.z,
            (W.m[2] * d.x + W.m[6] * d.y) + W.m[10] * d.z);
  }

  // test one triangle of one instance; updates best (closest) or returns on first (any)
  inline bool test_prim(const RayPre& rp, uint32_t inst, uint32_t prim, float tmin, float tmax, bool any, Hit& best) const {
    uint32_t a, b, c; tri_indices(instances[inst], prim, a, b, c);
    float t, u, v;
    if (!tri_test(rp, pos(a), pos(b), pos(c), t, u, v)) return false;
    if (!(t > tmin && t < tmax)) return false;
    if (any || better(t, inst, prim, best)) { best.t = t; best.u = u; best.v = v; best.prim = prim; best.inst = inst; }
    return true;
  }

  Hit trace_brute(const Ray& r, bool any) const {
    Hit best{-1.0f, 0, 0, INVALID, INVALID};
    const vec3 o = v3(r.o[0], r.o[1], r.o[2]), d = v3(r.d[0], r.d[1], r.d[2]);
    for (uint32_t i = 0; i < instances.size(); ++i) {
      if (blas_of_instance[i] < 0) continue;
      vec3 oo, od; ray_to_object(i, o, d, oo, od);
      RayPre rp = ray_pre(oo, od);
      const uint32_t nt = instances[i].num_mesh_indices / 3;
      for (uint32_t p = 0; p < nt; ++p)
        if (test_prim(rp, i, p, r.tmin, r.tmax, any, best) && any) return best;
    }
    return best;
  }

  Hit trace(const Ray& r, bool any) const {
    Hit best{-1.0f, 0, 0, INVALID, INVALID};
    if (tlas.nodes.empty()) return best;
    const vec3 o = v3(r.o[0], r.o[1], r.o[2]), d = v3(r.d[0], r.d[1], r.d[2]);
    const vec3 inv_d = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
      const BNode& n = tlas.nodes[stack[--sp]];
      const float tfar = (best.t >= 0.0f) ? best.t : r.tmax;
      if (!slab(n.box, o, inv_d, r.tmin, tfar)) continue;
      if (n.left >= 0) { stack[sp++] = n.left; stack[sp++] = n.right; continue; }
      for (int k = n.first; k < n.first + n.count; ++k) {
        const uint32_t inst = tlas_items[tlas.order[k]];
        if (trace_instance(inst, o, d, r.tmin, r.tmax, any, best) && any) return best;
      }
    }
    return best;
  }

  bool trace_instance(uint32_t inst, vec3 o, vec3 d, float tmin, float tmax, bool any, Hit& best) const {
    const Blas& bl = *blas[blas_of_instance[inst]];
    vec3 oo, od; ray_to_object(inst, o, d, oo, od);
    const RayPre rp = ray_pre(oo, od);
    const vec3 inv_d = v3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
    bool found = false;
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
      const BNode& n = bl.bvh.nodes[stack[--sp]];
      const float tfar = (best.t >= 0.0f) ? best.t : tmax;
      if (!slab(n.box, oo, inv_d, tmin, tfar)) continue;
      if (n.left >= 0) { stack[sp++] = n.left; stack[sp++] = n.right; continue; }
      for (int k = n.first; k < n.first + n.count; ++k)
        if (test_prim(rp, inst, bl.bvh.order[k], tmin, tmax, any, best)) { found = true; if (any) return true; }
    }
    return found;
  }
};

}  // namespace snkor
