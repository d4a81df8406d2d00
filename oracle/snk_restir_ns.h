// oracle/snk_restir_ns.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// CPU statement of the north-star ReSTIR DI (BASELINE.json north_star item 4: initial candidate generation,
// temporal reuse with motion-vector reprojection, spatial reuse with a bias-corrected MIS weight, reservoirs staged
// as SoA) that snake-vk_b200/csrc/restir_ns.cu must reproduce bit for bit. It keeps the reference's data contract —
// the 48-B PixelReservoir with two independent reservoirs per pixel (point lights; emissive triangles), buffer
// half 0 = final / next frame's history, half 1 = temporal (RayGen.rgen:53-76, EditorLayer.cpp:299-302), the settings of
// RT::RtSettings and the shader constants (RayGen.rgen:95-97, :435, :488, :491) — and replaces the reference's estimator
// quirks:
//   * target function p^ = sum of the RGB contribution (BSDF x cos x emitted / falloff; emissive samples carry the
//     light-side cosine and 1/d^2, source pdf 1 / (n_em x n_tri x area)), evaluated on the G-buffer surface
//   * RNG = PCG hash of (pixel, frame, dimension) instead of gold_noise
//   * W = w_sum / (M p^(y)) with the initial sample's visibility folded in (occluded -> W = 0)
//   * temporal: history at the motion-vector-reprojected pixel, accepted on the previous normal AND the previous
//     world position (from prev depth), M clamped to m_cap x M_initial; the initial sample is kept when history fails
//   * spatial: k neighbours within the radius, accepted on current-frame normal and hit distance; the selected
//     sample's weight is corrected with the MIS weight m_s(y) = p^_s(y) / sum_j p^_j(y) M_j over the pixels that
//     contributed (Bitterli et al. 2020, Alg. 6 with MIS weights). By default the p^_j are unshadowed (biased: darkens where
//     a neighbour could not have produced the sample); with RT_UNBIASED_REUSE every neighbour term carries the visibility of
//     the selected sample from THAT neighbour's surface (one shadow ray per neighbour and light type), which removes the bias
//   * the final contribution is evaluated on the exact primary-hit surface (not the quantised G-buffer), one shadow
//     ray per reservoir + one for the directional light
// Stages (each a loop over pixels here, a kernel + a wavefront shadow-ray launch in the CUDA path):
//   initial -> [trace] -> temporal -> spatial + final evaluation -> [trace] -> resolve
// Numerics contract as in snk_math.h (IEEE binary32 in the order written; no trigonometry is used here).
#pragma once
#include "snk_northstar.h"  // (included from its tail)

namespace snkor {

inline uint32_t rs_seed(uint32_t pixel, uint32_t frame) { return pcg_hash(pcg_hash(pixel) + frame); }
inline float rs_rand(uint32_t h0, uint32_t dim) { return (float)(pcg_hash(h0 + dim) >> 8) * (1.0f / 16777216.0f); }
constexpr uint32_t RS_DIM_INITIAL = 0x1000u, RS_DIM_TEMPORAL = 0x2000u, RS_DIM_SPATIAL = 0x3000u;

// shading point of the reservoir passes
struct RsSurf {
  vec3 n, wo, o, albedo;
  float roughness, metallic;
};
// from the G-buffer planes (what the reference's ray-gen shader reads, RayGen.rgen:534-546)
inline bool rs_gsurf(const Frame& F, const mat4& ivp, vec3 cam, uint32_t x, uint32_t y, bool prev_planes, RsSurf& s, vec3& pos, vec3& n_raw) {
  const GBuffer& g = F.g;
  const size_t px = (size_t)y * g.w + x;
  if (!prev_planes && (g.prim[px] == INVALID || (g.flags[px] & MAT_FLAG_EMISSIVE))) return false;
  const uint16_t* np = prev_planes ? &g.prev_normal[4 * px] : &g.normal[4 * px];
  n_raw = v3(f16_to_f32(np[0]), f16_to_f32(np[1]), f16_to_f32(np[2]));
  const float u = ((float)x + 0.5f) / (float)g.w, v = ((float)y + 0.5f) / (float)g.h;
  pos = world_pos_from_depth(ivp, prev_planes ? g.prev_depth[px] : g.depth[px], u, v);
  if (prev_planes) return true;
  s.albedo = v3(snorm8_to_f32(g.albedo[4 * px]), snorm8_to_f32(g.albedo[4 * px + 1]), snorm8_to_f32(g.albedo[4 * px + 2]));
  s.roughness = gmax(f16_to_f32(g.rma[4 * px]), 0.05f);
  s.metallic = f16_to_f32(g.rma[4 * px + 1]);
  s.wo = normalize(cam - pos);
  s.n = n_raw;
  if (dot(s.n, s.wo) < 0.0f) s.n = -s.n;
  s.o = pos + s.n * 0.01f;
  return true;
}
// from the primary-hit ids (exact surface; the path stages start from the same point)
inline RsSurf rs_xsurf(const NsSurf& h, vec3 cam) {
  RsSurf s;
  s.albedo = h.albedo; s.roughness = h.roughness; s.metallic = h.metallic;
  s.wo = normalize(cam - h.p);
  s.n = h.n;
  if (dot(s.n, s.wo) < 0.0f) s.n = -s.n;
  s.o = h.p + s.n * 0.01f;
  return s;
}

struct RsEval { vec3 C, wi; float tmax; };  // C = unshadowed contribution, tmax < 0: no contribution
inline float rs_phat(const RsEval& e) { return e.tmax > 0.0f ? (e.C.x + e.C.y) + e.C.z : 0.0f; }
inline RsEval rs_eval_pl(const RsSurf& s, const Pointlight& pl) {
  RsEval e; e.C = v3(0.0f); e.wi = v3(0.0f, 1.0f, 0.0f); e.tmax = -1.0f;
  const vec3 d = v3(pl.pos[0], pl.pos[1], pl.pos[2]) - s.o;
  const float dist = length(d);
  if (!(dist > 0.002f)) return e;
  const vec3 wi = inv_scale(d, dist);
  if (!(dot(s.n, wi) > 0.0f)) return e;
  const float atten = (pl.atten_constant + pl.atten_linear * dist) + pl.atten_exp * (dist * dist);
  e.C = inv_scale(ns_bsdf_cos(s.n, s.wo, wi, s.albedo, s.roughness, s.metallic) * v3(pl.colour[0], pl.colour[1], pl.colour[2]), atten);
  e.wi = wi; e.tmax = dist - 0.001f;
  return e;
}
inline RsEval rs_eval_em(const RsSurf& s, const NsSurf& q) {
  RsEval e; e.C = v3(0.0f); e.wi = v3(0.0f, 1.0f, 0.0f); e.tmax = -1.0f;
  const vec3 d = q.p - s.o;
  const float dist = length(d);
  if (!(dist > 0.01f)) return e;
  const vec3 wi = inv_scale(d, dist);
  const float cos_l = fabsf(dot(q.n, wi));
  if (!(dot(s.n, wi) > 0.0f) || !(cos_l > 0.000001f)) return e;
  e.C = inv_scale((ns_bsdf_cos(s.n, s.wo, wi, s.albedo, s.roughness, s.metallic) * q.albedo) * cos_l, dist * dist);
  e.wi = wi; e.tmax = dist - 0.005f;
  return e;
}
inline bool rs_em_valid(const Scene& sc, uint32_t obj, uint32_t tri) {
  return obj < sc.instances.size() && tri < sc.instances[obj].num_mesh_indices / 3u;
}

inline void rs_put_ray(Ray& r, vec3 o, vec3 d, float tmax) {
  r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z; r.tmin = 0.001f; r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z; r.tmax = tmax;
}
inline void rs_dead_ray(Ray& r) { r.o[0] = r.o[1] = r.o[2] = 0.0f; r.tmin = 0.0f; r.d[0] = 0.0f; r.d[1] = 1.0f; r.d[2] = 0.0f; r.tmax = -1.0f; }
inline void rs_empty(PixelReservoir& r) {
  r.w_sum_pl = 0.0f; r.w_pl = 0.0f; r.pl_idx = INVALID; r.pl_m = 0; r.object_instance_idx = INVALID; r.triangle_idx = INVALID;
  r.bary[0] = r.bary[1] = 0.0f; r.w_sum_em = 0.0f; r.w_em = 0.0f; r.m_em = 0; r.pad = 0;
}

// ---- stage 1: initial candidates (RIS) + shadow ray of the selected sample ----------------------------------------
inline void rs_initial_pixel(Frame& F, RsScratch& S, uint32_t x, uint32_t y) {
  const Scene& sc = *F.sc;
  const size_t px = (size_t)y * F.g.w + x;
  const vec3 cam = v3(F.ubo.cam_pos[0], F.ubo.cam_pos[1], F.ubo.cam_pos[2]);
  PixelReservoir r; rs_empty(r);
  rs_dead_ray(S.rays[3 * px]); rs_dead_ray(S.rays[3 * px + 1]);
  S.phat[2 * px] = S.phat[2 * px + 1] = 0.0f;
  RsSurf s; vec3 pos, n_raw;
  if (rs_gsurf(F, F.ivp, cam, x, y, false, s, pos, n_raw)) {
    const uint32_t h0 = rs_seed((uint32_t)px, F.ubo.frame_idx);
    const uint32_t iters = F.settings.ris_iters;
    const uint32_t n_pl = sc.lights.num_pointlights, n_em = (uint32_t)sc.emissive_idx.size();
    if (n_pl > 0) {
      // few lights: every light is a candidate exactly once (stratified RIS); else `iters` uniform picks
      const bool all = n_pl <= iters;
      const uint32_t m = all ? n_pl : iters;
      float p_sel = 0.0f; RsEval e_sel = RsEval{v3(0.0f), v3(0.0f, 1.0f, 0.0f), -1.0f};
      for (uint32_t i = 0; i < m; ++i) {
        const uint32_t idx = all ? i : ns_pick(rs_rand(h0, RS_DIM_INITIAL + 2u * i), n_pl);
        const RsEval e = rs_eval_pl(s, sc.lights.pointlights[idx]);
        const float p_hat = rs_phat(e);
        const float w = p_hat * (float)n_pl;
        r.w_sum_pl += w;
        if (w > 0.0f && rs_rand(h0, RS_DIM_INITIAL + 2u * i + 1u) * r.w_sum_pl < w) { r.pl_idx = idx; p_sel = p_hat; e_sel = e; }
      }
      r.pl_m = m;
      if (r.pl_idx != INVALID) {
        r.w_pl = (r.w_sum_pl / (float)r.pl_m) / p_sel;
        S.phat[2 * px] = p_sel;
        rs_put_ray(S.rays[3 * px], s.o, e_sel.wi, e_sel.tmax);
      }
    }
    if (n_em > 0) {
      float p_sel = 0.0f; RsEval e_sel = RsEval{v3(0.0f), v3(0.0f, 1.0f, 0.0f), -1.0f};
      for (uint32_t i = 0; i < iters; ++i) {
        const uint32_t d = RS_DIM_INITIAL + 0x800u + 5u * i;
        const uint32_t obj = sc.emissive_idx[ns_pick(rs_rand(h0, d), n_em)];
        const uint32_t ntri = sc.instances[obj].num_mesh_indices / 3u;
        const uint32_t tri = ns_pick(rs_rand(h0, d + 1u), ntri);
        float bu = rs_rand(h0, d + 2u), bv = rs_rand(h0, d + 3u);
        if (bu + bv >= 1.0f) { bu = 1.0f - bu; bv = 1.0f - bv; }
        const NsSurf q = ns_light_point(sc, obj, tri, bu, bv);
        const RsEval e = rs_eval_em(s, q);
        const float p_hat = rs_phat(e);
        const float w = (p_hat * ((float)n_em * (float)ntri)) * q.area;
        r.w_sum_em += w;
        if (w > 0.0f && rs_rand(h0, d + 4u) * r.w_sum_em < w) {
          r.object_instance_idx = obj; r.triangle_idx = tri; r.bary[0] = bu; r.bary[1] = bv; p_sel = p_hat; e_sel = e;
        }
      }
      r.m_em = iters;
      if (r.object_instance_idx != INVALID) {
        r.w_em = (r.w_sum_em / (float)r.m_em) / p_sel;
        S.phat[2 * px + 1] = p_sel;
        rs_put_ray(S.rays[3 * px + 1], s.o, e_sel.wi, e_sel.tmax);
      }
    }
  }
  F.reservoirs[(size_t)F.g.w * F.g.h + px] = r;
}

// ---- stage 2: visibility of the initial sample, temporal reuse -> buffer half 1 ------------------------------------
inline void rs_temporal_pixel(Frame& F, RsScratch& S, const mat4& prev_ivp, bool history_ok, uint32_t x, uint32_t y) {
  const Scene& sc = *F.sc;
  const uint32_t W = F.g.w, H = F.g.h;
  const size_t px = (size_t)y * W + x, n = (size_t)W * H;
  const vec3 cam = v3(F.ubo.cam_pos[0], F.ubo.cam_pos[1], F.ubo.cam_pos[2]);
  RsSurf s; vec3 pos, n_raw;
  if (!rs_gsurf(F, F.ivp, cam, x, y, false, s, pos, n_raw)) return;  // stage 1 wrote the empty reservoir
  PixelReservoir r = F.reservoirs[n + px];
  float p_pl = S.phat[2 * px], p_em = S.phat[2 * px + 1];
  if (S.vis[3 * px] == 0.0f) { r.w_sum_pl = 0.0f; r.w_pl = 0.0f; }
  if (S.vis[3 * px + 1] == 0.0f) { r.w_sum_em = 0.0f; r.w_em = 0.0f; }
  if (history_ok && F.settings.num_temporal_resamples != 0) {
    const float hx = ((float)x + 0.5f) / (float)W + f16_to_f32(F.g.motion[2 * px]);
    const float hy = ((float)y + 0.5f) / (float)H + f16_to_f32(F.g.motion[2 * px + 1]);
    if ((hx > 0.0f && hy > 0.0f) && (hx < 1.0f && hy < 1.0f)) {
      uint32_t qx = to_uint(hx * (float)W), qy = to_uint(hy * (float)H);
      if (qx > W - 1) qx = W - 1;
      if (qy > H - 1) qy = H - 1;
      RsSurf unused; vec3 ppos, pn;
      rs_gsurf(F, prev_ivp, cam, qx, qy, true, unused, ppos, pn);
      const vec3 dp = ppos - pos;
      const float tol = 0.02f * length(cam - pos) + 0.02f;
      if (dot(pn, n_raw) > F.settings.normal_threshold && length(dp) < tol) {
        PixelReservoir h = F.reservoirs[(size_t)qy * W + qx];
        const uint32_t h0 = rs_seed((uint32_t)px, F.ubo.frame_idx);
        if (h.pl_m > F.settings.m_cap * r.pl_m) h.pl_m = F.settings.m_cap * r.pl_m;
        if (h.m_em > F.settings.m_cap * r.m_em) h.m_em = F.settings.m_cap * r.m_em;
        if (h.pl_idx < sc.lights.num_pointlights) {
          const float ph = rs_phat(rs_eval_pl(s, sc.lights.pointlights[h.pl_idx]));
          const float w = (ph * h.w_pl) * (float)h.pl_m;
          r.w_sum_pl += w;
          if (w > 0.0f && rs_rand(h0, RS_DIM_TEMPORAL) * r.w_sum_pl < w) { r.pl_idx = h.pl_idx; p_pl = ph; }
        }
        r.pl_m += h.pl_m;
        if (rs_em_valid(sc, h.object_instance_idx, h.triangle_idx)) {
          const NsSurf q = ns_light_point(sc, h.object_instance_idx, h.triangle_idx, h.bary[0], h.bary[1]);
          const float ph = rs_phat(rs_eval_em(s, q));
          const float w = (ph * h.w_em) * (float)h.m_em;
          r.w_sum_em += w;
          if (w > 0.0f && rs_rand(h0, RS_DIM_TEMPORAL + 1u) * r.w_sum_em < w) {
            r.object_instance_idx = h.object_instance_idx; r.triangle_idx = h.triangle_idx; r.bary[0] = h.bary[0]; r.bary[1] = h.bary[1]; p_em = ph;
          }
        }
        r.m_em += h.m_em;
      }
    }
  }
  r.w_pl = (r.pl_idx != INVALID && p_pl > 0.0f && r.pl_m > 0) ? (r.w_sum_pl / (float)r.pl_m) / p_pl : 0.0f;
  r.w_em = (r.object_instance_idx != INVALID && p_em > 0.0f && r.m_em > 0) ? (r.w_sum_em / (float)r.m_em) / p_em : 0.0f;
  F.reservoirs[n + px] = r;
}

// ---- stage 3: spatial reuse with MIS bias correction -> buffer half 0; final contributions + shadow rays -----------
// visibility of a light sample from a neighbour's G-buffer surface (unbiased reuse): the shadow ray rs_eval_* describes
inline float rs_mis_visibility(const Frame& F, const RsSurf& sj, const RsEval& e, uint32_t& rays) {
  Ray r;
  rs_put_ray(r, sj.o, e.wi, e.tmax);
  ++rays;
  return F.sc->trace(r, true).t >= 0.0f ? 0.0f : 1.0f;
}

// returns the shadow rays it traced itself (unbiased reuse only)
inline uint32_t rs_spatial_pixel(Frame& F, RsScratch& S, uint32_t x, uint32_t y) {
  uint32_t mis_rays = 0;
  const bool unbiased = (F.settings.flags & RT_UNBIASED_REUSE) != 0u;
  const Scene& sc = *F.sc;
  const uint32_t W = F.g.w, H = F.g.h;
  const size_t px = (size_t)y * W + x, n = (size_t)W * H;
  const vec3 cam = v3(F.ubo.cam_pos[0], F.ubo.cam_pos[1], F.ubo.cam_pos[2]);
  for (int k = 0; k < 3; ++k) rs_dead_ray(S.rays[3 * px + k]);
  for (int k = 0; k < 9; ++k) S.C[9 * px + k] = 0.0f;
  PixelReservoir out; rs_empty(out);
  RsSurf s; vec3 pos, n_raw;
  if (!rs_gsurf(F, F.ivp, cam, x, y, false, s, pos, n_raw)) { F.reservoirs[px] = out; return 0u; }
  const bool restir = !(F.settings.flags & RT_NO_RESTIR);
  const uint32_t n_pl = sc.lights.num_pointlights;
  if (restir) {
    const uint32_t h0 = rs_seed((uint32_t)px, F.ubo.frame_idx);
    // contributing pixels: slot 0 = this pixel, then the accepted neighbours
    constexpr int MAXN = 9;
    uint32_t cx[MAXN], cy[MAXN];
    int nc = 0;
    cx[nc] = x; cy[nc] = y; ++nc;
    int ns = F.settings.num_spatial_resamples;
    if (ns > MAXN - 1) ns = MAXN - 1;
    const float t_q = F.g.hit_t[px];
    for (int i = 0; i < ns; ++i) {
      const float ua = rs_rand(h0, RS_DIM_SPATIAL + 2u * (uint32_t)i), ub = rs_rand(h0, RS_DIM_SPATIAL + 2u * (uint32_t)i + 1u);
      const int ox = (int)((ua * 2.0f - 1.0f) * F.settings.neighbour_radius), oy = (int)((ub * 2.0f - 1.0f) * F.settings.neighbour_radius);
      const int qx = (int)x + ox, qy = (int)y + oy;
      if ((ox == 0 && oy == 0) || qx < 0 || qy < 0 || qx >= (int)W || qy >= (int)H) continue;
      const size_t qp = (size_t)qy * W + qx;
      if (F.g.prim[qp] == INVALID || (F.g.flags[qp] & MAT_FLAG_EMISSIVE)) continue;
      const vec3 nq = v3(f16_to_f32(F.g.normal[4 * qp]), f16_to_f32(F.g.normal[4 * qp + 1]), f16_to_f32(F.g.normal[4 * qp + 2]));
      if (!(dot(nq, n_raw) >= F.settings.normal_threshold)) continue;
      if (!(fabsf(F.g.hit_t[qp] - t_q) <= 0.1f * t_q)) continue;
      cx[nc] = (uint32_t)qx; cy[nc] = (uint32_t)qy; ++nc;
    }
    // resample the contributors' temporal reservoirs at this pixel
    int src_pl = -1, src_em = -1;
    float p_pl = 0.0f, p_em = 0.0f;
    for (int j = 0; j < nc; ++j) {
      const PixelReservoir& c = F.reservoirs[n + (size_t)cy[j] * W + cx[j]];
      if (c.pl_idx < n_pl) {
        const float ph = rs_phat(rs_eval_pl(s, sc.lights.pointlights[c.pl_idx]));
        const float w = (ph * c.w_pl) * (float)c.pl_m;
        out.w_sum_pl += w;
        if (w > 0.0f && rs_rand(h0, RS_DIM_SPATIAL + 0x100u + 2u * (uint32_t)j) * out.w_sum_pl < w) { out.pl_idx = c.pl_idx; p_pl = ph; src_pl = j; }
      }
      out.pl_m += c.pl_m;
      if (rs_em_valid(sc, c.object_instance_idx, c.triangle_idx)) {
        const NsSurf q = ns_light_point(sc, c.object_instance_idx, c.triangle_idx, c.bary[0], c.bary[1]);
        const float ph = rs_phat(rs_eval_em(s, q));
        const float w = (ph * c.w_em) * (float)c.m_em;
        out.w_sum_em += w;
        if (w > 0.0f && rs_rand(h0, RS_DIM_SPATIAL + 0x100u + 2u * (uint32_t)j + 1u) * out.w_sum_em < w) {
          out.object_instance_idx = c.object_instance_idx; out.triangle_idx = c.triangle_idx; out.bary[0] = c.bary[0]; out.bary[1] = c.bary[1];
          p_em = ph; src_em = j;
        }
      }
      out.m_em += c.m_em;
    }
    // MIS weight of the selected sample (one of the M_s candidates behind source s): m_s(y) = p^_s(y) / sum_j p^_j(y) M_j
    if (src_pl >= 0) {
      float num = 0.0f, den = 0.0f;
      for (int j = 0; j < nc; ++j) {
        float ph = p_pl;
        bool visible = true;
        if (j > 0) {
          RsSurf sj; vec3 pj, nj;
          rs_gsurf(F, F.ivp, cam, cx[j], cy[j], false, sj, pj, nj);
          const RsEval e = rs_eval_pl(sj, sc.lights.pointlights[out.pl_idx]);
          ph = rs_phat(e);
          if (unbiased && ph > 0.0f) visible = rs_mis_visibility(F, sj, e, mis_rays) != 0.0f;
        }
        const float term = ph * (float)F.reservoirs[n + (size_t)cy[j] * W + cx[j]].pl_m;
        den += visible ? term : 0.0f;
        if (j == src_pl) num = visible ? ph : 0.0f;
      }
      out.w_pl = (den > 0.0f && p_pl > 0.0f) ? (out.w_sum_pl * (num / den)) / p_pl : 0.0f;
    }
    if (src_em >= 0) {
      const NsSurf q = ns_light_point(sc, out.object_instance_idx, out.triangle_idx, out.bary[0], out.bary[1]);
      float num = 0.0f, den = 0.0f;
      for (int j = 0; j < nc; ++j) {
        float ph = p_em;
        bool visible = true;
        if (j > 0) {
          RsSurf sj; vec3 pj, nj;
          rs_gsurf(F, F.ivp, cam, cx[j], cy[j], false, sj, pj, nj);
          const RsEval e = rs_eval_em(sj, q);
          ph = rs_phat(e);
          if (unbiased && ph > 0.0f) visible = rs_mis_visibility(F, sj, e, mis_rays) != 0.0f;
        }
        const float term = ph * (float)F.reservoirs[n + (size_t)cy[j] * W + cx[j]].m_em;
        den += visible ? term : 0.0f;
        if (j == src_em) num = visible ? ph : 0.0f;
      }
      out.w_em = (den > 0.0f && p_em > 0.0f) ? (out.w_sum_em * (num / den)) / p_em : 0.0f;
    }
  }
  F.reservoirs[px] = out;
  // final contributions on the exact primary-hit surface
  const NsSurf hs = ns_surface(sc, F.g.inst[px], F.g.prim[px], F.g.bary[2 * px], F.g.bary[2 * px + 1]);
  const RsSurf xs = rs_xsurf(hs, cam);
  float* C = &S.C[9 * px];
  if (restir && out.pl_idx != INVALID && out.w_pl > 0.0f) {
    const RsEval e = rs_eval_pl(xs, sc.lights.pointlights[out.pl_idx]);
    if (e.tmax > 0.0f) {
      const vec3 c = e.C * out.w_pl;
      C[0] = c.x; C[1] = c.y; C[2] = c.z;
      rs_put_ray(S.rays[3 * px], xs.o, e.wi, e.tmax);
    }
  }
  if (restir && out.object_instance_idx != INVALID && out.w_em > 0.0f) {
    const NsSurf q = ns_light_point(sc, out.object_instance_idx, out.triangle_idx, out.bary[0], out.bary[1]);
    const RsEval e = rs_eval_em(xs, q);
    if (e.tmax > 0.0f) {
      const vec3 c = e.C * out.w_em;
      C[3] = c.x; C[4] = c.y; C[5] = c.z;
      rs_put_ray(S.rays[3 * px + 1], xs.o, e.wi, e.tmax);
    }
  }
  {
    const vec3 wi = v3(sc.lights.dir_light.dir[0], sc.lights.dir_light.dir[1], sc.lights.dir_light.dir[2]);
    const vec3 col = v3(sc.lights.dir_light.colour[0], sc.lights.dir_light.colour[1], sc.lights.dir_light.colour[2]);
    if ((col.x > 0.0f || col.y > 0.0f || col.z > 0.0f) && dot(xs.n, wi) > 0.0f) {
      const vec3 c = ns_bsdf_cos(xs.n, xs.wo, wi, xs.albedo, xs.roughness, xs.metallic) * col;
      C[6] = c.x; C[7] = c.y; C[8] = c.z;
      rs_put_ray(S.rays[3 * px + 2], xs.o, wi, 10000.0f);
    }
  }
  return mis_rays;
}

inline uint64_t rs_trace(const Frame& F, RsScratch& S, uint32_t nplanes, uint32_t y0, uint32_t y1, int threads) {
  std::atomic<uint64_t> total{0};
  parallel_rows(y0, y1, threads, [&](uint32_t y) {
    uint64_t na = 0;
    for (uint32_t x = 0; x < F.g.w; ++x)
      for (uint32_t k = 0; k < nplanes; ++k) {
        const size_t i = 3 * ((size_t)y * F.g.w + x) + k;
        const Ray& r = S.rays[i];
        S.vis[i] = 1.0f;
        if (r.tmax > r.tmin) {
          ++na;
          S.vis[i] = F.sc->trace(r, true).t >= 0.0f ? 0.0f : 1.0f;
        }
      }
    total += na;
  });
  return total.load();
}

// direct term of rows [y0,y1): `halo` rows around the band also run stages 1-2 (their temporal reservoirs are what the
// band's spatial reuse reads); F.direct_rgba and both reservoir halves are written like the reference-mode passes do
inline void render_restir_ns(Frame& F, RsScratch& S, bool history_ok, uint32_t y0, uint32_t y1, uint32_t halo, int threads) {
  const uint32_t W = F.g.w, H = F.g.h;
  if (S.phat.size() != 2 * (size_t)W * H) S.resize((size_t)W * H);
  F.ivp = inv_view_proj(F.ubo);
  const mat4 prev_ivp = inv_view_proj(F.prev_ubo);
  const uint32_t a0 = y0 > halo ? y0 - halo : 0u, a1 = y1 + halo < H ? y1 + halo : H;
  const bool restir = !(F.settings.flags & RT_NO_RESTIR);
  if (restir) {
    parallel_rows(a0, a1, threads, [&](uint32_t y) { for (uint32_t x = 0; x < W; ++x) rs_initial_pixel(F, S, x, y); });
    F.rays_any += rs_trace(F, S, 2, a0, a1, threads);
    parallel_rows(a0, a1, threads, [&](uint32_t y) { for (uint32_t x = 0; x < W; ++x) rs_temporal_pixel(F, S, prev_ivp, history_ok, x, y); });
  }
  std::atomic<uint64_t> mis_total{0};
  parallel_rows(y0, y1, threads, [&](uint32_t y) {
    uint64_t c = 0;
    for (uint32_t x = 0; x < W; ++x) c += rs_spatial_pixel(F, S, x, y);
    mis_total += c;
  });
  F.rays_any += mis_total.load();
  F.rays_any += rs_trace(F, S, 3, y0, y1, threads);
  parallel_rows(y0, y1, threads, [&](uint32_t y) {
    for (uint32_t x = 0; x < W; ++x) {
      const size_t px = (size_t)y * W + x;
      const float* C = &S.C[9 * px];
      const float* v = &S.vis[3 * px];
      float* d = &F.direct_rgba[4 * px];
      for (int c = 0; c < 3; ++c) d[c] = (C[c] * v[0] + C[3 + c] * v[1]) + C[6 + c] * v[2];
      d[3] = 1.0f;
    }
  });
}

}  // namespace snkor
