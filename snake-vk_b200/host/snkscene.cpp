// snkscene.cpp — host-side scene front-end (include/snkscene.h): entity hierarchy -> the flat buffers of the path-tracing pass.
// C++ restatement of the reference's scene systems (file:line citations in the header and at each function). glm's float32
// arithmetic is followed operation by operation (quat from Euler angles, mat3_cast, perspectiveRH_ZO, lookAtRH, mat4 product)
// so the matrices agree with the engine's to rounding; tests/test_cpu_scene.py checks them against the reference's vendored
// glm (oracle/_ref) and the committed golden vectors.
#include "snkscene.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }  // glm: v * inversesqrt(dot(v, v))

struct M4 { float m[16]; float& at(int c, int r) { return m[c * 4 + r]; } float at(int c, int r) const { return m[c * 4 + r]; } };
inline M4 identity() { M4 r{}; r.at(0, 0) = r.at(1, 1) = r.at(2, 2) = r.at(3, 3) = 1.0f; return r; }
// glm mat4 * mat4: Result[c] = A[0]*B[c][0] + A[1]*B[c][1] + A[2]*B[c][2] + A[3]*B[c][3]
inline M4 mul(const M4& A, const M4& B) {
  M4 R;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      R.at(c, r) = ((A.at(0, r) * B.at(c, 0) + A.at(1, r) * B.at(c, 1)) + A.at(2, r) * B.at(c, 2)) + A.at(3, r) * B.at(c, 3);
  return R;
}
// ExtraMath::Rotate3D (ExtraMath.cpp:148-151): mat3_cast(quat(radians(euler)))
inline void rotate3d(V3 euler_deg, float R[3][3]) {
  const float k = 0.01745329251994329576923690768489f;
  const V3 e{euler_deg.x * k, euler_deg.y * k, euler_deg.z * k};
  const V3 c{std::cos(e.x * 0.5f), std::cos(e.y * 0.5f), std::cos(e.z * 0.5f)}, s{std::sin(e.x * 0.5f), std::sin(e.y * 0.5f), std::sin(e.z * 0.5f)};
  const float w = c.x * c.y * c.z + s.x * s.y * s.z, x = s.x * c.y * c.z - c.x * s.y * s.z, y = c.x * s.y * c.z + s.x * c.y * s.z,
              z = c.x * c.y * s.z - s.x * s.y * c.z;
  const float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
  R[0][0] = 1.0f - 2.0f * (qyy + qzz); R[0][1] = 2.0f * (qxy + qwz); R[0][2] = 2.0f * (qxz - qwy);
  R[1][0] = 2.0f * (qxy - qwz); R[1][1] = 1.0f - 2.0f * (qxx + qzz); R[1][2] = 2.0f * (qyz + qwx);
  R[2][0] = 2.0f * (qxz + qwy); R[2][1] = 2.0f * (qyz - qwx); R[2][2] = 1.0f - 2.0f * (qxx + qyy);
}
inline M4 perspective_vk(float fov, float aspect, float z_near, float z_far) {  // CameraComponent.h:15-19
  const float fovy = (fov * 0.5f) * 0.01745329251994329576923690768489f;
  const float t = std::tan(fovy / 2.0f);
  M4 r{};
  r.at(0, 0) = 1.0f / (aspect * t);
  r.at(1, 1) = 1.0f / t;
  r.at(2, 2) = z_far / (z_near - z_far);
  r.at(2, 3) = -1.0f;
  r.at(3, 2) = -(z_far * z_near) / (z_far - z_near);
  r.at(1, 1) *= -1.0f;  // flip y for Vulkan
  return r;
}
inline M4 look_at(V3 eye, V3 center, V3 up) {  // glm::lookAtRH
  const V3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
  M4 r = identity();
  r.at(0, 0) = s.x; r.at(1, 0) = s.y; r.at(2, 0) = s.z;
  r.at(0, 1) = u.x; r.at(1, 1) = u.y; r.at(2, 1) = u.z;
  r.at(0, 2) = -f.x; r.at(1, 2) = -f.y; r.at(2, 2) = -f.z;
  r.at(3, 0) = -dot(s, eye); r.at(3, 1) = -dot(u, eye); r.at(3, 2) = dot(f, eye);
  return r;
}
inline M4 ortho_zo(float l, float r_, float b, float t, float n, float f) {  // glm::orthoRH_ZO
  M4 r = identity();
  r.at(0, 0) = 2.0f / (r_ - l); r.at(1, 1) = 2.0f / (t - b); r.at(2, 2) = -1.0f / (f - n);
  r.at(3, 0) = -(r_ + l) / (r_ - l); r.at(3, 1) = -(t + b) / (t - b); r.at(3, 2) = -n / (f - n);
  return r;
}

struct Entity {
  uint32_t parent = SNKSCENE_NONE;
  std::vector<uint32_t> children;
  // TransformComponent
  V3 pos{0, 0, 0}, orientation{0, 0, 0}, scale{1, 1, 1};
  V3 abs_pos{0, 0, 0}, abs_orientation{0, 0, 0}, abs_scale{1, 1, 1};
  V3 forward{0, 0, -1}, right{1, 0, 0}, up{0, 1, 0};
  bool is_absolute = false;
  M4 transform = identity();
  // StaticMeshComponent + RaytracingInstanceBufferIdxComponent
  bool has_mesh = false;
  uint32_t mesh = 0, rt_instance_idx = 0;
  std::vector<uint32_t> material_slots;
  // PointlightComponent
  bool has_pointlight = false;
  V3 pl_colour{1, 1, 1};
  float pl_exp = 0.05f, pl_linear = 0.25f, pl_constant = 0.5f;  // Lights.h:6-10 defaults
};
struct MeshRec { SnkMeshEntry entry; std::vector<SnkSubmesh> subs; };

}  // namespace

struct SnkScene {
  std::string err;
  std::vector<Entity> ents;
  std::vector<MeshRec> meshes;
  std::vector<SnkMaterial> materials;
  std::vector<uint32_t> mesh_order, pointlight_order;  // insertion order of the component pools (views iterate it backwards)
  uint32_t next_instance_slot = 0;
  V3 dir_colour{0, 0, 0};
  float dir_theta = 0.f, dir_phi = 0.f;
  uint32_t camera = SNKSCENE_NONE;
  float fov = 90.f, aspect = 16.f / 9.f, z_near = 0.1f, z_far = 2500.f;
  uint32_t frame_idx = 0;
  bool prev_frame_invalid = true;
  // flat buffers
  std::vector<float> transforms, prev_transforms;
  std::vector<SnkInstanceData> instances;
  std::vector<uint32_t> blas_of_instance, emissive;
  std::vector<uint8_t> lights;
  SnkCommonUBO ubo{}, prev_ubo{};
  bool ubo_valid = false;
};

namespace {

#define CHECK(cond, msg) do { if (!(cond)) { sc->err = msg; return SNKPT_E_INVALID; } } while (0)

// TransformSystem::UpdateAbsTransforms (TransformSystem.cpp:31-48)
void update_abs(SnkScene* sc, uint32_t e) {
  Entity& c = sc->ents[e];
  c.abs_scale = c.scale;
  c.abs_orientation = c.orientation;
  c.abs_pos = V3{c.transform.at(3, 0), c.transform.at(3, 1), c.transform.at(3, 2)};
  if (c.is_absolute) return;
  for (uint32_t p = c.parent; p != SNKSCENE_NONE; p = sc->ents[p].parent) {
    c.abs_orientation = c.abs_orientation + sc->ents[p].orientation;
    c.abs_scale = c.abs_scale * sc->ents[p].scale;
  }
}

// TransformSystem::RebuildMatrix (TransformSystem.cpp:50-114) + UpdateChildren (:116-121)
void rebuild_matrix(SnkScene* sc, uint32_t e) {
  Entity& c = sc->ents[e];
  float R[3][3];
  rotate3d(c.orientation, R);
  M4 t = identity();
  for (int col = 0; col < 3; ++col)
    for (int r = 0; r < 3; ++r) t.at(col, r) = R[col][r];
  const bool rooted = c.is_absolute || c.parent == SNKSCENE_NONE;
  if (rooted) {
    const float s[3] = {c.scale.x, c.scale.y, c.scale.z};
    for (int col = 0; col < 3; ++col)
      for (int r = 0; r < 3; ++r) t.at(col, r) *= s[col];
  } else {
    // apply the total scale while undoing the parent's, to prevent shearing (:79-97)
    const V3 ps = sc->ents[c.parent].abs_scale;
    const float inv_det = 1.f / (ps.x * ps.y * ps.z);
    const V3 total = c.scale * ps;
    const float inv_scale[3] = {ps.y * ps.z * inv_det, ps.x * ps.z * inv_det, ps.y * ps.x * inv_det};
    const float tot[3] = {total.x, total.y, total.z};
    for (int col = 0; col < 3; ++col)
      for (int r = 0; r < 3; ++r) t.at(col, r) *= tot[col] * inv_scale[r];
  }
  t.at(3, 0) = c.pos.x; t.at(3, 1) = c.pos.y; t.at(3, 2) = c.pos.z; t.at(3, 3) = 1.0f;
  c.transform = rooted ? t : mul(sc->ents[c.parent].transform, t);
  c.forward = normalize(V3{-c.transform.at(2, 0), -c.transform.at(2, 1), -c.transform.at(2, 2)});
  c.right = normalize(cross(c.forward, V3{0.f, 1.f, 0.f}));
  c.up = normalize(cross(c.right, c.forward));
  update_abs(sc, e);
  for (size_t k = 0; k < sc->ents[e].children.size(); ++k) rebuild_matrix(sc, sc->ents[e].children[k]);
}

bool valid_ent(const SnkScene* sc, uint32_t e) { return e < sc->ents.size(); }

}  // namespace

extern "C" {

SnkScene* snkscene_create(void) { return new SnkScene(); }
void snkscene_destroy(SnkScene* sc) { delete sc; }
const char* snkscene_last_error(const SnkScene* sc) { return sc ? sc->err.c_str() : "null scene"; }

uint32_t snkscene_create_entity(SnkScene* sc, uint32_t parent) {
  if (!sc || (parent != SNKSCENE_NONE && !valid_ent(sc, parent))) return SNKSCENE_NONE;
  const uint32_t id = (uint32_t)sc->ents.size();
  sc->ents.emplace_back();
  sc->ents[id].parent = parent;
  if (parent != SNKSCENE_NONE) sc->ents[parent].children.push_back(id);
  rebuild_matrix(sc, id);
  return id;
}

int snkscene_set_parent(SnkScene* sc, uint32_t ent, uint32_t parent) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(valid_ent(sc, ent) && (parent == SNKSCENE_NONE || valid_ent(sc, parent)), "snkscene_set_parent: unknown entity");
  for (uint32_t p = parent; p != SNKSCENE_NONE; p = sc->ents[p].parent) CHECK(p != ent, "snkscene_set_parent: cycle");
  Entity& e = sc->ents[ent];
  if (e.parent != SNKSCENE_NONE) {
    auto& ch = sc->ents[e.parent].children;
    for (size_t k = 0; k < ch.size(); ++k) if (ch[k] == ent) { ch.erase(ch.begin() + k); break; }
  }
  e.parent = parent;
  if (parent != SNKSCENE_NONE) sc->ents[parent].children.push_back(ent);
  rebuild_matrix(sc, ent);  // RelationshipComponent UPDATED -> RebuildMatrix (TransformSystem.cpp:14-18)
  return SNKPT_OK;
}

int snkscene_set_transform(SnkScene* sc, uint32_t ent, const float* pos3, const float* euler_deg3, const float* scale3) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(valid_ent(sc, ent), "snkscene_set_transform: unknown entity");
  Entity& e = sc->ents[ent];
  if (pos3) e.pos = V3{pos3[0], pos3[1], pos3[2]};
  if (euler_deg3) e.orientation = V3{euler_deg3[0], euler_deg3[1], euler_deg3[2]};
  if (scale3) e.scale = V3{scale3[0], scale3[1], scale3[2]};
  rebuild_matrix(sc, ent);
  return SNKPT_OK;
}

int snkscene_set_absolute(SnkScene* sc, uint32_t ent, int absolute) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(valid_ent(sc, ent), "snkscene_set_absolute: unknown entity");
  sc->ents[ent].is_absolute = absolute != 0;
  rebuild_matrix(sc, ent);
  return SNKPT_OK;
}

int snkscene_get_matrix(const SnkScene* sc, uint32_t ent, float* out16) {
  if (!sc || !valid_ent(sc, ent) || !out16) return SNKPT_E_INVALID;
  memcpy(out16, sc->ents[ent].transform.m, 64);
  return SNKPT_OK;
}

int snkscene_get_abs(const SnkScene* sc, uint32_t ent, float* pos3, float* scale3, float* orientation3, float* forward3) {
  if (!sc || !valid_ent(sc, ent)) return SNKPT_E_INVALID;
  const Entity& e = sc->ents[ent];
  if (pos3) { pos3[0] = e.abs_pos.x; pos3[1] = e.abs_pos.y; pos3[2] = e.abs_pos.z; }
  if (scale3) { scale3[0] = e.abs_scale.x; scale3[1] = e.abs_scale.y; scale3[2] = e.abs_scale.z; }
  if (orientation3) { orientation3[0] = e.abs_orientation.x; orientation3[1] = e.abs_orientation.y; orientation3[2] = e.abs_orientation.z; }
  if (forward3) { forward3[0] = e.forward.x; forward3[1] = e.forward.y; forward3[2] = e.forward.z; }
  return SNKPT_OK;
}

uint32_t snkscene_register_mesh(SnkScene* sc, const SnkMeshEntry* entry, const SnkSubmesh* submeshes, uint32_t num_submeshes) {
  if (!sc || !entry || !submeshes || num_submeshes == 0 || num_submeshes > SNKPT_MAX_SUBMESHES) return SNKSCENE_NONE;
  MeshRec m;
  m.entry = *entry;
  m.subs.assign(submeshes, submeshes + num_submeshes);
  sc->meshes.push_back(m);
  return (uint32_t)sc->meshes.size() - 1;
}

int snkscene_set_material(SnkScene* sc, uint32_t slot, const SnkMaterial* material) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(material && slot < 4096, "snkscene_set_material: slot");
  if (sc->materials.size() <= slot) {
    SnkMaterial blank;
    memset(&blank, 0, sizeof(blank));
    blank.albedo_tex_idx = blank.normal_tex_idx = blank.roughness_tex_idx = blank.metallic_tex_idx = blank.ao_tex_idx = SNKPT_INVALID_INDEX;
    sc->materials.resize(slot + 1, blank);
  }
  sc->materials[slot] = *material;
  return SNKPT_OK;
}

int snkscene_set_static_mesh(SnkScene* sc, uint32_t ent, uint32_t mesh, const uint32_t* material_slots, uint32_t num_materials) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(valid_ent(sc, ent) && mesh < sc->meshes.size() && (material_slots || num_materials == 0), "snkscene_set_static_mesh: arguments");
  const MeshRec& m = sc->meshes[mesh];
  for (const SnkSubmesh& s : m.subs) CHECK(s.material_index < num_materials, "snkscene_set_static_mesh: a submesh names a material the component does not have");
  CHECK((size_t)sc->next_instance_slot + m.subs.size() <= 4096, "snkscene_set_static_mesh: instance buffer full (4096 slots, never reclaimed)");
  Entity& e = sc->ents[ent];
  if (!e.has_mesh) sc->mesh_order.push_back(ent);
  e.has_mesh = true;
  e.mesh = mesh;
  e.material_slots.assign(material_slots, material_slots + num_materials);
  // add or update: always a fresh slot range (RaytracingBufferSystem.cpp:15-23)
  e.rt_instance_idx = sc->next_instance_slot;
  sc->next_instance_slot += (uint32_t)m.subs.size();
  return SNKPT_OK;
}

int snkscene_remove_static_mesh(SnkScene* sc, uint32_t ent) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(valid_ent(sc, ent), "snkscene_remove_static_mesh: unknown entity");
  Entity& e = sc->ents[ent];
  if (e.has_mesh)
    for (size_t k = 0; k < sc->mesh_order.size(); ++k) if (sc->mesh_order[k] == ent) { sc->mesh_order.erase(sc->mesh_order.begin() + k); break; }
  e.has_mesh = false;
  return SNKPT_OK;
}

int snkscene_set_directional_light(SnkScene* sc, const float* colour3, float theta, float phi) {
  if (!sc || !colour3) return SNKPT_E_INVALID;
  sc->dir_colour = V3{colour3[0], colour3[1], colour3[2]};
  sc->dir_theta = theta; sc->dir_phi = phi;
  return SNKPT_OK;
}

int snkscene_set_pointlight(SnkScene* sc, uint32_t ent, const float* colour3, float atten_exp, float atten_linear, float atten_constant) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(valid_ent(sc, ent) && colour3, "snkscene_set_pointlight: arguments");
  Entity& e = sc->ents[ent];
  if (!e.has_pointlight) {
    CHECK(sc->pointlight_order.size() < SNKPT_MAX_POINTLIGHTS, "snkscene_set_pointlight: more than 1024 point lights");
    sc->pointlight_order.push_back(ent);
  }
  e.has_pointlight = true;
  e.pl_colour = V3{colour3[0], colour3[1], colour3[2]};
  e.pl_exp = atten_exp; e.pl_linear = atten_linear; e.pl_constant = atten_constant;
  return SNKPT_OK;
}

int snkscene_remove_pointlight(SnkScene* sc, uint32_t ent) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(valid_ent(sc, ent), "snkscene_remove_pointlight: unknown entity");
  if (sc->ents[ent].has_pointlight)
    for (size_t k = 0; k < sc->pointlight_order.size(); ++k)
      if (sc->pointlight_order[k] == ent) { sc->pointlight_order.erase(sc->pointlight_order.begin() + k); break; }
  sc->ents[ent].has_pointlight = false;
  return SNKPT_OK;
}

int snkscene_set_camera(SnkScene* sc, uint32_t ent, float fov, float aspect_ratio, float z_near, float z_far) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(valid_ent(sc, ent) && fov > 0.f && aspect_ratio > 0.f && z_far > z_near && z_near > 0.f, "snkscene_set_camera: arguments");
  sc->camera = ent; sc->fov = fov; sc->aspect = aspect_ratio; sc->z_near = z_near; sc->z_far = z_far;
  return SNKPT_OK;
}

int snkscene_invalidate_history(SnkScene* sc) {
  if (!sc) return SNKPT_E_INVALID;
  sc->prev_frame_invalid = true;
  return SNKPT_OK;
}

int snkscene_frame_start(SnkScene* sc, float time_elapsed_ms, float delta_time_ms) {
  if (!sc) return SNKPT_E_INVALID;
  CHECK(sc->camera != SNKSCENE_NONE, "snkscene_frame_start: no active camera (SceneInfoBufferSystem::UpdateUBO returns early)");
  // ---- TransformBufferSystem: prev <- what cur held last frame; cur <- the matrices now (new slots: prev = cur)
  const size_t n = sc->ents.size();
  const size_t had = sc->transforms.size() / 16;
  sc->prev_transforms = sc->transforms;
  sc->prev_transforms.resize(16 * n);
  sc->transforms.resize(16 * n);
  for (size_t i = 0; i < n; ++i) {
    memcpy(&sc->transforms[16 * i], sc->ents[i].transform.m, 64);
    if (i >= had) memcpy(&sc->prev_transforms[16 * i], sc->ents[i].transform.m, 64);
  }
  // ---- RaytracingInstanceBufferSystem::UpdateInstanceBuffer: one record per (entity, submesh) at the entity's slot range
  SnkInstanceData blank;
  memset(&blank, 0, sizeof(blank));
  sc->instances.resize(sc->next_instance_slot, blank);
  sc->blas_of_instance.assign(sc->next_instance_slot, SNKPT_INVALID_INDEX);  // orphaned slots are not in the TLAS
  for (uint32_t ent : sc->mesh_order) {
    const Entity& e = sc->ents[ent];
    const MeshRec& m = sc->meshes[e.mesh];
    for (size_t j = 0; j < m.subs.size(); ++j) {
      SnkInstanceData d;
      d.transform_idx = ent;  // TransformBufferIdxComponent: one slot per entity, creation order
      d.material_idx = e.material_slots[m.subs[j].material_index];
      d.mesh_buffer_vertex_offset = m.entry.data_start_vertex_idx + m.subs[j].base_vertex;
      d.mesh_buffer_index_offset = m.entry.data_start_indices_idx + m.subs[j].base_index;
      d.num_mesh_indices = m.subs[j].num_indices;
      sc->instances[e.rt_instance_idx + j] = d;
      sc->blas_of_instance[e.rt_instance_idx + j] = m.entry.blas_id[j];
    }
  }
  // ---- UpdateEmissiveIdxBuffer: registry view order = reverse insertion
  sc->emissive.clear();
  for (size_t k = sc->mesh_order.size(); k-- > 0;) {
    const Entity& e = sc->ents[sc->mesh_order[k]];
    const MeshRec& m = sc->meshes[e.mesh];
    for (size_t j = 0; j < m.subs.size(); ++j) {
      const uint32_t slot = e.material_slots[m.subs[j].material_index];
      if (slot < sc->materials.size() && (sc->materials[slot].flags & SNKPT_MAT_FLAG_EMISSIVE)) sc->emissive.push_back(e.rt_instance_idx + (uint32_t)j);
    }
  }
  // ---- LightBufferSystem::UpdateLightSSBO
  const size_t npl = sc->pointlight_order.size();
  sc->lights.assign(112 + 48 * (npl ? npl : 1), 0);
  SnkDirectionalLight dl;
  memset(&dl, 0, sizeof(dl));
  dl.colour[0] = sc->dir_colour.x; dl.colour[1] = sc->dir_colour.y; dl.colour[2] = sc->dir_colour.z; dl.colour[3] = 1.f;
  {
    // ExtraMath::SphericalToCartesian(1, theta, phi) (ExtraMath.cpp:31-33), normalised
    const V3 d = normalize(V3{std::sin(sc->dir_phi) * std::cos(sc->dir_theta), std::sin(sc->dir_phi) * std::sin(sc->dir_theta), std::cos(sc->dir_phi)} * 1.f);
    dl.dir[0] = d.x; dl.dir[1] = d.y; dl.dir[2] = d.z; dl.dir[3] = 1.f;
    const M4 lt = mul(ortho_zo(-100.f, 100.f, -100.f, 100.f, 0.1f, 100.f), look_at(d * 50.f, V3{0, 0, 0}, V3{0, 1, 0}));
    memcpy(dl.light_transform, lt.m, 64);
  }
  memcpy(sc->lights.data(), &dl, 96);
  const uint32_t counts[2] = {(uint32_t)npl, 0u};
  memcpy(sc->lights.data() + 96, counts, 8);
  for (size_t k = 0; k < npl; ++k) {
    const Entity& e = sc->ents[sc->pointlight_order[npl - 1 - k]];
    SnkPointlight pl;
    pl.colour[0] = e.pl_colour.x; pl.colour[1] = e.pl_colour.y; pl.colour[2] = e.pl_colour.z; pl.colour[3] = 1.f;
    pl.pos[0] = e.abs_pos.x; pl.pos[1] = e.abs_pos.y; pl.pos[2] = e.abs_pos.z; pl.pos[3] = 1.f;
    pl.atten_exp = e.pl_exp; pl.atten_linear = e.pl_linear; pl.atten_constant = e.pl_constant; pl.padding = 1.f;
    memcpy(sc->lights.data() + 112 + 48 * k, &pl, 48);
  }
  // ---- SceneInfoBufferSystem: old UBO <- s_ubo, then UpdateUBO
  const Entity& cam = sc->ents[sc->camera];
  sc->prev_ubo = sc->ubo;
  SnkCommonUBO u;
  memset(&u, 0, sizeof(u));
  u.frame_idx = sc->frame_idx;
  const M4 view = look_at(cam.pos, cam.pos + cam.forward, V3{0.f, 1.f, 0.f});
  const M4 proj = perspective_vk(sc->fov, sc->aspect, sc->z_near, sc->z_far);
  const M4 pv = mul(proj, view);
  memcpy(u.view, view.m, 64); memcpy(u.proj, proj.m, 64); memcpy(u.proj_view, pv.m, 64);
  u.time_elapsed = time_elapsed_ms; u.delta_time = delta_time_ms;
  u.prev_frame_invalid = sc->prev_frame_invalid ? 1u : 0u;
  u.cam_pos[0] = cam.abs_pos.x; u.cam_pos[1] = cam.abs_pos.y; u.cam_pos[2] = cam.abs_pos.z; u.cam_pos[3] = 1.f;
  u.cam_forward[0] = cam.forward.x; u.cam_forward[1] = cam.forward.y; u.cam_forward[2] = cam.forward.z; u.cam_forward[3] = 1.f;
  sc->ubo = u;
  if (!sc->ubo_valid) sc->prev_ubo = u;  // first frame: the old UBO is the zero-initialised static (prev_frame_invalid covers it)
  sc->ubo_valid = true;
  return SNKPT_OK;
}

int snkscene_frame_end(SnkScene* sc) {
  if (!sc) return SNKPT_E_INVALID;
  sc->prev_frame_invalid = false;  // FrameEndEvent (SceneInfoBufferSystem.cpp:34-36)
  sc->frame_idx++;
  return SNKPT_OK;
}

const float* snkscene_transforms(const SnkScene* sc, int prev, uint32_t* count) {
  if (!sc) return nullptr;
  if (count) *count = (uint32_t)(sc->transforms.size() / 16);
  return prev ? sc->prev_transforms.data() : sc->transforms.data();
}
const SnkInstanceData* snkscene_instances(const SnkScene* sc, uint32_t* count) {
  if (!sc) return nullptr;
  if (count) *count = (uint32_t)sc->instances.size();
  return sc->instances.data();
}
const uint32_t* snkscene_blas_of_instance(const SnkScene* sc, uint32_t* count) {
  if (!sc) return nullptr;
  if (count) *count = (uint32_t)sc->blas_of_instance.size();
  return sc->blas_of_instance.data();
}
const uint32_t* snkscene_emissive(const SnkScene* sc, uint32_t* count) {
  if (!sc) return nullptr;
  if (count) *count = (uint32_t)sc->emissive.size();
  return sc->emissive.data();
}
const SnkMaterial* snkscene_materials(const SnkScene* sc, uint32_t* count) {
  if (!sc) return nullptr;
  if (count) *count = (uint32_t)sc->materials.size();
  return sc->materials.data();
}
const void* snkscene_lights(const SnkScene* sc, size_t* bytes) {
  if (!sc) return nullptr;
  if (bytes) *bytes = sc->lights.size();
  return sc->lights.data();
}
const SnkCommonUBO* snkscene_ubo(const SnkScene* sc, int prev) { return sc ? (prev ? &sc->prev_ubo : &sc->ubo) : nullptr; }

int snkscene_push(SnkScene* sc, SnkPt* pt, void* stream) {
  if (!sc || !pt) return SNKPT_E_INVALID;
  CHECK(sc->ubo_valid, "snkscene_push: call snkscene_frame_start first");
  int rc = SNKPT_OK;
  if (!sc->materials.empty()) rc = snkpt_set_materials(pt, sc->materials.data(), 0, (uint32_t)sc->materials.size());
  if (!rc) rc = snkpt_set_lights(pt, sc->lights.data(), sc->lights.size());
  if (!rc) rc = snkpt_set_transforms(pt, sc->transforms.data(), sc->prev_transforms.data(), 0, (uint32_t)(sc->transforms.size() / 16));
  if (!rc) rc = snkpt_set_instances(pt, sc->instances.data(), sc->blas_of_instance.data(), (uint32_t)sc->instances.size(), sc->emissive.data(),
                                    (uint32_t)sc->emissive.size());
  if (!rc) rc = snkpt_build_tlas(pt, stream);
  if (!rc) rc = snkpt_set_frame(pt, &sc->ubo, &sc->prev_ubo);
  if (rc) sc->err = std::string("snkscene_push: ") + snkpt_last_error(pt);
  return rc;
}

}  // extern "C"
