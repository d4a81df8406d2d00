// oracle/ref_glm.cpp — TEST INFRASTRUCTURE. Compiled against the reference's
// vendored glm (Core/extern/glm, include path only — no reference source is
// copied) to reproduce the host-side matrix producers exactly as the engine
// computes them:
//   CameraComponent::GetProjectionMatrix  (Core/headers/components/CameraComponent.h:15-19)
//   SceneInfoBufferSystem::UpdateUBO      (Core/src/scene/SceneInfoBufferSystem.cpp:78-81)
//   TransformSystem::RebuildMatrix        (Core/src/scene/TransformSystem.cpp:59-77) via ExtraMath::Rotate3D (ExtraMath.cpp:150-153)
//   BLAS::GenerateInstance                (Core/src/resources/TLAS.cpp:196-197)
//   ExtraMath::SphericalToCartesian       (ExtraMath.cpp:31-33)
#define GLM_FORCE_DEPTH_ZERO_TO_ONE  // Core/headers/pch/pch.h:3
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <glm/gtc/quaternion.hpp>
#include <cstring>

extern "C" {
__attribute__((visibility("default"))) void ref_projection(float fov, float aspect, float z_near, float z_far, float* out16) {
  auto proj = glm::perspective(glm::radians(fov * 0.5f), aspect, z_near, z_far);
  proj[1][1] *= -1;
  memcpy(out16, &proj, 64);
}
__attribute__((visibility("default"))) void ref_view(const float* pos, const float* forward, float* out16) {
  glm::vec3 p(pos[0], pos[1], pos[2]), f(forward[0], forward[1], forward[2]);
  auto v = glm::lookAt(p, p + f, glm::vec3(0.f, 1.f, 0.f));
  memcpy(out16, &v, 64);
}
__attribute__((visibility("default"))) void ref_mul(const float* a, const float* b, float* out16) {
  glm::mat4 A, B; memcpy(&A, a, 64); memcpy(&B, b, 64);
  auto C = A * B; memcpy(out16, &C, 64);
}
__attribute__((visibility("default"))) void ref_trs(const float* pos, const float* euler_deg, const float* scale, float* out16) {
  glm::quat q(glm::radians(glm::vec3(euler_deg[0], euler_deg[1], euler_deg[2])));
  glm::mat4 t = glm::mat3_cast(q);
  for (int r = 0; r < 3; ++r) { t[0][r] *= scale[0]; t[1][r] *= scale[1]; t[2][r] *= scale[2]; }
  t[3][0] = pos[0]; t[3][1] = pos[1]; t[3][2] = pos[2]; t[3][3] = 1.0f;
  memcpy(out16, &t, 64);
}
__attribute__((visibility("default"))) void ref_instance_transform(const float* mat16, float* out12) {
  glm::mat4 m; memcpy(&m, mat16, 64);
  auto t = glm::transpose(m);
  memcpy(out12, &t, sizeof(float) * 12);
}
__attribute__((visibility("default"))) void ref_inverse(const float* mat16, float* out16) {
  glm::mat4 m; memcpy(&m, mat16, 64);
  auto i = glm::inverse(m); memcpy(out16, &i, 64);
}
__attribute__((visibility("default"))) void ref_normal_matrix_det(const float* mat16, float* out9, float* det) {
  glm::mat4 m; memcpy(&m, mat16, 64);
  glm::mat3 nm = glm::transpose(glm::inverse(glm::mat3(m)));
  memcpy(out9, &nm, 36);
  *det = glm::determinant(glm::mat3(m));
}
__attribute__((visibility("default"))) void ref_spherical(float theta, float phi, float* out3) {
  glm::vec3 v = glm::normalize(1.f * glm::vec3{sinf(phi) * cosf(theta), sinf(phi) * sinf(theta), cosf(phi)});
  memcpy(out3, &v, 12);
}
}
