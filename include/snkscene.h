/* snkscene.h — host-side scene front-end for libsnkpt.
 *
 * A C++ restatement (C ABI here, implementation snake-vk_b200/host/snkscene.cpp) of the engine systems that FEED the
 * path-tracing pass: they turn an entity hierarchy into the six flat buffers the pass consumes. Nothing here touches the GPU;
 * snkscene_push hands the buffers to a SnkPt context through include/snkpt.h.
 *
 *   TransformSystem::RebuildMatrix / UpdateAbsTransforms      Core/src/scene/TransformSystem.cpp:31-114
 *   TransformBufferSystem::UpdateTransformBuffer              Core/headers/scene/TransformBufferSystem.h:92-122
 *   RaytracingInstanceBufferSystem::UpdateInstanceBuffer /
 *     UpdateEmissiveIdxBuffer                                 Core/src/scene/RaytracingBufferSystem.cpp:10-130
 *   LightBufferSystem::UpdateLightSSBO                        Core/src/scene/LightBufferSystem.cpp:65-108
 *   SceneInfoBufferSystem::UpdateUBO                          Core/src/scene/SceneInfoBufferSystem.cpp:25-33, :66-91
 *   CameraComponent::GetProjectionMatrix                      Core/headers/components/CameraComponent.h:15-19
 *
 * Semantics kept from the reference (and where they are simplified):
 *   - entity / transform-buffer / instance-buffer slots are handed out by bump counters and never reclaimed; updating a
 *     StaticMeshComponent moves the entity to NEW instance slots (RaytracingBufferSystem.cpp:15-23), the old ones stay in the
 *     buffer but are no longer referenced by the TLAS (blas_of_instance = SNKPT_INVALID_INDEX);
 *   - a child's matrix = parent matrix * (R * S with the parent's absolute scale divided out again, TransformSystem.cpp:78-106);
 *     absolute orientation / scale are the sum / product over the ancestors (:31-48), `absolute` entities ignore the parent;
 *   - the previous-frame transform buffer: the reference copies a transform into the "old" buffer once it is at least one frame
 *     old, per frame in flight (TransformBufferSystem.h:92-122); what the shaders of frame F see is prev = the value cur had in
 *     frame F-1, which is what frame_start produces. DEVIATION: a new transform slot starts with prev = cur (the reference
 *     leaves its old buffer uninitialised for the first two frames);
 *   - registry views iterate in reverse insertion order (entt's default pools): emissive list and point lights are emitted in
 *     that order;
 *   - view = lookAt(local position, local position + forward, +Y) but cam_pos = ABSOLUTE position (SceneInfoBufferSystem.cpp:
 *     78, :87) — identical for root cameras.
 */
#ifndef SNKSCENE_H
#define SNKSCENE_H
#include <stddef.h>
#include <stdint.h>

#include "snkpt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SNKSCENE_API __attribute__((visibility("default")))
#define SNKSCENE_NONE 0xFFFFFFFFu

typedef struct SnkScene SnkScene;

SNKSCENE_API SnkScene* snkscene_create(void);
SNKSCENE_API void snkscene_destroy(SnkScene* sc);
SNKSCENE_API const char* snkscene_last_error(const SnkScene* sc);

/* Scene::CreateEntity: entity + TransformComponent + RelationshipComponent; returns the entity id (= its transform-buffer slot) */
SNKSCENE_API uint32_t snkscene_create_entity(SnkScene* sc, uint32_t parent);
SNKSCENE_API int snkscene_set_parent(SnkScene* sc, uint32_t ent, uint32_t parent);
/* TransformComponent::SetPosition / SetOrientation (Euler angles in degrees) / SetScale; NULL keeps the current value */
SNKSCENE_API int snkscene_set_transform(SnkScene* sc, uint32_t ent, const float* pos3, const float* euler_deg3, const float* scale3);
SNKSCENE_API int snkscene_set_absolute(SnkScene* sc, uint32_t ent, int absolute);
SNKSCENE_API int snkscene_get_matrix(const SnkScene* sc, uint32_t ent, float* out16);
SNKSCENE_API int snkscene_get_abs(const SnkScene* sc, uint32_t ent, float* pos3, float* scale3, float* orientation3, float* forward3);

/* MeshBufferManager entry + submesh table of an uploaded mesh (what snkpt_upload_mesh returned); returns the mesh id */
SNKSCENE_API uint32_t snkscene_register_mesh(SnkScene* sc, const SnkMeshEntry* entry, const SnkSubmesh* submeshes, uint32_t num_submeshes);
/* GlobalMaterialBufferManager slot */
SNKSCENE_API int snkscene_set_material(SnkScene* sc, uint32_t slot, const SnkMaterial* material);
/* StaticMeshComponent add / update: material_slots[k] = global slot of the mesh's material index k */
SNKSCENE_API int snkscene_set_static_mesh(SnkScene* sc, uint32_t ent, uint32_t mesh, const uint32_t* material_slots, uint32_t num_materials);
SNKSCENE_API int snkscene_remove_static_mesh(SnkScene* sc, uint32_t ent);

SNKSCENE_API int snkscene_set_directional_light(SnkScene* sc, const float* colour3, float theta, float phi);
SNKSCENE_API int snkscene_set_pointlight(SnkScene* sc, uint32_t ent, const float* colour3, float atten_exp, float atten_linear, float atten_constant);
SNKSCENE_API int snkscene_remove_pointlight(SnkScene* sc, uint32_t ent);
/* CameraComponent on `ent`, made the active camera */
SNKSCENE_API int snkscene_set_camera(SnkScene* sc, uint32_t ent, float fov, float aspect_ratio, float z_near, float z_far);

/* FrameStartEvent: transform / instance / emissive / light buffers refreshed, old UBO <- UBO, UpdateUBO.
 * FrameEndEvent: prev_frame_invalid cleared, frame index advanced. */
SNKSCENE_API int snkscene_frame_start(SnkScene* sc, float time_elapsed_ms, float delta_time_ms);
SNKSCENE_API int snkscene_frame_end(SnkScene* sc);
SNKSCENE_API int snkscene_invalidate_history(SnkScene* sc); /* resize / scene load: prev_frame_invalid for the next frame */

/* the flat buffers as of the last frame_start (pointers stay valid until the next call that modifies the scene) */
SNKSCENE_API const float* snkscene_transforms(const SnkScene* sc, int prev, uint32_t* count);
SNKSCENE_API const SnkInstanceData* snkscene_instances(const SnkScene* sc, uint32_t* count);
SNKSCENE_API const uint32_t* snkscene_blas_of_instance(const SnkScene* sc, uint32_t* count);
SNKSCENE_API const uint32_t* snkscene_emissive(const SnkScene* sc, uint32_t* count);
SNKSCENE_API const SnkMaterial* snkscene_materials(const SnkScene* sc, uint32_t* count);
SNKSCENE_API const void* snkscene_lights(const SnkScene* sc, size_t* bytes);
SNKSCENE_API const SnkCommonUBO* snkscene_ubo(const SnkScene* sc, int prev);

/* the upload calls of one frame on a path-tracing context: materials, lights, transforms, instances, TLAS, frame UBOs */
SNKSCENE_API int snkscene_push(SnkScene* sc, SnkPt* pt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNKSCENE_H */
