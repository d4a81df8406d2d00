/* snkpt.h — C ABI of the B200-native path-tracing pass.
 *
 * Drop-in boundary for Snake-VK's pathtracing pass, i.e. `class RT`
 * (reference: Core/headers/rendering/Raytracing.h:11-31) plus the buffer
 * "systems" that feed it. The reference is a C++ class that pulls its inputs
 * out of engine singletons; this header states the same data contract as plain
 * pointers + sizes so that a binding (C++, ctypes, cgo ...) is a few lines.
 * INTEGRATION.md shows the engine-side shim.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SNKPT_E_* otherwise and
 *     never aborts (the reference's SNK_ASSERT/__debugbreak convention,
 *     Core/headers/util/util.h:7-31, is replaced by status codes);
 *     snkpt_last_error() gives the text of the last failure.
 *   - one context per GPU; a context is not thread-safe (the reference records
 *     the pass from a single render job, Core/src/core/App.cpp:98-112).
 *   - host pointers are copied before the call returns; the caller keeps
 *     ownership. `stream` arguments are a `cudaStream_t` passed as void*
 *     (NULL = the context's own stream).
 *   - matrices are column-major float[16] (glm layout); record layouts are the
 *     reference's wire formats and are static_assert'ed in csrc/layouts.h.
 *   - there is no CPU fallback: every entry point that computes needs the CUDA
 *     device named in SnkPtConfig and fails with SNKPT_E_CUDA without it.
 */
#ifndef SNKPT_H
#define SNKPT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNKPT_API __attribute__((visibility("default")))

/* ---- status codes ------------------------------------------------------- */
enum {
  SNKPT_OK = 0,
  SNKPT_E_INVALID = -1, /* bad argument / out of capacity                  */
  SNKPT_E_CUDA = -2,    /* CUDA runtime error (text in snkpt_last_error)   */
  SNKPT_E_STATE = -3,   /* call order violated (e.g. render before TLAS)   */
  SNKPT_E_NOMEM = -4
};

/* ---- wire formats --------------------------------------------------------- */

/* Core/headers/scene/RaytracingBufferSystem.h:14-28, RaytracingInstances.glsl:7-13 */
typedef struct SnkInstanceData {
  uint32_t transform_idx;
  uint32_t mesh_buffer_index_offset; /* in indices; the shaders divide by 3 */
  uint32_t mesh_buffer_vertex_offset;
  uint32_t material_idx;
  uint32_t num_mesh_indices;
} SnkInstanceData; /* 20 B */

/* TexMatBuffers.glsl:5-20, writer Core/src/assets/ResourceBufferManagers.cpp:64-101 */
#define SNKPT_INVALID_INDEX 0xFFFFFFFFu
#define SNKPT_MAT_FLAG_EMISSIVE 1u
typedef struct SnkMaterial {
  uint32_t albedo_tex_idx, normal_tex_idx, roughness_tex_idx, metallic_tex_idx, ao_tex_idx;
  float roughness, metallic, ao;
  float albedo[4];
  uint32_t flags;
  uint32_t pad0, pad1, pad2;
} SnkMaterial; /* 64 B */

/* LightBuffers.glsl:1-49 (std430), writer Core/src/scene/LightBufferSystem.cpp:65-108 */
typedef struct SnkDirectionalLight {
  float colour[4];
  float dir[4]; /* unit vector TOWARDS the light (ExtraMath.cpp:31-33) */
  float light_transform[16];
} SnkDirectionalLight; /* 96 B */
typedef struct SnkPointlight {
  float colour[4];
  float pos[4];
  float atten_exp, atten_linear, atten_constant, padding;
} SnkPointlight; /* 48 B */
typedef struct SnkSpotlight {
  float colour[4];
  float pos[4];
  float dir[4];
  float atten_exp, atten_linear, atten_constant, aperture;
} SnkSpotlight; /* 64 B */
#define SNKPT_MAX_POINTLIGHTS 1024
#define SNKPT_MAX_SPOTLIGHTS 1024
typedef struct SnkLightData {
  SnkDirectionalLight dir_light; /* @0   */
  uint32_t num_pointlights;      /* @96  */
  uint32_t num_spotlights;       /* @100 */
  float pad0, pad1;
  SnkPointlight pointlights[SNKPT_MAX_POINTLIGHTS]; /* @112   */
  SnkSpotlight spotlights[SNKPT_MAX_SPOTLIGHTS];    /* @49264 */
} SnkLightData;

/* CommonUBO.glsl:1-27 (std140, 240 B), writer Core/src/scene/SceneInfoBufferSystem.cpp:66-91 */
typedef struct SnkCommonUBO {
  float view[16];
  float proj[16];
  float proj_view[16];
  float cam_pos[4];
  float cam_forward[4];
  uint32_t frame_idx;
  float time_elapsed; /* milliseconds; doubles as an RNG seed in reference mode */
  float delta_time;   /* milliseconds                                           */
  uint32_t prev_frame_invalid;
} SnkCommonUBO; /* 240 B */

/* RayGen.rgen:53-76 (std430, 48 B). AoS view used only at the boundary; the
 * device keeps reservoirs as SoA planes in HBM. */
typedef struct SnkPixelReservoir {
  float w_sum_pl;
  float w_pl;
  uint32_t pl_idx;
  uint32_t pl_m;
  uint32_t object_instance_idx;
  uint32_t triangle_idx;
  float bary[2];
  float w_sum_em;
  float w_em;
  uint32_t m_em;
  uint32_t pad;
} SnkPixelReservoir; /* 48 B */

/* submesh record of MeshData (Core/headers/assets/MeshData.h) */
typedef struct SnkSubmesh {
  uint32_t num_indices;
  uint32_t num_vertices;
  uint32_t base_vertex;
  uint32_t base_index;
  uint32_t material_index;
} SnkSubmesh;

#define SNKPT_MAX_SUBMESHES 64
/* MeshBufferManager::MeshEntryData (Core/src/rendering/MeshBufferManager.cpp:83-87)
 * + the BLAS handles the reference keeps in MeshDataAsset (MeshData.h:57) */
typedef struct SnkMeshEntry {
  uint32_t data_start_indices_idx;
  uint32_t data_start_vertex_idx;
  uint32_t num_submeshes;
  uint32_t blas_id[SNKPT_MAX_SUBMESHES];
} SnkMeshEntry;

/* traceRayEXT as the shaders use it (RayGen.rgen:101-121, :164-176): 32 B ray
 * in, hit out. Directions are NOT normalised by the API. */
typedef struct SnkRay {
  float o[3];
  float tmin;
  float d[3];
  float tmax;
} SnkRay; /* 32 B */
typedef struct SnkHit {
  float t;       /* < 0 → miss (RayMiss.rmiss:24 sets distance = -1)          */
  float u, v;    /* barycentric weights of v1, v2 (RayClosestHit.rchit:88)    */
  uint32_t prim; /* gl_PrimitiveID inside the BLAS                             */
  uint32_t inst; /* gl_InstanceCustomIndexEXT (TLAS.cpp:193)                   */
} SnkHit; /* 20 B */
enum {
  SNKPT_RAY_CLOSEST = 0u,
  SNKPT_RAY_ANY = 1u,          /* TerminateOnFirstHit|Opaque|SkipClosestHit    */
  SNKPT_RAY_DEVICE_PTRS = 2u   /* rays / hits already live in device memory    */
};

/* ---- configuration -------------------------------------------------------- */
enum { SNKPT_MODE_REFERENCE = 0, SNKPT_MODE_NORTHSTAR = 1 };

typedef struct SnkPtConfig {
  uint32_t width, height;
  int32_t device;          /* CUDA ordinal                                       */
  uint32_t max_instances;  /* 0 → 4096 (RaytracingBufferSystem.cpp:60)            */
  uint32_t max_transforms; /* 0 → 4096 (TransformBufferSystem.h:70)               */
  uint32_t max_materials;  /* 0 → 4096 (ResourceBufferManagers.cpp:10)            */
  uint32_t row_begin;      /* rows [row_begin,row_end) are rendered by this       */
  uint32_t row_end;        /* context; 0,0 → all rows (multi-GPU band sharding)   */
  uint32_t reserved[4];
} SnkPtConfig;

/* RT::RtSettings (Raytracing.h:13-16) + the constants the shader hard-codes
 * (RayGen.rgen:95-97, :435, :488, :491). snkpt_default_settings() fills the
 * reference defaults. */
typedef struct SnkRtSettings {
  int32_t num_temporal_resamples; /* 1 */
  int32_t num_spatial_resamples;  /* 3 */
  uint32_t mode;                  /* SNKPT_MODE_*                 */
  uint32_t ris_iters;             /* 32  (LIGHT_RIS_ITERS)         */
  uint32_t m_cap;                 /* 20  (M_CAP)                   */
  float neighbour_radius;         /* 30 px                         */
  float normal_threshold;         /* 0.95                          */
  uint32_t max_bounces;           /* northstar: 4                  */
  uint32_t spp;                   /* northstar: 1                  */
  uint32_t flags;                 /* SNKPT_RT_*                    */
  uint32_t reserved[6];
} SnkRtSettings;
enum {
  SNKPT_RT_NO_RESTIR = 1u, /* skip both reservoir passes                         */
  SNKPT_RT_NO_BOUNCE = 2u, /* skip BounceDiffuse / path continuation             */
  SNKPT_RT_UNBIASED_REUSE = 4u /* north-star mode: the MIS weights of spatial reuse carry the visibility of the selected sample
                                * from every contributing pixel's surface (<= 2 shadow rays per accepted neighbour): removes the
                                * darkening bias of reuse across shadow boundaries (Bitterli et al. 2020, Alg. 6)            */
};

/* what snkpt_read_aux / snkpt_device_ptr can name */
enum {
  SNKPT_AUX_PRIM_ID = 0,         /* u32/pixel, 0xFFFFFFFF = sky                  */
  SNKPT_AUX_INST_ID = 1,         /* u32/pixel                                    */
  SNKPT_AUX_DEPTH = 2,           /* f32/pixel  (D32F, cleared to 1)              */
  SNKPT_AUX_NORMAL = 3,          /* 4×f16/pixel                                  */
  SNKPT_AUX_ALBEDO = 4,          /* 4×snorm8/pixel                               */
  SNKPT_AUX_RMA = 5,             /* 4×f16/pixel                                  */
  SNKPT_AUX_MOTION = 6,          /* 2×f16/pixel                                  */
  SNKPT_AUX_MAT_FLAGS = 7,       /* u16/pixel                                    */
  SNKPT_AUX_PREV_NORMAL = 8,     /* 4×f16/pixel                                  */
  SNKPT_AUX_PREV_DEPTH = 9,      /* f32/pixel                                    */
  SNKPT_AUX_RESERVOIRS_FINAL = 10,    /* SnkPixelReservoir/pixel — buffer half 0 */
  SNKPT_AUX_RESERVOIRS_TEMPORAL = 11, /* SnkPixelReservoir/pixel — buffer half 1 */
  SNKPT_AUX_BVH_NODES = 12,      /* 80-B wide nodes, all BLASes then the TLAS    */
  SNKPT_AUX_BVH_TRIS = 13,       /* 48-B triangle records                        */
  SNKPT_AUX_HIT_T = 14,          /* f32/pixel primary-hit distance, -1 = sky     */
  SNKPT_AUX_HIT_BARY = 15,       /* 2×f32/pixel                                  */
  SNKPT_AUX_RADIANCE_F32 = 16,   /* 4×f32/pixel un-quantised frame output        */
  SNKPT_AUX_ACCUM_F32 = 17,      /* 4×f32/pixel progressive mean                 */
  SNKPT_AUX_DIRECT_F32 = 18,     /* 4×f32/pixel direct part (dir light + ReSTIR) */
  SNKPT_AUX_TLAS_INSTANCES = 19, /* 80-B device instance records                 */
  SNKPT_AUX_BLAS_TABLE = 20,     /* SnkBlasInfo records                          */
  SNKPT_AUX_TAA_HISTORY = 21,    /* 4×f16/pixel TAA history image                */
  SNKPT_AUX_TAA_PREV_VELOCITY = 22 /* 2×f16/pixel velocity of the previous resolve */
};

enum { SNKPT_FMT_RGBA16F = 0, SNKPT_FMT_RGBA32F = 1 };

/* counters of the last snkpt_render / snkpt_trace_primary / snkpt_trace_rays */
typedef struct SnkPtStats {
  uint64_t rays_primary;
  uint64_t rays_shadow;      /* any-hit queries                                  */
  uint64_t rays_secondary;   /* closest-hit queries after the primary hit        */
  uint64_t kernel_launches;  /* launches of this library's kernels since create  */
  uint64_t blas_nodes, blas_tris, tlas_nodes;
  float ms_primary, ms_restir1, ms_restir2, ms_bounce, ms_total; /* CUDA events of last frame */
  float ms_build_last;       /* last BLAS / TLAS build                            */
  uint32_t stack_overflows;  /* traversal-stack overflows since create (must be 0) */
  uint64_t total_rays_primary, total_rays_shadow, total_rays_secondary; /* since create, folded on the device per frame */
} SnkPtStats; /* 112 B */

/* one kernel launch of the last frame (snkpt_trace_primary .. snkpt_render) or of
 * the last snkpt_trace_rays, timed with CUDA events on the launching stream */
typedef struct SnkLaunchTiming {
  char name[32];      /* kernel name, e.g. "trace_closest.B"                       */
  float ms;           /* event-to-event duration of this launch                    */
  uint32_t is_trace;  /* 1 = a traversal launch (units are rays)                   */
  uint64_t units;     /* rays or pixels handed to the launch (incl. dead slots)    */
  uint64_t live;      /* traversal launches at profiling level 2: rays with tmax >= tmin */
  uint64_t nodes, tris, insts; /* level 2: SURVEY §8(d) algorithmic work of the launch */
  uint64_t alg_bytes; /* level 2: 32·live + 16|4·live + 80·nodes + 48·tris + 64·insts  */
} SnkLaunchTiming; /* 88 B */

typedef struct SnkBlasInfo {
  uint32_t node_root;   /* index of the root wide node in the node pool          */
  uint32_t node_count;
  uint32_t tri_base;    /* first 48-B triangle record                            */
  uint32_t tri_count;
  float bounds_lo[3], bounds_hi[3];
} SnkBlasInfo;

typedef struct SnkPt SnkPt;

/* ---- lifetime --------------------------------------------------------------- */
/* RT::Init (Raytracing.cpp:269-278) + EditorLayer::InitGBuffer / InitRenderResources
 * (Editor/src/EditorLayer.cpp:227-303, 351-411): allocates output image,
 * G-buffer planes and the 2·W·H reservoir planes; history starts invalid. */
SNKPT_API int snkpt_create(const SnkPtConfig* cfg, SnkPt** out);
SNKPT_API void snkpt_destroy(SnkPt* ctx);
/* ctx may be NULL: returns the last error of a failed snkpt_create */
SNKPT_API const char* snkpt_last_error(const SnkPt* ctx);
SNKPT_API void snkpt_default_settings(SnkRtSettings* s);
/* EditorLayer::OnFrameStart resize path (EditorLayer.cpp:415-423): re-allocates
 * per-pixel state and raises prev_frame_invalid for the next frame */
SNKPT_API int snkpt_resize(SnkPt* ctx, uint32_t width, uint32_t height);
SNKPT_API int snkpt_set_rows(SnkPt* ctx, uint32_t row_begin, uint32_t row_end);
/* Row-band sharding (one band per GPU). The band [row_begin,row_end) is what pass 2, the bounce
 * stages and the output cover. Reservoir pass 1 additionally covers `pass1_halo` rows above and below the band
 * (spatial reuse reads temporal reservoirs up to neighbour_radius rows away) and the G-buffer `gbuffer_halo`
 * rows (pass 1 on the halo reads prev-frame normals / history one motion vector further). With
 * pass1_halo >= radius, gbuffer_halo >= pass1_halo + max motion + 1 and the FINAL reservoirs (buffer half 0)
 * of rows [band ± (pass1_halo + max motion)] exchanged between band neighbours after every frame, a band
 * renders bit-identically to the same rows of a whole-frame context. Defaults 0, 0 (whole frame). */
SNKPT_API int snkpt_set_halo(SnkPt* ctx, uint32_t pass1_halo, uint32_t gbuffer_halo);
/* Row-band GROUPS: the frame strong-scaled over the GPUs of one node (BASELINE config 4; no counterpart in the reference,
 * whose frame runs on one GPU — SURVEY §8e). Band `rank` of `world` equal bands (snkpt_group_band) is rendered by one
 * context created with those rows; instead of re-rendering halo rows the library exchanges reservoir rows between band
 * neighbours inside snkpt_render: the temporal reservoirs of `radius_rows` rows per edge after the temporal stage (spatial
 * reuse, RayGen.rgen:470-506) and the final reservoirs of `margin_rows` rows per edge after the spatial stage (next
 * frame's temporal reuse, :424-461). Only the G-buffer keeps a halo of max(radius, margin) + 1 rows. North-star mode.
 * A band then equals the same rows of a whole-frame context bit for bit while per-frame motion stays below margin_rows.
 *   snkpt_group_attach_nccl   one process per GPU: ncclSend / ncclRecv on the frame's stream (libnccl.so.2 resolved at run
 *                             time, SNKPT_NCCL_LIB overrides); rank 0 makes the id (snkpt_group_unique_id), the host
 *                             broadcasts the 128 bytes with whatever it has (torch.distributed in bench.py)
 *   snkpt_group_attach_local  the contexts of ONE process (several GPUs, or one GPU in tests), ordered by rank: direct
 *                             device-to-device copies; each context must then be driven by its own host thread
 *   snkpt_group_set_bands     optional, before attaching: bands of unequal height (row_starts[0] = 0 < ... < row_starts[world]
 *                             = height, the same table on every context of the group) — rows cost what they show (sky against
 *                             geometry), so equal bands leave GPUs idle at the two exchanges; bench.py balances them from the
 *                             rays each band traced. world = 0 / NULL returns to equal bands. */
SNKPT_API void snkpt_group_band(uint32_t height, int world, int rank, uint32_t* row_begin, uint32_t* row_end);
SNKPT_API int snkpt_group_set_bands(SnkPt* ctx, int world, const uint32_t* row_starts);
SNKPT_API int snkpt_group_unique_id(void* out128);
SNKPT_API int snkpt_group_attach_nccl(SnkPt* ctx, int rank, int world, const void* id128, uint32_t radius_rows, uint32_t margin_rows);
SNKPT_API int snkpt_group_attach_local(SnkPt** ctxs, int n, uint32_t radius_rows, uint32_t margin_rows);
SNKPT_API int snkpt_group_detach(SnkPt* ctx);
SNKPT_API int snkpt_group_info(SnkPt* ctx, int* rank, int* world, uint64_t* bytes_sent);
/* invalidates temporal history (SceneInfoBufferSystem m_prev_frame_invalid) */
SNKPT_API int snkpt_reset_history(SnkPt* ctx);

/* ---- scene upload ------------------------------------------------------------- */
/* MeshBufferManager::LoadMeshFromData (MeshBufferManager.cpp:22-88) + one
 * BLAS::GenerateFromMeshData per submesh (TLAS.cpp:109-186). positions, normals,
 * tangents are tightly packed float3, uvs float2 (normals/tangents/uvs may be
 * NULL → zeros), indices are mesh-local u32. Builds one wide BVH per submesh. */
SNKPT_API int snkpt_upload_mesh(SnkPt* ctx, const float* positions, const float* normals,
                                const float* tangents, const float* uvs, uint32_t num_vertices,
                                const uint32_t* indices, uint32_t num_indices,
                                const SnkSubmesh* submeshes, uint32_t num_submeshes,
                                SnkMeshEntry* out_entry, void* stream);
/* overwrite vertex positions of an uploaded mesh (animated geometry), then
 * refit its BLASes without changing topology (no reference counterpart: the
 * reference never updates a BLAS, TLAS.cpp:63) */
SNKPT_API int snkpt_update_positions(SnkPt* ctx, const SnkMeshEntry* entry, const float* positions,
                                     uint32_t num_vertices);
SNKPT_API int snkpt_refit_blas(SnkPt* ctx, uint32_t blas_id, void* stream);
SNKPT_API int snkpt_rebuild_blas(SnkPt* ctx, uint32_t blas_id, void* stream);

/* GlobalMaterialBufferManager::UpdateMaterialUBO (ResourceBufferManagers.cpp:64-101) */
SNKPT_API int snkpt_set_materials(SnkPt* ctx, const SnkMaterial* materials, uint32_t first,
                                  uint32_t count);
/* LightBufferSystem::UpdateLightSSBO (LightBufferSystem.cpp:65-108); `bytes`
 * may stop after the last used point light */
/* Texture2DAsset upload + GlobalTextureBufferManager slot (Core/src/assets/ResourceBufferManagers.cpp:12-62; the
 * `textures[4096]` descriptor array of TexMatBuffers.glsl): RGBA8 texels of the base level, row-major, top row first.
 * srgb != 0 = VK_FORMAT_R8G8B8A8_SRGB as base-colour textures are created (AssetLoader.cpp:570). Sampled with linear
 * filter + repeat addressing (Images.cpp:18-22) through Material::albedo_tex_idx by the hit shading
 * (RayClosestHit.rchit:115-116) and the G-buffer (GBuffer.frag:57-58). width == 0 frees the slot. Normal / roughness /
 * metallic / ao maps are not sampled by the ray-tracing shaders and are not uploaded. */
SNKPT_API int snkpt_set_texture(SnkPt* ctx, uint32_t tex_idx, uint32_t width, uint32_t height, const void* rgba8, uint32_t srgb);
SNKPT_API int snkpt_set_lights(SnkPt* ctx, const void* light_data, size_t bytes);
/* TransformBufferSystem::UpdateTransformBuffer (TransformBufferSystem.h:92-122) */
SNKPT_API int snkpt_set_transforms(SnkPt* ctx, const float* cur_mat4, const float* prev_mat4,
                                   uint32_t first, uint32_t count);
/* RaytracingInstanceBufferSystem::UpdateInstanceBuffer / UpdateEmissiveIdxBuffer
 * (RaytracingBufferSystem.cpp:69-130) + BLAS::GenerateInstance (TLAS.cpp:188-200):
 * slot i of `instances` is the instanceCustomIndex; blas_of_instance[i] is the
 * BLAS it references (SNKPT_INVALID_INDEX = slot unused). */
SNKPT_API int snkpt_set_instances(SnkPt* ctx, const SnkInstanceData* instances,
                                  const uint32_t* blas_of_instance, uint32_t count,
                                  const uint32_t* emissive_idx, uint32_t num_emissive);
/* TlasSystem::BuildTlas (TlasSystem.cpp:27-49) → TLAS::BuildFromInstances
 * (TLAS.cpp:10-87): full rebuild from the current instances + transforms.
 * Stream semantics (also snkpt_trace_primary): work enqueued on `stream` AFTER this call sees the rebuilt TLAS, as with any
 * stream-ordered call. The rebuild itself may start EARLIER than the work enqueued before it: when nothing it reads was changed
 * since the previous rebuild, it runs on an internal stream, into a second copy of the TLAS storage, under the tail of the frame in
 * flight (the camera rays likewise, into a second copy of the G-buffer planes). Results do not depend on it; SNKPT_TLAS_EARLY=0 /
 * SNKPT_PRIMARY_EARLY=0 in the environment keep both strictly in order. */
SNKPT_API int snkpt_build_tlas(SnkPt* ctx, void* stream);

/* ---- per frame ------------------------------------------------------------------ */
/* SceneInfoBufferSystem::UpdateUBO (SceneInfoBufferSystem.cpp:29-33, 66-91) */
SNKPT_API int snkpt_set_frame(SnkPt* ctx, const SnkCommonUBO* cur, const SnkCommonUBO* prev);
/* Replaces GBufferPass::RecordCommandBuffer (GBufferPass.cpp:70-196): copies
 * normal/depth to the prev planes, then fills the G-buffer with primary rays.
 * jitter = sub-pixel offset in pixels (NULL = pixel centres). */
SNKPT_API int snkpt_trace_primary(SnkPt* ctx, const float* jitter_xy, void* stream);
/* GetFrameJitter (Core/headers/rendering/RenderCommon.h:39-60): Halton(2,3) - 0.5 over round(8 * (output_w / internal_w)^2)
 * phases, the push constant GBufferPass hands GBuffer.vert (GBufferPass.cpp:136). GBuffer.vert:90-95 moves the raster
 * position by (jitter * w - 0.5) / resolution in clip space = jitter / 2 pixels (minus a 0.25 / w pixel term that depends on
 * the vertex depth); a camera ray cannot carry a per-vertex term, so the equivalent ray offset handed to
 * snkpt_trace_primary is -jitter / 2 pixels (DEVIATION: the 0.25 / w term is dropped). out_pixels receives that offset. */
SNKPT_API void snkpt_frame_jitter(uint32_t frame_idx, uint32_t internal_render_width, uint32_t output_render_width,
                                  float* out_jitter_xy, float* out_pixels_xy);
/* RT::RecordRenderCmdBuf (Raytracing.cpp:227-267): reservoir pass 1, pass 2,
 * shading and bounce rays; writes the output image */
SNKPT_API int snkpt_render(SnkPt* ctx, const SnkRtSettings* settings, void* stream);
/* progressive mean of frame outputs in fp32 (no reference counterpart) */
SNKPT_API int snkpt_accumulate(SnkPt* ctx, int reset, void* stream);
/* TAA_ResolvePass::RecordCommandBuffer (Core/src/renderpasses/TAA_ResolvePass.cpp:39-56),
 * enabled by RenderSettings::using_taa (Editor/src/EditorLayer.cpp:636-640): runs
 * TAA_Resolve.comp (Core/res/shaders/TAA_Resolve.comp:1-85) on the RGBA16F output image
 * of the last snkpt_render in place, against the history image and the G-buffer
 * velocity plane, then does the pass's two blits (output -> history, velocity ->
 * previous velocity). The first call after create / resize / snkpt_reset_history
 * seeds the history with the current frame (the reference leaves it undefined).
 * Whole frames only: returns SNKPT_E_STATE on a context that renders a row band. */
SNKPT_API int snkpt_taa_resolve(SnkPt* ctx, void* stream);

/* ---- readback ---------------------------------------------------------------------- */
/* internal_render_image (EditorLayer.cpp:389-395): RGBA16F, linear, alpha 1.
 * Copies rows [row_begin,row_end) × width into dst (device→host, synchronous). */
SNKPT_API int snkpt_read_output(SnkPt* ctx, void* dst, uint32_t format);
/* The same copy without blocking: enqueued behind the frame on the context's copy stream, so it overlaps the next
 * frame (whose kernels wait for it before they overwrite the image). dst should be page-locked. slot = 0..3: up to four reads may
 * be in flight; snkpt_read_output_wait(slot) returns when that read has landed (the engine presents frame N while N+1 renders,
 * as the reference's frames in flight do, EditorLayer.cpp:563-665). */
SNKPT_API int snkpt_read_output_async(SnkPt* ctx, void* dst, uint32_t format, uint32_t slot);
SNKPT_API int snkpt_read_output_wait(SnkPt* ctx, uint32_t slot);
SNKPT_API int snkpt_read_aux(SnkPt* ctx, uint32_t what, void* dst, size_t bytes);
SNKPT_API int snkpt_aux_size(SnkPt* ctx, uint32_t what, size_t* bytes);
/* overwrite a reservoir half from an AoS host array (tests, halo import) */
SNKPT_API int snkpt_write_reservoirs(SnkPt* ctx, uint32_t half, const SnkPixelReservoir* src,
                                     uint32_t first_pixel, uint32_t count);
/* raw device pointer + byte size of a per-pixel plane set, for NCCL plumbing by
 * the host (fp16 output: what = 0x100, fp32 accum: 0x101, reservoir SoA planes
 * of half h: 0x110 + h) */
SNKPT_API int snkpt_device_ptr(SnkPt* ctx, uint32_t what, void** ptr, size_t* bytes);
SNKPT_API int snkpt_get_stats(SnkPt* ctx, SnkPtStats* out);
/* level 0: off. 1: bracket every kernel launch with CUDA events (cheap; what
 * bench.py runs with). 2: additionally run the work-counting diagnostic after
 * every traversal launch (timings of those frames are then not representative).
 * 3: bracket only the traversal launches, on the stream each one runs on, and keep
 * the north-star frame's two chains overlapped (levels 1 and 2 run one launch at a
 * time so that every kernel is timed alone) — what bench.py's timed region uses.
 * 4: bracket every launch AND keep the chains overlapped: a timeline — `nodes` of a record then holds the start of the launch
 * in ns after the start of the first record (durations of overlapped launches include the time they shared the GPU).
 * Records accumulate (up to 2048) until the next snkpt_set_profiling call or
 * snkpt_get_launch_timings, which returns and clears them. */
SNKPT_API int snkpt_set_profiling(SnkPt* ctx, uint32_t level);
/* synchronises the context stream; returns up to `cap` records of the last frame */
SNKPT_API int snkpt_get_launch_timings(SnkPt* ctx, SnkLaunchTiming* out, uint32_t cap, uint32_t* count);
SNKPT_API int snkpt_synchronize(SnkPt* ctx);

/* ---- particle collision pass --------------------------------------------
 * The reference's ParticleSystem (Core/src/scene/ParticleSystem.cpp:224-299) on the same acceleration structures:
 * per sub-step cell keys (Particle.comp:51-55), key sort, start indices (:77-85), ray-traced collision update
 * (ParticleUpdate.rgen / .rchit), copy back. */
typedef struct SnkParticle {      /* Particle.glsl:3-7 */
  float position_radius[4];
  float velocity[4];              /* w: 0 = asleep, 1 = awake */
  float colour[4];
} SnkParticle;                    /* 48 B */
typedef struct SnkPhysicsParams { /* PhysicsParamsUBO, Core/headers/scene/ParticleSystem.h:16-24 */
  float particle_restitution, timestep, friction_coefficient, repulsion_factor, penetration_slop, restitution_slop, sleep_threshold;
} SnkPhysicsParams;
SNKPT_API void snkpt_particles_default_params(SnkPhysicsParams* p);
/* ParticleSystem::InitBuffers + InitParticles (:14-222): allocates `count` particles; init != 0 runs Particle.comp's Init */
SNKPT_API int snkpt_particles_create(SnkPt* ctx, uint32_t count, int init);
SNKPT_API int snkpt_particles_write(SnkPt* ctx, const SnkParticle* src, uint32_t first, uint32_t count);
SNKPT_API int snkpt_particles_read(SnkPt* ctx, SnkParticle* dst, uint32_t first, uint32_t count);
/* ParticleSystem::UpdateParticles: `substeps` sub-steps (the reference runs 3 per frame) against the current TLAS and the
 * current / previous transforms; params NULL = defaults. rays_traced (optional) receives the rays of this call after a sync. */
SNKPT_API int snkpt_particles_update(SnkPt* ctx, const SnkPhysicsParams* params, uint32_t substeps, void* stream);
SNKPT_API int snkpt_particles_rays(SnkPt* ctx, uint64_t* rays_traced);
/* SortUintsByKey (declared, never defined: Core/src/scene/ParticleSystem.cpp:14 — `extern "C" void SortUintsByKey(unsigned*
 * device_keys, unsigned* device_values, unsigned count)`; the engine's own key sort is the bitonic chain of
 * ParticleSystem::UpdateParticles): stable ascending sort of `count` DEVICE-resident (key, value) pairs in place, on `stream`
 * (NULL = the context's stream) — the one-sweep radix sort the BVH builder and the particle grid use. count < 2^30. */
SNKPT_API int snkpt_sort_uints_by_key(SnkPt* ctx, uint32_t* device_keys, uint32_t* device_values, uint32_t count, void* stream);

/* ---- ray queries --------------------------------------------------------------------- */
/* traceRayEXT semantics (closest hit in (tmin,tmax), watertight, no culling;
 * ties broken towards the lowest (inst, prim)) over the current TLAS. Used by
 * the parity tests, the traversal benchmark and the particle pass (next row). */
SNKPT_API int snkpt_trace_rays(SnkPt* ctx, const SnkRay* rays, uint64_t count, uint32_t flags,
                               SnkHit* hits, void* stream);
/* Diagnostic: per ray, the number of wide nodes / triangle records / instances
 * whose box overlaps [tmin, t_hit] — the traversal-order independent
 * "algorithmic bytes" of SURVEY §8(d). out3[3*i+{0,1,2}] = nodes, tris, instances. */
SNKPT_API int snkpt_count_work(SnkPt* ctx, const SnkRay* rays, const SnkHit* hits, uint64_t count,
                               uint32_t flags, uint32_t* out3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNKPT_H */
