// context.h — the per-GPU state behind the C ABI (include/snkpt.h).
//
// Mirrors what the reference spreads over RT (Raytracing.h:11-31), MeshBufferManager,
// the scene buffer systems and EditorLayer's RaytracingResources: global SoA mesh
// buffers, one wide BVH per submesh, instance / transform / material / light
// mirrors, the TLAS, the G-buffer planes and the two reservoir halves.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "bvh_build.cuh"
#include "layouts.h"
#include "traverse.cuh"

namespace snk {

// growable device buffer (MeshBufferManager::Resize semantics: new buffer + device copy)
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool view = false;  // p points into another DevBuf's allocation (never freed or grown through this object)
  void alias(void* q, size_t n) { if (!view) release(); p = q; bytes = n; view = true; }
  cudaError_t reserve(size_t need, bool keep, cudaStream_t st) {
    if (need <= bytes) return cudaSuccess;
    if (view) return cudaErrorInvalidValue;
    size_t nb = bytes ? bytes : 4096;
    while (nb < need) nb += nb / 2 + 4096;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, nb);
    if (e != cudaSuccess) return e;
    if (keep && p && bytes) {
      e = cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (p) cudaFree(p);
    p = q; bytes = nb;
    return e;
  }
  void release() { if (p && !view) cudaFree(p); p = nullptr; bytes = 0; view = false; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct BlasSrc { uint32_t index_offset, vertex_offset, num_tris; };
// object-space bounding sphere of a BLAS's vertex range (host side; the TLAS build turns it into a world-space sphere per instance)
struct BlasSphere { uint32_t v_begin = 0, v_count = 0; double c[3] = {0, 0, 0}; double r = -1.0; };  // r < 0: unknown (no culling)

// one timed kernel launch (snkpt_set_profiling / snkpt_get_launch_timings)
struct ProfRec {
  char name[32];
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  uint64_t units = 0;
  bool is_trace = false, any = false, counted = false, queue = false;
};
constexpr int PROF_MAX = 2048;

// per-frame constants handed to the kernels by value
struct FrameConsts {
  float ivp[16];        // inverse(view) * inverse(proj) hoisted (Util.glsl:3), float64 rounded once
  float prev_ivp[16];   // the same of the previous frame's UBO (north-star temporal reuse: previous world position)
  float proj_view[16];
  float prev_proj_view[16];
  float cam_pos[3];
  float time_elapsed, delta_time;
  uint32_t frame_idx, prev_frame_invalid;
  uint32_t W, H, row0, row1;
  float jx, jy;
};

struct PixelPlanes {
  // G-buffer (formats of EditorLayer.cpp:229-256)
  char4* albedo; ushort4* normal; float* depth; ushort4* rma; ushort2* motion; uint16_t* flags;
  ushort4* prev_normal; float* prev_depth;
  // aux
  uint32_t *prim, *inst; float* hit_t; float2* bary;
  // packed surface record, 2 x uint4 per pixel: what a NEIGHBOUR's gather needs of the planes above in one 32-byte sector
  // (shading.cuh: pack_surf / SurfRec) — written with the G-buffer, read by the spatial-reuse stage of the north-star ReSTIR
  uint4* surf;
  // reservoirs: 2 halves x 3 planes of 16 B (SoA); half 0 = final, half 1 = temporal
  float4* res[2][3];
  // images
  float4 *radiance, *direct, *accum; ushort4* out16;
  // TAA (taa.cu): history ping-pong (RGBA16F) and the velocity of the previous resolve (RG16F)
  ushort4* taa_hist[2]; ushort2* taa_prev_vel[2];
};

// one entry of the albedo texture table (snkpt_set_texture): RGBA8 texels, base level only
struct TexDev { const uchar4* texels; uint32_t w, h, srgb, pad; };

struct SceneDev {
  const float *positions, *normals, *uvs, *tangents;
  const uint32_t* indices;
  const SnkInstanceData* instances;
  const float* transforms;       // mat4 column-major
  const float* prev_transforms;
  const float* normal_mats;      // 3x3 row-major (transpose(inverse(mat3(M)))), 12 floats stride
  const float* det3;
  const SnkMaterial* materials;
  const SnkLightData* lights;
  const uint32_t* emissive_idx;
  uint32_t num_emissive, num_instances;
  // emissive-triangle table of the north-star stages (northstar.cuh:EmTri): record em_base_of_obj[obj] + tri, or
  // em_base_of_obj[obj] == 0xFFFFFFFF when the instance is not in the emissive list
  const float4* em_table;
  const uint32_t* em_base_of_obj;
  const uint32_t* em_base;  // per emissive-list entry k: first record of emissive_idx[k]
  // albedo textures (TexMatBuffers.glsl `textures[]`): table indexed by Material::albedo_tex_idx, sRGB decode table
  const TexDev* textures;
  const float* srgb_lut;
  uint32_t num_textures;
};

struct GroupState;  // group.cu: row-band group this context belongs to (halo exchange of reservoir rows)
}  // namespace snk

struct SnkPt {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint32_t W = 0, H = 0, row0 = 0, row1 = 0;
  uint32_t halo_pass1 = 0, halo_gbuffer = 0;  // extra rows around the band (snkpt_set_halo)
  snk::GroupState* group = nullptr;           // row-band group (snkpt_group_attach_*): reservoir rows are exchanged, not re-rendered
  uint32_t max_instances = 4096, max_transforms = 4096, max_materials = 4096;

  // mesh pools
  snk::DevBuf positions, normals, tangents, uvs, indices;
  uint32_t num_vertices = 0, num_indices = 0;
  snk::DevBuf blas_nodes, blas_tris, d_blas;
  uint32_t n_blas_nodes = 0, n_blas_tris = 0;
  std::vector<SnkBlasInfo> blas;
  std::vector<snk::BlasSrc> blas_src;
  std::vector<snk::BlasSphere> blas_sphere;   // per BLAS, from the vertices handed to snkpt_upload_mesh / snkpt_update_positions
  std::vector<snk::RefitState> blas_refit;  // side arrays of refitted BLASes (empty until the first refit)
  snk::BuildScratch scratch_blas, scratch_tlas, scratch_ptcl;
  // particle pass (particles.cu): two particle buffers (ping-pong), cell start table, ray counter
  snk::DevBuf ptcl[2], ptcl_start, ptcl_rays;
  uint32_t num_particles = 0, ptcl_cur = 0;

  // scene mirrors (host + device)
  std::vector<SnkInstanceData> h_instances;
  std::vector<uint32_t> h_blas_of_instance, h_emissive;
  std::vector<float> h_transforms, h_prev_transforms;
  uint32_t num_instances = 0, num_emissive = 0, num_pointlights = 0;
  bool has_dir_light = false;  // host mirror: the directional light has a colour (snkpt_set_lights)
  snk::DevBuf materials, lights, transforms, prev_transforms, normal_mats, det3, instances, blas_of_instance, emissive_idx;
  snk::DevBuf tex_table, srgb_lut;  // albedo textures
  std::vector<snk::DevBuf> tex_data;
  std::vector<snk::TexDev> h_tex;
  snk::DevBuf em_table, em_base_of_obj, em_base;  // emissive-triangle table (rebuilt every north-star frame)
  uint32_t num_emissive_tris = 0;

  // TLAS
  snk::DevBuf tlas_nodes, tlas_leaf_ids, inst_recs, valid_slots;  // inst_recs, valid_slots, inst_sphere: views into tlas_stage
  snk::DevBuf tlas_stage;  // instance records | valid slots | spheres, laid out like the pinned staging block: ONE upload per TLAS build
  // The per-frame TLAS rebuild runs AHEAD of the frame before it (api.cu:snkpt_build_tlas): it writes the other of two copies of the
  // TLAS storage on its own stream and only waits for the frames that read that copy (everything enqueued before the PREVIOUS
  // rebuild) and for scene mutations, not for the frame in flight. *_alt hold the copy the in-flight frame reads.
  snk::DevBuf tlas_nodes_alt, tlas_leaf_ids_alt, tlas_stage_alt;
  cudaStream_t tlas_stream = nullptr;
  cudaEvent_t ev_tlas_prev[2] = {nullptr, nullptr};  // recorded on the frame's stream at each rebuild call (alternating)
  cudaEvent_t ev_tlas_built = nullptr, ev_tlas_sync = nullptr;
  uint32_t tlas_builds = 0;
  uint64_t scene_epoch = 1, tlas_epoch = 0;  // scene_epoch: bumped by every call that changes what the rebuild reads
  snk::DevBuf inst_sphere;  // float4 per instance slot: world-space bounding sphere (xyz, r^2), r^2 < 0 = none (traverse.cuh)
  uint32_t tlas_prims = 0;
  bool tlas_built = false;
  // small per-frame uploads (transforms, instances, lights, materials) go through a ring of pinned chunks and are enqueued
  // on the stream the frame runs on: no stream synchronisation per call, ordering with the frame's kernels by stream order
  static constexpr int RING_CHUNKS = 8;
  static constexpr size_t RING_CHUNK_BYTES = 1u << 20;
  uint8_t* ring = nullptr;
  int ring_cur = 0;
  size_t ring_used = 0;
  cudaEvent_t ring_ev[RING_CHUNKS] = {};
  bool ring_armed[RING_CHUNKS] = {};
  cudaStream_t work_stream = nullptr;  // stream of the last frame call (snkpt_build_tlas / trace_primary / render), default: `stream`
  // asynchronous read-back of the output image (snkpt_read_output_async): own copy stream, two slots
  // north-star frame: low-priority stream of the ReSTIR chain (northstar.cu), forked from / joined to the frame's stream
  cudaStream_t aux_stream = nullptr, hi_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join_b = nullptr;
  cudaEvent_t ev_path[20] = {};  // path chain: after ns_gen, after every traversal launch (staggered start of the ReSTIR chain)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_frame_done = nullptr, ev_read_done[4] = {};
  bool read_pending[4] = {false, false, false, false};
  void* h_stage = nullptr;       // pinned staging of the per-frame instance records (snkpt_build_tlas)
  size_t h_stage_bytes = 0;
  cudaEvent_t ev_stage = nullptr;
  bool stage_in_flight = false, build_timing_pending = false;

  // per pixel
  snk::DevBuf px[32];
  // second copy of the G-buffer planes (albedo, normal, depth, rma, motion, flags, prim, inst, hit_t, bary, surf): snkpt_trace_primary
  // alternates between the two, so the previous frame's normals / depths are simply the other copy (no blit), and the camera rays of
  // frame f+1 may run under the tail of frame f (api.cu). gb_cur = the copy the last snkpt_trace_primary wrote.
  snk::DevBuf px_alt[11];
  uint32_t gb_cur = 0, gb_calls = 0;
  uint64_t gb_epoch = 0;
  cudaEvent_t ev_gb_prev[2] = {nullptr, nullptr}, ev_gb_done = nullptr, ev_gb_sync = nullptr;
  // the frame in flight reads the OTHER copy's normals / depths (its "previous frame") in its temporal-reuse stage: camera rays that run
  // ahead overwrite that copy, so they also wait for ev_prev_free (recorded after that stage); prev_free = 0: no such event for the last
  // render (reference mode reads the previous normals until its last pass) -> the camera rays stay in order
  cudaEvent_t ev_prev_free = nullptr;
  int prev_free = 1;  // 1 before any render: nobody is reading
  bool prev_free_event = false;
  snk::PixelPlanes planes{};
  bool history_valid = false;
  uint32_t history_mode = 0;    // estimator that wrote the reservoir history (the two modes' reservoirs are not interchangeable)
  bool gbuffer_valid = false;
  bool render_pending = false;  // a frame was enqueued since the last stats read
  uint32_t accum_count = 0;
  bool taa_valid = false;       // the TAA history holds a resolved frame
  uint32_t taa_cur = 0;         // ping-pong index of the current history / previous-velocity images

  // frame
  SnkCommonUBO ubo{}, prev_ubo{};
  bool frame_set = false;
  snk::FrameConsts fc{};

  // generic ray query scratch
  snk::DevBuf q_rays, q_hits, q_work;
  snk::DevBuf wave[12];      // wavefront queues of the bounce stages (rays, hits, shading records)
  snk::DevBuf wave_refs;     // compact ray-reference queue of the north-star frame (trace_q.cu)
  snk::DevBuf wave_paths;    // compact lists of the paths alive after each vertex stage (northstar.cu)
  // u32 view: [0] traversal stack overflows, [2..3] shadow rays of the frame, [4..5] closest-hit rays, [8..11] since-create
  // totals; u64 view [16..31]: ray tickets of the persistent traversal launches
  snk::DevBuf dev_counters;
  snk::DevBuf trav_spill;    // global overflow of the short traversal stacks (traverse.cuh)
  uint32_t ticket_rr = 0;
  uint32_t primary_parity = 0, shadow_parity = 0;
  std::vector<uint32_t> band_starts;  // snkpt_group_set_bands: row boundaries of the group's bands (empty: equal bands)  // which of the two tile tickets the next k_primary_packet launch draws from

  SnkPtStats stats{};
  cudaEvent_t ev[8] = {};

  // per-launch profiling
  uint32_t profiling = 0;
  std::vector<snk::ProfRec> prof = std::vector<snk::ProfRec>(snk::PROF_MAX);
  int prof_n = 0;
  snk::DevBuf prof_work;  // 4 x u64 per record: live rays, nodes, leaf records, instances
  snk::DevBuf prof_hits;  // level 2: hit distances of any-hit launches

  snk::TraceScene trace_scene() const;
  snk::SceneDev scene_dev() const;
};
