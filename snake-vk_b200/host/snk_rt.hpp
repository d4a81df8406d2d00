// snk_rt.hpp — C++ host-side mirror of the reference's pathtracing-pass class over the C ABI (header only).
//
// Same surface as `SNAKE::RT` (Core/headers/rendering/Raytracing.h:11-31): RtSettings {num_temporal_resamples,
// num_spatial_resamples} (:13-16), Init (:18; Raytracing.cpp:269-278) and RecordRenderCmdBuf (:20-24; Raytracing.cpp:227-267).
// Where the reference pulls scene data from engine singletons and records Vulkan commands, this class takes a
// snkscene front-end (include/snkscene.h) or raw snkpt_set_* calls and enqueues CUDA work on a stream. Errors throw
// std::runtime_error with the library's message (the reference logs and breaks, Core/headers/util/util.h:7-31).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

#include "snkpt.h"
#include "snkscene.h"

namespace snk {

class RT {
 public:
  struct RtSettings {  // Raytracing.h:13-16
    int32_t num_temporal_resamples = 1;
    int32_t num_spatial_resamples = 3;
  };
  enum class Estimator : uint32_t { Reference = SNKPT_MODE_REFERENCE, NorthStar = SNKPT_MODE_NORTHSTAR };

  RT() = default;
  RT(const RT&) = delete;
  RT& operator=(const RT&) = delete;
  ~RT() { if (m_ctx) snkpt_destroy(m_ctx); }

  // RT::Init + EditorLayer::InitGBuffer / InitRenderResources: output image, G-buffer planes, reservoirs
  void Init(uint32_t width, uint32_t height, int device = 0) {
    SnkPtConfig cfg{};
    cfg.width = width; cfg.height = height; cfg.device = device;
    if (snkpt_create(&cfg, &m_ctx) != SNKPT_OK) throw std::runtime_error(std::string("snkpt_create: ") + snkpt_last_error(nullptr));
  }

  // MeshBufferManager::LoadMeshFromData + BLAS::GenerateFromMeshData; registers the result with the scene front-end
  uint32_t LoadMesh(SnkScene* scene, const float* positions, const float* normals, const float* tangents, const float* uvs, uint32_t num_vertices,
                    const uint32_t* indices, uint32_t num_indices, const SnkSubmesh* submeshes, uint32_t num_submeshes) {
    SnkMeshEntry entry{};
    Check(snkpt_upload_mesh(m_ctx, positions, normals, tangents, uvs, num_vertices, indices, num_indices, submeshes, num_submeshes, &entry, nullptr),
          "snkpt_upload_mesh");
    return snkscene_register_mesh(scene, &entry, submeshes, num_submeshes);
  }

  // one frame: the FrameStart buffer updates (front-end), GBufferPass::RecordCommandBuffer, RT::RecordRenderCmdBuf, optional TAA
  void RecordRenderCmdBuf(SnkScene* scene, const RtSettings& settings, float time_elapsed_ms, float delta_ms, Estimator estimator = Estimator::NorthStar,
                          bool taa = false, void* stream = nullptr) {
    if (snkscene_frame_start(scene, time_elapsed_ms, delta_ms) != SNKPT_OK) throw std::runtime_error(snkscene_last_error(scene));
    if (snkscene_push(scene, m_ctx, stream) != SNKPT_OK) throw std::runtime_error(snkscene_last_error(scene));
    float jitter[2], pixels[2];
    snkpt_frame_jitter(snkscene_ubo(scene, 0)->frame_idx, m_width_hint, m_width_hint, jitter, pixels);
    Check(snkpt_trace_primary(m_ctx, m_jitter ? pixels : nullptr, stream), "snkpt_trace_primary");
    SnkRtSettings s;
    snkpt_default_settings(&s);
    s.num_temporal_resamples = settings.num_temporal_resamples;
    s.num_spatial_resamples = settings.num_spatial_resamples;
    s.mode = (uint32_t)estimator;
    Check(snkpt_render(m_ctx, &s, stream), "snkpt_render");
    if (taa) Check(snkpt_taa_resolve(m_ctx, stream), "snkpt_taa_resolve");
    snkscene_frame_end(scene);
  }

  void EnableJitter(bool on, uint32_t render_width) { m_jitter = on; m_width_hint = render_width; }
  void ReadOutput(void* rgba16f) { Check(snkpt_read_output(m_ctx, rgba16f, SNKPT_FMT_RGBA16F), "snkpt_read_output"); }
  SnkPt* context() const { return m_ctx; }

 private:
  void Check(int rc, const char* what) {
    if (rc != SNKPT_OK) throw std::runtime_error(std::string(what) + ": " + snkpt_last_error(m_ctx));
  }
  SnkPt* m_ctx = nullptr;
  bool m_jitter = false;
  uint32_t m_width_hint = 1920;
};

}  // namespace snk
