// example_rt.cpp — smallest C++ program over snk_rt.hpp + snkscene.h: one triangle-pair floor, a point light, a camera, two frames.
// Built by tests/test_cpu_scene.py (compile + link check on the CPU box) and runnable on a GPU box: prints the mean of the image.
#include <cstdio>
#include <vector>

#include "snk_rt.hpp"

int main() {
  try {
    const uint32_t W = 64, H = 48;
    snk::RT rt;
    rt.Init(W, H);
    SnkScene* scene = snkscene_create();
    const float pos[] = {-5, 0, -5, 5, 0, -5, 5, 0, 5, -5, 0, 5}, nrm[] = {0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0}, uv[] = {0, 0, 1, 0, 1, 1, 0, 1};
    const uint32_t idx[] = {0, 2, 1, 0, 3, 2};
    const SnkSubmesh sub{6, 4, 0, 0, 0};
    const uint32_t mesh = rt.LoadMesh(scene, pos, nrm, nrm, uv, 4, idx, 6, &sub, 1);
    SnkMaterial m{};
    m.albedo_tex_idx = m.normal_tex_idx = m.roughness_tex_idx = m.metallic_tex_idx = m.ao_tex_idx = SNKPT_INVALID_INDEX;
    m.roughness = 0.5f; m.ao = 1.0f; m.albedo[0] = m.albedo[1] = m.albedo[2] = 0.8f; m.albedo[3] = 1.0f;
    snkscene_set_material(scene, 0, &m);
    const uint32_t floor = snkscene_create_entity(scene, SNKSCENE_NONE), slot0 = 0;
    snkscene_set_static_mesh(scene, floor, mesh, &slot0, 1);
    const float lp[] = {0, 2, 0}, lc[] = {8, 8, 8}, cp[] = {0, 2, -6}, ce[] = {-15, 180, 0}, white[] = {1, 1, 1};
    const uint32_t lamp = snkscene_create_entity(scene, SNKSCENE_NONE);
    snkscene_set_transform(scene, lamp, lp, nullptr, nullptr);
    snkscene_set_pointlight(scene, lamp, lc, 0.05f, 0.25f, 0.5f);
    snkscene_set_directional_light(scene, white, 0.8f, 0.9f);
    const uint32_t cam = snkscene_create_entity(scene, SNKSCENE_NONE);
    snkscene_set_transform(scene, cam, cp, ce, nullptr);
    snkscene_set_camera(scene, cam, 90.f, (float)W / H, 0.1f, 2500.f);
    std::vector<uint16_t> img(4 * W * H);
    for (int f = 0; f < 2; ++f) rt.RecordRenderCmdBuf(scene, snk::RT::RtSettings{}, 1000.f * f / 60.f, 16.6667f);
    rt.ReadOutput(img.data());
    unsigned long long nonzero = 0;
    for (size_t i = 0; i < img.size(); i += 4) nonzero += (img[i] & 0x7FFF) != 0;
    std::printf("example_rt: %llu of %u pixels lit\n", nonzero, W * H);
    snkscene_destroy(scene);
    return nonzero ? 0 : 2;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "example_rt: %s\n", e.what());
    return 1;
  }
}
