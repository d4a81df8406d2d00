// gbuffer_nm.cu — normal maps of the G-buffer (GBuffer.vert:49-64 CalculateTbnMatrix, GBuffer.frag:48-52), as a second pass over
// the pixels that is launched only when textures are set: the primary-visibility kernel (gbuffer.cu) keeps its code and register
// budget for the untextured case. CPU statement: oracle/snk_scene.h Scene::mapped_normal. Compiled with -fmad=false.
#include "kernels.h"
#include "shading.cuh"

namespace snk {

namespace {

// normal = TBN (interpolated per-vertex) x texel, for the pixels whose material has a normal map
__global__ void __launch_bounds__(256) k_gbuffer_normal_map(SceneDev D, FrameConsts fc, PixelPlanes P) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t npix = fc.W * (fc.row1 - fc.row0);
  if (i >= npix) return;
  const size_t px = (size_t)fc.row0 * fc.W + i;
  const uint32_t prim = P.prim[px];
  if (prim == INVALID_IDX) return;
  const SnkInstanceData in = load_instance(D, P.inst[px]);
  const SnkMaterial* m = D.materials + in.material_idx;
  if (__ldg(&m->normal_tex_idx) == INVALID_IDX) return;
  uint32_t ia, ib, ic;
  tri_indices(D, in, prim, ia, ib, ic);
  const float2 bc = P.bary[px];
  const float b0 = (1.0f - bc.x) - bc.y;
  vec3 nm;
  if (mapped_normal(D, m, D.transforms + 16ull * in.transform_idx, ia, ib, ic, b0, bc.x, bc.y, nm)) {
    const ushort4 n16 = make_ushort4(f2h(nm.x), f2h(nm.y), f2h(nm.z), f2h(1.0f));
    P.normal[px] = n16;
    // the packed surface record repeats the normal (shading.cuh: pack_surf)
    uint4 a = P.surf[2 * px];
    a.x = (uint32_t)n16.x | ((uint32_t)n16.y << 16);
    a.y = (a.y & 0xFFFF0000u) | (uint32_t)n16.z;
    P.surf[2 * px] = a;
  }
}

}  // namespace

int launch_gbuffer_normal_map(const SceneDev& D, const FrameConsts& fc, const PixelPlanes& P, cudaStream_t st) {
  const uint32_t npix = fc.W * (fc.row1 - fc.row0);
  if (npix == 0) return 0;
  k_gbuffer_normal_map<<<(npix + 255) / 256, 256, 0, st>>>(D, fc, P);
  return 1;
}

}  // namespace snk
