// gbuffer.cu — primary visibility with camera rays.
//
// The reference rasterises its G-buffer (GBufferPass::RecordCommandBuffer,
// Core/src/renderpasses/GBufferPass.cpp:70-196; GBuffer.vert:66-96; GBuffer.frag:41-72);
// here one camera ray per pixel (Util.glsl:9-13 CalcRayDir) produces the same
// quantities in the same storage formats (Editor/src/EditorLayer.cpp:229-256) plus
// the primitive / instance ids the parity gate is read on. The prev-normal /
// prev-depth blit (GBufferPass.cpp:73-77) is folded into the same kernel: each
// thread moves its own texel to the prev planes before overwriting it.
// Compiled with -fmad=false (see vecmath.cuh).
#include <cstdlib>

#include "kernels.h"
#include "packet.cuh"
#include "shading.cuh"

namespace snk {

namespace {

constexpr float PRIMARY_TMIN = 0.1f;  // CameraComponent.h:11 z_near
constexpr float PRIMARY_TMAX = 10000.0f;

#ifndef SNK_PRIMARY_CTAS
#define SNK_PRIMARY_CTAS 6
#endif
constexpr int PRIMARY_CTAS_PER_SM = SNK_PRIMARY_CTAS;

// G-buffer planes of one pixel from its primary hit (GBuffer.vert:66-96, GBuffer.frag:41-72) + the ids of the parity gate
__device__ __forceinline__ void write_gbuffer(const SceneDev& D, const FrameConsts& fc, const PixelPlanes& P, size_t px, const HitRec& h) {
  P.prim[px] = h.prim; P.inst[px] = h.inst; P.hit_t[px] = h.t;
  if (h.t < 0.0f) {
    P.bary[px] = make_float2(0.f, 0.f);
    P.albedo[px] = make_char4(0, 0, 0, 0);
    P.normal[px] = make_ushort4(0, 0, 0, 0);
    P.rma[px] = make_ushort4(0, 0, 0, 0);
    P.depth[px] = 1.0f;
    P.motion[px] = make_ushort2(0, 0);
    P.flags[px] = 0;
    pack_surf(P.surf, px, make_ushort4(0, 0, 0, 0), 0u, h.t, 1.0f, make_char4(0, 0, 0, 0), make_ushort4(0, 0, 0, 0), false);
    return;
  }
  P.bary[px] = make_float2(h.u, h.v);
  const SnkInstanceData in = D.instances[h.inst];
  const uint32_t base = 3u * (h.prim + in.mesh_buffer_index_offset / 3u);
  const uint32_t ia = D.indices[base] + in.mesh_buffer_vertex_offset, ib = D.indices[base + 1] + in.mesh_buffer_vertex_offset,
                 ic = D.indices[base + 2] + in.mesh_buffer_vertex_offset;
  const float b0 = (1.0f - h.u) - h.v, b1 = h.u, b2 = h.v;
  const vec3 pa = v3(D.positions[3 * ia], D.positions[3 * ia + 1], D.positions[3 * ia + 2]);
  const vec3 pb = v3(D.positions[3 * ib], D.positions[3 * ib + 1], D.positions[3 * ib + 2]);
  const vec3 pc = v3(D.positions[3 * ic], D.positions[3 * ic + 1], D.positions[3 * ic + 2]);
  const vec3 na = v3(D.normals[3 * ia], D.normals[3 * ia + 1], D.normals[3 * ia + 2]);
  const vec3 nb = v3(D.normals[3 * ib], D.normals[3 * ib + 1], D.normals[3 * ib + 2]);
  const vec3 nc = v3(D.normals[3 * ic], D.normals[3 * ic + 1], D.normals[3 * ic + 2]);
  const vec3 p = (pa * b0 + pb * b1) + pc * b2;
  const vec3 n0 = (na * b0 + nb * b1) + nc * b2;
  const float* M = D.transforms + 16ull * in.transform_idx;
  const vec3 world = mul_point(M, p);
  const vec3 n = normalize(mul_mat3r(D.normal_mats + 9ull * in.transform_idx, n0));  // normal maps: k_gbuffer_normal_map below
  const float4 clip = mul_vec4(fc.proj_view, world.x, world.y, world.z, 1.0f);
  const vec3 pworld = mul_point(D.prev_transforms + 16ull * in.transform_idx, p);
  const float4 pclip = mul_vec4(fc.prev_proj_view, pworld.x, pworld.y, pworld.z, 1.0f);
  const float depth = clip.z / clip.w;
  P.depth[px] = depth;
  const float mx = (pclip.x / pclip.w - clip.x / clip.w) * 0.5f, my = (pclip.y / pclip.w - clip.y / clip.w) * 0.5f;
  P.motion[px] = make_ushort2(f2h(mx), f2h(my));
  const SnkMaterial* m = D.materials + in.material_idx;
  const vec3 alb = albedo_at(D, m, ia, ib, ic, b0, b1, b2);  // GBuffer.frag:56-58
  const char4 alb8 = make_char4(f2snorm8(alb.x), f2snorm8(alb.y), f2snorm8(alb.z), 127);
  const ushort4 n16 = make_ushort4(f2h(n.x), f2h(n.y), f2h(n.z), f2h(1.0f));
  const ushort4 rma16 = make_ushort4(f2h(m->roughness), f2h(m->metallic), f2h(m->ao), f2h(1.0f));
  const uint16_t mflags = (uint16_t)m->flags;
  P.albedo[px] = alb8;
  P.normal[px] = n16;
  P.rma[px] = rma16;
  P.flags[px] = mflags;
  pack_surf(P.surf, px, n16, mflags, h.t, depth, alb8, rma16, true);
}

__global__ void __launch_bounds__(128, PRIMARY_CTAS_PER_SM) k_primary(TraceScene S, SceneDev D, FrameConsts fc, PixelPlanes P) {
  // persistent CTAs walk 16x8 pixel tiles, 8x4 pixels per warp (coherent primary rays)
  __shared__ uint2 sm_stack[SHORT_STACK * 128];
  const uint32_t t = threadIdx.x, w = t >> 5, l = t & 31;
  const StackRef stk{sm_stack + t, S.spill + ((size_t)blockIdx.x * 4 + w) * (SPILL_STACK * 32) + l};
  const uint32_t tiles_x = (fc.W + 15u) / 16u, n_tiles = tiles_x * ((fc.row1 - fc.row0 + 7u) / 8u);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const uint32_t x = (tile % tiles_x) * 16 + (w & 1) * 8 + (l & 7);
  const uint32_t y = fc.row0 + (tile / tiles_x) * 8 + (w >> 1) * 4 + (l >> 3);
  const bool in_range = x < fc.W && y < fc.row1;
  const size_t px = in_range ? (size_t)y * fc.W + x : 0;
  // (the previous frame's normals / depths are the other copy of the G-buffer planes: P.prev_normal / P.prev_depth point there)

  const float u = (((float)x + 0.5f) + fc.jx) / (float)fc.W, v = (((float)y + 0.5f) + fc.jy) / (float)fc.H;
  const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
  const vec3 far = world_pos_from_depth(fc.ivp, 1.0f, u, v);
  const vec3 d = normalize(far - cam);
  HitRec h;
  trace_warp_body<false, 128>(S, make_float4(cam.x, cam.y, cam.z, PRIMARY_TMIN), make_float4(d.x, d.y, d.z, in_range ? PRIMARY_TMAX : -1.0f),
                              stk, h);
  if (!in_range) continue;
  write_gbuffer(D, fc, P, px, h);
  }
}

#ifndef SNK_PACKET_CTAS
#define SNK_PACKET_CTAS 8
#endif
__global__ void __launch_bounds__(128, SNK_PACKET_CTAS) k_primary_packet(TraceScene S, SceneDev D, FrameConsts fc, PixelPlanes P, uint32_t* ticket,
                                                                         uint32_t* next_ticket) {
  __shared__ uint32_t sm_stack[4 * PK_STACK];
  const uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const PkStack stk{sm_stack + w * PK_STACK, reinterpret_cast<uint32_t*>(S.spill + ((size_t)blockIdx.x * 4 + w) * (SPILL_STACK * 32))};
  // the warps are independent: each draws its 8x4 tiles from a ticket (four consecutive tickets = one 16x8 block, so
  // neighbouring warps work on neighbouring pixels); tiles of sky and of dense geometry differ 10x in cost
  const uint32_t tiles_x = (fc.W + 15u) / 16u, n_tiles = 4u * tiles_x * ((fc.row1 - fc.row0 + 7u) / 8u);
  if (blockIdx.x == 0 && threadIdx.x == 0) *next_ticket = 0u;  // the next launch's ticket (stream order: nobody reads it before)
  for (;;) {
    uint32_t tile = 0;
    if (l == 0) tile = atomicAdd(ticket, 1u);
    tile = __shfl_sync(PK_FULL, tile, 0);
    if (tile >= n_tiles) break;
    const uint32_t q = tile & 3u, blk = tile >> 2;
    const uint32_t x = (blk % tiles_x) * 16 + (q & 1) * 8 + (l & 7);
    const uint32_t y = fc.row0 + (blk / tiles_x) * 8 + (q >> 1) * 4 + (l >> 3);
    const bool in_range = x < fc.W && y < fc.row1;
    const size_t px = in_range ? (size_t)y * fc.W + x : 0;
    const float u = (((float)x + 0.5f) + fc.jx) / (float)fc.W, v = (((float)y + 0.5f) + fc.jy) / (float)fc.H;
    const vec3 cam = v3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const vec3 far = world_pos_from_depth(fc.ivp, 1.0f, u, v);
    const vec3 d = normalize(far - cam);
    HitRec h;
    pk_trace(S, cam.x, cam.y, cam.z, d.x, d.y, d.z, PRIMARY_TMIN, in_range ? PRIMARY_TMAX : -1.0f, stk, l, h);
    if (in_range) write_gbuffer(D, fc, P, px, h);
  }
}

__global__ void k_accumulate(const float4* __restrict__ radiance, float4* __restrict__ accum, ushort4* __restrict__ out16,
                             uint32_t npix, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float4 r = radiance[i];
  float4 a = count ? accum[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float inv = 1.0f / (float)(count + 1);
  a.x += (r.x - a.x) * inv; a.y += (r.y - a.y) * inv; a.z += (r.z - a.z) * inv; a.w = 1.0f;
  accum[i] = a;
  out16[i] = make_ushort4(f2h(a.x), f2h(a.y), f2h(a.z), f2h(1.0f));
}

}  // namespace

int launch_primary(const TraceScene& S, const SceneDev& D, const FrameConsts& fc, const PixelPlanes& P, uint32_t* tickets2, uint32_t parity,
                   cudaStream_t st) {
  if (fc.row1 <= fc.row0) return 0;
  const uint32_t n_tiles = ((fc.W + 15) / 16) * ((fc.row1 - fc.row0 + 7) / 8);
  // SNKPT_PRIMARY=ray selects the one-ray-per-lane kernel (A/B measurements; same bits either way)
  static const bool one_ray = getenv("SNKPT_PRIMARY") && getenv("SNKPT_PRIMARY")[0] == 'r';
  if (one_ray) {
    const uint32_t resident = (uint32_t)device_sm_count() * PRIMARY_CTAS_PER_SM;
    k_primary<<<n_tiles < resident ? n_tiles : resident, 128, 0, st>>>(S, D, fc, P);
  } else {
    // two tickets used in turn: a launch draws from one and zeroes the other for its successor (no memset between frames)
    const uint32_t resident = (uint32_t)device_sm_count() * SNK_PACKET_CTAS;
    k_primary_packet<<<n_tiles < resident ? n_tiles : resident, 128, 0, st>>>(S, D, fc, P, tickets2 + (parity & 1u), tickets2 + ((parity & 1u) ^ 1u));
  }
  return 1 + (D.num_textures ? launch_gbuffer_normal_map(D, fc, P, st) : 0);
}

int launch_accumulate(const PixelPlanes& P, uint32_t npix, uint32_t count, cudaStream_t st) {
  k_accumulate<<<(npix + 255) / 256, 256, 0, st>>>(P.radiance, P.accum, P.out16, npix, count);
  return 1;
}

}  // namespace snk
