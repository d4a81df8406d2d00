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
  P.depth[px] = clip.z / clip.w;
  const float mx = (pclip.x / pclip.w - clip.x / clip.w) * 0.5f, my = (pclip.y / pclip.w - clip.y / clip.w) * 0.5f;
  P.motion[px] = make_ushort2(f2h(mx), f2h(my));
  const SnkMaterial* m = D.materials + in.material_idx;
  const vec3 alb = albedo_at(D, m, ia, ib, ic, b0, b1, b2);  // GBuffer.frag:56-58
  P.albedo[px] = make_char4(f2snorm8(alb.x), f2snorm8(alb.y), f2snorm8(alb.z), 127);
  P.normal[px] = make_ushort4(f2h(n.x), f2h(n.y), f2h(n.z), f2h(1.0f));
  P.rma[px] = make_ushort4(f2h(m->roughness), f2h(m->metallic), f2h(m->ao), f2h(1.0f));
  P.flags[px] = (uint16_t)m->flags;
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
  if (in_range) {
  P.prev_normal[px] = P.normal[px];
  P.prev_depth[px] = P.depth[px];
  }

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

// ---- packet traversal of the camera rays (k_primary_packet) ---------------------------------------------------------------
// The 32 camera rays of an 8x4 pixel tile share their origin and span a thin pyramid. Instead of 32 independent state machines
// (k_primary: every lane decodes all 8 child boxes of its own node, 262 instructions per node and lane, 24 of 32 lanes busy) the
// WARP walks the BVH once with one shared stack, and the box work of a node is spread over the lanes: lane (j, k) = (lane & 7,
// lane >> 3) decodes ONE child box and tests it against ONE of the four side planes of the tile's pyramid; a ballot gives the
// children that lie outside some plane. Lanes with k = 0 also cull by distance: a box farther from the origin than every
// lane's current closest hit cannot matter. Children are visited in the octant order of the tile's centre ray. Only the
// triangle test stays per ray: every lane tests every triangle of the surviving leaves with exactly the arithmetic of the
// one-ray traversal (tri_test, the same instance transform), so the hits — closest t, lowest (instance, primitive) at equal t —
// are the ones the one-ray kernel and the brute-force oracle find: the plane and distance tests only have to be conservative
// (relative + absolute slack below), the set of triangles a ray sees only grows.
constexpr int PK_STACK = 128;               // shared stack entries per warp (u32)
constexpr uint32_t PK_INST = 0x80000000u;   // stack entry = a TLAS leaf record (instance) instead of a node

struct PkSpace {
  float ox, oy, oz;   // common ray origin in this space
  float nx, ny, nz;   // inward unit normal of this lane's side plane (k = lane >> 3)
  float eps_abs;      // absolute slack: float32 rounding of (coordinate - origin)
  uint32_t octinv;    // 7 - octant of the centre direction (children are visited in ascending (slot ^ octinv) ... descending, see below)
};
constexpr uint32_t PK_FULL = 0xFFFFFFFFu;

__device__ __forceinline__ float pk_warp_max(float v) {  // v >= 0
  return __uint_as_float(__reduce_max_sync(PK_FULL, __float_as_uint(v)));
}

// side planes of the pyramid spanned by the tile's corner rays (lanes 0, 7, 31, 24 in tile order), all through the origin
__device__ __forceinline__ void pk_setup_space(PkSpace& s, float ox, float oy, float oz, float dx, float dy, float dz, uint32_t lane) {
  const uint32_t k = lane >> 3;
  const int ca = k == 0 ? 0 : (k == 1 ? 7 : (k == 2 ? 31 : 24));
  const int cb = k == 0 ? 7 : (k == 1 ? 31 : (k == 2 ? 24 : 0));
  const float ax = __shfl_sync(PK_FULL, dx, ca), ay = __shfl_sync(PK_FULL, dy, ca), az = __shfl_sync(PK_FULL, dz, ca);
  const float bx = __shfl_sync(PK_FULL, dx, cb), by = __shfl_sync(PK_FULL, dy, cb), bz = __shfl_sync(PK_FULL, dz, cb);
  const float cx = __shfl_sync(PK_FULL, dx, 0) + __shfl_sync(PK_FULL, dx, 31), cy = __shfl_sync(PK_FULL, dy, 0) + __shfl_sync(PK_FULL, dy, 31),
              cz = __shfl_sync(PK_FULL, dz, 0) + __shfl_sync(PK_FULL, dz, 31);
  const float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
  const float len2 = nx * nx + ny * ny + nz * nz;
  float inv = len2 > 0.0f ? rsqrtf(len2) : 0.0f;  // a degenerate plane (n = 0) culls nothing
  if (nx * cx + ny * cy + nz * cz < 0.0f) inv = -inv;
  s.ox = ox; s.oy = oy; s.oz = oz;
  s.nx = nx * inv; s.ny = ny * inv; s.nz = nz * inv;
  s.eps_abs = 2e-6f * ((fabsf(ox) + fabsf(oy)) + fabsf(oz));
  const uint32_t oct = (cx < 0.f ? 1u : 0u) | (cy < 0.f ? 2u : 0u) | (cz < 0.f ? 4u : 0u);
  s.octinv = 7u - oct;
}

// One node for the whole warp. Lane j (= lane & 7) looks at child slot j ^ octinv, so ballot bits are in visiting order.
// Returns (warp-uniform): `inner` = surviving inner children by position, `leaf` = surviving leaf records (24-bit record mask);
// my_child = node index of this lane's child when it is a surviving inner one (lanes < 8).
__device__ __forceinline__ void pk_node(const uint4* __restrict__ nodes, uint32_t idx, const PkSpace& s, float far2, uint32_t lane,
                                        uint32_t& inner, uint32_t& leaf, uint32_t& rec_base, uint32_t& my_child) {
  uint4 n0, n1, n2, n3, n4;
  load_node(nodes, idx, n0, n1, n2, n3, n4);
  const uint32_t c = (lane & 7u) ^ s.octinv;
  // byte c of an 8-byte row = one PRMT over the word pair (selector nibble 0..7); the & 0xFF folds into the I2F.U8
  const uint32_t meta = __byte_perm(n1.z, n1.w, c) & 0xFFu;
  const float px = __uint_as_float(n0.x) - s.ox, py = __uint_as_float(n0.y) - s.oy, pz = __uint_as_float(n0.z) - s.oz;
  const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
              sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
  // child box relative to the origin (layouts.h: w2 = qlo_x | qlo_y, w3 = qlo_z | qhi_x, w4 = qhi_y | qhi_z)
  const float lox = fmaf((float)(__byte_perm(n2.x, n2.y, c) & 0xFFu), sx, px), hix = fmaf((float)(__byte_perm(n3.z, n3.w, c) & 0xFFu), sx, px);
  const float loy = fmaf((float)(__byte_perm(n2.z, n2.w, c) & 0xFFu), sy, py), hiy = fmaf((float)(__byte_perm(n4.x, n4.y, c) & 0xFFu), sy, py);
  const float loz = fmaf((float)(__byte_perm(n3.x, n3.y, c) & 0xFFu), sz, pz), hiz = fmaf((float)(__byte_perm(n4.z, n4.w, c) & 0xFFu), sz, pz);
  // largest signed distance of the box to this lane's plane; < 0 (beyond the slack): the box is outside the pyramid
  const float m = (fmaxf(s.nx * lox, s.nx * hix) + fmaxf(s.ny * loy, s.ny * hiy)) + fmaxf(s.nz * loz, s.nz * hiz);
  const float ext = ((fabsf(lox) + fabsf(hix)) + (fabsf(loy) + fabsf(hiy))) + (fabsf(loz) + fabsf(hiz));
  bool out = m < -(1e-5f * ext + s.eps_abs);
  // squared distance origin -> box against the farthest point any lane still cares about
  const float gx = fmaxf(fmaxf(lox, -hix), 0.0f), gy = fmaxf(fmaxf(loy, -hiy), 0.0f), gz = fmaxf(fmaxf(loz, -hiz), 0.0f);
  const float g2 = (gx * gx + gy * gy) + gz * gz;
  out = out || g2 * 0.9999f > far2 || meta == 0u;
  const uint32_t ob = __ballot_sync(PK_FULL, out);
  const uint32_t alive = ~(ob | (ob >> 8) | (ob >> 16) | (ob >> 24)) & 0xFFu;  // by position
  const bool is_inner = (meta & 0x18u) == 0x18u;
  const uint32_t ib = __ballot_sync(PK_FULL, is_inner) & 0xFFu;
  inner = alive & ib;
  const bool mine = lane < 8u && ((alive >> lane) & 1u);
  leaf = __reduce_or_sync(PK_FULL, (mine && !is_inner) ? ((meta >> 5) & 7u) << (meta & 0x1Fu) : 0u);
  rec_base = n1.y;
  my_child = n1.x + __popc((n0.w >> 24) & ~(0xFFFFFFFFu << c));
}

// parallel push of the surviving inner children: position 7 is visited first (trav_step_impl opens the highest bit first), so
// position p lands at stack offset rank(p) counted from below
__device__ __forceinline__ void pk_push_inner(uint32_t* stk, int& sp, uint32_t inner, uint32_t my_child, uint32_t lane, uint32_t* overflow) {
  const int n = __popc(inner);
  if (sp + n > PK_STACK) { if (lane == 0) atomicAdd(overflow, 1u); return; }
  if (lane < 8u && ((inner >> lane) & 1u)) stk[sp + __popc(inner & ~(0xFFFFFFFFu << lane))] = my_child;
  sp += n;
}

__device__ __forceinline__ void pk_trace(const TraceScene& S, float ox, float oy, float oz, float dx, float dy, float dz, float tmin, float tmax,
                                         uint32_t* stk, uint32_t lane, HitRec& best) {
  best.t = -1.0f; best.u = 0.f; best.v = 0.f; best.prim = 0xFFFFFFFFu; best.inst = 0xFFFFFFFFu;
  float tcur = tmax;
  if (S.tlas_node_count == 0u) return;
  PkSpace W;
  pk_setup_space(W, ox, oy, oz, dx, dy, dz, lane);
  const float ddw = (dx * dx + dy * dy) + dz * dz;
  float far2 = pk_warp_max(tcur > tmin ? (tcur * tcur) * ddw : 0.0f);
  if (!(far2 > 0.0f)) return;
  int sp = 0;
  stk[sp++] = 0u;  // TLAS root (every lane writes the same word)
  __syncwarp();
  while (sp > 0) {
    const uint32_t e = stk[--sp];
    __syncwarp();  // everybody has read the entry before anybody overwrites the slot
    if (!(e & PK_INST)) {
      uint32_t inner, leaf, rec_base, my_child;
      pk_node(S.tlas_nodes, e, W, far2, lane, inner, leaf, rec_base, my_child);
      pk_push_inner(stk, sp, inner, my_child, lane, S.overflow);
      // the instances of this node's leaf slots go on top: they are entered before the inner children are opened
      const int nl = __popc(leaf);
      if (sp + nl > PK_STACK) { if (lane == 0) atomicAdd(S.overflow, 1u); }
      else {
        if (lane < 24u && ((leaf >> lane) & 1u)) stk[sp + __popc(leaf & ~(0xFFFFFFFFu << lane))] = PK_INST | (rec_base + lane);
        sp += nl;
      }
      __syncwarp();
      continue;
    }
    // ---- one instance
    const uint32_t slot = __ldg(S.tlas_leaf_ids + (e & ~PK_INST));
    const bool may = tcur > tmin && (!S.inst_sphere || ray_may_hit_sphere(__ldg(S.inst_sphere + slot), ox, oy, oz, dx, dy, dz, tcur));
    if (!__any_sync(PK_FULL, may)) continue;
    const float4* M = reinterpret_cast<const float4*>(S.inst + slot);
    const float4 r0 = __ldg(M), r1 = __ldg(M + 1), r2 = __ldg(M + 2);
    const uint4 meta = __ldg(reinterpret_cast<const uint4*>(M + 3));
    RaySpace r;
    // o' = ((m0*x + m1*y) + m2*z) + m3 ; d' = (m0*x + m1*y) + m2*z   (unfused, as trav_step_impl and the oracle)
    r.ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, ox), __fmul_rn(r0.y, oy)), __fmul_rn(r0.z, oz)), r0.w);
    r.oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, ox), __fmul_rn(r1.y, oy)), __fmul_rn(r1.z, oz)), r1.w);
    r.oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, ox), __fmul_rn(r2.y, oy)), __fmul_rn(r2.z, oz)), r2.w);
    r.dx = __fadd_rn(__fadd_rn(__fmul_rn(r0.x, dx), __fmul_rn(r0.y, dy)), __fmul_rn(r0.z, dz));
    r.dy = __fadd_rn(__fadd_rn(__fmul_rn(r1.x, dx), __fmul_rn(r1.y, dy)), __fmul_rn(r1.z, dz));
    r.dz = __fadd_rn(__fadd_rn(__fmul_rn(r2.x, dx), __fmul_rn(r2.y, dy)), __fmul_rn(r2.z, dz));
    setup_tri_space(r);
    PkSpace B;
    pk_setup_space(B, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, lane);
    // t is a parameter, not a distance, in object space: lane l reaches at most |d'_l| * tcur_l from the origin
    const float ddo = (r.dx * r.dx + r.dy * r.dy) + r.dz * r.dz;
    float far2o = pk_warp_max(tcur > tmin ? ((tcur * tcur) * ddo) * 1.00001f : 0.0f);
    const uint32_t cur_inst = meta.y;
    const int base = sp;
    bool took_any = false;
    stk[sp++] = meta.x;
    __syncwarp();
    while (sp > base) {
      const uint32_t node = stk[--sp];
      __syncwarp();
      uint32_t inner, leaf, rec_base, my_child;
      pk_node(S.blas_nodes, node, B, far2o, lane, inner, leaf, rec_base, my_child);
      pk_push_inner(stk, sp, inner, my_child, lane, S.overflow);
      __syncwarp();
      bool took = false;
      while (leaf) {
        const uint32_t bit = __ffs(leaf) - 1;
        leaf &= leaf - 1u;
        const float4* tp = S.tris + 3ull * (rec_base + bit);
        const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
        float th, u, v;
        if (tri_test(r, p0, p1, p2, th, u, v) && th > tmin) {
          const uint32_t prim = __float_as_uint(p0.w);
          if (th < tcur || (th == tcur && best.t >= 0.0f && (cur_inst < best.inst || (cur_inst == best.inst && prim < best.prim)))) {
            best.t = th; best.u = u; best.v = v; best.prim = prim; best.inst = cur_inst;
            tcur = th;
            took = true;
          }
        }
      }
      if (__any_sync(PK_FULL, took)) {
        far2o = pk_warp_max(tcur > tmin ? ((tcur * tcur) * ddo) * 1.00001f : 0.0f);
        took_any = true;
      }
    }
    if (took_any) far2 = pk_warp_max(tcur > tmin ? ((tcur * tcur) * ddw) * 1.00001f : 0.0f);
  }
}

#ifndef SNK_PACKET_CTAS
#define SNK_PACKET_CTAS 8
#endif
__global__ void __launch_bounds__(128, SNK_PACKET_CTAS) k_primary_packet(TraceScene S, SceneDev D, FrameConsts fc, PixelPlanes P, uint32_t* ticket,
                                                                         uint32_t* next_ticket) {
  __shared__ uint32_t sm_stack[4 * PK_STACK];
  const uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31;
  uint32_t* stk = sm_stack + w * PK_STACK;
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
    if (in_range) {
      P.prev_normal[px] = P.normal[px];
      P.prev_depth[px] = P.depth[px];
    }
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
