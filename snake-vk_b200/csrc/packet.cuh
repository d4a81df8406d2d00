// packet.cuh — warp-wide BVH traversal for ray packets (device code): the camera rays of a pixel tile (gbuffer.cu) and the
// parallel shadow rays of a tile towards the directional light (restir_ns.cu).
//
// Replaces the same driver work as traverse.cuh (vkCmdTraceRaysKHR / traceRayEXT, Core/src/rendering/Raytracing.cpp:253, :260)
// for rays that are coherent by construction. Box culling is shared by the warp and only conservative; the triangle test is
// traverse.cuh's, per ray, so hits are bit-identical to the one-ray traversal and to the brute-force oracle.
#pragma once
#include "traverse.cuh"

namespace snk {

// ---- packets with a common origin: the camera rays of an 8x4 pixel tile ---------------------------------------------------
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
constexpr int PK_STACK = 128;               // shared stack entries per warp (u32); deeper entries go to the warp's global spill area
constexpr int PK_SPILL = SPILL_STACK * 32 * 2;  // u32 entries of the per-warp spill region of traverse.cuh (TraceScene::spill)
// the warp's stack: entry i < PK_STACK in shared memory, the rest (degenerate, chain-like hierarchies) in global memory
struct PkStack {
  uint32_t* sm;
  uint32_t* gl;
  __device__ __forceinline__ uint32_t& at(int i) const { return i < PK_STACK ? sm[i] : gl[i - PK_STACK]; }
};
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

// this lane's child box of a node, relative to the origin (ox, oy, oz): lane j (= lane & 7) looks at child slot c = j ^ octinv,
// so ballot bits come out in visiting order. Byte c of an 8-byte row = one PRMT over the word pair (selector nibble 0..7); the
// & 0xFF folds into the I2F.U8.
struct PkBox { float lox, loy, loz, hix, hiy, hiz; uint32_t meta; };
__device__ __forceinline__ PkBox pk_child_box(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4, uint32_t c, float ox, float oy,
                                              float oz) {
  PkBox b;
  b.meta = __byte_perm(n1.z, n1.w, c) & 0xFFu;
  const float px = __uint_as_float(n0.x) - ox, py = __uint_as_float(n0.y) - oy, pz = __uint_as_float(n0.z) - oz;
  const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
              sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
  // layouts.h: w2 = qlo_x | qlo_y, w3 = qlo_z | qhi_x, w4 = qhi_y | qhi_z
  b.lox = fmaf((float)(__byte_perm(n2.x, n2.y, c) & 0xFFu), sx, px); b.hix = fmaf((float)(__byte_perm(n3.z, n3.w, c) & 0xFFu), sx, px);
  b.loy = fmaf((float)(__byte_perm(n2.z, n2.w, c) & 0xFFu), sy, py); b.hiy = fmaf((float)(__byte_perm(n4.x, n4.y, c) & 0xFFu), sy, py);
  b.loz = fmaf((float)(__byte_perm(n3.x, n3.y, c) & 0xFFu), sz, pz); b.hiz = fmaf((float)(__byte_perm(n4.z, n4.w, c) & 0xFFu), sz, pz);
  return b;
}
// from the per-lane verdicts to the warp-uniform result: `inner` = surviving inner children by position, `leaf` = surviving
// leaf records (24-bit record mask), my_child = node index of this lane's child when it is a surviving inner one (lanes < 8)
__device__ __forceinline__ void pk_collect(const uint4 n0, const uint4 n1, uint32_t c, uint32_t meta, bool out, uint32_t lane, uint32_t& inner,
                                           uint32_t& leaf, uint32_t& rec_base, uint32_t& my_child) {
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

// One node for the whole warp (common-origin packets).
__device__ __forceinline__ void pk_node(const uint4* __restrict__ nodes, uint32_t idx, const PkSpace& s, float far2, uint32_t lane,
                                        uint32_t& inner, uint32_t& leaf, uint32_t& rec_base, uint32_t& my_child) {
  uint4 n0, n1, n2, n3, n4;
  load_node(nodes, idx, n0, n1, n2, n3, n4);
  const uint32_t c = (lane & 7u) ^ s.octinv;
  const PkBox b = pk_child_box(n0, n1, n2, n3, n4, c, s.ox, s.oy, s.oz);
  // largest signed distance of the box to this lane's plane; < 0 (beyond the slack): the box is outside the pyramid
  const float m = (fmaxf(s.nx * b.lox, s.nx * b.hix) + fmaxf(s.ny * b.loy, s.ny * b.hiy)) + fmaxf(s.nz * b.loz, s.nz * b.hiz);
  const float ext = ((fabsf(b.lox) + fabsf(b.hix)) + (fabsf(b.loy) + fabsf(b.hiy))) + (fabsf(b.loz) + fabsf(b.hiz));
  bool out = m < -(1e-5f * ext + s.eps_abs);
  // squared distance origin -> box against the farthest point any lane still cares about
  const float gx = fmaxf(fmaxf(b.lox, -b.hix), 0.0f), gy = fmaxf(fmaxf(b.loy, -b.hiy), 0.0f), gz = fmaxf(fmaxf(b.loz, -b.hiz), 0.0f);
  const float g2 = (gx * gx + gy * gy) + gz * gz;
  out = out || g2 * 0.9999f > far2 || b.meta == 0u;
  pk_collect(n0, n1, c, b.meta, out, lane, inner, leaf, rec_base, my_child);
}

// parallel push of the surviving inner children: position 7 is visited first (trav_step_impl opens the highest bit first), so
// position p lands at stack offset rank(p) counted from below
__device__ __forceinline__ void pk_push_inner(const PkStack& stk, int& sp, uint32_t inner, uint32_t my_child, uint32_t lane, uint32_t* overflow) {
  const int n = __popc(inner);
  if (sp + n > PK_STACK + PK_SPILL) { if (lane == 0) atomicAdd(overflow, 1u); return; }
  if (lane < 8u && ((inner >> lane) & 1u)) stk.at(sp + __popc(inner & ~(0xFFFFFFFFu << lane))) = my_child;
  sp += n;
}

__device__ __forceinline__ void pk_trace(const TraceScene& S, float ox, float oy, float oz, float dx, float dy, float dz, float tmin, float tmax,
                                         const PkStack& stk, uint32_t lane, HitRec& best) {
  best.t = -1.0f; best.u = 0.f; best.v = 0.f; best.prim = 0xFFFFFFFFu; best.inst = 0xFFFFFFFFu;
  float tcur = tmax;
  if (S.tlas_node_count == 0u) return;
  PkSpace W;
  pk_setup_space(W, ox, oy, oz, dx, dy, dz, lane);
  const float ddw = (dx * dx + dy * dy) + dz * dz;
  float far2 = pk_warp_max(tcur > tmin ? (tcur * tcur) * ddw : 0.0f);
  if (!(far2 > 0.0f)) return;
  int sp = 0;
  stk.at(sp++) = 0u;  // TLAS root (every lane writes the same word)
  __syncwarp();
  while (sp > 0) {
    const uint32_t e = stk.at(--sp);
    __syncwarp();  // everybody has read the entry before anybody overwrites the slot
    if (!(e & PK_INST)) {
      uint32_t inner, leaf, rec_base, my_child;
      pk_node(S.tlas_nodes, e, W, far2, lane, inner, leaf, rec_base, my_child);
      pk_push_inner(stk, sp, inner, my_child, lane, S.overflow);
      // the instances of this node's leaf slots go on top: they are entered before the inner children are opened
      const int nl = __popc(leaf);
      if (sp + nl > PK_STACK + PK_SPILL) { if (lane == 0) atomicAdd(S.overflow, 1u); }
      else {
        if (lane < 24u && ((leaf >> lane) & 1u)) stk.at(sp + __popc(leaf & ~(0xFFFFFFFFu << lane))) = PK_INST | (rec_base + lane);
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
    stk.at(sp++) = meta.x;
    __syncwarp();
    while (sp > base) {
      const uint32_t node = stk.at(--sp);
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

// ---- packets of parallel rays: the shadow rays of a pixel tile towards the directional light ------------------------------
// Up to 32 rays with the SAME direction d and different origins (the primary-hit points of an 8x4 tile), any-hit. The culling
// volume is the box around the origins in a frame (a, b, w = d / |d|), extruded along +w: four side planes (one per lane group
// k) and the plane behind the hindmost origin; only rays that still look for a hit span it. Everything else as above.
struct PkOrtho {
  float ox, oy, oz;     // reference origin (first ray of the packet that is still looking)
  float nx, ny, nz, h;  // this lane's side plane: a box is outside when max over the box of n.(x - o) < h
  float wx, wy, wz, wmin;
  float eps_abs;
  uint32_t octinv;
};
__device__ __forceinline__ int pk_order_int(float x) { const int i = __float_as_int(x); return i ^ ((i >> 31) & 0x7FFFFFFF); }
__device__ __forceinline__ float pk_order_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7FFFFFFF)); }
__device__ __forceinline__ float pk_warp_minf(float v) { return pk_order_float(__reduce_min_sync(PK_FULL, pk_order_int(v))); }
__device__ __forceinline__ float pk_warp_maxf(float v) { return pk_order_float(__reduce_max_sync(PK_FULL, pk_order_int(v))); }

// `want`: this lane's ray still looks for a hit (at least one lane must). Origins and the direction in the current space.
__device__ __forceinline__ void pk_setup_ortho(PkOrtho& s, bool want, float ox, float oy, float oz, float dx, float dy, float dz, uint32_t lane) {
  const int first = __ffs(__ballot_sync(PK_FULL, want)) - 1;
  s.ox = __shfl_sync(PK_FULL, ox, first); s.oy = __shfl_sync(PK_FULL, oy, first); s.oz = __shfl_sync(PK_FULL, oz, first);
  // every lane holds the same direction; the frame is built from the first lane's copy so that all lanes agree to the bit
  const float wx0 = __shfl_sync(PK_FULL, dx, first), wy0 = __shfl_sync(PK_FULL, dy, first), wz0 = __shfl_sync(PK_FULL, dz, first);
  const float inv = rsqrtf((wx0 * wx0 + wy0 * wy0) + wz0 * wz0);
  const float wx = wx0 * inv, wy = wy0 * inv, wz = wz0 * inv;
  // a: unit vector across the smallest component of w, b = w x a
  float ax, ay, az;
  if (fabsf(wx) <= fabsf(wy) && fabsf(wx) <= fabsf(wz)) { ax = 0.0f; ay = -wz; az = wy; }
  else if (fabsf(wy) <= fabsf(wz)) { ax = wz; ay = 0.0f; az = -wx; }
  else { ax = -wy; ay = wx; az = 0.0f; }
  const float ia = rsqrtf((ax * ax + ay * ay) + az * az);
  ax *= ia; ay *= ia; az *= ia;
  const float bx = wy * az - wz * ay, by = wz * ax - wx * az, bz = wx * ay - wy * ax;
  const float ex = ox - s.ox, ey = oy - s.oy, ez = oz - s.oz;
  const float ua = (ax * ex + ay * ey) + az * ez, ub = (bx * ex + by * ey) + bz * ez, uw = (wx * ex + wy * ey) + wz * ez;
  const float big = 3.0e38f;
  const float amin = pk_warp_minf(want ? ua : big), amax = pk_warp_maxf(want ? ua : -big);
  const float bmin = pk_warp_minf(want ? ub : big), bmax = pk_warp_maxf(want ? ub : -big);
  s.wmin = pk_warp_minf(want ? uw : big);
  const uint32_t k = lane >> 3;
  const float sg = (k & 1u) ? -1.0f : 1.0f;
  s.nx = sg * (k < 2u ? ax : bx); s.ny = sg * (k < 2u ? ay : by); s.nz = sg * (k < 2u ? az : bz);
  s.h = k == 0u ? amin : (k == 1u ? -amax : (k == 2u ? bmin : -bmax));
  s.wx = wx; s.wy = wy; s.wz = wz;
  const float spread = (fmaxf(fabsf(amin), fabsf(amax)) + fmaxf(fabsf(bmin), fabsf(bmax))) + fabsf(s.wmin);
  s.eps_abs = 2e-6f * ((fabsf(s.ox) + fabsf(s.oy)) + fabsf(s.oz)) + 1e-5f * spread;
  const uint32_t oct = (wx < 0.f ? 1u : 0u) | (wy < 0.f ? 2u : 0u) | (wz < 0.f ? 4u : 0u);
  s.octinv = 7u - oct;
}

__device__ __forceinline__ void pk_node_ortho(const uint4* __restrict__ nodes, uint32_t idx, const PkOrtho& s, uint32_t lane, uint32_t& inner,
                                              uint32_t& leaf, uint32_t& rec_base, uint32_t& my_child) {
  uint4 n0, n1, n2, n3, n4;
  load_node(nodes, idx, n0, n1, n2, n3, n4);
  const uint32_t c = (lane & 7u) ^ s.octinv;
  const PkBox b = pk_child_box(n0, n1, n2, n3, n4, c, s.ox, s.oy, s.oz);
  const float m = (fmaxf(s.nx * b.lox, s.nx * b.hix) + fmaxf(s.ny * b.loy, s.ny * b.hiy)) + fmaxf(s.nz * b.loz, s.nz * b.hiz);
  const float mw = (fmaxf(s.wx * b.lox, s.wx * b.hix) + fmaxf(s.wy * b.loy, s.wy * b.hiy)) + fmaxf(s.wz * b.loz, s.wz * b.hiz);
  const float ext = ((fabsf(b.lox) + fabsf(b.hix)) + (fabsf(b.loy) + fabsf(b.hiy))) + (fabsf(b.loz) + fabsf(b.hiz));
  const float eps = 1e-5f * ext + s.eps_abs;
  const bool out = m < s.h - eps || mw < s.wmin - eps || b.meta == 0u;
  pk_collect(n0, n1, c, b.meta, out, lane, inner, leaf, rec_base, my_child);
}

// any-hit for a packet of parallel rays; returns true when this lane's ray (tmin < t < tmax) hits something. Lanes without a
// ray pass tmax <= tmin. Every lane of the warp must call it.
__device__ __forceinline__ bool pk_trace_parallel_any(const TraceScene& S, float ox, float oy, float oz, float tmin, float dx, float dy, float dz, float tmax,
                                                      const PkStack& stk, uint32_t lane) {
  bool want = tmax > tmin && S.tlas_node_count != 0u;  // still looking for a hit
  if (!__any_sync(PK_FULL, want)) return false;
  bool hit = false;
  PkOrtho W;
  pk_setup_ortho(W, want, ox, oy, oz, dx, dy, dz, lane);
  int sp = 0;
  stk.at(sp++) = 0u;
  __syncwarp();
  while (sp > 0) {
    const uint32_t e = stk.at(--sp);
    __syncwarp();
    if (!(e & PK_INST)) {
      uint32_t inner, leaf, rec_base, my_child;
      pk_node_ortho(S.tlas_nodes, e, W, lane, inner, leaf, rec_base, my_child);
      pk_push_inner(stk, sp, inner, my_child, lane, S.overflow);
      const int nl = __popc(leaf);
      if (sp + nl > PK_STACK + PK_SPILL) { if (lane == 0) atomicAdd(S.overflow, 1u); }
      else {
        if (lane < 24u && ((leaf >> lane) & 1u)) stk.at(sp + __popc(leaf & ~(0xFFFFFFFFu << lane))) = PK_INST | (rec_base + lane);
        sp += nl;
      }
      __syncwarp();
      continue;
    }
    const uint32_t slot = __ldg(S.tlas_leaf_ids + (e & ~PK_INST));
    const bool may = want && (!S.inst_sphere || ray_may_hit_sphere(__ldg(S.inst_sphere + slot), ox, oy, oz, dx, dy, dz, tmax));
    if (!__any_sync(PK_FULL, may)) continue;
    const float4* M = reinterpret_cast<const float4*>(S.inst + slot);
    const float4 r0 = __ldg(M), r1 = __ldg(M + 1), r2 = __ldg(M + 2);
    const uint4 meta = __ldg(reinterpret_cast<const uint4*>(M + 3));
    RaySpace r;
    r.ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, ox), __fmul_rn(r0.y, oy)), __fmul_rn(r0.z, oz)), r0.w);
    r.oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, ox), __fmul_rn(r1.y, oy)), __fmul_rn(r1.z, oz)), r1.w);
    r.oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, ox), __fmul_rn(r2.y, oy)), __fmul_rn(r2.z, oz)), r2.w);
    r.dx = __fadd_rn(__fadd_rn(__fmul_rn(r0.x, dx), __fmul_rn(r0.y, dy)), __fmul_rn(r0.z, dz));
    r.dy = __fadd_rn(__fadd_rn(__fmul_rn(r1.x, dx), __fmul_rn(r1.y, dy)), __fmul_rn(r1.z, dz));
    r.dz = __fadd_rn(__fadd_rn(__fmul_rn(r2.x, dx), __fmul_rn(r2.y, dy)), __fmul_rn(r2.z, dz));
    setup_tri_space(r);
    PkOrtho B;
    pk_setup_ortho(B, may, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, lane);  // only the rays that can reach this instance span the volume
    const int base = sp;
    stk.at(sp++) = meta.x;
    __syncwarp();
    bool live = true;  // some lane that may hit this instance is still looking
    while (sp > base) {
      const uint32_t node = stk.at(--sp);
      __syncwarp();
      if (!live) continue;  // drain the instance's entries
      uint32_t inner, leaf, rec_base, my_child;
      pk_node_ortho(S.blas_nodes, node, B, lane, inner, leaf, rec_base, my_child);
      pk_push_inner(stk, sp, inner, my_child, lane, S.overflow);
      __syncwarp();
      bool took = false;
      while (leaf) {
        const uint32_t bit = __ffs(leaf) - 1;
        leaf &= leaf - 1u;
        const float4* tp = S.tris + 3ull * (rec_base + bit);
        const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
        float th, u, v;
        if (want && tri_test(r, p0, p1, p2, th, u, v) && th > tmin && th < tmax) { hit = true; want = false; took = true; }
      }
      if (__any_sync(PK_FULL, took)) live = __any_sync(PK_FULL, may && want);
    }
    if (!__any_sync(PK_FULL, want)) break;
  }
  return hit;
}

}  // namespace snk
