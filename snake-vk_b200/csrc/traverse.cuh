// traverse.cuh — two-level BVH8 traversal + watertight ray/triangle test (device code).
//
// Replaces the driver side of every traceRayEXT the reference issues
// (RayGen.rgen:101-121 ShadowTest, :164-176 / :195-207 closest-hit rays,
// RayClosestHit.rchit:64-84) and vkCmdTraceRaysKHR's traversal
// (Core/src/rendering/Raytracing.cpp:253, :260). Semantics pinned to the Vulkan
// ray-tracing contract the shaders rely on: closest hit with tmin < t < tmax,
// watertight shared edges, no culling, barycentrics = weights of v1, v2;
// equal-t ties resolve to the lowest (instance, primitive).
//
// The triangle test is written with explicit round-to-nearest intrinsics (never
// contracted to FMA) so that its results are bit-identical to the CPU oracle's
// restatement; the box tests use FMA and a relative pad and only need to be
// conservative.
#pragma once
#include <cuda_runtime.h>

#include "layouts.h"

namespace snk {

struct TraceScene {
  const uint4* tlas_nodes;        // 5 x uint4 per node, root at index 0
  const uint32_t* tlas_leaf_ids;  // instance slot per TLAS leaf record
  const uint4* blas_nodes;        // node pool of all BLASes
  const float4* tris;             // 3 x float4 per triangle record
  const SnkInstRec* inst;         // indexed by instance slot
  const float4* inst_sphere;      // world-space bounding sphere per instance slot (centre, r^2; r^2 < 0: none) or nullptr
  uint32_t* overflow;             // device counter: traversal stack overflows (must stay 0)
  uint32_t tlas_node_count;       // 0 = nothing to hit
  uint2* spill;                   // global overflow area of the short stacks (persistent kernels), SPILL_STACK x 32 per warp
};

// queue traversal (trace_q.cu): a ray reference = segment << REF_SLOT_BITS | slot; a segment names the ray plane the
// slot indexes and where that ray's result goes (closest: hit + inst planes; any: visibility plane)
constexpr uint32_t REF_SLOT_BITS = 28u, REF_SLOT_MASK = (1u << REF_SLOT_BITS) - 1u;
constexpr int TRACE_MAX_SEGS = 8;
struct TraceSeg {
  const float4* rays;  // 2 x float4 per slot (o,tmin | d,tmax)
  float4* hit;         // closest: t,u,v,prim
  uint32_t* inst;      // closest: instance custom index
  float* vis;          // any: 1 = unoccluded, 0 = occluded
  uint32_t any, pad;
};
struct TraceSegs { TraceSeg s[TRACE_MAX_SEGS]; };
struct RefQueue {
  uint32_t* refs;    // compact list of ray references
  uint32_t* count;   // device counter the producers append through
  uint32_t* ticket;  // device counter the traversal kernel draws from
  uint32_t cap;      // capacity in references
};

struct HitRec {
  float t, u, v;
  uint32_t prim, inst;
};

// Traversal stack entries are 8 bytes; the top SHORT_STACK entries of every thread live in SHARED memory,
// interleaved by thread (entry e of thread t at stk[e * STRIDE], STRIDE = CTA size) so a warp's pushes / pops
// are conflict-free 64-bit accesses; deeper entries spill to a per-warp global area. Deliberately not a
// local-memory array: an earlier version saw lanes that left the divergent loop early clobber the uniform
// register holding the local-frame base (compute-sanitizer, k_primary on C2).
constexpr float BOX_PAD = 1.000004f;

// Instance culling: a TLAS leaf only says that the ray crosses the instance's world-space BOX; before paying for the entry
// (ray transform, box + Woop constants, BLAS root test) the ray is tested against the instance's bounding SPHERE. For compact
// objects the box is mostly empty (a sphere fills 52 % of its box: half of all entries into the icospheres of C2 were misses).
// Conservative under float32: the radius is inflated on the host, the squared distances get a relative slack here, and a ray
// is only rejected when it clearly misses, starts outside and points away, or enters beyond the current closest hit.
__device__ __forceinline__ bool ray_may_hit_sphere(float4 sph, float ox, float oy, float oz, float dx, float dy, float dz, float tcur) {
  if (!(sph.w >= 0.0f)) return true;
  const float cx = sph.x - ox, cy = sph.y - oy, cz = sph.z - oz;
  const float oc2 = cx * cx + cy * cy + cz * cz, dd = dx * dx + dy * dy + dz * dz, b = cx * dx + cy * dy + cz * dz;
  const float cc = oc2 * 0.999996f - sph.w;          // > 0: the origin is outside
  if (!(cc > 0.0f)) return true;
  const float slack = 2e-6f * dd * oc2;
  const float disc = b * b - dd * cc + slack;        // upper bound of the discriminant
  if (disc < 0.0f) return false;                     // the line misses
  if (b < 0.0f && b * b > slack) return false;       // outside and pointing away
  const float e = b - tcur * dd;                     // > 0: the closest approach lies beyond tcur
  if (e > 0.0f && e * e > disc) return false;        // the entry point lies beyond the current closest hit / tmax
  return true;
}

struct RaySpace {
  float ox, oy, oz;
  float dx, dy, dz;
  float idx, idy, idz;  // 1 / clamped direction
  float Sx, Sy, Sz;     // Woop shear constants
  int kx, ky, kz;
  uint32_t octinv;
};

__device__ __forceinline__ float sel3(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

__device__ __forceinline__ void setup_ray_space(RaySpace& r, float ox, float oy, float oz, float dx, float dy, float dz) {
  r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
  const float cx = fabsf(dx) > 1e-20f ? dx : copysignf(1e-20f, dx);
  const float cy = fabsf(dy) > 1e-20f ? dy : copysignf(1e-20f, dy);
  const float cz = fabsf(dz) > 1e-20f ? dz : copysignf(1e-20f, dz);
  r.idx = 1.0f / cx; r.idy = 1.0f / cy; r.idz = 1.0f / cz;
  const uint32_t oct = (r.idx < 0.f ? 1u : 0u) | (r.idy < 0.f ? 2u : 0u) | (r.idz < 0.f ? 4u : 0u);
  r.octinv = 7u - oct;
  // Woop et al. 2013: dominant axis + shear (bit-identical to oracle/snk_scene.h ray_pre)
  const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
  int kz = 0;
  if (ay > ax) kz = 1;
  if (az > (kz == 0 ? ax : ay)) kz = 2;
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  const float dkz = sel3(dx, dy, dz, kz);
  if (dkz < 0.0f) { const int t = kx; kx = ky; ky = t; }
  r.kx = kx; r.ky = ky; r.kz = kz;
  r.Sx = __fdiv_rn(sel3(dx, dy, dz, kx), dkz);
  r.Sy = __fdiv_rn(sel3(dx, dy, dz, ky), dkz);
  r.Sz = __fdiv_rn(1.0f, dkz);
}

// returns true when the line hits the triangle; (t,u,v) valid then. No range test.
// DERIVE_K: kx / ky are recomputed from kz and the sign of the dominant direction component (the rule of setup_tri_space) instead
// of being read from r — two registers less across the traversal loop for ~6 instructions per test; the same values.
template <bool DERIVE_K = false>
__device__ __forceinline__ bool tri_test(const RaySpace& r0, float4 p0, float4 p1, float4 p2, float& t, float& u, float& v) {
  struct { float ox, oy, oz, Sx, Sy, Sz; int kx, ky, kz; } r = {r0.ox, r0.oy, r0.oz, r0.Sx, r0.Sy, r0.Sz, r0.kx, r0.ky, r0.kz};
  if (DERIVE_K) {
    int kx = r.kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (sel3(r0.dx, r0.dy, r0.dz, r.kz) < 0.0f) { const int k = kx; kx = ky; ky = k; }
    r.kx = kx; r.ky = ky;
  }
  const float Ax0 = __fsub_rn(p0.x, r.ox), Ay0 = __fsub_rn(p0.y, r.oy), Az0 = __fsub_rn(p0.z, r.oz);
  const float Bx0 = __fsub_rn(p1.x, r.ox), By0 = __fsub_rn(p1.y, r.oy), Bz0 = __fsub_rn(p1.z, r.oz);
  const float Cx0 = __fsub_rn(p2.x, r.ox), Cy0 = __fsub_rn(p2.y, r.oy), Cz0 = __fsub_rn(p2.z, r.oz);
  const float Akz = sel3(Ax0, Ay0, Az0, r.kz), Bkz = sel3(Bx0, By0, Bz0, r.kz), Ckz = sel3(Cx0, Cy0, Cz0, r.kz);
  const float Ax = __fsub_rn(sel3(Ax0, Ay0, Az0, r.kx), __fmul_rn(r.Sx, Akz)), Ay = __fsub_rn(sel3(Ax0, Ay0, Az0, r.ky), __fmul_rn(r.Sy, Akz));
  const float Bx = __fsub_rn(sel3(Bx0, By0, Bz0, r.kx), __fmul_rn(r.Sx, Bkz)), By = __fsub_rn(sel3(Bx0, By0, Bz0, r.ky), __fmul_rn(r.Sy, Bkz));
  const float Cx = __fsub_rn(sel3(Cx0, Cy0, Cz0, r.kx), __fmul_rn(r.Sx, Ckz)), Cy = __fsub_rn(sel3(Cx0, Cy0, Cz0, r.ky), __fmul_rn(r.Sy, Ckz));
  float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
  float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
  float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    const double CxBy = __dmul_rn((double)Cx, (double)By), CyBx = __dmul_rn((double)Cy, (double)Bx);
    U = (float)__dsub_rn(CxBy, CyBx);
    const double AxCy = __dmul_rn((double)Ax, (double)Cy), AyCx = __dmul_rn((double)Ay, (double)Cx);
    V = (float)__dsub_rn(AxCy, AyCx);
    const double BxAy = __dmul_rn((double)Bx, (double)Ay), ByAx = __dmul_rn((double)By, (double)Ax);
    W = (float)__dsub_rn(BxAy, ByAx);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  const float det = __fadd_rn(__fadd_rn(U, V), W);
  if (det == 0.0f) return false;
  const float Az = __fmul_rn(r.Sz, Akz), Bz = __fmul_rn(r.Sz, Bkz), Cz = __fmul_rn(r.Sz, Ckz);
  const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
  const float rcp = __fdiv_rn(1.0f, det);
  t = __fmul_rn(T, rcp);
  u = __fmul_rn(V, rcp);
  v = __fmul_rn(W, rcp);
  return true;
}

__device__ __forceinline__ uint32_t byte_of(uint32_t lo_word, uint32_t hi_word, int s) {
  return ((s < 4 ? lo_word : hi_word) >> ((s & 3) * 8)) & 0xFFu;
}

// replicate the sign bit of each byte over the byte (0x80 -> 0xFF)
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
  uint32_t v;
  asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(v) : "r"(x));
  return v;
}

// float 1 + q * 2^-15 of byte `k` of `w` (q = 0..255): the byte is dropped into mantissa bits 8..15 of 1.0f with ONE
// PRMT on the ALU pipe. The 48 I2F.U8 per node this replaces run on the quarter-rate XU pipe, which ncu showed as the
// busiest pipe of the traversal kernel (58 % of peak, profiles/r01q_prof_traceq.txt).
#ifndef SNK_NODE_PRMT
#define SNK_NODE_PRMT 0  // measured (r01r): no gain over I2F.U8 — the kernel is bound by issue slots, not by the XU pipe — and 6 more instructions per node
#endif
template <int K>
__device__ __forceinline__ float byte_to_unit_float(uint32_t w) {
  uint32_t v;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(w), "r"(0x3F800000u), "n"(0x7604 | (K << 4)));
  return __uint_as_float(v);
}

// intersect the 8 quantised child boxes of one node; returns the hit mask
// (bits 24..31: inner children in octant order, bits 0..23: leaf records). Branch-free: the per-child bit
// position / bit pattern are derived four children at a time from the meta bytes (Ylitie et al. 2017), an
// empty slot has meta 0 and contributes no bits whatever its box test says.
// Plane distance: t = q * a + b with a = step / d, b = (origin - o) / d is evaluated as f(q) * A + B with
// f(q) = 1 + q * 2^-15, A = 2^15 a, B = b - A. B carries one extra rounding of at most 2^-24 |B| <= 2^-24 |b| + 2^-9 |a|;
// the second term (1/512 of a quantisation step) is covered by moving near planes down and far planes up by
// 2^-22 |A| = 2^-7 |a| per axis, the first by the relative BOX_PAD as before. The tests only have to be conservative.
#if SNK_NODE_PRMT
__device__ __forceinline__ uint32_t intersect_node(const RaySpace& r, const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3,
                                                   const uint4 n4, float tmin, float tmax) {
  const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
  // step * 2^15: exponent bytes + 15
  const float sx = __uint_as_float(((n0.w & 0xFFu) + 15u) << 23), sy = __uint_as_float((((n0.w >> 8) & 0xFFu) + 15u) << 23),
              sz = __uint_as_float((((n0.w >> 16) & 0xFFu) + 15u) << 23);
  const float ax = sx * r.idx, ay = sy * r.idy, az = sz * r.idz;
  const float bx = fmaf(px - r.ox, r.idx, -ax), by = fmaf(py - r.oy, r.idy, -ay), bz = fmaf(pz - r.oz, r.idz, -az);
  const float ex = fabsf(ax) * 2.384185791015625e-07f, ey = fabsf(ay) * 2.384185791015625e-07f, ez = fabsf(az) * 2.384185791015625e-07f;
  const float bxn = bx - ex, bxf = bx + ex, byn = by - ey, byf = by + ey, bzn = bz - ez, bzf = bz + ez;
  // near / far plane words per axis
  const bool nx = r.idx < 0.f, ny = r.idy < 0.f, nz = r.idz < 0.f;
  const uint32_t xn[2] = {nx ? n3.z : n2.x, nx ? n3.w : n2.y}, xf[2] = {nx ? n2.x : n3.z, nx ? n2.y : n3.w};
  const uint32_t yn[2] = {ny ? n4.x : n2.z, ny ? n4.y : n2.w}, yf[2] = {ny ? n2.z : n4.x, ny ? n2.w : n4.y};
  const uint32_t zn[2] = {nz ? n4.z : n3.x, nz ? n4.w : n3.y}, zf[2] = {nz ? n3.x : n4.z, nz ? n3.y : n4.w};
  const uint32_t octinv4 = r.octinv * 0x01010101u;
  uint32_t hitmask = 0;
#define SNK_CHILD(J)                                                                                                          \
  {                                                                                                                           \
    const float tx0 = fmaf(byte_to_unit_float<J>(xn[h]), ax, bxn), tx1 = fmaf(byte_to_unit_float<J>(xf[h]), ax, bxf);           \
    const float ty0 = fmaf(byte_to_unit_float<J>(yn[h]), ay, byn), ty1 = fmaf(byte_to_unit_float<J>(yf[h]), ay, byf);           \
    const float tz0 = fmaf(byte_to_unit_float<J>(zn[h]), az, bzn), tz1 = fmaf(byte_to_unit_float<J>(zf[h]), az, bzf);           \
    const float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tmin));                                                                \
    const float tf = fminf(fminf(tx1, ty1), fminf(tz1, tmax));                                                                \
    const uint32_t bits = ((child_bits4 >> (8 * J)) & 0xFFu) << ((bit_index4 >> (8 * J)) & 0xFFu);                            \
    hitmask |= (tn <= tf * BOX_PAD) ? bits : 0u;                                                                              \
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t meta4 = h ? n1.w : n1.z;
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    SNK_CHILD(0) SNK_CHILD(1) SNK_CHILD(2) SNK_CHILD(3)
  }
#undef SNK_CHILD
  return hitmask;
}
#else
__device__ __forceinline__ uint32_t intersect_node(const RaySpace& r, const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3,
                                                   const uint4 n4, float tmin, float tmax) {
  const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
  const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
              sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
  const float ax = sx * r.idx, ay = sy * r.idy, az = sz * r.idz;
  const float bx = (px - r.ox) * r.idx, by = (py - r.oy) * r.idy, bz = (pz - r.oz) * r.idz;
  // near / far plane words per axis
  const bool nx = r.idx < 0.f, ny = r.idy < 0.f, nz = r.idz < 0.f;
  const uint32_t xn[2] = {nx ? n3.z : n2.x, nx ? n3.w : n2.y}, xf[2] = {nx ? n2.x : n3.z, nx ? n2.y : n3.w};
  const uint32_t yn[2] = {ny ? n4.x : n2.z, ny ? n4.y : n2.w}, yf[2] = {ny ? n2.z : n4.x, ny ? n2.w : n4.y};
  const uint32_t zn[2] = {nz ? n4.z : n3.x, nz ? n4.w : n3.y}, zf[2] = {nz ? n3.x : n4.z, nz ? n3.y : n4.w};
  const uint32_t octinv4 = r.octinv * 0x01010101u;
  uint32_t hitmask = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t meta4 = h ? n1.w : n1.z;
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int sh = 8 * j;
      const float tx0 = fmaf((float)((xn[h] >> sh) & 0xFFu), ax, bx), tx1 = fmaf((float)((xf[h] >> sh) & 0xFFu), ax, bx);
      const float ty0 = fmaf((float)((yn[h] >> sh) & 0xFFu), ay, by), ty1 = fmaf((float)((yf[h] >> sh) & 0xFFu), ay, by);
      const float tz0 = fmaf((float)((zn[h] >> sh) & 0xFFu), az, bz), tz1 = fmaf((float)((zf[h] >> sh) & 0xFFu), az, bz);
      const float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tmin));
      const float tf = fminf(fminf(tx1, ty1), fminf(tz1, tmax));
      const uint32_t bits = ((child_bits4 >> sh) & 0xFFu) << ((bit_index4 >> sh) & 0xFFu);
      hitmask |= (tn <= tf * BOX_PAD) ? bits : 0u;
    }
  }
  return hitmask;
}
#endif

__device__ __forceinline__ void load_node(const uint4* __restrict__ nodes, uint32_t idx, uint4& n0, uint4& n1, uint4& n2, uint4& n3,
                                          uint4& n4) {
  const uint4* p = nodes + 5ull * idx;
  n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3); n4 = __ldg(p + 4);
}

// ---- state-machine form (persistent wavefront kernels) ---------------------------------------------------
// Same traversal as trace_ray, cut into three warp-synchronous phases per iteration so that (a) lanes
// reconverge every iteration, (b) each phase runs with as many lanes as possible and (c) a warp can swap
// finished rays for new ones between iterations (dynamic fetch):
//   node phase  lanes with an inner-node hit pending open the nearest child (5 x 128-bit node load, 8 box tests)
//   leaf phase  lanes with leaf records pending take ONE record: a triangle test in a BLAS, an instance entry in the TLAS
//   pop phase   lanes whose group is exhausted pop the stack (a sentinel pops back to world space)
// Stack: SHORT_STACK entries per lane in shared memory (interleaved by thread), deeper entries in a per-warp
// global spill area (interleaved by lane) that stays L1/L2 resident.
#ifndef SNK_SHORT_STACK
#define SNK_SHORT_STACK 8
#endif
constexpr int SHORT_STACK = SNK_SHORT_STACK;
constexpr int SPILL_STACK = 40;

struct StackRef {
  uint2* sm;  // this thread's column of the shared short stack (stride = CTA threads)
  uint2* gl;  // this lane's column of the warp's spill area (stride = 32)
  float4* world = nullptr;  // optional: this thread's cached world-space box constants (1/d.xyz, octant) for leaving an instance
};

struct Trav {
  float ox, oy, oz, tmin, dx, dy, dz;  // world-space ray
  float tcur;                          // tmax, shrinking to the closest hit
  RaySpace r;                          // current space (world in the TLAS, object in a BLAS)
  uint2 G, T;                          // pending node group / leaf-record group
  int sp, blas_sp;                     // blas_sp: stack height at instance entry (leaving = popping down to it)
  uint32_t cur_inst;
  bool in_blas, tri_ready;             // tri_ready: the Woop constants of r are valid for the current space
  HitRec best;
};

template <int STRIDE>
__device__ __forceinline__ void stk_push(const StackRef& s, int& sp, uint2 v, uint32_t* overflow) {
  if (sp < SHORT_STACK) s.sm[sp * STRIDE] = v;
  else if (sp < SHORT_STACK + SPILL_STACK) s.gl[(sp - SHORT_STACK) * 32] = v;
  else atomicAdd(overflow, 1u);
  sp += sp < SHORT_STACK + SPILL_STACK ? 1 : 0;
}
template <int STRIDE>
__device__ __forceinline__ uint2 stk_pop(const StackRef& s, int& sp) {
  --sp;
  return sp < SHORT_STACK ? s.sm[sp * STRIDE] : s.gl[(sp - SHORT_STACK) * 32];
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// box-test part of the ray space: origin, reciprocal direction, octant. The reciprocal only feeds the
// (padded, conservative) box tests, so the 1-ulp approximation is enough.
__device__ __forceinline__ void setup_box_space(RaySpace& r, float ox, float oy, float oz, float dx, float dy, float dz) {
  r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
  const float cx = fabsf(dx) > 1e-20f ? dx : copysignf(1e-20f, dx);
  const float cy = fabsf(dy) > 1e-20f ? dy : copysignf(1e-20f, dy);
  const float cz = fabsf(dz) > 1e-20f ? dz : copysignf(1e-20f, dz);
  r.idx = rcp_approx(cx); r.idy = rcp_approx(cy); r.idz = rcp_approx(cz);
  const uint32_t oct = (r.idx < 0.f ? 1u : 0u) | (r.idy < 0.f ? 2u : 0u) | (r.idz < 0.f ? 4u : 0u);
  r.octinv = 7u - oct;
}
// triangle-test part (Woop et al. 2013 dominant axis + shear), bit-identical to oracle/snk_scene.h ray_pre
__device__ __forceinline__ void setup_tri_space(RaySpace& r) {
  const float dx = r.dx, dy = r.dy, dz = r.dz;
  const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
  int kz = 0;
  if (ay > ax) kz = 1;
  if (az > (kz == 0 ? ax : ay)) kz = 2;
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  const float dkz = sel3(dx, dy, dz, kz);
  if (dkz < 0.0f) { const int t = kx; kx = ky; ky = t; }
  r.kx = kx; r.ky = ky; r.kz = kz;
  r.Sx = __fdiv_rn(sel3(dx, dy, dz, kx), dkz);
  r.Sy = __fdiv_rn(sel3(dx, dy, dz, ky), dkz);
  r.Sz = __fdiv_rn(1.0f, dkz);
}

__device__ __forceinline__ void trav_begin(Trav& t, float4 a, float4 b) {
  t.ox = a.x; t.oy = a.y; t.oz = a.z; t.tmin = a.w; t.dx = b.x; t.dy = b.y; t.dz = b.z; t.tcur = b.w;
  setup_box_space(t.r, a.x, a.y, a.z, b.x, b.y, b.z);
  t.G = make_uint2(0u, 0x80000000u);
  t.T = make_uint2(0u, 0u);
  t.sp = 0; t.blas_sp = 0; t.cur_inst = 0; t.in_blas = false; t.tri_ready = false;
  t.best.t = -1.0f; t.best.u = 0.f; t.best.v = 0.f; t.best.prim = 0xFFFFFFFFu; t.best.inst = 0xFFFFFFFFu;
}

#ifndef SNK_TRI_REPS
#define SNK_TRI_REPS 1  // triangle records a lane may test per iteration (2: -5 %, 3: -8 %, profiles/r02r_variants_tri_reps_world_cache.txt)
#endif
#ifndef SNK_TRI_MIN
#define SNK_TRI_MIN 0  // triangle phase runs when at least this many lanes have a record pending (or nobody opened a node)
#endif

// One warp-synchronous iteration for the lanes in `warp` (all of them must call it together):
//   1. pick the node to open: TLAS lanes with an instance record pending enter the instance (transform the
//      ray, node = BLAS root); lanes with an inner hit pending take the nearest child of their group
//   2. common node phase: 5 x 128-bit load + 8 quantised box tests -> new inner-hit group G and leaf group T
//   3. triangle phase: BLAS lanes with leaf records pending test ONE triangle (watertight, bit-exact with the oracle)
//   4. pop phase: lanes with nothing pending pop the stack; popping below the instance-entry height returns
//      to world space; an empty stack finishes the ray
// Returns true when this lane's ray is finished (stack empty, or first hit for ANY).
// COLD (with CACHED_WORLD): the state a ray touches a few times in its life — the world-space origin and direction (needed again only
// when it leaves an instance) and u, v, primitive, instance of its closest hit (written when a hit is accepted, read at the end and on
// an exact distance tie) — lives in the thread's shared-memory column stk.world[1 .. 3] instead of ten registers: at 64 registers the
// compiler otherwise spills nine values around the node test of EVERY iteration (9 STL + 9 LDL per node, profiles/r04r_k_trace_q.sass).
#ifndef SNK_EARLY_NODE_LOAD
#define SNK_EARLY_NODE_LOAD 0  // 1: the node is requested inside the pick branch, before the rest of the group is pushed — 20 more live registers there bring
                                // the spills back: C2 4.41 -> 4.50 ms, reference mode 33.2 -> 35.8 ms (profiles/r07n_variants_load_order.txt)
#endif
template <int STRIDE, bool CACHED_WORLD = false, bool COLD = false>
__device__ __forceinline__ bool trav_step_impl(const TraceScene& S, Trav& t, const StackRef& stk, bool active, uint32_t warp, const bool ANY) {
  // ---- 1
  uint4 n0, n1, n2, n3, n4;
  uint32_t node_idx = 0;
  const uint4* nodes = S.blas_nodes;
  bool open = false;
  if (active) {
    if (t.T.y != 0u && !t.in_blas) {
      const uint32_t bit = __ffs(t.T.y) - 1;
      t.T.y &= ~(1u << bit);
      const uint32_t slot = __ldg(S.tlas_leaf_ids + t.T.x + bit);
      // instance culling: a ray that misses the bounding sphere skips the entry (this lane idles for the rest of the iteration)
      if (!S.inst_sphere || ray_may_hit_sphere(__ldg(S.inst_sphere + slot), t.r.ox, t.r.oy, t.r.oz, t.r.dx, t.r.dy, t.r.dz, t.tcur)) {
      // the instance record is requested before the (branchy) stack pushes so that its latency passes under them
      const float4* M = reinterpret_cast<const float4*>(S.inst + slot);
      const float4 r0 = __ldg(M), r1 = __ldg(M + 1), r2 = __ldg(M + 2);
      const uint4 meta = __ldg(reinterpret_cast<const uint4*>(M + 3));
      if (t.T.y) stk_push<STRIDE>(stk, t.sp, t.T, S.overflow);
      if (t.G.y & 0xFF000000u) stk_push<STRIDE>(stk, t.sp, t.G, S.overflow);
      t.blas_sp = t.sp;
      // o' = ((m0*x + m1*y) + m2*z) + m3 ; d' = (m0*x + m1*y) + m2*z   (unfused, as the oracle); in the TLAS t.r IS the world-space ray
      const float wox = t.r.ox, woy = t.r.oy, woz = t.r.oz, wdx = t.r.dx, wdy = t.r.dy, wdz = t.r.dz;
      const float oox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, wox), __fmul_rn(r0.y, woy)), __fmul_rn(r0.z, woz)), r0.w);
      const float ooy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, wox), __fmul_rn(r1.y, woy)), __fmul_rn(r1.z, woz)), r1.w);
      const float ooz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, wox), __fmul_rn(r2.y, woy)), __fmul_rn(r2.z, woz)), r2.w);
      const float odx = __fadd_rn(__fadd_rn(__fmul_rn(r0.x, wdx), __fmul_rn(r0.y, wdy)), __fmul_rn(r0.z, wdz));
      const float ody = __fadd_rn(__fadd_rn(__fmul_rn(r1.x, wdx), __fmul_rn(r1.y, wdy)), __fmul_rn(r1.z, wdz));
      const float odz = __fadd_rn(__fadd_rn(__fmul_rn(r2.x, wdx), __fmul_rn(r2.y, wdy)), __fmul_rn(r2.z, wdz));
      setup_box_space(t.r, oox, ooy, ooz, odx, ody, odz);
      t.in_blas = true; t.tri_ready = false;
      t.cur_inst = meta.y;
      if (SNK_EARLY_NODE_LOAD) load_node(S.blas_nodes, meta.x, n0, n1, n2, n3, n4);
      node_idx = meta.x;
      open = true;
      }
    } else if (t.T.y == 0u && (t.G.y & 0xFF000000u) != 0u) {
      const uint32_t hits_imask = t.G.y;
      const uint32_t bit = 31u - __clz(hits_imask);
      const uint32_t slot = (bit - 24u) ^ t.r.octinv;
      node_idx = t.G.x + __popc(hits_imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
      nodes = t.in_blas ? S.blas_nodes : S.tlas_nodes;
      if (SNK_EARLY_NODE_LOAD) load_node(nodes, node_idx, n0, n1, n2, n3, n4);
      t.G.y &= ~(1u << bit);
      if (t.G.y & 0xFF000000u) stk_push<STRIDE>(stk, t.sp, t.G, S.overflow);
      open = true;
    }
  }
  __syncwarp(warp);
  // ---- 2
  if (open) {
    if (!SNK_EARLY_NODE_LOAD) load_node(nodes, node_idx, n0, n1, n2, n3, n4);
    const uint32_t hm = intersect_node(t.r, n0, n1, n2, n3, n4, t.tmin, t.tcur);
    t.G = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
    t.T = make_uint2(n1.y, hm & 0x00FFFFFFu);
  }
  // ---- 3
  bool done = false;
  const bool has_tri = active && t.in_blas && t.T.y != 0u;
  bool tri_go = has_tri;
  if (SNK_TRI_MIN > 1) {
    const uint32_t n_tri = __popc(__ballot_sync(warp, has_tri)), n_open = __popc(__ballot_sync(warp, open));
    tri_go = has_tri && (n_tri >= SNK_TRI_MIN || n_open == 0u);
  } else {
    __syncwarp(warp);
  }
#pragma unroll
  for (int rep = 0; rep < SNK_TRI_REPS; ++rep) {
    // SNK_TRI_REPS > 1: lanes with more records pending test another one in the same iteration
    const bool go = rep == 0 ? tri_go : (active && !done && t.in_blas && t.T.y != 0u);
    if (rep && __ballot_sync(warp, go) == 0u) break;
    if (go) {
      const uint32_t bit = __ffs(t.T.y) - 1;
      t.T.y &= ~(1u << bit);
      // the record is requested before the (divergent, three-division) Woop setup of the lanes that test their first triangle in this
      // instance, so that its latency passes under the setup instead of after it
      const float4* tp = S.tris + 3ull * (t.T.x + bit);
      const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
      if (!t.tri_ready) { setup_tri_space(t.r); t.tri_ready = true; }
      float th, u, v;
      if (tri_test<COLD>(t.r, p0, p1, p2, th, u, v) && th > t.tmin) {
        // tcur == tmax until the first hit, afterwards the closest distance so far: a hit must be strictly inside
        // (tmin, tmax); at the current closest distance the lower (instance, primitive) wins
        const uint32_t prim = __float_as_uint(p0.w);
        bool take;
        if (ANY) take = th < t.tcur;
        else if (COLD) {
          take = th < t.tcur;
          if (th == t.tcur && t.best.t >= 0.0f) {  // exact tie with the closest hit so far (rare): its ids are in shared memory
            const float4 b = stk.world[3 * STRIDE];
            const uint32_t bprim = __float_as_uint(b.z), binst = __float_as_uint(b.w);
            take = t.cur_inst < binst || (t.cur_inst == binst && prim < bprim);
          }
        } else take = th < t.tcur || (th == t.tcur && t.best.t >= 0.0f &&
                                      (t.cur_inst < t.best.inst || (t.cur_inst == t.best.inst && prim < t.best.prim)));
        if (take) {
          if (COLD) { t.best.t = th; stk.world[3 * STRIDE] = make_float4(u, v, __uint_as_float(prim), __uint_as_float(t.cur_inst)); }
          else { t.best.t = th; t.best.u = u; t.best.v = v; t.best.prim = prim; t.best.inst = t.cur_inst; }
          t.tcur = th;
          done = ANY;
        }
      }
    }
    __syncwarp(warp);
  }
  // ---- 4
  if (active && !done && t.T.y == 0u && (t.G.y & 0xFF000000u) == 0u) {
    if (t.in_blas && t.sp == t.blas_sp) {
      // leave the instance: back to world space (no triangles are tested in the TLAS, so only the box part); kernels that
      // cache the world-space constants in shared memory reload them instead of recomputing three reciprocals
      if (CACHED_WORLD) {
        const float4 w = *stk.world;
        if (COLD) {
          const float4 wo = stk.world[STRIDE], wd = stk.world[2 * STRIDE];
          t.r.ox = wo.x; t.r.oy = wo.y; t.r.oz = wo.z; t.r.dx = wd.x; t.r.dy = wd.y; t.r.dz = wd.z;
        } else {
          t.r.ox = t.ox; t.r.oy = t.oy; t.r.oz = t.oz; t.r.dx = t.dx; t.r.dy = t.dy; t.r.dz = t.dz;
        }
        t.r.idx = w.x; t.r.idy = w.y; t.r.idz = w.z; t.r.octinv = __float_as_uint(w.w);
      } else {
        setup_box_space(t.r, t.ox, t.oy, t.oz, t.dx, t.dy, t.dz);
      }
      t.in_blas = false; t.tri_ready = false;
    }
    if (t.sp == 0) done = true;
    else {
      const uint2 e = stk_pop<STRIDE>(stk, t.sp);
      if (e.y & 0xFF000000u) t.G = e;
      else { t.T = e; t.G = make_uint2(0u, 0u); }
    }
  }
  return done;
}

template <bool ANY, int STRIDE>
__device__ __forceinline__ bool trav_step(const TraceScene& S, Trav& t, const StackRef& stk, bool active, uint32_t warp) {
  return trav_step_impl<STRIDE>(S, t, stk, active, warp, ANY);
}

// ---- phase-voting form (queue traversal kernel, trace_q.cu) ------------------------------------------------------------
// trav_step runs its node / triangle / instance-entry code every iteration with whatever subset of lanes wants
// it: ncu (profiles/r01m_prof_trace.txt) shows 13-14 of 32 threads active per issued instruction at ~78 % issue
// utilisation, i.e. the kernel is bound by issue slots spent on mostly-empty phases. Here every lane names the ONE
// phase it needs next and the warp executes only the phase with the most takers; the others wait (they cost no
// issue slots) until their phase wins the vote. Invariant between calls: an active lane has an inner-node group
// (G) or a leaf-record group (T) pending.
enum : uint32_t { WANT_NONE = 0u, WANT_ENTER = 1u, WANT_NODE = 2u, WANT_TRI = 3u };

#ifndef SNK_VOTE_TRI_BIAS
#define SNK_VOTE_TRI_BIAS 0  // the triangle phase also runs when it has at least (node takers - bias) takers
#endif

template <int STRIDE>
__device__ __forceinline__ bool trav_vote_step(const TraceScene& S, Trav& t, const StackRef& stk, bool active, bool any) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  uint32_t want = WANT_NONE;
  if (active) want = t.T.y != 0u ? (t.in_blas ? WANT_TRI : WANT_ENTER) : WANT_NODE;
  const int n_node = __popc(__ballot_sync(FULL, want == WANT_NODE));
  const int n_tri = __popc(__ballot_sync(FULL, want == WANT_TRI));
  const int n_enter = __popc(__ballot_sync(FULL, want == WANT_ENTER));
  uint32_t phase;
  if (n_node >= n_tri + SNK_VOTE_TRI_BIAS && n_node >= n_enter) phase = WANT_NODE;
  else if (n_tri >= n_enter) phase = WANT_TRI;
  else phase = WANT_ENTER;
  bool done = false;
  const bool mine = want == phase;
  if (phase == WANT_NODE) {
    if (mine) {
      const uint32_t hits_imask = t.G.y;
      const uint32_t bit = 31u - __clz(hits_imask);
      t.G.y &= ~(1u << bit);
      if (t.G.y & 0xFF000000u) stk_push<STRIDE>(stk, t.sp, t.G, S.overflow);
      const uint32_t slot = (bit - 24u) ^ t.r.octinv;
      const uint32_t node_idx = t.G.x + __popc(hits_imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
      const uint4* nodes = t.in_blas ? S.blas_nodes : S.tlas_nodes;
      uint4 n0, n1, n2, n3, n4;
      load_node(nodes, node_idx, n0, n1, n2, n3, n4);
      const uint32_t hm = intersect_node(t.r, n0, n1, n2, n3, n4, t.tmin, t.tcur);
      t.G = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
      t.T = make_uint2(n1.y, hm & 0x00FFFFFFu);
    }
  } else if (phase == WANT_TRI) {
    if (mine) {
      const uint32_t bit = __ffs(t.T.y) - 1;
      t.T.y &= ~(1u << bit);
      if (!t.tri_ready) { setup_tri_space(t.r); t.tri_ready = true; }
      const float4* tp = S.tris + 3ull * (t.T.x + bit);
      const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
      float th, u, v;
      if (tri_test(t.r, p0, p1, p2, th, u, v) && th > t.tmin) {
        const uint32_t prim = __float_as_uint(p0.w);
        const bool take = th < t.tcur || (!any && th == t.tcur && t.best.t >= 0.0f &&
                                          (t.cur_inst < t.best.inst || (t.cur_inst == t.best.inst && prim < t.best.prim)));
        if (take) {
          t.best.t = th; t.best.u = u; t.best.v = v; t.best.prim = prim; t.best.inst = t.cur_inst;
          t.tcur = th;
          done = any;
        }
      }
    }
  } else {
    if (mine) {
      // enter one instance: the rest of the leaf group and the pending inner group go on the stack, the ray moves
      // to object space and the BLAS root becomes the pending inner group (one inner bit, empty child mask)
      const uint32_t bit = __ffs(t.T.y) - 1;
      t.T.y &= ~(1u << bit);
      const uint32_t slot = __ldg(S.tlas_leaf_ids + t.T.x + bit);
      if (t.T.y) stk_push<STRIDE>(stk, t.sp, t.T, S.overflow);
      if (t.G.y & 0xFF000000u) stk_push<STRIDE>(stk, t.sp, t.G, S.overflow);
      t.blas_sp = t.sp;
      const float4* M = reinterpret_cast<const float4*>(S.inst + slot);
      const float4 r0 = __ldg(M), r1 = __ldg(M + 1), r2 = __ldg(M + 2);
      const uint4 meta = __ldg(reinterpret_cast<const uint4*>(M + 3));
      const float oox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, t.ox), __fmul_rn(r0.y, t.oy)), __fmul_rn(r0.z, t.oz)), r0.w);
      const float ooy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, t.ox), __fmul_rn(r1.y, t.oy)), __fmul_rn(r1.z, t.oz)), r1.w);
      const float ooz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, t.ox), __fmul_rn(r2.y, t.oy)), __fmul_rn(r2.z, t.oz)), r2.w);
      const float odx = __fadd_rn(__fadd_rn(__fmul_rn(r0.x, t.dx), __fmul_rn(r0.y, t.dy)), __fmul_rn(r0.z, t.dz));
      const float ody = __fadd_rn(__fadd_rn(__fmul_rn(r1.x, t.dx), __fmul_rn(r1.y, t.dy)), __fmul_rn(r1.z, t.dz));
      const float odz = __fadd_rn(__fadd_rn(__fmul_rn(r2.x, t.dx), __fmul_rn(r2.y, t.dy)), __fmul_rn(r2.z, t.dz));
      setup_box_space(t.r, oox, ooy, ooz, odx, ody, odz);
      t.in_blas = true; t.tri_ready = false;
      t.cur_inst = meta.y;
      t.G = make_uint2(meta.x, 0x01000000u);
      t.T = make_uint2(0u, 0u);
    }
  }
  __syncwarp(FULL);
  // pop: lanes of the executed phase that have nothing pending any more
  if (mine && !done && t.T.y == 0u && (t.G.y & 0xFF000000u) == 0u) {
    if (t.in_blas && t.sp == t.blas_sp) {
      setup_box_space(t.r, t.ox, t.oy, t.oz, t.dx, t.dy, t.dz);
      t.in_blas = false; t.tri_ready = false;
    }
    if (t.sp == 0) done = true;
    else {
      const uint2 e = stk_pop<STRIDE>(stk, t.sp);
      if (e.y & 0xFF000000u) t.G = e;
      else { t.T = e; t.G = make_uint2(0u, 0u); }
    }
  }
  return done;
}

// ---- warp-synchronous inline tracer (shading kernels) --------------------------------------------------------
// The lanes of a warp that are converged at the call trace one ray each through the same three phases;
// lanes without a ray pass tmax <= tmin. NOT inlined: the ReSTIR kernels call it from up to 11 program
// points, and the inlined copies (13 K SASS instructions per kernel) stalled 70-80 % of all issue slots on
// instruction-cache misses (profiles/r01c_prof_restir.txt).
template <bool ANY, int STRIDE>
__device__ __forceinline__ void trace_warp_body(const TraceScene& S, float4 a, float4 b, const StackRef& stk, HitRec& best) {
  const uint32_t warp = __activemask();
  Trav t;
  trav_begin(t, a, b);
  bool active = b.w > a.w && S.tlas_node_count != 0u;
  while (__ballot_sync(warp, active)) {
    if (trav_step<ANY, STRIDE>(S, t, stk, active, warp)) active = false;
  }
  best = t.best;
}
// returns 1.0 when a hit exists in (tmin, tmax), else 0.0 (RayGen.rgen:101-121 ShadowTest → payload.hit)
template <int STRIDE>
__device__ __noinline__ float trace_warp_any(const TraceScene* S, float ox, float oy, float oz, float tmin, float dx, float dy, float dz,
                                             float tmax, uint2* sm_col, uint2* gl_col) {
  HitRec h;
  trace_warp_body<true, STRIDE>(*S, make_float4(ox, oy, oz, tmin), make_float4(dx, dy, dz, tmax), StackRef{sm_col, gl_col}, h);
  return h.t >= 0.0f ? 1.0f : 0.0f;
}

}  // namespace snk
