// bvh_build.cu — LBVH -> BVH8 builder for sm_100a. See bvh_build.cuh for the
// reference entry points this replaces (TLAS.cpp:10-87, :109-186).
//
// All stages are HBM-bound integer / float32 streaming kernels; grids are sized
// from the primitive count with 256-thread CTAs (the collapse kernel is the one
// latency-bound stage: one thread per wide node, fed through a ticket queue).
#include "bvh_build.cuh"
#include "kernels.h"

#include <atomic>
#include <cstdlib>

#include <cfloat>
#include <cstdio>

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace snk {

namespace {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
// tile shape of the radix passes (profiles/r03f_variants_sort_shape.txt): 256 threads x 16 keys held 106 registers (25 % occupancy,
// ncu r03e: issue slots 7 % busy); 512 x 8 at 64 registers is 5 % faster on a 10 M-key sort, the other shapes are within +-3 %
#ifndef SNK_SORT_THREADS
#define SNK_SORT_THREADS 512
#endif
#ifndef SNK_SORT_ITEMS
#define SNK_SORT_ITEMS 8
#endif
#ifndef SNK_SORT_PREFETCH_VALS
#define SNK_SORT_PREFETCH_VALS 0
#endif
#ifndef SNK_SORT_MIN_CTAS
#define SNK_SORT_MIN_CTAS 2
#endif
constexpr int SORT_THREADS = SNK_SORT_THREADS;
constexpr int SORT_ITEMS = SNK_SORT_ITEMS;                      // keys per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;            // 4096 keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;
#ifndef SNK_LEAF_MAX_BLAS
#define SNK_LEAF_MAX_BLAS 2
#endif
#ifndef SNK_LEAF_MAX_TLAS
#define SNK_LEAF_MAX_TLAS 1
#endif
// primitives per leaf slot (<= 3: the slot's meta byte holds a unary count). Triangles: 3. Instances: 1 — entering an
// instance costs a ray transform plus at least the BLAS root test, so every instance gets its own box test first
// (with 3 per slot a ray entered 5.1 instances on C2, with 1 it enters 2.6: frame 8.08 -> 7.52 ms; 2 instead of 3 triangles per
// slot: 7.29 ms, 1: 7.27 ms with 30 % more nodes — profiles/r02d_variants_leaf_size.txt).
constexpr int LEAF_MAX_BLAS = SNK_LEAF_MAX_BLAS, LEAF_MAX_TLAS = SNK_LEAF_MAX_TLAS;

enum { C_NODES = 0, C_LEAVES = 1, C_DONE = 2, C_TICKET = 3, C_OVERFLOW = 4, C_PENDING = 5, C_BOUNDS = 8 };

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__device__ __forceinline__ void init_counters(uint32_t* counters, uint32_t* queue, int t);
__global__ void k_init(uint32_t* counters, uint32_t* queue) { init_counters(counters, queue, threadIdx.x); }
__device__ __forceinline__ void init_counters(uint32_t* counters, uint32_t* queue, int t) {
  if (t < 16) {
    uint32_t v = 0;
    if (t == C_NODES) v = 1;
    if (t >= C_BOUNDS && t < C_BOUNDS + 3) v = 0xFFFFFFFFu;
    counters[t] = v;
  }
  if (t == 0) queue[0] = 0;
}

__device__ __forceinline__ void reduce_centroid_bounds(float cx, float cy, float cz, bool valid, uint32_t* counters) {
  float lo[3] = {valid ? cx : FLT_MAX, valid ? cy : FLT_MAX, valid ? cz : FLT_MAX};
  float hi[3] = {valid ? cx : -FLT_MAX, valid ? cy : -FLT_MAX, valid ? cz : -FLT_MAX};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
  }
  // warp results -> shared -> one atomic set per CTA (per-warp atomics on six addresses serialised at L2:
  // 0.38 ms of a 3 M-triangle build)
  __shared__ float s_lo[3][8], s_hi[3][8];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { s_lo[k][wid] = lo[k]; s_hi[k][wid] = hi[k]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float l = FLT_MAX, h = -FLT_MAX;
    for (int w = 0; w < nw; ++w) { l = fminf(l, s_lo[threadIdx.x][w]); h = fmaxf(h, s_hi[threadIdx.x][w]); }
    if (l <= h) {
      atomicMin(&counters[C_BOUNDS + threadIdx.x], f2ord(l));
      atomicMax(&counters[C_BOUNDS + 3 + threadIdx.x], f2ord(h));
    }
  }
}

// widen by 2 ulp so that the (conservatively quantised) boxes also cover the rounding of the triangle test
__device__ __forceinline__ float pad_dn(float v) { return v - fabsf(v) * 2.4e-7f - 1e-30f; }
__device__ __forceinline__ float pad_up(float v) { return v + fabsf(v) * 2.4e-7f + 1e-30f; }

__global__ void k_triangle_bounds(const float* __restrict__ positions, const uint32_t* __restrict__ indices, uint32_t index_offset,
                                  uint32_t vertex_offset, uint32_t n, float4* __restrict__ prim_lo, float4* __restrict__ prim_hi,
                                  uint32_t* counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  float cx = 0, cy = 0, cz = 0;
  if (valid) {
    const uint32_t* ip = indices + index_offset + 3ull * i;
    const uint32_t a = ip[0] + vertex_offset, b = ip[1] + vertex_offset, c = ip[2] + vertex_offset;
    const float ax = positions[3ull * a], ay = positions[3ull * a + 1], az = positions[3ull * a + 2];
    const float bx = positions[3ull * b], by = positions[3ull * b + 1], bz = positions[3ull * b + 2];
    const float cx2 = positions[3ull * c], cy2 = positions[3ull * c + 1], cz2 = positions[3ull * c + 2];
    const float lx = fminf(ax, fminf(bx, cx2)), ly = fminf(ay, fminf(by, cy2)), lz = fminf(az, fminf(bz, cz2));
    const float hx = fmaxf(ax, fmaxf(bx, cx2)), hy = fmaxf(ay, fmaxf(by, cy2)), hz = fmaxf(az, fmaxf(bz, cz2));
    prim_lo[i] = make_float4(pad_dn(lx), pad_dn(ly), pad_dn(lz), 0.f);
    prim_hi[i] = make_float4(pad_up(hx), pad_up(hy), pad_up(hz), 0.f);
    cx = 0.5f * (lx + hx); cy = 0.5f * (ly + hy); cz = 0.5f * (lz + hz);
  }
  reduce_centroid_bounds(cx, cy, cz, valid, counters);
}

// dequantised box of child slot s of a wide node (exactly the planes the traversal tests)
__device__ __forceinline__ void child_box(const SnkNode8& nd, int s, float* lo, float* hi) {
  const float sx = __uint_as_float((uint32_t)nd.ex << 23), sy = __uint_as_float((uint32_t)nd.ey << 23), sz = __uint_as_float((uint32_t)nd.ez << 23);
  lo[0] = nd.px + (float)nd.qlox[s] * sx; lo[1] = nd.py + (float)nd.qloy[s] * sy; lo[2] = nd.pz + (float)nd.qloz[s] * sz;
  hi[0] = nd.px + (float)nd.qhix[s] * sx; hi[1] = nd.py + (float)nd.qhiy[s] * sy; hi[2] = nd.pz + (float)nd.qhiz[s] * sz;
}
// world AABB of an object-space box under the affine map M (column-major mat4): centre M*c, half extent |M|*e — the
// same box as transforming the 8 corners, at a quarter of the arithmetic (the caller pads the result)
__device__ __forceinline__ void grow_by_transformed_box(const float* __restrict__ M, const float* blo, const float* bhi, float* lo, float* hi) {
  const float cx = 0.5f * (blo[0] + bhi[0]), cy = 0.5f * (blo[1] + bhi[1]), cz = 0.5f * (blo[2] + bhi[2]);
  const float ex = 0.5f * (bhi[0] - blo[0]), ey = 0.5f * (bhi[1] - blo[1]), ez = 0.5f * (bhi[2] - blo[2]);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float c = M[r] * cx + M[4 + r] * cy + M[8 + r] * cz + M[12 + r];
    const float e = fabsf(M[r]) * ex + fabsf(M[4 + r]) * ey + fabsf(M[8 + r]) * ez;
    lo[r] = fminf(lo[r], c - e); hi[r] = fmaxf(hi[r], c + e);
  }
}

// World AABB of an instance = union of the transformed boxes of the BLAS's depth-2 children (up to 64 boxes)
// rather than of its root box: the transformed root box of a rotated mesh is up to sqrt(3) too wide per axis,
// which on C2 made every ray enter ~6 instances it could not hit.
// One WARP per instance: lane l takes child s = l / 4 of the BLAS root and, when that child is an inner node, grandchildren
// 2 (l % 4) and 2 (l % 4) + 1 (one thread per instance walked the 64 boxes serially: 53 us of the 0.2 ms per-frame TLAS rebuild
// for the 205 instances of C2). min / max are order independent: same boxes as before.
__global__ void __launch_bounds__(256) k_instance_bounds(const float* __restrict__ transforms, const SnkInstanceData* __restrict__ inst,
                                  const uint32_t* __restrict__ blas_of_instance, const SnkBlasInfo* __restrict__ blas,
                                  const SnkNode8* __restrict__ blas_nodes, const uint32_t* __restrict__ valid_slots, uint32_t n,
                                  float4* __restrict__ prim_lo, float4* __restrict__ prim_hi, uint32_t* counters) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  const bool valid = i < n;
  float cx = 0, cy = 0, cz = 0;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (valid) {
    const uint32_t slot = valid_slots[i];
    const SnkBlasInfo b = blas[blas_of_instance[slot]];
    const float* M = transforms + 16ull * inst[slot].transform_idx;
    const SnkNode8& root = blas_nodes[b.node_root];
    const int s = (int)(lane >> 2), part = (int)(lane & 3u);
    const uint32_t meta = root.meta[s];
    if (meta != 0) {
      float blo[3], bhi[3];
      child_box(root, s, blo, bhi);
      if ((meta & 0x18u) == 0x18u) {
        const uint32_t inner_rank = __popc((uint32_t)root.imask & ((1u << s) - 1u));
        const SnkNode8& ch = blas_nodes[root.child_base + inner_rank];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s2 = 2 * part + h;
          if (ch.meta[s2] == 0) continue;
          float clo[3], chi[3];
          child_box(ch, s2, clo, chi);
          // the grandchild lies inside the child box as well: intersect (both are conservative)
          for (int k = 0; k < 3; ++k) { clo[k] = fmaxf(clo[k], blo[k]); chi[k] = fminf(chi[k], bhi[k]); }
          grow_by_transformed_box(M, clo, chi, lo, hi);
        }
      } else if (part == 0) {
        grow_by_transformed_box(M, blo, bhi, lo, hi);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[r] = fminf(lo[r], __shfl_xor_sync(0xffffffffu, lo[r], o));
      hi[r] = fmaxf(hi[r], __shfl_xor_sync(0xffffffffu, hi[r], o));
    }
  }
  if (valid) {
    // instance boxes are transformed in float32: widen by a relative 1e-5 so they stay conservative
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float e = 1e-5f * fmaxf(fmaxf(fabsf(lo[r]), fabsf(hi[r])), 1e-3f);
      lo[r] -= e; hi[r] += e;
    }
    if (lane == 0) {
      prim_lo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
      prim_hi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    cx = 0.5f * (lo[0] + hi[0]); cy = 0.5f * (lo[1] + hi[1]); cz = 0.5f * (lo[2] + hi[2]);
  }
  reduce_centroid_bounds(cx, cy, cz, valid, counters);
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

#ifndef SNK_HOIST_LARGE
#define SNK_HOIST_LARGE 8  // 0 = off
#endif
__device__ __forceinline__ unsigned long long morton_key(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi, uint32_t i,
                                                         const uint32_t* __restrict__ counters, const bool hoist_large = false) {
  const float lx = ord2f(counters[C_BOUNDS]), ly = ord2f(counters[C_BOUNDS + 1]), lz = ord2f(counters[C_BOUNDS + 2]);
  const float hx = ord2f(counters[C_BOUNDS + 3]), hy = ord2f(counters[C_BOUNDS + 4]), hz = ord2f(counters[C_BOUNDS + 5]);
  const float4 a = prim_lo[i], b = prim_hi[i];
#ifdef SNK_MORTON_UNIFORM
  // one scale for the three axes (cubic cells): a flat scene does not spend its top bits on the short axis
  const float ext = fmaxf(fmaxf(hx - lx, hy - ly), hz - lz);
  const float sx = ext > 0.f ? 2097152.0f / ext : 0.f, sy = sx, sz = sx;
#else
  const float sx = hx > lx ? 2097152.0f / (hx - lx) : 0.f, sy = hy > ly ? 2097152.0f / (hy - ly) : 0.f,
              sz = hz > lz ? 2097152.0f / (hz - lz) : 0.f;
#endif
  const float qx = fminf(fmaxf((0.5f * (a.x + b.x) - lx) * sx, 0.f), 2097151.f);
  const float qy = fminf(fmaxf((0.5f * (a.y + b.y) - ly) * sy, 0.f), 2097151.f);
  const float qz = fminf(fmaxf((0.5f * (a.z + b.z) - lz) * sz, 0.f), 2097151.f);
  unsigned long long key = (expand21((unsigned long long)qx) << 2) | (expand21((unsigned long long)qy) << 1) | expand21((unsigned long long)qz);
#if SNK_HOIST_LARGE
  // Large-primitive hoisting (triangle builds of the multi-CTA path): a triangle much larger than its neighbours (a floor under a
  // field of small objects) inflates the box of every ancestor on its path to the root, and every ray then has to open those
  // ancestors. Bit 63 (above the 63 Morton bits; k_morton shifts it down with the key) moves such triangles into a subtree of
  // their own directly under the root: two Morton-ordered hierarchies, one of the few large triangles and one of everything else
  // with tight boxes. Large = longest box edge above 1/SNK_HOIST_LARGE of the longest edge of the centroid bounds.
  // Measured (profiles/r05m_variants_hoist_large.txt): 1 M-triangle flat scene 15.4 -> 14.9 nodes per ray, frame 6.87 -> 6.74 ms;
  // 3 M flat 33.7 -> 33.2 ms; thresholds 1/4, 1/8, 1/32 alike. For the TLAS (the floor INSTANCE) and the small meshes of the
  // single-CTA builder it costs nodes (C2: 7.8 -> 8.4 per ray, frame 5.10 -> 5.21 ms): not applied there.
  if (hoist_large) {
    const float ext = fmaxf(fmaxf(hx - lx, hy - ly), hz - lz), pe = fmaxf(fmaxf(b.x - a.x, b.y - a.y), b.z - a.z);
    if (pe * (float)SNK_HOIST_LARGE > ext && ext > 0.f) key |= 1ull << 63;
  }
#endif
  return key;
}
// `drop` low key bits are shifted out: the radix sort then needs (63 - drop) / 8 passes instead of 8 (launch_build)
__global__ void k_morton(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi, uint32_t n,
                         const uint32_t* __restrict__ counters, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, int drop,
                         bool hoist_large) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = morton_key(prim_lo, prim_hi, i, counters, hoist_large) >> drop;
  vals[i] = i;
}

// ---- radix sort: one pass = histogram, scan, scatter ------------------------------------
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift,
                                                            uint32_t* __restrict__ block_hist, uint32_t num_blocks) {
  __shared__ uint32_t sh[256];
  for (int d = threadIdx.x; d < 256; d += SORT_THREADS) sh[d] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * SORT_TILE;
#pragma unroll 4
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const uint32_t idx = base + it * SORT_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&sh[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < 256; d += SORT_THREADS) block_hist[d * num_blocks + blockIdx.x] = sh[d];
}

// block_hist is digit-major: row d holds the count of digit d in every tile. Offsets = (number of keys with a
// smaller digit) + (keys with digit d in earlier tiles): CTA d scans its own row in place and leaves the row
// total in digit_total[d]; k_scan_digits then turns the 256 totals into exclusive bases (added by the scatter).
__global__ void __launch_bounds__(256) k_scan_rows(uint32_t* __restrict__ block_hist, uint32_t num_blocks, uint32_t* __restrict__ digit_total) {
  __shared__ uint32_t warp_sums[8];
  __shared__ uint32_t carry;
  uint32_t* row = block_hist + (size_t)blockIdx.x * num_blocks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t base = 0; base < num_blocks; base += 256) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < num_blocks ? row[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < wid; ++w) wbase += warp_sums[w];
    const uint32_t ex = carry + wbase + inc - v;
    if (i < num_blocks) row[i] = ex;
    __syncthreads();
    if (threadIdx.x == 255) carry = ex + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}
__global__ void __launch_bounds__(256) k_scan_digits(uint32_t* __restrict__ digit_total) {
  __shared__ uint32_t warp_sums[8];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t v = digit_total[threadIdx.x];
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  uint32_t wbase = 0;
  for (uint32_t w = 0; w < wid; ++w) wbase += warp_sums[w];
  digit_total[threadIdx.x] = wbase + inc - v;
}

__global__ void __launch_bounds__(SORT_THREADS, SNK_SORT_MIN_CTAS) k_sort_scatter(const unsigned long long* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in, uint32_t n, int shift,
                                                               const uint32_t* __restrict__ block_offsets, uint32_t num_blocks,
                                                               const uint32_t* __restrict__ digit_base,
                                                               unsigned long long* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t wh[SORT_WARPS][256];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) wh[wid][d] = 0;
  __syncwarp();
  const uint32_t warp_base = blockIdx.x * SORT_TILE + wid * (SORT_ITEMS * 32);
  const uint32_t lt = (1u << lane) - 1u;
  unsigned long long key[SORT_ITEMS];
#if SNK_SORT_PREFETCH_VALS
  uint32_t val[SORT_ITEMS];  // fetched with the keys: phase 3 is a serial chain per warp and must not wait for a load per item
#endif
  // phase 1: per-warp digit histogram (each warp owns a contiguous run of the tile → stable)
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const uint32_t idx = warp_base + it * 32 + lane;
    const bool valid = idx < n;
    key[it] = valid ? keys_in[idx] : ~0ull;
#if SNK_SORT_PREFETCH_VALS
    val[it] = valid ? vals_in[idx] : 0u;
#endif
    const uint32_t d = valid ? ((uint32_t)(key[it] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (valid && (peers & lt) == 0) wh[wid][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // phase 2: turn per-warp counts into running output positions
  for (uint32_t d = threadIdx.x; d < 256u; d += SORT_THREADS) {  // one digit per thread
    uint32_t run = digit_base[d] + block_offsets[d * num_blocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      const uint32_t c = wh[w][d];
      wh[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // phase 3: rank + scatter
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const uint32_t idx = warp_base + it * 32 + lane;
    const bool valid = idx < n;
    const uint32_t d = valid ? ((uint32_t)(key[it] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & lt);
    uint32_t pos = 0;
    if (valid) pos = wh[wid][d] + rank;
    __syncwarp();
    if (valid && rank == 0) wh[wid][d] += __popc(peers);
    __syncwarp();
    if (valid) {
      keys_out[pos] = key[it];
#if SNK_SORT_PREFETCH_VALS
      vals_out[pos] = val[it];
#else
      vals_out[pos] = vals_in[idx];
#endif
    }
  }
}

// The same pass with the tile sorted in shared memory first: ranks are computed as above, but each (key, value) goes to its
// tile-local sorted position, and a second sweep writes the tile out in that order — keys of one digit leave the CTA as one
// contiguous run (16 keys = 128 B on average) instead of 32 unrelated 8-byte stores per warp instruction. ncu (r03e) had the
// direct version at 40 % L2 throughput with 7 % issue-slot use: bound by L2 write transactions, not by DRAM bytes.
#ifndef SNK_SORT_STAGED
#define SNK_SORT_STAGED 1
#endif
constexpr size_t SORT_STAGED_SMEM = (size_t)SORT_TILE * 12 + (size_t)SORT_WARPS * 256 * 4 + 256 * 4 * 2 + 64;
__global__ void __launch_bounds__(SORT_THREADS, SNK_SORT_MIN_CTAS) k_sort_scatter_staged(const unsigned long long* __restrict__ keys_in,
                                                                      const uint32_t* __restrict__ vals_in, uint32_t n, int shift,
                                                                      const uint32_t* __restrict__ block_offsets, uint32_t num_blocks,
                                                                      const uint32_t* __restrict__ digit_base,
                                                                      unsigned long long* __restrict__ keys_out,
                                                                      uint32_t* __restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(sort_smem);
  uint32_t* sval = reinterpret_cast<uint32_t*>(skey + SORT_TILE);
  uint32_t (*wh)[256] = reinterpret_cast<uint32_t (*)[256]>(sval + SORT_TILE);
  uint32_t* tile_start = reinterpret_cast<uint32_t*>(wh + SORT_WARPS);  // first tile-local position of digit d
  uint32_t* goff = tile_start + 256;                                    // global position of tile-local position j: goff[d] + j
  uint32_t* warp_sums = goff + 256;
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) wh[wid][d] = 0;
  __syncwarp();
  const uint32_t tile_base = blockIdx.x * SORT_TILE;
  const uint32_t warp_base = tile_base + wid * (SORT_ITEMS * 32);
  const uint32_t lt = (1u << lane) - 1u;
  unsigned long long key[SORT_ITEMS];
  uint32_t val[SORT_ITEMS];
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const uint32_t idx = warp_base + it * 32 + lane;
    const bool valid = idx < n;
    key[it] = valid ? keys_in[idx] : ~0ull;
    val[it] = valid ? vals_in[idx] : 0u;
  }
  // phase 1: per-warp digit histogram (each warp owns a contiguous run of the tile -> stable)
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const bool valid = warp_base + it * 32 + lane < n;
    const uint32_t d = valid ? ((uint32_t)(key[it] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (valid && (peers & lt) == 0) wh[wid][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // phase 2: per-warp counts -> tile-local start positions; tile totals -> exclusive scan over the digits
  uint32_t total = 0, d = threadIdx.x;
  if (d < 256u) {
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      const uint32_t c = wh[w][d];
      wh[w][d] = total;
      total += c;
    }
  }
  uint32_t inc = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31 && wid < 8) warp_sums[wid] = inc;
  __syncthreads();
  if (d < 256u) {
    uint32_t before = 0;
    for (uint32_t w = 0; w < wid; ++w) before += warp_sums[w];
    const uint32_t start = before + inc - total;
    tile_start[d] = start;
    goff[d] = digit_base[d] + block_offsets[d * num_blocks + blockIdx.x] - start;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) wh[w][d] += start;
  }
  __syncthreads();
  // phase 3: rank + move to the tile-local sorted position
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const bool valid = warp_base + it * 32 + lane < n;
    const uint32_t dg = valid ? ((uint32_t)(key[it] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, dg);
    const uint32_t rank = __popc(peers & lt);
    uint32_t pos = 0;
    if (valid) pos = wh[wid][dg] + rank;
    __syncwarp();
    if (valid && rank == 0) wh[wid][dg] += __popc(peers);
    __syncwarp();
    if (valid) { skey[pos] = key[it]; sval[pos] = val[it]; }
  }
  __syncthreads();
  // phase 4: write the tile out in sorted order
  const uint32_t tile_n = n - tile_base < (uint32_t)SORT_TILE ? n - tile_base : (uint32_t)SORT_TILE;
  for (uint32_t j = threadIdx.x; j < tile_n; j += SORT_THREADS) {
    const unsigned long long k = skey[j];
    const uint32_t g = goff[(uint32_t)(k >> shift) & 255u] + j;
    keys_out[g] = k;
    vals_out[g] = sval[j];
  }
}

static void launch_sort_scatter(const unsigned long long* keys_in, const uint32_t* vals_in, uint32_t n, int shift, const uint32_t* block_offsets,
                                uint32_t nb, const uint32_t* digit_base, unsigned long long* keys_out, uint32_t* vals_out, cudaStream_t st) {
#if SNK_SORT_STAGED
  static bool configured[64] = {};  // the opt-in to > 48 KB of dynamic shared memory is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(k_sort_scatter_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_STAGED_SMEM);
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  k_sort_scatter_staged<<<nb, SORT_THREADS, SORT_STAGED_SMEM, st>>>(keys_in, vals_in, n, shift, block_offsets, nb, digit_base, keys_out, vals_out);
#else
  k_sort_scatter<<<nb, SORT_THREADS, 0, st>>>(keys_in, vals_in, n, shift, block_offsets, nb, digit_base, keys_out, vals_out);
#endif
}

// ---- one-sweep radix sort (round 2) -------------------------------------------------------------------------------------
// The LSD passes above cost four launches each (tile histograms, two scans, scatter) and read the keys twice. Here a pass is ONE
// kernel (Adinets & Merrill's "onesweep"): the digit totals of ALL passes come from one histogram kernel up front (they do not
// depend on the order of the keys), and a tile learns how many keys with its digit lie in the tiles before it by a decoupled
// look-back over per-tile status words (count | AGGREGATE or INCLUSIVE flag in one 32-bit word, so no fence is needed): tile t
// publishes its own counts as soon as it has them, then walks back over t-1, t-2, ... adding aggregates until it meets an
// inclusive prefix. Tiles are handed out by a ticket, so every tile a CTA waits for has already started. The scatter itself is
// the staged one above with ONE ranking pass instead of two: the first pass over a warp's keys also remembers each key's place
// among the warp's keys of the same digit. Same permutation as the LSD passes (stable in warp, tile and ticket order).
#ifndef SNK_SORT_ONESWEEP
#define SNK_SORT_ONESWEEP 1
#endif
constexpr int SWEEP_MAX_PASSES = 8;
constexpr uint32_t SW_TICKETS = SWEEP_MAX_PASSES * 256;  // word offset of the per-pass tickets
constexpr uint32_t SW_HEAD = SW_TICKETS + 64;             // [8][256] digit totals | 8 tickets | pad; then nb x 256 status words per pass
constexpr uint32_t SW_AGG = 1u << 30, SW_INC = 2u << 30, SW_COUNT = (1u << 30) - 1u;

// digit totals of all passes in one read of the keys. Shared-memory atomics cost 2 cycles per lane on this part (64 per warp
// instruction): with them this kernel took 176 us for 10 M keys x 5 passes. Instead every warp keeps PRIVATE counters and its
// lanes agree on them with match.any — the leader of each group of equal digits adds the group's size with a plain load / store.
constexpr int SWH_THREADS = 512, SWH_WARPS = SWH_THREADS / 32;
__global__ void __launch_bounds__(SWH_THREADS) k_sweep_hist(const unsigned long long* __restrict__ keys, uint32_t n, int passes,
                                                            uint32_t* __restrict__ head) {
  extern __shared__ uint32_t swh[];  // [warp][pass][256]
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, lt = (1u << lane) - 1u;
  uint32_t* mine = swh + (size_t)wid * passes * 256;
  for (int i = lane; i < passes * 256; i += 32) mine[i] = 0u;
  __syncwarp();
  const uint32_t stride = gridDim.x * SWH_THREADS;
  for (uint32_t base = blockIdx.x * SWH_THREADS + wid * 32u; base < n; base += stride) {  // warp-uniform trip count
    const uint32_t idx = base + lane;
    const bool valid = idx < n;
    const unsigned long long k = valid ? keys[idx] : 0ull;
    for (int p = 0; p < passes; ++p) {
      const uint32_t d = valid ? ((uint32_t)(k >> (8 * p)) & 255u) : 256u;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      if (valid && (peers & lt) == 0u) mine[p * 256 + d] += (uint32_t)__popc(peers);
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += SWH_THREADS) {
    uint32_t t = 0;
    for (int w = 0; w < SWH_WARPS; ++w) t += swh[(size_t)w * passes * 256 + i];
    if (t) atomicAdd(&head[i], t);
  }
}

constexpr size_t SWEEP_SMEM = (size_t)SORT_TILE * 12 + (size_t)SORT_WARPS * 256 * 4 + 256 * 4 * 2 + 64 + 16;
__global__ void __launch_bounds__(SORT_THREADS, SNK_SORT_MIN_CTAS) k_sweep_pass(const unsigned long long* __restrict__ keys_in,
                                                                                  const uint32_t* __restrict__ vals_in, uint32_t n, int shift, int pass,
                                                                                  uint32_t* head, uint32_t* status,
                                                                                  unsigned long long* __restrict__ keys_out,
                                                                                  uint32_t* __restrict__ vals_out) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(sort_smem);
  uint32_t* sval = reinterpret_cast<uint32_t*>(skey + SORT_TILE);
  uint32_t (*wh)[256] = reinterpret_cast<uint32_t (*)[256]>(sval + SORT_TILE);
  uint32_t* goff = reinterpret_cast<uint32_t*>(wh + SORT_WARPS);  // global position of tile-local position j: goff[d] + j
  uint32_t* warp_sums = goff + 256;                                // 8 tile sums | 8 global sums
  uint32_t* s_tile = warp_sums + 16;
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) *s_tile = atomicAdd(&head[SW_TICKETS + pass], 1u);
  for (int d = lane; d < 256; d += 32) wh[wid][d] = 0;
  __syncthreads();
  const uint32_t tile = *s_tile;
  const uint32_t tile_base = tile * SORT_TILE;
  const uint32_t warp_base = tile_base + wid * (SORT_ITEMS * 32);
  const uint32_t lt = (1u << lane) - 1u;
  unsigned long long key[SORT_ITEMS];
  uint32_t val[SORT_ITEMS];
  uint32_t off[SORT_ITEMS];  // place of the key among the keys of the same digit of this warp
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const uint32_t idx = warp_base + it * 32 + lane;
    const bool valid = idx < n;
    key[it] = valid ? keys_in[idx] : ~0ull;
    val[it] = valid ? vals_in[idx] : 0u;
  }
  // phase 1: per-warp digit counts (each warp owns a contiguous run of the tile -> stable) + the key's place in its warp's run
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const bool valid = warp_base + it * 32 + lane < n;
    const uint32_t d = valid ? ((uint32_t)(key[it] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & lt);
    const uint32_t prev = valid ? wh[wid][d] : 0u;
    __syncwarp();
    if (valid && rank == 0u) wh[wid][d] = prev + __popc(peers);
    __syncwarp();
    off[it] = prev + rank;
  }
  __syncthreads();
  // phase 2: per-warp counts -> exclusive over the warps; tile totals published, scanned over the digits, looked back over the tiles
  uint32_t total = 0, gcount = 0;
  const uint32_t d = threadIdx.x;
  volatile uint32_t* st = status;
  if (d < 256u) {
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      const uint32_t c = wh[w][d];
      wh[w][d] = total;
      total += c;
    }
    st[(size_t)tile * 256u + d] = (tile == 0u ? SW_INC : SW_AGG) | total;
    gcount = head[pass * 256 + d];
  }
  uint32_t inc = total, ginc = gcount;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o), g = __shfl_up_sync(0xffffffffu, ginc, o);
    if (lane >= o) { inc += t; ginc += g; }
  }
  if (lane == 31 && wid < 8) { warp_sums[wid] = inc; warp_sums[8 + wid] = ginc; }
  __syncthreads();
  if (d < 256u) {
    uint32_t before = 0, gbefore = 0;
    for (uint32_t w = 0; w < wid; ++w) { before += warp_sums[w]; gbefore += warp_sums[8 + w]; }
    const uint32_t start = before + inc - total;          // first tile-local position of digit d
    const uint32_t gbase = gbefore + ginc - gcount;        // keys with a smaller digit in the whole input
    uint32_t excl = 0;                                     // keys with digit d in earlier tiles
    if (tile > 0u) {
      uint32_t t = tile - 1u;
      for (;;) {
        const uint32_t v = st[(size_t)t * 256u + d];
        if ((v >> 30) == 0u) continue;                     // that tile has not published yet (it has started: ticket order)
        excl += v & SW_COUNT;
        if (v & SW_INC) break;
        --t;
      }
      st[(size_t)tile * 256u + d] = SW_INC | (excl + total);
    }
    goff[d] = gbase + excl - start;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) wh[w][d] += start;
  }
  __syncthreads();
  // phase 3: move to the tile-local sorted position
#pragma unroll
  for (int it = 0; it < SORT_ITEMS; ++it) {
    const bool valid = warp_base + it * 32 + lane < n;
    if (valid) {
      const uint32_t pos = wh[wid][(uint32_t)(key[it] >> shift) & 255u] + off[it];
      skey[pos] = key[it]; sval[pos] = val[it];
    }
  }
  __syncthreads();
  // phase 4: write the tile out in sorted order
  const uint32_t tile_n = n - tile_base < (uint32_t)SORT_TILE ? n - tile_base : (uint32_t)SORT_TILE;
  for (uint32_t j = threadIdx.x; j < tile_n; j += SORT_THREADS) {
    const unsigned long long k = skey[j];
    const uint32_t g = goff[(uint32_t)(k >> shift) & 255u] + j;
    keys_out[g] = k;
    vals_out[g] = sval[j];
  }
}

// SNKPT_SORT=lsd selects the four-launch LSD passes (A/B runs)
static bool sort_onesweep() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SNKPT_SORT"); v = (SNK_SORT_ONESWEEP && !(e && e[0] == 'l')) ? 1 : 0; }
  return v == 1;
}
// sorts s.keys[cur] / s.vals[cur] by the key bits [shift0, shift0 + 8 * passes); returns the number of stream operations
static int launch_sort_sweep(BuildScratch& s, uint32_t n, int passes, int& cur, cudaStream_t st) {
  static bool configured[64] = {};  // the opt-in to > 48 KB of dynamic shared memory is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(k_sweep_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM);
    cudaFuncSetAttribute(k_sweep_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, SWH_WARPS * SWEEP_MAX_PASSES * 256 * (int)sizeof(uint32_t));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const uint32_t nb = (n + SORT_TILE - 1) / SORT_TILE;
  cudaMemsetAsync(s.sweep, 0, sizeof(uint32_t) * ((size_t)SW_HEAD + (size_t)passes * nb * 256u), st);
  const uint32_t hist_want = (n + 4u * SWH_THREADS - 1u) / (4u * SWH_THREADS), hist_res = (uint32_t)device_sm_count() * 2u;
  const size_t hist_smem = (size_t)SWH_WARPS * passes * 256 * sizeof(uint32_t);
  k_sweep_hist<<<hist_want < hist_res ? hist_want : hist_res, SWH_THREADS, hist_smem, st>>>(s.keys[cur], n, passes, s.sweep);
  for (int p = 0; p < passes; ++p) {
    k_sweep_pass<<<nb, SORT_THREADS, SWEEP_SMEM, st>>>(s.keys[cur], s.vals[cur], n, 8 * p, p, s.sweep, s.sweep + SW_HEAD + (size_t)p * nb * 256u,
                                                        s.keys[cur ^ 1], s.vals[cur ^ 1]);
    cur ^= 1;
  }
  return passes + 2;
}

// ---- Karras 2012 ------------------------------------------------------------------------
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}

__device__ __forceinline__ void karras_node(const unsigned long long* __restrict__ keys, int n, int i, int* __restrict__ left,
                                            int* __restrict__ right, int* __restrict__ parent, int* __restrict__ first, int* __restrict__ last) {
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
  const int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
  left[i] = lc; right[i] = rc;
  parent[lc] = i; parent[rc] = i;
  first[i] = lo; last[i] = hi;
  if (i == 0) parent[0] = -1;
}
__global__ void k_karras(const unsigned long long* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                         int* __restrict__ parent, int* __restrict__ first, int* __restrict__ last) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  karras_node(keys, n, i, left, right, parent, first, last);
}

__device__ __forceinline__ void refit_from_leaf(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
                                                const uint32_t* __restrict__ sorted_prims, int n, int k, const int* __restrict__ left,
                                                const int* __restrict__ right, const int* __restrict__ parent, float4* bb_lo, float4* bb_hi,
                                                uint32_t* flags) {
  const uint32_t p = sorted_prims[k];
  int node = n - 1 + k;
  float4 lo = prim_lo[p], hi = prim_hi[p];
  bb_lo[node] = lo; bb_hi[node] = hi;
  int par = parent[node];
  while (par >= 0) {
    __threadfence();
    // the first child to arrive leaves its index (+1) and stops; the second one finds its sibling there and continues
    const uint32_t seen = atomicExch(&flags[par], (uint32_t)node + 1u);
    if (seen == 0u) return;
    __threadfence();
    const int other = (int)seen - 1;
    const float4 olo = __ldcg(&bb_lo[other]), ohi = __ldcg(&bb_hi[other]);
    lo = make_float4(fminf(lo.x, olo.x), fminf(lo.y, olo.y), fminf(lo.z, olo.z), 0.f);
    hi = make_float4(fmaxf(hi.x, ohi.x), fmaxf(hi.y, ohi.y), fmaxf(hi.z, ohi.z), 0.f);
    bb_lo[par] = lo; bb_hi[par] = hi;
    node = par;
    par = parent[node];
  }
}
__global__ void k_refit_binary(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
                               const uint32_t* __restrict__ sorted_prims, int n, const int* __restrict__ left,
                               const int* __restrict__ right, const int* __restrict__ parent, float4* bb_lo, float4* bb_hi,
                               uint32_t* flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  refit_from_leaf(prim_lo, prim_hi, sorted_prims, n, k, left, right, parent, bb_lo, bb_hi, flags);
}

// The same boxes with the walk kept ON CHIP (round 2). The walk above pays two device-wide fences, one L2 atomic and three dependent
// L2 reads per level and node: 0.94 of 3.9 ms at 10 M triangles. Every internal node of the Karras hierarchy covers a contiguous
// range [first, last] of the sorted leaves and its own index lies inside that range, so a CTA that owns the leaves [c0, c1) also
// owns — exclusively — every internal node whose range lies inside [c0, c1): their arrival words and the boxes the siblings hand
// over live in shared memory (index - c0), ordered by block-level fences. A walk that reaches a node whose range crosses a chunk
// border (about 3 % of the nodes) is parked in a list and finished by a second kernel with the global protocol above — both
// children of such a node agree on that because the choice depends on the parent alone. (Finishing those walks inside the first
// kernel kept every CTA resident for its slowest 20-level global walk: 0.85 ms instead of 0.94.) min / max are exact: the boxes
// are the same bits as before.
#ifndef SNK_REFIT_CHUNKED
#define SNK_REFIT_CHUNKED 1
#endif
constexpr int RB_CHUNK = 512;
__global__ void __launch_bounds__(RB_CHUNK, 4) k_refit_binary_chunked(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
                                                                      const uint32_t* __restrict__ sorted_prims, int n,
                                                                      const int* __restrict__ parent, const int* __restrict__ first,
                                                                      const int* __restrict__ last, float4* bb_lo, float4* bb_hi,
                                                                      uint32_t* __restrict__ pending, uint32_t* pending_count) {
  // local ids: internal node i -> i - c0, leaf k -> RB_CHUNK + k - c0
  __shared__ float s_box[6][2 * RB_CHUNK];
  __shared__ uint32_t s_flag[RB_CHUNK];
  __shared__ uint32_t s_npend, s_base;
  const int c0 = blockIdx.x * RB_CHUNK, c1 = min(c0 + RB_CHUNK, n);
  s_flag[threadIdx.x] = 0u;
  if (threadIdx.x == 0) s_npend = 0u;
  __syncthreads();
  const int k = c0 + (int)threadIdx.x;
  int node = -1;  // >= 0 afterwards: this thread's walk goes on at parent[node], outside the chunk
  if (k < n) {
    const uint32_t p = sorted_prims[k];
    float4 lo = prim_lo[p], hi = prim_hi[p];
    int my = RB_CHUNK + (int)threadIdx.x;
    node = n - 1 + k;
    bb_lo[node] = lo; bb_hi[node] = hi;
    int par = parent[node];
    while (par >= 0) {
      if (first[par] < c0 || last[par] >= c1) break;
      const int j = par - c0;
      s_box[0][my] = lo.x; s_box[1][my] = lo.y; s_box[2][my] = lo.z; s_box[3][my] = hi.x; s_box[4][my] = hi.y; s_box[5][my] = hi.z;
      __threadfence_block();
      const uint32_t seen = atomicExch(&s_flag[j], (uint32_t)my + 1u);
      if (seen == 0u) { par = -1; break; }  // the sibling will find this box
      __threadfence_block();
      const int o = (int)seen - 1;
      lo = make_float4(fminf(lo.x, s_box[0][o]), fminf(lo.y, s_box[1][o]), fminf(lo.z, s_box[2][o]), 0.f);
      hi = make_float4(fmaxf(hi.x, s_box[3][o]), fmaxf(hi.y, s_box[4][o]), fmaxf(hi.z, s_box[5][o]), 0.f);
      bb_lo[par] = lo; bb_hi[par] = hi;
      node = par; my = j;
      par = parent[node];
    }
    if (par < 0) node = -1;  // handed over to the sibling, or the root is done
  }
  uint32_t slot = 0;
  if (node >= 0) slot = atomicAdd(&s_npend, 1u);
  __syncthreads();
  if (threadIdx.x == 0 && s_npend) s_base = atomicAdd(pending_count, s_npend);
  __syncthreads();
  if (node >= 0) pending[s_base + slot] = (uint32_t)node;
}
// the parked walks: node's box is in bb_lo / bb_hi (written by the kernel before), its parent is shared with another chunk
__global__ void __launch_bounds__(256) k_refit_binary_top(const uint32_t* __restrict__ pending, const uint32_t* __restrict__ pending_count,
                                                          const int* __restrict__ parent, float4* bb_lo, float4* bb_hi, uint32_t* flags) {
  const uint32_t count = *pending_count;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    int node = (int)pending[i];
    float4 lo = bb_lo[node], hi = bb_hi[node];
    int par = parent[node];
    while (par >= 0) {
      // the first child to arrive leaves its index (+1) and stops; the second one finds its sibling there and continues
      const uint32_t seen = atomicExch(&flags[par], (uint32_t)node + 1u);
      if (seen == 0u) break;
      __threadfence();
      const int other = (int)seen - 1;
      const float4 olo = __ldcg(&bb_lo[other]), ohi = __ldcg(&bb_hi[other]);
      lo = make_float4(fminf(lo.x, olo.x), fminf(lo.y, olo.y), fminf(lo.z, olo.z), 0.f);
      hi = make_float4(fmaxf(hi.x, ohi.x), fmaxf(hi.y, ohi.y), fmaxf(hi.z, ohi.z), 0.f);
      bb_lo[par] = lo; bb_hi[par] = hi;
      __threadfence();
      node = par;
      par = parent[node];
    }
  }
}

// ---- collapse to 8-wide nodes ---------------------------------------------------------------
struct CollapseArgs {
  int n;
  const int *left, *right, *first, *last;
  const float4 *bb_lo, *bb_hi;
  const uint32_t* sorted_prims;
  uint32_t* queue;
  uint32_t* counters;
  SnkNode8* nodes;
  uint32_t node_offset, node_capacity;
  SnkTri* tris;
  uint32_t* leaf_ids;
  const uint32_t* leaf_remap;
  uint32_t leaf_offset;
  const float* positions;
  const uint32_t* indices;
  uint32_t index_offset, vertex_offset;
  int leaf_max;
  int self_reset;  // single-CTA builds: leave counters / queue[0] as k_init would (the TLAS is rebuilt every frame: one launch less)
};

__device__ __forceinline__ float half_area(float4 lo, float4 hi) {
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ int node_count(const CollapseArgs& a, int node) {
  return node >= a.n - 1 ? 1 : (a.last[node] - a.first[node] + 1);
}

__device__ void emit_leaf_records(const CollapseArgs& a, int node, uint32_t dst) {
  int f, l;
  if (node >= a.n - 1) { f = l = node - (a.n - 1); }
  else { f = a.first[node]; l = a.last[node]; }
  for (int k = f; k <= l; ++k, ++dst) {
    const uint32_t prim = a.sorted_prims[k];
    if (a.tris) {
      const uint32_t* ip = a.indices + a.index_offset + 3ull * prim;
      const uint32_t ia = ip[0] + a.vertex_offset, ib = ip[1] + a.vertex_offset, ic = ip[2] + a.vertex_offset;
      float4* t = reinterpret_cast<float4*>(a.tris + dst);
      t[0] = make_float4(a.positions[3ull * ia], a.positions[3ull * ia + 1], a.positions[3ull * ia + 2], __uint_as_float(prim));
      t[1] = make_float4(a.positions[3ull * ib], a.positions[3ull * ib + 1], a.positions[3ull * ib + 2], 0.f);
      t[2] = make_float4(a.positions[3ull * ic], a.positions[3ull * ic + 1], a.positions[3ull * ic + 2], 0.f);
    } else {
      a.leaf_ids[dst] = a.leaf_remap ? a.leaf_remap[prim] : prim;
    }
  }
}

// quantisation exponent: smallest e with 2^e * 255 >= extent (biased, as a float exponent byte)
__device__ __forceinline__ int quant_exp(float extent) {
  if (!(extent > 0.f)) return 1;
  int ex;
  frexpf(extent / 255.0f, &ex);  // extent/255 = m * 2^ex, m in [0.5,1)  ->  2^ex >= extent/255
  int be = ex + 127;
  return be < 1 ? 1 : (be > 254 ? 254 : be);
}

// conservative 8-bit quantisation of one child box against the node grid (origin nlo, biased exponents be):
// floor / ceil, then nudged until the planes as the traversal dequantises them really enclose the child
__device__ __forceinline__ void quantise_child(const float4 nlo, const int* be, const float4 clo, const float4 chi, uint8_t* ql, uint8_t* qh) {
  const float base3[3] = {nlo.x, nlo.y, nlo.z};
  const float lo3[3] = {clo.x, clo.y, clo.z}, hi3[3] = {chi.x, chi.y, chi.z};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float scale = __uint_as_float((uint32_t)be[k] << 23), inv = __uint_as_float((uint32_t)(254 - be[k]) << 23);
    float fl = floorf((lo3[k] - base3[k]) * inv), fh = ceilf((hi3[k] - base3[k]) * inv);
    fl = fminf(fmaxf(fl, 0.f), 255.f); fh = fminf(fmaxf(fh, 0.f), 255.f);
    if (fl > 0.f && base3[k] + fl * scale > lo3[k]) fl -= 1.f;
    if (fh < 255.f && base3[k] + fh * scale < hi3[k]) fh += 1.f;
    ql[k] = (uint8_t)fl; qh[k] = (uint8_t)fh;
  }
}

__device__ void process_wide_node(const CollapseArgs& a, uint32_t w, int b) {
  // ---- choose up to 8 children by repeatedly opening the largest one
  int ch[8];
  float area[8];
  int cnt[8];
  int m = 0;
  if (a.n == 1) {
    ch[0] = 0; cnt[0] = 1; area[0] = 0.f; m = 1;  // single leaf (node index n-1+0 = 0)
  } else {
    const int l = a.left[b], r = a.right[b];
    ch[0] = l; ch[1] = r; m = 2;
    cnt[0] = node_count(a, l); cnt[1] = node_count(a, r);
    area[0] = half_area(a.bb_lo[l], a.bb_hi[l]); area[1] = half_area(a.bb_lo[r], a.bb_hi[r]);
    for (int phase = 0; phase < 2; ++phase) {
      const int min_open = phase == 0 ? a.leaf_max + 1 : 2;
      while (m < 8) {
        int best = -1; float barea = -1.f;
        for (int i = 0; i < m; ++i)
          if (cnt[i] >= min_open && area[i] > barea) { barea = area[i]; best = i; }
        if (best < 0) break;
        const int c = ch[best];
        const int cl = a.left[c], cr = a.right[c];
        ch[best] = cl; cnt[best] = node_count(a, cl); area[best] = half_area(a.bb_lo[cl], a.bb_hi[cl]);
        ch[m] = cr; cnt[m] = node_count(a, cr); area[m] = half_area(a.bb_lo[cr], a.bb_hi[cr]);
        ++m;
      }
    }
  }
  // ---- node box and child boxes
  float4 lo[8], hi[8];
  float4 nlo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0), nhi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0);
  const int leaf_base = a.n - 1;
  for (int i = 0; i < m; ++i) {
    const int node = (a.n == 1) ? 0 : ch[i];
    lo[i] = a.bb_lo[node]; hi[i] = a.bb_hi[node];
    nlo = make_float4(fminf(nlo.x, lo[i].x), fminf(nlo.y, lo[i].y), fminf(nlo.z, lo[i].z), 0);
    nhi = make_float4(fmaxf(nhi.x, hi[i].x), fmaxf(nhi.y, hi[i].y), fmaxf(nhi.z, hi[i].z), 0);
  }
  (void)leaf_base;
  // ---- octant-ordered slot assignment (greedy on the largest remaining |cost|)
  // slot s is visited first by rays whose direction is negative exactly on the axes whose bit is set in s,
  // so it should hold the child lying furthest towards +axis on those axes and -axis on the others.
  int slot_of[8];
  {
    const float ccx = 0.5f * (nlo.x + nhi.x), ccy = 0.5f * (nlo.y + nhi.y), ccz = 0.5f * (nlo.z + nhi.z);
    float cost[8][8];
    for (int i = 0; i < m; ++i) {
      const float dx = 0.5f * (lo[i].x + hi[i].x) - ccx, dy = 0.5f * (lo[i].y + hi[i].y) - ccy, dz = 0.5f * (lo[i].z + hi[i].z) - ccz;
      for (int s = 0; s < 8; ++s) cost[i][s] = ((s & 1) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 4) ? dz : -dz);
    }
    uint32_t used_slots = 0, used_children = 0;
    for (int k = 0; k < m; ++k) {
      float best = -FLT_MAX; int bi = 0, bs = 0;
      for (int i = 0; i < m; ++i) {
        if (used_children & (1u << i)) continue;
        for (int s = 0; s < 8; ++s) {
          if (used_slots & (1u << s)) continue;
          if (cost[i][s] > best) { best = cost[i][s]; bi = i; bs = s; }
        }
      }
      slot_of[bi] = bs; used_children |= 1u << bi; used_slots |= 1u << bs;
    }
  }
  int child_in_slot[8];
  for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
  for (int i = 0; i < m; ++i) child_in_slot[slot_of[i]] = i;
  // ---- allocate inner children + leaf records
  int n_inner = 0, n_leafrec = 0;
  for (int i = 0; i < m; ++i) {
    if (cnt[i] > a.leaf_max) ++n_inner; else n_leafrec += cnt[i];
  }
  uint32_t child_base = 0, rec_base = 0;
  if (n_inner) child_base = atomicAdd(&a.counters[C_NODES], (uint32_t)n_inner);
  if (n_leafrec) rec_base = atomicAdd(&a.counters[C_LEAVES], (uint32_t)n_leafrec);
  if (child_base + n_inner > a.node_capacity) {  // cannot happen with capacity >= n; flag and stop growing
    atomicExch(&a.counters[C_OVERFLOW], 1u);
    n_inner = 0;
  }
  // ---- quantise
  SnkNode8 nd;
  nd.px = nlo.x; nd.py = nlo.y; nd.pz = nlo.z;
  int be[3] = {quant_exp(nhi.x - nlo.x), quant_exp(nhi.y - nlo.y), quant_exp(nhi.z - nlo.z)};
  nd.imask = 0;
  nd.child_base = a.node_offset + child_base;
  nd.tri_base = a.leaf_offset + rec_base;
  uint32_t inner_rank = 0, rec_off = 0;
  for (int s = 0; s < 8; ++s) {
    const int i = child_in_slot[s];
    if (i < 0) {
      nd.meta[s] = 0;
      nd.qlox[s] = nd.qloy[s] = nd.qloz[s] = 255; nd.qhix[s] = nd.qhiy[s] = nd.qhiz[s] = 0;  // inverted: never hit
      continue;
    }
    uint8_t ql[3], qh[3];
    quantise_child(nlo, be, lo[i], hi[i], ql, qh);
    nd.qlox[s] = ql[0]; nd.qloy[s] = ql[1]; nd.qloz[s] = ql[2];
    nd.qhix[s] = qh[0]; nd.qhiy[s] = qh[1]; nd.qhiz[s] = qh[2];
    if (cnt[i] > a.leaf_max) {
      if (inner_rank < (uint32_t)n_inner) {
        nd.imask |= (uint8_t)(1u << s);
        nd.meta[s] = (uint8_t)((1u << 5) | (24u + s));
        // work item of the child wide node (read by the next level, after the grid-wide barrier)
        a.queue[child_base + inner_rank] = (uint32_t)ch[i];
        ++inner_rank;
      } else {
        nd.meta[s] = 0;
      }
    } else {
      const uint32_t unary = cnt[i] == 1 ? 1u : (cnt[i] == 2 ? 3u : 7u);
      nd.meta[s] = (uint8_t)((unary << 5) | rec_off);
      emit_leaf_records(a, (a.n == 1) ? 0 : ch[i], a.leaf_offset + rec_base + rec_off);
      rec_off += cnt[i];
    }
  }
  nd.ex = (uint8_t)be[0]; nd.ey = (uint8_t)be[1]; nd.ez = (uint8_t)be[2];
  // 5 x 128-bit stores
  const uint4* src = reinterpret_cast<const uint4*>(&nd);
  uint4* dst = reinterpret_cast<uint4*>(a.nodes + a.node_offset + w);
#pragma unroll
  for (int k = 0; k < 5; ++k) dst[k] = src[k];
  if (n_leafrec) {
    __threadfence();
    atomicAdd(&a.counters[C_DONE], (uint32_t)n_leafrec);
  }
}

// The same node construction by EIGHT lanes (an aligned octet of a warp, called by whole warps: four consecutive nodes w per warp): lane i owns child i
// while the children are chosen and placed, lane s owns slot s when the node is written. The one-thread version above is a
// ~3000-instruction serial chain with its 8x8 cost matrix in local memory, and because the collapse is level-synchronous its
// LATENCY (x tree depth), not its throughput, set the build time: k_collapse was 1.76 of 2.96 ms at 3 M triangles
// (profiles/r02y_build_breakdown.txt). Same children, same slots, same records as process_wide_node.
__device__ void process_wide_node8(const CollapseArgs& a, uint32_t w, const bool active) {
  const uint32_t lane = threadIdx.x & 31u, sub = lane & 7u, gshift = lane & 24u, gm = 0xFFu << gshift;
  int ch = 0, cnt = 0;
  float area = -1.f;
  float4 lo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), hi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
  int m = 1;
  if (!active) {
    m = 0;  // an octet past the end of the level: no children, takes part in the warp-wide allocation only
  } else if (a.n == 1) {
    if (sub == 0) { cnt = 1; lo = a.bb_lo[0]; hi = a.bb_hi[0]; area = 0.f; }  // single leaf (node index n-1+0 = 0)
  } else {
    const int b = (int)a.queue[w];
    m = 2;
    // a lane also fetches the children of its child, so that opening it later costs one memory round trip, not two
    int kid_l = 0, kid_r = 0;
    if (sub < 2u) {
      ch = sub ? a.right[b] : a.left[b];
      if (ch < a.n - 1) { kid_l = a.left[ch]; kid_r = a.right[ch]; }
      cnt = node_count(a, ch); lo = a.bb_lo[ch]; hi = a.bb_hi[ch]; area = half_area(lo, hi);
    }
    // ---- choose up to 8 children by repeatedly opening the largest one
    for (int phase = 0; phase < 2; ++phase) {
      const int min_open = phase == 0 ? a.leaf_max + 1 : 2;
      while (m < 8) {
        float key = ((int)sub < m && cnt >= min_open && area > -1.f) ? area : -1.f;
        int bi = (int)sub;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          const float ok = __shfl_xor_sync(gm, key, o, 8);
          const int oi = __shfl_xor_sync(gm, bi, o, 8);
          if (ok > key || (ok == key && oi < bi)) { key = ok; bi = oi; }
        }
        if (key < 0.f) break;
        const int cl = __shfl_sync(gm, kid_l, bi, 8), cr = __shfl_sync(gm, kid_r, bi, 8);
        if ((int)sub == bi || (int)sub == m) {
          ch = (int)sub == bi ? cl : cr;
          if (ch < a.n - 1) { kid_l = a.left[ch]; kid_r = a.right[ch]; }
          cnt = node_count(a, ch); lo = a.bb_lo[ch]; hi = a.bb_hi[ch]; area = half_area(lo, hi);
        }
        ++m;
      }
    }
  }
  const bool valid = (int)sub < m;
  // ---- node box
  float4 nlo = lo, nhi = hi;  // lanes without a child hold the empty box
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    nlo.x = fminf(nlo.x, __shfl_xor_sync(gm, nlo.x, o, 8)); nlo.y = fminf(nlo.y, __shfl_xor_sync(gm, nlo.y, o, 8));
    nlo.z = fminf(nlo.z, __shfl_xor_sync(gm, nlo.z, o, 8));
    nhi.x = fmaxf(nhi.x, __shfl_xor_sync(gm, nhi.x, o, 8)); nhi.y = fmaxf(nhi.y, __shfl_xor_sync(gm, nhi.y, o, 8));
    nhi.z = fmaxf(nhi.z, __shfl_xor_sync(gm, nhi.z, o, 8));
  }
  // ---- octant-ordered slot assignment (greedy on the largest remaining cost; ties: lowest child, then lowest slot)
  int my_slot = -1;
  {
    const float ccx = 0.5f * (nlo.x + nhi.x), ccy = 0.5f * (nlo.y + nhi.y), ccz = 0.5f * (nlo.z + nhi.z);
    const float dx = 0.5f * (lo.x + hi.x) - ccx, dy = 0.5f * (lo.y + hi.y) - ccy, dz = 0.5f * (lo.z + hi.z) - ccz;
    float cost[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) cost[s] = ((s & 1) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 4) ? dz : -dz);
    uint32_t used_slots = 0;
    for (int k = 0; k < m; ++k) {
      // this lane's best free slot; a placed or childless lane bids -inf, a child with NaN costs bids -FLT_MAX on the first free slot
      float lb = -FLT_MAX;
      int ls = __ffs(~used_slots) - 1;
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (!((used_slots >> s) & 1u) && cost[s] > lb) { lb = cost[s]; ls = s; }
      if (!valid || my_slot >= 0) lb = -INFINITY;
      int bi = (int)sub;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const float ob = __shfl_xor_sync(gm, lb, o, 8);
        const int oi = __shfl_xor_sync(gm, bi, o, 8), os = __shfl_xor_sync(gm, ls, o, 8);
        if (ob > lb || (ob == lb && oi < bi)) { lb = ob; bi = oi; ls = os; }
      }
      if ((int)sub == bi) my_slot = ls;
      used_slots |= 1u << ls;
    }
  }
  // ---- quantise (still one child per lane)
  int be[3] = {quant_exp(nhi.x - nlo.x), quant_exp(nhi.y - nlo.y), quant_exp(nhi.z - nlo.z)};
  uint32_t qa = 0, qb = 0;
  if (valid) {
    uint8_t ql[3], qh[3];
    quantise_child(nlo, be, lo, hi, ql, qh);
    qa = (uint32_t)ql[0] | ((uint32_t)ql[1] << 8) | ((uint32_t)ql[2] << 16);
    qb = (uint32_t)qh[0] | ((uint32_t)qh[1] << 8) | ((uint32_t)qh[2] << 16);
  }
  // ---- transpose: lane s takes the child placed in slot s
  int src = -1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int si = __shfl_sync(gm, my_slot, i, 8);
    if (si == (int)sub) src = i;
  }
  const bool occ = src >= 0;
  const int from = occ ? src : 0;
  const int s_ch = __shfl_sync(gm, ch, from, 8), s_cnt = __shfl_sync(gm, cnt, from, 8);
  uint32_t s_qa = __shfl_sync(gm, qa, from, 8), s_qb = __shfl_sync(gm, qb, from, 8);
  if (!occ) { s_qa = 0x00FFFFFFu; s_qb = 0u; }  // inverted: never hit
  // ---- allocate inner children + leaf records (ranks in slot order)
  const bool is_inner = occ && s_cnt > a.leaf_max, is_leaf = occ && !is_inner;
  uint32_t inner_mask = (__ballot_sync(gm, is_inner) >> gshift) & 0xFFu;
  const uint32_t c1 = (__ballot_sync(gm, is_leaf && (s_cnt & 1)) >> gshift) & 0xFFu;
  const uint32_t c2 = (__ballot_sync(gm, is_leaf && (s_cnt & 2)) >> gshift) & 0xFFu;
  int n_inner = __popc(inner_mask);
  const int n_leafrec = __popc(c1) + 2 * __popc(c2);
  // one atomic per WARP and counter: every wide node bumping the same two words made the L2 atomic unit the bottleneck of the
  // whole build (same-address atomics retire at ~1 per clock: 3 x 1.9 M of them = 2.6 of the 3.0 ms of k_collapse at 10 M triangles)
  uint32_t child_base = 0, rec_base = 0;
  {
    uint32_t before_i = 0, before_l = 0, total_i = 0, total_l = 0;
#pragma unroll
    for (uint32_t g = 0; g < 4; ++g) {
      const uint32_t gi = __shfl_sync(0xFFFFFFFFu, (uint32_t)n_inner, g * 8), gl = __shfl_sync(0xFFFFFFFFu, (uint32_t)n_leafrec, g * 8);
      if (g * 8 < gshift) { before_i += gi; before_l += gl; }
      total_i += gi; total_l += gl;
    }
    if (lane == 0u) {
      if (total_i) child_base = atomicAdd(&a.counters[C_NODES], total_i);
      if (total_l) rec_base = atomicAdd(&a.counters[C_LEAVES], total_l);
    }
    child_base = __shfl_sync(0xFFFFFFFFu, child_base, 0) + before_i;
    rec_base = __shfl_sync(0xFFFFFFFFu, rec_base, 0) + before_l;
  }
  if (!active) return;
  if (child_base + n_inner > a.node_capacity) {  // cannot happen with capacity >= n; flag and stop growing
    if (sub == 0u) atomicExch(&a.counters[C_OVERFLOW], 1u);
    n_inner = 0; inner_mask = 0;
  }
  const uint32_t below = (1u << sub) - 1u;
  uint32_t meta = 0;
  if (is_inner) {
    if ((inner_mask >> sub) & 1u) {
      meta = (1u << 5) | (24u + sub);
      // work item of the child wide node (read by the next level, after the grid-wide barrier)
      a.queue[child_base + __popc(inner_mask & below)] = (uint32_t)s_ch;
    }
  } else if (is_leaf) {
    const uint32_t rec_off = __popc(c1 & below) + 2u * __popc(c2 & below);
    const uint32_t unary = s_cnt == 1 ? 1u : (s_cnt == 2 ? 3u : 7u);
    meta = (unary << 5) | rec_off;
    emit_leaf_records(a, s_ch, a.leaf_offset + rec_base + rec_off);
  }
  // ---- write the node: lane 0 the 24-byte header, every lane the 7 bytes of its slot
  uint8_t* np = reinterpret_cast<uint8_t*>(a.nodes + a.node_offset + w);
  if (sub == 0u) {
    const uint32_t ebits = (uint32_t)be[0] | ((uint32_t)be[1] << 8) | ((uint32_t)be[2] << 16) | (inner_mask << 24);
    *reinterpret_cast<float4*>(np) = make_float4(nlo.x, nlo.y, nlo.z, __uint_as_float(ebits));
    *reinterpret_cast<uint2*>(np + 16) = make_uint2(a.node_offset + child_base, a.leaf_offset + rec_base);
  }
  np[24 + sub] = (uint8_t)meta;
  np[32 + sub] = (uint8_t)s_qa; np[40 + sub] = (uint8_t)(s_qa >> 8); np[48 + sub] = (uint8_t)(s_qa >> 16);
  np[56 + sub] = (uint8_t)s_qb; np[64 + sub] = (uint8_t)(s_qb >> 8); np[72 + sub] = (uint8_t)(s_qb >> 16);
}

#ifndef SNK_COLLAPSE_OCTET
#define SNK_COLLAPSE_OCTET 1  // 1: eight lanes per wide node (process_wide_node8); 0: one thread per node
#endif
// nodes [begin, end) of one level by `nthreads` threads of which this is number `tid` (whole warps call this together)
__device__ __forceinline__ void process_level(const CollapseArgs& a, uint32_t begin, uint32_t end, uint32_t tid, uint32_t nthreads) {
#if SNK_COLLAPSE_OCTET
  for (uint32_t w0 = begin + ((tid & ~31u) >> 3); w0 < end; w0 += nthreads >> 3) {  // warp-uniform trip count
    const uint32_t w = w0 + ((tid & 31u) >> 3);
    process_wide_node8(a, w, w < end);
  }
#else
  for (uint32_t w = begin + tid; w < end; w += nthreads) process_wide_node(a, w, (int)a.queue[w]);
#endif
}

// Level-synchronous collapse in ONE cooperative launch: the wide nodes of a level occupy a contiguous index range
// (children are allocated with an atomic bump, so level L+1 = everything allocated while level L was processed);
// all threads process the level with converged warps, then the grid synchronises and moves to the next range.
// Earlier versions let thread w spin until its parent published node w: lanes of a warp received their work at
// different times and ran the (long, serial) node construction one after the other — 2.5-2.9 of 4.2 ms at 3 M triangles.
#ifndef SNK_COLLAPSE_CTAS
#define SNK_COLLAPSE_CTAS 10  // 8 / 10 / 12 / 16 CTAs per SM: 4.48 / 4.48 / 4.53 / 4.63 ms at 10 M triangles (profiles/r03i_variants_collapse_occupancy.txt)
#endif
__global__ void __launch_bounds__(128, SNK_COLLAPSE_CTAS) k_collapse(CollapseArgs a) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  uint32_t begin = 0, end = 1;
  while (begin < end) {
    process_level(a, begin, end, gtid, gsize);
    grid.sync();
    begin = end;
    const uint32_t total = *((volatile uint32_t*)&a.counters[C_NODES]);
    end = total < a.node_capacity ? total : a.node_capacity;
    grid.sync();  // nobody may bump C_NODES for the next level before every thread has read this level's end
  }
}

// ---- whole build in ONE CTA for small inputs (the per-frame TLAS: TlasSystem::BuildTlas rebuilds <= 4096 instances every
// frame, TlasSystem.cpp:27-49). The general path above needs 37 launches (8 radix passes x 4 + Morton, Karras, refit,
// collapse ...) and was launch-latency bound: 0.38 ms for the 205 instances of C2. Here: Morton codes -> bitonic sort of
// (key, index) pairs in shared memory -> Karras hierarchy -> bottom-up boxes -> level-synchronous collapse, with
// __syncthreads() where the general path has kernel boundaries or grid syncs.
#ifndef SNK_SMALL_PLOC
#define SNK_SMALL_PLOC 3  // r02p: PLOC for every small build: nodes per ray 15.0 -> 14.2 but triangles per ray 4.0 -> 4.6, C2 +2 %. r04p (with instance
                          // culling the TLAS is most of a ray's nodes): PLOC for the TLAS only: C4 (2344 instances) 32.3 -> 30.7 ms, C2 (205) 5.71 -> 5.67 ms
                          // with a 0.1 ms slower rebuild: adopted for TLASes of >= 1024 instances (3)
#endif
constexpr int SMALL_MAX = 4096, SMALL_THREADS = 1024;
constexpr int SMALL_TREE_MAX = 1024;  // largest single-CTA build whose binary hierarchy lives in shared memory
constexpr size_t SMALL_TREE_SMEM = (size_t)SMALL_TREE_MAX * (7 * 4 + 2 * 2 * 16);  // 7 int arrays (parent twice) + 2 x 2n float4 boxes = 92 KB
#ifdef SNK_SMALL_TIMING  // phase stamps of the single-CTA builder (debug builds only)
#define SNK_STAMP(k) if (threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(snk_stamp[k])); }
#define SNK_STAMP_PRINT() if (threadIdx.x == 0 && n > 100) printf("k_build_small n=%d: keys %llu sort %llu karras %llu boxes %llu collapse %llu ns\n", n, snk_stamp[1]-snk_stamp[0], snk_stamp[2]-snk_stamp[1], snk_stamp[3]-snk_stamp[2], snk_stamp[4]-snk_stamp[3], snk_stamp[5]-snk_stamp[4]);
#else
#define SNK_STAMP(k)
#define SNK_STAMP_PRINT()
#endif
__global__ void __launch_bounds__(SMALL_THREADS) k_build_small(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi, int n,
                                                               uint32_t* counters, unsigned long long* keys_g, uint32_t* vals_g, int* left,
                                                               int* right, int* parent, int* first, int* last, float4* bb_lo, float4* bb_hi,
                                                               uint32_t* flags, CollapseArgs a, int tree_in_smem) {
  __shared__ unsigned long long sk[SMALL_MAX];
  __shared__ uint32_t sv[SMALL_MAX];
#ifdef SNK_SMALL_TIMING
  unsigned long long snk_stamp[6];
#endif
  const int t = threadIdx.x, NT = blockDim.x;  // launched with as many threads as there are elements to sort (128..1024)
  SNK_STAMP(0);
  // Builds of <= SMALL_TREE_MAX primitives (the per-frame TLAS of a few hundred instances) keep the whole binary hierarchy — child /
  // range / parent indices, node boxes, the collapse's work queue — in dynamic shared memory: the collapse opens up to seven
  // children per wide node one after the other, each step a dependent read of two nodes, and three levels of that through the
  // L2 were 40 of the kernel's 63 us on C2's 205 instances (phase stamps, SNK_SMALL_TIMING). Generic pointers: the device
  // functions below do not care which memory they walk.
  extern __shared__ __align__(16) unsigned char small_dyn[];
  float4 *const root_lo = bb_lo, *const root_hi = bb_hi;  // the host reads the root box of a BLAS from the global arrays
  if (tree_in_smem) {
    int* si = reinterpret_cast<int*>(small_dyn);
    left = si; right = si + SMALL_TREE_MAX; first = si + 2 * SMALL_TREE_MAX; last = si + 3 * SMALL_TREE_MAX; parent = si + 4 * SMALL_TREE_MAX;  // 2 n entries
    a.queue = reinterpret_cast<uint32_t*>(si + 6 * SMALL_TREE_MAX);
    bb_lo = reinterpret_cast<float4*>(si + 7 * SMALL_TREE_MAX);
    bb_hi = bb_lo + 2 * SMALL_TREE_MAX;
    a.left = left; a.right = right; a.first = first; a.last = last; a.bb_lo = bb_lo; a.bb_hi = bb_hi;
    if (t == 0) a.queue[0] = 0u;  // wide node 0 collapses the binary root (k_init does this for the global queue)
  }
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = t; i < np2; i += NT) {
    sk[i] = i < n ? morton_key(prim_lo, prim_hi, (uint32_t)i, counters) : ~0ull;
    sv[i] = (uint32_t)i;
  }
  for (int i = t; i < n; i += NT) flags[i] = 0u;
  __syncthreads();
  SNK_STAMP(1);
  // bitonic sort by (key, index): the index breaks ties the way the stable radix sort of the general path does
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < np2; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long ka = sk[i], kb = sk[ixj];
          const uint32_t va = sv[i], vb = sv[ixj];
          const bool a_gt_b = ka > kb || (ka == kb && va > vb);
          const bool up = (i & k) == 0;
          if (a_gt_b == up) { sk[i] = kb; sk[ixj] = ka; sv[i] = vb; sv[ixj] = va; }
        }
      }
      __syncthreads();
    }
  }
  SNK_STAMP(2);
  for (int i = t; i < n; i += NT) { keys_g[i] = sk[i]; vals_g[i] = sv[i]; }
  if (n > 1) {
    for (int i = t; i < n - 1; i += NT) karras_node(sk, n, i, left, right, parent, first, last);
  } else if (t == 0) {
    parent[0] = -1;  // the only leaf is the root
  }
  __threadfence();
  __syncthreads();
  SNK_STAMP(3);
  // Boxes of the binary nodes WITHOUT the bottom-up walk: every internal node covers a contiguous range [first, last] of the
  // sorted leaves, so its box is the union over that range — whole 32-leaf chunks through a chunk box, the ragged ends leaf by
  // leaf (<= 62 + n / 32 reads, all independent). The bottom-up refit of the multi-CTA path (refit_from_leaf: an L2 atomic, two
  // fences and three dependent loads per level) was ~35 of this kernel's 63 us on the 205 instances of C2: the depth of the tree
  // times the L2 latency. min / max are exact, so the boxes are the same bits.
  float* chunk = reinterpret_cast<float*>(sk);  // the keys are dead: 6 floats per chunk of 32 sorted leaves (<= 128 chunks)
  {
    const int lane = t & 31, nchunks = (n + 31) >> 5;
    for (int c = t >> 5; c < nchunks; c += NT >> 5) {
      const int k = c * 32 + lane;
      float4 lo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), hi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
      if (k < n) {
        const uint32_t p = sv[k];
        lo = prim_lo[p]; hi = prim_hi[p];
        bb_lo[n - 1 + k] = lo; bb_hi[n - 1 + k] = hi;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
      }
      if (lane == 0) { float* cb = chunk + 6 * c; cb[0] = lo.x; cb[1] = lo.y; cb[2] = lo.z; cb[3] = hi.x; cb[4] = hi.y; cb[5] = hi.z; }
    }
  }
  __threadfence();
  __syncthreads();
  for (int i = t; i < n - 1; i += NT) {
    const int f = first[i], l = last[i];
    float4 lo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), hi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
    int k = f;
    while (k <= l) {
      if ((k & 31) == 0 && k + 31 <= l) {
        const float* cb = chunk + 6 * (k >> 5);
        lo.x = fminf(lo.x, cb[0]); lo.y = fminf(lo.y, cb[1]); lo.z = fminf(lo.z, cb[2]);
        hi.x = fmaxf(hi.x, cb[3]); hi.y = fmaxf(hi.y, cb[4]); hi.z = fmaxf(hi.z, cb[5]);
        k += 32;
      } else {
        const float4 a0 = bb_lo[n - 1 + k], a1 = bb_hi[n - 1 + k];  // written by this CTA before the barrier; never cached before that
        lo.x = fminf(lo.x, a0.x); lo.y = fminf(lo.y, a0.y); lo.z = fminf(lo.z, a0.z);
        hi.x = fmaxf(hi.x, a1.x); hi.y = fmaxf(hi.y, a1.y); hi.z = fmaxf(hi.z, a1.z);
        k += 1;
      }
    }
    bb_lo[i] = lo; bb_hi[i] = hi;
  }
  __threadfence();
  __syncthreads();
  if (tree_in_smem && t == 0) { root_lo[0] = bb_lo[0]; root_hi[0] = bb_hi[0]; }
  SNK_STAMP(4);
  // collapse, level by level (queue[0] = binary root and counters[C_NODES] = 1 were set by k_init)
  volatile uint32_t& s_end = *reinterpret_cast<volatile uint32_t*>(&sk[0]);  // the keys are dead by now (static smem is at its 48 KB limit)
  uint32_t begin = 0, end = 1;
  while (begin < end) {
    process_level(a, begin, end, (uint32_t)t, (uint32_t)NT);
    __threadfence();
    __syncthreads();
    if (t == 0) {
      const uint32_t total = *((volatile uint32_t*)&a.counters[C_NODES]);
      s_end = total < a.node_capacity ? total : a.node_capacity;
    }
    __syncthreads();
    begin = end;
    end = s_end;
    __syncthreads();
  }
  SNK_STAMP(5);
  SNK_STAMP_PRINT();
  if (a.self_reset) init_counters(a.counters, tree_in_smem ? a.counters + 15 : a.queue, t);  // (queue[0] of the global queue was never touched: a spare counter word takes the store)
}

// ---- the same single-CTA build with a better binary tree: PLOC (parallel locally-ordered clustering, Meister & Bittner 2018)
// over the Morton-sorted primitives instead of the Karras hierarchy. Every round each cluster finds the neighbour within
// PLOC_RADIUS positions whose union with it has the smallest area, mutual pairs merge, the list is compacted in order. The tree
// is close to a full-sweep SAH tree. Built to test whether the LBVH costs the instanced scenes node visits (both their TLAS and
// their <= 4096-triangle BLASes are built here): it does not pay on these scenes (see SNK_SMALL_PLOC), so it is an option. Internal ids are handed out from n-2 downwards by rank,
// so the last merge is node 0 (the root) and the build is deterministic. Afterwards the leaves are renumbered so that every
// subtree covers a contiguous range of the (re-ordered) primitive list, which is what the collapse expects.
#ifndef SNK_PLOC_RADIUS
#define SNK_PLOC_RADIUS 12
#endif
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* scr, uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = (blockDim.x + 31u) >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += x;
  }
  if (lane == 31) scr[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < nw ? scr[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= (uint32_t)o) w += x;
    }
    scr[lane] = w;
  }
  __syncthreads();
  const uint32_t base = wid ? scr[wid - 1] : 0u;
  total = scr[nw - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(SMALL_THREADS) k_build_small_ploc(const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi, int n,
                                                                    uint32_t* counters, uint32_t* vals_g, uint32_t* vals_out, int* left, int* right,
                                                                    int* parent, int* first, int* last, float4* bb_lo, float4* bb_hi, uint32_t* flags,
                                                                    CollapseArgs a) {
  __shared__ unsigned long long sk[SMALL_MAX];
  __shared__ uint32_t sv[SMALL_MAX];
#ifdef SNK_SMALL_TIMING
  unsigned long long snk_stamp[6];
  int snk_rounds = 0;
#endif
  const int t = threadIdx.x, NT = blockDim.x;
  SNK_STAMP(0);
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = t; i < np2; i += NT) {
    sk[i] = i < n ? morton_key(prim_lo, prim_hi, (uint32_t)i, counters) : ~0ull;
    sv[i] = (uint32_t)i;
  }
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < np2; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long ka = sk[i], kb = sk[ixj];
          const uint32_t va = sv[i], vb = sv[ixj];
          const bool a_gt_b = ka > kb || (ka == kb && va > vb);
          const bool up = (i & k) == 0;
          if (a_gt_b == up) { sk[i] = kb; sk[ixj] = ka; sv[i] = vb; sv[ixj] = va; }
        }
      }
      __syncthreads();
    }
  }
  SNK_STAMP(1);
  // leaves: node n-1+i = sorted primitive i
  for (int i = t; i < n; i += NT) {
    const uint32_t prim = sv[i];
    vals_g[i] = prim;
    bb_lo[n - 1 + i] = prim_lo[prim];
    bb_hi[n - 1 + i] = prim_hi[prim];
  }
  __syncthreads();
  if (n == 1) {
    if (t == 0) { parent[0] = -1; vals_out[0] = vals_g[0]; }
  } else {
    // shared memory from here on: two cluster lists in the key array, neighbour / flag / scan scratch in the value array
    uint32_t* cl = reinterpret_cast<uint32_t*>(sk);
    uint32_t* cl2 = cl + SMALL_MAX;
    uint16_t* nb = reinterpret_cast<uint16_t*>(sv);
    uint8_t* fl = reinterpret_cast<uint8_t*>(nb + SMALL_MAX);
    uint32_t* scr = reinterpret_cast<uint32_t*>(fl + SMALL_MAX);
    for (int i = t; i < n; i += NT) cl[i] = (uint32_t)(n - 1 + i);
    __syncthreads();
    int c = n, next = n - 2;
    auto node_size = [&](int node) { return node >= n - 1 ? 1 : last[node]; };
    SNK_STAMP(2);
    while (c > 1) {
#ifdef SNK_SMALL_TIMING
      ++snk_rounds;
#endif
      for (int i = t; i < c; i += NT) {
        const float4 lo = bb_lo[cl[i]], hi = bb_hi[cl[i]];
        const int j0 = i - SNK_PLOC_RADIUS < 0 ? 0 : i - SNK_PLOC_RADIUS, j1 = i + SNK_PLOC_RADIUS > c - 1 ? c - 1 : i + SNK_PLOC_RADIUS;
        float best_area = FLT_MAX;
        int best = -1;
        for (int j = j0; j <= j1; ++j) {
          if (j == i) continue;
          const float4 l2 = bb_lo[cl[j]], h2 = bb_hi[cl[j]];
          const float ar = half_area(make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.f),
                                     make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.f));
          if (ar < best_area) { best_area = ar; best = j; }
        }
        if (best < 0) best = i + 1 < c ? i + 1 : i - 1;  // non-finite boxes: keep the list shrinking
        nb[i] = (uint16_t)best;
      }
      __syncthreads();
      for (int i = t; i < c; i += NT) {
        const int j = nb[i];
        fl[i] = nb[j] == i ? (i < j ? 1 : 2) : 0;  // 1: merges with its partner here, 2: absorbed by the partner, 0: stays
      }
      __syncthreads();
      const int per = (c + NT - 1) / NT, lo_i = t * per < c ? t * per : c, hi_i = lo_i + per < c ? lo_i + per : c;
      uint32_t n_lead = 0, n_keep = 0;
      for (int i = lo_i; i < hi_i; ++i) { n_lead += fl[i] == 1; n_keep += fl[i] != 2; }
      uint32_t tot_lead, tot_keep;
      uint32_t ex_lead = block_excl_scan(n_lead, scr, tot_lead);
      uint32_t ex_keep = block_excl_scan(n_keep, scr, tot_keep);
      for (int i = lo_i; i < hi_i; ++i) {
        if (fl[i] == 1) {
          const int id = next - (int)(ex_lead++);
          const int ca = (int)cl[i], cb = (int)cl[nb[i]];
          left[id] = ca; right[id] = cb;
          parent[ca] = id; parent[cb] = id;
          const float4 l1 = bb_lo[ca], h1 = bb_hi[ca], l2 = bb_lo[cb], h2 = bb_hi[cb];
          bb_lo[id] = make_float4(fminf(l1.x, l2.x), fminf(l1.y, l2.y), fminf(l1.z, l2.z), 0.f);
          bb_hi[id] = make_float4(fmaxf(h1.x, h2.x), fmaxf(h1.y, h2.y), fmaxf(h1.z, h2.z), 0.f);
          last[id] = node_size(ca) + node_size(cb);  // subtree size for now; becomes the range end below
          cl2[ex_keep++] = (uint32_t)id;
        } else if (fl[i] == 0) {
          cl2[ex_keep++] = cl[i];
        }
      }
      next -= (int)tot_lead;
      c = (int)tot_keep;
      uint32_t* tmp = cl; cl = cl2; cl2 = tmp;
      __threadfence_block();
      __syncthreads();
    }
    SNK_STAMP(3);
    if (t == 0) parent[0] = -1;
    __syncthreads();
    // position of every leaf in a depth-first order: sum of the sizes of the left siblings on the way to the root
    for (int i = t; i < n; i += NT) {
      int node = n - 1 + i, pos = 0, p;
      while ((p = parent[node]) >= 0) {
        if (right[p] == node) pos += node_size(left[p]);
        node = p;
      }
      flags[i] = (uint32_t)pos;
    }
    __syncthreads();
    for (int i = t; i < n; i += NT) {
      const uint32_t pos = flags[i], prim = vals_g[i];
      vals_out[pos] = prim;
      bb_lo[n - 1 + pos] = prim_lo[prim];
      bb_hi[n - 1 + pos] = prim_hi[prim];
    }
    for (int k = t; k < n - 1; k += NT) {
      const int l = left[k], r = right[k];
      if (l >= n - 1) left[k] = n - 1 + (int)flags[l - (n - 1)];
      if (r >= n - 1) right[k] = n - 1 + (int)flags[r - (n - 1)];
    }
    __syncthreads();
    for (int k = t; k < n - 1; k += NT) {
      int node = k;
      while (node < n - 1) node = left[node];
      const int f = node - (n - 1), sz = last[k];
      first[k] = f;
      last[k] = f + sz - 1;
    }
  }
  __threadfence();
  __syncthreads();
  SNK_STAMP(4);
  volatile uint32_t& s_end = *reinterpret_cast<volatile uint32_t*>(&sv[SMALL_MAX - 1]);
  uint32_t begin = 0, end = 1;
  while (begin < end) {
    process_level(a, begin, end, (uint32_t)t, (uint32_t)NT);
    __threadfence();
    __syncthreads();
    if (t == 0) {
      const uint32_t total = *((volatile uint32_t*)&a.counters[C_NODES]);
      s_end = total < a.node_capacity ? total : a.node_capacity;
    }
    __syncthreads();
    begin = end;
    end = s_end;
    __syncthreads();
  }
  SNK_STAMP(5);
#ifdef SNK_SMALL_TIMING
  if (threadIdx.x == 0 && n > 100) printf("k_build_small_ploc n=%d: sort %llu leaves %llu ploc %llu (%d rounds) order %llu collapse %llu ns\n", n, snk_stamp[1]-snk_stamp[0], snk_stamp[2]-snk_stamp[1], snk_stamp[3]-snk_stamp[2], snk_rounds, snk_stamp[4]-snk_stamp[3], snk_stamp[5]-snk_stamp[4]);
#endif
  if (a.self_reset) init_counters(a.counters, a.queue, t);
}

// ---- topology-preserving refit of a built BVH8 (animated vertices) -------------------------------------------
// (no reference counterpart: the reference never updates a BLAS, TLAS.cpp:63 — BASELINE config "build + refit sweep")
// k_refit_parents  once per BLAS: parent index of every wide node
// k_refit_tris     rewrites the 48-B triangle records from the current vertex positions (record keeps its primitive id)
// k_refit_nodes    bottom-up: a thread starts at every node without inner children, recomputes the child boxes
//                  (leaf slots from their records, inner slots from the boxes their children left in node_box),
//                  re-quantises the node in place and climbs to the parent when it is the last child to arrive
__global__ void k_refit_parents(const SnkNode8* __restrict__ nodes, uint32_t node_root, uint32_t node_count, uint32_t* __restrict__ parent) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= node_count) return;
  const SnkNode8& nd = nodes[node_root + w];
  if (w == 0) parent[0] = EMPTY;
  const uint32_t n_inner = __popc((uint32_t)nd.imask);
  for (uint32_t k = 0; k < n_inner; ++k) parent[nd.child_base - node_root + k] = w;
}

__global__ void k_refit_tris(SnkTri* __restrict__ tris, uint32_t tri_base, uint32_t tri_count, const float* __restrict__ positions,
                             const uint32_t* __restrict__ indices, uint32_t index_offset, uint32_t vertex_offset) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tri_count) return;
  float4* t = reinterpret_cast<float4*>(tris + tri_base + i);
  const uint32_t prim = __float_as_uint(t[0].w);
  const uint32_t* ip = indices + index_offset + 3ull * prim;
  const uint32_t ia = ip[0] + vertex_offset, ib = ip[1] + vertex_offset, ic = ip[2] + vertex_offset;
  t[0] = make_float4(positions[3ull * ia], positions[3ull * ia + 1], positions[3ull * ia + 2], __uint_as_float(prim));
  t[1] = make_float4(positions[3ull * ib], positions[3ull * ib + 1], positions[3ull * ib + 2], 0.f);
  t[2] = make_float4(positions[3ull * ic], positions[3ull * ic + 1], positions[3ull * ic + 2], 0.f);
}

__global__ void k_refit_nodes(SnkNode8* __restrict__ nodes, uint32_t node_root, uint32_t node_count, const SnkTri* __restrict__ tris,
                              const uint32_t* __restrict__ parent, uint32_t* __restrict__ arrive, float4* __restrict__ node_box) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= node_count) return;
  if (nodes[node_root + w].imask != 0) return;  // only nodes without inner children start a climb
  while (true) {
    SnkNode8 nd = nodes[node_root + w];
    float4 lo[8], hi[8];
    float4 nlo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0), nhi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0);
    uint32_t inner_rank = 0;
    for (int s = 0; s < 8; ++s) {
      const uint32_t meta = nd.meta[s];
      if (meta == 0) continue;
      if ((meta & 0x18u) == 0x18u) {
        const uint32_t c = nd.child_base - node_root + inner_rank++;
        lo[s] = __ldcg(&node_box[2ull * c]); hi[s] = __ldcg(&node_box[2ull * c + 1]);
      } else {
        const uint32_t first = nd.tri_base + (meta & 31u), cnt = __popc(meta >> 5);
        float4 l = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0), h = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0);
        for (uint32_t k = 0; k < cnt; ++k) {
          const float4* t = reinterpret_cast<const float4*>(tris + first + k);
          for (int v = 0; v < 3; ++v) {
            const float4 p = t[v];
            l = make_float4(fminf(l.x, p.x), fminf(l.y, p.y), fminf(l.z, p.z), 0);
            h = make_float4(fmaxf(h.x, p.x), fmaxf(h.y, p.y), fmaxf(h.z, p.z), 0);
          }
        }
        lo[s] = make_float4(pad_dn(l.x), pad_dn(l.y), pad_dn(l.z), 0);
        hi[s] = make_float4(pad_up(h.x), pad_up(h.y), pad_up(h.z), 0);
      }
      nlo = make_float4(fminf(nlo.x, lo[s].x), fminf(nlo.y, lo[s].y), fminf(nlo.z, lo[s].z), 0);
      nhi = make_float4(fmaxf(nhi.x, hi[s].x), fmaxf(nhi.y, hi[s].y), fmaxf(nhi.z, hi[s].z), 0);
    }
    nd.px = nlo.x; nd.py = nlo.y; nd.pz = nlo.z;
    const int be[3] = {quant_exp(nhi.x - nlo.x), quant_exp(nhi.y - nlo.y), quant_exp(nhi.z - nlo.z)};
    nd.ex = (uint8_t)be[0]; nd.ey = (uint8_t)be[1]; nd.ez = (uint8_t)be[2];
    for (int s = 0; s < 8; ++s) {
      if (nd.meta[s] == 0) continue;
      uint8_t ql[3], qh[3];
      quantise_child(nlo, be, lo[s], hi[s], ql, qh);
      nd.qlox[s] = ql[0]; nd.qloy[s] = ql[1]; nd.qloz[s] = ql[2];
      nd.qhix[s] = qh[0]; nd.qhiy[s] = qh[1]; nd.qhiz[s] = qh[2];
    }
    const uint4* src = reinterpret_cast<const uint4*>(&nd);
    uint4* dst = reinterpret_cast<uint4*>(nodes + node_root + w);
#pragma unroll
    for (int k = 0; k < 5; ++k) dst[k] = src[k];
    node_box[2ull * w] = nlo; node_box[2ull * w + 1] = nhi;
    const uint32_t par = parent[w];
    if (par == EMPTY) return;
    __threadfence();
    const uint32_t need = __popc((uint32_t)nodes[node_root + par].imask);
    if (atomicAdd(&arrive[par], 1u) + 1u < need) return;  // not the last child: the last one continues
    __threadfence();
    w = par;
  }
}


// Refit, 8 lanes per wide node (lane s = child slot s), bottom-up: every lane rebuilds ITS slot — a leaf slot gathers the moved
// vertices of its (<= 3) triangles, rewrites the 48-byte records and takes their box, an inner slot reads its child's box — then
// the node box is a shuffle reduction over the 8 lanes, every lane quantises and stores its own 6 bytes and lane 0 the header.
// Replaces k_refit_tris + one thread per node (local arrays of 8 boxes, 80-byte struct copies): 1.41 -> see profiles/ at 10 M.
__global__ void __launch_bounds__(256) k_refit_nodes8(SnkNode8* __restrict__ nodes, uint32_t node_root, uint32_t node_count, SnkTri* __restrict__ tris,
                                                      const float* __restrict__ positions, const uint32_t* __restrict__ indices, uint32_t index_offset,
                                                      uint32_t vertex_offset, const uint32_t* __restrict__ parent, uint32_t* __restrict__ arrive,
                                                      float4* __restrict__ node_box) {
  const uint32_t lane = threadIdx.x & 31u, s = lane & 7u, gshift = lane & 24u, gm = 0xFFu << gshift;
  uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (w >= node_count) return;
  {
    const uint4 h0 = *reinterpret_cast<const uint4*>(nodes + node_root + w);
    if ((h0.w >> 24) != 0u) return;  // only nodes without inner children start a climb (uniform over the 8 lanes)
  }
  while (true) {
    uint4* np = reinterpret_cast<uint4*>(nodes + node_root + w);
    const uint4 h0 = __ldcg(np), h1 = __ldcg(np + 1);
    const uint32_t imask = h0.w >> 24, child_base = h1.x, tri_base = h1.y;
    const uint32_t meta = ((s < 4 ? h1.z : h1.w) >> (8 * (s & 3))) & 0xFFu;
    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (meta != 0u) {
      if ((meta & 0x18u) == 0x18u) {
        const uint32_t c = child_base - node_root + __popc(imask & ((1u << s) - 1u));
        const float4 l = __ldcg(&node_box[2ull * c]), h = __ldcg(&node_box[2ull * c + 1]);
        lo = make_float3(l.x, l.y, l.z); hi = make_float3(h.x, h.y, h.z);
      } else {
        const uint32_t first = tri_base + (meta & 31u), cnt = __popc(meta >> 5);
        for (uint32_t k = 0; k < cnt; ++k) {
          float4* t = reinterpret_cast<float4*>(tris + first + k);
          const uint32_t prim = __float_as_uint(t[0].w);
          const uint32_t* ip = indices + index_offset + 3ull * prim;
          const uint32_t iv[3] = {ip[0] + vertex_offset, ip[1] + vertex_offset, ip[2] + vertex_offset};
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            const float x = positions[3ull * iv[v]], y = positions[3ull * iv[v] + 1], z = positions[3ull * iv[v] + 2];
            t[v] = make_float4(x, y, z, v == 0 ? __uint_as_float(prim) : 0.f);
            lo = make_float3(fminf(lo.x, x), fminf(lo.y, y), fminf(lo.z, z));
            hi = make_float3(fmaxf(hi.x, x), fmaxf(hi.y, y), fmaxf(hi.z, z));
          }
        }
        lo = make_float3(pad_dn(lo.x), pad_dn(lo.y), pad_dn(lo.z));
        hi = make_float3(pad_up(hi.x), pad_up(hi.y), pad_up(hi.z));
      }
    }
    float3 nlo = lo, nhi = hi;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      nlo.x = fminf(nlo.x, __shfl_xor_sync(gm, nlo.x, o)); nlo.y = fminf(nlo.y, __shfl_xor_sync(gm, nlo.y, o)); nlo.z = fminf(nlo.z, __shfl_xor_sync(gm, nlo.z, o));
      nhi.x = fmaxf(nhi.x, __shfl_xor_sync(gm, nhi.x, o)); nhi.y = fmaxf(nhi.y, __shfl_xor_sync(gm, nhi.y, o)); nhi.z = fmaxf(nhi.z, __shfl_xor_sync(gm, nhi.z, o));
    }
    const int be[3] = {quant_exp(nhi.x - nlo.x), quant_exp(nhi.y - nlo.y), quant_exp(nhi.z - nlo.z)};
    uint8_t* bytes = reinterpret_cast<uint8_t*>(np);
    if (meta != 0u) {
      uint8_t ql[3], qh[3];
      quantise_child(make_float4(nlo.x, nlo.y, nlo.z, 0.f), be, make_float4(lo.x, lo.y, lo.z, 0.f), make_float4(hi.x, hi.y, hi.z, 0.f), ql, qh);
      bytes[32 + s] = ql[0]; bytes[40 + s] = ql[1]; bytes[48 + s] = ql[2];
      bytes[56 + s] = qh[0]; bytes[64 + s] = qh[1]; bytes[72 + s] = qh[2];
    }
    uint32_t next = EMPTY;
    if (s == 0) {
      np[0] = make_uint4(__float_as_uint(nlo.x), __float_as_uint(nlo.y), __float_as_uint(nlo.z),
                         (uint32_t)be[0] | ((uint32_t)be[1] << 8) | ((uint32_t)be[2] << 16) | (imask << 24));
      node_box[2ull * w] = make_float4(nlo.x, nlo.y, nlo.z, 0.f); node_box[2ull * w + 1] = make_float4(nhi.x, nhi.y, nhi.z, 0.f);
      const uint32_t par = parent[w];
      if (par != EMPTY) {
        __threadfence();
        const uint32_t need = __popc((uint32_t)nodes[node_root + par].imask);
        if (atomicAdd(&arrive[par], 1u) + 1u >= need) { __threadfence(); next = par; }  // the last child continues with the parent
      }
    }
    next = __shfl_sync(gm, next, gshift);
    if (next == EMPTY) return;
    w = next;
  }
}

template <class T>
cudaError_t alloc(T*& p, size_t count) {
  return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
}

}  // namespace

void scratch_free(BuildScratch& s) {
  cudaFree(s.prim_lo); cudaFree(s.prim_hi);
  for (int i = 0; i < 2; ++i) { cudaFree(s.keys[i]); cudaFree(s.vals[i]); }
  cudaFree(s.block_hist); cudaFree(s.left); cudaFree(s.right); cudaFree(s.parent); cudaFree(s.first); cudaFree(s.last);
  cudaFree(s.bb_lo); cudaFree(s.bb_hi); cudaFree(s.flags); cudaFree(s.queue); cudaFree(s.counters); cudaFree(s.digit_total); cudaFree(s.sweep);
  s = BuildScratch();
}

cudaError_t scratch_reserve(BuildScratch& s, uint32_t n) {
  if (n <= s.capacity && s.capacity) return cudaSuccess;
  scratch_free(s);
  const size_t cap = n < 1024 ? 1024 : (size_t)n + n / 8;
  cudaError_t e;
  s.hist_blocks = (uint32_t)((cap + SORT_TILE - 1) / SORT_TILE);
#define SNK_TRY(x) if ((e = (x)) != cudaSuccess) { scratch_free(s); return e; }
  SNK_TRY(alloc(s.prim_lo, cap)); SNK_TRY(alloc(s.prim_hi, cap));
  for (int i = 0; i < 2; ++i) { SNK_TRY(alloc(s.keys[i], cap)); SNK_TRY(alloc(s.vals[i], cap)); }
  SNK_TRY(alloc(s.block_hist, (size_t)256 * s.hist_blocks));
  SNK_TRY(alloc(s.left, cap)); SNK_TRY(alloc(s.right, cap)); SNK_TRY(alloc(s.parent, 2 * cap));
  SNK_TRY(alloc(s.first, cap)); SNK_TRY(alloc(s.last, cap));
  SNK_TRY(alloc(s.bb_lo, 2 * cap)); SNK_TRY(alloc(s.bb_hi, 2 * cap));
  SNK_TRY(alloc(s.flags, cap)); SNK_TRY(alloc(s.queue, cap + 1)); SNK_TRY(alloc(s.counters, 16)); SNK_TRY(alloc(s.digit_total, 256));
  SNK_TRY(alloc(s.sweep, (size_t)SW_HEAD + (size_t)SWEEP_MAX_PASSES * 256 * s.hist_blocks));
#undef SNK_TRY
  s.capacity = (uint32_t)cap;
  return cudaSuccess;
}

int launch_refit(RefitState& r, SnkNode8* nodes, uint32_t node_root, uint32_t node_count, SnkTri* tris, uint32_t tri_base, uint32_t tri_count,
                 const float* positions, const uint32_t* indices, uint32_t index_offset, uint32_t vertex_offset, cudaStream_t st) {
  int launches = 0;
  if (r.node_count != node_count || !r.parent) {
    cudaFree(r.parent); cudaFree(r.arrive); cudaFree(r.node_box);
    r = RefitState();
    if (alloc(r.parent, node_count) != cudaSuccess || alloc(r.arrive, node_count) != cudaSuccess || alloc(r.node_box, 2ull * node_count) != cudaSuccess)
      return -1;
    r.node_count = node_count;
    k_refit_parents<<<(node_count + 255) / 256, 256, 0, st>>>(nodes, node_root, node_count, r.parent);
    ++launches;
  }
  cudaMemsetAsync(r.arrive, 0, sizeof(uint32_t) * node_count, st);
#ifdef SNK_REFIT_ONE_THREAD
  if (tri_count) { k_refit_tris<<<(tri_count + 255) / 256, 256, 0, st>>>(tris, tri_base, tri_count, positions, indices, index_offset, vertex_offset); ++launches; }
  k_refit_nodes<<<(node_count + 127) / 128, 128, 0, st>>>(nodes, node_root, node_count, tris, r.parent, r.arrive, r.node_box);
#else
  k_refit_nodes8<<<(unsigned)(((uint64_t)node_count * 8 + 255) / 256), 256, 0, st>>>(nodes, node_root, node_count, tris, positions, indices, index_offset, vertex_offset,
                                                                                    r.parent, r.arrive, r.node_box);
#endif
  return launches + 1;
}
void refit_free(RefitState& r) { cudaFree(r.parent); cudaFree(r.arrive); cudaFree(r.node_box); r = RefitState(); }

void launch_triangle_bounds(BuildScratch& s, const float* positions, const uint32_t* indices, uint32_t index_offset,
                            uint32_t vertex_offset, uint32_t n, cudaStream_t st) {
  k_init<<<1, 32, 0, st>>>(s.counters, s.queue);
  s.clean = false;
  if (n) k_triangle_bounds<<<(n + 255) / 256, 256, 0, st>>>(positions, indices, index_offset, vertex_offset, n, s.prim_lo, s.prim_hi, s.counters);
}

void launch_instance_bounds(BuildScratch& s, const float* transforms_cm, const SnkInstanceData* inst,
                            const uint32_t* blas_of_instance, const SnkBlasInfo* blas, const SnkNode8* blas_nodes, const uint32_t* valid_slots,
                            uint32_t n, cudaStream_t st) {
  if (!s.clean) k_init<<<1, 32, 0, st>>>(s.counters, s.queue);  // a single-CTA TLAS build leaves its scratch initialised
  s.clean = false;
  if (n) k_instance_bounds<<<(n + 7) / 8, 256, 0, st>>>(transforms_cm, inst, blas_of_instance, blas, blas_nodes, valid_slots, n, s.prim_lo, s.prim_hi, s.counters);
}

// stable LSD radix sort of the (key, value) pairs in s.keys[cur] / s.vals[cur] over the low `bits` key bits; `cur` names the
// buffer that holds the result. Also used by the particle pass (particles.cu) for its cell keys.
int launch_sort_pairs(BuildScratch& s, uint32_t n, int bits, int& cur, cudaStream_t st) {
  int launches = 0;
  if (n <= 1) return 0;
  if (sort_onesweep() && n < (1u << 30)) return launch_sort_sweep(s, n, (bits + 7) / 8, cur, st);
  const uint32_t nb = (n + SORT_TILE - 1) / SORT_TILE;
  for (int shift = 0; shift < bits; shift += 8) {
    k_sort_hist<<<nb, SORT_THREADS, 0, st>>>(s.keys[cur], n, shift, s.block_hist, nb);
    k_scan_rows<<<256, 256, 0, st>>>(s.block_hist, nb, s.digit_total);
    k_scan_digits<<<1, 256, 0, st>>>(s.digit_total);
    launch_sort_scatter(s.keys[cur], s.vals[cur], n, shift, s.block_hist, nb, s.digit_total, s.keys[cur ^ 1], s.vals[cur ^ 1], st);
    cur ^= 1;
    launches += 4;
  }
  return launches;
}

namespace {
__global__ void k_widen_pairs(const uint32_t* __restrict__ k, const uint32_t* __restrict__ v, uint32_t n, unsigned long long* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = k[i]; vals[i] = v[i]; }
}
__global__ void k_narrow_pairs(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n, uint32_t* __restrict__ k,
                               uint32_t* __restrict__ v) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { k[i] = (uint32_t)keys[i]; v[i] = vals[i]; }
}
}  // namespace
int launch_sort_uints(BuildScratch& s, uint32_t* keys, uint32_t* vals, uint32_t n, cudaStream_t st) {
  const uint32_t g = (n + 255) / 256;
  k_widen_pairs<<<g, 256, 0, st>>>(keys, vals, n, s.keys[0], s.vals[0]);
  int cur = 0;
  const int launches = launch_sort_pairs(s, n, 32, cur, st);
  k_narrow_pairs<<<g, 256, 0, st>>>(s.keys[cur], s.vals[cur], n, keys, vals);
  return launches + 2;
}

int launch_build(BuildScratch& s, uint32_t n, const BuildOutput& out, const float* positions, const uint32_t* indices,
                 uint32_t index_offset, uint32_t vertex_offset, cudaStream_t st) {
  int launches = 0;
  if (n == 0) return 0;
  const uint32_t g = (n + 255) / 256;
#ifndef SNK_NO_SMALL_BUILD
  if (n <= (uint32_t)SMALL_MAX) {
    const uint32_t cap = out.node_capacity < n ? out.node_capacity : n;
    CollapseArgs a;
    a.n = (int)n; a.left = s.left; a.right = s.right; a.first = s.first; a.last = s.last; a.bb_lo = s.bb_lo; a.bb_hi = s.bb_hi;
    a.sorted_prims = s.vals[0]; a.queue = s.queue; a.counters = s.counters; a.nodes = out.nodes; a.node_offset = out.node_offset;
    a.node_capacity = cap; a.tris = out.tris; a.leaf_ids = out.leaf_ids; a.leaf_remap = out.leaf_remap; a.leaf_offset = out.leaf_offset;
    a.positions = positions; a.indices = indices; a.index_offset = index_offset; a.vertex_offset = vertex_offset;
    a.leaf_max = out.tris ? LEAF_MAX_BLAS : LEAF_MAX_TLAS;
    a.self_reset = out.tris ? 0 : 1;
    s.clean = a.self_reset != 0;
    uint32_t threads = 128;
    while (threads < n && threads < (uint32_t)SMALL_THREADS) threads <<= 1;
    // SNK_SMALL_PLOC: 0 = Karras hierarchy, 1 = locally-ordered clustering for every small build, 2 = for the TLAS only, 3 = for large TLASes
    const bool ploc = SNK_SMALL_PLOC == 1 || (SNK_SMALL_PLOC == 2 && !out.tris) || (SNK_SMALL_PLOC == 3 && !out.tris && n >= 1024u);
    if (ploc) {
      a.sorted_prims = s.vals[1];
      k_build_small_ploc<<<1, threads, 0, st>>>(s.prim_lo, s.prim_hi, (int)n, s.counters, s.vals[0], s.vals[1], s.left, s.right, s.parent, s.first, s.last,
                                                 s.bb_lo, s.bb_hi, s.flags, a);
    } else {
      const bool in_smem = n <= (uint32_t)SMALL_TREE_MAX;
      if (in_smem) {
        static bool configured[64] = {};  // the opt-in to > 48 KB of dynamic shared memory is per device
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !configured[dev]) {
          cudaFuncSetAttribute(k_build_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_TREE_SMEM);
          if (dev >= 0 && dev < 64) configured[dev] = true;
        }
      }
      k_build_small<<<1, threads, in_smem ? SMALL_TREE_SMEM : 0, st>>>(s.prim_lo, s.prim_hi, (int)n, s.counters, s.keys[0], s.vals[0], s.left, s.right, s.parent, s.first,
                                                                       s.last, s.bb_lo, s.bb_hi, s.flags, a, in_smem ? 1 : 0);
    }
    return 1;
  }
#endif
  // Morton resolution: 63 bits (21 per axis) are far more than n primitives can use; the keys keep 3 * b bits with
  // 2^(3b) >= 4096 n cells (at least 13 bits per axis beyond 4 M primitives, 10 below), i.e. 5 or 4 radix passes instead of 8.
  // Equal keys are ordered by primitive index (stable sort, Karras's tie-break), so the hierarchy stays well defined.
  // SNKPT_SORT_PASSES=8 restores the full keys (A/B: profiles/r04m_*).
  static int forced_passes = -1;
  if (forced_passes < 0) { const char* e = getenv("SNKPT_SORT_PASSES"); forced_passes = e ? atoi(e) : 0; }
  int passes = forced_passes >= 1 && forced_passes <= 8 ? forced_passes : (n > (4u << 20) ? 5 : 4);
  const int axis_bits = passes == 8 ? 21 : (8 * passes) / 3;
  const int key_bits = 3 * axis_bits;
  k_morton<<<g, 256, 0, st>>>(s.prim_lo, s.prim_hi, n, s.counters, s.keys[0], s.vals[0], 63 - key_bits, out.tris != nullptr); ++launches;
  int cur = 0;
  if (n > 1) {
    const uint32_t nb = (n + SORT_TILE - 1) / SORT_TILE;
    const int sort_bits = key_bits + (SNK_HOIST_LARGE ? 1 : 0);  // + the large-primitive bit above the Morton bits
    if (sort_onesweep() && n < (1u << 30)) launches += launch_sort_sweep(s, n, (sort_bits + 7) / 8, cur, st);
    else
    for (int shift = 0; shift < sort_bits; shift += 8) {
      k_sort_hist<<<nb, SORT_THREADS, 0, st>>>(s.keys[cur], n, shift, s.block_hist, nb);
      k_scan_rows<<<256, 256, 0, st>>>(s.block_hist, nb, s.digit_total);
      k_scan_digits<<<1, 256, 0, st>>>(s.digit_total);
      launch_sort_scatter(s.keys[cur], s.vals[cur], n, shift, s.block_hist, nb, s.digit_total, s.keys[cur ^ 1], s.vals[cur ^ 1], st);
      cur ^= 1;
      launches += 4;
    }
    cudaMemsetAsync(s.flags, 0, sizeof(uint32_t) * n, st);
    k_karras<<<g, 256, 0, st>>>(s.keys[cur], (int)n, s.left, s.right, s.parent, s.first, s.last); ++launches;
  } else {
    cudaMemsetAsync(s.parent, 0xFF, sizeof(int), st);  // parent[0] = -1: the only leaf is the root
  }
#if SNK_REFIT_CHUNKED
  {
    uint32_t* pending = s.vals[cur ^ 1];  // the sort's other buffer is free again
    k_refit_binary_chunked<<<(n + RB_CHUNK - 1) / RB_CHUNK, RB_CHUNK, 0, st>>>(s.prim_lo, s.prim_hi, s.vals[cur], (int)n, s.parent, s.first, s.last, s.bb_lo,
                                                                            s.bb_hi, pending, s.counters + C_PENDING);
    const uint32_t want = (n / 16u + 255u) / 256u, resident = (uint32_t)device_sm_count() * 8u;
    k_refit_binary_top<<<want < resident ? (want ? want : 1u) : resident, 256, 0, st>>>(pending, s.counters + C_PENDING, s.parent, s.bb_lo, s.bb_hi, s.flags);
    launches += 2;
  }
#else
  k_refit_binary<<<g, 256, 0, st>>>(s.prim_lo, s.prim_hi, s.vals[cur], (int)n, s.left, s.right, s.parent, s.bb_lo, s.bb_hi, s.flags); ++launches;
#endif
  // queue[0] = root was set by k_init
  const uint32_t cap = out.node_capacity < n ? out.node_capacity : n;
  CollapseArgs a;
  a.n = (int)n; a.left = s.left; a.right = s.right; a.first = s.first; a.last = s.last; a.bb_lo = s.bb_lo; a.bb_hi = s.bb_hi;
  a.sorted_prims = s.vals[cur]; a.queue = s.queue; a.counters = s.counters; a.nodes = out.nodes; a.node_offset = out.node_offset;
  a.node_capacity = cap; a.tris = out.tris; a.leaf_ids = out.leaf_ids; a.leaf_remap = out.leaf_remap; a.leaf_offset = out.leaf_offset;
  a.positions = positions; a.indices = indices; a.index_offset = index_offset; a.vertex_offset = vertex_offset;
    a.leaf_max = out.tris ? LEAF_MAX_BLAS : LEAF_MAX_TLAS;
    a.self_reset = 0;
  {
    // resident grid of the cooperative launch, per device (a process may drive contexts on several GPUs, from several threads)
    static std::atomic<int> resident_of_device[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = (dev >= 0 && dev < 64) ? dev : 0;
    int resident_i = resident_of_device[slot].load(std::memory_order_relaxed);
    if (resident_i == 0) {
      int sms = 0, ctas_per_sm = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_collapse, 128, 0);
      if (ctas_per_sm < 1) ctas_per_sm = 1;
      resident_i = (sms > 0 ? sms : 1) * ctas_per_sm;
      resident_of_device[slot].store(resident_i, std::memory_order_relaxed);
    }
    const uint32_t want = (uint32_t)(((uint64_t)cap * (SNK_COLLAPSE_OCTET ? 8 : 1) + 127) / 128), resident = (uint32_t)resident_i;
    void* args[] = {&a};
    if (cudaLaunchCooperativeKernel((const void*)k_collapse, dim3(want < resident ? want : resident), dim3(128), args, 0, st) != cudaSuccess) return -1;
    ++launches;
  }
  return launches;
}

}  // namespace snk
