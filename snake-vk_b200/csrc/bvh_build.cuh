// bvh_build.cuh — interface of the sm_100a LBVH -> BVH8 builder (bvh_build.cu).
//
// Replaces the opaque driver builds of the reference:
//   BLAS::GenerateFromMeshData  (Core/src/resources/TLAS.cpp:109-186, vkCmdBuildAccelerationStructuresKHR :178)
//   TLAS::BuildFromInstances    (Core/src/resources/TLAS.cpp:10-87,  vkCmdBuildAccelerationStructuresKHR :74)
// Pipeline: primitive AABBs + centroid bounds -> 63-bit Morton codes -> LSD radix
// sort (8-bit digits, warp match ranking) -> Karras'12 hierarchy -> bottom-up AABB
// refit with atomic arrival flags -> single-launch treelet collapse to 8-wide nodes
// with quantised child boxes and octant-ordered slots (80-B nodes, 48-B triangles).
#pragma once
#include <cuda_runtime.h>

#include "layouts.h"

namespace snk {

struct BuildScratch {
  uint32_t capacity = 0;                                // primitives this scratch is sized for
  float4 *prim_lo = nullptr, *prim_hi = nullptr;        // primitive AABBs (input order)
  unsigned long long* keys[2] = {nullptr, nullptr};     // Morton codes (ping-pong)
  uint32_t* vals[2] = {nullptr, nullptr};               // primitive ids (ping-pong)
  uint32_t* block_hist = nullptr;                        // 256 x hist_blocks digit counts
  int *left = nullptr, *right = nullptr, *parent = nullptr;  // binary hierarchy: 2n-1 nodes, n-1 internal first
  int *first = nullptr, *last = nullptr;                // sorted range of every internal node
  float4 *bb_lo = nullptr, *bb_hi = nullptr;            // binary node boxes (2n-1)
  uint32_t* flags = nullptr;                             // arrival counters (n-1)
  uint32_t* queue = nullptr;                             // binary node handled by wide node w
  uint32_t* counters = nullptr;                          // 16 words, see bvh_build.cu
  uint32_t* digit_total = nullptr;                       // 256 words: per-digit totals / bases of the current sort pass
  uint32_t* sweep = nullptr;                             // one-sweep sort: digit totals and tickets of all passes, then the tiles' status words
  uint32_t hist_blocks = 0;
  bool clean = false;                                    // counters / queue[0] are in their initial state (left so by a single-CTA TLAS build)
};

cudaError_t scratch_reserve(BuildScratch& s, uint32_t n);
void scratch_free(BuildScratch& s);

struct BuildOutput {
  SnkNode8* nodes;         // node pool base
  uint32_t node_offset;    // absolute index of this hierarchy's root
  uint32_t node_capacity;  // nodes available from node_offset
  SnkTri* tris;            // triangle pool (BLAS builds) or nullptr
  uint32_t* leaf_ids;      // leaf id list (TLAS builds) or nullptr
  const uint32_t* leaf_remap;  // TLAS: primitive -> instance slot (nullptr = identity)
  uint32_t leaf_offset;    // absolute index of the first free leaf record
};

// side arrays of a topology-preserving refit, kept per BLAS after its first refit
struct RefitState {
  uint32_t node_count = 0;
  uint32_t* parent = nullptr;   // parent wide node (relative index), EMPTY for the root
  uint32_t* arrive = nullptr;   // arrival counters of the bottom-up pass
  float4* node_box = nullptr;   // 2 x float4 per node: its float box (read by the parent)
};
// rewrites the triangle records from `positions` and refits / re-quantises every node bottom-up; node_box[0..1] then
// holds the root box. Returns the number of kernels launched, -1 on allocation failure.
int launch_refit(RefitState& r, SnkNode8* nodes, uint32_t node_root, uint32_t node_count, SnkTri* tris, uint32_t tri_base, uint32_t tri_count,
                 const float* positions, const uint32_t* indices, uint32_t index_offset, uint32_t vertex_offset, cudaStream_t st);
void refit_free(RefitState& r);

// triangle AABBs of one submesh -> s.prim_lo / s.prim_hi (+ centroid bounds in s.counters)
void launch_triangle_bounds(BuildScratch& s, const float* positions, const uint32_t* indices, uint32_t index_offset,
                            uint32_t vertex_offset, uint32_t n, cudaStream_t st);
// world AABBs of n instances (primitive i = slot valid_slots[i]) -> s.prim_lo / s.prim_hi
void launch_instance_bounds(BuildScratch& s, const float* transforms_cm, const SnkInstanceData* inst,
                            const uint32_t* blas_of_instance, const SnkBlasInfo* blas, const SnkNode8* blas_nodes, const uint32_t* valid_slots,
                            uint32_t n, cudaStream_t st);
// stable LSD radix sort of s.keys[cur] / s.vals[cur] over the low `bits` key bits (8 bits per pass); updates `cur`
int launch_sort_pairs(BuildScratch& s, uint32_t n, int bits, int& cur, cudaStream_t st);
// snkpt_sort_uints_by_key: 32-bit keys / values of the caller, sorted in place through the scratch's 64-bit key buffers
int launch_sort_uints(BuildScratch& s, uint32_t* keys, uint32_t* vals, uint32_t n, cudaStream_t st);
// generic build over s.prim_lo/hi. Returns the number of kernels launched. After completion s.counters[0] holds
// the number of wide nodes written, s.counters[1] the number of leaf records, s.counters[4] an overflow flag.
int launch_build(BuildScratch& s, uint32_t n, const BuildOutput& out, const float* positions, const uint32_t* indices,
                 uint32_t index_offset, uint32_t vertex_offset, cudaStream_t st);

}  // namespace snk
