// layouts.h — wire formats (asserted against the reference's GLSL/C++ layouts)
// and the device-side acceleration-structure records.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/snkpt.h"

// ---- reference wire formats -------------------------------------------------------
static_assert(sizeof(SnkInstanceData) == 20, "RaytracingInstances.glsl:7-13 InstanceData is 5 x u32");
static_assert(sizeof(SnkMaterial) == 64 && offsetof(SnkMaterial, albedo) == 32 && offsetof(SnkMaterial, flags) == 48,
              "TexMatBuffers.glsl:5-20 Material is 64 B");
static_assert(sizeof(SnkDirectionalLight) == 96 && sizeof(SnkPointlight) == 48 && sizeof(SnkSpotlight) == 64,
              "LightBuffers.glsl:1-27");
static_assert(offsetof(SnkLightData, num_pointlights) == 96 && offsetof(SnkLightData, pointlights) == 112 &&
                  offsetof(SnkLightData, spotlights) == 49264,
              "LightBuffers.glsl:41-49 LightData offsets");
static_assert(sizeof(SnkCommonUBO) == 240 && offsetof(SnkCommonUBO, frame_idx) == 224, "CommonUBO.glsl:1-13 (std140)");
static_assert(sizeof(SnkPixelReservoir) == 48 && offsetof(SnkPixelReservoir, w_sum_em) == 32, "RayGen.rgen:53-76 (std430)");
static_assert(sizeof(SnkRay) == 32 && sizeof(SnkHit) == 20, "ray / hit records");
static_assert(sizeof(SnkRtSettings) == 64, "SnkRtSettings");

// ---- acceleration structure ---------------------------------------------------------
// 80-byte compressed wide node (8 children, quantised child boxes), read as 5
// coalesced 128-bit loads. Layout:
//   w0: p.x p.y p.z | ex ey ez imask           origin of the 8-bit grid, biased exponents, inner-child mask
//   w1: child_base | tri_base | meta[0..3] | meta[4..7]
//   w2: qlo_x[0..7] | qlo_y[0..7]
//   w3: qlo_z[0..7] | qhi_x[0..7]
//   w4: qhi_y[0..7] | qhi_z[0..7]
// meta[i]: 0 = empty; inner child: 0b001_11000 | slot; leaf: (unary count << 5) | first record offset (0..23)
struct alignas(16) SnkNode8 {
  float px, py, pz;
  uint8_t ex, ey, ez, imask;
  uint32_t child_base, tri_base;
  uint8_t meta[8];
  uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(SnkNode8) == 80, "wide node is 80 B");

// 48-byte triangle record (object space), 3 x 128-bit loads; v0w carries the primitive id
struct alignas(16) SnkTri {
  float v0x, v0y, v0z;
  uint32_t prim;
  float v1x, v1y, v1z;
  uint32_t pad1;
  float v2x, v2y, v2z;
  uint32_t pad2;
};
static_assert(sizeof(SnkTri) == 48, "triangle record is 48 B");

// device TLAS instance record: world->object 3x4 (rows), BLAS root, 64 B = 4 x 128-bit
struct alignas(16) SnkInstRec {
  float w2o[12];        // row-major 3x4: o' = R*o + t
  uint32_t blas_root;   // node index of the BLAS root
  uint32_t custom_idx;  // instanceCustomIndex (slot in the InstanceData buffer)
  uint32_t blas_id;
  uint32_t flags;
};
static_assert(sizeof(SnkInstRec) == 64, "instance record is 64 B");
