// vecmath.cuh — float32 vector helpers for the shading kernels.
//
// Numerics contract of the shading path: every kernel that
// includes this header is compiled with -fmad=false, so + - * / sqrt are IEEE
// binary32 operations evaluated in the order written; sin / cos / tan are
// evaluated in float64 and rounded once; pow(x,5) is ((x*x)*(x*x))*x; min / max
// follow the GLSL definitions. This makes the GLSL restatement reproducible
// bit-for-bit between this CUDA path and the CPU oracle used by the tests.
#pragma once
#include <cuda_runtime.h>

#include <cuda_fp16.h>

#include <cstdint>

#if defined(__FAST_MATH__) || defined(__USE_FAST_MATH__)
#error "shading kernels must not be compiled with fast math"
#endif

namespace snk {

struct vec3 { float x, y, z; };

__device__ __forceinline__ vec3 v3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ vec3 v3(float s) { return v3(s, s, s); }
__device__ __forceinline__ vec3 operator+(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ vec3 operator-(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ vec3 operator*(vec3 a, vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ vec3 operator/(vec3 a, vec3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ vec3 operator*(vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ vec3 operator*(float s, vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ vec3 operator/(vec3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
// north-star estimator only (its numerics are this repository's own, oracle/snk_math.h:inv_scale is the same expression): one
// IEEE reciprocal and three products instead of three IEEE divisions — the light-sampling loops evaluate 4 of these per candidate
__device__ __forceinline__ vec3 inv_scale(vec3 a, float s) { const float r = 1.0f / s; return v3(a.x * r, a.y * r, a.z * r); }
__device__ __forceinline__ vec3 operator-(vec3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
__device__ __forceinline__ vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
__device__ __forceinline__ vec3& operator*=(vec3& a, float s) { a = a * s; return a; }

__device__ __forceinline__ float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ vec3 cross(vec3 a, vec3 b) {
  return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float length(vec3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float length2(float x, float y) { return sqrtf(x * x + y * y); }
__device__ __forceinline__ vec3 normalize(vec3 a) { const float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
__device__ __forceinline__ float gmax(float x, float y) { return x < y ? y : x; }
__device__ __forceinline__ float gmin(float x, float y) { return y < x ? y : x; }
__device__ __forceinline__ float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
__device__ __forceinline__ float fract(float x) { return x - floorf(x); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ vec3 mix(vec3 a, vec3 b, float t) { return v3(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
// real calls, not inlined: each float64 sin / cos / tan expands to ~300 SASS instructions and the shading kernels
// use them at dozens of program points; inlined they overflow the 32 KB instruction cache (profiles/r01c_prof_restir.txt)
static __device__ __noinline__ float sin_c(float x) { return (float)sin((double)x); }
static __device__ __noinline__ float cos_c(float x) { return (float)cos((double)x); }
static __device__ __noinline__ float tan_c(float x) { return (float)tan((double)x); }
__device__ __forceinline__ float pow5(float x) { const float x2 = x * x; return (x2 * x2) * x; }

// GLSL `M * vec4(p, 1)`.xyz for a column-major mat4
__device__ __forceinline__ vec3 mul_point(const float* __restrict__ M, vec3 p) {
  return v3(((M[0] * p.x + M[4] * p.y) + M[8] * p.z) + M[12], ((M[1] * p.x + M[5] * p.y) + M[9] * p.z) + M[13],
            ((M[2] * p.x + M[6] * p.y) + M[10] * p.z) + M[14]);
}
__device__ __forceinline__ float4 mul_vec4(const float* __restrict__ M, float x, float y, float z, float w) {
  float4 r;
  r.x = ((M[0] * x + M[4] * y) + M[8] * z) + M[12] * w;
  r.y = ((M[1] * x + M[5] * y) + M[9] * z) + M[13] * w;
  r.z = ((M[2] * x + M[6] * y) + M[10] * z) + M[14] * w;
  r.w = ((M[3] * x + M[7] * y) + M[11] * z) + M[15] * w;
  return r;
}
// row-major 3x3 with a row stride of 3 floats
__device__ __forceinline__ vec3 mul_mat3r(const float* __restrict__ A, vec3 v) {
  return v3((A[0] * v.x + A[1] * v.y) + A[2] * v.z, (A[3] * v.x + A[4] * v.y) + A[5] * v.z, (A[6] * v.x + A[7] * v.y) + A[8] * v.z);
}

// ---- storage formats of the G-buffer (EditorLayer.cpp:229-256) -----------------------------
__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ signed char f2snorm8(float f) {
  if (!(f == f)) return 0;
  const float c = f < -1.0f ? -1.0f : (f > 1.0f ? 1.0f : f);
  return (signed char)__float2int_rn(c * 127.0f);
}
__device__ __forceinline__ float snorm8_2f(signed char v) { const float f = (float)v / 127.0f; return f < -1.0f ? -1.0f : f; }

// Util.glsl:1-7 with inverse(view)*inverse(proj) hoisted to the host
__device__ __forceinline__ vec3 world_pos_from_depth(const float* __restrict__ ivp, float depth, float u, float v) {
  const float4 ws = mul_vec4(ivp, u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f);
  const float w = gmax(ws.w, 1e-6f);
  return v3(ws.x / w, ws.y / w, ws.z / w);
}

}  // namespace snk
