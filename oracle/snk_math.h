// oracle/snk_math.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Scalar float32 vector helpers used by the CPU restatement of Snake-VK's
// ray-tracing shaders. The oracle is compiled with -ffp-contract=off so every
// expression below is evaluated exactly as written (IEEE binary32, round to
// nearest even, no fused multiply-add). That is the "numerics contract" the
// CUDA kernels are written against:
//   * + - * / sqrt are IEEE float32, evaluated left to right as written;
//   * sin / cos / tan are evaluated in float64 and rounded once to float32
//     (GLSL leaves their precision implementation-defined; Util.glsl:77-82's
//     gold_noise is chaotic in the last ulp of tan, so a definition is needed);
//   * pow(x, 5.0) is ((x*x)*(x*x))*x and pow(x, 2) is x*x (LightBuffers.glsl:66, :153);
//   * min/max/clamp follow the GLSL definitions (y < x ? y : x), not fmin/fmax.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace snkor {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

inline vec3 v3(float x, float y, float z) { return vec3{x, y, z}; }
inline vec3 v3(float s) { return vec3{s, s, s}; }
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// north-star estimator only (csrc/vecmath.cuh:inv_scale): one reciprocal, three products
inline vec3 inv_scale(vec3 a, float s) { const float r = 1.0f / s; return {a.x * r, a.y * r, a.z * r}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }

inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) {
  return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline float length2(float x, float y) { return sqrtf(x * x + y * y); }
inline vec3 normalize(vec3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline float gmax(float x, float y) { return x < y ? y : x; }   // GLSL max
inline float gmin(float x, float y) { return y < x ? y : x; }   // GLSL min
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float fract(float x) { return x - floorf(x); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, float t) { return {mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)}; }
inline vec3 reflect(vec3 i, vec3 n) { return i - n * (2.0f * dot(n, i)); }
inline float sin_c(float x) { return (float)sin((double)x); }
inline float cos_c(float x) { return (float)cos((double)x); }
inline float tan_c(float x) { return (float)tan((double)x); }
inline float pow5(float x) { float x2 = x * x; return (x2 * x2) * x; }

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// column-major 4x4 (glm / GLSL): m[c*4 + r]
struct mat4 { float m[16]; float at(int c, int r) const { return m[c * 4 + r]; } };

// GLSL `M * vec4(p, 1)` restricted to xyz: ((c0*x + c1*y) + c2*z) + c3
inline vec3 mul_point(const mat4& M, vec3 p) {
  return {((M.m[0] * p.x + M.m[4] * p.y) + M.m[8] * p.z) + M.m[12],
          ((M.m[1] * p.x + M.m[5] * p.y) + M.m[9] * p.z) + M.m[13],
          ((M.m[2] * p.x + M.m[6] * p.y) + M.m[10] * p.z) + M.m[14]};
}
inline vec4 mul_vec4(const mat4& M, vec4 p) {
  vec4 r;
  r.x = ((M.m[0] * p.x + M.m[4] * p.y) + M.m[8] * p.z) + M.m[12] * p.w;
  r.y = ((M.m[1] * p.x + M.m[5] * p.y) + M.m[9] * p.z) + M.m[13] * p.w;
  r.z = ((M.m[2] * p.x + M.m[6] * p.y) + M.m[10] * p.z) + M.m[14] * p.w;
  r.w = ((M.m[3] * p.x + M.m[7] * p.y) + M.m[11] * p.z) + M.m[15] * p.w;
  return r;
}

// 3x3 stored as rows r[i][j] (math convention), applied as ((r0*x + r1*y) + r2*z)
struct mat3r { float r[3][3]; };
inline vec3 mul(const mat3r& A, vec3 v) {
  return {(A.r[0][0] * v.x + A.r[0][1] * v.y) + A.r[0][2] * v.z,
          (A.r[1][0] * v.x + A.r[1][1] * v.y) + A.r[1][2] * v.z,
          (A.r[2][0] * v.x + A.r[2][1] * v.y) + A.r[2][2] * v.z};
}

// float64 general 4x4 inverse (Gauss-Jordan, partial pivoting); callers round
// the result once to float32. Used for inverse(view), inverse(proj)
// (Util.glsl:3) hoisted out of the pixel loop, and for world->object matrices.
inline bool invert4_d(const double* a_cm, double* inv_cm) {
  double a[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { a[r][c] = a_cm[c * 4 + r]; a[r][c + 4] = (r == c) ? 1.0 : 0.0; }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    for (int r = col + 1; r < 4; ++r) if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
    if (a[piv][col] == 0.0) return false;
    if (piv != col) for (int c = 0; c < 8; ++c) { double t = a[piv][c]; a[piv][c] = a[col][c]; a[col][c] = t; }
    double d = 1.0 / a[col][col];
    for (int c = 0; c < 8; ++c) a[col][c] *= d;
    for (int r = 0; r < 4; ++r) {
      if (r == col) continue;
      double f = a[r][col];
      if (f != 0.0) for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
    }
  }
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) inv_cm[c * 4 + r] = a[r][c + 4];
  return true;
}
inline mat4 inverse_f(const mat4& M) {
  double a[16], inv[16];
  for (int i = 0; i < 16; ++i) a[i] = M.m[i];
  mat4 R;
  if (!invert4_d(a, inv)) { for (int i = 0; i < 16; ++i) R.m[i] = NAN; return R; }
  for (int i = 0; i < 16; ++i) R.m[i] = (float)inv[i];
  return R;
}
// mat4 product evaluated in float64, rounded once (hoisted inverse(view)*inverse(proj))
inline mat4 mul_d(const mat4& A, const mat4& B) {
  mat4 R;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += (double)A.m[k * 4 + r] * (double)B.m[c * 4 + k];
      R.m[c * 4 + r] = (float)s;
    }
  return R;
}

}  // namespace snkor
