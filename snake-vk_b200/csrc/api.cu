// api.cu — implementation of the C ABI declared in include/snkpt.h.
//
// Host side of the pass: owns the device mirrors of the reference's buffers,
// builds BLAS / TLAS, derives the per-transform matrices the shaders recompute
// per use (inverse, normal matrix, determinant: RayClosestHit.rchit:111,
// RayGen.rgen:251,:268,:333 — hoisted here in float64 and rounded once), and
// enqueues the kernels. No CPU compute path exists: every entry point that
// produces results launches CUDA kernels on the context's device.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "kernels.h"

#include <cstdlib>

using namespace snk;

static thread_local std::string g_create_error;

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
      return SNKPT_E_CUDA;                                                               \
    }                                                                                    \
  } while (0)
#define REQUIRE(cond, code, msg)                                                         \
  do {                                                                                   \
    if (!(cond)) { ctx->err = msg; return code; }                                        \
  } while (0)

// ---- float64 matrix helpers (host) -----------------------------------------------------
static bool invert4(const double* a_cm, double* inv_cm) {
  // cofactor expansion, column-major in / out
  const double* m = a_cm;
  double inv[16];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  if (det == 0.0) return false;
  const double id = 1.0 / det;
  for (int i = 0; i < 16; ++i) inv_cm[i] = inv[i] * id;
  return true;
}

static void derive_transform(const float* M, float* w2o12, float* nmat9, float* det3) {
  double a[16], inv[16];
  for (int i = 0; i < 16; ++i) a[i] = M[i];
  if (!invert4(a, inv)) for (int i = 0; i < 16; ++i) inv[i] = NAN;
  // rows of the inverse as 3x4 (row-major): row r = inv[c*4 + r], c = 0..3
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) w2o12[r * 4 + c] = (float)inv[c * 4 + r];
  const double m00 = M[0], m01 = M[4], m02 = M[8], m10 = M[1], m11 = M[5], m12 = M[9], m20 = M[2], m21 = M[6], m22 = M[10];
  const double det = m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) + m02 * (m10 * m21 - m11 * m20);
  *det3 = (float)det;
  const double id = 1.0 / det;
  // inverse(mat3) entries i[r][c]; normal matrix = transpose → n[r][c] = i[c][r]
  const double i3[3][3] = {{(m11 * m22 - m12 * m21) * id, (m02 * m21 - m01 * m22) * id, (m01 * m12 - m02 * m11) * id},
                           {(m12 * m20 - m10 * m22) * id, (m00 * m22 - m02 * m20) * id, (m02 * m10 - m00 * m12) * id},
                           {(m10 * m21 - m11 * m20) * id, (m01 * m20 - m00 * m21) * id, (m00 * m11 - m01 * m10) * id}};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) nmat9[r * 3 + c] = (float)i3[c][r];
}

// ---- per-launch profiling ---------------------------------------------------------------------
namespace snk {
void prof_reset(SnkPt*) {}  // records accumulate until read (snkpt_get_launch_timings) or re-armed (snkpt_set_profiling)
int prof_begin(SnkPt* ctx, const char* name, uint64_t units, bool is_trace, bool any, cudaStream_t st) {
  if (!ctx->profiling || ctx->prof_n >= PROF_MAX) return -1;
  if (ctx->profiling == 3 && !is_trace) return -1;  // level 3: traversal launches only
  ProfRec& r = ctx->prof[ctx->prof_n];
  if (!r.e0 && (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess)) return -1;
  memset(r.name, 0, sizeof(r.name));
  snprintf(r.name, sizeof(r.name), "%s", name);
  r.units = units; r.is_trace = is_trace; r.any = any; r.counted = false; r.queue = false;
  cudaEventRecord(r.e0, st);
  return ctx->prof_n++;
}
void prof_end(SnkPt* ctx, int rec, cudaStream_t st) {
  if (rec >= 0) cudaEventRecord(ctx->prof[rec].e1, st);
}
int traced_launch(SnkPt* ctx, const char* name, const float4* rays, uint64_t n, bool any, SnkHit* hits_aos, float4* hits_soa,
                  uint32_t* inst_soa, float* vis_out, cudaStream_t st) {
  if (n == 0) return 0;
  const bool count = ctx->profiling == 2 && ctx->prof_n < PROF_MAX;
  if (count && !hits_soa) {
    // the work counter needs the reported hit distance of every ray
    if (ctx->prof_hits.reserve(16 * n, false, st) == cudaSuccess) hits_soa = ctx->prof_hits.as<float4>();
  }
  const TraceScene S = ctx->trace_scene();
  const int rec = prof_begin(ctx, name, n, true, any, st);
  // every launch draws from its own ticket word (launches of one frame may overlap on different streams later)
  unsigned long long* ticket = ctx->dev_counters.as<unsigned long long>() + 16 + (ctx->ticket_rr++ & 15u);
  const int launches = launch_trace(S, rays, n, any, hits_aos, hits_soa, inst_soa, vis_out, ticket, st);
  prof_end(ctx, rec, st);
  if (count && rec >= 0 && hits_soa && ctx->prof_work.reserve(64 * PROF_MAX, false, st) == cudaSuccess) {
    unsigned long long* sums = ctx->prof_work.as<unsigned long long>() + 8 * rec;
    cudaMemsetAsync(sums, 0, 64, st);
    launch_count_work_sum(S, rays, hits_soa, n, sums, st);
    ctx->prof[rec].counted = true;
  }
  return launches;
}

// ---- compact ray-reference queues of the frame (trace_q.cu) ----
// dev_counters u32 view [64 + 2k] = entries appended to queue k, [65 + 2k] = ticket of its traversal launch
RefQueue ref_queue(SnkPt* ctx, int k, uint32_t cap, uint32_t first_ref) {
  RefQueue q;
  q.refs = ctx->wave_refs.as<uint32_t>() + first_ref;
  q.count = ctx->dev_counters.as<uint32_t>() + 64 + 2 * k;
  q.ticket = q.count + 1;
  q.cap = cap;
  return q;
}
int reserve_ref_queues(SnkPt* ctx, uint32_t cap, cudaStream_t st) {
  const cudaError_t e = ctx->wave_refs.reserve(4ull * cap, false, st);
  if (e != cudaSuccess) { ctx->err = std::string("ray reference queue: ") + cudaGetErrorString(e); return SNKPT_E_NOMEM; }
  return SNKPT_OK;
}
// 24 queues (u32 [64..112)) and the per-stage path-list counters of northstar.cu (u32 [112..136))
void reset_ref_queues(SnkPt* ctx, cudaStream_t st) { cudaMemsetAsync(ctx->dev_counters.as<uint32_t>() + 64, 0, 4 * (136 - 64), st); }

// north-star frame: the ReSTIR chain may run on the auxiliary stream (off under profiling, whose per-launch timings want one
// launch at a time, and with SNKPT_NS_OVERLAP=0)
bool ns_overlap_enabled(const SnkPt* ctx) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SNKPT_NS_OVERLAP"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1 && (ctx->profiling == 0 || ctx->profiling >= 3);
}

int traced_launch_q(SnkPt* ctx, const char* name, const TraceSegs& segs, const RefQueue& q, cudaStream_t st, int spill_region) {
  const bool count = ctx->profiling == 2 && ctx->prof_n < PROF_MAX;
  float* prof_t = nullptr;
  if (count && ctx->prof_hits.reserve(4ull * q.cap, false, st) == cudaSuccess) prof_t = ctx->prof_hits.as<float>();
  TraceScene S = ctx->trace_scene();
  // two traversal launches may be in flight at once (north-star chains): each gets its own stack spill region
  S.spill += (size_t)spill_region * (trace_spill_bytes() / sizeof(uint2));
  const int rec = prof_begin(ctx, name, q.cap, true, false, st);
  const int launches = launch_trace_q(S, segs, q, prof_t, st);
  prof_end(ctx, rec, st);
  if (rec >= 0) ctx->prof[rec].queue = true;
  if (count && rec >= 0 && prof_t && ctx->prof_work.reserve(64 * PROF_MAX, false, st) == cudaSuccess) {
    unsigned long long* sums = ctx->prof_work.as<unsigned long long>() + 8 * rec;
    cudaMemsetAsync(sums, 0, 64, st);
    launch_count_work_q(S, segs, q, prof_t, sums, st);
    ctx->prof[rec].counted = true;
  }
  return launches;
}
}  // namespace snk

TraceScene SnkPt::trace_scene() const {
  TraceScene S;
  S.tlas_nodes = tlas_nodes.as<uint4>();
  S.tlas_leaf_ids = tlas_leaf_ids.as<uint32_t>();
  S.blas_nodes = blas_nodes.as<uint4>();
  S.tris = blas_tris.as<float4>();
  S.inst = inst_recs.as<SnkInstRec>();
  S.inst_sphere = inst_sphere.as<float4>();
  S.overflow = dev_counters.as<uint32_t>();
  S.tlas_node_count = (tlas_built && tlas_prims) ? 1u : 0u;
  S.spill = trav_spill.as<uint2>();
  return S;
}

SceneDev SnkPt::scene_dev() const {
  SceneDev D;
  D.positions = positions.as<float>(); D.normals = normals.as<float>(); D.uvs = uvs.as<float>(); D.tangents = tangents.as<float>();
  D.indices = indices.as<uint32_t>();
  D.instances = instances.as<SnkInstanceData>();
  D.transforms = transforms.as<float>(); D.prev_transforms = prev_transforms.as<float>();
  D.normal_mats = normal_mats.as<float>(); D.det3 = det3.as<float>();
  D.materials = materials.as<SnkMaterial>();
  D.lights = lights.as<SnkLightData>();
  D.emissive_idx = emissive_idx.as<uint32_t>();
  D.num_emissive = num_emissive; D.num_instances = num_instances;
  D.textures = tex_table.as<TexDev>(); D.srgb_lut = srgb_lut.as<float>(); D.num_textures = (uint32_t)h_tex.size();
  D.em_table = em_table.as<float4>(); D.em_base_of_obj = em_base_of_obj.as<uint32_t>(); D.em_base = em_base.as<uint32_t>();
  return D;
}

// ---- per-pixel planes ----------------------------------------------------------------------
// the G-buffer planes exist twice (context.h): index into px[] of the planes of copy 0; copy 1 lives in px_alt[] in the same order
static const int GB_PLANES = 11;
static const int GB_PLANE_IDX[GB_PLANES] = {0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 26};
static void set_gbuffer_planes(SnkPt* ctx) {
  PixelPlanes& P = ctx->planes;
  void* cur[GB_PLANES]; void* oth[GB_PLANES];
  for (int g = 0; g < GB_PLANES; ++g) {
    void* a = ctx->px[GB_PLANE_IDX[g]].p; void* b = ctx->px_alt[g].p;
    cur[g] = ctx->gb_cur ? b : a; oth[g] = ctx->gb_cur ? a : b;
  }
  P.albedo = (char4*)cur[0]; P.normal = (ushort4*)cur[1]; P.depth = (float*)cur[2]; P.rma = (ushort4*)cur[3]; P.motion = (ushort2*)cur[4];
  P.flags = (uint16_t*)cur[5]; P.prim = (uint32_t*)cur[6]; P.inst = (uint32_t*)cur[7]; P.hit_t = (float*)cur[8]; P.bary = (float2*)cur[9];
  P.surf = (uint4*)cur[10];
  P.prev_normal = (ushort4*)oth[1]; P.prev_depth = (float*)oth[2];  // what the previous snkpt_trace_primary wrote (GBufferPass.cpp:73-77)
}
static int alloc_pixels(SnkPt* ctx) {
  const size_t n = (size_t)ctx->W * ctx->H;
  // planes 6, 7 were the previous frame's normals / depths: they are the OTHER copy of the G-buffer planes now (px_alt)
  const size_t sizes[] = {4, 8, 4, 8, 4, 2, 0, 0, 4, 4, 4, 8, 16, 16, 16, 16, 16, 16, 16, 16, 16, 8, 8, 8, 4, 4, 32};
  const int count = (int)(sizeof(sizes) / sizeof(sizes[0]));
  for (int i = 0; i < count; ++i) {
    ctx->px[i].release();
    if (!sizes[i]) continue;
    CK(ctx->px[i].reserve(sizes[i] * n, false, ctx->stream));
    CK(cudaMemsetAsync(ctx->px[i].p, 0, sizes[i] * n, ctx->stream));
  }
  for (int g = 0; g < GB_PLANES; ++g) {
    ctx->px_alt[g].release();
    CK(ctx->px_alt[g].reserve(sizes[GB_PLANE_IDX[g]] * n, false, ctx->stream));
    CK(cudaMemsetAsync(ctx->px_alt[g].p, 0, sizes[GB_PLANE_IDX[g]] * n, ctx->stream));
  }
  ctx->gb_cur = 0; ctx->gb_calls = 0;
  set_gbuffer_planes(ctx);
  PixelPlanes& P = ctx->planes;
  for (int h = 0; h < 2; ++h) for (int k = 0; k < 3; ++k) P.res[h][k] = ctx->px[12 + h * 3 + k].as<float4>();
  P.radiance = ctx->px[18].as<float4>(); P.direct = ctx->px[19].as<float4>(); P.accum = ctx->px[20].as<float4>();
  P.out16 = ctx->px[21].as<ushort4>();
  P.taa_hist[0] = ctx->px[22].as<ushort4>(); P.taa_hist[1] = ctx->px[23].as<ushort4>();
  P.taa_prev_vel[0] = ctx->px[24].as<ushort2>(); P.taa_prev_vel[1] = ctx->px[25].as<ushort2>();
  // depth planes clear to 1.0 (GBufferPass.cpp:79); prim ids to "sky"; reservoirs to empty
  std::vector<float> ones(n, 1.0f);
  CK(cudaMemcpyAsync(P.depth, ones.data(), 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(P.prev_depth, ones.data(), 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(P.prim, 0xFF, 4 * n, ctx->stream));
  CK(cudaMemsetAsync(P.inst, 0xFF, 4 * n, ctx->stream));
  CK(cudaMemsetAsync(ctx->px_alt[6].p, 0xFF, 4 * n, ctx->stream));  // prim / inst of the other copy
  CK(cudaMemsetAsync(ctx->px_alt[7].p, 0xFF, 4 * n, ctx->stream));
  std::vector<SnkPixelReservoir> empty(n);
  memset(empty.data(), 0, sizeof(SnkPixelReservoir) * n);
  for (auto& r : empty) { r.pl_idx = r.object_instance_idx = r.triangle_idx = SNKPT_INVALID_INDEX; }
  CK(cudaStreamSynchronize(ctx->stream));
  for (uint32_t h = 0; h < 2; ++h) {
    int rc = snkpt_write_reservoirs(ctx, h, empty.data(), 0, (uint32_t)n);
    if (rc) return rc;
  }
  ctx->history_valid = false;
  ctx->gbuffer_valid = false;
  ctx->accum_count = 0;
  ctx->taa_valid = false; ctx->taa_cur = 0;
  return SNKPT_OK;
}

// ---- asynchronous small uploads ---------------------------------------------------------------------------------
static cudaStream_t work_stream(SnkPt* ctx) { return ctx->work_stream ? ctx->work_stream : ctx->stream; }
// a frame call on another stream than the previous one: order it after everything enqueued so far (uploads, earlier frames)
static void note_stream(SnkPt* ctx, cudaStream_t st) {
  cudaStream_t prev = work_stream(ctx);
  if (prev != st) {
    if (!ctx->ev_frame_done) cudaEventCreateWithFlags(&ctx->ev_frame_done, cudaEventDisableTiming);
    cudaEventRecord(ctx->ev_frame_done, prev);
    cudaStreamWaitEvent(st, ctx->ev_frame_done, 0);
  }
  ctx->work_stream = st;
}

// object-space bounding sphere of a vertex range: centre of its box, largest distance to it (float64, rounded outwards later)
static void sphere_of_vertices(const float* pos, uint32_t count, BlasSphere& sp) {
  sp.r = -1.0;
  if (!pos || count == 0) return;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (uint32_t v = 0; v < count; ++v)
    for (int k = 0; k < 3; ++k) {
      const double x = pos[3ull * v + k];
      if (!(x == x) || x > 1e30 || x < -1e30) return;  // NaN / huge coordinates: no sphere, no culling
      lo[k] = x < lo[k] ? x : lo[k]; hi[k] = x > hi[k] ? x : hi[k];
    }
  for (int k = 0; k < 3; ++k) sp.c[k] = 0.5 * (lo[k] + hi[k]);
  double r2 = 0.0;
  for (uint32_t v = 0; v < count; ++v) {
    double d2 = 0.0;
    for (int k = 0; k < 3; ++k) { const double d = pos[3ull * v + k] - sp.c[k]; d2 += d * d; }
    r2 = d2 > r2 ? d2 : r2;
  }
  sp.r = sqrt(r2);
}

// world-space bounding sphere of an instance: centre = M c, radius = r * (upper bound of the largest singular value of mat3(M))
static void world_sphere(const float* M /* column-major 4x4 */, const BlasSphere& sp, float out[4]) {
  out[0] = out[1] = out[2] = 0.f; out[3] = -1.f;
  if (sp.r < 0.0) return;
  double B[3][3];  // B = A^T A, A = upper 3x3 (columns M[0..2], M[4..6], M[8..10])
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += (double)M[4 * i + k] * (double)M[4 * j + k];
      B[i][j] = s;
    }
  // largest eigenvalue of B <= min(largest Gershgorin disc, trace): exact for a rotation times a uniform scale
  double g = 0.0, tr = B[0][0] + B[1][1] + B[2][2];
  for (int i = 0; i < 3; ++i) { const double row = fabs(B[i][0]) + fabs(B[i][1]) + fabs(B[i][2]); g = row > g ? row : g; }
  const double lam = g < tr ? g : tr;
  if (!(lam == lam) || lam > 1e30) return;
  double c[3];
  for (int k = 0; k < 3; ++k) c[k] = (double)M[k] * sp.c[0] + (double)M[4 + k] * sp.c[1] + (double)M[8 + k] * sp.c[2] + (double)M[12 + k];
  const double cmax = fmax(fabs(c[0]), fmax(fabs(c[1]), fabs(c[2])));
  // outward slack: float32 rounding of the centre and of the device arithmetic (relative), vertex transform error (absolute)
  const double r = sp.r * sqrt(lam) * (1.0 + 2e-4) + 4e-6 * cmax + 1e-30;
  out[0] = (float)c[0]; out[1] = (float)c[1]; out[2] = (float)c[2];
  out[3] = (float)(r * r * (1.0 + 1e-6));
}

// Entry points that rewrite data an in-flight frame may still read (vertex positions, BVH nodes, textures, reservoirs) first wait
// for the stream the frames run on: the host is allowed to run a frame ahead of the GPU (the internal chain streams of the
// north-star frame always join that stream before the frame ends).
static cudaError_t quiesce_frames(SnkPt* ctx) {
  cudaError_t e = cudaStreamSynchronize(work_stream(ctx));
  if (e == cudaSuccess && work_stream(ctx) != ctx->stream) e = cudaStreamSynchronize(ctx->stream);
  return e;
}

static int upload_async(SnkPt* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return SNKPT_OK;
  cudaStream_t st = work_stream(ctx);
  if (bytes > SnkPt::RING_CHUNK_BYTES) {  // large and rare: plain copy, caller's memory must outlive it
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return SNKPT_OK;
  }
  if (!ctx->ring) {
    CK(cudaMallocHost((void**)&ctx->ring, SnkPt::RING_CHUNK_BYTES * SnkPt::RING_CHUNKS));
    for (int i = 0; i < SnkPt::RING_CHUNKS; ++i) CK(cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
  }
  if (ctx->ring_used + bytes > SnkPt::RING_CHUNK_BYTES) {
    // leave the chunk: everything copied out of it so far is ordered before this event
    CK(cudaEventRecord(ctx->ring_ev[ctx->ring_cur], st));
    ctx->ring_armed[ctx->ring_cur] = true;
    ctx->ring_cur = (ctx->ring_cur + 1) % SnkPt::RING_CHUNKS;
    if (ctx->ring_armed[ctx->ring_cur]) { CK(cudaEventSynchronize(ctx->ring_ev[ctx->ring_cur])); ctx->ring_armed[ctx->ring_cur] = false; }
    ctx->ring_used = 0;
  }
  uint8_t* slot = ctx->ring + (size_t)ctx->ring_cur * SnkPt::RING_CHUNK_BYTES + ctx->ring_used;
  memcpy(slot, src, bytes);
  CK(cudaMemcpyAsync(dst, slot, bytes, cudaMemcpyHostToDevice, st));
  ctx->ring_used += (bytes + 255) & ~(size_t)255;
  return SNKPT_OK;
}

namespace snk {
void wait_readback(SnkPt* ctx, cudaStream_t st) {
  for (int k = 0; k < 4; ++k)
    if (ctx->read_pending[k]) cudaStreamWaitEvent(st, ctx->ev_read_done[k], 0);
}
}  // namespace snk

extern "C" {

int snkpt_create(const SnkPtConfig* cfg, SnkPt** out) {
  if (!cfg || !out || cfg->width == 0 || cfg->height == 0) { g_create_error = "snkpt_create: bad config"; return SNKPT_E_INVALID; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = std::string("snkpt_create: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range") +
                     "); this library has no CPU path";
    return SNKPT_E_CUDA;
  }
  SnkPt* ctx = new (std::nothrow) SnkPt();
  if (!ctx) return SNKPT_E_NOMEM;
  ctx->device = cfg->device;
  e = cudaSetDevice(cfg->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return SNKPT_E_CUDA; }
  ctx->W = cfg->width; ctx->H = cfg->height;
  ctx->row0 = cfg->row_begin; ctx->row1 = cfg->row_end ? cfg->row_end : cfg->height;
  if (ctx->row1 > ctx->H || ctx->row0 >= ctx->row1) { g_create_error = "snkpt_create: bad row band"; delete ctx; return SNKPT_E_INVALID; }
  if (cfg->max_instances) ctx->max_instances = cfg->max_instances;
  if (cfg->max_transforms) ctx->max_transforms = cfg->max_transforms;
  if (cfg->max_materials) ctx->max_materials = cfg->max_materials;
  int rc = alloc_pixels(ctx);
  if (rc == SNKPT_OK) {
    // zero-filled light block: no lights until snkpt_set_lights
    cudaError_t e2 = ctx->lights.reserve(sizeof(SnkLightData), false, ctx->stream);
    if (e2 == cudaSuccess) e2 = cudaMemsetAsync(ctx->lights.p, 0, sizeof(SnkLightData), ctx->stream);
    if (e2 == cudaSuccess) e2 = ctx->materials.reserve(sizeof(SnkMaterial) * ctx->max_materials, false, ctx->stream);
    if (e2 == cudaSuccess) e2 = cudaMemsetAsync(ctx->materials.p, 0, sizeof(SnkMaterial) * ctx->max_materials, ctx->stream);
    if (e2 == cudaSuccess) e2 = ctx->dev_counters.reserve(1024, false, ctx->stream);
    if (e2 == cudaSuccess) e2 = cudaMemsetAsync(ctx->dev_counters.p, 0, 1024, ctx->stream);
    if (e2 == cudaSuccess) e2 = ctx->trav_spill.reserve(3 * trace_spill_bytes(), false, ctx->stream);  // regions: path chain | ReSTIR chain | camera rays running ahead
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
    if (e2 != cudaSuccess) { ctx->err = cudaGetErrorString(e2); rc = SNKPT_E_CUDA; }
  }
  if (rc != SNKPT_OK) { g_create_error = ctx->err; snkpt_destroy(ctx); return rc; }
  *out = ctx;
  return SNKPT_OK;
}

void snkpt_destroy(SnkPt* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->tlas_stream) { cudaStreamSynchronize(ctx->tlas_stream); cudaStreamDestroy(ctx->tlas_stream); ctx->tlas_stream = nullptr; }
  for (cudaEvent_t* e : {&ctx->ev_tlas_prev[0], &ctx->ev_tlas_prev[1], &ctx->ev_tlas_built, &ctx->ev_tlas_sync, &ctx->ev_gb_prev[0], &ctx->ev_gb_prev[1], &ctx->ev_gb_done, &ctx->ev_gb_sync, &ctx->ev_prev_free})
    if (*e) { cudaEventDestroy(*e); *e = nullptr; }
  DevBuf* bufs[] = {&ctx->positions, &ctx->normals, &ctx->tangents, &ctx->uvs, &ctx->indices, &ctx->blas_nodes, &ctx->blas_tris,
                    &ctx->d_blas, &ctx->materials, &ctx->lights, &ctx->transforms, &ctx->prev_transforms, &ctx->normal_mats,
                    &ctx->det3, &ctx->instances, &ctx->blas_of_instance, &ctx->emissive_idx, &ctx->tlas_nodes, &ctx->tlas_leaf_ids,
                    &ctx->inst_recs, &ctx->inst_sphere, &ctx->valid_slots, &ctx->tlas_stage, &ctx->tlas_nodes_alt, &ctx->tlas_leaf_ids_alt, &ctx->tlas_stage_alt, &ctx->q_rays, &ctx->q_hits, &ctx->q_work, &ctx->dev_counters, &ctx->trav_spill,
                    &ctx->wave_refs, &ctx->wave_paths, &ctx->prof_work, &ctx->prof_hits, &ctx->em_table, &ctx->em_base_of_obj, &ctx->em_base, &ctx->tex_table, &ctx->srgb_lut};
  for (auto& b : ctx->tex_data) b.release();
  for (DevBuf* b : bufs) b->release();
  for (auto& b : ctx->wave) b.release();
  for (auto& b : ctx->px) b.release();
  for (auto& b : ctx->px_alt) b.release();
  for (auto& r : ctx->blas_refit) refit_free(r);
  scratch_free(ctx->scratch_blas);
  scratch_free(ctx->scratch_tlas);
  scratch_free(ctx->scratch_ptcl);
  for (auto& b : ctx->ptcl) b.release();
  ctx->ptcl_start.release(); ctx->ptcl_rays.release();
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
  if (ctx->ring) cudaFreeHost(ctx->ring);
  for (auto& ev : ctx->ring_ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev_read_done) if (ev) cudaEventDestroy(ev);
  if (ctx->ev_frame_done) cudaEventDestroy(ctx->ev_frame_done);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
  group_free(ctx);
  if (ctx->hi_stream) { cudaStreamSynchronize(ctx->hi_stream); cudaStreamDestroy(ctx->hi_stream); }
  if (ctx->ev_join_b) cudaEventDestroy(ctx->ev_join_b);
  for (cudaEvent_t& e : ctx->ev_path) if (e) { cudaEventDestroy(e); e = nullptr; }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  for (auto& r : ctx->prof) { if (r.e0) cudaEventDestroy(r.e0); if (r.e1) cudaEventDestroy(r.e1); }
  ctx->prof_work.release(); ctx->prof_hits.release();
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* snkpt_last_error(const SnkPt* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void snkpt_default_settings(SnkRtSettings* s) {
  if (!s) return;
  memset(s, 0, sizeof(*s));
  s->num_temporal_resamples = 1;  // Raytracing.h:14
  s->num_spatial_resamples = 3;   // Raytracing.h:15
  s->mode = SNKPT_MODE_REFERENCE;
  s->ris_iters = 32;              // RayGen.rgen:95
  s->m_cap = 20;                  // RayGen.rgen:97
  s->neighbour_radius = 30.0f;    // RayGen.rgen:488
  s->normal_threshold = 0.95f;    // RayGen.rgen:435, :491
  s->max_bounces = 4;
  s->spp = 1;
}

int snkpt_resize(SnkPt* ctx, uint32_t width, uint32_t height) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(width && height, SNKPT_E_INVALID, "snkpt_resize: zero extent");
  CK(cudaSetDevice(ctx->device));
  // the planes are freed and reallocated: nothing that reads or writes them may still be in flight — frames on the caller's
  // stream, uploads, asynchronous read-backs on the copy stream (EditorLayer.cpp:351-366 waits for the device as well)
  CK(cudaStreamSynchronize(work_stream(ctx)));
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->W = width; ctx->H = height; ctx->row0 = 0; ctx->row1 = height;
  const int rc = alloc_pixels(ctx);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));  // the cleared planes are in place before a frame on any other stream starts
  return SNKPT_OK;
}

int snkpt_set_rows(SnkPt* ctx, uint32_t row_begin, uint32_t row_end) {
  if (!ctx) return SNKPT_E_INVALID;
  if (row_end == 0) row_end = ctx->H;
  REQUIRE(row_begin < row_end && row_end <= ctx->H, SNKPT_E_INVALID, "snkpt_set_rows: bad band");
  ctx->row0 = row_begin; ctx->row1 = row_end;
  return SNKPT_OK;
}

int snkpt_set_halo(SnkPt* ctx, uint32_t pass1_halo, uint32_t gbuffer_halo) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(!ctx->group, SNKPT_E_STATE, "snkpt_set_halo: the context belongs to a row-band group, which sets its own halo");
  REQUIRE(gbuffer_halo >= pass1_halo, SNKPT_E_INVALID, "snkpt_set_halo: the G-buffer halo must cover the pass-1 halo");
  ctx->halo_pass1 = pass1_halo; ctx->halo_gbuffer = gbuffer_halo;
  return SNKPT_OK;
}

int snkpt_reset_history(SnkPt* ctx) {
  if (!ctx) return SNKPT_E_INVALID;
  ctx->history_valid = false;
  ctx->taa_valid = false;
  return SNKPT_OK;
}

// ---- meshes + BLAS ------------------------------------------------------------------------------
static int build_blas(SnkPt* ctx, uint32_t blas_id, bool fresh, cudaStream_t st) {
  const BlasSrc src = ctx->blas_src[blas_id];
  const uint32_t n = src.num_tris;
  SnkBlasInfo& info = ctx->blas[blas_id];
  if (fresh) {
    info.node_root = ctx->n_blas_nodes;
    info.tri_base = ctx->n_blas_tris;
  }
  const uint32_t node_cap = n ? n : 1;
  CK(ctx->blas_nodes.reserve(sizeof(SnkNode8) * ((size_t)info.node_root + node_cap), true, st));
  CK(ctx->blas_tris.reserve(sizeof(SnkTri) * ((size_t)info.tri_base + n + 1), true, st));
  CK(scratch_reserve(ctx->scratch_blas, n ? n : 1));
  BuildScratch& s = ctx->scratch_blas;
  cudaEventRecord(ctx->ev[6], st);
  launch_triangle_bounds(s, ctx->positions.as<float>(), ctx->indices.as<uint32_t>(), src.index_offset, src.vertex_offset, n, st);
  ctx->stats.kernel_launches += n ? 2 : 1;
  if (n == 0) {
    // empty BLAS: a root with no children
    SnkNode8 nd; memset(&nd, 0, sizeof(nd));
    for (int k = 0; k < 8; ++k) { nd.qlox[k] = nd.qloy[k] = nd.qloz[k] = 255; }
    nd.ex = nd.ey = nd.ez = 127;
    CK(cudaMemcpyAsync(ctx->blas_nodes.as<SnkNode8>() + info.node_root, &nd, sizeof(nd), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    info.node_count = 1; info.tri_count = 0;
    for (int k = 0; k < 3; ++k) { info.bounds_lo[k] = 0; info.bounds_hi[k] = 0; }
    if (fresh) { ctx->n_blas_nodes += 1; }
    return SNKPT_OK;
  }
  BuildOutput out;
  out.nodes = ctx->blas_nodes.as<SnkNode8>(); out.node_offset = info.node_root; out.node_capacity = node_cap;
  out.tris = ctx->blas_tris.as<SnkTri>(); out.leaf_ids = nullptr; out.leaf_remap = nullptr; out.leaf_offset = info.tri_base;
  const int built = launch_build(s, n, out, ctx->positions.as<float>(), ctx->indices.as<uint32_t>(), src.index_offset,
                                             src.vertex_offset, st);
  REQUIRE(built >= 0, SNKPT_E_CUDA, "BLAS build: the cooperative collapse kernel could not be launched");
  ctx->stats.kernel_launches += built;
  cudaEventRecord(ctx->ev[7], st);
  uint32_t counters[16];
  CK(cudaMemcpyAsync(counters, s.counters, sizeof(counters), cudaMemcpyDeviceToHost, st));
  float4 root_lo, root_hi;
  // binary root box = box of node 0 (n > 1) or of the only leaf (also index 0)
  CK(cudaMemcpyAsync(&root_lo, s.bb_lo, sizeof(float4), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&root_hi, s.bb_hi, sizeof(float4), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  cudaEventElapsedTime(&ctx->stats.ms_build_last, ctx->ev[6], ctx->ev[7]);
  REQUIRE(counters[4] == 0, SNKPT_E_NOMEM, "BLAS build: node pool overflow");
  REQUIRE(counters[1] == n, SNKPT_E_STATE, "BLAS build: not every primitive reached a leaf");
  info.node_count = counters[0];
  info.tri_count = counters[1];
  info.bounds_lo[0] = root_lo.x; info.bounds_lo[1] = root_lo.y; info.bounds_lo[2] = root_lo.z;
  info.bounds_hi[0] = root_hi.x; info.bounds_hi[1] = root_hi.y; info.bounds_hi[2] = root_hi.z;
  if (fresh) { ctx->n_blas_nodes += info.node_count; ctx->n_blas_tris += info.tri_count; }
  return SNKPT_OK;
}

static int upload_blas_table(SnkPt* ctx, cudaStream_t st) {
  if (ctx->blas.empty()) return SNKPT_OK;
  CK(ctx->d_blas.reserve(sizeof(SnkBlasInfo) * ctx->blas.size(), false, st));
  CK(cudaMemcpyAsync(ctx->d_blas.p, ctx->blas.data(), sizeof(SnkBlasInfo) * ctx->blas.size(), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  return SNKPT_OK;
}

int snkpt_upload_mesh(SnkPt* ctx, const float* positions, const float* normals, const float* tangents, const float* uvs,
                      uint32_t num_vertices, const uint32_t* indices, uint32_t num_indices, const SnkSubmesh* submeshes,
                      uint32_t num_submeshes, SnkMeshEntry* out_entry, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;  // the next TLAS rebuild must be ordered after this call's device work (snkpt_build_tlas)
  REQUIRE(positions && indices && submeshes && out_entry, SNKPT_E_INVALID, "snkpt_upload_mesh: null argument");
  REQUIRE(num_submeshes >= 1 && num_submeshes <= SNKPT_MAX_SUBMESHES, SNKPT_E_INVALID, "snkpt_upload_mesh: submesh count");
  REQUIRE(num_indices % 3 == 0, SNKPT_E_INVALID, "snkpt_upload_mesh: indices not a multiple of 3");
  for (uint32_t s = 0; s < num_submeshes; ++s) {
    REQUIRE(submeshes[s].num_indices % 3 == 0 && (uint64_t)submeshes[s].base_index + submeshes[s].num_indices <= num_indices &&
                (uint64_t)submeshes[s].base_vertex + submeshes[s].num_vertices <= num_vertices,
            SNKPT_E_INVALID, "snkpt_upload_mesh: submesh range out of bounds");
    for (uint32_t k = 0; k < submeshes[s].num_indices; ++k)
      REQUIRE((uint64_t)indices[submeshes[s].base_index + k] + submeshes[s].base_vertex < num_vertices, SNKPT_E_INVALID,
              "snkpt_upload_mesh: index out of range");
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  CK(quiesce_frames(ctx));  // the pools may be re-allocated (MeshBufferManager::Resize) under frames in flight
  note_stream(ctx, st);
  // append to the five global SoA buffers (MeshBufferManager.cpp:49-70)
  const size_t v0 = ctx->num_vertices, i0 = ctx->num_indices, nv = num_vertices, ni = num_indices;
  CK(ctx->positions.reserve(12 * (v0 + nv), true, st)); CK(ctx->normals.reserve(12 * (v0 + nv), true, st));
  CK(ctx->tangents.reserve(12 * (v0 + nv), true, st)); CK(ctx->uvs.reserve(8 * (v0 + nv), true, st));
  CK(ctx->indices.reserve(4 * (i0 + ni), true, st));
  CK(cudaMemcpyAsync(ctx->positions.as<char>() + 12 * v0, positions, 12 * nv, cudaMemcpyHostToDevice, st));
  if (normals) CK(cudaMemcpyAsync(ctx->normals.as<char>() + 12 * v0, normals, 12 * nv, cudaMemcpyHostToDevice, st));
  else CK(cudaMemsetAsync(ctx->normals.as<char>() + 12 * v0, 0, 12 * nv, st));
  if (tangents) CK(cudaMemcpyAsync(ctx->tangents.as<char>() + 12 * v0, tangents, 12 * nv, cudaMemcpyHostToDevice, st));
  else CK(cudaMemsetAsync(ctx->tangents.as<char>() + 12 * v0, 0, 12 * nv, st));
  if (uvs) CK(cudaMemcpyAsync(ctx->uvs.as<char>() + 8 * v0, uvs, 8 * nv, cudaMemcpyHostToDevice, st));
  else CK(cudaMemsetAsync(ctx->uvs.as<char>() + 8 * v0, 0, 8 * nv, st));
  CK(cudaMemcpyAsync(ctx->indices.as<char>() + 4 * i0, indices, 4 * ni, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  ctx->num_vertices += num_vertices; ctx->num_indices += num_indices;
  memset(out_entry, 0, sizeof(*out_entry));
  out_entry->data_start_indices_idx = (uint32_t)i0;
  out_entry->data_start_vertex_idx = (uint32_t)v0;
  out_entry->num_submeshes = num_submeshes;
  for (uint32_t s = 0; s < num_submeshes; ++s) {
    const uint32_t id = (uint32_t)ctx->blas.size();
    ctx->blas.push_back(SnkBlasInfo{});
    ctx->blas_src.push_back(BlasSrc{(uint32_t)i0 + submeshes[s].base_index, (uint32_t)v0 + submeshes[s].base_vertex, submeshes[s].num_indices / 3});
    BlasSphere sph;
    sph.v_begin = (uint32_t)v0 + submeshes[s].base_vertex; sph.v_count = submeshes[s].num_vertices;
    sphere_of_vertices(positions + 3ull * submeshes[s].base_vertex, submeshes[s].num_vertices, sph);
    ctx->blas_sphere.push_back(sph);
    out_entry->blas_id[s] = id;
    int rc = build_blas(ctx, id, true, st);
    if (rc) return rc;
  }
  ctx->stats.blas_nodes = ctx->n_blas_nodes; ctx->stats.blas_tris = ctx->n_blas_tris;
  ctx->tlas_built = false;
  return upload_blas_table(ctx, st);
}

int snkpt_update_positions(SnkPt* ctx, const SnkMeshEntry* entry, const float* positions, uint32_t num_vertices) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;  // the next TLAS rebuild must be ordered after this call's device work (snkpt_build_tlas)
  REQUIRE(entry && positions, SNKPT_E_INVALID, "snkpt_update_positions: null argument");
  REQUIRE((uint64_t)entry->data_start_vertex_idx + num_vertices <= ctx->num_vertices, SNKPT_E_INVALID, "snkpt_update_positions: range");
  CK(cudaSetDevice(ctx->device));
  CK(quiesce_frames(ctx));  // the frame's shading and emissive sampling read these positions
  cudaStream_t st = work_stream(ctx);
  CK(cudaMemcpyAsync(ctx->positions.as<char>() + 12ull * entry->data_start_vertex_idx, positions, 12ull * num_vertices, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  // the bounding spheres of the BLASes whose vertices moved (instance culling in the traversal)
  for (BlasSphere& sph : ctx->blas_sphere) {
    const uint64_t b0 = entry->data_start_vertex_idx, b1 = b0 + num_vertices;
    if (sph.v_begin + (uint64_t)sph.v_count <= b0 || sph.v_begin >= b1) continue;
    if (sph.v_begin >= b0 && sph.v_begin + (uint64_t)sph.v_count <= b1) sphere_of_vertices(positions + 3ull * (sph.v_begin - b0), sph.v_count, sph);
    else sph.r = -1.0;  // partially updated range: unknown
  }
  ctx->tlas_built = false;
  return SNKPT_OK;
}

int snkpt_rebuild_blas(SnkPt* ctx, uint32_t blas_id, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;  // the next TLAS rebuild must be ordered after this call's device work (snkpt_build_tlas)
  REQUIRE(blas_id < ctx->blas.size(), SNKPT_E_INVALID, "snkpt_rebuild_blas: bad id");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  CK(quiesce_frames(ctx));  // frames in flight traverse these nodes
  note_stream(ctx, st);
  // rebuild in place: the node/triangle regions of this BLAS were sized for its worst case at first build
  // only when it is the last BLAS in the pools; otherwise append a fresh copy
  const SnkBlasInfo old = ctx->blas[blas_id];
  const bool last = old.node_root + old.node_count == ctx->n_blas_nodes && old.tri_base + old.tri_count == ctx->n_blas_tris;
  if (last) { ctx->n_blas_nodes = old.node_root; ctx->n_blas_tris = old.tri_base; }
  int rc = build_blas(ctx, blas_id, true, st);
  if (rc) return rc;
  if (blas_id < ctx->blas_refit.size()) refit_free(ctx->blas_refit[blas_id]);  // topology changed
  ctx->stats.blas_nodes = ctx->n_blas_nodes; ctx->stats.blas_tris = ctx->n_blas_tris;
  ctx->tlas_built = false;
  return upload_blas_table(ctx, st);
}

int snkpt_refit_blas(SnkPt* ctx, uint32_t blas_id, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;  // the next TLAS rebuild must be ordered after this call's device work (snkpt_build_tlas)
  REQUIRE(blas_id < ctx->blas.size(), SNKPT_E_INVALID, "snkpt_refit_blas: bad id");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  CK(quiesce_frames(ctx));  // frames in flight traverse these nodes and triangles
  note_stream(ctx, st);
  SnkBlasInfo& info = ctx->blas[blas_id];
  if (info.tri_count == 0) return SNKPT_OK;
  if (ctx->blas_refit.size() < ctx->blas.size()) ctx->blas_refit.resize(ctx->blas.size());
  RefitState& r = ctx->blas_refit[blas_id];
  const BlasSrc src = ctx->blas_src[blas_id];
  cudaEventRecord(ctx->ev[6], st);
  const int n = launch_refit(r, ctx->blas_nodes.as<SnkNode8>(), info.node_root, info.node_count, ctx->blas_tris.as<SnkTri>(), info.tri_base,
                             info.tri_count, ctx->positions.as<float>(), ctx->indices.as<uint32_t>(), src.index_offset, src.vertex_offset, st);
  REQUIRE(n >= 0, SNKPT_E_NOMEM, "snkpt_refit_blas: out of device memory");
  ctx->stats.kernel_launches += n;
  cudaEventRecord(ctx->ev[7], st);
  float4 box[2];
  CK(cudaMemcpyAsync(box, r.node_box, sizeof(box), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  cudaEventElapsedTime(&ctx->stats.ms_build_last, ctx->ev[6], ctx->ev[7]);
  info.bounds_lo[0] = box[0].x; info.bounds_lo[1] = box[0].y; info.bounds_lo[2] = box[0].z;
  info.bounds_hi[0] = box[1].x; info.bounds_hi[1] = box[1].y; info.bounds_hi[2] = box[1].z;
  ctx->tlas_built = false;  // instance boxes derive from the BLAS bounds
  return upload_blas_table(ctx, st);
}

// ---- scene buffers -----------------------------------------------------------------------------------
int snkpt_set_materials(SnkPt* ctx, const SnkMaterial* materials, uint32_t first, uint32_t count) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;
  REQUIRE(materials && (uint64_t)first + count <= ctx->max_materials, SNKPT_E_INVALID, "snkpt_set_materials: range");
  CK(cudaSetDevice(ctx->device));
  int rc = upload_async(ctx, ctx->materials.as<SnkMaterial>() + first, materials, sizeof(SnkMaterial) * count);
  if (rc) return rc;
  return SNKPT_OK;
}

int snkpt_set_texture(SnkPt* ctx, uint32_t tex_idx, uint32_t width, uint32_t height, const void* rgba8, uint32_t srgb) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;
  REQUIRE(tex_idx < 4096, SNKPT_E_INVALID, "snkpt_set_texture: index (the descriptor array has 4096 slots)");
  REQUIRE((width == 0 && height == 0) || (rgba8 && width <= 16384 && height <= 16384 && width && height), SNKPT_E_INVALID, "snkpt_set_texture: size / null texels");
  CK(cudaSetDevice(ctx->device));
  CK(quiesce_frames(ctx));  // frames in flight sample the table and the texels
  if (ctx->h_tex.size() <= tex_idx) { ctx->h_tex.resize(tex_idx + 1, TexDev{nullptr, 0, 0, 0, 0}); ctx->tex_data.resize(tex_idx + 1); }
  DevBuf& buf = ctx->tex_data[tex_idx];
  buf.release();
  TexDev& td = ctx->h_tex[tex_idx];
  td = TexDev{nullptr, 0, 0, 0, 0};
  if (width) {
    const size_t bytes = 4ull * width * height;
    CK(buf.reserve(bytes, false, ctx->stream));
    CK(cudaMemcpyAsync(buf.p, rgba8, bytes, cudaMemcpyHostToDevice, ctx->stream));
    td = TexDev{buf.as<uchar4>(), width, height, srgb ? 1u : 0u, 0u};
  }
  if (!ctx->srgb_lut.p) {
    // sRGB EOTF in float64, rounded once (oracle/snk_scene.h:srgb8_to_linear evaluates the same expression)
    float lut[256];
    for (int c = 0; c < 256; ++c) {
      const double x = (double)c / 255.0;
      lut[c] = (float)(x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4));
    }
    CK(ctx->srgb_lut.reserve(sizeof(lut), false, ctx->stream));
    CK(cudaMemcpyAsync(ctx->srgb_lut.p, lut, sizeof(lut), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(ctx->tex_table.reserve(sizeof(TexDev) * 4096, false, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tex_table.p, ctx->h_tex.data(), sizeof(TexDev) * ctx->h_tex.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SNKPT_OK;
}

int snkpt_set_lights(SnkPt* ctx, const void* light_data, size_t bytes) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;
  REQUIRE(light_data && bytes >= 112 && bytes <= sizeof(SnkLightData), SNKPT_E_INVALID, "snkpt_set_lights: size");
  const SnkLightData* L = (const SnkLightData*)light_data;
  REQUIRE(L->num_pointlights <= SNKPT_MAX_POINTLIGHTS && 112 + (size_t)L->num_pointlights * 48 <= bytes, SNKPT_E_INVALID,
          "snkpt_set_lights: point-light count exceeds the supplied bytes");
  CK(cudaSetDevice(ctx->device));
  int rc = upload_async(ctx, ctx->lights.p, light_data, bytes);
  if (rc) return rc;
  ctx->num_pointlights = L->num_pointlights;
  ctx->has_dir_light = L->dir_light.colour[0] > 0.0f || L->dir_light.colour[1] > 0.0f || L->dir_light.colour[2] > 0.0f;  // as ns_light_counts
  return SNKPT_OK;
}

int snkpt_set_transforms(SnkPt* ctx, const float* cur, const float* prev, uint32_t first, uint32_t count) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;  // the next TLAS rebuild must be ordered after this call's device work (snkpt_build_tlas)
  REQUIRE(cur && (uint64_t)first + count <= ctx->max_transforms, SNKPT_E_INVALID, "snkpt_set_transforms: range");
  CK(cudaSetDevice(ctx->device));
  const size_t need = (size_t)first + count;
  if (ctx->h_transforms.size() < need * 16) { ctx->h_transforms.resize(need * 16, 0.f); ctx->h_prev_transforms.resize(need * 16, 0.f); }
  memcpy(ctx->h_transforms.data() + 16 * (size_t)first, cur, 64 * (size_t)count);
  memcpy(ctx->h_prev_transforms.data() + 16 * (size_t)first, prev ? prev : cur, 64 * (size_t)count);
  CK(ctx->transforms.reserve(64 * (size_t)ctx->max_transforms, true, ctx->stream));
  CK(ctx->prev_transforms.reserve(64 * (size_t)ctx->max_transforms, true, ctx->stream));
  CK(ctx->normal_mats.reserve(36 * (size_t)ctx->max_transforms, true, ctx->stream));
  CK(ctx->det3.reserve(4 * (size_t)ctx->max_transforms, true, ctx->stream));
  std::vector<float> nm(9 * (size_t)count), det(count), w2o(12);
  for (uint32_t i = 0; i < count; ++i) derive_transform(cur + 16 * (size_t)i, w2o.data(), nm.data() + 9 * (size_t)i, &det[i]);
  int rc = upload_async(ctx, ctx->transforms.as<float>() + 16 * (size_t)first, cur, 64 * (size_t)count);
  if (!rc) rc = upload_async(ctx, ctx->prev_transforms.as<float>() + 16 * (size_t)first, prev ? prev : cur, 64 * (size_t)count);
  if (!rc) rc = upload_async(ctx, ctx->normal_mats.as<float>() + 9 * (size_t)first, nm.data(), 36 * (size_t)count);
  if (!rc) rc = upload_async(ctx, ctx->det3.as<float>() + first, det.data(), 4 * (size_t)count);
  if (rc) return rc;
  ctx->tlas_built = false;
  return SNKPT_OK;
}

int snkpt_set_instances(SnkPt* ctx, const SnkInstanceData* instances, const uint32_t* blas_of_instance, uint32_t count,
                        const uint32_t* emissive_idx, uint32_t num_emissive) {
  if (!ctx) return SNKPT_E_INVALID;
  ++ctx->scene_epoch;  // the next TLAS rebuild must be ordered after this call's device work (snkpt_build_tlas)
  REQUIRE(count <= ctx->max_instances && num_emissive <= ctx->max_instances, SNKPT_E_INVALID, "snkpt_set_instances: capacity (4096 by default)");
  REQUIRE((instances && blas_of_instance) || count == 0, SNKPT_E_INVALID, "snkpt_set_instances: null argument");
  for (uint32_t i = 0; i < count; ++i) {
    if (blas_of_instance[i] == SNKPT_INVALID_INDEX) continue;
    REQUIRE(blas_of_instance[i] < ctx->blas.size(), SNKPT_E_INVALID, "snkpt_set_instances: unknown BLAS id");
    const BlasSrc& b = ctx->blas_src[blas_of_instance[i]];
    REQUIRE(instances[i].num_mesh_indices / 3 == b.num_tris && instances[i].mesh_buffer_index_offset == b.index_offset &&
                instances[i].mesh_buffer_vertex_offset == b.vertex_offset,
            SNKPT_E_INVALID, "snkpt_set_instances: InstanceData does not describe the BLAS it references");
    REQUIRE(instances[i].transform_idx < ctx->max_transforms && instances[i].material_idx < ctx->max_materials, SNKPT_E_INVALID,
            "snkpt_set_instances: transform / material index out of range");
  }
  for (uint32_t i = 0; i < num_emissive; ++i) {
    REQUIRE(emissive_idx[i] < count, SNKPT_E_INVALID, "snkpt_set_instances: emissive index out of range");
    // the light-sampling stages pick a triangle of every listed instance (RayGen.rgen:311-342): it needs geometry
    const uint32_t e = emissive_idx[i];
    REQUIRE(blas_of_instance[e] != SNKPT_INVALID_INDEX && instances[e].num_mesh_indices >= 3, SNKPT_E_INVALID,
            "snkpt_set_instances: an emissive-list entry references an instance without a BLAS / without triangles");
  }
  // stored reservoir samples name (instance, triangle) pairs: an emissive instance whose mesh range changed invalidates them
  if (ctx->history_valid)
    for (uint32_t i = 0; i < num_emissive && ctx->history_valid; ++i) {
      const uint32_t e = emissive_idx[i];
      if (e >= ctx->h_instances.size() || ctx->h_instances[e].num_mesh_indices != instances[e].num_mesh_indices ||
          ctx->h_instances[e].mesh_buffer_index_offset != instances[e].mesh_buffer_index_offset ||
          ctx->h_instances[e].mesh_buffer_vertex_offset != instances[e].mesh_buffer_vertex_offset)
        ctx->history_valid = false;
    }
  CK(cudaSetDevice(ctx->device));
  ctx->h_instances.assign(instances, instances + count);
  ctx->h_blas_of_instance.assign(blas_of_instance, blas_of_instance + count);
  ctx->num_instances = count; ctx->num_emissive = num_emissive;
  CK(ctx->instances.reserve(sizeof(SnkInstanceData) * (size_t)ctx->max_instances, false, ctx->stream));
  CK(ctx->blas_of_instance.reserve(4 * (size_t)ctx->max_instances, false, ctx->stream));
  CK(ctx->emissive_idx.reserve(4 * (size_t)ctx->max_instances, false, ctx->stream));
  int rc = SNKPT_OK;
  if (count) {
    rc = upload_async(ctx, ctx->instances.p, instances, sizeof(SnkInstanceData) * count);
    if (!rc) rc = upload_async(ctx, ctx->blas_of_instance.p, blas_of_instance, 4 * (size_t)count);
  }
  if (!rc && num_emissive) rc = upload_async(ctx, ctx->emissive_idx.p, emissive_idx, 4 * (size_t)num_emissive);
  if (rc) return rc;
  // emissive-triangle table: record ranges per emissive-list entry (the first entry of an instance listed twice wins)
  std::vector<uint32_t> base_of_obj(ctx->max_instances, SNKPT_INVALID_INDEX), base(num_emissive ? num_emissive : 1, 0u);
  uint32_t total = 0;
  for (uint32_t k = 0; k < num_emissive; ++k) {
    base[k] = total;
    if (base_of_obj[emissive_idx[k]] == SNKPT_INVALID_INDEX) base_of_obj[emissive_idx[k]] = total;
    total += instances[emissive_idx[k]].num_mesh_indices / 3u;
  }
  ctx->num_emissive_tris = total;
  // stored reservoir samples name (instance, triangle) pairs of the emissive list: a different list invalidates them
  if (ctx->h_emissive != std::vector<uint32_t>(emissive_idx, emissive_idx + num_emissive)) {
    if (!ctx->h_emissive.empty() || num_emissive) ctx->history_valid = false;
    ctx->h_emissive.assign(emissive_idx, emissive_idx + num_emissive);
  }
  CK(ctx->em_base_of_obj.reserve(4 * (size_t)ctx->max_instances, false, ctx->stream));
  CK(ctx->em_base.reserve(4 * (size_t)ctx->max_instances, false, ctx->stream));
  CK(ctx->em_table.reserve(96ull * (total ? total : 1), false, ctx->stream));  // EM_REC_F4 x 16 B per emissive triangle
  rc = upload_async(ctx, ctx->em_base_of_obj.p, base_of_obj.data(), 4 * (size_t)ctx->max_instances);
  if (!rc && num_emissive) rc = upload_async(ctx, ctx->em_base.p, base.data(), 4 * (size_t)num_emissive);
  if (rc) return rc;
  ctx->tlas_built = false;
  return SNKPT_OK;
}

// the stream (and its events) on which the per-frame TLAS rebuild and the camera rays run ahead of the frame in flight
static int ensure_early_stream(SnkPt* ctx) {
  if (ctx->tlas_stream) return SNKPT_OK;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  CK(cudaStreamCreateWithPriority(&ctx->tlas_stream, cudaStreamNonBlocking, hi));
  for (cudaEvent_t* e : {&ctx->ev_tlas_prev[0], &ctx->ev_tlas_prev[1], &ctx->ev_tlas_built, &ctx->ev_tlas_sync, &ctx->ev_gb_prev[0], &ctx->ev_gb_prev[1], &ctx->ev_gb_done, &ctx->ev_gb_sync, &ctx->ev_prev_free})
    CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  return SNKPT_OK;
}

int snkpt_build_tlas(SnkPt* ctx, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);
  // BLAS::GenerateInstance (TLAS.cpp:188-200): one record per instance slot that references a BLAS.
  // The records are staged in pinned host memory owned by the context so that the upload is a true asynchronous copy:
  // no stream synchronisation per frame (the host may run a frame ahead of the GPU). ev_stage guards the reuse.
  const size_t need = (sizeof(SnkInstRec) + 4 + 16) * (size_t)ctx->max_instances;
  if (ctx->h_stage_bytes < need) {
    if (ctx->h_stage) { CK(cudaStreamSynchronize(st)); cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; ctx->h_stage_bytes = 0; }
    CK(cudaMallocHost(&ctx->h_stage, need));
    ctx->h_stage_bytes = need;
    if (!ctx->ev_stage) CK(cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming));
  } else if (ctx->stage_in_flight) {
    CK(cudaEventSynchronize(ctx->ev_stage));
  }
  ctx->stage_in_flight = false;
  // compact layout for THIS frame's instance count ni: records (64 B) | spheres (16 B) | valid slots (4 B), one transfer of 84 ni bytes
  const size_t ni = ((size_t)ctx->num_instances + 3) & ~(size_t)3;
  SnkInstRec* recs = reinterpret_cast<SnkInstRec*>(ctx->h_stage);
  float* spheres = reinterpret_cast<float*>(recs + ni);  // 4 floats per instance slot
  uint32_t* valid = reinterpret_cast<uint32_t*>(spheres + 4 * ni);
  static const bool cull_instances = !(getenv("SNKPT_NO_INSTANCE_CULL") && getenv("SNKPT_NO_INSTANCE_CULL")[0] == '1');
  uint32_t n = 0;
  float nm[9], det;
  for (uint32_t i = 0; i < ctx->num_instances; ++i) {
    SnkInstRec& r = recs[i];
    memset(&r, 0, sizeof(r));
    spheres[4 * i] = spheres[4 * i + 1] = spheres[4 * i + 2] = 0.f; spheres[4 * i + 3] = -1.f;
    const uint32_t b = ctx->h_blas_of_instance[i];
    if (b == SNKPT_INVALID_INDEX || ctx->blas[b].tri_count == 0) continue;
    const uint32_t t = ctx->h_instances[i].transform_idx;
    REQUIRE((size_t)t * 16 + 16 <= ctx->h_transforms.size(), SNKPT_E_STATE, "snkpt_build_tlas: instance references a transform that was never set");
    derive_transform(ctx->h_transforms.data() + 16 * (size_t)t, r.w2o, nm, &det);
    r.blas_root = ctx->blas[b].node_root; r.custom_idx = i; r.blas_id = b; r.flags = 0;
    if (cull_instances && b < ctx->blas_sphere.size()) world_sphere(ctx->h_transforms.data() + 16 * (size_t)t, ctx->blas_sphere[b], spheres + 4 * (size_t)i);
    valid[n++] = i;
  }
  // The rebuild writes the copy of the TLAS storage that the frame BEFORE the one in flight read, on its own stream: it waits for
  // everything enqueued before the previous rebuild call (the last readers of that copy) and — when the scene was changed since —
  // for the caller's stream as it stands, but not for the frame in flight; the caller's stream then waits for the rebuild.
  // A 40-us single-CTA kernel (+ its upload) no longer sits between two frames. SNKPT_TLAS_EARLY=0 restores the in-order rebuild.
  static const bool early_env = !(getenv("SNKPT_TLAS_EARLY") && getenv("SNKPT_TLAS_EARLY")[0] == '0');
  { const int rc = ensure_early_stream(ctx); if (rc) return rc; }
  std::swap(ctx->tlas_nodes, ctx->tlas_nodes_alt);
  std::swap(ctx->tlas_leaf_ids, ctx->tlas_leaf_ids_alt);
  std::swap(ctx->tlas_stage, ctx->tlas_stage_alt);
  cudaStream_t bs = ctx->tlas_stream;
  const uint32_t k = ctx->tlas_builds++;
  const bool early = early_env && k >= 2 && ctx->tlas_epoch == ctx->scene_epoch && (ctx->profiling == 0 || ctx->profiling == 3);
  if (early) {
    CK(cudaStreamWaitEvent(bs, ctx->ev_tlas_prev[(k + 1) & 1], 0));  // recorded at the previous rebuild call: frames <= f-1 are behind it
  } else {
    CK(cudaEventRecord(ctx->ev_tlas_sync, st));                      // in order with everything the caller enqueued so far
    CK(cudaStreamWaitEvent(bs, ctx->ev_tlas_sync, 0));
  }
  ctx->tlas_epoch = ctx->scene_epoch;
  CK(cudaEventRecord(ctx->ev_tlas_prev[k & 1], st));  // before this frame's consumers are enqueued
  // device copy of the staging block, same layout; inst_recs / inst_sphere / valid_slots are views into it
  CK(ctx->tlas_stage.reserve(need, false, bs));
  {
    char* base = static_cast<char*>(ctx->tlas_stage.p);
    ctx->inst_recs.alias(base, sizeof(SnkInstRec) * ni);
    ctx->inst_sphere.alias(base + sizeof(SnkInstRec) * ni, 16 * ni);
    ctx->valid_slots.alias(base + (sizeof(SnkInstRec) + 16) * ni, 4 * ni);
  }
  CK(ctx->tlas_nodes.reserve(sizeof(SnkNode8) * (size_t)(ctx->max_instances + 1), false, bs));
  CK(ctx->tlas_leaf_ids.reserve(4 * (size_t)(ctx->max_instances + 1), false, bs));
  CK(scratch_reserve(ctx->scratch_tlas, ctx->max_instances));
  ctx->tlas_prims = n;
  ctx->tlas_built = true;
  if (n == 0) return SNKPT_OK;
  cudaEventRecord(ctx->ev[6], bs);
  CK(cudaMemcpyAsync(ctx->tlas_stage.p, ctx->h_stage, (sizeof(SnkInstRec) + 16 + 4) * ni, cudaMemcpyHostToDevice, bs));
  CK(cudaEventRecord(ctx->ev_stage, bs));
  ctx->stage_in_flight = true;
  BuildScratch& s = ctx->scratch_tlas;
  launch_instance_bounds(s, ctx->transforms.as<float>(), ctx->instances.as<SnkInstanceData>(), ctx->blas_of_instance.as<uint32_t>(),
                         ctx->d_blas.as<SnkBlasInfo>(), ctx->blas_nodes.as<SnkNode8>(), ctx->valid_slots.as<uint32_t>(), n, bs);
  BuildOutput out;
  out.nodes = ctx->tlas_nodes.as<SnkNode8>(); out.node_offset = 0; out.node_capacity = ctx->max_instances;
  out.tris = nullptr; out.leaf_ids = ctx->tlas_leaf_ids.as<uint32_t>(); out.leaf_remap = ctx->valid_slots.as<uint32_t>(); out.leaf_offset = 0;
  const int built = launch_build(s, n, out, nullptr, nullptr, 0, 0, bs);
  REQUIRE(built >= 0, SNKPT_E_CUDA, "snkpt_build_tlas: the cooperative collapse kernel could not be launched");
  ctx->stats.kernel_launches += 2 + built;
  cudaEventRecord(ctx->ev[7], bs);
  ctx->build_timing_pending = true;  // ms_build_last is read from ev[6..7] by snkpt_get_stats
  CK(cudaEventRecord(ctx->ev_tlas_built, bs));
  CK(cudaStreamWaitEvent(st, ctx->ev_tlas_built, 0));  // the frame's primary rays, traversal launches and queries come after the rebuild
  CK(cudaGetLastError());
  return SNKPT_OK;
}

// ---- per frame --------------------------------------------------------------------------------------------
int snkpt_set_frame(SnkPt* ctx, const SnkCommonUBO* cur, const SnkCommonUBO* prev) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(cur, SNKPT_E_INVALID, "snkpt_set_frame: null UBO");
  ctx->ubo = *cur;
  ctx->prev_ubo = prev ? *prev : *cur;
  FrameConsts& fc = ctx->fc;
  double v[16], p[16], pv[16], inv[16];
  for (int i = 0; i < 16; ++i) { v[i] = cur->view[i]; p[i] = cur->proj[i]; }
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += p[k * 4 + r] * v[c * 4 + k];
      pv[c * 4 + r] = s;
    }
  REQUIRE(invert4(pv, inv), SNKPT_E_INVALID, "snkpt_set_frame: proj*view is singular");
  {
    double pv2[16], inv2[16];
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += (double)ctx->prev_ubo.proj[k * 4 + r] * (double)ctx->prev_ubo.view[c * 4 + k];
        pv2[c * 4 + r] = s;
      }
    const bool ok = invert4(pv2, inv2);
    for (int i = 0; i < 16; ++i) fc.prev_ivp[i] = ok ? (float)inv2[i] : (float)inv[i];
  }
  for (int i = 0; i < 16; ++i) { fc.ivp[i] = (float)inv[i]; fc.proj_view[i] = cur->proj_view[i]; fc.prev_proj_view[i] = ctx->prev_ubo.proj_view[i]; }
  for (int i = 0; i < 3; ++i) fc.cam_pos[i] = cur->cam_pos[i];
  fc.time_elapsed = cur->time_elapsed; fc.delta_time = cur->delta_time;
  fc.frame_idx = cur->frame_idx; fc.prev_frame_invalid = cur->prev_frame_invalid;
  ctx->frame_set = true;
  return SNKPT_OK;
}

static void fill_frame_dims(SnkPt* ctx) {
  ctx->fc.W = ctx->W; ctx->fc.H = ctx->H; ctx->fc.row0 = ctx->row0; ctx->fc.row1 = ctx->row1;
}
// band widened by `halo` rows on both sides, clamped to the image
static void widen_rows(const SnkPt* ctx, uint32_t halo, uint32_t& r0, uint32_t& r1) {
  r0 = ctx->row0 > halo ? ctx->row0 - halo : 0u;
  r1 = ctx->row1 + halo < ctx->H ? ctx->row1 + halo : ctx->H;
}

static float halton(unsigned idx, unsigned base) {  // RenderCommon.h:39-48
  float f = 1.f, r = 0.f;
  while (idx > 0) {
    f *= (float)base;
    r += (float)(idx % base) / f;
    idx /= base;
  }
  return r;
}
void snkpt_frame_jitter(uint32_t frame_idx, uint32_t internal_render_width, uint32_t output_render_width, float* out_jitter_xy, float* out_pixels_xy) {
  float jx = 0.f, jy = 0.f;
  if (internal_render_width != 0 && output_render_width != 0) {  // RenderCommon.h:50-59
    const float scale_ratio = (float)output_render_width / (float)internal_render_width;
    const unsigned phases = (unsigned)(8 * scale_ratio * scale_ratio + 0.5f);
    const unsigned idx = (frame_idx % (phases ? phases : 1u)) + 1;
    jx = halton(idx, 2) - 0.5f; jy = halton(idx, 3) - 0.5f;
  }
  if (out_jitter_xy) { out_jitter_xy[0] = jx; out_jitter_xy[1] = jy; }
  if (out_pixels_xy) { out_pixels_xy[0] = -0.5f * jx; out_pixels_xy[1] = -0.5f * jy; }
}

int snkpt_trace_primary(SnkPt* ctx, const float* jitter_xy, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(ctx->frame_set, SNKPT_E_STATE, "snkpt_trace_primary: call snkpt_set_frame first");
  REQUIRE(ctx->tlas_built, SNKPT_E_STATE, "snkpt_trace_primary: call snkpt_build_tlas first");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);
  fill_frame_dims(ctx);
  ctx->fc.jx = jitter_xy ? jitter_xy[0] : 0.f; ctx->fc.jy = jitter_xy ? jitter_xy[1] : 0.f;
  prof_reset(ctx);
  // The camera rays write the copy of the G-buffer planes that the frame BEFORE the one in flight read, on the stream the TLAS rebuild
  // runs on (after it): like the rebuild they wait for everything enqueued before the previous snkpt_trace_primary call — the last
  // readers of that copy — and for scene changes, not for the frame in flight, whose last launches are latency-bound tails; the caller's
  // stream then waits for them. SNKPT_PRIMARY_EARLY=0 keeps them in order with the caller's stream.
  { const int rc = ensure_early_stream(ctx); if (rc) return rc; }
  static const bool early_env = !(getenv("SNKPT_PRIMARY_EARLY") && getenv("SNKPT_PRIMARY_EARLY")[0] == '0');
  ctx->gb_cur ^= 1u;
  set_gbuffer_planes(ctx);
  cudaStream_t es = ctx->tlas_stream;
  const uint32_t k = ctx->gb_calls++;
  const bool early = early_env && k >= 2 && ctx->gb_epoch == ctx->scene_epoch && (ctx->profiling == 0 || ctx->profiling == 3) && !ctx->group && ctx->prev_free == 1;
  if (early) {
    CK(cudaStreamWaitEvent(es, ctx->ev_gb_prev[(k + 1) & 1], 0));
    if (ctx->prev_free_event) CK(cudaStreamWaitEvent(es, ctx->ev_prev_free, 0));  // the frame in flight is done with the copy written here
  } else {
    CK(cudaEventRecord(ctx->ev_gb_sync, st));
    CK(cudaStreamWaitEvent(es, ctx->ev_gb_sync, 0));
  }
  ctx->gb_epoch = ctx->scene_epoch;
  CK(cudaEventRecord(ctx->ev_gb_prev[k & 1], st));
  cudaEventRecord(ctx->ev[0], es);
  FrameConsts fg = ctx->fc;
  widen_rows(ctx, ctx->halo_gbuffer, fg.row0, fg.row1);
  const int rec = prof_begin(ctx, "primary", (uint64_t)ctx->W * (fg.row1 - fg.row0), false, false, es);
  TraceScene S = ctx->trace_scene();
  S.spill += 2 * (trace_spill_bytes() / sizeof(uint2));  // a spill region of their own: traversal launches of the frame in flight use the other two
  ctx->stats.kernel_launches += launch_primary(S, ctx->scene_dev(), fg, ctx->planes, ctx->dev_counters.as<uint32_t>() + 200, ctx->primary_parity++, es);
  prof_end(ctx, rec, es);
  cudaEventRecord(ctx->ev[1], es);
  CK(cudaEventRecord(ctx->ev_gb_done, es));
  CK(cudaStreamWaitEvent(st, ctx->ev_gb_done, 0));
  CK(cudaGetLastError());
  ctx->stats.rays_primary = (uint64_t)ctx->W * (fg.row1 - fg.row0);
  ctx->stats.total_rays_primary += ctx->stats.rays_primary;
  ctx->gbuffer_valid = true;
  return SNKPT_OK;
}

int snkpt_render(SnkPt* ctx, const SnkRtSettings* settings, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(settings, SNKPT_E_INVALID, "snkpt_render: null settings");
  REQUIRE(ctx->gbuffer_valid, SNKPT_E_STATE, "snkpt_render: call snkpt_trace_primary first");
  REQUIRE(ctx->tlas_built, SNKPT_E_STATE, "snkpt_render: call snkpt_build_tlas first");
  REQUIRE(settings->mode <= SNKPT_MODE_NORTHSTAR, SNKPT_E_INVALID, "snkpt_render: unknown mode");
  REQUIRE(settings->ris_iters >= 1 && settings->ris_iters <= 1024, SNKPT_E_INVALID, "snkpt_render: ris_iters");
  REQUIRE(settings->num_spatial_resamples >= 0 && settings->num_spatial_resamples <= 64, SNKPT_E_INVALID, "snkpt_render: num_spatial_resamples");
  REQUIRE(settings->num_temporal_resamples >= 0, SNKPT_E_INVALID, "snkpt_render: num_temporal_resamples");
  REQUIRE(settings->spp == 1, SNKPT_E_INVALID, "snkpt_render: spp must be 1 (more samples per pixel = more frames + snkpt_accumulate)");
  REQUIRE(settings->m_cap >= 1 && settings->m_cap <= 65535, SNKPT_E_INVALID, "snkpt_render: m_cap");
  REQUIRE(settings->neighbour_radius >= 0.0f && settings->neighbour_radius <= 4096.0f, SNKPT_E_INVALID, "snkpt_render: neighbour_radius (NaN / negative / beyond 4096 pixels)");
  REQUIRE(settings->normal_threshold >= -1.0f && settings->normal_threshold <= 1.0f, SNKPT_E_INVALID, "snkpt_render: normal_threshold is a cosine");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);
  fill_frame_dims(ctx);
  if (ctx->history_mode != settings->mode) ctx->history_valid = false;
  ctx->history_mode = settings->mode;
  if (ctx->group) {
    REQUIRE(settings->mode == SNKPT_MODE_NORTHSTAR, SNKPT_E_STATE, "snkpt_render: a row-band group renders the north-star mode (reference mode: snkpt_set_halo)");
    REQUIRE(settings->neighbour_radius >= 0.0f && (uint32_t)ceilf(settings->neighbour_radius) + 1u <= group_gbuffer_halo(ctx), SNKPT_E_INVALID,
            "snkpt_render: neighbour_radius exceeds the rows the group exchanges");
  }
  { const int erc = ensure_early_stream(ctx); if (erc) return erc; }
  ctx->prev_free = 0; ctx->prev_free_event = false;  // launch_northstar_frame says when the previous-frame normals / depths are no longer read
  int rc = settings->mode == SNKPT_MODE_REFERENCE ? launch_reference_frame(ctx, *settings, st) : launch_northstar_frame(ctx, *settings, st);
  if (rc < 0) return rc;
  CK(cudaGetLastError());
  // a frame without the reservoir stages writes no history: the next ReSTIR frame must not reuse older reservoirs as "previous frame"
  ctx->history_valid = !(settings->flags & SNKPT_RT_NO_RESTIR);
  return SNKPT_OK;
}

int snkpt_accumulate(SnkPt* ctx, int reset, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);
  if (reset) ctx->accum_count = 0;
  wait_readback(ctx, st);
  ctx->stats.kernel_launches += launch_accumulate(ctx->planes, ctx->W * ctx->H, ctx->accum_count, st);
  ctx->accum_count++;
  CK(cudaGetLastError());
  return SNKPT_OK;
}

int snkpt_taa_resolve(SnkPt* ctx, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(ctx->gbuffer_valid, SNKPT_E_STATE, "snkpt_taa_resolve: render a frame first");
  REQUIRE(ctx->row0 == 0 && ctx->row1 == ctx->H, SNKPT_E_STATE, "snkpt_taa_resolve: whole frames only (gather the row bands first)");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);
  const size_t n = (size_t)ctx->W * ctx->H;
  PixelPlanes& P = ctx->planes;
  const uint32_t cur = ctx->taa_cur, nxt = cur ^ 1u;
  if (!ctx->taa_valid) {
    // the reference never initialises the history images: define them as the current frame and its velocity
    CK(cudaMemcpyAsync(P.taa_hist[cur], P.out16, 8 * n, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(P.taa_prev_vel[cur], P.motion, 4 * n, cudaMemcpyDeviceToDevice, st));
  }
  wait_readback(ctx, st);
  const int rec = prof_begin(ctx, "taa_resolve", n, false, false, st);
  ctx->stats.kernel_launches += launch_taa_resolve(P.out16, P.motion, P.taa_hist[cur], P.taa_prev_vel[cur], P.taa_hist[nxt], P.taa_prev_vel[nxt],
                                                   ctx->W, ctx->H, st);
  // the shader stores into the image it reads: the resolved frame becomes the output image (and stays the new history)
  CK(cudaMemcpyAsync(P.out16, P.taa_hist[nxt], 8 * n, cudaMemcpyDeviceToDevice, st));
  prof_end(ctx, rec, st);
  CK(cudaGetLastError());
  ctx->taa_cur = nxt;
  ctx->taa_valid = true;
  return SNKPT_OK;
}

// ---- readback ------------------------------------------------------------------------------------------------
static int aux_desc(SnkPt* ctx, uint32_t what, const void** src, size_t* bytes) {
  const size_t n = (size_t)ctx->W * ctx->H;
  const PixelPlanes& P = ctx->planes;
  switch (what) {
    case SNKPT_AUX_PRIM_ID: *src = P.prim; *bytes = 4 * n; break;
    case SNKPT_AUX_INST_ID: *src = P.inst; *bytes = 4 * n; break;
    case SNKPT_AUX_DEPTH: *src = P.depth; *bytes = 4 * n; break;
    case SNKPT_AUX_NORMAL: *src = P.normal; *bytes = 8 * n; break;
    case SNKPT_AUX_ALBEDO: *src = P.albedo; *bytes = 4 * n; break;
    case SNKPT_AUX_RMA: *src = P.rma; *bytes = 8 * n; break;
    case SNKPT_AUX_MOTION: *src = P.motion; *bytes = 4 * n; break;
    case SNKPT_AUX_MAT_FLAGS: *src = P.flags; *bytes = 2 * n; break;
    case SNKPT_AUX_PREV_NORMAL: *src = P.prev_normal; *bytes = 8 * n; break;
    case SNKPT_AUX_PREV_DEPTH: *src = P.prev_depth; *bytes = 4 * n; break;
    case SNKPT_AUX_BVH_NODES: *src = ctx->blas_nodes.p; *bytes = sizeof(SnkNode8) * (size_t)ctx->n_blas_nodes; break;
    case SNKPT_AUX_BVH_TRIS: *src = ctx->blas_tris.p; *bytes = sizeof(SnkTri) * (size_t)ctx->n_blas_tris; break;
    case SNKPT_AUX_HIT_T: *src = P.hit_t; *bytes = 4 * n; break;
    case SNKPT_AUX_HIT_BARY: *src = P.bary; *bytes = 8 * n; break;
    case SNKPT_AUX_RADIANCE_F32: *src = P.radiance; *bytes = 16 * n; break;
    case SNKPT_AUX_ACCUM_F32: *src = P.accum; *bytes = 16 * n; break;
    case SNKPT_AUX_DIRECT_F32: *src = P.direct; *bytes = 16 * n; break;
    case SNKPT_AUX_TLAS_INSTANCES: *src = ctx->inst_recs.p; *bytes = sizeof(SnkInstRec) * (size_t)ctx->num_instances; break;
    case SNKPT_AUX_RESERVOIRS_FINAL:
    case SNKPT_AUX_RESERVOIRS_TEMPORAL: *src = nullptr; *bytes = 48 * n; break;
    case SNKPT_AUX_TAA_HISTORY: *src = P.taa_hist[ctx->taa_cur]; *bytes = 8 * n; break;
    case SNKPT_AUX_TAA_PREV_VELOCITY: *src = P.taa_prev_vel[ctx->taa_cur]; *bytes = 4 * n; break;
    case SNKPT_AUX_BLAS_TABLE: *src = nullptr; *bytes = sizeof(SnkBlasInfo) * ctx->blas.size(); break;
    default: return SNKPT_E_INVALID;
  }
  return SNKPT_OK;
}

int snkpt_aux_size(SnkPt* ctx, uint32_t what, size_t* bytes) {
  if (!ctx || !bytes) return SNKPT_E_INVALID;
  const void* src;
  int rc = aux_desc(ctx, what, &src, bytes);
  if (rc) ctx->err = "snkpt_aux_size: unknown aux id";
  return rc;
}

int snkpt_read_aux(SnkPt* ctx, uint32_t what, void* dst, size_t bytes) {
  if (!ctx) return SNKPT_E_INVALID;
  const void* src = nullptr;
  size_t need = 0;
  REQUIRE(aux_desc(ctx, what, &src, &need) == SNKPT_OK, SNKPT_E_INVALID, "snkpt_read_aux: unknown aux id");
  REQUIRE(dst && bytes >= need, SNKPT_E_INVALID, "snkpt_read_aux: destination too small");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(work_stream(ctx)));
  if (what == SNKPT_AUX_BLAS_TABLE) { memcpy(dst, ctx->blas.data(), need); return SNKPT_OK; }
  if (what == SNKPT_AUX_RESERVOIRS_FINAL || what == SNKPT_AUX_RESERVOIRS_TEMPORAL) {
    // SoA planes → the reference's 48-B AoS record
    const uint32_t half = what == SNKPT_AUX_RESERVOIRS_FINAL ? 0 : 1;
    const size_t n = (size_t)ctx->W * ctx->H;
    std::vector<float4> pl(3 * n);
    for (int k = 0; k < 3; ++k) CK(cudaMemcpy(pl.data() + k * n, ctx->planes.res[half][k], 16 * n, cudaMemcpyDeviceToHost));
    SnkPixelReservoir* o = (SnkPixelReservoir*)dst;
    for (size_t i = 0; i < n; ++i) {
      memcpy(&o[i].w_sum_pl, &pl[i], 16);
      memcpy(&o[i].object_instance_idx, &pl[n + i], 16);
      memcpy(&o[i].w_sum_em, &pl[2 * n + i], 16);
    }
    return SNKPT_OK;
  }
  if (need) CK(cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost));
  return SNKPT_OK;
}

int snkpt_write_reservoirs(SnkPt* ctx, uint32_t half, const SnkPixelReservoir* src, uint32_t first_pixel, uint32_t count) {
  if (!ctx) return SNKPT_E_INVALID;
  const size_t n = (size_t)ctx->W * ctx->H;
  REQUIRE(half < 2 && src && (size_t)first_pixel + count <= n, SNKPT_E_INVALID, "snkpt_write_reservoirs: range");
  CK(cudaSetDevice(ctx->device));
  std::vector<float4> pl(3 * (size_t)count);
  for (size_t i = 0; i < count; ++i) {
    memcpy(&pl[i], &src[i].w_sum_pl, 16);
    memcpy(&pl[count + i], &src[i].object_instance_idx, 16);
    memcpy(&pl[2 * (size_t)count + i], &src[i].w_sum_em, 16);
  }
  CK(quiesce_frames(ctx));
  for (int k = 0; k < 3; ++k)
    CK(cudaMemcpy(ctx->planes.res[half][k] + first_pixel, pl.data() + (size_t)k * count, 16 * (size_t)count, cudaMemcpyHostToDevice));
  return SNKPT_OK;
}

int snkpt_read_output(SnkPt* ctx, void* dst, uint32_t format) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(dst && format <= SNKPT_FMT_RGBA32F, SNKPT_E_INVALID, "snkpt_read_output: arguments");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = work_stream(ctx);
  const size_t off = (size_t)ctx->row0 * ctx->W, cnt = (size_t)(ctx->row1 - ctx->row0) * ctx->W;
  if (format == SNKPT_FMT_RGBA16F) CK(cudaMemcpyAsync(dst, ctx->planes.out16 + off, 8 * cnt, cudaMemcpyDeviceToHost, st));
  else CK(cudaMemcpyAsync(dst, ctx->planes.radiance + off, 16 * cnt, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SNKPT_OK;
}

int snkpt_read_output_async(SnkPt* ctx, void* dst, uint32_t format, uint32_t slot) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(dst && format <= SNKPT_FMT_RGBA32F && slot < 4, SNKPT_E_INVALID, "snkpt_read_output_async: arguments");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 4; ++k) CK(cudaEventCreateWithFlags(&ctx->ev_read_done[k], cudaEventDisableTiming));
    if (!ctx->ev_frame_done) CK(cudaEventCreateWithFlags(&ctx->ev_frame_done, cudaEventDisableTiming));
  }
  // the copy starts when everything enqueued on the frame's stream so far is done and runs beside the next frame; kernels that
  // overwrite the image wait for it (snk::wait_readback)
  CK(cudaEventRecord(ctx->ev_frame_done, work_stream(ctx)));
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_frame_done, 0));
  const size_t off = (size_t)ctx->row0 * ctx->W, cnt = (size_t)(ctx->row1 - ctx->row0) * ctx->W;
  // in 1 MB pieces: one 16 MB transfer holds its copy engine for 0.3 ms, and the small copies / memsets with which the NEXT frame
  // starts (TLAS staging, per-frame uploads) queued up behind it — measured +0.07 ms per frame for a single transfer beside the
  // frame, +0.01 ms in pieces (tools/e2e_bisect.py, profiles/r05s_e2e_bisect.txt)
  const char* src = format == SNKPT_FMT_RGBA16F ? reinterpret_cast<const char*>(ctx->planes.out16 + off) : reinterpret_cast<const char*>(ctx->planes.radiance + off);
  const size_t total = (format == SNKPT_FMT_RGBA16F ? 8 : 16) * cnt, piece = 1u << 20;
  for (size_t o = 0; o < total; o += piece)
    CK(cudaMemcpyAsync(static_cast<char*>(dst) + o, src + o, total - o < piece ? total - o : piece, cudaMemcpyDeviceToHost, ctx->copy_stream));
  CK(cudaEventRecord(ctx->ev_read_done[slot], ctx->copy_stream));
  ctx->read_pending[slot] = true;
  return SNKPT_OK;
}

int snkpt_read_output_wait(SnkPt* ctx, uint32_t slot) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(slot < 4, SNKPT_E_INVALID, "snkpt_read_output_wait: slot");
  if (!ctx->read_pending[slot]) return SNKPT_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->ev_read_done[slot]));
  ctx->read_pending[slot] = false;
  return SNKPT_OK;
}

int snkpt_device_ptr(SnkPt* ctx, uint32_t what, void** ptr, size_t* bytes) {
  if (!ctx || !ptr || !bytes) return SNKPT_E_INVALID;
  const size_t n = (size_t)ctx->W * ctx->H;
  if (what == 0x100) { *ptr = ctx->planes.out16; *bytes = 8 * n; return SNKPT_OK; }
  if (what == 0x101) { *ptr = ctx->planes.accum; *bytes = 16 * n; return SNKPT_OK; }
  if (what == 0x102) { *ptr = ctx->planes.radiance; *bytes = 16 * n; return SNKPT_OK; }
  if (what >= 0x110 && what < 0x116) { *ptr = ctx->planes.res[(what - 0x110) / 3][(what - 0x110) % 3]; *bytes = 16 * n; return SNKPT_OK; }
  ctx->err = "snkpt_device_ptr: unknown id";
  return SNKPT_E_INVALID;
}

int snkpt_get_stats(SnkPt* ctx, SnkPtStats* out) {
  if (!ctx || !out) return SNKPT_E_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->gbuffer_valid && cudaEventQuery(ctx->ev[1]) == cudaSuccess) cudaEventElapsedTime(&ctx->stats.ms_primary, ctx->ev[0], ctx->ev[1]);
  if (ctx->build_timing_pending && cudaEventSynchronize(ctx->ev[7]) == cudaSuccess) {
    cudaEventElapsedTime(&ctx->stats.ms_build_last, ctx->ev[6], ctx->ev[7]);
    ctx->build_timing_pending = false;
  }
  uint32_t c[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaDeviceSynchronize() == cudaSuccess && cudaMemcpy(c, ctx->dev_counters.p, sizeof(c), cudaMemcpyDeviceToHost) == cudaSuccess) {
    ctx->stats.stack_overflows = c[0];
    ctx->stats.total_rays_shadow = ((uint64_t)c[9] << 32) | c[8];
    ctx->stats.total_rays_secondary = ((uint64_t)c[11] << 32) | c[10];
    if (ctx->render_pending) {
      ctx->stats.rays_shadow = ((uint64_t)c[3] << 32) | c[2];
      ctx->stats.rays_secondary = ((uint64_t)c[5] << 32) | c[4];
      cudaEventElapsedTime(&ctx->stats.ms_restir1, ctx->ev[2], ctx->ev[3]);
      cudaEventElapsedTime(&ctx->stats.ms_restir2, ctx->ev[3], ctx->ev[4]);
      cudaEventElapsedTime(&ctx->stats.ms_bounce, ctx->ev[4], ctx->ev[5]);
      cudaEventElapsedTime(&ctx->stats.ms_total, ctx->ev[2], ctx->ev[5]);
      ctx->render_pending = false;
    }
  }
  *out = ctx->stats;
  return SNKPT_OK;
}

int snkpt_set_profiling(SnkPt* ctx, uint32_t level) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(level <= 4, SNKPT_E_INVALID, "snkpt_set_profiling: level");
  ctx->profiling = level;
  ctx->prof_n = 0;
  return SNKPT_OK;
}

int snkpt_get_launch_timings(SnkPt* ctx, SnkLaunchTiming* out, uint32_t cap, uint32_t* count) {
  if (!ctx || !count) return SNKPT_E_INVALID;
  REQUIRE(out || cap == 0, SNKPT_E_INVALID, "snkpt_get_launch_timings: null output");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaDeviceSynchronize());  // launches may have been enqueued on a caller stream
  std::vector<unsigned long long> work(8 * PROF_MAX, 0ull);
  if (ctx->prof_work.p) CK(cudaMemcpy(work.data(), ctx->prof_work.p, 64 * PROF_MAX, cudaMemcpyDeviceToHost));
  uint32_t n = 0;
  for (int i = 0; i < ctx->prof_n && n < cap; ++i, ++n) {
    const ProfRec& r = ctx->prof[i];
    SnkLaunchTiming& o = out[n];
    memset(&o, 0, sizeof(o));
    memcpy(o.name, r.name, sizeof(o.name));
    cudaEventElapsedTime(&o.ms, r.e0, r.e1);
    if (ctx->profiling == 4) {  // timeline: start of the launch after the first record's start, in ns
      float off = 0.f;
      cudaEventElapsedTime(&off, ctx->prof[0].e0, r.e0);
      o.nodes = (uint64_t)((double)(off > 0.f ? off : 0.f) * 1e6);
    }
    o.is_trace = r.is_trace ? 1u : 0u;
    o.units = r.units;
    if (r.counted) {
      o.live = work[8 * i]; o.nodes = work[8 * i + 1]; o.tris = work[8 * i + 2]; o.insts = work[8 * i + 3];
      // ray in 32 B, hit out 16 B (closest) / 4 B (any), 80 B per node, 48 B per triangle record, 64 B per instance;
      // a queue launch (trace_q.cu) mixes both kinds and also reads the 4-B reference and writes the 4-B instance plane
      const uint64_t n_any = r.queue ? work[8 * i + 4] : (r.any ? o.live : 0ull);
      o.alg_bytes = o.live * (32ull + (r.queue ? 4ull : 0ull)) + n_any * 4ull + (o.live - n_any) * (r.queue ? 20ull : 16ull) + 80ull * o.nodes +
                    48ull * o.tris + 64ull * o.insts;
      if (r.queue) o.units = o.live;
    }
  }
  *count = n;
  ctx->prof_n = 0;
  return SNKPT_OK;
}

int snkpt_synchronize(SnkPt* ctx) {
  if (!ctx) return SNKPT_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(work_stream(ctx)));
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
  CK(cudaGetLastError());
  return SNKPT_OK;
}

// ---- particle collision pass -------------------------------------------------------------------------------------
void snkpt_particles_default_params(SnkPhysicsParams* p) {
  if (!p) return;
  p->particle_restitution = 0.2f; p->timestep = 1.f / 120.f; p->friction_coefficient = 0.5f; p->repulsion_factor = 0.08f;
  p->penetration_slop = 0.005f; p->restitution_slop = 0.05f; p->sleep_threshold = 0.05f;
}

int snkpt_particles_create(SnkPt* ctx, uint32_t count, int init) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(count > 0 && count <= (1u << 24), SNKPT_E_INVALID, "snkpt_particles_create: 1 .. 2^24 particles");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = work_stream(ctx);
  CK(cudaStreamSynchronize(st));
  for (int k = 0; k < 2; ++k) CK(ctx->ptcl[k].reserve(sizeof(SnkParticle) * (size_t)count, false, st));
  CK(ctx->ptcl_start.reserve(4ull * count, false, st));
  CK(ctx->ptcl_rays.reserve(8, false, st));
  CK(cudaMemsetAsync(ctx->ptcl_rays.p, 0, 8, st));
  CK(scratch_reserve(ctx->scratch_ptcl, count));
  ctx->num_particles = count;
  ctx->ptcl_cur = 0;
  CK(cudaMemsetAsync(ctx->ptcl[0].p, 0, sizeof(SnkParticle) * (size_t)count, st));
  if (init) ctx->stats.kernel_launches += launch_particles_init(ctx->ptcl[0].as<SnkParticle>(), count, st);
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  return SNKPT_OK;
}

int snkpt_particles_write(SnkPt* ctx, const SnkParticle* src, uint32_t first, uint32_t count) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(src && (uint64_t)first + count <= ctx->num_particles, SNKPT_E_INVALID, "snkpt_particles_write: range");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = work_stream(ctx);
  CK(cudaMemcpyAsync(ctx->ptcl[ctx->ptcl_cur].as<SnkParticle>() + first, src, sizeof(SnkParticle) * (size_t)count, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  return SNKPT_OK;
}

int snkpt_particles_read(SnkPt* ctx, SnkParticle* dst, uint32_t first, uint32_t count) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(dst && (uint64_t)first + count <= ctx->num_particles, SNKPT_E_INVALID, "snkpt_particles_read: range");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = work_stream(ctx);
  CK(cudaMemcpyAsync(dst, ctx->ptcl[ctx->ptcl_cur].as<SnkParticle>() + first, sizeof(SnkParticle) * (size_t)count, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SNKPT_OK;
}

int snkpt_particles_update(SnkPt* ctx, const SnkPhysicsParams* params, uint32_t substeps, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(ctx->num_particles > 0, SNKPT_E_STATE, "snkpt_particles_update: call snkpt_particles_create first");
  REQUIRE(ctx->tlas_built, SNKPT_E_STATE, "snkpt_particles_update: call snkpt_build_tlas first");
  REQUIRE(substeps >= 1 && substeps <= 64, SNKPT_E_INVALID, "snkpt_particles_update: substeps");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);
  SnkPhysicsParams prm;
  if (params) prm = *params; else snkpt_particles_default_params(&prm);
  REQUIRE(prm.timestep > 0.f, SNKPT_E_INVALID, "snkpt_particles_update: timestep");
  for (uint32_t s = 0; s < substeps; ++s) {
    const uint32_t cur = ctx->ptcl_cur;
    ctx->stats.kernel_launches += launch_particles_substep(ctx, ctx->ptcl[cur].as<SnkParticle>(), ctx->ptcl[cur ^ 1].as<SnkParticle>(), ctx->num_particles,
                                                           ctx->ptcl_start.as<uint32_t>(), prm, ctx->ptcl_rays.as<unsigned long long>(), st);
    ctx->ptcl_cur = cur ^ 1;  // the reference copies the result buffer back; here the two buffers swap roles
  }
  CK(cudaGetLastError());
  return SNKPT_OK;
}

int snkpt_particles_rays(SnkPt* ctx, uint64_t* rays_traced) {
  if (!ctx || !rays_traced) return SNKPT_E_INVALID;
  REQUIRE(ctx->num_particles > 0, SNKPT_E_STATE, "snkpt_particles_rays: no particles");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(work_stream(ctx)));
  unsigned long long v = 0;
  CK(cudaMemcpy(&v, ctx->ptcl_rays.p, 8, cudaMemcpyDeviceToHost));
  *rays_traced = v;
  return SNKPT_OK;
}

int snkpt_sort_uints_by_key(SnkPt* ctx, uint32_t* device_keys, uint32_t* device_values, uint32_t count, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE((device_keys && device_values) || count == 0, SNKPT_E_INVALID, "snkpt_sort_uints_by_key: null argument");
  REQUIRE(count < (1u << 30), SNKPT_E_INVALID, "snkpt_sort_uints_by_key: count must be below 2^30");
  if (count < 2) return SNKPT_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);
  CK(scratch_reserve(ctx->scratch_ptcl, count));
  ctx->stats.kernel_launches += launch_sort_uints(ctx->scratch_ptcl, device_keys, device_values, count, st);
  CK(cudaGetLastError());
  return SNKPT_OK;
}

// ---- ray queries ---------------------------------------------------------------------------------------------------
int snkpt_trace_rays(SnkPt* ctx, const SnkRay* rays, uint64_t count, uint32_t flags, SnkHit* hits, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(ctx->tlas_built, SNKPT_E_STATE, "snkpt_trace_rays: call snkpt_build_tlas first");
  REQUIRE((rays && hits) || count == 0, SNKPT_E_INVALID, "snkpt_trace_rays: null argument");
  if (count == 0) return SNKPT_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);  // after the TLAS build / uploads enqueued on the previous stream
  const bool any = flags & SNKPT_RAY_ANY;
  const float4* d_rays; SnkHit* d_hits;
  if (flags & SNKPT_RAY_DEVICE_PTRS) { d_rays = (const float4*)rays; d_hits = hits; }
  else {
    CK(ctx->q_rays.reserve(sizeof(SnkRay) * count, false, st));
    CK(ctx->q_hits.reserve(sizeof(SnkHit) * count, false, st));
    CK(cudaMemcpyAsync(ctx->q_rays.p, rays, sizeof(SnkRay) * count, cudaMemcpyHostToDevice, st));
    d_rays = ctx->q_rays.as<float4>(); d_hits = ctx->q_hits.as<SnkHit>();
  }
  prof_reset(ctx);
  ctx->stats.kernel_launches += traced_launch(ctx, any ? "trace_any.query" : "trace_closest.query", d_rays, count, any, d_hits, nullptr,
                                              nullptr, nullptr, st);
  CK(cudaGetLastError());
  if (any) ctx->stats.rays_shadow = count; else ctx->stats.rays_secondary = count;
  if (!(flags & SNKPT_RAY_DEVICE_PTRS)) {
    CK(cudaMemcpyAsync(hits, d_hits, sizeof(SnkHit) * count, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return SNKPT_OK;
}

int snkpt_count_work(SnkPt* ctx, const SnkRay* rays, const SnkHit* hits, uint64_t count, uint32_t flags, uint32_t* out3, void* stream) {
  if (!ctx) return SNKPT_E_INVALID;
  REQUIRE(ctx->tlas_built, SNKPT_E_STATE, "snkpt_count_work: call snkpt_build_tlas first");
  REQUIRE((rays && hits && out3) || count == 0, SNKPT_E_INVALID, "snkpt_count_work: null argument");
  if (count == 0) return SNKPT_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  note_stream(ctx, st);
  const float4* d_rays; const SnkHit* d_hits; uint32_t* d_out;
  if (flags & SNKPT_RAY_DEVICE_PTRS) { d_rays = (const float4*)rays; d_hits = hits; d_out = out3; }
  else {
    CK(ctx->q_rays.reserve(sizeof(SnkRay) * count, false, st));
    CK(ctx->q_hits.reserve(sizeof(SnkHit) * count, false, st));
    CK(ctx->q_work.reserve(12 * count, false, st));
    CK(cudaMemcpyAsync(ctx->q_rays.p, rays, sizeof(SnkRay) * count, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->q_hits.p, hits, sizeof(SnkHit) * count, cudaMemcpyHostToDevice, st));
    d_rays = ctx->q_rays.as<float4>(); d_hits = ctx->q_hits.as<SnkHit>(); d_out = ctx->q_work.as<uint32_t>();
  }
  launch_count_work(ctx->trace_scene(), d_rays, d_hits, count, flags & SNKPT_RAY_ANY, d_out, st);
  CK(cudaGetLastError());
  if (!(flags & SNKPT_RAY_DEVICE_PTRS)) {
    CK(cudaMemcpyAsync(out3, d_out, 12 * count, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return SNKPT_OK;
}

}  // extern "C"
