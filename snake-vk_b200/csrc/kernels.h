// kernels.h — host-side launchers of the CUDA kernels (one .cu per stage).
#pragma once
#include "context.h"

namespace snk {

// trace.cu — wavefront traversal over a ray queue
// rays: 2 x float4 per ray (o,tmin | d,tmax). hits_aos: SnkHit records (20 B) or nullptr;
// hits_soa / inst_soa: float4 (t,u,v,prim) + u32 planes or nullptr.
// vis_out: 1.0 = unoccluded / 0.0 = occluded per ray (any-hit queues) or nullptr.
// ticket: device u64 the persistent kernel draws ray indices from (zeroed by the launcher)
int launch_trace(const TraceScene& S, const float4* rays, uint64_t n, bool any, SnkHit* hits_aos, float4* hits_soa,
                 uint32_t* inst_soa, float* vis_out, unsigned long long* ticket, cudaStream_t st);
size_t trace_spill_bytes();
int launch_count_work(const TraceScene& S, const float4* rays, const SnkHit* hits, uint64_t n, bool any, uint32_t* out3,
                      cudaStream_t st);
// sums {live rays, nodes, leaf records, instances} of a whole queue into sums4 (hits as float4 with t in .x)
int launch_count_work_sum(const TraceScene& S, const float4* rays, const float4* hits_soa, uint64_t n, unsigned long long* sums4,
                          cudaStream_t st);

// trace_q.cu — queue traversal: drains a compact queue of ray references (closest-hit and any-hit rays of several
// segments mixed); prof_t (optional) receives the reported hit distance per queue position
int launch_trace_q(const TraceScene& S, const TraceSegs& segs, const RefQueue& q, float* prof_t, cudaStream_t st);
// trace_mx.cu — the same queue drained with several rays multiplexed per lane (state in shared memory, one voted phase per iteration)
int launch_trace_mx(const TraceScene& S, const TraceSegs& segs, const RefQueue& q, float* prof_t, int sms, cudaStream_t st);
size_t trace_mx_spill_bytes(int sms);
int device_sm_count();
int launch_count_work_q(const TraceScene& S, const TraceSegs& segs, const RefQueue& q, const float* prof_t, unsigned long long* sums8,
                        cudaStream_t st);
// launch_trace_q bracketed by a profiling record (+ counted replay at level 2)
int traced_launch_q(SnkPt* ctx, const char* name, const TraceSegs& segs, const RefQueue& q, cudaStream_t st, int spill_region = 0);
bool ns_overlap_enabled(const SnkPt* ctx);
// queue k (0..23) of the frame: refs in ctx->wave_refs (shared: a queue is drained before the next one is filled), counters in
// dev_counters; all are reset by reset_ref_queues
RefQueue ref_queue(SnkPt* ctx, int k, uint32_t cap, uint32_t first_ref = 0);
int reserve_ref_queues(SnkPt* ctx, uint32_t cap, cudaStream_t st);
void reset_ref_queues(SnkPt* ctx, cudaStream_t st);

// group.cu — row-band groups: exchange of reservoir rows between band neighbours (kind 0: temporal half after the temporal
// stage, kind 1: final half after the spatial stage); no-ops without a group
int group_exchange(SnkPt* ctx, int kind, cudaStream_t st);
uint32_t group_gbuffer_halo(const SnkPt* ctx);
void group_free(SnkPt* ctx);

// particles.cu — Particle.comp Init; one sub-step of ParticleSystem::UpdateParticles (keys, sort, start indices, update): P -> out
int launch_particles_init(SnkParticle* P, uint32_t n, cudaStream_t st);
int launch_particles_substep(SnkPt* ctx, const SnkParticle* P, SnkParticle* out, uint32_t n, uint32_t* cell_start, const SnkPhysicsParams& prm,
                             unsigned long long* ray_counter, cudaStream_t st);

// api.cu — a kernel about to overwrite the output image first waits for pending asynchronous read-backs of it
void wait_readback(SnkPt* ctx, cudaStream_t st);

// api.cu — per-launch profiling (no-ops at level 0)
void prof_reset(SnkPt* ctx);
int prof_begin(SnkPt* ctx, const char* name, uint64_t units, bool is_trace, bool any, cudaStream_t st);
void prof_end(SnkPt* ctx, int rec, cudaStream_t st);
// launch_trace bracketed by a profiling record; at level 2 also counts the queue's algorithmic work
int traced_launch(SnkPt* ctx, const char* name, const float4* rays, uint64_t n, bool any, SnkHit* hits_aos, float4* hits_soa,
                  uint32_t* inst_soa, float* vis_out, cudaStream_t st);

// gbuffer.cu — primary visibility (replaces GBufferPass::RecordCommandBuffer, GBufferPass.cpp:70-196)
// tickets2: two device words used in turn by the packet kernel's tile fetch (dev_counters u32 [200, 201])
int launch_primary(const TraceScene& S, const SceneDev& D, const FrameConsts& fc, const PixelPlanes& P, uint32_t* tickets2, uint32_t parity, cudaStream_t st);
// gbuffer_nm.cu — normal maps applied to the G-buffer normal plane (second pass, only when textures are set)
int launch_gbuffer_normal_map(const SceneDev& D, const FrameConsts& fc, const PixelPlanes& P, cudaStream_t st);

// restir.cu / bounce.cu — RT::RecordRenderCmdBuf (Raytracing.cpp:227-267) in reference mode
int launch_restir_passes(SnkPt* ctx, const SnkRtSettings& s, cudaStream_t st);
int launch_reference_frame(SnkPt* ctx, const SnkRtSettings& s, cudaStream_t st);
int launch_northstar_frame(SnkPt* ctx, const SnkRtSettings& s, cudaStream_t st);
// restir_ns.cu — north-star ReSTIR DI stages (initial / temporal / spatial + final evaluation / resolve) → planes.direct
int restir_ns_prepare(SnkPt* ctx, TraceSegs& segs, cudaStream_t st);
int build_emissive_table(SnkPt* ctx, cudaStream_t st);
int restir_ns_initial(SnkPt* ctx, const SnkRtSettings& s, const RefQueue& q, cudaStream_t st);
int restir_ns_reuse(SnkPt* ctx, const SnkRtSettings& s, const RefQueue& q, cudaStream_t st, int spill_region);
int restir_ns_resolve(SnkPt* ctx, cudaStream_t st);
int launch_accumulate(const PixelPlanes& P, uint32_t npix, uint32_t count, cudaStream_t st);
// taa.cu — TAA_Resolve.comp: resolved frame -> out (next history), this frame's velocity -> vel_out
int launch_taa_resolve(const ushort4* cur, const ushort2* velocity, const ushort4* history, const ushort2* prev_vel, ushort4* out,
                       ushort2* vel_out, uint32_t W, uint32_t H, cudaStream_t st);

}  // namespace snk
