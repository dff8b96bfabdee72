#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "gpu_types.h"

namespace drt {

struct drt_hit_rec {  // == drt_hit of include/drt.h
  float t, b1, b2;
  int32_t prim;
};

// Launches the closest-hit (any = false) or any-hit (any = true) traversal for n rays whose two
// float4 arrays live in device memory.  `out` is drt_hit_rec[n] or uint8_t[n].
// `range` / `nDev`: the renderer's per-ray f64 intervals and device-resident ray count (n is then the queue capacity);
// out == nullptr with count = true only counts.
// `extras`: what scenes with TransformedPrimitives add — the rays' times, the f64 tHit and the instance of each closest hit.
struct ExactExtras {
  const double* times = nullptr;  // ray times: indexed by the ray, or (timesBySlot) by the wavefront slot id carried in the bits of rayO.w
  int32_t timesBySlot = 0;
  double* tOut = nullptr;         // closest hit: tHit in f64 (+inf on a miss)
  int32_t* instOut = nullptr;     // closest hit: index of the TransformedPrimitive the hit came through, -1 for a top-level primitive
};
cudaError_t launchTrace(const TraceScene& sc, bool any, bool count, const void* rayO, const void* rayD, uint64_t n,
                        void* out, DeviceCounters* counters, cudaStream_t stream, const double2* range = nullptr,
                        const uint32_t* nDev = nullptr, const ExactExtras* extras = nullptr);

// Optional inputs/outputs of the production kernel used by the wavefront renderer.
struct TraceExtras {
  const uint32_t* nDev = nullptr;  // ray count in device memory (overrides n)
  const double2* range = nullptr;  // per ray (minDistance, maxDistance) in f64 instead of the float4 .w lanes
  double* tOut = nullptr;          // closest hit: tHit in f64 (+inf on a miss)
  int32_t noUV = 0;                // closest hit: the caller reads only t and the primitive (the renderer rebuilds the hit geometry
                                   // from t): a quadric's (u, v) — an atan2 and an acos in binary64 — are not computed, b1 = b2 = 0
};

// Production path: persistent warps, while-while traversal, float32-filtered slab test with exact
// float64 fallback (trace_fast.cu).  `nextRay` is a device counter owned by the context.
cudaError_t launchTraceFast(const TraceScene& sc, bool any, const void* rayO, const void* rayD, uint64_t n, void* out,
                            unsigned long long* nextRay, int numSMs, cudaStream_t stream,
                            const TraceExtras* extras = nullptr);

// Second generation (trace_fast2.cu): quantised 64-byte wide nodes, postponed leaves, every leaf box decided exactly in
// the leaf phase.  launchTraceFast forwards here whenever the scene carries quantised nodes (TraceScene::wideQ).
cudaError_t launchTraceQ(const TraceScene& sc, bool any, const void* rayO, const void* rayD, uint64_t n, void* out,
                         unsigned long long* nextRay, int numSMs, cudaStream_t stream, const TraceExtras* extras = nullptr);

// Float32 builds of the production kernels (trace_fast_f32.cu, trace_q_f32.cu: trace_fast.cu / trace_fast2.cu lowered by gen_f32.py): what
// the path integrator's ray queues run under DRT_PRECISION_F32.  Box filter, leaf box, triangle and quadric tests in float32; the
// visiting order is the reference's, the decisions are float32 ones (a hit within rounding of an edge or of the interval's end may
// differ), which is the tolerance that mode states.  Same kernel choice by scene as launchTraceFast.
cudaError_t launchTraceFastF32(const TraceScene& sc, bool any, const void* rayO, const void* rayD, uint64_t n, void* out,
                               unsigned long long* nextRay, int numSMs, cudaStream_t stream, const TraceExtras* extras);

}  // namespace drt
