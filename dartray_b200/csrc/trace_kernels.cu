// Closest-hit / any-hit BVH traversal kernels for sm_100a.
//
// Arithmetic contract (bit-exact with the reference's Dart VM semantics): float32 storage, IEEE
// binary64 arithmetic without FMA contraction (this file is compiled with -fmad=false; every
// fused operation below is an explicit fma()/fmaf() call that is proven not to change a result).
//
//   slab test   /root/reference/lib/accelerators/bvh_accel.dart:439-472
//   traversal   bvh_accel.dart:101-165 (closest), :167-226 (any)
//   triangle    lib/shapes/triangle.dart:44-98 (intersect, f64), :162-194 (intersectP, f32-rounded vectors)
//   sphere      lib/shapes/sphere.dart:39-116, :169-241; Quadratic lib/core/common.dart:140-167
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>

#include "gpu_types.h"
#include "trace_device.cuh"
#include "trace_kernels.h"

namespace drt {

// One thread per ray.  ANY = false: closest hit (bvh_accel.dart:101-165); ANY = true: first hit
// found in the reference's visiting order ends the walk (bvh_accel.dart:167-226).
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(128) traceKernel(TraceScene sc, const float4* __restrict__ rayO,
                                                   const float4* __restrict__ rayD, uint64_t n, drt_hit_rec* hits,
                                                   uint8_t* occluded, DeviceCounters* counters, const double2* __restrict__ range,
                                                   const uint32_t* __restrict__ nDev, ExactExtras xx) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (nDev) n = *nDev;  // a wavefront queue of the renderer: the count lives in device memory
  unsigned long long nodesVisited = 0, primsTested = 0;
  bool found = false;
  HitState hit;
  hit.t = CUDART_INF; hit.b1 = 0.0; hit.b2 = 0.0; hit.prim = -1; hit.inst = -1;
  if (i < n && !sc.empty) {
    RayState r;
    initRay(r, rayO[i], rayD[i]);
    // the ray's time (scenes with TransformedPrimitives): per ray, or per wavefront slot with the slot id in the bits of rayO.w
    // (slot id 0xffffffff: a ray built without a time, i.e. at time 0 — the ambient-occlusion rays, ambient_occlusion_integrator.dart:45)
    double time = 0.0;
    if (xx.times) {
      const uint32_t ti = xx.timesBySlot ? __float_as_uint(rayO[i].w) : (uint32_t)i;
      if (!xx.timesBySlot) time = xx.times[i];
      else if (ti != 0xffffffffu) time = xx.times[ti];
    }
    if (range) { r.mint = range[i].x; r.maxt = range[i].y; }  // renderer rays: f64 minDistance / maxDistance
    int32_t stackRef[DRT_STACK];
    double stackT[DRT_STACK];
    int sp = 0;
    int32_t cur = 0;
    bool haveCur = false;
    {  // reference node 0: its own box is tested first (bvh_accel.dart:123-125)
      double tmin, tmax;
      if (COUNT) nodesVisited++;
      if (slabs(r, sc.rootMin[0], sc.rootMin[1], sc.rootMin[2], sc.rootMax[0], sc.rootMax[1], sc.rootMax[2], &tmin,
                &tmax) &&
          (tmin < r.maxt) && (tmax > r.mint)) {
        cur = sc.rootRef;
        haveCur = true;
      }
    }
    while (haveCur) {
      if (cur >= 0) {
        // interior node: decide both children from one 64-byte record
        const GNode* nd = sc.nodes + cur;
        float4 q0 = ldg4(&nd->c0min[0]), q1 = ldg4(&nd->c0max[1]), q2 = ldg4(&nd->c1min[2]);
        int4 q3 = __ldg(reinterpret_cast<const int4*>(&nd->ref0));
        int neg = q3.z == 0 ? r.negx : (q3.z == 1 ? r.negy : r.negz);
        double tmin0, tmax0, tmin1, tmax1;
        bool h0 = slabs(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tmin0, &tmax0) && (tmax0 > r.mint);
        bool h1 = slabs(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tmin1, &tmax1) && (tmax1 > r.mint);
        // near child first (bvh_accel.dart:147-153): dirIsNeg[axis] ? second : first
        int32_t nearRef = neg ? q3.y : q3.x, farRef = neg ? q3.x : q3.y;
        bool hn = neg ? h1 : h0, hf = neg ? h0 : h1;
        double tn = neg ? tmin1 : tmin0, tf = neg ? tmin0 : tmin1;
        if (COUNT) nodesVisited++;  // the near child's slab test happens now in the reference
        // The far child is pushed with its entry distance; the reference tests it when it is
        // popped, against the maxDistance of THAT moment -> the `tmin < maxDistance` half of the
        // test is re-evaluated at pop time (the other conditions do not change).
        if (hf && tf < r.maxt) {
          stackRef[sp] = farRef;
          stackT[sp] = tf;
          sp++;
        } else if (COUNT) {
          stackRef[sp] = farRef;  // counted when the reference would pop and reject it
          stackT[sp] = CUDART_INF;
          sp++;
        }
        if (hn && tn < r.maxt) {
          cur = nearRef;
          continue;
        }
      } else {
        // leaf
        uint32_t off = refLeafOffset(cur), cnt = refLeafCountField(cur);
        const GPrim* pr = sc.prims + off;
        if (cnt == 15u) cnt = (uint32_t)__ldg(&pr->leafCount);
        for (uint32_t k = 0; k < cnt; ++k) {
          float4 a = ldg4(&pr[k].p1[0]), b = ldg4(&pr[k].p2[0]), c = ldg4(&pr[k].p3[0]);
          int kind = __float_as_int(c.w);
          if (COUNT) primsTested++;
          if ((kind & 1) == 0) {
            if (ANY) {
              if (triangleAny(r, a, b, c)) { found = true; break; }
            } else {
              if (triangleClosest(r, a, b, c, &hit)) { found = true; hit.inst = -1; }
            }
          } else {
            const GSphere& s = sc.spheres[kind >> 1];
            double th, u, v;
            if (s.shape == 6) {  // a TransformedPrimitive: the object's own accelerator sees the transformed ray
              HitState ih;
              if (instanceTestCold(sc, s.instance, ANY, &r, time, &ih)) {
                found = true;
                if (ANY) break;
                hit.t = ih.t; hit.b1 = ih.b1; hit.b2 = ih.b2; hit.prim = ih.prim; hit.inst = s.instance;
                r.maxt = ih.t;
              }
            } else if (ANY) {
              if (sphereTest<true>(s, r, true, &th, nullptr, nullptr)) { found = true; break; }
            } else if (sphereTest<true>(s, r, false, &th, &u, &v)) {
              hit.t = th; hit.b1 = u; hit.b2 = v; hit.prim = __float_as_int(a.w); hit.inst = -1;
              r.maxt = th;
              found = true;
            }
          }
        }
        if (ANY && found) break;
      }
      // pop (bvh_accel.dart:139-143,156-159)
      haveCur = false;
      while (sp > 0) {
        --sp;
        if (COUNT) nodesVisited++;
        if (stackT[sp] < r.maxt) {
          cur = stackRef[sp];
          haveCur = true;
          break;
        }
      }
    }
  }
  if (i < n && (ANY ? (occluded != nullptr) : (hits != nullptr))) {  // null output: count only
    if (ANY) {
      occluded[i] = found ? 1 : 0;
    } else {
      drt_hit_rec o;
      o.t = found ? __double2float_rn(hit.t) : CUDART_INF_F;
      o.b1 = __double2float_rn(hit.b1);
      o.b2 = __double2float_rn(hit.b2);
      o.prim = hit.prim;
      reinterpret_cast<float4*>(hits)[i] = make_float4(o.t, o.b1, o.b2, __int_as_float(o.prim));
      if (xx.tOut) xx.tOut[i] = found ? hit.t : CUDART_INF;
      if (xx.instOut) xx.instOut[i] = found ? hit.inst : -1;
    }
  }
  if (COUNT) {
    unsigned long long nv = nodesVisited, pt = primsTested, hh = (i < n && found) ? 1ull : 0ull,
                       rr = (i < n) ? 1ull : 0ull;
    for (int o = 16; o > 0; o >>= 1) {
      nv += __shfl_down_sync(0xffffffffu, nv, o);
      pt += __shfl_down_sync(0xffffffffu, pt, o);
      hh += __shfl_down_sync(0xffffffffu, hh, o);
      rr += __shfl_down_sync(0xffffffffu, rr, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&counters->nodes_visited, nv);
      atomicAdd(&counters->prims_tested, pt);
      atomicAdd(&counters->hits, hh);
      atomicAdd(&counters->rays, rr);
    }
  }
}

cudaError_t launchTrace(const TraceScene& sc, bool any, bool count, const void* rayO, const void* rayD, uint64_t n,
                        void* out, DeviceCounters* counters, cudaStream_t stream, const double2* range, const uint32_t* nDev,
                        const ExactExtras* extras) {
  if (n == 0) return cudaSuccess;
  ExactExtras xx{};
  if (extras) xx = *extras;
  const int block = 128;
  uint64_t grid64 = (n + block - 1) / block;
  if (grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  dim3 grid((unsigned)grid64);
  const float4* o = static_cast<const float4*>(rayO);
  const float4* d = static_cast<const float4*>(rayD);
  if (any) {
    if (count) traceKernel<true, true><<<grid, block, 0, stream>>>(sc, o, d, n, nullptr, (uint8_t*)out, counters, range, nDev, xx);
    else traceKernel<true, false><<<grid, block, 0, stream>>>(sc, o, d, n, nullptr, (uint8_t*)out, counters, range, nDev, xx);
  } else {
    if (count) traceKernel<false, true><<<grid, block, 0, stream>>>(sc, o, d, n, (drt_hit_rec*)out, nullptr, counters, range, nDev, xx);
    else traceKernel<false, false><<<grid, block, 0, stream>>>(sc, o, d, n, (drt_hit_rec*)out, nullptr, counters, range, nDev, xx);
  }
  return cudaGetLastError();
}

}  // namespace drt
