// Production traversal kernels: persistent warps, while-while traversal, float32-FILTERED slab
// tests with an exact float64 fallback.
//
// Every decision the reference makes (bvh_accel.dart:439-472 slab test, :139-159 pop order,
// triangle.dart:44-98 / :162-194, sphere.dart) is still made with the reference's arithmetic; the
// float32 filter only answers when its answer provably equals the float64 one:
//
//   reference   t = ((double)b - (double)o) * (double)invDir        (b, o, invDir are float32)
//   filter      t' = (b - o) * invDir in float32  ->  |t - t'| <= 2^-23 |t'| (+ underflow)
//
// With eps = 2^-22 and an absolute floor `tiny`, hi(x) = x + eps|x| + tiny and lo(x) = x - eps|x| - tiny
// are monotone, so hi(max_i t'_i) >= max_i t_i etc.  A box is accepted/rejected by the filter only if
// all three reference conditions (max near <= min far, tmin < maxDistance, tmax > minDistance) are
// decided with that margin; anything else — including NaN/inf from zero direction components and
// float32 overflow — takes the exact path (slabs() of trace_device.cuh).
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>

#include "gpu_types.h"
#include "trace_device.cuh"
#include "trace_kernels.h"

namespace drt {

#define FULL_MASK 0xffffffffu
#define RAY_CHUNK 256  // rays a warp reserves per atomicAdd on the global ray counter

struct FastRay {
  float ox, oy, oz;  // ray.origin
  float dx, dy, dz;  // ray.direction
  float ix, iy, iz;  // invDir: (float)(1.0 / (double)d), bvh_accel.dart:109-111
  double mint, maxt;
  float mintLo, mintHi, maxtLo, maxtHi;  // float32 brackets of the f64 interval ends
  bool slow;                             // non-finite origin / invDir: exact path for every box
};

struct StackEntry {
  int32_t ref;
  float tmin;  // float32 image of the box entry distance (exact value re-derived when it matters)
};

static __device__ __forceinline__ void widen(const FastRay& f, RayState& r) {
  r.ox = f.ox; r.oy = f.oy; r.oz = f.oz;
  r.dx = f.dx; r.dy = f.dy; r.dz = f.dz;
  r.ix = f.ix; r.iy = f.iy; r.iz = f.iz;
  r.mint = f.mint; r.maxt = f.maxt;
  r.negx = f.ix < 0.f; r.negy = f.iy < 0.f; r.negz = f.iz < 0.f;
}

static __device__ __forceinline__ void setMaxt(FastRay& f, double t) {
  f.maxt = t;
  f.maxtLo = __double2float_rd(t);
  f.maxtHi = __double2float_ru(t);
}

#define DRT_EPS 2.384185791015625e-07f  // 2^-22
#define DRT_TINY 1.0e-37f

// 1 = the reference accepts the box, 0 = it rejects it, 2 = not provable in float32.
static __device__ __forceinline__ int slabFilter(const FastRay& r, float lox, float loy, float loz, float hix,
                                                 float hiy, float hiz, float* tminOut) {
  float t0x = (lox - r.ox) * r.ix, t1x = (hix - r.ox) * r.ix;
  float t0y = (loy - r.oy) * r.iy, t1y = (hiy - r.oy) * r.iy;
  float t0z = (loz - r.oz) * r.iz, t1z = (hiz - r.oz) * r.iz;
  float tmin = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
  float tmax = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  float dmin = fmaf(DRT_EPS, fabsf(tmin), DRT_TINY), dmax = fmaf(DRT_EPS, fabsf(tmax), DRT_TINY);
  float tminHi = tmin + dmin, tminLo = tmin - dmin, tmaxHi = tmax + dmax, tmaxLo = tmax - dmax;
  bool pass = (tminHi <= tmaxLo) && (tminHi < r.maxtLo) && (tmaxLo > r.mintHi);
  bool fail = (tminLo > tmaxHi) || (tminLo >= r.maxtHi) || (tmaxHi <= r.mintLo);
  *tminOut = tmin;
  return pass ? 1 : (fail ? 0 : 2);
}

// Exact evaluation of one box (rare).  Returns the reference's decision; *tminOut = f64 tmin.
static __device__ __noinline__ bool slabExact(const FastRay& f, float lox, float loy, float loz, float hix, float hiy,
                                              float hiz, double* tminOut) {
  RayState r;
  widen(f, r);
  double tmin, tmax;
  if (!slabs(r, lox, loy, loz, hix, hiy, hiz, &tmin, &tmax)) return false;
  *tminOut = tmin;
  return (tmin < r.maxt) && (tmax > r.mint);
}

// Exact re-test of a popped LEAF whose stored entry distance is too close to maxDistance to call:
// rebuild the leaf box from its primitives (triangle.dart:39-42 / sphere world bound) and evaluate
// `tmin < ray.maxDistance` exactly as the reference does when it visits the leaf node.
static __device__ __noinline__ bool leafStillReachable(const TraceScene& sc, const FastRay& f, int32_t ref) {
  uint32_t off = refLeafOffset(ref), cnt = refLeafCountField(ref);
  const GPrim* pr = sc.prims + off;
  if (cnt == 15u) cnt = (uint32_t)__ldg(&pr->leafCount);
  float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (uint32_t k = 0; k < cnt; ++k) {
    float4 a = ldg4(&pr[k].p1[0]), b = ldg4(&pr[k].p2[0]), c = ldg4(&pr[k].p3[0]);
    int kind = __float_as_int(c.w);
    if ((kind & 1) == 0) {
      lo[0] = fminf(lo[0], fminf(a.x, fminf(b.x, c.x))); hi[0] = fmaxf(hi[0], fmaxf(a.x, fmaxf(b.x, c.x)));
      lo[1] = fminf(lo[1], fminf(a.y, fminf(b.y, c.y))); hi[1] = fmaxf(hi[1], fmaxf(a.y, fmaxf(b.y, c.y)));
      lo[2] = fminf(lo[2], fminf(a.z, fminf(b.z, c.z))); hi[2] = fmaxf(hi[2], fmaxf(a.z, fmaxf(b.z, c.z)));
    } else {
      const GSphere& s = sc.spheres[kind >> 1];
      for (int x = 0; x < 3; ++x) { lo[x] = fminf(lo[x], s.wmin[x]); hi[x] = fmaxf(hi[x], s.wmax[x]); }
    }
  }
  double tmin;
  return slabExact(f, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], &tmin);
}

// Pops until an entry survives the reference's pop-time test `tmin < ray.maxDistance`
// (bvh_accel.dart:139-143,156-159 + :471).  Interior entries inside the undecidable band are
// entered (their children are culled by the same comparison); leaf entries in the
// band are re-tested exactly.
static __device__ __forceinline__ bool popNext(const TraceScene& sc, const FastRay& r, const StackEntry* stack, int& sp,
                                               int32_t& cur) {
  while (sp > 0) {
    --sp;
    int32_t ref = stack[sp].ref;
    float t = stack[sp].tmin;
    float dt = fmaf(DRT_EPS, fabsf(t), DRT_TINY);
    bool take = (t + dt) < r.maxtLo;
    if (!take && !((t - dt) >= r.maxtHi)) take = ref >= 0 ? true : leafStillReachable(sc, r, ref);
    if (take) {
      cur = ref;
      return true;
    }
  }
  return false;
}

template <bool ANY>
__global__ void __launch_bounds__(128, 4) traceFastKernel(TraceScene sc, const float4* __restrict__ rayO,
                                                          const float4* __restrict__ rayD, uint64_t n,
                                                          float4* __restrict__ hits, uint8_t* __restrict__ occluded,
                                                          unsigned long long* __restrict__ nextRay) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned ltMask = (1u << lane) - 1u;
  unsigned long long warpNext = 0, warpEnd = 0;  // warp-uniform: the chunk of rays this warp owns
  bool exhausted = false;                        // warp-uniform: the global counter ran past n
  bool alive = false;
  unsigned long long rayIdx = 0;
  FastRay r;
  StackEntry stack[DRT_STACK];
  int sp = 0;
  int32_t cur = 0;
  float hb1 = 0.f, hb2 = 0.f;
  int hprim = -1;
  bool found = false;

  auto retire = [&]() {
    alive = false;
    if (ANY) occluded[rayIdx] = found ? 1 : 0;
    else hits[rayIdx] = make_float4(found ? __double2float_rn(r.maxt) : CUDART_INF_F, hb1, hb2, __int_as_float(hprim));
  };

  for (;;) {
    // ---- refill idle lanes from the warp's chunk ----------------------------------------------
    unsigned dead = __ballot_sync(FULL_MASK, !alive);
    if (dead) {
      if (warpNext == warpEnd && !exhausted) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(nextRay, (unsigned long long)RAY_CHUNK);
        base = __shfl_sync(FULL_MASK, base, 0);
        if (base >= n) {
          exhausted = true;
        } else {
          warpNext = base;
          warpEnd = (base + RAY_CHUNK) < n ? (base + RAY_CHUNK) : n;
        }
      }
      unsigned avail = (unsigned)(warpEnd - warpNext);
      unsigned nDead = __popc(dead);
      unsigned rank = __popc(dead & ltMask);
      if (!alive && rank < avail) {
        rayIdx = warpNext + rank;
        float4 o = __ldg(rayO + rayIdx), d = __ldg(rayD + rayIdx);
        r.ox = o.x; r.oy = o.y; r.oz = o.z;
        r.dx = d.x; r.dy = d.y; r.dz = d.z;
        r.ix = __double2float_rn(1.0 / (double)d.x);
        r.iy = __double2float_rn(1.0 / (double)d.y);
        r.iz = __double2float_rn(1.0 / (double)d.z);
        r.mint = o.w;
        r.mintLo = r.mintHi = o.w;
        setMaxt(r, (double)d.w);
        // any inf/NaN among origin / invDir components -> every box of this ray takes the exact path
        r.slow = !(fabsf(r.ox) <= 3.0e38f) || !(fabsf(r.oy) <= 3.0e38f) || !(fabsf(r.oz) <= 3.0e38f) ||
                 !(fabsf(r.ix) <= 3.0e38f) || !(fabsf(r.iy) <= 3.0e38f) || !(fabsf(r.iz) <= 3.0e38f);
        sp = 0;
        found = false;
        hprim = -1;
        hb1 = hb2 = 0.f;
        alive = true;
        // reference node 0: its own box is tested first (bvh_accel.dart:123-125)
        double t0 = 0.0;
        if (sc.empty || !slabExact(r, sc.rootMin[0], sc.rootMin[1], sc.rootMin[2], sc.rootMax[0], sc.rootMax[1],
                                   sc.rootMax[2], &t0))
          retire();
        else
          cur = sc.rootRef;
      }
      warpNext += nDead < avail ? nDead : avail;
      if (exhausted && __all_sync(FULL_MASK, !alive)) break;
    }

    // ---- interior phase: walk until this lane holds a leaf (or runs out of nodes) -------------
    while (alive && cur >= 0) {
      const GNode* nd = sc.nodes + cur;
      float4 q0 = ldg4(&nd->c0min[0]), q1 = ldg4(&nd->c0max[1]), q2 = ldg4(&nd->c1min[2]);
      int4 q3 = __ldg(reinterpret_cast<const int4*>(&nd->ref0));
      float tm0 = 0.f, tm1 = 0.f;
      int c0 = r.slow ? 2 : slabFilter(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tm0);
      int c1 = r.slow ? 2 : slabFilter(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tm1);
      if (c0 == 2) {
        double t = 0.0;
        c0 = slabExact(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t) ? 1 : 0;
        tm0 = __double2float_rn(t);
      }
      if (c1 == 2) {
        double t = 0.0;
        c1 = slabExact(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t) ? 1 : 0;
        tm1 = __double2float_rn(t);
      }
      // near child first: dirIsNeg[axis] ? second : first (bvh_accel.dart:147-153)
      float iax = q3.z == 0 ? r.ix : (q3.z == 1 ? r.iy : r.iz);
      bool neg = iax < 0.f;
      int32_t nearRef = neg ? q3.y : q3.x, farRef = neg ? q3.x : q3.y;
      int hn = neg ? c1 : c0, hf = neg ? c0 : c1;
      if (hf) {
        stack[sp].ref = farRef;
        stack[sp].tmin = neg ? tm0 : tm1;
        sp++;
      }
      if (hn) cur = nearRef;
      else if (!popNext(sc, r, stack, sp, cur)) retire();
    }

    // ---- leaf phase ------------------------------------------------------------------------------
    if (alive) {
      uint32_t off = refLeafOffset(cur), cnt = refLeafCountField(cur);
      const GPrim* pr = sc.prims + off;
      if (cnt == 15u) cnt = (uint32_t)__ldg(&pr->leafCount);
      RayState rs;
      widen(r, rs);
      bool stop = false;
      for (uint32_t k = 0; k < cnt && !stop; ++k) {
        float4 a = ldg4(&pr[k].p1[0]), b = ldg4(&pr[k].p2[0]), c = ldg4(&pr[k].p3[0]);
        int kind = __float_as_int(c.w);
        if ((kind & 1) == 0) {
          if (ANY) {
            if (triangleAny(rs, a, b, c)) { found = true; stop = true; }
          } else {
            HitState h;
            if (triangleClosest(rs, a, b, c, &h)) {
              found = true;
              hb1 = __double2float_rn(h.b1); hb2 = __double2float_rn(h.b2); hprim = h.prim;
            }
          }
        } else {
          const GSphere& s = sc.spheres[kind >> 1];
          double th, u, v;
          if (ANY) {
            if (sphereTest(s, rs, true, &th, nullptr, nullptr)) { found = true; stop = true; }
          } else if (sphereTest(s, rs, false, &th, &u, &v)) {
            found = true;
            hb1 = __double2float_rn(u); hb2 = __double2float_rn(v); hprim = __float_as_int(a.w);
            rs.maxt = th;
          }
        }
      }
      if (!ANY && rs.maxt != r.maxt) setMaxt(r, rs.maxt);
      if ((ANY && found) || !popNext(sc, r, stack, sp, cur)) retire();
    }
  }
}

cudaError_t launchTraceFast(const TraceScene& sc, bool any, const void* rayO, const void* rayD, uint64_t n, void* out,
                            unsigned long long* nextRay, int numSMs, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(nextRay, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  const int block = 128;
  static int perSm[2] = {0, 0};
  if (!perSm[any ? 1 : 0]) {
    int b = 0;
    e = any ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, traceFastKernel<true>, block, 0)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, traceFastKernel<false>, block, 0);
    if (e != cudaSuccess) return e;
    perSm[any ? 1 : 0] = b > 0 ? b : 1;
  }
  uint64_t want = (n + RAY_CHUNK - 1) / RAY_CHUNK;  // one warp per chunk is enough
  uint64_t blocksWanted = (want + 3) / 4;
  uint64_t persistent = (uint64_t)numSMs * perSm[any ? 1 : 0];  // one resident wave: persistent warps
  dim3 grid((unsigned)(blocksWanted < persistent ? blocksWanted : persistent));
  const float4* o = static_cast<const float4*>(rayO);
  const float4* d = static_cast<const float4*>(rayD);
  if (any) traceFastKernel<true><<<grid, block, 0, stream>>>(sc, o, d, n, nullptr, (uint8_t*)out, nextRay);
  else traceFastKernel<false><<<grid, block, 0, stream>>>(sc, o, d, n, (float4*)out, nullptr, nextRay);
  return cudaGetLastError();
}

}  // namespace drt
