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

#ifndef DRT_MIN_BLOCKS
#define DRT_MIN_BLOCKS 5
#endif
#ifndef DRT_LEAF_BATCH
#define DRT_LEAF_BATCH 16  // lanes holding an untested leaf before the warp runs the exact leaf phase
#endif

// Hot per-ray state, kept in registers (never address-taken: the exact helpers take it by value).
struct FastRay {
  float ox, oy, oz;  // ray.origin
  float ix, iy, iz;  // invDir: (float)(1.0 / (double)d), bvh_accel.dart:109-111
  float mintLo, mintHi, maxtLo, maxtHi;  // float32 brackets of the f64 interval ends
  double mint, maxt;
  unsigned negMask;  // bit a = invDir[a] < 0 (dirIsNeg, bvh_accel.dart:113-115); bit 3 = "slow" ray
};

struct StackEntry {
  int32_t ref;
  float tmin;  // float32 image of the box entry distance (exact value re-derived when it matters)
};

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

// Exact evaluation of one box (rare), everything passed BY VALUE so the caller's ray state stays in
// registers.  Returns the reference's decision (bvh_accel.dart:439-472); *tminOut = float32(tmin).
static __device__ __noinline__ bool slabExact(float ox, float oy, float oz, float ix, float iy, float iz, double mint,
                                              double maxt, float lox, float loy, float loz, float hix, float hiy,
                                              float hiz, float* tminOut) {
  RayState r;
  r.ox = ox; r.oy = oy; r.oz = oz;
  r.ix = ix; r.iy = iy; r.iz = iz;
  r.negx = ix < 0.f; r.negy = iy < 0.f; r.negz = iz < 0.f;
  double tmin, tmax;
  if (!slabs(r, lox, loy, loz, hix, hiy, hiz, &tmin, &tmax)) return false;
  *tminOut = __double2float_rn(tmin);
  return (tmin < maxt) && (tmax > mint);
}
#define SLAB_EXACT(r, lox, loy, loz, hix, hiy, hiz, tout) \
  slabExact((r).ox, (r).oy, (r).oz, (r).ix, (r).iy, (r).iz, (r).mint, (r).maxt, lox, loy, loz, hix, hiy, hiz, tout)

// Exact re-test of a popped LEAF whose stored entry distance is too close to maxDistance to call:
// rebuild the leaf box from its primitives (triangle.dart:39-42 / sphere world bound) and evaluate
// the slab test exactly as the reference does when it visits the leaf node.
static __device__ __noinline__ bool leafStillReachable(const GPrim* prims, const GSphere* spheres, float ox, float oy,
                                                       float oz, float ix, float iy, float iz, double mint, double maxt,
                                                       int32_t ref) {
  uint32_t off = refLeafOffset(ref), cnt = refLeafCountField(ref);
  const GPrim* pr = prims + off;
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
      const GSphere& s = spheres[kind >> 1];
      for (int x = 0; x < 3; ++x) { lo[x] = fminf(lo[x], s.wmin[x]); hi[x] = fmaxf(hi[x], s.wmax[x]); }
    }
  }
  float t;
  return slabExact(ox, oy, oz, ix, iy, iz, mint, maxt, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], &t);
}

// Pops until an entry survives the reference's pop-time test `tmin < ray.maxDistance`
// (bvh_accel.dart:139-143,156-159 + :471).  Interior entries inside the undecidable band are
// entered (their children are culled by the same comparison); leaf entries in the
// band are re-tested exactly.
#define POP_NEXT(ok)                                                                                       \
  do {                                                                                                     \
    ok = false;                                                                                            \
    while (sp > 0) {                                                                                       \
      --sp;                                                                                                \
      int32_t ref_ = stack[sp].ref;                                                                        \
      float t_ = stack[sp].tmin;                                                                           \
      float dt_ = fmaf(DRT_EPS, fabsf(t_), DRT_TINY);                                                      \
      bool take_ = (t_ + dt_) < r.maxtLo;                                                                  \
      if (!take_ && !((t_ - dt_) >= r.maxtHi))                                                             \
        take_ = ref_ >= 0 ? true                                                                           \
                          : leafStillReachable(sc.prims, sc.spheres, r.ox, r.oy, r.oz, r.ix, r.iy, r.iz,   \
                                               r.mint, r.maxt, ref_);                                      \
      if (take_) { cur = ref_; ok = true; break; }                                                         \
    }                                                                                                      \
  } while (0)

template <bool ANY>
__global__ void __launch_bounds__(128, DRT_MIN_BLOCKS) traceFastKernel(TraceScene sc, const float4* __restrict__ rayO,
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

#define RETIRE()                                                                                             \
  do {                                                                                                       \
    alive = false;                                                                                           \
    if (ANY) occluded[rayIdx] = found ? 1 : 0;                                                               \
    else hits[rayIdx] = make_float4(found ? __double2float_rn(r.maxt) : CUDART_INF_F, hb1, hb2,              \
                                    __int_as_float(hprim));                                                  \
  } while (0)

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
        r.ix = __double2float_rn(1.0 / (double)d.x);
        r.iy = __double2float_rn(1.0 / (double)d.y);
        r.iz = __double2float_rn(1.0 / (double)d.z);
        r.mint = o.w;
        r.mintLo = r.mintHi = o.w;
        r.maxt = d.w;
        r.maxtLo = r.maxtHi = d.w;
        // any inf/NaN among origin / invDir components -> every box of this ray takes the exact path
        bool slow = !(fabsf(r.ox) <= 3.0e38f) || !(fabsf(r.oy) <= 3.0e38f) || !(fabsf(r.oz) <= 3.0e38f) ||
                    !(fabsf(r.ix) <= 3.0e38f) || !(fabsf(r.iy) <= 3.0e38f) || !(fabsf(r.iz) <= 3.0e38f);
        r.negMask = (r.ix < 0.f ? 1u : 0u) | (r.iy < 0.f ? 2u : 0u) | (r.iz < 0.f ? 4u : 0u) | (slow ? 8u : 0u);
        sp = 0;
        found = false;
        hprim = -1;
        hb1 = hb2 = 0.f;
        alive = true;
        // reference node 0: its own box is tested first (bvh_accel.dart:123-125)
        float t0;
        if (sc.empty || !SLAB_EXACT(r, sc.rootMin[0], sc.rootMin[1], sc.rootMin[2], sc.rootMax[0], sc.rootMax[1],
                                    sc.rootMax[2], &t0))
          RETIRE();
        else
          cur = sc.rootRef;
      }
      warpNext += nDead < avail ? nDead : avail;
      if (exhausted && __all_sync(FULL_MASK, !alive)) break;
    }

    // ---- one interior step for every lane that holds an interior node ---------------------------
    if (alive && cur >= 0) {
      const GNode* nd = sc.nodes + cur;
      float4 q0 = ldg4(&nd->c0min[0]), q1 = ldg4(&nd->c0max[1]), q2 = ldg4(&nd->c1min[2]);
      int4 q3 = __ldg(reinterpret_cast<const int4*>(&nd->ref0));
      float tm0 = 0.f, tm1 = 0.f;
      const bool slow = (r.negMask & 8u) != 0;
      int c0 = slow ? 2 : slabFilter(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tm0);
      int c1 = slow ? 2 : slabFilter(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tm1);
      if (c0 == 2) c0 = SLAB_EXACT(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tm0) ? 1 : 0;
      if (c1 == 2) c1 = SLAB_EXACT(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tm1) ? 1 : 0;
      // near child first: dirIsNeg[axis] ? second : first (bvh_accel.dart:147-153)
      const bool neg = ((r.negMask >> q3.z) & 1u) != 0;
      int32_t nearRef = neg ? q3.y : q3.x, farRef = neg ? q3.x : q3.y;
      int hn = neg ? c1 : c0, hf = neg ? c0 : c1;
      if (hf) {
        stack[sp].ref = farRef;
        stack[sp].tmin = neg ? tm0 : tm1;
        sp++;
      }
      if (hn) {
        cur = nearRef;
      } else {
        bool ok;
        POP_NEXT(ok);
        if (!ok) RETIRE();
      }
    }

    // ---- exact leaf phase, batched: run it when enough lanes wait on a leaf, or nobody can walk ----
    const bool atLeaf = alive && cur < 0;
    unsigned waiting = __ballot_sync(FULL_MASK, atLeaf);
    unsigned walking = __ballot_sync(FULL_MASK, alive && cur >= 0);
    if (waiting && (__popc(waiting) >= DRT_LEAF_BATCH || walking == 0)) {
      if (atLeaf) {
        uint32_t off = refLeafOffset(cur), cnt = refLeafCountField(cur);
        const GPrim* pr = sc.prims + off;
        if (cnt == 15u) cnt = (uint32_t)__ldg(&pr->leafCount);
        float4 o = __ldg(rayO + rayIdx), d = __ldg(rayD + rayIdx);  // direction is only needed here
        RayState rs;
        rs.ox = o.x; rs.oy = o.y; rs.oz = o.z;
        rs.dx = d.x; rs.dy = d.y; rs.dz = d.z;
        rs.mint = r.mint; rs.maxt = r.maxt;
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
        if (!ANY && rs.maxt != r.maxt) {
          r.maxt = rs.maxt;
          r.maxtLo = __double2float_rd(rs.maxt);
          r.maxtHi = __double2float_ru(rs.maxt);
        }
        if (ANY && found) {
          RETIRE();
        } else {
          bool ok;
          POP_NEXT(ok);
          if (!ok) RETIRE();
        }
      }
    }
  }
#undef RETIRE
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
