// Production traversal kernels: persistent warps over a two-level-collapsed (4-wide) BVH, float32
// CONSERVATIVE interior tests, exact leaf decisions, batched exact leaf phase.
//
// Why this returns exactly what the reference's walk (bvh_accel.dart:101-226) returns:
//  (1) The reference tests a leaf's primitives iff the leaf's own slab test passes at the moment the
//      leaf node is visited; every ancestor test is implied: child boxes are contained in parent
//      boxes (exact float32 min/max unions) and IEEE rounding is monotone, so tmin_parent <=
//      tmin_child and tmax_parent >= tmax_child for the reference's own f64 expressions, and
//      maxDistance only shrinks.  Hence visiting a SUPERSET of interior nodes, in the same depth-
//      first near/far order, cannot change which primitives are tested, in which order, with which
//      maxDistance — provided each LEAF box decision is the reference's.
//  (2) Interior boxes are therefore tested in float32 with an outward margin (never a false miss).
//  (3) A leaf box is decided by the float32 filter only when the margin proves the f64 decision;
//      otherwise the leaf is marked "undecided" and the leaf phase evaluates the reference's slab
//      test in f64 (slabs(), trace_device.cuh) on the leaf box rebuilt from its primitives.
//  (4) Primitive tests are the reference's arithmetic (triangle.dart:44-98 / 162-194, sphere.dart).
//
//   reference   t = ((double)b - (double)o) * (double)invDir        (b, o, invDir are float32)
//   filter      t' = (b - o) * invDir in float32  ->  |t - t'| <= 2^-23 |t'| (+ underflow)
//   margins     hi(x) = x + 2^-22|x| + tiny,  lo(x) = x - 2^-22|x| - tiny   (monotone in x)
//
// Rays with a non-finite origin or invDir component (zero direction components) use the exact f64
// test for every box instead of the filter.
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
  float tmin;  // float32 image of the box entry distance
};

#define DRT_EPS 2.384185791015625e-07f  // 2^-22
#define DRT_TINY 1.0e-37f

// 0 = the reference surely rejects the box, 1 = it surely accepts it, 2 = not provable in float32.
static __device__ __forceinline__ int slabFilter(const FastRay& r, const float lox, const float loy, const float loz,
                                                 const float hix, const float hiy, const float hiz, float* tminOut) {
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
  return fail ? 0 : (pass ? 1 : 2);
}

// The reference's slab test in f64 (bvh_accel.dart:439-472), arguments BY VALUE so the caller's ray
// state stays in registers.  *tminOut = float32(tmin).
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

// One slot of a wide node: conservative for interior children, exact-or-marked for leaf children.
// Returns whether to visit; may mark a leaf reference "undecided".
static __device__ __forceinline__ bool testSlot(const FastRay& r, int32_t& ref, const float lox, const float loy,
                                                const float loz, const float hix, const float hiy, const float hiz,
                                                float* tmin) {
  if (ref == DRT_REF_EMPTY) return false;
  if (r.negMask & 8u)  // slow ray: exact decision for every box
    return slabExact(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, r.mint, r.maxt, lox, loy, loz, hix, hiy, hiz, tmin);
  int c = slabFilter(r, lox, loy, loz, hix, hiy, hiz, tmin);
  if (c == 2 && ref < 0) ref = refMarkUndecided(ref);
  return c != 0;
}

// Pops until an entry survives the reference's pop-time test `tmin < ray.maxDistance`
// (bvh_accel.dart:139-143,156-159 + :471).  Entries inside the undecidable band are entered:
// interior ones conservatively, leaf ones marked "undecided" for the exact leaf phase.
#define POP_NEXT(ok)                                                                \
  do {                                                                              \
    ok = false;                                                                     \
    while (sp > 0) {                                                                \
      --sp;                                                                         \
      int32_t ref_ = stack[sp].ref;                                                 \
      float t_ = stack[sp].tmin;                                                    \
      float dt_ = fmaf(DRT_EPS, fabsf(t_), DRT_TINY);                               \
      if ((t_ - dt_) >= r.maxtHi) continue; /* surely culled */                     \
      if (!((t_ + dt_) < r.maxtLo) && ref_ < 0) ref_ = refMarkUndecided(ref_);      \
      cur = ref_;                                                                   \
      ok = true;                                                                    \
      break;                                                                        \
    }                                                                               \
  } while (0)

template <bool ANY>
__global__ void __launch_bounds__(128, DRT_MIN_BLOCKS)
    traceFastKernel(TraceScene sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD, uint64_t n,
                    float4* __restrict__ hits, uint8_t* __restrict__ occluded, unsigned long long* __restrict__ nextRay,
                    TraceExtras ex) {
  const unsigned lane = threadIdx.x & 31u;
  if (ex.nDev) n = *ex.nDev;  // wavefront queues: the ray count lives in device memory
  const unsigned ltMask = (1u << lane) - 1u;
  unsigned long long warpNext = 0, warpEnd = 0;  // warp-uniform: the chunk of rays this warp owns
  bool exhausted = false;                        // warp-uniform: the global counter ran past n
  bool alive = false;
  unsigned long long rayIdx = 0;
  FastRay r;
  StackEntry stack[100];  // <= 3 pushes per wide level, <= 32 wide levels (binary depth < 64)
  int sp = 0;
  int32_t cur = 0;
  float hb1 = 0.f, hb2 = 0.f;
  int hprim = -1;
  bool found = false;

#define RETIRE()                                                                                             \
  do {                                                                                                       \
    alive = false;                                                                                           \
    if (ANY) occluded[rayIdx] = found ? 1 : 0;                                                               \
    else {                                                                                                   \
      hits[rayIdx] = make_float4(found ? __double2float_rn(r.maxt) : CUDART_INF_F, hb1, hb2,                 \
                                 __int_as_float(hprim));                                                     \
      if (ex.tOut) ex.tOut[rayIdx] = found ? r.maxt : CUDART_INF;                                            \
    }                                                                                                        \
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
        if (ex.range) {  // renderer rays: the reference's f64 minDistance / maxDistance (ray.dart:34-36)
          double2 mm = __ldg(ex.range + rayIdx);
          r.mint = mm.x; r.mintLo = __double2float_rd(mm.x); r.mintHi = __double2float_ru(mm.x);
          r.maxt = mm.y; r.maxtLo = __double2float_rd(mm.y); r.maxtHi = __double2float_ru(mm.y);
        } else {
          r.mint = o.w;
          r.mintLo = r.mintHi = o.w;
          r.maxt = d.w;
          r.maxtLo = r.maxtHi = d.w;
        }
        bool slow = !(fabsf(r.ox) <= 3.0e38f) || !(fabsf(r.oy) <= 3.0e38f) || !(fabsf(r.oz) <= 3.0e38f) ||
                    !(fabsf(r.ix) <= 3.0e38f) || !(fabsf(r.iy) <= 3.0e38f) || !(fabsf(r.iz) <= 3.0e38f);
        r.negMask = (r.ix < 0.f ? 1u : 0u) | (r.iy < 0.f ? 2u : 0u) | (r.iz < 0.f ? 4u : 0u) | (slow ? 8u : 0u);
        sp = 0;
        found = false;
        hprim = -1;
        hb1 = hb2 = 0.f;
        alive = true;
        if (sc.empty) {
          RETIRE();
        } else {
          // Reference node 0's own box (bvh_accel.dart:123-125): implied by its children's boxes when
          // the root is interior; a root LEAF gets its box decided exactly in the leaf phase.
          cur = sc.wideRootRef;
          if (cur < 0) cur = refMarkUndecided(cur);
        }
      }
      warpNext += nDead < avail ? nDead : avail;
      if (exhausted && __all_sync(FULL_MASK, !alive)) break;
    }

    // ---- one wide-node step for every lane that holds an interior node --------------------------
    if (alive && cur >= 0) {
      const float4* nd = reinterpret_cast<const float4*>(sc.wide + cur);
      float4 q0 = __ldg(nd + 0), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3), q4 = __ldg(nd + 4),
             q5 = __ldg(nd + 5);
      int4 qr = __ldg(reinterpret_cast<const int4*>(nd + 6));
      int4 qa = __ldg(reinterpret_cast<const int4*>(nd + 7));
      float t0, t1, t2, t3;
      int32_t r0 = qr.x, r1 = qr.y, r2 = qr.z, r3 = qr.w;
      bool p0 = testSlot(r, r0, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0);
      bool p1 = testSlot(r, r1, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1);
      bool p2 = testSlot(r, r2, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, &t2);
      bool p3 = testSlot(r, r3, q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, &t3);
      // visiting order of the reference's depth-first walk (bvh_accel.dart:147-153)
      const bool sA = ((r.negMask >> qa.y) & 1u) != 0, sB = ((r.negMask >> qa.z) & 1u) != 0,
                 sP = ((r.negMask >> qa.x) & 1u) != 0;
      if (sA) { int32_t tr = r0; r0 = r1; r1 = tr; float tt = t0; t0 = t1; t1 = tt; bool tp = p0; p0 = p1; p1 = tp; }
      if (sB) { int32_t tr = r2; r2 = r3; r3 = tr; float tt = t2; t2 = t3; t3 = tt; bool tp = p2; p2 = p3; p3 = tp; }
      if (sP) {
        int32_t tr = r0; r0 = r2; r2 = tr; tr = r1; r1 = r3; r3 = tr;
        float tt = t0; t0 = t2; t2 = tt; tt = t1; t1 = t3; t3 = tt;
        bool tp = p0; p0 = p2; p2 = tp; tp = p1; p1 = p3; p3 = tp;
      }
      // first passing slot becomes current, the later ones are pushed so that they pop in order
      if (p3 && (p0 || p1 || p2)) { stack[sp].ref = r3; stack[sp].tmin = t3; sp++; }
      if (p2 && (p0 || p1)) { stack[sp].ref = r2; stack[sp].tmin = t2; sp++; }
      if (p1 && p0) { stack[sp].ref = r1; stack[sp].tmin = t1; sp++; }
      if (p0) cur = r0;
      else if (p1) cur = r1;
      else if (p2) cur = r2;
      else if (p3) cur = r3;
      else {
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
        bool boxOk = true;
        if (refLeafUndecided(cur)) {
          // the reference's own slab test of this leaf node, on the box rebuilt from its primitives
          // (triangle.dart:39-42 / the sphere's world bound)
          float lo0 = CUDART_INF_F, lo1 = CUDART_INF_F, lo2 = CUDART_INF_F, hi0 = -CUDART_INF_F, hi1 = -CUDART_INF_F,
                hi2 = -CUDART_INF_F;
          for (uint32_t k = 0; k < cnt; ++k) {
            float4 a = ldg4(&pr[k].p1[0]), b = ldg4(&pr[k].p2[0]), c = ldg4(&pr[k].p3[0]);
            int kind = __float_as_int(c.w);
            if ((kind & 1) == 0) {
              lo0 = fminf(lo0, fminf(a.x, fminf(b.x, c.x))); hi0 = fmaxf(hi0, fmaxf(a.x, fmaxf(b.x, c.x)));
              lo1 = fminf(lo1, fminf(a.y, fminf(b.y, c.y))); hi1 = fmaxf(hi1, fmaxf(a.y, fmaxf(b.y, c.y)));
              lo2 = fminf(lo2, fminf(a.z, fminf(b.z, c.z))); hi2 = fmaxf(hi2, fmaxf(a.z, fmaxf(b.z, c.z)));
            } else {
              const GSphere& s = sc.spheres[kind >> 1];
              lo0 = fminf(lo0, s.wmin[0]); hi0 = fmaxf(hi0, s.wmax[0]);
              lo1 = fminf(lo1, s.wmin[1]); hi1 = fmaxf(hi1, s.wmax[1]);
              lo2 = fminf(lo2, s.wmin[2]); hi2 = fmaxf(hi2, s.wmax[2]);
            }
          }
          float tt;
          boxOk = slabExact(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz, r.mint, r.maxt, lo0, lo1, lo2, hi0, hi1, hi2, &tt);
        }
        bool stop = false;
        for (uint32_t k = 0; boxOk && k < cnt && !stop; ++k) {
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
                            unsigned long long* nextRay, int numSMs, cudaStream_t stream, const TraceExtras* extras) {
  TraceExtras ex{};
  if (extras) ex = *extras;
  if (n == 0 && !ex.nDev) return cudaSuccess;
  if (ex.nDev) n = ~0ull >> 8;  // unknown on the host: launch the full persistent grid
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
  if (any) traceFastKernel<true><<<grid, block, 0, stream>>>(sc, o, d, n, nullptr, (uint8_t*)out, nextRay, ex);
  else traceFastKernel<false><<<grid, block, 0, stream>>>(sc, o, d, n, (float4*)out, nullptr, nextRay, ex);
  return cudaGetLastError();
}

}  // namespace drt
