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
#include <cstdlib>
#include <algorithm>

#include "gpu_types.h"
#include "trace_device.cuh"
#include "trace_kernels.h"

namespace drt {

#define FULL_MASK 0xffffffffu
#ifndef RAY_CHUNK
#define RAY_CHUNK 64  // measured on B200: 64 beats 128 / 256 (shorter tail, tools/variant_sweep.sh)
#endif  // rays a warp reserves per atomicAdd on the global ray counter
#ifndef DRT_MIN_BLOCKS
#define DRT_MIN_BLOCKS 5
#endif
#ifndef DRT_SMEM_STACK
#define DRT_SMEM_STACK 24  // stack entries per thread kept in shared memory (24 KB per 128-thread CTA)
#endif
#ifndef DRT_REFILL_MIN
#define DRT_REFILL_MIN 4  // idle lanes before a warp refills from its chunk: 4 beats 1 / 8 / 12 (tools/variant_sweep.sh, +2 %)
#endif
#ifndef DRT_LEAF_BATCH
#define DRT_LEAF_BATCH 16  // lanes holding an untested leaf before the warp runs the exact leaf phase
#endif

// Hot per-ray state, kept in registers (never address-taken: the exact helpers take it by value).
struct FastRay {
  // ray.origin and invDir = (float)(1.0 / (double)d) (bvh_accel.dart:109-111), each value duplicated into both
  // halves of an f32x2 operand
  unsigned long long ox2, oy2, oz2, ix2, iy2, iz2;
  float mintLo, mintHi, maxtLo, maxtHi;  // float32 brackets of the f64 interval ends
  double mint, maxt;
  unsigned negMask;  // bit a = invDir[a] < 0 (dirIsNeg, bvh_accel.dart:113-115); bit 3 = "slow" ray; bits 8.. = 3 * octant
};

struct StackEntry {
  int32_t ref;
  float tmin;  // float32 image of the box entry distance
};

#define DRT_EPS 2.384185791015625e-07f  // 2^-22
#define DRT_TINY 1.0e-37f

// Blackwell packed float32 arithmetic (FADD2 / FMUL2): both lanes are IEEE round-to-nearest, i.e. the
// same values two scalar instructions give.
static __device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
static __device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
// Traversal-stack accesses with explicit 32-bit shared-space addresses: the generic-pointer forms cost an S2R + address
// arithmetic per access (the compiler has to pick between the shared window and local memory).
static __device__ __forceinline__ void sts64(unsigned addr, int32_t ref, float t) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(ref), "r"(__float_as_uint(t)) : "memory");
}
static __device__ __forceinline__ uint2 lds64(unsigned addr) {
  uint2 e;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(addr) : "memory");
  return e;
}
static __device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
static __device__ __forceinline__ void slabPair(unsigned long long box, unsigned long long o2, unsigned long long i2, float* t0,
                                                float* t1) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(box), "l"(o2));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(d), "l"(i2));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(*t0), "=f"(*t1) : "l"(d));
}

// Float32 image of the reference's slab test on one box given as three (lo, hi) pairs.
// Returns 0 = the reference surely rejects the box, 1 = it surely accepts it, 2 = not provable in float32.
static __device__ __forceinline__ int slabFilter(const FastRay& r, unsigned long long bx, unsigned long long by,
                                                 unsigned long long bz, float* tminOut) {
  float t0x, t1x, t0y, t1y, t0z, t1z;
  slabPair(bx, r.ox2, r.ix2, &t0x, &t1x);
  slabPair(by, r.oy2, r.iy2, &t0y, &t1y);
  slabPair(bz, r.oz2, r.iz2, &t0z, &t1z);
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

// One slot of a wide node for a regular ray: conservative for interior children, exact-or-marked for leaf children.
// Returns the reference to visit (a leaf reference possibly marked "undecided") or DRT_REF_EMPTY.
static __device__ __forceinline__ int32_t testSlot(const FastRay& r, int32_t ref, const float2 bx, const float2 by,
                                                   const float2 bz, float* tmin) {
  int c = slabFilter(r, pack2(bx.x, bx.y), pack2(by.x, by.y), pack2(bz.x, bz.y), tmin);
  int32_t marked = (c == 2 && ref < 0) ? refMarkUndecided(ref) : ref;  // an EMPTY slot is positive: never marked
  return c != 0 ? marked : DRT_REF_EMPTY;
}

// The same slot for a "slow" ray (non-finite origin or invDir component; warp-divergent, rare): the reference's own
// f64 decision for every box.
static __device__ __forceinline__ int32_t testSlotExact(const FastRay& r, int32_t ref, const float2 bx, const float2 by,
                                                        const float2 bz, float* tmin) {
  *tmin = 0.f;
  if (ref == DRT_REF_EMPTY) return ref;
  float te = 0.f;  // only this temporary is address-taken: the caller's t0..t3 stay in registers
  const bool ok = slabExact(lo32(r.ox2), lo32(r.oy2), lo32(r.oz2), lo32(r.ix2), lo32(r.iy2), lo32(r.iz2), r.mint, r.maxt, bx.x, by.x,
                            bz.x, bx.y, by.y, bz.y, &te);
  *tmin = te;
  return ok ? ref : DRT_REF_EMPTY;
}

// Pops until an entry survives the reference's pop-time test `tmin < ray.maxDistance`
// (bvh_accel.dart:139-143,156-159 + :471).  Entries inside the undecidable band are entered:
// interior ones conservatively, leaf ones marked "undecided" for the exact leaf phase.
#define POP_NEXT(ok)                                                                \
  do {                                                                              \
    ok = false;                                                                     \
    while (sp > 0) {                                                                \
      --sp;                                                                         \
      const uint2 e_ = sp < DRT_SMEM_STACK ? lds64(smBase + (unsigned)sp * 1024u) : deepStack[sp - DRT_SMEM_STACK]; \
      int32_t ref_ = (int32_t)e_.x;                                                 \
      if (ANY) { /* maxDistance never shrinks: neither culled nor re-marked */      \
        cur = ref_;                                                                 \
        ok = true;                                                                  \
        break;                                                                      \
      }                                                                             \
      float t_ = __uint_as_float(e_.y);                                             \
      float dt_ = fmaf(DRT_EPS, fabsf(t_), DRT_TINY);                               \
      if ((t_ - dt_) >= r.maxtHi) continue; /* surely culled */                     \
      if (!((t_ + dt_) < r.maxtLo) && ref_ < 0) ref_ = refMarkUndecided(ref_);      \
      cur = ref_;                                                                   \
      ok = true;                                                                    \
      break;                                                                        \
    }                                                                               \
  } while (0)

// QUAD: the leaf code the scene needs (TraceScene::quadMode): 0 triangles only, 1 + spheres / disks, 2 + the remaining quadrics.
template <bool ANY, int QUAD>
__global__ void __launch_bounds__(128, DRT_MIN_BLOCKS)
    traceFastKernel(TraceScene sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD, uint32_t n,
                    float4* __restrict__ hits, uint8_t* __restrict__ occluded, unsigned int* __restrict__ nextRay,
                    TraceExtras ex) {
  const unsigned lane = threadIdx.x & 31u;
  if (ex.nDev) n = *ex.nDev;  // wavefront queues: the ray count lives in device memory
  const unsigned ltMask = (1u << lane) - 1u;
  uint32_t warpNext = 0, warpEnd = 0;  // warp-uniform: the chunk of rays this warp owns (launches hold < 2^31 rays)
  bool exhausted = false;                        // warp-uniform: the global counter ran past n
  bool alive = false;
  uint32_t rayIdx = 0;
  FastRay r;
  // Traversal stack: the first DRT_SMEM_STACK entries of every thread live in shared memory, laid out
  // [entry][thread] so that a warp's accesses are bank-conflict free whatever the lanes' depths; deeper
  // entries (rare) overflow to local memory.  <= 3 pushes per wide level, <= 32 wide levels.
  extern __shared__ uint2 smStack[];
  float* smDir = reinterpret_cast<float*>(smStack + DRT_SMEM_STACK * 128);  // [3][128]: ray.direction, read by the leaf phase
  uint2 deepStack[104 - DRT_SMEM_STACK];
#define STACK_STORE(i, refv, tv)                                                                   \
  do {                                                                                             \
    const uint2 e__ = make_uint2((unsigned)(refv), __float_as_uint(tv));                           \
    if ((i) < DRT_SMEM_STACK) smStack[(i) * 128 + threadIdx.x] = e__;                              \
    else deepStack[(i) - DRT_SMEM_STACK] = e__;                                                    \
  } while (0)
#define STACK_LOAD(i) ((i) < DRT_SMEM_STACK ? smStack[(i) * 128 + threadIdx.x] : deepStack[(i) - DRT_SMEM_STACK])
  // entry i of this thread: smBase + i * 1024.  Computed by a volatile asm so that the compiler keeps it in a register
  // instead of rematerialising it (S2R SR_CgaCtaId + LEA) at every stack access.
  unsigned smBase;
  asm volatile("{ .reg .u64 t64; cvta.to.shared.u64 t64, %1; cvt.u32.u64 %0, t64; }" : "=r"(smBase) : "l"(smStack));
  smBase += threadIdx.x * 8u;
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
    // ray set-up runs at a handful of lanes: wait until DRT_REFILL_MIN lanes are idle (or nothing is left to wait for)
    if (dead && (__popc(dead) >= DRT_REFILL_MIN || dead == FULL_MASK || (exhausted && warpNext == warpEnd))) {
      if (warpNext == warpEnd && !exhausted) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(nextRay, (unsigned)RAY_CHUNK);
        base = __shfl_sync(FULL_MASK, base, 0);
        if (base >= n) {
          exhausted = true;
        } else {
          warpNext = base;
          warpEnd = (base + RAY_CHUNK) < n ? (base + RAY_CHUNK) : n;
#ifndef DRT_NO_RAY_PREFETCH
          // the chunk's ray records (2 x 16 B per ray) are streamed from HBM: start them towards L1 now so
          // that the lane-by-lane refills below do not each wait for DRAM
          {
            const uint32_t linesPerArray = (RAY_CHUNK * 16 + 127) / 128;
            if (lane < 2 * linesPerArray) {
              const float4* basePtr = (lane < linesPerArray ? rayO : rayD) + base;
              const char* pf = reinterpret_cast<const char*>(basePtr) + 128 * (lane % linesPerArray);
              asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
            }
          }
#endif
        }
      }
      unsigned avail = (unsigned)(warpEnd - warpNext);
      unsigned nDead = __popc(dead);
      unsigned rank = __popc(dead & ltMask);
      if (!alive && rank < avail) {
        rayIdx = warpNext + rank;
        float4 o = __ldg(rayO + rayIdx), d = __ldg(rayD + rayIdx);
        // invDir = (float)(1.0 / (double)d) (bvh_accel.dart:109-111).  The correctly rounded float32 quotient is the same
        // value: rounding a quotient to 53 bits and then to 24 cannot differ from rounding it to 24 directly
        // (53 >= 2 * 24 + 2), subnormal and infinite results included.  rcp.rn.f32 is that quotient.
        const float ix = __frcp_rn(d.x), iy = __frcp_rn(d.y), iz = __frcp_rn(d.z);
        smDir[threadIdx.x] = d.x; smDir[128 + threadIdx.x] = d.y; smDir[256 + threadIdx.x] = d.z;
        r.ox2 = pack2(o.x, o.x); r.oy2 = pack2(o.y, o.y); r.oz2 = pack2(o.z, o.z);
        r.ix2 = pack2(ix, ix); r.iy2 = pack2(iy, iy); r.iz2 = pack2(iz, iz);
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
        bool slow = !(fabsf(o.x) <= 3.0e38f) || !(fabsf(o.y) <= 3.0e38f) || !(fabsf(o.z) <= 3.0e38f) ||
                    !(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f) || !(fabsf(iz) <= 3.0e38f);
        r.negMask = (ix < 0.f ? 1u : 0u) | (iy < 0.f ? 2u : 0u) | (iz < 0.f ? 4u : 0u) | (slow ? 8u : 0u);
        r.negMask |= (3u * (r.negMask & 7u)) << 8;  // bits 8..: shift of this ray's octant in a node's orderLut
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
    bool needPop = false;
    if (alive && cur >= 0) {
      // the 128-byte node in four 256-bit loads (LDG.E.256, sm_100): half the L1 requests of 128-bit loads
      const GNode4* nd = sc.wide + cur;
      float4 q0, q1, q2, q3, q4, q5;
      int4 qr, qa;
      ldg256(reinterpret_cast<const float*>(nd), q0, q1);
      ldg256(reinterpret_cast<const float*>(nd) + 8, q2, q3);
      ldg256(reinterpret_cast<const float*>(nd) + 16, q4, q5);
      {
        float4 fr, fa;
        ldg256(reinterpret_cast<const float*>(nd) + 24, fr, fa);
        qr = make_int4(__float_as_int(fr.x), __float_as_int(fr.y), __float_as_int(fr.z), __float_as_int(fr.w));
        qa = make_int4(__float_as_int(fa.x), __float_as_int(fa.y), __float_as_int(fa.z), __float_as_int(fa.w));
      }
      float t0, t1, t2, t3;
      // slot k: (lo.x, hi.x) (lo.y, hi.y) (lo.z, hi.z), six consecutive floats
      int32_t r0, r1, r2, r3;
      if (!(r.negMask & 8u)) {
        r0 = testSlot(r, qr.x, make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y), &t0);
        r1 = testSlot(r, qr.y, make_float2(q1.z, q1.w), make_float2(q2.x, q2.y), make_float2(q2.z, q2.w), &t1);
        r2 = testSlot(r, qr.z, make_float2(q3.x, q3.y), make_float2(q3.z, q3.w), make_float2(q4.x, q4.y), &t2);
        r3 = testSlot(r, qr.w, make_float2(q4.z, q4.w), make_float2(q5.x, q5.y), make_float2(q5.z, q5.w), &t3);
      } else {
        r0 = testSlotExact(r, qr.x, make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y), &t0);
        r1 = testSlotExact(r, qr.y, make_float2(q1.z, q1.w), make_float2(q2.x, q2.y), make_float2(q2.z, q2.w), &t1);
        r2 = testSlotExact(r, qr.z, make_float2(q3.x, q3.y), make_float2(q3.z, q3.w), make_float2(q4.x, q4.y), &t2);
        r3 = testSlotExact(r, qr.w, make_float2(q4.z, q4.w), make_float2(q5.x, q5.y), make_float2(q5.z, q5.w), &t3);
      }
      // visiting order of the reference's depth-first walk (bvh_accel.dart:147-153), branch-free
      // closest hit: the node carries its three near/far decisions for each dirIsNeg octant (orderLut).
      // any hit: the answer is the OR over all leaves whose box test passes, whatever the visiting order, so the
      // slots are walked as stored (the builder put the larger boxes first).
      const unsigned dec = ANY ? 0u : ((unsigned)qa.w >> (r.negMask >> 8));  // the node's decisions for this ray's octant
      const bool sP = (dec & 1u) != 0, sA = (dec & 2u) != 0, sB = (dec & 4u) != 0;
      const int32_t a0 = sA ? r1 : r0, a1 = sA ? r0 : r1, b0 = sB ? r3 : r2, b1 = sB ? r2 : r3;
      const float ta0 = sA ? t1 : t0, ta1 = sA ? t0 : t1, tb0 = sB ? t3 : t2, tb1 = sB ? t2 : t3;
      const int32_t s0 = sP ? b0 : a0, s1 = sP ? b1 : a1, s2 = sP ? a0 : b0, s3 = sP ? a1 : b1;
      const float u1 = sP ? tb1 : ta1, u2 = sP ? ta0 : tb0, u3 = sP ? ta1 : tb1;
      const bool v0 = s0 != DRT_REF_EMPTY, v1 = s1 != DRT_REF_EMPTY, v2 = s2 != DRT_REF_EMPTY, v3 = s3 != DRT_REF_EMPTY;
      // the first passing slot becomes current; the later ones are pushed so that they pop in order.
      // Stores are unconditional, the stack pointer moves only for real pushes.
      const int p3 = (v3 && (v0 || v1 || v2)) ? 1 : 0, p2 = (v2 && (v0 || v1)) ? 1 : 0, p1 = (v1 && v0) ? 1 : 0;
      if (sp <= DRT_SMEM_STACK - 3) {  // all three stores land in shared memory (the common case)
        unsigned a = smBase + (unsigned)sp * 1024u;
        sts64(a, s3, u3); a += p3 ? 1024u : 0u;
        sts64(a, s2, u2); a += p2 ? 1024u : 0u;
        sts64(a, s1, u1);
        sp += p3 + p2 + p1;
      } else {
        STACK_STORE(sp, s3, u3); sp += p3;
        STACK_STORE(sp, s2, u2); sp += p2;
        STACK_STORE(sp, s1, u1); sp += p1;
      }
      cur = v0 ? s0 : (v1 ? s1 : (v2 ? s2 : s3));
      needPop = !(v0 || v1 || v2 || v3);
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
        RayState rs;  // origin from the packed registers, direction from shared memory (only needed here)
        rs.ox = lo32(r.ox2); rs.oy = lo32(r.oy2); rs.oz = lo32(r.oz2);
        rs.dx = smDir[threadIdx.x]; rs.dy = smDir[128 + threadIdx.x]; rs.dz = smDir[256 + threadIdx.x];
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
          boxOk = slabExact(lo32(r.ox2), lo32(r.oy2), lo32(r.oz2), lo32(r.ix2), lo32(r.iy2), lo32(r.iz2), r.mint, r.maxt, lo0, lo1, lo2, hi0, hi1, hi2, &tt);
        }
        bool stop = false;
        for (uint32_t k = 0; boxOk && k < cnt && !stop; ++k) {
          float4 a = ldg4(&pr[k].p1[0]), b = ldg4(&pr[k].p2[0]), c = ldg4(&pr[k].p3[0]);
          int kind = __float_as_int(c.w);
          if (QUAD == 0 || (kind & 1) == 0) {
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
              if (sphereTest<QUAD == 2>(s, rs, true, &th, nullptr, nullptr)) { found = true; stop = true; }
            } else if (sphereTest<QUAD == 2>(s, rs, false, &th, ex.noUV ? nullptr : &u, &v)) {
              found = true;
              hb1 = ex.noUV ? 0.f : __double2float_rn(u); hb2 = ex.noUV ? 0.f : __double2float_rn(v); hprim = __float_as_int(a.w);
              rs.maxt = th;
            }
          }
        }
        if (!ANY && rs.maxt != r.maxt) {
          r.maxt = rs.maxt;
          r.maxtLo = __double2float_rd(rs.maxt);
          r.maxtHi = __double2float_ru(rs.maxt);
        }
        if (ANY && found) RETIRE();
        else needPop = true;
      }
    }

    // ---- next stack entry for every lane that finished a node without a child or a leaf ----------
    if (alive && needPop) {
      bool ok;
      POP_NEXT(ok);
      if (!ok) RETIRE();
    }
  }
#undef RETIRE
}

// ---------------------------------------------------------------------------------------------------------------------
// Small scenes (<= DRT_SMALL_MAX_LEAVES leaves, no instances: BASELINE.json config 4 has 25 primitives in 11 leaves): no tree, no
// stack, no persistent warps.  By argument (1) above the reference's answer is fixed by (a) the ORDER in which its walk reaches the
// leaves, which depends on the ray's dirIsNeg octant only (bvh_accel.dart:147-153) and is tabulated by the host (GSmallScene::order),
// and (b) each leaf's own box decision at the moment it is reached, every interior test being implied by it.  One thread per ray:
//   pass 1  the float32 filter of every leaf box, the leaves read in storage order (one shared-memory broadcast per leaf for the whole
//           warp, no divergence); survivors are marked in a 32-bit mask at their POSITION in the octant's visiting order;
//   pass 2  the marked leaves in that order: the filter again with the ray's maxDistance of this moment, the reference's binary64 slab
//           test where the filter cannot prove the decision, then the reference's primitive tests (trace_device.cuh).
// A leaf dropped in pass 1 fails the reference's test at any later moment too (maxDistance only shrinks), so the two passes test the
// primitives of exactly the leaves the reference enters, in its order, with its maxDistance.  "Slow" rays (non-finite origin or
// invDir) mark every leaf and decide each box in binary64.  Any hit: the order is immaterial, octant 0's is used.
#ifndef DRT_SMALL_BLOCK
#define DRT_SMALL_BLOCK 256
#endif
#ifndef DRT_SMALL_MIN_BLOCKS
#define DRT_SMALL_MIN_BLOCKS 4  // 64 registers.  Config 4 on B200 (profiles/r02z5_variants_ab.log): 3 (80 registers) 0.5648 s, 4 0.5552 s
#endif
#ifndef DRT_SMALL_QUAD_BATCH
#define DRT_SMALL_QUAD_BATCH 1  // lanes waiting at a quadric before the warp runs the quadric test.  Measured (same log): 1 (no
                                // waiting) 0.5478 s, 8 0.5648 s, 16 0.5645 s — the waiting lanes' extra trips cost more than the fuller test
#endif
template <bool ANY, int QUAD>
__global__ void __launch_bounds__(DRT_SMALL_BLOCK, DRT_SMALL_MIN_BLOCKS)
    traceSmallKernel(TraceScene sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD, uint32_t n,
                     float4* __restrict__ hits, uint8_t* __restrict__ occluded, TraceExtras ex) {
  __shared__ GSmallScene sm;
  {
    const uint4* src = reinterpret_cast<const uint4*>(sc.small);
    uint4* dst = reinterpret_cast<uint4*>(&sm);
    for (unsigned i = threadIdx.x; i < sizeof(GSmallScene) / 16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  if (ex.nDev) n = *ex.nDev;
  const int nLeaves = sm.nLeaves;
  // the loop runs the same number of trips for every lane of a warp (a lane beyond n idles): the leaf loop below votes
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t rayIdx = base + threadIdx.x;
    const bool valid = rayIdx < n;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(1.f, 1.f, 1.f, 0.f);
    if (valid) { o = __ldg(rayO + rayIdx); d = __ldg(rayD + rayIdx); }
    const float ix = __frcp_rn(d.x), iy = __frcp_rn(d.y), iz = __frcp_rn(d.z);  // invDir, see traceFastKernel
    FastRay r;
    r.ox2 = pack2(o.x, o.x); r.oy2 = pack2(o.y, o.y); r.oz2 = pack2(o.z, o.z);
    r.ix2 = pack2(ix, ix); r.iy2 = pack2(iy, iy); r.iz2 = pack2(iz, iz);
    if (valid && ex.range) {
      const double2 mm = __ldg(ex.range + rayIdx);
      r.mint = mm.x; r.mintLo = __double2float_rd(mm.x); r.mintHi = __double2float_ru(mm.x);
      r.maxt = mm.y; r.maxtLo = __double2float_rd(mm.y); r.maxtHi = __double2float_ru(mm.y);
    } else {
      r.mint = o.w; r.mintLo = r.mintHi = o.w;
      r.maxt = d.w; r.maxtLo = r.maxtHi = d.w;
    }
    const bool slow = !(fabsf(o.x) <= 3.0e38f) || !(fabsf(o.y) <= 3.0e38f) || !(fabsf(o.z) <= 3.0e38f) ||
                      !(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f) || !(fabsf(iz) <= 3.0e38f);
    const unsigned oct = ANY ? 0u : ((ix < 0.f ? 1u : 0u) | (iy < 0.f ? 2u : 0u) | (iz < 0.f ? 4u : 0u));
    r.negMask = oct;
    // ---- pass 1: which leaves can the reference enter at all (storage order: one broadcast per leaf) --------------------
    unsigned mask = 0u;
    if (!valid || sc.empty) {
      mask = 0u;
    } else if (slow) {
      mask = nLeaves >= 32 ? 0xffffffffu : ((1u << nLeaves) - 1u);
    } else {
      for (int l = 0; l < nLeaves; ++l) {
        const GSmallLeaf& L = sm.leaf[l];
        float tm;
        const int c = slabFilter(r, pack2(L.lo[0], L.hi[0]), pack2(L.lo[1], L.hi[1]), pack2(L.lo[2], L.hi[2]), &tm);
        if (c) mask |= 1u << sm.position[oct][l];
      }
    }
    RayState rs;
    rs.ox = o.x; rs.oy = o.y; rs.oz = o.z;
    rs.dx = d.x; rs.dy = d.y; rs.dz = d.z;
    rs.mint = r.mint; rs.maxt = r.maxt;
    bool found = false;
    float hb1 = 0.f, hb2 = 0.f;
    int hprim = -1;
    // ---- pass 2: the marked leaves in the octant's order.  Every trip of the warp-uniform loop first gives each lane that has
    //      finished its leaf the next one it may enter (filter with the maxDistance of this moment, binary64 where undecided), then
    //      tests ONE primitive per lane, so that the long binary64 primitive tests run with as many lanes as have work ------------
    const GPrim* pr = sc.prims;
    uint32_t k = 0, cnt = 0;
    for (;;) {
      if (k >= cnt && !(ANY && found)) {
        while (mask) {
          const int j = __ffs((int)mask) - 1;
          mask &= mask - 1u;
          const GSmallLeaf& L = sm.leaf[sm.order[oct][j]];
          const float lo0 = L.lo[0], lo1 = L.lo[1], lo2 = L.lo[2], hi0 = L.hi[0], hi1 = L.hi[1], hi2 = L.hi[2];
          const int32_t ref = L.ref;
          int c = 2;
          float tm;
          if (!slow) c = slabFilter(r, pack2(lo0, hi0), pack2(lo1, hi1), pack2(lo2, hi2), &tm);
          if (c == 0) continue;
          if (c == 2 && !slabExact(o.x, o.y, o.z, ix, iy, iz, r.mint, r.maxt, lo0, lo1, lo2, hi0, hi1, hi2, &tm)) continue;
          pr = sc.prims + refLeafOffset(ref);
          cnt = refLeafCountField(ref);
          if (cnt == 15u) cnt = (uint32_t)__ldg(&pr->leafCount);
          k = 0;
          break;
        }
      }
      __syncwarp();
      const bool work = k < cnt && !(ANY && found);
      if (!__any_sync(FULL_MASK, work)) break;
      // triangles at once; a lane whose next primitive is a quadric (a long, rarely taken path: a handful of lanes per warp reach
      // the sphere's leaf in a given trip) waits until DRT_SMALL_QUAD_BATCH lanes do, or nobody has a triangle left to test
      bool atQuad = false;
      int quadIdx = 0, quadPrim = 0;
      if (work) {
        const float4 a = ldg4(&pr[k].p1[0]), b = ldg4(&pr[k].p2[0]), cc = ldg4(&pr[k].p3[0]);
        const int kind = __float_as_int(cc.w);
        if (QUAD == 0 || (kind & 1) == 0) {
          ++k;
          if (ANY) {
            if (triangleAny(rs, a, b, cc)) found = true;
          } else {
            HitState h;
            if (triangleClosest(rs, a, b, cc, &h)) {
              found = true;
              hb1 = __double2float_rn(h.b1); hb2 = __double2float_rn(h.b2); hprim = h.prim;
            }
          }
        } else {
          atQuad = true;
          quadIdx = kind >> 1;
          quadPrim = __float_as_int(a.w);
        }
      }
      if (QUAD != 0) {
        const unsigned qb = __ballot_sync(FULL_MASK, atQuad), tb = __ballot_sync(FULL_MASK, work && !atQuad);
        if (atQuad && (__popc(qb) >= DRT_SMALL_QUAD_BATCH || tb == 0u)) {
          ++k;
          const GSphere& s = sc.spheres[quadIdx];
          double th, u, v;
          if (ANY) {
            if (sphereTest<QUAD == 2>(s, rs, true, &th, nullptr, nullptr)) found = true;
          } else if (sphereTest<QUAD == 2>(s, rs, false, &th, ex.noUV ? nullptr : &u, &v)) {
            found = true;
            hb1 = ex.noUV ? 0.f : __double2float_rn(u); hb2 = ex.noUV ? 0.f : __double2float_rn(v); hprim = quadPrim;
            rs.maxt = th;
          }
        }
      }
      if (!ANY && rs.maxt != r.maxt) {
        r.maxt = rs.maxt;
        r.maxtLo = __double2float_rd(rs.maxt);
        r.maxtHi = __double2float_ru(rs.maxt);
      }
    }
    if (valid) {
      if (ANY) {
        occluded[rayIdx] = found ? 1 : 0;
      } else {
        hits[rayIdx] = make_float4(found ? __double2float_rn(r.maxt) : CUDART_INF_F, hb1, hb2, __int_as_float(hprim));
        if (ex.tOut) ex.tOut[rayIdx] = found ? r.maxt : CUDART_INF;
      }
    }
  }
}

static cudaError_t launchSmall(const TraceScene& sc, bool any, const float4* o, const float4* d, uint32_t n, bool nUnknown, void* out,
                               int numSMs, cudaStream_t stream, const TraceExtras& ex) {
  typedef void (*KernelFn)(TraceScene, const float4*, const float4*, uint32_t, float4*, uint8_t*, TraceExtras);
  static const KernelFn kKernels[6] = {traceSmallKernel<false, 0>, traceSmallKernel<false, 1>, traceSmallKernel<false, 2>,
                                       traceSmallKernel<true, 0>,  traceSmallKernel<true, 1>,  traceSmallKernel<true, 2>};
  const int variant = (any ? 3 : 0) + (sc.quadMode < 0 ? 0 : (sc.quadMode > 2 ? 2 : sc.quadMode));
  const KernelFn kernel = kKernels[variant];
  static int perSm[6] = {0, 0, 0, 0, 0, 0};
  if (!perSm[variant]) {
    int b = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, DRT_SMALL_BLOCK, 0);
    if (e != cudaSuccess) return e;
    perSm[variant] = b > 0 ? b : 1;
  }
  const uint64_t resident = (uint64_t)numSMs * perSm[variant];
  const uint64_t wanted = nUnknown ? resident : ((uint64_t)n + DRT_SMALL_BLOCK - 1) / DRT_SMALL_BLOCK;
  static const int waves = std::getenv("DRT_SMALL_WAVES") ? std::max(1, std::atoi(std::getenv("DRT_SMALL_WAVES"))) : 8;  // A/B knob: config 4
  // with 1 / 2 / 4 / 8 / 16 / 32 waves 0.4629 / 0.4551 / 0.4500 / 0.4484 / 0.4487 / 0.4526 s (profiles/r02z22_small_waves_ab.log)
  const uint64_t cap = resident * (uint64_t)waves;  // a few waves of blocks: the tail of one wave overlaps the head of the next
  dim3 grid((unsigned)(wanted < 1 ? 1 : (wanted < cap ? wanted : cap)));
  if (nUnknown) grid.x = (unsigned)cap;
  if (any) kernel<<<grid, DRT_SMALL_BLOCK, 0, stream>>>(sc, o, d, n, nullptr, (uint8_t*)out, ex);
  else kernel<<<grid, DRT_SMALL_BLOCK, 0, stream>>>(sc, o, d, n, (float4*)out, nullptr, ex);
  return cudaGetLastError();
}

static cudaError_t launchOne(const TraceScene& sc, bool any, const float4* o, const float4* d, uint32_t n, bool nUnknown, void* out,
                             unsigned long long* nextRay, int numSMs, cudaStream_t stream, const TraceExtras& ex) {
  cudaError_t e = cudaMemsetAsync(nextRay, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  const int block = 128;
  const size_t smem = (size_t)DRT_SMEM_STACK * block * sizeof(uint2) + 3 * block * sizeof(float);
  typedef void (*KernelFn)(TraceScene, const float4*, const float4*, uint32_t, float4*, uint8_t*, unsigned int*, TraceExtras);
  static const KernelFn kKernels[6] = {traceFastKernel<false, 0>, traceFastKernel<false, 1>, traceFastKernel<false, 2>,
                                       traceFastKernel<true, 0>,  traceFastKernel<true, 1>,  traceFastKernel<true, 2>};
  const int variant = (any ? 3 : 0) + (sc.quadMode < 0 ? 0 : (sc.quadMode > 2 ? 2 : sc.quadMode));
  const KernelFn kernel = kKernels[variant];
  static int perSm[6] = {0, 0, 0, 0, 0, 0};
  if (!perSm[variant]) {
    int b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, block, smem);
    if (e != cudaSuccess) return e;
    perSm[variant] = b > 0 ? b : 1;
  }
  uint64_t want = nUnknown ? ~0ull >> 8 : ((uint64_t)n + RAY_CHUNK - 1) / RAY_CHUNK;  // one warp per chunk is enough
  uint64_t blocksWanted = (want + 3) / 4;
  uint64_t persistent = (uint64_t)numSMs * perSm[variant];  // one resident wave: persistent warps
  dim3 grid((unsigned)(blocksWanted < persistent ? blocksWanted : persistent));
  unsigned int* ctr = reinterpret_cast<unsigned int*>(nextRay);
  if (any) kernel<<<grid, block, smem, stream>>>(sc, o, d, n, nullptr, (uint8_t*)out, ctr, ex);
  else kernel<<<grid, block, smem, stream>>>(sc, o, d, n, (float4*)out, nullptr, ctr, ex);
  return cudaGetLastError();
}

cudaError_t launchTraceFast(const TraceScene& sc, bool any, const void* rayO, const void* rayD, uint64_t n, void* out,
                            unsigned long long* nextRay, int numSMs, cudaStream_t stream, const TraceExtras* extras) {
  if (sc.wideQ && !sc.small) return launchTraceQ(sc, any, rayO, rayD, n, out, nextRay, numSMs, stream, extras);
  TraceExtras ex{};
  if (extras) ex = *extras;
  const float4* o = static_cast<const float4*>(rayO);
  const float4* d = static_cast<const float4*>(rayD);
  if (sc.small) {  // a handful of leaves: the leaf-list kernel
    if (ex.nDev) return launchSmall(sc, any, o, d, 0, true, out, numSMs, stream, ex);
    const uint64_t kMaxS = 1ull << 30;
    for (uint64_t first = 0; first < n; first += kMaxS) {
      const uint32_t m = (uint32_t)(n - first < kMaxS ? n - first : kMaxS);
      TraceExtras e2 = ex;
      if (e2.range) e2.range += first;
      if (e2.tOut) e2.tOut += first;
      void* o2 = any ? (void*)((uint8_t*)out + first) : (void*)((float4*)out + first);
      cudaError_t e = launchSmall(sc, any, o + first, d + first, m, false, o2, numSMs, stream, e2);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  if (ex.nDev) return launchOne(sc, any, o, d, 0, true, out, nextRay, numSMs, stream, ex);  // count lives on the device (< 2^31)
  const uint64_t kMax = 1ull << 30;  // rays per launch: 32-bit ray indices inside the kernel
  for (uint64_t first = 0; first < n; first += kMax) {
    const uint32_t m = (uint32_t)(n - first < kMax ? n - first : kMax);
    TraceExtras e2 = ex;
    if (e2.range) e2.range += first;
    if (e2.tOut) e2.tOut += first;
    void* o2 = any ? (void*)((uint8_t*)out + first) : (void*)((float4*)out + first);
    cudaError_t e = launchOne(sc, any, o + first, d + first, m, false, o2, nextRay, numSMs, stream, e2);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace drt
