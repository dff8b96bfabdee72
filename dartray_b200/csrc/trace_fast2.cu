// Production traversal kernels, second generation: persistent warps over the 4-wide collapsed BVH with QUANTISED
// 64-byte nodes (GNode4Q: two 256-bit loads per node step instead of four — the L1TEX wavefront rate was the limiter of
// the float32-box kernel in trace_fast.cu), conservative float32 box tests, a short per-lane queue of postponed leaves,
// and an exact leaf phase.
//
// Why this returns exactly what the reference's walk (bvh_accel.dart:101-226) returns — the argument of trace_fast.cu,
// with the leaf decision moved entirely into the leaf phase:
//  (1) The reference tests a leaf's primitives iff the leaf's own slab test passes at the moment the leaf is visited;
//      every ancestor test is implied (child boxes are contained in parent boxes, IEEE rounding is monotone, and
//      maxDistance only shrinks).  Visiting a SUPERSET of the nodes, with the leaves kept in the reference's depth-first
//      near/far order, therefore cannot change which primitives are tested, in which order, with which maxDistance —
//      provided each LEAF box decision is the reference's.
//  (2) Node boxes here are only conservative: decoded box >= true box + guard (checked by the builder), evaluated in
//      float32 with an outward margin.  A box the reference enters is always entered.
//  (3) EVERY leaf that is reached gets the reference's own binary64 slab test (slabs(), trace_device.cuh) on the box
//      rebuilt from its primitives, with the ray's maxDistance of that moment, before its primitives are tested.
//  (4) Leaves are postponed, never reordered: a lane may walk on while up to DRT_PEND leaves wait in its queue, and the
//      leaf phase drains the queue in order.  What was walked in between is a superset walk (maxDistance was only
//      staler, i.e. larger); each postponed leaf still gets its exact box test with the up-to-date maxDistance, and the
//      reference's cull of an ancestor implies the leaf's own test fails.
//  (5) Primitive tests are the reference's arithmetic (triangle.dart:44-98 / 162-194, sphere.dart, ...).
//
// Float32 evaluation of a quantised plane.  Plane position p = O + q * s (s = 2^e, q = 0..255).  Reference value for a
// true plane b:  t = ((double)b - (double)o) * (double)invDir.  Here, per node and axis:
//      S' = (s * 2^15) * invDir                      (exact: a power of two)
//      C  = fl(fl(O - o) * invDir),  C'' = fl(C - S')
//      f  = 1 + q * 2^-15                            (a float32 assembled by one PRMT: 0x3F80'qq'00)
//      t' = fma(f, S', C'')                          = q * s * invDir + (O - o) * invDir up to rounding
//   |t' - t_exact(p)| <= 2^-22 |t'| + 2^-8 |s * invDir|   (two roundings of C, one of C'', one of the fma; |C| <= |t| + 255 |s invDir|)
// The relative part is covered by the margin DRT_EPS2 = 2^-21 applied to the slot's entry / exit distances, the absolute
// part by the builder's guard of 2^-6 grid steps.  Preconditions (otherwise the ray is "slow" and every box is decided in
// binary64 on the decoded box): |o| <= 2^62, 2^-60 <= |invDir| <= 2^40, all finite; builder: |coordinates| <= 2^62,
// -60 <= e <= 62 — no overflow, and underflow errors (<= 2^-149) stay below the guard (2^-6 * 2^-60 * 2^-60).
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>

#include "gpu_types.h"
#include "trace_device.cuh"
#include "trace_kernels.h"

namespace drt {

#define FULL_MASK 0xffffffffu
#ifndef RAY_CHUNK
#define RAY_CHUNK 64
#endif
#ifndef DRT_Q_MIN_BLOCKS
#define DRT_Q_MIN_BLOCKS 6  // 80 registers, 24 warps per SM: 5 / 6 / 7 / 8 measured 4.20 / 3.81 / 4.09 / 5.87 ms (incoherent closest)
#endif
#ifndef DRT_Q_BLOCK
#define DRT_Q_BLOCK 128  // threads per CTA
#endif
#define DRT_Q_STRIDE (DRT_Q_BLOCK * 8u)  // bytes between two stack entries of one thread
#ifndef DRT_SMEM_STACK
#define DRT_SMEM_STACK 24
#endif
#ifndef DRT_REFILL_MIN
#define DRT_REFILL_MIN 4
#endif
#ifndef DRT_Q_LEAF_BATCH
#define DRT_Q_LEAF_BATCH 10  // blocked lanes before the warp runs the leaf phase (6 / 8 / 10: 3.93 / 3.81 / 3.76 ms)
#endif
// leaves a lane may hold before it has to wait for the leaf phase (1 = wait at every leaf).  Measured on B200 (config 2,
// profiles/r02_summary.md): postponing one leaf helps the closest-hit walk (+3 %), the any-hit walk loses 10 % to the nodes
// it visits for nothing after the hit that would have ended it
#ifndef DRT_PEND_CLOSEST
#define DRT_PEND_CLOSEST 2
#endif
#ifndef DRT_PEND_ANY
#define DRT_PEND_ANY 1
#endif
#define DRT_PEND (ANY ? DRT_PEND_ANY : DRT_PEND_CLOSEST)

#ifndef DRT_Q_LAZY_BOX
#define DRT_Q_LAZY_BOX 0  // 1: the leaf's f64 box test only once a triangle of the leaf reports a hit.  REJECTED on measurement
                         // (profiles/r02q_trace_ab.log: 4.45 vs 3.59 ms): the box test with `tmin < maxDistance` is what culls most
                         // leaves a conservative walk reaches, far cheaper than the f64 triangle tests it saves
#endif
#ifndef DRT_Q_LEAF_PRE
#define DRT_Q_LEAF_PRE 0  // 1: a conservative float32 slab test on the leaf's own (unquantised) box in front of the binary64 one (a leaf
                         // it rejects would fail the reference's test too, so only the survivors pay for f64).  Exact (full-size tests
                         // green) but REJECTED on measurement (profiles/r02s_trace_ab.log: incoherent closest 3.70 vs 3.58 ms, any hit
                         // 2.11 vs 2.13): the leaf phase waits for its loads, not for the binary64 pipe
#endif
#ifndef DRT_Q_STEPS
#define DRT_Q_STEPS 3  // pop + node steps per round of refill / leaf-phase checks (1 / 2 / 3: 4.01 / 3.63 / 3.56 ms)
#endif

#define DRT_EPS2 4.76837158203125e-07f  // 2^-21
#define DRT_REF_NONE ((int32_t)0x7ffffffe)

static __device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
static __device__ __forceinline__ void unpack2(unsigned long long v, float* lo, float* hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(*lo), "=f"(*hi) : "l"(v));
}
static __device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
static __device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
static __device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
static __device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
static __device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
static __device__ __forceinline__ float min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
static __device__ __forceinline__ void ldg256u(const void* p, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
static __device__ __forceinline__ void sts64(unsigned addr, int32_t ref, float t) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(ref), "r"(__float_as_uint(t)) : "memory");
}
static __device__ __forceinline__ uint2 lds64(unsigned addr) {
  uint2 e;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(addr) : "memory");
  return e;
}
static __device__ __forceinline__ void sts32(unsigned addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
static __device__ __forceinline__ uint32_t lds32(unsigned addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// Hot per-ray state (registers).
struct QRay {
  unsigned long long noxy2, ixy2;  // (-o.x, -o.y), (invDir.x, invDir.y): operands of the packed float32 instructions
  float noz, iz;
  float mintLo, maxtHi;  // float32 brackets of the f64 interval ends (mintLo <= mint, maxtHi >= maxt)
  double maxt;  // minDistance (f64) is parked in shared memory
  unsigned selXY;  // PRMT selectors that put the NEAR byte of each (lo, hi) pair first (0x3210 or 0x2301): x in bits 0-15, y in 16-31
  unsigned flags;  // bits 0-2: dirIsNeg (bvh_accel.dart:113-115); bit 3: slow ray; bits 8-12: 3 * octant; bits 16-31: the z selector
};

// The reference's slab test in f64 (bvh_accel.dart:439-472), arguments BY VALUE so the caller's ray state stays in
// registers.  Returns (ok << 32) | bits of float32(tmin).
static __device__ __noinline__ unsigned long long slabExactQ(float ox, float oy, float oz, float ix, float iy, float iz, double mint,
                                                             double maxt, float lox, float loy, float loz, float hix, float hiy,
                                                             float hiz) {
  RayState r;
  r.ox = ox; r.oy = oy; r.oz = oz;
  r.ix = ix; r.iy = iy; r.iz = iz;
  r.negx = ix < 0.f; r.negy = iy < 0.f; r.negz = iz < 0.f;
  double tmin, tmax;
  if (!slabs(r, lox, loy, loz, hix, hiy, hiz, &tmin, &tmax)) return 0ull;
  const bool ok = (tmin < maxt) && (tmax > mint);
  return ((unsigned long long)(ok ? 1u : 0u) << 32) | (unsigned long long)__float_as_uint(__double2float_rn(tmin));
}

// f32-keep-begin (gen_f32.py: the decode of a quantised box stays binary64 in the float32 unit too — a float32 decode could shrink it)
// One slot of a quantised node for a "slow" ray (zero / tiny / huge direction components, far-away origins; rare and
// warp-divergent): the reference's own binary64 decision on the DECODED box, which contains the true box, so that a box
// the reference enters is entered (for zero direction components: origin strictly inside the true slab -> strictly
// inside the decoded one; on a true boundary -> NaN or inside, handled like the reference handles its NaN).
static __device__ __noinline__ unsigned long long slowSlotQ(float ox, float oy, float oz, float ix, float iy, float iz, double mint,
                                                            double maxt, float Ox, float Oy, float Oz, float sx, float sy, float sz,
                                                            unsigned bx, unsigned by, unsigned bz) {
  // bx / by / bz: (lo, hi) bytes of this slot in bits 0-15; s* = 2^(e+15)
  const double k = 1.0 / 32768.0;
  const float lox = __double2float_rd((double)Ox + (double)(bx & 255u) * ((double)sx * k));
  const float hix = __double2float_ru((double)Ox + (double)((bx >> 8) & 255u) * ((double)sx * k));
  const float loy = __double2float_rd((double)Oy + (double)(by & 255u) * ((double)sy * k));
  const float hiy = __double2float_ru((double)Oy + (double)((by >> 8) & 255u) * ((double)sy * k));
  const float loz = __double2float_rd((double)Oz + (double)(bz & 255u) * ((double)sz * k));
  const float hiz = __double2float_ru((double)Oz + (double)((bz >> 8) & 255u) * ((double)sz * k));
  return slabExactQ(ox, oy, oz, ix, iy, iz, mint, maxt, lox, loy, loz, hix, hiy, hiz);
}

// f32-keep-end

// QUAD: the leaf code the scene needs (TraceScene::quadMode): 0 triangles only, 1 + spheres / disks, 2 + the remaining quadrics.
template <bool ANY, int QUAD>
__global__ void __launch_bounds__(DRT_Q_BLOCK, DRT_Q_MIN_BLOCKS)
    traceQKernel(TraceScene sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD, uint32_t n,
                 float4* __restrict__ hits, uint8_t* __restrict__ occluded, unsigned int* __restrict__ nextRay, TraceExtras ex) {
  const unsigned lane = threadIdx.x & 31u;
  if (ex.nDev) n = *ex.nDev;  // wavefront queues: the ray count lives in device memory
  const unsigned ltMask = (1u << lane) - 1u;
  uint32_t warpNext = 0, warpEnd = 0;  // warp-uniform: the chunk of rays this warp owns
  bool exhausted = false;              // warp-uniform: the global counter ran past n
  bool alive = false;
  QRay r;
  r.noxy2 = r.ixy2 = 0ull; r.noz = r.iz = 0.f; r.mintLo = r.maxtHi = 0.f; r.maxt = 0.0;
  r.selXY = 0x32103210u; r.flags = 0u;
  // Traversal stack as in trace_fast.cu: [entry][thread] in shared memory, deep entries in local memory.
  extern __shared__ uint2 smStack[];
  // Cold per-ray state is parked in shared memory behind the stack, in the stack's own [entry][thread] layout (touched at
  // set-up, in the leaf phase and at retirement only; addressed from the one base register like the stack entries):
  // words 0-2 ray.direction, 3 ray index, 4-5 minDistance (f64), 6 b1, 7 b2, 8 primitive id of the closest hit so far (-1: none)
#define COLD_ADDR(i) (smBase + (unsigned)(DRT_SMEM_STACK + ((i) >> 1)) * DRT_Q_STRIDE + (unsigned)((i) & 1) * 4u)
#define COLD_LD(i) lds32(COLD_ADDR(i))
#define COLD_ST(i, v) sts32(COLD_ADDR(i), (v))
  uint2 deepStack[104 - DRT_SMEM_STACK];
  unsigned smBase;
  asm volatile("{ .reg .u64 t64; cvta.to.shared.u64 t64, %1; cvt.u32.u64 %0, t64; }" : "=r"(smBase) : "l"(smStack));
  smBase += threadIdx.x * 8u;
  int sp = 0;
  int32_t cur = DRT_REF_NONE;
  int32_t pend0 = 0, pend1 = 0;  // postponed leaves, oldest first
  int npend = 0;
  bool found = false;  // any hit only; closest hit: COLD(5) >= 0

#define STACK_STORE(i, refv, tv)                                                                   \
  do {                                                                                             \
    if ((i) < DRT_SMEM_STACK) sts64(smBase + (unsigned)(i) * DRT_Q_STRIDE, (refv), (tv));                 \
    else deepStack[(i) - DRT_SMEM_STACK] = make_uint2((unsigned)(refv), __float_as_uint(tv));      \
  } while (0)
#define RETIRE()                                                                                             \
  do {                                                                                                       \
    alive = false;                                                                                           \
    const uint32_t rayIdx_ = COLD_LD(3);                                                                     \
    if (ANY) occluded[rayIdx_] = found ? 1 : 0;                                                              \
    else {                                                                                                   \
      const int hprim_ = (int)COLD_LD(8);                                                                    \
      hits[rayIdx_] = make_float4(hprim_ >= 0 ? __double2float_rn(r.maxt) : CUDART_INF_F, __uint_as_float(COLD_LD(6)), \
                                  __uint_as_float(COLD_LD(7)), __int_as_float(hprim_));                      \
      if (ex.tOut) ex.tOut[rayIdx_] = hprim_ >= 0 ? r.maxt : CUDART_INF;                                     \
    }                                                                                                        \
  } while (0)
#ifdef DRT_Q_PREFETCH_LEAF  // start the postponed leaf's first record towards L1: the leaf phase comes a few node steps later
#define PREFETCH_LEAF(refv) asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.prims + refLeafOffset(refv)))
#else
#define PREFETCH_LEAF(refv)
#endif
#define ENQUEUE_LEAF(refv)                        \
  do {                                            \
    if (npend == 0) pend0 = (refv);               \
    else pend1 = (refv);                          \
    ++npend;                                      \
    PREFETCH_LEAF(refv);                          \
  } while (0)

  for (;;) {
    // ---- refill idle lanes from the warp's chunk ----------------------------------------------
    unsigned dead = __ballot_sync(FULL_MASK, !alive);
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
          {
            const uint32_t linesPerArray = (RAY_CHUNK * 16 + 127) / 128;
            if (lane < 2 * linesPerArray) {
              const float4* basePtr = (lane < linesPerArray ? rayO : rayD) + base;
              const char* pf = reinterpret_cast<const char*>(basePtr) + 128 * (lane % linesPerArray);
              asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
            }
          }
        }
      }
      unsigned avail = (unsigned)(warpEnd - warpNext);
      unsigned nDead = __popc(dead);
      unsigned rank = __popc(dead & ltMask);
      if (!alive && rank < avail) {
        const uint32_t rayIdx = warpNext + rank;
        COLD_ST(3, rayIdx);
        float4 o = __ldg(rayO + rayIdx), d = __ldg(rayD + rayIdx);
        // invDir = (float)(1.0 / (double)d) (bvh_accel.dart:109-111) == the correctly rounded float32 quotient (trace_fast.cu)
        const float ix = __frcp_rn(d.x), iy = __frcp_rn(d.y), iz = __frcp_rn(d.z);
        COLD_ST(0, __float_as_uint(d.x)); COLD_ST(1, __float_as_uint(d.y)); COLD_ST(2, __float_as_uint(d.z));
        r.noxy2 = pack2(-o.x, -o.y); r.noz = -o.z;
        r.ixy2 = pack2(ix, iy); r.iz = iz;
        if (ex.range) {  // renderer rays: the reference's f64 minDistance / maxDistance (ray.dart:34-36)
          double2 mm = __ldg(ex.range + rayIdx);
          COLD_ST(4, (uint32_t)__double2loint(mm.x)); COLD_ST(5, (uint32_t)__double2hiint(mm.x));
          r.mintLo = __double2float_rd(mm.x);
          r.maxt = mm.y; r.maxtHi = __double2float_ru(mm.y);
        } else {
          const double m0 = (double)o.w;
          COLD_ST(4, (uint32_t)__double2loint(m0)); COLD_ST(5, (uint32_t)__double2hiint(m0));
          r.mintLo = o.w;
          r.maxt = d.w; r.maxtHi = d.w;
        }
        const float kOMax = 4.611686018427388e18f;                   // 2^62
        const float kIMin = 8.673617379884035e-19f, kIMax = 1.099511627776e12f;  // 2^-60, 2^40
        const bool slow = !(fabsf(o.x) <= kOMax) || !(fabsf(o.y) <= kOMax) || !(fabsf(o.z) <= kOMax) ||
                          !(fabsf(ix) <= kIMax) || !(fabsf(iy) <= kIMax) || !(fabsf(iz) <= kIMax) ||
                          !(fabsf(ix) >= kIMin) || !(fabsf(iy) >= kIMin) || !(fabsf(iz) >= kIMin);
        const unsigned neg = (ix < 0.f ? 1u : 0u) | (iy < 0.f ? 2u : 0u) | (iz < 0.f ? 4u : 0u);
        r.flags = neg | (slow ? 8u : 0u) | ((3u * neg) << 8) | ((neg & 4u) ? 0x23010000u : 0x32100000u);
        r.selXY = ((neg & 1u) ? 0x2301u : 0x3210u) | ((neg & 2u) ? 0x23010000u : 0x32100000u);
        sp = 0;
        npend = 0;
        found = false;
        COLD_ST(6, 0u); COLD_ST(7, 0u); COLD_ST(8, 0xffffffffu);
        alive = true;
        cur = DRT_REF_NONE;
        if (sc.empty) {
          RETIRE();
        } else if (sc.wideRootRef < 0) {
          ENQUEUE_LEAF(sc.wideRootRef);  // a root LEAF: its box is decided in the leaf phase like any other
        } else {
          cur = sc.wideRootRef;  // reference node 0's own box (bvh_accel.dart:123-125) is implied by its children's
        }
      }
      warpNext += nDead < avail ? nDead : avail;
      if (exhausted && __all_sync(FULL_MASK, !alive)) break;
    }

    // ---- exact leaf phase: run it when enough lanes cannot walk on, or nobody can ----------------
    {
      const bool blocked = alive && cur == DRT_REF_NONE && (npend == DRT_PEND || (sp == 0 && npend > 0));
      const unsigned bl = __ballot_sync(FULL_MASK, blocked);
      const unsigned wk = __ballot_sync(FULL_MASK, alive && !blocked);
      if (bl && (__popc(bl) >= DRT_Q_LEAF_BATCH || wk == 0)) {
        if (alive && npend > 0) {
          RayState rs;  // origin from the packed registers, direction from shared memory (only needed here)
          float nox, noy;
          unpack2(r.noxy2, &nox, &noy);
          float ixf, iyf;
          unpack2(r.ixy2, &ixf, &iyf);
          rs.ox = -nox; rs.oy = -noy; rs.oz = -r.noz;
          rs.dx = __uint_as_float(COLD_LD(0)); rs.dy = __uint_as_float(COLD_LD(1)); rs.dz = __uint_as_float(COLD_LD(2));
          rs.mint = __hiloint2double((int)COLD_LD(5), (int)COLD_LD(4)); rs.maxt = r.maxt;
          bool stop = false;
          for (int j = 0; j < npend && !stop; ++j) {
            const int32_t leaf = j == 0 ? pend0 : pend1;
            uint32_t off = refLeafOffset(leaf), cnt = refLeafCountField(leaf);
            const GPrim* pr = sc.prims + off;
            float4 a = ldg4(&pr[0].p1[0]), b = ldg4(&pr[0].p2[0]), c = ldg4(&pr[0].p3[0]);
            if (cnt == 15u) cnt = (uint32_t)__float_as_int(b.w);
            // the reference's own slab test of this leaf node (bvh_accel.dart:125 / 187), on the box rebuilt from its
            // primitives (triangle.dart:39-42 / the quadric's world bound), with the maxDistance of this moment
#define DRT_LEAF_BOX(okVar, maxtArg)                                                                                       \
            {                                                                                                              \
              float lo0 = CUDART_INF_F, lo1 = CUDART_INF_F, lo2 = CUDART_INF_F, hi0 = -CUDART_INF_F, hi1 = -CUDART_INF_F,  \
                    hi2 = -CUDART_INF_F;                                                                                   \
              for (uint32_t k2 = 0; k2 < cnt; ++k2) {                                                                      \
                const float4 a2 = ldg4(&pr[k2].p1[0]), b2 = ldg4(&pr[k2].p2[0]), c2 = ldg4(&pr[k2].p3[0]);                 \
                const int kind2 = __float_as_int(c2.w);                                                                    \
                if (QUAD == 0 || (kind2 & 1) == 0) {                                                                       \
                  lo0 = fminf(lo0, fminf(a2.x, fminf(b2.x, c2.x))); hi0 = fmaxf(hi0, fmaxf(a2.x, fmaxf(b2.x, c2.x)));      \
                  lo1 = fminf(lo1, fminf(a2.y, fminf(b2.y, c2.y))); hi1 = fmaxf(hi1, fmaxf(a2.y, fmaxf(b2.y, c2.y)));      \
                  lo2 = fminf(lo2, fminf(a2.z, fminf(b2.z, c2.z))); hi2 = fmaxf(hi2, fmaxf(a2.z, fmaxf(b2.z, c2.z)));      \
                } else {                                                                                                   \
                  const GSphere& s2 = sc.spheres[kind2 >> 1];                                                              \
                  lo0 = fminf(lo0, s2.wmin[0]); hi0 = fmaxf(hi0, s2.wmax[0]);                                              \
                  lo1 = fminf(lo1, s2.wmin[1]); hi1 = fmaxf(hi1, s2.wmax[1]);                                              \
                  lo2 = fminf(lo2, s2.wmin[2]); hi2 = fmaxf(hi2, s2.wmax[2]);                                              \
                }                                                                                                          \
              }                                                                                                            \
              okVar = (slabExactQ(-nox, -noy, -r.noz, ixf, iyf, r.iz, rs.mint, (maxtArg), lo0, lo1, lo2, hi0, hi1, hi2) >> 32) != 0ull; \
            }
            if (QUAD == 0 && DRT_Q_LAZY_BOX) {
              // Triangles only: the box test decides nothing unless a triangle of the leaf reports a hit (a miss changes no
              // state), so it is evaluated at the FIRST hit, with the maxDistance the ray had when the reference would have
              // reached the leaf — which is still rs.maxt's value then, no hit of this leaf having been taken before.  A
              // failing box drops the hit and the rest of the leaf, as if the reference had not entered it.
              const double maxtEntry = rs.maxt;
              bool boxKnown = false;
              for (uint32_t k = 0; k < cnt && !stop; ++k) {
                if (k) { a = ldg4(&pr[k].p1[0]); b = ldg4(&pr[k].p2[0]); c = ldg4(&pr[k].p3[0]); }
                HitState h;
                const bool hit = ANY ? triangleAny(rs, a, b, c) : triangleClosest(rs, a, b, c, &h);
                if (hit) {
                  if (!boxKnown) {
                    bool boxOk;
                    DRT_LEAF_BOX(boxOk, maxtEntry);
                    if (!boxOk) { rs.maxt = maxtEntry; break; }
                    boxKnown = true;
                  }
                  if (ANY) { found = true; stop = true; }
                  else {
                    COLD_ST(6, __float_as_uint(__double2float_rn(h.b1))); COLD_ST(7, __float_as_uint(__double2float_rn(h.b2)));
                    COLD_ST(8, (uint32_t)h.prim);
                  }
                }
              }
              continue;
            }
            bool boxOk;
#if DRT_Q_LEAF_PRE
            if (QUAD == 0 && cnt == 1u && !(r.flags & 8u)) {
              // one triangle (the usual leaf: ranges of <= 4 primitives are always split): its box in float32.  t' = fl(fl(b - o) *
              // invDir) = t (1 + e), |e| <= 2^-23, against the reference's binary64 t; DRT_EPS2 = 2^-21 covers it.  Every comparison
              // is written so that a NaN (0 * inf cannot occur here: slow rays skip this) never rejects.
              const float bx0 = fminf(a.x, fminf(b.x, c.x)), bx1 = fmaxf(a.x, fmaxf(b.x, c.x));
              const float by0 = fminf(a.y, fminf(b.y, c.y)), by1 = fmaxf(a.y, fmaxf(b.y, c.y));
              const float bz0 = fminf(a.z, fminf(b.z, c.z)), bz1 = fmaxf(a.z, fmaxf(b.z, c.z));
              const float tx0 = __fmul_rn(__fadd_rn(bx0, nox), ixf), tx1 = __fmul_rn(__fadd_rn(bx1, nox), ixf);
              const float ty0 = __fmul_rn(__fadd_rn(by0, noy), iyf), ty1 = __fmul_rn(__fadd_rn(by1, noy), iyf);
              const float tz0 = __fmul_rn(__fadd_rn(bz0, r.noz), r.iz), tz1 = __fmul_rn(__fadd_rn(bz1, r.noz), r.iz);
              const float N_ = max3(fminf(tx0, tx1), fminf(ty0, ty1), fminf(tz0, tz1));
              const float F_ = min3(fmaxf(tx0, tx1), fmaxf(ty0, ty1), fmaxf(tz0, tz1));
              // the absolute 1e-36 pays for products that fall into the float32 denormal range, where the relative margin vanishes
              const float Nlo_ = fmaf(-DRT_EPS2, fabsf(N_), N_) - 1.0e-36f, Fhi_ = fmaf(DRT_EPS2, fabsf(F_), F_) + 1.0e-36f;
              const float mh_ = j == 0 ? r.maxtHi : __double2float_ru(rs.maxt);  // the first leaf of the queue may have shortened the ray
              if (Nlo_ > Fhi_ || Nlo_ >= mh_ || Fhi_ < r.mintLo) continue;
            }
#endif
            DRT_LEAF_BOX(boxOk, rs.maxt);
            for (uint32_t k = 0; boxOk && k < cnt && !stop; ++k) {
              if (k) { a = ldg4(&pr[k].p1[0]); b = ldg4(&pr[k].p2[0]); c = ldg4(&pr[k].p3[0]); }
              const int kind = __float_as_int(c.w);
              if (QUAD == 0 || (kind & 1) == 0) {
                if (ANY) {
                  if (triangleAny(rs, a, b, c)) { found = true; stop = true; }
                } else {
                  HitState h;
                  if (triangleClosest(rs, a, b, c, &h)) {
                    COLD_ST(6, __float_as_uint(__double2float_rn(h.b1))); COLD_ST(7, __float_as_uint(__double2float_rn(h.b2)));
                    COLD_ST(8, (uint32_t)h.prim);
                  }
                }
              } else {
                const GSphere& s = sc.spheres[kind >> 1];
                double th, u, v;
                if (ANY) {
                  if (sphereTest<QUAD == 2>(s, rs, true, &th, nullptr, nullptr)) { found = true; stop = true; }
                } else if (sphereTest<QUAD == 2>(s, rs, false, &th, ex.noUV ? nullptr : &u, &v)) {
                  COLD_ST(6, ex.noUV ? 0u : __float_as_uint(__double2float_rn(u))); COLD_ST(7, ex.noUV ? 0u : __float_as_uint(__double2float_rn(v)));
                  COLD_ST(8, (uint32_t)__float_as_int(a.w));
                  rs.maxt = th;
                }
              }
            }
          }
#undef DRT_LEAF_BOX
          npend = 0;
          if (!ANY && rs.maxt != r.maxt) {
            r.maxt = rs.maxt;
            r.maxtHi = __double2float_ru(rs.maxt);
          }
          if (ANY && found) RETIRE();
        }
      }
    }

    // ---- next node for every lane without one: pop until an interior node comes up, the leaf queue is full, or the
    //      stack is empty (bvh_accel.dart:139-143,156-159; the pop-time cull is the reference's `tmin < maxDistance`
    //      on a lower bound of tmin) ------------------------------------------------------------------
#pragma unroll
    for (int step_ = 0; step_ < DRT_Q_STEPS; ++step_) {
    if (alive && cur == DRT_REF_NONE) {
#ifdef DRT_Q_SINGLE_POP
      for (int once_ = 0; once_ < 1; ++once_) {  // one entry per step: a culled or leaf entry costs the lane this step, not the warp a loop trip
#else
      for (;;) {
#endif
        if (npend == DRT_PEND || sp == 0) break;
        --sp;
        const uint2 e_ = sp < DRT_SMEM_STACK ? lds64(smBase + (unsigned)sp * DRT_Q_STRIDE) : deepStack[sp - DRT_SMEM_STACK];
        const int32_t ref_ = (int32_t)e_.x;
        if (!ANY && __uint_as_float(e_.y) >= r.maxtHi) continue;  // surely culled (NaN: never)
        if (ref_ < 0) { ENQUEUE_LEAF(ref_); continue; }
        cur = ref_;
        break;
      }
      if (cur == DRT_REF_NONE && sp == 0 && npend == 0) RETIRE();  // nothing left: outside the pop loop, one copy of the stores
    }

    // ---- one wide-node step for every lane that holds an interior node --------------------------
    if (alive && cur != DRT_REF_NONE) {
      const GNode4Q* nd = sc.wideQ + cur;
      uint32_t A[8], B[8];
      ldg256u(nd, A);
      ldg256u(reinterpret_cast<const uint32_t*>(nd) + 8, B);
      float t0, t1, t2, t3;
      int32_t r0, r1, r2, r3;  // the slot's reference when its box is entered, DRT_REF_EMPTY otherwise
      if (!(r.flags & 8u)) {
        // per node: S' and C'' (see the header)
        const float sx = __uint_as_float(A[3] << 16), sy = __uint_as_float(A[3] & 0xffff0000u), sz = __uint_as_float(A[4] << 16);
        const unsigned long long Sxy = mul2(pack2(sx, sy), r.ixy2);
        const float Sz = __fmul_rn(sz, r.iz);
        const unsigned long long Cxy =
            sub2(mul2(add2(pack2(__uint_as_float(A[0]), __uint_as_float(A[1])), r.noxy2), r.ixy2), Sxy);
        const float Cz = __fsub_rn(__fmul_rn(__fadd_rn(__uint_as_float(A[2]), r.noz), r.iz), Sz);
        const unsigned long long Sz2 = pack2(Sz, Sz), Cz2 = pack2(Cz, Cz);
        // near byte first in every (lo, hi) pair
        const unsigned selY = r.selXY >> 16, selZ = r.flags >> 16;
        const unsigned wx0 = __byte_perm(A[6], 0u, r.selXY), wx1 = __byte_perm(A[7], 0u, r.selXY);
        const unsigned wy0 = __byte_perm(B[0], 0u, selY), wy1 = __byte_perm(B[1], 0u, selY);
        const unsigned wz0 = __byte_perm(B[2], 0u, selZ), wz1 = __byte_perm(B[3], 0u, selZ);
        const float maxtHi = r.maxtHi, mintLo = r.mintLo;
#define Q_SLOT(WX, WY, WZ, SELN, SELF, REF, TOUT, ROUT)                                                                  \
  do {                                                                                                                   \
    const float nx_ = __uint_as_float(__byte_perm(WX, 0x3F800000u, SELN)), fx_ = __uint_as_float(__byte_perm(WX, 0x3F800000u, SELF)); \
    const float ny_ = __uint_as_float(__byte_perm(WY, 0x3F800000u, SELN)), fy_ = __uint_as_float(__byte_perm(WY, 0x3F800000u, SELF)); \
    const float nz_ = __uint_as_float(__byte_perm(WZ, 0x3F800000u, SELN)), fz_ = __uint_as_float(__byte_perm(WZ, 0x3F800000u, SELF)); \
    float tnx_, tny_, tfx_, tfy_, tnz_, tfz_;                                                                            \
    unpack2(fma2(pack2(nx_, ny_), Sxy, Cxy), &tnx_, &tny_);                                                              \
    unpack2(fma2(pack2(fx_, fy_), Sxy, Cxy), &tfx_, &tfy_);                                                              \
    unpack2(fma2(pack2(nz_, fz_), Sz2, Cz2), &tnz_, &tfz_);                                                              \
    const float N_ = max3(tnx_, tny_, tnz_), F_ = min3(tfx_, tfy_, tfz_);                                                \
    const float Nlo_ = fmaf(-DRT_EPS2, fabsf(N_), N_), Fhi_ = fmaf(DRT_EPS2, fabsf(F_), F_);                             \
    TOUT = Nlo_;                                                                                                         \
    ROUT = ((Nlo_ <= fminf(Fhi_, maxtHi)) && (Fhi_ >= mintLo)) ? (REF) : DRT_REF_EMPTY;                                  \
  } while (0)
        Q_SLOT(wx0, wy0, wz0, 0x7604, 0x7614, (int32_t)B[4], t0, r0);
        Q_SLOT(wx0, wy0, wz0, 0x7624, 0x7634, (int32_t)B[5], t1, r1);
        Q_SLOT(wx1, wy1, wz1, 0x7604, 0x7614, (int32_t)B[6], t2, r2);
        Q_SLOT(wx1, wy1, wz1, 0x7624, 0x7634, (int32_t)B[7], t3, r3);
#undef Q_SLOT
      } else {
        float nox, noy, ixf, iyf;
        unpack2(r.noxy2, &nox, &noy);
        unpack2(r.ixy2, &ixf, &iyf);
        const double mintS = __hiloint2double((int)COLD_LD(5), (int)COLD_LD(4));
        const float sx = __uint_as_float(A[3] << 16), sy = __uint_as_float(A[3] & 0xffff0000u), sz = __uint_as_float(A[4] << 16);
#define Q_SLOW(K, REF, TOUT, ROUT)                                                                                        \
  do {                                                                                                                    \
    TOUT = 0.f;                                                                                                           \
    ROUT = DRT_REF_EMPTY;                                                                                                 \
    if ((REF) != DRT_REF_EMPTY) {                                                                                         \
      const unsigned sh_ = ((K) & 1) * 16;                                                                                \
      const unsigned long long v_ = slowSlotQ(-nox, -noy, -r.noz, ixf, iyf, r.iz, mintS, r.maxt, __uint_as_float(A[0]),   \
                                              __uint_as_float(A[1]), __uint_as_float(A[2]), sx, sy, sz,                    \
                                              (A[6 + ((K) >> 1)] >> sh_) & 0xffffu, (B[0 + ((K) >> 1)] >> sh_) & 0xffffu,  \
                                              (B[2 + ((K) >> 1)] >> sh_) & 0xffffu);                                       \
      const float tm_ = __uint_as_float((unsigned)(v_ & 0xffffffffull));                                                  \
      TOUT = fmaf(-DRT_EPS2, fabsf(tm_), tm_);                                                                            \
      ROUT = (v_ >> 32) != 0ull ? (REF) : DRT_REF_EMPTY;                                                                  \
    }                                                                                                                     \
  } while (0)
        Q_SLOW(0, (int32_t)B[4], t0, r0);
        Q_SLOW(1, (int32_t)B[5], t1, r1);
        Q_SLOW(2, (int32_t)B[6], t2, r2);
        Q_SLOW(3, (int32_t)B[7], t3, r3);
#undef Q_SLOW
      }
      // visiting order of the reference's depth-first walk (bvh_accel.dart:147-153) from the node's per-octant decisions;
      // any hit: the answer does not depend on the order, the slots are walked as stored (larger boxes first)
      const unsigned dec = ANY ? 0u : (A[5] >> ((r.flags >> 8) & 31u));
      const bool sP = (dec & 1u) != 0, sA = (dec & 2u) != 0, sB = (dec & 4u) != 0;
      const int32_t a0 = sA ? r1 : r0, a1 = sA ? r0 : r1, b0 = sB ? r3 : r2, b1 = sB ? r2 : r3;
      const float ta0 = sA ? t1 : t0, ta1 = sA ? t0 : t1, tb0 = sB ? t3 : t2, tb1 = sB ? t2 : t3;
      const int32_t s0 = sP ? b0 : a0, s1 = sP ? b1 : a1, s2 = sP ? a0 : b0, s3 = sP ? a1 : b1;
      const float u1 = sP ? tb1 : ta1, u2 = sP ? ta0 : tb0, u3 = sP ? ta1 : tb1;
      const bool v0 = s0 != DRT_REF_EMPTY, v1 = s1 != DRT_REF_EMPTY, v2 = s2 != DRT_REF_EMPTY, v3 = s3 != DRT_REF_EMPTY;
      const int p3 = (v3 && (v0 || v1 || v2)) ? 1 : 0, p2 = (v2 && (v0 || v1)) ? 1 : 0, p1 = (v1 && v0) ? 1 : 0;
      if (sp <= DRT_SMEM_STACK - 3) {
        unsigned a = smBase + (unsigned)sp * DRT_Q_STRIDE;
        sts64(a, s3, u3); a += p3 ? DRT_Q_STRIDE : 0u;
        sts64(a, s2, u2); a += p2 ? DRT_Q_STRIDE : 0u;
        sts64(a, s1, u1);
        sp += p3 + p2 + p1;
      } else {
        STACK_STORE(sp, s3, u3); sp += p3;
        STACK_STORE(sp, s2, u2); sp += p2;
        STACK_STORE(sp, s1, u1); sp += p1;
      }
      cur = v0 ? s0 : (v1 ? s1 : (v2 ? s2 : (v3 ? s3 : DRT_REF_NONE)));
#ifdef DRT_Q_PREFETCH_TOP  // the entry that pops next (the second passing child): towards L1 while the first child's subtree is walked
      if (p1 | p2 | p3) {
        const int32_t nxt_ = p1 ? s1 : (p2 ? s2 : s3);
        if (nxt_ >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.wideQ + nxt_));
      }
#endif
      if (cur < 0) {  // a leaf: postpone it (the lane walked, so its queue had room)
        ENQUEUE_LEAF(cur);
        cur = DRT_REF_NONE;
      }
    }
    }  // DRT_Q_STEPS
  }
#undef RETIRE
#undef COLD_ADDR
#undef COLD_LD
#undef COLD_ST
#undef STACK_STORE
#undef ENQUEUE_LEAF
}

static cudaError_t launchOneQ(const TraceScene& sc, bool any, const float4* o, const float4* d, uint32_t n, bool nUnknown, void* out,
                              unsigned long long* nextRay, int numSMs, cudaStream_t stream, const TraceExtras& ex) {
  cudaError_t e = cudaMemsetAsync(nextRay, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  const int block = DRT_Q_BLOCK;
  const size_t smem = (size_t)(DRT_SMEM_STACK + 5) * block * sizeof(uint2);  // stack + 5 entries of parked ray state per thread
  typedef void (*KernelFn)(TraceScene, const float4*, const float4*, uint32_t, float4*, uint8_t*, unsigned int*, TraceExtras);
  static const KernelFn kKernels[6] = {traceQKernel<false, 0>, traceQKernel<false, 1>, traceQKernel<false, 2>,
                                       traceQKernel<true, 0>,  traceQKernel<true, 1>,  traceQKernel<true, 2>};
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
  const uint64_t warpsPerBlock = DRT_Q_BLOCK / 32;
  uint64_t blocksWanted = (want + warpsPerBlock - 1) / warpsPerBlock;
  uint64_t persistent = (uint64_t)numSMs * perSm[variant];  // one resident wave: persistent warps
  dim3 grid((unsigned)(blocksWanted < persistent ? blocksWanted : persistent));
  unsigned int* ctr = reinterpret_cast<unsigned int*>(nextRay);
  if (any) kernel<<<grid, block, smem, stream>>>(sc, o, d, n, nullptr, (uint8_t*)out, ctr, ex);
  else kernel<<<grid, block, smem, stream>>>(sc, o, d, n, (float4*)out, nullptr, ctr, ex);
  return cudaGetLastError();
}

cudaError_t launchTraceQ(const TraceScene& sc, bool any, const void* rayO, const void* rayD, uint64_t n, void* out,
                         unsigned long long* nextRay, int numSMs, cudaStream_t stream, const TraceExtras* extras) {
  TraceExtras ex{};
  if (extras) ex = *extras;
  const float4* o = static_cast<const float4*>(rayO);
  const float4* d = static_cast<const float4*>(rayD);
  if (ex.nDev) return launchOneQ(sc, any, o, d, 0, true, out, nextRay, numSMs, stream, ex);  // count lives on the device (< 2^31)
  const uint64_t kMax = 1ull << 30;  // rays per launch: 32-bit ray indices inside the kernel
  for (uint64_t first = 0; first < n; first += kMax) {
    const uint32_t m = (uint32_t)(n - first < kMax ? n - first : kMax);
    TraceExtras e2 = ex;
    if (e2.range) e2.range += first;
    if (e2.tOut) e2.tOut += first;
    void* o2 = any ? (void*)((uint8_t*)out + first) : (void*)((float4*)out + first);
    cudaError_t e = launchOneQ(sc, any, o + first, d + first, m, false, o2, nextRay, numSMs, stream, e2);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace drt
