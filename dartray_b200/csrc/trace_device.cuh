// Device helpers shared by the traversal kernels: the reference's slab test, triangle and sphere
// tests restated op for op (see trace_kernels.cu header for the arithmetic contract).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>

#include "anim_transform.h"
#include "gpu_types.h"

namespace drt {

#define DRT_STACK 64  // bvh_accel.dart:120 — the reference's todo stack has 64 entries

struct RayState {
  double ox, oy, oz;     // ray.origin (float32 values widened)
  double dx, dy, dz;     // ray.direction
  double ix, iy, iz;     // invDir: 1.0/d evaluated in f64, then ROUNDED TO FLOAT32 (bvh_accel.dart:109-111)
  double mint, maxt;     // ray.minDistance / ray.maxDistance (f64 in the reference, ray.dart:34-36)
  int negx, negy, negz;  // dirIsNeg (bvh_accel.dart:113-115)
};

// bvh_accel.dart:439-470 without the final ray-interval comparison: returns whether the three
// slabs overlap and the [tmin, tmax] the reference would compare against the ray interval.
static __device__ __forceinline__ bool slabs(const RayState& r, float lox, float loy, float loz, float hix, float hiy,
                                      float hiz, double* tminOut, double* tmaxOut) {
  double tmin = ((double)(r.negx ? hix : lox) - r.ox) * r.ix;
  double tmax = ((double)(r.negx ? lox : hix) - r.ox) * r.ix;
  double tymin = ((double)(r.negy ? hiy : loy) - r.oy) * r.iy;
  double tymax = ((double)(r.negy ? loy : hiy) - r.oy) * r.iy;
  if ((tmin > tymax) || (tymin > tmax)) return false;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  double tzmin = ((double)(r.negz ? hiz : loz) - r.oz) * r.iz;
  double tzmax = ((double)(r.negz ? loz : hiz) - r.oz) * r.iz;
  if ((tmin > tzmax) || (tzmin > tmax)) return false;
  if (tzmin > tmin) tmin = tzmin;
  if (tzmax < tmax) tmax = tzmax;
  *tminOut = tmin;
  *tmaxOut = tmax;
  return true;
}

struct HitState {
  double t, b1, b2;
  int prim;
  int inst;  // the TransformedPrimitive the hit came through (only the literal walk of trace_kernels.cu sets and reads it)
};

// triangle.dart:52-98, all f64 on the float32 vertices; updates r.maxt like
// geometric_primitive.dart:59.
static __device__ __forceinline__ bool triangleClosest(RayState& r, const float4 q0, const float4 q1, const float4 q2,
                                                HitState* h) {
  double p1x = q0.x, p1y = q0.y, p1z = q0.z;
  double e1x = (double)q1.x - p1x, e1y = (double)q1.y - p1y, e1z = (double)q1.z - p1z;
  double e2x = (double)q2.x - p1x, e2y = (double)q2.y - p1y, e2z = (double)q2.z - p1z;
  double s1x = (r.dy * e2z) - (r.dz * e2y);
  double s1y = (r.dz * e2x) - (r.dx * e2z);
  double s1z = (r.dx * e2y) - (r.dy * e2x);
  double divisor = (s1x * e1x) + (s1y * e1y) + (s1z * e1z);
  if (divisor == 0.0) return false;
  double invDivisor = 1.0 / divisor;
  double sx = r.ox - p1x, sy = r.oy - p1y, sz = r.oz - p1z;
  double b1 = (sx * s1x + sy * s1y + sz * s1z) * invDivisor;
  if (b1 < 0.0 || b1 > 1.0) return false;
  double s2x = (sy * e1z) - (sz * e1y);
  double s2y = (sz * e1x) - (sx * e1z);
  double s2z = (sx * e1y) - (sy * e1x);
  double b2 = ((r.dx * s2x) + (r.dy * s2y) + (r.dz * s2z)) * invDivisor;
  if (b2 < 0.0 || b1 + b2 > 1.0) return false;
  double t = (e2x * s2x + e2y * s2y + e2z * s2z) * invDivisor;
  if (t < r.mint || t > r.maxt) return false;
  h->t = t;
  h->b1 = b1;
  h->b2 = b2;
  h->prim = __float_as_int(q0.w);
  r.maxt = t;
  return true;
}

// Rounds an f64 expression to float32 and widens it again: what constructing a Dart
// Vector/Point (Float32List storage) does to each component.
static __device__ __forceinline__ double rf(double v) { return (double)__double2float_rn(v); }

// triangle.dart:162-194: e1, e2, s1, s, s2 are Vectors, i.e. float32-rounded.
static __device__ __forceinline__ bool triangleAny(const RayState& r, const float4 q0, const float4 q1, const float4 q2) {
  double p1x = q0.x, p1y = q0.y, p1z = q0.z;
  double e1x = rf((double)q1.x - p1x), e1y = rf((double)q1.y - p1y), e1z = rf((double)q1.z - p1z);
  double e2x = rf((double)q2.x - p1x), e2y = rf((double)q2.y - p1y), e2z = rf((double)q2.z - p1z);
  double s1x = rf((r.dy * e2z) - (r.dz * e2y));
  double s1y = rf((r.dz * e2x) - (r.dx * e2z));
  double s1z = rf((r.dx * e2y) - (r.dy * e2x));
  double divisor = s1x * e1x + s1y * e1y + s1z * e1z;
  if (divisor == 0.0) return false;
  double invDivisor = 1.0 / divisor;
  double sx = rf(r.ox - p1x), sy = rf(r.oy - p1y), sz = rf(r.oz - p1z);
  double b1 = (sx * s1x + sy * s1y + sz * s1z) * invDivisor;
  if (b1 < 0.0 || b1 > 1.0) return false;
  double s2x = rf((sy * e1z) - (sz * e1y));
  double s2y = rf((sz * e1x) - (sx * e1z));
  double s2z = rf((sx * e1y) - (sy * e1x));
  double b2 = (r.dx * s2x + r.dy * s2y + r.dz * s2z) * invDivisor;
  if (b2 < 0.0 || b1 + b2 > 1.0) return false;
  double t = (e2x * s2x + e2y * s2y + e2z * s2z) * invDivisor;
  if (t < r.mint || t > r.maxt) return false;
  return true;
}

// Cylinder / Cone / Paraboloid / Hyperboloid intersect == intersectP up to the hit decision (cylinder.dart:39-104 /
// :153-221, cone.dart:35-98 / :155-214, paraboloid.dart:37-100 / :158-217, hyperboloid.dart:57-122 / :178-244): one
// decision sequence, different coefficients, z range and phi.  Kept out of line: these shapes are rare next to
// triangles and must not grow the leaf phase of the traversal kernels.
static __device__ __forceinline__ double quadricPhi(const GSphere& s, double px, double py, double pz, double* vOut) {
  double ax = px, ay = py;
  if (s.shape == 5) {  // hyperboloid.dart:96-103: pr = p1 * (1 - v) + p2 * v in float32 Points
    double v = (pz - (double)s.hp1[2]) / ((double)s.hp2[2] - (double)s.hp1[2]);
    double prx = rf(rf((double)s.hp1[0] * (1.0 - v)) + rf((double)s.hp2[0] * v));
    double pry = rf(rf((double)s.hp1[1] * (1.0 - v)) + rf((double)s.hp2[1] * v));
    ay = prx * py - px * pry;
    ax = px * prx + py * pry;
    *vOut = v;
  }
  double phi = atan2(ay, ax);
  if (phi < 0.0) phi += 2.0 * 3.141592653589793;
  return phi;
}
static __device__ __noinline__ bool quadricTest(const GSphere& s, const RayState& r, double* thitOut, double* uOut, double* vOut) {
  const float* m = s.w2o;
  double ox = rf((double)m[0] * r.ox + (double)m[1] * r.oy + (double)m[2] * r.oz + (double)m[3]);
  double oy = rf((double)m[4] * r.ox + (double)m[5] * r.oy + (double)m[6] * r.oz + (double)m[7]);
  double oz = rf((double)m[8] * r.ox + (double)m[9] * r.oy + (double)m[10] * r.oz + (double)m[11]);
  double w = (double)s.w2oRow3[0] * r.ox + (double)s.w2oRow3[1] * r.oy + (double)s.w2oRow3[2] * r.oz + (double)s.w2oRow3[3];
  if (w != 1.0) { ox = rf(ox / w); oy = rf(oy / w); oz = rf(oz / w); }
  double dx = rf((double)m[0] * r.dx + (double)m[1] * r.dy + (double)m[2] * r.dz);
  double dy = rf((double)m[4] * r.dx + (double)m[5] * r.dy + (double)m[6] * r.dz);
  double dz = rf((double)m[8] * r.dx + (double)m[9] * r.dy + (double)m[10] * r.dz);
  double A, B, C, zlo = s.zmin, zhi = s.zmax;
  if (s.shape == 2) {
    A = dx * dx + dy * dy;
    B = 2.0 * (dx * ox + dy * oy);
    C = ox * ox + oy * oy - s.radius * s.radius;
  } else if (s.shape == 3) {
    double k = s.radius / s.height;
    k = k * k;
    A = dx * dx + dy * dy - k * dz * dz;
    B = 2.0 * (dx * ox + dy * oy - k * dz * (oz - s.height));
    C = ox * ox + oy * oy - k * (oz - s.height) * (oz - s.height);
    zlo = 0.0;
    zhi = s.height;
  } else if (s.shape == 4) {
    double k = s.zmax / (s.radius * s.radius);
    A = k * (dx * dx + dy * dy);
    B = 2 * k * (dx * ox + dy * oy) - dz;
    C = k * (ox * ox + oy * oy) - oz;
  } else {
    double a = s.ha, c = s.hc;
    A = a * dx * dx + a * dy * dy - c * dz * dz;
    B = 2.0 * (a * dx * ox + a * dy * oy - c * dz * oz);
    C = a * ox * ox + a * oy * oy - c * oz * oz - 1;
  }
  double discrim = B * B - 4.0 * A * C;  // common.dart:140-167
  if (discrim < 0.0) return false;
  double rootDiscrim = sqrt(discrim);
  double q = (B < 0.0) ? -0.5 * (B - rootDiscrim) : -0.5 * (B + rootDiscrim);
  double t0 = q / A, t1 = C / q;
  if (t0 > t1) { double tt = t0; t0 = t1; t1 = tt; }
  if (t0 > r.maxt || t1 < r.mint) return false;
  double thit = t0;
  if (t0 < r.mint) {
    thit = t1;
    if (thit > r.maxt) return false;
  }
  double px = rf(ox + rf(dx * thit)), py = rf(oy + rf(dy * thit)), pz = rf(oz + rf(dz * thit));
  double v = 0.0;
  double phi = quadricPhi(s, px, py, pz, &v);
  if (pz < zlo || pz > zhi || phi > s.phiMax) {
    if (thit == t1) return false;
    thit = t1;
    if (t1 > r.maxt) return false;
    px = rf(ox + rf(dx * thit)); py = rf(oy + rf(dy * thit)); pz = rf(oz + rf(dz * thit));
    phi = quadricPhi(s, px, py, pz, &v);
    if (pz < zlo || pz > zhi || phi > s.phiMax) return false;
  }
  *thitOut = thit;
  if (uOut) {
    if (s.shape == 2 || s.shape == 4) v = (pz - s.zmin) / (s.zmax - s.zmin);
    else if (s.shape == 3) v = pz / s.height;
    *uOut = phi / s.phiMax;
    *vOut = v;
  }
  return true;
}

// sphere.dart:39-116 / :169-241.  `shadow` selects intersectP, whose `thit == t1` comparison is
// between a double and a List and therefore never true (sphere.dart:210).
// GENERAL = false: the scene holds no cylinder / cone / paraboloid / hyperboloid, their out-of-line test is not even linked
// into the kernel (a call site in the leaf phase costs the traversal kernels ~40 % on triangle scenes, measured on B200).
template <bool GENERAL>
static __device__ bool sphereTest(const GSphere& s, const RayState& r, bool shadow, double* thitOut, double* uOut,
                           double* vOut) {
  // transform.dart:110-145,180-195: object-space origin/direction are float32 Points/Vectors
  if (GENERAL && s.shape >= 2) return quadricTest(s, r, thitOut, uOut, vOut);
  const float* m = s.w2o;
  if (s.shape == 1) {  // Disk.intersect / intersectP, lib/shapes/disk.dart:39-75 / :107-140 (same decisions)
    double ox = rf((double)m[0] * r.ox + (double)m[1] * r.oy + (double)m[2] * r.oz + (double)m[3]);
    double oy = rf((double)m[4] * r.ox + (double)m[5] * r.oy + (double)m[6] * r.oz + (double)m[7]);
    double oz = rf((double)m[8] * r.ox + (double)m[9] * r.oy + (double)m[10] * r.oz + (double)m[11]);
    double w = (double)s.w2oRow3[0] * r.ox + (double)s.w2oRow3[1] * r.oy + (double)s.w2oRow3[2] * r.oz + (double)s.w2oRow3[3];
    if (w != 1.0) { ox = rf(ox / w); oy = rf(oy / w); oz = rf(oz / w); }
    double dx = rf((double)m[0] * r.dx + (double)m[1] * r.dy + (double)m[2] * r.dz);
    double dy = rf((double)m[4] * r.dx + (double)m[5] * r.dy + (double)m[6] * r.dz);
    double dz = rf((double)m[8] * r.dx + (double)m[9] * r.dy + (double)m[10] * r.dz);
    if (fabs(dz) < 1.0e-7) return false;
    double thit = (s.height - oz) / dz;
    if (thit < r.mint || thit > r.maxt) return false;
    double px = rf(ox + rf(dx * thit)), py = rf(oy + rf(dy * thit));
    double dist2 = px * px + py * py;
    if (dist2 > s.radius * s.radius || dist2 < s.innerRadius * s.innerRadius) return false;
    // phi <= fl(2 pi) whatever the hit point (atan2 >= -pi; a negative value gets 2.0 * pi added): with phiMax at its full-circle
    // value the cut below never fires, and a caller that does not ask for (u, v) needs no atan2 at all
    double phi = 0.0;
    if (uOut || !(s.phiMax >= 6.283185307179586)) {
      phi = atan2(py, px);
      if (phi < 0) phi += 2.0 * 3.141592653589793;
    }
    if (phi > s.phiMax) return false;
    *thitOut = thit;
    if (uOut) {
      double oneMinusV = (sqrt(dist2) - s.innerRadius) / (s.radius - s.innerRadius);
      *uOut = phi / s.phiMax;
      *vOut = 1.0 - oneMinusV;
    }
    return true;
  }
  double ox = rf((double)m[0] * r.ox + (double)m[1] * r.oy + (double)m[2] * r.oz + (double)m[3]);
  double oy = rf((double)m[4] * r.ox + (double)m[5] * r.oy + (double)m[6] * r.oz + (double)m[7]);
  double oz = rf((double)m[8] * r.ox + (double)m[9] * r.oy + (double)m[10] * r.oz + (double)m[11]);
  double w = (double)s.w2oRow3[0] * r.ox + (double)s.w2oRow3[1] * r.oy + (double)s.w2oRow3[2] * r.oz +
             (double)s.w2oRow3[3];
  if (w != 1.0) { ox = rf(ox / w); oy = rf(oy / w); oz = rf(oz / w); }
  double dx = rf((double)m[0] * r.dx + (double)m[1] * r.dy + (double)m[2] * r.dz);
  double dy = rf((double)m[4] * r.dx + (double)m[5] * r.dy + (double)m[6] * r.dz);
  double dz = rf((double)m[8] * r.dx + (double)m[9] * r.dy + (double)m[10] * r.dz);
  double A = dx * dx + dy * dy + dz * dz;
  double B = 2 * (dx * ox + dy * oy + dz * oz);
  double C = ox * ox + oy * oy + oz * oz - s.radius * s.radius;
  // common.dart:140-167
  double discrim = B * B - 4.0 * A * C;
  if (discrim < 0.0) return false;
  double rootDiscrim = sqrt(discrim);
  double q = (B < 0.0) ? -0.5 * (B - rootDiscrim) : -0.5 * (B + rootDiscrim);
  double t0 = q / A, t1 = C / q;
  if (t0 > t1) { double tt = t0; t0 = t1; t1 = tt; }
  if (t0 > r.maxt || t1 < r.mint) return false;
  double thit = t0;
  if (thit < r.mint) {
    thit = t1;
    if (thit > r.maxt) return false;
  }
  // ray.pointAt: origin + (direction * t), two float32 roundings (ray.dart:70-71)
  double px = rf(ox + rf(dx * thit)), py = rf(oy + rf(dy * thit)), pz = rf(oz + rf(dz * thit));
  if (px == 0.0 && py == 0.0) px = rf(1.0e-5 * s.radius);
  const bool needPhi = uOut || !(s.phiMax >= 6.283185307179586);  // see the disk above
  double phi = 0.0;
  if (needPhi) {
    phi = atan2(py, px);
    if (phi < 0.0) phi += 2.0 * 3.141592653589793;
  }
  if ((s.zmin > -s.radius && pz < s.zmin) || (s.zmax < s.radius && pz > s.zmax) || phi > s.phiMax) {
    if (!shadow && thit == t1) return false;
    if (t1 > r.maxt) return false;
    thit = t1;
    px = rf(ox + rf(dx * thit)); py = rf(oy + rf(dy * thit)); pz = rf(oz + rf(dz * thit));
    if (px == 0.0 && py == 0.0) px = rf(1.0e-5 * s.radius);
    if (needPhi) {
      phi = atan2(py, px);
      if (phi < 0.0) phi += 2.0 * 3.141592653589793;
    }
    if ((s.zmin > -s.radius && pz < s.zmin) || (s.zmax < s.radius && pz > s.zmax) || phi > s.phiMax) return false;
  }
  *thitOut = thit;
  if (uOut) {
    double cz = pz / s.radius;
    cz = cz < -1.0 ? -1.0 : (cz > 1.0 ? 1.0 : cz);
    double theta = acos(cz);
    *uOut = phi / s.phiMax;
    *vOut = (theta - s.thetaMin) / (s.thetaMax - s.thetaMin);
  }
  return true;
}

static __device__ __forceinline__ void initRay(RayState& r, const float4 o, const float4 d) {
  r.ox = o.x; r.oy = o.y; r.oz = o.z;
  r.dx = d.x; r.dy = d.y; r.dz = d.z;
  r.mint = o.w;
  r.maxt = d.w;
  float ixf = __double2float_rn(1.0 / r.dx), iyf = __double2float_rn(1.0 / r.dy), izf = __double2float_rn(1.0 / r.dz);
  r.ix = ixf; r.iy = iyf; r.iz = izf;
  r.negx = ixf < 0.f; r.negy = iyf < 0.f; r.negz = izf < 0.f;
}

static __device__ __forceinline__ float4 ldg4(const void* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// BVHAccel.intersectP (bvh_accel.dart:167-226) for ONE ray inside another kernel: the literal walk of traceKernel<ANY = true>
// (trace_kernels.cu) over the binary nodes, f64 slab test at every node.  For stages that trace a few rays per thread in a
// data-dependent loop (the single-scattering volume integrator's shadow rays), where a wavefront queue per step would cost a
// launch per step; the batched kernels of trace_fast*.cu remain the production path.
static __device__ __noinline__ bool instanceTestCold(const TraceScene& sc, int inst, bool any, const RayState* rWorld, double time,
                                                     HitState* hit);
static __device__ __noinline__ bool anyHitWalk(const TraceScene& sc, float4 o4, float4 d4, double mint, double maxt, double time = 0.0) {
  if (sc.empty) return false;
  RayState r;
  initRay(r, o4, d4);
  r.mint = mint;
  r.maxt = maxt;
  int32_t stackRef[DRT_STACK];
  int sp = 0;
  int32_t cur = 0;
  {
    double tmin, tmax;
    if (!(slabs(r, sc.rootMin[0], sc.rootMin[1], sc.rootMin[2], sc.rootMax[0], sc.rootMax[1], sc.rootMax[2], &tmin, &tmax) &&
          (tmin < r.maxt) && (tmax > r.mint)))
      return false;
    cur = sc.rootRef;
  }
  for (;;) {
    if (cur >= 0) {
      const GNode* nd = sc.nodes + cur;
      float4 q0 = ldg4(&nd->c0min[0]), q1 = ldg4(&nd->c0max[1]), q2 = ldg4(&nd->c1min[2]);
      int4 q3 = __ldg(reinterpret_cast<const int4*>(&nd->ref0));
      const int neg = q3.z == 0 ? r.negx : (q3.z == 1 ? r.negy : r.negz);
      double tmin0, tmax0, tmin1, tmax1;
      const bool h0 = slabs(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tmin0, &tmax0) && (tmin0 < r.maxt) && (tmax0 > r.mint);
      const bool h1 = slabs(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tmin1, &tmax1) && (tmin1 < r.maxt) && (tmax1 > r.mint);
      const int32_t nearRef = neg ? q3.y : q3.x, farRef = neg ? q3.x : q3.y;
      const bool hn = neg ? h1 : h0, hf = neg ? h0 : h1;
      if (hf) stackRef[sp++] = farRef;  // maxDistance never shrinks in intersectP: the pop-time test cannot change
      if (hn) { cur = nearRef; continue; }
    } else {
      uint32_t off = refLeafOffset(cur), cnt = refLeafCountField(cur);
      const GPrim* pr = sc.prims + off;
      if (cnt == 15u) cnt = (uint32_t)__ldg(&pr->leafCount);
      for (uint32_t k = 0; k < cnt; ++k) {
        float4 a = ldg4(&pr[k].p1[0]), b = ldg4(&pr[k].p2[0]), c = ldg4(&pr[k].p3[0]);
        const int kind = __float_as_int(c.w);
        if ((kind & 1) == 0) {
          if (triangleAny(r, a, b, c)) return true;
        } else {
          double th;
          const GSphere& s = sc.spheres[kind >> 1];
          if (s.shape == 6) {  // a TransformedPrimitive at the ray's time
            HitState ih;
            if (instanceTestCold(sc, s.instance, true, &r, time, &ih)) return true;
          } else if (sphereTest<true>(s, r, true, &th, nullptr, nullptr)) {
            return true;
          }
        }
      }
    }
    if (sp == 0) return false;
    cur = stackRef[--sp];
  }
}

// ---- TransformedPrimitive (lib/core/primitive/transformed_primitive.dart:30-62) -------------------------------------------------
// Transform.transformRay (transform.dart:180-196): the origin as a Point (homogeneous divide when w != 1), the direction as a
// Vector, both float32; the interval is kept.  invDir / dirIsNeg as the nested BVHAccel.intersect derives them (bvh_accel.dart:109-115).
static __device__ inline void transformRayState(const M4& m, const RayState& r, RayState* o) {
  double x = rf((double)m.d[0] * r.ox + (double)m.d[1] * r.oy + (double)m.d[2] * r.oz + (double)m.d[3]);
  double y = rf((double)m.d[4] * r.ox + (double)m.d[5] * r.oy + (double)m.d[6] * r.oz + (double)m.d[7]);
  double z = rf((double)m.d[8] * r.ox + (double)m.d[9] * r.oy + (double)m.d[10] * r.oz + (double)m.d[11]);
  const double w = (double)m.d[12] * r.ox + (double)m.d[13] * r.oy + (double)m.d[14] * r.oz + (double)m.d[15];
  if (w != 1.0) { x = rf(x / w); y = rf(y / w); z = rf(z / w); }
  o->ox = x; o->oy = y; o->oz = z;
  o->dx = rf((double)m.d[0] * r.dx + (double)m.d[1] * r.dy + (double)m.d[2] * r.dz);
  o->dy = rf((double)m.d[4] * r.dx + (double)m.d[5] * r.dy + (double)m.d[6] * r.dz);
  o->dz = rf((double)m.d[8] * r.dx + (double)m.d[9] * r.dy + (double)m.d[10] * r.dz);
  o->mint = r.mint; o->maxt = r.maxt;
  const float ixf = __double2float_rn(1.0 / o->dx), iyf = __double2float_rn(1.0 / o->dy), izf = __double2float_rn(1.0 / o->dz);
  o->ix = ixf; o->iy = iyf; o->iz = izf;
  o->negx = ixf < 0.f; o->negy = iyf < 0.f; o->negz = izf < 0.f;
}

// The primitives of one leaf against a ray in the space the leaf lives in (triangles and quadrics: an object holds no instance)
template <bool ANY>
static __device__ inline bool objectLeaf(const TraceScene& sc, int32_t leaf, RayState& r, HitState* hit) {
  uint32_t off = refLeafOffset(leaf), cnt = refLeafCountField(leaf);
  const GPrim* pr = sc.prims + off;
  if (cnt == 15u) cnt = (uint32_t)__ldg(&pr->leafCount);
  bool found = false;
  for (uint32_t k = 0; k < cnt; ++k) {
    const float4 a = ldg4(&pr[k].p1[0]), b = ldg4(&pr[k].p2[0]), c = ldg4(&pr[k].p3[0]);
    const int kind = __float_as_int(c.w);
    if ((kind & 1) == 0) {
      if (ANY) { if (triangleAny(r, a, b, c)) return true; }
      else if (triangleClosest(r, a, b, c, hit)) found = true;
    } else {
      const GSphere& s = sc.spheres[kind >> 1];
      double th, u, v;
      if (ANY) { if (sphereTest<true>(s, r, true, &th, nullptr, nullptr)) return true; }
      else if (sphereTest<true>(s, r, false, &th, &u, &v)) {
        hit->t = th; hit->b1 = u; hit->b2 = v; hit->prim = __float_as_int(a.w);
        r.maxt = th;
        found = true;
      }
    }
  }
  return found;
}

// BVHAccel.intersect / intersectP of the object's own accelerator (bvh_accel.dart:101-226): the literal walk of traceKernel
template <bool ANY>
static __device__ bool objectWalk(const TraceScene& sc, const GObject& ob, RayState& r, HitState* hit) {
  if (ob.single) return objectLeaf<ANY>(sc, ob.rootRef, r, hit);  // one GeometricPrimitive: no accelerator, no box
  {
    double tmin, tmax;
    if (!(slabs(r, ob.rootMin[0], ob.rootMin[1], ob.rootMin[2], ob.rootMax[0], ob.rootMax[1], ob.rootMax[2], &tmin, &tmax) &&
          (tmin < r.maxt) && (tmax > r.mint)))
      return false;
  }
  int32_t stackRef[DRT_STACK];
  double stackT[DRT_STACK];
  int sp = 0;
  int32_t cur = ob.rootRef;
  bool found = false;
  for (;;) {
    if (cur >= 0) {
      const GNode* nd = sc.nodes + cur;
      const float4 q0 = ldg4(&nd->c0min[0]), q1 = ldg4(&nd->c0max[1]), q2 = ldg4(&nd->c1min[2]);
      const int4 q3 = __ldg(reinterpret_cast<const int4*>(&nd->ref0));
      const int neg = q3.z == 0 ? r.negx : (q3.z == 1 ? r.negy : r.negz);
      double tmin0, tmax0, tmin1, tmax1;
      const bool h0 = slabs(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tmin0, &tmax0) && (tmax0 > r.mint);
      const bool h1 = slabs(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tmin1, &tmax1) && (tmax1 > r.mint);
      const int32_t nearRef = neg ? q3.y : q3.x, farRef = neg ? q3.x : q3.y;
      const bool hn = neg ? h1 : h0, hf = neg ? h0 : h1;
      const double tn = neg ? tmin1 : tmin0, tf = neg ? tmin0 : tmin1;
      if (hf && tf < r.maxt) { stackRef[sp] = farRef; stackT[sp] = tf; sp++; }  // re-tested against maxDistance when popped
      if (hn && tn < r.maxt) { cur = nearRef; continue; }
    } else {
      if (objectLeaf<ANY>(sc, cur, r, hit)) {
        if (ANY) return true;
        found = true;
      }
    }
    bool have = false;
    while (sp > 0) {
      --sp;
      if (stackT[sp] < r.maxt) { cur = stackRef[sp]; have = true; break; }
    }
    if (!have) return found;
  }
}

// TransformedPrimitive.intersect / intersectP for the ray `r` (world space) at `time`: interpolate, transform the ray, query the
// object.  A closest hit leaves t / b1 / b2 / prim in *hit; the caller sets r.maxDistance = t (transformed_primitive.dart:39).
static __device__ __noinline__ bool instanceTestCold(const TraceScene& sc, int inst, bool any, const RayState* rWorld, double time,
                                                     HitState* hit) {
  const GInstance& in = sc.instances[inst];
  M4 m, inv;
  animInterpolate(in, time, &m, &inv);
  RayState r;
  transformRayState(m, *rWorld, &r);
  const GObject& ob = sc.objects[in.object];
  if (any) return objectWalk<true>(sc, ob, r, hit);
  return objectWalk<false>(sc, ob, r, hit);
}

}  // namespace drt
