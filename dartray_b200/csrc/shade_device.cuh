// Shading arithmetic of the render path, in the reference's numeric model: Vector/Point/Normal/
// RGBColor objects STORE float32 and every expression is evaluated in IEEE binary64, rounded to
// binary32 when an object is built (lib/core/vector.dart:26-74, rgb_color.dart:23-169,
// spectrum.dart:1145).  The whole library is compiled with -fmad=false, so nothing here contracts.
//
// Each function cites the reference code it replaces (paths relative to /root/reference).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cmath>
#include <cstdint>

#include "anim_transform.h"
#include "gpu_types.h"
#include "render_types.h"

namespace drt {

// DRT_EXTRA = 0 compiles the shading code WITHOUT per-vertex mesh attributes and the cylinder / cone / paraboloid /
// hyperboloid shapes: render_kernels.cu is built twice (plain / extra, see render_kernels_extra.cu) and the launchers pick
// by RenderScene::extra, so the kernels the benchmark scenes run carry no trace of the rare features.
#ifndef DRT_EXTRA
#define DRT_EXTRA 1
#endif
#define DRT_PI 3.141592653589793
#define DRT_INV_PI 0.31830988618379067154
#define DRT_INV_TWOPI 0.15915494309189533577
#define DRT_ONE_MINUS_EPS 0.9999999403953552  // montecarlo.dart:23

// ---- Vector / Point / Normal (vector.dart:26-218) ---------------------------------------------------
struct V3 {
  float x, y, z;
};
static DRT_HD inline V3 mkv(double x, double y, double z) { return V3{(float)x, (float)y, (float)z}; }
// ONE operation on two float32 operands, evaluated in binary64 and rounded to float32 — what `new Vector(a.x + b.x, ...)` does —
// equals the float32 operation itself: double rounding is innocuous for +, -, *, / when the wide format has >= 2 * 24 + 2 bits
// (Figueroa).  The float32 forms save two widening and one narrowing conversion per component (the XU pipe was the busiest pipe of
// shadePathKernel) — tried as an A/B build, see below.  Longer expressions (Dot, Cross, a * double) keep their binary64 evaluation.
#ifndef DRT_F32_SINGLE_OPS
#define DRT_F32_SINGLE_OPS 0  // measured on B200 (profiles/r02t_f32ops_ab.log): config 4 0.9715 s with the float32 forms, 0.9690 s
                             // without; cornell_materials 0.405 vs 0.399 s — no gain (results identical, 122 GPU tests green), off
#endif
#if DRT_F32_SINGLE_OPS
#ifdef __CUDA_ARCH__
#define DRT_FADD(a, b) __fadd_rn((a), (b))
#define DRT_FSUB(a, b) __fsub_rn((a), (b))
#define DRT_FMUL(a, b) __fmul_rn((a), (b))
#define DRT_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define DRT_FADD(a, b) ((a) + (b))
#define DRT_FSUB(a, b) ((a) - (b))
#define DRT_FMUL(a, b) ((a) * (b))
#define DRT_FDIV(a, b) ((a) / (b))
#endif
static DRT_HD inline V3 operator+(const V3& a, const V3& b) { return V3{DRT_FADD(a.x, b.x), DRT_FADD(a.y, b.y), DRT_FADD(a.z, b.z)}; }
static DRT_HD inline V3 operator-(const V3& a, const V3& b) { return V3{DRT_FSUB(a.x, b.x), DRT_FSUB(a.y, b.y), DRT_FSUB(a.z, b.z)}; }
#else
static DRT_HD inline V3 operator+(const V3& a, const V3& b) { return mkv((double)a.x + b.x, (double)a.y + b.y, (double)a.z + b.z); }
static DRT_HD inline V3 operator-(const V3& a, const V3& b) { return mkv((double)a.x - b.x, (double)a.y - b.y, (double)a.z - b.z); }
#endif
static DRT_HD inline V3 operator*(const V3& a, double f) { return mkv((double)a.x * f, (double)a.y * f, (double)a.z * f); }
static DRT_HD inline V3 operator/(const V3& a, double f) { return mkv((double)a.x / f, (double)a.y / f, (double)a.z / f); }
static DRT_HD inline V3 operator-(const V3& a) { return V3{-a.x, -a.y, -a.z}; }
static DRT_HD inline double Dot(const V3& a, const V3& b) { return (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z; }
static DRT_HD inline double AbsDot(const V3& a, const V3& b) { return fabs(Dot(a, b)); }
static DRT_HD inline V3 Cross(const V3& a, const V3& b) {
  double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
  return mkv((ay * bz) - (az * by), (az * bx) - (ax * bz), (ax * by) - (ay * bx));
}
static DRT_HD inline double LengthSquared(const V3& v) { return (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z; }
static DRT_HD inline double Length(const V3& v) { return sqrt(LengthSquared(v)); }
static DRT_HD inline V3 Normalize(const V3& v) { return v / Length(v); }
static DRT_HD inline double DistanceSquared(const V3& a, const V3& b) { return LengthSquared(b - a); }
static DRT_HD inline double Distance(const V3& a, const V3& b) { return Length(b - a); }
static DRT_HD inline void CoordinateSystem(const V3& v1, V3* v2, V3* v3) {  // vector.dart:198-214
  if (fabs((double)v1.x) > fabs((double)v1.y)) {
    double invLen = 1.0 / sqrt((double)v1.x * v1.x + (double)v1.z * v1.z);
    *v2 = mkv(-(double)v1.z * invLen, 0.0, (double)v1.x * invLen);
  } else {
    double invLen = 1.0 / sqrt((double)v1.y * v1.y + (double)v1.z * v1.z);
    *v2 = mkv(0.0, (double)v1.z * invLen, -(double)v1.y * invLen);
  }
  *v3 = Cross(v1, *v2);
}
static DRT_HD inline V3 FaceForward(const V3& n, const V3& v) { return (Dot(n, v) < 0.0) ? -n : n; }
static DRT_HD inline double clampD(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static DRT_HD inline double LerpD(double t, double a, double b) { return (1.0 - t) * a + t * b; }  // common.dart:80-81

// ray.dart:70-71: origin + (direction * t), a new float32 object at each step
static DRT_HD inline V3 RayAt(const V3& o, const V3& d, double t) { return o + (d * t); }

// transform.dart:110-161 on a float32 row-major 4x4
static DRT_HD inline V3 XfPoint(const float* m, const V3& p) {
  double x = p.x, y = p.y, z = p.z;
  V3 out = mkv((double)m[0] * x + (double)m[1] * y + (double)m[2] * z + (double)m[3],
               (double)m[4] * x + (double)m[5] * y + (double)m[6] * z + (double)m[7],
               (double)m[8] * x + (double)m[9] * y + (double)m[10] * z + (double)m[11]);
  double w = (double)m[12] * x + (double)m[13] * y + (double)m[14] * z + (double)m[15];
  if (w != 1.0) out = mkv((double)out.x / w, (double)out.y / w, (double)out.z / w);
  return out;
}
static DRT_HD inline V3 XfVector(const float* m, const V3& p) {
  double x = p.x, y = p.y, z = p.z;
  return mkv((double)m[0] * x + (double)m[1] * y + (double)m[2] * z, (double)m[4] * x + (double)m[5] * y + (double)m[6] * z,
             (double)m[8] * x + (double)m[9] * y + (double)m[10] * z);
}
// normal: transpose of the inverse (transform.dart:147-161); `mInv` is the inverse matrix
static DRT_HD inline V3 XfNormal(const float* mInv, const V3& p) {
  double x = p.x, y = p.y, z = p.z;
  return mkv((double)mInv[0] * x + (double)mInv[4] * y + (double)mInv[8] * z,
             (double)mInv[1] * x + (double)mInv[5] * y + (double)mInv[9] * z,
             (double)mInv[2] * x + (double)mInv[6] * y + (double)mInv[10] * z);
}

// ---- RGBColor (rgb_color.dart:23-169) ---------------------------------------------------------------
struct Spec {
  float r, g, b;
};
static DRT_HD inline Spec mks(double r, double g, double b) { return Spec{(float)r, (float)g, (float)b}; }
static DRT_HD inline Spec mks1(double v) { return Spec{(float)v, (float)v, (float)v}; }
#if DRT_F32_SINGLE_OPS
static DRT_HD inline Spec operator+(const Spec& a, const Spec& b) { return Spec{DRT_FADD(a.r, b.r), DRT_FADD(a.g, b.g), DRT_FADD(a.b, b.b)}; }
static DRT_HD inline Spec operator*(const Spec& a, const Spec& b) { return Spec{DRT_FMUL(a.r, b.r), DRT_FMUL(a.g, b.g), DRT_FMUL(a.b, b.b)}; }
#else
static DRT_HD inline Spec operator+(const Spec& a, const Spec& b) { return mks((double)a.r + b.r, (double)a.g + b.g, (double)a.b + b.b); }
static DRT_HD inline Spec operator*(const Spec& a, const Spec& b) { return mks((double)a.r * b.r, (double)a.g * b.g, (double)a.b * b.b); }
#endif
static DRT_HD inline Spec operator*(const Spec& a, double s) { return mks((double)a.r * s, (double)a.g * s, (double)a.b * s); }
static DRT_HD inline Spec operator/(const Spec& a, double s) { return mks((double)a.r / s, (double)a.g / s, (double)a.b / s); }
static DRT_HD inline bool IsBlack(const Spec& s) { return !(s.r != 0.f || s.g != 0.f || s.b != 0.f); }
static DRT_HD inline double Luminance(const Spec& s) { return 0.212671 * s.r + 0.715160 * s.g + 0.072169 * s.b; }

// ---- counter-based random streams (the "keyed" layout: one stream per (pixel, array) for the
// sampler and per (pixel, sample) for the integrators; draw order inside a stream is the
// reference's, lib/core/rng.dart:27-43) ---------------------------------------------------------------
static DRT_HD inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static DRT_HD inline uint64_t streamKey(uint64_t seed, int32_t x, int32_t y, uint32_t sampleIdx, uint32_t streamId) {
  // the seed is hashed before it meets the pixel, so that seeds s and s^1 do not merely swap neighbouring pixels' streams
  uint64_t k1 = mix64(mix64(seed + 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)y << 32)));
  return mix64(k1 ^ ((uint64_t)sampleIdx | ((uint64_t)streamId << 32)) ^ 0xD1B54A32D192ED03ull);
}
#define DRT_STREAM_PIXEL 4095u
#define DRT_STREAM_INTEGRATOR 0x80000000u
#define DRT_STREAM_TRANSMITTANCE 0x80000001u  // the draws of VolumeIntegrator.transmittance inside the surface integrator, in call order
#define DRT_STREAM_VOLUME_LI 0x80000002u      // the draws of VolumeIntegrator.Li along the camera ray
// d-th draw of a stream, d = 1, 2, ...
static DRT_HD inline uint64_t draw64(uint64_t key, uint64_t d) { return mix64(key + d * 0x9E3779B97F4A7C15ull); }
#if DRT_REAL32  // the float32 build (gen_f32.py): 24 bits, so that the draw stays below 1
static DRT_HD inline double drawFloat(uint64_t key, uint64_t d) { return (double)(uint32_t)(draw64(key, d) >> 40) * 5.9604644775390625e-8; }
#else
static DRT_HD inline double drawFloat(uint64_t key, uint64_t d) { return (double)(draw64(key, d) >> 11) * (1.0 / 9007199254740992.0); }
#endif
static DRT_HD inline uint32_t drawUint(uint64_t key, uint64_t d) {  // (draw >> 32) % 0xffffffff: the value itself unless it is 2^32 - 1
  const uint32_t x = (uint32_t)(draw64(key, d) >> 32);
  return x == 0xffffffffu ? 0u : x;
}

struct Stream {
  uint64_t key;
  uint64_t ctr;
  DRT_HD double randomFloat() { return drawFloat(key, ++ctr); }
  DRT_HD uint32_t randomUint() { return drawUint(key, ++ctr); }
};

// ---- montecarlo.dart ----------------------------------------------------------------------------------
static DRT_HD inline double Sobol2(uint32_t n, uint32_t scramble) {  // :486-493
  for (uint32_t v = 1u << 31; n != 0; n >>= 1, v ^= v >> 1)
    if (n & 0x1) scramble ^= v;
  return fmin(((scramble >> 8) & 0xffffff) / (double)(1 << 24), DRT_ONE_MINUS_EPS);
}
static DRT_HD inline double VanDerCorput(uint32_t n, uint32_t scramble) {  // :495-504
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
  n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
  n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
  n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
  n ^= scramble;
  return fmin(((n >> 8) & 0xffffff) / (double)(1 << 24), DRT_ONE_MINUS_EPS);
}
static DRT_HD inline V3 UniformSampleSphere(double u1, double u2) {  // :113-120
  double z = 1.0 - 2.0 * u1;
  double r = sqrt(fmax(0.0, 1.0 - z * z));
  double phi = 2.0 * DRT_PI * u2;
  return mkv(r * cos(phi), r * sin(phi), z);
}
static DRT_HD inline void ConcentricSampleDisk(double u1, double u2, double* dx, double* dy) {  // :155-201
  double r, theta;
  double sx = 2 * u1 - 1, sy = 2 * u2 - 1;
  if (sx == 0.0 && sy == 0.0) { *dx = 0.0; *dy = 0.0; return; }
  if (sx >= -sy) {
    if (sx > sy) { r = sx; theta = (sy > 0.0) ? sy / r : 8.0 + sy / r; }
    else { r = sy; theta = 2.0 - sx / r; }
  } else {
    if (sx <= sy) { r = -sx; theta = 4.0 - sy / r; }
    else { r = -sy; theta = 6.0 + sx / r; }
  }
  theta *= DRT_PI / 4.0;
  *dx = r * cos(theta);
  *dy = r * sin(theta);
}
static DRT_HD inline V3 CosineSampleHemisphere(double u1, double u2) {  // :203-209
  double dx, dy;
  ConcentricSampleDisk(u1, u2, &dx, &dy);
  double z = sqrt(fmax(0.0, 1.0 - dx * dx - dy * dy));
  return mkv(dx, dy, z);
}
static DRT_HD inline double PowerHeuristic(int nf, double fPdf, int ng, double gPdf) {  // :480-484
  double f = nf * fPdf, g = ng * gPdf;
  return (f * f) / (f * f + g * g);
}
static DRT_HD inline V3 UniformSampleCone2(double u1, double u2, double costhetamax, const V3& x, const V3& y, const V3& z) {
  double costheta = LerpD(u1, costhetamax, 1.0);  // :135-142
  double sintheta = sqrt(1.0 - costheta * costheta);
  double phi = u2 * 2.0 * DRT_PI;
  return x * (cos(phi) * sintheta) + y * (sin(phi) * sintheta) + z * costheta;
}
static DRT_HD inline double UniformConePdf(double cosThetaMax) { return 1.0 / (2.0 * DRT_PI * (1.0 - cosThetaMax)); }

// ---- shapes as the shading code sees them ---------------------------------------------------------------
struct ShapeHit {  // what Shape.intersect leaves behind: tHit, rayEpsilon and the part of dg the matte path reads
  double t, rayEps;
  V3 p, nn, dpdu;
};

struct TriVerts {
  V3 p1, p2, p3;
};

static __device__ inline TriVerts loadTri(const RenderScene& rs, uint32_t prim) {
  const GPrim* g = rs.ts.prims + rs.primToRec[prim];
  float4 a = __ldg(reinterpret_cast<const float4*>(&g->p1[0])), b = __ldg(reinterpret_cast<const float4*>(&g->p2[0])),
         c = __ldg(reinterpret_cast<const float4*>(&g->p3[0]));
  TriVerts t;
  t.p1 = V3{a.x, a.y, a.z}; t.p2 = V3{b.x, b.y, b.z}; t.p3 = V3{c.x, c.y, c.z};
  return t;
}

// triangle.dart:100-154 with the default uvs (0,0),(1,0),(1,1) (:255-262): du1=-1, du2=0, dv1=-1,
// dv2=-1, determinant 1, so dpdu = (dp1*dv2 - dp2*dv1) * 1 and dpdv = (dp1*-du2 + dp2*du1) * 1.
static DRT_HD inline void triPartials(const TriVerts& t, V3* dpdu, V3* dpdv) {
  V3 dp1 = t.p1 - t.p3, dp2 = t.p2 - t.p3;
  *dpdu = ((dp1 * -1.0) - (dp2 * -1.0)) * 1.0;
  *dpdv = ((dp1 * -0.0) + (dp2 * -1.0)) * 1.0;
}
// triangle.dart:104-131 with the mesh's own uvs (getUVs, :246-254)
static DRT_HD inline void triPartialsUV(const TriVerts& t, const double uv[6], V3* dpdu, V3* dpdv) {
  double du1 = uv[0] - uv[4], du2 = uv[2] - uv[4], dv1 = uv[1] - uv[5], dv2 = uv[3] - uv[5];
  V3 dp1 = t.p1 - t.p3, dp2 = t.p2 - t.p3;
  double determinant = du1 * dv2 - dv1 * du2;
  if (determinant == 0.0) {
    double e1x = (double)t.p2.x - t.p1.x, e1y = (double)t.p2.y - t.p1.y, e1z = (double)t.p2.z - t.p1.z;
    double e2x = (double)t.p3.x - t.p1.x, e2y = (double)t.p3.y - t.p1.y, e2z = (double)t.p3.z - t.p1.z;
    double e3x = (e2y * e1z) - (e2z * e1y), e3y = (e2z * e1x) - (e2x * e1z), e3z = (e2x * e1y) - (e2y * e1x);
    double len = sqrt(e3x * e3x + e3y * e3y + e3z * e3z);
    CoordinateSystem(mkv(e3x / len, e3y / len, e3z / len), dpdu, dpdv);
  } else {
    double invdet = 1.0 / determinant;
    *dpdu = ((dp1 * dv2) - (dp2 * dv1)) * invdet;
    *dpdv = ((dp1 * -du2) + (dp2 * du1)) * invdet;
  }
}
static DRT_HD inline V3 shapeNormal(const V3& dpdu, const V3& dpdv, bool reverse) {  // differential_geometry.dart:77-99
  V3 nn = Normalize(Cross(dpdu, dpdv));
  if (reverse) nn = nn * -1.0;
  return nn;
}

// triangle.dart:44-98, f64 throughout
static DRT_HD inline bool triIntersectT(const TriVerts& tv, const V3& o, const V3& d, double mint, double maxt, double* tOut) {
  double p1x = tv.p1.x, p1y = tv.p1.y, p1z = tv.p1.z;
  double e1x = (double)tv.p2.x - p1x, e1y = (double)tv.p2.y - p1y, e1z = (double)tv.p2.z - p1z;
  double e2x = (double)tv.p3.x - p1x, e2y = (double)tv.p3.y - p1y, e2z = (double)tv.p3.z - p1z;
  double dx = d.x, dy = d.y, dz = d.z;
  double s1x = (dy * e2z) - (dz * e2y);
  double s1y = (dz * e2x) - (dx * e2z);
  double s1z = (dx * e2y) - (dy * e2x);
  double divisor = (s1x * e1x) + (s1y * e1y) + (s1z * e1z);
  if (divisor == 0.0) return false;
  double invDivisor = 1.0 / divisor;
  double sx = (double)o.x - p1x, sy = (double)o.y - p1y, sz = (double)o.z - p1z;
  double b1 = (sx * s1x + sy * s1y + sz * s1z) * invDivisor;
  if (b1 < 0.0 || b1 > 1.0) return false;
  double s2x = (sy * e1z) - (sz * e1y);
  double s2y = (sz * e1x) - (sx * e1z);
  double s2z = (sx * e1y) - (sy * e1x);
  double b2 = ((dx * s2x) + (dy * s2y) + (dz * s2z)) * invDivisor;
  if (b2 < 0.0 || b1 + b2 > 1.0) return false;
  double t = (e2x * s2x + e2y * s2y + e2z * s2z) * invDivisor;
  if (t < mint || t > maxt) return false;
  *tOut = t;
  return true;
}

// sphere.dart:39-116: tHit and the object-space hit point (after the 1e-5*r nudge)
static DRT_HD inline bool sphereIntersectT(const GSphere& s, const V3& o, const V3& d, double mint, double maxt, double* tOut,
                                           V3* phitOut) {
  float w2o[16];
  for (int i = 0; i < 12; ++i) w2o[i] = s.w2o[i];
  for (int i = 0; i < 4; ++i) w2o[12 + i] = s.w2oRow3[i];
  V3 ro = XfPoint(w2o, o), rd = XfVector(w2o, d);
  double dx = rd.x, dy = rd.y, dz = rd.z, ox = ro.x, oy = ro.y, oz = ro.z;
  double A = dx * dx + dy * dy + dz * dz;
  double B = 2 * (dx * ox + dy * oy + dz * oz);
  double C = ox * ox + oy * oy + oz * oz - s.radius * s.radius;
  double discrim = B * B - 4.0 * A * C;  // common.dart:140-167
  if (discrim < 0.0) return false;
  double rootDiscrim = sqrt(discrim);
  double q = (B < 0.0) ? -0.5 * (B - rootDiscrim) : -0.5 * (B + rootDiscrim);
  double t0 = q / A, t1 = C / q;
  if (t0 > t1) { double tt = t0; t0 = t1; t1 = tt; }
  if (t0 > maxt || t1 < mint) return false;
  double thit = t0;
  if (thit < mint) {
    thit = t1;
    if (thit > maxt) return false;
  }
  V3 phit = RayAt(ro, rd, thit);
  if (phit.x == 0.0f && phit.y == 0.0f) phit.x = (float)(1.0e-5 * s.radius);
  double phi = atan2((double)phit.y, (double)phit.x);
  if (phi < 0.0) phi += 2.0 * DRT_PI;
  if ((s.zmin > -s.radius && phit.z < s.zmin) || (s.zmax < s.radius && phit.z > s.zmax) || phi > s.phiMax) {
    if (thit == t1) return false;
    if (t1 > maxt) return false;
    thit = t1;
    phit = RayAt(ro, rd, thit);
    if (phit.x == 0.0f && phit.y == 0.0f) phit.x = (float)(1.0e-5 * s.radius);
    phi = atan2((double)phit.y, (double)phit.x);
    if (phi < 0.0) phi += 2.0 * DRT_PI;
    if ((s.zmin > -s.radius && phit.z < s.zmin) || (s.zmax < s.radius && phit.z > s.zmax) || phi > s.phiMax) return false;
  }
  *tOut = thit;
  *phitOut = phit;
  return true;
}

// cylinder.dart:39-104, cone.dart:35-98, paraboloid.dart:37-100, hyperboloid.dart:57-122: tHit, the object-space hit
// point and phi (see quadricTest in trace_device.cuh for the decision sequence the four files share)
static DRT_HD inline double quadricPhiV(const GSphere& s, const V3& phit) {
  double ay = phit.y, ax = phit.x;
  if (s.shape == 5) {  // hyperboloid.dart:96-103
    V3 hp1 = V3{s.hp1[0], s.hp1[1], s.hp1[2]}, hp2 = V3{s.hp2[0], s.hp2[1], s.hp2[2]};
    double v = ((double)phit.z - hp1.z) / ((double)hp2.z - hp1.z);
    V3 pr = (hp1 * (1.0 - v)) + (hp2 * v);
    ay = (double)pr.x * phit.y - (double)phit.x * pr.y;
    ax = (double)phit.x * pr.x + (double)phit.y * pr.y;
  }
  double phi = atan2(ay, ax);
  if (phi < 0.0) phi += 2.0 * DRT_PI;
  return phi;
}
static DRT_HD inline bool quadricIntersectT(const GSphere& s, const V3& o, const V3& d, double mint, double maxt, double* tOut,
                                            V3* phitOut) {
  float w2o[16];
  for (int i = 0; i < 12; ++i) w2o[i] = s.w2o[i];
  for (int i = 0; i < 4; ++i) w2o[12 + i] = s.w2oRow3[i];
  V3 ro = XfPoint(w2o, o), rd = XfVector(w2o, d);
  double dx = rd.x, dy = rd.y, dz = rd.z, ox = ro.x, oy = ro.y, oz = ro.z;
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
  if (t0 > maxt || t1 < mint) return false;
  double thit = t0;
  if (t0 < mint) {
    thit = t1;
    if (thit > maxt) return false;
  }
  V3 phit = RayAt(ro, rd, thit);
  double phi = quadricPhiV(s, phit);
  if (phit.z < zlo || phit.z > zhi || phi > s.phiMax) {
    if (thit == t1) return false;
    thit = t1;
    if (t1 > maxt) return false;
    phit = RayAt(ro, rd, thit);
    phi = quadricPhiV(s, phit);
    if (phit.z < zlo || phit.z > zhi || phi > s.phiMax) return false;
  }
  *tOut = thit;
  *phitOut = phit;
  return true;
}

// cylinder.dart:111-112, cone.dart:101-107, paraboloid.dart:107-109, hyperboloid.dart:125-133: p, dpdu, dpdv in world space
static DRT_HD inline void quadricPartials(const GSphere& s, const V3& phit, V3* p, V3* dpdu, V3* dpdv) {
  float o2w[16];
  for (int i = 0; i < 12; ++i) o2w[i] = s.o2w[i];
  for (int i = 0; i < 4; ++i) o2w[12 + i] = s.o2wRow3[i];
  V3 du = mkv(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);
  V3 dv;
  if (s.shape == 2) {
    dv = mkv(0.0, 0.0, s.zmax - s.zmin);
  } else if (s.shape == 3) {
    double v = (double)phit.z / s.height;
    dv = mkv(-(double)phit.x / (1.0 - v), -(double)phit.y / (1.0 - v), s.height);
  } else if (s.shape == 4) {
    dv = mkv((double)phit.x / (2.0 * phit.z), (double)phit.y / (2.0 * phit.z), 1.0) * (s.zmax - s.zmin);
  } else {
    double phi = quadricPhiV(s, phit);
    double cosphi = cos(phi), sinphi = sin(phi);
    double ex = (double)s.hp2[0] - (double)s.hp1[0], ey = (double)s.hp2[1] - (double)s.hp1[1];
    dv = mkv(ex * cosphi - ey * sinphi, ex * sinphi + ey * cosphi, (double)s.hp2[2] - (double)s.hp1[2]);
  }
  *p = XfPoint(o2w, phit);
  *dpdu = XfVector(o2w, du);
  *dpdv = XfVector(o2w, dv);
}

#ifdef __CUDACC__
// Out-of-line entry points: the rare quadrics must not grow the shading kernels (instruction-cache bound).
static __device__ __noinline__ void quadricPartialsCold(const GSphere& s, V3 phit, V3* p, V3* dpdu, V3* dpdv) {
  quadricPartials(s, phit, p, dpdu, dpdv);
}
static __device__ __noinline__ bool quadricIntersectCold(const GSphere& s, V3 o, V3 d, double mint, double maxt, double* tOut, V3* p,
                                                         V3* dpdu, V3* dpdv) {
  V3 phit;
  if (!quadricIntersectT(s, o, d, mint, maxt, tOut, &phit)) return false;
  quadricPartials(s, phit, p, dpdu, dpdv);
  return true;
}
// shape.dart:96-98 -> cylinder.dart:230-240
static __device__ __noinline__ V3 cylinderSampleCold(const GSphere& s, bool rev, double u1, double u2, V3* ns) {
  float o2w[16], w2o[16];
  for (int i = 0; i < 12; ++i) { o2w[i] = s.o2w[i]; w2o[i] = s.w2o[i]; }
  for (int i = 0; i < 4; ++i) { o2w[12 + i] = s.o2wRow3[i]; w2o[12 + i] = s.w2oRow3[i]; }
  double z = s.zmin * (1.0 - u1) + s.zmax * u1;  // Lerp, common.dart:80-81
  double t = u2 * s.phiMax;
  V3 pc = mkv(s.radius * cos(t), s.radius * sin(t), z);
  V3 n = XfNormal(w2o, V3{pc.x, pc.y, 0.f});
  n = n / Length(n);
  if (rev) n = n * -1.0;
  *ns = n;
  return XfPoint(o2w, pc);
}
#endif

// disk.dart:39-67: tHit and the object-space hit point
static DRT_HD inline bool diskIntersectT(const GSphere& s, const V3& o, const V3& d, double mint, double maxt, double* tOut,
                                         V3* phitOut) {
  float w2o[16];
  for (int i = 0; i < 12; ++i) w2o[i] = s.w2o[i];
  for (int i = 0; i < 4; ++i) w2o[12 + i] = s.w2oRow3[i];
  V3 ro = XfPoint(w2o, o), rd = XfVector(w2o, d);
  if (fabs((double)rd.z) < 1.0e-7) return false;
  double thit = (s.height - ro.z) / rd.z;
  if (thit < mint || thit > maxt) return false;
  V3 phit = RayAt(ro, rd, thit);
  double dist2 = (double)phit.x * phit.x + (double)phit.y * phit.y;
  if (dist2 > s.radius * s.radius || dist2 < s.innerRadius * s.innerRadius) return false;
  double phi = atan2((double)phit.y, (double)phit.x);
  if (phi < 0) phi += 2.0 * DRT_PI;
  if (phi > s.phiMax) return false;
  *tOut = thit;
  *phitOut = phit;
  return true;
}

// disk.dart:69-97: p, dpdu, dpdv in world space from the object-space hit point
static DRT_HD inline void diskPartials(const GSphere& s, const V3& phit, V3* p, V3* dpdu, V3* dpdv) {
  float o2w[16];
  for (int i = 0; i < 12; ++i) o2w[i] = s.o2w[i];
  for (int i = 0; i < 4; ++i) o2w[12 + i] = s.o2wRow3[i];
  double dist2 = (double)phit.x * phit.x + (double)phit.y * phit.y;
  double oneMinusV = (sqrt(dist2) - s.innerRadius) / (s.radius - s.innerRadius);
  double invOneMinusV = (oneMinusV > 0.0) ? (1.0 / oneMinusV) : 0.0;
  V3 du = mkv(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);
  V3 dv = mkv(-(double)phit.x * invOneMinusV, -(double)phit.y * invOneMinusV, 0.0);
  du = du * (s.phiMax * 0.15915494309189533577);  // INV_TWOPI, common.dart:24
  dv = dv * ((s.radius - s.innerRadius) / s.radius);
  *p = XfPoint(o2w, phit);
  *dpdu = XfVector(o2w, du);
  *dpdv = XfVector(o2w, dv);
}

// sphere.dart:118-160: p, dpdu, dpdv in world space from the object-space hit point
static DRT_HD inline void spherePartials(const GSphere& s, const V3& phit, V3* p, V3* dpdu, V3* dpdv) {
  float o2w[16];
  for (int i = 0; i < 12; ++i) o2w[i] = s.o2w[i];
  for (int i = 0; i < 4; ++i) o2w[12 + i] = s.o2wRow3[i];
  double theta = acos(clampD((double)phit.z / s.radius, -1.0, 1.0));
  double zradius = sqrt((double)phit.x * phit.x + (double)phit.y * phit.y);
  double invzradius = 1.0 / zradius;
  double cosphi = phit.x * invzradius, sinphi = phit.y * invzradius;
  V3 du = mkv(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);
  V3 dv = mkv(phit.z * cosphi, phit.z * sinphi, -s.radius * sin(theta)) * (s.thetaMax - s.thetaMin);
  *p = XfPoint(o2w, phit);
  *dpdu = XfVector(o2w, du);
  *dpdv = XfVector(o2w, dv);
}

static __device__ inline bool primReverse(const RenderScene& rs, uint32_t prim) { return (__ldg(rs.primAttr + prim) >> 31) != 0; }
static __device__ inline int primLight(const RenderScene& rs, uint32_t prim) { return (int)((__ldg(rs.primAttr + prim) >> 16) & 0x7fffu) - 1; }
static __device__ inline int primMaterial(const RenderScene& rs, uint32_t prim) { return (int)(__ldg(rs.primAttr + prim) & 0xffffu); }

// A triangle of a mesh that carries uv / N / S (drt_set_mesh_shading).  Two out-of-line pieces — meshes without attributes
// (the benchmark scenes) never come here, and no address of the caller's ShapeHit escapes into them:
//   triMeshDgCold       dg.dpdu / dg.nn from the mesh's own uvs (triangle.dart:104-131, getUVs :246-254)
//   triMeshShadingCold  Triangle.getShadingGeometry (:271-364): dgShading.nn and dgShading.dpdu.  The barycentrics are
//                       re-derived from the ray with the expressions of Triangle.intersect (:52-95), i.e. the values the
//                       reference holds at this point.
struct MeshGeom {
  V3 a, b;
};
static __device__ inline void meshTriUVs(const RenderScene& rs, const GMesh& mesh, uint32_t i0, uint32_t i1, uint32_t i2, double uv[6]) {
  uv[0] = 0.0; uv[1] = 0.0; uv[2] = 1.0; uv[3] = 0.0; uv[4] = 1.0; uv[5] = 1.0;
  if (mesh.flags & 4u) {
    uv[0] = rs.vertUV[2 * (size_t)i0]; uv[1] = rs.vertUV[2 * (size_t)i0 + 1];
    uv[2] = rs.vertUV[2 * (size_t)i1]; uv[3] = rs.vertUV[2 * (size_t)i1 + 1];
    uv[4] = rs.vertUV[2 * (size_t)i2]; uv[5] = rs.vertUV[2 * (size_t)i2 + 1];
  }
}
static __device__ __noinline__ void triMeshDgCold(const RenderScene& rs, uint32_t prim, bool rev, MeshGeom* out) {
  const GMesh& mesh = rs.meshes[__ldg(rs.meshOfTri + prim)];
  const uint32_t i0 = __ldg(rs.triIdx + 3 * (size_t)prim), i1 = __ldg(rs.triIdx + 3 * (size_t)prim + 1),
                 i2 = __ldg(rs.triIdx + 3 * (size_t)prim + 2);
  const TriVerts tv = loadTri(rs, prim);
  double uv[6];
  meshTriUVs(rs, mesh, i0, i1, i2, uv);
  V3 dpdu, dpdv;
  triPartialsUV(tv, uv, &dpdu, &dpdv);
  out->a = dpdu;
  out->b = shapeNormal(dpdu, dpdv, rev);
}
static __device__ __noinline__ void triMeshShadingCold(const RenderScene& rs, uint32_t prim, V3 o, V3 d, bool rev, V3 nn, V3 dpdu,
                                                       MeshGeom* out) {
  out->a = nn;    // dgShading.nn
  out->b = dpdu;  // dgShading.dpdu
  const GMesh& mesh = rs.meshes[__ldg(rs.meshOfTri + prim)];
  if (!(mesh.flags & 3u)) return;  // triangle.dart:273-276
  const uint32_t i0 = __ldg(rs.triIdx + 3 * (size_t)prim), i1 = __ldg(rs.triIdx + 3 * (size_t)prim + 1),
                 i2 = __ldg(rs.triIdx + 3 * (size_t)prim + 2);
  const TriVerts tv = loadTri(rs, prim);
  double uv[6];
  meshTriUVs(rs, mesh, i0, i1, i2, uv);
  // b1, b2 of Triangle.intersect (:52-95)
  double p1x = tv.p1.x, p1y = tv.p1.y, p1z = tv.p1.z;
  double e1x = (double)tv.p2.x - p1x, e1y = (double)tv.p2.y - p1y, e1z = (double)tv.p2.z - p1z;
  double e2x = (double)tv.p3.x - p1x, e2y = (double)tv.p3.y - p1y, e2z = (double)tv.p3.z - p1z;
  double dx = d.x, dy = d.y, dz = d.z;
  double s1x = (dy * e2z) - (dz * e2y), s1y = (dz * e2x) - (dx * e2z), s1z = (dx * e2y) - (dy * e2x);
  double invDivisor = 1.0 / ((s1x * e1x) + (s1y * e1y) + (s1z * e1z));
  double sx = (double)o.x - p1x, sy = (double)o.y - p1y, sz = (double)o.z - p1z;
  double b1 = (sx * s1x + sy * s1y + sz * s1z) * invDivisor;
  double s2x = (sy * e1z) - (sz * e1y), s2y = (sz * e1x) - (sx * e1z), s2z = (sx * e1y) - (sy * e1x);
  double b2 = ((dx * s2x) + (dy * s2y) + (dz * s2z)) * invDivisor;
  double b0 = 1.0 - b1 - b2;
  double tu = b0 * uv[0] + b1 * uv[2] + b2 * uv[4], tv_ = b0 * uv[1] + b1 * uv[3] + b2 * uv[5];
  // getShadingGeometry: barycentrics back from (u, v) by SolveLinearSystem2x2 (common.dart:170-185)
  double A0 = uv[2] - uv[0], A1 = uv[4] - uv[0], A2 = uv[3] - uv[1], A3 = uv[5] - uv[1];
  double C0 = tu - uv[0], C1 = tv_ - uv[1];
  double det = A0 * A3 - A1 * A2, bx, by = 0.0, bz = 0.0;
  bool ok = !(fabs(det) < 1.0e-10);
  if (ok) {
    by = (A3 * C0 - A1 * C1) / det;
    bz = (A0 * C1 - A2 * C0) / det;
    if (isnan(by) || isnan(bz)) ok = false;
  }
  if (!ok) bx = by = bz = 1.0 / 3.0;
  else bx = 1.0 - by - bz;
  V3 ns, ss, ts;
  if (mesh.flags & 1u) {
    const float* n = rs.vertN;
    V3 n0 = V3{n[3 * (size_t)i0], n[3 * (size_t)i0 + 1], n[3 * (size_t)i0 + 2]}, n1 = V3{n[3 * (size_t)i1], n[3 * (size_t)i1 + 1], n[3 * (size_t)i1 + 2]},
       n2 = V3{n[3 * (size_t)i2], n[3 * (size_t)i2 + 1], n[3 * (size_t)i2 + 2]};
    V3 ni = ((n0 * bx) + (n1 * by)) + (n2 * bz);
    const float* w = mesh.w2o;  // transformNormal: transpose of the inverse (transform.dart:147-161)
    ns = Normalize(mkv((double)w[0] * ni.x + (double)w[3] * ni.y + (double)w[6] * ni.z, (double)w[1] * ni.x + (double)w[4] * ni.y + (double)w[7] * ni.z,
                       (double)w[2] * ni.x + (double)w[5] * ni.y + (double)w[8] * ni.z));
  } else {
    ns = nn;
  }
  if (mesh.flags & 2u) {
    const float* sv = rs.vertS;
    V3 s0 = V3{sv[3 * (size_t)i0], sv[3 * (size_t)i0 + 1], sv[3 * (size_t)i0 + 2]}, s1 = V3{sv[3 * (size_t)i1], sv[3 * (size_t)i1 + 1], sv[3 * (size_t)i1 + 2]},
       s2 = V3{sv[3 * (size_t)i2], sv[3 * (size_t)i2 + 1], sv[3 * (size_t)i2 + 2]};
    V3 si = ((s0 * bx) + (s1 * by)) + (s2 * bz);
    const float* m = mesh.o2w;
    ss = Normalize(mkv((double)m[0] * si.x + (double)m[1] * si.y + (double)m[2] * si.z, (double)m[3] * si.x + (double)m[4] * si.y + (double)m[5] * si.z,
                       (double)m[6] * si.x + (double)m[7] * si.y + (double)m[8] * si.z));
  } else {
    ss = Normalize(dpdu);
  }
  ts = Cross(ss, ns);
  if (LengthSquared(ts) > 0.0) {
    ts = Normalize(ts);
    ss = Cross(ts, ns);
  } else {
    CoordinateSystem(ns, &ss, &ts);
  }
  out->a = shapeNormal(ss, ts, rev);  // dgShading.set(dg.p, ss, ts, ...): differential_geometry.dart:77-99
  out->b = ss;
}

// Differential geometry of a hit found by the traversal kernels (lib/core/intersection.dart:27-72):
// the shape is re-evaluated at the known tHit, which reproduces what Shape.intersect stored.
// EXTRA = false: the scene has neither per-vertex mesh attributes nor cylinder / cone / paraboloid / hyperboloid shapes
// (RenderScene::extra == 0); their out-of-line calls are compiled out of the kernels config 3 / 4 run.
template <bool EXTRA = (DRT_EXTRA != 0)>
static __device__ inline void hitGeometry(const RenderScene& rs, uint32_t prim, const V3& o, const V3& d, double t, ShapeHit* h) {
  h->t = t;
  V3 dpdv;
  if (prim < rs.ntris) {
    TriVerts tv = loadTri(rs, prim);
    h->p = RayAt(o, d, t);
    h->rayEps = 1.0e-3 * t;  // triangle.dart:157
    if (EXTRA && rs.meshOfTri) {  // through a temporary: the address of *h must not escape into an out-of-line call
      MeshGeom mg;
      triMeshDgCold(rs, prim, primReverse(rs, prim), &mg);
      h->dpdu = mg.a; h->nn = mg.b;
      return;
    }
    triPartials(tv, &h->dpdu, &dpdv);
  } else {
    const GSphere& s = rs.ts.spheres[prim - rs.ntris];
    // object-space hit point at tHit: ray.pointAt on the transformed ray (sphere.dart:62-64 / :95-97)
    float w2o[16];
    for (int i = 0; i < 12; ++i) w2o[i] = s.w2o[i];
    for (int i = 0; i < 4; ++i) w2o[12 + i] = s.w2oRow3[i];
    V3 ro = XfPoint(w2o, o), rd = XfVector(w2o, d);
    V3 phit = RayAt(ro, rd, t);
    if (EXTRA && s.shape >= 2) {  // results through temporaries: the address of *h must not escape into an out-of-line call
      V3 qp, qu, qv;
      quadricPartialsCold(s, phit, &qp, &qu, &qv);
      h->p = qp; h->dpdu = qu; dpdv = qv;
    } else if (s.shape == 1) {
      diskPartials(s, phit, &h->p, &h->dpdu, &dpdv);
    } else {
      if (phit.x == 0.0f && phit.y == 0.0f) phit.x = (float)(1.0e-5 * s.radius);
      spherePartials(s, phit, &h->p, &h->dpdu, &dpdv);
    }
    h->rayEps = 5.0e-4 * t;  // sphere.dart:164, disk.dart:100
  }
  h->nn = shapeNormal(h->dpdu, dpdv, primReverse(rs, prim));
}

// A hit that came through a TransformedPrimitive (transformed_primitive.dart:30-58): the shape saw the ray in primitive space
// (worldToPrimitive.interpolate(ray.time) applied to it), so its differential geometry is rebuilt there and moved to world space
// with Inverse(w2p) — p as a point, nn as a normal (then normalised), dpdu as a vector — unless w2p is the identity.
static __device__ __noinline__ void instanceHitCold(const RenderScene& rs, uint32_t prim, int inst, double time, V3 o, V3 d, double t,
                                                    ShapeHit* out) {
  M4 m, inv;
  animInterpolate(rs.ts.instances[inst], time, &m, &inv);
  const V3 o2 = XfPoint(m.d, o), d2 = XfVector(m.d, d);
  ShapeHit h;
  hitGeometry<true>(rs, prim, o2, d2, t, &h);
  if (!m4IsIdentity(m)) {
    h.p = XfPoint(inv.d, h.p);
    h.nn = Normalize(XfNormal(m.d, h.nn));  // p2w = Transform(inv, m): its transformNormal reads ITS inverse, m
    h.dpdu = XfVector(inv.d, h.dpdu);
  }
  *out = h;
}
// hitGeometry for entry q of the extension queue (slot = its wavefront slot): Wavefront::extInst names the instance, if any
template <bool EXTRA = (DRT_EXTRA != 0)>
static __device__ inline void hitGeometryQ(const RenderScene& rs, const Wavefront& wf, uint32_t q, uint32_t slot, uint32_t prim, const V3& o,
                                           const V3& d, double t, ShapeHit* h) {
  if (DRT_EXTRA && EXTRA && wf.extInst) {
    const int inst = wf.extInst[q];
    if (inst >= 0) {
      ShapeHit tmp;
      instanceHitCold(rs, prim, inst, wf.slotTime[slot], o, d, t, &tmp);
      *h = tmp;
      return;
    }
  }
  hitGeometry<EXTRA>(rs, prim, o, d, t, h);
}

// Shape.intersect on one shape with an explicit interval (ShapeSet / Shape.pdf2 use it directly,
// shape_set.dart:65-79, shape.dart:100-121)
#ifndef DRT_SHAPE_INLINE
#define DRT_SHAPE_INLINE inline  // measured on B200 (tools/shade_sweep.sh): inlined + 4 CTAs/SM is fastest
#endif
static __device__ DRT_SHAPE_INLINE bool shapeIntersect(const RenderScene& rs, uint32_t prim, const V3& o, const V3& d, double mint,
                                             double maxt, ShapeHit* h) {
  V3 dpdv;
  if (prim < rs.ntris) {
    TriVerts tv = loadTri(rs, prim);
    double t;
    if (!triIntersectT(tv, o, d, mint, maxt, &t)) return false;
    h->t = t;
    h->p = RayAt(o, d, t);
    h->rayEps = 1.0e-3 * t;
    if (DRT_EXTRA && rs.meshOfTri) {
      MeshGeom mg;
      triMeshDgCold(rs, prim, primReverse(rs, prim), &mg);
      h->dpdu = mg.a; h->nn = mg.b;
      return true;
    }
    triPartials(tv, &h->dpdu, &dpdv);
  } else {
    const GSphere& s = rs.ts.spheres[prim - rs.ntris];
    double t;
    V3 phit;
    if (DRT_EXTRA && s.shape >= 2) {
      V3 qp, qu, qv;
      double qt;
      if (!quadricIntersectCold(s, o, d, mint, maxt, &qt, &qp, &qu, &qv)) return false;
      t = qt; h->p = qp; h->dpdu = qu; dpdv = qv;
    } else if (s.shape == 1) {
      if (!diskIntersectT(s, o, d, mint, maxt, &t, &phit)) return false;
      diskPartials(s, phit, &h->p, &h->dpdu, &dpdv);
    } else {
      if (!sphereIntersectT(s, o, d, mint, maxt, &t, &phit)) return false;
      spherePartials(s, phit, &h->p, &h->dpdu, &dpdv);
    }
    h->t = t;
    h->rayEps = 5.0e-4 * t;
  }
  h->nn = shapeNormal(h->dpdu, dpdv, primReverse(rs, prim));
  return true;
}

// Out-of-line copy for the light-sampling code, where the general shape test only runs for sphere / disk light shapes
// (triangle light shapes are tested from their resident GLightShape record): inlining it at every call site of
// shapeSetSample / shapeSetPdf made shadePathKernel 15,152 instructions (242 KB), far beyond the instruction caches.
static __device__ __noinline__ bool shapeIntersectCold(const RenderScene& rs, uint32_t prim, V3 o, V3 d, double mint, double maxt,
                                                       ShapeHit* h) {
  return shapeIntersect(rs, prim, o, d, mint, maxt, h);
}

// ---- BSDF: one diffuse lobe (bsdf.dart:41-255, bxdf.dart:28-91, lambertian.dart:30-48,
// oren_nayar.dart:24-58, matte_material.dart:41-65) ------------------------------------------------------
enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };

struct Bsdf {
  V3 nn, ng, sn, tn;
  bool hasBxdf, orenNayar;
  Spec R;
  double A, B;
};

static __device__ inline V3 bsdfToLocal(const Bsdf& b, const V3& v) { return mkv(Dot(v, b.sn), Dot(v, b.tn), Dot(v, b.nn)); }
static __device__ inline V3 bsdfToWorld(const Bsdf& b, const V3& v) {
  return mkv((double)b.sn.x * v.x + (double)b.tn.x * v.y + (double)b.nn.x * v.z,
             (double)b.sn.y * v.x + (double)b.tn.y * v.y + (double)b.nn.y * v.z,
             (double)b.sn.z * v.x + (double)b.tn.z * v.y + (double)b.nn.z * v.z);
}
static __device__ inline double AbsCosTheta(const V3& v) { return fabs((double)v.z); }
static __device__ inline double SinTheta2(const V3& v) { return fmax(0.0, 1.0 - (double)v.z * v.z); }
static __device__ inline double SinTheta(const V3& v) { return sqrt(SinTheta2(v)); }
static __device__ inline double CosPhi(const V3& v) { double s = SinTheta(v); return s == 0.0 ? 1.0 : clampD((double)v.x / s, -1.0, 1.0); }
static __device__ inline double SinPhi(const V3& v) { double s = SinTheta(v); return s == 0.0 ? 0.0 : clampD((double)v.y / s, -1.0, 1.0); }

// dgShading.nn / dgShading.dpdu of the hit: dg's own for meshes without N / S (triangle.dart:273-276) and quadrics
// (shape.dart:73-77).  o, d: the ray that found the hit.
template <bool EXTRA>
static __device__ inline void shadingFrame(const RenderScene& rs, uint32_t prim, const ShapeHit& h, const V3& o, const V3& d, V3* nsh,
                                           V3* ssh) {
  *nsh = h.nn;
  *ssh = h.dpdu;
  if (EXTRA && rs.meshOfTri && prim < rs.ntris) {
    MeshGeom mg;
    triMeshShadingCold(rs, prim, o, d, primReverse(rs, prim), h.nn, h.dpdu, &mg);
    *nsh = mg.a;
    *ssh = mg.b;
  }
}
template <bool EXTRA = (DRT_EXTRA != 0)>
static __device__ inline Bsdf makeBsdf(const RenderScene& rs, uint32_t prim, const ShapeHit& h, const V3& o, const V3& d) {
  Bsdf b;
  V3 ssh;
  shadingFrame<EXTRA>(rs, prim, h, o, d, &b.nn, &ssh);
  b.ng = h.nn;
  b.sn = Normalize(ssh);
  b.tn = Cross(b.nn, b.sn);
  const GMaterial m = rs.materials[primMaterial(rs, prim)];
  Spec r = mks(clampD(m.kd[0], 0.0, CUDART_INF), clampD(m.kd[1], 0.0, CUDART_INF), clampD(m.kd[2], 0.0, CUDART_INF));
  double sig = clampD(m.sigma, 0.0, 90.0);
  b.hasBxdf = !IsBlack(r);
  b.orenNayar = false;
  b.R = r;
  b.A = b.B = 0.0;
  if (b.hasBxdf && sig != 0.0) {  // oren_nayar.dart:24-31
    b.orenNayar = true;
    double sigma = (DRT_PI / 180.0) * sig, sigma2 = sigma * sigma;
    b.A = 1.0 - (sigma2 / (2.0 * (sigma2 + 0.33)));
    b.B = 0.45 * sigma2 / (sigma2 + 0.09);
  }
  return b;
}
static __device__ inline Spec bxdfF(const Bsdf& b, const V3& wo, const V3& wi) {
  if (!b.orenNayar) return b.R * DRT_INV_PI;  // lambertian.dart:35-37
  double sinthetai = SinTheta(wi), sinthetao = SinTheta(wo);  // oren_nayar.dart:33-58
  double maxcos = 0.0;
  if (sinthetai > 1e-4 && sinthetao > 1e-4) {
    double dcos = CosPhi(wi) * CosPhi(wo) + SinPhi(wi) * SinPhi(wo);
    maxcos = fmax(0.0, dcos);
  }
  double sinalpha, tanbeta;
  if (AbsCosTheta(wi) > AbsCosTheta(wo)) { sinalpha = sinthetao; tanbeta = sinthetai / AbsCosTheta(wi); }
  else { sinalpha = sinthetai; tanbeta = sinthetao / AbsCosTheta(wo); }
  return b.R * (DRT_INV_PI * (b.A + b.B * maxcos * sinalpha * tanbeta));
}
static __device__ inline double bxdfPdf(const V3& wo, const V3& wi) {  // bxdf.dart:84-88
  return ((double)wo.z * wi.z > 0.0) ? AbsCosTheta(wi) * DRT_INV_PI : 0.0;
}
static __device__ inline bool bxdfMatches(int flags) { const int type = BSDF_REFLECTION | BSDF_DIFFUSE; return (type & flags) == type; }
static __device__ inline Spec bsdfF(const Bsdf& b, const V3& woW, const V3& wiW, int flags) {  // bsdf.dart:177-198
  V3 wi = bsdfToLocal(b, wiW), wo = bsdfToLocal(b, woW);
  if (Dot(wiW, b.ng) * Dot(woW, b.ng) > 0) flags &= ~BSDF_TRANSMISSION;
  else flags &= ~BSDF_REFLECTION;
  Spec r = mks1(0.0);
  if (b.hasBxdf && bxdfMatches(flags)) r = r + bxdfF(b, wo, wi);
  return r;
}
static __device__ inline double bsdfPdf(const Bsdf& b, const V3& woW, const V3& wiW, int flags) {  // bsdf.dart:128-146
  if (!b.hasBxdf) return 0.0;
  V3 wo = bsdfToLocal(b, woW), wi = bsdfToLocal(b, wiW);
  return bxdfMatches(flags) ? bxdfPdf(wo, wi) / 1 : 0.0;
}
// bsdf.dart:53-126 with one BxDF; u0/u1 are the float32 direction samples (the component sample picks
// among matching BxDFs: one candidate, nothing to pick)
static __device__ inline Spec bsdfSampleF(const Bsdf& b, const V3& woW, V3* wiW, float u0, float u1, double* pdfOut, int flags,
                                          int* sampledType) {
  *sampledType = 0;
  *pdfOut = 0.0;
  if (!(b.hasBxdf && bxdfMatches(flags))) return mks1(0.0);
  V3 wo = bsdfToLocal(b, woW);
  V3 wi = CosineSampleHemisphere(u0, u1);  // bxdf.dart:37-48
  if (wo.z < 0.0f) wi.z = (float)((double)wi.z * -1.0);
  *pdfOut = bxdfPdf(wo, wi);
  if (*pdfOut == 0.0) return mks1(0.0);
  *sampledType = BSDF_REFLECTION | BSDF_DIFFUSE;
  *wiW = bsdfToWorld(b, wi);
  Spec r = mks1(0.0);
  if (Dot(*wiW, b.ng) * Dot(woW, b.ng) > 0) flags &= ~BSDF_TRANSMISSION;
  else flags &= ~BSDF_REFLECTION;
  if (bxdfMatches(flags)) r = r + bxdfF(b, wo, wi);
  return r;
}

// The single-lobe BSDF ignores the component sample (one candidate, nothing to pick)
static __device__ inline Spec bsdfSampleF(const Bsdf& b, const V3& woW, V3* wiW, float u0, float u1, double comp, double* pdfOut,
                                          int flags, int* sampledType) {
  (void)comp;
  return bsdfSampleF(b, woW, wiW, u0, u1, pdfOut, flags, sampledType);
}

// ---- general BSDF: an ordered list of BxDFs per material (bsdf.dart:41-255; lambertian / oren_nayar / microfacet +
// blinn / specular_reflection / specular_transmission / fresnel_*.dart).  Used when drt_set_material_lobes defined
// the materials; scenes with matte materials only keep the single-lobe code above.
struct BsdfG {
  V3 nn, ng, sn, tn;
  int n;
  const GLobe* lobes;
};
static __device__ inline V3 bsdfToLocal(const BsdfG& b, const V3& v) { return mkv(Dot(v, b.sn), Dot(v, b.tn), Dot(v, b.nn)); }
static __device__ inline V3 bsdfToWorld(const BsdfG& b, const V3& v) {
  return mkv((double)b.sn.x * v.x + (double)b.tn.x * v.y + (double)b.nn.x * v.z,
             (double)b.sn.y * v.x + (double)b.tn.y * v.y + (double)b.nn.y * v.z,
             (double)b.sn.z * v.x + (double)b.tn.z * v.y + (double)b.nn.z * v.z);
}
template <bool EXTRA = (DRT_EXTRA != 0)>
static __device__ inline BsdfG makeBsdfG(const RenderScene& rs, uint32_t prim, const ShapeHit& h, const V3& o, const V3& d) {
  BsdfG b;
  V3 ssh;
  shadingFrame<EXTRA>(rs, prim, h, o, d, &b.nn, &ssh);
  b.ng = h.nn;
  b.sn = Normalize(ssh);
  b.tn = Cross(b.nn, b.sn);
  const uint2 ml = __ldg(rs.matLobes + primMaterial(rs, prim));
  b.lobes = rs.lobes + ml.x;
  b.n = (int)ml.y;
  return b;
}
// dart:math min / max: NaN when either argument is NaN
static __device__ inline double dartMin(double a, double b) { return (isnan(a) || isnan(b)) ? CUDART_NAN : (a < b ? a : b); }
static __device__ inline double dartMax(double a, double b) { return (isnan(a) || isnan(b)) ? CUDART_NAN : (a > b ? a : b); }
#if DRT_F32_SINGLE_OPS
static __device__ inline Spec operator-(const Spec& a, const Spec& b) { return Spec{DRT_FSUB(a.r, b.r), DRT_FSUB(a.g, b.g), DRT_FSUB(a.b, b.b)}; }
static __device__ inline Spec operator/(const Spec& a, const Spec& b) { return Spec{DRT_FDIV(a.r, b.r), DRT_FDIV(a.g, b.g), DRT_FDIV(a.b, b.b)}; }
#else
static __device__ inline Spec operator-(const Spec& a, const Spec& b) { return mks((double)a.r - b.r, (double)a.g - b.g, (double)a.b - b.b); }
static __device__ inline Spec operator/(const Spec& a, const Spec& b) { return mks((double)a.r / b.r, (double)a.g / b.g, (double)a.b / b.b); }
#endif
static __device__ inline bool SameHemisphere(const V3& w, const V3& wp) { return (double)w.z * wp.z > 0.0; }  // vector.dart:194-196

static __device__ inline int lobeType(int kind) {
  return kind <= 1 ? (BSDF_REFLECTION | BSDF_DIFFUSE)
                   : ((kind == 2 || kind >= 5) ? (BSDF_REFLECTION | BSDF_GLOSSY)
                                : (kind == 3 ? (BSDF_REFLECTION | BSDF_SPECULAR) : (BSDF_TRANSMISSION | BSDF_SPECULAR)));
}
static __device__ inline int lobeTypeOf(const GLobe& l) {  // brdf_to_btdf.dart:27-29 flips reflection <-> transmission
  const int t = lobeType(l.kind);
  return (l.wrap & 1) ? (t ^ (BSDF_REFLECTION | BSDF_TRANSMISSION)) : t;
}
static __device__ inline bool lobeMatches(const GLobe& l, int flags) { const int t = lobeTypeOf(l); return (t & flags) == t; }

static __device__ inline Spec fresnelDielectric(double cosi, double eta_i, double eta_t) {  // fresnel_dielectric.dart:24-56
  if (!isnan(cosi)) cosi = clampD(cosi, -1.0, 1.0);
  const bool entering = cosi > 0.0;
  double ei = eta_i, et = eta_t;
  if (!entering) { const double t = ei; ei = et; et = t; }
  const double sint = ei / et * sqrt(dartMax(0.0, 1.0 - cosi * cosi));
  if (sint >= 1.0) return mks1(1.0);
  const double cost = sqrt(dartMax(0.0, 1.0 - sint * sint));
  cosi = fabs(cosi);
  const double Rparl = ((et * cosi) - (ei * cost)) / ((et * cosi) + (ei * cost));
  const double Rperp = ((ei * cosi) - (et * cost)) / ((ei * cosi) + (et * cost));
  return mks1((Rparl * Rparl + Rperp * Rperp) / 2.0);
}
static __device__ inline Spec lobeFresnel(const GLobe& l, double cosi) {
  if (l.fresnel == 0) return mks1(1.0);  // fresnel_no_op.dart
  if (l.fresnel == 1) return fresnelDielectric(cosi, l.ei, l.et);
  cosi = fabs(cosi);  // fresnel_conductor.dart:24-49
  const Spec ONE = mks1(1.0), cosSqr = mks1(cosi * cosi);
  const Spec eta = Spec{l.eta[0], l.eta[1], l.eta[2]}, k = Spec{l.k[0], l.k[1], l.k[2]};
  const Spec tmp = (eta * eta + k * k) * (cosi * cosi);
  Spec r1 = (tmp - (eta * (2.0 * cosi)) + ONE);
  Spec r2 = (tmp + (eta * (2.0 * cosi)) + ONE);
  const Spec Rparl2 = r1 / r2;
  const Spec tmp_f = eta * eta + k * k;
  r1 = (tmp_f - (eta * (2.0 * cosi)) + cosSqr);
  r2 = (tmp_f + (eta * (2.0 * cosi)) + cosSqr);
  const Spec Rperp2 = r1 / r2;
  return (Rparl2 + Rperp2) / 2.0;
}
static __device__ inline double blinnPdfOf(double exponent, double costheta, double woDotWh) {  // blinn.dart:50-56,62-68
  double pdf = ((exponent + 1.0) * pow(costheta, exponent)) / (2.0 * DRT_PI * 4.0 * woDotWh);
  if (woDotWh <= 0.0) pdf = 0.0;
  return pdf;
}
// FresnelBlend (fresnel_blend.dart:24-90) over an Anisotropic distribution (anisotropic.dart:27-121): lobe kind 5 with Rd = rgb,
// Rs = eta, ex = param, ey = ei.  SubstrateMaterial's only BxDF; kept out of line (rare, pow / atan / tan heavy).
static __device__ inline double anisoPdfOf(const GLobe& l, const V3& wo, const V3& wh) {
  const double costhetah = AbsCosTheta(wh), ds = 1.0 - costhetah * costhetah;
  double p = 0.0;
  if (ds > 0.0 && Dot(wo, wh) > 0.0) {
    const double e = (l.param * wh.x * wh.x + l.ei * wh.y * wh.y) / ds;
    const double dd = sqrt((l.param + 1.0) * (l.ei + 1.0)) * DRT_INV_TWOPI * pow(costhetah, e);
    p = dd / (4.0 * Dot(wo, wh));
  }
  return p;
}
static __device__ __noinline__ void blendFCold(const GLobe& l, V3 wo, V3 wi, Spec* out) {
  const Spec ONE = mks1(1.0), Rd = Spec{l.rgb[0], l.rgb[1], l.rgb[2]}, Rs = Spec{l.eta[0], l.eta[1], l.eta[2]};
  const Spec diffuse = Rd * ((28.0 / (23.0 * DRT_PI))) * (ONE - Rs) *
                       ((1.0 - pow(1.0 - 0.5 * AbsCosTheta(wi), 5.0)) * (1.0 - pow(1.0 - 0.5 * AbsCosTheta(wo), 5.0)));
  V3 wh = wi + wo;
  if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) { *out = mks1(0.0); return; }
  wh = Normalize(wh);
  const double costhetah = fabs((double)wh.z), d1 = 1.0 - costhetah * costhetah;
  double D = 0.0;
  if (d1 != 0.0) {
    const double e = (l.param * wh.x * wh.x + l.ei * wh.y * wh.y) / d1;
    D = sqrt((l.param + 2.0) * (l.ei + 2.0)) * DRT_INV_TWOPI * pow(costhetah, e);
  }
  const double a = D / (4.0 * AbsDot(wi, wh) * dartMax(AbsCosTheta(wi), AbsCosTheta(wo)));
  const Spec b = Rs + (ONE - Rs) * (pow(1.0 - Dot(wi, wh), 5.0));
  *out = diffuse + b * a;
}
static __device__ __noinline__ double blendPdfCold(const GLobe& l, V3 wo, V3 wi) {
  if (!SameHemisphere(wo, wi)) return 0.0;
  return 0.5 * (AbsCosTheta(wi) * DRT_INV_PI + anisoPdfOf(l, wo, Normalize(wo + wi)));
}
static __device__ inline void anisoFirstQuadrant(const GLobe& l, double u1, double u2, double* phi, double* costheta) {
  const double ex = l.param, ey = l.ei;
  if (ex == ey) *phi = DRT_PI * u1 * 0.5;
  else *phi = atan(sqrt((ex + 1.0) / (ey + 1.0)) * tan(DRT_PI * u1 * 0.5));
  const double cosphi = cos(*phi), sinphi = sin(*phi);
  *costheta = pow(u2, 1.0 / (ex * cosphi * cosphi + ey * sinphi * sinphi + 1.0));
}
static __device__ __noinline__ void blendSampleCold(const GLobe& l, V3 wo, double u1, double u2, V3* wiOut, double* pdfOut, Spec* fOut) {
  V3 wi;
  if (u1 < 0.5) {
    u1 = 2.0 * u1;
    wi = CosineSampleHemisphere(u1, u2);
    if (wo.z < 0.0f) wi.z = (float)((double)wi.z * -1.0);
  } else {
    u1 = 2.0 * (u1 - 0.5);
    double phi, cosTheta;
    if (u1 < 0.25) {
      anisoFirstQuadrant(l, 4.0 * u1, u2, &phi, &cosTheta);
    } else if (u1 < 0.5) {
      u1 = 4.0 * (0.5 - u1);
      anisoFirstQuadrant(l, u1, u2, &phi, &cosTheta);
      phi = DRT_PI - phi;
    } else if (u1 < 0.75) {
      u1 = 4.0 * (u1 - 0.5);
      anisoFirstQuadrant(l, u1, u2, &phi, &cosTheta);
      phi += DRT_PI;
    } else {
      u1 = 4.0 * (1.0 - u1);
      anisoFirstQuadrant(l, u1, u2, &phi, &cosTheta);
      phi = 2.0 * DRT_PI - phi;
    }
    const double sintheta = sqrt(dartMax(0.0, 1.0 - cosTheta * cosTheta));
    V3 wh = mkv(sintheta * cos(phi), sintheta * sin(phi), cosTheta);
    if (!SameHemisphere(wo, wh)) wh = -wh;
    wi = -wo + wh * 2.0 * Dot(wo, wh);
    *pdfOut = anisoPdfOf(l, wo, wh);
    if (!SameHemisphere(wo, wi)) { *wiOut = wi; *fOut = mks1(0.0); return; }
  }
  *wiOut = wi;
  *pdfOut = blendPdfCold(l, wo, wi);
  blendFCold(l, wo, wi, fOut);
}

// MeasuredMaterial's two BxDFs (lobe kinds 6 / 7; render_types.h GMeasured: the table's device pointer travels in the bits of `et`,
// its dimensions in the bits of `k`).  Both keep BxDF's cosine sampling and density (bxdf.dart:37-48,84-88).  Out of line: rare.
static __device__ inline const float* measuredData(const GLobe& l) { return (const float*)(uintptr_t)__double_as_longlong(l.et); }
static __device__ inline double SphericalPhi(const V3& v) {  // vector.dart:189-192
  const double p = atan2((double)v.y, (double)v.x);
  return p < 0.0 ? p + 2.0 * DRT_PI : p;
}
// regular_halfangle_brdf.dart:27-75.  REMAP(V, MAX, COUNT) => ((V / MAX).toInt() * COUNT).clamp(0, COUNT - 1) AS WRITTEN: the quotient is
// truncated BEFORE it is multiplied, so an index is 0 until V / MAX reaches 1
static __device__ inline int measuredRemap(double v, double mx, int count) {
  const long long q = (long long)trunc(v / mx) * count;
  return (int)(q < 0 ? 0 : (q > count - 1 ? count - 1 : q));
}
static __device__ __noinline__ void regularHalfangleFCold(const GLobe& l, V3 wo, V3 wi, Spec* out) {
  const int nThetaH = __float_as_int(l.k[0]), nThetaD = __float_as_int(l.k[1]), nPhiD = __float_as_int(l.k[2]);
  V3 wh = wo + wi;
  if (wh.z < 0.0f) { wo = -wo; wi = -wi; wh = -wh; }
  if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) { *out = mks1(0.0); return; }
  wh = Normalize(wh);
  const double whTheta = acos(clampD((double)wh.z, -1.0, 1.0));
  const double whCosPhi = CosPhi(wh), whSinPhi = SinPhi(wh), whCosTheta = (double)wh.z, whSinTheta = SinTheta(wh);
  const V3 whx = mkv(whCosPhi * whCosTheta, whSinPhi * whCosTheta, -whSinTheta);
  const V3 why = mkv(-whSinPhi, whCosPhi, 0.0);
  const V3 wd = mkv(Dot(wi, whx), Dot(wi, why), Dot(wi, wh));
  const double wdTheta = acos(clampD((double)wd.z, -1.0, 1.0));
  double wdPhi = SphericalPhi(wd);
  if (wdPhi > DRT_PI) wdPhi -= DRT_PI;
  const int whThetaIndex = measuredRemap(sqrt(dartMax(0.0, whTheta / (DRT_PI / 2.0))), 1.0, nThetaH);
  const int wdThetaIndex = measuredRemap(wdTheta, DRT_PI / 2.0, nThetaD);
  const int wdPhiIndex = measuredRemap(wdPhi, DRT_PI, nPhiD);
  const size_t index = (size_t)wdPhiIndex + (size_t)nPhiD * ((size_t)wdThetaIndex + (size_t)whThetaIndex * nThetaD);
  const float* brdf = measuredData(l);
  *out = Spec{__ldg(brdf + 3 * index), __ldg(brdf + 3 * index + 1), __ldg(brdf + 3 * index + 2)};
}
// irregular_isotropic_brdf.dart:36-62 over BRDFRemap (brdf_remap.dart:23-47).  KdTree.lookup (kdtree.dart:86-112) hands proc() exactly
// the samples with DistanceSquared(sample.p, m) < maxDist2, in an order that depends on object hash codes (kdtree.dart:120-124): the
// reference does not fix the order of its float32 sums, the oracle visits the samples in file order, and so does this loop.
static __device__ __noinline__ void irregularIsotropicFCold(const GLobe& l, V3 wo, V3 wi, Spec* out) {
  const int n = __float_as_int(l.k[0]);
  const float* data = measuredData(l);
  const double cosi = (double)wi.z, coso = (double)wo.z, sini = SinTheta(wi), sino = SinTheta(wo);
  double dphi = SphericalPhi(wi) - SphericalPhi(wo);
  if (dphi < 0.0) dphi += 2.0 * DRT_PI;
  if (dphi > 2.0 * DRT_PI) dphi -= 2.0 * DRT_PI;
  if (dphi > DRT_PI) dphi = 2.0 * DRT_PI - dphi;
  const V3 m = mkv(sini * sino, dphi / DRT_PI, cosi * coso);
  double lastMaxDist2 = 0.001;
  for (;;) {
    Spec v = mks1(0.0);
    double sumWeights = 0.0;
    int nFound = 0;
    for (int i = 0; i < n; ++i) {
      const float* q = data + 6 * (size_t)i;
      const V3 sp = V3{__ldg(q), __ldg(q + 1), __ldg(q + 2)};
      const double d2 = DistanceSquared(sp, m);
      if (d2 < lastMaxDist2) {
        const double weight = exp(-100.0 * d2);
        v = v + Spec{__ldg(q + 3), __ldg(q + 4), __ldg(q + 5)} * weight;
        sumWeights += weight;
        ++nFound;
      }
    }
    if (nFound > 2 || lastMaxDist2 > 1.5) {
      *out = mks(clampD((double)v.r, 0.0, CUDART_INF), clampD((double)v.g, 0.0, CUDART_INF), clampD((double)v.b, 0.0, CUDART_INF)) / sumWeights;
      return;
    }
    lastMaxDist2 *= 2.0;
  }
}

static __device__ inline Spec lobeBaseF(const GLobe& l, const V3& wo, const V3& wi) {
  if (DRT_EXTRA && l.kind >= 5) {
    Spec r;
    if (l.kind == 5) blendFCold(l, wo, wi, &r);
    else if (l.kind == 6) regularHalfangleFCold(l, wo, wi, &r);
    else irregularIsotropicFCold(l, wo, wi, &r);
    return r;
  }
  const Spec R = Spec{l.rgb[0], l.rgb[1], l.rgb[2]};
  if (l.kind == 0) return R * DRT_INV_PI;  // lambertian.dart:35-37
  if (l.kind == 1) {                       // oren_nayar.dart:24-58
    const double sigma = (DRT_PI / 180.0) * l.param, sigma2 = sigma * sigma;
    const double A = 1.0 - (sigma2 / (2.0 * (sigma2 + 0.33))), B = 0.45 * sigma2 / (sigma2 + 0.09);
    const double sinthetai = SinTheta(wi), sinthetao = SinTheta(wo);
    double maxcos = 0.0;
    if (sinthetai > 1e-4 && sinthetao > 1e-4) {
      const double dcos = CosPhi(wi) * CosPhi(wo) + SinPhi(wi) * SinPhi(wo);
      maxcos = fmax(0.0, dcos);
    }
    double sinalpha, tanbeta;
    if (AbsCosTheta(wi) > AbsCosTheta(wo)) { sinalpha = sinthetao; tanbeta = sinthetai / AbsCosTheta(wi); }
    else { sinalpha = sinthetai; tanbeta = sinthetao / AbsCosTheta(wo); }
    return R * (DRT_INV_PI * (A + B * maxcos * sinalpha * tanbeta));
  }
  if (l.kind == 2) {  // microfacet.dart:28-57, blinn.dart:31-34
    const double cosThetaO = AbsCosTheta(wo), cosThetaI = AbsCosTheta(wi);
    if (cosThetaI == 0.0 || cosThetaO == 0.0) return mks1(0.0);
    V3 wh = wi + wo;
    if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return mks1(0.0);
    wh = Normalize(wh);
    const double cosThetaH = Dot(wi, wh);
    const Spec F = lobeFresnel(l, cosThetaH);
    const double NdotWh = AbsCosTheta(wh), WOdotWh = AbsDot(wo, wh);
    const double G = dartMin(1.0, dartMin((2.0 * NdotWh * cosThetaO / WOdotWh), (2.0 * NdotWh * cosThetaI / WOdotWh)));
    const double D = (l.param + 2.0) * DRT_INV_TWOPI * pow(NdotWh, l.param);
    return R * (D * G) * F / (4.0 * cosThetaI * cosThetaO);
  }
  return mks1(0.0);  // specular BxDFs: f == 0
}
static __device__ inline double lobeBasePdf(const GLobe& l, const V3& wo, const V3& wi) {
  if (DRT_EXTRA && l.kind == 5) return blendPdfCold(l, wo, wi);
  if (l.kind <= 1 || l.kind >= 6) return SameHemisphere(wo, wi) ? AbsCosTheta(wi) * DRT_INV_PI : 0.0;  // bxdf.dart:84-88
  if (l.kind == 2) {                                                                     // microfacet.dart:68-73
    if (!SameHemisphere(wo, wi)) return 0.0;
    const V3 wh = Normalize(wo + wi);
    return blinnPdfOf(l.param, AbsCosTheta(wh), Dot(wo, wh));
  }
  return 0.0;
}
// *pdfOut is left untouched when the BxDF returns without setting it (specular_transmission.dart:52-54)
static __device__ inline Spec lobeBaseSampleF(const GLobe& l, const V3& wo, V3* wi, double u1, double u2, double* pdfOut) {
  if (DRT_EXTRA && l.kind == 5) {
    V3 w;
    Spec f;
    double p = *pdfOut;
    blendSampleCold(l, wo, u1, u2, &w, &p, &f);
    *wi = w;
    *pdfOut = p;
    return f;
  }
  if (l.kind <= 1 || l.kind >= 6) {  // bxdf.dart:37-48
    *wi = CosineSampleHemisphere(u1, u2);
    if (wo.z < 0.0f) wi->z = (float)((double)wi->z * -1.0);
    *pdfOut = lobeBasePdf(l, wo, *wi);
    return lobeBaseF(l, wo, *wi);
  }
  if (l.kind == 2) {  // microfacet.dart:59-66 + blinn.dart:36-60
    const double exponent = l.param;
    const double costheta = pow(u1, 1.0 / (exponent + 1.0));
    const double sintheta = sqrt(dartMax(0.0, 1.0 - costheta * costheta));
    const double phi = u2 * 2.0 * DRT_PI;
    V3 wh = mkv(sintheta * cos(phi), sintheta * sin(phi), costheta);
    if (!SameHemisphere(wo, wh)) wh = -wh;
    *wi = -wo + wh * 2.0 * Dot(wo, wh);
    *pdfOut = blinnPdfOf(exponent, costheta, Dot(wo, wh));
    if (!SameHemisphere(wo, *wi)) return mks1(0.0);
    return lobeBaseF(l, wo, *wi);
  }
  const Spec R = Spec{l.rgb[0], l.rgb[1], l.rgb[2]};
  if (l.kind == 3) {  // specular_reflection.dart:34-41
    *wi = V3{-wo.x, -wo.y, wo.z};
    *pdfOut = 1.0;
    return (lobeFresnel(l, (double)wo.z) * R) / AbsCosTheta(*wi);
  }
  // specular_transmission.dart:37-66
  const bool entering = (double)wo.z > 0.0;
  double ei = l.ei, et = l.et;
  if (!entering) { const double t = ei; ei = et; et = t; }
  const double sini2 = SinTheta2(wo);
  const double eta = ei / et;
  const double sint2 = eta * eta * sini2;
  if (sint2 >= 1.0) return mks1(0.0);
  double cost = sqrt(dartMax(0.0, 1.0 - sint2));
  if (entering) cost = -cost;
  *wi = mkv(eta * -(double)wo.x, eta * -(double)wo.y, cost);
  *pdfOut = 1.0;
  const Spec F = fresnelDielectric((double)wo.z, l.ei, l.et);
  return ((mks1(1.0) - F) * R) / AbsCosTheta(*wi);
}

// The BxDF the BSDF holds: the lobe itself, BRDFToBTDF(lobe) (brdf_to_btdf.dart:31-58: the other hemisphere of wi) and / or
// ScaledBxDF(.., s) (scaled_bxdf.dart:24-52: s * f; it does not override pdf, so BxDF.pdf's cosine density answers, bxdf.dart:84-88)
static __device__ inline V3 OtherHemisphere(const V3& w) { return V3{w.x, w.y, -w.z}; }
static __device__ inline Spec lobeF(const GLobe& l, const V3& wo, const V3& wi) {
  if (l.wrap == 0) return lobeBaseF(l, wo, wi);
  const Spec r = lobeBaseF(l, wo, (l.wrap & 1) ? OtherHemisphere(wi) : wi);
  return (l.wrap & 2) ? Spec{l.scale[0], l.scale[1], l.scale[2]} * r : r;
}
static __device__ inline double lobePdf(const GLobe& l, const V3& wo, const V3& wi) {
  if (l.wrap & 2) return SameHemisphere(wo, wi) ? AbsCosTheta(wi) * DRT_INV_PI : 0.0;
  return lobeBasePdf(l, wo, (l.wrap & 1) ? OtherHemisphere(wi) : wi);
}
static __device__ inline Spec lobeSampleF(const GLobe& l, const V3& wo, V3* wi, double u1, double u2, double* pdfOut) {
  const Spec r = lobeBaseSampleF(l, wo, wi, u1, u2, pdfOut);
  if (l.wrap & 1) *wi = OtherHemisphere(*wi);
  return (l.wrap & 2) ? Spec{l.scale[0], l.scale[1], l.scale[2]} * r : r;
}

// The three BSDF entry points of a BxDF list are OUT OF LINE: inlined at their five call sites they made shadePathKernel<GENERAL>
// 23.5 K instructions, and ncu showed the kernel waiting for instruction fetch (stall no_instruction 59 per issue, issue active
// 6 %: profiles/r01n_summary.md).  One copy of each keeps the divergent warps inside the instruction cache.
static __device__ __noinline__ Spec bsdfF(const BsdfG& b, const V3& woW, const V3& wiW, int flags) {  // bsdf.dart:177-198
  const V3 wi = bsdfToLocal(b, wiW), wo = bsdfToLocal(b, woW);
  if (Dot(wiW, b.ng) * Dot(woW, b.ng) > 0) flags &= ~BSDF_TRANSMISSION;
  else flags &= ~BSDF_REFLECTION;
  Spec r = mks1(0.0);
  for (int i = 0; i < b.n; ++i) {
    const GLobe& l = b.lobes[i];
    if (lobeMatches(l, flags)) r = r + lobeF(l, wo, wi);
  }
  return r;
}
static __device__ __noinline__ double bsdfPdf(const BsdfG& b, const V3& woW, const V3& wiW, int flags) {  // bsdf.dart:128-146
  if (b.n == 0) return 0.0;
  const V3 wo = bsdfToLocal(b, woW), wi = bsdfToLocal(b, wiW);
  double p = 0.0;
  int matching = 0;
  for (int i = 0; i < b.n; ++i) {
    const GLobe& l = b.lobes[i];
    if (lobeMatches(l, flags)) { ++matching; p += lobePdf(l, wo, wi); }
  }
  return matching > 0 ? p / matching : 0.0;
}
static __device__ __noinline__ Spec bsdfSampleF(const BsdfG& b, const V3& woW, V3* wiW, float u0, float u1, double comp, double* pdfOut,
                                          int flags, int* sampledType) {  // bsdf.dart:53-126
  *sampledType = 0;
  *pdfOut = 0.0;
  int matching = 0;
  for (int i = 0; i < b.n; ++i) matching += lobeMatches(b.lobes[i], flags) ? 1 : 0;
  if (matching == 0) return mks1(0.0);
  const int which = min((int)floor(comp * matching), matching - 1);
  int chosen = 0, count = which;
  for (int i = 0; i < b.n; ++i)
    if (lobeMatches(b.lobes[i], flags) && count-- == 0) { chosen = i; break; }
  const GLobe& lc = b.lobes[chosen];
  const int type = lobeTypeOf(lc);
  const V3 wo = bsdfToLocal(b, woW);
  V3 wi = V3{0.f, 0.f, 0.f};
  Spec f = lobeSampleF(lc, wo, &wi, (double)u0, (double)u1, pdfOut);
  if (*pdfOut == 0.0) return mks1(0.0);
  *sampledType = type;
  *wiW = bsdfToWorld(b, wi);
  if (!((type & BSDF_SPECULAR) != 0) && matching > 1)
    for (int i = 0; i < b.n; ++i) {
      const GLobe& l = b.lobes[i];
      if (i != chosen && lobeMatches(l, flags)) *pdfOut += lobePdf(l, wo, wi);
    }
  if (matching > 1) *pdfOut /= matching;
  if ((type & BSDF_SPECULAR) == 0) {
    f = mks1(0.0);
    if (Dot(*wiW, b.ng) * Dot(woW, b.ng) > 0) flags &= ~BSDF_TRANSMISSION;
    else flags &= ~BSDF_REFLECTION;
    for (int i = 0; i < b.n; ++i) {
      const GLobe& l = b.lobes[i];
      if (lobeMatches(l, flags)) f = f + lobeF(l, wo, wi);
    }
  }
  return f;
}

template <bool GENERAL> struct BsdfOf { typedef Bsdf type; };
template <> struct BsdfOf<true> { typedef BsdfG type; };
template <bool GENERAL, bool EXTRA>
struct MakeBsdf;
template <bool EXTRA>
struct MakeBsdf<false, EXTRA> {
  static __device__ inline Bsdf make(const RenderScene& rs, uint32_t prim, const ShapeHit& h, const V3& o, const V3& d) {
    return makeBsdf<EXTRA>(rs, prim, h, o, d);
  }
};
template <bool EXTRA>
struct MakeBsdf<true, EXTRA> {
  static __device__ inline BsdfG make(const RenderScene& rs, uint32_t prim, const ShapeHit& h, const V3& o, const V3& d) {
    return makeBsdfG<EXTRA>(rs, prim, h, o, d);
  }
};
template <bool GENERAL, bool EXTRA = (DRT_EXTRA != 0)>
static __device__ inline typename BsdfOf<GENERAL>::type makeBsdfT(const RenderScene& rs, uint32_t prim, const ShapeHit& h, const V3& o,
                                                                  const V3& d) {
  return MakeBsdf<GENERAL, EXTRA>::make(rs, prim, h, o, d);
}

// Materials with a program (drt_set_material_programs: textures that read the hit point, bump maps): the BSDF the texture pass
// (texture_kernels.cu) built for the slot's current vertex replaces the flattened one — frame = dgShading after Material.Bump
// (bsdf.dart:45-51), lobes as the material's getBSDF added them.  ng stays the geometric normal.
static __device__ inline void applyHitBsdf(const RenderScene& rs, const Wavefront& wf, uint32_t slot, BsdfG* b) {
#if DRT_EXTRA
  if (rs.nPrograms <= 0) return;
  const int n = wf.hitCount[slot];
  if (n < 0) return;
  const size_t cap = wf.cap;
  b->nn = V3{wf.hitFrame[slot], wf.hitFrame[cap + slot], wf.hitFrame[2 * cap + slot]};
  b->sn = V3{wf.hitFrame[3 * cap + slot], wf.hitFrame[4 * cap + slot], wf.hitFrame[5 * cap + slot]};
  b->tn = Cross(b->nn, b->sn);
  b->lobes = wf.hitLobes + (size_t)slot * 8;
  b->n = n;
#endif
}
static __device__ inline void applyHitBsdf(const RenderScene&, const Wavefront&, uint32_t, Bsdf*) {}

// ---- lights (diffuse_area_light.dart:44-70, shape_set.dart:43-96, shape.dart:100-121,
// triangle.dart:265-269,366-383, sphere.dart:243-311, point_light.dart:41-47) --------------------------
static __device__ inline Spec lightRadiance(const GLight& l) { return Spec{l.L[0], l.L[1], l.L[2]}; }
static __device__ inline Spec areaL(const GLight& l, const V3& n, const V3& w) { return Dot(n, w) > 0.0 ? lightRadiance(l) : mks1(0.0); }

static DRT_HD inline double triArea(const TriVerts& t) { return 0.5 * Length(Cross(t.p2 - t.p1, t.p3 - t.p1)); }

static __device__ inline V3 sphereCenter(const GSphere& s) {
  float o2w[16];
  for (int i = 0; i < 12; ++i) o2w[i] = s.o2w[i];
  for (int i = 0; i < 4; ++i) o2w[12 + i] = s.o2wRow3[i];
  return XfPoint(o2w, V3{0.f, 0.f, 0.f});
}

// Shape.intersect on one shape of a ShapeSet: tHit and dg.nn (all the light code reads).
static __device__ inline bool lightShapeIntersect(const RenderScene& rs, const GLightShape& ls, const V3& o, const V3& d,
                                                  double mint, double maxt, double* tOut, V3* nnOut) {
  if (ls.prim < rs.ntris) {
    TriVerts tv;
    tv.p1 = V3{ls.p1[0], ls.p1[1], ls.p1[2]}; tv.p2 = V3{ls.p2[0], ls.p2[1], ls.p2[2]}; tv.p3 = V3{ls.p3[0], ls.p3[1], ls.p3[2]};
    if (!triIntersectT(tv, o, d, mint, maxt, tOut)) return false;
    *nnOut = V3{ls.nn[0], ls.nn[1], ls.nn[2]};
    return true;
  }
  ShapeHit h;
  if (!shapeIntersectCold(rs, ls.prim, o, d, mint, maxt, &h)) return false;
  *tOut = h.t;
  *nnOut = h.nn;
  return true;
}

// Shape.sample(p, u1, u2) -> point on the shape and its normal
static __device__ __noinline__ void quadricSample2Cold(const RenderScene& rs, uint32_t prim, V3 p, double u1, double u2, V3* ptOut, V3* ns);
static __device__ inline V3 quadricSample2(const RenderScene& rs, uint32_t prim, const V3& p, double u1, double u2, V3* ns);
static __device__ inline V3 shapeSample2(const RenderScene& rs, const GLightShape& ls, const V3& p, double u1, double u2, V3* ns) {
  const uint32_t prim = ls.prim;
  if (prim < rs.ntris) {  // shape.dart:96-98 -> triangle.dart:366-383, UniformSampleTriangle montecarlo.dart:215-220
    double su1 = sqrt(u1);
    double b1 = 1.0 - su1, b2 = u2 * su1;
    V3 p1 = V3{ls.p1[0], ls.p1[1], ls.p1[2]}, p2 = V3{ls.p2[0], ls.p2[1], ls.p2[2]}, p3 = V3{ls.p3[0], ls.p3[1], ls.p3[2]};
    V3 pt = p1 * b1 + p2 * b2 + p3 * (1.0 - b1 - b2);
    *ns = V3{ls.ns[0], ls.ns[1], ls.ns[2]};
    return pt;
  }
  // sphere / disk / cylinder light shapes: out of line, like their intersection test (shapeIntersectCold), so that the
  // kernels of a triangle-light scene (config 4) do not carry them; results through temporaries
#ifdef DRT_QSAMPLE_INLINE
  return quadricSample2(rs, prim, p, u1, u2, ns);
#else
  V3 pt, nq;
  quadricSample2Cold(rs, prim, p, u1, u2, &pt, &nq);
  *ns = nq;
  return pt;
#endif
}
static __device__ inline V3 quadricSample2(const RenderScene& rs, uint32_t prim, const V3& p, double u1, double u2, V3* ns) {
  const GSphere& s = rs.ts.spheres[prim - rs.ntris];
  const bool rev = primReverse(rs, prim);
  float o2w[16], w2o[16];
  for (int i = 0; i < 12; ++i) { o2w[i] = s.o2w[i]; w2o[i] = s.w2o[i]; }
  for (int i = 0; i < 4; ++i) { o2w[12 + i] = s.o2wRow3[i]; w2o[12 + i] = s.w2oRow3[i]; }
  if (s.shape == 1) {  // shape.dart:96-98 -> disk.dart:147-159
    double t0, t1;
    ConcentricSampleDisk(u1, u2, &t0, &t1);
    V3 pd = mkv(t0 * s.radius, t1 * s.radius, s.height);
    V3 n = XfNormal(w2o, V3{0.f, 0.f, 1.f});
    n = n / Length(n);
    if (rev) n = n * -1.0;
    *ns = n;
    return XfPoint(o2w, pd);
  }
  if (DRT_EXTRA && s.shape == 2) return cylinderSampleCold(s, rev, u1, u2, ns);
  // sphere.dart:261-297
  V3 Pcenter = XfPoint(o2w, V3{0.f, 0.f, 0.f});
  V3 wc = Normalize(Pcenter - p);
  V3 wcX, wcY;
  CoordinateSystem(wc, &wcX, &wcY);
  if (DistanceSquared(p, Pcenter) - s.radius * s.radius < 1.0e-4) {  // sphere.dart:247-259
    V3 ps = V3{0.f, 0.f, 0.f} + UniformSampleSphere(u1, u2) * s.radius;
    V3 n = XfNormal(w2o, ps);
    n = n / Length(n);
    if (rev) n = -n;
    *ns = n;
    return XfPoint(o2w, ps);
  }
  double sinThetaMax2 = s.radius * s.radius / DistanceSquared(p, Pcenter);
  double cosThetaMax = sqrt(fmax(0.0, 1.0 - sinThetaMax2));
  V3 rd = UniformSampleCone2(u1, u2, cosThetaMax, wcX, wcY, wc);
  double thit;
  ShapeHit h;
  if (shapeIntersectCold(rs, prim, p, rd, 1.0e-3, CUDART_INF, &h)) thit = h.t;
  else thit = Dot(Pcenter - p, Normalize(rd));
  V3 ps = RayAt(p, rd, thit);
  V3 n = Normalize(ps - Pcenter);
  if (rev) n = -n;
  *ns = n;
  return ps;
}

static __device__ __noinline__ void quadricSample2Cold(const RenderScene& rs, uint32_t prim, V3 p, double u1, double u2, V3* ptOut, V3* ns) {
  *ptOut = quadricSample2(rs, prim, p, u1, u2, ns);
}

static __device__ inline double shapePdf2(const RenderScene& rs, const GLightShape& ls, const V3& p, const V3& wi) {
  const uint32_t prim = ls.prim;
  if (prim >= rs.ntris && rs.ts.spheres[prim - rs.ntris].shape == 0) {  // sphere.dart:299-311
    const GSphere& s = rs.ts.spheres[prim - rs.ntris];
    V3 Pcenter = sphereCenter(s);
    if (!(DistanceSquared(p, Pcenter) - s.radius * s.radius < 1.0e-4)) {
      double sinThetaMax2 = s.radius * s.radius / DistanceSquared(p, Pcenter);
      double cosThetaMax = sqrt(fmax(0.0, 1.0 - sinThetaMax2));
      return UniformConePdf(cosThetaMax);
    }
  }
  double t;  // shape.dart:100-121
  V3 nn;
  if (!lightShapeIntersect(rs, ls, p, wi, 1.0e-3, CUDART_INF, &t, &nn)) return 0.0;
  double pdf = DistanceSquared(p, RayAt(p, wi, t)) / (AbsDot(nn, -wi) * ls.area);
  if (isinf(pdf)) pdf = 0.0;
  return pdf;
}

static __device__ inline double shapeSetPdf(const RenderScene& rs, const GLight& l, const V3& p, const V3& wi) {  // shape_set.dart:81-89
  double pdf = 0.0;
  for (uint32_t i = 0; i < l.nShapes; ++i) {
    const GLightShape& ls = rs.lightShapes[l.shapeOffset + i];
    pdf += ls.area * shapePdf2(rs, ls, p, wi);
  }
  return pdf / l.area;
}

// Distribution1D.sampleDiscrete (montecarlo.dart:82-92): upper_bound over the float32 cdf
static __device__ inline int sampleDiscrete(const float* cdf, int count, double u) {
  int lo = 0, hi = count + 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (u < (double)cdf[mid]) hi = mid;
    else lo = mid + 1;
  }
  int r = lo - 1;
  return r < 0 ? 0 : r;
}

static __device__ inline V3 shapeSetSample(const RenderScene& rs, const GLight& l, const V3& p, float u0, float u1, double comp,
                                           V3* Ns) {  // shape_set.dart:53-79
  int sn = sampleDiscrete(rs.lightCdf + l.cdfOffset, (int)l.nShapes, comp) % (int)l.nShapes;
  V3 pt = shapeSample2(rs, rs.lightShapes[l.shapeOffset + sn], p, u0, u1, Ns);
  V3 rd = pt - p;
  double thit = 1.0;
  bool anyHit = false;
  V3 nnLast = *Ns;
  for (uint32_t i = 0; i < l.nShapes; ++i) {
    double t;
    V3 nn;
    if (lightShapeIntersect(rs, rs.lightShapes[l.shapeOffset + i], p, rd, 1.0e-3, CUDART_INF, &t, &nn)) {
      anyHit = true;
      thit = t;
      nnLast = nn;
    }
  }
  if (anyHit) *Ns = nnLast;
  return RayAt(p, rd, thit);
}

// ---- InfiniteAreaLight (infinite_area_light.dart) ----------------------------------------------------------------------------
// MIPMap.lookup with width 0 is triangle(0, s, t) (mipmap.dart:206-212,341-355; TEXTURE_REPEAT, Dart's % is never negative)
static __device__ inline Spec envTexel(const float* tex, int W, int H, long long s, long long t) {
  s = ((s % W) + W) % W;
  t = ((t % H) + H) % H;
  const float* q = tex + 3 * (size_t)(t * W + s);
  return Spec{q[0], q[1], q[2]};
}
static __device__ __noinline__ void mapLookupCold(const RenderScene& rs, const GLight& l, double u, double v, Spec* out) {
  const float* tex = rs.envData + l.envOffset;
  const int W = l.mapW, H = l.mapH;
  double s = u * W - 0.5, t = v * H - 0.5;
  const long long s0 = (long long)floor(s), t0 = (long long)floor(t);
  const double ds = s - s0, dt = t - t0;
  *out = envTexel(tex, W, H, s0, t0) * ((1.0 - ds) * (1.0 - dt)) + envTexel(tex, W, H, s0, t0 + 1) * ((1.0 - ds) * dt) +
         envTexel(tex, W, H, s0 + 1, t0) * (ds * (1.0 - dt)) + envTexel(tex, W, H, s0 + 1, t0 + 1) * (ds * dt);
}
static __device__ inline void envRadianceCold(const RenderScene& rs, const GLight& l, double u, double v, Spec* out) {
  Spec r;
  mapLookupCold(rs, l, u, v, &r);
  *out = r * lightRadiance(l);  // _radiance = lookup * L (infinite_area_light.dart:240-242)
}
static __device__ inline double SphericalTheta(const V3& v) { return acos(clampD((double)v.z, -1.0, 1.0)); }  // vector.dart:185-187
static __device__ inline V3 xf3(const float* m, const V3& w) {
  return mkv((double)m[0] * w.x + (double)m[1] * w.y + (double)m[2] * w.z, (double)m[3] * w.x + (double)m[4] * w.y + (double)m[5] * w.z,
             (double)m[6] * w.x + (double)m[7] * w.y + (double)m[8] * w.z);
}
// Light.Le(ray) of an infinite light (:86-91)
static __device__ inline Spec infiniteLe(const RenderScene& rs, const GLight& l, const V3& d) {
  const V3 wh = Normalize(xf3(l.w2l, d));
  const double s = SphericalPhi(wh) * DRT_INV_TWOPI, t = SphericalTheta(wh) * DRT_INV_PI;
  Spec r;
  envRadianceCold(rs, l, s, t, &r);
  return r;
}
// Distribution1D.sampleContinuous (montecarlo.dart:50-80) over float32 func / cdf tables
static __device__ inline double dist1DSample(const float* func, const float* cdf, double funcInt, int count, double u, double* pdf, int* off) {
  int lo = 0, hi = count + 1;  // upper_bound(cdf, u, last: count + 1)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (u < (double)cdf[mid]) hi = mid;
    else lo = mid + 1;
  }
  int offset = lo - 1 < 0 ? 0 : lo - 1;
  if (offset == count) offset = count - 1;
  if (off) *off = offset;
  const double dc = (double)cdf[offset + 1] - (double)cdf[offset];
  double du = 0.0;
  if (dc != 0.0) du = (u - (double)cdf[offset]) / dc;
  *pdf = (double)func[offset] / funcInt;
  return (offset + du) / count;
}
struct EnvTables {
  const float *condFunc, *condCdf, *condInt, *margFunc, *margCdf;
  double margInt;
};
static __device__ inline EnvTables envTables(const RenderScene& rs, const GLight& l) {
  const size_t W = l.mapW, H = l.mapH;
  EnvTables t;
  t.condFunc = rs.envData + l.envOffset + 3 * W * H;
  t.condCdf = t.condFunc + W * H;
  t.condInt = t.condCdf + H * (W + 1);
  t.margFunc = t.condInt + H;
  t.margCdf = t.margFunc + H;
  t.margInt = (double)t.margCdf[H + 1];
  return t;
}
// InfiniteAreaLight.sampleLAtPoint (:93-131): direction, pdf and radiance; false when the sample carries nothing
static __device__ __noinline__ bool infiniteSampleCold(const RenderScene& rs, const GLight& l, double u0, double u1, V3* wi, double* pdf,
                                                       Spec* Li) {
  const EnvTables t = envTables(rs, l);
  const int W = l.mapW, H = l.mapH;
  double pdfs1, pdfs0;
  int v;
  const double uvv = dist1DSample(t.margFunc, t.margCdf, t.margInt, H, u1, &pdfs1, &v);
  const double uvu = dist1DSample(t.condFunc + (size_t)v * W, t.condCdf + (size_t)v * (W + 1), (double)t.condInt[v], W, u0, &pdfs0, nullptr);
  const double mapPdf = pdfs0 * pdfs1;
  *pdf = 0.0;
  if (mapPdf == 0.0) return false;
  const double theta = uvv * DRT_PI, phi = uvu * 2.0 * DRT_PI;
  const double costheta = cos(theta), sintheta = sin(theta), sinphi = sin(phi), cosphi = cos(phi);
  *wi = xf3(l.l2w, mkv(sintheta * cosphi, sintheta * sinphi, costheta));
  if (sintheta == 0.0) *pdf = 0.0;
  else *pdf = mapPdf / (2.0 * DRT_PI * DRT_PI * sintheta);
  envRadianceCold(rs, l, uvu, uvv, Li);
  return true;
}
// InfiniteAreaLight.pdf (:244-259) with Distribution2D.pdf (montecarlo.dart:250-263)
static __device__ __noinline__ double infinitePdfCold(const RenderScene& rs, const GLight& l, V3 w) {
  const V3 wi = xf3(l.w2l, w);
  const double theta = SphericalTheta(wi), phi = SphericalPhi(wi);
  const double sintheta = sin(theta);
  if (sintheta == 0.0) return 0.0;
  const EnvTables t = envTables(rs, l);
  const int W = l.mapW, H = l.mapH;
  const double u = phi * DRT_INV_TWOPI, v = theta * DRT_INV_PI;
  const int iu = (int)fmin(fmax(trunc(u * W), 0.0), (double)(W - 1)), iv = (int)fmin(fmax(trunc(v * H), 0.0), (double)(H - 1));
  const double ci = (double)t.condInt[iv];
  if (ci * t.margInt == 0.0) return 0.0;
  const double p = ((double)t.condFunc[(size_t)iv * W + iu] * (double)t.margFunc[iv]) / (ci * t.margInt);
  return p / (2.0 * DRT_PI * DRT_PI * sintheta);
}
// ProjectionLight.projection(w) (projection_light.dart:115-139) / GoniometricLight.scale(w) (goniometric_light.dart:71-86) times
// the intensity over the squared distance, as sampleLAtPoint writes it
static __device__ __noinline__ void mappedPointLightCold(const RenderScene& rs, const GLight& l, V3 w, double dist2, Spec* out) {
  const Spec I = lightRadiance(l);
  if (l.kind == 5) {
    const V3 wl = xf3(l.w2l, w);
    Spec proj = mks1(0.0);
    if (!((double)wl.z < l.hither)) {
      const V3 Pl = XfPoint(l.proj, wl);
      if (!((double)Pl.x < l.screen[0] || (double)Pl.x > l.screen[1] || (double)Pl.y < l.screen[2] || (double)Pl.y > l.screen[3])) {
        if (l.mapW == 0) proj = mks1(1.0);
        else mapLookupCold(rs, l, ((double)Pl.x - l.screen[0]) / (l.screen[1] - l.screen[0]),
                           ((double)Pl.y - l.screen[2]) / (l.screen[3] - l.screen[2]), &proj);
      }
    }
    *out = I * proj / dist2;
    return;
  }
  V3 wp = Normalize(xf3(l.w2l, w));
  const float tmp = wp.y;
  wp.y = wp.z;
  wp.z = tmp;
  const double theta = SphericalTheta(wp), phi = SphericalPhi(wp);
  if (l.mapW == 0) { *out = I * 1.0 / dist2; return; }
  Spec sc;
  mapLookupCold(rs, l, phi * DRT_INV_TWOPI, theta * DRT_INV_PI, &sc);
  *out = I * sc / dist2;
}

// Light.pdf(p, wi) for the non-delta lights
static __device__ inline double lightPdfAny(const RenderScene& rs, const GLight& l, const V3& p, const V3& wi) {
  if (DRT_EXTRA && l.kind == 4) return infinitePdfCold(rs, l, wi);
  return shapeSetPdf(rs, l, p, wi);
}

// What Integrator.EstimateDirect (integrator.dart:119-185) needs traced before it can finish: the
// shadow ray of the light sample and the closest-hit ray of the BSDF sample, each with the
// contribution it carries if the query comes out right.
struct DirectWork {
  bool hasShadow, hasMis;
  V3 shO, shD;
  double shMin, shMax;
  Spec shContribution;  // f * Li * (|wi.n| * weight / lightPdf)
  V3 misD;
  Spec misF;
  double misScale;  // |wi.n| * weight / bsdfPdf
};

template <typename BSDF>
static __device__ inline void estimateDirectSetup(const RenderScene& rs, int lightIndex, const V3& p, const V3& n, const V3& wo,
                                                  double rayEps, const BSDF& bsdf, float lu0, float lu1, double lcomp, float bu0,
                                                  float bu1, double bcomp, int flags, DirectWork* w) {
  const GLight l = rs.lights[lightIndex];
  w->hasShadow = false;
  w->hasMis = false;
  V3 wi;
  double lightPdf = 0.0, bPdf = 0.0;
  Spec Li;
  V3 segTo;
  double eps2;
  const bool infinite = DRT_EXTRA && l.kind == 4;
  const bool delta = l.kind != 0 && !infinite;  // isDeltaLight: point, distant, spot
  const bool distant = l.kind == 2 || infinite;  // the shadow ray runs to infinity (visibility_tester.dart:31-33)
  if (infinite) {  // infinite_area_light.dart:93-131
    segTo = p;
    eps2 = 0.0;
    Li = mks1(0.0);
    infiniteSampleCold(rs, l, lu0, lu1, &wi, &lightPdf, &Li);
  } else if (distant) {  // distant_light.dart:41-48
    wi = V3{l.pos[0], l.pos[1], l.pos[2]};
    lightPdf = 1.0;
    segTo = p;
    eps2 = 0.0;
    Li = lightRadiance(l);
  } else if (delta) {  // point_light.dart:41-47, spot_light.dart:36-70
    V3 pos = V3{l.pos[0], l.pos[1], l.pos[2]};
    wi = Normalize(pos - p);
    lightPdf = 1.0;
    segTo = pos;
    eps2 = 0.0;
    if (l.kind == 1) {
      Li = lightRadiance(l) / DistanceSquared(pos, p);
    } else if (DRT_EXTRA && l.kind >= 5) {
      mappedPointLightCold(rs, l, -wi, DistanceSquared(pos, p), &Li);
    } else {
      const V3 w = -wi;
      const V3 wl = Normalize(mkv((double)l.w2l[0] * w.x + (double)l.w2l[1] * w.y + (double)l.w2l[2] * w.z,
                                  (double)l.w2l[3] * w.x + (double)l.w2l[4] * w.y + (double)l.w2l[5] * w.z,
                                  (double)l.w2l[6] * w.x + (double)l.w2l[7] * w.y + (double)l.w2l[8] * w.z));
      const double costheta = wl.z;
      double falloff;
      if (costheta < l.cosTotalWidth) falloff = 0.0;
      else if (costheta > l.cosFalloffStart) falloff = 1.0;
      else {
        const double dl = (costheta - l.cosTotalWidth) / (l.cosFalloffStart - l.cosTotalWidth);
        falloff = dl * dl * dl * dl;
      }
      Li = lightRadiance(l) * falloff / DistanceSquared(pos, p);
    }
  } else {  // diffuse_area_light.dart:59-70
    V3 ns;
    V3 ps = shapeSetSample(rs, l, p, lu0, lu1, lcomp, &ns);
    wi = Normalize(ps - p);
    lightPdf = shapeSetPdf(rs, l, p, wi);
    segTo = ps;
    eps2 = 1.0e-3;
    Li = areaL(l, ns, -wi);
  }
  if (lightPdf > 0.0 && !IsBlack(Li)) {
    Spec f = bsdfF(bsdf, wo, wi, flags);
    if (!IsBlack(f)) {
      w->hasShadow = true;
      w->shO = p;
      w->shMin = rayEps;
      if (distant) {
        w->shD = wi;
        w->shMax = CUDART_INF;
      } else {
        double dist = Distance(p, segTo);  // visibility_tester.dart:26-29
        w->shD = (segTo - p) / dist;
        w->shMax = dist * (1.0 - eps2);
      }
      Li = Li * mks1(1.0);  // transmittance
      if (delta) {
        w->shContribution = f * Li * (AbsDot(wi, n) / lightPdf);
      } else {
        bPdf = bsdfPdf(bsdf, wo, wi, flags);
        double weight = PowerHeuristic(1, lightPdf, 1, bPdf);
        w->shContribution = f * Li * ((AbsDot(wi, n) * weight / lightPdf));
      }
    }
  }
  if (!delta) {
    int sampledType = 0;
    Spec f = bsdfSampleF(bsdf, wo, &wi, bu0, bu1, bcomp, &bPdf, flags, &sampledType);
    if (!IsBlack(f) && bPdf > 0.0) {
      double weight = 1.0;
      if ((sampledType & BSDF_SPECULAR) == 0) {
        lightPdf = lightPdfAny(rs, l, p, wi);
        if (lightPdf == 0.0) return;
        weight = PowerHeuristic(1, bPdf, 1, lightPdf);
      }
      w->hasMis = true;
      w->misD = wi;
      w->misF = f;
      w->misScale = AbsDot(wi, n) * weight / bPdf;
    }
  }
}

// WhittedIntegrator.Li's light loop body (whitted_integrator.dart:46-63): one light sample, every BxDF, no multiple
// importance sampling.  Fills the shadow-ray part of DirectWork.
template <typename BSDF>
static __device__ inline void whittedLightSetup(const RenderScene& rs, int lightIndex, const V3& p, const V3& n, const V3& wo,
                                                double rayEps, const BSDF& bsdf, float lu0, float lu1, double lcomp, DirectWork* w) {
  const GLight l = rs.lights[lightIndex];
  w->hasShadow = false;
  w->hasMis = false;
  V3 wi, segTo = p;
  double lightPdf = 1.0, eps2 = 0.0;
  Spec Li;
  const bool infinite = DRT_EXTRA && l.kind == 4;
  const bool distant = l.kind == 2 || infinite;
  if (infinite) {  // infinite_area_light.dart:93-131
    Li = mks1(0.0);
    infiniteSampleCold(rs, l, lu0, lu1, &wi, &lightPdf, &Li);
  } else if (distant) {  // distant_light.dart:41-48
    wi = V3{l.pos[0], l.pos[1], l.pos[2]};
    Li = lightRadiance(l);
  } else if (l.kind != 0) {  // point_light.dart:41-47, spot_light.dart:36-70
    const V3 pos = V3{l.pos[0], l.pos[1], l.pos[2]};
    wi = Normalize(pos - p);
    segTo = pos;
    if (l.kind == 1) {
      Li = lightRadiance(l) / DistanceSquared(pos, p);
    } else if (DRT_EXTRA && l.kind >= 5) {
      mappedPointLightCold(rs, l, -wi, DistanceSquared(pos, p), &Li);
    } else {
      const V3 wn = -wi;
      const V3 wl = Normalize(mkv((double)l.w2l[0] * wn.x + (double)l.w2l[1] * wn.y + (double)l.w2l[2] * wn.z,
                                  (double)l.w2l[3] * wn.x + (double)l.w2l[4] * wn.y + (double)l.w2l[5] * wn.z,
                                  (double)l.w2l[6] * wn.x + (double)l.w2l[7] * wn.y + (double)l.w2l[8] * wn.z));
      const double costheta = wl.z;
      double falloff;
      if (costheta < l.cosTotalWidth) falloff = 0.0;
      else if (costheta > l.cosFalloffStart) falloff = 1.0;
      else {
        const double dl = (costheta - l.cosTotalWidth) / (l.cosFalloffStart - l.cosTotalWidth);
        falloff = dl * dl * dl * dl;
      }
      Li = lightRadiance(l) * falloff / DistanceSquared(pos, p);
    }
  } else {  // diffuse_area_light.dart:59-70
    V3 ns;
    const V3 ps = shapeSetSample(rs, l, p, lu0, lu1, lcomp, &ns);
    wi = Normalize(ps - p);
    lightPdf = shapeSetPdf(rs, l, p, wi);
    segTo = ps;
    eps2 = 1.0e-3;
    Li = areaL(l, ns, -wi);
  }
  if (IsBlack(Li) || lightPdf == 0.0) return;
  const Spec f = bsdfF(bsdf, wo, wi, BSDF_ALL);
  if (IsBlack(f)) return;
  w->hasShadow = true;
  w->shO = p;
  w->shMin = rayEps;
  if (distant) {
    w->shD = wi;
    w->shMax = CUDART_INF;
  } else {
    const double dist = Distance(p, segTo);  // visibility_tester.dart:26-29
    w->shD = (segTo - p) / dist;
    w->shMax = dist * (1.0 - eps2);
  }
  w->shContribution = f * Li * AbsDot(wi, n) * mks1(1.0) / lightPdf;  // :59-61, transmittance == 1
}

}  // namespace drt
