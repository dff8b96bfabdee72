// AnimatedTransform of the reference for the GPU path (host + device): TransformedPrimitive.intersect / intersectP evaluate
// worldToPrimitive.interpolate(ray.time) per ray (lib/core/primitive/transformed_primitive.dart:30-62), so the interpolation runs on
// the device with the reference's arithmetic — float32 storage in Matrix4x4 / Vector, binary64 expressions, no contraction:
//   lib/core/matrix4x4.dart:190-210 (Transpose, Mul), :242-357 (determinant, invert)
//   lib/core/quaternion.dart:39-77 (fromMatrix), :120-149 (toTransform), :151-173 (Slerp, Dot, Normalize)
//   lib/core/transform.dart:31-35 (Transform(m) inverts), :83-86 (operator *), :214-227 (Translate)
//   lib/core/animated_transform.dart:35-43 (constructor), :61-105 (Decompose), :107-136 (interpolate), :183-200 (motionBounds)
// Decompose and motionBounds run once per instance on the host (as the reference's constructors do); interpolate on both sides.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "gpu_types.h"

namespace drt {

struct M4 {
  float d[16];
};

// One TransformedPrimitive: its AnimatedTransform after the constructor ran, and the object it wraps.
struct GInstance {
  M4 startM, startInv, endM, endInv;
  double startTime, endTime;
  double Rw[2];       // quaternion w (a Dart double); v is a float32 Vector
  float Rv[2][3];
  float T[2][3];
  M4 S[2];
  int32_t animated;   // actuallyAnimated: the two Transforms differ element-wise (transform.dart:67-69, matrix4x4.dart:86-93)
  int32_t object;
};

// The `primitive` of a TransformedPrimitive: a BVHAccel over the object's primitives (binary nodes appended to TraceScene::nodes,
// leaf records to TraceScene::prims) or, single != 0, one GeometricPrimitive (rootRef = the leaf reference of its record; no box).
struct GObject {
  float rootMin[3], rootMax[3];
  int32_t rootRef;
  int32_t single;
};

static DRT_HD inline float anim_f32(double v) { return (float)v; }

static DRT_HD inline M4 m4Identity() {
  M4 r;
  for (int i = 0; i < 16; ++i) r.d[i] = (i % 5 == 0) ? 1.f : 0.f;
  return r;
}
static DRT_HD inline M4 m4Transpose(const M4& m) {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) r.d[4 * i + j] = m.d[4 * j + i];
  return r;
}
static DRT_HD inline M4 m4Mul(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0, k = 0; i < 4; ++i, k += 4)
    for (int j = 0; j < 4; ++j)
      r.d[k + j] = anim_f32((double)a.d[k] * (double)b.d[j] + (double)a.d[k + 1] * (double)b.d[4 + j] +
                            (double)a.d[k + 2] * (double)b.d[8 + j] + (double)a.d[k + 3] * (double)b.d[12 + j]);
  return r;
}
// invert(): the reference reads the elements with the indices of the transpose (its n12 is data[4]) and stores the cofactors
// transposed back; every chain of products is one left-to-right binary64 expression, each result one float32 store.
static DRT_HD inline M4 m4Inverse(const M4& m) {
  const double a11 = m.d[0], a12 = m.d[4], a13 = m.d[8], a14 = m.d[12];
  const double a21 = m.d[1], a22 = m.d[5], a23 = m.d[9], a24 = m.d[13];
  const double a31 = m.d[2], a32 = m.d[6], a33 = m.d[10], a34 = m.d[14];
  const double a41 = m.d[3], a42 = m.d[7], a43 = m.d[11], a44 = m.d[15];
  const double det = (a14 * a23 * a32 * a41) - (a13 * a24 * a32 * a41) - (a14 * a22 * a33 * a41) + (a12 * a24 * a33 * a41) +
                     (a13 * a22 * a34 * a41) - (a12 * a23 * a34 * a41) - (a14 * a23 * a31 * a42) + (a13 * a24 * a31 * a42) +
                     (a14 * a21 * a33 * a42) - (a11 * a24 * a33 * a42) - (a13 * a21 * a34 * a42) + (a11 * a23 * a34 * a42) +
                     (a14 * a22 * a31 * a43) - (a12 * a24 * a31 * a43) - (a14 * a21 * a32 * a43) + (a11 * a24 * a32 * a43) +
                     (a12 * a21 * a34 * a43) - (a11 * a22 * a34 * a43) - (a13 * a22 * a31 * a44) + (a12 * a23 * a31 * a44) +
                     (a13 * a21 * a32 * a44) - (a11 * a23 * a32 * a44) - (a12 * a21 * a33 * a44) + (a11 * a22 * a33 * a44);
  if (det == 0.0) return m;
  const double k = 1.0 / det;
  M4 r;
  r.d[0] = anim_f32((a23 * a34 * a42 - a24 * a33 * a42 + a24 * a32 * a43 - a22 * a34 * a43 - a23 * a32 * a44 + a22 * a33 * a44) * k);
  r.d[4] = anim_f32((a14 * a33 * a42 - a13 * a34 * a42 - a14 * a32 * a43 + a12 * a34 * a43 + a13 * a32 * a44 - a12 * a33 * a44) * k);
  r.d[8] = anim_f32((a13 * a24 * a42 - a14 * a23 * a42 + a14 * a22 * a43 - a12 * a24 * a43 - a13 * a22 * a44 + a12 * a23 * a44) * k);
  r.d[12] = anim_f32((a14 * a23 * a32 - a13 * a24 * a32 - a14 * a22 * a33 + a12 * a24 * a33 + a13 * a22 * a34 - a12 * a23 * a34) * k);
  r.d[1] = anim_f32((a24 * a33 * a41 - a23 * a34 * a41 - a24 * a31 * a43 + a21 * a34 * a43 + a23 * a31 * a44 - a21 * a33 * a44) * k);
  r.d[5] = anim_f32((a13 * a34 * a41 - a14 * a33 * a41 + a14 * a31 * a43 - a11 * a34 * a43 - a13 * a31 * a44 + a11 * a33 * a44) * k);
  r.d[9] = anim_f32((a14 * a23 * a41 - a13 * a24 * a41 - a14 * a21 * a43 + a11 * a24 * a43 + a13 * a21 * a44 - a11 * a23 * a44) * k);
  r.d[13] = anim_f32((a13 * a24 * a31 - a14 * a23 * a31 + a14 * a21 * a33 - a11 * a24 * a33 - a13 * a21 * a34 + a11 * a23 * a34) * k);
  r.d[2] = anim_f32((a22 * a34 * a41 - a24 * a32 * a41 + a24 * a31 * a42 - a21 * a34 * a42 - a22 * a31 * a44 + a21 * a32 * a44) * k);
  r.d[6] = anim_f32((a14 * a32 * a41 - a12 * a34 * a41 - a14 * a31 * a42 + a11 * a34 * a42 + a12 * a31 * a44 - a11 * a32 * a44) * k);
  r.d[10] = anim_f32((a12 * a24 * a41 - a14 * a22 * a41 + a14 * a21 * a42 - a11 * a24 * a42 - a12 * a21 * a44 + a11 * a22 * a44) * k);
  r.d[14] = anim_f32((a14 * a22 * a31 - a12 * a24 * a31 - a14 * a21 * a32 + a11 * a24 * a32 + a12 * a21 * a34 - a11 * a22 * a34) * k);
  r.d[3] = anim_f32((a23 * a32 * a41 - a22 * a33 * a41 - a23 * a31 * a42 + a21 * a33 * a42 + a22 * a31 * a43 - a21 * a32 * a43) * k);
  r.d[7] = anim_f32((a12 * a33 * a41 - a13 * a32 * a41 + a13 * a31 * a42 - a11 * a33 * a42 - a12 * a31 * a43 + a11 * a32 * a43) * k);
  r.d[11] = anim_f32((a13 * a22 * a41 - a12 * a23 * a41 - a13 * a21 * a42 + a11 * a23 * a42 + a12 * a21 * a43 - a11 * a22 * a43) * k);
  r.d[15] = anim_f32((a12 * a23 * a31 - a13 * a22 * a31 + a13 * a21 * a32 - a11 * a23 * a32 - a12 * a21 * a33 + a11 * a22 * a33) * k);
  return r;
}

struct Quat4 {
  float x, y, z;  // v: a Vector
  double w;
};
static DRT_HD inline double quatDot(const Quat4& a, const Quat4& b) {
  return ((double)a.x * (double)b.x + (double)a.y * (double)b.y + (double)a.z * (double)b.z) + a.w * b.w;
}
static DRT_HD inline Quat4 quatScale(const Quat4& q, double f) {
  return Quat4{anim_f32((double)q.x * f), anim_f32((double)q.y * f), anim_f32((double)q.z * f), q.w * f};
}
static DRT_HD inline Quat4 quatAdd(const Quat4& a, const Quat4& b) {
  return Quat4{anim_f32((double)a.x + (double)b.x), anim_f32((double)a.y + (double)b.y), anim_f32((double)a.z + (double)b.z), a.w + b.w};
}
static DRT_HD inline Quat4 quatSub(const Quat4& a, const Quat4& b) {
  return Quat4{anim_f32((double)a.x - (double)b.x), anim_f32((double)a.y - (double)b.y), anim_f32((double)a.z - (double)b.z), a.w - b.w};
}
static DRT_HD inline Quat4 quatNormalize(const Quat4& q) {
  const double l = sqrt(quatDot(q, q));
  return Quat4{anim_f32((double)q.x / l), anim_f32((double)q.y / l), anim_f32((double)q.z / l), q.w / l};
}
static DRT_HD inline Quat4 quatSlerp(double t, const Quat4& q1, const Quat4& q2) {
  const double cosTheta = quatDot(q1, q2);
  if (cosTheta > 0.9995) return quatNormalize(quatAdd(quatScale(q1, 1.0 - t), quatScale(q2, t)));
  const double cc = cosTheta < -1.0 ? -1.0 : (cosTheta > 1.0 ? 1.0 : cosTheta);  // num.clamp: a NaN stays a NaN
  const double thetap = acos(cc) * t;
  const Quat4 qperp = quatNormalize(quatSub(q2, quatScale(q1, cosTheta)));
  return quatAdd(quatScale(q1, cos(thetap)), quatScale(qperp, sin(thetap)));
}
// toTransform: m holds the rotation's transpose; the Transform is (Transpose(m), m)
static DRT_HD inline M4 quatMatrix(const Quat4& q) {
  const double x = q.x, y = q.y, z = q.z, w = q.w;
  const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = x * w, wy = y * w, wz = z * w;
  M4 m = m4Identity();
  m.d[0] = anim_f32(1.0 - 2.0 * (yy + zz)); m.d[1] = anim_f32(2.0 * (xy + wz)); m.d[2] = anim_f32(2.0 * (xz - wy));
  m.d[4] = anim_f32(2.0 * (xy - wz)); m.d[5] = anim_f32(1.0 - 2.0 * (xx + zz)); m.d[6] = anim_f32(2.0 * (yz + wx));
  m.d[8] = anim_f32(2.0 * (xz + wy)); m.d[9] = anim_f32(2.0 * (yz - wx)); m.d[10] = anim_f32(1.0 - 2.0 * (xx + yy));
  return m;
}

// AnimatedTransform.interpolate(time): worldToPrimitive at the ray's time as (m, mInv)
static DRT_HD inline void animInterpolate(const GInstance& a, double time, M4* mOut, M4* invOut) {
  if (!a.animated || time <= a.startTime) { *mOut = a.startM; *invOut = a.startInv; return; }
  if (time >= a.endTime) { *mOut = a.endM; *invOut = a.endInv; return; }
  const double dt = (time - a.startTime) / (a.endTime - a.startTime);
  float tr[3];
  for (int k = 0; k < 3; ++k)  // T[0] * (1 - dt) + T[1] * dt: three float32 Vectors
    tr[k] = anim_f32((double)anim_f32((double)a.T[0][k] * (1.0 - dt)) + (double)anim_f32((double)a.T[1][k] * dt));
  const Quat4 rot = quatSlerp(dt, Quat4{a.Rv[0][0], a.Rv[0][1], a.Rv[0][2], a.Rw[0]}, Quat4{a.Rv[1][0], a.Rv[1][1], a.Rv[1][2], a.Rw[1]});
  M4 scale;
  for (int i = 0; i < 16; ++i) scale.d[i] = anim_f32((double)a.S[0].d[i] * (1.0 - dt) + (double)a.S[1].d[i] * dt);  // Lerp
  M4 tm = m4Identity(), ti = m4Identity();
  tm.d[3] = tr[0]; tm.d[7] = tr[1]; tm.d[11] = tr[2];
  ti.d[3] = anim_f32(-(double)tr[0]); ti.d[7] = anim_f32(-(double)tr[1]); ti.d[11] = anim_f32(-(double)tr[2]);
  const M4 rInv = quatMatrix(rot), rM = m4Transpose(rInv);
  const M4 sInv = m4Inverse(scale);
  // (Translate * rotate) * Transform(scale): m = Mul(Mul(tm, rM), scale), mInv = Mul(sInv, Mul(rInv, ti))
  *mOut = m4Mul(m4Mul(tm, rM), scale);
  *invOut = m4Mul(sInv, m4Mul(rInv, ti));
}

static DRT_HD inline bool m4IsIdentity(const M4& m) {
  for (int i = 0; i < 16; ++i)
    if (m.d[i] != ((i % 5 == 0) ? 1.f : 0.f)) return false;
  return true;
}

// ---- host only: what the AnimatedTransform constructor and TransformedPrimitive.worldBound compute once ----------------------
static inline Quat4 quatFromMatrix(const M4& m) {
  Quat4 q{0.f, 0.f, 0.f, 1.0};
  const double trace = (double)m.d[0] + (double)m.d[5] + (double)m.d[10];
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    q.w = s / 2.0;
    s = 0.5 / s;
    q.x = anim_f32(((double)m.d[9] - (double)m.d[6]) * s);
    q.y = anim_f32(((double)m.d[2] - (double)m.d[8]) * s);
    q.z = anim_f32(((double)m.d[4] - (double)m.d[1]) * s);
  } else {
    const int nxt[3] = {1, 2, 0};
    double v[3] = {0.0, 0.0, 0.0};
    int i = 0;
    if (m.d[5] > m.d[0]) i = 1;
    if (m.d[10] > m.d[i * 4 + i]) i = 2;
    const int j = nxt[i], k = nxt[j];
    double s = std::sqrt(((double)m.d[i * 4 + i] - ((double)m.d[j * 4 + j] + (double)m.d[k * 4 + k])) + 1.0);
    v[i] = s * 0.5;
    if (s != 0.0) s = 0.5 / s;
    q.w = ((double)m.d[k * 4 + j] - (double)m.d[j * 4 + k]) * s;
    v[j] = ((double)m.d[j * 4 + i] + (double)m.d[i * 4 + j]) * s;
    v[k] = ((double)m.d[k * 4 + i] + (double)m.d[i * 4 + k]) * s;
    q.x = anim_f32(v[0]); q.y = anim_f32(v[1]); q.z = anim_f32(v[2]);
  }
  return q;
}
static inline double dartMaxHost(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? std::nan("") : (a > b ? a : b); }
// Decompose(m) -> T, R, S: polar decomposition by averaging R with its inverse transpose until the rows stop moving
static inline void animDecompose(const M4& m, float T[3], Quat4* Rq, M4* S) {
  T[0] = m.d[3]; T[1] = m.d[7]; T[2] = m.d[11];
  M4 M = m;
  for (int i = 0; i < 3; ++i) M.d[i * 4 + 3] = M.d[12 + i] = 0.f;
  M.d[15] = 1.f;
  M4 R = M;
  double norm;
  int count = 0;
  do {
    const M4 Rit = m4Inverse(m4Transpose(R));
    M4 Rnext;
    for (int i = 0; i < 16; ++i) Rnext.d[i] = anim_f32(0.5 * ((double)R.d[i] + (double)Rit.d[i]));
    norm = 0.0;
    for (int i = 0, j = 0; i < 3; ++i, j += 4) {
      const double n = std::fabs((double)R.d[j] - (double)Rnext.d[j]) + std::fabs((double)R.d[j + 1] - (double)Rnext.d[j + 1]) +
                       std::fabs((double)R.d[j + 2] - (double)Rnext.d[j + 2]);
      norm = dartMaxHost(norm, n);
    }
    R = Rnext;
  } while (++count < 100 && norm > 0.0001);
  *Rq = quatFromMatrix(R);
  *S = m4Mul(m4Inverse(R), M);
}
static inline void animInit(GInstance* a, const float* startM, const float* startInv, const float* endM, const float* endInv, double t0,
                            double t1) {
  std::memcpy(a->startM.d, startM, 64); std::memcpy(a->startInv.d, startInv, 64);
  std::memcpy(a->endM.d, endM, 64); std::memcpy(a->endInv.d, endInv, 64);
  a->startTime = t0; a->endTime = t1;
  bool same = true;
  for (int i = 0; i < 16; ++i) same = same && a->startM.d[i] == a->endM.d[i] && a->startInv.d[i] == a->endInv.d[i];
  a->animated = same ? 0 : 1;
  for (int k = 0; k < 2; ++k) {
    Quat4 q;
    animDecompose(k == 0 ? a->startM : a->endM, a->T[k], &q, &a->S[k]);
    a->Rv[k][0] = q.x; a->Rv[k][1] = q.y; a->Rv[k][2] = q.z; a->Rw[k] = q.w;
  }
}
// Transform.transformPoint with the homogeneous divide (transform.dart:110-129)
static inline void m4Point(const M4& m, const float p[3], float out[3]) {
  const double x = p[0], y = p[1], z = p[2];
  float q[3] = {anim_f32((double)m.d[0] * x + (double)m.d[1] * y + (double)m.d[2] * z + (double)m.d[3]),
                anim_f32((double)m.d[4] * x + (double)m.d[5] * y + (double)m.d[6] * z + (double)m.d[7]),
                anim_f32((double)m.d[8] * x + (double)m.d[9] * y + (double)m.d[10] * z + (double)m.d[11])};
  const double w = (double)m.d[12] * x + (double)m.d[13] * y + (double)m.d[14] * z + (double)m.d[15];
  if (w != 1.0) for (int k = 0; k < 3; ++k) q[k] = anim_f32((double)q[k] / w);
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2];
}
// worldToPrimitive.motionBounds(b, true): the union over 128 times of Inverse(interpolate(time)).transformBBox(b)
static inline void animMotionBounds(const GInstance& a, const float bmin[3], const float bmax[3], float outMin[3], float outMax[3]) {
  for (int k = 0; k < 3; ++k) { outMin[k] = INFINITY; outMax[k] = -INFINITY; }
  const int nSteps = a.animated ? 128 : 1;
  for (int i = 0; i < nSteps; ++i) {
    M4 m, inv;
    if (a.animated) {
      const double s = (double)i / (128 - 1);
      animInterpolate(a, a.startTime * (1.0 - s) + a.endTime * s, &m, &inv);
    } else {
      m = a.startM; inv = a.startInv;
    }
    for (int c = 0; c < 8; ++c) {  // transform.dart:163-178 (the union does not depend on the corner order)
      const float p[3] = {(c & 1) ? bmax[0] : bmin[0], (c & 2) ? bmax[1] : bmin[1], (c & 4) ? bmax[2] : bmin[2]};
      float q[3];
      m4Point(inv, p, q);
      for (int k = 0; k < 3; ++k) {
        outMin[k] = q[k] < outMin[k] ? q[k] : outMin[k];
        outMax[k] = q[k] > outMax[k] ? q[k] : outMax[k];
      }
    }
  }
}
}  // namespace drt
