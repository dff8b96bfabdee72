// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.h header).  PARITY UNPINNED.
//
// AnimatedTransform and what it stands on, restated from the reference:
//   lib/core/matrix4x4.dart (Float32List storage: every element store rounds to binary32; Mul :197-210, Transpose :190-195,
//   determinant :242-290, invert :295-357), lib/core/quaternion.dart (v is a float32 Vector, w a Dart double; fromMatrix :39-77,
//   toTransform :120-149, Slerp :151-161), lib/core/transform.dart (Transform(m) inverts, operator * :83-86, Translate :214-227),
//   lib/core/animated_transform.dart (Decompose :61-105, interpolate :107-136, motionBounds :183-200).
#pragma once
#include "ref_core.h"

namespace orc {

struct Mat4 {
  float d[16];
  Mat4() {
    for (int i = 0; i < 16; ++i) d[i] = (i % 5 == 0) ? 1.f : 0.f;
  }
  explicit Mat4(const float* m) { std::memcpy(d, m, sizeof(d)); }
  bool operator==(const Mat4& o) const {  // matrix4x4.dart:86-93 (element-wise !=: a NaN makes the matrices differ)
    for (int i = 0; i < 16; ++i)
      if (d[i] != o.d[i]) return false;
    return true;
  }
};

static inline Mat4 Mat4Transpose(const Mat4& m) {
  Mat4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) r.d[4 * i + j] = m.d[4 * j + i];
  return r;
}
static inline Mat4 Mat4Mul(const Mat4& a, const Mat4& b) {  // matrix4x4.dart:197-210: one f64 expression per element, left to right
  Mat4 r;
  for (int i = 0, k = 0; i < 4; ++i, k += 4)
    for (int j = 0; j < 4; ++j)
      r.d[k + j] = f32((double)a.d[k] * b.d[j] + (double)a.d[k + 1] * b.d[4 + j] + (double)a.d[k + 2] * b.d[8 + j] +
                       (double)a.d[k + 3] * b.d[12 + j]);
  return r;
}
// matrix4x4.dart:242-357.  The names follow the file: n<row><col> read with the indices of the TRANSPOSE (n12 = data[4]); the
// results are stored transposed back, so the function inverts the matrix.  Every product chain is evaluated left to right in f64.
static inline Mat4 Mat4Inverse(const Mat4& m) {
  const double n11 = m.d[0], n12 = m.d[4], n13 = m.d[8], n14 = m.d[12];
  const double n21 = m.d[1], n22 = m.d[5], n23 = m.d[9], n24 = m.d[13];
  const double n31 = m.d[2], n32 = m.d[6], n33 = m.d[10], n34 = m.d[14];
  const double n41 = m.d[3], n42 = m.d[7], n43 = m.d[11], n44 = m.d[15];
  const double det = (n14 * n23 * n32 * n41) - (n13 * n24 * n32 * n41) - (n14 * n22 * n33 * n41) + (n12 * n24 * n33 * n41) +
                     (n13 * n22 * n34 * n41) - (n12 * n23 * n34 * n41) - (n14 * n23 * n31 * n42) + (n13 * n24 * n31 * n42) +
                     (n14 * n21 * n33 * n42) - (n11 * n24 * n33 * n42) - (n13 * n21 * n34 * n42) + (n11 * n23 * n34 * n42) +
                     (n14 * n22 * n31 * n43) - (n12 * n24 * n31 * n43) - (n14 * n21 * n32 * n43) + (n11 * n24 * n32 * n43) +
                     (n12 * n21 * n34 * n43) - (n11 * n22 * n34 * n43) - (n13 * n22 * n31 * n44) + (n12 * n23 * n31 * n44) +
                     (n13 * n21 * n32 * n44) - (n11 * n23 * n32 * n44) - (n12 * n21 * n33 * n44) + (n11 * n22 * n33 * n44);
  if (det == 0.0) return m;
  const double invDet = 1.0 / det;
  Mat4 r;
  r.d[0] = f32((n23 * n34 * n42 - n24 * n33 * n42 + n24 * n32 * n43 - n22 * n34 * n43 - n23 * n32 * n44 + n22 * n33 * n44) * invDet);
  r.d[4] = f32((n14 * n33 * n42 - n13 * n34 * n42 - n14 * n32 * n43 + n12 * n34 * n43 + n13 * n32 * n44 - n12 * n33 * n44) * invDet);
  r.d[8] = f32((n13 * n24 * n42 - n14 * n23 * n42 + n14 * n22 * n43 - n12 * n24 * n43 - n13 * n22 * n44 + n12 * n23 * n44) * invDet);
  r.d[12] = f32((n14 * n23 * n32 - n13 * n24 * n32 - n14 * n22 * n33 + n12 * n24 * n33 + n13 * n22 * n34 - n12 * n23 * n34) * invDet);
  r.d[1] = f32((n24 * n33 * n41 - n23 * n34 * n41 - n24 * n31 * n43 + n21 * n34 * n43 + n23 * n31 * n44 - n21 * n33 * n44) * invDet);
  r.d[5] = f32((n13 * n34 * n41 - n14 * n33 * n41 + n14 * n31 * n43 - n11 * n34 * n43 - n13 * n31 * n44 + n11 * n33 * n44) * invDet);
  r.d[9] = f32((n14 * n23 * n41 - n13 * n24 * n41 - n14 * n21 * n43 + n11 * n24 * n43 + n13 * n21 * n44 - n11 * n23 * n44) * invDet);
  r.d[13] = f32((n13 * n24 * n31 - n14 * n23 * n31 + n14 * n21 * n33 - n11 * n24 * n33 - n13 * n21 * n34 + n11 * n23 * n34) * invDet);
  r.d[2] = f32((n22 * n34 * n41 - n24 * n32 * n41 + n24 * n31 * n42 - n21 * n34 * n42 - n22 * n31 * n44 + n21 * n32 * n44) * invDet);
  r.d[6] = f32((n14 * n32 * n41 - n12 * n34 * n41 - n14 * n31 * n42 + n11 * n34 * n42 + n12 * n31 * n44 - n11 * n32 * n44) * invDet);
  r.d[10] = f32((n12 * n24 * n41 - n14 * n22 * n41 + n14 * n21 * n42 - n11 * n24 * n42 - n12 * n21 * n44 + n11 * n22 * n44) * invDet);
  r.d[14] = f32((n14 * n22 * n31 - n12 * n24 * n31 - n14 * n21 * n32 + n11 * n24 * n32 + n12 * n21 * n34 - n11 * n22 * n34) * invDet);
  r.d[3] = f32((n23 * n32 * n41 - n22 * n33 * n41 - n23 * n31 * n42 + n21 * n33 * n42 + n22 * n31 * n43 - n21 * n32 * n43) * invDet);
  r.d[7] = f32((n12 * n33 * n41 - n13 * n32 * n41 + n13 * n31 * n42 - n11 * n33 * n42 - n12 * n31 * n43 + n11 * n32 * n43) * invDet);
  r.d[11] = f32((n13 * n22 * n41 - n12 * n23 * n41 - n13 * n21 * n42 + n11 * n23 * n42 + n12 * n21 * n43 - n11 * n22 * n43) * invDet);
  r.d[15] = f32((n12 * n23 * n31 - n13 * n22 * n31 + n13 * n21 * n32 - n11 * n23 * n32 - n12 * n21 * n33 + n11 * n22 * n33) * invDet);
  return r;
}

static inline Transform XfFrom(const Mat4& m, const Mat4& inv) { return Transform(m.d, inv.d); }
static inline Transform XfMul(const Transform& a, const Transform& b) {  // transform.dart:83-86
  return XfFrom(Mat4Mul(Mat4(a.m), Mat4(b.m)), Mat4Mul(Mat4(b.mInv), Mat4(a.mInv)));
}
static inline Transform XfInverse(const Transform& t) { return Transform(t.mInv, t.m); }  // transform.dart:58-60
static inline bool XfIsIdentity(const Transform& t) { return Mat4(t.m) == Mat4(); }       // transform.dart:47-56

struct Quat {
  Vec v;
  double w = 1.0;
};
static inline double QDot(const Quat& a, const Quat& b) { return Dot(a.v, b.v) + a.w * b.w; }
static inline Quat QScale(const Quat& q, double f) { Quat r; r.v = q.v * f; r.w = q.w * f; return r; }
static inline Quat QAdd(const Quat& a, const Quat& b) { Quat r; r.v = a.v + b.v; r.w = a.w + b.w; return r; }
static inline Quat QSub(const Quat& a, const Quat& b) { Quat r; r.v = a.v - b.v; r.w = a.w - b.w; return r; }
static inline Quat QNormalize(const Quat& q) {
  const double l = std::sqrt(QDot(q, q));
  Quat r; r.v = q.v / l; r.w = q.w / l;
  return r;
}
static inline Quat QFromMatrix(const Mat4& m) {  // quaternion.dart:39-77
  Quat q;
  const double trace = (double)m.d[0] + m.d[5] + m.d[10];
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    q.w = s / 2.0;
    s = 0.5 / s;
    q.v = Vec(((double)m.d[9] - m.d[6]) * s, ((double)m.d[2] - m.d[8]) * s, ((double)m.d[4] - m.d[1]) * s);
  } else {
    static const int nxt[3] = {1, 2, 0};
    double qq[3] = {0.0, 0.0, 0.0};
    int i = 0;
    if (m.d[5] > m.d[0]) i = 1;
    if (m.d[10] > m.d[i * 4 + i]) i = 2;
    const int j = nxt[i], k = nxt[j];
    double s = std::sqrt(((double)m.d[i * 4 + i] - ((double)m.d[j * 4 + j] + m.d[k * 4 + k])) + 1.0);
    qq[i] = s * 0.5;
    if (s != 0.0) s = 0.5 / s;
    q.w = ((double)m.d[k * 4 + j] - m.d[j * 4 + k]) * s;
    qq[j] = ((double)m.d[j * 4 + i] + m.d[i * 4 + j]) * s;
    qq[k] = ((double)m.d[k * 4 + i] + m.d[i * 4 + k]) * s;
    q.v = Vec(qq[0], qq[1], qq[2]);
  }
  return q;
}
static inline Transform QToTransform(const Quat& q) {  // quaternion.dart:120-149
  const double x = q.v.x, y = q.v.y, z = q.v.z, w = q.w;
  const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = x * w, wy = y * w, wz = z * w;
  Mat4 m;
  m.d[0] = f32(1.0 - 2.0 * (yy + zz)); m.d[1] = f32(2.0 * (xy + wz)); m.d[2] = f32(2.0 * (xz - wy));
  m.d[4] = f32(2.0 * (xy - wz)); m.d[5] = f32(1.0 - 2.0 * (xx + zz)); m.d[6] = f32(2.0 * (yz + wx));
  m.d[8] = f32(2.0 * (xz + wy)); m.d[9] = f32(2.0 * (yz - wx)); m.d[10] = f32(1.0 - 2.0 * (xx + yy));
  return XfFrom(Mat4Transpose(m), m);
}
static inline Quat QSlerp(double t, const Quat& q1, const Quat& q2) {  // quaternion.dart:151-161
  const double cosTheta = QDot(q1, q2);
  if (cosTheta > 0.9995) return QNormalize(QAdd(QScale(q1, 1.0 - t), QScale(q2, t)));
  const double theta = std::acos(clampd(cosTheta, -1.0, 1.0));
  const double thetap = theta * t;
  const Quat qperp = QNormalize(QSub(q2, QScale(q1, cosTheta)));
  return QAdd(QScale(q1, std::cos(thetap)), QScale(qperp, std::sin(thetap)));
}

struct AnimatedTransform {
  Transform start, end;
  double startTime = 0.0, endTime = 1.0;
  bool actuallyAnimated = false;
  Vec T[2];
  Quat R[2];
  Mat4 S[2];

  static void Decompose(const Mat4& m, Vec* T, Quat* Rq, Mat4* S) {  // animated_transform.dart:61-105
    *T = Vec(m.d[3], m.d[7], m.d[11]);
    Mat4 M = m;
    for (int i = 0; i < 3; ++i) M.d[i * 4 + 3] = M.d[12 + i] = 0.f;
    M.d[15] = 1.f;
    double norm;
    int count = 0;
    Mat4 R = M;
    do {
      Mat4 Rnext;
      const Mat4 Rit = Mat4Inverse(Mat4Transpose(R));
      for (int i = 0; i < 16; ++i) Rnext.d[i] = f32(0.5 * ((double)R.d[i] + Rit.d[i]));
      norm = 0.0;
      for (int i = 0, j = 0; i < 3; ++i, j += 4) {
        const double n = std::fabs((double)R.d[j] - Rnext.d[j]) + std::fabs((double)R.d[j + 1] - Rnext.d[j + 1]) +
                         std::fabs((double)R.d[j + 2] - Rnext.d[j + 2]);
        norm = (std::isnan(norm) || std::isnan(n)) ? std::nan("") : (norm > n ? norm : n);  // dart:math max
      }
      R = Rnext;
    } while (++count < 100 && norm > 0.0001);
    *Rq = QFromMatrix(R);
    *S = Mat4Mul(Mat4Inverse(R), M);
  }

  void init(const Transform& t1, double time1, const Transform& t2, double time2) {  // animated_transform.dart:35-43
    start = t1; end = t2;
    startTime = time1; endTime = time2;
    actuallyAnimated = !(Mat4(t1.m) == Mat4(t2.m) && Mat4(t1.mInv) == Mat4(t2.mInv));  // transform.dart:67-69
    Decompose(Mat4(start.m), &T[0], &R[0], &S[0]);
    Decompose(Mat4(end.m), &T[1], &R[1], &S[1]);
  }

  Transform interpolate(double time) const {  // animated_transform.dart:107-136
    if (!actuallyAnimated || time <= startTime) return start;
    if (time >= endTime) return end;
    const double dt = (time - startTime) / (endTime - startTime);
    const Vec trans = T[0] * (1.0 - dt) + T[1] * dt;
    const Quat rotate = QSlerp(dt, R[0], R[1]);
    Mat4 scale;
    for (int i = 0; i < 16; ++i) scale.d[i] = f32((double)S[0].d[i] * (1.0 - dt) + (double)S[1].d[i] * dt);  // Lerp, common.dart:80-81
    Mat4 tm, tinv;  // Transform.Translate, transform.dart:214-227
    tm.d[3] = trans.x; tm.d[7] = trans.y; tm.d[11] = trans.z;
    tinv.d[3] = f32(-(double)trans.x); tinv.d[7] = f32(-(double)trans.y); tinv.d[11] = f32(-(double)trans.z);
    const Transform scaleT = XfFrom(scale, Mat4Inverse(scale));  // new Transform(scale): transform.dart:31-35
    return XfMul(XfMul(XfFrom(tm, tinv), QToTransform(rotate)), scaleT);
  }

  // AnimatedTransform.transformRay (animated_transform.dart:138-154) picks the same three cases as interpolate
  BBox motionBounds(const BBox& b, bool useInverse) const {  // animated_transform.dart:183-200
    if (!actuallyAnimated) return XfInverse(start).bbox(b);
    BBox ret;
    const int nSteps = 128;
    for (int i = 0; i < nSteps; ++i) {
      const double s = (double)i / (nSteps - 1);
      const double time = startTime * (1.0 - s) + endTime * s;
      Transform t = interpolate(time);
      if (useInverse) t = XfInverse(t);
      ret = Union(ret, t.bbox(b));
    }
    return ret;
  }
};

}  // namespace orc
