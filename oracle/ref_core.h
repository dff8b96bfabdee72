// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of DartRay's core numeric types (all citations relative to
// /root/reference).  PARITY UNPINNED: the reference ships no golden vectors or
// working tests (test/spectrum_test.dart:14-39 is commented out) and no Dart VM
// exists in this image, so this restatement is pinned only by analytic
// known-answer tests and brute-force cross-checks (tests/).
//
// Numeric model honoured everywhere in oracle/:
//   * Vector/Point/Normal/Matrix4x4/RGBColor STORE float32
//     (lib/core/vector.dart:26-34, matrix4x4.dart:26-27) but every arithmetic
//     expression is evaluated in IEEE binary64 and rounded to binary32 only when
//     a new object is constructed or a component is stored.
//   * No FMA contraction (build with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

static const double kInf = std::numeric_limits<double>::infinity();  // common.dart:26 (1.0e500)
static const double kPi = 3.141592653589793;                         // dart:math pi
static const double INV_PI = 0.31830988618379067154;                 // common.dart:23
static const double INV_TWOPI = 0.15915494309189533577;              // common.dart:24
static const double ONE_MINUS_EPSILON = 0.9999999403953552;          // montecarlo.dart:23

static inline float f32(double v) { return (float)v; }

// lib/core/vector.dart:26 / point.dart:26 / normal.dart:26 — float32 storage.
struct Vec {
  float x = 0.f, y = 0.f, z = 0.f;
  Vec() {}
  Vec(double X, double Y, double Z) : x(f32(X)), y(f32(Y)), z(f32(Z)) {}
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  void set(int i, double v) { (i == 0 ? x : (i == 1 ? y : z)) = f32(v); }
};

// vector.dart:57-74 (operators build a new f32 object from f64 expressions)
static inline Vec operator+(const Vec& a, const Vec& b) {
  return Vec((double)a.x + b.x, (double)a.y + b.y, (double)a.z + b.z);
}
static inline Vec operator-(const Vec& a, const Vec& b) {
  return Vec((double)a.x - b.x, (double)a.y - b.y, (double)a.z - b.z);
}
static inline Vec operator*(const Vec& a, double f) {
  return Vec((double)a.x * f, (double)a.y * f, (double)a.z * f);
}
static inline Vec operator/(const Vec& a, double f) {
  return Vec((double)a.x / f, (double)a.y / f, (double)a.z / f);
}
static inline Vec operator-(const Vec& a) { return Vec(-(double)a.x, -(double)a.y, -(double)a.z); }

// vector.dart:150-152
static inline double Dot(const Vec& a, const Vec& b) {
  return (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z;
}
static inline double AbsDot(const Vec& a, const Vec& b) { return std::fabs(Dot(a, b)); }
// vector.dart:158-168
static inline Vec Cross(const Vec& a, const Vec& b) {
  double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
  return Vec((ay * bz) - (az * by), (az * bx) - (ax * bz), (ax * by) - (ay * bx));
}
static inline double LengthSquared(const Vec& v) {
  return (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z;
}
static inline double Length(const Vec& v) { return std::sqrt(LengthSquared(v)); }
// vector.dart:170
static inline Vec Normalize(const Vec& v) { return v / Length(v); }
static inline double DistanceSquared(const Vec& a, const Vec& b) { return LengthSquared(b - a); }
static inline double Distance(const Vec& a, const Vec& b) { return Length(b - a); }

// vector.dart:198-214
static inline void CoordinateSystem(const Vec& v1, Vec* v2, Vec* v3) {
  if (std::fabs((double)v1.x) > std::fabs((double)v1.y)) {
    double invLen = 1.0 / std::sqrt((double)v1.x * v1.x + (double)v1.z * v1.z);
    *v2 = Vec(-(double)v1.z * invLen, 0.0, (double)v1.x * invLen);
  } else {
    double invLen = 1.0 / std::sqrt((double)v1.y * v1.y + (double)v1.z * v1.z);
    *v2 = Vec(0.0, (double)v1.z * invLen, -(double)v1.y * invLen);
  }
  *v3 = Cross(v1, *v2);
}

// vector.dart:216-218 FaceForward
static inline Vec FaceForward(const Vec& n, const Vec& v) { return (Dot(n, v) < 0.0) ? -n : n; }

// vector.dart:172-183
static inline Vec SphericalDirection(double sintheta, double costheta, double phi) {
  return Vec(sintheta * std::cos(phi), sintheta * std::sin(phi), costheta);
}

// lib/core/ray.dart:27-75 — origin/direction f32, min/max distance f64.
struct Ray {
  Vec o, d;
  double mint = 0.0, maxt = kInf;
  double time = 0.0;
  int depth = 0;
  Ray() {}
  Ray(const Vec& O, const Vec& D, double mn = 0.0, double mx = kInf, double tm = 0.0, int dp = 0)
      : o(O), d(D), mint(mn), maxt(mx), time(tm), depth(dp) {}
  // ray.dart:70-71: origin + (direction * t), each step a new f32 object
  Vec at(double t) const { return o + (d * t); }
};

// lib/core/bbox.dart:27-205
struct BBox {
  Vec pMin, pMax;
  BBox() {
    pMin.x = pMin.y = pMin.z = std::numeric_limits<float>::infinity();
    pMax.x = pMax.y = pMax.z = -std::numeric_limits<float>::infinity();
  }
  explicit BBox(const Vec& p) : pMin(p), pMax(p) {}
  BBox(const Vec& p1, const Vec& p2) {  // bbox.dart:35-41
    pMin = Vec(std::fmin((double)p1.x, (double)p2.x), std::fmin((double)p1.y, (double)p2.y),
               std::fmin((double)p1.z, (double)p2.z));
    pMax = Vec(std::fmax((double)p1.x, (double)p2.x), std::fmax((double)p1.y, (double)p2.y),
               std::fmax((double)p1.z, (double)p2.z));
  }
  const Vec& operator[](int i) const { return i == 0 ? pMin : pMax; }
  // bbox.dart:68  (pMin * 0.5) + (pMax * 0.5), f32 rounding at each Point op
  Vec center() const { return (pMin * 0.5) + (pMax * 0.5); }
  // bbox.dart:164-167: d is a Vector (f32), the area expression is f64
  double surfaceArea() const {
    Vec d = pMax - pMin;
    return 2.0 * ((double)d.x * d.y + (double)d.x * d.z + (double)d.y * d.z);
  }
  // bbox.dart:174-183
  int maximumExtent() const {
    Vec diag = pMax - pMin;
    if (diag.x > diag.y && diag.x > diag.z) return 0;
    if (diag.y > diag.z) return 1;
    return 2;
  }
  void expand(double delta) {  // bbox.dart:155-162
    pMin = Vec((double)pMin.x - delta, (double)pMin.y - delta, (double)pMin.z - delta);
    pMax = Vec((double)pMax.x + delta, (double)pMax.y + delta, (double)pMax.z + delta);
  }
};
static inline float fminf32(float a, float b) { return a < b ? a : b; }
static inline float fmaxf32(float a, float b) { return a > b ? a : b; }
static inline BBox Union(const BBox& b, const BBox& b2) {  // bbox.dart:143-153,203-205
  BBox r = b;
  r.pMin.x = fminf32(r.pMin.x, b2.pMin.x); r.pMin.y = fminf32(r.pMin.y, b2.pMin.y);
  r.pMin.z = fminf32(r.pMin.z, b2.pMin.z);
  r.pMax.x = fmaxf32(r.pMax.x, b2.pMax.x); r.pMax.y = fmaxf32(r.pMax.y, b2.pMax.y);
  r.pMax.z = fmaxf32(r.pMax.z, b2.pMax.z);
  return r;
}
static inline BBox UnionPoint(const BBox& b, const Vec& p) {  // bbox.dart:131-141,199-201
  BBox r = b;
  r.pMin.x = fminf32(r.pMin.x, p.x); r.pMin.y = fminf32(r.pMin.y, p.y); r.pMin.z = fminf32(r.pMin.z, p.z);
  r.pMax.x = fmaxf32(r.pMax.x, p.x); r.pMax.y = fmaxf32(r.pMax.y, p.y); r.pMax.z = fmaxf32(r.pMax.z, p.z);
  return r;
}

// lib/core/matrix4x4.dart:26 (Float32List(16), row-major) + transform.dart:27
struct Transform {
  float m[16];
  float mInv[16];
  Transform() {
    static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::memcpy(m, I, sizeof(I));
    std::memcpy(mInv, I, sizeof(I));
  }
  Transform(const float* M, const float* Minv) {
    std::memcpy(m, M, sizeof(m));
    std::memcpy(mInv, Minv, sizeof(mInv));
  }
  // transform.dart:110-129
  Vec point(const Vec& p) const {
    double x = p.x, y = p.y, z = p.z;
    Vec out((double)m[0] * x + (double)m[1] * y + (double)m[2] * z + (double)m[3],
            (double)m[4] * x + (double)m[5] * y + (double)m[6] * z + (double)m[7],
            (double)m[8] * x + (double)m[9] * y + (double)m[10] * z + (double)m[11]);
    double w = (double)m[12] * x + (double)m[13] * y + (double)m[14] * z + (double)m[15];
    if (w != 1.0) out = Vec((double)out.x / w, (double)out.y / w, (double)out.z / w);  // invScale
    return out;
  }
  // transform.dart:131-145
  Vec vector(const Vec& p) const {
    double x = p.x, y = p.y, z = p.z;
    return Vec((double)m[0] * x + (double)m[1] * y + (double)m[2] * z,
               (double)m[4] * x + (double)m[5] * y + (double)m[6] * z,
               (double)m[8] * x + (double)m[9] * y + (double)m[10] * z);
  }
  // transform.dart:147-161 (transpose of the inverse)
  Vec normal(const Vec& p) const {
    double x = p.x, y = p.y, z = p.z;
    return Vec((double)mInv[0] * x + (double)mInv[4] * y + (double)mInv[8] * z,
               (double)mInv[1] * x + (double)mInv[5] * y + (double)mInv[9] * z,
               (double)mInv[2] * x + (double)mInv[6] * y + (double)mInv[10] * z);
  }
  // transform.dart:163-178
  BBox bbox(const BBox& b) const {
    BBox out(point(b.pMin));
    out = UnionPoint(out, point(Vec(b.pMax.x, b.pMin.y, b.pMin.z)));
    out = UnionPoint(out, point(Vec(b.pMin.x, b.pMax.y, b.pMin.z)));
    out = UnionPoint(out, point(Vec(b.pMin.x, b.pMin.y, b.pMax.z)));
    out = UnionPoint(out, point(Vec(b.pMin.x, b.pMax.y, b.pMax.z)));
    out = UnionPoint(out, point(Vec(b.pMax.x, b.pMax.y, b.pMin.z)));
    out = UnionPoint(out, point(Vec(b.pMax.x, b.pMin.y, b.pMax.z)));
    out = UnionPoint(out, point(b.pMax));
    return out;
  }
  // transform.dart:180-195
  Ray ray(const Ray& r) const {
    Ray tr;
    tr.o = point(r.o);
    tr.d = vector(r.d);
    tr.mint = r.mint; tr.maxt = r.maxt; tr.time = r.time; tr.depth = r.depth;
    return tr;
  }
};

static inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline double Radians(double deg) { return (kPi / 180.0) * deg; }

}  // namespace orc
