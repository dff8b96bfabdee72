// ORACLE — TEST INFRASTRUCTURE ONLY.  See ref_texture.h for what this restates; citations relative to /root/reference.
#include "ref_texture.h"

#include <algorithm>
#include <cmath>

#include "ref_render.h"

namespace orc {

static inline double Log2d(double x) { return std::log(x) * (1.0 / std::log(2.0)); }  // common.dart:98-103
static inline int64_t dartMod(int64_t a, int64_t n) { return ((a % n) + n) % n; }     // Dart's % is never negative

// ---- RayDifferential --------------------------------------------------------------------------------
void RayDiff::scale(const Vec& o, const Vec& d, double s) {  // ray_differential.dart:56-61
  rxo = o + (rxo - o) * s;
  ryo = o + (ryo - o) * s;
  rxd = d + (rxd - d) * s;
  ryd = d + (ryd - d) * s;
}

// ---- DifferentialGeometry.computeDifferentials (differential_geometry.dart:122-205) -----------------
static bool solve2x2(const double A[4], const double B[2], double* x0, double* x1) {  // common.dart:170-185
  double det = A[0] * A[3] - A[1] * A[2];
  if (std::fabs(det) < 1.0e-10) return false;
  *x0 = (A[3] * B[0] - A[1] * B[1]) / det;
  *x1 = (A[0] * B[1] - A[2] * B[0]) / det;
  if (std::isnan(*x0) || std::isnan(*x1)) return false;
  return true;
}

void computeDifferentials(DG* dg, const RayDiff& rd) {
  auto zero = [&]() {
    dg->dudx = dg->dvdx = dg->dudy = dg->dvdy = 0.0;
    dg->dpdx = Vec(0.0, 0.0, 0.0);
    dg->dpdy = Vec(0.0, 0.0, 0.0);
  };
  if (!rd.has) { zero(); return; }
  const Vec& nn = dg->nn;
  const Vec& p = dg->p;
  double d = -Dot(nn, Vec(p.x, p.y, p.z));
  Vec rxv(rd.rxo.x, rd.rxo.y, rd.rxo.z);
  double tx = -(Dot(nn, rxv) + d) / Dot(nn, rd.rxd);
  if (std::isnan(tx)) { zero(); return; }
  Vec px = rd.rxo + rd.rxd * tx;
  Vec ryv = rd.ryo;
  double ty = -(Dot(nn, ryv) + d) / Dot(nn, rd.ryd);
  if (std::isnan(ty)) { zero(); return; }  // dpdx keeps whatever it held: a fresh DifferentialGeometry holds zeros
  Vec py = rd.ryo + rd.ryd * ty;
  dg->dpdx = px - p;
  dg->dpdy = py - p;
  int a0, a1;
  if (std::fabs((double)nn.x) > std::fabs((double)nn.y) && std::fabs((double)nn.x) > std::fabs((double)nn.z)) { a0 = 1; a1 = 2; }
  else if (std::fabs((double)nn.y) > std::fabs((double)nn.z)) { a0 = 0; a1 = 2; }
  else { a0 = 0; a1 = 1; }
  double A[4] = {dg->dpdu[a0], dg->dpdv[a0], dg->dpdu[a1], dg->dpdv[a1]};
  double Bx[2] = {px[a0] - p[a0], px[a1] - p[a1]};
  double By[2] = {py[a0] - p[a0], py[a1] - p[a1]};
  double du, dv;
  if (!solve2x2(A, Bx, &du, &dv)) { dg->dudx = 0.0; dg->dvdx = 0.0; }
  else { dg->dudx = du; dg->dvdx = dv; }
  if (!solve2x2(A, By, &du, &dv)) { dg->dudy = 0.0; dg->dvdy = 0.0; }
  else { dg->dudy = du; dg->dvdy = dv; }
}

// ---- MIPMap (mipmap.dart) -----------------------------------------------------------------------------
static float gEwaLut[128];
static bool gEwaLutReady = false;
static void ewaLutInit() {  // :168-176: a Float32List
  if (gEwaLutReady) return;
  for (int i = 0; i < 128; ++i) {
    double alpha = 2.0, r2 = (double)i / (128 - 1);
    gEwaLut[i] = f32(std::exp(-alpha * r2) - std::exp(-alpha));
  }
  gEwaLutReady = true;
}

void TexImage::init(int width, int height, int channels_, const float* texels, int wrap_, bool trilinear_, double maxAniso_) {
  channels = channels_; wrap = wrap_; trilinear = trilinear_; maxAniso = maxAniso_;
  ewaLutInit();
  levels = 1 + (int)Log2d((double)std::max(width, height));  // :143
  w.assign(levels, 1); h.assign(levels, 1);
  data.assign(levels, {});
  w[0] = width; h[0] = height;
  data[0].assign(texels, texels + (size_t)width * height * channels);
  for (int i = 1; i < levels; ++i) {  // :152-166
    const int sRes = std::max(1, w[i - 1] / 2), tRes = std::max(1, h[i - 1] / 2);
    w[i] = sRes; h[i] = tRes;
    data[i].resize((size_t)sRes * tRes * channels);
    for (int t = 0; t < tRes; ++t)
      for (int s = 0; s < sRes; ++s) {
        float* out = &data[i][((size_t)t * sRes + s) * channels];
        if (channels == 1) {  // doubles by value: the plain average, stored into a Float32List
          out[0] = f32((texelF(i - 1, 2 * s, 2 * t) + texelF(i - 1, 2 * s + 1, 2 * t) + texelF(i - 1, 2 * s, 2 * t + 1) +
                        texelF(i - 1, 2 * s + 1, 2 * t + 1)) * 0.25);
        } else {
          // AS WRITTEN: SpectrumImage.operator[] hands out ONE shared RGBColor (spectrum_image.dart:103-112,133-135), so in
          // `texel(a) + texel(b)` both operands are the same object by the time operator+ runs and the first texel is lost:
          // the level holds (2 * b + c + d) * 0.25, every operation a new float32 RGBColor.
          float b[3], c[3], d[3];
          texelS(i - 1, 2 * s + 1, 2 * t, b);
          texelS(i - 1, 2 * s, 2 * t + 1, c);
          texelS(i - 1, 2 * s + 1, 2 * t + 1, d);
          for (int k = 0; k < 3; ++k) {
            float acc = f32((double)b[k] + (double)b[k]);
            acc = f32((double)acc + (double)c[k]);
            acc = f32((double)acc + (double)d[k]);
            out[k] = f32((double)acc * 0.25);
          }
        }
      }
  }
}

// texel(), :183-204.  `black` reports TEXTURE_BLACK's out-of-range case.
static inline bool wrapST(int wrap, int64_t W, int64_t H, int64_t* s, int64_t* t) {
  if (wrap == 0) { *s = dartMod(*s, W); *t = dartMod(*t, H); }
  else if (wrap == 2) { *s = std::min(std::max(*s, (int64_t)0), W - 1); *t = std::min(std::max(*t, (int64_t)0), H - 1); }
  else if (*s < 0 || *s >= W || *t < 0 || *t >= H) return false;
  return true;
}
void TexImage::texelS(int level, int64_t s, int64_t t, float out[3]) const {
  if (!wrapST(wrap, w[level], h[level], &s, &t)) { out[0] = out[1] = out[2] = 0.f; return; }
  const float* q = &data[level][(size_t)(t * w[level] + s) * 3];
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2];
}
double TexImage::texelF(int level, int64_t s, int64_t t) const {
  // TEXTURE_BLACK returns `new Spectrum(0.0)` even for a float image (:196-199); the arithmetic on it would throw in Dart —
  // a float image with wrap "black" read outside [0, 1) is not a case the reference can evaluate; 0 here.
  if (!wrapST(wrap, w[level], h[level], &s, &t)) return 0.0;
  return data[level][(size_t)(t * w[level] + s)];
}

// triangle(), :341-355
void TexImage::triangleS(int level, double s, double t, float out[3]) const {
  level = std::min(std::max(level, 0), levels - 1);
  s = s * w[level] - 0.5;
  t = t * h[level] - 0.5;
  const int64_t s0 = (int64_t)std::floor(s), t0 = (int64_t)std::floor(t);
  const double ds = s - s0, dt = t - t0;
  float a[3], b[3], c[3], d[3];
  texelS(level, s0, t0, a); texelS(level, s0, t0 + 1, b); texelS(level, s0 + 1, t0, c); texelS(level, s0 + 1, t0 + 1, d);
  for (int k = 0; k < 3; ++k) {
    float acc = f32((double)a[k] * ((1.0 - ds) * (1.0 - dt)));
    acc = f32((double)acc + (double)f32((double)b[k] * ((1.0 - ds) * dt)));
    acc = f32((double)acc + (double)f32((double)c[k] * (ds * (1.0 - dt))));
    out[k] = f32((double)acc + (double)f32((double)d[k] * (ds * dt)));
  }
}
double TexImage::triangleF(int level, double s, double t) const {
  level = std::min(std::max(level, 0), levels - 1);
  s = s * w[level] - 0.5;
  t = t * h[level] - 0.5;
  const int64_t s0 = (int64_t)std::floor(s), t0 = (int64_t)std::floor(t);
  const double ds = s - s0, dt = t - t0;
  return texelF(level, s0, t0) * ((1.0 - ds) * (1.0 - dt)) + texelF(level, s0, t0 + 1) * ((1.0 - ds) * dt) +
         texelF(level, s0 + 1, t0) * (ds * (1.0 - dt)) + texelF(level, s0 + 1, t0 + 1) * (ds * dt);
}

// lookup(), :206-222
void TexImage::lookupS(double s, double t, double width, float out[3]) const {
  const double level = levels - 1 + Log2d(std::fmax(width, 1.0e-8));
  if (level < 0) { triangleS(0, s, t, out); return; }
  if (level >= levels - 1) { texelS(levels - 1, 0, 0, out); return; }
  const int iLevel = (int)std::floor(level);
  const double delta = level - iLevel;
  float a[3], b[3];
  triangleS(iLevel, s, t, a);
  triangleS(iLevel + 1, s, t, b);
  for (int k = 0; k < 3; ++k) out[k] = f32((double)f32((double)a[k] * (1.0 - delta)) + (double)f32((double)b[k] * delta));
}
double TexImage::lookupF(double s, double t, double width) const {
  const double level = levels - 1 + Log2d(std::fmax(width, 1.0e-8));
  if (level < 0) return triangleF(0, s, t);
  if (level >= levels - 1) return texelF(levels - 1, 0, 0);
  const int iLevel = (int)std::floor(level);
  const double delta = level - iLevel;
  return triangleF(iLevel, s, t) * (1.0 - delta) + triangleF(iLevel + 1, s, t) * delta;
}

// EWA(), :270-339.  The ellipse set-up is shared; `S` accumulates a Spectrum (float32 per operation), `F` a double.
namespace {
struct Ellipse {
  double s, t, A, B, C;
  int64_t s0, s1, t0, t1;
};
Ellipse ewaSetup(int W, int H, double s, double t, double ds0, double dt0, double ds1, double dt1) {
  Ellipse e;
  e.s = s * W - 0.5;
  e.t = t * H - 0.5;
  ds0 *= W; dt0 *= H; ds1 *= W; dt1 *= H;
  double A = dt0 * dt0 + dt1 * dt1 + 1;
  double B = -2.0 * (ds0 * dt0 + ds1 * dt1);
  double C = ds0 * ds0 + ds1 * ds1 + 1;
  const double invF = 1.0 / (A * C - B * B * 0.25);
  A *= invF; B *= invF; C *= invF;
  const double det = -B * B + 4.0 * A * C;
  const double invDet = 1.0 / det;
  const double uSqrt = std::sqrt(det * C), vSqrt = std::sqrt(A * det);
  e.s0 = (int64_t)std::ceil(e.s - 2.0 * invDet * uSqrt);
  e.s1 = (int64_t)std::floor(e.s + 2.0 * invDet * uSqrt);
  e.t0 = (int64_t)std::ceil(e.t - 2.0 * invDet * vSqrt);
  e.t1 = (int64_t)std::floor(e.t + 2.0 * invDet * vSqrt);
  e.A = A; e.B = B; e.C = C;
  return e;
}
inline double ewaWeight(double r2) {
  return gEwaLut[(int)std::fmin(r2 * 128, 127.0)];  // :316-317: min(r2 * SIZE, SIZE - 1).toInt()
}
}  // namespace

void TexImage::ewaS(int level, double s, double t, double ds0, double dt0, double ds1, double dt1, float out[3]) const {
  if (level >= levels) { texelS(levels - 1, 0, 0, out); return; }
  const Ellipse e = ewaSetup(w[level], h[level], s, t, ds0, dt0, ds1, dt1);
  float sum[3] = {0.f, 0.f, 0.f};
  double sumWts = 0.0;
  for (int64_t it = e.t0; it <= e.t1; ++it) {
    const double tt = it - e.t;
    for (int64_t si = e.s0; si <= e.s1; ++si) {
      const double ss = si - e.s;
      const double r2 = e.A * ss * ss + e.B * ss * tt + e.C * tt * tt;
      if (r2 < 1.0) {
        const double weight = ewaWeight(r2);
        float tx[3];
        texelS(level, si, it, tx);
        for (int k = 0; k < 3; ++k) sum[k] = f32((double)sum[k] + (double)f32((double)tx[k] * weight));
        sumWts += weight;
      }
    }
  }
  for (int k = 0; k < 3; ++k) out[k] = f32((double)sum[k] / sumWts);
}
double TexImage::ewaF(int level, double s, double t, double ds0, double dt0, double ds1, double dt1) const {
  if (level >= levels) return texelF(levels - 1, 0, 0);
  const Ellipse e = ewaSetup(w[level], h[level], s, t, ds0, dt0, ds1, dt1);
  double sum = 0.0, sumWts = 0.0;
  for (int64_t it = e.t0; it <= e.t1; ++it) {
    const double tt = it - e.t;
    for (int64_t si = e.s0; si <= e.s1; ++si) {
      const double ss = si - e.s;
      const double r2 = e.A * ss * ss + e.B * ss * tt + e.C * tt * tt;
      if (r2 < 1.0) {
        const double weight = ewaWeight(r2);
        sum += texelF(level, si, it) * weight;
        sumWts += weight;
      }
    }
  }
  return sum / sumWts;
}

// lookup2(), :224-268.  Returns through `lod == -1` the cases that end in a plain lookup.
namespace {
struct Lookup2 {
  int mode;  // 0: trilinear lookup(width), 1: triangle(0), 2: EWA blend
  double width, ds0, dt0, ds1, dt1, d;
  int ilod;
};
Lookup2 lookup2Setup(bool trilinear, double maxAniso, int levels, double ds0, double dt0, double ds1, double dt1) {
  Lookup2 r{};
  if (trilinear) {
    r.mode = 0;
    r.width = 2.0 * std::fmax(std::fmax(std::fabs(ds0), std::fabs(dt0)), std::fmax(std::fabs(ds1), std::fabs(dt1)));
    return r;
  }
  if (ds0 * ds0 + dt0 * dt0 < ds1 * ds1 + dt1 * dt1) { std::swap(ds0, ds1); std::swap(dt0, dt1); }
  const double majorLength = std::sqrt(ds0 * ds0 + dt0 * dt0);
  double minorLength = std::sqrt(ds1 * ds1 + dt1 * dt1);
  if (minorLength * maxAniso < majorLength && minorLength > 0.0) {
    const double scale = majorLength / (minorLength * maxAniso);
    ds1 *= scale; dt1 *= scale; minorLength *= scale;
  }
  if (minorLength == 0.0) { r.mode = 1; return r; }
  const double lod = std::fmax(0.0, levels - 1.0 + Log2d(minorLength));
  r.mode = 2;
  r.ilod = (int)std::floor(lod);
  r.d = lod - r.ilod;
  r.ds0 = ds0; r.dt0 = dt0; r.ds1 = ds1; r.dt1 = dt1;
  return r;
}
}  // namespace

void TexImage::lookup2S(double s, double t, double ds0, double dt0, double ds1, double dt1, float out[3]) const {
  const Lookup2 q = lookup2Setup(trilinear, maxAniso, levels, ds0, dt0, ds1, dt1);
  if (q.mode == 0) { lookupS(s, t, q.width, out); return; }
  if (q.mode == 1) { triangleS(0, s, t, out); return; }
  float a[3], b[3];
  ewaS(q.ilod, s, t, q.ds0, q.dt0, q.ds1, q.dt1, a);
  ewaS(q.ilod + 1, s, t, q.ds0, q.dt0, q.ds1, q.dt1, b);
  for (int k = 0; k < 3; ++k) out[k] = f32((double)f32((double)a[k] * (1.0 - q.d)) + (double)f32((double)b[k] * q.d));
}
double TexImage::lookup2F(double s, double t, double ds0, double dt0, double ds1, double dt1) const {
  const Lookup2 q = lookup2Setup(trilinear, maxAniso, levels, ds0, dt0, ds1, dt1);
  if (q.mode == 0) return lookupF(s, t, q.width);
  if (q.mode == 1) return triangleF(0, s, t);
  return ewaF(q.ilod, s, t, q.ds0, q.dt0, q.ds1, q.dt1) * (1.0 - q.d) + ewaF(q.ilod + 1, s, t, q.ds0, q.dt0, q.ds1, q.dt1) * q.d;
}

// ---- texture mappings -----------------------------------------------------------------------------------
namespace {
struct ST {
  double s, t, dsdx, dtdx, dsdy, dtdy;
};
inline double SphericalThetaV(const Vec& v) { return std::acos(clampd((double)v.z, -1.0, 1.0)); }  // vector.dart:185-187
inline double SphericalPhiV(const Vec& v) {                                                       // vector.dart:189-192
  double p = std::atan2((double)v.y, (double)v.x);
  return (p < 0.0) ? p + 2.0 * kPi : p;
}
void sphereST(const TextureNode& n, const Vec& p, double* s, double* t) {  // spherical_mapping_2d.dart:58-64
  Vec vec = Normalize(n.worldToTexture.point(p));
  *s = SphericalThetaV(vec) * INV_PI;
  *t = SphericalPhiV(vec) * INV_TWOPI;
}
void cylinderST(const TextureNode& n, const Vec& p, double* s, double* t) {  // cylindrical_mapping_2d.dart:56-60
  Vec vec = Normalize(n.worldToTexture.point(p));
  *s = (kPi + std::atan2((double)vec.y, (double)vec.x)) / (2.0 * kPi);
  *t = vec.z;
}
ST mapST(const TextureNode& n, const DG& dg) {
  ST r{};
  if (n.mapping == 0) {  // uv_mapping_2d.dart:25-36
    r.s = n.su * dg.u + n.du;
    r.t = n.sv * dg.v + n.dv;
    r.dsdx = n.su * dg.dudx; r.dtdx = n.sv * dg.dvdx;
    r.dsdy = n.su * dg.dudy; r.dtdy = n.sv * dg.dvdy;
  } else if (n.mapping == 1 || n.mapping == 2) {  // spherical / cylindrical: forward differences, delta 0.1 / 0.01
    const double delta = n.mapping == 1 ? 0.1 : 0.01;
    auto f = n.mapping == 1 ? sphereST : cylinderST;
    f(n, dg.p, &r.s, &r.t);
    double sx, tx, sy, ty;
    f(n, dg.p + dg.dpdx * delta, &sx, &tx);
    r.dsdx = (sx - r.s) / delta;
    r.dtdx = (tx - r.t) / delta;
    if (r.dtdx > 0.5) r.dtdx = 1.0 - r.dtdx;
    else if (r.dtdx < -0.5) r.dtdx = -(r.dtdx + 1.0);
    f(n, dg.p + dg.dpdy * delta, &sy, &ty);
    r.dsdy = (sy - r.s) / delta;
    r.dtdy = (ty - r.t) / delta;
    if (r.dtdy > 0.5) r.dtdy = 1.0 - r.dtdy;
    else if (r.dtdy < -0.5) r.dtdy = -(r.dtdy + 1.0);
  } else {  // planar_mapping_2d.dart:29-38
    r.s = n.du + Dot(dg.p, n.v1);
    r.t = n.dv + Dot(dg.p, n.v2);
    r.dsdx = Dot(dg.dpdx, n.v1); r.dtdx = Dot(dg.dpdx, n.v2);
    r.dsdy = Dot(dg.dpdy, n.v1); r.dtdy = Dot(dg.dpdy, n.v2);
  }
  return r;
}
// checkerboard_texture.dart:29-75: the weight of tex2 (0 or 1 when a single check is sampled)
double checkerWeight(const TextureNode& n, const ST& m, bool* single) {
  auto point = [&]() { return dartMod((int64_t)std::floor(m.s) + (int64_t)std::floor(m.t), 2) == 0 ? 0.0 : 1.0; };
  *single = true;
  if (n.aaMethod == 0) return point();
  const double ds = std::fmax(std::fabs(m.dsdx), std::fabs(m.dsdy)), dt = std::fmax(std::fabs(m.dtdx), std::fabs(m.dtdy));
  const double s0 = m.s - ds, s1 = m.s + ds, t0 = m.t - dt, t1 = m.t + dt;
  if (std::floor(s0) == std::floor(s1) && std::floor(t0) == std::floor(t1)) return point();
  auto BUMPINT = [](double x) { return std::floor(x / 2) + 2.0 * std::fmax((x / 2) - std::floor(x / 2) - 0.5, 0.0); };
  const double sint = (BUMPINT(s1) - BUMPINT(s0)) / (2.0 * ds), tint = (BUMPINT(t1) - BUMPINT(t0)) / (2.0 * dt);
  double area2 = sint + tint - 2.0 * sint * tint;
  if (ds > 1.0 || dt > 1.0) area2 = 0.5;
  *single = false;
  return area2;
}
}  // namespace

// ---- Perlin noise (lib/core/texture.dart:40-140) ----------------------------------------------------------------------------------
namespace {
// Ken Perlin's published permutation (the table of his 2002 reference implementation), which texture.dart:142-203 holds twice over
const uint8_t kNoisePerm[256] = {
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148,
    247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175,
    74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54,
    65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64,
    52, 217, 226, 250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213,
    119, 248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104,
    218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157,
    184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180};
inline int P(int i) { return kNoisePerm[i & 255]; }  // _NOISE_PERM has 512 entries: the second half repeats the first
inline double Grad(int x, int y, int z, double dx, double dy, double dz) {  // :119-125
  int h = P(P(P(x) + y) + z);
  h &= 15;
  double u = (h < 8 || h == 12 || h == 13) ? dx : dy;
  double v = (h < 4 || h == 12 || h == 13) ? dy : dz;
  return ((h & 1) != 0 ? -u : u) + ((h & 2) != 0 ? -v : v);
}
inline double NoiseWeight(double t) {  // :128-132
  double t3 = t * t * t, t4 = t3 * t;
  return 6.0 * t4 * t - 15.0 * t4 + 10.0 * t3;
}
inline double LerpN(double t, double a, double b) { return (1.0 - t) * a + t * b; }  // common.dart:80-81
double Noise(double x, double y = 0.5, double z = 0.5) {  // :40-77
  int ix = (int)std::floor(x), iy = (int)std::floor(y), iz = (int)std::floor(z);
  double dx = x - ix, dy = y - iy, dz = z - iz;
  ix &= 255; iy &= 255; iz &= 255;
  double w000 = Grad(ix, iy, iz, dx, dy, dz), w100 = Grad(ix + 1, iy, iz, dx - 1, dy, dz);
  double w010 = Grad(ix, iy + 1, iz, dx, dy - 1, dz), w110 = Grad(ix + 1, iy + 1, iz, dx - 1, dy - 1, dz);
  double w001 = Grad(ix, iy, iz + 1, dx, dy, dz - 1), w101 = Grad(ix + 1, iy, iz + 1, dx - 1, dy, dz - 1);
  double w011 = Grad(ix, iy + 1, iz + 1, dx, dy - 1, dz - 1), w111 = Grad(ix + 1, iy + 1, iz + 1, dx - 1, dy - 1, dz - 1);
  double wx = NoiseWeight(dx), wy = NoiseWeight(dy), wz = NoiseWeight(dz);
  double x00 = LerpN(wx, w000, w100), x10 = LerpN(wx, w010, w110), x01 = LerpN(wx, w001, w101), x11 = LerpN(wx, w011, w111);
  double y0 = LerpN(wy, x00, x10), y1 = LerpN(wy, x01, x11);
  return LerpN(wz, y0, y1);
}
inline double NoisePoint(const Vec& p) { return Noise(p.x, p.y, p.z); }
inline double SmoothStep(double mn, double mx, double value) {  // common.dart:131-134
  double v = clampd((value - mn) / (mx - mn), 0.0, 1.0);
  return v * v * (-2.0 * v + 3.0);
}
double FBm(const Vec& Pt, const Vec& dpdx, const Vec& dpdy, double omega, int maxOctaves) {  // :81-102
  double s2 = std::fmax(LengthSquared(dpdx), LengthSquared(dpdy));
  double log2_s2 = Log2d(s2);
  double foctaves = std::fmin((double)maxOctaves, std::fmax(0.0, -1.0 - 0.5 * log2_s2));
  int octaves = (int)std::floor(foctaves);
  double sum = 0.0, lambda = 1.0, o = 1.0;
  for (int i = 0; i < octaves; ++i) {
    sum += o * NoisePoint(Pt * lambda);
    lambda *= 1.99;
    o *= omega;
  }
  double partialOctave = foctaves - octaves;
  sum += o * SmoothStep(0.3, 0.7, partialOctave) * NoisePoint(Pt * lambda);
  return sum;
}
double Turbulence(const Vec& Pt, const Vec& dpdx, const Vec& dpdy, double omega, int maxOctaves) {  // :104-129
  double s2 = std::fmax(LengthSquared(dpdx), LengthSquared(dpdy));
  double foctaves = std::fmin((double)maxOctaves, std::fmax(0.0, -1.0 - 0.5 * Log2d(s2)));
  int octaves = (int)std::floor(foctaves);
  double sum = 0.0, lambda = 1.0, o = 1.0;
  for (int i = 0; i < octaves; ++i) {
    sum += o * std::fabs(NoisePoint(Pt * lambda));
    lambda *= 1.99;
    o *= omega;
  }
  double partialOctave = foctaves - octaves;
  sum += o * SmoothStep(0.3, 0.7, partialOctave) * std::fabs(NoisePoint(Pt * lambda));
  sum += (maxOctaves - foctaves) * 0.2;
  return sum;
}
// IdentityMapping3D.map (identity_mapping_3d.dart:25-29)
inline Vec map3D(const TextureNode& n, const DG& dg, Vec* dpdx, Vec* dpdy) {
  *dpdx = n.worldToTexture.vector(dg.dpdx);
  *dpdy = n.worldToTexture.vector(dg.dpdy);
  return n.worldToTexture.point(dg.p);
}
// the scalar of the noise textures: 7 fbm (fbm_texture.dart:26-32), 8 wrinkled (wrinkled_texture.dart:26-32), 9 windy (windy_texture.dart:26-38)
double noiseScalar(const TextureNode& n, const DG& dg) {
  Vec dpdx, dpdy;
  Vec Pt = map3D(n, dg, &dpdx, &dpdy);
  if (n.kind == 7) return FBm(Pt, dpdx, dpdy, n.value[0], n.aaMethod);
  if (n.kind == 8) return Turbulence(Pt, dpdx, dpdy, n.value[0], n.aaMethod);
  double windStrength = FBm(Pt * 0.1, dpdx * 0.1, dpdy * 0.1, 0.5, 3);
  double waveHeight = FBm(Pt, dpdx, dpdy, 0.5, 6);
  return std::fabs(windStrength) * waveHeight;
}
// dots_texture.dart:26-52: true = inside a dot
bool insideDot(double s, double t) {
  int sCell = (int)std::floor(s + 0.5), tCell = (int)std::floor(t + 0.5);
  if (Noise(sCell + 0.5, tCell + 0.5) > 0) {
    double radius = 0.35, maxShift = 0.5 - radius;
    double sCenter = sCell + maxShift * Noise(sCell + 1.5, tCell + 2.8);
    double tCenter = tCell + maxShift * Noise(sCell + 4.5, tCell + 9.8);
    double ds = s - sCenter, dt = t - tCenter;
    if (ds * ds + dt * dt < radius * radius) return true;
  }
  return false;
}
bool checker3D(const TextureNode& n, const DG& dg) {  // checkerboard_3d_texture.dart:26-35: true = tex1
  Vec dpdx, dpdy;
  Vec p = map3D(n, dg, &dpdx, &dpdy);
  return dartMod((int64_t)std::floor((double)p.x) + (int64_t)std::floor((double)p.y) + (int64_t)std::floor((double)p.z), 2) == 0;
}
}  // namespace

// ---- textures -----------------------------------------------------------------------------------------------
double TextureSet::evalFloat(int id, const DG& dg) const {
  const TextureNode& n = nodes[(size_t)id];
  switch (n.kind) {
    case 0: return n.value[0];  // constant_texture.dart
    case 1: {                   // scale_texture.dart:26-33: t2 * t1
      const double t1 = evalFloat(n.tex1, dg), t2 = evalFloat(n.tex2, dg);
      return t2 * t1;
    }
    case 2: {  // mix_texture.dart:26-31
      const double t1 = evalFloat(n.tex1, dg), t2 = evalFloat(n.tex2, dg), amt = evalFloat(n.amount, dg);
      return t1 * (1.0 - amt) + t2 * amt;
    }
    case 3: {  // image_texture.dart:76-86
      const ST m = mapST(n, dg);
      return images[(size_t)n.image].lookup2F(m.s, m.t, m.dsdx, m.dtdx, m.dsdy, m.dtdy);
    }
    case 4: {
      bool single;
      const double w2 = checkerWeight(n, mapST(n, dg), &single);
      if (single) return w2 == 0.0 ? evalFloat(n.tex1, dg) : evalFloat(n.tex2, dg);
      return evalFloat(n.tex1, dg) * (1.0 - w2) + evalFloat(n.tex2, dg) * w2;
    }
    case 6: {  // bilerp_texture.dart:26-40
      const ST m = mapST(n, dg);
      const double s = m.s, t = m.t;
      return n.value[0] * ((1.0 - s) * (1 - t)) + n.value2[0] * (1.0 - s) * t + n.value2[3] * s * (1.0 - t) + n.value2[6] * s * t;
    }
    case 7: case 8: case 9: return noiseScalar(n, dg);
    case 11: {  // DotsTexture(mapping, outsideDot = tex1, insideDot = tex2)
      const ST m = mapST(n, dg);
      return evalFloat(insideDot(m.s, m.t) ? n.tex2 : n.tex1, dg);
    }
    case 12: return evalFloat(checker3D(n, dg) ? n.tex1 : n.tex2, dg);
    default: return 0.0;
  }
}

void TextureSet::evalSpec(int id, const DG& dg, float out[3]) const {
  const TextureNode& n = nodes[(size_t)id];
  auto mulS = [](const float a[3], double s, float o[3]) { for (int k = 0; k < 3; ++k) o[k] = f32((double)a[k] * s); };
  auto addS = [](const float a[3], const float b[3], float o[3]) { for (int k = 0; k < 3; ++k) o[k] = f32((double)a[k] + (double)b[k]); };
  switch (n.kind) {
    case 0:
      for (int k = 0; k < 3; ++k) out[k] = f32(n.value[k]);
      return;
    case 1: {  // t1 * t2, RGBColor * RGBColor
      float a[3], b[3];
      evalSpec(n.tex1, dg, a);
      evalSpec(n.tex2, dg, b);
      for (int k = 0; k < 3; ++k) out[k] = f32((double)a[k] * (double)b[k]);
      return;
    }
    case 2: {
      float a[3], b[3], x[3], y[3];
      evalSpec(n.tex1, dg, a);
      evalSpec(n.tex2, dg, b);
      const double amt = evalFloat(n.amount, dg);
      mulS(a, 1.0 - amt, x);
      mulS(b, amt, y);
      addS(x, y, out);
      return;
    }
    case 3: {
      const ST m = mapST(n, dg);
      images[(size_t)n.image].lookup2S(m.s, m.t, m.dsdx, m.dtdx, m.dsdy, m.dtdy, out);
      return;
    }
    case 4: {
      bool single;
      const double w2 = checkerWeight(n, mapST(n, dg), &single);
      if (single) { evalSpec(w2 == 0.0 ? n.tex1 : n.tex2, dg, out); return; }
      float a[3], b[3], x[3], y[3];
      evalSpec(n.tex1, dg, a);
      evalSpec(n.tex2, dg, b);
      mulS(a, 1.0 - w2, x);
      mulS(b, w2, y);
      addS(x, y, out);
      return;
    }
    case 5: {  // uv_texture.dart:26-37
      const ST m = mapST(n, dg);
      out[0] = f32(m.s - std::floor(m.s)); out[1] = f32(m.t - std::floor(m.t)); out[2] = 0.f;
      return;
    }
    case 6: {  // v00 * ((1 - s) * (1 - t)) + v01 * (1 - s) * t + v10 * s * (1 - t) + v11 * s * t
      const ST m = mapST(n, dg);
      const double s = m.s, t = m.t;
      float v00[3], v01[3], v10[3], v11[3], a[3], b[3], c[3];
      for (int k = 0; k < 3; ++k) { v00[k] = f32(n.value[k]); v01[k] = f32(n.value2[k]); v10[k] = f32(n.value2[3 + k]); v11[k] = f32(n.value2[6 + k]); }
      mulS(v00, (1.0 - s) * (1 - t), a);
      mulS(v01, 1.0 - s, b); mulS(b, t, b);
      addS(a, b, a);
      mulS(v10, s, c); mulS(c, 1.0 - t, c);
      addS(a, c, a);
      mulS(v11, s, c); mulS(c, t, c);
      addS(a, c, out);
      return;
    }
    case 7: case 8: case 9: {  // new Spectrum(n)
      out[0] = out[1] = out[2] = f32(noiseScalar(n, dg));
      return;
    }
    case 10: {  // marble_texture.dart:27-66
      Vec dpdx, dpdy;
      Vec Pt = map3D(n, dg, &dpdx, &dpdy);
      const double scale = n.value[1], variation = n.value[2];
      Pt = Pt * scale;
      double marble = (double)Pt.y + variation * FBm(Pt, dpdx * scale, dpdy * scale, n.value[0], n.aaMethod);
      double t = 0.5 + 0.5 * std::sin(marble);
      static const double c[27] = {0.58, 0.58, 0.6, 0.58, 0.58, 0.6, 0.58, 0.58, 0.6, 0.5, 0.5, 0.5, 0.6, 0.59, 0.58,
                                   0.58, 0.58, 0.6, 0.58, 0.58, 0.6, 0.2, 0.2, 0.33, 0.58, 0.58, 0.6};
      const int NSEG = 9 - 3;
      int first = (int)std::floor(t * NSEG);
      t = (t * NSEG - first);
      const int ci = first * 3;
      float c0[3], c1[3], c2[3], c3[3], s0[3], s1[3], s2[3], a[3], b[3];
      for (int k = 0; k < 3; ++k) { c0[k] = f32(c[ci + k]); c1[k] = f32(c[ci + 3 + k]); c2[k] = f32(c[ci + 6 + k]); c3[k] = f32(c[ci + 9 + k]); }
      auto bez = [&](const float* x, const float* y, float* o) { mulS(x, 1.0 - t, a); mulS(y, t, b); addS(a, b, o); };
      bez(c0, c1, s0); bez(c1, c2, s1); bez(c2, c3, s2);
      float r0[3], r1[3], r[3];
      bez(s0, s1, r0); bez(s1, s2, r1);
      bez(r0, r1, r);
      mulS(r, 1.5, out);
      return;
    }
    case 11: {
      const ST m = mapST(n, dg);
      evalSpec(insideDot(m.s, m.t) ? n.tex2 : n.tex1, dg, out);
      return;
    }
    case 12: evalSpec(checker3D(n, dg) ? n.tex1 : n.tex2, dg, out); return;
    default:
      out[0] = out[1] = out[2] = 0.f;
  }
}

// ---- Material.Bump (material.dart:35-88) --------------------------------------------------------------------
void Bump(const TextureSet& ts, int d, const DG& dgGeom, const DG& dgs, DG* dgBump) {
  DG dgEval = dgs;
  double du = 0.5 * (std::fabs(dgs.dudx) + std::fabs(dgs.dudy));
  if (du == 0.0) du = 0.01;
  dgEval.p = dgs.p + dgs.dpdu * du;
  dgEval.u = dgs.u + du;
  dgEval.nn = Normalize(Cross(dgs.dpdu, dgs.dpdv) + dgs.dndu * du);
  const double uDisplace = ts.evalFloat(d, dgEval);
  double dv = 0.5 * (std::fabs(dgs.dvdx) + std::fabs(dgs.dvdy));
  if (dv == 0.0) dv = 0.01;
  dgEval.p = dgs.p + dgs.dpdv * dv;
  dgEval.u = dgs.u;
  dgEval.v = dgs.v + dv;
  dgEval.nn = Normalize(Cross(dgs.dpdu, dgs.dpdv) + dgs.dndv * dv);
  const double vDisplace = ts.evalFloat(d, dgEval);
  const double displace = ts.evalFloat(d, dgs);
  *dgBump = dgs;
  dgBump->dpdu = dgs.dpdu + dgs.nn * (uDisplace - displace) / du + dgs.dndu * displace;
  dgBump->dpdv = dgs.dpdv + dgs.nn * (vDisplace - displace) / dv + dgs.dndv * displace;
  dgBump->nn = Normalize(Cross(dgBump->dpdu, dgBump->dpdv));
  if (dgs.reverse) dgBump->nn = dgBump->nn * -1.0;
  dgBump->nn = FaceForward(dgBump->nn, dgGeom.nn);
}

}  // namespace orc
