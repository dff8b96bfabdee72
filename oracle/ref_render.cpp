// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.h / ref_render.h headers).  PARITY UNPINNED.
#include "ref_render.h"

#include <algorithm>
#include <cassert>
#include <thread>

namespace orc {

// BSDF flags, lib/core/reflection/bsdf.dart:23-31
enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };

static inline double Lerp(double t, double a, double b) { return (1.0 - t) * a + t * b; }  // common.dart:80-81

// ---------------------------------------------------------------------------------------------------
// montecarlo.dart
void Distribution1D::init(const std::vector<double>& f) {  // :26-48 (func, cdf are Float32Lists)
  count = (int)f.size();
  func.resize(count);
  for (int i = 0; i < count; ++i) func[i] = f32(f[i]);
  cdf.assign(count + 1, 0.f);
  for (int i = 1; i < count + 1; ++i) cdf[i] = f32((double)cdf[i - 1] + (double)func[i - 1] / count);
  funcInt = cdf[count];
  if (funcInt == 0.0) {
    for (int i = 1; i < count + 1; ++i) cdf[i] = f32((double)i / count);
  } else {
    for (int i = 1; i < count + 1; ++i) cdf[i] = f32((double)cdf[i] / funcInt);
  }
}
int Distribution1D::sampleDiscrete(double u) const {  // :82-92, upper_bound over cdf[0..count]
  int ptr = (int)(std::upper_bound(cdf.begin(), cdf.begin() + count + 1, u,
                                   [](double v, float c) { return v < (double)c; }) - cdf.begin());
  return std::max(0, ptr - 1);
}

double Distribution1D::sampleContinuous(double u, double* pdf, int* off) const {  // :50-80
  int ptr = (int)(std::upper_bound(cdf.begin(), cdf.begin() + count + 1, u,
                                   [](double v, float c) { return v < (double)c; }) - cdf.begin());
  int offset = std::max(0, ptr - 1);
  if (offset == count) offset = count - 1;
  if (off) *off = offset;
  double dc = (double)cdf[offset + 1] - (double)cdf[offset];
  double du = 0.0;
  if (dc != 0.0) du = (u - (double)cdf[offset]) / dc;
  if (pdf) *pdf = (double)func[offset] / funcInt;
  return (offset + du) / count;
}

void Distribution2D::init(const std::vector<float>& data, int nu, int nv) {
  pConditionalV.assign(nv, Distribution1D());
  std::vector<double> marginal(nv);
  for (int v = 0; v < nv; ++v) {
    std::vector<double> l(data.begin() + (size_t)v * nu, data.begin() + (size_t)v * nu + nu);
    pConditionalV[v].init(l);
    marginal[v] = f32(pConditionalV[v].funcInt);  // Float32List marginalFunc
  }
  pMarginal.init(marginal);
}
void Distribution2D::sampleContinuous(double u0, double u1, double uv[2], double* pdf) const {
  double pdfs1, pdfs0;
  int v;
  uv[1] = pMarginal.sampleContinuous(u1, &pdfs1, &v);
  uv[0] = pConditionalV[v].sampleContinuous(u0, &pdfs0, nullptr);
  *pdf = pdfs0 * pdfs1;
}
double Distribution2D::pdf(double u, double v) const {
  const int cu = pConditionalV[0].count, cv = pMarginal.count;
  // (u * count).toInt().clamp(0, count - 1): toInt truncates toward zero (and throws on NaN / inf, which a sampled
  // direction cannot produce here)
  int iu = (int)std::min<double>(std::max<double>(std::trunc(u * cu), 0.0), cu - 1);
  int iv = (int)std::min<double>(std::max<double>(std::trunc(v * cv), 0.0), cv - 1);
  if (pConditionalV[iv].funcInt * pMarginal.funcInt == 0.0) return 0.0;
  return ((double)pConditionalV[iv].func[iu] * (double)pMarginal.func[iv]) / (pConditionalV[iv].funcInt * pMarginal.funcInt);
}

static inline double Log2(double x) { return std::log(x) * (1.0 / std::log(2.0)); }  // common.dart:98-103

void MipMap::init(int width, int height, const float* rgb) {
  levels = 1 + (int)Log2(std::max(width, height));  // mipmap.dart:143
  pyramid.assign(levels, {});
  w.assign(levels, 1);
  h.assign(levels, 1);
  w[0] = width; h[0] = height;
  pyramid[0].resize((size_t)width * height);
  for (size_t i = 0; i < pyramid[0].size(); ++i) pyramid[0][i] = Spec(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
  for (int i = 1; i < levels; ++i) {  // :152-166
    int sRes = std::max(1, w[i - 1] / 2), tRes = std::max(1, h[i - 1] / 2);
    w[i] = sRes; h[i] = tRes;
    pyramid[i].resize((size_t)sRes * tRes);
    for (int t = 0, p = 0; t < tRes; ++t)
      for (int s2 = 0; s2 < sRes; ++s2, ++p)
        pyramid[i][p] = (texel(i - 1, 2 * s2, 2 * t) + texel(i - 1, 2 * s2 + 1, 2 * t) + texel(i - 1, 2 * s2, 2 * t + 1) +
                         texel(i - 1, 2 * s2 + 1, 2 * t + 1)) * 0.25;
  }
}
Spec MipMap::texel(int level, int64_t s, int64_t t) const {  // :183-204, TEXTURE_REPEAT; Dart's % is never negative
  const int64_t W = w[level], H = h[level];
  s = ((s % W) + W) % W;
  t = ((t % H) + H) % H;
  return pyramid[level][(size_t)(t * W + s)];
}
Spec MipMap::triangle(int level, double s, double t) const {  // :341-355
  level = std::min(std::max(level, 0), levels - 1);
  s = s * w[level] - 0.5;
  t = t * h[level] - 0.5;
  int64_t s0 = (int64_t)std::floor(s), t0 = (int64_t)std::floor(t);
  double ds = s - s0, dt = t - t0;
  return texel(level, s0, t0) * ((1.0 - ds) * (1.0 - dt)) + texel(level, s0, t0 + 1) * ((1.0 - ds) * dt) +
         texel(level, s0 + 1, t0) * (ds * (1.0 - dt)) + texel(level, s0 + 1, t0 + 1) * (ds * dt);
}
Spec MipMap::lookup(double s, double t, double width) const {  // :206-222
  double level = levels - 1 + Log2(std::fmax(width, 1.0e-8));
  if (level < 0) return triangle(0, s, t);
  if (level >= levels - 1) return texel(levels - 1, 0, 0);
  int iLevel = (int)std::floor(level);
  double delta = level - iLevel;
  return triangle(iLevel, s, t) * (1.0 - delta) + triangle(iLevel + 1, s, t) * delta;
}

void Light::setRadianceMap(int width, int height, const float* rgb) {  // infinite_area_light.dart:276-306
  radianceMap.init(width, height, rgb);
  double filter = 1.0 / std::max(width, height);
  std::vector<float> img((size_t)width * height);
  for (int v = 0; v < height; ++v) {
    double vp = (double)v / height;
    double sinTheta = std::sin(kPi * (v + 0.5) / height);
    for (int u = 0; u < width; ++u) {
      double up = (double)u / width;
      img[u + (size_t)v * width] = f32(radiance(up, vp, filter).luminance());
      img[u + (size_t)v * width] = f32((double)img[u + (size_t)v * width] * sinTheta);
    }
  }
  distribution.init(img, width, height);
}

static inline Vec UniformSampleSphere(double u1, double u2) {  // :113-120
  double z = 1.0 - 2.0 * u1;
  double r = std::sqrt(std::fmax(0.0, 1.0 - z * z));
  double phi = 2.0 * kPi * u2;
  return Vec(r * std::cos(phi), r * std::sin(phi), z);
}
static inline Vec UniformSampleCone2(double u1, double u2, double costhetamax, const Vec& x, const Vec& y, const Vec& z) {
  double costheta = Lerp(u1, costhetamax, 1.0);  // :135-142
  double sintheta = std::sqrt(1.0 - costheta * costheta);
  double phi = u2 * 2.0 * kPi;
  return x * (std::cos(phi) * sintheta) + y * (std::sin(phi) * sintheta) + z * costheta;
}
static inline double UniformConePdf(double cosThetaMax) { return 1.0 / (2.0 * kPi * (1.0 - cosThetaMax)); }
static void ConcentricSampleDisk(double u1, double u2, double* dx, double* dy) {  // :155-201
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
  theta *= kPi / 4.0;
  *dx = r * std::cos(theta);
  *dy = r * std::sin(theta);
}
static inline Vec CosineSampleHemisphere(double u1, double u2) {  // :203-209
  double dx, dy;
  ConcentricSampleDisk(u1, u2, &dx, &dy);
  double z = std::sqrt(std::fmax(0.0, 1.0 - dx * dx - dy * dy));
  return Vec(dx, dy, z);
}
static inline double PowerHeuristic(int nf, double fPdf, int ng, double gPdf) {  // :480-484
  double f = nf * fPdf, g = ng * gPdf;
  return (f * f) / (f * f + g * g);
}
static inline double Sobol2(uint32_t n, uint32_t scramble) {  // :486-493
  for (uint32_t v = 1u << 31; n != 0; n >>= 1, v ^= v >> 1)
    if (n & 0x1) scramble ^= v;
  return std::fmin(((scramble >> 8) & 0xffffff) / (double)(1 << 24), ONE_MINUS_EPSILON);
}
static inline double VanDerCorput(uint32_t n, uint32_t scramble) {  // :495-504
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
  n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
  n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
  n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
  n ^= scramble;
  return std::fmin(((n >> 8) & 0xffffff) / (double)(1 << 24), ONE_MINUS_EPSILON);
}
static void Shuffle(float* samples, int offset, int count, int dims, Rng& rng) {  // :294-303
  for (int i = 0; i < count; ++i) {
    int other = i + (int)(rng.randomUint() % (uint32_t)(count - i));
    for (int j = 0; j < dims; ++j) std::swap(samples[offset + dims * i + j], samples[offset + dims * other + j]);
  }
}
static void LDShuffleScrambled1D(int nSamples, int nPixel, float* samples, Rng& rng) {  // :524-536
  uint32_t scramble = rng.randomUint();
  for (int i = 0; i < nSamples * nPixel; ++i) samples[i] = f32(VanDerCorput(i, scramble));
  for (int i = 0; i < nPixel; ++i) Shuffle(samples, i * nSamples, nSamples, 1, rng);
  Shuffle(samples, 0, nPixel, nSamples, rng);
}
static void LDShuffleScrambled2D(int nSamples, int nPixel, float* samples, Rng& rng) {  // :539-551
  uint32_t s0 = rng.randomUint(), s1 = rng.randomUint();
  for (int i = 0; i < nSamples * nPixel; ++i) {
    samples[2 * i] = f32(VanDerCorput(i, s0));
    samples[2 * i + 1] = f32(Sobol2(i, s1));
  }
  for (int i = 0; i < nPixel; ++i) Shuffle(samples, 2 * i * nSamples, nSamples, 2, rng);
  Shuffle(samples, 0, nPixel, 2 * nSamples, rng);
}
static void StratifiedSample1D(float* s, int n, Rng& rng, bool jitter) {  // :270-277
  double invTot = 1.0 / n;
  for (int i = 0; i < n; ++i) {
    double delta = jitter ? rng.randomFloat() : 0.5;
    s[i] = f32(std::fmin((i + delta) * invTot, ONE_MINUS_EPSILON));
  }
}
static void StratifiedSample2D(float* s, int nx, int ny, Rng& rng, bool jitter) {  // :279-292
  double dx = 1.0 / nx, dy = 1.0 / ny;
  int si = 0;
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x) {
      double jx = jitter ? rng.randomFloat() : 0.5;
      double jy = jitter ? rng.randomFloat() : 0.5;
      s[si++] = f32(std::fmin((x + jx) * dx, ONE_MINUS_EPSILON));
      s[si++] = f32(std::fmin((y + jy) * dy, ONE_MINUS_EPSILON));
    }
}
static void LatinHypercube(float* samples, int nSamples, int nDim, Rng& rng) {  // :305-325
  double delta = 1.0 / nSamples;
  for (int i = 0; i < nSamples; ++i)
    for (int j = 0; j < nDim; ++j) samples[nDim * i + j] = f32(std::fmin((i + rng.randomFloat()) * delta, ONE_MINUS_EPSILON));
  for (int i = 0; i < nDim; ++i)
    for (int j = 0; j < nSamples; ++j) {
      int other = j + (int)(rng.randomUint() % (uint32_t)(nSamples - j));
      std::swap(samples[nDim * j + i], samples[nDim * other + i]);
    }
}
static inline int RoundUpPow2(int v) {  // common.dart:117-125
  v--;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v + 1;
}

// matte_material.dart:41-65 with constant textures
Material Material::matte(const Spec& kd, double sigma) {
  Material m;
  Spec r(clampd(kd.c[0], 0.0, kInf), clampd(kd.c[1], 0.0, kInf), clampd(kd.c[2], 0.0, kInf));
  double sig = clampd(sigma, 0.0, 90.0);
  if (!r.isBlack()) {
    Lobe l;
    l.R = r;
    if (sig == 0.0) l.kind = 0;
    else { l.kind = 1; l.param = sig; }
    m.lobes.push_back(l);
  }
  return m;
}

// ---------------------------------------------------------------------------------------------------
// film, lib/film/image_film.dart
void Film::configure() {  // :51-97
  left = (int)std::ceil(xres * crop[0]);
  width = std::max(1, (int)std::ceil(xres * crop[1]) - left);
  top = (int)std::ceil(yres * crop[2]);
  height = std::max(1, (int)std::ceil(yres * crop[3]) - top);
  invXWidth = 1.0 / xWidth;  // filter.dart:26-39
  invYWidth = 1.0 / yWidth;
  Lxyz.assign((size_t)width * height * 3, 0.f);
  weightSum.assign((size_t)width * height, 0.f);
}
void Film::getSampleExtent(int e[4]) const {  // :247-252
  e[0] = (int)std::floor(left + 0.5 - xWidth);
  e[1] = (int)std::ceil(left + 0.5 + width + xWidth);
  e[2] = (int)std::floor(top + 0.5 - yWidth);
  e[3] = (int)std::ceil(top + 0.5 + height + yWidth);
}
void Film::addSample(double imageX, double imageY, const Spec& L) {  // :99-150
  double dimageX = imageX - 0.5, dimageY = imageY - 0.5;
  int x0 = (int)std::ceil(dimageX - xWidth), x1 = (int)std::floor(dimageX + xWidth);
  int y0 = (int)std::ceil(dimageY - yWidth), y1 = (int)std::floor(dimageY + yWidth);
  x0 = std::max(x0, left); x1 = std::min(x1, left + width - 1);
  y0 = std::max(y0, top); y1 = std::min(y1, top + height - 1);
  if ((x1 - x0) < 0 || (y1 - y0) < 0) return;
  // L.toXYZ(): XYZColor (float32) via spectrum.dart:293-297
  double r = L.c[0], g = L.c[1], b = L.c[2];
  float xyz[3] = {f32(0.412453 * r + 0.357580 * g + 0.180423 * b), f32(0.212671 * r + 0.715160 * g + 0.072169 * b),
                  f32(0.019334 * r + 0.119193 * g + 0.950227 * b)};
  for (int y = y0; y <= y1; ++y) {
    double fy = std::fabs((y - dimageY) * invYWidth * 16);
    int iy = std::min((int)std::floor(fy), 15);
    for (int x = x0; x <= x1; ++x) {
      double fx = std::fabs((x - dimageX) * invXWidth * 16);
      int ix = std::min((int)std::floor(fx), 15);
      double filterWt = table[iy * 16 + ix];
      size_t pi = (size_t)(y - top) * width + (x - left);
      Lxyz[3 * pi] = f32((double)Lxyz[3 * pi] + filterWt * xyz[0]);
      Lxyz[3 * pi + 1] = f32((double)Lxyz[3 * pi + 1] + filterWt * xyz[1]);
      Lxyz[3 * pi + 2] = f32((double)Lxyz[3 * pi + 2] + filterWt * xyz[2]);
      weightSum[pi] = f32((double)weightSum[pi] + filterWt);
    }
  }
}
void Film::writeImage(float* rgb) const {  // :268-299 (OutputImage.rgb is a Float32List)
  for (size_t pi = 0; pi < (size_t)width * height; ++pi) {
    double x = Lxyz[3 * pi], y = Lxyz[3 * pi + 1], z = Lxyz[3 * pi + 2];
    double c0 = 3.240479 * x - 1.537150 * y - 0.498535 * z;
    double c1 = -0.969256 * x + 1.875991 * y + 0.041556 * z;
    double c2 = 0.055648 * x - 0.204043 * y + 1.057311 * z;
    double w = weightSum[pi];
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    if (w != 0.0) {
      double invWt = 1.0 / w;
      o0 = f32(std::fmax(0.0, c0 * invWt)); o1 = f32(std::fmax(0.0, c1 * invWt)); o2 = f32(std::fmax(0.0, c2 * invWt));
    }
    rgb[3 * pi] = o0; rgb[3 * pi + 1] = o1; rgb[3 * pi + 2] = o2;  // + splatScale * 0
  }
}

// ---------------------------------------------------------------------------------------------------
// shading geometry
namespace {

// struct DG: the whole DifferentialGeometry, ref_texture.h

struct Isect {
  DG dg;
  int32_t prim = -1;
  double rayEpsilon = 0;
  int32_t inst = -1;  // the TransformedPrimitive the hit came through and its worldToPrimitive at the ray's time
  Transform w2p;
};

struct SampleLayout {
  std::vector<int> n1D, n2D;
  int add1D(int n) { n1D.push_back(n); return (int)n1D.size() - 1; }
  int add2D(int n) { n2D.push_back(n); return (int)n2D.size() - 1; }
  int floatsPerSample() const {
    int n = 5;
    for (int v : n1D) n += v;
    for (int v : n2D) n += 2 * v;
    return n;
  }
};

struct SampleVals {
  double imageX = 0, imageY = 0, lensU = 0, lensV = 0, time = 0;
  std::vector<std::vector<float>> oneD, twoD;
  void alloc(const SampleLayout& l) {
    oneD.resize(l.n1D.size());
    twoD.resize(l.n2D.size());
    for (size_t i = 0; i < l.n1D.size(); ++i) oneD[i].assign(l.n1D[i], 0.f);
    for (size_t i = 0; i < l.n2D.size(); ++i) twoD[i].assign(2 * l.n2D[i], 0.f);
  }
};

struct SampleOffsets {  // LightSampleOffsets / BSDFSampleOffsets
  int nSamples = 0, componentOffset = 0, posOffset = 0;
};

struct U3 {  // LightSample / BSDFSample: two float32 values + one f64 component
  float u0 = 0, u1 = 0;
  double comp = 0;
  static U3 random(Rng& rng) {  // bsdf_sample.dart:38-44, light_sample.dart:45-51
    U3 s;
    s.u0 = f32(rng.randomFloat());
    s.u1 = f32(rng.randomFloat());
    s.comp = rng.randomFloat();
    return s;
  }
  static U3 fromSample(const SampleVals& sv, const SampleOffsets& o, int n) {
    U3 s;
    s.u0 = sv.twoD[o.posOffset][2 * n];
    s.u1 = sv.twoD[o.posOffset][2 * n + 1];
    s.comp = sv.oneD[o.componentOffset][n];
    return s;
  }
};

// dart:math min / max return NaN when either argument is NaN
static inline double dmin(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? std::nan("") : (a < b ? a : b); }
static inline double dmax(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? std::nan("") : (a > b ? a : b); }

// One BxDF with its per-hit constants (lib/core/reflection/*.dart).  Local frame: z = shading normal.
struct LobeEval {
  Lobe l;
  int type = 0;
  double A = 0, B = 0;  // OrenNayar (oren_nayar.dart:24-31)

  static double CosTheta(const Vec& v) { return v.z; }
  static double AbsCosTheta(const Vec& v) { return std::fabs((double)v.z); }
  static double SinTheta2(const Vec& v) { return std::fmax(0.0, 1.0 - CosTheta(v) * CosTheta(v)); }
  static double SinTheta(const Vec& v) { return std::sqrt(SinTheta2(v)); }
  static double CosPhi(const Vec& v) { double s = SinTheta(v); return s == 0.0 ? 1.0 : clampd((double)v.x / s, -1.0, 1.0); }
  static double SinPhi(const Vec& v) { double s = SinTheta(v); return s == 0.0 ? 0.0 : clampd((double)v.y / s, -1.0, 1.0); }
  static bool SameHemisphere(const Vec& w, const Vec& wp) { return (double)w.z * wp.z > 0.0; }  // vector.dart:194-196

  void init(const Lobe& lobe) {
    l = lobe;
    switch (l.kind) {
      case 0: type = BSDF_REFLECTION | BSDF_DIFFUSE; break;
      case 1: {
        type = BSDF_REFLECTION | BSDF_DIFFUSE;
        double sigma = Radians(l.param), sigma2 = sigma * sigma;
        A = 1.0 - (sigma2 / (2.0 * (sigma2 + 0.33)));
        B = 0.45 * sigma2 / (sigma2 + 0.09);
        break;
      }
      case 2: type = BSDF_REFLECTION | BSDF_GLOSSY; break;
      case 3: type = BSDF_REFLECTION | BSDF_SPECULAR; break;
      case 5: case 6: case 7: type = BSDF_REFLECTION | BSDF_GLOSSY; break;
      default: type = BSDF_TRANSMISSION | BSDF_SPECULAR; break;
    }
    if (l.wrap & 1) type ^= (BSDF_REFLECTION | BSDF_TRANSMISSION);  // brdf_to_btdf.dart:27-29; ScaledBxDF keeps the type
  }
  bool matches(int flags) const { return (type & flags) == type; }  // bxdf.dart:31-33

  // fresnel_dielectric.dart:24-56, fresnel_conductor.dart:24-49, fresnel_no_op.dart
  static Spec dielectric(double cosi, double eta_i, double eta_t) {
    cosi = std::isnan(cosi) ? cosi : clampd(cosi, -1.0, 1.0);
    bool entering = cosi > 0.0;
    double ei = eta_i, et = eta_t;
    if (!entering) std::swap(ei, et);
    double sint = ei / et * std::sqrt(dmax(0.0, 1.0 - cosi * cosi));
    if (sint >= 1.0) return Spec(1.0);
    double cost = std::sqrt(dmax(0.0, 1.0 - sint * sint));
    cosi = std::fabs(cosi);
    double Rparl = ((et * cosi) - (ei * cost)) / ((et * cosi) + (ei * cost));
    double Rperp = ((ei * cosi) - (et * cost)) / ((ei * cosi) + (et * cost));
    return Spec((Rparl * Rparl + Rperp * Rperp) / 2.0);
  }
  Spec fresnel(double cosi) const {
    if (l.fresnel == 0) return Spec(1.0);
    if (l.fresnel == 1) return dielectric(cosi, l.ei, l.et);
    cosi = std::fabs(cosi);
    const Spec ONE(1.0), cosSqr(cosi * cosi);
    const Spec& eta = l.eta;
    const Spec& k = l.k;
    Spec tmp = (eta * eta + k * k) * (cosi * cosi);
    Spec r1 = (tmp - (eta * (2.0 * cosi)) + ONE);
    Spec r2 = (tmp + (eta * (2.0 * cosi)) + ONE);
    Spec Rparl2 = r1 / r2;
    Spec tmp_f = eta * eta + k * k;
    r1 = (tmp_f - (eta * (2.0 * cosi)) + cosSqr);
    r2 = (tmp_f + (eta * (2.0 * cosi)) + cosSqr);
    Spec Rperp2 = r1 / r2;
    return (Rparl2 + Rperp2) / 2.0;
  }

  // blinn.dart:31-71
  double blinnD(const Vec& wh) const {
    double costhetah = AbsCosTheta(wh);
    return (l.param + 2.0) * INV_TWOPI * std::pow(costhetah, l.param);
  }
  double blinnSample(const Vec& wo, Vec* wi, double u1, double u2) const {
    const double exponent = l.param;
    double costheta = std::pow(u1, 1.0 / (exponent + 1.0));
    double sintheta = std::sqrt(dmax(0.0, 1.0 - costheta * costheta));
    double phi = u2 * 2.0 * kPi;
    Vec wh(sintheta * std::cos(phi), sintheta * std::sin(phi), costheta);  // vector.dart:172-176
    if (!SameHemisphere(wo, wh)) wh = -wh;
    *wi = -wo + wh * 2.0 * Dot(wo, wh);
    double pdf = ((exponent + 1.0) * std::pow(costheta, exponent)) / (2.0 * kPi * 4.0 * Dot(wo, wh));
    if (Dot(wo, wh) <= 0.0) pdf = 0.0;
    return pdf;
  }
  double blinnPdf(const Vec& wo, const Vec& wi) const {
    const double exponent = l.param;
    Vec wh = Normalize(wo + wi);
    double costheta = AbsCosTheta(wh);
    double pdf = ((exponent + 1.0) * std::pow(costheta, exponent)) / (2.0 * kPi * 4.0 * Dot(wo, wh));
    if (Dot(wo, wh) <= 0.0) pdf = 0.0;
    return pdf;
  }

  // anisotropic.dart:27-121 (ex = l.param, ey = l.ei, both already clamped to 10000 by the constructor, :30-37)
  double anisoD(const Vec& wh) const {
    double costhetah = std::fabs((double)wh.z);
    double d = 1.0 - costhetah * costhetah;
    if (d == 0.0) return 0.0;
    double e = (l.param * wh.x * wh.x + l.ei * wh.y * wh.y) / d;
    return std::sqrt((l.param + 2.0) * (l.ei + 2.0)) * INV_TWOPI * std::pow(costhetah, e);
  }
  double anisoPdfOf(const Vec& wo, const Vec& wh) const {
    double costhetah = AbsCosTheta(wh);
    double ds = 1.0 - costhetah * costhetah;
    double p = 0.0;
    if (ds > 0.0 && Dot(wo, wh) > 0.0) {
      double e = (l.param * wh.x * wh.x + l.ei * wh.y * wh.y) / ds;
      double d = std::sqrt((l.param + 1.0) * (l.ei + 1.0)) * INV_TWOPI * std::pow(costhetah, e);
      p = d / (4.0 * Dot(wo, wh));
    }
    return p;
  }
  double anisoPdf(const Vec& wo, const Vec& wi) const { return anisoPdfOf(wo, Normalize(wo + wi)); }
  void anisoFirstQuadrant(double u1, double u2, double* phi, double* costheta) const {
    const double ex = l.param, ey = l.ei;
    if (ex == ey) *phi = kPi * u1 * 0.5;
    else *phi = std::atan(std::sqrt((ex + 1.0) / (ey + 1.0)) * std::tan(kPi * u1 * 0.5));
    double cosphi = std::cos(*phi), sinphi = std::sin(*phi);
    *costheta = std::pow(u2, 1.0 / (ex * cosphi * cosphi + ey * sinphi * sinphi + 1.0));
  }
  double anisoSample(const Vec& wo, Vec* wi, double u1, double u2) const {
    double phi, cosTheta;
    if (u1 < 0.25) {
      anisoFirstQuadrant(4.0 * u1, u2, &phi, &cosTheta);
    } else if (u1 < 0.5) {
      u1 = 4.0 * (0.5 - u1);
      anisoFirstQuadrant(u1, u2, &phi, &cosTheta);
      phi = kPi - phi;
    } else if (u1 < 0.75) {
      u1 = 4.0 * (u1 - 0.5);
      anisoFirstQuadrant(u1, u2, &phi, &cosTheta);
      phi += kPi;
    } else {
      u1 = 4.0 * (1.0 - u1);
      anisoFirstQuadrant(u1, u2, &phi, &cosTheta);
      phi = 2.0 * kPi - phi;
    }
    double sintheta = std::sqrt(dmax(0.0, 1.0 - cosTheta * cosTheta));
    Vec wh = SphericalDirection(sintheta, cosTheta, phi);
    if (!SameHemisphere(wo, wh)) wh = -wh;
    *wi = -wo + wh * 2.0 * Dot(wo, wh);
    return anisoPdfOf(wo, wh);
  }
  // fresnel_blend.dart:30-58
  Spec blendF(const Vec& wo, const Vec& wi) const {
    const Spec ONE(1.0);
    const Spec& Rd = l.R;
    const Spec& Rs = l.eta;
    Spec diffuse = Rd * ((28.0 / (23.0 * kPi))) * (ONE - Rs) *
                   ((1.0 - std::pow(1.0 - 0.5 * AbsCosTheta(wi), 5)) * (1.0 - std::pow(1.0 - 0.5 * AbsCosTheta(wo), 5)));
    Vec wh = wi + wo;
    if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spec(0.0);
    wh = Normalize(wh);
    double a = anisoD(wh) / (4.0 * AbsDot(wi, wh) * dmax(AbsCosTheta(wi), AbsCosTheta(wo)));
    Spec b = Rs + (ONE - Rs) * (std::pow(1.0 - Dot(wi, wh), 5.0));
    return diffuse + b * a;
  }
  double blendPdf(const Vec& wo, const Vec& wi) const {  // :84-90
    if (!SameHemisphere(wo, wi)) return 0.0;
    return 0.5 * (AbsCosTheta(wi) * INV_PI + anisoPdf(wo, wi));
  }

  // regular_halfangle_brdf.dart:27-75
  Spec regularHalfangleF(const Vec& WO, const Vec& WI) const {
    const MeasuredTable& tb = *l.measured;
    Vec wo = WO, wi = WI;
    Vec wh = wo + wi;
    if (wh.z < 0.0f) { wo = -wo; wi = -wi; wh = -wh; }
    if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spec(0.0);
    wh = Normalize(wh);
    double whTheta = std::acos(clampd((double)wh.z, -1.0, 1.0));  // SphericalTheta
    double whCosPhi = CosPhi(wh), whSinPhi = SinPhi(wh), whCosTheta = CosTheta(wh), whSinTheta = SinTheta(wh);
    Vec whx(whCosPhi * whCosTheta, whSinPhi * whCosTheta, -whSinTheta);
    Vec why(-whSinPhi, whCosPhi, 0.0);
    Vec wd(Dot(wi, whx), Dot(wi, why), Dot(wi, wh));
    double wdTheta = std::acos(clampd((double)wd.z, -1.0, 1.0));
    double wdPhi = std::atan2((double)wd.y, (double)wd.x);  // SphericalPhi, vector.dart:189-192
    if (wdPhi < 0.0) wdPhi += 2.0 * kPi;
    if (wdPhi > kPi) wdPhi -= kPi;
    // REMAP(V, MAX, COUNT) => ((V / MAX).toInt() * COUNT).clamp(0, COUNT - 1) AS WRITTEN: the truncation comes BEFORE the multiplication,
    // so every index is 0 except where V / MAX reaches 1 (toInt of a NaN throws in Dart: not a value the BRDF is evaluated at)
    auto REMAP = [](double V, double MAX, int COUNT) {
      long long q = (long long)std::trunc(V / MAX) * COUNT;
      return (int)std::min<long long>(std::max<long long>(q, 0), COUNT - 1);
    };
    int whThetaIndex = REMAP(std::sqrt(std::fmax(0.0, whTheta / (kPi / 2.0))), 1.0, tb.dims[0]);
    int wdThetaIndex = REMAP(wdTheta, kPi / 2.0, tb.dims[1]);
    int wdPhiIndex = REMAP(wdPhi, kPi, tb.dims[2]);
    size_t index = (size_t)wdPhiIndex + (size_t)tb.dims[2] * ((size_t)wdThetaIndex + (size_t)whThetaIndex * tb.dims[1]);
    return Spec(tb.data[3 * index], tb.data[3 * index + 1], tb.data[3 * index + 2]);
  }
  // brdf_remap.dart:23-47
  static Vec BRDFRemap(const Vec& wo, const Vec& wi) {
    double cosi = CosTheta(wi), coso = CosTheta(wo), sini = SinTheta(wi), sino = SinTheta(wo);
    auto sphPhi = [](const Vec& v) { double p = std::atan2((double)v.y, (double)v.x); return p < 0.0 ? p + 2.0 * kPi : p; };
    double dphi = sphPhi(wi) - sphPhi(wo);
    if (dphi < 0.0) dphi += 2.0 * kPi;
    if (dphi > 2.0 * kPi) dphi -= 2.0 * kPi;
    if (dphi > kPi) dphi = 2.0 * kPi - dphi;
    return Vec(sini * sino, dphi / kPi, cosi * coso);
  }
  // irregular_isotropic_brdf.dart:36-62.  KdTree.lookup (kdtree.dart:86-112) hands proc() exactly the samples with
  // DistanceSquared(sample.p, m) < maxDist2; in which order depends on nth_element and on object hash codes (kdtree.dart:120-124), so
  // the reference itself does not fix the order of the float32 sums: here the samples are visited in file order.
  Spec irregularIsotropicF(const Vec& wo, const Vec& wi) const {
    const MeasuredTable& tb = *l.measured;
    Vec m = BRDFRemap(wo, wi);
    double lastMaxDist2 = 0.001;
    for (;;) {
      Spec v(0.0);
      double sumWeights = 0.0;
      int nFound = 0;
      for (int i = 0; i < tb.dims[0]; ++i) {
        const float* q = &tb.data[6 * (size_t)i];
        Vec sp;
        sp.x = q[0]; sp.y = q[1]; sp.z = q[2];
        double d2 = DistanceSquared(sp, m);
        if (d2 < lastMaxDist2) {
          double weight = std::exp(-100.0 * d2);
          Spec sv;
          sv.c[0] = q[3]; sv.c[1] = q[4]; sv.c[2] = q[5];
          v = v + sv * weight;
          sumWeights += weight;
          ++nFound;
        }
      }
      if (nFound > 2 || lastMaxDist2 > 1.5) return Spec(clampd(v.c[0], 0.0, kInf), clampd(v.c[1], 0.0, kInf), clampd(v.c[2], 0.0, kInf)) / sumWeights;
      lastMaxDist2 *= 2.0;
    }
  }

  Spec baseF(const Vec& wo, const Vec& wi) const {
    switch (l.kind) {
      case 0: return l.R * INV_PI;  // lambertian.dart:35-37
      case 1: {                     // oren_nayar.dart:33-58
        double sinthetai = SinTheta(wi), sinthetao = SinTheta(wo);
        double maxcos = 0.0;
        if (sinthetai > 1e-4 && sinthetao > 1e-4) {
          double dcos = CosPhi(wi) * CosPhi(wo) + SinPhi(wi) * SinPhi(wo);
          maxcos = std::fmax(0.0, dcos);
        }
        double sinalpha, tanbeta;
        if (AbsCosTheta(wi) > AbsCosTheta(wo)) { sinalpha = sinthetao; tanbeta = sinthetai / AbsCosTheta(wi); }
        else { sinalpha = sinthetai; tanbeta = sinthetao / AbsCosTheta(wo); }
        return l.R * (INV_PI * (A + B * maxcos * sinalpha * tanbeta));
      }
      case 2: {  // microfacet.dart:28-57
        double cosThetaO = AbsCosTheta(wo), cosThetaI = AbsCosTheta(wi);
        if (cosThetaI == 0.0 || cosThetaO == 0.0) return Spec(0.0);
        Vec wh = wi + wo;
        if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return Spec(0.0);
        wh = Normalize(wh);
        double cosThetaH = Dot(wi, wh);
        Spec F = fresnel(cosThetaH);
        double NdotWh = AbsCosTheta(wh), NdotWo = AbsCosTheta(wo), NdotWi = AbsCosTheta(wi), WOdotWh = AbsDot(wo, wh);
        double G = dmin(1.0, dmin((2.0 * NdotWh * NdotWo / WOdotWh), (2.0 * NdotWh * NdotWi / WOdotWh)));
        return l.R * (blinnD(wh) * G) * F / (4.0 * cosThetaI * cosThetaO);
      }
      case 5: return blendF(wo, wi);
      case 6: return regularHalfangleF(wo, wi);
      case 7: return irregularIsotropicF(wo, wi);
      default: return Spec(0.0);  // specular_reflection.dart:30-32, specular_transmission.dart:33-35
    }
  }
  double basePdf(const Vec& wo, const Vec& wi) const {
    switch (l.kind) {
      case 0:
      case 1:
      case 6:
      case 7: return SameHemisphere(wo, wi) ? AbsCosTheta(wi) * INV_PI : 0.0;  // bxdf.dart:84-88
      case 2: return SameHemisphere(wo, wi) ? blinnPdf(wo, wi) : 0.0;          // microfacet.dart:68-73
      case 5: return blendPdf(wo, wi);
      default: return 0.0;
    }
  }
  // pdf is untouched when a BxDF returns without setting it (specular_transmission.dart:52-54)
  Spec baseSampleF(const Vec& wo, Vec* wi, double u1, double u2, double* pdfOut) const {
    switch (l.kind) {
      case 0:
      case 1:
      case 6:
      case 7: {  // bxdf.dart:37-48
        *wi = CosineSampleHemisphere(u1, u2);
        if (wo.z < 0.0f) wi->z = f32((double)wi->z * -1.0);
        *pdfOut = basePdf(wo, *wi);
        return baseF(wo, *wi);
      }
      case 2: {  // microfacet.dart:59-66
        *pdfOut = blinnSample(wo, wi, u1, u2);
        if (!SameHemisphere(wo, *wi)) return Spec(0.0);
        return baseF(wo, *wi);
      }
      case 5: {  // fresnel_blend.dart:60-82
        if (u1 < 0.5) {
          u1 = 2.0 * u1;
          *wi = CosineSampleHemisphere(u1, u2);
          if (wo.z < 0.0f) wi->z = f32((double)wi->z * -1.0);
        } else {
          u1 = 2.0 * (u1 - 0.5);
          *pdfOut = anisoSample(wo, wi, u1, u2);
          if (!SameHemisphere(wo, *wi)) return Spec(0.0);
        }
        *pdfOut = blendPdf(wo, *wi);
        return blendF(wo, *wi);
      }
      case 3: {  // specular_reflection.dart:34-41
        *wi = Vec(-(double)wo.x, -(double)wo.y, wo.z);
        *pdfOut = 1.0;
        return (fresnel(CosTheta(wo)) * l.R) / AbsCosTheta(*wi);
      }
      default: {  // specular_transmission.dart:37-66
        bool entering = CosTheta(wo) > 0.0;
        double ei = l.ei, et = l.et;
        if (!entering) std::swap(ei, et);
        double sini2 = SinTheta2(wo);
        double eta = ei / et;
        double sint2 = eta * eta * sini2;
        if (sint2 >= 1.0) return Spec(0.0);
        double cost = std::sqrt(dmax(0.0, 1.0 - sint2));
        if (entering) cost = -cost;
        double sintOverSini = eta;
        *wi = Vec(sintOverSini * -(double)wo.x, sintOverSini * -(double)wo.y, cost);
        *pdfOut = 1.0;
        Spec F = dielectric(CosTheta(wo), l.ei, l.et);
        return ((Spec(1.0) - F) * l.R) / AbsCosTheta(*wi);
      }
    }
  }

  // The BxDF the BSDF holds: the lobe itself, BRDFToBTDF(lobe) (brdf_to_btdf.dart:31-58: the other hemisphere of wi) and / or
  // ScaledBxDF(.., s) (scaled_bxdf.dart:24-52: s * f; it does NOT override pdf, so BxDF.pdf's cosine density answers, bxdf.dart:84-88)
  static Vec OtherHemisphere(const Vec& w) { return Vec(w.x, w.y, -(double)w.z); }
  Spec f(const Vec& wo, const Vec& wi) const {
    Spec r = baseF(wo, (l.wrap & 1) ? OtherHemisphere(wi) : wi);
    return (l.wrap & 2) ? l.scale * r : r;
  }
  double pdf(const Vec& wo, const Vec& wi) const {
    if (l.wrap & 2) return SameHemisphere(wo, wi) ? AbsCosTheta(wi) * INV_PI : 0.0;
    return basePdf(wo, (l.wrap & 1) ? OtherHemisphere(wi) : wi);
  }
  Spec sample_f(const Vec& wo, Vec* wi, double u1, double u2, double* pdfOut) const {
    Spec r = baseSampleF(wo, wi, u1, u2, pdfOut);
    if (l.wrap & 1) *wi = OtherHemisphere(*wi);
    return (l.wrap & 2) ? l.scale * r : r;
  }
};

struct Bsdf {  // bsdf.dart:41-255
  Vec p, nn, ng, sn, tn;
  DG dgs;            // bsdf.dgShading
  double eta = 1.0;  // bsdf.eta (glass: index, translucent: 1.5)
  Vec dpdx, dpdy;    // isect.dg.dpdx / dpdy after Intersection.getBSDF's computeDifferentials
  bool framed = false;
  Lobe lobes[8];     // what the BxDFs were built from (MixMaterial re-wraps them)
  int nBxDFs = 0;
  LobeEval bxdfs[8];
  Vec worldToLocal(const Vec& v) const { return Vec(Dot(v, sn), Dot(v, tn), Dot(v, nn)); }
  Vec localToWorld(const Vec& v) const {
    return Vec((double)sn.x * v.x + (double)tn.x * v.y + (double)nn.x * v.z,
               (double)sn.y * v.x + (double)tn.y * v.y + (double)nn.y * v.z,
               (double)sn.z * v.x + (double)tn.z * v.y + (double)nn.z * v.z);
  }
  int numComponents(int flags) const {
    int n = 0;
    for (int i = 0; i < nBxDFs; ++i) n += bxdfs[i].matches(flags) ? 1 : 0;
    return n;
  }
  Spec f(const Vec& woW, const Vec& wiW, int flags) const {  // bsdf.dart:177-198
    Vec wi = worldToLocal(wiW), wo = worldToLocal(woW);
    if (Dot(wiW, ng) * Dot(woW, ng) > 0) flags &= ~BSDF_TRANSMISSION;
    else flags &= ~BSDF_REFLECTION;
    Spec r(0.0);
    for (int i = 0; i < nBxDFs; ++i)
      if (bxdfs[i].matches(flags)) r = r + bxdfs[i].f(wo, wi);
    return r;
  }
  double pdf(const Vec& woW, const Vec& wiW, int flags) const {  // bsdf.dart:128-146
    if (nBxDFs == 0) return 0.0;
    Vec wo = worldToLocal(woW), wi = worldToLocal(wiW);
    double p = 0.0;
    int matching = 0;
    for (int i = 0; i < nBxDFs; ++i)
      if (bxdfs[i].matches(flags)) { ++matching; p += bxdfs[i].pdf(wo, wi); }
    return matching > 0 ? p / matching : 0.0;
  }
  Spec sample_f(const Vec& woW, Vec* wiW, const U3& s, double* pdfOut, int flags, int* sampledType) const {  // :53-126
    int matching = numComponents(flags);
    if (matching == 0) { *pdfOut = 0.0; if (sampledType) *sampledType = 0; return Spec(0.0); }
    int which = std::min((int)std::floor(s.comp * matching), matching - 1);
    const LobeEval* bxdf = nullptr;
    int count = which;
    for (int i = 0; i < nBxDFs; ++i)
      if (bxdfs[i].matches(flags) && count-- == 0) { bxdf = &bxdfs[i]; break; }
    Vec wo = worldToLocal(woW);
    Vec wi;
    *pdfOut = 0.0;
    Spec f = bxdf->sample_f(wo, &wi, s.u0, s.u1, pdfOut);
    if (*pdfOut == 0.0) { if (sampledType) *sampledType = 0; return Spec(0.0); }
    if (sampledType) *sampledType = bxdf->type;
    *wiW = localToWorld(wi);
    if (!((bxdf->type & BSDF_SPECULAR) != 0) && matching > 1)
      for (int i = 0; i < nBxDFs; ++i)
        if (&bxdfs[i] != bxdf && bxdfs[i].matches(flags)) *pdfOut += bxdfs[i].pdf(wo, wi);
    if (matching > 1) *pdfOut /= matching;
    if ((bxdf->type & BSDF_SPECULAR) == 0) {
      f = Spec(0.0);
      if (Dot(*wiW, ng) * Dot(woW, ng) > 0) flags &= ~BSDF_TRANSMISSION;
      else flags &= ~BSDF_REFLECTION;
      for (int i = 0; i < nBxDFs; ++i)
        if (bxdfs[i].matches(flags)) f = f + bxdfs[i].f(wo, wi);
    }
    return f;
  }
};

struct Ctx {
  RenderScene& rs;
  const Scene& g;
  SampleLayout layout;
  // path integrator offsets (path_integrator.dart:124-131)
  SampleOffsets lightOff[3], bsdfOff[3], pathOff[3];
  int lightNumOff[3] = {-1, -1, -1};
  // direct lighting offsets (direct_lighting_integrator.dart:70-96)
  std::vector<SampleOffsets> dlLight, dlBsdf;
  int dlLightNum = -1;
  RenderStats stats;

  explicit Ctx(RenderScene& r) : rs(r), g(*r.geom) {}

  SampleOffsets lightOffsets(int n) { SampleOffsets o; o.nSamples = n; o.componentOffset = layout.add1D(n); o.posOffset = layout.add2D(n); return o; }
  SampleOffsets bsdfOffsets(int n) { return lightOffsets(n); }  // componentOffset + dirOffset, same order

  void requestSamples() {
    const IntegratorCfg& ic = rs.integ;
    if (ic.kind == 0) {
      for (int i = 0; i < 3; ++i) {
        lightOff[i] = lightOffsets(1);
        lightNumOff[i] = layout.add1D(1);
        bsdfOff[i] = bsdfOffsets(1);
        pathOff[i] = bsdfOffsets(1);
      }
    } else if (ic.kind == 2) {
      if (ic.strategy == 0) {
        for (size_t i = 0; i < rs.lights.size(); ++i) {
          int n = rs.lights[i].nSamples;
          if (rs.sampler.kind == 0 || rs.sampler.kind >= 4) n = RoundUpPow2(n);  // sampler.roundSize (LD, adaptive, bestcandidate)
          dlLight.push_back(lightOffsets(n));
          dlBsdf.push_back(bsdfOffsets(n));
        }
        dlLightNum = -1;
      } else {
        dlLight.push_back(lightOffsets(1));
        dlLightNum = layout.add1D(1);
        dlBsdf.push_back(bsdfOffsets(1));
      }
    }
    // volume integrator (emission by default, render_options.dart:24-39): emission_integrator.dart:26-29,
    // single_scatter_integrator.dart:47-50
    tauSampleOffset = layout.add1D(1);
    scatterSampleOffset = layout.add1D(1);
  }

  // ---- participating media (lib/core/volume/*.dart, lib/volume_regions/*.dart, lib/volume_integrators/*.dart) ----------------
  int tauSampleOffset = -1, scatterSampleOffset = -1;
  Rng* trRng = nullptr;   // transmittance() draws; set per camera sample (keyed: kStreamTransmittance; serial: the one RNG)
  Rng* volRng = nullptr;  // volume Li draws (keyed: kStreamVolumeLi)
  bool hasVolume() const { return !rs.volume.regions.empty(); }

  static bool boxIntersectP(const Vec& pMin, const Vec& pMax, const Ray& ray, double* hitt0, double* hitt1) {  // bbox.dart:81-114
    double t0 = ray.mint, t1 = ray.maxt;
    const float o[3] = {ray.o.x, ray.o.y, ray.o.z}, d[3] = {ray.d.x, ray.d.y, ray.d.z};
    const float lo[3] = {pMin.x, pMin.y, pMin.z}, hi[3] = {pMax.x, pMax.y, pMax.z};
    for (int i = 0; i < 3; ++i) {
      double invRayDir = 1.0 / (double)d[i];
      double tNear = ((double)lo[i] - (double)o[i]) * invRayDir;
      double tFar = ((double)hi[i] - (double)o[i]) * invRayDir;
      if (tNear > tFar) std::swap(tNear, tFar);
      t0 = tNear > t0 ? tNear : t0;
      t1 = tFar < t1 ? tFar : t1;
      if (t0 > t1) return false;
    }
    *hitt0 = t0;
    *hitt1 = t1;
    return true;
  }
  static bool boxInside(const Vec& pMin, const Vec& pMax, const Vec& pt) {  // bbox.dart:123-127
    return pt.x >= pMin.x && pt.x <= pMax.x && pt.y >= pMin.y && pt.y <= pMax.y && pt.z >= pMin.z && pt.z <= pMax.z;
  }
  static void extentOf(const VolumeRegionCfg& v, Vec* pMin, Vec* pMax) {
    BBox b(v.p0, v.p1);
    *pMin = b.pMin;
    *pMax = b.pMax;
  }
  bool regionIntersectP(const VolumeRegionCfg& v, const Ray& r, double* t0, double* t1) const {
    Ray ray = v.worldToVolume.ray(r);  // homogenous_volume_region.dart:32-35
    Vec lo, hi;
    extentOf(v, &lo, &hi);
    return boxIntersectP(lo, hi, ray, t0, t1);
  }
  double regionDensity(const VolumeRegionCfg& v, const Vec& Pobj) const {
    Vec lo, hi;
    extentOf(v, &lo, &hi);
    if (!boxInside(lo, hi, Pobj)) return 0.0;
    if (v.kind == 0) return 1.0;
    if (v.kind == 1) {  // exponential_density_region.dart:42-50
      double height = Dot(Pobj - lo, v.upDir);
      return v.a * std::exp(-v.b * height);
    }
    // volume_grid.dart:39-66
    Vec vox(((double)Pobj.x - lo.x) / ((double)hi.x - lo.x), ((double)Pobj.y - lo.y) / ((double)hi.y - lo.y),
            ((double)Pobj.z - lo.z) / ((double)hi.z - lo.z));  // bbox.dart:193-197, a float32 Vector
    vox.x = f32((double)vox.x * v.nx - 0.5);
    vox.y = f32((double)vox.y * v.ny - 0.5);
    vox.z = f32((double)vox.z * v.nz - 0.5);
    int vx = (int)std::floor((double)vox.x), vy = (int)std::floor((double)vox.y), vz = (int)std::floor((double)vox.z);
    double dx = (double)vox.x - vx, dy = (double)vox.y - vy, dz = (double)vox.z - vz;
    auto D = [&](int x, int y, int z) {
      x = std::min(std::max(x, 0), v.nx - 1);
      y = std::min(std::max(y, 0), v.ny - 1);
      z = std::min(std::max(z, 0), v.nz - 1);
      return v.density[(size_t)z * v.nx * v.ny + (size_t)y * v.nx + x];
    };
    double d00 = Lerp(dx, D(vx, vy, vz), D(vx + 1, vy, vz));
    double d10 = Lerp(dx, D(vx, vy + 1, vz), D(vx + 1, vy + 1, vz));
    double d01 = Lerp(dx, D(vx, vy, vz + 1), D(vx + 1, vy, vz + 1));
    double d11 = Lerp(dx, D(vx, vy + 1, vz + 1), D(vx + 1, vy + 1, vz + 1));
    double d0 = Lerp(dy, d00, d10), d1 = Lerp(dy, d01, d11);
    return Lerp(dz, d0, d1);
  }
  // sigma_a / sigma_s / sigma_t / Lve at a world point: homogenous_volume_region.dart:37-56, density_region.dart:33-47
  enum { kSigA, kSigS, kSigT, kLve };
  Spec regionCoeff(const VolumeRegionCfg& v, const Vec& p, int which) const {
    const Spec c = which == kSigA ? v.sigA : which == kSigS ? v.sigS : which == kSigT ? (v.sigA + v.sigS) : v.le;
    const Vec q = v.worldToVolume.point(p);
    if (v.kind == 0) {
      Vec lo, hi;
      extentOf(v, &lo, &hi);
      return boxInside(lo, hi, q) ? c : Spec(0.0);
    }
    return c * regionDensity(v, q);
  }
  double regionPhase(const VolumeRegionCfg& v, const Vec& p, const Vec& w, const Vec& wp) const {
    if (v.kind == 0) {  // homogenous_volume_region.dart:58-63
      Vec lo, hi;
      extentOf(v, &lo, &hi);
      if (!boxInside(lo, hi, v.worldToVolume.point(p))) return 0.0;
    }
    double costheta = Dot(w, wp);  // PhaseHG, volume.dart:84-88
    return 1.0 / (4.0 * kPi) * (1.0 - v.g * v.g) / std::pow(1.0 + v.g * v.g - 2.0 * v.g * costheta, 1.5);
  }
  Spec regionTau(const VolumeRegionCfg& v, const Ray& r, double stepSize, double u) const {
    double t0 = 0.0, t1 = 0.0;
    if (v.kind == 0) {  // homogenous_volume_region.dart:65-73: analytic, step and offset unused
      if (!regionIntersectP(v, r, &t0, &t1)) return Spec(0.0);
      return (v.sigA + v.sigS) * Distance(r.at(t0), r.at(t1));
    }
    // density_region.dart:53-77
    double length = Length(r.d);
    if (length == 0.0) return Spec(0.0);
    Ray rn(r.o, r.d / length, r.mint * length, r.maxt * length, r.time);
    if (!regionIntersectP(v, rn, &t0, &t1)) return Spec(0.0);
    Spec tau(0.0);
    t0 += u * stepSize;
    while (t0 < t1) {
      tau = tau + regionCoeff(v, rn.at(t0), kSigT);
      t0 += stepSize;
    }
    return tau * stepSize;
  }
  // the scene's volumeRegion: the region itself, or an AggregateVolume over all of them (aggregate_volume.dart:23-103)
  bool volIntersectP(const Ray& ray, double* t0, double* t1) const {
    const auto& R = rs.volume.regions;
    if (R.size() == 1) return regionIntersectP(R[0], ray, t0, t1);
    *t0 = kInf;
    *t1 = -kInf;
    for (const VolumeRegionCfg& v : R) {
      double a = 0.0, b = 0.0;
      if (regionIntersectP(v, ray, &a, &b)) { *t0 = dmin(*t0, a); *t1 = dmax(*t1, b); }
    }
    return *t0 < *t1;
  }
  Spec volCoeff(const Vec& p, int which) const {
    const auto& R = rs.volume.regions;
    if (R.size() == 1) return regionCoeff(R[0], p, which);
    Spec s(0.0);
    for (const VolumeRegionCfg& v : R) s = s + regionCoeff(v, p, which);
    return s;
  }
  double volPhase(const Vec& p, const Vec& w, const Vec& wp) const {
    const auto& R = rs.volume.regions;
    if (R.size() == 1) return regionPhase(R[0], p, w, wp);
    double ph = 0.0, sumWt = 0.0;  // aggregate_volume.dart:71-80
    for (const VolumeRegionCfg& v : R) {
      double wt = regionCoeff(v, p, kSigS).luminance();
      sumWt += wt;
      ph += wt * regionPhase(v, p, w, wp);
    }
    return ph / sumWt;
  }
  Spec volTau(const Ray& ray, double step, double offset) const {
    const auto& R = rs.volume.regions;
    if (R.size() == 1) return regionTau(R[0], ray, step, offset);
    Spec t(0.0);
    for (const VolumeRegionCfg& v : R) t = t + regionTau(v, ray, step, offset);
    return t;
  }
  static Spec expNeg(const Spec& tau) { return Spec(std::exp(-(double)tau.c[0]), std::exp(-(double)tau.c[1]), std::exp(-(double)tau.c[2])); }

  // EmissionIntegrator.transmittance == SingleScatteringIntegrator.transmittance (emission_integrator.dart:85-105,
  // single_scatter_integrator.dart:26-45): with a Sample the step is stepSize and the offset the tau sample; without, 4 x stepSize
  // and a draw.  No volume region: 1, and NO draw.
  // `rng`: where the sample-less call draws its offset — the transmittance stream for the surface integrators' calls, the volume
  // Li stream for the calls SingleScatteringIntegrator.Li makes itself (keyed mode; serial mode has one RNG for everything).
  Spec transmittance(const Ray& ray, const SampleVals* sample, Rng* rng = nullptr) {
    if (!hasVolume()) return Spec(1.0);
    double step, offset;
    if (sample) {
      step = rs.volume.stepSize;
      offset = sample->oneD[tauSampleOffset][0];
    } else {
      step = 4.0 * rs.volume.stepSize;
      offset = (rng ? rng : trRng)->randomFloat();
    }
    return expNeg(volTau(ray, step, offset));
  }

  // EmissionIntegrator.Li (emission_integrator.dart:31-83) / SingleScatteringIntegrator.Li (single_scatter_integrator.dart:52-133)
  Spec volumeLi(const Ray& ray, const SampleVals& sample, Spec* T) {
    double t0 = 0.0, t1 = 0.0;
    if (!hasVolume() || !volIntersectP(ray, &t0, &t1) || (t1 - t0) == 0.0) {
      *T = Spec(1.0);
      return Spec(0.0);
    }
    Rng& rng = *volRng;
    const double stepSize = rs.volume.stepSize;
    const bool single = rs.volume.integrator == 1;
    Spec Lv(0.0);
    int nSamples = (int)std::ceil((t1 - t0) / stepSize);
    double step = (t1 - t0) / nSamples;
    Spec Tr(1.0);
    Vec p = ray.at(t0), pPrev;
    Vec w = -ray.d;
    t0 += (double)sample.oneD[scatterSampleOffset][0] * step;
    std::vector<float> lightNum, lightComp, lightPos;
    if (single) {
      lightNum.assign(nSamples, 0.f);
      LDShuffleScrambled1D(1, nSamples, lightNum.data(), rng);
      lightComp.assign(nSamples, 0.f);
      LDShuffleScrambled1D(1, nSamples, lightComp.data(), rng);
      lightPos.assign(2 * (size_t)nSamples, 0.f);
      LDShuffleScrambled2D(1, nSamples, lightPos.data(), rng);
    }
    int sampOffset = 0;
    for (int i = 0; i < nSamples; ++i, t0 += step) {
      pPrev = p;
      p = ray.at(t0);
      Ray tauRay(pPrev, p - pPrev, 0.0, 1.0, ray.time, ray.depth);
      Spec stepTau = volTau(tauRay, 0.5 * stepSize, rng.randomFloat());
      Tr = Tr * expNeg(stepTau);
      if (Tr.luminance() < 1.0e-3) {  // possibly terminate the march
        const double continueProb = 0.5;
        if (rng.randomFloat() > continueProb) {
          Tr = Spec(0.0);
          break;
        }
        Tr = Tr / continueProb;
      }
      Lv = Lv + Tr * volCoeff(p, kLve);
      if (single) {
        Spec ss = volCoeff(p, kSigS);
        if (!ss.isBlack() && !rs.lights.empty()) {
          int nLights = (int)rs.lights.size();
          int ln = std::min((int)std::floor((double)lightNum[sampOffset] * nLights), nLights - 1);
          const Light& light = rs.lights[ln];
          double pdf = 0.0;
          Vis vis;
          Vec wo;
          U3 ls;  // LightSample(lightComp, lightPos0, lightPos1): light_sample.dart:28-36
          ls.comp = lightComp[sampOffset];
          ls.u0 = lightPos[2 * (size_t)sampOffset];
          ls.u1 = lightPos[2 * (size_t)sampOffset + 1];
          Spec L = sampleLAtPoint(light, p, 0.0, ls, ray.time, &wo, &pdf, &vis);
          if (!L.isBlack() && pdf > 0.0 && !intersectP(vis.r)) {
            Spec Ld = L * transmittance(vis.r, nullptr, volRng);
            Lv = Lv + Tr * ss * volPhase(p, w, -wo) * Ld * (double)nLights / pdf;
          }
        }
        ++sampOffset;
      }
    }
    *T = Tr;
    return Lv * step;
  }

  // ---- scene queries with stats ----
  bool intersect(Ray& ray, Isect* is) {
    Hit h;
    Counters c;
    bool hit = g.intersect(ray, &h, &c);
    stats.closestRays++;
    stats.nodesVisited += c.nodes_visited;
    stats.primsTested += c.prims_tested;
    if (!hit) return false;
    fillIsect(ray, h, is);
    return true;
  }
  bool intersectP(const Ray& ray) {
    Counters c;
    bool hit = g.intersectP(ray, &c);
    stats.shadowRays++;
    stats.nodesVisited += c.nodes_visited;
    stats.primsTested += c.prims_tested;
    return hit;
  }

  // dndu / dndv from the fundamental forms (sphere.dart:138-153 and the same block in the other quadrics).  `perFactor`: the
  // sphere multiplies dpdu by (f F - e G) and then by invEGF2 (two float32 Vectors), the others by their product.
  static void weingarten(const Vec& dpdu, const Vec& dpdv, const Vec& d2Pduu, const Vec& d2Pduv, const Vec& d2Pdvv, bool perFactor,
                         Vec* dndu, Vec* dndv) {
    double E = Dot(dpdu, dpdu), F = Dot(dpdu, dpdv), G = Dot(dpdv, dpdv);
    Vec N = Normalize(Cross(dpdu, dpdv));
    double e = Dot(N, d2Pduu), f = Dot(N, d2Pduv), gg = Dot(N, d2Pdvv);
    double invEGF2 = 1.0 / (E * G - F * F);
    if (perFactor) {
      *dndu = dpdu * (f * F - e * G) * invEGF2 + dpdv * (e * F - f * E) * invEGF2;
      *dndv = dpdu * (gg * F - f * G) * invEGF2 + dpdv * (f * F - gg * E) * invEGF2;
    } else {
      *dndu = dpdu * ((f * F - e * G) * invEGF2) + dpdv * ((e * F - f * E) * invEGF2);
      *dndv = dpdu * ((gg * F - f * G) * invEGF2) + dpdv * ((f * F - gg * E) * invEGF2);
    }
  }

  // dg.set(...) of triangle.dart:100-154 / sphere.dart:118-160 + differential_geometry.dart:77-99.
  // `rayAtHit` is the ray the shape test ran on (its maxt already holds tHit for closest-hit queries).
  void shapeDG(uint32_t prim, const Ray& ray, const Hit& h, DG* dg) const {
    if (prim < g.ntris()) {
      Vec p1, p2, p3;
      g.triVerts(prim, &p1, &p2, &p3);
      double uvs[6];  // triangle.dart:105 getUVs: the mesh's uvs or (0,0),(1,0),(1,1) (:246-262)
      g.triUVs(prim, uvs);
      {  // triangle.dart:133-136
        double b0 = 1.0 - h.b1 - h.b2;
        dg->u = b0 * uvs[0] + h.b1 * uvs[2] + h.b2 * uvs[4];
        dg->v = b0 * uvs[1] + h.b1 * uvs[3] + h.b2 * uvs[5];
      }
      double du1 = uvs[0] - uvs[4], du2 = uvs[2] - uvs[4], dv1 = uvs[1] - uvs[5], dv2 = uvs[3] - uvs[5];
      Vec dp1 = p1 - p3, dp2 = p2 - p3;
      double determinant = du1 * dv2 - dv1 * du2;
      Vec dpdu, dpdv;
      if (determinant == 0.0) {
        double e1x = (double)p2.x - p1.x, e1y = (double)p2.y - p1.y, e1z = (double)p2.z - p1.z;
        double e2x = (double)p3.x - p1.x, e2y = (double)p3.y - p1.y, e2z = (double)p3.z - p1.z;
        double e3x = (e2y * e1z) - (e2z * e1y), e3y = (e2z * e1x) - (e2x * e1z), e3z = (e2x * e1y) - (e2y * e1x);
        double len = std::sqrt(e3x * e3x + e3y * e3y + e3z * e3z);
        CoordinateSystem(Vec(e3x / len, e3y / len, e3z / len), &dpdu, &dpdv);
      } else {
        double invdet = 1.0 / determinant;
        dpdu = (dp1 * dv2 - dp2 * dv1) * invdet;
        dpdv = (dp1 * -du2 + dp2 * du1) * invdet;
      }
      dg->p = ray.at(h.t);
      dg->dpdu = dpdu;
      dg->dpdv = dpdv;
      dg->dndu = Vec();  // dg.set(..., Normal.ZERO, Normal.ZERO, ...), triangle.dart:150
      dg->dndv = Vec();
    } else if (g.spheres[prim - g.ntris()].shape >= 2) {
      const Sphere& s = g.spheres[prim - g.ntris()];
      Vec phit = h.phitObj;
      Vec dpdu(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);  // the same expression in all four files
      Vec dpdv;
      if (s.shape == 2) {  // cylinder.dart:111-112
        dpdv = Vec(0.0, 0.0, s.zmax - s.zmin);
      } else if (s.shape == 3) {  // cone.dart:101-107
        double v = (double)phit.z / s.height;
        dpdv = Vec(-(double)phit.x / (1.0 - v), -(double)phit.y / (1.0 - v), s.height);
      } else if (s.shape == 4) {  // paraboloid.dart:107-109
        dpdv = Vec((double)phit.x / (2.0 * phit.z), (double)phit.y / (2.0 * phit.z), 1.0) * (s.zmax - s.zmin);
      } else {  // hyperboloid.dart:128-133
        double cosphi = std::cos(h.phi), sinphi = std::sin(h.phi);
        dpdv = Vec(((double)s.hp2.x - s.hp1.x) * cosphi - ((double)s.hp2.y - s.hp1.y) * sinphi,
                   ((double)s.hp2.x - s.hp1.x) * sinphi + ((double)s.hp2.y - s.hp1.y) * cosphi, (double)s.hp2.z - s.hp1.z);
      }
      // second derivatives and the Weingarten equations (cylinder.dart:114-135, cone.dart:109-131, paraboloid.dart:111-137,
      // hyperboloid.dart:138-157)
      Vec d2Pduu = Vec(phit.x, phit.y, 0.0) * (-s.phiMax * s.phiMax), d2Pduv, d2Pdvv;
      if (s.shape == 3) {
        double v = (double)phit.z / s.height;
        d2Pduv = Vec(phit.y, -(double)phit.x, 0.0) * (s.phiMax / (1.0 - v));
      } else if (s.shape == 4) {
        d2Pduv = Vec(-(double)phit.y / (2.0 * phit.z), (double)phit.x / (2.0 * phit.z), 0.0) * (s.zmax - s.zmin) * s.phiMax;
        d2Pdvv = Vec((double)phit.x / (4.0 * phit.z * phit.z), (double)phit.y / (4.0 * phit.z * phit.z), 0.0) *
                 (-(s.zmax - s.zmin) * (s.zmax - s.zmin));
      } else if (s.shape == 5) {
        d2Pduv = Vec(-(double)dpdv.y, dpdv.x, 0.0) * s.phiMax;
      }
      Vec dndu, dndv;
      weingarten(dpdu, dpdv, d2Pduu, d2Pduv, d2Pdvv, false, &dndu, &dndv);
      dg->p = s.o2w.point(phit);
      dg->dpdu = s.o2w.vector(dpdu);
      dg->dpdv = s.o2w.vector(dpdv);
      dg->dndu = s.o2w.normal(dndu);
      dg->dndv = s.o2w.normal(dndv);
      dg->u = h.b1; dg->v = h.b2;
    } else if (g.spheres[prim - g.ntris()].shape == 1) {  // disk.dart:69-97
      const Sphere& s = g.spheres[prim - g.ntris()];
      Vec phit = h.phitObj;
      double dist2 = (double)phit.x * phit.x + (double)phit.y * phit.y;
      double oneMinusV = (std::sqrt(dist2) - s.innerRadius) / (s.radius - s.innerRadius);
      double invOneMinusV = (oneMinusV > 0.0) ? (1.0 / oneMinusV) : 0.0;
      Vec dpdu(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);
      Vec dpdv(-(double)phit.x * invOneMinusV, -(double)phit.y * invOneMinusV, 0.0);
      dpdu = dpdu * (s.phiMax * INV_TWOPI);
      dpdv = dpdv * ((s.radius - s.innerRadius) / s.radius);
      dg->p = s.o2w.point(phit);
      dg->dpdu = s.o2w.vector(dpdu);
      dg->dpdv = s.o2w.vector(dpdv);
      dg->dndu = s.o2w.normal(Vec());  // disk.dart:81-82
      dg->dndv = s.o2w.normal(Vec());
      dg->u = h.b1; dg->v = h.b2;
    } else {
      const Sphere& s = g.spheres[prim - g.ntris()];
      Vec phit = h.phitObj;
      double phi = h.phi;
      double theta = std::acos(clampd((double)phit.z / s.radius, -1.0, 1.0));
      double zradius = std::sqrt((double)phit.x * phit.x + (double)phit.y * phit.y);
      double invzradius = 1.0 / zradius;
      double cosphi = phit.x * invzradius, sinphi = phit.y * invzradius;
      (void)phi;
      Vec dpdu(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);
      Vec dpdv = Vec(phit.z * cosphi, phit.z * sinphi, -s.radius * std::sin(theta)) * (s.thetaMax - s.thetaMin);
      // sphere.dart:131-153 (this file multiplies the vectors factor by factor)
      Vec d2Pduu = Vec(phit.x, phit.y, 0.0) * -s.phiMax * s.phiMax;
      Vec d2Pduv = Vec(-sinphi, cosphi, 0.0) * (s.thetaMax - s.thetaMin) * phit.z * s.phiMax;
      Vec d2Pdvv = Vec(phit.x, phit.y, phit.z) * -(s.thetaMax - s.thetaMin) * (s.thetaMax - s.thetaMin);
      Vec dndu, dndv;
      weingarten(dpdu, dpdv, d2Pduu, d2Pduv, d2Pdvv, true, &dndu, &dndv);
      dg->p = s.o2w.point(phit);
      dg->dpdu = s.o2w.vector(dpdu);
      dg->dpdv = s.o2w.vector(dpdv);
      dg->dndu = s.o2w.normal(dndu);
      dg->dndv = s.o2w.normal(dndv);
      dg->u = h.b1; dg->v = h.b2;
    }
    dg->nn = Normalize(Cross(dg->dpdu, dg->dpdv));
    dg->reverse = g.reverseOf[prim] != 0;
    if (g.reverseOf[prim]) dg->nn = dg->nn * -1.0;  // transformSwapsHandedness is never set (shape.dart:30)
  }
  void fillIsect(const Ray& ray, const Hit& h, Isect* is) const {
    is->prim = h.prim;
    is->rayEpsilon = h.rayEpsilon;
    is->inst = h.inst;
    // the shape computed dg.p with the ray it was given; maxt == tHit after a closest-hit query
    if (h.inst < 0) {
      shapeDG((uint32_t)h.prim, ray, h, &is->dg);
      return;
    }
    // transformed_primitive.dart:30-58: the shape saw the ray in primitive space; its differential geometry goes back to world space
    const Transform w2p = g.instances[(size_t)h.inst].worldToPrimitive.interpolate(ray.time);
    const Ray r2(w2p.point(ray.o), w2p.vector(ray.d), ray.mint, ray.maxt, ray.time, ray.depth);
    shapeDG((uint32_t)h.prim, r2, h, &is->dg);
    is->w2p = w2p;
    if (!XfIsIdentity(w2p)) {
      const Transform p2w = XfInverse(w2p);
      DG& dg = is->dg;
      dg.p = p2w.point(dg.p);
      dg.nn = Normalize(p2w.normal(dg.nn));
      dg.dpdu = p2w.vector(dg.dpdu);
      dg.dpdv = p2w.vector(dg.dpdv);
      dg.dndu = p2w.normal(dg.dndu);
      dg.dndv = p2w.normal(dg.dndv);
    }
  }

  // Shape.intersect without the GeometricPrimitive side effect (ray.maxDistance unchanged)
  bool shapeIntersect(uint32_t prim, const Ray& ray, double* thit, DG* dg) const {
    Ray r = ray;
    Hit h;
    if (!g.primIntersect(prim, r, &h)) return false;
    *thit = h.t;
    shapeDG(prim, ray, h, dg);
    return true;
  }

  // ---- BSDF (intersection.dart:44-50 -> geometric_primitive.dart:71-75 -> the material's getBSDF) ----
  static Spec specOf(const float v[3]) { Spec r; r.c[0] = v[0]; r.c[1] = v[1]; r.c[2] = v[2]; return r; }
  static Spec clampS(const Spec& a, double lo = 0.0, double hi = kInf) {  // rgb_color.dart:189-192
    auto cl = [&](float x) { double v = x; return v < lo ? lo : (v > hi ? hi : v); };
    return Spec(cl(a.c[0]), cl(a.c[1]), cl(a.c[2]));
  }
  static double blinnExp(double rough) {  // 1 / roughness, then blinn.dart:24-28
    double e = 1.0 / rough;
    return (e > 10000.0 || std::isnan(e)) ? 10000.0 : e;
  }
  static Spec approxEta(const Spec& fr) {  // shiny_metal_material.dart:75-79
    Spec refl = clampS(fr, 0.0, 0.999);
    Spec sq(std::sqrt((double)refl.c[0]), std::sqrt((double)refl.c[1]), std::sqrt((double)refl.c[2]));
    return (Spec(1.0) + sq) / (Spec(1.0) - sq);
  }
  Spec texS(int id, const DG& dg) const { float v[3]; rs.textures.evalSpec(id, dg, v); return specOf(v); }
  double texF(int id, const DG& dg) const { return rs.textures.evalFloat(id, dg); }
  void addLobe(Bsdf* b, const Lobe& lobe) const {
    Lobe l = lobe;
    if (l.kind == 6 || l.kind == 7) l.measured = &rs.measured[(size_t)l.param];
    b->bxdfs[b->nBxDFs++].init(l);
    b->lobes[b->nBxDFs - 1] = l;
  }
  static Lobe mkLobe(int kind, const Spec& R, int fresnel = 0, double param = 0.0, double ei = 1.0, double et = 1.0) {
    Lobe l; l.kind = kind; l.R = R; l.fresnel = fresnel; l.param = param; l.ei = ei; l.et = et; return l;
  }
  // new BSDF(dgs, dgGeom.nn, eta): bsdf.dart:45-51
  static void frameBsdf(Bsdf* b, const DG& dgs, const Vec& ng, double eta) {
    b->dgs = dgs; b->eta = eta;
    b->p = dgs.p; b->ng = ng; b->nn = dgs.nn;
    b->sn = Normalize(dgs.dpdu);
    b->tn = Cross(b->nn, b->sn);
    b->nBxDFs = 0;
    b->framed = true;
  }
  // Material.getBSDF(dgGeom, dgShading) of material `mat` into *b (lib/materials/*.dart)
  void materialBSDF(uint32_t mat, const DG& dgGeom, const DG& dgShading, Bsdf* b) const {
    static const Material kDefault = Material::matte(Spec(0.5), 0.0);  // no material table: the default matte, Kd = 0.5
    const MaterialProgram* prog = (mat < rs.programs.size() && rs.programs[mat].kind >= 0) ? &rs.programs[mat] : nullptr;
    if (!prog) {  // constant parameters, no bump map: the lobe list flattened by the caller
      frameBsdf(b, dgShading, dgGeom.nn, 1.0);
      const Material& m = rs.materials.empty() ? kDefault : rs.materials[mat];
      for (const Lobe& l : m.lobes) addLobe(b, l);
      return;
    }
    const int* t = prog->tex;
    if (prog->kind == 9) {  // mix_material.dart:36-50
      materialBSDF((uint32_t)prog->m1, dgGeom, dgShading, b);
      Bsdf b2;
      materialBSDF((uint32_t)prog->m2, dgGeom, dgShading, &b2);
      Spec s1 = clampS(texS(t[0], dgShading));
      Spec s2 = clampS(Spec(1.0) - s1);
      const int n1 = b->nBxDFs, n2 = b2.nBxDFs;
      Lobe all[8];
      int n = 0;
      for (int i = 0; i < n1 && n < 8; ++i) { all[n] = b->lobes[i]; all[n].wrap |= 2; all[n].scale = s1; ++n; }
      for (int i = 0; i < n2 && n < 8; ++i) { all[n] = b2.lobes[i]; all[n].wrap |= 2; all[n].scale = s2; ++n; }
      b->nBxDFs = 0;
      for (int i = 0; i < n; ++i) addLobe(b, all[i]);
      return;
    }
    DG dgs = dgShading;
    if (prog->bump >= 0) Bump(rs.textures, prog->bump, dgGeom, dgShading, &dgs);
    switch (prog->kind) {
      case 0: {  // matte_material.dart:41-65
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        Spec r = clampS(texS(t[0], dgs));
        double sig = clampd(texF(t[1], dgs), 0.0, 90.0);
        if (!r.isBlack()) addLobe(b, sig == 0.0 ? mkLobe(0, r) : mkLobe(1, r, 0, sig));
        break;
      }
      case 1: {  // mirror_material.dart:38-55
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        Spec R = clampS(texS(t[0], dgs));
        if (!R.isBlack()) addLobe(b, mkLobe(3, R, 0));
        break;
      }
      case 2: {  // glass_material.dart:44-70
        double ior = texF(t[2], dgs);
        frameBsdf(b, dgs, dgGeom.nn, ior);
        Spec R = clampS(texS(t[0], dgs)), T = clampS(texS(t[1], dgs));
        if (!R.isBlack()) addLobe(b, mkLobe(3, R, 1, 0.0, 1.0, ior));
        if (!T.isBlack()) addLobe(b, mkLobe(4, T, 1, 0.0, 1.0, ior));
        break;
      }
      case 3: {  // plastic_material.dart:43-72
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        Spec kd = clampS(texS(t[0], dgs));
        if (!kd.isBlack()) addLobe(b, mkLobe(0, kd));
        Spec ks = clampS(texS(t[1], dgs));
        if (!ks.isBlack()) addLobe(b, mkLobe(2, ks, 1, blinnExp(texF(t[2], dgs)), 1.5, 1.0));
        break;
      }
      case 4: {  // metal_material.dart:44-64
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        double rough = texF(t[2], dgs);
        Lobe l = mkLobe(2, Spec(1.0), 2, blinnExp(rough));
        l.eta = texS(t[0], dgs);
        l.k = texS(t[1], dgs);
        addLobe(b, l);
        break;
      }
      case 5: {  // shiny_metal_material.dart:44-73
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        Spec spec = clampS(texS(t[0], dgs));
        double rough = texF(t[2], dgs);
        Spec R = clampS(texS(t[1], dgs));
        if (!spec.isBlack()) { Lobe l = mkLobe(2, Spec(1.0), 2, blinnExp(rough)); l.eta = approxEta(spec); l.k = Spec(0.0); addLobe(b, l); }
        if (!R.isBlack()) { Lobe l = mkLobe(3, Spec(1.0), 2); l.eta = approxEta(R); l.k = Spec(0.0); addLobe(b, l); }
        break;
      }
      case 6: {  // substrate_material.dart:48-70: FresnelBlend(d, s, Anisotropic(1 / u, 1 / v))
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        Spec d = clampS(texS(t[0], dgs)), sp = clampS(texS(t[1], dgs));
        double u = texF(t[2], dgs), v = texF(t[3], dgs);
        if (!d.isBlack() || !sp.isBlack()) { Lobe l = mkLobe(5, d, 0, blinnExp(u), blinnExp(v)); l.eta = sp; addLobe(b, l); }
        break;
      }
      case 7: {  // translucent_material.dart:47-103
        frameBsdf(b, dgs, dgGeom.nn, 1.5);
        Spec r = clampS(texS(t[2], dgs)), tr = clampS(texS(t[3], dgs));
        if (r.isBlack() && tr.isBlack()) break;
        Spec kd = clampS(texS(t[0], dgs));
        if (!kd.isBlack()) {
          if (!r.isBlack()) addLobe(b, mkLobe(0, r * kd));
          if (!tr.isBlack()) { Lobe l = mkLobe(0, tr * kd); l.wrap = 1; addLobe(b, l); }
        }
        Spec ks = clampS(texS(t[1], dgs));
        if (!ks.isBlack()) {
          double e = blinnExp(texF(t[4], dgs));
          if (!r.isBlack()) addLobe(b, mkLobe(2, r * ks, 1, e, 1.5, 1.0));
          if (!tr.isBlack()) { Lobe l = mkLobe(2, tr * ks, 1, e, 1.5, 1.0); l.wrap = 1; addLobe(b, l); }
        }
        break;
      }
      case 11: {  // measured_material.dart:219-238 (m1 = table index; regularHalfangleData or thetaPhiData picks the BxDF)
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        addLobe(b, mkLobe(rs.measured.at((size_t)prog->m1).kind == 0 ? 6 : 7, Spec(1.0), 0, (double)prog->m1));
        break;
      }
      case 10: {  // subsurface_material.dart:52-69, kd_subsurface_material.dart:48-67 (the BSSRDF is the dipole integrator's)
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        Spec R = clampS(texS(t[0], dgs));
        double e = texF(t[1], dgs);
        if (!R.isBlack()) addLobe(b, mkLobe(3, R, 1, 0.0, 1.0, e));
        break;
      }
      default: {  // 8 uber_material.dart:56-104
        frameBsdf(b, dgs, dgGeom.nn, 1.0);
        Spec op = clampS(texS(t[5], dgs));
        if (!(op.c[0] == 1.f && op.c[1] == 1.f && op.c[2] == 1.f)) {
          Spec neg(-(double)op.c[0], -(double)op.c[1], -(double)op.c[2]);
          addLobe(b, mkLobe(4, neg + Spec(1.0), 1, 0.0, 1.0, 1.0));
        }
        Spec kd = op * clampS(texS(t[0], dgs));
        if (!kd.isBlack()) addLobe(b, mkLobe(0, kd));
        double e = texF(t[6], dgs);
        Spec ks = op * clampS(texS(t[1], dgs));
        if (!ks.isBlack()) addLobe(b, mkLobe(2, ks, 1, blinnExp(texF(t[4], dgs)), e, 1.0));
        Spec kr = op * clampS(texS(t[2], dgs));
        if (!kr.isBlack()) addLobe(b, mkLobe(3, kr, 1, 0.0, e, 1.0));
        Spec kt = op * clampS(texS(t[3], dgs));
        if (!kt.isBlack()) addLobe(b, mkLobe(4, kt, 1, 0.0, e, 1.0));
        break;
      }
    }
  }

  // Intersection.getBSDF(ray): computeDifferentials, Shape.getShadingGeometry, Material.getBSDF.  `rd` null: a ray without
  // differentials (RayDifferential.child, ray_differential.dart:42-44)
  Bsdf getBSDF(const Isect& is, const RayDiff* rd = nullptr) const {
    Bsdf b;
    DG dg = is.dg;
    computeDifferentials(&dg, rd ? *rd : RayDiff());
    b.dpdx = dg.dpdx; b.dpdy = dg.dpdy;
    DG dgs = dg;  // dgShading == dg: no per-vertex normals (triangle.dart:273-276), quadrics (shape.dart:73-77)
    const Scene::MeshInfo* mesh = (uint32_t)is.prim < g.ntris() ? g.meshOf((uint32_t)is.prim) : nullptr;
    if (mesh && (mesh->hasN || mesh->hasS)) {  // Triangle.getShadingGeometry, triangle.dart:271-364
      const uint32_t tri = (uint32_t)is.prim;
      // obj2world: the Intersection's objectToWorld — the shape's, or through a TransformedPrimitive Inverse(worldToObject * w2p)
      // (transformed_primitive.dart:44-45); the dndu / dndv of :354-355 use the SHAPE's own objectToWorld either way
      const Transform obj2world = is.inst >= 0 && !XfIsIdentity(is.w2p) ? XfInverse(XfMul(XfInverse(mesh->o2w), is.w2p)) : mesh->o2w;
      double uv[6];
      g.triUVs(tri, uv);
      double A0 = uv[2] - uv[0], A1 = uv[4] - uv[0], A2 = uv[3] - uv[1], A3 = uv[5] - uv[1];
      double C0 = dg.u - uv[0], C1 = dg.v - uv[1];
      double bx, by, bz;
      double det = A0 * A3 - A1 * A2;  // SolveLinearSystem2x2, common.dart:170-185
      bool ok = !(std::fabs(det) < 1.0e-10);
      if (ok) {
        by = (A3 * C0 - A1 * C1) / det;
        bz = (A0 * C1 - A2 * C0) / det;
        if (std::isnan(by) || std::isnan(bz)) ok = false;
      }
      if (!ok) bx = by = bz = 1.0 / 3.0;
      else bx = 1.0 - by - bz;
      auto vert = [&](const std::vector<float>& a, int k) {
        const float* q = &a[3 * (size_t)g.idx[3 * (size_t)tri + k]];
        return Vec(q[0], q[1], q[2]);
      };
      Vec ns, ss, ts;
      if (mesh->hasN) ns = Normalize(obj2world.normal(((vert(g.vertN, 0) * bx) + (vert(g.vertN, 1) * by)) + (vert(g.vertN, 2) * bz)));
      else ns = dg.nn;
      if (mesh->hasS) ss = Normalize(obj2world.vector(((vert(g.vertS, 0) * bx) + (vert(g.vertS, 1) * by)) + (vert(g.vertS, 2) * bz)));
      else ss = Normalize(dg.dpdu);
      ts = Cross(ss, ns);
      if (LengthSquared(ts) > 0.0) {
        ts = Normalize(ts);
        ss = Cross(ts, ns);
      } else {
        CoordinateSystem(ns, &ss, &ts);
      }
      Vec dndu, dndv;  // :328-351
      if (mesh->hasN) {
        double du1 = uv[0] - uv[4], du2 = uv[2] - uv[4], dv1 = uv[1] - uv[5], dv2 = uv[3] - uv[5];
        Vec dn1 = vert(g.vertN, 0) - vert(g.vertN, 2), dn2 = vert(g.vertN, 1) - vert(g.vertN, 2);
        double determinant = du1 * dv2 - dv1 * du2;
        if (determinant != 0.0) {
          double invdet = 1.0 / determinant;
          dndu = (dn1 * dv2 - dn2 * dv1) * invdet;
          dndv = (dn1 * -du2 + dn2 * du1) * invdet;
        }
      }
      // dgShading.set(dg.p, ss, ts, o2w(dndu), o2w(dndv), dg.u, dg.v, dg.shape): nn = normalize(cross(dpdu, dpdv)), flipped by
      // reverseOrientation (differential_geometry.dart:77-99); the differentials are copied over (:354-363)
      dgs.dpdu = ss;
      dgs.dpdv = ts;
      dgs.dndu = mesh->o2w.normal(dndu);
      dgs.dndv = mesh->o2w.normal(dndv);
      dgs.nn = Normalize(Cross(ss, ts));
      if (g.reverseOf[is.prim]) dgs.nn = dgs.nn * -1.0;
    }
    materialBSDF(rs.materials.empty() ? 0u : (uint32_t)g.materialOf[is.prim], dg, dgs, &b);
    return b;
  }

  // ---- lights ----
  Spec areaL(const Light& l, const Vec& n, const Vec& w) const { return Dot(n, w) > 0.0 ? l.L : Spec(0.0); }
  Spec isectLe(const Isect& is, const Vec& wo) const {  // intersection.dart:62-65
    int li = g.lightOf[is.prim];
    return li >= 0 ? areaL(rs.lights[li], is.dg.nn, wo) : Spec(0.0);
  }
  double shapeArea(uint32_t prim) const {
    if (prim < g.ntris()) {  // triangle.dart:265-269
      Vec p1, p2, p3;
      g.triVerts(prim, &p1, &p2, &p3);
      return 0.5 * Length(Cross(p2 - p1, p3 - p1));
    }
    const Sphere& s = g.spheres[prim - g.ntris()];
    if (s.shape == 1) return s.phiMax * 0.5 * (s.radius * s.radius - s.innerRadius * s.innerRadius);  // disk.dart:142-145
    if (s.shape == 2) return (s.zmax - s.zmin) * s.phiMax * s.radius;                                // cylinder.dart:226-228
    if (s.shape == 3) return s.radius * std::sqrt((s.height * s.height) + (s.radius * s.radius)) * s.phiMax / 2.0;  // cone.dart:211-214
    if (s.shape == 4) return s.phiMax / 12.0 * (std::pow(1 + 4 * s.zmin, 1.5) - std::pow(1 + 4 * s.zmax, 1.5));  // paraboloid.dart:215-218 (as written)
    if (s.shape == 5) {  // hyperboloid.dart:246-261
      double p1x = s.hp1.x, p1y = s.hp1.y, p1z = s.hp1.z, p2x = s.hp2.x, p2y = s.hp2.y, p2z = s.hp2.z;
      auto SQR = [](double a) { return a * a; };
      auto QUAD = [](double a) { return a * a * a * a; };
      return s.phiMax / 6.0 *
             (2.0 * QUAD(p1x) - 2.0 * p1x * p1x * p1x * p2x + 2.0 * QUAD(p2x) +
              2.0 * (p1y * p1y + p1y * p2y + p2y * p2y) * (SQR(p1y - p2y) + SQR(p1z - p2z)) +
              p2x * p2x * (5.0 * p1y * p1y + 2.0 * p1y * p2y - 4.0 * p2y * p2y + 2.0 * SQR(p1z - p2z)) +
              p1x * p1x * (-4.0 * p1y * p1y + 2.0 * p1y * p2y + 5.0 * p2y * p2y + 2.0 * SQR(p1z - p2z)) -
              2.0 * p1x * p2x * (p2x * p2x - p1y * p1y + 5.0 * p1y * p2y - p2y * p2y - p1z * p1z + 2.0 * p1z * p2z - p2z * p2z));
    }
    return s.phiMax * s.radius * (s.zmax - s.zmin);  // sphere.dart:243-245
  }
  Vec sphereSample(const Sphere& s, double u1, double u2, Vec* ns) const {  // sphere.dart:247-259
    Vec p = Vec() + UniformSampleSphere(u1, u2) * s.radius;
    Vec n = s.o2w.normal(Vec(p.x, p.y, p.z));
    n = n / Length(n);
    if (s.reverseOrientation) n = Vec(-(double)n.x, -(double)n.y, -(double)n.z);
    *ns = n;
    return s.o2w.point(p);
  }
  Vec shapeSample2(uint32_t prim, const Vec& p, double u1, double u2, Vec* ns) const {
    if (prim < g.ntris()) {  // shape.dart:96-98 -> triangle.dart:366-383
      double su1 = std::sqrt(u1);
      double b1 = 1.0 - su1, b2 = u2 * su1;
      Vec t0, t1, t2;
      g.triVerts(prim, &t0, &t1, &t2);
      Vec pt = t0 * b1 + t1 * b2 + t2 * (1.0 - b1 - b2);
      Vec n = Normalize(Cross(t1 - t0, t2 - t0));
      if (g.reverseOf[prim]) n = Vec((double)n.x * -1.0, (double)n.y * -1.0, (double)n.z * -1.0);
      *ns = n;
      return pt;
    }
    const Sphere& s = g.spheres[prim - g.ntris()];
    if (s.shape == 1) {  // shape.dart:96-98 -> disk.dart:147-159
      double t0, t1;
      ConcentricSampleDisk(u1, u2, &t0, &t1);
      Vec pd(t0 * s.radius, t1 * s.radius, s.height);
      Vec n = s.o2w.normal(Vec(0.0, 0.0, 1.0));
      n = n / Length(n);
      if (s.reverseOrientation) n = n * -1.0;
      *ns = n;
      return s.o2w.point(pd);
    }
    if (s.shape == 2) {  // shape.dart:96-98 -> cylinder.dart:230-240 (the only other quadric with a sample())
      double z = s.zmin * (1.0 - u1) + s.zmax * u1;  // Lerp, common.dart:80-81
      double t = u2 * s.phiMax;
      Vec pc(s.radius * std::cos(t), s.radius * std::sin(t), z);
      Vec n = s.o2w.normal(Vec(pc.x, pc.y, 0.0));
      n = n / Length(n);
      if (s.reverseOrientation) n = n * -1.0;
      *ns = n;
      return s.o2w.point(pc);
    }
    // sphere.dart:261-297
    Vec Pcenter = s.o2w.point(Vec());
    Vec wc = Normalize(Pcenter - p);
    Vec wcX, wcY;
    CoordinateSystem(wc, &wcX, &wcY);
    if (DistanceSquared(p, Pcenter) - s.radius * s.radius < 1.0e-4) return sphereSample(s, u1, u2, ns);
    double sinThetaMax2 = s.radius * s.radius / DistanceSquared(p, Pcenter);
    double cosThetaMax = std::sqrt(std::fmax(0.0, 1.0 - sinThetaMax2));
    Ray r(p, UniformSampleCone2(u1, u2, cosThetaMax, wcX, wcY, wc), 1.0e-3);
    double thit;
    DG dgs;
    if (!shapeIntersect(prim, r, &thit, &dgs)) thit = Dot(Pcenter - p, Normalize(r.d));
    Vec ps = r.at(thit);
    Vec n = Normalize(ps - Pcenter);
    if (s.reverseOrientation) n = Vec(-(double)n.x, -(double)n.y, -(double)n.z);
    *ns = n;
    return ps;
  }
  double shapePdf2(uint32_t prim, const Vec& p, const Vec& wi) const {
    if (prim >= g.ntris() && g.spheres[prim - g.ntris()].shape == 0) {  // sphere.dart:299-311
      const Sphere& s = g.spheres[prim - g.ntris()];
      Vec Pcenter = s.o2w.point(Vec());
      if (!(DistanceSquared(p, Pcenter) - s.radius * s.radius < 1.0e-4)) {
        double sinThetaMax2 = s.radius * s.radius / DistanceSquared(p, Pcenter);
        double cosThetaMax = std::sqrt(std::fmax(0.0, 1.0 - sinThetaMax2));
        return UniformConePdf(cosThetaMax);
      }
    }
    // shape.dart:100-121
    Ray ray(p, wi, 1.0e-3);
    ray.depth = -1;
    double thit;
    DG dgl;
    if (!shapeIntersect(prim, ray, &thit, &dgl)) return 0.0;
    double pdf = DistanceSquared(p, ray.at(thit)) / (AbsDot(dgl.nn, -wi) * shapeArea(prim));
    if (std::isinf(pdf)) pdf = 0.0;
    return pdf;
  }
  Vec shapeSetSample(const Light& l, const U3& ls, Vec* Ns, const Vec& p) const {  // shape_set.dart:53-79
    int sn = l.areaDistribution.sampleDiscrete(ls.comp) % (int)l.shapes.size();
    Vec pt = shapeSample2(l.shapes[sn], p, ls.u0, ls.u1, Ns);
    Ray r(p, pt - p, 1.0e-3, kInf);
    double thit = 1.0;
    bool anyHit = false;
    DG dg;
    for (uint32_t sh : l.shapes) anyHit = shapeIntersect(sh, r, &thit, &dg) || anyHit;
    if (anyHit) *Ns = dg.nn;
    return r.at(thit);
  }
  double shapeSetPdf(const Light& l, const Vec& p, const Vec& wi) const {  // shape_set.dart:81-89
    double pdf = 0.0;
    for (size_t i = 0; i < l.shapes.size(); ++i) pdf += l.areas[i] * shapePdf2(l.shapes[i], p, wi);
    return pdf / l.area;
  }

  // Light.Le(ray): zero for every light but the infinite one (light.dart:70-72, infinite_area_light.dart:86-91)
  static double SphericalTheta(const Vec& v) { return std::acos(clampd((double)v.z, -1.0, 1.0)); }  // vector.dart:185-187
  static double SphericalPhi(const Vec& v) {                                                       // vector.dart:189-192
    double p = std::atan2((double)v.y, (double)v.x);
    return (p < 0.0) ? p + 2.0 * kPi : p;
  }
  Spec lightLe(const Light& l, const Ray& r) const {
    if (l.kind != 4) return Spec(0.0);
    Vec wh = Normalize(l.worldToLight.vector(r.d));
    double s = SphericalPhi(wh) * INV_TWOPI, t = SphericalTheta(wh) * INV_PI;
    return l.radiance(s, t, 0.0);
  }
  Spec allLightsLe(const Ray& r) const {  // sampler_renderer.dart:88-92, path_integrator.dart:108-112
    Spec L(0.0);
    for (const Light& l : rs.lights) L = L + lightLe(l, r);
    return L;
  }
  double lightPdf(const Light& l, const Vec& p, const Vec& w) const {
    if (l.kind == 4) {  // infinite_area_light.dart:244-259
      Vec wi = l.worldToLight.vector(w);
      double theta = SphericalTheta(wi), phi = SphericalPhi(wi);
      double sintheta = std::sin(theta);
      if (sintheta == 0.0) return 0.0;
      return l.distribution.pdf(phi * INV_TWOPI, theta * INV_PI) / (2.0 * kPi * kPi * sintheta);
    }
    return shapeSetPdf(l, p, w);
  }

  struct Vis { Ray r; };
  static void setSegment(Vis* v, const Vec& p1, double eps1, const Vec& p2, double eps2, double time) {
    double dist = Distance(p1, p2);  // visibility_tester.dart:26-29
    v->r = Ray(p1, (p2 - p1) / dist, eps1, dist * (1.0 - eps2), time);
  }
  Spec sampleLAtPoint(const Light& l, const Vec& p, double pEps, const U3& ls, double time, Vec* wi, double* pdf, Vis* vis) const {
    if (l.kind == 1) {  // point_light.dart:41-47
      *wi = Normalize(l.pos - p);
      *pdf = 1.0;
      setSegment(vis, p, pEps, l.pos, 0.0, time);
      return l.L / DistanceSquared(l.pos, p);
    }
    if (l.kind == 2) {  // distant_light.dart:41-48
      *wi = l.pos;
      *pdf = 1.0;
      vis->r = Ray(p, *wi, pEps, kInf, time);  // visibility_tester.dart:31-33
      return l.L;
    }
    if (l.kind == 3) {  // spot_light.dart:62-70 with falloff (:36-53)
      *wi = Normalize(l.pos - p);
      *pdf = 1.0;
      setSegment(vis, p, pEps, l.pos, 0.0, time);
      Vec wl = Normalize(l.worldToLight.vector(-*wi));
      double costheta = wl.z, falloff;
      if (costheta < l.cosTotalWidth) falloff = 0.0;
      else if (costheta > l.cosFalloffStart) falloff = 1.0;
      else {
        double delta = (costheta - l.cosTotalWidth) / (l.cosFalloffStart - l.cosTotalWidth);
        falloff = delta * delta * delta * delta;
      }
      return l.L * falloff / DistanceSquared(l.pos, p);
    }
    if (l.kind == 5) {  // projection_light.dart:102-109,115-139
      *wi = Normalize(l.pos - p);
      *pdf = 1.0;
      setSegment(vis, p, pEps, l.pos, 0.0, time);
      Vec wl = l.worldToLight.vector(-*wi);
      Spec proj(0.0);
      if (!(wl.z < l.hither)) {
        Vec Pl = l.lightProjection.point(wl);
        if (!(Pl.x < l.screen[0] || Pl.x > l.screen[1] || Pl.y < l.screen[2] || Pl.y > l.screen[3])) {
          if (l.radianceMap.levels == 0) proj = Spec(1.0);
          else proj = l.radianceMap.lookup(((double)Pl.x - l.screen[0]) / (l.screen[1] - l.screen[0]),
                                           ((double)Pl.y - l.screen[2]) / (l.screen[3] - l.screen[2]), 0.0);
        }
      }
      return l.L * proj / DistanceSquared(l.pos, p);
    }
    if (l.kind == 6) {  // goniometric_light.dart:58-86
      *wi = Normalize(l.pos - p);
      *pdf = 1.0;
      setSegment(vis, p, pEps, l.pos, 0.0, time);
      Vec wp = Normalize(l.worldToLight.vector(-*wi));
      std::swap(wp.y, wp.z);
      double theta = SphericalTheta(wp), phi = SphericalPhi(wp);
      if (l.radianceMap.levels == 0) return l.L * 1.0 / DistanceSquared(l.pos, p);
      return l.L * l.radianceMap.lookup(phi * INV_TWOPI, theta * INV_PI, 0.0) / DistanceSquared(l.pos, p);
    }
    if (l.kind == 4) {  // infinite_area_light.dart:93-131
      double uv[2], mapPdf;
      l.distribution.sampleContinuous(ls.u0, ls.u1, uv, &mapPdf);
      *pdf = 0.0;
      if (mapPdf == 0.0) return Spec(0.0);
      double theta = uv[1] * kPi, phi = uv[0] * 2.0 * kPi;
      double costheta = std::cos(theta), sintheta = std::sin(theta);
      double sinphi = std::sin(phi), cosphi = std::cos(phi);
      *wi = l.lightToWorld.vector(Vec(sintheta * cosphi, sintheta * sinphi, costheta));
      if (sintheta == 0.0) *pdf = 0.0;
      else *pdf = mapPdf / (2.0 * kPi * kPi * sintheta);
      vis->r = Ray(p, *wi, pEps, kInf, time);  // visibility_tester.dart:31-33
      return l.radiance(uv[0], uv[1], 0.0);
    }
    Vec ns;  // diffuse_area_light.dart:59-70
    Vec ps = shapeSetSample(l, ls, &ns, p);
    *wi = Normalize(ps - p);
    *pdf = shapeSetPdf(l, p, *wi);
    setSegment(vis, p, pEps, ps, 1.0e-3, time);
    return areaL(l, ns, -*wi);
  }

  // integrator.dart:119-185
  Spec EstimateDirect(const Light& light, int lightIndex, const Vec& p, const Vec& n, const Vec& wo, double rayEpsilon, double time,
                      const Bsdf& bsdf, const U3& lightSample, const U3& bsdfSample, int flags) {
    Spec Ld(0.0);
    Vec wi;
    double lightPdf = 0.0, bsdfPdf = 0.0;
    Vis vis;
    Spec Li = sampleLAtPoint(light, p, rayEpsilon, lightSample, time, &wi, &lightPdf, &vis);
    const bool delta = light.kind != 0 && light.kind != 4;  // isDeltaLight: point, distant, spot
    if (lightPdf > 0.0 && !Li.isBlack()) {
      Spec f = bsdf.f(wo, wi, flags);
      if (!f.isBlack() && !intersectP(vis.r)) {
        Li = Li * transmittance(vis.r, nullptr);  // integrator.dart:137
        if (delta) {
          Ld = Ld + f * Li * (AbsDot(wi, n) / lightPdf);
        } else {
          bsdfPdf = bsdf.pdf(wo, wi, flags);
          double weight = PowerHeuristic(1, lightPdf, 1, bsdfPdf);
          Ld = Ld + f * Li * ((AbsDot(wi, n) * weight / lightPdf));
        }
      }
    }
    if (!delta) {
      int sampledType = 0;
      Spec f = bsdf.sample_f(wo, &wi, bsdfSample, &bsdfPdf, flags, &sampledType);
      if (!f.isBlack() && bsdfPdf > 0.0) {
        double weight = 1.0;
        if ((sampledType & BSDF_SPECULAR) == 0) {
          lightPdf = this->lightPdf(light, p, wi);
          if (lightPdf == 0.0) return Ld;
          weight = PowerHeuristic(1, bsdfPdf, 1, lightPdf);
        }
        Isect lightIsect;
        Spec Li2(0.0);
        Ray ray(p, wi, rayEpsilon, kInf, time);
        if (intersect(ray, &lightIsect)) {
          if (g.lightOf[lightIsect.prim] == lightIndex) Li2 = isectLe(lightIsect, -wi);
        } else {
          Li2 = lightLe(light, ray);  // zero unless the light is infinite
        }
        if (!Li2.isBlack()) {
          Li2 = Li2 * transmittance(ray, nullptr);  // integrator.dart:178 (the ray ends at the light's surface, or never)
          Ld = Ld + f * Li2 * (AbsDot(wi, n) * weight / bsdfPdf);
        }
      }
    }
    return Ld;
  }

  // integrator.dart:79-117
  Spec UniformSampleOneLight(const Vec& p, const Vec& n, const Vec& wo, double rayEpsilon, double time, const Bsdf& bsdf,
                             const SampleVals& sample, Rng& rng, int lightNumOffset, const SampleOffsets* lightOff_,
                             const SampleOffsets* bsdfOff_) {
    int nLights = (int)rs.lights.size();
    if (nLights == 0) return Spec(0.0);
    int lightNum;
    if (lightNumOffset != -1) lightNum = (int)std::floor(sample.oneD[lightNumOffset][0] * (double)nLights);
    else lightNum = (int)std::floor(rng.randomFloat() * nLights);
    lightNum = std::min(lightNum, nLights - 1);
    U3 ls, bs;
    if (lightOff_ && bsdfOff_) { ls = U3::fromSample(sample, *lightOff_, 0); bs = U3::fromSample(sample, *bsdfOff_, 0); }
    else { ls = U3::random(rng); bs = U3::random(rng); }
    return EstimateDirect(rs.lights[lightNum], lightNum, p, n, wo, rayEpsilon, time, bsdf, ls, bs, BSDF_ALL & ~BSDF_SPECULAR) *
           (double)nLights;
  }

  // integrator.dart:39-77
  Spec UniformSampleAllLights(const Vec& p, const Vec& n, const Vec& wo, double rayEpsilon, double time, const Bsdf& bsdf,
                              const SampleVals& sample, Rng& rng) {
    Spec L(0.0);
    for (size_t i = 0; i < rs.lights.size(); ++i) {
      int nSamples = dlLight[i].nSamples;
      Spec Ld(0.0);
      for (int j = 0; j < nSamples; ++j) {
        U3 ls = U3::fromSample(sample, dlLight[i], j), bs = U3::fromSample(sample, dlBsdf[i], j);
        Ld = Ld + EstimateDirect(rs.lights[i], (int)i, p, n, wo, rayEpsilon, time, bsdf, ls, bs, BSDF_ALL & ~BSDF_SPECULAR);
      }
      L = L + Ld / (double)nSamples;
    }
    (void)rng;
    return L;
  }

  // path_integrator.dart:29-122
  Spec pathLi(const Ray& r, const Isect& isect, const SampleVals& sample, Rng& rng, const RayDiff* rd = nullptr) {
    Spec pathThroughput(1.0), L(0.0);
    Ray ray = r;
    bool specularBounce = false;
    Isect isectP = isect, localIsect;
    for (int bounces = 0;; ++bounces) {
      if (bounces == 0 || specularBounce) L = L + pathThroughput * isectLe(isectP, -ray.d);
      // the rays after the first are RayDifferential.child: no differentials (path_integrator.dart:100)
      Bsdf bsdf = getBSDF(isectP, bounces == 0 ? rd : nullptr);
      Vec p = bsdf.p, n = bsdf.nn, wo = -ray.d;
      if (bounces < 3)
        L = L + pathThroughput * UniformSampleOneLight(p, n, wo, isectP.rayEpsilon, ray.time, bsdf, sample, rng,
                                                       lightNumOff[bounces], &lightOff[bounces], &bsdfOff[bounces]);
      else
        L = L + pathThroughput * UniformSampleOneLight(p, n, wo, isectP.rayEpsilon, ray.time, bsdf, sample, rng, -1, nullptr, nullptr);
      U3 outgoing = bounces < 3 ? U3::fromSample(sample, pathOff[bounces], 0) : U3::random(rng);
      Vec wi;
      double pdf = 0.0;
      int flags = 0;
      Spec f = bsdf.sample_f(wo, &wi, outgoing, &pdf, BSDF_ALL, &flags);
      if (f.isBlack() || pdf == 0.0) break;
      specularBounce = (flags & BSDF_SPECULAR) != 0;
      pathThroughput = pathThroughput * (f * AbsDot(wi, n) / pdf);
      ray = Ray(p, wi, isectP.rayEpsilon, kInf, ray.time, ray.depth + 1);  // RayDifferential.child
      if (bounces > 3) {
        double continueProbability = dmin(0.5, pathThroughput.luminance());  // Math.min: NaN propagates (path_integrator.dart:94)
        if (rng.randomFloat() > continueProbability) break;
        pathThroughput = pathThroughput / continueProbability;
      }
      if (bounces == rs.integ.maxDepth) break;
      if (!intersect(ray, &localIsect)) {  // path_integrator.dart:106-114
        if (specularBounce)
          for (const Light& lt : rs.lights) L = L + pathThroughput * lightLe(lt, ray);
        break;
      }
      pathThroughput = pathThroughput * transmittance(ray, nullptr);  // path_integrator.dart:116
      isectP = localIsect;
    }
    return L;
  }

  // ambient_occlusion_integrator.dart:28-53
  Spec aoLi(const Ray& ray, const Isect& isect, Rng& rng, const RayDiff* rd = nullptr) {
    Bsdf bsdf = getBSDF(isect, rd);
    Vec p = bsdf.p;
    Vec n = FaceForward(isect.dg.nn, -ray.d);
    int nSamples = RoundUpPow2(rs.integ.aoSamples);
    uint32_t s0 = rng.randomUint(), s1 = rng.randomUint();
    int nClear = 0;
    for (int i = 0; i < nSamples; ++i) {
      double u0 = VanDerCorput(i, s0), u1 = Sobol2(i, s1);
      Vec w = UniformSampleSphere(u0, u1);
      if (Dot(w, n) < 0.0) w = -w;
      Ray r(p, w, rs.integ.aoMinDist, rs.integ.aoMaxDist);
      if (!intersectP(r)) ++nClear;
    }
    return Spec((double)nClear / nSamples);
  }

  // direct_lighting_integrator.dart:30-68
  Spec directLi(const Ray& ray, const Isect& isect, const SampleVals& sample, Rng& rng, const RayDiff* rd = nullptr) {
    Spec L(0.0);
    Bsdf bsdf = getBSDF(isect, rd);
    Vec wo = -ray.d, p = bsdf.p, n = bsdf.nn;
    L = L + isectLe(isect, wo);
    if (!rs.lights.empty()) {
      if (rs.integ.strategy == 0) L = L + UniformSampleAllLights(p, n, wo, isect.rayEpsilon, ray.time, bsdf, sample, rng);
      else L = L + UniformSampleOneLight(p, n, wo, isect.rayEpsilon, ray.time, bsdf, sample, rng, dlLightNum, &dlLight[0], &dlBsdf[0]);
    }
    if (ray.depth + 1 < rs.integ.maxDepth) {
      L = L + specularBranch(ray, bsdf, isect, sample, rng, BSDF_REFLECTION | BSDF_SPECULAR, rd);
      L = L + specularBranch(ray, bsdf, isect, sample, rng, BSDF_TRANSMISSION | BSDF_SPECULAR, rd);
    }
    return L;
  }

  // whitted_integrator.dart:26-78: emitted light, one LightSample.random(rng) per light (no multiple importance
  // sampling, every BxDF), then the specular recursion
  Spec whittedLi(const Ray& ray, const Isect& isect, const SampleVals& sample, Rng& rng, const RayDiff* rd = nullptr) {
    Spec L(0.0);
    Bsdf bsdf = getBSDF(isect, rd);
    Vec p = bsdf.p, n = bsdf.nn, wo = -ray.d;
    L = L + isectLe(isect, wo);
    for (size_t i = 0; i < rs.lights.size(); ++i) {
      Vec wi;
      double pdf = 0.0;
      Vis vis;
      Spec Li = sampleLAtPoint(rs.lights[i], p, isect.rayEpsilon, U3::random(rng), ray.time, &wi, &pdf, &vis);
      if (Li.isBlack() || pdf == 0.0) continue;
      Spec f = bsdf.f(wo, wi, BSDF_ALL);
      if (!f.isBlack() && !intersectP(vis.r)) L = L + f * Li * AbsDot(wi, n) * transmittance(vis.r, &sample) / pdf;  // whitted_integrator.dart:56-58
    }
    if (ray.depth + 1 < rs.integ.maxDepth) {
      L = L + specularBranch(ray, bsdf, isect, sample, rng, BSDF_REFLECTION | BSDF_SPECULAR, rd);
      L = L + specularBranch(ray, bsdf, isect, sample, rng, BSDF_TRANSMISSION | BSDF_SPECULAR, rd);
    }
    return L;
  }

  // Integrator.SpecularReflect / SpecularTransmit (integrator.dart:187-290): one BSDFSample.random(rng) is drawn whether or
  // not the BSDF has such a component; the child ray carries differentials when its parent does.
  Spec specularBranch(const Ray& ray, const Bsdf& bsdf, const Isect& isect, const SampleVals& sample, Rng& rng, int flags,
                      const RayDiff* rdIn = nullptr) {
    Vec wo = -ray.d, wi;
    double pdf = 0.0;
    Vec p = bsdf.p, n = bsdf.nn;
    U3 u = U3::random(rng);
    Spec f = bsdf.sample_f(wo, &wi, u, &pdf, flags, nullptr);
    Spec L(0.0);
    if (pdf > 0.0 && !f.isBlack() && AbsDot(wi, n) != 0.0) {
      Ray rd(p, wi, isect.rayEpsilon, kInf, ray.time, ray.depth + 1);  // RayDifferential.child
      RayDiff cd;
      if (rdIn && rdIn->has) {
        const DG& ds = bsdf.dgs;
        cd.has = true;
        cd.rxo = p + bsdf.dpdx;
        cd.ryo = p + bsdf.dpdy;
        Vec dndx = ds.dndu * ds.dudx + ds.dndv * ds.dvdx;
        Vec dndy = ds.dndu * ds.dudy + ds.dndv * ds.dvdy;
        Vec dwodx = -rdIn->rxd - wo, dwody = -rdIn->ryd - wo;
        double dDNdx = Dot(dwodx, n) + Dot(wo, dndx), dDNdy = Dot(dwody, n) + Dot(wo, dndy);
        if (flags & BSDF_REFLECTION) {  // :203-221
          cd.rxd = wi - dwodx + (dndx * Dot(wo, n) + n * dDNdx) * 2.0;
          cd.ryd = wi - dwody + (dndy * Dot(wo, n) + n * dDNdy) * 2.0;
        } else {  // :250-280
          double eta = bsdf.eta;
          Vec w = -wo;
          if (Dot(wo, n) < 0.0) eta = 1.0 / eta;
          double mu = eta * Dot(w, n) - Dot(wi, n);
          double dmudx = (eta - (eta * eta * Dot(w, n)) / Dot(wi, n)) * dDNdx;
          double dmudy = (eta - (eta * eta * Dot(w, n)) / Dot(wi, n)) * dDNdy;
          cd.rxd = wi + dwodx * eta - (dndx * mu + n * dmudx);
          cd.ryd = wi + dwody * eta - (dndy * mu + n * dmudy);
        }
      }
      Spec Li = LiRay(rd, sample, rng, cd.has ? &cd : nullptr);
      L = f * Li * (AbsDot(wi, n) / pdf);
    }
    return L;
  }

  // Camera.generateRayDifferential: perspective_camera.dart:93-132, orthographic_camera.dart:86-117, and the generic one-pixel
  // shifts of camera.dart:37-62 for the environment camera.  `rd` null: the main ray only.
  Ray cameraRay(const SampleVals& s, RayDiff* rd = nullptr) const {
    const Camera& c = rs.camera;
    // AnimatedTransform.transformRay / transformRayDifferential (animated_transform.dart:138-169): interpolate(ray.time), then the
    // Transform's own method; a static camera keeps startTransform
    const Transform c2w = c.c2w(s.time);
    if (c.kind == 2) {  // environment_camera.dart:42-52
      auto gen = [&](double imageX, double imageY) {
        double theta = kPi * imageY / rs.film.yres;
        double phi = 2 * kPi * imageX / rs.film.xres;
        Ray er(Vec(), Vec(std::sin(theta) * std::cos(phi), std::cos(theta), std::sin(theta) * std::sin(phi)), 0.0, kInf);
        er.time = s.time;
        Ray ew = c2w.ray(er);
        ew.time = s.time;
        return ew;
      };
      Ray ew = gen(s.imageX, s.imageY);
      if (rd) {  // camera.dart:40-58: sshift.imageX++, then imageX--, imageY++ (Dart doubles)
        Ray rx = gen(s.imageX + 1.0, s.imageY), ry = gen((s.imageX + 1.0) - 1.0, s.imageY + 1.0);
        rd->has = true;
        rd->rxo = rx.o; rd->rxd = rx.d; rd->ryo = ry.o; rd->ryd = ry.d;
      }
      return ew;
    }
    Vec Pras(s.imageX, s.imageY, 0.0);
    Vec Pcamera = c.rasterToCamera.point(Pras);
    Ray ray(Vec(0.0, 0.0, 0.0), Normalize(Pcamera), 0.0, kInf);
    if (c.kind == 1) ray = Ray(Pcamera, Vec(0.0, 0.0, 1.0), 0.0, kInf);  // orthographic_camera.dart:52-58
    if (c.lensRadius > 0.0) {
      double lu, lv;
      ConcentricSampleDisk(s.lensU, s.lensV, &lu, &lv);
      lu *= c.lensRadius;
      lv *= c.lensRadius;
      double ft = c.focalDistance / ray.d.z;
      Vec Pfocus = ray.at(ft);
      ray.o = Vec(lu, lv, 0.0);
      ray.d = Normalize(Pfocus - ray.o);
    }
    ray.time = s.time;
    Ray w = c2w.ray(ray);
    w.time = s.time;
    if (rd) {
      rd->has = true;
      if (c.kind == 0) {  // perspective_camera.dart:50-56,122-128: the offsets ignore the lens; transformRayDifferential
        Vec dxCamera = c.rasterToCamera.point(Vec(1.0, 0.0, 0.0)) - c.rasterToCamera.point(Vec(0.0, 0.0, 0.0));
        Vec dyCamera = c.rasterToCamera.point(Vec(0.0, 1.0, 0.0)) - c.rasterToCamera.point(Vec(0.0, 0.0, 0.0));
        rd->rxo = c2w.point(ray.o);
        rd->ryo = c2w.point(ray.o);
        rd->rxd = c2w.vector(Normalize(Pcamera + dxCamera));
        rd->ryd = c2w.vector(Normalize(Pcamera + dyCamera));
      } else {
        // orthographic_camera.dart:111-115 AS WRITTEN: rxOrigin / ryOrigin are built in camera space and the call that follows is
        // transformRay, not transformRayDifferential, so they STAY in camera space; rxDirection and ryDirection are the very
        // object ray.direction is, which transformRay overwrites in place: they end up as the world-space direction.
        Vec dxCamera = c.rasterToCamera.vector(Vec(1.0, 0.0, 0.0)), dyCamera = c.rasterToCamera.vector(Vec(0.0, 1.0, 0.0));
        rd->rxo = ray.o + dxCamera;
        rd->ryo = ray.o + dyCamera;
        rd->rxd = w.d;
        rd->ryd = w.d;
      }
    }
    return w;
  }
  // Sampler.samplesPerPixel as each sampler's constructor hands it to the base class (lib/samplers/*.dart)
  int samplesPerPixel() const {
    const SamplerCfg& sc = rs.sampler;
    switch (sc.kind) {
      case 0: return RoundUpPow2(sc.spp);
      case 1: return sc.xs * sc.ys;
      case 4: return RoundUpPow2(std::max(sc.xs, sc.ys));
      default: return sc.spp;
    }
  }

  // SamplerRenderer.Li (sampler_renderer.dart:67-98): also what the specular recursion calls
  int32_t lastCameraPrim = -1;  // primitive the last camera ray hit (adaptive sampler, shapeid method)
  Spec LiRay(const Ray& rayIn, const SampleVals& s, Rng& rng, const RayDiff* rd = nullptr) {
    Ray ray = rayIn;  // Scene.intersect shrinks ray.maxDistance to the hit (geometric_primitive.dart:47-61)
    Isect isect;
    Spec L(0.0);
    const bool hit = intersect(ray, &isect);
    if (rayIn.depth == 0) lastCameraPrim = hit ? isect.prim : -1;  // camera rays only: the recursion's rays have depth > 0
    if (hit) {
      switch (rs.integ.kind) {
        case 0: L = pathLi(ray, isect, s, rng, rd); break;
        case 1: L = aoLi(ray, isect, rng, rd); break;
        case 3: L = whittedLi(ray, isect, s, rng, rd); break;
        default: L = directLi(ray, isect, s, rng, rd); break;
      }
    } else {
      L = allLightsLe(ray);  // sampler_renderer.dart:86-92
    }
    // T * Li + Lvi (sampler_renderer.dart:93-97); `ray` ends at the surface hit
    Spec T(1.0);
    Spec Lvi = volumeLi(ray, s, &T);
    return T * L + Lvi;
  }

  // sampler_renderer.dart:173-193
  Spec Li(const SampleVals& s, Rng& rng) {
    RayDiff rd;
    const bool textured = !rs.programs.empty();  // the differentials only feed texture filtering and bump mapping
    Ray ray = cameraRay(s, textured ? &rd : nullptr);
    if (textured) rd.scale(ray.o, ray.d, 1.0 / std::sqrt((double)samplesPerPixel()));  // sampler_renderer.dart:166
    stats.cameraSamples++;
    Spec L = LiRay(ray, s, rng, textured ? &rd : nullptr);
    L = L * 1.0;  // rayWeight
    if (L.hasNaNs()) L = Spec(0.0);
    else if (L.luminance() < -1e-5) L = Spec(0.0);
    else if (std::isinf(L.luminance())) L = Spec(0.0);
    return L;
  }
};

// pixel visitation order, lib/pixel_samplers/*.dart
std::vector<int32_t> pixelOrder(const SamplerCfg& sc, int x, int y, int width, int height) {
  std::vector<int32_t> px;
  px.reserve((size_t)width * height * 2);
  int right = x + width - 1, bottom = y + height - 1;
  if (sc.pixelOrder == 0) {  // linear_pixel_sampler.dart:29-40
    for (int yy = y; yy <= bottom; ++yy)
      for (int xx = x; xx <= right; ++xx) { px.push_back(xx); px.push_back(yy); }
    return px;
  }
  // tile_pixel_sampler.dart:39-100
  int ts = sc.tileSize;
  int nx = width / ts + ((width % ts == 0) ? 0 : 1), ny = height / ts + ((height % ts == 0) ? 0 : 1);
  std::vector<int32_t> tiles;
  for (int yi = 0; yi < ny; ++yi)
    for (int xi = 0; xi < nx; ++xi) { tiles.push_back(xi); tiles.push_back(yi); }
  int numTiles = (int)tiles.size() / 2;
  DartRandom rng(5489);  // rng.dart:29
  for (int ti = 1; ti < numTiles; ++ti) {
    int lx = ti * 2, rx = (int)(rng.randomUint() % (uint32_t)numTiles) * 2;
    std::swap(tiles[lx], tiles[rx]);
    std::swap(tiles[lx + 1], tiles[rx + 1]);
  }
  for (int i = 0; i < numTiles; ++i) {
    int sx = x + tiles[2 * i] * ts, sy = y + tiles[2 * i + 1] * ts;
    for (int yi = 0; yi < ts; ++yi) {
      int yy = sy + yi;
      if (yy > bottom) break;
      for (int xi = 0; xi < ts; ++xi) {
        int xx = sx + xi;
        if (xx > right) break;
        px.push_back(xx);
        px.push_back(yy);
      }
    }
  }
  return px;
}

// GetSubWindow, common.dart:52-73
void GetSubWindow(int w, int h, int num, int count, int e[4]) {
  int nx = count, ny = 1;
  while ((nx & 0x1) == 0 && 2 * w * ny < h * nx) { nx >>= 1; ny <<= 1; }
  int xo = num % nx, yo = num / nx;
  double tx0 = (double)xo / nx, tx1 = (double)(xo + 1) / nx, ty0 = (double)yo / ny, ty1 = (double)(yo + 1) / ny;
  e[0] = (int)std::floor(Lerp(tx0, 0, w));
  e[1] = std::min((int)std::floor(Lerp(tx1, 0, w)), w);
  e[2] = (int)std::floor(Lerp(ty0, 0, h));
  e[3] = std::min((int)std::floor(Lerp(ty1, 0, h)), h);
}

// montecarlo.dart:327-339: `n = (n * invBase).toInt()` is a double product truncated, not an integer division
static double RadicalInverse(uint64_t n, int base) {
  double val = 0.0;
  const double invBase = 1.0 / base;
  double invBi = invBase;
  while (n > 0) {
    const int d_i = (int)(n % (uint64_t)base);
    val += d_i * invBi;
    n = (uint64_t)((double)n * invBase);
    invBi *= invBase;
  }
  return val;
}

// AdaptiveSampler's constructor (adaptive_sampler.dart:40-84): xs / ys carry minsamples / maxsamples
void adaptiveCounts(const SamplerCfg& sc, int* mn, int* mx) {
  int mins = sc.xs, maxs = sc.ys;
  if (mins > maxs) std::swap(mins, maxs);
  int a = RoundUpPow2(mins), b = RoundUpPow2(maxs);  // IsPowerOf2 ? itself : rounded up
  if (a < 2) a = 2;
  if (a == b) b *= 2;
  *mn = a;
  *mx = b;
}
// adaptive_sampler.dart:160-190.  `prims`: the camera ray's primitive id per sample, -1 for a miss — the reference compares
// Intersection.shapeId / primitiveId, which for an escaped ray are whatever an EARLIER sample left in the reused Intersection
// object (sampler_renderer.dart:147-176); that staleness only exists in a serial run and is not restated: a miss compares as -1.
bool needsSupersampling(int method, const std::vector<Spec>& Ls, const std::vector<int32_t>& prims, int count) {
  if (method == 0) {
    for (int i = 0; i < count - 1; ++i)
      if (prims[i] != prims[i + 1]) return true;
    return false;
  }
  double Lavg = 0.0;
  for (int i = 0; i < count; ++i) Lavg += Ls[i].luminance();
  Lavg /= count;
  const double maxContrast = 0.5;
  for (int i = 0; i < count; ++i)
    if (std::fabs(Ls[i].luminance() - Lavg) / Lavg > maxContrast) return true;
  return false;
}

// Samples of one pixel visit.  `pass` only matters for the random sampler (one visit per pass).
// Serial mode draws everything from `serial`; keyed mode derives streams from (seed, x, y).
int pixelSamples(const SamplerCfg& sc, const SampleLayout& layout, const Camera& cam, int x, int y, int pass, Rng* serial,
                 std::vector<SampleVals>& out) {
  const bool keyed = sc.rngMode == 1;
  if (sc.kind == 0 || sc.kind == 4) {  // low_discrepancy_sampler.dart:64-88 + montecarlo.dart:407-473
    // adaptive (adaptive_sampler.dart:101-131): the same LDPixelSample with minSamples (pass 0) or maxSamples (pass 1: the
    // pixel is being supersampled); the keyed streams of the second visit carry the pass so that it is a fresh draw
    int n = RoundUpPow2(sc.spp);
    if (sc.kind == 4) {
      int mn, mx;
      adaptiveCounts(sc, &mn, &mx);
      n = pass == 0 ? mn : mx;
    }
    const uint32_t keyPass = sc.kind == 4 ? (uint32_t)pass : 0u;
    out.resize(n);
    for (auto& s : out) s.alloc(layout);
    uint32_t arr = 0;
    auto stream = [&](KeyedRng& k) -> Rng& {
      if (!keyed) return *serial;
      k = KeyedRng(sc.seed, x, y, keyPass, arr);
      return k;
    };
    KeyedRng k;
    std::vector<float> img(2 * n), lens(2 * n), tm(n);
    LDShuffleScrambled2D(1, n, img.data(), stream(k)); arr++;
    LDShuffleScrambled2D(1, n, lens.data(), stream(k)); arr++;
    LDShuffleScrambled1D(1, n, tm.data(), stream(k)); arr++;
    std::vector<std::vector<float>> one(layout.n1D.size()), two(layout.n2D.size());
    for (size_t i = 0; i < layout.n1D.size(); ++i) {
      one[i].resize((size_t)layout.n1D[i] * n);
      LDShuffleScrambled1D(layout.n1D[i], n, one[i].data(), stream(k)); arr++;
    }
    for (size_t i = 0; i < layout.n2D.size(); ++i) {
      two[i].resize((size_t)2 * layout.n2D[i] * n);
      LDShuffleScrambled2D(layout.n2D[i], n, two[i].data(), stream(k)); arr++;
    }
    for (int i = 0; i < n; ++i) {
      out[i].imageX = x + (double)img[2 * i];
      out[i].imageY = y + (double)img[2 * i + 1];
      out[i].time = Lerp(tm[i], cam.shutterOpen, cam.shutterClose);
      out[i].lensU = lens[2 * i];
      out[i].lensV = lens[2 * i + 1];
      for (size_t j = 0; j < layout.n1D.size(); ++j)
        for (int q = 0; q < layout.n1D[j]; ++q) out[i].oneD[j][q] = one[j][(size_t)layout.n1D[j] * i + q];
      for (size_t j = 0; j < layout.n2D.size(); ++j)
        for (int q = 0; q < 2 * layout.n2D[j]; ++q) out[i].twoD[j][q] = two[j][(size_t)2 * layout.n2D[j] * i + q];
    }
    return n;
  }
  KeyedRng k(sc.seed, x, y, (uint32_t)pass, kStreamPixel);
  Rng& rng = keyed ? (Rng&)k : *serial;
  if (sc.kind == 5) {  // best_candidate_sampler.dart:36-52,74-132: (x, y) carry n = (tile index) * 4096 + tableOffset
    const uint64_t n = ((uint64_t)(uint32_t)y << 30) | (uint32_t)x;
    const double tableWidth = 64 / std::sqrt((double)sc.spp);
    const int right = sc.winX + sc.winW - 1, bottom = sc.winY + sc.winH - 1;
    const int xTileStart = (int)std::floor(sc.winX / tableWidth), xTileEnd = (int)std::floor(right / tableWidth);
    const int yTileStart = (int)std::floor(sc.winY / tableWidth);
    (void)bottom;
    const int nx = xTileEnd - xTileStart + 1;
    const uint64_t tile = n / 4096;
    const int to = (int)(n % 4096) * 5;
    const int xTile = xTileStart + (int)(tile % (uint64_t)nx), yTile = yTileStart + (int)(tile / (uint64_t)nx);
    DartRandom tileRng((int64_t)xTile + ((int64_t)yTile << 8));  // the tile's own RNG (:44-47,91-94): not the shared stream
    double so[3];
    for (int i = 0; i < 3; ++i) so[i] = tileRng.randomFloat();
    auto WRAP = [](double v) { return v > 1 ? (v - 1) : v; };
    const double* T = sc.sampleTable.data();
    const double imageX = (xTile + T[to]) * tableWidth, imageY = (yTile + T[to + 1]) * tableWidth;
    // as written (:117-118): BOTH coordinates are compared with left and right
    if (imageX < sc.winX || imageX > right || imageY < sc.winX || imageY > right) { out.clear(); return 0; }
    out.resize(1);
    out[0].alloc(layout);
    out[0].imageX = imageX;
    out[0].imageY = imageY;
    out[0].time = Lerp(WRAP(so[0] + T[to + 2]), cam.shutterOpen, cam.shutterClose);
    out[0].lensU = WRAP(so[1] + T[to + 3]);
    out[0].lensV = WRAP(so[2] + T[to + 4]);
    // integrator samples: LDShuffleScrambled1D / 2D with one pixel sample (:125-131); keyed like the lowdiscrepancy arrays
    uint32_t arr = 3;
    auto stream = [&](KeyedRng& kk) -> Rng& {
      if (!keyed) return *serial;
      kk = KeyedRng(sc.seed, x, y, 0, arr);
      return kk;
    };
    KeyedRng kk;
    for (size_t i = 0; i < layout.n1D.size(); ++i) { LDShuffleScrambled1D(layout.n1D[i], 1, out[0].oneD[i].data(), stream(kk)); arr++; }
    for (size_t i = 0; i < layout.n2D.size(); ++i) { LDShuffleScrambled2D(layout.n2D[i], 1, out[0].twoD[i].data(), stream(kk)); arr++; }
    return 1;
  }
  if (sc.kind == 3) {  // halton_sampler.dart:59-104: (x, y) carry the sequence index n = y * 2^30 + x, one sample per index
    const uint64_t n = ((uint64_t)(uint32_t)y << 30) | (uint32_t)x;
    const double u = RadicalInverse(n, 3), v = RadicalInverse(n, 2);
    const double lerpDelta = (double)std::max(sc.winW, sc.winH);
    const double left = sc.winX, top = sc.winY;
    const double imageX = left * (1.0 - u) + (left + lerpDelta) * u;  // Lerp, common.dart:80-81
    const double imageY = top * (1.0 - v) + (top + lerpDelta) * v;
    // `right` and `bottom` are the INCLUSIVE last pixel coordinates (sampler.dart:53-55), so the last column and row of the
    // window only receive samples on their left / top edge: as written
    if (imageX > sc.winX + sc.winW - 1 || imageY > sc.winY + sc.winH - 1) { out.clear(); return 0; }
    out.resize(1);
    out[0].alloc(layout);
    out[0].imageX = imageX;
    out[0].imageY = imageY;
    out[0].lensU = RadicalInverse(n + 1, 5);  // currentSample has been incremented by now (:82-89)
    out[0].lensV = RadicalInverse(n + 1, 7);
    out[0].time = Lerp(RadicalInverse(n + 1, 11), cam.shutterOpen, cam.shutterClose);
    for (size_t j = 0; j < layout.n1D.size(); ++j) LatinHypercube(out[0].oneD[j].data(), layout.n1D[j], 1, rng);
    for (size_t j = 0; j < layout.n2D.size(); ++j) LatinHypercube(out[0].twoD[j].data(), layout.n2D[j], 2, rng);
    return 1;
  }
  if (sc.kind == 1) {  // stratified_sampler.dart:67-124
    int n = sc.xs * sc.ys;
    out.resize(n);
    for (auto& s : out) s.alloc(layout);
    std::vector<float> img(2 * n), lens(2 * n), tm(n);
    StratifiedSample2D(img.data(), sc.xs, sc.ys, rng, sc.jitter != 0);
    StratifiedSample2D(lens.data(), sc.xs, sc.ys, rng, sc.jitter != 0);
    StratifiedSample1D(tm.data(), n, rng, sc.jitter != 0);
    for (int o = 0; o < 2 * n; o += 2) {  // float32 += int
      img[o] = f32((double)img[o] + x);
      img[o + 1] = f32((double)img[o + 1] + y);
    }
    Shuffle(lens.data(), 0, n, 2, rng);
    Shuffle(tm.data(), 0, n, 1, rng);
    for (int i = 0; i < n; ++i) {
      out[i].imageX = img[2 * i];
      out[i].imageY = img[2 * i + 1];
      out[i].lensU = lens[2 * i];
      out[i].lensV = lens[2 * i + 1];
      out[i].time = Lerp(tm[i], cam.shutterOpen, cam.shutterClose);
      for (size_t j = 0; j < layout.n1D.size(); ++j) LatinHypercube(out[i].oneD[j].data(), layout.n1D[j], 1, rng);
      for (size_t j = 0; j < layout.n2D.size(); ++j) LatinHypercube(out[i].twoD[j].data(), layout.n2D[j], 2, rng);
    }
    return n;
  }
  // random_sampler.dart:47-88 (FULL_SAMPLING: samplesPerPixel samples per visit)
  int n = sc.spp;
  out.resize(n);
  for (auto& s : out) s.alloc(layout);
  for (int si = 0; si < n; ++si) {
    out[si].imageX = rng.randomFloat() + x;
    out[si].imageY = rng.randomFloat() + y;
    out[si].lensU = rng.randomFloat();
    out[si].lensV = rng.randomFloat();
    out[si].time = Lerp(rng.randomFloat(), cam.shutterOpen, cam.shutterClose);
    for (size_t i = 0; i < layout.n1D.size(); ++i)
      for (int j = 0; j < layout.n1D[i]; ++j) out[si].oneD[i][j] = f32(rng.randomFloat());
    for (size_t i = 0; i < layout.n2D.size(); ++i)
      for (int j = 0; j < 2 * layout.n2D[i]; ++j) out[si].twoD[i][j] = f32(rng.randomFloat());
  }
  return n;
}

}  // namespace

static Bsdf canonicalBsdf(const RenderScene& rs, uint32_t material) {
  Bsdf b;
  b.p = Vec(0, 0, 0);
  b.nn = b.ng = Vec(0, 0, 1);
  b.sn = Vec(1, 0, 0);
  b.tn = Vec(0, 1, 0);
  for (Lobe l : rs.materials.at(material).lobes) {
    if (l.kind == 6 || l.kind == 7) l.measured = &rs.measured.at((size_t)l.param);
    b.bxdfs[b.nBxDFs++].init(l);
  }
  return b;
}

void RenderScene::bsdfEval(uint32_t material, uint32_t n, const double* wo, const double* wi, int flags, float* f, double* pdf) const {
  const Bsdf b = canonicalBsdf(*this, material);
  for (uint32_t i = 0; i < n; ++i) {
    const Vec o(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), w(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
    const Spec v = b.f(o, w, flags);
    for (int k = 0; k < 3; ++k) f[3 * i + k] = v.c[k];
    pdf[i] = b.pdf(o, w, flags);
  }
}

void RenderScene::bsdfSample(uint32_t material, uint32_t n, const double* wo, const double* u, int flags, double* wi, float* f,
                             double* pdf, int32_t* sampledType) const {
  const Bsdf b = canonicalBsdf(*this, material);
  for (uint32_t i = 0; i < n; ++i) {
    const Vec o(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
    U3 s;
    s.u0 = f32(u[3 * i]);  // BSDFSample keeps its two direction values in a Float32List (bsdf_sample.dart)
    s.u1 = f32(u[3 * i + 1]);
    s.comp = u[3 * i + 2];
    Vec w(0, 0, 0);
    int type = 0;
    double p = 0.0;
    const Spec v = b.sample_f(o, &w, s, &p, flags, &type);
    for (int k = 0; k < 3; ++k) f[3 * i + k] = v.c[k];
    wi[3 * i] = w.x; wi[3 * i + 1] = w.y; wi[3 * i + 2] = w.z;
    pdf[i] = p;
    sampledType[i] = type;
  }
}

void RenderScene::finalizeLights() {
  Ctx c(*this);
  for (Light& l : lights) {
    if (l.kind != 0) continue;
    l.areas.clear();
    l.area = 0.0;
    for (uint32_t sh : l.shapes) {  // shape_set.dart:43-50
      double a = c.shapeArea(sh);
      l.areas.push_back(a);
      l.area += a;
    }
    l.areaDistribution.init(l.areas);
  }
}

int RenderScene::samplesForPixel(int px, int py, std::vector<float>* out) {
  Ctx c(*this);
  c.requestSamples();
  std::vector<SampleVals> sv;
  DartRandom serial((int64_t)sampler.seed);
  int n = pixelSamples(sampler, c.layout, camera, px, py, 0, &serial, sv);
  int per = c.layout.floatsPerSample();
  out->clear();
  for (int i = 0; i < n; ++i) {
    out->push_back((float)(sv[i].imageX - px));
    out->push_back((float)(sv[i].imageY - py));
    out->push_back((float)sv[i].lensU);
    out->push_back((float)sv[i].lensV);
    out->push_back((float)sv[i].time);
    for (auto& a : sv[i].oneD) for (float v : a) out->push_back(v);
    for (auto& a : sv[i].twoD) for (float v : a) out->push_back(v);
  }
  return per;
}

void RenderScene::render(int taskNum, int taskCount, int nthreads) {
  finalizeLights();
  int ext[4];
  film.getSampleExtent(ext);
  // dartray.dart:1009-1023: the task renders a sub-window of the SAMPLE extent
  int x = ext[0], y = ext[2], w = ext[1] - ext[0], h = ext[3] - ext[2];
  if (taskCount > 1) {
    int e[4];
    GetSubWindow(w, h, taskNum, taskCount, e);
    x = ext[0] + e[0]; w = e[1] - e[0];
    y = ext[2] + e[2]; h = e[3] - e[2];
  }
  std::vector<int32_t> px;
  if (sampler.kind == 5) {  // best_candidate_sampler.dart:38-44: every table entry of every tile the window touches
    sampler.winX = x; sampler.winY = y; sampler.winW = w; sampler.winH = h;
    const double tableWidth = 64 / std::sqrt((double)sampler.spp);
    const int64_t nx = (int64_t)std::floor((x + w - 1) / tableWidth) - (int64_t)std::floor(x / tableWidth) + 1;
    const int64_t ny = (int64_t)std::floor((y + h - 1) / tableWidth) - (int64_t)std::floor(y / tableWidth) + 1;
    const uint64_t wanted = (uint64_t)(nx * ny) * 4096;
    px.reserve(2 * wanted);
    for (uint64_t n = 0; n < wanted; ++n) { px.push_back((int32_t)(n & 0x3fffffffu)); px.push_back((int32_t)(n >> 30)); }
  } else if (sampler.kind == 3) {  // halton_sampler.dart:32-38: spp * delta^2 sequence indices instead of pixels
    sampler.winX = x; sampler.winY = y; sampler.winW = w; sampler.winH = h;
    const uint64_t delta = (uint64_t)std::max(w, h), wanted = (uint64_t)sampler.spp * delta * delta;
    px.reserve(2 * wanted);
    for (uint64_t n = 0; n < wanted; ++n) { px.push_back((int32_t)(n & 0x3fffffffu)); px.push_back((int32_t)(n >> 30)); }
  } else {
    px = pixelOrder(sampler, x, y, w, h);
  }
  const size_t nPix = px.size() / 2;
  const int passes = sampler.kind == 2 ? sampler.spp : 1;  // random sampler quirk: spp visits of spp samples

  if (sampler.rngMode == 0 || nthreads <= 1) {
    Ctx c(*this);
    c.requestSamples();
    DartRandom serial(taskNum);  // sampler_renderer.dart:137
    std::vector<SampleVals> sv;
    for (int pass = 0; pass < passes; ++pass)
      for (size_t i = 0; i < nPix; ++i) {
        // adaptive sampler: the pixel is visited with minSamples and, if reportResults asks for it, again with maxSamples,
        // the first visit's samples being dropped (adaptive_sampler.dart:133-158, sampler_renderer.dart:199-207)
        for (int visit = 0; visit < (sampler.kind == 4 ? 2 : 1); ++visit) {
          const int ps = sampler.kind == 4 ? visit : pass;
          int n = pixelSamples(sampler, c.layout, camera, px[2 * i], px[2 * i + 1], ps, &serial, sv);
          std::vector<Spec> Ls(n);
          std::vector<int32_t> prims(n);
          for (int s = 0; s < n; ++s) {
            KeyedRng k(sampler.seed, px[2 * i], px[2 * i + 1], (uint32_t)(ps * n + s), kStreamIntegrator);
            KeyedRng kt(sampler.seed, px[2 * i], px[2 * i + 1], (uint32_t)(ps * n + s), kStreamTransmittance);
            KeyedRng kv(sampler.seed, px[2 * i], px[2 * i + 1], (uint32_t)(ps * n + s), kStreamVolumeLi);
            Rng& rng = sampler.rngMode == 1 ? (Rng&)k : (Rng&)serial;
            c.trRng = sampler.rngMode == 1 ? (Rng*)&kt : (Rng*)&serial;
            c.volRng = sampler.rngMode == 1 ? (Rng*)&kv : (Rng*)&serial;
            Ls[s] = c.Li(sv[s], rng);
            prims[s] = c.lastCameraPrim;
          }
          if (sampler.kind == 4 && visit == 0 && needsSupersampling(sampler.jitter, Ls, prims, n)) continue;
          for (int s = 0; s < n; ++s) film.addSample(sv[s].imageX, sv[s].imageY, Ls[s]);
          break;
        }
      }
    stats = c.stats;
    return;
  }
  // keyed mode, threaded: pixels are independent; film updates are applied in pixel order afterwards
  struct Contribution { double ix, iy; Spec L; };
  std::vector<std::vector<Contribution>> perThread(nthreads);
  std::vector<RenderStats> st(nthreads);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t] {
      Ctx c(*this);
      c.requestSamples();
      std::vector<SampleVals> sv;
      size_t b = nPix * t / nthreads, e = nPix * (t + 1) / nthreads;
      for (int pass = 0; pass < passes; ++pass)
        for (size_t i = b; i < e; ++i) {
          for (int visit = 0; visit < (sampler.kind == 4 ? 2 : 1); ++visit) {
            const int ps = sampler.kind == 4 ? visit : pass;
            int n = pixelSamples(sampler, c.layout, camera, px[2 * i], px[2 * i + 1], ps, nullptr, sv);
            std::vector<Spec> Ls(n);
            std::vector<int32_t> prims(n);
            for (int s = 0; s < n; ++s) {
              KeyedRng k(sampler.seed, px[2 * i], px[2 * i + 1], (uint32_t)(ps * n + s), kStreamIntegrator);
              KeyedRng kt(sampler.seed, px[2 * i], px[2 * i + 1], (uint32_t)(ps * n + s), kStreamTransmittance);
              KeyedRng kv(sampler.seed, px[2 * i], px[2 * i + 1], (uint32_t)(ps * n + s), kStreamVolumeLi);
              c.trRng = &kt;
              c.volRng = &kv;
              Ls[s] = c.Li(sv[s], k);
              prims[s] = c.lastCameraPrim;
            }
            if (sampler.kind == 4 && visit == 0 && needsSupersampling(sampler.jitter, Ls, prims, n)) continue;
            for (int s = 0; s < n; ++s) perThread[t].push_back({sv[s].imageX, sv[s].imageY, Ls[s]});
            break;
          }
        }
      st[t] = c.stats;
    });
  for (auto& t : th) t.join();
  for (int t = 0; t < nthreads; ++t) {
    for (const Contribution& c : perThread[t]) film.addSample(c.ix, c.iy, c.L);
    stats.cameraSamples += st[t].cameraSamples; stats.closestRays += st[t].closestRays; stats.shadowRays += st[t].shadowRays;
    stats.nodesVisited += st[t].nodesVisited; stats.primsTested += st[t].primsTested;
  }
}

}  // namespace orc
