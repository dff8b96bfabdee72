// See env_map.h.  Spectrum values are float32 triples and every Spectrum operation rounds to float32, as RGBColor's Float32List
// storage does in the reference (lib/core/rgb_color.dart:23-169); scalars are doubles.
#include "env_map.h"

#include <algorithm>
#include <cmath>
#include <cstddef>

namespace drt {
namespace {

struct Rgb {
  float v[3];
};
inline Rgb add(const Rgb& a, const Rgb& b) {
  return Rgb{{(float)((double)a.v[0] + b.v[0]), (float)((double)a.v[1] + b.v[1]), (float)((double)a.v[2] + b.v[2])}};
}
inline Rgb scale(const Rgb& a, double s) { return Rgb{{(float)((double)a.v[0] * s), (float)((double)a.v[1] * s), (float)((double)a.v[2] * s)}}; }
inline Rgb mul(const Rgb& a, const float b[3]) {
  return Rgb{{(float)((double)a.v[0] * b[0]), (float)((double)a.v[1] * b[1]), (float)((double)a.v[2] * b[2])}};
}
inline double log2Dart(double x) { return std::log(x) * (1.0 / std::log(2.0)); }  // common.dart:98-103

class Pyramid {  // mipmap.dart:142-166: each level averages 2 x 2 texels of the finer one (TEXTURE_REPEAT addressing)
 public:
  Pyramid(int width, int height, const float* rgb) {
    const int nLevels = 1 + (int)log2Dart((double)std::max(width, height));
    w_.push_back(width);
    h_.push_back(height);
    texels_.emplace_back((size_t)width * height);
    for (size_t i = 0; i < texels_[0].size(); ++i) texels_[0][i] = Rgb{{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]}};
    for (int lv = 1; lv < nLevels; ++lv) {
      const int sw = std::max(1, w_[lv - 1] / 2), sh = std::max(1, h_[lv - 1] / 2);
      std::vector<Rgb> next((size_t)sw * sh);
      for (int t = 0; t < sh; ++t)
        for (int s = 0; s < sw; ++s) {
          Rgb sum = add(add(add(at(lv - 1, 2 * s, 2 * t), at(lv - 1, 2 * s + 1, 2 * t)), at(lv - 1, 2 * s, 2 * t + 1)),
                        at(lv - 1, 2 * s + 1, 2 * t + 1));
          next[(size_t)t * sw + s] = scale(sum, 0.25);
        }
      w_.push_back(sw);
      h_.push_back(sh);
      texels_.push_back(std::move(next));
    }
  }
  int levels() const { return (int)texels_.size(); }
  const Rgb& at(int level, long long s, long long t) const {  // texel(), :183-204 (Dart's % is never negative)
    const long long W = w_[level], H = h_[level];
    s = ((s % W) + W) % W;
    t = ((t % H) + H) % H;
    return texels_[level][(size_t)(t * W + s)];
  }
  Rgb bilinear(int level, double s, double t) const {  // triangle(), :341-355
    level = std::min(std::max(level, 0), levels() - 1);
    s = s * w_[level] - 0.5;
    t = t * h_[level] - 0.5;
    const long long s0 = (long long)std::floor(s), t0 = (long long)std::floor(t);
    const double ds = s - s0, dt = t - t0;
    Rgb r = scale(at(level, s0, t0), (1.0 - ds) * (1.0 - dt));
    r = add(r, scale(at(level, s0, t0 + 1), (1.0 - ds) * dt));
    r = add(r, scale(at(level, s0 + 1, t0), ds * (1.0 - dt)));
    return add(r, scale(at(level, s0 + 1, t0 + 1), ds * dt));
  }
  Rgb trilinear(double s, double t, double width) const {  // lookup(), :206-222
    const double level = levels() - 1 + log2Dart(std::fmax(width, 1.0e-8));
    if (level < 0) return bilinear(0, s, t);
    if (level >= levels() - 1) return at(levels() - 1, 0, 0);
    const int il = (int)std::floor(level);
    const double delta = level - il;
    return add(scale(bilinear(il, s, t), 1.0 - delta), scale(bilinear(il + 1, s, t), delta));
  }

 private:
  std::vector<int> w_, h_;
  std::vector<std::vector<Rgb>> texels_;
};

// Distribution1D (montecarlo.dart:25-48) of `n` float32 function values: appends the cdf (n + 1 floats) to `cdf` and returns
// funcInt (a float32 value read back as a double)
float stepCdf(const float* func, int n, std::vector<float>* cdf) {
  const size_t base = cdf->size();
  cdf->resize(base + n + 1);
  float* c = cdf->data() + base;
  c[0] = 0.f;
  for (int i = 1; i <= n; ++i) c[i] = (float)((double)c[i - 1] + (double)func[i - 1] / n);
  const float total = c[n];
  for (int i = 1; i <= n; ++i) c[i] = total == 0.f ? (float)((double)i / n) : (float)((double)c[i] / (double)total);
  return total;
}

}  // namespace

void appendEnvTables(int width, int height, const float* rgb, const float L[3], std::vector<float>* out) {
  const Pyramid mip(width, height, rgb);
  out->insert(out->end(), rgb, rgb + 3 * (size_t)width * height);
  // scalar image the light is sampled from: luminance of the filtered radiance times sin(theta), :284-299
  std::vector<float> img((size_t)width * height);
  const double filter = 1.0 / std::max(width, height);
  for (int v = 0; v < height; ++v) {
    const double vp = (double)v / height, sinTheta = std::sin(3.141592653589793 * (v + 0.5) / height);
    for (int u = 0; u < width; ++u) {
      const Rgb r = mul(mip.trilinear((double)u / width, vp, filter), L);
      const float lum = (float)(0.212671 * r.v[0] + 0.715160 * r.v[1] + 0.072169 * r.v[2]);  // rgb_color.dart:167-169
      img[(size_t)v * width + u] = (float)((double)lum * sinTheta);
    }
  }
  out->insert(out->end(), img.begin(), img.end());
  std::vector<float> condCdf, condInt(height);
  for (int v = 0; v < height; ++v) condInt[v] = stepCdf(img.data() + (size_t)v * width, width, &condCdf);
  out->insert(out->end(), condCdf.begin(), condCdf.end());
  out->insert(out->end(), condInt.begin(), condInt.end());
  out->insert(out->end(), condInt.begin(), condInt.end());  // the marginal's function values ARE the rows' integrals (:230-236)
  std::vector<float> margCdf;
  const float margInt = stepCdf(condInt.data(), height, &margCdf);
  out->insert(out->end(), margCdf.begin(), margCdf.end());
  out->push_back(margInt);
}

}  // namespace drt
