// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.h header).  PARITY UNPINNED.
//
// CPU restatement of the rendering half of the DartRay hot path:
//   RNG                lib/core/rng.dart:27-43 (wraps dart:math Random — Dart SDK, NOT in /root/reference)
//   pixel samplers     lib/pixel_samplers/{linear,tile}_pixel_sampler.dart
//   samplers           lib/samplers/{low_discrepancy,stratified,random}_sampler.dart + lib/core/montecarlo.dart
//   camera             lib/cameras/perspective_camera.dart:93-132
//   shading geometry   lib/shapes/triangle.dart:100-154, lib/shapes/sphere.dart:118-160,
//                      lib/core/differential_geometry.dart:77-99
//   BSDF               lib/core/reflection/{bsdf,bxdf,lambertian,oren_nayar}.dart, lib/materials/matte_material.dart
//   lights             lib/lights/{diffuse_area,point}_light.dart, lib/core/light/shape_set.dart, lib/core/shape.dart:100-121
//   integrators        lib/core/integrator.dart:39-185, lib/surface_integrators/{path,ambient_occlusion,direct_lighting}_integrator.dart
//   renderer           lib/renderers/sampler_renderer.dart:67-98,118-218
//   film               lib/film/image_film.dart:51-185,247-299
//
// Two random-stream modes:
//   SERIAL — one sequential generator per task shared by sampler and integrators, as the reference
//            does (sampler_renderer.dart:137).  The generator restates the Dart VM's `Random` from its
//            published algorithm (sdk/lib/_internal/vm/lib/math_patch.dart + runtime/lib/math.cc;
//            version unpinned: pubspec.yaml has no SDK constraint) — a documented assumption.
//   KEYED  — counter-based streams keyed by (pixel, array) for the sampler and (pixel, sample) for the
//            integrators.  This is the layout the GPU replays; same algorithms, same draw order
//            inside each stream.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "ref_anim.h"
#include "ref_scene.h"
#include "ref_texture.h"

namespace orc {

// ---- spectrum: RGBColor, Float32List(3) storage (rgb_color.dart:23-169, spectrum.dart:1145) ----
struct Spec {
  float c[3] = {0.f, 0.f, 0.f};
  Spec() {}
  explicit Spec(double v) { c[0] = c[1] = c[2] = f32(v); }
  Spec(double r, double g, double b) { c[0] = f32(r); c[1] = f32(g); c[2] = f32(b); }
  bool isBlack() const { return !(c[0] != 0.f || c[1] != 0.f || c[2] != 0.f); }
  bool hasNaNs() const { return std::isnan(c[0]) || std::isnan(c[1]) || std::isnan(c[2]); }
  double luminance() const { return 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]; }  // rgb_color.dart:167-169
};
static inline Spec operator+(const Spec& a, const Spec& b) { return Spec((double)a.c[0] + b.c[0], (double)a.c[1] + b.c[1], (double)a.c[2] + b.c[2]); }
static inline Spec operator*(const Spec& a, const Spec& b) { return Spec((double)a.c[0] * b.c[0], (double)a.c[1] * b.c[1], (double)a.c[2] * b.c[2]); }
static inline Spec operator*(const Spec& a, double s) { return Spec((double)a.c[0] * s, (double)a.c[1] * s, (double)a.c[2] * s); }
static inline Spec operator-(const Spec& a, const Spec& b) { return Spec((double)a.c[0] - b.c[0], (double)a.c[1] - b.c[1], (double)a.c[2] - b.c[2]); }
static inline Spec operator/(const Spec& a, const Spec& b) { return Spec((double)a.c[0] / b.c[0], (double)a.c[1] / b.c[1], (double)a.c[2] / b.c[2]); }
static inline Spec operator/(const Spec& a, double s) { return Spec((double)a.c[0] / s, (double)a.c[1] / s, (double)a.c[2] / s); }

// ---- random streams -----------------------------------------------------------------------------
struct Rng {
  virtual ~Rng() {}
  virtual uint32_t next32() = 0;  // uniform 32-bit word
  virtual double randomFloat() = 0;
  // rng.dart:40-42: Random.nextInt(0xffffffff) -> [0, 2^32 - 2]
  virtual uint32_t randomUint() = 0;
};

// Restatement of the Dart VM `Random` (see header comment; unpinned).
struct DartRandom : Rng {
  uint32_t lo, hi;
  static uint64_t mix64(uint64_t n) {
    n = (~n) + (n << 21);
    n = n ^ (n >> 24);
    n = n * 265;
    n = n ^ (n >> 14);
    n = n * 21;
    n = n ^ (n >> 28);
    n = n + (n << 31);
    return n;
  }
  explicit DartRandom(int64_t seed) {
    uint64_t s = mix64((uint64_t)seed);
    if (s == 0) s = 0x5A17;
    lo = (uint32_t)s;
    hi = (uint32_t)(s >> 32);
    for (int i = 0; i < 4; ++i) nextState();
  }
  void nextState() {
    uint64_t st = 0xffffda61ull * lo + hi;
    lo = (uint32_t)st;
    hi = (uint32_t)(st >> 32);
  }
  uint32_t nextInt(uint64_t max) {
    if ((max & (~max + 1)) == max) {
      nextState();
      return lo & (uint32_t)(max - 1);
    }
    uint64_t rnd32, result;
    do {
      nextState();
      rnd32 = lo;
      result = rnd32 % max;
    } while ((rnd32 - result + max) > 4294967296ull);
    return (uint32_t)result;
  }
  uint32_t next32() override { nextState(); return lo; }
  double randomFloat() override {
    double a = nextInt(1u << 26), b = nextInt(1u << 27);
    return (a * 134217728.0 + b) / 9007199254740992.0;
  }
  uint32_t randomUint() override { return nextInt(0xffffffffull); }
};

// Counter-based stream (splitmix64 finaliser); the GPU implements exactly this.
struct KeyedRng : Rng {
  uint64_t key = 0, ctr = 0;
  static uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  static uint64_t makeKey(uint64_t seed, int32_t x, int32_t y, uint32_t sampleIdx, uint32_t streamId) {
    // the seed is hashed before it meets the pixel, so that seeds s and s^1 do not merely swap neighbouring pixels' streams
    uint64_t k1 = mix(mix(seed + 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)y << 32)));
    return mix(k1 ^ ((uint64_t)sampleIdx | ((uint64_t)streamId << 32)) ^ 0xD1B54A32D192ED03ull);
  }
  KeyedRng() {}
  KeyedRng(uint64_t seed, int32_t x, int32_t y, uint32_t sampleIdx, uint32_t streamId)
      : key(makeKey(seed, x, y, sampleIdx, streamId)), ctr(0) {}
  uint64_t next64() { ++ctr; return mix(key + ctr * 0x9E3779B97F4A7C15ull); }
  uint32_t next32() override { return (uint32_t)(next64() >> 32); }
  double randomFloat() override { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t randomUint() override { return (uint32_t)(next64() >> 32) % 0xffffffffu; }
};
static const uint32_t kStreamPixel = 4095;          // stratified / random sampler: one stream per pixel
static const uint32_t kStreamIntegrator = 0x80000000u;
// keyed mode with participating media: the draws the volume integrator's transmittance() makes inside the surface integrator
// (one per call, in call order) and the draws of its Li() — its own transmittance() calls along the single-scattering shadow rays
// included — have their own per-camera-sample streams, so that the wavefront can run the camera rays' volume integration and the
// surface integrator's bounces in either order.  Serial mode: the one RNG.
static const uint32_t kStreamTransmittance = 0x80000001u;
static const uint32_t kStreamVolumeLi = 0x80000002u;

// ---- scene description beyond geometry ----------------------------------------------------------------
// One BxDF of a material's BSDF with constant textures.  The materials of lib/materials/*.dart decompose into
// ordered lists of these (the order is the order of their bsdf.add calls, which BSDF.sample_f's component choice
// depends on, bsdf.dart:68-79):
//   matte    matte_material.dart:41-65      Lambertian | OrenNayar
//   mirror   mirror_material.dart:26-43     SpecularReflection(Kr, FresnelNoOp)
//   glass    glass_material.dart:26-52      SpecularReflection(Kr, FresnelDielectric(1, ior)) + SpecularTransmission(Kt, 1, ior)
//   plastic  plastic_material.dart:26-53    Lambertian(Kd) + Microfacet(Ks, FresnelDielectric(1.5, 1), Blinn(1/roughness))
//   metal    metal_material.dart:26-46      Microfacet(1, FresnelConductor(eta, k), Blinn(1/roughness))
//   uber     uber_material.dart:27-75       SpecularTransmission(1-op, 1, 1) + Lambertian + Microfacet + SpecularReflection + SpecularTransmission
// MeasuredMaterial's data (measured_material.dart:76-205): kind 0 = RegularHalfangleBRDF table (3 floats per cell, dims = nThetaH,
// nThetaD, nPhiD), kind 1 = the IrregIsotropicBRDFSamples of a .brdf file (dims[0] samples of 6 floats: BRDFRemap point, RGB value)
struct MeasuredTable {
  int kind = 0;
  int dims[3] = {0, 0, 0};
  std::vector<float> data;
};

struct Lobe {
  int kind = 0;     // 0 Lambertian, 1 OrenNayar, 2 Microfacet(Blinn), 3 SpecularReflection, 4 SpecularTransmission,
                    // 5 FresnelBlend(Rd = R, Rs = eta, Anisotropic(ex = param, ey = ei)) (fresnel_blend.dart, anisotropic.dart)
  Spec R;           // R / T of the BxDF (already clamped by the material)
  int fresnel = 0;  // 0 FresnelNoOp, 1 FresnelDielectric(ei, et), 2 FresnelConductor(eta, k)
  Spec eta, k;      // conductor
  double ei = 1.0, et = 1.0;  // FresnelDielectric / SpecularTransmission indices
  double param = 0.0;         // Blinn exponent (blinn.dart:24-28: clamped to 10000) | OrenNayar sigma in degrees
  // wrappers (lib/core/reflection/brdf_to_btdf.dart, scaled_bxdf.dart): bit 0 = BRDFToBTDF(bxdf) (TranslucentMaterial),
  // bit 1 = ScaledBxDF(<that>, scale) (MixMaterial)
  int wrap = 0;
  Spec scale = Spec(1.0);
  // kinds 6 RegularHalfangleBRDF / 7 IrregularIsotropicBRDF (regular_halfangle_brdf.dart, irregular_isotropic_brdf.dart): param = index
  // of the table in RenderScene::measured; the pointer is resolved when the BSDF is built
  const MeasuredTable* measured = nullptr;
};

struct Material {
  std::vector<Lobe> lobes;  // <= 8 (bsdf.dart:253)
  // matte_material.dart:41-65 with constant Kd / sigma
  static Material matte(const Spec& kd, double sigma);
};

struct Distribution1D {  // montecarlo.dart:25-98
  std::vector<float> func, cdf;
  double funcInt = 0;
  int count = 0;
  void init(const std::vector<double>& f);
  int sampleDiscrete(double u) const;
  double sampleContinuous(double u, double* pdf, int* off) const;  // :50-80
};

struct Distribution2D {  // montecarlo.dart:222-268
  std::vector<Distribution1D> pConditionalV;
  Distribution1D pMarginal;
  void init(const std::vector<float>& data, int nu, int nv);
  void sampleContinuous(double u0, double u1, double uv[2], double* pdf) const;
  double pdf(double u, double v) const;
};

// lib/core/mipmap.dart restricted to what InfiniteAreaLight uses: TEXTURE_REPEAT, the box pyramid (:142-166), texel (:183-204),
// triangle (:341-355) and the trilinear lookup (:206-222).  Level 0 arrives at power-of-two resolution: the reference's own
// constructor has already resampled it (:72-139) when the Dart side reads `radianceMap.pyramid[0]`.
struct MipMap {
  int levels = 0;
  std::vector<int> w, h;
  std::vector<std::vector<Spec>> pyramid;
  void init(int width, int height, const float* rgb);
  Spec texel(int level, int64_t s, int64_t t) const;
  Spec triangle(int level, double s, double t) const;
  Spec lookup(double s, double t, double width) const;
};

struct Light {
  int kind = 0;  // 0 = DiffuseAreaLight, 1 = PointLight, 2 = DistantLight, 3 = SpotLight, 4 = InfiniteAreaLight,
                 // 5 = ProjectionLight (projection_light.dart), 6 = GoniometricLight (goniometric_light.dart)
  Spec L;        // Lemit / intensity / radiance
  Vec pos;       // point / spot light position (world); distant light: lightDir (distant_light.dart:26)
  Transform worldToLight;                            // spot light (spot_light.dart:38-53)
  double cosTotalWidth = 0, cosFalloffStart = 0;     // spot light (spot_light.dart:29-30)
  int nSamples = 1;
  std::vector<uint32_t> shapes;  // ShapeSet order (shape_set.dart:26-41)
  std::vector<double> areas;
  double area = 0;
  Distribution1D areaDistribution;
  // InfiniteAreaLight (infinite_area_light.dart:37-69,276-306): worldToLight above + lightToWorld, the radiance map, its
  // sampling distribution
  Transform lightToWorld;
  MipMap radianceMap;
  Distribution2D distribution;
  Spec radiance(double u, double v, double width) const { return radianceMap.lookup(u, v, width) * L; }  // :240-242
  void setRadianceMap(int width, int height, const float* rgb);
  // ProjectionLight (projection_light.dart:38-100): lightProjection, the screen window and hither; both it and the
  // GoniometricLight look their map up at width 0 (radianceMap above; levels == 0: no map -> 1)
  Transform lightProjection;
  double screen[4] = {-1, 1, -1, 1}, hither = 1.0e-3;
};

struct Camera {  // perspective_camera.dart:46-57 + projective_camera.dart:34-53
  Transform rasterToCamera, cameraToWorld;
  // Camera.cameraToWorld is an AnimatedTransform (camera.dart:27, dartray.dart:971-975): `animated` when the end-time CTM differs
  bool animated = false;
  AnimatedTransform cameraMotion;
  Transform c2w(double time) const { return animated ? cameraMotion.interpolate(time) : cameraToWorld; }
  double lensRadius = 0, focalDistance = 1e30, shutterOpen = 0, shutterClose = 1;
  int kind = 0;  // 0 perspective, 1 orthographic (orthographic_camera.dart:52-80), 2 environment (environment_camera.dart:42-52)
};

struct Film {  // image_film.dart:51-97
  int xres = 0, yres = 0;
  double crop[4] = {0, 1, 0, 1};
  double xWidth = 0.5, yWidth = 0.5, invXWidth = 2, invYWidth = 2;
  float table[256];
  int left = 0, top = 0, width = 0, height = 0;
  std::vector<float> Lxyz, weightSum;
  void configure();
  void getSampleExtent(int e[4]) const;
  void addSample(double imageX, double imageY, const Spec& L);  // image_film.dart:99-185
  void writeImage(float* rgb) const;                            // image_film.dart:268-299
};

struct SamplerCfg {
  int kind = 0;  // 0 lowdiscrepancy, 1 stratified, 2 random, 3 halton, 4 adaptive (xs = minsamples, ys = maxsamples, jitter = method),
                 // 5 bestcandidate (best_candidate_sampler.dart; its 4096 x 5 pattern arrives through orc_set_sample_table)
  std::vector<double> sampleTable;  // bestcandidate: _SAMPLE_TABLE (best_candidate_sampler.dart:163-4258), handed over by the caller
  // halton: the sampler's window (left, top, width, height), filled in by render() (halton_sampler.dart:32-38)
  int winX = 0, winY = 0, winW = 0, winH = 0;
  int xs = 2, ys = 2;
  int spp = 4;
  int jitter = 1;
  int pixelOrder = 1;  // 0 linear, 1 tile
  int tileSize = 32;
  uint64_t seed = 0;
  int rngMode = 1;  // 0 serial (reference), 1 keyed (what the GPU replays)
};

struct IntegratorCfg {
  int kind = 0;  // 0 path, 1 ambientocclusion, 2 directlighting, 3 whitted
  int maxDepth = 5;
  int strategy = 0;  // directlighting: 0 = all, 1 = one
  int aoSamples = 2048;
  double aoMinDist = 1e-4, aoMaxDist = kInf;
};

// VolumeRegion plugins (lib/volume_regions/*.dart) and the volume integrator (lib/volume_integrators/*.dart)
struct VolumeRegionCfg {
  int kind = 0;  // 0 homogeneous, 1 exponential, 2 volumegrid
  Spec sigA, sigS, le;
  double g = 0.0;
  Vec p0, p1;            // extent = BBox(p0, p1) in volume space
  Transform worldToVolume;
  double a = 1.0, b = 1.0;  // exponential: density = a * exp(-b * height)
  Vec upDir;                // normalised (exponential_density_region.dart:29)
  int nx = 1, ny = 1, nz = 1;
  std::vector<double> density;  // volumegrid: Float64List, z-major (volume_grid.dart:73)
};
struct VolumeCfg {
  std::vector<VolumeRegionCfg> regions;  // > 1: AggregateVolume (dartray.dart:604-612)
  int integrator = 0;                    // 0 emission (the default, render_options.dart:24-39), 1 single
  double stepSize = 1.0;
};

struct RenderStats {
  uint64_t cameraSamples = 0, closestRays = 0, shadowRays = 0, nodesVisited = 0, primsTested = 0;
};

struct RenderScene {
  Scene* geom = nullptr;
  std::vector<Material> materials;
  // textures that read the hit point and the materials built from them (ref_texture.h); programs is empty or one per material
  TextureSet textures;
  std::vector<MaterialProgram> programs;
  std::vector<MeasuredTable> measured;
  std::vector<Light> lights;
  Camera camera;
  Film film;
  SamplerCfg sampler;
  IntegratorCfg integ;
  VolumeCfg volume;
  RenderStats stats;

  void finalizeLights();
  // _SamplerRendererTask.run for task taskNum of taskCount (sampler_renderer.dart:118-218 with the
  // sub-window of dartray.dart:1009-1023); adds into film.
  void render(int taskNum, int taskCount, int nthreads);
  // The camera samples only (imageX, imageY, lensU, lensV, time + integrator arrays) of one pixel,
  // for sampler parity tests.  Returns floats per sample.
  int samplesForPixel(int px, int py, std::vector<float>* out);
  // Test probes of the BSDF a material builds (Material.getBSDF with the canonical frame sn = +x, tn = +y, nn = ng = +z):
  // BSDF.f / BSDF.pdf for n (wo, wi) pairs, and BSDF.sample_f for n (wo, (u0, u1, component)) pairs — bsdf.dart:53-198.
  void bsdfEval(uint32_t material, uint32_t n, const double* wo, const double* wi, int flags, float* f, double* pdf) const;
  void bsdfSample(uint32_t material, uint32_t n, const double* wo, const double* u, int flags, double* wi, float* f, double* pdf,
                  int32_t* sampledType) const;
};

}  // namespace orc
