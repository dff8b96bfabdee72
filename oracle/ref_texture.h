// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.h header).  PARITY UNPINNED.
//
// CPU restatement of what DartRay evaluates at a hit point before it can build the BSDF:
//   DifferentialGeometry        lib/core/differential_geometry.dart:27-205 (computeDifferentials :122-205)
//   RayDifferential             lib/core/ray_differential.dart:27-68
//   MIPMap                      lib/core/mipmap.dart:63-355 (pyramid :142-166, texel :183-204, lookup :206-222,
//                               lookup2 / EWA :224-339, triangle :341-355)
//   texture mappings            lib/core/texture/{uv,spherical,cylindrical,planar}_mapping_2d.dart
//   textures                    lib/core/texture/constant_texture.dart, lib/textures/{scale,mix,image}_texture.dart,
//                               checkerboard_texture.dart, uv_texture.dart, bilerp_texture.dart, and the noise textures
//                               {fbm,wrinkled,windy,marble,dots}_texture.dart, checkerboard_3d_texture.dart over Noise / FBm /
//                               Turbulence (lib/core/texture.dart:40-140) and IdentityMapping3D
//   Material.Bump               lib/core/material.dart:35-88
//   the materials' getBSDF      lib/materials/{matte,mirror,glass,plastic,metal,shiny_metal,substrate,translucent,uber,mix}
//                               _material.dart with textures that read the hit point
#pragma once
#include <cstdint>
#include <vector>

#include "ref_core.h"

namespace orc {

struct Spec;
struct Lobe;

// The whole of DifferentialGeometry (differential_geometry.dart:27-42).
struct DG {
  Vec p, nn, dpdu, dpdv, dndu, dndv, dpdx, dpdy;
  double u = 0.0, v = 0.0;
  double dudx = 0.0, dvdx = 0.0, dudy = 0.0, dvdy = 0.0;
  bool reverse = false;  // shape.reverseOrientation (transformSwapsHandedness is never set, shape.dart:30)
};

// The differential part of a RayDifferential in world space (ray_differential.dart:27-33)
struct RayDiff {
  bool has = false;
  Vec rxo, ryo, rxd, ryd;
  void scale(const Vec& o, const Vec& d, double s);  // scaleDifferentials, :56-61
};

// differential_geometry.dart:122-205
void computeDifferentials(DG* dg, const RayDiff& rd);

// One MIPMap built by MIPMap.texture (mipmap.dart:63-181) from its level 0 (power-of-two resolution: the reference's own
// constructor has resampled the file by then, :72-139); `channels` 1 = a float image, 3 = a spectrum image.
struct TexImage {
  int channels = 3, levels = 0, wrap = 0;  // wrap: 0 repeat, 1 black, 2 clamp (mipmap.dart:24-26)
  bool trilinear = false;
  double maxAniso = 8.0;
  std::vector<int> w, h;
  std::vector<std::vector<float>> data;  // per level, `channels` floats per texel
  void init(int width, int height, int channels_, const float* texels, int wrap_, bool trilinear_, double maxAniso_);
  // channels == 3
  void texelS(int level, int64_t s, int64_t t, float out[3]) const;
  void triangleS(int level, double s, double t, float out[3]) const;
  void lookupS(double s, double t, double width, float out[3]) const;
  void ewaS(int level, double s, double t, double ds0, double dt0, double ds1, double dt1, float out[3]) const;
  void lookup2S(double s, double t, double ds0, double dt0, double ds1, double dt1, float out[3]) const;
  // channels == 1: Dart doubles
  double texelF(int level, int64_t s, int64_t t) const;
  double triangleF(int level, double s, double t) const;
  double lookupF(double s, double t, double width) const;
  double ewaF(int level, double s, double t, double ds0, double dt0, double ds1, double dt1) const;
  double lookup2F(double s, double t, double ds0, double dt0, double ds1, double dt1) const;
};

struct TextureNode {
  int kind = 0;       // 0 constant, 1 scale, 2 mix, 3 imagemap, 4 checkerboard (2D), 5 uv, 6 bilerp, 7 fbm, 8 wrinkled, 9 windy,
                      // 10 marble, 11 dots, 12 checkerboard (3D): aaMethod = octaves, value = (roughness, scale, variation)
  int spectrum = 0;   // 0: Texture<double>, 1: Texture<Spectrum>
  int tex1 = -1, tex2 = -1, amount = -1;
  double value[3] = {0, 0, 0};  // constant; bilerp: v00 (then value2: v01, v10, v11)
  double value2[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int mapping = 0;    // 0 uv, 1 spherical, 2 cylindrical, 3 planar
  double su = 1, sv = 1, du = 0, dv = 0;  // uv mapping; planar: du, dv = ds, dt
  Transform worldToTexture;
  Vec v1, v2;         // planar
  int image = -1;     // index into TextureSet::images
  int aaMethod = 0;   // checkerboard: 0 none, 1 closedform
};

// A material whose parameters are textures (lib/materials/*.dart).  tex[] by kind:
//   0 matte        Kd, sigma                         5 shinymetal   Ks, Kr, roughness
//   1 mirror       Kr                                6 substrate    Kd, Ks, uroughness, vroughness
//   2 glass        Kr, Kt, index                     7 translucent  Kd, Ks, reflect, transmit, roughness
//   3 plastic      Kd, Ks, roughness                 8 uber         Kd, Ks, Kr, Kt, roughness, opacity, index
//   4 metal        eta, k, roughness                 9 mix          amount; m1 / m2 = material indices
//   10 subsurface / kdsubsurface   Kr, index
// kind -1: the material keeps its flattened lobe list (constant parameters, no bump map).
struct MaterialProgram {
  int kind = -1;
  int tex[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  int bump = -1;
  int m1 = -1, m2 = -1;
};

struct TextureSet {
  std::vector<TextureNode> nodes;
  std::vector<TexImage> images;
  double evalFloat(int id, const DG& dg) const;
  void evalSpec(int id, const DG& dg, float out[3]) const;
};

// material.dart:35-88
void Bump(const TextureSet& ts, int d, const DG& dgGeom, const DG& dgs, DG* dgBump);

}  // namespace orc
