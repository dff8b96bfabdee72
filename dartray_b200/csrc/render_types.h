// Device-side description of what the wavefront renderer reads: materials, lights, camera, film,
// sampler layout and integrator parameters (see DESIGN.md "Render pipeline").
#pragma once
#include <cstdint>

#include "anim_transform.h"
#include "gpu_types.h"

namespace drt {

struct GMaterial {  // matte_material.dart:41-65 with constant textures
  float kd[3];
  float sigma;
};

// One BxDF of a material's BSDF (drt_set_material_lobes; lib/core/reflection/*.dart with constant textures):
// kind 0 Lambertian, 1 OrenNayar (param = sigma in degrees), 2 Microfacet with a Blinn distribution (param = exponent),
// 3 SpecularReflection, 4 SpecularTransmission (ei, et); fresnel 0 FresnelNoOp, 1 FresnelDielectric(ei, et),
// 2 FresnelConductor(eta, k).
struct GLobe {
  int32_t kind, fresnel;
  float rgb[3];
  float eta[3];
  float k[3];
  int32_t wrap;  // bit 0: BRDFToBTDF(bxdf) (brdf_to_btdf.dart), bit 1: ScaledBxDF(.., scale) (scaled_bxdf.dart); drt_set_lobe_wrappers
  double param, ei, et;
  float scale[3];
  float pad_;
};
static_assert(sizeof(GLobe) == 88, "GLobe layout");
// Lobe kinds 6 RegularHalfangleBRDF / 7 IrregularIsotropicBRDF (regular_halfangle_brdf.dart, irregular_isotropic_brdf.dart: the two BxDFs
// of MeasuredMaterial): `param` is the table index the caller gave (drt_set_measured); when the scene tables are uploaded the library
// writes where the table lives INTO the lobe — the bits of the device pointer in `et`, the three dimensions as int bits in `k` — so
// the BxDF code needs nothing but the lobe, like every other kind (texture-pass lobes copy the same fields from GMeasured).
struct GMeasured {
  const float* data;  // kind 0: nThetaH x nThetaD x nPhiD cells of RGB; kind 1: n samples of (BRDFRemap point, RGB)
  int32_t dims[3];
  int32_t kind;
};
static_assert(sizeof(GMeasured) == 24, "GMeasured layout");

struct GLight {
  int32_t kind;  // 0 = DiffuseAreaLight (diffuse_area_light.dart), 1 = PointLight (point_light.dart),
                 // 2 = DistantLight (distant_light.dart; pos = lightDir), 3 = SpotLight (spot_light.dart),
                 // 4 = InfiniteAreaLight (infinite_area_light.dart), 5 = ProjectionLight (projection_light.dart),
                 // 6 = GoniometricLight (goniometric_light.dart)
  float L[3];    // Lemit / intensity / radiance
  float pos[3];
  float w2l[9];  // spot: rows of worldToLight's upper 3x3 (Transform.transformVector, transform.dart:139-146)
  double cosTotalWidth, cosFalloffStart;
  int32_t nSamples;
  uint32_t shapeOffset, nShapes;  // ShapeSet (shape_set.dart:26-50): slice of lightShapes / lightShapeAreas
  uint32_t cdfOffset;             // slice of lightCdf: nShapes + 1 floats (Distribution1D, montecarlo.dart:25-48)
  double area;
  // InfiniteAreaLight: rows of lightToWorld's upper 3x3 (w2l above holds worldToLight's), the radiance map's resolution and
  // where its tables start in RenderScene::envData (floats):
  //   texels 3 x W x H | conditional func W x H | conditional cdf H x (W + 1) | conditional funcInt H |
  //   marginal func H | marginal cdf H + 1 | marginal funcInt 1          (Distribution2D, montecarlo.dart:222-268)
  float l2w[9];
  int32_t mapW, mapH;  // projection / goniometric lights: their map (texels only at envOffset), 0 x 0 = no map
  uint32_t envOffset;
  // ProjectionLight (projection_light.dart:38-100)
  float proj[16];  // lightProjection
  double screen[4], hither;
};

// Per direct-lighting light: where its LightSampleOffsets / BSDFSampleOffsets live in a sample record
// (direct_lighting_integrator.dart:70-96); value indices into the integrator part of the record.
struct DirectOffsets {
  int32_t nSamples;
  int32_t lightComp, lightPos, bsdfComp, bsdfPos;
};

// One shape of a light's ShapeSet with everything the light-sampling code needs resident in one record.  For a
// triangle the two normals are constants of the shape (triangle.dart:100-154 with the default uvs; :366-383), so
// they are computed once on the host with the same arithmetic instead of at every hit.
struct GLightShape {
  float p1[3], p2[3], p3[3];  // triangle vertices (unused for quadrics)
  float nn[3];                // dg.nn of any hit: normalize(cross(dpdu, dpdv)), flipped by reverseOrientation
  float ns[3];                // Triangle.sample's normal: normalize(cross(p2 - p1, p3 - p1)), flipped likewise
  uint32_t prim;              // primitive id (>= ntris: sphere / disk, evaluated by the general shape code)
  double area;
};

// One TriangleMesh of drt_set_mesh_shading: the upper 3x3 of objectToWorld (Transform.transformVector) and of
// worldToObject (Transform.transformNormal multiplies by its transpose, transform.dart:147-161), and the attribute flags.
struct GMesh {
  float o2w[9], w2o[9];
  uint32_t flags;  // bit 0 N, 1 S, 2 uv
  float pad_;
};

// One VolumeRegion (lib/volume_regions/*.dart): kind 0 homogeneous, 1 exponential density, 2 volumegrid.
struct GVolume {
  int32_t kind, nx, ny, nz;
  float sigA[3], sigS[3], sigT[3], le[3];  // sigT = the Spectrum sig_a + sig_s
  float lo[3], hi[3];                      // extent = BBox(p0, p1) in volume space
  float w2v[16];                           // worldToVolume
  double g, a, b;
  float up[3];                             // normalised up direction (exponential)
  uint32_t densityOffset;                  // first value of the grid in RenderScene::volDensity
};

// One texture node.  Image levels live in RenderScene::texData (floats): level l of the image at levelOffset[l], w >> l x h >> l
// (at least 1) texels of `channels` floats; texData[0..127] is MIPMap.weightLut (mipmap.dart:168-176).
struct GTex {
  int32_t kind, spectrum, tex1, tex2, amount, mapping;
  int32_t w, h, channels, wrap, trilinear, aa, levels, pad_;
  uint32_t levelOffset[16];
  double value[3], value2[9];
  double su, sv, du, dv, maxAniso;
  float w2t[16];
  float v1[3], v2[3];
};
struct GProgram {
  int32_t kind;
  int32_t tex[8];
  int32_t bump, m1, m2;
};

struct RenderScene {
  TraceScene ts;
  uint32_t ntris, nprims;
  const uint32_t* primToRec;  // primitive id -> GPrim record
  const uint32_t* primAttr;   // bits 0-15 material, 16-30 light + 1, 31 reverseOrientation
  const GMaterial* materials;
  // general materials (any BxDF list): material m owns lobes [matLobes[m].x, matLobes[m].x + matLobes[m].y)
  int32_t general;  // 0: every material is matte and the shading kernels take the single-lobe path
  int32_t nMaterials;
  const uint2* matLobes;
  const GLobe* lobes;
  const GLight* lights;
  int32_t nLights;
  const GLightShape* lightShapes;
  const float* lightCdf;
  // per-vertex shading attributes (null when no mesh carries any: triangle.dart:273-276 copies dg)
  const uint32_t* meshOfTri;  // ntris
  const uint32_t* triIdx;     // ntris x 3 vertex indices
  const GMesh* meshes;
  const float* vertN; const float* vertS; const float* vertUV;
  const float* envData;  // radiance maps and sampling tables of the infinite lights (GLight::envOffset)
  int32_t nInfinite;     // number of InfiniteAreaLights: escaped rays pick up their Le (sampler_renderer.dart:86-92)
  int32_t extra;  // 1: mesh attributes or quadrics of shape >= 2 are present (selects the kernels compiled with EXTRA)
  // participating media (scene.volumeRegion + the volume integrator; nVolumes == 0: none, transmittance == 1 without a draw)
  const GVolume* volumes;
  int32_t nVolumes;
  const double* volDensity;
  int32_t volIntegrator;  // 0 emission, 1 single
  double volStep;
  // textures that read the hit point and the materials built from them (drt_set_textures / drt_set_material_programs):
  // evaluated by the texture pass (texture_kernels.cu) into Wavefront::hitLobes before a queue is shaded
  const GTex* textures;
  const float* texData;
  const GProgram* programs;  // nPrograms == 0 or one per material
  int32_t nPrograms;
  const GMeasured* measured;  // drt_set_measured (material program kind 11 reads it; flattened lobes carry their table themselves)
  int32_t nMeasured;
};

struct RenderParams {
  // camera: perspective_camera.dart:46-57,93-132 + projective_camera.dart:34-53
  float rasterToCamera[16], cameraToWorld[16];
  // Camera.cameraToWorld is an AnimatedTransform (camera.dart:27): non-null when the camera moves (drt_set_camera_motion) — the
  // camera kernels then use cameraMotion->interpolate(sample time) instead of cameraToWorld (animated_transform.dart:138-169)
  const GInstance* cameraMotion;
  double lensRadius, focalDistance, shutterOpen, shutterClose;
  int32_t cameraKind;  // 0 perspective, 1 orthographic (orthographic_camera.dart:52-80), 2 environment (environment_camera.dart:42-52)
  // film: image_film.dart:51-97
  int32_t xres, yres, left, top, width, height;
  double xWidth, yWidth, invXWidth, invYWidth;
  const float* filterTable;  // 256 floats
  double* film;              // width*height x (X, Y, Z, weight)
  // sampler
  int32_t samplerKind;  // 0 lowdiscrepancy, 1 stratified, 2 random, 3 halton, 4 adaptive (lowdiscrepancy samples, two visits),
                        // 5 bestcandidate
  // bestcandidate (best_candidate_sampler.dart:36-52): the 4096 x 5 pattern, the three shifts of every tile the window touches,
  // the tile grid and the table's width in pixels
  const double* bcTable;
  const double* bcTileShifts;
  int32_t bcXTileStart, bcYTileStart, bcTilesX;
  double bcTableWidth;
  int32_t adaptiveMethod;  // adaptive_sampler.dart:37-38: 0 compare shape ids, 1 contrast threshold
  int32_t winX, winY, winW, winH;  // halton: the sampler's window (halton_sampler.dart:32-38), set per render call
  int32_t xs, ys, jitter;
  int32_t nPixelSamples;  // samples per pixel visit
  uint64_t seed;
  double diffScale;  // 1 / sqrt(sampler.samplesPerPixel): RayDifferential.scaleDifferentials, sampler_renderer.dart:166
  int32_t nVals;  // integrator values per sample (sum n1D + 2 sum n2D)
  int32_t ldAllSingle;  // lowdiscrepancy: every array holds one value per camera sample (path / AO): index-shuffle fast path
  // integrator
  int32_t integKind;  // 0 path, 1 ambientocclusion, 2 directlighting
  int32_t maxDepth, strategy, aoSamples;
  double aoMinDist, aoMaxDist;
  // path_integrator.dart:124-131: value indices for bounces 0..2
  int32_t pLightComp[3], pLightPos[3], pLightNum[3], pBsdfComp[3], pBsdfPos[3], pPathComp[3], pPathPos[3];
  // volume integrator's two one-D samples (emission_integrator.dart:26-29): value indices
  int32_t pTauSample, pScatterSample;
  // direct lighting, strategy "one": light number value; per-light offsets in `direct`
  int32_t dlLightNum;
  const DirectOffsets* direct;
};

// One sampler array (montecarlo.dart:524-551 works array by array)
struct SampleArray {
  int32_t dims;      // 1 or 2
  int32_t nSamples;  // values per camera sample
  int32_t dest;      // >= 0: first value index in the integrator record; -1: image (x,y); -2: lens (u,v); -3: time
  uint32_t streamId; // position in the reference's generation order (keys the stream)
};

// Work queues and per-sample state of one batch of camera samples.  `cap` slots.
struct Wavefront {
  uint32_t cap;
  // per slot
  int32_t* pixX; int32_t* pixY;      // pixel of the slot
  uint32_t* sampleIdx;               // index of the sample inside its pixel (keys the integrator stream)
  double2* camXY; double2* camLens;  // imageX, imageY / lensU, lensV
  double* camTime;                   // time sample in [0,1): a float32 value for the samplers that keep a Float32List of them, binary64 otherwise
  float* vals;                       // nVals x cap, value-major
  float* L;                          // 3 x cap, channel-major: radiance accumulated so far
  float* T;                          // 3 x cap: path throughput
  // what the last shaded vertex of a slot waits for (integrator.dart:119-185)
  float* pendSh;                     // 3 x cap: contribution if the shadow ray is unoccluded
  float* pendMisF;                   // 3 x cap: f of the BSDF sample
  double* pendMisScale;
  float* pendT;                      // 3 x cap: throughput the direct estimate is multiplied by
  int32_t* shIdx; int32_t* misIdx; int32_t* misLight;
  uint8_t* specBounce;               // path integrator: the last sampled BxDF was specular (path_integrator.dart:85)
  // directlighting with specular BxDFs (integrator.dart:187-290): the recursion is evaluated chain by chain (one
  // reflect / transmit choice per level), see renderBatch.  All null unless the scene needs it.
  uint32_t* specCtr;                 // per slot: draws of the integrator stream consumed so far (BSDFSample.random per branch call)
  uint32_t* specCtrAt;               // [level][slot]: the counter the branch call of that level used (chains re-walk their prefix)
  float4* bakO; float4* bakD; double2* bakRange; uint32_t* bakSlot; float4* bakHit; double* bakT;  // the camera-ray queue
  // AO / direct-lighting per-slot state
  float* hitP; float* hitN;          // 3 x cap each
  uint32_t* aoScramble;              // 2 x cap
  int32_t* nClear;
  float* Ld;                         // 3 x cap: per-light accumulator of UniformSampleAllLights
  // queues: extension rays (double-buffered), shadow rays, MIS rays
  float4* extO[2]; float4* extD[2]; double2* extRange[2]; uint32_t* extSlot[2];
  float4* extHit; double* extT;
  float4* shO; float4* shD; double2* shRange; uint8_t* shOcc;
  float4* misO; float4* misD; double2* misRange; float4* misHit; double* misT;
  uint32_t* counts;  // [0],[1] = extension queue sizes, [2] = shadow, [3] = MIS, [4] = hit list size
  uint32_t* hitList; // slots whose camera ray hit (AO / direct lighting)
  // BxDF-list scenes: the extension queue is shaded in material order (a warp then runs one BxDF list instead of up to 32):
  // shadeOrder[i] = queue index, built per bounce by a counting sort over matHist (nMaterials + 1 bins, misses last)
  uint32_t* shadeOrder;
  uint32_t* matHist;
  // participating media: position of the slot's transmittance stream; T and Lvi of the camera ray (sampler_renderer.dart:93-97)
  uint32_t* trCtr;
  float* volT; float* volL;  // 3 x cap each
  float* volScratch;         // single scattering: 4 x volMaxSteps x cap (lightNum, lightComp, lightPos x 2 per step), [value][slot]
  uint32_t volMaxSteps;
  // per slot, textured scenes only: the BSDF the texture pass built for the slot's current vertex (<= 8 lobes, bsdf.dart:253;
  // frame = dgShading.nn, normalize(dgShading.dpdu) after bump mapping); hitCount < 0: the material has no program
  GLobe* hitLobes;
  int32_t* hitCount;
  float* hitFrame;  // 6 x cap
  // scenes with TransformedPrimitives only (null otherwise): Ray.time of every ray the slot's camera sample spawns (camera_sample.dart:33,
  // ray.dart:59; the traversal reads it through the slot id the rays carry in the bits of rayO.w), and the instance each closest hit of
  // the extension / MIS queue came through (-1: a top-level primitive)
  double* slotTime;
  int32_t* extInst; int32_t* misInst; int32_t* bakInst;
  int32_t* camPrim;  // adaptive sampler: primitive the slot's camera ray hit (-1: none)
  uint8_t* adaptFlag;  // adaptive sampler, per pixel of the batch: 1 = supersample (the first visit's samples are dropped)
};

struct RenderCounters {  // mirrors the reference's ray counters (stats.dart:541-555)
  unsigned long long cameraSamples, closestRays, shadowRays, zeroedSamples;  // cameraSamples: halton only (accepted samples)
};

}  // namespace drt
