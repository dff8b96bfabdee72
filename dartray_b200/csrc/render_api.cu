// C ABI, render half (include/drt.h): what _SamplerRendererTask.run does for one task
// (lib/renderers/sampler_renderer.dart:118-218), as a wavefront pipeline over batches of camera samples.
// Host orchestration only; every stage is a kernel of render_kernels.cu / trace_fast.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "drt_ctx.h"

#include <thread>
#include "dart_random.h"
#include "env_map.h"
#include "render_kernels.h"
#include "texture_kernels.h"
#include "shade_device.cuh"

namespace {

struct HostLight {
  int kind = 0;
  float L[3] = {0, 0, 0}, pos[3] = {0, 0, 0};
  int nSamples = 1;
  std::vector<uint32_t> shapes;
  float w2l[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // spot lights (drt_set_spot_params)
  double cosTotalWidth = 0, cosFalloffStart = 0;
  bool haveSpot = false;
  // infinite lights (drt_set_infinite_light): rows of lightToWorld's upper 3x3 (w2l holds worldToLight's) and the radiance map
  float l2w[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  int mapW = 0, mapH = 0;
  std::vector<float> texels;
  // projection / goniometric lights (drt_set_light_map)
  bool haveMapParams = false;
  float proj[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  double screen[4] = {-1, 1, -1, 1}, hither = 1.0e-3;
};

struct ByteArena {  // one cudaMalloc per wavefront, carved into aligned arrays
  char* base = nullptr;
  size_t size = 0, used = 0;
  template <class T>
  T* take(size_t n) {
    used = (used + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + used) : nullptr;
    used += n * sizeof(T);
    return p;
  }
};

}  // namespace

// Stage launchers: the `extra` build of render_kernels.cu when the scene carries mesh attributes or the rarer quadrics
// (RenderScene::extra), the `plain` build otherwise (render_kernels.h).
#define STAGE(fn) (r->rs.extra ? drt::extra::fn : drt::plain::fn)

struct RenderState {
  // host-side description
  std::vector<GMaterial> materials{GMaterial{{0.5f, 0.5f, 0.5f}, 0.f}};
  // general materials (drt_set_material_lobes): BxDF lists; empty when every material is matte
  bool general = false, hasSpecular = false, hasBlend = false;  // hasBlend: a FresnelBlend lobe (kind 5) -> the `extra` kernels
  std::vector<uint2> matLobes;
  std::vector<GLobe> lobes;
  std::vector<HostLight> lights;
  // participating media (drt_set_volumes / drt_set_volume_integrator)
  std::vector<GVolume> volumes;
  std::vector<double> volDensity;
  int volIntegrator = 0;
  double volStep = 1.0;
  // drt_set_shading_precision: DRT_PRECISION_F32 runs the path integrator's vertex / resolve kernels from the float32 build
  // (render_kernels_f32.cu / render_kernels_f32x.cu) on scenes without media, texture programs or instances
  int shadingPrecision = 0;
  bool f32Trace = false;  // set while the path integrator's queues of a float32 render are traced: they run the float32 builds of the
                          // traversal kernels (trace_fast_f32.cu, trace_q_f32.cu); env DRT_F32_TRACE=0 keeps the binary64 traversal (A/B runs)
  std::vector<float> volV2W;     // n x 16: volumeToWorld, for the regions' world bound
  uint32_t volMaxSteps = 0;      // single scattering: bound on a camera ray's march steps (regions' bound diagonal / stepsize)
  DevBuf<GVolume> dVolumes;
  DevBuf<double> dVolDensity;
  // textures that read the hit point, and the materials built from them (drt_set_textures / drt_set_material_programs)
  std::vector<GTex> textures;
  std::vector<float> texData;
  std::vector<GProgram> programs;
  bool programsMaySpecular = false;
  // an animated camera (drt_set_camera_motion): Camera.cameraToWorld as an AnimatedTransform
  bool cameraMoves = false;
  GInstance cameraMotion{};
  DevBuf<GInstance> dCameraMotion;
  // MeasuredMaterial tables (drt_set_measured): descriptors with data == offset into measuredData until they are uploaded
  std::vector<GMeasured> measured;
  std::vector<uint64_t> measuredOffsets;
  std::vector<float> measuredData;
  bool hasMeasured = false;
  DevBuf<GTex> dTextures;
  DevBuf<float> dTexData;
  DevBuf<GProgram> dPrograms;
  bool haveCamera = false, haveFilm = false;
  RenderParams rp{};
  double crop[4] = {0, 1, 0, 1};
  float table[256];
  int spp = 4, pixelOrder = 1, tileSize = 32;
  uint64_t batchSlots = 0;  // 0 = default
  // device-side scene tables
  DevBuf<uint32_t> dPrimToRec, dPrimAttr;
  DevBuf<GLightShape> dLightShapes;
  DevBuf<GMaterial> dMaterials;
  DevBuf<uint2> dMatLobes;
  DevBuf<GLobe> dLobes;
  DevBuf<GMeasured> dMeasured;
  DevBuf<float> dMeasuredData;
  DevBuf<GLight> dLights;
  DevBuf<float> dLightCdf, dTable;
  DevBuf<uint32_t> dMeshOfTri, dTriIdx;
  DevBuf<GMesh> dMeshes;
  DevBuf<float> dVertN, dVertS, dVertUV, dEnv;
  DevBuf<uint32_t> dAdaptList, dAdaptCount;  // adaptive sampler: pixels to supersample
  std::vector<double> sampleTable;           // bestcandidate sampler: the 4096 x 5 pattern (drt_set_sample_table)
  DevBuf<double> dBcTable, dBcShifts;
  DevBuf<DirectOffsets> dDirect;
  DevBuf<SampleArray> dArrays;
  DevBuf<double> dFilm;
  DevBuf<RenderCounters> dCounters;
  DevBuf<float> dRgb, dXyz, dWeight;
  size_t filmPixels = 0;
  // drt_set_render_profiling: one event after every launch (class of the launch it follows), spans summed after the render
  int profFlags = 0;
  std::vector<cudaEvent_t> profEv;
  std::vector<int> profCls;
  size_t profUsed = 0;
  drt_render_profile prof{};
  DevBuf<DeviceCounters> dWork;  // [0] closest, [1] any
  bool sceneTablesValid = false;
  uint64_t buildSerial = 0;
  // sample layout
  std::vector<SampleArray> arrays;
  std::vector<DirectOffsets> direct;
  int maxVals = 0, maxOthers = 0;
  // wavefront storage
  char* wfMem = nullptr;
  size_t wfBytes = 0;
  Wavefront wf{};
  uint32_t shCap = 0;
  RenderScene rs{};
  drt_render_stats stats{};
};

static RenderState* state(drt_ctx* c) {
  if (!c->render) {
    c->render = new RenderState();
    std::memset(c->render->table, 0, sizeof(c->render->table));
  }
  return c->render;
}

void drtRenderStateDestroy(drt_ctx* c) {
  RenderState* r = c->render;
  if (!r) return;
  r->dPrimToRec.release(); r->dPrimAttr.release(); r->dLightShapes.release(); r->dMaterials.release(); r->dLights.release();
  r->dLightCdf.release(); r->dTable.release(); r->dDirect.release(); r->dArrays.release();
  r->dMeshOfTri.release(); r->dTriIdx.release(); r->dMeshes.release(); r->dVertN.release(); r->dVertS.release(); r->dVertUV.release(); r->dEnv.release(); r->dAdaptList.release(); r->dAdaptCount.release(); r->dBcTable.release(); r->dBcShifts.release();
  r->dFilm.release(); r->dCounters.release(); r->dRgb.release(); r->dXyz.release(); r->dWeight.release(); r->dWork.release();
  r->dVolumes.release(); r->dVolDensity.release();
  r->dTextures.release(); r->dTexData.release(); r->dPrograms.release();
  for (cudaEvent_t e : r->profEv) cudaEventDestroy(e);
  if (r->wfMem) cudaFree(r->wfMem);
  delete r;
  c->render = nullptr;
}

static inline int roundUpPow2(int v) {  // common.dart:117-125
  v--;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v + 1;
}

// ImageFilm constructor (image_film.dart:51-97)
static void configureFilm(RenderState* r) {
  RenderParams& p = r->rp;
  p.left = (int)std::ceil(p.xres * r->crop[0]);
  p.width = std::max(1, (int)std::ceil(p.xres * r->crop[1]) - p.left);
  p.top = (int)std::ceil(p.yres * r->crop[2]);
  p.height = std::max(1, (int)std::ceil(p.yres * r->crop[3]) - p.top);
  p.invXWidth = 1.0 / p.xWidth;
  p.invYWidth = 1.0 / p.yWidth;
}

// Sample layout: what the integrators request (path_integrator.dart:124-131,
// direct_lighting_integrator.dart:70-96) + the default volume integrator's two 1D samples
// (emission_integrator.dart:26-29), in the reference's order: 1D arrays first, then 2D.
static void buildLayout(RenderState* r) {
  RenderParams& p = r->rp;
  std::vector<int> n1D, n2D;
  struct Off { int comp1D, pos2D; };
  auto offsets = [&](int n) { Off o; n1D.push_back(n); o.comp1D = (int)n1D.size() - 1; n2D.push_back(n); o.pos2D = (int)n2D.size() - 1; return o; };
  Off pl[3], pb[3], pp[3];
  int pn[3] = {-1, -1, -1};
  std::vector<Off> dl, db;
  std::vector<int> dn;
  int dlNum = -1;
  if (p.integKind == 0) {
    for (int i = 0; i < 3; ++i) {
      pl[i] = offsets(1);
      n1D.push_back(1); pn[i] = (int)n1D.size() - 1;
      pb[i] = offsets(1);
      pp[i] = offsets(1);
    }
  } else if (p.integKind == 2) {
    if (p.strategy == 0) {
      for (const HostLight& l : r->lights) {
        int n = l.nSamples;
        if (p.samplerKind == 0 || p.samplerKind >= 4) n = roundUpPow2(n);  // LowDiscrepancy / Adaptive / BestCandidate Sampler.roundSize
        dn.push_back(n);
        dl.push_back(offsets(n));
        db.push_back(offsets(n));
      }
    } else {
      dn.push_back(1);
      dl.push_back(offsets(1));
      n1D.push_back(1); dlNum = (int)n1D.size() - 1;
      db.push_back(offsets(1));
    }
  }
  n1D.push_back(1);  // volume integrator: tau sample, scatter sample (emission_integrator.dart:26-29)
  n1D.push_back(1);
  std::vector<int> v1(n1D.size()), v2(n2D.size());
  int v = 0;
  for (size_t i = 0; i < n1D.size(); ++i) { v1[i] = v; v += n1D[i]; }
  for (size_t i = 0; i < n2D.size(); ++i) { v2[i] = v; v += 2 * n2D[i]; }
  p.nVals = v;
  for (int i = 0; i < 3; ++i) {
    if (p.integKind == 0) {
      p.pLightComp[i] = v1[pl[i].comp1D]; p.pLightPos[i] = v2[pl[i].pos2D]; p.pLightNum[i] = v1[pn[i]];
      p.pBsdfComp[i] = v1[pb[i].comp1D]; p.pBsdfPos[i] = v2[pb[i].pos2D];
      p.pPathComp[i] = v1[pp[i].comp1D]; p.pPathPos[i] = v2[pp[i].pos2D];
    } else {
      p.pLightComp[i] = p.pLightPos[i] = p.pLightNum[i] = p.pBsdfComp[i] = p.pBsdfPos[i] = p.pPathComp[i] = p.pPathPos[i] = 0;
    }
  }
  p.dlLightNum = dlNum >= 0 ? v1[dlNum] : 0;
  p.pTauSample = v1[n1D.size() - 2];
  p.pScatterSample = v1[n1D.size() - 1];
  r->direct.clear();
  for (size_t i = 0; i < dl.size(); ++i)
    r->direct.push_back(DirectOffsets{dn[i], v1[dl[i].comp1D], v2[dl[i].pos2D], v1[db[i].comp1D], v2[db[i].pos2D]});
  // generation order of the sampler arrays (low_discrepancy_sampler.dart:64-88 -> montecarlo.dart:407-473)
  r->arrays.clear();
  uint32_t sid = 0;
  r->arrays.push_back(SampleArray{2, 1, -1, sid++});
  r->arrays.push_back(SampleArray{2, 1, -2, sid++});
  r->arrays.push_back(SampleArray{1, 1, -3, sid++});
  for (size_t i = 0; i < n1D.size(); ++i) r->arrays.push_back(SampleArray{1, n1D[i], v1[i], sid++});
  for (size_t i = 0; i < n2D.size(); ++i) r->arrays.push_back(SampleArray{2, n2D[i], v2[i], sid++});
  r->maxVals = r->maxOthers = 0;
  p.ldAllSingle = 1;
  for (const SampleArray& a : r->arrays) {
    if (a.nSamples != 1) p.ldAllSingle = 0;
    r->maxVals = std::max(r->maxVals, a.dims * a.nSamples * p.nPixelSamples);
    r->maxOthers = std::max(r->maxOthers, a.nSamples * p.nPixelSamples + p.nPixelSamples);
  }
}

// Scene tables the shading kernels read: primitive -> leaf record, per-primitive attributes, materials,
// lights with their ShapeSet areas and area distribution (shape_set.dart:43-50, montecarlo.dart:26-48).
static int uploadSceneTables(drt_ctx* c, RenderState* r) {
  const uint32_t nt = c->ntris(), np = c->nprims();
  std::vector<uint32_t> primToRec(std::max<uint32_t>(np, 1), 0), attr(std::max<uint32_t>(np, 1), 0);
  const std::vector<uint32_t>& recIds = (c->owner ? c->owner : c)->recPrimIds;  // the top level's records, then the objects'; instance stand-ins skipped
  for (size_t i = 0; i < recIds.size(); ++i)
    if (recIds[i] < np) primToRec[recIds[i]] = (uint32_t)i;
  const int nMat = (int)r->materials.size(), nLights = (int)r->lights.size();
  for (uint32_t i = 0; i < np; ++i) {
    int m = i < nt ? c->matOf[i] : c->sphMat[i - nt], l = i < nt ? c->lightOf[i] : c->sphLight[i - nt];
    int rev = i < nt ? c->revOf[i] : c->sphRev[i - nt];
    if (m < 0 || m >= nMat) return fail(c, DRT_E_INVALID, "a primitive refers to a material index that drt_set_materials did not define");
    if (l >= nLights || l < -1) return fail(c, DRT_E_INVALID, "a primitive refers to a light index that drt_set_lights did not define");
    if (m > 0xffff || l + 1 > 0x7fff) return fail(c, DRT_E_INVALID, "too many materials (65535) or lights (32766)");
    attr[i] = (uint32_t)m | ((uint32_t)(l + 1) << 16) | (rev ? 0x80000000u : 0u);
  }
  std::vector<GLight> gl(std::max(nLights, 1));
  std::vector<float> env;  // radiance maps + sampling tables of the infinite lights
  int nInfinite = 0, nMapped = 0;
  std::vector<GLightShape> shapes;
  std::vector<float> cdf;
  auto pushShape = [&](uint32_t sh, double area) {
    GLightShape ls;
    std::memset(&ls, 0, sizeof(ls));
    ls.prim = sh;
    ls.area = area;
    if (sh < nt) {
      TriVerts t;
      const float* p1 = &c->P[3 * (size_t)c->idx[3 * (size_t)sh]];
      const float* p2 = &c->P[3 * (size_t)c->idx[3 * (size_t)sh + 1]];
      const float* p3 = &c->P[3 * (size_t)c->idx[3 * (size_t)sh + 2]];
      t.p1 = V3{p1[0], p1[1], p1[2]}; t.p2 = V3{p2[0], p2[1], p2[2]}; t.p3 = V3{p3[0], p3[1], p3[2]};
      std::memcpy(ls.p1, p1, 12); std::memcpy(ls.p2, p2, 12); std::memcpy(ls.p3, p3, 12);
      const bool rev = c->revOf[sh] != 0;
      V3 dpdu, dpdv;
      double uv[6] = {0.0, 0.0, 1.0, 0.0, 1.0, 1.0};
      if (!c->meshOfTri.empty() && (c->meshFlags[c->meshOfTri[sh]] & 4))  // triangle.dart:246-254: the mesh's own uvs
        for (int k = 0; k < 3; ++k) {
          uv[2 * k] = c->vertUV[2 * (size_t)c->idx[3 * (size_t)sh + k]];
          uv[2 * k + 1] = c->vertUV[2 * (size_t)c->idx[3 * (size_t)sh + k] + 1];
        }
      triPartialsUV(t, uv, &dpdu, &dpdv);
      V3 nn = shapeNormal(dpdu, dpdv, rev);  // the dg.nn Triangle.intersect leaves behind
      V3 ns = Normalize(Cross(t.p2 - t.p1, t.p3 - t.p1));  // triangle.dart:374-381
      if (rev) ns = mkv((double)ns.x * -1.0, (double)ns.y * -1.0, (double)ns.z * -1.0);
      ls.nn[0] = nn.x; ls.nn[1] = nn.y; ls.nn[2] = nn.z;
      ls.ns[0] = ns.x; ls.ns[1] = ns.y; ls.ns[2] = ns.z;
    }
    shapes.push_back(ls);
  };
  for (int i = 0; i < nLights; ++i) {
    const HostLight& hl = r->lights[i];
    GLight& g = gl[i];
    g.kind = hl.kind;
    if (hl.kind == 3 && !hl.haveSpot) return fail(c, DRT_E_STATE, "a spot light (kind 3) needs drt_set_spot_params after drt_set_lights");
    std::memcpy(g.l2w, hl.l2w, sizeof(g.l2w));
    g.mapW = g.mapH = 0;
    g.envOffset = 0;
    if (hl.kind == 4) {
      if (hl.mapW == 0) return fail(c, DRT_E_STATE, "an infinite light (kind 4) needs drt_set_infinite_light after drt_set_lights");
      g.mapW = hl.mapW;
      g.mapH = hl.mapH;
      g.envOffset = (uint32_t)env.size();
      appendEnvTables(hl.mapW, hl.mapH, hl.texels.data(), hl.L, &env);
      ++nInfinite;
    }
    std::memcpy(g.proj, hl.proj, sizeof(g.proj));
    std::memcpy(g.screen, hl.screen, sizeof(g.screen));
    g.hither = hl.hither;
    if (hl.kind >= 5) {
      if (!hl.haveMapParams) return fail(c, DRT_E_STATE, "a projection / goniometric light (kind 5 / 6) needs drt_set_light_map after drt_set_lights");
      g.mapW = hl.mapW;
      g.mapH = hl.mapH;
      g.envOffset = (uint32_t)env.size();
      env.insert(env.end(), hl.texels.begin(), hl.texels.end());
      ++nMapped;
    }
    std::memcpy(g.w2l, hl.w2l, sizeof(g.w2l));
    g.cosTotalWidth = hl.cosTotalWidth;
    g.cosFalloffStart = hl.cosFalloffStart;
    std::memcpy(g.L, hl.L, 12);
    std::memcpy(g.pos, hl.pos, 12);
    g.nSamples = hl.nSamples;
    g.shapeOffset = (uint32_t)shapes.size();
    g.nShapes = (uint32_t)hl.shapes.size();
    g.cdfOffset = (uint32_t)cdf.size();
    g.area = 0.0;
    if (hl.kind == 0 && hl.shapes.empty()) return fail(c, DRT_E_INVALID, "an area light has no shapes");
    std::vector<double> a;
    for (uint32_t sh : hl.shapes) {
      if (sh >= np) return fail(c, DRT_E_INVALID, "light shape id out of range");
      // the MIS test `lightIsect.primitive.getAreaLight() == light` (integrator.dart:170) reads the primitive's own light index:
      // a shape listed under light i has to carry i, or the BSDF-sampled half of the estimate silently never matches
      if ((sh < nt ? c->lightOf[sh] : c->sphLight[sh - nt]) != i)
        return fail(c, DRT_E_INVALID, "a light's shape list names a primitive whose light index is another light (or none)");
      double area;
      if (sh < nt) {  // triangle.dart:265-269
        TriVerts t;
        const float* p1 = &c->P[3 * (size_t)c->idx[3 * (size_t)sh]];
        const float* p2 = &c->P[3 * (size_t)c->idx[3 * (size_t)sh + 1]];
        const float* p3 = &c->P[3 * (size_t)c->idx[3 * (size_t)sh + 2]];
        t.p1 = V3{p1[0], p1[1], p1[2]}; t.p2 = V3{p2[0], p2[1], p2[2]}; t.p3 = V3{p3[0], p3[1], p3[2]};
        area = triArea(t);
      } else {  // sphere.dart:243-245 with the constructor's clamps (sphere.dart:24-32)
        const HostSphere& s = c->spheres[sh - nt];
        if (s.shape == DRT_QUADRIC_CYLINDER) {  // cylinder.dart:226-228 with the constructor's min / max / clamp (:24-31)
          double zmin = std::fmin(s.prm[1], s.prm[2]), zmax = std::fmax(s.prm[1], s.prm[2]);
          double phiMaxC = (DRT_PI / 180.0) * clampD(s.prm[3], 0.0, 360.0);
          area = (zmax - zmin) * phiMaxC * s.prm[0];
          a.push_back(area);
          g.area += area;
          pushShape(sh, area);
          continue;
        }
        if (s.shape > DRT_QUADRIC_CYLINDER)  // shape.dart:83-86: Shape.sample is unimplemented for cone / paraboloid / hyperboloid
          return fail(c, DRT_E_INVALID, "cone / paraboloid / hyperboloid shapes cannot be area lights (no Shape.sample in the reference)");
        if (s.shape == 1) {  // disk.dart:142-145
          double phiMaxD = (DRT_PI / 180.0) * clampD(s.phiMaxDeg, 0.0, 360.0);
          area = phiMaxD * 0.5 * (s.radius * s.radius - s.innerRadius * s.innerRadius);
          a.push_back(area);
          g.area += area;
          pushShape(sh, area);
          continue;
        }
        double zmin = clampD(std::fmin(s.zmin, s.zmax), -s.radius, s.radius), zmax = clampD(std::fmax(s.zmin, s.zmax), -s.radius, s.radius);
        double phiMax = (DRT_PI / 180.0) * clampD(s.phiMaxDeg, 0.0, 360.0);
        area = phiMax * s.radius * (zmax - zmin);
      }
      a.push_back(area);
      g.area += area;
      pushShape(sh, area);
    }
    if (hl.kind == 0) {  // Distribution1D(areas): float32 func and cdf (montecarlo.dart:26-48)
      const int count = (int)a.size();
      std::vector<float> func(count), cd(count + 1, 0.f);
      for (int k = 0; k < count; ++k) func[k] = (float)a[k];
      for (int k = 1; k < count + 1; ++k) cd[k] = (float)((double)cd[k - 1] + (double)func[k - 1] / count);
      double funcInt = cd[count];
      for (int k = 1; k < count + 1; ++k) cd[k] = funcInt == 0.0 ? (float)((double)k / count) : (float)((double)cd[k] / funcInt);
      cdf.insert(cdf.end(), cd.begin(), cd.end());
    }
  }
  CK(c, r->dPrimToRec.ensure(primToRec.size()));
  CK(c, r->dPrimAttr.ensure(attr.size()));
  CK(c, r->dMaterials.ensure(r->materials.size()));
  CK(c, r->dLights.ensure(gl.size()));
  CK(c, r->dLightShapes.ensure(std::max<size_t>(1, shapes.size())));
  CK(c, r->dLightCdf.ensure(std::max<size_t>(1, cdf.size())));
  CK(c, cudaMemcpy(r->dPrimToRec.p, primToRec.data(), primToRec.size() * 4, cudaMemcpyHostToDevice));
  CK(c, cudaMemcpy(r->dPrimAttr.p, attr.data(), attr.size() * 4, cudaMemcpyHostToDevice));
  CK(c, cudaMemcpy(r->dMaterials.p, r->materials.data(), r->materials.size() * sizeof(GMaterial), cudaMemcpyHostToDevice));
  CK(c, r->dMatLobes.ensure(std::max<size_t>(1, r->matLobes.size())));
  CK(c, r->dLobes.ensure(std::max<size_t>(1, r->lobes.size())));
  if (!r->matLobes.empty()) CK(c, cudaMemcpy(r->dMatLobes.p, r->matLobes.data(), r->matLobes.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  // MeasuredMaterial tables; lobes of kind 6 / 7 get their table's device address and dimensions written into them (render_types.h)
  std::vector<GMeasured> gmeas(r->measured);
  if (!gmeas.empty()) {
    CK(c, r->dMeasuredData.ensure(std::max<size_t>(1, r->measuredData.size())));
    CK(c, cudaMemcpy(r->dMeasuredData.p, r->measuredData.data(), r->measuredData.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (size_t i = 0; i < gmeas.size(); ++i) gmeas[i].data = r->dMeasuredData.p + r->measuredOffsets[i];
    CK(c, r->dMeasured.ensure(gmeas.size()));
    CK(c, cudaMemcpy(r->dMeasured.p, gmeas.data(), gmeas.size() * sizeof(GMeasured), cudaMemcpyHostToDevice));
  }
  if (!r->lobes.empty()) {
    std::vector<GLobe> ls(r->lobes);
    for (GLobe& l : ls) {
      if (l.kind != 6 && l.kind != 7) continue;
      const double ti = l.param;
      if (!(ti >= 0.0) || ti >= (double)gmeas.size() || ti != std::floor(ti))
        return fail(c, DRT_E_INVALID, "a measured BxDF (lobe kind 6 / 7) names a table drt_set_measured did not define");
      const GMeasured& tb = gmeas[(size_t)ti];
      if (tb.kind != l.kind - 6) return fail(c, DRT_E_INVALID, "lobe kind 6 needs a regular-halfangle table, kind 7 an irregular-isotropic one");
      const long long bits = (long long)(uintptr_t)tb.data;
      std::memcpy(&l.et, &bits, 8);
      for (int k = 0; k < 3; ++k) std::memcpy(&l.k[k], &tb.dims[k], 4);
    }
    CK(c, cudaMemcpy(r->dLobes.p, ls.data(), ls.size() * sizeof(GLobe), cudaMemcpyHostToDevice));
  }
  CK(c, cudaMemcpy(r->dLights.p, gl.data(), gl.size() * sizeof(GLight), cudaMemcpyHostToDevice));
  if (!shapes.empty()) CK(c, cudaMemcpy(r->dLightShapes.p, shapes.data(), shapes.size() * sizeof(GLightShape), cudaMemcpyHostToDevice));
  if (!cdf.empty()) CK(c, cudaMemcpy(r->dLightCdf.p, cdf.data(), cdf.size() * 4, cudaMemcpyHostToDevice));
  RenderScene& rs = r->rs;
  rs.meshOfTri = nullptr; rs.triIdx = nullptr; rs.meshes = nullptr; rs.vertN = rs.vertS = rs.vertUV = nullptr;
  if (!c->meshOfTri.empty() && nt > 0) {  // drt_set_mesh_shading
    const size_t nm = c->meshFlags.size();
    std::vector<GMesh> gm(nm);
    for (size_t m = 0; m < nm; ++m) {
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          gm[m].o2w[3 * a + b] = c->meshO2W[16 * m + 4 * a + b];
          gm[m].w2o[3 * a + b] = c->meshW2O[16 * m + 4 * a + b];
        }
      gm[m].flags = c->meshFlags[m];
      gm[m].pad_ = 0.f;
    }
    CK(c, r->dMeshOfTri.ensure(nt));
    CK(c, r->dTriIdx.ensure(3 * (size_t)nt));
    CK(c, r->dMeshes.ensure(nm));
    CK(c, cudaMemcpy(r->dMeshOfTri.p, c->meshOfTri.data(), (size_t)nt * 4, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(r->dTriIdx.p, c->idx.data(), 3 * (size_t)nt * 4, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(r->dMeshes.p, gm.data(), nm * sizeof(GMesh), cudaMemcpyHostToDevice));
    rs.meshOfTri = r->dMeshOfTri.p; rs.triIdx = r->dTriIdx.p; rs.meshes = r->dMeshes.p;
    if (!c->vertN.empty()) {
      CK(c, r->dVertN.ensure(c->vertN.size()));
      CK(c, cudaMemcpy(r->dVertN.p, c->vertN.data(), c->vertN.size() * 4, cudaMemcpyHostToDevice));
      rs.vertN = r->dVertN.p;
    }
    if (!c->vertS.empty()) {
      CK(c, r->dVertS.ensure(c->vertS.size()));
      CK(c, cudaMemcpy(r->dVertS.p, c->vertS.data(), c->vertS.size() * 4, cudaMemcpyHostToDevice));
      rs.vertS = r->dVertS.p;
    }
    if (!c->vertUV.empty()) {
      CK(c, r->dVertUV.ensure(c->vertUV.size()));
      CK(c, cudaMemcpy(r->dVertUV.p, c->vertUV.data(), c->vertUV.size() * 4, cudaMemcpyHostToDevice));
      rs.vertUV = r->dVertUV.p;
    }
  }
  rs.ts = c->ts;
  rs.envData = nullptr;
  rs.nInfinite = nInfinite;
  if (!env.empty()) {
    CK(c, r->dEnv.ensure(env.size()));
    CK(c, cudaMemcpy(r->dEnv.p, env.data(), env.size() * 4, cudaMemcpyHostToDevice));
    rs.envData = r->dEnv.p;
  }
  rs.volumes = nullptr;
  rs.volDensity = nullptr;
  rs.nVolumes = (int32_t)r->volumes.size();
  rs.volIntegrator = r->volIntegrator;
  rs.volStep = r->volStep;
  if (!r->volumes.empty()) {
    CK(c, r->dVolumes.ensure(r->volumes.size()));
    CK(c, cudaMemcpy(r->dVolumes.p, r->volumes.data(), r->volumes.size() * sizeof(GVolume), cudaMemcpyHostToDevice));
    CK(c, r->dVolDensity.ensure(std::max<size_t>(1, r->volDensity.size())));
    if (!r->volDensity.empty())
      CK(c, cudaMemcpy(r->dVolDensity.p, r->volDensity.data(), r->volDensity.size() * sizeof(double), cudaMemcpyHostToDevice));
    rs.volumes = r->dVolumes.p;
    rs.volDensity = r->dVolDensity.p;
  }
  rs.textures = nullptr;
  rs.texData = nullptr;
  rs.programs = nullptr;
  rs.nPrograms = 0;
  if (!r->programs.empty()) {
    if ((int)r->programs.size() != nMat) return fail(c, DRT_E_INVALID, "drt_set_material_programs: one entry per material of the material table");
    if (!r->general) return fail(c, DRT_E_STATE, "material programs go with drt_set_material_lobes (an empty lobe list for a program's material)");
    // which parameter of which plugin is a spectrum texture (the tex[] order documented in include/drt.h)
    static const int kSlots[12] = {2, 1, 3, 3, 3, 3, 4, 5, 7, 1, 2, 0};
    static const unsigned kSpectrumMask[12] = {0x1, 0x1, 0x3, 0x3, 0x3, 0x3, 0x3, 0xf, 0x2f, 0x1, 0x1, 0x0};
    const int nTex = (int)r->textures.size();
    r->programsMaySpecular = false;
    for (int m = 0; m < nMat; ++m) {
      const GProgram& pr = r->programs[m];
      if (pr.kind < 0) continue;
      if (pr.kind > 11) return fail(c, DRT_E_INVALID, "material program kind out of range");
      if (pr.kind == 11 && (pr.m1 < 0 || pr.m1 >= (int)gmeas.size()))
        return fail(c, DRT_E_INVALID, "measured: m1 must name a table of drt_set_measured");
      for (int k = 0; k < kSlots[pr.kind]; ++k) {
        if (pr.tex[k] < 0 || pr.tex[k] >= nTex) return fail(c, DRT_E_INVALID, "a material program names a texture node drt_set_textures did not define");
        if (r->textures[pr.tex[k]].spectrum != (int)((kSpectrumMask[pr.kind] >> k) & 1u))
          return fail(c, DRT_E_INVALID, "a material program binds a float texture to a spectrum parameter or the reverse");
      }
      if (pr.bump >= nTex || (pr.bump >= 0 && r->textures[pr.bump].spectrum != 0)) return fail(c, DRT_E_INVALID, "bumpmap must be a float texture node");
      if (pr.kind == 9) {
        for (int sub : {pr.m1, pr.m2}) {
          if (sub < 0 || sub >= nMat || sub == m) return fail(c, DRT_E_INVALID, "mix: m1 / m2 must be other materials of the table");
          if (r->programs[sub].kind == 9) return fail(c, DRT_E_UNSUPPORTED, "a mix of mixes (ScaledBxDF of a ScaledBxDF) is not representable");
        }
      }
      // can this program add a specular BxDF at some hit?  (a constant-valued parameter decides it for good)
      auto constantIs = [&](int id, float v) {
        const GTex& t = r->textures[id];
        return t.kind == 0 && (float)t.value[0] == v && (float)t.value[1] == v && (float)t.value[2] == v;
      };
      bool spec = false;
      if (pr.kind == 1 || pr.kind == 10) spec = !constantIs(pr.tex[0], 0.f);                    // mirror, subsurface: Kr
      else if (pr.kind == 2) spec = !constantIs(pr.tex[0], 0.f) || !constantIs(pr.tex[1], 0.f);  // glass: Kr, Kt
      else if (pr.kind == 5) spec = !constantIs(pr.tex[1], 0.f);                                 // shinymetal: Kr
      else if (pr.kind == 8) spec = !constantIs(pr.tex[2], 0.f) || !constantIs(pr.tex[3], 0.f) || !constantIs(pr.tex[5], 1.f);  // uber: Kr, Kt, opacity
      if (spec) r->programsMaySpecular = true;  // mix: its two materials are entries of the same table
    }
    CK(c, r->dTextures.ensure(std::max<size_t>(1, r->textures.size())));
    if (!r->textures.empty()) CK(c, cudaMemcpy(r->dTextures.p, r->textures.data(), r->textures.size() * sizeof(GTex), cudaMemcpyHostToDevice));
    CK(c, r->dTexData.ensure(std::max<size_t>(128, r->texData.size())));
    if (!r->texData.empty()) CK(c, cudaMemcpy(r->dTexData.p, r->texData.data(), r->texData.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(c, r->dPrograms.ensure(r->programs.size()));
    CK(c, cudaMemcpy(r->dPrograms.p, r->programs.data(), r->programs.size() * sizeof(GProgram), cudaMemcpyHostToDevice));
    rs.textures = r->dTextures.p;
    rs.texData = r->dTexData.p;
    rs.programs = r->dPrograms.p;
    rs.nPrograms = nMat;
  }
  rs.measured = gmeas.empty() ? nullptr : r->dMeasured.p;
  rs.nMeasured = (int32_t)gmeas.size();
  rs.extra = (rs.meshOfTri != nullptr || c->ts.quadMode == 2 || c->ts.nInstances > 0 || nInfinite > 0 || nMapped > 0 || (r->general && (r->hasBlend || r->hasMeasured)) ||
              rs.nVolumes > 0 || rs.nPrograms > 0) ? 1 : 0;
  rs.ntris = nt;
  rs.nprims = np;
  rs.primToRec = r->dPrimToRec.p;
  rs.primAttr = r->dPrimAttr.p;
  rs.materials = r->dMaterials.p;
  rs.general = r->general ? 1 : 0;
  rs.nMaterials = nMat;
  rs.matLobes = r->dMatLobes.p;
  rs.lobes = r->dLobes.p;
  rs.lights = r->dLights.p;
  rs.nLights = nLights;
  rs.lightShapes = r->dLightShapes.p;
  rs.lightCdf = r->dLightCdf.p;
  r->sceneTablesValid = true;
  r->buildSerial = c->buildSerial;
  return DRT_OK;
}

static const int kMaxChainLevels = 16;  // specular recursion depth the chain evaluation covers (maxdepth <= 17)

static void carve(ByteArena& a, Wavefront& wf, uint32_t cap, uint32_t shCap, int nVals, bool chains, bool volumes, uint32_t volMaxSteps,
                  bool programs, bool instances) {
  wf.cap = cap;
  wf.slotTime = nullptr; wf.extInst = wf.misInst = wf.bakInst = nullptr;
  if (instances) {
    wf.slotTime = a.take<double>(cap);
    wf.extInst = a.take<int32_t>(cap); wf.misInst = a.take<int32_t>(cap);
    if (chains) wf.bakInst = a.take<int32_t>(cap);
  }
  if (programs) {  // the BSDF the texture pass builds per slot: 8 lobes (bsdf.dart:253) + the shading frame
    wf.hitLobes = a.take<GLobe>(8 * (size_t)cap);
    wf.hitCount = a.take<int32_t>(cap);
    wf.hitFrame = a.take<float>(6 * (size_t)cap);
  } else {
    wf.hitLobes = nullptr; wf.hitCount = nullptr; wf.hitFrame = nullptr;
  }
  wf.volScratch = nullptr;
  wf.volMaxSteps = volMaxSteps;
  if (volumes) {
    wf.trCtr = a.take<uint32_t>(cap);
    wf.volT = a.take<float>(3 * (size_t)cap); wf.volL = a.take<float>(3 * (size_t)cap);
    if (volMaxSteps) wf.volScratch = a.take<float>(4 * (size_t)volMaxSteps * cap);
  } else {
    wf.trCtr = nullptr;
    wf.volT = wf.volL = nullptr;
  }
  wf.pixX = a.take<int32_t>(cap); wf.pixY = a.take<int32_t>(cap); wf.sampleIdx = a.take<uint32_t>(cap);
  wf.camXY = a.take<double2>(cap); wf.camLens = a.take<double2>(cap); wf.camTime = a.take<double>(cap);
  wf.vals = a.take<float>((size_t)std::max(nVals, 1) * cap);
  wf.L = a.take<float>(3 * (size_t)cap); wf.T = a.take<float>(3 * (size_t)cap);
  wf.pendSh = a.take<float>(3 * (size_t)cap); wf.pendMisF = a.take<float>(3 * (size_t)cap); wf.pendMisScale = a.take<double>(cap);
  wf.pendT = a.take<float>(3 * (size_t)cap);
  wf.shIdx = a.take<int32_t>(cap); wf.misIdx = a.take<int32_t>(cap); wf.misLight = a.take<int32_t>(cap);
  wf.specBounce = a.take<uint8_t>(cap);
  wf.hitP = a.take<float>(3 * (size_t)cap); wf.hitN = a.take<float>(3 * (size_t)cap);
  wf.aoScramble = a.take<uint32_t>(2 * (size_t)cap); wf.nClear = a.take<int32_t>(cap);
  wf.Ld = a.take<float>(3 * (size_t)cap);
  for (int k = 0; k < 2; ++k) {
    wf.extO[k] = a.take<float4>(cap); wf.extD[k] = a.take<float4>(cap); wf.extRange[k] = a.take<double2>(cap);
    wf.extSlot[k] = a.take<uint32_t>(cap);
  }
  wf.extHit = a.take<float4>(cap); wf.extT = a.take<double>(cap);
  wf.shO = a.take<float4>(shCap); wf.shD = a.take<float4>(shCap); wf.shRange = a.take<double2>(shCap); wf.shOcc = a.take<uint8_t>(shCap);
  wf.misO = a.take<float4>(cap); wf.misD = a.take<float4>(cap); wf.misRange = a.take<double2>(cap);
  wf.misHit = a.take<float4>(cap); wf.misT = a.take<double>(cap);
  wf.counts = a.take<uint32_t>(Q_COUNT);
  wf.hitList = a.take<uint32_t>(cap);
  wf.shadeOrder = a.take<uint32_t>(cap);
  wf.matHist = a.take<uint32_t>(1024);
  wf.camPrim = a.take<int32_t>(cap);
  wf.adaptFlag = a.take<uint8_t>(cap);
  if (chains) {  // directlighting with specular BxDFs: counters per recursion level + a copy of the camera-ray queue
    wf.specCtr = a.take<uint32_t>(cap);
    wf.specCtrAt = a.take<uint32_t>((size_t)kMaxChainLevels * cap);
    wf.bakO = a.take<float4>(cap); wf.bakD = a.take<float4>(cap); wf.bakRange = a.take<double2>(cap);
    wf.bakSlot = a.take<uint32_t>(cap); wf.bakHit = a.take<float4>(cap); wf.bakT = a.take<double>(cap);
  } else {
    wf.specCtr = wf.specCtrAt = nullptr;
    wf.bakO = wf.bakD = nullptr; wf.bakRange = nullptr; wf.bakSlot = nullptr; wf.bakHit = nullptr; wf.bakT = nullptr;
  }
}

static int ensureWavefront(drt_ctx* c, RenderState* r, uint32_t cap, uint32_t shCap) {
  // the whitted integrator draws its light samples from the per-slot stream counter even without specular BxDFs
  const bool chains = (r->rp.integKind == 2 && r->hasSpecular && r->rp.maxDepth > 1) || r->rp.integKind == 3;
  ByteArena probe;
  Wavefront tmp{};
  const bool volumes = !r->volumes.empty();
  const uint32_t volMaxSteps = (volumes && r->volIntegrator == 1) ? r->volMaxSteps : 0;
  const bool programs = !r->programs.empty();
  const bool instances = c->ts.nInstances > 0;
  carve(probe, tmp, cap, shCap, r->rp.nVals, chains, volumes, volMaxSteps, programs, instances);
  size_t need = probe.used + 256;
  if (need > r->wfBytes) {
    if (r->wfMem) cudaFree(r->wfMem);
    r->wfMem = nullptr;
    r->wfBytes = 0;
    CK(c, cudaMalloc((void**)&r->wfMem, need));
    r->wfBytes = need;
  }
  ByteArena a;
  a.base = r->wfMem;
  a.size = r->wfBytes;
  carve(a, r->wf, cap, shCap, r->rp.nVals, chains, volumes, volMaxSteps, programs, instances);
  r->shCap = shCap;
  return DRT_OK;
}

static int ensureFilm(drt_ctx* c, RenderState* r) {
  const size_t n = (size_t)r->rp.width * r->rp.height;
  if (n != r->filmPixels || !r->dFilm.p) {
    CK(c, r->dFilm.ensure(4 * n));
    CK(c, cudaMemset(r->dFilm.p, 0, 4 * n * sizeof(double)));
    r->filmPixels = n;
  }
  CK(c, r->dTable.ensure(256));
  CK(c, cudaMemcpy(r->dTable.p, r->table, sizeof(r->table), cudaMemcpyHostToDevice));
  if (!r->dCounters.p) {  // cudaMalloc does not clear: a fresh context must not inherit a freed buffer's bytes as ray counts
    CK(c, r->dCounters.ensure(1));
    CK(c, cudaMemset(r->dCounters.p, 0, sizeof(RenderCounters)));
  }
  r->rp.film = r->dFilm.p;
  r->rp.filterTable = r->dTable.p;
  return DRT_OK;
}

// GetSubWindow, common.dart:52-73
static void getSubWindow(int w, int h, int num, int count, int e[4]) {
  int nx = count, ny = 1;
  while ((nx & 0x1) == 0 && 2 * w * ny < h * nx) { nx >>= 1; ny <<= 1; }
  int xo = num % nx, yo = num / nx;
  double tx0 = (double)xo / nx, tx1 = (double)(xo + 1) / nx, ty0 = (double)yo / ny, ty1 = (double)(yo + 1) / ny;
  auto lerp = [](double t, double a, double b) { return (1.0 - t) * a + t * b; };
  e[0] = (int)std::floor(lerp(tx0, 0, w));
  e[1] = std::min((int)std::floor(lerp(tx1, 0, w)), w);
  e[2] = (int)std::floor(lerp(ty0, 0, h));
  e[3] = std::min((int)std::floor(lerp(ty1, 0, h)), h);
}

// AdaptiveSampler's constructor (adaptive_sampler.dart:40-84): min / max swapped into order, rounded up to powers of two, at least
// two initial samples, and more maximum than minimum samples
static void adaptiveCounts(int mins, int maxs, int* mn, int* mx) {
  if (mins > maxs) std::swap(mins, maxs);
  int a = roundUpPow2(mins), b = roundUpPow2(maxs);
  if (a < 2) a = 2;
  if (a == b) b *= 2;
  *mn = a;
  *mx = b;
}

static int prepare(drt_ctx* c, RenderState* r) {
  if (c->device == DRT_DEVICE_NONE) return fail(c, DRT_E_NODEVICE, kNoDevice);
  if (!c->built) return fail(c, DRT_E_STATE, "drt_build_bvh must be called before rendering");
  if (!r->haveCamera || !r->haveFilm) return fail(c, DRT_E_STATE, "drt_set_camera and drt_set_film must be called before rendering");
  CK(c, cudaSetDevice(c->device));
  RenderParams& p = r->rp;
  p.nPixelSamples = p.samplerKind == 0 ? roundUpPow2(r->spp) : (p.samplerKind == 1 ? p.xs * p.ys : ((p.samplerKind == 3 || p.samplerKind == 5) ? 1 : r->spp));
  if (p.samplerKind == 5 && r->sampleTable.size() != 5 * 4096)
    return fail(c, DRT_E_STATE, "the bestcandidate sampler (kind 5) needs drt_set_sample_table");
  if (p.samplerKind == 4) {  // the first visit of every pixel: minSamples (renderWindow switches to maxSamples for the second)
    int mn, mx;
    adaptiveCounts(p.xs, p.ys, &mn, &mx);
    p.nPixelSamples = mn;
    p.adaptiveMethod = p.jitter ? 1 : 0;
  }
  if (p.nPixelSamples < 1) return fail(c, DRT_E_INVALID, "sampler produces no samples per pixel");
  buildLayout(r);
  if (p.integKind == 2 && p.strategy == 0 && r->direct.size() != r->lights.size()) return fail(c, DRT_E_STATE, "light table out of date");
  if (!r->sceneTablesValid || r->buildSerial != c->buildSerial) {
    int rc = uploadSceneTables(c, r);
    if (rc != DRT_OK) return rc;
  }
  r->rs.ts = c->ts;
  p.cameraMotion = nullptr;
  if (r->cameraMoves) {
    CK(c, r->dCameraMotion.ensure(1));
    CK(c, cudaMemcpy(r->dCameraMotion.p, &r->cameraMotion, sizeof(GInstance), cudaMemcpyHostToDevice));
    p.cameraMotion = r->dCameraMotion.p;
  }
  if (c->ts.nInstances > 0) {  // TransformedPrimitives: what the renderer carries for them so far
    std::vector<uint8_t> inObject(c->nprims(), 0);
    for (const drt_ctx::HostObject& ob : c->objects)
      for (uint32_t id : ob.order) inObject[id] = 1;
    for (uint32_t id = 0; id < c->nprims(); ++id) {
      if (!inObject[id]) continue;
      const int32_t light = id < c->ntris() ? (c->lightOf.empty() ? -1 : c->lightOf[id]) : (c->sphLight.empty() ? -1 : c->sphLight[id - c->ntris()]);
      if (light >= 0) return fail(c, DRT_E_UNSUPPORTED, "an area light on an instanced / animated shape (the reference drops it with a warning, dartray.dart:407-410)");
      if (id < c->ntris() && !c->meshOfTri.empty() && (c->meshFlags[c->meshOfTri[id]] & 3u))
        return fail(c, DRT_E_UNSUPPORTED, "per-vertex N / S on a mesh inside an instanced / animated object");
    }
  }
  CK(c, r->dArrays.ensure(r->arrays.size()));
  CK(c, cudaMemcpy(r->dArrays.p, r->arrays.data(), r->arrays.size() * sizeof(SampleArray), cudaMemcpyHostToDevice));
  CK(c, r->dDirect.ensure(std::max<size_t>(1, r->direct.size())));
  if (!r->direct.empty())
    CK(c, cudaMemcpy(r->dDirect.p, r->direct.data(), r->direct.size() * sizeof(DirectOffsets), cudaMemcpyHostToDevice));
  p.direct = r->dDirect.p;
  int rc = ensureFilm(c, r);
  if (rc != DRT_OK) return rc;
  if (p.samplerKind == 0 || p.samplerKind >= 4) {
    size_t maxVals = (size_t)r->maxVals;
    if (p.samplerKind == 4) {  // the second visit of a supersampled pixel uses the maxSamples layout: check THAT one before any work
      int mn, mx;
      adaptiveCounts(p.xs, p.ys, &mn, &mx);
      maxVals = maxVals / (size_t)mn * (size_t)mx;
    }
    size_t smem = 4 * (size_t)(maxVals | 1) * sizeof(float);  // G = 32: four tasks per block
    if (smem > 200 * 1024) return fail(c, DRT_E_INVALID, "lowdiscrepancy sampler: pixelsamples x light nsamples too large for one warp's shared memory");
  }
  if (!r->volumes.empty() && r->volIntegrator == 1) {
    // a camera ray (unit direction: every camera normalises it) spends at most the diagonal of the regions' world bound inside them
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = 0; i < r->volumes.size(); ++i) {
      const GVolume& v = r->volumes[i];
      for (int k = 0; k < 8; ++k) {
        const V3 q = XfPoint(&r->volV2W[16 * i], V3{(k & 1) ? v.hi[0] : v.lo[0], (k & 2) ? v.hi[1] : v.lo[1], (k & 4) ? v.hi[2] : v.lo[2]});
        const double qq[3] = {q.x, q.y, q.z};
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], qq[a]); hi[a] = std::max(hi[a], qq[a]); }
      }
    }
    const double diag = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
    const double steps = std::ceil(diag * 1.001 / r->volStep) + 2.0;
    if (!(steps <= 65536.0)) return fail(c, DRT_E_UNSUPPORTED, "single-scattering volume integrator: more than 65536 steps across the volume regions");
    r->volMaxSteps = (uint32_t)steps;
  }
  if (!r->volumes.empty() && p.integKind >= 2 && r->hasSpecular && p.maxDepth > 1)
    return fail(c, DRT_E_UNSUPPORTED, "participating media with the specular recursion of directlighting / whitted (renderer.Li along every "
                                      "reflected ray, integrator.dart:187-290) is not on the GPU path");
  {  // Sampler.samplesPerPixel as each sampler's constructor hands it to the base class (lib/samplers/*.dart)
    int sppBase = r->spp;
    if (p.samplerKind == 0) sppBase = roundUpPow2(r->spp);
    else if (p.samplerKind == 1) sppBase = p.xs * p.ys;
    else if (p.samplerKind == 4) sppBase = roundUpPow2(std::max(p.xs, p.ys));
    p.diffScale = 1.0 / std::sqrt((double)sppBase);  // sampler_renderer.dart:166
  }
  if (!r->programs.empty() && p.integKind >= 2 && (r->hasSpecular || r->programsMaySpecular) && p.maxDepth > 1)
    return fail(c, DRT_E_UNSUPPORTED, "textured / bump-mapped materials with the specular recursion of directlighting / whitted (ray "
                                      "differentials through SpecularReflect / SpecularTransmit, integrator.dart:203-280) are not on the GPU path");
  if (!r->volumes.empty() && (p.samplerKind == 3 || p.samplerKind == 4 || p.samplerKind == 5))
    return fail(c, DRT_E_UNSUPPORTED, "participating media with the halton / adaptive / bestcandidate samplers is not on the GPU path");
  return DRT_OK;
}

// Records the end of the launch(es) just enqueued; cls = DRT_PK_*, -1 = start of a render (nothing before it is counted).
static void profMark(drt_ctx* c, int cls) {
  RenderState* r = c->render;
  if (!r || !(r->profFlags & DRT_PROFILE_TIME)) return;
  if (r->profUsed == r->profEv.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    r->profEv.push_back(e);
    r->profCls.push_back(0);
  }
  r->profCls[r->profUsed] = cls;
  cudaEventRecord(r->profEv[r->profUsed++], c->stream);
}

static int profCollect(drt_ctx* c) {  // after the render's stream synchronize
  RenderState* r = c->render;
  if (r->profFlags & DRT_PROFILE_TIME) {
    for (size_t i = 1; i < r->profUsed; ++i) {
      if (r->profCls[i] < 0) continue;
      float ms = 0.f;
      CK(c, cudaEventElapsedTime(&ms, r->profEv[i - 1], r->profEv[i]));
      r->prof.ms[r->profCls[i]] += ms;
      r->prof.launches[r->profCls[i]]++;
    }
    r->profUsed = 0;
  }
  if (r->profFlags & DRT_PROFILE_WORK) {
    DeviceCounters h[2];
    CK(c, cudaMemcpy(h, r->dWork.p, sizeof(h), cudaMemcpyDeviceToHost));
    CK(c, cudaMemset(r->dWork.p, 0, sizeof(h)));
    drt_counters* dst[2] = {&r->prof.closest, &r->prof.any};
    for (int k = 0; k < 2; ++k) {
      dst[k]->rays += h[k].rays; dst[k]->nodes_visited += h[k].nodes_visited;
      dst[k]->prims_tested += h[k].prims_tested; dst[k]->hits += h[k].hits;
    }
  }
  return DRT_OK;
}

static int traceQueue(drt_ctx* c, bool any, const float4* o, const float4* d, const double2* range, const uint32_t* nDev, void* out,
                      double* tOut, cudaStream_t st) {
  static const bool exactEnv = std::getenv("DRT_RENDER_EXACT_WALK") != nullptr;  // A/B: the literal walk for every renderer queue
  if (c->ts.nInstances > 0 || exactEnv) {
    // scenes with TransformedPrimitives: the literal walk, which descends into the objects at each ray's time.  One thread per
    // queue slot (the live count is on the device; the surplus threads leave at once).
    const Wavefront& wf = c->render->wf;
    ExactExtras xx;
    xx.times = wf.slotTime;
    xx.timesBySlot = 1;
    xx.tOut = tOut;
    xx.instOut = any ? nullptr : (out == (void*)wf.misHit ? wf.misInst : wf.extInst);
    const uint64_t capQ = (o == wf.shO) ? c->render->shCap : wf.cap;
    CK(c, launchTrace(c->ts, any, false, o, d, capQ, out, nullptr, st, range, nDev, &xx));
    c->launches++;
    profMark(c, any ? DRT_PK_TRACE_ANY : DRT_PK_TRACE_CLOSEST);
    return DRT_OK;
  }
  TraceExtras ex;
  ex.nDev = nDev;
  ex.range = range;
  ex.tOut = tOut;
  ex.noUV = 1;  // the shading stages read the primitive and tHit only
  const bool f32Q = c->render && c->render->f32Trace && c->render->rp.integKind == 0;
  if (f32Q) CK(c, launchTraceFastF32(c->ts, any, o, d, 0, out, c->dNextRay.p, c->numSMs, st, &ex));
  else CK(c, launchTraceFast(c->ts, any, o, d, 0, out, c->dNextRay.p, c->numSMs, st, &ex));
  c->launches++;
  profMark(c, any ? DRT_PK_TRACE_ANY : DRT_PK_TRACE_CLOSEST);
  RenderState* r = c->render;
  if (r && (r->profFlags & DRT_PROFILE_WORK)) {  // the reference's walk over the same queue, counting only
    uint32_t n = 0;
    CK(c, cudaMemcpyAsync(&n, nDev, sizeof(n), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    CK(c, launchTrace(c->ts, any, true, o, d, n, nullptr, r->dWork.p + (any ? 1 : 0), st, range, nDev));
    profMark(c, -1);
  }
  return DRT_OK;
}

#define RK(call)                         \
  do {                                   \
    int rc__ = (call);                   \
    if (rc__ != DRT_OK) return rc__;     \
  } while (0)

// DirectLightingIntegrator.Li without its recursion (direct_lighting_integrator.dart:30-55) on the vertices of extension
// queue `cur`: emitted light, then UniformSampleAllLights / UniformSampleOneLight, one launch group per (light, sample).
// `weighted`: the vertices are the ends of a specular chain; their radiance is added with the chain weight.
static int directStage(drt_ctx* c, RenderState* r, int cur, bool weighted) {
  const RenderParams& p = r->rp;
  const RenderScene& rs = r->rs;
  const Wavefront& wf = r->wf;
  cudaStream_t st = c->stream;
  const int sms = c->numSMs;
  RenderCounters* rc = r->dCounters.p;
  if (rs.nPrograms > 0) {  // camera vertices only: prepare() rejects programs together with the specular recursion
    CK(c, launchTexturePass(p, rs, wf, cur, weighted ? 0 : 1, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
    c->launches++;
  }
  CK(c, STAGE(launchDirectSetup)(p, rs, wf, cur, weighted ? 1 : 0, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
  c->launches++;
  if (rs.nLights <= 0) return DRT_OK;
  const bool one = p.strategy != 0;
  const int nL = one ? 1 : rs.nLights;
  // BxDF-list scenes: one material sort of the queue serves every (light, sample) launch below (64 -> 46.5 ms on cornell_materials,
  // 1080p x 16 spp; DRT_NO_DIRECT_SORT keeps the queue order for A/B runs)
  int sorted = 0;
  static const bool directSort = std::getenv("DRT_NO_DIRECT_SORT") == nullptr;
  if (directSort) {
    CK(c, STAGE(launchMaterialSort)(rs, wf, cur, sms, &sorted, st)); profMark(c, DRT_PK_OTHER);
    if (sorted) c->launches += 3;
  }
  for (int li = 0; li < nL; ++li) {
    const int nS = one ? 1 : r->direct[li].nSamples;
    for (int j = 0; j < nS; ++j) {
      CK(c, STAGE(launchResetCounts)(wf, (1u << Q_SHADOW) | (1u << Q_MIS), st)); profMark(c, DRT_PK_OTHER);
      CK(c, STAGE(launchDirectSample)(p, rs, wf, one ? -1 : li, j, cur, rc, sorted, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
      RK(traceQueue(c, true, wf.shO, wf.shD, wf.shRange, wf.counts + Q_SHADOW, wf.shOcc, nullptr, st));
      RK(traceQueue(c, false, wf.misO, wf.misD, wf.misRange, wf.counts + Q_MIS, wf.misHit, wf.misT, st));
      int mode = RESOLVE_DIRECT | (weighted ? RESOLVE_WEIGHTED : 0);
      if (one) mode |= RESOLVE_ONE;
      else {
        if (j == 0) mode |= RESOLVE_FIRST_OF_LIGHT;
        if (j == nS - 1) mode |= RESOLVE_LAST_OF_LIGHT;
        if (j == nS - 1 && li == nL - 1) mode |= RESOLVE_FINAL;
      }
      CK(c, STAGE(launchResolveDirect)(p, rs, wf, cur, mode, nS, sms, st)); profMark(c, DRT_PK_RESOLVE);
      c->launches += 3;
    }
  }
  return DRT_OK;
}

// The specular recursion of DirectLightingIntegrator.Li (direct_lighting_integrator.dart:56-64 -> Integrator.
// SpecularReflect / SpecularTransmit, integrator.dart:187-290 -> renderer.Li) evaluated chain by chain.  Radiance is
// linear in the chain weights, so L = sum over chains (one reflect / transmit choice per level) of weight x
// (Le + direct lighting at the chain's last vertex).  Chains are visited in the recursion's own depth-first order
// (reflect before transmit), each re-walking its prefix from the saved camera-ray queue; the LAST step of a chain is a
// call the recursion makes at that moment, so the per-slot stream counters advance exactly as the reference's shared
// RNG does (one BSDFSample.random per call, valid component or not).  A prefix nobody survives prunes its subtree.
// WhittedIntegrator.Li without its recursion (whitted_integrator.dart:26-63) on the vertices of queue `cur`.
static int whittedStage(drt_ctx* c, RenderState* r, int cur, bool weighted) {
  const RenderParams& p = r->rp;
  const RenderScene& rs = r->rs;
  const Wavefront& wf = r->wf;
  cudaStream_t st = c->stream;
  const int sms = c->numSMs;
  RenderCounters* rc = r->dCounters.p;
  if (rs.nPrograms > 0) {
    CK(c, launchTexturePass(p, rs, wf, cur, weighted ? 0 : 1, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
    c->launches++;
  }
  CK(c, STAGE(launchWhittedSetup)(p, rs, wf, cur, weighted ? 1 : 0, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
  c->launches++;
  int sorted = 0;  // material order measured no gain here (33.7 vs 34.1 ms, cornell_materials 1080p x 16 spp): off unless DRT_WHITTED_SORT is set
  static const bool whittedSort = std::getenv("DRT_WHITTED_SORT") != nullptr;
  if (whittedSort && rs.nLights > 0) {
    CK(c, STAGE(launchMaterialSort)(rs, wf, cur, sms, &sorted, st)); profMark(c, DRT_PK_OTHER);
    if (sorted) c->launches += 3;
  }
  for (int li = 0; li < rs.nLights; ++li) {
    CK(c, STAGE(launchResetCounts)(wf, (1u << Q_SHADOW) | (1u << Q_MIS), st)); profMark(c, DRT_PK_OTHER);
    CK(c, STAGE(launchWhittedSample)(p, rs, wf, li, cur, rc, sorted, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
    RK(traceQueue(c, true, wf.shO, wf.shD, wf.shRange, wf.counts + Q_SHADOW, wf.shOcc, nullptr, st));
    CK(c, STAGE(launchResolveDirect)(p, rs, wf, cur, RESOLVE_DIRECT | RESOLVE_WHITTED | (weighted ? RESOLVE_WEIGHTED : 0), 1, sms, st)); profMark(c, DRT_PK_RESOLVE);
    c->launches += 3;
  }
  return DRT_OK;
}

static int integratorStage(drt_ctx* c, RenderState* r, int cur, bool weighted) {
  return r->rp.integKind == 3 ? whittedStage(c, r, cur, weighted) : directStage(c, r, cur, weighted);
}

static int specularChains(drt_ctx* c, RenderState* r) {
  const RenderParams& p = r->rp;
  const RenderScene& rs = r->rs;
  const Wavefront& wf = r->wf;
  cudaStream_t st = c->stream;
  const int sms = c->numSMs;
  RenderCounters* rc = r->dCounters.p;
  const size_t cap = wf.cap;
  const int maxLevel = std::min(p.maxDepth - 1, kMaxChainLevels - 1);  // a vertex at depth d recurses iff d + 1 < maxDepth
  if (maxLevel < 1) return DRT_OK;
  // save the camera-ray queue (buffer 0 and the hit results of its trace)
  CK(c, cudaMemcpyAsync(wf.bakO, wf.extO[0], cap * sizeof(float4), cudaMemcpyDeviceToDevice, st));
  CK(c, cudaMemcpyAsync(wf.bakD, wf.extD[0], cap * sizeof(float4), cudaMemcpyDeviceToDevice, st));
  CK(c, cudaMemcpyAsync(wf.bakRange, wf.extRange[0], cap * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  CK(c, cudaMemcpyAsync(wf.bakSlot, wf.extSlot[0], cap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  CK(c, cudaMemcpyAsync(wf.bakHit, wf.extHit, cap * sizeof(float4), cudaMemcpyDeviceToDevice, st));
  CK(c, cudaMemcpyAsync(wf.bakT, wf.extT, cap * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (wf.bakInst) CK(c, cudaMemcpyAsync(wf.bakInst, wf.extInst, cap * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  uint32_t n0 = 0;
  CK(c, cudaMemcpyAsync(&n0, wf.counts + Q_EXT0, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  if (n0 == 0) return DRT_OK;
  const int kFlags[2] = {1 | 16, 2 | 16};  // BSDF_REFLECTION | BSDF_SPECULAR, BSDF_TRANSMISSION | BSDF_SPECULAR
  std::vector<int> chain;  // branch choice per level
  // explicit depth-first walk: `next[level]` = the next branch to try below the current prefix of that length
  std::vector<int> next(1, 0);
  while (!next.empty()) {
    const int len = (int)next.size() - 1;  // current prefix length
    if (next.back() >= 2) {                // both branches of this prefix done
      next.pop_back();
      if (!chain.empty()) chain.pop_back();
      continue;
    }
    const int b = next.back()++;
    chain.push_back(b);
    // re-walk the chain from the camera-ray queue
    CK(c, cudaMemcpyAsync(wf.extO[0], wf.bakO, n0 * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    CK(c, cudaMemcpyAsync(wf.extD[0], wf.bakD, n0 * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    CK(c, cudaMemcpyAsync(wf.extRange[0], wf.bakRange, n0 * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    CK(c, cudaMemcpyAsync(wf.extSlot[0], wf.bakSlot, n0 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    CK(c, cudaMemcpyAsync(wf.extHit, wf.bakHit, n0 * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    CK(c, cudaMemcpyAsync(wf.extT, wf.bakT, n0 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (wf.bakInst) CK(c, cudaMemcpyAsync(wf.extInst, wf.bakInst, n0 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    CK(c, cudaMemcpyAsync(wf.counts + Q_EXT0, &n0, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    int cur = 0;
    for (int k = 1; k <= len + 1; ++k) {
      CK(c, STAGE(launchResetCounts)(wf, 1u << (cur ^ 1), st)); profMark(c, DRT_PK_OTHER);
      CK(c, STAGE(launchSpecularStep)(p, rs, wf, cur, kFlags[chain[k - 1]], k, k == len + 1 ? 1 : 0, rc, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
      c->launches += 2;
      cur ^= 1;
      RK(traceQueue(c, false, wf.extO[cur], wf.extD[cur], wf.extRange[cur], wf.counts + cur, wf.extHit, wf.extT, st));
    }
    if (rs.nInfinite > 0) {  // the new rays of this chain that escape: renderer.Li = sum of Le, times the chain weight
      CK(c, STAGE(launchEscape)(rs, wf, cur, ESCAPE_WEIGHTED, sms, st)); profMark(c, DRT_PK_OTHER);
      c->launches++;
    }
    uint32_t live = 0;
    CK(c, cudaMemcpyAsync(&live, wf.counts + cur, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    if (live == 0) {  // nobody took this branch: no vertex, no deeper calls
      chain.pop_back();
      continue;
    }
    RK(integratorStage(c, r, cur, true));
    if (len + 1 < maxLevel) next.push_back(0);  // the new vertices recurse themselves
    else chain.pop_back();
  }
  return DRT_OK;
}

// One batch of camera samples through the whole pipeline; everything is enqueued on c->stream.
static int renderBatch(drt_ctx* c, RenderState* r, const PixelBatch& pb) {
  const RenderParams& p = r->rp;
  const RenderScene& rs = r->rs;
  const Wavefront& wf = r->wf;
  cudaStream_t st = c->stream;
  const int sms = c->numSMs;
  const uint32_t nSlots = pb.nPixels * (uint32_t)p.nPixelSamples;
  RenderCounters* rc = r->dCounters.p;
  CK(c, STAGE(launchSampler)(p, wf, r->dArrays.p, (int)r->arrays.size(), r->maxVals, r->maxOthers, pb, sms, st)); profMark(c, DRT_PK_SAMPLER);
  CK(c, STAGE(launchResetCounts)(wf, 0xffu, st)); profMark(c, DRT_PK_OTHER);
  CK(c, STAGE(launchRaygen)(p, wf, pb, rc, st)); profMark(c, DRT_PK_SAMPLER);
  c->launches += 3;
  // camera rays: Scene.intersect (sampler_renderer.dart:84)
  RK(traceQueue(c, false, wf.extO[0], wf.extD[0], wf.extRange[0], wf.counts + Q_EXT0, wf.extHit, wf.extT, st));
  if (rs.nInfinite > 0) {  // escaped camera rays see the infinite lights (sampler_renderer.dart:86-92), whatever the integrator
    CK(c, STAGE(launchEscape)(rs, wf, 0, ESCAPE_CAMERA, sms, st)); profMark(c, DRT_PK_OTHER);
    c->launches++;
  }
  if (p.samplerKind == 4) {
    CK(c, STAGE(launchSaveCameraPrims)(wf, nSlots, st)); profMark(c, DRT_PK_OTHER);
    c->launches++;
  }
  if (rs.nVolumes > 0) {  // VolumeIntegrator.Li along the camera rays (sampler_renderer.dart:93-95): T and Lvi per slot
    CK(c, cudaMemsetAsync(wf.trCtr, 0, (size_t)wf.cap * sizeof(uint32_t), st));
    CK(c, drt::extra::launchVolumeLi(p, rs, wf, rc, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
    c->launches++;
  }
  if (p.samplerKind != 3 && p.samplerKind != 5) {  // halton / bestcandidate: the accepted samples are counted on the device (raygenKernel)
    r->stats.camera_samples += nSlots;
    r->stats.closest_rays += nSlots;
  }
  if (p.integKind == 0) {
    int cur = 0;
    // float32 shading (drt_set_shading_precision; env DRT_SHADE_F32=1 / 0 overrides for A/B runs): same queues, same sample values, the
    // camera rays traced in binary64 as always; the vertex / resolve kernels and the traversal of the integrator's own ray queues run
    // their float32 builds
    static const char* f32Env = std::getenv("DRT_SHADE_F32");
    const bool wantF32 = f32Env ? f32Env[0] == '1' : r->shadingPrecision == DRT_PRECISION_F32;
    const bool f32 = wantF32 && rs.nVolumes == 0 && rs.nPrograms == 0 && c->ts.nInstances == 0 && wf.slotTime == nullptr;
    static const bool f32TraceOff = std::getenv("DRT_F32_TRACE") != nullptr && std::getenv("DRT_F32_TRACE")[0] == '0';
    r->f32Trace = f32 && !f32TraceOff;
    const auto shadeF32 = rs.extra ? drt::extraf::launchShadePath : drt::plainf::launchShadePath;
    const auto resolveF32 = rs.extra ? drt::extraf::launchResolveDirect : drt::plainf::launchResolveDirect;
    for (int bounce = 0; bounce <= p.maxDepth; ++bounce) {
      CK(c, STAGE(launchResetCounts)(wf, (1u << (cur ^ 1)) | (1u << Q_SHADOW) | (1u << Q_MIS), st)); profMark(c, DRT_PK_OTHER);
      if (rs.nPrograms > 0) {  // only the camera ray carries differentials: the later rays are RayDifferential.child (path_integrator.dart:100)
        CK(c, launchTexturePass(p, rs, wf, cur, bounce == 0 ? 1 : 0, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
        c->launches++;
      }
      CK(c, (f32 ? shadeF32 : STAGE(launchShadePath))(p, rs, wf, bounce, cur, rc, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
      c->launches += 2;
      if (rs.nLights > 0) {
        // float32 shading: minDistance / maxDistance are the float32 values in the ray records' .w lanes (no binary64 range arrays)
        RK(traceQueue(c, true, wf.shO, wf.shD, f32 ? nullptr : wf.shRange, wf.counts + Q_SHADOW, wf.shOcc, nullptr, st));
        RK(traceQueue(c, false, wf.misO, wf.misD, f32 ? nullptr : wf.misRange, wf.counts + Q_MIS, wf.misHit, f32 ? nullptr : wf.misT, st));
        CK(c, (f32 ? resolveF32 : STAGE(launchResolveDirect))(p, rs, wf, cur, RESOLVE_PATH, 1, sms, st)); profMark(c, DRT_PK_RESOLVE);
        c->launches++;
      }
      if (bounce == p.maxDepth) break;
      cur ^= 1;
      RK(traceQueue(c, false, wf.extO[cur], wf.extD[cur], f32 ? nullptr : wf.extRange[cur], wf.counts + cur, wf.extHit, f32 ? nullptr : wf.extT, st));
      if (rs.nInfinite > 0 && rs.general) {  // path_integrator.dart:106-114: only after a specular bounce, which matte scenes never take
        CK(c, STAGE(launchEscape)(rs, wf, cur, ESCAPE_PATH, sms, st)); profMark(c, DRT_PK_OTHER);
        c->launches++;
      }
    }
    r->f32Trace = false;
  } else if (p.integKind == 1) {
    CK(c, STAGE(launchAoSetup)(p, rs, wf, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
    c->launches++;
    const int nS = roundUpPow2(p.aoSamples);
    const uint32_t hitsPerChunk = std::max<uint32_t>(1, r->shCap / (uint32_t)nS);
    for (uint32_t first = 0; first < nSlots; first += hitsPerChunk) {
      CK(c, STAGE(launchAoGen)(p, wf, first, hitsPerChunk, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
      RK(traceQueue(c, true, wf.shO, wf.shD, wf.shRange, wf.counts + Q_SHADOW, wf.shOcc, nullptr, st));
      CK(c, STAGE(launchAoCount)(p, wf, first, hitsPerChunk, rc, sms, st)); profMark(c, DRT_PK_INTEGRATOR);
      c->launches += 2;
    }
  } else {  // directlighting / whitted: the integrator at the camera vertices, then its specular recursion
    if (wf.specCtr) CK(c, cudaMemsetAsync(wf.specCtr, 0, (size_t)wf.cap * sizeof(uint32_t), st));
    RK(integratorStage(c, r, 0, false));
    if (wf.specCtr && r->hasSpecular) RK(specularChains(c, r));
  }
  const bool firstVisit = p.samplerKind == 4 && pb.pass == 0;
  if (firstVisit) {  // reportResults (adaptive_sampler.dart:133-158): supersampled pixels drop this visit's samples
    CK(c, STAGE(launchAdaptiveDecide)(p, wf, pb, r->dAdaptList.p, r->dAdaptCount.p, st)); profMark(c, DRT_PK_OTHER);
    c->launches++;
  }
  if (rs.nVolumes > 0) {  // T * Li + Lvi (sampler_renderer.dart:97)
    CK(c, drt::extra::launchVolumeCombine(wf, nSlots, st)); profMark(c, DRT_PK_OTHER);
    c->launches++;
  }
  CK(c, STAGE(launchFilm)(p, wf, nSlots, firstVisit ? 1 : 0, rc, st)); profMark(c, DRT_PK_FILM);
  c->launches++;
  return DRT_OK;
}

static int renderWindow(drt_ctx* c, int x, int y, int w, int h, uint32_t shard, uint32_t nShards) {
  RenderState* r = state(c);
  RK(prepare(c, r));
  const RenderParams& p = r->rp;
  if (p.integKind >= 2 && r->hasSpecular && p.maxDepth - 1 > kMaxChainLevels - 1)
    return fail(c, DRT_E_UNSUPPORTED, "directlighting / whitted with specular BxDFs: maxdepth above 16 is not on the GPU path");
  if (w <= 0 || h <= 0) return DRT_OK;
  uint64_t total = (uint64_t)w * h;
  const bool halton = p.samplerKind == 3 || p.samplerKind == 5;  // a global sequence of sample indices instead of pixels
  if (p.samplerKind == 5) {  // best_candidate_sampler.dart:36-52: every entry of the pattern in every tile the window touches
    const double tableWidth = 64 / std::sqrt((double)r->spp);
    const int xs0 = (int)std::floor(x / tableWidth), xs1 = (int)std::floor((x + w - 1) / tableWidth);
    const int ys0 = (int)std::floor(y / tableWidth), ys1 = (int)std::floor((y + h - 1) / tableWidth);
    const int nx = xs1 - xs0 + 1, ny = ys1 - ys0 + 1;
    total = (uint64_t)nx * ny * 4096;
    std::vector<double> shifts(3 * (size_t)nx * ny);
    for (int ty = 0; ty < ny; ++ty)
      for (int tx = 0; tx < nx; ++tx) {
        DartRandom tileRng((int64_t)(xs0 + tx) + ((int64_t)(ys0 + ty) << 8));  // :44-47,91-94
        for (int k = 0; k < 3; ++k) shifts[3 * ((size_t)ty * nx + tx) + k] = tileRng.nextDouble();
      }
    CK(c, r->dBcTable.ensure(r->sampleTable.size()));
    CK(c, r->dBcShifts.ensure(shifts.size()));
    CK(c, cudaMemcpy(r->dBcTable.p, r->sampleTable.data(), r->sampleTable.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(r->dBcShifts.p, shifts.data(), shifts.size() * sizeof(double), cudaMemcpyHostToDevice));
    r->rp.bcTable = r->dBcTable.p;
    r->rp.bcTileShifts = r->dBcShifts.p;
    r->rp.bcXTileStart = xs0; r->rp.bcYTileStart = ys0; r->rp.bcTilesX = nx;
    r->rp.bcTableWidth = tableWidth;
    r->rp.winX = x; r->rp.winY = y; r->rp.winW = w; r->rp.winH = h;
  } else if (halton) {  // halton_sampler.dart:32-38: spp * delta^2 indices of the sequence take the place of the window's pixels
    const uint64_t delta = (uint64_t)std::max(w, h);
    total = (uint64_t)r->spp * delta * delta;
    r->rp.winX = x; r->rp.winY = y; r->rp.winW = w; r->rp.winH = h;
  }
  const uint32_t blockPixels = 1024;
  uint64_t mine = total;
  if (nShards > 1) {  // pixels of the blocks shard, shard + nShards, ...
    const uint64_t nBlocks = (total + blockPixels - 1) / blockPixels;
    const uint64_t owned = nBlocks > shard ? (nBlocks - shard + nShards - 1) / nShards : 0;
    mine = owned * blockPixels;
    if (owned && (nBlocks - 1) % nShards == shard) mine -= nBlocks * blockPixels - total;
  }
  // 16 Mi camera samples in flight (~10 GB of wavefront state): measured on B200, config 4 takes 1063 / 1022 / 1002 / 991 ms at
  // 4 / 8 / 16 / 32 Mi slots (tools/batch_sweep.sh) — fewer, longer launches per bounce
  uint64_t slots = r->batchSlots ? r->batchSlots : (1ull << 24);
  if (!r->programs.empty()) slots = std::min<uint64_t>(slots, 1ull << 21);  // 8 x 88 B of per-slot BSDF records (texture pass): 1.5 GB
  if (const char* e = std::getenv("DRT_BATCH_SLOTS")) slots = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));
  if (!r->volumes.empty() && r->volIntegrator == 1)  // the march's sample arrays: 16 bytes x steps per slot, at most ~8 GB
    slots = std::max<uint64_t>(4096, std::min<uint64_t>(slots, (8ull << 30) / (16ull * std::max<uint32_t>(r->volMaxSteps, 1))));
  // One visit of `count` pixels (of this shard's part of the window, or of `list`) with the current p.nPixelSamples per pixel
  if (r->profFlags & DRT_PROFILE_WORK) {
    const bool fresh = r->dWork.p == nullptr;
    CK(c, r->dWork.ensure(2));
    if (fresh) CK(c, cudaMemset(r->dWork.p, 0, 2 * sizeof(DeviceCounters)));
  }
  profMark(c, -1);
  auto runVisit = [&](uint64_t count, uint32_t visit, const uint32_t* list) -> int {
    uint64_t pixelsPerBatch = std::max<uint64_t>(1, slots / (uint64_t)p.nPixelSamples);
    pixelsPerBatch = std::min<uint64_t>(pixelsPerBatch, std::max<uint64_t>(count, 1));
    const uint64_t cap64 = pixelsPerBatch * (uint64_t)p.nPixelSamples;
    if (cap64 > 0x7fffffffull) return fail(c, DRT_E_INVALID, "samples per pixel too large for one batch");
    const uint32_t cap = (uint32_t)cap64;
    uint32_t shCap = cap;
    if (p.integKind == 1) {
      const uint64_t nS = (uint64_t)roundUpPow2(p.aoSamples);
      shCap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(cap, std::max<uint64_t>(nS, 1ull << 24)), (uint64_t)cap * nS);
      shCap = std::max<uint32_t>(shCap, (uint32_t)nS);
    }
    RK(ensureWavefront(c, r, cap, shCap));
    const int passes = p.samplerKind == 2 ? r->spp : 1;  // random sampler: spp visits of spp samples (random_sampler.dart:47-88)
    for (int pass = 0; pass < passes; ++pass)
      for (uint64_t first = 0; first < count; first += pixelsPerBatch) {
        PixelBatch pb;
        pb.x0 = x; pb.y0 = y; pb.w = w;
        if (halton) { pb.x0 = 0; pb.y0 = 0; pb.w = 1 << 30; }  // pixelOf: (n mod 2^30, n / 2^30)
        pb.firstPixel = first;
        pb.nPixels = (uint32_t)std::min<uint64_t>(pixelsPerBatch, count - first);
        pb.pass = p.samplerKind == 4 ? visit : (uint32_t)pass;
        pb.shard = shard; pb.nShards = nShards; pb.blockPixels = blockPixels;
        pb.list = list;
        RK(renderBatch(c, r, pb));
      }
    return DRT_OK;
  };
  if (p.samplerKind == 4) {  // adaptive_sampler.dart:101-158: every pixel with minSamples, the flagged ones again with maxSamples
    if (total > 0xffffffffull) return fail(c, DRT_E_INVALID, "adaptive sampler: window too large");
    CK(c, r->dAdaptList.ensure(std::max<uint64_t>(mine, 1)));
    CK(c, r->dAdaptCount.ensure(1));
    CK(c, cudaMemsetAsync(r->dAdaptCount.p, 0, sizeof(uint32_t), c->stream));
    RK(runVisit(mine, 0, nullptr));
    uint32_t nSuper = 0;
    CK(c, cudaMemcpyAsync(&nSuper, r->dAdaptCount.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (nSuper > 0) {
      int mn, mx;
      adaptiveCounts(p.xs, p.ys, &mn, &mx);
      r->rp.nPixelSamples = mx;
      buildLayout(r);  // the sample arrays hold maxSamples values per pixel now
      CK(c, r->dArrays.ensure(r->arrays.size()));
      CK(c, cudaMemcpy(r->dArrays.p, r->arrays.data(), r->arrays.size() * sizeof(SampleArray), cudaMemcpyHostToDevice));
      int rc2 = runVisit(nSuper, 1, r->dAdaptList.p);
      r->rp.nPixelSamples = mn;
      buildLayout(r);
      CK(c, cudaMemcpy(r->dArrays.p, r->arrays.data(), r->arrays.size() * sizeof(SampleArray), cudaMemcpyHostToDevice));
      if (rc2 != DRT_OK) return rc2;
    }
  } else {
    RK(runVisit(mine, 0, nullptr));
  }
  CK(c, cudaStreamSynchronize(c->stream));
  RenderCounters hc;
  CK(c, cudaMemcpy(&hc, r->dCounters.p, sizeof(hc), cudaMemcpyDeviceToHost));
  r->stats.closest_rays += hc.closestRays + hc.cameraSamples;  // halton: one camera ray per accepted sample
  r->stats.camera_samples += hc.cameraSamples;
  r->stats.shadow_rays += hc.shadowRays;
  r->stats.zeroed_samples += hc.zeroedSamples;
  CK(c, cudaMemset(r->dCounters.p, 0, sizeof(RenderCounters)));
  if (r->profFlags) RK(profCollect(c));
  return DRT_OK;
}

// ---- multi-device contexts (drt_create_multi) ------------------------------------------------------------------------------
// What lib/dartray_web/render_manager.dart:100-141 does with one isolate per image region and a copy of each region's pixels:
// here one GPU per set of interleaved pixel blocks, and a SUM of the films (filter footprints cross block borders).
struct FilmPtrs {
  const double* src[15];
};

// dst += sum of the peers' films, read through peer-mapped pointers over NVLink (one pass, no staging copy)
__global__ void filmSumKernel(double* __restrict__ dst, FilmPtrs peers, int nPeers, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double v = dst[i];
    for (int k = 0; k < nPeers; ++k) v += peers.src[k][i];
    dst[i] = v;
  }
}

static int reduceFilms(drt_ctx* c) {
  RenderState* r = state(c);
  const size_t n = 4 * r->filmPixels;
  if (n == 0) return DRT_OK;
  CK(c, cudaSetDevice(c->device));
  bool direct = true;
  for (drt_ctx* p : c->peers) {
    if (4 * state(p)->filmPixels != n) return fail(c, DRT_E_STATE, "multi-device render: the devices' films differ in size");
    int can = 1;
    if (p->device != c->device) CK(c, cudaDeviceCanAccessPeer(&can, c->device, p->device));
    if (!can) direct = false;
  }
  if (direct && !c->peerAccessTried) {
    for (drt_ctx* p : c->peers) {
      if (p->device == c->device) continue;
      cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) { cudaGetLastError(); direct = false; }
    }
    c->peerAccessTried = true;
  }
  const int grid = c->numSMs * 8;
  if (direct) {
    for (size_t first = 0; first < c->peers.size(); first += 15) {
      FilmPtrs fp{};
      const int cnt = (int)std::min<size_t>(15, c->peers.size() - first);
      for (int k = 0; k < cnt; ++k) fp.src[k] = state(c->peers[first + k])->dFilm.p;
      filmSumKernel<<<grid, 256, 0, c->stream>>>(r->dFilm.p, fp, cnt, n);
      CK(c, cudaGetLastError());
      c->launches++;
    }
  } else {  // no peer mapping between these devices: copy each film over and add it
    DevBuf<double> stage;
    CK(c, stage.ensure(n));
    for (drt_ctx* p : c->peers) {
      CK(c, cudaMemcpyPeerAsync(stage.p, c->device, state(p)->dFilm.p, p->device, n * sizeof(double), c->stream));
      FilmPtrs fp{};
      fp.src[0] = stage.p;
      filmSumKernel<<<grid, 256, 0, c->stream>>>(r->dFilm.p, fp, 1, n);
      CK(c, cudaGetLastError());
      c->launches++;
    }
    CK(c, cudaStreamSynchronize(c->stream));
    stage.release();
  }
  CK(c, cudaStreamSynchronize(c->stream));
  // the peers' films hold only what they rendered since the last sum: a later render into the same film adds its own delta
  for (drt_ctx* p : c->peers) {
    CK(c, cudaSetDevice(p->device));
    CK(c, cudaMemset(state(p)->dFilm.p, 0, n * sizeof(double)));
  }
  CK(c, cudaSetDevice(c->device));
  return DRT_OK;
}

static int renderWindow(drt_ctx* c, int x, int y, int w, int h, uint32_t shard, uint32_t nShards);

// One render call of a multi-device context: device k renders shard (shard * D + k) of (nShards * D) of the window — the
// interleaved 1024-pixel blocks of drt_render_shard, so that the union of the devices' samples is exactly the one-device
// sample set (streams are keyed by pixel) — each on its own host thread; then the films are summed into this device's.
static int renderMulti(drt_ctx* c, int x, int y, int w, int h, uint32_t shard, uint32_t nShards) {
  const uint32_t D = 1 + (uint32_t)c->peers.size();
  std::vector<int> rcs(D, DRT_OK);
  std::vector<std::thread> th;
  for (uint32_t k = 1; k < D; ++k)
    th.emplace_back([&, k] { rcs[k] = renderWindow(c->peers[k - 1], x, y, w, h, shard * D + k, nShards * D); });
  rcs[0] = renderWindow(c, x, y, w, h, shard * D, nShards * D);
  for (auto& t : th) t.join();
  for (uint32_t k = 1; k < D; ++k)
    if (rcs[k] != DRT_OK) { c->err = "device " + std::to_string(c->peers[k - 1]->device) + ": " + c->peers[k - 1]->err; return rcs[k]; }
  if (rcs[0] != DRT_OK) return rcs[0];
  return reduceFilms(c);
}

static void sampleExtent(const RenderParams& p, int e[4]) {  // image_film.dart:247-252
  e[0] = (int)std::floor(p.left + 0.5 - p.xWidth);
  e[1] = (int)std::ceil(p.left + 0.5 + p.width + p.xWidth);
  e[2] = (int)std::floor(p.top + 0.5 - p.yWidth);
  e[3] = (int)std::ceil(p.top + 0.5 + p.height + p.yWidth);
}

extern "C" {

int drt_set_materials(drt_ctx* c, uint32_t n, const int32_t* kind, const float* kd, const float* sigma) {
  DRT_FORWARD_TO_PEERS(c, drt_set_materials(p_, n, kind, kd, sigma));
  if (!c) return DRT_E_INVALID;
  if (n && !kd) return fail(c, DRT_E_INVALID, "null Kd array");
  RenderState* r = state(c);
  r->materials.resize(std::max<uint32_t>(n, 1));
  if (n == 0) r->materials[0] = GMaterial{{0.5f, 0.5f, 0.5f}, 0.f};
  for (uint32_t i = 0; i < n; ++i) {
    if (kind && kind[i] != 0) return fail(c, DRT_E_INVALID, "only material kind 0 (matte) is on the GPU path");
    r->materials[i] = GMaterial{{kd[3 * i], kd[3 * i + 1], kd[3 * i + 2]}, sigma ? sigma[i] : 0.f};
  }
  r->general = r->hasSpecular = r->hasBlend = false;
  r->matLobes.clear();
  r->lobes.clear();
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_material_lobes(drt_ctx* c, uint32_t n, const uint32_t* lobe_offsets, const int32_t* lobe_kind, const float* lobe_rgb,
                           const int32_t* fresnel_kind, const float* fresnel_eta, const float* fresnel_k,
                           const double* lobe_scalars) {
  DRT_FORWARD_TO_PEERS(c, drt_set_material_lobes(p_, n, lobe_offsets, lobe_kind, lobe_rgb, fresnel_kind, fresnel_eta, fresnel_k, lobe_scalars));
  if (!c) return DRT_E_INVALID;
  if (n == 0 || !lobe_offsets) return fail(c, DRT_E_INVALID, "drt_set_material_lobes needs at least one material and its offsets");
  const uint32_t nl = lobe_offsets[n];
  if (nl && (!lobe_kind || !lobe_rgb || !lobe_scalars)) return fail(c, DRT_E_INVALID, "null lobe arrays");
  RenderState* r = state(c);
  std::vector<uint2> ml(n);
  std::vector<GLobe> ls(nl);
  bool spec = false, blend = false, measuredLobe = false;
  for (uint32_t i = 0; i < n; ++i) {
    if (lobe_offsets[i + 1] < lobe_offsets[i] || lobe_offsets[i + 1] - lobe_offsets[i] > 8)
      return fail(c, DRT_E_INVALID, "a BSDF holds at most 8 BxDFs (bsdf.dart:253) and the offsets must not decrease");
    ml[i] = make_uint2(lobe_offsets[i], lobe_offsets[i + 1] - lobe_offsets[i]);
  }
  for (uint32_t j = 0; j < nl; ++j) {
    GLobe& l = ls[j];
    l.kind = lobe_kind[j];
    l.fresnel = fresnel_kind ? fresnel_kind[j] : 0;
    if (l.kind < 0 || l.kind > 7 || l.fresnel < 0 || l.fresnel > 2) return fail(c, DRT_E_INVALID, "unknown BxDF or Fresnel kind");
    measuredLobe = measuredLobe || l.kind >= 6;
    if (l.kind == 5 && !fresnel_eta) return fail(c, DRT_E_INVALID, "FresnelBlend (kind 5) carries Rs in the eta array");
    blend = blend || l.kind == 5;
    if (l.fresnel == 2 && (!fresnel_eta || !fresnel_k)) return fail(c, DRT_E_INVALID, "FresnelConductor needs eta and k");
    for (int k = 0; k < 3; ++k) {
      l.rgb[k] = lobe_rgb[3 * j + k];
      l.eta[k] = fresnel_eta ? fresnel_eta[3 * j + k] : 0.f;
      l.k[k] = fresnel_k ? fresnel_k[3 * j + k] : 0.f;
    }
    l.pad_ = 0.f;
    l.wrap = 0;
    l.scale[0] = l.scale[1] = l.scale[2] = 1.f;
    l.param = lobe_scalars[3 * j];
    l.ei = lobe_scalars[3 * j + 1];
    l.et = lobe_scalars[3 * j + 2];
    spec = spec || l.kind == 3 || l.kind == 4;
  }
  r->materials.assign(n, GMaterial{{0.f, 0.f, 0.f}, 0.f});  // the single-lobe table is not read when `general` is set
  r->matLobes.swap(ml);
  r->lobes.swap(ls);
  r->general = true;
  r->hasSpecular = spec;
  r->hasBlend = blend;
  r->hasMeasured = measuredLobe;
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_measured(drt_ctx* c, uint32_t n_tables, const int32_t* kind, const int32_t* dims, const uint64_t* offsets, const float* data,
                     uint64_t n_floats) {
  DRT_FORWARD_TO_PEERS(c, drt_set_measured(p_, n_tables, kind, dims, offsets, data, n_floats));
  if (!c) return DRT_E_INVALID;
  if (n_tables && (!kind || !dims || !offsets || !data)) return fail(c, DRT_E_INVALID, "drt_set_measured: null arrays");
  RenderState* r = state(c);
  std::vector<GMeasured> ms(n_tables);
  for (uint32_t i = 0; i < n_tables; ++i) {
    GMeasured& m = ms[i];
    m.data = nullptr;
    m.kind = kind[i];
    for (int k = 0; k < 3; ++k) m.dims[k] = dims[3 * i + k];
    if (m.kind != 0 && m.kind != 1) return fail(c, DRT_E_INVALID, "measured table kind: 0 = regular halfangle (.merl), 1 = irregular isotropic (.brdf)");
    if (m.dims[0] < 1 || (m.kind == 0 && (m.dims[1] < 1 || m.dims[2] < 1))) return fail(c, DRT_E_INVALID, "measured table dimensions must be positive");
    const uint64_t need = m.kind == 0 ? 3ull * (uint64_t)m.dims[0] * (uint64_t)m.dims[1] * (uint64_t)m.dims[2] : 6ull * (uint64_t)m.dims[0];
    if (offsets[i] > n_floats || need > n_floats - offsets[i]) return fail(c, DRT_E_INVALID, "a measured table reaches beyond the data array");
  }
  r->measured.swap(ms);
  r->measuredOffsets.assign(offsets, offsets + n_tables);
  r->measuredData.assign(data, data + (n_tables ? n_floats : 0));
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_lobe_wrappers(drt_ctx* c, uint32_t n_lobes, const int32_t* wrap, const float* scale_rgb) {
  DRT_FORWARD_TO_PEERS(c, drt_set_lobe_wrappers(p_, n_lobes, wrap, scale_rgb));
  if (!c) return DRT_E_INVALID;
  RenderState* r = state(c);
  if (!r->general || n_lobes != r->lobes.size())
    return fail(c, DRT_E_STATE, "drt_set_lobe_wrappers: the lobe count must equal that of the last drt_set_material_lobes");
  if (n_lobes && !wrap) return fail(c, DRT_E_INVALID, "null wrapper array");
  for (uint32_t j = 0; j < n_lobes; ++j) {
    if (wrap[j] < 0 || wrap[j] > 3) return fail(c, DRT_E_INVALID, "wrapper bits: 1 = BRDFToBTDF, 2 = ScaledBxDF");
    if ((wrap[j] & 2) && !scale_rgb) return fail(c, DRT_E_INVALID, "ScaledBxDF needs the scale array");
    GLobe& l = r->lobes[j];
    l.wrap = wrap[j];
    for (int k = 0; k < 3; ++k) l.scale[k] = (wrap[j] & 2) ? scale_rgb[3 * j + k] : 1.f;
  }
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_textures(drt_ctx* c, uint32_t n, const drt_texture* nodes, const float* texels, uint64_t n_texel_floats) {
  if (!c) return DRT_E_INVALID;
  if (n > 0 && !nodes) return fail(c, DRT_E_INVALID, "drt_set_textures: nodes is NULL");
  if (n_texel_floats > 0 && !texels) return fail(c, DRT_E_INVALID, "drt_set_textures: texels is NULL");
  RenderState* r = state(c);
  std::vector<GTex> tex;
  std::vector<float> data;
  std::string err;
  if (!buildTextureTables(n, nodes, texels, n_texel_floats, &tex, &data, &err)) {
    return fail(c, DRT_E_INVALID, ("drt_set_textures: " + err).c_str());
  }
  r->textures.swap(tex);
  r->texData.swap(data);
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_material_programs(drt_ctx* c, uint32_t n, const drt_material_program* programs) {
  if (!c) return DRT_E_INVALID;
  if (n > 0 && !programs) return fail(c, DRT_E_INVALID, "drt_set_material_programs: programs is NULL");
  RenderState* r = state(c);
  r->programs.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    GProgram& g = r->programs[i];
    g.kind = programs[i].kind;
    for (int k = 0; k < 8; ++k) g.tex[k] = programs[i].tex[k];
    g.bump = programs[i].bump;
    g.m1 = programs[i].m1;
    g.m2 = programs[i].m2;
  }
  r->sceneTablesValid = false;  // validated against the material and texture tables when the render starts
  return DRT_OK;
}

int drt_set_lights(drt_ctx* c, uint32_t n, const int32_t* kind, const float* L, const float* pos, const int32_t* nsamples,
                   const uint32_t* shape_offsets, const uint32_t* shape_prims) {
  DRT_FORWARD_TO_PEERS(c, drt_set_lights(p_, n, kind, L, pos, nsamples, shape_offsets, shape_prims));
  if (!c) return DRT_E_INVALID;
  if (n && (!kind || !L)) return fail(c, DRT_E_INVALID, "null light arrays");
  RenderState* r = state(c);
  std::vector<HostLight> ls(n);
  for (uint32_t i = 0; i < n; ++i) {
    HostLight& l = ls[i];
    l.kind = kind[i];
    if (l.kind < 0 || l.kind > 6)
      return fail(c, DRT_E_INVALID, "light kind must be 0 (diffuse area), 1 (point), 2 (distant), 3 (spot), 4 (infinite), 5 (projection) or 6 (goniometric)");
    if (l.kind != 0 && !pos) return fail(c, DRT_E_INVALID, "point / distant / spot lights need the pos array");
    std::memcpy(l.L, L + 3 * i, 12);
    if (pos) std::memcpy(l.pos, pos + 3 * i, 12);
    l.nSamples = nsamples ? std::max(1, nsamples[i]) : 1;
    if (shape_offsets && shape_prims)
      for (uint32_t k = shape_offsets[i]; k < shape_offsets[i + 1]; ++k) l.shapes.push_back(shape_prims[k]);
  }
  r->lights.swap(ls);
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_spot_params(drt_ctx* c, uint32_t n, const float* world_to_light, const double* cos_total_falloff) {
  DRT_FORWARD_TO_PEERS(c, drt_set_spot_params(p_, n, world_to_light, cos_total_falloff));
  if (!c) return DRT_E_INVALID;
  RenderState* r = state(c);
  if (n != r->lights.size()) return fail(c, DRT_E_STATE, "drt_set_spot_params: n must equal the light count of the last drt_set_lights");
  if (n && (!world_to_light || !cos_total_falloff)) return fail(c, DRT_E_INVALID, "null spot-light arrays");
  for (uint32_t i = 0; i < n; ++i) {
    HostLight& l = r->lights[i];
    const float* m = world_to_light + 16 * i;
    for (int row = 0; row < 3; ++row)
      for (int col = 0; col < 3; ++col) l.w2l[3 * row + col] = m[4 * row + col];
    l.cosTotalWidth = cos_total_falloff[2 * i];
    l.cosFalloffStart = cos_total_falloff[2 * i + 1];
    l.haveSpot = true;
  }
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_infinite_light(drt_ctx* c, uint32_t index, int width, int height, const float* rgb, const float* light_to_world,
                           const float* world_to_light) {
  DRT_FORWARD_TO_PEERS(c, drt_set_infinite_light(p_, index, width, height, rgb, light_to_world, world_to_light));
  if (!c) return DRT_E_INVALID;
  RenderState* r = state(c);
  if (index >= r->lights.size() || r->lights[index].kind != 4)
    return fail(c, DRT_E_STATE, "drt_set_infinite_light: the light must have been declared with kind 4 by the last drt_set_lights");
  if (!rgb || !light_to_world || !world_to_light) return fail(c, DRT_E_INVALID, "null infinite-light arrays");
  if (width < 1 || height < 1 || (width & (width - 1)) || (height & (height - 1)) || (uint64_t)width * height > (1u << 26))
    return fail(c, DRT_E_INVALID, "the radiance map must have power-of-two resolution (level 0 of the reference's MIPMap), at most 64 Mi texels");
  HostLight& l = r->lights[index];
  for (int row = 0; row < 3; ++row)
    for (int col = 0; col < 3; ++col) {
      l.l2w[3 * row + col] = light_to_world[4 * row + col];
      l.w2l[3 * row + col] = world_to_light[4 * row + col];
    }
  l.mapW = width;
  l.mapH = height;
  l.texels.assign(rgb, rgb + 3 * (size_t)width * height);
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_light_map(drt_ctx* c, uint32_t index, int width, int height, const float* rgb, const float* world_to_light,
                      const float* light_projection, const double* screen_window, double hither) {
  DRT_FORWARD_TO_PEERS(c, drt_set_light_map(p_, index, width, height, rgb, world_to_light, light_projection, screen_window, hither));
  if (!c) return DRT_E_INVALID;
  RenderState* r = state(c);
  if (index >= r->lights.size() || (r->lights[index].kind != 5 && r->lights[index].kind != 6))
    return fail(c, DRT_E_STATE, "drt_set_light_map: the light must have been declared with kind 5 or 6 by the last drt_set_lights");
  if (!world_to_light) return fail(c, DRT_E_INVALID, "null world_to_light");
  HostLight& l = r->lights[index];
  if (l.kind == 5 && (!light_projection || !screen_window)) return fail(c, DRT_E_INVALID, "a projection light needs its projection and screen window");
  if (rgb) {
    if (width < 1 || height < 1 || (width & (width - 1)) || (height & (height - 1)) || (uint64_t)width * height > (1u << 26))
      return fail(c, DRT_E_INVALID, "the map must have power-of-two resolution (level 0 of the reference's MIPMap), at most 64 Mi texels");
    l.mapW = width;
    l.mapH = height;
    l.texels.assign(rgb, rgb + 3 * (size_t)width * height);
  } else {
    l.mapW = l.mapH = 0;
    l.texels.clear();
  }
  for (int row = 0; row < 3; ++row)
    for (int col = 0; col < 3; ++col) l.w2l[3 * row + col] = world_to_light[4 * row + col];
  if (l.kind == 5) {
    std::memcpy(l.proj, light_projection, sizeof(l.proj));
    std::memcpy(l.screen, screen_window, sizeof(l.screen));
    l.hither = hither;
  }
  l.haveMapParams = true;
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_volumes(drt_ctx* c, uint32_t n, const int32_t* kind, const float* sigma_a, const float* sigma_s, const float* le, const double* g,
                    const float* p0_p1, const float* volume_to_world, const float* world_to_volume, const double* exp_a_b,
                    const float* up_dir, const int32_t* grid_dims, const uint64_t* density_offsets, const double* density) {
  DRT_FORWARD_TO_PEERS(c, drt_set_volumes(p_, n, kind, sigma_a, sigma_s, le, g, p0_p1, volume_to_world, world_to_volume, exp_a_b, up_dir,
                                          grid_dims, density_offsets, density));
  if (!c) return DRT_E_INVALID;
  if (n && (!kind || !sigma_a || !sigma_s || !le || !g || !p0_p1 || !world_to_volume)) return fail(c, DRT_E_INVALID, "null volume arrays");
  RenderState* r = state(c);
  std::vector<GVolume> vols(n);
  std::vector<double> dens;
  for (uint32_t i = 0; i < n; ++i) {
    GVolume& v = vols[i];
    std::memset(&v, 0, sizeof(v));
    v.kind = kind[i];
    if (v.kind < 0 || v.kind > 2) return fail(c, DRT_E_INVALID, "volume kind must be 0 (homogeneous), 1 (exponential) or 2 (volumegrid)");
    for (int k = 0; k < 3; ++k) {
      v.sigA[k] = sigma_a[3 * i + k];
      v.sigS[k] = sigma_s[3 * i + k];
      v.sigT[k] = (float)((double)v.sigA[k] + (double)v.sigS[k]);  // the Spectrum sig_a + sig_s
      v.le[k] = le[3 * i + k];
      // BBox(p0, p1): component-wise min / max (bbox.dart:35-41)
      v.lo[k] = std::fmin(p0_p1[6 * i + k], p0_p1[6 * i + 3 + k]);
      v.hi[k] = std::fmax(p0_p1[6 * i + k], p0_p1[6 * i + 3 + k]);
    }
    std::memcpy(v.w2v, world_to_volume + 16 * i, 64);
    v.g = g[i];
    v.a = v.b = 1.0;
    v.nx = v.ny = v.nz = 1;
    if (v.kind == 1) {
      if (!exp_a_b || !up_dir) return fail(c, DRT_E_INVALID, "exponential volume needs a, b and updir");
      v.a = exp_a_b[2 * i];
      v.b = exp_a_b[2 * i + 1];
      const V3 up = Normalize(V3{up_dir[3 * i], up_dir[3 * i + 1], up_dir[3 * i + 2]});  // exponential_density_region.dart:29
      v.up[0] = up.x; v.up[1] = up.y; v.up[2] = up.z;
    }
    if (v.kind == 2) {
      if (!grid_dims || !density_offsets || !density) return fail(c, DRT_E_INVALID, "volumegrid needs nx, ny, nz and the density values");
      v.nx = grid_dims[3 * i]; v.ny = grid_dims[3 * i + 1]; v.nz = grid_dims[3 * i + 2];
      const uint64_t cnt = density_offsets[i + 1] - density_offsets[i];
      if (v.nx < 1 || v.ny < 1 || v.nz < 1 || cnt != (uint64_t)v.nx * v.ny * v.nz)
        return fail(c, DRT_E_INVALID, "volumegrid: the number of density values is not nx * ny * nz (volume_grid.dart:95-99)");
      v.densityOffset = (uint32_t)dens.size();
      dens.insert(dens.end(), density + density_offsets[i], density + density_offsets[i + 1]);
    }
  }
  r->volumes.swap(vols);
  r->volDensity.swap(dens);
  r->volV2W.assign(volume_to_world ? volume_to_world : world_to_volume, (volume_to_world ? volume_to_world : world_to_volume) + 16 * (size_t)n);
  if (!volume_to_world && n) return fail(c, DRT_E_INVALID, "null volume_to_world");
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_volume_integrator(drt_ctx* c, int32_t kind, double step_size) {
  DRT_FORWARD_TO_PEERS(c, drt_set_volume_integrator(p_, kind, step_size));
  if (!c) return DRT_E_INVALID;
  if (kind < 0 || kind > 1) return fail(c, DRT_E_INVALID, "volume integrator must be 0 (emission) or 1 (single)");
  if (!(step_size > 0.0)) return fail(c, DRT_E_INVALID, "stepsize must be positive");
  RenderState* r = state(c);
  r->volIntegrator = kind;
  r->volStep = step_size;
  r->sceneTablesValid = false;
  return DRT_OK;
}

int drt_set_shading_precision(drt_ctx* c, int32_t precision) {
  DRT_FORWARD_TO_PEERS(c, drt_set_shading_precision(p_, precision));
  if (!c) return DRT_E_INVALID;
  if (precision != DRT_PRECISION_F64 && precision != DRT_PRECISION_F32) return fail(c, DRT_E_INVALID, "precision must be DRT_PRECISION_F64 or DRT_PRECISION_F32");
  state(c)->shadingPrecision = precision;
  return DRT_OK;
}

int drt_set_camera(drt_ctx* c, const float* raster_to_camera, const float* camera_to_world, double lens_radius,
                   double focal_distance, double shutter_open, double shutter_close) {
  DRT_FORWARD_TO_PEERS(c, drt_set_camera(p_, raster_to_camera, camera_to_world, lens_radius, focal_distance, shutter_open, shutter_close));
  if (!c) return DRT_E_INVALID;
  if (!raster_to_camera || !camera_to_world) return fail(c, DRT_E_INVALID, "null camera matrix");
  RenderState* r = state(c);
  std::memcpy(r->rp.rasterToCamera, raster_to_camera, 64);
  std::memcpy(r->rp.cameraToWorld, camera_to_world, 64);
  r->cameraMoves = false;  // drt_set_camera_motion follows for an animated camera
  r->rp.cameraMotion = nullptr;
  r->rp.lensRadius = lens_radius; r->rp.focalDistance = focal_distance;
  r->rp.shutterOpen = shutter_open; r->rp.shutterClose = shutter_close;
  r->haveCamera = true;
  return DRT_OK;
}

int drt_set_camera_motion(drt_ctx* c, const float* camera_to_world_end, double start_time, double end_time) {
  DRT_FORWARD_TO_PEERS(c, drt_set_camera_motion(p_, camera_to_world_end, start_time, end_time));
  if (!c) return DRT_E_INVALID;
  RenderState* r = state(c);
  if (!r->haveCamera) return fail(c, DRT_E_STATE, "drt_set_camera_motion follows drt_set_camera");
  r->cameraMoves = false;
  r->rp.cameraMotion = nullptr;
  if (!camera_to_world_end) return DRT_OK;
  // AnimatedTransform(cam2world[0], start, cam2world[1], end) (dartray.dart:971-975): only the matrices m are read by the camera
  animInit(&r->cameraMotion, r->rp.cameraToWorld, r->rp.cameraToWorld, camera_to_world_end, camera_to_world_end, start_time, end_time);
  r->cameraMotion.object = -1;
  r->cameraMoves = r->cameraMotion.animated != 0;
  return DRT_OK;
}

int drt_set_camera_kind(drt_ctx* c, int kind) {
  DRT_FORWARD_TO_PEERS(c, drt_set_camera_kind(p_, kind));
  if (!c) return DRT_E_INVALID;
  if (kind < 0 || kind > 2) return fail(c, DRT_E_INVALID, "camera kind must be 0 (perspective), 1 (orthographic) or 2 (environment)");
  state(c)->rp.cameraKind = kind;
  return DRT_OK;
}

int drt_set_film(drt_ctx* c, int xres, int yres, const double* crop, double xwidth, double ywidth, const float* table) {
  DRT_FORWARD_TO_PEERS(c, drt_set_film(p_, xres, yres, crop, xwidth, ywidth, table));
  if (!c) return DRT_E_INVALID;
  if (xres < 1 || yres < 1 || !(xwidth > 0.0) || !(ywidth > 0.0) || !table) return fail(c, DRT_E_INVALID, "bad film parameters");
  RenderState* r = state(c);
  r->rp.xres = xres; r->rp.yres = yres;
  for (int i = 0; i < 4; ++i) r->crop[i] = crop ? crop[i] : ((i & 1) ? 1.0 : 0.0);
  r->rp.xWidth = xwidth; r->rp.yWidth = ywidth;
  std::memcpy(r->table, table, sizeof(r->table));
  configureFilm(r);
  r->haveFilm = true;
  r->filmPixels = 0;  // new film: cleared on the next render / clear
  return DRT_OK;
}

int drt_set_sampler(drt_ctx* c, int kind, int xs, int ys, int spp, int jitter, int pixel_order, int tile_size, uint64_t seed) {
  DRT_FORWARD_TO_PEERS(c, drt_set_sampler(p_, kind, xs, ys, spp, jitter, pixel_order, tile_size, seed));
  if (!c) return DRT_E_INVALID;
  if (kind < 0 || kind > 5)
    return fail(c, DRT_E_INVALID, "sampler kind must be 0 (lowdiscrepancy), 1 (stratified), 2 (random), 3 (halton), 4 (adaptive) or 5 (bestcandidate)");
  if (spp < 1 || xs < 1 || ys < 1) return fail(c, DRT_E_INVALID, "sample counts must be >= 1");
  RenderState* r = state(c);
  r->rp.samplerKind = kind; r->rp.xs = xs; r->rp.ys = ys; r->rp.jitter = jitter; r->rp.seed = seed;
  r->spp = spp; r->pixelOrder = pixel_order; r->tileSize = tile_size;
  return DRT_OK;
}

int drt_set_sample_table(drt_ctx* c, const double* table, uint32_t n_entries) {
  DRT_FORWARD_TO_PEERS(c, drt_set_sample_table(p_, table, n_entries));
  if (!c) return DRT_E_INVALID;
  if (!table || n_entries != 4096) return fail(c, DRT_E_INVALID, "the bestcandidate pattern holds 4096 entries of 5 values (best_candidate_sampler.dart:32-34)");
  state(c)->sampleTable.assign(table, table + 5 * (size_t)n_entries);
  return DRT_OK;
}

int drt_set_integrator(drt_ctx* c, int kind, int maxdepth, int strategy, int ao_nsamples, double ao_mindist, double ao_maxdist) {
  DRT_FORWARD_TO_PEERS(c, drt_set_integrator(p_, kind, maxdepth, strategy, ao_nsamples, ao_mindist, ao_maxdist));
  if (!c) return DRT_E_INVALID;
  if (kind < 0 || kind > 3)
    return fail(c, DRT_E_INVALID, "integrator kind must be 0 (path), 1 (ambientocclusion), 2 (directlighting) or 3 (whitted)");
  if (kind == 1 && ao_nsamples < 1) return fail(c, DRT_E_INVALID, "ambientocclusion nsamples must be >= 1");
  RenderState* r = state(c);
  r->rp.integKind = kind; r->rp.maxDepth = maxdepth; r->rp.strategy = strategy; r->rp.aoSamples = std::max(1, ao_nsamples);
  r->rp.aoMinDist = ao_mindist; r->rp.aoMaxDist = ao_maxdist;
  return DRT_OK;
}

int drt_set_batch_slots(drt_ctx* c, uint64_t slots) {
  DRT_FORWARD_TO_PEERS(c, drt_set_batch_slots(p_, slots));
  if (!c) return DRT_E_INVALID;
  state(c)->batchSlots = slots;
  return DRT_OK;
}

int drt_render(drt_ctx* c, int task_num, int task_count) {
  if (!c) return DRT_E_INVALID;
  if (task_count < 1 || task_num < 0 || task_num >= task_count) return fail(c, DRT_E_INVALID, "bad task_num / task_count");
  RenderState* r = state(c);
  if (!r->haveFilm) return fail(c, DRT_E_STATE, "drt_set_film must be called before rendering");
  int ext[4];
  sampleExtent(r->rp, ext);
  int x = ext[0], y = ext[2], w = ext[1] - ext[0], h = ext[3] - ext[2];
  if (task_count > 1) {  // dartray.dart:1009-1023: the task's sub-window of the SAMPLE extent
    int e[4];
    getSubWindow(w, h, task_num, task_count, e);
    x = ext[0] + e[0]; w = e[1] - e[0];
    y = ext[2] + e[2]; h = e[3] - e[2];
  }
  if (!c->peers.empty()) return renderMulti(c, x, y, w, h, 0, 1);
  return renderWindow(c, x, y, w, h, 0, 1);
}

int drt_render_shard(drt_ctx* c, int shard, int n_shards) {
  if (!c) return DRT_E_INVALID;
  if (n_shards < 1 || shard < 0 || shard >= n_shards) return fail(c, DRT_E_INVALID, "bad shard / n_shards");
  RenderState* r = state(c);
  if (!r->haveFilm) return fail(c, DRT_E_STATE, "drt_set_film must be called before rendering");
  int ext[4];
  sampleExtent(r->rp, ext);
  if (!c->peers.empty()) return renderMulti(c, ext[0], ext[2], ext[1] - ext[0], ext[3] - ext[2], (uint32_t)shard, (uint32_t)n_shards);
  return renderWindow(c, ext[0], ext[2], ext[1] - ext[0], ext[3] - ext[2], (uint32_t)shard, (uint32_t)n_shards);
}

int drt_film_clear(drt_ctx* c) {
  DRT_FORWARD_TO_PEERS(c, drt_film_clear(p_));
  if (!c) return DRT_E_INVALID;
  RenderState* r = state(c);
  if (c->device == DRT_DEVICE_NONE) return fail(c, DRT_E_NODEVICE, kNoDevice);
  if (!r->haveFilm) return fail(c, DRT_E_STATE, "drt_set_film must be called first");
  CK(c, cudaSetDevice(c->device));
  r->filmPixels = 0;
  int rc = ensureFilm(c, r);
  r->stats = drt_render_stats{};
  r->prof = drt_render_profile{};
  return rc;
}

int drt_film_size(const drt_ctx* c, int32_t out[4]) {
  if (!c || !out || !c->render || !c->render->haveFilm) return DRT_E_STATE;
  const RenderParams& p = c->render->rp;
  out[0] = p.left; out[1] = p.top; out[2] = p.width; out[3] = p.height;
  return DRT_OK;
}

int drt_film_device(drt_ctx* c, void** d_film, uint64_t* n_doubles) {
  if (!c || !d_film || !n_doubles) return DRT_E_INVALID;
  RenderState* r = state(c);
  if (c->device == DRT_DEVICE_NONE) return fail(c, DRT_E_NODEVICE, kNoDevice);
  if (!r->haveFilm) return fail(c, DRT_E_STATE, "drt_set_film must be called first");
  CK(c, cudaSetDevice(c->device));
  if ((size_t)r->rp.width * r->rp.height != r->filmPixels) RK(ensureFilm(c, r));
  *d_film = r->dFilm.p;
  *n_doubles = 4ull * r->filmPixels;
  return DRT_OK;
}

int drt_film_read(drt_ctx* c, float* rgb, float* xyz, float* weight) {
  if (!c) return DRT_E_INVALID;
  RenderState* r = state(c);
  if (c->device == DRT_DEVICE_NONE) return fail(c, DRT_E_NODEVICE, kNoDevice);
  if (!r->haveFilm) return fail(c, DRT_E_STATE, "drt_set_film must be called first");
  CK(c, cudaSetDevice(c->device));
  if ((size_t)r->rp.width * r->rp.height != r->filmPixels) RK(ensureFilm(c, r));
  const size_t n = r->filmPixels;
  CK(c, r->dRgb.ensure(3 * n));
  CK(c, r->dXyz.ensure(3 * n));
  CK(c, r->dWeight.ensure(n));
  CK(c, cudaDeviceSynchronize());  // the film may have been summed across GPUs on another stream
  CK(c, STAGE(launchFilmConvert)(r->rp, rgb ? r->dRgb.p : nullptr, xyz ? r->dXyz.p : nullptr, weight ? r->dWeight.p : nullptr, c->stream));
  c->launches++;
  if (rgb) CK(c, cudaMemcpyAsync(rgb, r->dRgb.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (xyz) CK(c, cudaMemcpyAsync(xyz, r->dXyz.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (weight) CK(c, cudaMemcpyAsync(weight, r->dWeight.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return DRT_OK;
}

int drt_pixel_samples(drt_ctx* c, int x, int y, float* out, int cap, int32_t* n_samples, int32_t* floats_per_sample) {
  if (!c || !out) return DRT_E_INVALID;
  RenderState* r = state(c);
  RK(prepare(c, r));
  const RenderParams& p = r->rp;
  const uint32_t n = (uint32_t)p.nPixelSamples;
  RK(ensureWavefront(c, r, n, n));
  PixelBatch pb{x, y, 1, 0, 1, 0, 0, 1, 1024};
  CK(c, STAGE(launchSampler)(p, r->wf, r->dArrays.p, (int)r->arrays.size(), r->maxVals, r->maxOthers, pb, c->numSMs, c->stream)); profMark(c, DRT_PK_SAMPLER);
  c->launches++;
  std::vector<double2> xy(n), lens(n);
  std::vector<double> tm(n);
  std::vector<float> vals((size_t)std::max(p.nVals, 1) * n);
  CK(c, cudaMemcpyAsync(xy.data(), r->wf.camXY, n * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaMemcpyAsync(lens.data(), r->wf.camLens, n * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaMemcpyAsync(tm.data(), r->wf.camTime, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaMemcpyAsync(vals.data(), r->wf.vals, (size_t)p.nVals * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  const int per = 5 + p.nVals;
  if (n_samples) *n_samples = (int32_t)n;
  for (uint32_t i = 0; i < n; ++i) {
    std::vector<float> rec(per);
    rec[0] = (float)(xy[i].x - x); rec[1] = (float)(xy[i].y - y);
    rec[2] = (float)lens[i].x; rec[3] = (float)lens[i].y;
    rec[4] = (float)((1.0 - tm[i]) * p.shutterOpen + tm[i] * p.shutterClose);
    for (int v = 0; v < p.nVals; ++v) rec[5 + v] = vals[(size_t)v * n + i];
    for (int k = 0; k < per; ++k)
      if ((size_t)i * per + k < (size_t)cap) out[(size_t)i * per + k] = rec[k];
  }
  if (floats_per_sample) *floats_per_sample = per;
  return DRT_OK;
}

int drt_set_render_profiling(drt_ctx* c, int flags) {
  if (!c) return DRT_E_INVALID;
  if (flags & ~(DRT_PROFILE_TIME | DRT_PROFILE_WORK)) return fail(c, DRT_E_INVALID, "unknown profiling flags");
  state(c)->profFlags = flags;
  return DRT_OK;
}

int drt_render_profile_get(drt_ctx* c, drt_render_profile* out) {
  if (!c || !out) return DRT_E_INVALID;
  *out = state(c)->prof;
  return DRT_OK;
}

int drt_render_stats_get(drt_ctx* c, drt_render_stats* out) {
  if (!c || !out) return DRT_E_INVALID;
  *out = state(c)->stats;
  for (drt_ctx* p : c->peers) {  // a multi-device context reports the rays of all its devices
    const drt_render_stats& s = state(p)->stats;
    out->camera_samples += s.camera_samples; out->closest_rays += s.closest_rays;
    out->shadow_rays += s.shadow_rays; out->zeroed_samples += s.zeroed_samples;
  }
  return DRT_OK;
}

}  // extern "C"
