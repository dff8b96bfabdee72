// Participating media of the render path, in the reference's numeric model (float32 Point / Vector / Spectrum storage,
// binary64 expressions): the VolumeRegion plugins (lib/volume_regions/{homogenous_volume_region,exponential_density_region,
// volume_grid}.dart over lib/core/volume/{density_region,aggregate_volume,volume}.dart) and the transmittance() both volume
// integrators share (lib/volume_integrators/emission_integrator.dart:85-105, single_scatter_integrator.dart:26-45).
// Only compiled into the `extra` build of the stage kernels (DRT_EXTRA): scenes without a Volume statement never see it.
#pragma once
#include "shade_device.cuh"

namespace drt {

struct VRay {  // Ray (ray.dart:27-75): float32 origin / direction, f64 interval
  V3 o, d;
  double mint, maxt;
};
static __device__ inline V3 VRayAt(const VRay& r, double t) { return RayAt(r.o, r.d, t); }

// BBox.intersectP (bbox.dart:81-114)
static __device__ inline bool volBoxIntersectP(const float* lo, const float* hi, const VRay& ray, double* hitt0, double* hitt1) {
  double t0 = ray.mint, t1 = ray.maxt;
  const float o[3] = {ray.o.x, ray.o.y, ray.o.z}, d[3] = {ray.d.x, ray.d.y, ray.d.z};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double invRayDir = 1.0 / (double)d[i];
    double tNear = ((double)lo[i] - (double)o[i]) * invRayDir;
    double tFar = ((double)hi[i] - (double)o[i]) * invRayDir;
    if (tNear > tFar) { const double t = tNear; tNear = tFar; tFar = t; }
    t0 = tNear > t0 ? tNear : t0;
    t1 = tFar < t1 ? tFar : t1;
    if (t0 > t1) return false;
  }
  *hitt0 = t0;
  *hitt1 = t1;
  return true;
}
static __device__ inline bool volBoxInside(const float* lo, const float* hi, const V3& pt) {  // bbox.dart:123-127
  return pt.x >= lo[0] && pt.x <= hi[0] && pt.y >= lo[1] && pt.y <= hi[1] && pt.z >= lo[2] && pt.z <= hi[2];
}
static __device__ inline bool regionIntersectP(const GVolume& v, const VRay& r, double* t0, double* t1) {
  VRay ray;  // worldToVolume.transformRay (transform.dart:180-195)
  ray.o = XfPoint(v.w2v, r.o);
  ray.d = XfVector(v.w2v, r.d);
  ray.mint = r.mint;
  ray.maxt = r.maxt;
  return volBoxIntersectP(v.lo, v.hi, ray, t0, t1);
}
static __device__ inline double regionDensity(const RenderScene& rs, const GVolume& v, const V3& Pobj) {
  if (!volBoxInside(v.lo, v.hi, Pobj)) return 0.0;
  if (v.kind == 0) return 1.0;
  if (v.kind == 1) {  // exponential_density_region.dart:42-50
    const V3 rel = Pobj - V3{v.lo[0], v.lo[1], v.lo[2]};
    const double height = Dot(rel, V3{v.up[0], v.up[1], v.up[2]});
    return v.a * exp(-v.b * height);
  }
  // volume_grid.dart:39-66; extent.offset() is a float32 Vector (bbox.dart:193-197), scaled in place
  V3 vox = mkv(((double)Pobj.x - v.lo[0]) / ((double)v.hi[0] - v.lo[0]), ((double)Pobj.y - v.lo[1]) / ((double)v.hi[1] - v.lo[1]),
               ((double)Pobj.z - v.lo[2]) / ((double)v.hi[2] - v.lo[2]));
  vox.x = (float)((double)vox.x * v.nx - 0.5);
  vox.y = (float)((double)vox.y * v.ny - 0.5);
  vox.z = (float)((double)vox.z * v.nz - 0.5);
  const int vx = (int)floor((double)vox.x), vy = (int)floor((double)vox.y), vz = (int)floor((double)vox.z);
  const double dx = (double)vox.x - vx, dy = (double)vox.y - vy, dz = (double)vox.z - vz;
  const auto* dens = rs.volDensity + v.densityOffset;  // binary64 storage in either build
  auto D = [&](int x, int y, int z) {
    x = min(max(x, 0), v.nx - 1);
    y = min(max(y, 0), v.ny - 1);
    z = min(max(z, 0), v.nz - 1);
    return __ldg(dens + ((size_t)z * v.nx * v.ny + (size_t)y * v.nx + x));
  };
  const double d00 = LerpD(dx, D(vx, vy, vz), D(vx + 1, vy, vz));
  const double d10 = LerpD(dx, D(vx, vy + 1, vz), D(vx + 1, vy + 1, vz));
  const double d01 = LerpD(dx, D(vx, vy, vz + 1), D(vx + 1, vy, vz + 1));
  const double d11 = LerpD(dx, D(vx, vy + 1, vz + 1), D(vx + 1, vy + 1, vz + 1));
  const double d0 = LerpD(dy, d00, d10), d1 = LerpD(dy, d01, d11);
  return LerpD(dz, d0, d1);
}
// sigma_a / sigma_s / sigma_t / Lve at a world point (homogenous_volume_region.dart:37-56, density_region.dart:33-47)
enum { VOL_SIG_A = 0, VOL_SIG_S = 1, VOL_SIG_T = 2, VOL_LVE = 3 };
static __device__ inline Spec regionCoeff(const RenderScene& rs, const GVolume& v, const V3& p, int which) {
  const float* c = which == VOL_SIG_A ? v.sigA : which == VOL_SIG_S ? v.sigS : which == VOL_SIG_T ? v.sigT : v.le;
  const Spec s = Spec{c[0], c[1], c[2]};
  const V3 q = XfPoint(v.w2v, p);
  if (v.kind == 0) return volBoxInside(v.lo, v.hi, q) ? s : mks1(0.0);
  return s * regionDensity(rs, v, q);
}
static __device__ inline double regionPhase(const GVolume& v, const V3& p, const V3& w, const V3& wp) {
  if (v.kind == 0 && !volBoxInside(v.lo, v.hi, XfPoint(v.w2v, p))) return 0.0;  // homogenous_volume_region.dart:58-63
  const double costheta = Dot(w, wp);  // PhaseHG, volume.dart:84-88
  return 1.0 / (4.0 * DRT_PI) * (1.0 - v.g * v.g) / pow(1.0 + v.g * v.g - 2.0 * v.g * costheta, 1.5);
}
static __device__ inline Spec regionTau(const RenderScene& rs, const GVolume& v, const VRay& r, double stepSize, double u) {
  double t0 = 0.0, t1 = 0.0;
  if (v.kind == 0) {  // homogenous_volume_region.dart:65-73: analytic
    if (!regionIntersectP(v, r, &t0, &t1)) return mks1(0.0);
    return Spec{v.sigT[0], v.sigT[1], v.sigT[2]} * Distance(VRayAt(r, t0), VRayAt(r, t1));
  }
  const double length = Length(r.d);  // density_region.dart:53-77
  if (length == 0.0) return mks1(0.0);
  VRay rn;
  rn.o = r.o;
  rn.d = r.d / length;
  rn.mint = r.mint * length;
  rn.maxt = r.maxt * length;
  if (!regionIntersectP(v, rn, &t0, &t1)) return mks1(0.0);
  Spec tau = mks1(0.0);
  t0 += u * stepSize;
  while (t0 < t1) {
    tau = tau + regionCoeff(rs, v, VRayAt(rn, t0), VOL_SIG_T);
    t0 += stepSize;
  }
  return tau * stepSize;
}
// scene.volumeRegion: the one region, or the AggregateVolume over all of them (aggregate_volume.dart:23-103)
static __device__ inline bool volIntersectP(const RenderScene& rs, const VRay& ray, double* t0, double* t1) {
  if (rs.nVolumes == 1) return regionIntersectP(rs.volumes[0], ray, t0, t1);
  *t0 = CUDART_INF;
  *t1 = -CUDART_INF;
  for (int i = 0; i < rs.nVolumes; ++i) {
    double a = 0.0, b = 0.0;
    if (regionIntersectP(rs.volumes[i], ray, &a, &b)) { *t0 = dartMin(*t0, a); *t1 = dartMax(*t1, b); }
  }
  return *t0 < *t1;
}
static __device__ inline Spec volCoeff(const RenderScene& rs, const V3& p, int which) {
  if (rs.nVolumes == 1) return regionCoeff(rs, rs.volumes[0], p, which);
  Spec s = mks1(0.0);
  for (int i = 0; i < rs.nVolumes; ++i) s = s + regionCoeff(rs, rs.volumes[i], p, which);
  return s;
}
static __device__ inline double volPhase(const RenderScene& rs, const V3& p, const V3& w, const V3& wp) {
  if (rs.nVolumes == 1) return regionPhase(rs.volumes[0], p, w, wp);
  double ph = 0.0, sumWt = 0.0;  // aggregate_volume.dart:71-80
  for (int i = 0; i < rs.nVolumes; ++i) {
    const double wt = Luminance(regionCoeff(rs, rs.volumes[i], p, VOL_SIG_S));
    sumWt += wt;
    ph += wt * regionPhase(rs.volumes[i], p, w, wp);
  }
  return ph / sumWt;
}
static __device__ inline Spec volTau(const RenderScene& rs, const VRay& ray, double step, double offset) {
  if (rs.nVolumes == 1) return regionTau(rs, rs.volumes[0], ray, step, offset);
  Spec t = mks1(0.0);
  for (int i = 0; i < rs.nVolumes; ++i) t = t + regionTau(rs, rs.volumes[i], ray, step, offset);
  return t;
}
static __device__ inline Spec expNeg(const Spec& tau) { return mks(exp(-(double)tau.r), exp(-(double)tau.g), exp(-(double)tau.b)); }

// VolumeIntegrator.transmittance without a Sample (step = 4 x stepSize, one draw for the offset): the calls the surface
// integrators make (integrator.dart:137,178, path_integrator.dart:116).  The draw comes from the camera sample's transmittance
// stream (DRT_STREAM_TRANSMITTANCE), whose position is kept per slot.  Out of line: rare next to the surface shading.
static __device__ __noinline__ void volTransmittanceDrawCold(const RenderScene& rs, uint64_t key, uint32_t* ctr, V3 o, V3 d, double mint,
                                                             double maxt, Spec* out) {
  const uint32_t k = ++*ctr;
  const double offset = drawFloat(key, k);
  VRay ray{o, d, mint, maxt};
  *out = expNeg(volTau(rs, ray, 4.0 * rs.volStep, offset));
}
// ... with the Sample (whitted_integrator.dart:56-58): step = stepSize, offset = the tau sample, no draw
static __device__ __noinline__ void volTransmittanceSampleCold(const RenderScene& rs, double tauSample, V3 o, V3 d, double mint, double maxt,
                                                               Spec* out) {
  VRay ray{o, d, mint, maxt};
  *out = expNeg(volTau(rs, ray, rs.volStep, tauSample));
}

}  // namespace drt
