// What DartRay evaluates at a hit point before it can build the BSDF of a material whose parameters are textures that read the
// hit point, or that carries a bump map — the texture pass of the wavefront (texture_kernels.cu):
//   DifferentialGeometry.computeDifferentials   lib/core/differential_geometry.dart:122-205
//   camera ray differentials                    lib/cameras/perspective_camera.dart:50-56,122-128, orthographic_camera.dart:111-115,
//                                               lib/core/camera.dart:37-62, ray_differential.dart:56-61 (scaleDifferentials)
//   full dg of every shape (u, v, dndu, dndv)   lib/shapes/{triangle,sphere,disk,cylinder,cone,paraboloid,hyperboloid}.dart
//   Triangle.getShadingGeometry                 lib/shapes/triangle.dart:271-364
//   MIPMap.lookup2 / lookup / EWA / triangle    lib/core/mipmap.dart:183-355
//   the texture mappings and textures           lib/core/texture/*.dart, lib/textures/*.dart (every texture plugin), Noise / FBm /
//                                               Turbulence lib/core/texture.dart:40-140
//   Material.Bump                               lib/core/material.dart:35-88
//   the materials' getBSDF                      lib/materials/*.dart
// Arithmetic model as everywhere in the shading code (shade_device.cuh header): float32 objects, binary64 expressions.
#pragma once
#include "shade_device.cuh"

namespace drt {

struct FullDG {  // differential_geometry.dart:27-42
  V3 p, nn, dpdu, dpdv, dndu, dndv, dpdx, dpdy;
  double u, v, dudx, dvdx, dudy, dvdy;
  bool reverse;
};

struct RayDiffs {  // ray_differential.dart:27-33, world space
  bool has;
  V3 rxo, ryo, rxd, ryd;
};

static __device__ inline double Log2d(double x) { return log(x) * (1.0 / 0.6931471805599453); }  // common.dart:98-103 (1 / Math.log(2))
static __device__ inline long long dartModLL(long long a, long long n) { return ((a % n) + n) % n; }

// ---- camera ray differentials, regenerated from the slot's camera sample -----------------------------------------------
static __device__ inline void cameraGenerate(const RenderParams& rp, double imageX, double imageY, double lensU, double lensV, V3* oC, V3* dC,
                                             V3* pCam) {
  // camera space: the ray before cameraToWorld (perspective_camera.dart:59-91, orthographic_camera.dart:52-80)
  const V3 Pcamera = XfPoint(rp.rasterToCamera, mkv(imageX, imageY, 0.0));
  V3 o = V3{0.f, 0.f, 0.f}, d = Normalize(Pcamera);
  if (rp.cameraKind == 1) { o = Pcamera; d = V3{0.f, 0.f, 1.f}; }
  if (rp.lensRadius > 0.0) {
    double lu, lv;
    ConcentricSampleDisk(lensU, lensV, &lu, &lv);
    lu *= rp.lensRadius;
    lv *= rp.lensRadius;
    const double ft = rp.focalDistance / d.z;
    const V3 Pfocus = RayAt(o, d, ft);
    o = mkv(lu, lv, 0.0);
    d = Normalize(Pfocus - o);
  }
  *oC = o; *dC = d; *pCam = Pcamera;
}
static __device__ inline void environmentRay(const RenderParams& rp, const float* c2w, double imageX, double imageY, V3* o, V3* d) {  // environment_camera.dart:42-52
  const double theta = DRT_PI * imageY / rp.yres, phi = 2 * DRT_PI * imageX / rp.xres;
  *o = XfPoint(c2w, V3{0.f, 0.f, 0.f});
  *d = XfVector(c2w, mkv(sin(theta) * cos(phi), cos(theta), sin(theta) * sin(phi)));
}
// o, d: the world-space camera ray (as raygenKernel stored it); scale = 1 / sqrt(samplesPerPixel) (sampler_renderer.dart:166)
// `time`: the camera sample's time, for a moving camera (cameraToWorld.interpolate(time), animated_transform.dart:158-169)
static __device__ __noinline__ void cameraDifferentialsCold(const RenderParams& rp, double imageX, double imageY, double lensU, double lensV,
                                                            V3 o, V3 d, double scale, double time, RayDiffs* out) {
  RayDiffs r;
  r.has = true;
  const float* c2w = rp.cameraToWorld;
  M4 camM, camInv;
  if (rp.cameraMotion) {
    animInterpolate(*rp.cameraMotion, time, &camM, &camInv);
    c2w = camM.d;
  }
  if (rp.cameraKind == 2) {  // camera.dart:40-58: imageX++, then imageX--, imageY++
    environmentRay(rp, c2w, imageX + 1.0, imageY, &r.rxo, &r.rxd);
    environmentRay(rp, c2w, (imageX + 1.0) - 1.0, imageY + 1.0, &r.ryo, &r.ryd);
  } else {
    V3 oC, dC, pCam;
    cameraGenerate(rp, imageX, imageY, lensU, lensV, &oC, &dC, &pCam);
    if (rp.cameraKind == 0) {  // perspective_camera.dart:50-56,122-128
      const V3 p0 = XfPoint(rp.rasterToCamera, V3{0.f, 0.f, 0.f});
      const V3 dxCamera = XfPoint(rp.rasterToCamera, V3{1.f, 0.f, 0.f}) - p0, dyCamera = XfPoint(rp.rasterToCamera, V3{0.f, 1.f, 0.f}) - p0;
      r.rxo = XfPoint(c2w, oC);
      r.ryo = r.rxo;
      r.rxd = XfVector(c2w, Normalize(pCam + dxCamera));
      r.ryd = XfVector(c2w, Normalize(pCam + dyCamera));
    } else {
      // orthographic_camera.dart:111-115 AS WRITTEN: the offset origins are built in camera space and transformRay (not
      // transformRayDifferential) follows, so they stay there; rxDirection / ryDirection are the object ray.direction is, which
      // transformRay overwrites in place: the world-space direction.
      const V3 dxCamera = XfVector(rp.rasterToCamera, V3{1.f, 0.f, 0.f}), dyCamera = XfVector(rp.rasterToCamera, V3{0.f, 1.f, 0.f});
      r.rxo = oC + dxCamera;
      r.ryo = oC + dyCamera;
      r.rxd = d;
      r.ryd = d;
    }
  }
  // scaleDifferentials (ray_differential.dart:56-61)
  r.rxo = o + (r.rxo - o) * scale;
  r.ryo = o + (r.ryo - o) * scale;
  r.rxd = d + (r.rxd - d) * scale;
  r.ryd = d + (r.ryd - d) * scale;
  *out = r;
}

// ---- DifferentialGeometry.computeDifferentials (:122-205) --------------------------------------------------------------
static __device__ inline double comp(const V3& v, int i) { return i == 0 ? (double)v.x : (i == 1 ? (double)v.y : (double)v.z); }
static __device__ inline bool solve2x2(const double A[4], const double B[2], double* x0, double* x1) {  // common.dart:170-185
  const double det = A[0] * A[3] - A[1] * A[2];
  if (fabs(det) < 1.0e-10) return false;
  *x0 = (A[3] * B[0] - A[1] * B[1]) / det;
  *x1 = (A[0] * B[1] - A[2] * B[0]) / det;
  if (isnan(*x0) || isnan(*x1)) return false;
  return true;
}
static __device__ inline void computeDifferentials(FullDG* dg, const RayDiffs& rd) {
  dg->dudx = dg->dvdx = dg->dudy = dg->dvdy = 0.0;
  dg->dpdx = dg->dpdy = V3{0.f, 0.f, 0.f};
  if (!rd.has) return;
  const V3 nn = dg->nn, p = dg->p;
  const double d = -Dot(nn, p);
  const double tx = -(Dot(nn, rd.rxo) + d) / Dot(nn, rd.rxd);
  if (isnan(tx)) return;
  const V3 px = rd.rxo + rd.rxd * tx;
  const double ty = -(Dot(nn, rd.ryo) + d) / Dot(nn, rd.ryd);
  if (isnan(ty)) return;
  const V3 py = rd.ryo + rd.ryd * ty;
  dg->dpdx = px - p;
  dg->dpdy = py - p;
  int a0, a1;
  if (fabs((double)nn.x) > fabs((double)nn.y) && fabs((double)nn.x) > fabs((double)nn.z)) { a0 = 1; a1 = 2; }
  else if (fabs((double)nn.y) > fabs((double)nn.z)) { a0 = 0; a1 = 2; }
  else { a0 = 0; a1 = 1; }
  const double A[4] = {comp(dg->dpdu, a0), comp(dg->dpdv, a0), comp(dg->dpdu, a1), comp(dg->dpdv, a1)};
  const double Bx[2] = {comp(px, a0) - comp(p, a0), comp(px, a1) - comp(p, a1)};
  const double By[2] = {comp(py, a0) - comp(p, a0), comp(py, a1) - comp(p, a1)};
  double du, dv;
  if (solve2x2(A, Bx, &du, &dv)) { dg->dudx = du; dg->dvdx = dv; }
  if (solve2x2(A, By, &du, &dv)) { dg->dudy = du; dg->dvdy = dv; }
}

// ---- the whole DifferentialGeometry of a hit (Shape.intersect's dg.set) ---------------------------------------------------
// dndu / dndv from the fundamental forms (sphere.dart:138-153 and the same block in the other quadrics).  perFactor: the sphere
// multiplies dpdu by (f F - e G) and then by invEGF2 (two float32 Vectors), the others by their product.
static __device__ inline void weingarten(const V3& dpdu, const V3& dpdv, const V3& d2Pduu, const V3& d2Pduv, const V3& d2Pdvv, bool perFactor,
                                         V3* dndu, V3* dndv) {
  const double E = Dot(dpdu, dpdu), F = Dot(dpdu, dpdv), G = Dot(dpdv, dpdv);
  const V3 N = Normalize(Cross(dpdu, dpdv));
  const double e = Dot(N, d2Pduu), f = Dot(N, d2Pduv), g = Dot(N, d2Pdvv);
  const double invEGF2 = 1.0 / (E * G - F * F);
  if (perFactor) {
    *dndu = dpdu * (f * F - e * G) * invEGF2 + dpdv * (e * F - f * E) * invEGF2;
    *dndv = dpdu * (g * F - f * G) * invEGF2 + dpdv * (f * F - g * E) * invEGF2;
  } else {
    *dndu = dpdu * ((f * F - e * G) * invEGF2) + dpdv * ((e * F - f * E) * invEGF2);
    *dndv = dpdu * ((g * F - f * G) * invEGF2) + dpdv * ((f * F - g * E) * invEGF2);
  }
}
static __device__ inline V3 xfNormalRows(const float* w2o, const V3& n) {  // transformNormal: transpose of worldToObject's 3x3 (rows of 4)
  return mkv((double)w2o[0] * n.x + (double)w2o[4] * n.y + (double)w2o[8] * n.z, (double)w2o[1] * n.x + (double)w2o[5] * n.y + (double)w2o[9] * n.z,
             (double)w2o[2] * n.x + (double)w2o[6] * n.y + (double)w2o[10] * n.z);
}

// dg (geometric) and dgs (Shape.getShadingGeometry) of the hit of ray (o, d) at tHit on primitive prim
static __device__ __noinline__ void fullGeometryCold(const RenderScene& rs, uint32_t prim, V3 o, V3 d, double t, FullDG* dgOut, FullDG* dgsOut) {
  FullDG dg;
  dg.dudx = dg.dvdx = dg.dudy = dg.dvdy = 0.0;
  dg.dpdx = dg.dpdy = dg.dndu = dg.dndv = V3{0.f, 0.f, 0.f};
  dg.reverse = primReverse(rs, prim);
  bool shadingDiffers = false;
  FullDG dgs;
  if (prim < rs.ntris) {
    const TriVerts tv = loadTri(rs, prim);
    double uv[6] = {0.0, 0.0, 1.0, 0.0, 1.0, 1.0};
    const GMesh* mesh = nullptr;
    uint32_t i0 = 0, i1 = 0, i2 = 0;
    if (rs.meshOfTri) {
      mesh = &rs.meshes[__ldg(rs.meshOfTri + prim)];
      i0 = __ldg(rs.triIdx + 3 * (size_t)prim); i1 = __ldg(rs.triIdx + 3 * (size_t)prim + 1); i2 = __ldg(rs.triIdx + 3 * (size_t)prim + 2);
      meshTriUVs(rs, *mesh, i0, i1, i2, uv);
    }
    // b1, b2 of Triangle.intersect (triangle.dart:52-95), from the ray that found the hit
    const double p1x = tv.p1.x, p1y = tv.p1.y, p1z = tv.p1.z;
    const double e1x = (double)tv.p2.x - p1x, e1y = (double)tv.p2.y - p1y, e1z = (double)tv.p2.z - p1z;
    const double e2x = (double)tv.p3.x - p1x, e2y = (double)tv.p3.y - p1y, e2z = (double)tv.p3.z - p1z;
    const double dx = d.x, dy = d.y, dz = d.z;
    const double s1x = (dy * e2z) - (dz * e2y), s1y = (dz * e2x) - (dx * e2z), s1z = (dx * e2y) - (dy * e2x);
    const double invDivisor = 1.0 / ((s1x * e1x) + (s1y * e1y) + (s1z * e1z));
    const double sx = (double)o.x - p1x, sy = (double)o.y - p1y, sz = (double)o.z - p1z;
    const double b1 = (sx * s1x + sy * s1y + sz * s1z) * invDivisor;
    const double s2x = (sy * e1z) - (sz * e1y), s2y = (sz * e1x) - (sx * e1z), s2z = (sx * e1y) - (sy * e1x);
    const double b2 = ((dx * s2x) + (dy * s2y) + (dz * s2z)) * invDivisor;
    const double b0 = 1.0 - b1 - b2;
    triPartialsUV(tv, uv, &dg.dpdu, &dg.dpdv);
    dg.u = b0 * uv[0] + b1 * uv[2] + b2 * uv[4];
    dg.v = b0 * uv[1] + b1 * uv[3] + b2 * uv[5];
    dg.p = RayAt(o, d, t);
    dg.nn = shapeNormal(dg.dpdu, dg.dpdv, dg.reverse);
    if (mesh && (mesh->flags & 3u)) {  // Triangle.getShadingGeometry, :271-364
      shadingDiffers = true;
      const double A0 = uv[2] - uv[0], A1 = uv[4] - uv[0], A2 = uv[3] - uv[1], A3 = uv[5] - uv[1];
      const double C0 = dg.u - uv[0], C1 = dg.v - uv[1];
      const double det = A0 * A3 - A1 * A2;
      double bx, by = 0.0, bz = 0.0;
      bool ok = !(fabs(det) < 1.0e-10);
      if (ok) {
        by = (A3 * C0 - A1 * C1) / det;
        bz = (A0 * C1 - A2 * C0) / det;
        if (isnan(by) || isnan(bz)) ok = false;
      }
      if (!ok) bx = by = bz = 1.0 / 3.0;
      else bx = 1.0 - by - bz;
      auto vert = [](const float* a, uint32_t i) { return V3{a[3 * (size_t)i], a[3 * (size_t)i + 1], a[3 * (size_t)i + 2]}; };
      const float* m = mesh->o2w;
      const float* w = mesh->w2o;  // 3 x 3, rows of 3
      auto xfN = [&](const V3& n) {
        return mkv((double)w[0] * n.x + (double)w[3] * n.y + (double)w[6] * n.z, (double)w[1] * n.x + (double)w[4] * n.y + (double)w[7] * n.z,
                   (double)w[2] * n.x + (double)w[5] * n.y + (double)w[8] * n.z);
      };
      V3 ns, ss, ts;
      if (mesh->flags & 1u) ns = Normalize(xfN(((vert(rs.vertN, i0) * bx) + (vert(rs.vertN, i1) * by)) + (vert(rs.vertN, i2) * bz)));
      else ns = dg.nn;
      if (mesh->flags & 2u) {
        const V3 si = ((vert(rs.vertS, i0) * bx) + (vert(rs.vertS, i1) * by)) + (vert(rs.vertS, i2) * bz);
        ss = Normalize(mkv((double)m[0] * si.x + (double)m[1] * si.y + (double)m[2] * si.z, (double)m[3] * si.x + (double)m[4] * si.y + (double)m[5] * si.z,
                           (double)m[6] * si.x + (double)m[7] * si.y + (double)m[8] * si.z));
      } else {
        ss = Normalize(dg.dpdu);
      }
      ts = Cross(ss, ns);
      if (LengthSquared(ts) > 0.0) {
        ts = Normalize(ts);
        ss = Cross(ts, ns);
      } else {
        CoordinateSystem(ns, &ss, &ts);
      }
      V3 dndu = V3{0.f, 0.f, 0.f}, dndv = V3{0.f, 0.f, 0.f};  // :328-351
      if (mesh->flags & 1u) {
        const double du1 = uv[0] - uv[4], du2 = uv[2] - uv[4], dv1 = uv[1] - uv[5], dv2 = uv[3] - uv[5];
        const V3 dn1 = vert(rs.vertN, i0) - vert(rs.vertN, i2), dn2 = vert(rs.vertN, i1) - vert(rs.vertN, i2);
        const double determinant = du1 * dv2 - dv1 * du2;
        if (determinant != 0.0) {
          const double invdet = 1.0 / determinant;
          dndu = (dn1 * dv2 - dn2 * dv1) * invdet;
          dndv = (dn1 * -du2 + dn2 * du1) * invdet;
        }
      }
      dgs = dg;
      dgs.dpdu = ss;
      dgs.dpdv = ts;
      dgs.dndu = xfN(dndu);
      dgs.dndv = xfN(dndv);
      dgs.nn = shapeNormal(ss, ts, dg.reverse);
    }
  } else {
    const GSphere& s = rs.ts.spheres[prim - rs.ntris];
    float w2o[16], o2w[16];
    for (int i = 0; i < 12; ++i) { w2o[i] = s.w2o[i]; o2w[i] = s.o2w[i]; }
    for (int i = 0; i < 4; ++i) { w2o[12 + i] = s.w2oRow3[i]; o2w[12 + i] = s.o2wRow3[i]; }
    const V3 ro = XfPoint(w2o, o), rdir = XfVector(w2o, d);
    V3 phit = RayAt(ro, rdir, t);
    V3 dpdu, dpdv, dndu = V3{0.f, 0.f, 0.f}, dndv = V3{0.f, 0.f, 0.f};
    double u, v;
    if (s.shape == 0) {  // sphere.dart:62-160
      if (phit.x == 0.0f && phit.y == 0.0f) phit.x = (float)(1.0e-5 * s.radius);
      double phi = atan2((double)phit.y, (double)phit.x);
      if (phi < 0.0) phi += 2.0 * DRT_PI;
      u = phi / s.phiMax;
      const double theta = acos(clampD((double)phit.z / s.radius, -1.0, 1.0));
      v = (theta - s.thetaMin) / (s.thetaMax - s.thetaMin);
      const double zradius = sqrt((double)phit.x * phit.x + (double)phit.y * phit.y);
      const double invzradius = 1.0 / zradius;
      const double cosphi = phit.x * invzradius, sinphi = phit.y * invzradius;
      dpdu = mkv(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);
      dpdv = mkv(phit.z * cosphi, phit.z * sinphi, -s.radius * sin(theta)) * (s.thetaMax - s.thetaMin);
      const V3 d2Pduu = mkv(phit.x, phit.y, 0.0) * -s.phiMax * s.phiMax;
      const V3 d2Pduv = mkv(-sinphi, cosphi, 0.0) * (s.thetaMax - s.thetaMin) * phit.z * s.phiMax;
      const V3 d2Pdvv = mkv(phit.x, phit.y, phit.z) * -(s.thetaMax - s.thetaMin) * (s.thetaMax - s.thetaMin);
      weingarten(dpdu, dpdv, d2Pduu, d2Pduv, d2Pdvv, true, &dndu, &dndv);
    } else if (s.shape == 1) {  // disk.dart:60-97
      double phi = atan2((double)phit.y, (double)phit.x);
      if (phi < 0.0) phi += 2.0 * DRT_PI;
      u = phi / s.phiMax;
      const double dist2 = (double)phit.x * phit.x + (double)phit.y * phit.y;
      const double oneMinusV = (sqrt(dist2) - s.innerRadius) / (s.radius - s.innerRadius);
      const double invOneMinusV = (oneMinusV > 0.0) ? (1.0 / oneMinusV) : 0.0;
      v = 1.0 - oneMinusV;
      dpdu = mkv(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);
      dpdv = mkv(-(double)phit.x * invOneMinusV, -(double)phit.y * invOneMinusV, 0.0);
      dpdu = dpdu * (s.phiMax * DRT_INV_TWOPI);
      dpdv = dpdv * ((s.radius - s.innerRadius) / s.radius);
    } else {  // cylinder.dart:106-135, cone.dart:101-131, paraboloid.dart:102-137, hyperboloid.dart:125-157
      double vh = 0.0;
      double ay = phit.y, ax = phit.x;
      if (s.shape == 5) {  // hyperboloid.dart:96-103
        const V3 hp1 = V3{s.hp1[0], s.hp1[1], s.hp1[2]}, hp2 = V3{s.hp2[0], s.hp2[1], s.hp2[2]};
        vh = ((double)phit.z - hp1.z) / ((double)hp2.z - hp1.z);
        const V3 pr = (hp1 * (1.0 - vh)) + (hp2 * vh);
        ay = (double)pr.x * phit.y - (double)phit.x * pr.y;
        ax = (double)phit.x * pr.x + (double)phit.y * pr.y;
      }
      double phi = atan2(ay, ax);
      if (phi < 0.0) phi += 2.0 * DRT_PI;
      u = phi / s.phiMax;
      dpdu = mkv(-s.phiMax * phit.y, s.phiMax * phit.x, 0.0);
      const V3 d2Pduu = mkv(phit.x, phit.y, 0.0) * (-s.phiMax * s.phiMax);
      V3 d2Pduv = V3{0.f, 0.f, 0.f}, d2Pdvv = V3{0.f, 0.f, 0.f};
      if (s.shape == 2) {
        v = ((double)phit.z - s.zmin) / (s.zmax - s.zmin);
        dpdv = mkv(0.0, 0.0, s.zmax - s.zmin);
      } else if (s.shape == 3) {
        v = (double)phit.z / s.height;
        dpdv = mkv(-(double)phit.x / (1.0 - v), -(double)phit.y / (1.0 - v), s.height);
        d2Pduv = mkv(phit.y, -(double)phit.x, 0.0) * (s.phiMax / (1.0 - v));
      } else if (s.shape == 4) {
        v = ((double)phit.z - s.zmin) / (s.zmax - s.zmin);
        dpdv = mkv((double)phit.x / (2.0 * phit.z), (double)phit.y / (2.0 * phit.z), 1.0) * (s.zmax - s.zmin);
        d2Pduv = mkv(-(double)phit.y / (2.0 * phit.z), (double)phit.x / (2.0 * phit.z), 0.0) * (s.zmax - s.zmin) * s.phiMax;
        d2Pdvv = mkv((double)phit.x / (4.0 * phit.z * phit.z), (double)phit.y / (4.0 * phit.z * phit.z), 0.0) * (-(s.zmax - s.zmin) * (s.zmax - s.zmin));
      } else {
        v = vh;
        const double cosphi = cos(phi), sinphi = sin(phi);
        const double ex = (double)s.hp2[0] - (double)s.hp1[0], ey = (double)s.hp2[1] - (double)s.hp1[1];
        dpdv = mkv(ex * cosphi - ey * sinphi, ex * sinphi + ey * cosphi, (double)s.hp2[2] - (double)s.hp1[2]);
        d2Pduv = mkv(-(double)dpdv.y, dpdv.x, 0.0) * s.phiMax;
      }
      weingarten(dpdu, dpdv, d2Pduu, d2Pduv, d2Pdvv, false, &dndu, &dndv);
    }
    dg.u = u; dg.v = v;
    dg.p = XfPoint(o2w, phit);
    dg.dpdu = XfVector(o2w, dpdu);
    dg.dpdv = XfVector(o2w, dpdv);
    dg.dndu = xfNormalRows(w2o, dndu);
    dg.dndv = xfNormalRows(w2o, dndv);
    dg.nn = shapeNormal(dg.dpdu, dg.dpdv, dg.reverse);
  }
  *dgOut = dg;
  *dgsOut = shadingDiffers ? dgs : dg;
}

// ---- MIPMap ----------------------------------------------------------------------------------------------------------------
struct TexCtx {
  const GTex* nodes;
  const float* data;
};

static __device__ inline int levelW(const GTex& t, int l) { return max(1, t.w >> l); }
static __device__ inline int levelH(const GTex& t, int l) { return max(1, t.h >> l); }
// texel(), mipmap.dart:183-204; returns false for TEXTURE_BLACK's outside
static __device__ inline bool texelAddr(const GTex& t, int level, long long s, long long tt, size_t* idx) {
  const long long W = levelW(t, level), H = levelH(t, level);
  if (t.wrap == 0) { s = dartModLL(s, W); tt = dartModLL(tt, H); }
  else if (t.wrap == 2) { s = min(max(s, 0ll), W - 1); tt = min(max(tt, 0ll), H - 1); }
  else if (s < 0 || s >= W || tt < 0 || tt >= H) return false;
  *idx = (size_t)t.levelOffset[level] + (size_t)(tt * W + s) * (size_t)t.channels;
  return true;
}
static __device__ inline Spec texelS(const TexCtx& c, const GTex& t, int level, long long s, long long tt) {
  size_t i;
  if (!texelAddr(t, level, s, tt, &i)) return Spec{0.f, 0.f, 0.f};
  return Spec{c.data[i], c.data[i + 1], c.data[i + 2]};
}
static __device__ inline double texelF(const TexCtx& c, const GTex& t, int level, long long s, long long tt) {
  size_t i;
  if (!texelAddr(t, level, s, tt, &i)) return 0.0;
  return (double)c.data[i];
}
// triangle(), :341-355
static __device__ inline Spec triangleS(const TexCtx& c, const GTex& t, int level, double s, double tt) {
  level = min(max(level, 0), t.levels - 1);
  s = s * levelW(t, level) - 0.5;
  tt = tt * levelH(t, level) - 0.5;
  const long long s0 = (long long)floor(s), t0 = (long long)floor(tt);
  const double ds = s - s0, dt = tt - t0;
  return texelS(c, t, level, s0, t0) * ((1.0 - ds) * (1.0 - dt)) + texelS(c, t, level, s0, t0 + 1) * ((1.0 - ds) * dt) +
         texelS(c, t, level, s0 + 1, t0) * (ds * (1.0 - dt)) + texelS(c, t, level, s0 + 1, t0 + 1) * (ds * dt);
}
static __device__ inline double triangleF(const TexCtx& c, const GTex& t, int level, double s, double tt) {
  level = min(max(level, 0), t.levels - 1);
  s = s * levelW(t, level) - 0.5;
  tt = tt * levelH(t, level) - 0.5;
  const long long s0 = (long long)floor(s), t0 = (long long)floor(tt);
  const double ds = s - s0, dt = tt - t0;
  return texelF(c, t, level, s0, t0) * ((1.0 - ds) * (1.0 - dt)) + texelF(c, t, level, s0, t0 + 1) * ((1.0 - ds) * dt) +
         texelF(c, t, level, s0 + 1, t0) * (ds * (1.0 - dt)) + texelF(c, t, level, s0 + 1, t0 + 1) * (ds * dt);
}
// lookup(), :206-222
static __device__ inline Spec lookupS(const TexCtx& c, const GTex& t, double s, double tt, double width) {
  const double level = t.levels - 1 + Log2d(fmax(width, 1.0e-8));
  if (level < 0) return triangleS(c, t, 0, s, tt);
  if (level >= t.levels - 1) return texelS(c, t, t.levels - 1, 0, 0);
  const int iLevel = (int)floor(level);
  const double delta = level - iLevel;
  return triangleS(c, t, iLevel, s, tt) * (1.0 - delta) + triangleS(c, t, iLevel + 1, s, tt) * delta;
}
static __device__ inline double lookupF(const TexCtx& c, const GTex& t, double s, double tt, double width) {
  const double level = t.levels - 1 + Log2d(fmax(width, 1.0e-8));
  if (level < 0) return triangleF(c, t, 0, s, tt);
  if (level >= t.levels - 1) return texelF(c, t, t.levels - 1, 0, 0);
  const int iLevel = (int)floor(level);
  const double delta = level - iLevel;
  return triangleF(c, t, iLevel, s, tt) * (1.0 - delta) + triangleF(c, t, iLevel + 1, s, tt) * delta;
}
// EWA(), :270-339
struct Ellipse {
  double s, t, A, B, C;
  long long s0, s1, t0, t1;
};
static __device__ inline Ellipse ewaSetup(int W, int H, double s, double t, double ds0, double dt0, double ds1, double dt1) {
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
  const double uSqrt = sqrt(det * C), vSqrt = sqrt(A * det);
  e.s0 = (long long)ceil(e.s - 2.0 * invDet * uSqrt);
  e.s1 = (long long)floor(e.s + 2.0 * invDet * uSqrt);
  e.t0 = (long long)ceil(e.t - 2.0 * invDet * vSqrt);
  e.t1 = (long long)floor(e.t + 2.0 * invDet * vSqrt);
  e.A = A; e.B = B; e.C = C;
  return e;
}
static __device__ inline double ewaWeight(const TexCtx& c, double r2) { return (double)c.data[(int)fmin(r2 * 128, 127.0)]; }
static __device__ __noinline__ void ewaSCold(const TexCtx& c, const GTex& t, int level, double s, double tt, double ds0, double dt0, double ds1,
                                             double dt1, Spec* out) {
  if (level >= t.levels) { *out = texelS(c, t, t.levels - 1, 0, 0); return; }
  const Ellipse e = ewaSetup(levelW(t, level), levelH(t, level), s, tt, ds0, dt0, ds1, dt1);
  Spec sum = Spec{0.f, 0.f, 0.f};
  double sumWts = 0.0;
  for (long long it = e.t0; it <= e.t1; ++it) {
    const double y = it - e.t;
    for (long long si = e.s0; si <= e.s1; ++si) {
      const double x = si - e.s;
      const double r2 = e.A * x * x + e.B * x * y + e.C * y * y;
      if (r2 < 1.0) {
        const double weight = ewaWeight(c, r2);
        sum = sum + texelS(c, t, level, si, it) * weight;
        sumWts += weight;
      }
    }
  }
  *out = sum / sumWts;
}
static __device__ __noinline__ double ewaFCold(const TexCtx& c, const GTex& t, int level, double s, double tt, double ds0, double dt0, double ds1,
                                               double dt1) {
  if (level >= t.levels) return texelF(c, t, t.levels - 1, 0, 0);
  const Ellipse e = ewaSetup(levelW(t, level), levelH(t, level), s, tt, ds0, dt0, ds1, dt1);
  double sum = 0.0, sumWts = 0.0;
  for (long long it = e.t0; it <= e.t1; ++it) {
    const double y = it - e.t;
    for (long long si = e.s0; si <= e.s1; ++si) {
      const double x = si - e.s;
      const double r2 = e.A * x * x + e.B * x * y + e.C * y * y;
      if (r2 < 1.0) {
        const double weight = ewaWeight(c, r2);
        sum += texelF(c, t, level, si, it) * weight;
        sumWts += weight;
      }
    }
  }
  return sum / sumWts;
}
// lookup2(), :224-268
struct Lookup2 {
  int mode;  // 0 trilinear lookup(width), 1 triangle(0), 2 EWA blend
  double width, ds0, dt0, ds1, dt1, d;
  int ilod;
};
static __device__ inline Lookup2 lookup2Setup(const GTex& t, double ds0, double dt0, double ds1, double dt1) {
  Lookup2 r;
  r.width = r.ds0 = r.dt0 = r.ds1 = r.dt1 = r.d = 0.0;
  r.ilod = 0;
  if (t.trilinear) {
    r.mode = 0;
    r.width = 2.0 * fmax(fmax(fabs(ds0), fabs(dt0)), fmax(fabs(ds1), fabs(dt1)));
    return r;
  }
  if (ds0 * ds0 + dt0 * dt0 < ds1 * ds1 + dt1 * dt1) {
    double x = ds0; ds0 = ds1; ds1 = x;
    x = dt0; dt0 = dt1; dt1 = x;
  }
  const double majorLength = sqrt(ds0 * ds0 + dt0 * dt0);
  double minorLength = sqrt(ds1 * ds1 + dt1 * dt1);
  if (minorLength * t.maxAniso < majorLength && minorLength > 0.0) {
    const double scale = majorLength / (minorLength * t.maxAniso);
    ds1 *= scale; dt1 *= scale; minorLength *= scale;
  }
  if (minorLength == 0.0) { r.mode = 1; return r; }
  const double lod = fmax(0.0, t.levels - 1.0 + Log2d(minorLength));
  r.mode = 2;
  r.ilod = (int)floor(lod);
  r.d = lod - r.ilod;
  r.ds0 = ds0; r.dt0 = dt0; r.ds1 = ds1; r.dt1 = dt1;
  return r;
}

// ---- texture mappings ----------------------------------------------------------------------------------------------------
struct ST {
  double s, t, dsdx, dtdx, dsdy, dtdy;
};
static __device__ inline void sphereST(const GTex& n, const V3& p, double* s, double* t) {  // spherical_mapping_2d.dart:58-64
  const V3 vec = Normalize(XfPoint(n.w2t, p));
  *s = SphericalTheta(vec) * DRT_INV_PI;
  *t = SphericalPhi(vec) * DRT_INV_TWOPI;
}
static __device__ inline void cylinderST(const GTex& n, const V3& p, double* s, double* t) {  // cylindrical_mapping_2d.dart:56-60
  const V3 vec = Normalize(XfPoint(n.w2t, p));
  *s = (DRT_PI + atan2((double)vec.y, (double)vec.x)) / (2.0 * DRT_PI);
  *t = vec.z;
}
static __device__ __noinline__ void mapSTCold(const GTex& n, const FullDG& dg, ST* out) {
  ST r;
  if (n.mapping == 0) {  // uv_mapping_2d.dart:25-36
    r.s = n.su * dg.u + n.du;
    r.t = n.sv * dg.v + n.dv;
    r.dsdx = n.su * dg.dudx; r.dtdx = n.sv * dg.dvdx;
    r.dsdy = n.su * dg.dudy; r.dtdy = n.sv * dg.dvdy;
  } else if (n.mapping == 1 || n.mapping == 2) {
    const double delta = n.mapping == 1 ? 0.1 : 0.01;
    double sx, tx, sy, ty;
    if (n.mapping == 1) { sphereST(n, dg.p, &r.s, &r.t); sphereST(n, dg.p + dg.dpdx * delta, &sx, &tx); sphereST(n, dg.p + dg.dpdy * delta, &sy, &ty); }
    else { cylinderST(n, dg.p, &r.s, &r.t); cylinderST(n, dg.p + dg.dpdx * delta, &sx, &tx); cylinderST(n, dg.p + dg.dpdy * delta, &sy, &ty); }
    r.dsdx = (sx - r.s) / delta;
    r.dtdx = (tx - r.t) / delta;
    if (r.dtdx > 0.5) r.dtdx = 1.0 - r.dtdx;
    else if (r.dtdx < -0.5) r.dtdx = -(r.dtdx + 1.0);
    r.dsdy = (sy - r.s) / delta;
    r.dtdy = (ty - r.t) / delta;
    if (r.dtdy > 0.5) r.dtdy = 1.0 - r.dtdy;
    else if (r.dtdy < -0.5) r.dtdy = -(r.dtdy + 1.0);
  } else {  // planar_mapping_2d.dart:29-38
    const V3 v1 = V3{n.v1[0], n.v1[1], n.v1[2]}, v2 = V3{n.v2[0], n.v2[1], n.v2[2]};
    r.s = n.du + Dot(dg.p, v1);
    r.t = n.dv + Dot(dg.p, v2);
    r.dsdx = Dot(dg.dpdx, v1); r.dtdx = Dot(dg.dpdx, v2);
    r.dsdy = Dot(dg.dpdy, v1); r.dtdy = Dot(dg.dpdy, v2);
  }
  *out = r;
}
// checkerboard_texture.dart:29-75: the weight of tex2; *single: one check is point sampled (weight 0 or 1)
static __device__ inline double checkerWeight(const GTex& n, const ST& m, bool* single) {
  const double pt = dartModLL((long long)floor(m.s) + (long long)floor(m.t), 2) == 0 ? 0.0 : 1.0;
  *single = true;
  if (n.aa == 0) return pt;
  const double ds = fmax(fabs(m.dsdx), fabs(m.dsdy)), dt = fmax(fabs(m.dtdx), fabs(m.dtdy));
  const double s0 = m.s - ds, s1 = m.s + ds, t0 = m.t - dt, t1 = m.t + dt;
  if (floor(s0) == floor(s1) && floor(t0) == floor(t1)) return pt;
#define DRT_BUMPINT(x) (floor((x) / 2) + 2.0 * fmax(((x) / 2) - floor((x) / 2) - 0.5, 0.0))
  const double sint = (DRT_BUMPINT(s1) - DRT_BUMPINT(s0)) / (2.0 * ds), tint = (DRT_BUMPINT(t1) - DRT_BUMPINT(t0)) / (2.0 * dt);
#undef DRT_BUMPINT
  double area2 = sint + tint - 2.0 * sint * tint;
  if (ds > 1.0 || dt > 1.0) area2 = 0.5;
  *single = false;
  return area2;
}

// ---- Perlin noise (lib/core/texture.dart:40-140) ----------------------------------------------------------------------------------
// Ken Perlin's published permutation (the table of his 2002 reference implementation), which texture.dart:142-203 holds twice over
static __device__ const uint8_t kNoisePerm[256] = {
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148,
    247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175,
    74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54,
    65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64,
    52, 217, 226, 250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213,
    119, 248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104,
    218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157,
    184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180};
static __device__ inline int noiseP(int i) { return kNoisePerm[i & 255]; }  // _NOISE_PERM has 512 entries: the second half repeats the first
static __device__ inline double noiseGrad(int x, int y, int z, double dx, double dy, double dz) {  // :119-125
  int h = noiseP(noiseP(noiseP(x) + y) + z);
  h &= 15;
  const double u = (h < 8 || h == 12 || h == 13) ? dx : dy;
  const double v = (h < 4 || h == 12 || h == 13) ? dy : dz;
  return ((h & 1) != 0 ? -u : u) + ((h & 2) != 0 ? -v : v);
}
static __device__ inline double noiseWeight(double t) {  // :128-132
  const double t3 = t * t * t, t4 = t3 * t;
  return 6.0 * t4 * t - 15.0 * t4 + 10.0 * t3;
}
static __device__ __noinline__ double NoiseCold(double x, double y, double z) {  // :40-77
  int ix = (int)floor(x), iy = (int)floor(y), iz = (int)floor(z);
  const double dx = x - ix, dy = y - iy, dz = z - iz;
  ix &= 255; iy &= 255; iz &= 255;
  const double w000 = noiseGrad(ix, iy, iz, dx, dy, dz), w100 = noiseGrad(ix + 1, iy, iz, dx - 1, dy, dz);
  const double w010 = noiseGrad(ix, iy + 1, iz, dx, dy - 1, dz), w110 = noiseGrad(ix + 1, iy + 1, iz, dx - 1, dy - 1, dz);
  const double w001 = noiseGrad(ix, iy, iz + 1, dx, dy, dz - 1), w101 = noiseGrad(ix + 1, iy, iz + 1, dx - 1, dy, dz - 1);
  const double w011 = noiseGrad(ix, iy + 1, iz + 1, dx, dy - 1, dz - 1), w111 = noiseGrad(ix + 1, iy + 1, iz + 1, dx - 1, dy - 1, dz - 1);
  const double wx = noiseWeight(dx), wy = noiseWeight(dy), wz = noiseWeight(dz);
  const double x00 = LerpD(wx, w000, w100), x10 = LerpD(wx, w010, w110), x01 = LerpD(wx, w001, w101), x11 = LerpD(wx, w011, w111);
  const double y0 = LerpD(wy, x00, x10), y1 = LerpD(wy, x01, x11);
  return LerpD(wz, y0, y1);
}
static __device__ inline double NoisePoint(const V3& p) { return NoiseCold((double)p.x, (double)p.y, (double)p.z); }
static __device__ inline double SmoothStep(double mn, double mx, double value) {  // common.dart:131-134
  const double v = clampD((value - mn) / (mx - mn), 0.0, 1.0);
  return v * v * (-2.0 * v + 3.0);
}
// FBm (:81-102) and Turbulence (:104-129): `turbulence` sums |noise| and adds 0.2 per octave that was cut
static __device__ __noinline__ double fbmCold(V3 Pt, V3 dpdx, V3 dpdy, double omega, int maxOctaves, bool turbulence) {
  const double s2 = fmax(LengthSquared(dpdx), LengthSquared(dpdy));
  const double foctaves = fmin((double)maxOctaves, fmax(0.0, -1.0 - 0.5 * Log2d(s2)));
  const int octaves = (int)floor(foctaves);
  double sum = 0.0, lambda = 1.0, o = 1.0;
  for (int i = 0; i < octaves; ++i) {
    const double nz = NoisePoint(Pt * lambda);
    sum += o * (turbulence ? fabs(nz) : nz);
    lambda *= 1.99;
    o *= omega;
  }
  const double partialOctave = foctaves - octaves;
  const double nz = NoisePoint(Pt * lambda);
  sum += o * SmoothStep(0.3, 0.7, partialOctave) * (turbulence ? fabs(nz) : nz);
  if (turbulence) sum += (maxOctaves - foctaves) * 0.2;
  return sum;
}
// IdentityMapping3D.map (identity_mapping_3d.dart:25-29)
static __device__ inline V3 map3D(const GTex& n, const FullDG& dg, V3* dpdx, V3* dpdy) {
  *dpdx = XfVector(n.w2t, dg.dpdx);
  *dpdy = XfVector(n.w2t, dg.dpdy);
  return XfPoint(n.w2t, dg.p);
}
// 7 fbm (fbm_texture.dart:26-32), 8 wrinkled (wrinkled_texture.dart:26-32), 9 windy (windy_texture.dart:26-38)
static __device__ inline double noiseScalar(const GTex& n, const FullDG& dg) {
  V3 dpdx, dpdy;
  const V3 Pt = map3D(n, dg, &dpdx, &dpdy);
  if (n.kind == 7) return fbmCold(Pt, dpdx, dpdy, n.value[0], n.aa, false);
  if (n.kind == 8) return fbmCold(Pt, dpdx, dpdy, n.value[0], n.aa, true);
  const double windStrength = fbmCold(Pt * 0.1, dpdx * 0.1, dpdy * 0.1, 0.5, 3, false);
  const double waveHeight = fbmCold(Pt, dpdx, dpdy, 0.5, 6, false);
  return fabs(windStrength) * waveHeight;
}
static __device__ inline bool insideDot(double s, double t) {  // dots_texture.dart:26-52
  const int sCell = (int)floor(s + 0.5), tCell = (int)floor(t + 0.5);
  if (NoiseCold(sCell + 0.5, tCell + 0.5, 0.5) > 0) {
    const double radius = 0.35, maxShift = 0.5 - radius;
    const double sCenter = sCell + maxShift * NoiseCold(sCell + 1.5, tCell + 2.8, 0.5);
    const double tCenter = tCell + maxShift * NoiseCold(sCell + 4.5, tCell + 9.8, 0.5);
    const double ds = s - sCenter, dt = t - tCenter;
    if (ds * ds + dt * dt < radius * radius) return true;
  }
  return false;
}
static __device__ inline bool checker3D(const GTex& n, const FullDG& dg) {  // checkerboard_3d_texture.dart:26-35: true = tex1
  V3 dpdx, dpdy;
  const V3 p = map3D(n, dg, &dpdx, &dpdy);
  return dartModLL((long long)floor((double)p.x) + (long long)floor((double)p.y) + (long long)floor((double)p.z), 2) == 0;
}

// ---- Texture.evaluate (recursive over the node table: depth bounded by the host, children precede parents) ------------------
static __device__ double texEvalF(const TexCtx& c, int id, const FullDG& dg);
static __device__ void texEvalS(const TexCtx& c, int id, const FullDG& dg, Spec* out);

static __device__ __noinline__ double texEvalF(const TexCtx& c, int id, const FullDG& dg) {
  const GTex& n = c.nodes[id];
  switch (n.kind) {
    case 0: return n.value[0];
    case 1: {
      const double t1 = texEvalF(c, n.tex1, dg), t2 = texEvalF(c, n.tex2, dg);
      return t2 * t1;
    }
    case 2: {
      const double t1 = texEvalF(c, n.tex1, dg), t2 = texEvalF(c, n.tex2, dg), amt = texEvalF(c, n.amount, dg);
      return t1 * (1.0 - amt) + t2 * amt;
    }
    case 3: {
      ST m;
      mapSTCold(n, dg, &m);
      const Lookup2 q = lookup2Setup(n, m.dsdx, m.dtdx, m.dsdy, m.dtdy);
      if (q.mode == 0) return lookupF(c, n, m.s, m.t, q.width);
      if (q.mode == 1) return triangleF(c, n, 0, m.s, m.t);
      return ewaFCold(c, n, q.ilod, m.s, m.t, q.ds0, q.dt0, q.ds1, q.dt1) * (1.0 - q.d) +
             ewaFCold(c, n, q.ilod + 1, m.s, m.t, q.ds0, q.dt0, q.ds1, q.dt1) * q.d;
    }
    case 4: {
      ST m;
      mapSTCold(n, dg, &m);
      bool single;
      const double w2 = checkerWeight(n, m, &single);
      if (single) return texEvalF(c, w2 == 0.0 ? n.tex1 : n.tex2, dg);
      return texEvalF(c, n.tex1, dg) * (1.0 - w2) + texEvalF(c, n.tex2, dg) * w2;
    }
    case 6: {
      ST m;
      mapSTCold(n, dg, &m);
      const double s = m.s, t = m.t;
      return n.value[0] * ((1.0 - s) * (1 - t)) + n.value2[0] * (1.0 - s) * t + n.value2[3] * s * (1.0 - t) + n.value2[6] * s * t;
    }
    case 7: case 8: case 9: return noiseScalar(n, dg);
    case 11: {  // DotsTexture(mapping, outsideDot = tex1, insideDot = tex2)
      ST m;
      mapSTCold(n, dg, &m);
      return texEvalF(c, insideDot(m.s, m.t) ? n.tex2 : n.tex1, dg);
    }
    case 12: return texEvalF(c, checker3D(n, dg) ? n.tex1 : n.tex2, dg);
    default: return 0.0;
  }
}
static __device__ __noinline__ void texEvalS(const TexCtx& c, int id, const FullDG& dg, Spec* out) {
  const GTex& n = c.nodes[id];
  switch (n.kind) {
    case 0: *out = mks(n.value[0], n.value[1], n.value[2]); return;
    case 1: {
      Spec a, b;
      texEvalS(c, n.tex1, dg, &a);
      texEvalS(c, n.tex2, dg, &b);
      *out = a * b;
      return;
    }
    case 2: {
      Spec a, b;
      texEvalS(c, n.tex1, dg, &a);
      texEvalS(c, n.tex2, dg, &b);
      const double amt = texEvalF(c, n.amount, dg);
      *out = a * (1.0 - amt) + b * amt;
      return;
    }
    case 3: {
      ST m;
      mapSTCold(n, dg, &m);
      const Lookup2 q = lookup2Setup(n, m.dsdx, m.dtdx, m.dsdy, m.dtdy);
      if (q.mode == 0) { *out = lookupS(c, n, m.s, m.t, q.width); return; }
      if (q.mode == 1) { *out = triangleS(c, n, 0, m.s, m.t); return; }
      Spec a, b;
      ewaSCold(c, n, q.ilod, m.s, m.t, q.ds0, q.dt0, q.ds1, q.dt1, &a);
      ewaSCold(c, n, q.ilod + 1, m.s, m.t, q.ds0, q.dt0, q.ds1, q.dt1, &b);
      *out = a * (1.0 - q.d) + b * q.d;
      return;
    }
    case 4: {
      ST m;
      mapSTCold(n, dg, &m);
      bool single;
      const double w2 = checkerWeight(n, m, &single);
      if (single) { texEvalS(c, w2 == 0.0 ? n.tex1 : n.tex2, dg, out); return; }
      Spec a, b;
      texEvalS(c, n.tex1, dg, &a);
      texEvalS(c, n.tex2, dg, &b);
      *out = a * (1.0 - w2) + b * w2;
      return;
    }
    case 5: {
      ST m;
      mapSTCold(n, dg, &m);
      *out = mks(m.s - floor(m.s), m.t - floor(m.t), 0.0);
      return;
    }
    case 6: {
      ST m;
      mapSTCold(n, dg, &m);
      const double s = m.s, t = m.t;
      const Spec v00 = mks(n.value[0], n.value[1], n.value[2]), v01 = mks(n.value2[0], n.value2[1], n.value2[2]),
                 v10 = mks(n.value2[3], n.value2[4], n.value2[5]), v11 = mks(n.value2[6], n.value2[7], n.value2[8]);
      *out = v00 * ((1.0 - s) * (1 - t)) + v01 * (1.0 - s) * t + v10 * s * (1.0 - t) + v11 * s * t;
      return;
    }
    case 7: case 8: case 9: *out = mks1(noiseScalar(n, dg)); return;  // new Spectrum(n)
    case 10: {  // marble_texture.dart:27-66
      V3 dpdx, dpdy;
      V3 Pt = map3D(n, dg, &dpdx, &dpdy);
      const double scale = n.value[1], variation = n.value[2];
      Pt = Pt * scale;
      const double marble = (double)Pt.y + variation * fbmCold(Pt, dpdx * scale, dpdy * scale, n.value[0], n.aa, false);
      double t = 0.5 + 0.5 * sin(marble);
      const double cs[27] = {0.58, 0.58, 0.6, 0.58, 0.58, 0.6, 0.58, 0.58, 0.6, 0.5, 0.5, 0.5, 0.6, 0.59, 0.58,
                             0.58, 0.58, 0.6, 0.58, 0.58, 0.6, 0.2, 0.2, 0.33, 0.58, 0.58, 0.6};
      const int NSEG = 9 - 3;
      const int first = (int)floor(t * NSEG);
      t = (t * NSEG - first);
      const int ci = first * 3;
      const Spec c0 = mks(cs[ci], cs[ci + 1], cs[ci + 2]), c1 = mks(cs[ci + 3], cs[ci + 4], cs[ci + 5]), c2 = mks(cs[ci + 6], cs[ci + 7], cs[ci + 8]),
                 c3 = mks(cs[ci + 9], cs[ci + 10], cs[ci + 11]);
      Spec s0 = c0 * (1.0 - t) + c1 * t;
      Spec s1 = c1 * (1.0 - t) + c2 * t;
      const Spec s2 = c2 * (1.0 - t) + c3 * t;
      s0 = s0 * (1.0 - t) + s1 * t;
      s1 = s1 * (1.0 - t) + s2 * t;
      *out = (s0 * (1.0 - t) + s1 * t) * 1.5;
      return;
    }
    case 11: {
      ST m;
      mapSTCold(n, dg, &m);
      texEvalS(c, insideDot(m.s, m.t) ? n.tex2 : n.tex1, dg, out);
      return;
    }
    case 12: texEvalS(c, checker3D(n, dg) ? n.tex1 : n.tex2, dg, out); return;
    default: *out = Spec{0.f, 0.f, 0.f};
  }
}

// ---- Material.Bump (material.dart:35-88) -------------------------------------------------------------------------------------
static __device__ __noinline__ void bumpCold(const TexCtx& c, int d, const FullDG& dgGeom, const FullDG& dgs, FullDG* dgBump) {
  FullDG dgEval = dgs;
  double du = 0.5 * (fabs(dgs.dudx) + fabs(dgs.dudy));
  if (du == 0.0) du = 0.01;
  dgEval.p = dgs.p + dgs.dpdu * du;
  dgEval.u = dgs.u + du;
  dgEval.nn = Normalize(Cross(dgs.dpdu, dgs.dpdv) + dgs.dndu * du);
  const double uDisplace = texEvalF(c, d, dgEval);
  double dv = 0.5 * (fabs(dgs.dvdx) + fabs(dgs.dvdy));
  if (dv == 0.0) dv = 0.01;
  dgEval.p = dgs.p + dgs.dpdv * dv;
  dgEval.u = dgs.u;
  dgEval.v = dgs.v + dv;
  dgEval.nn = Normalize(Cross(dgs.dpdu, dgs.dpdv) + dgs.dndv * dv);
  const double vDisplace = texEvalF(c, d, dgEval);
  const double displace = texEvalF(c, d, dgs);
  FullDG b = dgs;
  b.dpdu = dgs.dpdu + dgs.nn * (uDisplace - displace) / du + dgs.dndu * displace;
  b.dpdv = dgs.dpdv + dgs.nn * (vDisplace - displace) / dv + dgs.dndv * displace;
  b.nn = Normalize(Cross(b.dpdu, b.dpdv));
  if (dgs.reverse) b.nn = b.nn * -1.0;
  b.nn = FaceForward(b.nn, dgGeom.nn);
  *dgBump = b;
}

// ---- the materials' getBSDF (lib/materials/*.dart) ---------------------------------------------------------------------------
struct HitBsdf {  // new BSDF(dgs, dgGeom.nn): nn = dgs.nn, sn = normalize(dgs.dpdu) (bsdf.dart:45-51); tn follows
  V3 nn, sn;
  int n;
  GLobe lobes[8];
};
static __device__ inline Spec clampS(const Spec& a, double lo, double hi) {  // rgb_color.dart:189-192
  return mks(clampD((double)a.r, lo, hi), clampD((double)a.g, lo, hi), clampD((double)a.b, lo, hi));
}
static __device__ inline Spec clampS(const Spec& a) { return clampS(a, 0.0, CUDART_INF); }
static __device__ inline double blinnExp(double rough) {  // 1 / roughness, then blinn.dart:24-28
  const double e = 1.0 / rough;
  return (e > 10000.0 || isnan(e)) ? 10000.0 : e;
}
static __device__ inline Spec approxEta(const Spec& fr) {  // shiny_metal_material.dart:75-79
  const Spec refl = clampS(fr, 0.0, 0.999);
  const Spec sq = mks(sqrt((double)refl.r), sqrt((double)refl.g), sqrt((double)refl.b));
  return (mks1(1.0) + sq) / (mks1(1.0) - sq);
}
static __device__ inline GLobe mkLobe(int kind, const Spec& R, int fresnel = 0, double param = 0.0, double ei = 1.0, double et = 1.0) {
  GLobe l;
  l.kind = kind; l.fresnel = fresnel;
  l.rgb[0] = R.r; l.rgb[1] = R.g; l.rgb[2] = R.b;
  l.eta[0] = l.eta[1] = l.eta[2] = 0.f;
  l.k[0] = l.k[1] = l.k[2] = 0.f;
  l.wrap = 0;
  l.param = param; l.ei = ei; l.et = et;
  l.scale[0] = l.scale[1] = l.scale[2] = 1.f;
  l.pad_ = 0.f;
  return l;
}
static __device__ inline void setSpec3(float* dst, const Spec& s) { dst[0] = s.r; dst[1] = s.g; dst[2] = s.b; }
static __device__ inline void frameHit(HitBsdf* b, const FullDG& dgs) {
  b->nn = dgs.nn;
  b->sn = Normalize(dgs.dpdu);
  b->n = 0;
}

static __device__ __noinline__ void materialBsdfCold(const RenderScene& rs, const TexCtx& c, const GProgram* programs, uint32_t mat,
                                                     const FullDG& dgGeom, const FullDG& dgShading, HitBsdf* b, int depth) {
  const GProgram pr = programs[mat];
  if (pr.kind < 0) {  // constant parameters, no bump map: the flattened list
    frameHit(b, dgShading);
    const uint2 ml = __ldg(rs.matLobes + mat);
    for (uint32_t i = 0; i < ml.y && i < 8u; ++i) b->lobes[b->n++] = rs.lobes[ml.x + i];
    return;
  }
  const int* t = pr.tex;
  auto S = [&](int id, const FullDG& dg) { Spec v; texEvalS(c, id, dg, &v); return v; };
  if (pr.kind == 9) {  // mix_material.dart:36-50 (one level: the host rejects a mix of mixes)
    if (depth > 0) { frameHit(b, dgShading); return; }
    materialBsdfCold(rs, c, programs, (uint32_t)pr.m1, dgGeom, dgShading, b, 1);
    HitBsdf b2;
    materialBsdfCold(rs, c, programs, (uint32_t)pr.m2, dgGeom, dgShading, &b2, 1);
    const Spec s1 = clampS(S(t[0], dgShading));
    const Spec s2 = clampS(mks1(1.0) - s1);
    for (int i = 0; i < b->n; ++i) { b->lobes[i].wrap |= 2; setSpec3(b->lobes[i].scale, s1); }
    for (int i = 0; i < b2.n && b->n < 8; ++i) { GLobe l = b2.lobes[i]; l.wrap |= 2; setSpec3(l.scale, s2); b->lobes[b->n++] = l; }
    return;
  }
  FullDG dgs = dgShading;
  if (pr.bump >= 0) bumpCold(c, pr.bump, dgGeom, dgShading, &dgs);
  frameHit(b, dgs);
  auto add = [&](const GLobe& l) { if (b->n < 8) b->lobes[b->n++] = l; };
  switch (pr.kind) {
    case 0: {  // matte_material.dart:41-65
      const Spec r = clampS(S(t[0], dgs));
      const double sig = clampD(texEvalF(c, t[1], dgs), 0.0, 90.0);
      if (!IsBlack(r)) add(sig == 0.0 ? mkLobe(0, r) : mkLobe(1, r, 0, sig));
      break;
    }
    case 1: {  // mirror_material.dart:38-55
      const Spec R = clampS(S(t[0], dgs));
      if (!IsBlack(R)) add(mkLobe(3, R, 0));
      break;
    }
    case 2: {  // glass_material.dart:44-70
      const double ior = texEvalF(c, t[2], dgs);
      const Spec R = clampS(S(t[0], dgs)), T = clampS(S(t[1], dgs));
      if (!IsBlack(R)) add(mkLobe(3, R, 1, 0.0, 1.0, ior));
      if (!IsBlack(T)) add(mkLobe(4, T, 1, 0.0, 1.0, ior));
      break;
    }
    case 3: {  // plastic_material.dart:43-72
      const Spec kd = clampS(S(t[0], dgs));
      if (!IsBlack(kd)) add(mkLobe(0, kd));
      const Spec ks = clampS(S(t[1], dgs));
      if (!IsBlack(ks)) add(mkLobe(2, ks, 1, blinnExp(texEvalF(c, t[2], dgs)), 1.5, 1.0));
      break;
    }
    case 4: {  // metal_material.dart:44-64
      const double rough = texEvalF(c, t[2], dgs);
      GLobe l = mkLobe(2, mks1(1.0), 2, blinnExp(rough));
      setSpec3(l.eta, S(t[0], dgs));
      setSpec3(l.k, S(t[1], dgs));
      add(l);
      break;
    }
    case 5: {  // shiny_metal_material.dart:44-73
      const Spec spec = clampS(S(t[0], dgs));
      const double rough = texEvalF(c, t[2], dgs);
      const Spec R = clampS(S(t[1], dgs));
      if (!IsBlack(spec)) { GLobe l = mkLobe(2, mks1(1.0), 2, blinnExp(rough)); setSpec3(l.eta, approxEta(spec)); add(l); }
      if (!IsBlack(R)) { GLobe l = mkLobe(3, mks1(1.0), 2); setSpec3(l.eta, approxEta(R)); add(l); }
      break;
    }
    case 6: {  // substrate_material.dart:48-70
      const Spec d = clampS(S(t[0], dgs)), sp = clampS(S(t[1], dgs));
      const double u = texEvalF(c, t[2], dgs), v = texEvalF(c, t[3], dgs);
      if (!IsBlack(d) || !IsBlack(sp)) { GLobe l = mkLobe(5, d, 0, blinnExp(u), blinnExp(v)); setSpec3(l.eta, sp); add(l); }
      break;
    }
    case 7: {  // translucent_material.dart:47-103
      const Spec r = clampS(S(t[2], dgs)), tr = clampS(S(t[3], dgs));
      if (IsBlack(r) && IsBlack(tr)) break;
      const Spec kd = clampS(S(t[0], dgs));
      if (!IsBlack(kd)) {
        if (!IsBlack(r)) add(mkLobe(0, r * kd));
        if (!IsBlack(tr)) { GLobe l = mkLobe(0, tr * kd); l.wrap = 1; add(l); }
      }
      const Spec ks = clampS(S(t[1], dgs));
      if (!IsBlack(ks)) {
        const double e = blinnExp(texEvalF(c, t[4], dgs));
        if (!IsBlack(r)) add(mkLobe(2, r * ks, 1, e, 1.5, 1.0));
        if (!IsBlack(tr)) { GLobe l = mkLobe(2, tr * ks, 1, e, 1.5, 1.0); l.wrap = 1; add(l); }
      }
      break;
    }
    case 11: {  // measured_material.dart:219-238: m1 = table of drt_set_measured (its kind picks the BxDF)
      const GMeasured tb = rs.measured[pr.m1];
      GLobe l = mkLobe(tb.kind == 0 ? 6 : 7, mks1(1.0), 0, (double)pr.m1);
      l.et = __longlong_as_double((long long)(uintptr_t)tb.data);
      for (int k = 0; k < 3; ++k) l.k[k] = __int_as_float(tb.dims[k]);
      add(l);
      break;
    }
    case 10: {  // subsurface_material.dart:52-69, kd_subsurface_material.dart:48-67 (the BSSRDF is the dipole integrator's)
      const Spec R = clampS(S(t[0], dgs));
      const double e = texEvalF(c, t[1], dgs);
      if (!IsBlack(R)) add(mkLobe(3, R, 1, 0.0, 1.0, e));
      break;
    }
    default: {  // 8 uber_material.dart:56-104
      const Spec op = clampS(S(t[5], dgs));
      if (!(op.r == 1.f && op.g == 1.f && op.b == 1.f)) add(mkLobe(4, mks(-(double)op.r, -(double)op.g, -(double)op.b) + mks1(1.0), 1, 0.0, 1.0, 1.0));
      const Spec kd = op * clampS(S(t[0], dgs));
      if (!IsBlack(kd)) add(mkLobe(0, kd));
      const double e = texEvalF(c, t[6], dgs);
      const Spec ks = op * clampS(S(t[1], dgs));
      if (!IsBlack(ks)) add(mkLobe(2, ks, 1, blinnExp(texEvalF(c, t[4], dgs)), e, 1.0));
      const Spec kr = op * clampS(S(t[2], dgs));
      if (!IsBlack(kr)) add(mkLobe(3, kr, 1, 0.0, e, 1.0));
      const Spec kt = op * clampS(S(t[3], dgs));
      if (!IsBlack(kt)) add(mkLobe(4, kt, 1, 0.0, e, 1.0));
      break;
    }
  }
}

}  // namespace drt
