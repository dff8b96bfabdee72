// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.h header).  PARITY UNPINNED.
//
// Shapes (Triangle, Sphere) and BVHAccel restated from the reference:
//   lib/shapes/triangle.dart, lib/shapes/sphere.dart, lib/core/common.dart (Quadratic,
//   partition, nth_element), lib/accelerators/bvh_accel.dart.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "ref_anim.h"
#include "ref_core.h"

namespace orc {

// lib/shapes/sphere.dart:24-32,314-323 — radius/zmin/zmax/phiMax are Dart doubles.
// The same record also carries the other quadric on the path, Disk (lib/shapes/disk.dart:24-31,157-166):
// shape == 1, with height / radius / innerRadius / phiMax.  Quadrics share one id range after the triangles.
struct Sphere {
  Transform o2w;  // objectToWorld (m) / worldToObject is o2w.mInv kept separately below
  Transform w2o;
  double radius, zmin, zmax, phiMax, thetaMin, thetaMax;
  bool reverseOrientation = false;
  int shape = 0;  // 0 sphere, 1 disk, 2 cylinder, 3 cone, 4 paraboloid, 5 hyperboloid
  double height = 0.0, innerRadius = 0.0;
  Vec hp1, hp2;               // hyperboloid.dart:24 (Points, float32)
  double ha = 0.0, hc = 0.0;  // hyperboloid.dart:36-49 implicit coefficients
  Sphere() {}
  // The remaining quadrics.  prm: the ParamSet values in the order of each Create():
  //   2 cylinder   radius, zmin, zmax, phimax            (cylinder.dart:24-31,239-247)
  //   3 cone       height, radius, phimax                (cone.dart:23-27,216-222)
  //   4 paraboloid radius, zmin, zmax, phimax            (paraboloid.dart:23-29,220-228)
  //   5 hyperboloid p1.xyz, p2.xyz, phimax               (hyperboloid.dart:23-49,263-268)
  static Sphere makeQuadric(int kind, const float* O2W, const float* W2O, const double* prm, bool ro) {
    Sphere q;
    q.o2w = Transform(O2W, W2O);
    q.w2o = Transform(W2O, O2W);
    q.shape = kind;
    q.reverseOrientation = ro;
    q.thetaMin = q.thetaMax = 0.0;
    double pm = 360.0;
    if (kind == 2 || kind == 4) {
      q.radius = prm[0];
      q.zmin = std::fmin(prm[1], prm[2]);
      q.zmax = std::fmax(prm[1], prm[2]);
      pm = prm[3];
    } else if (kind == 3) {
      q.height = prm[0];
      q.radius = prm[1];
      q.zmin = 0.0;
      q.zmax = q.height;
      pm = prm[2];
    } else {
      Vec p1(prm[0], prm[1], prm[2]), p2(prm[3], prm[4], prm[5]);
      pm = prm[6];
      double radius1 = std::sqrt((double)p1.x * p1.x + (double)p1.y * p1.y);
      double radius2 = std::sqrt((double)p2.x * p2.x + (double)p2.y * p2.y);
      q.radius = std::fmax(radius1, radius2);  // rmax
      q.zmin = std::fmin((double)p1.z, (double)p2.z);
      q.zmax = std::fmax((double)p1.z, (double)p2.z);
      if (p2.z == 0.0f) std::swap(p1, p2);
      Vec pp = p1;
      double a, c;
      do {  // hyperboloid.dart:40-48
        pp = pp + ((p2 - p1) * 2.0);
        double xy1 = (double)pp.x * pp.x + (double)pp.y * pp.y;
        double xy2 = (double)p2.x * p2.x + (double)p2.y * p2.y;
        a = (1.0 / xy1 - ((double)pp.z * pp.z) / (xy1 * p2.z * p2.z)) / (1.0 - (xy2 * pp.z * pp.z) / (xy1 * p2.z * p2.z));
        c = (a * xy2 - 1.0) / ((double)p2.z * p2.z);
      } while (std::isinf(a) || std::isnan(a));
      q.hp1 = p1;
      q.hp2 = p2;
      q.ha = a;
      q.hc = c;
    }
    q.phiMax = Radians(clampd(pm, 0.0, 360.0));
    return q;
  }
  static Sphere makeDisk(const float* O2W, const float* W2O, double h, double r, double ri, double pm, bool ro) {
    Sphere d;
    d.o2w = Transform(O2W, W2O);
    d.w2o = Transform(W2O, O2W);
    d.shape = 1;
    d.height = h;
    d.radius = r;
    d.innerRadius = ri;
    d.phiMax = Radians(clampd(pm, 0.0, 360.0));  // disk.dart:28
    d.zmin = d.zmax = h;
    d.thetaMin = d.thetaMax = 0.0;
    d.reverseOrientation = ro;
    return d;
  }
  Sphere(const float* O2W, const float* W2O, double r, double z0, double z1, double pm, bool ro) {
    o2w = Transform(O2W, W2O);
    w2o = Transform(W2O, O2W);
    radius = r;
    zmin = clampd(std::fmin(z0, z1), -radius, radius);
    zmax = clampd(std::fmax(z0, z1), -radius, radius);
    thetaMin = std::acos(clampd(zmin / radius, -1.0, 1.0));
    thetaMax = std::acos(clampd(zmax / radius, -1.0, 1.0));
    phiMax = Radians(clampd(pm, 0.0, 360.0));
    reverseOrientation = ro;
  }
  // sphere.dart:34-37 + lib/core/shape.dart:38-40
  BBox worldBound() const {
    if (shape == 1) return o2w.bbox(BBox(Vec(-radius, -radius, height), Vec(radius, radius, height)));  // disk.dart:32-35
    // shape >= 2: cylinder.dart:33-37, cone.dart:29-33 (z in [0, height]), paraboloid.dart:31-35, hyperboloid.dart:51-55 (rmax)
    BBox ob(Vec(-radius, -radius, zmin), Vec(radius, radius, zmax));
    return o2w.bbox(ob);
  }
};

// Hit record shared by closest-hit queries.  `prim` is the UPLOAD-ORDER id
// (SURVEY §8b "primitive id convention"): triangles 0..ntris-1, spheres after.
struct Hit {
  double t = 0.0;
  double b1 = 0.0, b2 = 0.0;  // triangle barycentrics / sphere (u, v)
  int32_t prim = -1;
  double rayEpsilon = 0.0;
  Vec phitObj;       // sphere: object-space hit point after the 1e-5*r nudge
  double phi = 0.0;  // sphere
  int32_t inst = -1;  // TransformedPrimitive the hit came through (transformed_primitive.dart:30-62), -1: a top-level primitive
};

// lib/core/common.dart:140-167
static inline bool Quadratic(double A, double B, double C, double* t0, double* t1) {
  double discrim = B * B - 4.0 * A * C;
  if (discrim < 0.0) return false;
  double rootDiscrim = std::sqrt(discrim);
  double q;
  if (B < 0.0) q = -0.5 * (B - rootDiscrim);
  else q = -0.5 * (B + rootDiscrim);
  *t0 = q / A;
  *t1 = C / q;
  if (*t0 > *t1) std::swap(*t0, *t1);
  return true;
}

struct Counters {
  uint64_t nodes_visited = 0;  // every _intersectP call (bvh_accel.dart:125/187)
  uint64_t prims_tested = 0;   // every primitive test (bvh_accel.dart:131/193)
  uint64_t rays = 0;
  void add(const Counters& o) { nodes_visited += o.nodes_visited; prims_tested += o.prims_tested; rays += o.rays; }
};

struct LinearNode {  // bvh_accel.dart:533-538
  BBox bounds;
  int32_t offset = 0;       // primitivesOffset / secondChildOffset
  int32_t nPrimitives = 0;  // 0 -> interior
  int32_t axis = 0;
};

struct Scene {
  // world-space triangle soup, lib/shapes/triangle_mesh.dart:24-60
  std::vector<float> P;       // nverts*3
  std::vector<uint32_t> idx;  // ntris*3
  std::vector<Sphere> spheres;
  std::vector<uint32_t> buildOrder;  // refined order handed to BVHAccel (ids in upload numbering)

  // per-vertex shading attributes of the meshes the soup was merged from (triangle_mesh.dart:24-28: n, s, uvs are kept
  // as given, i.e. in OBJECT space; triangle.dart:271-364 transforms them with the mesh's objectToWorld)
  struct MeshInfo {
    Transform o2w;
    bool hasN = false, hasS = false, hasUV = false;
  };
  std::vector<float> vertN, vertS, vertUV;  // nverts*3, nverts*3, nverts*2 (empty when no mesh has them)
  std::vector<uint32_t> meshOfTri;          // ntris (empty: no mesh carries attributes)
  std::vector<MeshInfo> meshes;
  const MeshInfo* meshOf(uint32_t tri) const { return meshOfTri.empty() ? nullptr : &meshes[meshOfTri[tri]]; }
  // triangle.dart:246-262 getUVs
  void triUVs(uint32_t tri, double uv[6]) const {
    const MeshInfo* m = meshOf(tri);
    if (m && m->hasUV) {
      for (int k = 0; k < 3; ++k) {
        uv[2 * k] = vertUV[2 * (size_t)idx[3 * (size_t)tri + k]];
        uv[2 * k + 1] = vertUV[2 * (size_t)idx[3 * (size_t)tri + k] + 1];
      }
    } else {
      uv[0] = 0.0; uv[1] = 0.0; uv[2] = 1.0; uv[3] = 0.0; uv[4] = 1.0; uv[5] = 1.0;
    }
  }

  // per-primitive attributes (upload numbering)
  std::vector<int32_t> materialOf;
  std::vector<int32_t> lightOf;
  std::vector<uint8_t> reverseOf;

  // BVH (bvh_accel.dart:484-487)
  int maxPrimsInNode = 4;
  int splitMethod = 2;
  std::vector<uint32_t> ordered;  // primitives after the build (ids in upload numbering)
  std::vector<LinearNode> nodes;

  // TransformedPrimitive (lib/core/primitive/transformed_primitive.dart): what DartRay.shape builds for an animated shape
  // (dartray.dart:404-452) and DartRay.objectInstance for an instance (:505-546).  An Object is the `primitive` it wraps: one
  // GeometricPrimitive, or the BVHAccel over the refined primitives of the shape / the instance's primitive list (built with the
  // caller's split method and maxnodeprims: BVHAccel's defaults for an animated shape, the scene's accelerator for an instance).
  // Object primitives are geometric primitives of this scene (ids < nprims()) that the top-level build order leaves out; an
  // instance is a top-level primitive with id nprims() + its index.
  struct Object {
    std::vector<uint32_t> order;  // refined order handed to the nested BVHAccel (one entry: the primitive itself)
    int split = 2, maxPrims = 1;
    std::vector<uint32_t> ordered;
    std::vector<LinearNode> nodes;
    BBox worldBound(const Scene& sc) const {  // bvh_accel.dart:93-95 / geometric_primitive.dart:35-37
      if (order.size() == 1) return sc.primBound(order[0]);
      return nodes.empty() ? BBox() : nodes[0].bounds;
    }
  };
  struct Instance {
    uint32_t object = 0;
    AnimatedTransform worldToPrimitive;
    BBox bound;  // worldToPrimitive.motionBounds(primitive.worldBound(), true), transformed_primitive.dart:76-78
  };
  std::vector<Object> objects;
  std::vector<Instance> instances;

  uint32_t ntris() const { return (uint32_t)(idx.size() / 3); }
  uint32_t nprims() const { return ntris() + (uint32_t)spheres.size(); }

  void triVerts(uint32_t tri, Vec* p1, Vec* p2, Vec* p3) const {
    const float* a = &P[3 * (size_t)idx[3 * (size_t)tri + 0]];
    const float* b = &P[3 * (size_t)idx[3 * (size_t)tri + 1]];
    const float* c = &P[3 * (size_t)idx[3 * (size_t)tri + 2]];
    p1->x = a[0]; p1->y = a[1]; p1->z = a[2];
    p2->x = b[0]; p2->y = b[1]; p2->z = b[2];
    p3->x = c[0]; p3->y = c[1]; p3->z = c[2];
  }
  BBox primBound(uint32_t prim) const {
    if (prim < ntris()) {  // triangle.dart:39-42
      Vec p1, p2, p3;
      triVerts(prim, &p1, &p2, &p3);
      return UnionPoint(BBox(p1, p2), p3);
    }
    if (prim >= nprims()) return instances[prim - nprims()].bound;
    return spheres[prim - ntris()].worldBound();
  }

  // ---- primitive tests -------------------------------------------------
  bool triIntersect(uint32_t tri, Ray& ray, Hit* hit) const;   // triangle.dart:44-160 (+ geometric_primitive.dart:47-61)
  bool triIntersectP(uint32_t tri, const Ray& ray) const;      // triangle.dart:162-240
  bool sphIntersect(const Sphere& s, Ray& r, Hit* hit) const;  // sphere.dart:39-167
  bool sphIntersectP(const Sphere& s, const Ray& r) const;     // sphere.dart:169-241
  bool primIntersect(uint32_t prim, Ray& ray, Hit* hit, Counters* c = nullptr) const {
    if (prim >= nprims()) return instanceIntersect(prim - nprims(), ray, hit, c);
    bool h = prim < ntris() ? triIntersect(prim, ray, hit) : sphIntersect(spheres[prim - ntris()], ray, hit);
    if (h) { hit->prim = (int32_t)prim; hit->inst = -1; }
    return h;
  }
  bool primIntersectP(uint32_t prim, const Ray& ray, Counters* c = nullptr) const {
    if (prim >= nprims()) return instanceIntersectP(prim - nprims(), ray, c);
    return prim < ntris() ? triIntersectP(prim, ray) : sphIntersectP(spheres[prim - ntris()], ray);
  }
  bool instanceIntersect(uint32_t inst, Ray& r, Hit* hit, Counters* c) const;   // transformed_primitive.dart:30-58
  bool instanceIntersectP(uint32_t inst, const Ray& r, Counters* c) const;     // transformed_primitive.dart:60-62
  void setInstances(std::vector<Object>&& objs, std::vector<Instance>&& insts);  // builds the nested accelerators and the bounds

  // ---- BVH ---------------------------------------------------------------
  void buildBVH(int split, int maxPrims);                       // bvh_accel.dart:41-91
  void buildInto(const std::vector<uint32_t>& order, int split, int maxPrims, std::vector<uint32_t>* ordered,
                 std::vector<LinearNode>* nodes) const;
  bool walk(const std::vector<LinearNode>& nodes, const std::vector<uint32_t>& ordered, Ray& ray, Hit* hit, Counters* c) const;
  bool walkP(const std::vector<LinearNode>& nodes, const std::vector<uint32_t>& ordered, const Ray& ray, Counters* c) const;
  bool intersect(Ray& ray, Hit* hit, Counters* c) const;        // bvh_accel.dart:101-165
  bool intersectP(const Ray& ray, Counters* c) const;           // bvh_accel.dart:167-226
  // exhaustive loop in upload order (the pattern of aggregate_test_renderer.dart:82-96)
  bool intersectBrute(Ray& ray, Hit* hit, int* nTies, double* secondT) const;
};

}  // namespace orc
