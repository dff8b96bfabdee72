// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.h header).  PARITY UNPINNED.
#include "ref_scene.h"

#include <cassert>

namespace orc {

// ---------------------------------------------------------------------------
// lib/shapes/triangle.dart:44-98 + lib/core/primitive/geometric_primitive.dart:47-61
// Pure f64 arithmetic on f32 vertex / ray components.
bool Scene::triIntersect(uint32_t tri, Ray& ray, Hit* hit) const {
  Vec p1, p2, p3;
  triVerts(tri, &p1, &p2, &p3);
  double e1x = (double)p2.x - p1.x, e1y = (double)p2.y - p1.y, e1z = (double)p2.z - p1.z;
  double e2x = (double)p3.x - p1.x, e2y = (double)p3.y - p1.y, e2z = (double)p3.z - p1.z;
  double dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
  double s1x = (dy * e2z) - (dz * e2y);
  double s1y = (dz * e2x) - (dx * e2z);
  double s1z = (dx * e2y) - (dy * e2x);
  double divisor = (s1x * e1x) + (s1y * e1y) + (s1z * e1z);
  if (divisor == 0.0) return false;
  double invDivisor = 1.0 / divisor;
  double sx = (double)ray.o.x - p1.x, sy = (double)ray.o.y - p1.y, sz = (double)ray.o.z - p1.z;
  double b1 = (sx * s1x + sy * s1y + sz * s1z) * invDivisor;
  if (b1 < 0.0 || b1 > 1.0) return false;
  double s2x = (sy * e1z) - (sz * e1y);
  double s2y = (sz * e1x) - (sx * e1z);
  double s2z = (sx * e1y) - (sy * e1x);
  double b2 = ((dx * s2x) + (dy * s2y) + (dz * s2z)) * invDivisor;
  if (b2 < 0.0 || b1 + b2 > 1.0) return false;
  double t = (e2x * s2x + e2y * s2y + e2z * s2z) * invDivisor;
  if (t < ray.mint || t > ray.maxt) return false;
  hit->t = t;
  hit->b1 = b1;
  hit->b2 = b2;
  hit->rayEpsilon = 1.0e-3 * t;  // triangle.dart:157
  ray.maxt = t;                  // geometric_primitive.dart:59
  return true;
}

// lib/shapes/triangle.dart:162-194 — every intermediate Vector is rounded to f32.
bool Scene::triIntersectP(uint32_t tri, const Ray& ray) const {
  Vec p1, p2, p3;
  triVerts(tri, &p1, &p2, &p3);
  Vec e1 = p2 - p1;
  Vec e2 = p3 - p1;
  Vec s1 = Cross(ray.d, e2);
  double divisor = Dot(s1, e1);
  if (divisor == 0.0) return false;
  double invDivisor = 1.0 / divisor;
  Vec s = ray.o - p1;
  double b1 = Dot(s, s1) * invDivisor;
  if (b1 < 0.0 || b1 > 1.0) return false;
  Vec s2 = Cross(s, e1);
  double b2 = Dot(ray.d, s2) * invDivisor;
  if (b2 < 0.0 || b1 + b2 > 1.0) return false;
  double t = Dot(e2, s2) * invDivisor;
  if (t < ray.mint || t > ray.maxt) return false;
  return true;
}

// lib/shapes/sphere.dart:39-116 (t, phit, phi, clipping); the differential
// geometry of :118-160 is derived from (phitObj, phi) by the shading code.
static bool sphereCore(const Sphere& s, const Ray& r, bool shadowVariant, double* thitOut, Vec* phitOut,
                       double* phiOut) {
  Ray ray = s.w2o.ray(r);
  double dx = ray.d.x, dy = ray.d.y, dz = ray.d.z, ox = ray.o.x, oy = ray.o.y, oz = ray.o.z;
  double A = dx * dx + dy * dy + dz * dz;
  double B = 2 * (dx * ox + dy * oy + dz * oz);
  double C = ox * ox + oy * oy + oz * oz - s.radius * s.radius;
  double t0, t1;
  if (!Quadratic(A, B, C, &t0, &t1)) return false;
  if (t0 > ray.maxt || t1 < ray.mint) return false;
  double thit = t0;
  if (thit < ray.mint) {
    thit = t1;
    if (thit > ray.maxt) return false;
  }
  Vec phit = ray.at(thit);
  if (phit.x == 0.0f && phit.y == 0.0f) phit.x = f32(1.0e-5 * s.radius);
  double phi = std::atan2((double)phit.y, (double)phit.x);
  if (phi < 0.0) phi += 2.0 * kPi;
  if ((s.zmin > -s.radius && phit.z < s.zmin) || (s.zmax < s.radius && phit.z > s.zmax) || phi > s.phiMax) {
    // sphere.dart:91 compares thit == t1[0]; sphere.dart:210 compares a double with
    // the List `t1` itself, which is always false in Dart -> no early-out for intersectP.
    if (!shadowVariant && thit == t1) return false;
    if (t1 > ray.maxt) return false;
    thit = t1;
    phit = ray.at(thit);
    if (phit.x == 0.0f && phit.y == 0.0f) phit.x = f32(1.0e-5 * s.radius);
    phi = std::atan2((double)phit.y, (double)phit.x);
    if (phi < 0.0) phi += 2.0 * kPi;
    if ((s.zmin > -s.radius && phit.z < s.zmin) || (s.zmax < s.radius && phit.z > s.zmax) || phi > s.phiMax)
      return false;
  }
  *thitOut = thit;
  *phitOut = phit;
  *phiOut = phi;
  return true;
}

// lib/shapes/disk.dart:39-67 (intersect) == :107-140 (intersectP) up to the hit decision
static bool diskCore(const Sphere& s, const Ray& r, double* thitOut, Vec* phitOut, double* phiOut) {
  Ray ray = s.w2o.ray(r);
  if (std::fabs((double)ray.d.z) < 1.0e-7) return false;
  double thit = (s.height - ray.o.z) / ray.d.z;
  if (thit < ray.mint || thit > ray.maxt) return false;
  Vec phit = ray.at(thit);
  double dist2 = (double)phit.x * phit.x + (double)phit.y * phit.y;
  if (dist2 > s.radius * s.radius || dist2 < s.innerRadius * s.innerRadius) return false;
  double phi = std::atan2((double)phit.y, (double)phit.x);
  if (phi < 0) phi += 2.0 * kPi;
  if (phi > s.phiMax) return false;
  *thitOut = thit;
  *phitOut = phit;
  *phiOut = phi;
  return true;
}

// cylinder.dart:39-104 / :153-221, cone.dart:35-98 / :155-214, paraboloid.dart:37-100 / :158-217,
// hyperboloid.dart:57-122 / :178-244: the four quadrics share one decision sequence (quadratic, nearer root in
// range, clip by z and phi, retry at the farther root) and differ in the coefficients, the z range and phi.
static inline double quadricPhi(const Sphere& s, const Vec& phit, double* vOut) {
  if (s.shape == 5) {  // hyperboloid.dart:96-103
    double v = ((double)phit.z - s.hp1.z) / ((double)s.hp2.z - s.hp1.z);
    Vec pr = (s.hp1 * (1.0 - v)) + (s.hp2 * v);
    double phi = std::atan2((double)pr.x * phit.y - (double)phit.x * pr.y, (double)phit.x * pr.x + (double)phit.y * pr.y);
    if (phi < 0.0) phi += 2.0 * kPi;
    *vOut = v;
    return phi;
  }
  double phi = std::atan2((double)phit.y, (double)phit.x);
  if (phi < 0.0) phi += 2.0 * kPi;
  return phi;
}
static bool quadricCore(const Sphere& s, const Ray& r, double* thitOut, Vec* phitOut, double* phiOut, double* vOut) {
  Ray ray = s.w2o.ray(r);
  double dx = ray.d.x, dy = ray.d.y, dz = ray.d.z, ox = ray.o.x, oy = ray.o.y, oz = ray.o.z;
  double A, B, C, zlo = s.zmin, zhi = s.zmax;
  if (s.shape == 2) {  // cylinder.dart:46-52
    A = dx * dx + dy * dy;
    B = 2.0 * (dx * ox + dy * oy);
    C = ox * ox + oy * oy - s.radius * s.radius;
  } else if (s.shape == 3) {  // cone.dart:42-51
    double k = s.radius / s.height;
    k = k * k;
    A = dx * dx + dy * dy - k * dz * dz;
    B = 2.0 * (dx * ox + dy * oy - k * dz * (oz - s.height));
    C = ox * ox + oy * oy - k * (oz - s.height) * (oz - s.height);
    zlo = 0.0;
    zhi = s.height;
  } else if (s.shape == 4) {  // paraboloid.dart:44-50
    double k = s.zmax / (s.radius * s.radius);
    A = k * (dx * dx + dy * dy);
    B = 2 * k * (dx * ox + dy * oy) - dz;
    C = k * (ox * ox + oy * oy) - oz;
  } else {  // hyperboloid.dart:64-72
    double a = s.ha, c = s.hc;
    A = a * dx * dx + a * dy * dy - c * dz * dz;
    B = 2.0 * (a * dx * ox + a * dy * oy - c * dz * oz);
    C = a * ox * ox + a * oy * oy - c * oz * oz - 1;
  }
  double t0, t1;
  if (!Quadratic(A, B, C, &t0, &t1)) return false;
  if (t0 > ray.maxt || t1 < ray.mint) return false;
  double thit = t0;
  if (t0 < ray.mint) {
    thit = t1;
    if (thit > ray.maxt) return false;
  }
  Vec phit = ray.at(thit);
  double v = 0.0;
  double phi = quadricPhi(s, phit, &v);
  if (phit.z < zlo || phit.z > zhi || phi > s.phiMax) {
    if (thit == t1) return false;
    thit = t1;
    if (t1 > ray.maxt) return false;
    phit = ray.at(thit);
    phi = quadricPhi(s, phit, &v);
    if (phit.z < zlo || phit.z > zhi || phi > s.phiMax) return false;
  }
  if (s.shape == 2 || s.shape == 4) v = ((double)phit.z - s.zmin) / (s.zmax - s.zmin);  // cylinder.dart:108, paraboloid.dart:104
  else if (s.shape == 3) v = (double)phit.z / s.height;                                   // cone.dart:102
  *thitOut = thit;
  *phitOut = phit;
  *phiOut = phi;
  *vOut = v;
  return true;
}

bool Scene::sphIntersect(const Sphere& s, Ray& r, Hit* hit) const {
  double thit, phi;
  Vec phit;
  if (s.shape >= 2) {
    double v;
    if (!quadricCore(s, r, &thit, &phit, &phi, &v)) return false;
    hit->t = thit;
    hit->phitObj = phit;
    hit->phi = phi;
    hit->b1 = phi / s.phiMax;
    hit->b2 = v;
    hit->rayEpsilon = 5.0e-4 * thit;  // cylinder.dart:147, cone.dart:149, paraboloid.dart:152, hyperboloid.dart:172
    r.maxt = thit;
    return true;
  }
  if (s.shape == 1) {
    if (!diskCore(s, r, &thit, &phit, &phi)) return false;
    hit->t = thit;
    hit->phitObj = phit;
    hit->phi = phi;
    double dist2 = (double)phit.x * phit.x + (double)phit.y * phit.y;  // disk.dart:69-75 parametric (u, v)
    double oneMinusV = (std::sqrt(dist2) - s.innerRadius) / (s.radius - s.innerRadius);
    hit->b1 = phi / s.phiMax;
    hit->b2 = 1.0 - oneMinusV;
    hit->rayEpsilon = 5.0e-4 * thit;  // disk.dart:100
    r.maxt = thit;
    return true;
  }
  if (!sphereCore(s, r, false, &thit, &phit, &phi)) return false;
  hit->t = thit;
  hit->phitObj = phit;
  hit->phi = phi;
  // sphere.dart:119-121 parametric (u, v)
  double theta = std::acos(clampd((double)phit.z / s.radius, -1.0, 1.0));
  hit->b1 = phi / s.phiMax;
  hit->b2 = (theta - s.thetaMin) / (s.thetaMax - s.thetaMin);
  hit->rayEpsilon = 5.0e-4 * thit;  // sphere.dart:164
  r.maxt = thit;                    // geometric_primitive.dart:59
  return true;
}

bool Scene::sphIntersectP(const Sphere& s, const Ray& r) const {
  double thit, phi;
  Vec phit;
  if (s.shape == 1) return diskCore(s, r, &thit, &phit, &phi);
  if (s.shape >= 2) {
    double v;
    return quadricCore(s, r, &thit, &phit, &phi, &v);
  }
  return sphereCore(s, r, true, &thit, &phit, &phi);
}

// ---------------------------------------------------------------------------
// BVH build, lib/accelerators/bvh_accel.dart:41-91, 228-437.
namespace {

struct PrimInfo {  // bvh_accel.dart:490-501
  uint32_t primitiveNumber;
  Vec centroid;
  BBox bounds;
};

struct BuildNode {  // bvh_accel.dart:508-531
  BBox bounds;
  BuildNode* children[2] = {nullptr, nullptr};
  int splitAxis = 0, firstPrimOffset = 0, nPrimitives = 0;
};

// lib/core/common.dart:256-284
template <class Pred>
int partitionRef(std::vector<PrimInfo>& list, Pred pred, int first, int last) {
  while (first < last) {
    while (pred(list[first])) {
      ++first;
      if (first == last) return first;
    }
    do {
      --last;
      if (first == last) return first;
    } while (!pred(list[last]));
    std::swap(list[first], list[last]);
    ++first;
  }
  return first;
}

// lib/core/common.dart:289-297: copy the range, List.sort with a comparator that
// returns -1 when a<b and +1 otherwise (never 0), copy back.  Dart's List.sort
// (sdk/lib/internal/sort.dart, NOT in /root/reference — restated from its published
// algorithm) uses insertion sort for <= 32 elements:
//   for i in left+1..right: el=a[i]; j=i; while (j>left && compare(a[j-1], el) > 0) {a[j]=a[j-1]; j--;} a[j]=el;
// and a dual-pivot quicksort above that.  The SAH path only sorts ranges of <= 4
// elements, so the insertion sort is exact; larger ranges (splitmethod middle/equal)
// fall back to the same insertion sort here (O(n^2), unpinned ordering of equal keys).
void nthElementRef(std::vector<PrimInfo>& list, int first, int last, int dim) {
  for (int i = first + 1; i < last; ++i) {
    PrimInfo el = list[i];
    int j = i;
    // compare(a[j-1], el) > 0  <=>  !(a[j-1].centroid[dim] < el.centroid[dim])
    while (j > first && !(list[j - 1].centroid[dim] < el.centroid[dim])) {
      list[j] = list[j - 1];
      --j;
    }
    list[j] = el;
  }
}

struct Builder {
  const Scene& sc;
  int splitMethod, maxPrimsInNode;
  std::vector<PrimInfo> buildData;
  std::vector<uint32_t> ordered;
  std::vector<BuildNode*> pool;
  int totalNodes = 0;

  BuildNode* leaf(BuildNode* node, int start, int end, const BBox& bbox) {
    int first = (int)ordered.size();
    for (int i = start; i < end; ++i) ordered.push_back(buildData[i].primitiveNumber);
    node->firstPrimOffset = first;
    node->nPrimitives = end - start;
    node->bounds = bbox;
    return node;
  }

  BuildNode* recursiveBuild(int start, int end) {  // bvh_accel.dart:228-417
    assert(start != end);
    totalNodes++;
    BuildNode* node = new BuildNode();
    pool.push_back(node);
    BBox bbox;
    for (int i = start; i < end; ++i) bbox = Union(bbox, buildData[i].bounds);
    int nPrimitives = end - start;
    if (nPrimitives == 1) return leaf(node, start, end, bbox);

    BBox centroidBounds;
    for (int i = start; i < end; ++i) centroidBounds = UnionPoint(centroidBounds, buildData[i].centroid);
    int dim = centroidBounds.maximumExtent();
    int mid = (start + end) / 2;
    if (centroidBounds.pMax[dim] == centroidBounds.pMin[dim]) return leaf(node, start, end, bbox);

    const double cmin = centroidBounds.pMin[dim], cmax = centroidBounds.pMax[dim];
    bool fallthroughEqual = false;
    switch (splitMethod) {
      case 0: {  // SPLIT_MIDDLE, :282-304
        double pmid = 0.5 * (cmin + cmax);
        mid = partitionRef(buildData, [&](const PrimInfo& a) { return a.centroid[dim] < pmid; }, start, end);
        if (mid != start && mid != end) break;
        fallthroughEqual = true;
      }
      // fallthrough
      case 1: {  // SPLIT_EQUAL_COUNTS, :305-309
        (void)fallthroughEqual;
        mid = (start + end) / 2;
        nthElementRef(buildData, start, end, dim);
        break;
      }
      case 2:
      default: {  // SPLIT_SAH, :310-404
        if (nPrimitives <= 4) {
          mid = (start + end) / 2;
          nthElementRef(buildData, start, end, dim);
        } else {
          const int nBuckets = 12;
          int count[nBuckets] = {0};
          BBox bounds[nBuckets];
          for (int i = start; i < end; ++i) {
            int b = (int)(nBuckets * ((buildData[i].centroid[dim] - cmin) / (cmax - cmin)));  // .toInt()
            if (b == nBuckets) b = nBuckets - 1;
            count[b]++;
            bounds[b] = Union(bounds[b], buildData[i].bounds);
          }
          float cost[nBuckets - 1];  // Float32List, :345
          for (int i = 0; i < nBuckets - 1; ++i) {
            BBox b0, b1;
            int count0 = 0, count1 = 0;
            for (int j = 0; j <= i; ++j) { b0 = Union(b0, bounds[j]); count0 += count[j]; }
            for (int j = i + 1; j < nBuckets; ++j) { b1 = Union(b1, bounds[j]); count1 += count[j]; }
            cost[i] = f32(0.125 + (count0 * b0.surfaceArea() + count1 * b1.surfaceArea()) / bbox.surfaceArea());
          }
          double minCost = cost[0];
          int minCostSplit = 0;
          for (int i = 1; i < nBuckets - 1; ++i) {
            if (cost[i] < minCost) { minCost = cost[i]; minCostSplit = i; }
          }
          if (nPrimitives > maxPrimsInNode || minCost < nPrimitives) {
            mid = partitionRef(buildData,
                               [&](const PrimInfo& p) {
                                 int b = (int)std::floor(nBuckets * ((p.centroid[dim] - cmin) / (cmax - cmin)));
                                 if (b == nBuckets) b = nBuckets - 1;
                                 return b <= minCostSplit;
                               },
                               start, end);
          } else {
            return leaf(node, start, end, bbox);
          }
        }
        break;
      }
    }
    // :407-411 — the SECOND child is built first (affects leaf order in `ordered`).
    BuildNode* c2 = recursiveBuild(mid, end);
    BuildNode* c1 = recursiveBuild(start, mid);
    node->children[0] = c1;
    node->children[1] = c2;
    node->bounds = Union(c1->bounds, c2->bounds);  // :521
    node->splitAxis = dim;
    node->nPrimitives = 0;
    return node;
  }

  int flatten(BuildNode* node, std::vector<LinearNode>& nodes, int* offset) {  // :419-437
    LinearNode& ln = nodes[*offset];
    ln.bounds = node->bounds;
    int myOffset = (*offset)++;
    if (node->nPrimitives > 0) {
      ln.offset = node->firstPrimOffset;
      ln.nPrimitives = node->nPrimitives;
    } else {
      ln.axis = node->splitAxis;
      ln.nPrimitives = 0;
      flatten(node->children[0], nodes, offset);
      int second = flatten(node->children[1], nodes, offset);
      nodes[myOffset].offset = second;
    }
    return myOffset;
  }
};

}  // namespace

void Scene::buildInto(const std::vector<uint32_t>& order, int split, int maxPrims, std::vector<uint32_t>* orderedOut,
                      std::vector<LinearNode>* nodesOut) const {
  nodesOut->clear();
  orderedOut->clear();
  const uint32_t n = (uint32_t)order.size();
  if (n == 0) return;
  Builder b{*this, split, std::min(255, maxPrims), {}, {}, {}, 0};  // :44
  b.buildData.resize(n);
  for (uint32_t i = 0; i < n; ++i) {  // :59-65
    const uint32_t prim = order[i];
    b.buildData[i].primitiveNumber = prim;
    b.buildData[i].bounds = primBound(prim);
    b.buildData[i].centroid = b.buildData[i].bounds.center();
  }
  b.ordered.reserve(n);
  BuildNode* root = b.recursiveBuild(0, (int)n);
  nodesOut->resize(b.totalNodes);
  int off = 0;
  b.flatten(root, *nodesOut, &off);
  assert(off == b.totalNodes);
  orderedOut->swap(b.ordered);
  for (BuildNode* p : b.pool) delete p;
}

void Scene::buildBVH(int split, int maxPrims) {
  splitMethod = split;
  maxPrimsInNode = std::min(255, maxPrims);  // :44
  std::vector<uint32_t> order = buildOrder;
  if (order.empty()) {  // upload order; with instances the caller always gives the top-level order
    order.resize(nprims());
    for (uint32_t i = 0; i < nprims(); ++i) order[i] = i;
  }
  buildInto(order, splitMethod, maxPrimsInNode, &ordered, &nodes);
}

// lib/accelerators/bvh_accel.dart:439-472
static inline bool slab(const BBox& bounds, const Ray& ray, const Vec& invDir, const int dirIsNeg[3]) {
  double tmin = ((double)bounds[dirIsNeg[0]].x - ray.o.x) * invDir.x;
  double tmax = ((double)bounds[1 - dirIsNeg[0]].x - ray.o.x) * invDir.x;
  double tymin = ((double)bounds[dirIsNeg[1]].y - ray.o.y) * invDir.y;
  double tymax = ((double)bounds[1 - dirIsNeg[1]].y - ray.o.y) * invDir.y;
  if ((tmin > tymax) || (tymin > tmax)) return false;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  double tzmin = ((double)bounds[dirIsNeg[2]].z - ray.o.z) * invDir.z;
  double tzmax = ((double)bounds[1 - dirIsNeg[2]].z - ray.o.z) * invDir.z;
  if ((tmin > tzmax) || (tzmin > tmax)) return false;
  if (tzmin > tmin) tmin = tzmin;
  if (tzmax < tmax) tmax = tzmax;
  return (tmin < ray.maxt) && (tmax > ray.mint);
}

// lib/accelerators/bvh_accel.dart:101-165
bool Scene::intersect(Ray& ray, Hit* hit, Counters* c) const { return walk(nodes, ordered, ray, hit, c); }
bool Scene::walk(const std::vector<LinearNode>& nodes, const std::vector<uint32_t>& ordered, Ray& ray, Hit* hit, Counters* c) const {
  if (nodes.empty()) return false;
  bool any = false;
  Vec invDir(1.0 / (double)ray.d.x, 1.0 / (double)ray.d.y, 1.0 / (double)ray.d.z);  // f32-rounded, :109-111
  int dirIsNeg[3] = {invDir.x < 0 ? 1 : 0, invDir.y < 0 ? 1 : 0, invDir.z < 0 ? 1 : 0};
  int todoOffset = 0, nodeNum = 0;
  uint32_t todo[64];
  while (true) {
    const LinearNode& node = nodes[nodeNum];
    if (c) c->nodes_visited++;
    if (slab(node.bounds, ray, invDir, dirIsNeg)) {
      if (node.nPrimitives > 0) {
        for (int i = 0; i < node.nPrimitives; ++i) {
          if (c) c->prims_tested++;
          if (primIntersect(ordered[node.offset + i], ray, hit, c)) any = true;
        }
        if (todoOffset == 0) break;
        nodeNum = todo[--todoOffset];
      } else {
        if (dirIsNeg[node.axis] != 0) {
          todo[todoOffset++] = nodeNum + 1;
          nodeNum = node.offset;
        } else {
          todo[todoOffset++] = node.offset;
          nodeNum = nodeNum + 1;
        }
      }
    } else {
      if (todoOffset == 0) break;
      nodeNum = todo[--todoOffset];
    }
  }
  return any;
}

// lib/accelerators/bvh_accel.dart:167-226
bool Scene::intersectP(const Ray& ray, Counters* c) const { return walkP(nodes, ordered, ray, c); }
bool Scene::walkP(const std::vector<LinearNode>& nodes, const std::vector<uint32_t>& ordered, const Ray& ray, Counters* c) const {
  if (nodes.empty()) return false;
  Vec invDir(1.0 / (double)ray.d.x, 1.0 / (double)ray.d.y, 1.0 / (double)ray.d.z);
  int dirIsNeg[3] = {invDir.x < 0 ? 1 : 0, invDir.y < 0 ? 1 : 0, invDir.z < 0 ? 1 : 0};
  int todoOffset = 0, nodeNum = 0;
  uint32_t todo[64];
  while (true) {
    const LinearNode& node = nodes[nodeNum];
    if (c) c->nodes_visited++;
    if (slab(node.bounds, ray, invDir, dirIsNeg)) {
      if (node.nPrimitives > 0) {
        for (int i = 0; i < node.nPrimitives; ++i) {
          if (c) c->prims_tested++;
          if (primIntersectP(ordered[node.offset + i], ray, c)) return true;
        }
        if (todoOffset == 0) break;
        nodeNum = todo[--todoOffset];
      } else {
        if (dirIsNeg[node.axis] != 0) {
          todo[todoOffset++] = nodeNum + 1;
          nodeNum = node.offset;
        } else {
          todo[todoOffset++] = node.offset;
          nodeNum = nodeNum + 1;
        }
      }
    } else {
      if (todoOffset == 0) break;
      nodeNum = todo[--todoOffset];
    }
  }
  return false;
}

// transformed_primitive.dart:30-58.  The Intersection's differential geometry is moved to world space by the renderer side
// (ref_render.cpp fillIsect), which repeats interpolate(r.time) for the instance the hit names.
bool Scene::instanceIntersect(uint32_t inst, Ray& r, Hit* hit, Counters* c) const {
  const Instance& in = instances[inst];
  const Object& ob = objects[in.object];
  const Transform w2p = in.worldToPrimitive.interpolate(r.time);
  Ray ray(w2p.point(r.o), w2p.vector(r.d), r.mint, r.maxt, r.time, r.depth);  // transform.dart:180-196
  bool h;
  if (ob.order.size() == 1) {
    if (c) c->prims_tested++;
    h = primIntersect(ob.order[0], ray, hit);
  } else {
    h = walk(ob.nodes, ob.ordered, ray, hit, c);
  }
  if (!h) return false;
  r.maxt = ray.maxt;
  hit->inst = (int32_t)inst;
  return true;
}
// transformed_primitive.dart:60-62 over AnimatedTransform.transformRay (animated_transform.dart:138-154: the same three cases)
bool Scene::instanceIntersectP(uint32_t inst, const Ray& r, Counters* c) const {
  const Instance& in = instances[inst];
  const Object& ob = objects[in.object];
  const Transform w2p = in.worldToPrimitive.interpolate(r.time);
  const Ray ray(w2p.point(r.o), w2p.vector(r.d), r.mint, r.maxt, r.time, r.depth);
  if (ob.order.size() == 1) {
    if (c) c->prims_tested++;
    return primIntersectP(ob.order[0], ray);
  }
  return walkP(ob.nodes, ob.ordered, ray, c);
}

void Scene::setInstances(std::vector<Object>&& objs, std::vector<Instance>&& insts) {
  objects = std::move(objs);
  instances = std::move(insts);
  for (Object& ob : objects)
    if (ob.order.size() > 1) buildInto(ob.order, ob.split, ob.maxPrims, &ob.ordered, &ob.nodes);
  for (Instance& in : instances) in.bound = in.worldToPrimitive.motionBounds(objects[in.object].worldBound(*this), true);
}

// Exhaustive differential check in the style of aggregate_test_renderer.dart:82-96:
// every primitive in upload order, same shrinking ray.maxDistance.  Also reports how
// many primitives hit at exactly the winning t (tie set size) and the runner-up t.
bool Scene::intersectBrute(Ray& ray, Hit* hit, int* nTies, double* secondT) const {
  bool any = false;
  const double mint = ray.mint, maxt0 = ray.maxt;
  // the top-level primitives in upload order (with instances: the geometric primitives the build order names, then the instances)
  std::vector<uint32_t> top = buildOrder;
  if (top.empty()) { top.resize(nprims()); for (uint32_t i = 0; i < nprims(); ++i) top[i] = i; }
  std::sort(top.begin(), top.end());
  for (uint32_t p : top)
    if (primIntersect(p, ray, hit)) any = true;
  if (nTies || secondT) {
    int ties = 0;
    double second = kInf;
    if (any) {
      for (uint32_t p : top) {
        Ray r2 = ray;
        r2.mint = mint;
        r2.maxt = maxt0;
        Hit h2;
        if (primIntersect(p, r2, &h2)) {
          if (h2.t == hit->t) ties++;
          else if (h2.t < second) second = h2.t;
        }
      }
    }
    if (nTies) *nTies = ties;
    if (secondT) *secondT = second;
  }
  return any;
}

}  // namespace orc
