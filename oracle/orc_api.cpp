// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.h header).  PARITY UNPINNED.
//
// C entry points of the CPU oracle (liboracle.so).  They mirror include/drt.h one for one with an
// `orc_` prefix so tests can drive both libraries with the same arguments.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <string>
#include <thread>

#include "../include/drt.h"  // the plain-data structs drt_texture / drt_material_program only
#include "ref_render.h"
#include "ref_scene.h"

using namespace orc;

struct orc_ctx {
  Scene scene;
  RenderScene rs;
  Counters counters;
  std::string err;
  double buildSeconds = 0;
  std::vector<double> rayTimes;  // orc_set_ray_times
};

struct orc_hit {
  float t, b1, b2;
  int32_t prim;
};

template <class F>
static void parallelFor(uint64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n < 1024) {
    f(0, 0, n);
    return;
  }
  std::vector<std::thread> th;
  for (int k = 0; k < nthreads; ++k) {
    uint64_t b = n * k / nthreads, e = n * (k + 1) / nthreads;
    th.emplace_back([=] { f(k, b, e); });
  }
  for (auto& t : th) t.join();
}

extern "C" {

orc_ctx* orc_create() { return new orc_ctx(); }
void orc_destroy(orc_ctx* c) { delete c; }
const char* orc_last_error(const orc_ctx* c) { return c ? c->err.c_str() : ""; }

int orc_set_triangles(orc_ctx* c, const float* P, uint32_t nverts, const uint32_t* idx, uint32_t ntris,
                      const int32_t* mat, const int32_t* light, const uint8_t* rev) {
  Scene& s = c->scene;
  for (uint64_t i = 0; i < (uint64_t)ntris * 3; ++i)
    if (idx[i] >= nverts) { c->err = "triangle index out of range"; return -1; }
  s.P.assign(P, P + (size_t)nverts * 3);
  s.idx.assign(idx, idx + (size_t)ntris * 3);
  s.vertN.clear(); s.vertS.clear(); s.vertUV.clear(); s.meshOfTri.clear(); s.meshes.clear();
  size_t np = s.nprims();
  s.materialOf.resize(np, 0);
  s.lightOf.resize(np, -1);
  s.reverseOf.resize(np, 0);
  for (uint32_t i = 0; i < ntris; ++i) {
    s.materialOf[i] = mat ? mat[i] : 0;
    s.lightOf[i] = light ? light[i] : -1;
    s.reverseOf[i] = rev ? rev[i] : 0;
  }
  return 0;
}

// Per-vertex shading attributes (triangle_mesh.dart:24-28, triangle.dart:246-262,271-364): N / S in object space and uv,
// each nverts long or NULL; mesh_of_tri names the mesh (objectToWorld + which attributes it has: bit 0 N, 1 S, 2 uv).
int orc_set_mesh_shading(orc_ctx* c, const float* N, const float* S, const float* uv, const uint32_t* meshOfTri, uint32_t nmeshes,
                         const float* o2w, const float* w2o, const uint8_t* flags) {
  Scene& s = c->scene;
  s.vertN.clear(); s.vertS.clear(); s.vertUV.clear(); s.meshOfTri.clear(); s.meshes.clear();
  if (nmeshes == 0 || !meshOfTri) return 0;
  size_t nv = s.P.size() / 3;
  if (N) s.vertN.assign(N, N + 3 * nv);
  if (S) s.vertS.assign(S, S + 3 * nv);
  if (uv) s.vertUV.assign(uv, uv + 2 * nv);
  s.meshOfTri.assign(meshOfTri, meshOfTri + s.ntris());
  for (uint32_t t = 0; t < s.ntris(); ++t)
    if (meshOfTri[t] >= nmeshes) { c->err = "mesh index out of range"; return -1; }
  for (uint32_t m = 0; m < nmeshes; ++m) {
    Scene::MeshInfo mi;
    mi.o2w = Transform(o2w + 16 * m, w2o + 16 * m);
    mi.hasN = (flags[m] & 1) && N;
    mi.hasS = (flags[m] & 2) && S;
    mi.hasUV = (flags[m] & 4) && uv;
    s.meshes.push_back(mi);
  }
  return 0;
}

int orc_set_spheres(orc_ctx* c, uint32_t n, const float* o2w, const float* w2o, const double* prm, const int32_t* mat,
                    const int32_t* light, const uint8_t* rev) {
  Scene& s = c->scene;
  s.spheres.clear();
  uint32_t nt = s.ntris();
  s.materialOf.resize(nt + n, 0);
  s.lightOf.resize(nt + n, -1);
  s.reverseOf.resize(nt + n, 0);
  for (uint32_t i = 0; i < n; ++i) {
    s.spheres.emplace_back(o2w + 16 * i, w2o + 16 * i, prm[4 * i], prm[4 * i + 1], prm[4 * i + 2], prm[4 * i + 3],
                           rev ? rev[i] != 0 : false);
    s.materialOf[nt + i] = mat ? mat[i] : 0;
    s.lightOf[nt + i] = light ? light[i] : -1;
    s.reverseOf[nt + i] = rev ? rev[i] : 0;
  }
  return 0;
}

// Disks (lib/shapes/disk.dart) share the quadric id range: they are appended after the spheres, so call this after
// orc_set_spheres (which resets the range).  prm: n x 4 doubles height, radius, innerradius, phimax(degrees).
int orc_set_disks(orc_ctx* c, uint32_t n, const float* o2w, const float* w2o, const double* prm, const int32_t* mat,
                  const int32_t* light, const uint8_t* rev) {
  Scene& s = c->scene;
  uint32_t base = s.nprims();
  s.materialOf.resize(base + n, 0);
  s.lightOf.resize(base + n, -1);
  s.reverseOf.resize(base + n, 0);
  for (uint32_t i = 0; i < n; ++i) {
    s.spheres.push_back(Sphere::makeDisk(o2w + 16 * i, w2o + 16 * i, prm[4 * i], prm[4 * i + 1], prm[4 * i + 2], prm[4 * i + 3],
                                         rev ? rev[i] != 0 : false));
    s.materialOf[base + i] = mat ? mat[i] : 0;
    s.lightOf[base + i] = light ? light[i] : -1;
    s.reverseOf[base + i] = rev ? rev[i] : 0;
  }
  return 0;
}

// Cylinders / cones / paraboloids / hyperboloids (kind 2..5) join the same quadric id range, appended in call order.
// prm: n x 8 doubles, see Sphere::makeQuadric.
int orc_set_quadrics(orc_ctx* c, int kind, uint32_t n, const float* o2w, const float* w2o, const double* prm, const int32_t* mat,
                     const int32_t* light, const uint8_t* rev) {
  if (kind < 2 || kind > 5) return -1;
  Scene& s = c->scene;
  uint32_t base = s.nprims();
  s.materialOf.resize(base + n, 0);
  s.lightOf.resize(base + n, -1);
  s.reverseOf.resize(base + n, 0);
  for (uint32_t i = 0; i < n; ++i) {
    s.spheres.push_back(Sphere::makeQuadric(kind, o2w + 16 * i, w2o + 16 * i, prm + 8 * i, rev ? rev[i] != 0 : false));
    s.materialOf[base + i] = mat ? mat[i] : 0;
    s.lightOf[base + i] = light ? light[i] : -1;
    s.reverseOf[base + i] = rev ? rev[i] : 0;
  }
  return 0;
}

int orc_set_build_order(orc_ctx* c, const uint32_t* ids, uint32_t n) {
  if (!ids) { c->scene.buildOrder.clear(); return 0; }
  c->scene.buildOrder.assign(ids, ids + n);
  return 0;
}

// mirrors drt_set_instances (include/drt.h): objects = the aggregates TransformedPrimitives wrap, instances = the TransformedPrimitives
int orc_set_instances(orc_ctx* c, uint32_t nObjects, const uint32_t* objectOffsets, const uint32_t* objectPrims, const int32_t* objectSplit,
                      const int32_t* objectMaxPrims, uint32_t nInstances, const uint32_t* instanceObject, const float* startM,
                      const float* startMInv, const float* endM, const float* endMInv, const double* times) {
  Scene& s = c->scene;
  std::vector<Scene::Object> objs(nObjects);
  for (uint32_t i = 0; i < nObjects; ++i) {
    if (objectOffsets[i + 1] <= objectOffsets[i]) { c->err = "an object holds at least one primitive"; return -1; }
    objs[i].order.assign(objectPrims + objectOffsets[i], objectPrims + objectOffsets[i + 1]);
    for (uint32_t id : objs[i].order)
      if (id >= s.nprims()) { c->err = "object primitive id out of range"; return -1; }
    objs[i].split = objectSplit ? objectSplit[i] : 2;
    objs[i].maxPrims = objectMaxPrims ? objectMaxPrims[i] : 1;
  }
  std::vector<Scene::Instance> insts(nInstances);
  for (uint32_t i = 0; i < nInstances; ++i) {
    if (instanceObject[i] >= nObjects) { c->err = "instance names an object that was not defined"; return -1; }
    insts[i].object = instanceObject[i];
    insts[i].worldToPrimitive.init(Transform(startM + 16 * i, startMInv + 16 * i), times ? times[2 * i] : 0.0,
                                   Transform(endM + 16 * i, endMInv + 16 * i), times ? times[2 * i + 1] : 1.0);
  }
  s.setInstances(std::move(objs), std::move(insts));
  return 0;
}

// times of the rays of the following orc_trace_* calls (ray i travels at times[i]; NULL: every ray at time 0)
int orc_set_ray_times(orc_ctx* c, const double* times, uint64_t n) {
  if (times) c->rayTimes.assign(times, times + n);
  else c->rayTimes.clear();
  return 0;
}

// AnimatedTransform probes for the tests: Decompose of instance `inst`'s two transforms (T 2 x 3, R 2 x 4 as x y z w, S 2 x 16),
// interpolate(time) (m, mInv) and the instance's world bound
int orc_instance_probe(const orc_ctx* c, uint32_t inst, double time, double* T, double* R, float* S, float* m, float* mInv, float* bound,
                       int32_t* animated) {
  const Scene& s = c->scene;
  if (inst >= s.instances.size()) return -1;
  const AnimatedTransform& a = s.instances[inst].worldToPrimitive;
  for (int k = 0; k < 2; ++k) {
    if (T) { T[3 * k] = a.T[k].x; T[3 * k + 1] = a.T[k].y; T[3 * k + 2] = a.T[k].z; }
    if (R) { R[4 * k] = a.R[k].v.x; R[4 * k + 1] = a.R[k].v.y; R[4 * k + 2] = a.R[k].v.z; R[4 * k + 3] = a.R[k].w; }
    if (S) std::memcpy(S + 16 * k, a.S[k].d, 64);
  }
  const Transform t = a.interpolate(time);
  if (m) std::memcpy(m, t.m, 64);
  if (mInv) std::memcpy(mInv, t.mInv, 64);
  const BBox& b = s.instances[inst].bound;
  if (bound) { bound[0] = b.pMin.x; bound[1] = b.pMin.y; bound[2] = b.pMin.z; bound[3] = b.pMax.x; bound[4] = b.pMax.y; bound[5] = b.pMax.z; }
  if (animated) *animated = a.actuallyAnimated ? 1 : 0;
  return 0;
}

int orc_build_bvh(orc_ctx* c, int split, int maxPrims) {
  Scene& s = c->scene;
  if (s.instances.empty() && !s.buildOrder.empty() && s.buildOrder.size() != s.nprims()) { c->err = "build order size mismatch"; return -1; }
  if (!s.instances.empty()) {  // the top-level order names geometric primitives outside every object and the instances
    if (s.buildOrder.empty()) { c->err = "a scene with instances needs the top-level build order"; return -1; }
    for (uint32_t id : s.buildOrder)
      if (id >= s.nprims() + s.instances.size()) { c->err = "build order id out of range"; return -1; }
  }
  auto t0 = std::chrono::steady_clock::now();
  s.buildBVH(split, maxPrims);
  c->buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

double orc_build_seconds(const orc_ctx* c) { return c->buildSeconds; }
uint32_t orc_bvh_num_nodes(const orc_ctx* c) { return (uint32_t)c->scene.nodes.size(); }
uint32_t orc_num_prims(const orc_ctx* c) { return c->scene.nprims(); }

int orc_bvh_export(const orc_ctx* c, float* bounds, int32_t* offset, int32_t* nprims, int32_t* axis, uint32_t* ordered) {
  const Scene& s = c->scene;
  for (size_t i = 0; i < s.nodes.size(); ++i) {
    const LinearNode& n = s.nodes[i];
    if (bounds) {
      float* b = bounds + 6 * i;
      b[0] = n.bounds.pMin.x; b[1] = n.bounds.pMin.y; b[2] = n.bounds.pMin.z;
      b[3] = n.bounds.pMax.x; b[4] = n.bounds.pMax.y; b[5] = n.bounds.pMax.z;
    }
    if (offset) offset[i] = n.offset;
    if (nprims) nprims[i] = n.nPrimitives;
    if (axis) axis[i] = n.nPrimitives > 0 ? 0 : n.axis;
  }
  if (ordered)
    for (size_t i = 0; i < s.ordered.size(); ++i) ordered[i] = s.ordered[i];
  return 0;
}

static inline Ray makeRay(const float* o, const float* d) {
  Ray r;
  r.o.x = o[0]; r.o.y = o[1]; r.o.z = o[2];
  r.d.x = d[0]; r.d.y = d[1]; r.d.z = d[2];
  r.mint = o[3];
  r.maxt = d[3];
  return r;
}
static inline void storeHit(orc_hit* out, bool h, const Hit& hit) {
  if (h) {
    out->t = (float)hit.t; out->b1 = (float)hit.b1; out->b2 = (float)hit.b2; out->prim = hit.prim;
  } else {
    out->t = std::numeric_limits<float>::infinity(); out->b1 = 0; out->b2 = 0; out->prim = -1;
  }
}

// t64 (optional): the un-rounded f64 tHit per ray (+inf on miss).
int orc_trace_closest(orc_ctx* c, const float* o, const float* d, uint64_t n, orc_hit* hits, double* t64, int nthreads) {
  const Scene& s = c->scene;
  std::vector<Counters> cs(std::max(1, nthreads));
  parallelFor(n, nthreads, [&](int k, uint64_t b, uint64_t e) {
    Counters cc;
    for (uint64_t i = b; i < e; ++i) {
      Ray r = makeRay(o + 4 * i, d + 4 * i);
      if (i < c->rayTimes.size()) r.time = c->rayTimes[i];
      Hit hit;
      bool h = s.intersect(r, &hit, &cc);
      storeHit(&hits[i], h, hit);
      if (t64) t64[i] = h ? hit.t : kInf;
    }
    cc.rays = e - b;
    cs[k] = cc;
  });
  c->counters = Counters();
  for (auto& x : cs) c->counters.add(x);
  return 0;
}

int orc_trace_any(orc_ctx* c, const float* o, const float* d, uint64_t n, uint8_t* occluded, int nthreads) {
  const Scene& s = c->scene;
  std::vector<Counters> cs(std::max(1, nthreads));
  parallelFor(n, nthreads, [&](int k, uint64_t b, uint64_t e) {
    Counters cc;
    for (uint64_t i = b; i < e; ++i) {
      Ray r = makeRay(o + 4 * i, d + 4 * i);
      if (i < c->rayTimes.size()) r.time = c->rayTimes[i];
      occluded[i] = s.intersectP(r, &cc) ? 1 : 0;
    }
    cc.rays = e - b;
    cs[k] = cc;
  });
  c->counters = Counters();
  for (auto& x : cs) c->counters.add(x);
  return 0;
}

// Exhaustive closest hit (no BVH), plus per-ray tie-set size and runner-up t (may be NULL).
int orc_trace_closest_brute(orc_ctx* c, const float* o, const float* d, uint64_t n, orc_hit* hits, int32_t* nties,
                            double* second_t, int nthreads) {
  const Scene& s = c->scene;
  parallelFor(n, nthreads, [&](int, uint64_t b, uint64_t e) {
    for (uint64_t i = b; i < e; ++i) {
      Ray r = makeRay(o + 4 * i, d + 4 * i);
      if (i < c->rayTimes.size()) r.time = c->rayTimes[i];
      Hit hit;
      int ties = 0;
      double sec = kInf;
      bool h = s.intersectBrute(r, &hit, nties ? &ties : nullptr, second_t ? &sec : nullptr);
      storeHit(&hits[i], h, hit);
      if (nties) nties[i] = ties;
      if (second_t) second_t[i] = sec;
    }
  });
  return 0;
}

int orc_trace_any_brute(orc_ctx* c, const float* o, const float* d, uint64_t n, uint8_t* occluded, int nthreads) {
  const Scene& s = c->scene;
  parallelFor(n, nthreads, [&](int, uint64_t b, uint64_t e) {
    for (uint64_t i = b; i < e; ++i) {
      Ray r = makeRay(o + 4 * i, d + 4 * i);
      if (i < c->rayTimes.size()) r.time = c->rayTimes[i];
      uint8_t occ = 0;
      for (uint32_t p = 0; p < s.nprims() && !occ; ++p) occ = s.primIntersectP(p, r) ? 1 : 0;
      occluded[i] = occ;
    }
  });
  return 0;
}

// counters of the last trace call: rays, nodes_visited, prims_tested
int orc_get_counters(const orc_ctx* c, uint64_t out[3]) {
  out[0] = c->counters.rays;
  out[1] = c->counters.nodes_visited;
  out[2] = c->counters.prims_tested;
  return 0;
}

// ---- renderer (mirrors drt_set_materials ... drt_film_read) -------------------------------------
int orc_set_materials(orc_ctx* c, uint32_t n, const int32_t* kind, const float* kd, const float* sigma) {
  c->rs.materials.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    if (kind && kind[i] != 0) return -1;
    c->rs.materials[i] = Material::matte(Spec(kd[3 * i], kd[3 * i + 1], kd[3 * i + 2]), sigma ? sigma[i] : 0.0);
  }
  return 0;
}

// mirrors drt_set_material_lobes (include/drt.h): material i owns lobes [offsets[i], offsets[i + 1])
int orc_set_material_lobes(orc_ctx* c, uint32_t n, const uint32_t* offsets, const int32_t* kind, const float* rgb,
                           const int32_t* fresnel, const float* eta, const float* k, const double* scalars) {
  c->rs.materials.assign(n, Material());
  for (uint32_t i = 0; i < n; ++i) {
    if (offsets[i + 1] - offsets[i] > 8) return -1;
    for (uint32_t j = offsets[i]; j < offsets[i + 1]; ++j) {
      Lobe l;
      l.kind = kind[j];
      l.R = Spec(rgb[3 * j], rgb[3 * j + 1], rgb[3 * j + 2]);
      l.fresnel = fresnel ? fresnel[j] : 0;
      if (eta) l.eta = Spec(eta[3 * j], eta[3 * j + 1], eta[3 * j + 2]);
      if (k) l.k = Spec(k[3 * j], k[3 * j + 1], k[3 * j + 2]);
      l.param = scalars[3 * j];
      l.ei = scalars[3 * j + 1];
      l.et = scalars[3 * j + 2];
      c->rs.materials[i].lobes.push_back(l);
    }
  }
  return 0;
}

int orc_set_lights(orc_ctx* c, uint32_t n, const int32_t* kind, const float* L, const float* pos, const int32_t* nsamples,
                   const uint32_t* shape_offsets, const uint32_t* shape_prims) {
  c->rs.lights.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    Light& l = c->rs.lights[i];
    l.kind = kind[i];
    l.L = Spec(L[3 * i], L[3 * i + 1], L[3 * i + 2]);
    if (pos) { l.pos.x = pos[3 * i]; l.pos.y = pos[3 * i + 1]; l.pos.z = pos[3 * i + 2]; }
    l.nSamples = nsamples ? nsamples[i] : 1;
    l.shapes.clear();
    if (shape_offsets && shape_prims)
      for (uint32_t k = shape_offsets[i]; k < shape_offsets[i + 1]; ++k) l.shapes.push_back(shape_prims[k]);
  }
  return 0;
}

// mirrors drt_set_light_map: the map (level 0, power-of-two, or NULL) and transforms of a projection (kind 5) or goniometric
// (kind 6) light
int orc_set_light_map(orc_ctx* c, uint32_t index, int width, int height, const float* rgb, const float* w2l, const float* proj,
                      const double* screen, double hither) {
  if (index >= c->rs.lights.size() || (c->rs.lights[index].kind != 5 && c->rs.lights[index].kind != 6)) { c->err = "not a projection / goniometric light"; return -1; }
  Light& l = c->rs.lights[index];
  l.worldToLight = Transform(w2l, w2l);  // only m is used (vector)
  if (rgb) {
    if (width < 1 || height < 1 || (width & (width - 1)) || (height & (height - 1))) { c->err = "map resolution must be a power of two"; return -1; }
    l.radianceMap.init(width, height, rgb);
  } else {
    l.radianceMap = MipMap();
  }
  if (l.kind == 5) {
    l.lightProjection = Transform(proj, proj);
    for (int k = 0; k < 4; ++k) l.screen[k] = screen[k];
    l.hither = hither;
  }
  return 0;
}

// mirrors drt_set_lobe_wrappers: BRDFToBTDF / ScaledBxDF around the lobes of the last orc_set_material_lobes, in lobe order
int orc_set_lobe_wrappers(orc_ctx* c, uint32_t nLobes, const int32_t* wrap, const float* scale) {
  uint32_t k = 0;
  for (Material& m : c->rs.materials)
    for (Lobe& l : m.lobes) {
      if (k >= nLobes) { c->err = "lobe count differs from the last orc_set_material_lobes"; return -1; }
      l.wrap = wrap[k];
      l.scale = Spec(scale[3 * k], scale[3 * k + 1], scale[3 * k + 2]);
      ++k;
    }
  if (k != nLobes) { c->err = "lobe count differs from the last orc_set_material_lobes"; return -1; }
  return 0;
}

// mirrors drt_set_measured: the tables a MeasuredMaterial holds after loading its file (measured_material.dart:76-205)
int orc_set_measured(orc_ctx* c, uint32_t n, const int32_t* kind, const int32_t* dims, const uint64_t* offsets, const float* data,
                     uint64_t nFloats) {
  c->rs.measured.assign(n, MeasuredTable());
  for (uint32_t i = 0; i < n; ++i) {
    MeasuredTable& t = c->rs.measured[i];
    t.kind = kind[i];
    for (int k = 0; k < 3; ++k) t.dims[k] = dims[3 * i + k];
    if (t.kind != 0 && t.kind != 1) { c->err = "measured table kind must be 0 (regular halfangle) or 1 (irregular isotropic)"; return -1; }
    if (t.dims[0] < 1 || (t.kind == 0 && (t.dims[1] < 1 || t.dims[2] < 1))) { c->err = "measured table dimensions"; return -1; }
    const uint64_t need = t.kind == 0 ? 3ull * t.dims[0] * t.dims[1] * t.dims[2] : 6ull * t.dims[0];
    if (offsets[i] + need > nFloats) { c->err = "measured table beyond the data array"; return -1; }
    t.data.assign(data + offsets[i], data + offsets[i] + need);
  }
  return 0;
}

// mirrors drt_set_textures (include/drt.h)
int orc_set_textures(orc_ctx* c, uint32_t n, const drt_texture* nodes, const float* texels, uint64_t nTexelFloats) {
  TextureSet& ts = c->rs.textures;
  ts.nodes.clear();
  ts.images.clear();
  for (uint32_t i = 0; i < n; ++i) {
    const drt_texture& d = nodes[i];
    TextureNode t;
    t.kind = d.kind; t.spectrum = d.spectrum;
    t.tex1 = d.tex1; t.tex2 = d.tex2; t.amount = d.amount;
    for (int k = 0; k < 3; ++k) t.value[k] = d.value[k];
    for (int k = 0; k < 9; ++k) t.value2[k] = d.value2[k];
    t.mapping = d.mapping;
    t.su = d.su; t.sv = d.sv; t.du = d.du; t.dv = d.dv;
    t.worldToTexture = Transform(d.world_to_texture, d.world_to_texture);  // only m is used (transformPoint)
    t.v1 = Vec(d.v1[0], d.v1[1], d.v1[2]);
    t.v2 = Vec(d.v2[0], d.v2[1], d.v2[2]);
    t.aaMethod = d.aa_method;
    for (int child : {d.tex1, d.tex2, d.amount})
      if (child >= (int)i) { c->err = "a texture node may only reference earlier nodes"; return -1; }
    if (d.kind == 3) {
      const int W = d.image_width, H = d.image_height, ch = d.image_channels;
      if (W < 1 || H < 1 || (W & (W - 1)) || (H & (H - 1)) || (ch != 1 && ch != 3)) { c->err = "image: power-of-two resolution, 1 or 3 channels"; return -1; }
      if (ch != (d.spectrum ? 3 : 1)) { c->err = "image channels must match the texture's type"; return -1; }
      if (d.image_offset + (uint64_t)W * H * ch > nTexelFloats) { c->err = "image beyond the texel array"; return -1; }
      t.image = (int)ts.images.size();
      ts.images.emplace_back();
      ts.images.back().init(W, H, ch, texels + d.image_offset, d.image_wrap, d.image_trilinear != 0, d.max_anisotropy);
    }
    ts.nodes.push_back(t);
  }
  return 0;
}

// mirrors drt_set_material_programs
int orc_set_material_programs(orc_ctx* c, uint32_t n, const drt_material_program* programs) {
  c->rs.programs.clear();
  if (n == 0) return 0;
  if (n != c->rs.materials.size()) { c->err = "one program per material"; return -1; }
  for (uint32_t i = 0; i < n; ++i) {
    MaterialProgram p;
    p.kind = programs[i].kind;
    for (int k = 0; k < 8; ++k) p.tex[k] = programs[i].tex[k];
    p.bump = programs[i].bump;
    p.m1 = programs[i].m1; p.m2 = programs[i].m2;
    c->rs.programs.push_back(p);
  }
  return 0;
}

// Test probes: Texture.evaluate at n DifferentialGeometry records (p 3, u, v, dudx, dvdx, dudy, dvdy, dpdx 3, dpdy 3 = 15 doubles
// each; Points / Vectors are rounded to float32 as their Dart objects would hold them) -> 3 doubles per record (a float
// texture: the value in [0]); and MIPMap.lookup2 of image texture `node` at n (s, t, ds0, dt0, ds1, dt1) records.
int orc_texture_eval(orc_ctx* c, int32_t node, uint32_t n, const double* dgs, double* out) {
  const TextureSet& ts = c->rs.textures;
  if (node < 0 || (size_t)node >= ts.nodes.size()) return -1;
  for (uint32_t i = 0; i < n; ++i) {
    const double* q = dgs + 15 * (size_t)i;
    DG dg;
    dg.p = Vec(q[0], q[1], q[2]);
    dg.u = q[3]; dg.v = q[4]; dg.dudx = q[5]; dg.dvdx = q[6]; dg.dudy = q[7]; dg.dvdy = q[8];
    dg.dpdx = Vec(q[9], q[10], q[11]);
    dg.dpdy = Vec(q[12], q[13], q[14]);
    if (ts.nodes[(size_t)node].spectrum) {
      float v[3];
      ts.evalSpec(node, dg, v);
      out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
    } else {
      out[3 * i] = ts.evalFloat(node, dg);
      out[3 * i + 1] = out[3 * i + 2] = 0.0;
    }
  }
  return 0;
}
int orc_image_level(orc_ctx* c, int32_t node, int32_t level, int32_t* w, int32_t* h, float* out) {
  const TextureSet& ts = c->rs.textures;
  if (node < 0 || (size_t)node >= ts.nodes.size() || ts.nodes[(size_t)node].image < 0) return -1;
  const TexImage& im = ts.images[(size_t)ts.nodes[(size_t)node].image];
  if (level < 0) return im.levels;
  if (level >= im.levels) return -1;
  *w = im.w[level]; *h = im.h[level];
  if (out) std::copy(im.data[level].begin(), im.data[level].end(), out);
  return im.levels;
}

// mirrors drt_set_infinite_light: light `index` (kind 4 in orc_set_lights) gets its transforms and radiance map — level 0 of
// the reference's MIPMap, power-of-two resolution, RGB float32 (a 1x1 white texel when the scene names no map)
int orc_set_infinite_light(orc_ctx* c, uint32_t index, int width, int height, const float* rgb, const float* l2w, const float* w2l) {
  if (index >= c->rs.lights.size() || c->rs.lights[index].kind != 4) { c->err = "not an infinite light"; return -1; }
  if (width < 1 || height < 1 || (width & (width - 1)) || (height & (height - 1))) { c->err = "map resolution must be a power of two"; return -1; }
  Light& l = c->rs.lights[index];
  l.lightToWorld = Transform(l2w, w2l);
  l.worldToLight = Transform(w2l, l2w);
  l.setRadianceMap(width, height, rgb);
  return 0;
}

// mirrors drt_set_spot_params: worldToLight matrices and the two cosines of the spot lights set by orc_set_lights
int orc_set_spot_params(orc_ctx* c, uint32_t n, const float* w2l, const double* cosines) {
  if (n != c->rs.lights.size()) return -1;
  for (uint32_t i = 0; i < n; ++i) {
    Light& l = c->rs.lights[i];
    l.worldToLight = Transform(w2l + 16 * i, w2l + 16 * i);  // only m is used (vector)
    l.cosTotalWidth = cosines[2 * i];
    l.cosFalloffStart = cosines[2 * i + 1];
  }
  return 0;
}

// mirrors drt_set_camera_motion: the camera's end-time CTM (NULL: a static camera) and the two transform times
int orc_set_camera_motion(orc_ctx* c, const float* cameraToWorldEnd, double startTime, double endTime) {
  Camera& cam = c->rs.camera;
  cam.animated = false;
  if (!cameraToWorldEnd) return 0;
  cam.cameraMotion.init(cam.cameraToWorld, startTime, Transform(cameraToWorldEnd, cameraToWorldEnd), endTime);  // only m is read
  cam.animated = cam.cameraMotion.actuallyAnimated;
  return 0;
}

int orc_set_camera(orc_ctx* c, const float* rasterToCamera, const float* cameraToWorld, double lensRadius,
                   double focalDistance, double shutterOpen, double shutterClose) {
  Camera& cam = c->rs.camera;
  cam.rasterToCamera = Transform(rasterToCamera, rasterToCamera);  // only m is used (point/vector)
  cam.cameraToWorld = Transform(cameraToWorld, cameraToWorld);
  cam.animated = false;  // orc_set_camera_motion follows for an animated camera
  cam.lensRadius = lensRadius; cam.focalDistance = focalDistance;
  cam.shutterOpen = shutterOpen; cam.shutterClose = shutterClose;
  return 0;
}

int orc_set_camera_kind(orc_ctx* c, int kind) { c->rs.camera.kind = kind; return 0; }

int orc_set_film(orc_ctx* c, int xres, int yres, const double* crop, double xw, double yw, const float* table) {
  Film& f = c->rs.film;
  f.xres = xres; f.yres = yres;
  for (int i = 0; i < 4; ++i) f.crop[i] = crop ? crop[i] : (i & 1 ? 1.0 : 0.0);
  f.xWidth = xw; f.yWidth = yw;
  for (int i = 0; i < 256; ++i) f.table[i] = table[i];
  f.configure();
  return 0;
}

int orc_set_sampler(orc_ctx* c, int kind, int xs, int ys, int spp, int jitter, int pixelOrder, int tileSize, uint64_t seed,
                    int rngMode) {
  SamplerCfg& s = c->rs.sampler;
  s.kind = kind; s.xs = xs; s.ys = ys; s.spp = spp; s.jitter = jitter; s.pixelOrder = pixelOrder; s.tileSize = tileSize;
  s.seed = seed; s.rngMode = rngMode;
  return 0;
}

// mirrors drt_set_sample_table: the bestcandidate sampler's 4096 x 5 pattern
int orc_set_sample_table(orc_ctx* c, const double* table, uint32_t nEntries) {
  if (nEntries != 4096 || !table) { c->err = "the sample table holds 4096 x 5 values"; return -1; }
  c->rs.sampler.sampleTable.assign(table, table + 5 * (size_t)nEntries);
  return 0;
}

int orc_set_integrator(orc_ctx* c, int kind, int maxDepth, int strategy, int aoSamples, double aoMin, double aoMax) {
  IntegratorCfg& i = c->rs.integ;
  i.kind = kind; i.maxDepth = maxDepth; i.strategy = strategy; i.aoSamples = aoSamples; i.aoMinDist = aoMin; i.aoMaxDist = aoMax;
  return 0;
}

int orc_render(orc_ctx* c, int taskNum, int taskCount, int nthreads) {
  c->rs.geom = &c->scene;
  c->rs.stats = RenderStats();
  c->rs.render(taskNum, taskCount, nthreads);
  return 0;
}

int orc_film_clear(orc_ctx* c) { c->rs.film.configure(); return 0; }
int orc_film_size(const orc_ctx* c, int out[4]) {
  out[0] = c->rs.film.left; out[1] = c->rs.film.top; out[2] = c->rs.film.width; out[3] = c->rs.film.height;
  return 0;
}
int orc_film_read(const orc_ctx* c, float* rgb, float* xyz, float* weight) {
  const Film& f = c->rs.film;
  if (rgb) f.writeImage(rgb);
  if (xyz) std::copy(f.Lxyz.begin(), f.Lxyz.end(), xyz);
  if (weight) std::copy(f.weightSum.begin(), f.weightSum.end(), weight);
  return 0;
}

// Sample values of pixel (x, y): per sample imageX-x, imageY-y, lensU, lensV, time, 1D arrays, 2D arrays.
// Returns floats per sample; *nSamplesOut = samples generated.
int orc_pixel_samples(orc_ctx* c, int x, int y, float* out, int cap, int* nSamplesOut) {
  c->rs.geom = &c->scene;
  std::vector<float> v;
  int per = c->rs.samplesForPixel(x, y, &v);
  int n = per ? (int)v.size() / per : 0;
  if (nSamplesOut) *nSamplesOut = n;
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return per;
}

int orc_render_stats(const orc_ctx* c, uint64_t out[5]) {
  const RenderStats& s = c->rs.stats;
  out[0] = s.cameraSamples; out[1] = s.closestRays; out[2] = s.shadowRays; out[3] = s.nodesVisited; out[4] = s.primsTested;
  return 0;
}

// dart:math Random restatement, for documentation/tests of the serial stream
// VolumeRegion plugins + the volume integrator; same arrays as drt_set_volumes / drt_set_volume_integrator (include/drt.h)
int orc_set_volumes(orc_ctx* c, uint32_t n, const int32_t* kind, const float* sigA, const float* sigS, const float* le, const double* g,
                    const float* p0p1, const float* v2w, const float* w2v, const double* ab, const float* up, const int32_t* dims,
                    const uint64_t* densOff, const double* dens) {
  c->rs.volume.regions.clear();
  for (uint32_t i = 0; i < n; ++i) {
    VolumeRegionCfg v;
    v.kind = kind[i];
    if (v.kind < 0 || v.kind > 2) return -1;
    v.sigA = Spec(sigA[3 * i], sigA[3 * i + 1], sigA[3 * i + 2]);
    v.sigS = Spec(sigS[3 * i], sigS[3 * i + 1], sigS[3 * i + 2]);
    v.le = Spec(le[3 * i], le[3 * i + 1], le[3 * i + 2]);
    v.g = g[i];
    v.p0 = Vec(p0p1[6 * i], p0p1[6 * i + 1], p0p1[6 * i + 2]);
    v.p1 = Vec(p0p1[6 * i + 3], p0p1[6 * i + 4], p0p1[6 * i + 5]);
    v.worldToVolume = Transform(w2v + 16 * i, v2w + 16 * i);
    if (v.kind == 1) {
      v.a = ab[2 * i];
      v.b = ab[2 * i + 1];
      v.upDir = Normalize(Vec(up[3 * i], up[3 * i + 1], up[3 * i + 2]));  // exponential_density_region.dart:29
    }
    if (v.kind == 2) {
      v.nx = dims[3 * i]; v.ny = dims[3 * i + 1]; v.nz = dims[3 * i + 2];
      if (v.nx < 1 || v.ny < 1 || v.nz < 1 || densOff[i + 1] - densOff[i] != (uint64_t)v.nx * v.ny * v.nz) return -1;
      v.density.assign(dens + densOff[i], dens + densOff[i + 1]);
    }
    c->rs.volume.regions.push_back(v);
  }
  return 0;
}
int orc_set_volume_integrator(orc_ctx* c, int kind, double stepSize) {
  if (kind < 0 || kind > 1) return -1;
  c->rs.volume.integrator = kind;
  c->rs.volume.stepSize = stepSize;
  return 0;
}

// test probes: BSDF.f / pdf / sample_f of a material in the canonical shading frame (see RenderScene::bsdfEval)
int orc_bsdf_eval(orc_ctx* c, uint32_t material, uint32_t n, const double* wo, const double* wi, int flags, float* f, double* pdf) {
  if (material >= c->rs.materials.size()) return -1;
  c->rs.bsdfEval(material, n, wo, wi, flags, f, pdf);
  return 0;
}
int orc_bsdf_sample(orc_ctx* c, uint32_t material, uint32_t n, const double* wo, const double* u, int flags, double* wi, float* f,
                    double* pdf, int32_t* sampledType) {
  if (material >= c->rs.materials.size()) return -1;
  c->rs.bsdfSample(material, n, wo, u, flags, wi, f, pdf, sampledType);
  return 0;
}

int orc_dart_random(int64_t seed, int n, double* floats, uint32_t* uints) {
  DartRandom r(seed);
  for (int i = 0; i < n; ++i) {
    if (floats) floats[i] = r.randomFloat();
    if (uints) uints[i] = r.randomUint();
  }
  return 0;
}

}  // extern "C"
