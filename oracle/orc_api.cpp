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

#include "ref_scene.h"

using namespace orc;

struct orc_ctx {
  Scene scene;
  Counters counters;
  std::string err;
  double buildSeconds = 0;
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

int orc_set_build_order(orc_ctx* c, const uint32_t* ids, uint32_t n) {
  if (!ids) { c->scene.buildOrder.clear(); return 0; }
  c->scene.buildOrder.assign(ids, ids + n);
  return 0;
}

int orc_build_bvh(orc_ctx* c, int split, int maxPrims) {
  Scene& s = c->scene;
  if (!s.buildOrder.empty() && s.buildOrder.size() != s.nprims()) { c->err = "build order size mismatch"; return -1; }
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

}  // extern "C"
