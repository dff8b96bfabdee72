// C ABI of libdartray_gpu.so (include/drt.h).  Host orchestration only: scene staging, BVH build,
// device residency, launches.  There is deliberately no CPU execution path for any query.
#include <cuda_runtime.h>

#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "drt_ctx.h"

#include <thread>

namespace {
thread_local std::string g_createError;
}

// transform.dart:110-129 on a float32 matrix with float64 accumulation (sphere world bounds).
static void xformPoint(const float* m, const float p[3], float out[3]) {
  double x = p[0], y = p[1], z = p[2];
  float ox = (float)((double)m[0] * x + (double)m[1] * y + (double)m[2] * z + (double)m[3]);
  float oy = (float)((double)m[4] * x + (double)m[5] * y + (double)m[6] * z + (double)m[7]);
  float oz = (float)((double)m[8] * x + (double)m[9] * y + (double)m[10] * z + (double)m[11]);
  double w = (double)m[12] * x + (double)m[13] * y + (double)m[14] * z + (double)m[15];
  if (w != 1.0) { ox = (float)((double)ox / w); oy = (float)((double)oy / w); oz = (float)((double)oz / w); }
  out[0] = ox; out[1] = oy; out[2] = oz;
}

// Splits [0, n) over the host threads (scene set-up loops over millions of primitives).
template <class F>
static void parallelFor(size_t n, F&& f) {
  const unsigned nt = n < (1u << 16) ? 1u : std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  if (nt == 1) { f((size_t)0, n); return; }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t] { f(n * t / nt, n * (t + 1) / nt); });
  for (auto& x : th) x.join();
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

extern "C" {

int drt_version(void) { return DRT_VERSION; }

drt_ctx* drt_create(int device_id) {
  if (device_id == DRT_DEVICE_NONE) {  // host-only staging context: scene + BVH build/export, no queries
    drt_ctx* c = new drt_ctx();
    c->device = DRT_DEVICE_NONE;
    return c;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_createError = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0") +
                    "); libdartray_gpu has no CPU fallback";
    return nullptr;
  }
  if (device_id < 0 || device_id >= count) { g_createError = "device id out of range"; return nullptr; }
  if ((e = cudaSetDevice(device_id)) != cudaSuccess) { g_createError = cudaGetErrorString(e); return nullptr; }
  drt_ctx* c = new drt_ctx();
  c->device = device_id;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
      (e = c->dCounters.ensure(1)) != cudaSuccess || (e = c->dNextRay.ensure(1 + drt_ctx::kPipe + drt_ctx::kRing)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, device_id)) != cudaSuccess) {
    g_createError = cudaGetErrorString(e);
    delete c;
    return nullptr;
  }
  for (int i = 0; i < drt_ctx::kPipe && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking);
  for (int i = 0; i < 2 * drt_ctx::kMaxChunks && e == cudaSuccess; ++i) e = cudaEventCreate(&c->chunkEv[i]);
  for (int i = 0; i < drt_ctx::kRing && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ringEv[i], cudaEventDisableTiming);
  c->fastV1 = std::getenv("DRT_TRACE_V1") != nullptr;
  if (const char* q = std::getenv("DRT_Q_MIN_PRIMS")) c->qMinPrims = (uint32_t)std::strtoul(q, nullptr, 10);
  c->smallOff = std::getenv("DRT_NO_SMALL") != nullptr;
  if (e != cudaSuccess) {
    g_createError = cudaGetErrorString(e);
    drt_destroy(c);
    return nullptr;
  }
  return c;
}

drt_ctx* drt_create_multi(const int* device_ids, int n_devices) {
  if (!device_ids || n_devices < 1) { g_createError = "drt_create_multi needs at least one device id"; return nullptr; }
  for (int i = 0; i < n_devices; ++i)
    for (int j = 0; j < i; ++j)
      if (device_ids[i] == device_ids[j] && !std::getenv("DRT_ALLOW_DUPLICATE_DEVICES")) {
        // (the env knob lets a one-GPU box exercise the whole multi-device path: two contexts of one device)
        g_createError = "drt_create_multi: a device id is listed twice";
        return nullptr;
      }
  drt_ctx* c = drt_create(device_ids[0]);
  if (!c) return nullptr;
  for (int i = 1; i < n_devices; ++i) {
    drt_ctx* p = drt_create(device_ids[i]);
    if (!p) { drt_destroy(c); return nullptr; }
    p->owner = c;
    c->peers.push_back(p);
  }
  return c;
}

int drt_device_count(const drt_ctx* c) { return c ? 1 + (int)c->peers.size() : 0; }

void drt_destroy(drt_ctx* c) {
  if (!c) return;
  for (drt_ctx* p : c->peers) drt_destroy(p);
  c->peers.clear();
  if (c->device == DRT_DEVICE_NONE) { drtRenderStateDestroy(c); delete c; return; }
  cudaSetDevice(c->device);
  drtRenderStateDestroy(c);
  c->dNodes.release(); c->dWide.release(); c->dWideQ.release(); c->dPrims.release(); c->dSpheres.release(); c->dCounters.release(); c->dNextRay.release();
  c->dRayO.release(); c->dRayD.release(); c->dHits.release(); c->dOcc.release();
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  for (int i = 0; i < drt_ctx::kPipe; ++i) if (c->pipe[i]) cudaStreamDestroy(c->pipe[i]);
  for (int i = 0; i < 2 * drt_ctx::kMaxChunks; ++i) if (c->chunkEv[i]) cudaEventDestroy(c->chunkEv[i]);
  for (int i = 0; i < drt_ctx::kRing; ++i) if (c->ringEv[i]) cudaEventDestroy(c->ringEv[i]);
  delete c;
}

const char* drt_last_error(const drt_ctx* c) { return c ? c->err.c_str() : g_createError.c_str(); }

// The float32 wide nodes of trace_fast.cu, uploaded on demand (scenes that run the quantised kernel never need them).
static int uploadWideV1(drt_ctx* c) {
  const BuiltBvh& B = c->hostBvh();
  CK(c, cudaSetDevice(c->device));
  CK(c, c->dWide.ensure(std::max<size_t>(1, B.wide.size())));
  if (!B.wide.empty())
    CK(c, cudaMemcpy(c->dWide.p, B.wide.data(), B.wide.size() * sizeof(GNode4), cudaMemcpyHostToDevice));
  c->wideUploaded = true;
  c->ts.wide = c->dWide.p;
  return DRT_OK;
}

int drt_set_triangles(drt_ctx* c, const float* P, uint32_t nverts, const uint32_t* idx, uint32_t ntris,
                      const int32_t* mat, const int32_t* light, const uint8_t* rev) {
  DRT_FORWARD_TO_PEERS(c, drt_set_triangles(p_, P, nverts, idx, ntris, mat, light, rev));
  if (!c) return DRT_E_INVALID;
  if ((ntris && (!P || !idx)) || (ntris && !nverts)) return fail(c, DRT_E_INVALID, "null triangle arrays");
  for (uint64_t i = 0; i < (uint64_t)ntris * 3; ++i)
    if (idx[i] >= nverts) return fail(c, DRT_E_INVALID, "triangle index out of range");  // triangle_mesh.dart:168-174
  c->P.assign(P, P + (size_t)nverts * 3);
  c->idx.assign(idx, idx + (size_t)ntris * 3);
  c->vertN.clear(); c->vertS.clear(); c->vertUV.clear(); c->meshOfTri.clear();
  c->meshO2W.clear(); c->meshW2O.clear(); c->meshFlags.clear();
  c->matOf.assign(ntris, 0);
  c->lightOf.assign(ntris, -1);
  c->revOf.assign(ntris, 0);
  if (mat) c->matOf.assign(mat, mat + ntris);
  if (light) c->lightOf.assign(light, light + ntris);
  if (rev) c->revOf.assign(rev, rev + ntris);
  c->built = false;
  return DRT_OK;
}

int drt_set_mesh_shading(drt_ctx* c, const float* N, const float* S, const float* uv, const uint32_t* meshOfTri, uint32_t nmeshes,
                         const float* o2w, const float* w2o, const uint8_t* flags) {
  DRT_FORWARD_TO_PEERS(c, drt_set_mesh_shading(p_, N, S, uv, meshOfTri, nmeshes, o2w, w2o, flags));
  if (!c) return DRT_E_INVALID;
  c->vertN.clear(); c->vertS.clear(); c->vertUV.clear(); c->meshOfTri.clear();
  c->meshO2W.clear(); c->meshW2O.clear(); c->meshFlags.clear();
  c->buildSerial++;  // the render tables depend on it
  if (nmeshes == 0) return DRT_OK;
  if (!meshOfTri || !o2w || !w2o || !flags) return fail(c, DRT_E_INVALID, "null mesh arrays");
  const size_t nv = c->P.size() / 3, nt = c->ntris();
  for (size_t t = 0; t < nt; ++t)
    if (meshOfTri[t] >= nmeshes) return fail(c, DRT_E_INVALID, "mesh index out of range");
  if (N) c->vertN.assign(N, N + 3 * nv);
  if (S) c->vertS.assign(S, S + 3 * nv);
  if (uv) c->vertUV.assign(uv, uv + 2 * nv);
  c->meshOfTri.assign(meshOfTri, meshOfTri + nt);
  c->meshO2W.assign(o2w, o2w + 16 * (size_t)nmeshes);
  c->meshW2O.assign(w2o, w2o + 16 * (size_t)nmeshes);
  c->meshFlags.resize(nmeshes);
  for (uint32_t m = 0; m < nmeshes; ++m)
    c->meshFlags[m] = (uint8_t)(((flags[m] & 1) && N ? 1 : 0) | ((flags[m] & 2) && S ? 2 : 0) | ((flags[m] & 4) && uv ? 4 : 0));
  return DRT_OK;
}

int drt_set_spheres(drt_ctx* c, uint32_t n, const float* o2w, const float* w2o, const double* prm, const int32_t* mat,
                    const int32_t* light, const uint8_t* rev) {
  DRT_FORWARD_TO_PEERS(c, drt_set_spheres(p_, n, o2w, w2o, prm, mat, light, rev));
  if (!c) return DRT_E_INVALID;
  if (n && (!o2w || !w2o || !prm)) return fail(c, DRT_E_INVALID, "null sphere arrays");
  c->spheres.resize(n);
  c->sphMat.assign(n, 0);
  c->sphLight.assign(n, -1);
  c->sphRev.assign(n, 0);
  for (uint32_t i = 0; i < n; ++i) {
    HostSphere& s = c->spheres[i];
    std::memcpy(s.o2w, o2w + 16 * i, 64);
    std::memcpy(s.w2o, w2o + 16 * i, 64);
    s.radius = prm[4 * i]; s.zmin = prm[4 * i + 1]; s.zmax = prm[4 * i + 2]; s.phiMaxDeg = prm[4 * i + 3];
    s.shape = 0; s.height = 0.0; s.innerRadius = 0.0;
    if (mat) c->sphMat[i] = mat[i];
    if (light) c->sphLight[i] = light[i];
    if (rev) c->sphRev[i] = rev[i];
  }
  c->built = false;
  return DRT_OK;
}

int drt_set_disks(drt_ctx* c, uint32_t n, const float* o2w, const float* w2o, const double* prm, const int32_t* mat,
                  const int32_t* light, const uint8_t* rev) {
  DRT_FORWARD_TO_PEERS(c, drt_set_disks(p_, n, o2w, w2o, prm, mat, light, rev));
  if (!c) return DRT_E_INVALID;
  if (n && (!o2w || !w2o || !prm)) return fail(c, DRT_E_INVALID, "null disk arrays");
  for (uint32_t i = 0; i < n; ++i) {  // appended to the quadric range, after the spheres
    HostSphere s;
    std::memcpy(s.o2w, o2w + 16 * i, 64);
    std::memcpy(s.w2o, w2o + 16 * i, 64);
    s.shape = 1;
    s.height = prm[4 * i]; s.radius = prm[4 * i + 1]; s.innerRadius = prm[4 * i + 2]; s.phiMaxDeg = prm[4 * i + 3];
    s.zmin = s.zmax = s.height;
    c->spheres.push_back(s);
    c->sphMat.push_back(mat ? mat[i] : 0);
    c->sphLight.push_back(light ? light[i] : -1);
    c->sphRev.push_back(rev ? rev[i] : 0);
  }
  c->built = false;
  return DRT_OK;
}

int drt_set_quadrics(drt_ctx* c, int kind, uint32_t n, const float* o2w, const float* w2o, const double* prm, const int32_t* mat,
                     const int32_t* light, const uint8_t* rev) {
  DRT_FORWARD_TO_PEERS(c, drt_set_quadrics(p_, kind, n, o2w, w2o, prm, mat, light, rev));
  if (!c) return DRT_E_INVALID;
  if (kind < DRT_QUADRIC_CYLINDER || kind > DRT_QUADRIC_HYPERBOLOID) return fail(c, DRT_E_INVALID, "unknown quadric kind");
  if (n && (!o2w || !w2o || !prm)) return fail(c, DRT_E_INVALID, "null quadric arrays");
  for (uint32_t i = 0; i < n; ++i) {  // appended to the quadric range in call order
    HostSphere s;
    std::memcpy(s.o2w, o2w + 16 * i, 64);
    std::memcpy(s.w2o, w2o + 16 * i, 64);
    s.shape = kind;
    std::memcpy(s.prm, prm + 8 * i, sizeof(s.prm));
    c->spheres.push_back(s);
    c->sphMat.push_back(mat ? mat[i] : 0);
    c->sphLight.push_back(light ? light[i] : -1);
    c->sphRev.push_back(rev ? rev[i] : 0);
  }
  c->built = false;
  return DRT_OK;
}

int drt_set_instances(drt_ctx* c, uint32_t n_objects, const uint32_t* object_offsets, const uint32_t* object_prims,
                      const int32_t* object_split, const int32_t* object_max_node_prims, uint32_t n_instances,
                      const uint32_t* instance_object, const float* start_m, const float* start_minv, const float* end_m,
                      const float* end_minv, const double* times) {
  DRT_FORWARD_TO_PEERS(c, drt_set_instances(p_, n_objects, object_offsets, object_prims, object_split, object_max_node_prims, n_instances,
                                            instance_object, start_m, start_minv, end_m, end_minv, times));
  if (!c) return DRT_E_INVALID;
  if (n_objects && (!object_offsets || !object_prims)) return fail(c, DRT_E_INVALID, "drt_set_instances: null object arrays");
  if (n_instances && (!instance_object || !start_m || !start_minv || !end_m || !end_minv))
    return fail(c, DRT_E_INVALID, "drt_set_instances: null instance arrays");
  if (n_instances && !n_objects) return fail(c, DRT_E_INVALID, "drt_set_instances: instances without objects");
  std::vector<drt_ctx::HostObject> objs(n_objects);
  for (uint32_t i = 0; i < n_objects; ++i) {
    if (object_offsets[i + 1] <= object_offsets[i]) return fail(c, DRT_E_INVALID, "an object holds at least one primitive (dartray.dart:514-516)");
    objs[i].order.assign(object_prims + object_offsets[i], object_prims + object_offsets[i + 1]);
    objs[i].split = object_split ? object_split[i] : 2;
    objs[i].maxPrims = object_max_node_prims ? object_max_node_prims[i] : 1;  // BVHAccel's constructor defaults (bvh_accel.dart:41)
    if (objs[i].split < 0 || objs[i].split > 2 || objs[i].maxPrims < 1) return fail(c, DRT_E_INVALID, "object accelerator parameters out of range");
  }
  std::vector<GInstance> insts(n_instances);
  for (uint32_t i = 0; i < n_instances; ++i) {
    if (instance_object[i] >= n_objects) return fail(c, DRT_E_INVALID, "an instance names an object that was not defined");
    std::memset(&insts[i], 0, sizeof(GInstance));
    animInit(&insts[i], start_m + 16 * (size_t)i, start_minv + 16 * (size_t)i, end_m + 16 * (size_t)i, end_minv + 16 * (size_t)i,
             times ? times[2 * i] : 0.0, times ? times[2 * i + 1] : 1.0);
    insts[i].object = (int32_t)instance_object[i];
  }
  c->objects.swap(objs);
  c->instances.swap(insts);
  c->built = false;
  return DRT_OK;
}

int drt_set_ray_times(drt_ctx* c, const double* times, uint64_t n) {
  if (!c) return DRT_E_INVALID;
  if (times) c->rayTimes.assign(times, times + n);
  else c->rayTimes.clear();
  c->rayTimesDirty = true;
  return DRT_OK;
}

int drt_set_build_order(drt_ctx* c, const uint32_t* ids, uint32_t n) {
  DRT_FORWARD_TO_PEERS(c, drt_set_build_order(p_, ids, n));
  if (!c) return DRT_E_INVALID;
  if (!ids) c->order.clear();
  else c->order.assign(ids, ids + n);
  c->built = false;
  return DRT_OK;
}

// Constructors of the remaining quadrics: cylinder.dart:24-31, cone.dart:23-27, paraboloid.dart:23-29,
// hyperboloid.dart:23-49 (Points are float32, the expressions f64).  The object bound of each is
// (-radius, -radius, zmin) .. (radius, radius, zmax) with the record's radius / zmin / zmax (cone: 0 .. height;
// hyperboloid: rmax).
// Returns false for a hyperboloid whose implicit coefficients never become finite (hyperboloid.dart:40-48 loops until they do:
// p1 == p2, non-finite points, ...): an ABI entry point must not hang where the reference's constructor would.
static bool quadricSetup(const HostSphere& s, GSphere* gp) {
  GSphere& g = *gp;
  const double* prm = s.prm;
  double pm = 360.0;
  g.thetaMin = g.thetaMax = 0.0;
  g.height = 0.0;
  g.innerRadius = 0.0;
  if (s.shape == DRT_QUADRIC_CYLINDER || s.shape == DRT_QUADRIC_PARABOLOID) {
    g.radius = prm[0];
    g.zmin = std::fmin(prm[1], prm[2]);
    g.zmax = std::fmax(prm[1], prm[2]);
    pm = prm[3];
  } else if (s.shape == DRT_QUADRIC_CONE) {
    g.height = prm[0];
    g.radius = prm[1];
    g.zmin = 0.0;
    g.zmax = g.height;
    pm = prm[2];
  } else {
    float p1[3] = {(float)prm[0], (float)prm[1], (float)prm[2]}, p2[3] = {(float)prm[3], (float)prm[4], (float)prm[5]};
    pm = prm[6];
    double radius1 = std::sqrt((double)p1[0] * p1[0] + (double)p1[1] * p1[1]);
    double radius2 = std::sqrt((double)p2[0] * p2[0] + (double)p2[1] * p2[1]);
    g.radius = std::fmax(radius1, radius2);
    g.zmin = std::fmin((double)p1[2], (double)p2[2]);
    g.zmax = std::fmax((double)p1[2], (double)p2[2]);
    if (p2[2] == 0.0f)
      for (int k = 0; k < 3; ++k) std::swap(p1[k], p2[k]);
    float pp[3] = {p1[0], p1[1], p1[2]};
    double a, cc;
    int tries = 0;
    do {
      if (++tries > 4096) return false;
      for (int k = 0; k < 3; ++k) {
        float dk = (float)((double)p2[k] - (double)p1[k]);  // Vector p2 - p1
        float d2 = (float)((double)dk * 2.0);               // * 2.0
        pp[k] = (float)((double)pp[k] + (double)d2);        // Point + Vector
      }
      double xy1 = (double)pp[0] * pp[0] + (double)pp[1] * pp[1];
      double xy2 = (double)p2[0] * p2[0] + (double)p2[1] * p2[1];
      a = (1.0 / xy1 - ((double)pp[2] * pp[2]) / (xy1 * p2[2] * p2[2])) / (1.0 - (xy2 * pp[2] * pp[2]) / (xy1 * p2[2] * p2[2]));
      cc = (a * xy2 - 1.0) / ((double)p2[2] * p2[2]);
    } while (std::isinf(a) || std::isnan(a));
    for (int k = 0; k < 3; ++k) { g.hp1[k] = p1[k]; g.hp2[k] = p2[k]; }
    g.ha = a;
    g.hc = cc;
  }
  g.phiMax = (3.141592653589793 / 180.0) * clampd(pm, 0.0, 360.0);
  return true;
}

// Leaf lists of a small scene (GSmallScene, gpu_types.h): the leaves in storage order with their boxes, and for each dirIsNeg octant
// the order in which the reference's walk reaches them (bvh_accel.dart:147-153: the second child first when the ray's direction is
// negative along the node's split axis).  Returns false when the scene is not small.
static bool buildSmallScene(const BuiltBvh& B, const std::vector<GNode>& nodes, const std::vector<GPrim>& prims, const std::vector<GSphere>& gs,
                            GSmallScene* out) {
  if (B.nLeaves == 0 || B.nLeaves > DRT_SMALL_MAX_LEAVES) return false;
  std::memset(out, 0, sizeof(*out));
  std::vector<int32_t> walk[8];
  for (int oct = 0; oct < 8; ++oct) {
    std::vector<int32_t> todo{B.rootRef};
    while (!todo.empty()) {
      const int32_t ref = todo.back();
      todo.pop_back();
      if (ref == DRT_REF_EMPTY) continue;
      if (ref < 0) { walk[oct].push_back(ref); continue; }
      if ((size_t)ref >= nodes.size()) return false;
      const GNode& n = nodes[(size_t)ref];
      const bool neg = ((oct >> (n.axis & 3)) & 1) != 0;
      todo.push_back(neg ? n.ref0 : n.ref1);  // the far child waits on the stack
      todo.push_back(neg ? n.ref1 : n.ref0);
    }
    if (walk[oct].size() != B.nLeaves) return false;
  }
  out->nLeaves = (int32_t)B.nLeaves;
  for (uint32_t l = 0; l < B.nLeaves; ++l) {
    const int32_t ref = walk[0][l];
    GSmallLeaf& L = out->leaf[l];
    L.ref = ref;
    const uint32_t off = refLeafOffset(ref);
    uint32_t cnt = refLeafCountField(ref);
    if (off >= prims.size()) return false;
    if (cnt == 15u) cnt = (uint32_t)prims[off].leafCount;
    if (cnt == 0 || off + cnt > prims.size()) return false;
    for (int a = 0; a < 3; ++a) { L.lo[a] = INFINITY; L.hi[a] = -INFINITY; }
    for (uint32_t k = 0; k < cnt; ++k) {  // the box the traversal kernels rebuild from a leaf's records (trace_fast.cu, leaf phase)
      const GPrim& pr = prims[off + k];
      if ((pr.kindSphere & 1) == 0) {
        for (int a = 0; a < 3; ++a) {
          L.lo[a] = std::fmin(L.lo[a], std::fmin(pr.p1[a], std::fmin(pr.p2[a], pr.p3[a])));
          L.hi[a] = std::fmax(L.hi[a], std::fmax(pr.p1[a], std::fmax(pr.p2[a], pr.p3[a])));
        }
      } else {
        const size_t si = (size_t)(pr.kindSphere >> 1);
        if (si >= gs.size()) return false;
        for (int a = 0; a < 3; ++a) {
          L.lo[a] = std::fmin(L.lo[a], gs[si].wmin[a]);
          L.hi[a] = std::fmax(L.hi[a], gs[si].wmax[a]);
        }
      }
    }
  }
  for (int oct = 0; oct < 8; ++oct)
    for (uint32_t k = 0; k < B.nLeaves; ++k) {
      uint32_t l = 0;
      while (l < B.nLeaves && walk[0][l] != walk[oct][k]) ++l;
      if (l == B.nLeaves) return false;
      out->order[oct][k] = (uint8_t)l;
      out->position[oct][l] = (uint8_t)k;
    }
  return true;
}

// Device half of drt_build_bvh: the built tree's arrays go to c's device.  B / prims / gs may belong to another context (the
// owner of a multi-device context builds once on the host and every device uploads the same arrays).
// `src`: a context of ANOTHER device that already holds this build: the arrays then come over NVLink from its memory
// (cudaMemcpyPeer) instead of over PCIe from pageable host memory — soup_10m on 8 GPUs: 1.2 s of host uploads -> one upload.
static int uploadBuilt(drt_ctx* c, const BuiltBvh& B, const std::vector<GNode>& nodes, const std::vector<GPrim>& prims,
                       const std::vector<GSphere>& gs, const std::vector<GObject>& gobjs, const std::vector<GInstance>& ginsts, uint32_t np,
                       const drt_ctx* src = nullptr) {
  CK(c, cudaSetDevice(c->device));
  auto put = [&](void* dst, const void* host, const void* peer, size_t bytes) -> cudaError_t {
    if (bytes == 0) return cudaSuccess;
    if (src && peer && src->device != c->device) return cudaMemcpyPeer(dst, c->device, peer, src->device, bytes);
    return cudaMemcpy(dst, host, bytes, cudaMemcpyHostToDevice);
  };
  CK(c, c->dNodes.ensure(std::max<size_t>(1, nodes.size())));
  c->wideQOk = B.wideQOk && B.wideQ.size() == B.wide.size();
  c->wideUploaded = false;
  if (!c->useQ()) {  // the float32 nodes go to the device only when their kernel is the one that runs (or is asked for later)
    int rc = uploadWideV1(c);
    if (rc != DRT_OK) return rc;
  }
  if (c->wideQOk) {
    CK(c, c->dWideQ.ensure(std::max<size_t>(1, B.wideQ.size())));
    CK(c, put(c->dWideQ.p, B.wideQ.data(), src ? src->dWideQ.p : nullptr, B.wideQ.size() * sizeof(GNode4Q)));
  }
  CK(c, c->dPrims.ensure(prims.size()));
  CK(c, c->dSpheres.ensure(std::max<size_t>(1, gs.size())));
  CK(c, put(c->dNodes.p, nodes.data(), src ? src->dNodes.p : nullptr, nodes.size() * sizeof(GNode)));
  c->ts.instances = nullptr; c->ts.objects = nullptr; c->ts.nInstances = 0;
  if (!ginsts.empty()) {
    CK(c, c->dInstances.ensure(ginsts.size()));
    CK(c, c->dObjects.ensure(gobjs.size()));
    CK(c, cudaMemcpy(c->dInstances.p, ginsts.data(), ginsts.size() * sizeof(GInstance), cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->dObjects.p, gobjs.data(), gobjs.size() * sizeof(GObject), cudaMemcpyHostToDevice));
  }
  CK(c, put(c->dPrims.p, prims.data(), src ? src->dPrims.p : nullptr, prims.size() * sizeof(GPrim)));
  CK(c, put(c->dSpheres.p, gs.data(), src ? src->dSpheres.p : nullptr, gs.size() * sizeof(GSphere)));
  c->smallOk = false;
  if (ginsts.empty()) {
    GSmallScene small;
    if (buildSmallScene(B, nodes, prims, gs, &small)) {
      CK(c, c->dSmall.ensure(1));
      CK(c, cudaMemcpy(c->dSmall.p, &small, sizeof(small), cudaMemcpyHostToDevice));
      c->smallOk = true;
    }
  }
  c->ts.nodes = c->dNodes.p;
  c->ts.wide = c->dWide.p;
  c->ts.wideQ = c->useQ() ? c->dWideQ.p : nullptr;
  c->ts.small = c->useSmall() ? c->dSmall.p : nullptr;
  c->ts.wideRootRef = B.wideRootRef;
  c->ts.prims = c->dPrims.p;
  c->ts.spheres = c->dSpheres.p;
  std::memcpy(c->ts.rootMin, B.rootMin, 12);
  std::memcpy(c->ts.rootMax, B.rootMax, 12);
  c->ts.rootRef = B.rootRef;
  c->ts.empty = 0;
  c->ts.quadMode = 0;
  for (const HostSphere& hs : c->spheres) c->ts.quadMode = std::max(c->ts.quadMode, hs.shape >= 2 ? 2 : 1);
  if (!ginsts.empty()) {
    c->ts.instances = c->dInstances.p;
    c->ts.objects = c->dObjects.p;
    c->ts.nInstances = (int32_t)ginsts.size();
    c->ts.quadMode = 2;
  }
  c->info.n_nodes = (uint32_t)B.refNodes.size();
  c->info.n_prims = np;
  c->info.n_leaves = B.nLeaves;
  c->info.max_leaf_prims = B.maxLeafPrims;
  c->info.max_depth = B.maxDepth;
  c->info.device_bytes = (c->wideQOk ? B.wideQ.size() * sizeof(GNode4Q) : 0) + (c->wideUploaded ? B.wide.size() * sizeof(GNode4) : 0) +
                         prims.size() * sizeof(GPrim) + gs.size() * sizeof(GSphere);
  c->built = true;
  return DRT_OK;
}


int drt_build_bvh(drt_ctx* c, int split, int maxPrims) {
  if (!c) return DRT_E_INVALID;
  if (split < 0 || split > 2) return fail(c, DRT_E_INVALID, "split method must be 0 (middle), 1 (equal) or 2 (sah)");
  if (maxPrims < 1) return fail(c, DRT_E_INVALID, "maxnodeprims must be >= 1");
  const bool hostOnly = c->device == DRT_DEVICE_NONE;
  if (!hostOnly) CK(c, cudaSetDevice(c->device));
  auto t0 = std::chrono::steady_clock::now();
  const uint32_t nt = c->ntris(), np = c->nprims();
  c->built = false;
  c->buildSerial++;
  c->ts = TraceScene{};
  c->info = drt_bvh_info{};
  const uint32_t ninst = (uint32_t)c->instances.size();
  if (np == 0) {  // bvh_accel.dart:50-53: nodes == null, every query misses
    c->ts.empty = 1;
    c->built = true;
    for (drt_ctx* p_ : c->peers) { p_->ts = TraceScene{}; p_->ts.empty = 1; p_->built = true; p_->buildSerial++; }
    return DRT_OK;
  }
  std::vector<uint32_t> order = c->order;
  if (order.empty()) {
    order.resize(np);
    for (uint32_t i = 0; i < np; ++i) order[i] = i;
    if (ninst) return fail(c, DRT_E_INVALID, "a scene with instances needs the top-level build order (drt_set_build_order)");
  } else if (ninst == 0) {
    if (order.size() != np) return fail(c, DRT_E_INVALID, "build order length != primitive count");
    std::vector<uint8_t> seen(np, 0);
    for (uint32_t id : order) {
      if (id >= np || seen[id]) return fail(c, DRT_E_INVALID, "build order is not a permutation of the primitive ids");
      seen[id] = 1;
    }
  } else {
    // every geometric primitive belongs to the top level or to exactly one object; every instance is a top-level primitive
    std::vector<uint8_t> seen(np + ninst, 0);
    for (const drt_ctx::HostObject& ob : c->objects)
      for (uint32_t id : ob.order) {
        if (id >= np || seen[id]) return fail(c, DRT_E_INVALID, "object primitive ids must be distinct geometric primitives");
        seen[id] = 1;
      }
    for (uint32_t id : order) {
      if (id >= np + ninst || seen[id]) return fail(c, DRT_E_INVALID, "build order: ids of top-level primitives and instances, each once, none owned by an object");
      seen[id] = 1;
    }
    for (uint32_t id = 0; id < np + ninst; ++id)
      if (!seen[id]) return fail(c, DRT_E_INVALID, "a primitive or instance is neither in an object nor in the build order");
  }
  // world bounds: triangle.dart:39-42, sphere.dart:34-37 + shape.dart:38-40
  std::vector<PrimBounds> bounds(np + ninst);
  parallelFor(nt, [&](size_t t0, size_t t1) {
    for (size_t t = t0; t < t1; ++t) {
      PrimBounds& b = bounds[t];
      for (int a = 0; a < 3; ++a) {
        float v0 = c->P[3 * (size_t)c->idx[3 * t] + a], v1 = c->P[3 * (size_t)c->idx[3 * t + 1] + a],
              v2 = c->P[3 * (size_t)c->idx[3 * t + 2] + a];
        b.bmin[a] = std::fmin(std::fmin(v0, v1), v2);
        b.bmax[a] = std::fmax(std::fmax(v0, v1), v2);
      }
    }
  });
  std::vector<GSphere> gs(c->spheres.size() + ninst);
  for (size_t i = 0; i < c->spheres.size(); ++i) {
    const HostSphere& s = c->spheres[i];
    GSphere& g = gs[i];
    std::memcpy(g.w2o, s.w2o, 48);
    std::memcpy(g.w2oRow3, s.w2o + 12, 16);
    std::memcpy(g.o2w, s.o2w, 48);
    std::memcpy(g.o2wRow3, s.o2w + 12, 16);
    g.radius = s.radius;  // sphere.dart:24-32
    g.shape = s.shape;
    g.instance = -1;
    g.height = s.height;
    g.innerRadius = s.innerRadius;
    g.phiMax = (3.141592653589793 / 180.0) * clampd(s.phiMaxDeg, 0.0, 360.0);
    g.ha = g.hc = 0.0;
    for (int k = 0; k < 3; ++k) g.hp1[k] = g.hp2[k] = 0.f;
    if (s.shape == 1) {  // disk.dart:24-35: object bound (-r, -r, h) .. (r, r, h)
      g.zmin = g.zmax = s.height;
      g.thetaMin = g.thetaMax = 0.0;
    } else if (s.shape >= 2) {
      if (!quadricSetup(s, &g)) return fail(c, DRT_E_INVALID, "hyperboloid: degenerate points (the implicit form has no finite coefficients)");
    } else {
      g.zmin = clampd(std::fmin(s.zmin, s.zmax), -s.radius, s.radius);
      g.zmax = clampd(std::fmax(s.zmin, s.zmax), -s.radius, s.radius);
      g.thetaMin = std::acos(clampd(g.zmin / s.radius, -1.0, 1.0));
      g.thetaMax = std::acos(clampd(g.zmax / s.radius, -1.0, 1.0));
    }
    float lo[3] = {(float)-g.radius, (float)-g.radius, (float)g.zmin}, hi[3] = {(float)g.radius, (float)g.radius, (float)g.zmax};
    PrimBounds& b = bounds[nt + i];
    for (int k = 0; k < 8; ++k) {  // transform.dart:163-178
      float p[3] = {(k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2]}, q[3];
      xformPoint(s.o2w, p, q);
      for (int a = 0; a < 3; ++a) {
        b.bmin[a] = k == 0 ? q[a] : std::fmin(b.bmin[a], q[a]);
        b.bmax[a] = k == 0 ? q[a] : std::fmax(b.bmax[a], q[a]);
      }
    }
    std::memcpy(g.wmin, b.bmin, 12);
    std::memcpy(g.wmax, b.bmax, 12);
  }
  std::string err;
  // the objects TransformedPrimitives wrap: their own accelerators first (an instance's world bound needs its object's),
  // binary nodes and leaf records collected behind the top level's (references rebased below)
  std::vector<BuiltBvh> objBvh(c->objects.size());
  std::vector<GObject> gobjs(c->objects.size());
  for (size_t k = 0; k < c->objects.size(); ++k) {
    const drt_ctx::HostObject& ob = c->objects[k];
    if (!buildBvh(bounds, ob.order, ob.split, ob.maxPrims, &objBvh[k], &err)) return fail(c, DRT_E_INVALID, err.c_str());
    GObject& g = gobjs[k];
    std::memcpy(g.rootMin, objBvh[k].rootMin, 12);
    std::memcpy(g.rootMax, objBvh[k].rootMax, 12);
    g.single = ob.order.size() == 1 ? 1 : 0;
    if (g.single) {  // primitive.worldBound() of the GeometricPrimitive itself
      std::memcpy(g.rootMin, bounds[ob.order[0]].bmin, 12);
      std::memcpy(g.rootMax, bounds[ob.order[0]].bmax, 12);
    }
    g.rootRef = objBvh[k].rootRef;
  }
  for (uint32_t i = 0; i < ninst; ++i) {  // TransformedPrimitive.worldBound (transformed_primitive.dart:76-78)
    const GInstance& in = c->instances[i];
    const GObject& ob = gobjs[(size_t)in.object];
    PrimBounds& b = bounds[np + i];
    animMotionBounds(in, ob.rootMin, ob.rootMax, b.bmin, b.bmax);
    GSphere& g = gs[c->spheres.size() + i];
    std::memset(&g, 0, sizeof(g));
    g.shape = 6;
    g.instance = (int32_t)i;
    std::memcpy(g.wmin, b.bmin, 12);
    std::memcpy(g.wmax, b.bmax, 12);
  }
  static const bool phaseTiming = std::getenv("DRT_BUILD_TIMING") != nullptr;
  auto tPhase = std::chrono::steady_clock::now();
  auto phase = [&](const char* what) {
    if (!phaseTiming) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[drt build] %-28s %.3f s\n", what, std::chrono::duration<double>(now - tPhase).count());
    tPhase = now;
  };
  phase("(api) bounds, quadrics, objects");
  if (!buildBvh(bounds, order, split, maxPrims, &c->bvh, &err)) return fail(c, DRT_E_INVALID, err.c_str());
  phase("(api) buildBvh");
  const BuiltBvh& B = c->bvh;
  // the top level's binary nodes and leaf counts are used in place unless objects append theirs (soup_10m: 650 MB not copied)
  std::vector<GNode> nodesOwned;
  std::vector<uint32_t> recCountsOwned;
  if (!objBvh.empty()) {
    nodesOwned = B.nodes;
    recCountsOwned = B.leafCounts;
  }
  std::vector<GNode>& nodesMut = nodesOwned;
  std::vector<uint32_t>& recCountsMut = recCountsOwned;
  c->recPrimIds = B.leafPrimIds;
  for (size_t k = 0; k < objBvh.size(); ++k) {
    const BuiltBvh& O = objBvh[k];
    const int32_t nodeBase = (int32_t)nodesMut.size();
    const uint32_t recBase = (uint32_t)c->recPrimIds.size();
    auto rebase = [&](int32_t ref) -> int32_t {
      if (ref == DRT_REF_EMPTY) return ref;
      if (ref >= 0) return ref + nodeBase;
      const uint32_t bits = (uint32_t)~ref;
      return (int32_t)~((((bits >> 5) + recBase) << 5) | (bits & 31u));
    };
    for (GNode n : O.nodes) { n.ref0 = rebase(n.ref0); n.ref1 = rebase(n.ref1); nodesMut.push_back(n); }
    gobjs[k].rootRef = rebase(O.rootRef);
    c->recPrimIds.insert(c->recPrimIds.end(), O.leafPrimIds.begin(), O.leafPrimIds.end());
    recCountsMut.insert(recCountsMut.end(), O.leafCounts.begin(), O.leafCounts.end());
  }
  const std::vector<GNode>& nodes = objBvh.empty() ? B.nodes : nodesOwned;
  const std::vector<uint32_t>& recCounts = objBvh.empty() ? B.leafCounts : recCountsOwned;

  phase("(api) node / record copies");
  // leaf records
  std::vector<GPrim> prims(c->recPrimIds.size());
  parallelFor(prims.size(), [&](size_t i0, size_t i1) {
    for (size_t i = i0; i < i1; ++i) {
      uint32_t id = c->recPrimIds[i];
      GPrim& g = prims[i];
      std::memset(&g, 0, sizeof(g));
      g.primId = (int32_t)id;
      g.leafCount = (int32_t)recCounts[i];
      if (id < nt) {
        const float* a = &c->P[3 * (size_t)c->idx[3 * (size_t)id]];
        const float* b = &c->P[3 * (size_t)c->idx[3 * (size_t)id + 1]];
        const float* d = &c->P[3 * (size_t)c->idx[3 * (size_t)id + 2]];
        std::memcpy(g.p1, a, 12);
        std::memcpy(g.p2, b, 12);
        std::memcpy(g.p3, d, 12);
        g.kindSphere = 0;
      } else {  // quadrics, then the instances' stand-in records (GSphere::shape 6)
        g.kindSphere = (int32_t)(((id - nt) << 1) | 1u);
      }
    }
  });
  phase("(api) leaf records");
  if (hostOnly) {
    c->info.n_nodes = (uint32_t)B.refNodes.size();
    c->info.n_prims = np;
    c->info.n_leaves = B.nLeaves;
    c->info.max_leaf_prims = B.maxLeafPrims;
    c->info.max_depth = B.maxDepth;
    c->info.device_bytes = 0;
    c->info.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    c->built = true;
    return DRT_OK;
  }
  int rc = uploadBuilt(c, B, nodes, prims, gs, gobjs, c->instances, np);
  if (rc != DRT_OK) return rc;
  phase("(api) upload, first device");
  // a multi-device context: the same arrays to every other device, one host thread each
  if (!c->peers.empty()) {
    std::vector<int> rcs(c->peers.size(), DRT_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < c->peers.size(); ++i)
      th.emplace_back([&, i] {
        drt_ctx* p_ = c->peers[i];
        p_->built = false;
        p_->buildSerial++;
        p_->ts = TraceScene{};
        p_->info = drt_bvh_info{};
        p_->recPrimIds = c->recPrimIds;
        rcs[i] = uploadBuilt(p_, B, nodes, prims, gs, gobjs, c->instances, np, c);
      });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < rcs.size(); ++i)
      if (rcs[i] != DRT_OK) { c->err = "device " + std::to_string(c->peers[i]->device) + ": " + c->peers[i]->err; return rcs[i]; }
  }
  c->info.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (drt_ctx* p_ : c->peers) p_->info.build_seconds = c->info.build_seconds;
  return DRT_OK;
}

int drt_bvh_info_get(const drt_ctx* c, drt_bvh_info* out) {
  if (!c || !out) return DRT_E_INVALID;
  if (!c->built) return DRT_E_STATE;
  *out = c->info;
  return DRT_OK;
}

int drt_bvh_export(const drt_ctx* c, float* bounds, int32_t* offset, int32_t* nprims, int32_t* axis, uint32_t* ordered) {
  if (!c) return DRT_E_INVALID;
  if (!c->built) return DRT_E_STATE;
  const BuiltBvh& B = c->hostBvh();
  for (size_t i = 0; i < B.refNodes.size(); ++i) {
    const RefNode& n = B.refNodes[i];
    if (bounds) { std::memcpy(bounds + 6 * i, n.bmin, 12); std::memcpy(bounds + 6 * i + 3, n.bmax, 12); }
    if (offset) offset[i] = n.offset;
    if (nprims) nprims[i] = n.nPrimitives;
    if (axis) axis[i] = n.axis;
  }
  if (ordered) std::memcpy(ordered, B.refOrdered.data(), B.refOrdered.size() * sizeof(uint32_t));
  return DRT_OK;
}


// counterSlot < 0: a launch on a caller's stream (drt_trace_*_device) — takes the next slot of the counter ring.
static int traceDevice(drt_ctx* c, bool any, const void* o, const void* d, uint64_t n, void* out, cudaStream_t st,
                       int counterSlot = -1, uint64_t firstRay = 0) {
  if (c->device == DRT_DEVICE_NONE) return fail(c, DRT_E_NODEVICE, kNoDevice);
  if (!c->built) return fail(c, DRT_E_STATE, "drt_build_bvh must be called before tracing");
  if (n && (!o || !d || !out)) return fail(c, DRT_E_INVALID, "null ray or output buffer");
  CK(c, cudaSetDevice(c->device));
  int ring = -1;
  if (counterSlot < 0) {
    ring = c->ringNext;
    c->ringNext = (c->ringNext + 1) % drt_ctx::kRing;
    counterSlot = 1 + drt_ctx::kPipe + ring;
    if (c->ringUsed[ring]) CK(c, cudaStreamWaitEvent(st, c->ringEv[ring], 0));  // the slot's previous launch, maybe on another stream
  }
  if (c->counting || c->exactWalk || c->ts.nInstances > 0) {
    // reference-walk kernel: the slab test in f64 for every node; also the counting variant, and the kernel that descends into
    // TransformedPrimitives
    if (c->counting) CK(c, cudaMemsetAsync(c->dCounters.p, 0, sizeof(DeviceCounters), st));
    ExactExtras xx{};
    if (c->ts.nInstances > 0 && !c->rayTimes.empty()) {
      if (firstRay + n > c->rayTimes.size()) return fail(c, DRT_E_INVALID, "drt_set_ray_times gave fewer times than this call has rays");
      if (c->rayTimesDirty) {
        CK(c, c->dRayTimes.ensure(c->rayTimes.size()));
        CK(c, cudaMemcpy(c->dRayTimes.p, c->rayTimes.data(), c->rayTimes.size() * sizeof(double), cudaMemcpyHostToDevice));
        c->rayTimesDirty = false;
      }
      xx.times = c->dRayTimes.p + firstRay;
    }
    CK(c, launchTrace(c->ts, any, c->counting, o, d, n, out, c->dCounters.p, st, nullptr, nullptr, &xx));
  } else {
    CK(c, launchTraceFast(c->ts, any, o, d, n, out, c->dNextRay.p + counterSlot, c->numSMs, st));
  }
  if (ring >= 0) {
    CK(c, cudaEventRecord(c->ringEv[ring], st));
    c->ringUsed[ring] = true;
  }
  if (n) c->launches++;
  return DRT_OK;
}

static int traceHost(drt_ctx* c, bool any, const float* o, const float* d, uint64_t n, void* out) {
  if (!c) return DRT_E_INVALID;
  if (c->device == DRT_DEVICE_NONE) return fail(c, DRT_E_NODEVICE, kNoDevice);
  if (!c->built) return fail(c, DRT_E_STATE, "drt_build_bvh must be called before tracing");
  if (n == 0) return DRT_OK;
  if (!o || !d || !out) return fail(c, DRT_E_INVALID, "null ray or output buffer");
  CK(c, cudaSetDevice(c->device));
  CK(c, c->dRayO.ensure(n));
  CK(c, c->dRayD.ensure(n));
  if (any) CK(c, c->dOcc.ensure(n));
  else CK(c, c->dHits.ensure(n));
  void* dout = any ? (void*)c->dOcc.p : (void*)c->dHits.p;
  const size_t outSize = any ? 1 : sizeof(drt_hit_rec);
  // Chunks of >= 1 Mi rays, at most kMaxChunks, round-robin over the pipeline streams: with pinned host buffers
  // the upload of the next chunk and the download of the previous one overlap the traversal of this one.
  uint64_t chunk = 1ull << 20;  // measured (tools/e2e_sweep.py): 1 Mi beats 128 Ki .. 512 Ki (small launches underfill the GPU) and 2 Mi
  if (const char* e = std::getenv("DRT_E2E_CHUNK")) chunk = std::max<uint64_t>(1024, std::strtoull(e, nullptr, 10));
  if ((n + chunk - 1) / chunk > (uint64_t)drt_ctx::kMaxChunks) chunk = (n + drt_ctx::kMaxChunks - 1) / drt_ctx::kMaxChunks;
  if (c->counting || c->exactWalk) chunk = n;  // one launch: the counters describe the whole batch
  int nChunks = 0;
  for (uint64_t first = 0; first < n; first += chunk, ++nChunks) {
    const uint64_t m = n - first < chunk ? n - first : chunk;
    const int slot = nChunks % drt_ctx::kPipe;
    cudaStream_t st = c->pipe[slot];
    CK(c, cudaMemcpyAsync(c->dRayO.p + first, o + 4 * first, m * 16, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(c->dRayD.p + first, d + 4 * first, m * 16, cudaMemcpyHostToDevice, st));
    CK(c, cudaEventRecord(c->chunkEv[2 * nChunks], st));
    int rc = traceDevice(c, any, c->dRayO.p + first, c->dRayD.p + first, m, (char*)dout + first * outSize, st, 1 + slot, first);
    if (rc != DRT_OK) return rc;
    CK(c, cudaEventRecord(c->chunkEv[2 * nChunks + 1], st));
    CK(c, cudaMemcpyAsync((char*)out + first * outSize, (char*)dout + first * outSize, m * outSize, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < drt_ctx::kPipe; ++i) CK(c, cudaStreamSynchronize(c->pipe[i]));
  double ms = 0.0;
  for (int k = 0; k < nChunks; ++k) {
    float t = 0.f;
    CK(c, cudaEventElapsedTime(&t, c->chunkEv[2 * k], c->chunkEv[2 * k + 1]));
    ms += t;
  }
  c->lastKernelMs = ms;  // sum of the chunks' launch-to-finish spans (they overlap copies of their neighbours)
  return DRT_OK;
}

int drt_trace_closest(drt_ctx* c, const float* o, const float* d, uint64_t n, drt_hit* hits) {
  return traceHost(c, false, o, d, n, hits);
}
int drt_trace_any(drt_ctx* c, const float* o, const float* d, uint64_t n, uint8_t* occluded) {
  return traceHost(c, true, o, d, n, occluded);
}
int drt_trace_closest_device(drt_ctx* c, const void* o, const void* d, uint64_t n, void* hits, void* stream) {
  if (!c) return DRT_E_INVALID;
  return traceDevice(c, false, o, d, n, hits, (cudaStream_t)stream);
}
int drt_trace_any_device(drt_ctx* c, const void* o, const void* d, uint64_t n, void* occ, void* stream) {
  if (!c) return DRT_E_INVALID;
  return traceDevice(c, true, o, d, n, occ, (cudaStream_t)stream);
}

int drt_set_counting(drt_ctx* c, int enabled) {
  if (!c) return DRT_E_INVALID;
  c->counting = enabled != 0;
  return DRT_OK;
}

int drt_set_kernel_variant(drt_ctx* c, int variant) {
  DRT_FORWARD_TO_PEERS(c, drt_set_kernel_variant(p_, variant));
  if (!c) return DRT_E_INVALID;
  if (variant != DRT_KERNEL_FAST && variant != DRT_KERNEL_EXACT_WALK && variant != DRT_KERNEL_FAST_V1 && variant != DRT_KERNEL_FAST_Q)
    return fail(c, DRT_E_INVALID, "unknown kernel variant");
  c->exactWalk = variant == DRT_KERNEL_EXACT_WALK;
  c->fastV1 = variant == DRT_KERNEL_FAST_V1 || (variant != DRT_KERNEL_FAST_Q && std::getenv("DRT_TRACE_V1") != nullptr);
  if (variant == DRT_KERNEL_FAST_Q) c->qMinPrims = 0;
  c->variantForced = variant != DRT_KERNEL_FAST;
  if (c->built && c->device != DRT_DEVICE_NONE && !c->useQ() && !c->wideUploaded) {
    int rc = uploadWideV1(c);
    if (rc != DRT_OK) return rc;
  }
  c->ts.wideQ = (c->built && c->useQ()) ? c->dWideQ.p : nullptr;
  c->ts.small = (c->built && c->useSmall()) ? c->dSmall.p : nullptr;
  return DRT_OK;
}

int drt_get_counters(drt_ctx* c, drt_counters* out) {
  if (!c || !out) return DRT_E_INVALID;
  if (c->device == DRT_DEVICE_NONE) return fail(c, DRT_E_NODEVICE, kNoDevice);
  CK(c, cudaSetDevice(c->device));
  DeviceCounters h;
  CK(c, cudaDeviceSynchronize());
  CK(c, cudaMemcpy(&h, c->dCounters.p, sizeof(h), cudaMemcpyDeviceToHost));
  out->rays = h.rays;
  out->nodes_visited = h.nodes_visited;
  out->prims_tested = h.prims_tested;
  out->hits = h.hits;
  return DRT_OK;
}

double drt_last_kernel_ms(const drt_ctx* c) { return c ? c->lastKernelMs : 0.0; }
uint64_t drt_kernel_launches(const drt_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
