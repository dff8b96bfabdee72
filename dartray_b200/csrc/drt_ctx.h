// Internal: the context object behind the C ABI (include/drt.h), shared by drt_api.cu (scene, BVH,
// ray queries) and render_api.cu (wavefront renderer).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/drt.h"
#include "anim_transform.h"
#include "bvh_builder.h"
#include "gpu_types.h"
#include "render_types.h"
#include "trace_kernels.h"

using namespace drt;

static_assert(sizeof(drt_hit) == sizeof(drt_hit_rec), "hit record layout");

// a quadric: sphere (shape 0), disk (1: height, radius, innerRadius, phiMaxDeg), cylinder (2), cone (3: height, radius),
// paraboloid (4), hyperboloid (5: prm = p1, p2 as given)
struct HostSphere {
  float o2w[16], w2o[16];
  double radius = 0.0, zmin = 0.0, zmax = 0.0, phiMaxDeg = 360.0;
  int shape = 0;
  double height = 0.0, innerRadius = 0.0;
  double prm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct drt_ctx {
  int device = 0;
  std::string err;
  // staged scene (host)
  std::vector<float> P;
  std::vector<uint32_t> idx;
  std::vector<int32_t> matOf, lightOf;
  std::vector<uint8_t> revOf;
  std::vector<HostSphere> spheres;
  std::vector<int32_t> sphMat, sphLight;
  std::vector<uint8_t> sphRev;
  std::vector<uint32_t> order;
  // BVH
  bool built = false;
  uint64_t buildSerial = 0;  // bumped by every drt_build_bvh
  BuiltBvh bvh;
  drt_bvh_info info{};
  DevBuf<GNode> dNodes;
  DevBuf<GNode4> dWide;
  DevBuf<GNode4Q> dWideQ;
  bool wideQOk = false;  // the scene's wide nodes could be quantised (bvh_builder.cpp quantiseNode)
  bool wideUploaded = false;  // dWide holds this build's float32 wide nodes
  bool fastV1 = false;   // DRT_KERNEL_FAST_V1 / env DRT_TRACE_V1: the float32-box kernel of trace_fast.cu
  // Scenes below this many primitives keep the float32-box kernel: with a handful of nodes per ray the walk is all leaf
  // phase, where the second generation pays a binary64 box test per leaf (config 4, 25 primitives: 549 vs 417 M samples/s;
  // soup(8), 16 K triangles: 10.7 vs 8.5 Grays/s; soup(64), 127 K: 4.2 vs 4.6; soup_1m: 1.88 vs 2.36 Grays/s — tools/size_sweep.sh).  Env DRT_Q_MIN_PRIMS overrides (A/B runs).
  uint32_t qMinPrims = 65536;
  bool useQ() const { return wideQOk && !fastV1 && nprims() >= qMinPrims; }
  // Scenes of <= DRT_SMALL_MAX_LEAVES leaves without instances: the leaf-list kernel (traceSmallKernel) unless a variant is forced
  // (drt_set_kernel_variant other than DRT_KERNEL_FAST) or env DRT_NO_SMALL is set (A/B runs)
  DevBuf<GSmallScene> dSmall;
  bool smallOk = false, variantForced = false, smallOff = false;
  bool useSmall() const { return smallOk && !variantForced && !fastV1 && !smallOff; }
  DevBuf<GPrim> dPrims;
  DevBuf<GSphere> dSpheres;
  DevBuf<DeviceCounters> dCounters;
  // work counters of the persistent traversal kernels: slot 0 = the renderer (ctx stream), 1..kPipe = the host-buffer
  // pipeline streams, then a ring for drt_trace_*_device launches on caller streams (each launch takes the next slot and
  // waits for the slot's previous user, so concurrent launches never share a counter)
  static const int kRing = 32;
  DevBuf<unsigned long long> dNextRay;
  cudaEvent_t ringEv[kRing] = {};
  bool ringUsed[kRing] = {};
  int ringNext = 0;
  int numSMs = 148;
  TraceScene ts{};
  // ray staging for host-buffer calls
  DevBuf<float4> dRayO, dRayD;
  DevBuf<drt_hit_rec> dHits;
  DevBuf<uint8_t> dOcc;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host-buffer queries are pipelined over three streams: H2D of chunk k+1, kernel k, D2H of chunk k-1 overlap
  static const int kPipe = 3, kMaxChunks = 64;
  cudaStream_t pipe[kPipe] = {nullptr, nullptr, nullptr};
  cudaEvent_t chunkEv[2 * kMaxChunks] = {};
  bool counting = false;
  bool exactWalk = false;
  double lastKernelMs = 0.0;
  uint64_t launches = 0;
  // drt_set_mesh_shading: per-vertex N / S (object space) / uv, mesh of each triangle, per-mesh transforms + flags
  std::vector<float> vertN, vertS, vertUV;
  std::vector<uint32_t> meshOfTri;
  std::vector<float> meshO2W, meshW2O;  // nmeshes x 16
  std::vector<uint8_t> meshFlags;

  // drt_set_instances: TransformedPrimitives (transformed_primitive.dart).  objects = the aggregates they wrap (primitive ids in the
  // refined order handed to the nested accelerator + its parameters), instances = AnimatedTransform after its constructor + object
  struct HostObject {
    std::vector<uint32_t> order;
    int split = 2, maxPrims = 1;
  };
  std::vector<HostObject> objects;
  std::vector<GInstance> instances;
  std::vector<uint32_t> recPrimIds;  // primitive id of every GPrim record on the device (the top level's, then the objects')
  DevBuf<GInstance> dInstances;
  DevBuf<GObject> dObjects;
  // drt_set_ray_times: times of the rays of the host-buffer / device-buffer trace calls (scenes with instances; empty: time 0)
  std::vector<double> rayTimes;
  DevBuf<double> dRayTimes;
  bool rayTimesDirty = false;

  struct RenderState* render = nullptr;  // render_api.cu

  // drt_create_multi: this context drives device_ids[0]; `peers` are the contexts of the other devices (owned: destroyed with
  // this one).  Every scene / camera / film / sampler setter is applied to the peers too; the BVH is built once, here, and its
  // arrays uploaded to every device; renders are split over all devices and the films summed into this one's.
  std::vector<drt_ctx*> peers;
  drt_ctx* owner = nullptr;  // set on a peer: the context that holds the host copy of the built BVH
  bool peerAccessTried = false;
  const BuiltBvh& hostBvh() const { return owner ? owner->bvh : bvh; }

  uint32_t ntris() const { return (uint32_t)(idx.size() / 3); }
  uint32_t nprims() const { return ntris() + (uint32_t)spheres.size(); }
};

#define CK(ctx, call)                                                                    \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
      return e__ == cudaErrorMemoryAllocation ? DRT_E_NOMEM : DRT_E_CUDA;                \
    }                                                                                    \
  } while (0)

// First line of every setter: apply the call to the other devices of a multi-device context (`p_` names the peer).
#define DRT_FORWARD_TO_PEERS(ctx, call)                                                              \
  do {                                                                                               \
    if (ctx)                                                                                         \
      for (drt_ctx* p_ : (ctx)->peers) {                                                             \
        int rc_ = (call);                                                                            \
        if (rc_ != DRT_OK) {                                                                         \
          (ctx)->err = "device " + std::to_string(p_->device) + ": " + p_->err;                      \
          return rc_;                                                                                \
        }                                                                                            \
      }                                                                                              \
  } while (0)

static inline int fail(drt_ctx* c, int code, const char* msg) {
  if (c) c->err = msg;
  return code;
}


static const char* const kNoDevice =
    "context has no CUDA device (DRT_DEVICE_NONE): queries need a GPU, there is no CPU fallback";

// render_api.cu
void drtRenderStateDestroy(drt_ctx* c);
