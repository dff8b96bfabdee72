// Launchers of the wavefront render stages (render_kernels.cu).  All launches are asynchronous on
// `st`; queue sizes live in device memory (Wavefront::counts), so a whole batch is enqueued without a
// host round trip.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "render_types.h"

namespace drt {

// A run of pixels of the sample window [x0, x0+w) x [y0, ...), row-major, starting at `firstPixel`.
// Sharding: the window's pixels are cut into blocks of `blockPixels`; shard s of n owns blocks s, s+n, ...
// `firstPixel` counts pixels inside the shard (n = 1: plain row-major order).
struct PixelBatch {
  int32_t x0, y0, w;
  uint64_t firstPixel;
  uint32_t nPixels;
  uint32_t pass;  // visit number (random sampler: one visit per pass, random_sampler.dart:47-88)
  uint32_t shard, nShards, blockPixels;
};

enum { Q_EXT0 = 0, Q_EXT1 = 1, Q_SHADOW = 2, Q_MIS = 3, Q_HITS = 4, Q_COUNT = 8 };

cudaError_t launchSampler(const RenderParams& rp, const Wavefront& wf, const SampleArray* dArrays, int nArrays, int maxVals,
                          int maxOthers, const PixelBatch& pb, int numSMs, cudaStream_t st);
cudaError_t launchRaygen(const RenderParams& rp, const Wavefront& wf, const PixelBatch& pb, cudaStream_t st);
cudaError_t launchResetCounts(const Wavefront& wf, unsigned mask, cudaStream_t st);
// path integrator
cudaError_t launchShadePath(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int bounce, int cur,
                            RenderCounters* rc, int numSMs, cudaStream_t st);
cudaError_t launchResolveDirect(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int mode,
                                int nSamplesOfLight, int numSMs, cudaStream_t st);
// ambient occlusion
cudaError_t launchAoSetup(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int numSMs, cudaStream_t st);
cudaError_t launchAoGen(const RenderParams& rp, const Wavefront& wf, uint32_t firstHit, uint32_t maxHits, int numSMs,
                        cudaStream_t st);
cudaError_t launchAoCount(const RenderParams& rp, const Wavefront& wf, uint32_t firstHit, uint32_t maxHits, RenderCounters* rc,
                          int numSMs, cudaStream_t st);
// direct lighting
cudaError_t launchDirectSetup(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int weighted, int numSMs,
                              cudaStream_t st);
cudaError_t launchDirectSample(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int light, int j, int cur,
                               RenderCounters* rc, int numSMs, cudaStream_t st);
// whitted integrator (whitted_integrator.dart:26-78): emitted light + stream positions, then one launch per light
cudaError_t launchWhittedSetup(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int weighted, int numSMs,
                               cudaStream_t st);
cudaError_t launchWhittedSample(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int light, int cur,
                                RenderCounters* rc, int numSMs, cudaStream_t st);
// one SpecularReflect / SpecularTransmit call per vertex of queue `cur` (flags: BSDF_REFLECTION | BSDF_SPECULAR = 17 or
// BSDF_TRANSMISSION | BSDF_SPECULAR = 18); children go to queue cur ^ 1
cudaError_t launchSpecularStep(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int flags, int level,
                               int isNew, RenderCounters* rc, int numSMs, cudaStream_t st);
// film
cudaError_t launchFilm(const RenderParams& rp, const Wavefront& wf, uint32_t nSlots, RenderCounters* rc, cudaStream_t st);
cudaError_t launchFilmConvert(const RenderParams& rp, float* rgb, float* xyz, float* weight, cudaStream_t st);

// resolve mode bits: 1 = direct-lighting integrator (0 = path), 2 = first sample of a light, 4 = last sample of a
// light, 8 = last light (add the sum to L), 16 = strategy "one", 32 = the vertices belong to a specular chain: weight the
// sum by the chain weight (pendT)
enum { RESOLVE_PATH = 0, RESOLVE_DIRECT = 1, RESOLVE_FIRST_OF_LIGHT = 2, RESOLVE_LAST_OF_LIGHT = 4, RESOLVE_FINAL = 8, RESOLVE_ONE = 16,
       RESOLVE_WEIGHTED = 32, RESOLVE_WHITTED = 64 /* one unweighted light sample: add it as it is */ };

}  // namespace drt
