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
  const uint32_t* list;  // adaptive sampler, second visit: the window-linear indices of the pixels to supersample (device), or null
};

enum { ESCAPE_CAMERA = 0, ESCAPE_PATH = 1, ESCAPE_WEIGHTED = 2 };
enum { Q_EXT0 = 0, Q_EXT1 = 1, Q_SHADOW = 2, Q_MIS = 3, Q_HITS = 4, Q_COUNT = 8 };

// render_kernels.cu is compiled twice (shade_device.cuh, DRT_EXTRA): `plain` without per-vertex mesh attributes and the
// cylinder / cone / paraboloid / hyperboloid shapes — the kernels the benchmark scenes run, byte for byte what they were
// before those features existed — and `extra` with them.  render_api.cu picks by RenderScene::extra.
namespace plain {
#include "render_launchers.inc"
}
namespace extra {
#include "render_launchers.inc"
}
// float32 arithmetic (render_kernels_f32.cu): launchShadePath and launchResolveDirect only are defined
namespace plainf {
#include "render_launchers.inc"
}
namespace extraf {  // the same two launchers from the `extra` build (render_kernels_f32x.cu)
#include "render_launchers.inc"
}

// resolve mode bits: 1 = direct-lighting integrator (0 = path), 2 = first sample of a light, 4 = last sample of a
// light, 8 = last light (add the sum to L), 16 = strategy "one", 32 = the vertices belong to a specular chain: weight the
// sum by the chain weight (pendT)
enum { RESOLVE_PATH = 0, RESOLVE_DIRECT = 1, RESOLVE_FIRST_OF_LIGHT = 2, RESOLVE_LAST_OF_LIGHT = 4, RESOLVE_FINAL = 8, RESOLVE_ONE = 16,
       RESOLVE_WEIGHTED = 32, RESOLVE_WHITTED = 64 /* one unweighted light sample: add it as it is */ };

}  // namespace drt
