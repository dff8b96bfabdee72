// The texture pass of the wavefront renderer (texture_kernels.cu) and the host-side set-up of its tables.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/drt.h"
#include "render_types.h"

namespace drt {

// drt_set_textures: validates the nodes, rebuilds every image's MIPMap pyramid the way MIPMap.texture does (lib/core/mipmap.dart:
// 142-166, including what SpectrumImage's shared return object does to the spectrum levels) and lays them out in one float array
// behind the 128-entry EWA weight table (mipmap.dart:168-176).  Returns false with *err set on invalid input.
bool buildTextureTables(uint32_t n, const drt_texture* nodes, const float* texels, uint64_t nTexelFloats, std::vector<GTex>* out,
                        std::vector<float>* data, std::string* err);

// For every vertex of extension queue `cur` whose material has a program: the whole DifferentialGeometry, computeDifferentials
// (hasDiff: the queue holds camera rays, whose differentials are regenerated from the slot's camera sample), shading geometry,
// Material.Bump, the textures, the material's getBSDF -> wf.hitLobes / hitCount / hitFrame of the slot.
cudaError_t launchTexturePass(const RenderParams& rp, const RenderScene& rs, const Wavefront& wf, int cur, int hasDiff, int numSMs,
                              cudaStream_t st);

}  // namespace drt
