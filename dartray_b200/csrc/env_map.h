// Host-side set-up of an InfiniteAreaLight's radiance map (lib/lights/infinite_area_light.dart:276-306): the MIPMap pyramid
// the reference filters the map with (lib/core/mipmap.dart:142-222,341-355) and the Distribution2D it is importance-sampled
// from (lib/core/montecarlo.dart:25-48,222-237), flattened into one float array the device code indexes (GLight::envOffset).
#pragma once
#include <vector>

namespace drt {

// Appends, for a map of width x height RGB float32 texels (level 0 of the reference's MIPMap: power-of-two resolution) scaled by
// L at every lookup (infinite_area_light.dart:240-242):
//   texels 3 x W x H | conditional func W x H | conditional cdf H x (W + 1) | conditional funcInt H |
//   marginal func H | marginal cdf H + 1 | marginal funcInt 1
void appendEnvTables(int width, int height, const float* rgb, const float L[3], std::vector<float>* out);

}  // namespace drt
