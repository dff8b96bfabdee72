// Host-side BVH construction for libdartray_gpu.so.
//
// Produces the SAME tree as the reference's BVHAccel constructor
// (/root/reference/lib/accelerators/bvh_accel.dart:41-91, 228-437): same split decisions (float32
// SAH cost array, 12 buckets, equal-count split at <= 4 primitives), same in-leaf primitive order.
// The output is (a) the GPU layout of gpu_types.h and (b) the reference's linear numbering for
// drt_bvh_export.  It shares no code with oracle/.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "gpu_types.h"

namespace drt {

struct PrimBounds {
  float bmin[3], bmax[3];
};

struct RefNode {  // the reference's _LinearBVHNode numbering, for drt_bvh_export
  float bmin[3], bmax[3];
  int32_t offset, nPrimitives, axis;
};

struct BuiltBvh {
  std::vector<GNode> nodes;            // interior nodes, DFS order
  std::vector<GNode4> wide;            // two-level collapsed nodes, DFS order
  std::vector<GNode4Q> wideQ;          // wide[i] with quantised boxes (same indices, same references)
  bool wideQOk = true;                 // false: some node cannot be quantised (non-finite or > 2^62 coordinates)
  int32_t wideRootRef = 0;
  std::vector<uint32_t> leafPrimIds;   // primitive ids in GPU leaf order (DFS, left first)
  std::vector<uint32_t> leafCounts;    // for record i: count of its leaf if i is the leaf's first record, else 0
  std::vector<RefNode> refNodes;       // reference numbering
  std::vector<uint32_t> refOrdered;    // reference `primitives` order after the build
  float rootMin[3], rootMax[3];
  int32_t rootRef = 0;
  uint32_t nLeaves = 0, maxLeafPrims = 0, maxDepth = 0;
};

// `order[i]` = primitive id of the i-th refined primitive (bvh_accel.dart:59-65 buildData[i]);
// `bounds` is indexed by primitive id.  Returns false and fills `err` on failure.
bool buildBvh(const std::vector<PrimBounds>& bounds, const std::vector<uint32_t>& order, int splitMethod,
              int maxPrimsInNode, BuiltBvh* out, std::string* err);

}  // namespace drt
