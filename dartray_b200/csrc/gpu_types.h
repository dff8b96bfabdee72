// Device data layout of libdartray_gpu.so (see DESIGN.md "Data layout in HBM").
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define DRT_HD __host__ __device__
#else
#define DRT_HD
#endif

namespace drt {

// Child reference: >= 0 -> interior node index; < 0 -> leaf, bits = ~ref:
//   bits & 15        = primitive count (1..14), 15 = count stored in the first record's `leafCount`
//   (bits >> 4) & 1  = "box undecided": the float32 filter could not prove the leaf-box decision, the
//                      leaf phase must evaluate the reference's slab test exactly before the primitives
//   bits >> 5        = offset of the leaf's first record in GPrim[]
#define DRT_REF_EMPTY ((int32_t)0x7fffffff)  // unused slot of a wide node
static inline DRT_HD bool refIsLeaf(int32_t r) { return r < 0; }
static inline DRT_HD uint32_t refLeafOffset(int32_t r) { return ((uint32_t)~r) >> 5; }
static inline DRT_HD uint32_t refLeafCountField(int32_t r) { return ((uint32_t)~r) & 15u; }
static inline DRT_HD bool refLeafUndecided(int32_t r) { return (((uint32_t)~r) >> 4) & 1u; }
static inline DRT_HD int32_t refMarkUndecided(int32_t r) { return r & ~16; }
static inline int32_t makeLeafRef(uint32_t offset, uint32_t count) {
  uint32_t c = count < 15u ? count : 15u;
  return (int32_t)~((offset << 5) | c);
}

// One interior node = 64 bytes = four 128-bit loads.  It carries BOTH children's boxes (the
// reference's _LinearBVHNode, bvh_accel.dart:533-538, carries its own box; moving the boxes up one
// level lets one fetch decide near/far and is provably the same traversal).
//   q0 = c0.min.xyz, c0.max.x     q1 = c0.max.yz, c1.min.xy
//   q2 = c1.min.z, c1.max.xyz     q3 = ref0, ref1, axis, reference node number (for export/debug)
struct alignas(16) GNode {
  float c0min[3], c0max[3];
  float c1min[3], c1max[3];
  int32_t ref0, ref1;
  int32_t axis;
  int32_t refNode;
};
static_assert(sizeof(GNode) == 64, "GNode must be 64 bytes");

// Wide node = 128 bytes = eight 128-bit loads: two levels of the reference's binary tree collapsed.
// For the binary node P with children A (first) and B (second): slots 0,1 hold A's children (or A
// itself in slot 0 when A is a leaf), slots 2,3 hold B's.  The reference's visiting order is
// recovered from the three split axes: side = dirIsNeg[axisP] ? B : A first, and inside a side
// dirIsNeg[axisSide] ? second : first (bvh_accel.dart:147-153).  A's and B's own boxes are not
// stored: a child box that passes the slab test implies its parent box passes (monotone rounding,
// DESIGN.md), so skipping them cannot change which leaves are tested.
struct alignas(16) GNode4 {
  float box[4][6];   // per slot: (min.x, max.x), (min.y, max.y), (min.z, max.z) — each pair is one f32x2 operand
  int32_t ref[4];    // child reference or DRT_REF_EMPTY
  int32_t axisP, axisA, axisB;  // bits 0-1: split axis; bit 2: the pair is stored swapped (larger box first)
  int32_t orderLut;  // closest-hit visiting decisions per dirIsNeg octant, 3 bits each (see bvh_builder.cpp)
};
static_assert(sizeof(GNode4) == 128, "GNode4 must be 128 bytes");

// Quantised wide node = 64 bytes = two 256-bit loads (half the L1TEX wavefronts of GNode4): the same four slots, each box on
// an 8-bit grid local to the node.  Plane position = origin[a] + q * 2^e[a] (exact in binary64); the builder
// (bvh_builder.cpp, quantiseNode) guarantees  decoded lo <= true lo - g * 2^e  and  decoded hi >= true hi + g * 2^e  with
// g = 2^-6 of a grid step, which pays for the float32 evaluation error of the traversal kernel (trace_fast2.cu).  The
// boxes are only ever CONSERVATIVE: every leaf-box decision is made exactly, in binary64 on the box rebuilt from the
// leaf's primitives, in the leaf phase.
//   words 0-2  origin.xyz (float32)
//   word  3    high halves of the float32 values 2^(e+15): x in bits 0-15, y in bits 16-31
//   word  4    the same for z in bits 0-15
//   word  5    orderLut (as GNode4)
//   words 6-11 qx[2], qy[2], qz[2]: bytes (lo, hi) of slot 2j, (lo, hi) of slot 2j+1; an EMPTY slot is (255, 0)
//   words 12-15 the four child references (as GNode4)
struct alignas(32) GNode4Q {
  float origin[3];
  uint32_t scaleXY, scaleZ, orderLut;
  uint32_t q[3][2];
  int32_t ref[4];
};
static_assert(sizeof(GNode4Q) == 64, "GNode4Q must be 64 bytes");
#define DRT_Q_GUARD 0.015625   // g: 2^-6 grid steps
#define DRT_Q_EXP_MIN (-60)    // grid step 2^e, e in [DRT_Q_EXP_MIN, DRT_Q_EXP_MAX]
#define DRT_Q_EXP_MAX 62
#define DRT_Q_COORD_MAX 4.611686018427388e18  // 2^62: |coordinate| bound of a quantisable scene and of a "fast" ray origin

// One leaf primitive record = 48 bytes = three 128-bit loads, stored in leaf order.
// Triangle: the three ORIGINAL float32 world-space vertices (triangle.dart:47-50 reads exactly
// these; edges are formed in f64 on the fly so the arithmetic matches the reference bit for bit).
//   q0 = p1.xyz, primId     q1 = p2.xyz, leafCount     q2 = p3.xyz, kind|sphereIndex<<1
struct alignas(16) GPrim {
  float p1[3];
  int32_t primId;
  float p2[3];
  int32_t leafCount;
  float p3[3];
  int32_t kindSphere;  // bit0: 0 = triangle, 1 = sphere; bits 1.. = index into GSphere[]
};
static_assert(sizeof(GPrim) == 48, "GPrim must be 48 bytes");

// sphere.dart:24-32: transforms are float32 matrices, the shape parameters are doubles.  The record also carries
// the path's other quadric, Disk (disk.dart:24-31): shape == 1 with height / radius / innerRadius / phiMax.
struct alignas(16) GSphere {
  float w2o[12];  // rows 0..2 of worldToObject (affine)
  float w2oRow3[4];
  float o2w[12];
  float o2wRow3[4];
  double radius, zmin, zmax, phiMax, thetaMin, thetaMax;
  float wmin[3], wmax[3];  // world bound (sphere.dart:34-37 + shape.dart:38-40), = the leaf box of a 1-sphere leaf
  int32_t shape;           // 0 sphere, 1 disk, 2 cylinder, 3 cone, 4 paraboloid, 5 hyperboloid, 6 a TransformedPrimitive (instance)
  int32_t instance;        // shape 6: index into TraceScene::instances (wmin / wmax hold its world bound)
  double height, innerRadius;  // disk (both), cone (height)
  float hp1[3], hp2[3];        // hyperboloid.dart:24: the two Points after the constructor's swap (:36-39)
  double ha, hc;               // hyperboloid.dart:40-48 implicit coefficients; `radius` holds rmax
};

// Scenes of at most DRT_SMALL_MAX_LEAVES leaves (BASELINE.json config 4: 25 primitives) skip the tree: the reference's walk visits
// the leaves in an order that depends on the ray's dirIsNeg octant only (bvh_accel.dart:147-153), tests a leaf's primitives iff the
// leaf's own box test passes at that moment, and every interior test is implied by a leaf's (trace_fast.cu, argument (1)).  So the
// small-scene kernel (traceSmallKernel, trace_fast.cu) walks the octant's leaf list and decides every leaf box exactly.
#define DRT_SMALL_MAX_LEAVES 32
struct alignas(16) GSmallLeaf {
  float lo[3];
  int32_t ref;   // the leaf reference (offset / count of its GPrim records)
  float hi[3];   // box = union of the primitives' world bounds: what the reference's leaf node holds (bvh_accel.dart:262-281)
  int32_t pad;
};
struct alignas(16) GSmallScene {
  int32_t nLeaves, pad[3];
  GSmallLeaf leaf[DRT_SMALL_MAX_LEAVES];         // in storage (depth-first, first child first) order
  uint8_t order[8][DRT_SMALL_MAX_LEAVES];        // order[octant][k] = k-th leaf the reference's walk reaches for rays of that octant
  uint8_t position[8][DRT_SMALL_MAX_LEAVES];     // the inverse: position[octant][leaf]
};

struct GInstance;  // anim_transform.h
struct GObject;

struct TraceScene {
  const GNode* nodes;    // binary layout (exact-walk / counting kernel)
  const GNode4* wide;    // collapsed layout, float32 boxes (v1 production kernel; fallback)
  const GNode4Q* wideQ;  // the same nodes with quantised boxes (v2 production kernel); nullptr when a node is not quantisable
  int32_t wideRootRef;
  const GPrim* prims;
  const GSphere* spheres;
  float rootMin[3], rootMax[3];  // box of reference node 0, tested first (bvh_accel.dart:123-125)
  int32_t rootRef;
  int32_t empty;  // 1 -> no primitives (bvh_accel.dart:102-104)
  int32_t quadMode;  // which leaf code the scene needs: 0 triangles only, 1 + spheres / disks, 2 + the other quadrics
  // TransformedPrimitives (transformed_primitive.dart): a top-level leaf record of kind "quadric" whose GSphere has shape 6 stands
  // for instance GSphere::instance; the objects' binary nodes / leaf records sit behind the top level's in `nodes` / `prims`.
  // Scenes with instances run the literal walk (traceKernel), which descends into the object with the transformed ray.
  const GSmallScene* small;  // leaf lists of a scene with <= DRT_SMALL_MAX_LEAVES leaves and no instances, else nullptr
  const GInstance* instances;
  const GObject* objects;
  int32_t nInstances;
};

struct DeviceCounters {
  unsigned long long rays, nodes_visited, prims_tested, hits;
};

}  // namespace drt
