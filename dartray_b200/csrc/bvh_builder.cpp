// Host BVH builder: reference-identical topology, GPU-oriented output.  See bvh_builder.h.
#include "bvh_builder.h"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <future>
#include <thread>
#include <limits>

namespace drt {
namespace {

constexpr float kFInf = std::numeric_limits<float>::infinity();
constexpr int kBuckets = 12;  // bvh_accel.dart:319

struct Box {
  float lo[3], hi[3];
  void clear() {
    for (int a = 0; a < 3; ++a) { lo[a] = kFInf; hi[a] = -kFInf; }
  }
  void grow(const float* l, const float* h) {
    for (int a = 0; a < 3; ++a) { lo[a] = l[a] < lo[a] ? l[a] : lo[a]; hi[a] = h[a] > hi[a] ? h[a] : hi[a]; }
  }
  void growPoint(const float* p) { grow(p, p); }
  // bbox.dart:164-167: the diagonal is a float32 Vector, the area expression is float64.
  double area() const {
    double dx = (float)((double)hi[0] - lo[0]), dy = (float)((double)hi[1] - lo[1]), dz = (float)((double)hi[2] - lo[2]);
    return 2.0 * (dx * dy + dx * dz + dy * dz);
  }
};

struct PrimRef {
  float lo[3], hi[3];
  float cen[3];
  uint32_t id;
};

struct TNode {
  Box box;
  int32_t left = -1, right = -1;  // pool indices; -1 -> leaf
  uint32_t first = 0, count = 0;  // leaf: range in the builder's PrimRef array
  int32_t axis = 0;
  // subtree totals, filled when the children return: they fix every output offset, so the flattening passes below can
  // write disjoint ranges of pre-sized arrays from concurrent tasks
  uint32_t nNodes = 1, nPrims = 0, nInterior = 0, nWide = 0, nLeaves = 1, maxLeaf = 0, depthBelow = 0;
};

// The node pool: 2n slots, handed out by primitive count (TreeBuilder::build).  Raw storage: a slot is constructed by the task
// that builds its node (value-initialising 1.5 GB up front for 10 M primitives cost half a second, serially).
struct NodePool {
  TNode* p = nullptr;
  size_t n = 0;
  explicit NodePool(size_t count) : p(static_cast<TNode*>(std::malloc(std::max<size_t>(count, 1) * sizeof(TNode)))), n(count) {}
  ~NodePool() { std::free(p); }
  NodePool(const NodePool&) = delete;
  NodePool& operator=(const NodePool&) = delete;
  TNode& operator[](size_t i) { return p[i]; }
  const TNode& operator[](size_t i) const { return p[i]; }
  size_t size() const { return n; }
};
struct Arena {
  NodePool pool;
  explicit Arena(size_t count) : pool(count) {}
};

class TreeBuilder {
 public:
  TreeBuilder(std::vector<PrimRef>& refs, int split, int maxPrims) : refs_(refs), split_(split), maxPrims_(maxPrims) {}

  // Builds [start, end) into pool slot `me`.  The subtree of a node over n primitives owns the slots [me, me + 2n - 1): its first
  // child sits at me + 1, its second at me + 2 * (primitives of the first) — fixed by the split alone, so subtrees can be
  // built by concurrent tasks straight into the one pre-sized pool and the result does not depend on thread timing.
  int32_t build(Arena& arena, int32_t me, uint32_t start, uint32_t end, int depth) {
    Box box;
    box.clear();
    for (uint32_t i = start; i < end; ++i) box.grow(refs_[i].lo, refs_[i].hi);
    new (&arena.pool[me]) TNode();
    arena.pool[me].box = box;
    uint32_t n = end - start;
    if (n == 1) return makeLeaf(arena, me, start, end);

    Box cb;
    cb.clear();
    for (uint32_t i = start; i < end; ++i) cb.growPoint(refs_[i].cen);
    int dim = maxExtent(cb);
    if (cb.hi[dim] == cb.lo[dim]) return makeLeaf(arena, me, start, end);  // bvh_accel.dart:265-274

    uint32_t mid = (start + end) / 2;
    const double cmin = cb.lo[dim], cmax = cb.hi[dim];
    bool needEqual = false;
    if (split_ == 0) {  // SPLIT_MIDDLE, bvh_accel.dart:282-304
      double pmid = 0.5 * (cmin + cmax);
      mid = hoarePartition(start, end, [&](const PrimRef& r) { return (double)r.cen[dim] < pmid; });
      if (mid == start || mid == end) needEqual = true;
    } else if (split_ == 1) {
      needEqual = true;
    } else if (n <= 4) {  // bvh_accel.dart:313-316
      needEqual = true;
    } else {
      // 12-bucket SAH, bvh_accel.dart:318-403
      uint32_t cnt[kBuckets] = {0};
      Box bb[kBuckets];
      for (auto& b : bb) b.clear();
      const double inv = cmax - cmin;
      for (uint32_t i = start; i < end; ++i) {
        int b = (int)(kBuckets * (((double)refs_[i].cen[dim] - cmin) / inv));
        if (b == kBuckets) b = kBuckets - 1;
        cnt[b]++;
        bb[b].grow(refs_[i].lo, refs_[i].hi);
      }
      // prefix / suffix sweeps give the same unions as the reference's O(B^2) loops (min/max are exact)
      Box pre[kBuckets], suf[kBuckets];
      uint32_t preN[kBuckets], sufN[kBuckets];
      Box acc;
      acc.clear();
      uint32_t accN = 0;
      for (int i = 0; i < kBuckets; ++i) { acc.grow(bb[i].lo, bb[i].hi); accN += cnt[i]; pre[i] = acc; preN[i] = accN; }
      acc.clear();
      accN = 0;
      for (int i = kBuckets - 1; i >= 0; --i) { acc.grow(bb[i].lo, bb[i].hi); accN += cnt[i]; suf[i] = acc; sufN[i] = accN; }
      const double total = box.area();
      float cost[kBuckets - 1];  // Float32List in the reference (:345): costs are rounded to float32
      for (int i = 0; i < kBuckets - 1; ++i)
        cost[i] = (float)(0.125 + ((double)preN[i] * pre[i].area() + (double)sufN[i + 1] * suf[i + 1].area()) / total);
      double minCost = cost[0];
      int minSplit = 0;
      for (int i = 1; i < kBuckets - 1; ++i)
        if ((double)cost[i] < minCost) { minCost = cost[i]; minSplit = i; }
      if (n > (uint32_t)maxPrims_ || minCost < (double)n) {
        mid = hoarePartition(start, end, [&](const PrimRef& r) {
          int b = (int)std::floor(kBuckets * (((double)r.cen[dim] - cmin) / inv));
          if (b == kBuckets) b = kBuckets - 1;
          return b <= minSplit;
        });
      } else {
        return makeLeaf(arena, me, start, end);
      }
    }
    if (needEqual) {
      mid = (start + end) / 2;
      sortByCentroid(start, end, dim);
    }

    // Large subtrees are built concurrently (at most a few tasks per hardware thread in flight).
    int32_t l, r;
    const int32_t lSlot = me + 1, rSlot = me + 2 * (int32_t)(mid - start);
    if (n >= (1u << 14) && liveTasks_.load(std::memory_order_relaxed) < maxTasks_) {
      liveTasks_.fetch_add(1, std::memory_order_relaxed);
      auto fut = std::async(std::launch::async, [&] {
        int32_t v = build(arena, rSlot, mid, end, depth + 1);
        liveTasks_.fetch_sub(1, std::memory_order_relaxed);
        return v;
      });
      l = build(arena, lSlot, start, mid, depth + 1);
      r = fut.get();
    } else {
      l = build(arena, lSlot, start, mid, depth + 1);
      r = build(arena, rSlot, mid, end, depth + 1);
    }
    TNode& node = arena.pool[me];
    node.left = l;
    node.right = r;
    node.axis = dim;
    {
      const TNode &a = arena.pool[l], &b = arena.pool[r];
      node.nNodes = 1 + a.nNodes + b.nNodes;
      node.nPrims = a.nPrims + b.nPrims;
      node.nInterior = 1 + a.nInterior + b.nInterior;
      node.nLeaves = a.nLeaves + b.nLeaves;
      node.maxLeaf = std::max(a.maxLeaf, b.maxLeaf);
      node.depthBelow = 1 + std::max(a.depthBelow, b.depthBelow);
      // wide nodes of the two-level collapse rooted here (Flattener::emitWide): this one + those rooted at the grandchildren
      auto under = [&](const TNode& side) { return side.left < 0 ? 0u : arena.pool[side.left].nWide + arena.pool[side.right].nWide; };
      node.nWide = 1 + under(a) + under(b);
    }
    // bvh_accel.dart:521 — union of the children's boxes (equals `box`: min/max are exact)
    return me;
  }

 private:
  std::vector<PrimRef>& refs_;
  int split_, maxPrims_;
  std::atomic<int> liveTasks_{0};
  const int maxTasks_ = 4 * (int)std::max(1u, std::thread::hardware_concurrency());

  static int maxExtent(const Box& b) {  // bbox.dart:174-183 on the float32 diagonal
    float dx = (float)((double)b.hi[0] - b.lo[0]), dy = (float)((double)b.hi[1] - b.lo[1]),
          dz = (float)((double)b.hi[2] - b.lo[2]);
    if (dx > dy && dx > dz) return 0;
    if (dy > dz) return 1;
    return 2;
  }

  int32_t makeLeaf(Arena& arena, int32_t me, uint32_t start, uint32_t end) {
    arena.pool[me].first = start;
    arena.pool[me].count = end - start;
    arena.pool[me].nPrims = arena.pool[me].maxLeaf = end - start;
    return me;
  }

  // common.dart:256-284: element order after the call matters for later equal-count sorts.
  template <class Pred>
  uint32_t hoarePartition(uint32_t first, uint32_t last, Pred pred) {
    while (first < last) {
      while (pred(refs_[first])) {
        if (++first == last) return first;
      }
      do {
        if (first == --last) return first;
      } while (!pred(refs_[last]));
      std::swap(refs_[first], refs_[last]);
      ++first;
    }
    return first;
  }

  // common.dart:289-297 + Dart List.sort on a comparator that never returns 0: for the <= 32
  // element ranges the SAH path produces this is Dart's insertion sort, which moves an element in
  // front of every predecessor that is NOT strictly smaller.
  void sortByCentroid(uint32_t first, uint32_t last, int dim) {
    if (last - first > 32) {
      // splitmethod middle/equal on big ranges: same ordering rule, O(n log n).  Elements with
      // equal keys end up in reverse input order under the insertion rule above.
      std::reverse(refs_.begin() + first, refs_.begin() + last);
      std::stable_sort(refs_.begin() + first, refs_.begin() + last,
                       [dim](const PrimRef& a, const PrimRef& b) { return a.cen[dim] < b.cen[dim]; });
      return;
    }
    for (uint32_t i = first + 1; i < last; ++i) {
      PrimRef el = refs_[i];
      uint32_t j = i;
      while (j > first && !(refs_[j - 1].cen[dim] < el.cen[dim])) {
        refs_[j] = refs_[j - 1];
        --j;
      }
      refs_[j] = el;
    }
  }
};

struct Flattener {
  const NodePool& pool;
  const std::vector<PrimRef>& refs;
  BuiltBvh* out;

  // Two subtrees at once when they are big enough to pay for a task.
  template <class FA, class FB>
  static void both(bool parallel, FA&& fa, FB&& fb) {
    if (parallel) {
      auto fut = std::async(std::launch::async, [&] { fa(); });
      fb();
      fut.get();
    } else {
      fa();
      fb();
    }
  }
  static constexpr uint32_t kTaskNodes = 1u << 15;

  // One pass over the finished tree writes the three things that are numbered differently:
  //  * the reference numbering: depth first, first child at n+1, second child after the first one's subtree (bvh_accel.dart:419-437);
  //    node t gets index `myRef`;
  //  * the reference's `primitives` order: it appends a leaf's primitives when the leaf is created and builds the SECOND child first
  //    (bvh_accel.dart:407-411) -> right-first leaf order; `off` is the first slot of t's primitives;
  //  * the binary GPU layout: interior nodes in DFS order (`my`: index of t among them), leaf records in DFS (left-first) order, which
  //    is the order the in-place partitions left the PrimRef array in.
  // Every index follows from the subtree totals, so the two subtrees of a node are independent tasks.
  void emitBinary(int32_t t, int32_t my, int32_t myRef, uint32_t off) {
    const TNode& n = pool[t];
    RefNode rn;
    std::memcpy(rn.bmin, n.box.lo, 12);
    std::memcpy(rn.bmax, n.box.hi, 12);
    rn.axis = 0;
    if (n.left < 0) {
      rn.nPrimitives = (int32_t)n.count;
      rn.offset = (int32_t)off;
      out->refNodes[myRef] = rn;
      for (uint32_t i = 0; i < n.count; ++i) {
        const uint32_t id = refs[n.first + i].id;
        out->refOrdered[off + i] = id;
        out->leafPrimIds[n.first + i] = id;
        out->leafCounts[n.first + i] = i == 0 ? n.count : 0;
      }
      return;
    }
    const TNode &a = pool[n.left], &b = pool[n.right];
    const int32_t secondRef = myRef + 1 + (int32_t)a.nNodes;
    rn.axis = n.axis;
    rn.nPrimitives = 0;
    rn.offset = secondRef;
    out->refNodes[myRef] = rn;
    const int32_t myLeft = my + 1, myRight = my + 1 + (int32_t)a.nInterior;
    GNode g;
    std::memcpy(g.c0min, a.box.lo, 12);
    std::memcpy(g.c0max, a.box.hi, 12);
    std::memcpy(g.c1min, b.box.lo, 12);
    std::memcpy(g.c1max, b.box.hi, 12);
    g.ref0 = a.left < 0 ? leafRefOf(n.left) : myLeft;
    g.ref1 = b.left < 0 ? leafRefOf(n.right) : myRight;
    g.axis = n.axis;
    g.refNode = myRef;
    out->nodes[my] = g;
    both(n.nNodes >= kTaskNodes, [&] { emitBinary(n.left, myLeft, myRef + 1, off + b.nPrims); },
         [&] { emitBinary(n.right, myRight, secondRef, off); });
  }

  // Wide layout: collapse P with its two children (see GNode4).  Independent of emitBinary(): both derive a leaf's reference from
  // its range in the PrimRef array, so the two layouts index the same GPrim[] order.
  // A leaf's records start at its range in the builder's PrimRef array: the in-place partitions leave that array in leaf
  // (depth-first, first child first) order, which is the order of the GPrim records.
  int32_t leafRefOf(int32_t t) const { return makeLeafRef(pool[t].first, pool[t].count); }
  int32_t emitWide(int32_t t, int32_t my) {
    const TNode& n = pool[t];
    if (n.left < 0) return leafRefOf(t);
    GNode4 g;
    std::memset(&g, 0, sizeof(g));
    for (int k = 0; k < 4; ++k) {
      g.ref[k] = DRT_REF_EMPTY;
      for (int a = 0; a < 6; ++a) g.box[k][a] = 0.f;
    }
    g.axisP = n.axis;
    // Slot order = any-hit visiting order (that kernel walks the slots as stored: its answer does not depend on
    // the order): the larger box first, inside each side and between the sides.  A swap is recorded in bit 2 of
    // the axis field, which the closest-hit kernel XORs into dirIsNeg[axis] to recover the reference's order.
    int32_t sides[2] = {n.left, n.right};
    if (pool[sides[1]].box.area() > pool[sides[0]].box.area()) {
      std::swap(sides[0], sides[1]);
      g.axisP |= 4;
    }
    int32_t slotKid[4] = {-1, -1, -1, -1}, slotBase[4] = {0, 0, 0, 0};
    int32_t next = my + 1;  // the kids' wide subtrees follow in slot order
    for (int sIdx = 0; sIdx < 2; ++sIdx) {
      const TNode& side = pool[sides[sIdx]];
      int base = 2 * sIdx;
      int32_t kids[2];
      int nk;
      if (side.left < 0) { kids[0] = sides[sIdx]; nk = 1; }
      else {
        kids[0] = side.left; kids[1] = side.right; nk = 2;
        int32_t ax = side.axis;
        if (pool[kids[1]].box.area() > pool[kids[0]].box.area()) {
          std::swap(kids[0], kids[1]);
          ax |= 4;
        }
        (sIdx == 0 ? g.axisA : g.axisB) = ax;
      }
      for (int k = 0; k < nk; ++k) {
        const TNode& c = pool[kids[k]];
        for (int a = 0; a < 3; ++a) {  // (lo, hi) pairs per axis: one packed f32x2 operand each
          g.box[base + k][2 * a] = c.box.lo[a];
          g.box[base + k][2 * a + 1] = c.box.hi[a];
        }
        slotKid[base + k] = kids[k];
        slotBase[base + k] = next;
        if (c.left >= 0) next += (int32_t)c.nWide;
      }
    }
    {
      auto run = [&](int k) { if (slotKid[k] >= 0) g.ref[k] = emitWide(slotKid[k], slotBase[k]); };
      both(n.nNodes >= kTaskNodes, [&] { run(0); run(1); }, [&] { run(2); run(3); });
    }
    // visiting decisions of the closest-hit walk for each of the 8 dirIsNeg octants, 3 bits each:
    // bit 0 = slots 2,3 before 0,1; bit 1 = slot 1 before 0; bit 2 = slot 3 before 2
    uint32_t lut = 0;
    for (uint32_t o = 0; o < 8; ++o) {
      auto dec = [&](int32_t ax) { return ((o >> (ax & 3)) ^ (uint32_t)(ax >> 2)) & 1u; };
      lut |= (dec(g.axisP) | (dec(g.axisA) << 1) | (dec(g.axisB) << 2)) << (3 * o);
    }
    g.orderLut = (int32_t)lut;
    out->wide[my] = g;
    // quantised while the node is in cache (a separate pass over the 128-byte nodes cost soup_10m 0.9 s)
    if (!quantiseNode(g, &out->wideQ[my])) quantOk.store(false, std::memory_order_relaxed);
    return my;
  }
  std::atomic<bool> quantOk{true};

  // GNode4 -> GNode4Q: an 8-bit grid per axis, origin just below the node's box, step 2^e.  Every inequality the
  // traversal kernel relies on is CHECKED here with the decode expression the device uses (binary64, exact).
  static bool quantiseNode(const GNode4& g, GNode4Q* q) {
    std::memset(q, 0, sizeof(*q));
    for (int k = 0; k < 4; ++k) q->ref[k] = g.ref[k];
    q->orderLut = (uint32_t)g.orderLut;
    const double guard = DRT_Q_GUARD;
    uint32_t scaleHi[3] = {0, 0, 0};
    uint8_t qb[3][4][2];
    for (int a = 0; a < 3; ++a) {
      double lo = std::numeric_limits<double>::infinity(), hi = -lo;
      for (int k = 0; k < 4; ++k) {
        if (g.ref[k] == DRT_REF_EMPTY) continue;
        lo = std::min(lo, (double)g.box[k][2 * a]);
        hi = std::max(hi, (double)g.box[k][2 * a + 1]);
      }
      if (!(lo <= hi) || !(std::fabs(lo) <= DRT_Q_COORD_MAX) || !(std::fabs(hi) <= DRT_Q_COORD_MAX)) return false;
      int e = DRT_Q_EXP_MIN;
      {
        double ext = std::max(hi - lo, (double)std::fabs((float)lo) * 1.2e-7);
        if (ext > 0.0) {  // a starting exponent a little below log2(ext / 255); the loop below raises it until everything fits
          int ex = 0;
          std::frexp(ext, &ex);  // ext = m * 2^ex, m in [0.5, 1)
          e = std::max(e, ex - 10);
        }
      }
      for (;; ++e) {
        if (e > DRT_Q_EXP_MAX) return false;
        const double s = std::ldexp(1.0, e);
        float O = (float)(lo - 2.0 * guard * s);
        while (!(lo - (double)O >= guard * s)) O = std::nextafterf(O, -std::numeric_limits<float>::infinity());
        if (!(std::fabs((double)O) <= DRT_Q_COORD_MAX)) return false;
        bool fits = true;
        for (int k = 0; k < 4 && fits; ++k) {
          if (g.ref[k] == DRT_REF_EMPTY) { qb[a][k][0] = 255; qb[a][k][1] = 0; continue; }
          const double blo = g.box[k][2 * a], bhi = g.box[k][2 * a + 1];
          double ql = std::floor((blo - (double)O) / s - guard), qh = std::ceil((bhi - (double)O) / s + guard);
          while (ql >= 0.0 && !((double)O + ql * s <= blo - guard * s)) ql -= 1.0;
          while (qh <= 255.0 && !((double)O + qh * s >= bhi + guard * s)) qh += 1.0;
          if (ql < 0.0 || qh > 255.0) { fits = false; break; }
          qb[a][k][0] = (uint8_t)ql;
          qb[a][k][1] = (uint8_t)qh;
        }
        if (!fits) continue;
        q->origin[a] = O;
        const float sp = std::ldexp(1.0f, e + 15);  // what the kernel multiplies by: 1 + q * 2^-15 is a float32 built by PRMT
        uint32_t bits;
        std::memcpy(&bits, &sp, 4);
        scaleHi[a] = bits >> 16;  // a power of two: the low mantissa half is zero
        break;
      }
      for (int j = 0; j < 2; ++j)
        q->q[a][j] = (uint32_t)qb[a][2 * j][0] | ((uint32_t)qb[a][2 * j][1] << 8) | ((uint32_t)qb[a][2 * j + 1][0] << 16) |
                     ((uint32_t)qb[a][2 * j + 1][1] << 24);
    }
    q->scaleXY = scaleHi[0] | (scaleHi[1] << 16);
    q->scaleZ = scaleHi[2];
    return true;
  }

};

}  // namespace

bool buildBvh(const std::vector<PrimBounds>& bounds, const std::vector<uint32_t>& order, int splitMethod,
              int maxPrimsInNode, BuiltBvh* out, std::string* err) {
  const bool timing = std::getenv("DRT_BUILD_TIMING") != nullptr;
  auto tPrev = std::chrono::steady_clock::now();
  *out = BuiltBvh();
  const size_t n = order.size();
  if (n == 0) { *err = "no primitives"; return false; }
  if (n >= (1u << 26)) { *err = "too many primitives for the 26-bit leaf offset"; return false; }
  int maxPrims = std::min(255, maxPrimsInNode);  // bvh_accel.dart:44
  std::vector<PrimRef> refs(n);
  {
    const unsigned nt = n < (1u << 16) ? 1u : std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        for (size_t i = n * t / nt; i < n * (t + 1) / nt; ++i) {
          const PrimBounds& b = bounds[order[i]];
          PrimRef& r = refs[i];
          std::memcpy(r.lo, b.bmin, 12);
          std::memcpy(r.hi, b.bmax, 12);
          // bbox.dart:68: (pMin * 0.5) + (pMax * 0.5), each Point operation rounds to float32
          for (int a = 0; a < 3; ++a) {
            float h0 = (float)((double)b.bmin[a] * 0.5), h1 = (float)((double)b.bmax[a] * 0.5);
            r.cen[a] = (float)((double)h0 + (double)h1);
          }
          r.id = order[i];
        }
      });
    for (auto& x : th) x.join();
  }
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[drt build] %-28s %.3f s\n", what, std::chrono::duration<double>(now - tPrev).count());
    tPrev = now;
  };
  lap("primitive references");
  Arena arena(2 * n);  // slot ranges by primitive count (TreeBuilder::build); unused slots are never touched
  if (!arena.pool.p) { *err = "out of host memory for the BVH build"; return false; }
  TreeBuilder tb(refs, splitMethod, maxPrims);
  int32_t root = tb.build(arena, 0, 0, (uint32_t)n, 0);
  lap("tree (SAH recursion)");

  Flattener fl{arena.pool, refs, out};
  const TNode& rt = arena.pool[root];
  out->nLeaves = rt.nLeaves;
  out->maxLeafPrims = rt.maxLeaf;
  out->maxDepth = rt.depthBelow;
  // two independent passes over the finished tree: reference numbering + reference primitive order + binary GPU layout (emitBinary)
  // on a second thread, wide layout + quantised nodes here
  std::thread binaryChain([&] {
    out->refNodes.resize(rt.nNodes);
    out->refOrdered.resize(n);
    out->leafPrimIds.resize(n);
    out->leafCounts.resize(n);
    out->nodes.resize(rt.nInterior);
    fl.emitBinary(root, 0, 0, 0);
    out->rootRef = rt.left < 0 ? fl.leafRefOf(root) : 0;
  });
  out->wide.resize(rt.left < 0 ? 0 : rt.nWide);
  out->wideQ.resize(out->wide.size());
  out->wideRootRef = fl.emitWide(root, 0);
  out->wideQOk = fl.quantOk.load();
  lap("wide layout + quantised nodes");
  binaryChain.join();
  lap("reference numbering + binary layout (second thread)");
  std::memcpy(out->rootMin, arena.pool[root].box.lo, 12);
  std::memcpy(out->rootMax, arena.pool[root].box.hi, 12);
  if (out->maxDepth >= 64) {
    // The reference walks the tree with a fixed 64-entry todo stack (bvh_accel.dart:120).
    *err = "BVH depth exceeds the reference's 64-entry traversal stack";
    return false;
  }
  return true;
}

}  // namespace drt
