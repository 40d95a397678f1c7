// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_math.hpp header).
// CPU statement of the device's PLOC hierarchy (Meister & Bittner 2018, "Parallel Locally-Ordered Clustering for Bounding
// Volume Hierarchy Construction"): an SAH-quality binary tree over the Morton-sorted primitives of lbvh_ref.hpp, the
// device-side replacement for the QUALITY of the reference's top-down SAH builder (implementations/src/acceleration/
// mod.rs:97-160, split.rs:78-187 — whose cost is surface area weighted, split.rs:161-163,176). Defined operation by
// operation so that the device tree can be compared bit for bit.
//
//   clusters = the primitives in Morton order, each with its box and a reference (PTB_LEAF_BIT | sorted position)
//   repeat until one cluster is left:
//     1. nearest neighbour: for cluster i, among j in [i-R, i+R] (j != i, inside the array), the j that minimises
//        key(i,j) = (half area of box_i U box_j, min(i,j), max(i,j)) lexicographically. The key is symmetric, so the
//        pair with the globally smallest key is mutual: every round merges at least one pair.
//        half area = dx*dy + dy*dz + dz*dx with d = max - min of the union (f32, this operation order, no fma).
//     2. merge: if nn[nn[i]] == i and i < nn[i], clusters i and nn[i] become ONE cluster at position i (box = union,
//        reference = a new internal node whose left child is cluster i, right child cluster nn[i]); position nn[i] is
//        vacated. Merges of a round are numbered in position order (exclusive prefix sum over the merge flags).
//     3. compaction: the surviving clusters keep their relative order.
//   node numbering: the k-th node created overall (k = 0, 1, ...) is node (n-2) - k, so the root (created last) is node 0
//   and children always have larger indices than their parent.
#pragma once
#include <cstdint>
#include <vector>

#include "lbvh_ref.hpp"

namespace ref {

struct PlocBox {
  float mn[3], mx[3];
};
static inline float ploc_half_area(const PlocBox& a, const PlocBox& b) {
  const float dx = fmax_(a.mx[0], b.mx[0]) - fmin_(a.mn[0], b.mn[0]);
  const float dy = fmax_(a.mx[1], b.mx[1]) - fmin_(a.mn[1], b.mn[1]);
  const float dz = fmax_(a.mx[2], b.mx[2]) - fmin_(a.mn[2], b.mn[2]);
  return (dx * dy + dy * dz) + dz * dx;
}

// Replaces l.nodes (the Karras hierarchy) by the PLOC hierarchy over the same sorted primitives. Returns the rounds taken.
static inline int ploc_rebuild(Lbvh& l, int radius) {
  const size_t n = l.prim_sorted.size();
  if (n < 2) return 0;  // 0 / 1 primitive: the LBVH's own single node stands
  std::vector<PlocBox> box(n), box2;
  std::vector<uint32_t> ref(n), ref2;
  for (size_t i = 0; i < n; ++i) {
    Vec3 a, b;
    l.prims[l.prim_sorted[i]].aabb(a, b);
    box[i] = PlocBox{{a.x, a.y, a.z}, {b.x, b.y, b.z}};
    ref[i] = PTB_LEAF_BIT | (uint32_t)i;
  }
  l.nodes.assign(n - 1, ptb_bvh_node{});
  size_t created = 0;
  int rounds = 0;
  std::vector<uint32_t> nn;
  while (box.size() > 1) {
    const int64_t c = (int64_t)box.size();
    nn.assign(c, 0);
    for (int64_t i = 0; i < c; ++i) {
      float best = 0.0f;
      int64_t bj = -1;
      const int64_t lo = i - radius < 0 ? 0 : i - radius, hi = i + radius >= c ? c - 1 : i + radius;
      for (int64_t j = lo; j <= hi; ++j) {
        if (j == i) continue;
        const float a = ploc_half_area(box[i], box[j]);
        // key (a, min(i,j), max(i,j)): for j < i the pair is (j, i), for j > i it is (i, j). Scanning j upwards, pairs with
        // j < i come first and have increasing min; pairs with j > i share min = i > every earlier min and have
        // increasing max: a strict '<' on the area keeps the lexicographically smallest key.
        if (bj < 0 || a < best) { best = a; bj = j; }
      }
      nn[i] = (uint32_t)bj;
    }
    box2.clear();
    ref2.clear();
    for (int64_t i = 0; i < c; ++i) {
      const int64_t j = nn[i];
      if ((int64_t)nn[j] == i) {
        if (i < j) {
          const uint32_t id = (uint32_t)(n - 2 - created);
          ++created;
          ptb_bvh_node& nd = l.nodes[id];
          for (int k = 0; k < 3; ++k) { nd.lmin[k] = box[i].mn[k]; nd.lmax[k] = box[i].mx[k]; nd.rmin[k] = box[j].mn[k]; nd.rmax[k] = box[j].mx[k]; }
          nd.left = ref[i];
          nd.right = ref[j];
          nd.parent = 0xFFFFFFFFu;
          if (!(ref[i] & PTB_LEAF_BIT)) l.nodes[ref[i]].parent = id;
          if (!(ref[j] & PTB_LEAF_BIT)) l.nodes[ref[j]].parent = id;
          PlocBox u;
          for (int k = 0; k < 3; ++k) { u.mn[k] = fmin_(box[i].mn[k], box[j].mn[k]); u.mx[k] = fmax_(box[i].mx[k], box[j].mx[k]); }
          box2.push_back(u);
          ref2.push_back(id);
        }  // else: the pair's upper position is vacated
      } else {
        box2.push_back(box[i]);
        ref2.push_back(ref[i]);
      }
    }
    box.swap(box2);
    ref.swap(ref2);
    ++rounds;
  }
  return rounds;
}

}  // namespace ref
