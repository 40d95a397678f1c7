// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_math.hpp header).
// The reference's acceleration structure: top-down binary BVH (SAH 12-bucket / Middle / EqualCounts),
// BFS un-ordered, un-culled candidate collection, test-all closest hit.
// Follows implementations/src/acceleration/{mod.rs,aabb.rs,split.rs} and utility/mod.rs:119-134.
//
// Parity status: sort_by_indices is pinned by the reference's KAT (utility/mod.rs:141-149). BVH topology
// and closest-hit results are NOT pinned by any reference test ("parity unpinned" in the reference); they
// are cross-checked here against a brute-force all-primitives scan and analytic ray-sphere answers.
#pragma once
#include <algorithm>
#include <deque>
#include <utility>
#include <vector>

#include "ref_scene.hpp"

namespace ref {

// acceleration/aabb.rs:9-87
struct AABB {
  Vec3 min, max;
  bool valid = false;
  // aabb.rs:22-57
  bool does_int(const Ray& ray) const {
    const Float k = 1.0f + 2.0f * gamma(3);
    Float t1 = (min.x - ray.origin.x) * ray.d_inverse.x;
    Float t2 = (max.x - ray.origin.x) * ray.d_inverse.x;
    if (t1 > t2) std::swap(t1, t2);
    t2 *= k;
    Float tmin = fmin_(t1, t2);
    Float tmax = fmax_(t1, t2);

    t1 = (min.y - ray.origin.y) * ray.d_inverse.y;
    t2 = (max.y - ray.origin.y) * ray.d_inverse.y;
    if (t1 > t2) std::swap(t1, t2);
    t2 *= k;
    tmin = fmax_(tmin, fmin_(t1, t2));
    tmax = fmin_(tmax, fmax_(t1, t2));

    t1 = (min.z - ray.origin.z) * ray.d_inverse.z;
    t2 = (max.z - ray.origin.z) * ray.d_inverse.z;
    if (t1 > t2) std::swap(t1, t2);
    t2 *= k;
    tmin = fmax_(tmin, fmin_(t1, t2));
    tmax = fmin_(tmax, fmax_(t1, t2));

    return tmax > fmax_(tmin, 0.0f);
  }
  void merge(const Vec3& mn, const Vec3& mx) {  // aabb.rs:59-67
    if (valid) {
      min = min.min_by_component(mn);
      max = max.max_by_component(mx);
    } else {
      min = mn; max = mx; valid = true;
    }
  }
  void merge(const AABB& o) { if (o.valid) merge(o.min, o.max); }
  void extend_contains(const Vec3& p) { merge(p, p); }  // aabb.rs:69-77
  Vec3 get_extent() const { return max - min; }
  Float surface_area() const {  // aabb.rs:83-86
    Vec3 e = get_extent();
    return 2.0f * (e.x * e.y + e.x * e.z + e.y * e.z);
  }
};

enum SplitType { SPLIT_SAH = 0, SPLIT_MIDDLE = 1, SPLIT_EQUAL_COUNTS = 2 };

// acceleration/mod.rs:21-41
struct PrimitiveInfo {
  size_t index;
  Vec3 min, max, center;
};

static inline Float axis_value(int axis, const Vec3& v) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }
// primitives/mod.rs:52-60
static inline int get_max_axis(const Vec3& v) {
  if (v.x > v.y && v.x > v.z) return 0;
  if (v.y > v.z) return 1;
  return 2;
}

// utility/mod.rs:119-134 (cycle-following in-place permutation: new[i] = old[indices[i]])
template <class T>
static inline void sort_by_indices(std::vector<T>& vec, std::vector<size_t> indices) {
  for (size_t index = 0; index < vec.size(); ++index) {
    if (indices[index] != index) {
      size_t current_index = index;
      for (;;) {
        size_t target_index = indices[current_index];
        indices[current_index] = current_index;
        if (indices[target_index] == target_index) break;
        std::swap(vec[current_index], vec[target_index]);
        current_index = target_index;
      }
    }
  }
}

// split.rs:9-32 (`partition!` macro)
template <class Pred>
static inline size_t hoare_partition(PrimitiveInfo* a, size_t len, Pred pred) {
  size_t left = 0, right = len - 1;
  for (;;) {
    while (left < len && pred(a[left])) ++left;
    while (right > 0 && !pred(a[right])) --right;
    if (left >= right) return left;
    std::swap(a[left], a[right]);
  }
}

// split.rs:201-210 (Rust's sort_by is a stable merge sort)
static inline size_t split_equal(int axis, PrimitiveInfo* a, size_t len) {
  std::stable_sort(a, a + len, [axis](const PrimitiveInfo& l, const PrimitiveInfo& r) {
    return axis_value(axis, l.center) < axis_value(axis, r.center);
  });
  return len / 2;
}

// split.rs:189-199
static inline size_t calculate_b(int axis, const PrimitiveInfo& info, Float min, Float extent) {
  const size_t NUM_BUCKETS = 12;
  Float absolute_value = axis_value(axis, info.center);
  size_t b = sat_usize((Float)NUM_BUCKETS * (absolute_value - min) / extent);
  if (b == NUM_BUCKETS) b -= 1;
  return b;
}

// split.rs:78-187
static inline size_t split(SplitType type, const AABB& bounds, const AABB& center_bounds, int axis, PrimitiveInfo* a,
                           size_t len) {
  const size_t NUM_BUCKETS = 12, MAX_IN_NODE = 255;
  switch (type) {
    case SPLIT_MIDDLE: {
      Float point_mid = 0.5f * (axis_value(axis, center_bounds.min) + axis_value(axis, center_bounds.max));
      size_t mid = hoare_partition(a, len, [&](const PrimitiveInfo& p) { return axis_value(axis, p.center) < point_mid; });
      if (mid == 0 || mid == len - 1) {
        std::stable_sort(a, a + len, [axis](const PrimitiveInfo& l, const PrimitiveInfo& r) {
          return axis_value(axis, l.center) < axis_value(axis, r.center);
        });
      }
      return mid;
    }
    case SPLIT_EQUAL_COUNTS:
      return split_equal(axis, a, len);
    case SPLIT_SAH:
    default: {
      if (len <= 4) return split_equal(axis, a, len);
      struct Bucket { uint32_t count = 0; AABB bounds; };
      Bucket buckets[NUM_BUCKETS];
      Float max_val = axis_value(axis, center_bounds.max);
      Float min_val = axis_value(axis, center_bounds.min);
      Float centroid_extent = max_val - min_val;
      for (size_t i = 0; i < len; ++i) {
        size_t b = calculate_b(axis, a[i], min_val, centroid_extent);
        if (b >= NUM_BUCKETS) b = NUM_BUCKETS - 1;  // Rust would panic (index out of bounds); unreachable for finite input
        buckets[b].count += 1;
        buckets[b].bounds.merge(a[i].min, a[i].max);
      }
      Float costs[NUM_BUCKETS - 1];
      for (size_t i = 0; i < NUM_BUCKETS - 1; ++i) {
        AABB bl, br;
        uint32_t cl = 0, cr = 0;
        for (size_t j = 0; j <= i; ++j)
          if (buckets[j].bounds.valid) { bl.merge(buckets[j].bounds); cl += buckets[j].count; }
        for (size_t j = i + 1; j < NUM_BUCKETS; ++j)
          if (buckets[j].bounds.valid) { br.merge(buckets[j].bounds); cr += buckets[j].count; }
        Float left_sa = bl.valid ? bl.surface_area() : 0.0f;
        Float right_sa = br.valid ? br.surface_area() : 0.0f;
        costs[i] = 0.125f + ((Float)cl * left_sa + (Float)cr * right_sa) / bounds.surface_area();
      }
      Float min_cost = costs[0];
      size_t min_cost_index = 0;
      for (size_t i = 1; i < NUM_BUCKETS - 1; ++i)
        if (costs[i] < min_cost) { min_cost = costs[i]; min_cost_index = i; }
      if (len > MAX_IN_NODE || min_cost < (Float)len) {
        return hoare_partition(a, len, [&](const PrimitiveInfo& p) {
          return calculate_b(axis, p, min_val, centroid_extent) <= min_cost_index;
        });
      }
      return 0;
    }
  }
}

// acceleration/mod.rs:330-361
struct Node {
  AABB bounds;
  bool has_children = false;
  size_t children[2] = {0, 0};
  size_t primitive_offset = 0, number_primitives = 0;
};

struct Bvh {
  SplitType split_type = SPLIT_SAH;
  std::vector<Node> nodes;
  std::vector<Prim> primitives;  // in BVH order after build
  std::vector<size_t> lights;    // indices into `primitives` (BVH order)
  Sky sky;
  size_t max_depth_seen = 0;

  // acceleration/mod.rs:58-93
  void build(std::vector<Prim> prims, SplitType st) {
    split_type = st;
    nodes.clear();
    lights.clear();
    std::vector<PrimitiveInfo> infos(prims.size());
    for (size_t i = 0; i < prims.size(); ++i) {
      Vec3 mn, mx;
      prims[i].aabb(mn, mx);
      infos[i].index = i;
      infos[i].min = mn;
      infos[i].max = mx;
      infos[i].center = 0.5f * (mn + mx);
    }
    if (!prims.empty()) build_bvh(0, infos.data(), infos.size(), 1);
    std::vector<size_t> idx(infos.size());
    for (size_t i = 0; i < infos.size(); ++i) idx[i] = infos[i].index;
    sort_by_indices(prims, idx);
    for (size_t i = 0; i < prims.size(); ++i)
      if (prims[i].material->is_light()) lights.push_back(i);
    primitives.swap(prims);
  }

  // acceleration/mod.rs:97-160
  size_t build_bvh(size_t offset, PrimitiveInfo* infos, size_t n, size_t depth) {
    if (depth > max_depth_seen) max_depth_seen = depth;
    AABB bounds;
    for (size_t i = 0; i < n; ++i) bounds.merge(infos[i].min, infos[i].max);
    size_t node_index = nodes.size();
    Node nd;
    nd.bounds = bounds;
    nd.primitive_offset = offset;
    nd.number_primitives = n;
    nodes.push_back(nd);
    bool has_children = false;
    size_t c0 = 0, c1 = 0;
    if (n != 1) {
      AABB center_bounds;
      for (size_t i = 0; i < n; ++i) center_bounds.extend_contains(infos[i].center);
      int axis = get_max_axis(center_bounds.get_extent());
      if (std::fabs(axis_value(axis, center_bounds.min) - axis_value(axis, center_bounds.max)) < 100.0f * F32_EPS) {
        // leaf holding all n primitives
      } else {
        size_t mid = split(split_type, bounds, center_bounds, axis, infos, n);
        if (mid != 0) {
          c0 = build_bvh(offset, infos, mid, depth + 1);
          c1 = build_bvh(offset + mid, infos + mid, n - mid, depth + 1);
          has_children = true;
        }
      }
    }
    if (has_children) {
      nodes[node_index].has_children = true;
      nodes[node_index].children[0] = c0;
      nodes[node_index].children[1] = c1;
    }
    return node_index;
  }

  // acceleration/mod.rs:199-224 — BFS, no ordering, no t-max culling
  void get_intersection_candidates(const Ray& ray, std::vector<std::pair<size_t, size_t>>& offset_len,
                                   uint64_t* nodes_visited = nullptr) const {
    offset_len.clear();
    if (nodes.empty()) return;
    std::deque<size_t> node_stack;
    node_stack.push_back(0);
    while (!node_stack.empty()) {
      size_t index = node_stack.front();
      node_stack.pop_front();
      const Node& node = nodes[index];
      if (nodes_visited) ++*nodes_visited;
      if (!node.bounds.does_int(ray)) continue;
      if (node.has_children) {
        node_stack.push_back(node.children[0]);
        node_stack.push_back(node.children[1]);
      } else {
        offset_len.push_back(std::make_pair(node.primitive_offset, node.number_primitives));
      }
    }
  }

  static const size_t MISS = (size_t)-1;

  // acceleration/mod.rs:265-298; miss -> sky.get_si (sky.rs:79-91): zero Hit + sky material, index usize::MAX
  size_t check_hit(const Ray& ray, Hit& hit, const Material*& mat, uint64_t* nodes_visited = nullptr,
                   uint64_t* prims_tested = nullptr) const {
    thread_local std::vector<std::pair<size_t, size_t>> offset_lens;
    get_intersection_candidates(ray, offset_lens, nodes_visited);
    bool have = false;
    size_t best = MISS;
    Hit cur;
    for (const auto& ol : offset_lens) {
      for (size_t index = ol.first; index < ol.first + ol.second; ++index) {
        if (prims_tested) ++*prims_tested;
        if (primitives[index].get_int(ray, cur)) {
          if (cur.t > 0.0f) {
            if (have) {
              if (cur.t < hit.t) { hit = cur; best = index; }
              continue;
            }
            hit = cur;
            best = index;
            have = true;
          }
        }
      }
    }
    if (!have) {
      hit = Hit();
      mat = &sky.mat;
      return MISS;
    }
    mat = primitives[best].material;
    return best;
  }

  // acceleration/mod.rs:226-263
  bool check_hit_index(const Ray& ray, size_t index, Hit& hit) const {
    thread_local std::vector<std::pair<size_t, size_t>> offset_lens;
    get_intersection_candidates(ray, offset_lens);
    Hit light_hit;
    if (!primitives[index].get_int(ray, light_hit)) return false;
    if (!(light_hit.t > 0.0f)) return false;
    Float light_t = light_hit.t;
    Hit cur;
    for (const auto& ol : offset_lens) {
      for (size_t ci = ol.first; ci < ol.first + ol.second; ++ci) {
        if (ci == index) continue;
        if (primitives[ci].get_int(ray, cur)) {
          if (cur.t > 0.0f && cur.t < light_t) return false;
        }
      }
    }
    hit = light_hit;
    return true;
  }

  // acceleration/mod.rs:299-318
  Float get_pdf_from_index(const Hit& last_hit, const Hit& light_hit, const Vec3& sampled_dir, size_t index) const {
    bool sky_samplable = sky.can_sample();
    Float divisor = (Float)(sky_samplable ? lights.size() + 1 : lights.size());
    if (index == MISS) return sky.pdf(sampled_dir) / divisor;
    return primitives[index].scattering_pdf(last_hit.point, sampled_dir, light_hit) / divisor;
  }
  bool is_samplable(size_t index) const {  // `bvh.get_samplable().contains(&index)` (mis.rs:58)
    for (size_t l : lights)
      if (l == index) return true;
    return false;
  }
};

}  // namespace ref
