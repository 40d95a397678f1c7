// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_math.hpp header).
// CPU statement of the device's SAH builder (csrc/sah_build.cu): a top-down surface-area-heuristic hierarchy in the LBVH
// node format, the device-side replacement for the QUALITY of the reference's builder (implementations/src/acceleration/
// mod.rs:97-160 `build_bvh`, split.rs:78-187 `Split::Sah`: buckets along an axis of the centroid box, cost
// 0.125 + (n_l * area_l + n_r * area_r) / area, split.rs:161-163,176). Defined operation by operation so that the device
// tree can be compared bit for bit (f32, no fused multiply-adds on either side; min / max / integer counts are exact and
// order independent).
//
//   start: the primitives in the Morton order of lbvh_ref.hpp; root task = all positions, box = union of all boxes.
//   LARGE tasks (more than kSahSmall positions), one level of the tree at a time, tasks in position order:
//     bins: on each axis a with a finite nb / extent, primitive p falls into bin min((int)((c_a - mn_a) * (nb / ext_a)), nb-1)
//       of the TASK's box (c = 0.5 * (min + max), acceleration/mod.rs:29-41); a bin holds the union of its primitives' boxes
//       and their number.
//     split: over the axes in order and the nb-1 planes in order, cost = half_area(left) * n_l + half_area(right) * n_r for
//       planes with primitives on both sides; the FIRST minimum wins (the reference's constant and divisor do not change the
//       argmin; leaves hold one primitive, so there is no "do not split" branch). No valid plane (all centroids in one bin
//       on every axis): the task is halved in its current order and both halves keep its box.
//     partition: stable (the primitives of the left side keep their order, then those of the right side).
//   SMALL tasks (2..kSahSmall positions; on the device one warp each): exact sweep. Repeatedly, for every segment of two
//     or more positions: candidate (i, a) sends to the left every member j whose (c_a[j], j) is lexicographically
//     <= (c_a[i], i); candidates that send everything to the left are skipped; cost as above; the minimum by
//     (cost, i, a) wins; stable partition. No candidate with a comparable cost (NaN): halved in order.
//   depth bound: the traversal kernels' stacks hold 64 entries (ptb_traverse.cuh kStackDepth; the Karras tree is at most
//     62 levels deep). A split — in either regime — whose children could not both be finished by halving within
//     kSahMaxDepth levels (child depth + ceil(log2(child size)) > kSahMaxDepth) is replaced by the halving split, which
//     keeps that invariant; no leaf lies deeper than kSahMaxDepth.
//   node numbering: a node is named after the gap it splits at (the last position of its left child), which is unique in
//     a binary tree over a sequence — except that the root is node 0 and the node of gap 0 takes the root's gap.
#pragma once
#include <cstdint>
#include <vector>

#include "lbvh_ref.hpp"

namespace ref {

static const uint32_t kSahSmall = 32;
static const uint32_t kSahNone = 0xFFFFFFFFu;
static const uint32_t kSahMaxDepth = 60;
static inline uint32_t sah_ceil_log2(uint32_t n) { return n <= 1u ? 0u : 32u - (uint32_t)__builtin_clz(n - 1u); }
// can a subtree of n positions whose root sits at `depth` be finished by halving without exceeding max_depth?
static inline bool sah_fits(uint32_t depth, uint32_t n, uint32_t max_depth) { return depth + sah_ceil_log2(n) <= max_depth; }

struct SahBox {
  float mn[3], mx[3];
  void reset() { for (int k = 0; k < 3; ++k) { mn[k] = INF_F; mx[k] = -INF_F; } }
  void grow(const SahBox& o) { for (int k = 0; k < 3; ++k) { mn[k] = fmin_(mn[k], o.mn[k]); mx[k] = fmax_(mx[k], o.mx[k]); } }
  float half_area() const {
    const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    return (dx * dy + dy * dz) + dz * dx;
  }
};
struct SahTask {
  uint32_t lo, hi;  // positions [lo, hi)
  uint32_t parent, side;
  uint32_t depth;   // of the task's node (root: 0)
  SahBox box;
};
struct SahStats {
  uint32_t levels = 0, max_tasks = 0, small_tasks = 0, fallbacks = 0, max_depth = 0, depth_limited = 0;
};

struct SahBuilder {
  Lbvh& l;
  int nb;
  std::vector<SahBox> pbox;     // by ORIGINAL primitive id
  std::vector<uint32_t> order;  // position -> original primitive
  uint32_t root_gap = kSahNone;
  SahStats stats;
  uint32_t max_depth;  // kSahMaxDepth, or lower (test hook; never below what halving alone needs)
  SahBuilder(Lbvh& l_, int nbins, uint32_t max_depth_) : l(l_), nb(nbins), max_depth(max_depth_) {}

  uint32_t node_id(uint32_t gap) {
    if (root_gap == kSahNone) { root_gap = gap; return 0u; }
    return gap == root_gap ? 0u : (gap == 0u ? root_gap : gap);
  }
  float centroid(uint32_t p, int a) const { return 0.5f * (pbox[p].mn[a] + pbox[p].mx[a]); }
  void link(uint32_t parent, uint32_t side, uint32_t ref) {
    if (parent == kSahNone) return;
    (side ? l.nodes[parent].right : l.nodes[parent].left) = ref;
  }
  // a node over [lo, hi) that sends cl positions to the left: names it, links it, returns its id
  uint32_t make_node(uint32_t lo, uint32_t cl, uint32_t parent, uint32_t side) {
    const uint32_t id = node_id(lo + cl - 1u);
    l.nodes[id].parent = parent;
    link(parent, side, id);
    return id;
  }
  static bool axis_scale(const SahBox& b, int a, int nb, float& scale) {
    const float ext = b.mx[a] - b.mn[a];
    if (!(ext > 0.0f)) return false;
    scale = (float)nb / ext;
    return scale < 1.0e30f;
  }
  static int bin_of(float c, float mn, float scale, int nb) { return (int)fmin_((c - mn) * scale, (float)(nb - 1)); }

  void run_large(std::vector<SahTask>& large, std::vector<SahTask>& small) {
    std::vector<SahTask> next;
    std::vector<uint32_t> tmp;
    std::vector<SahBox> bins(3 * nb);
    std::vector<uint32_t> cnt(3 * nb);
    std::vector<float> ra(nb);
    std::vector<uint32_t> rc(nb);
    while (!large.empty()) {
      ++stats.levels;
      if (large.size() > stats.max_tasks) stats.max_tasks = (uint32_t)large.size();
      next.clear();
      for (const SahTask& t : large) {
        const uint32_t len = t.hi - t.lo;
        float scale[3];
        bool ok[3];
        for (int a = 0; a < 3; ++a) ok[a] = axis_scale(t.box, a, nb, scale[a]);
        for (int i = 0; i < 3 * nb; ++i) { bins[i].reset(); cnt[i] = 0; }
        for (uint32_t k = t.lo; k < t.hi; ++k) {
          const uint32_t p = order[k];
          for (int a = 0; a < 3; ++a) {
            const int bi = ok[a] ? bin_of(centroid(p, a), t.box.mn[a], scale[a], nb) : 0;
            bins[a * nb + bi].grow(pbox[p]);
            cnt[a * nb + bi]++;
          }
        }
        float best = INF_F;
        int best_axis = -1, best_bin = 0;
        uint32_t best_cl = 0;
        for (int a = 0; a < 3; ++a) {
          if (!ok[a]) continue;
          SahBox b;
          b.reset();
          uint32_t c = 0;
          for (int i = nb - 1; i > 0; --i) { b.grow(bins[a * nb + i]); c += cnt[a * nb + i]; ra[i] = c ? b.half_area() : 0.0f; rc[i] = c; }
          b.reset();
          c = 0;
          for (int i = 0; i < nb - 1; ++i) {
            b.grow(bins[a * nb + i]);
            c += cnt[a * nb + i];
            if (c == 0 || rc[i + 1] == 0) continue;
            const float cost = b.half_area() * (float)c + ra[i + 1] * (float)rc[i + 1];
            if (cost < best) { best = cost; best_axis = a; best_bin = i; best_cl = c; }
          }
        }
        if (best_axis >= 0 && !(sah_fits(t.depth + 1u, best_cl, max_depth) && sah_fits(t.depth + 1u, len - best_cl, max_depth))) {
          best_axis = -1;
          ++stats.depth_limited;
        }
        SahTask L, R;
        uint32_t cl;
        if (best_axis < 0) {
          ++stats.fallbacks;
          cl = len >> 1;
          L.box = t.box;
          R.box = t.box;
        } else {
          cl = best_cl;
          L.box.reset();
          R.box.reset();
          for (int i = 0; i < nb; ++i) (i <= best_bin ? L : R).box.grow(bins[best_axis * nb + i]);
          tmp.clear();
          for (int side = 0; side < 2; ++side)
            for (uint32_t k = t.lo; k < t.hi; ++k) {
              const uint32_t p = order[k];
              const bool left = bin_of(centroid(p, best_axis), t.box.mn[best_axis], scale[best_axis], nb) <= best_bin;
              if (left == (side == 0)) tmp.push_back(p);
            }
          for (uint32_t k = 0; k < len; ++k) order[t.lo + k] = tmp[k];
        }
        const uint32_t id = make_node(t.lo, cl, t.parent, t.side);
        L.lo = t.lo; L.hi = t.lo + cl; L.parent = id; L.side = 0;
        R.lo = t.lo + cl; R.hi = t.hi; R.parent = id; R.side = 1;
        L.depth = R.depth = t.depth + 1u;
        for (const SahTask* c : {&L, &R}) {
          const uint32_t cn = c->hi - c->lo;
          if (cn == 1) { link(id, c->side, PTB_LEAF_BIT | c->lo); if (c->depth > stats.max_depth) stats.max_depth = c->depth; }
          else if (cn <= kSahSmall) small.push_back(*c);
          else next.push_back(*c);
        }
      }
      large.swap(next);
    }
  }

  void run_small(const SahTask& t) {
    ++stats.small_tasks;
    // segments of the task, processed until every one is a single position; a segment's split depends on its members only
    struct Seg { uint32_t lo, hi, parent, side, depth; };
    std::vector<Seg> segs{Seg{t.lo, t.hi, t.parent, t.side, t.depth}}, nxt;
    std::vector<uint32_t> tmp;
    while (!segs.empty()) {
      nxt.clear();
      for (const Seg& s : segs) {
        const uint32_t len = s.hi - s.lo;
        float best = INF_F;
        int best_a = -1;
        uint32_t best_i = 0, best_cl = 0;
        for (uint32_t i = s.lo; i < s.hi; ++i)
          for (int a = 0; a < 3; ++a) {
            const float ci = centroid(order[i], a);
            SahBox L, R;
            L.reset();
            R.reset();
            uint32_t cl = 0;
            for (uint32_t j = s.lo; j < s.hi; ++j) {
              const float cj = centroid(order[j], a);
              const bool left = cj < ci || (cj == ci && j <= i);
              if (left) { L.grow(pbox[order[j]]); ++cl; } else R.grow(pbox[order[j]]);
            }
            if (cl == len) continue;
            const float cost = L.half_area() * (float)cl + R.half_area() * (float)(len - cl);
            if (cost < best) { best = cost; best_a = a; best_i = i; best_cl = cl; }
          }
        if (best_a >= 0 && !(sah_fits(s.depth + 1u, best_cl, max_depth) && sah_fits(s.depth + 1u, len - best_cl, max_depth))) {
          best_a = -1;
          ++stats.depth_limited;
        }
        uint32_t cl;
        if (best_a < 0) {
          ++stats.fallbacks;
          cl = len >> 1;
        } else {
          cl = best_cl;
          const float ci = centroid(order[best_i], best_a);
          tmp.clear();
          for (int side = 0; side < 2; ++side)
            for (uint32_t j = s.lo; j < s.hi; ++j) {
              const float cj = centroid(order[j], best_a);
              const bool left = cj < ci || (cj == ci && j <= best_i);
              if (left == (side == 0)) tmp.push_back(order[j]);
            }
          for (uint32_t k = 0; k < len; ++k) order[s.lo + k] = tmp[k];
        }
        const uint32_t id = make_node(s.lo, cl, s.parent, s.side);
        if (s.depth + 1u > stats.max_depth) stats.max_depth = s.depth + 1u;
        if (cl == 1) link(id, 0, PTB_LEAF_BIT | s.lo);
        else nxt.push_back(Seg{s.lo, s.lo + cl, id, 0, s.depth + 1u});
        if (len - cl == 1) link(id, 1, PTB_LEAF_BIT | (s.hi - 1u));
        else nxt.push_back(Seg{s.lo + cl, s.hi, id, 1, s.depth + 1u});
      }
      segs.swap(nxt);
    }
  }

  SahBox refit(uint32_t ref) {
    if (ref & PTB_LEAF_BIT) return pbox[order[ref & ~PTB_LEAF_BIT]];
    ptb_bvh_node& nd = l.nodes[ref];
    const SahBox a = refit(nd.left), b = refit(nd.right);
    for (int k = 0; k < 3; ++k) { nd.lmin[k] = a.mn[k]; nd.lmax[k] = a.mx[k]; nd.rmin[k] = b.mn[k]; nd.rmax[k] = b.mx[k]; }
    SahBox u = a;
    u.grow(b);
    return u;
  }

  void run() {
    const size_t n = l.prim_sorted.size();
    if (n < 2) return;  // 0 / 1 primitive: the LBVH's own single node stands
    if (max_depth == 0u || max_depth > kSahMaxDepth) max_depth = kSahMaxDepth;
    if (max_depth < sah_ceil_log2((uint32_t)n)) max_depth = sah_ceil_log2((uint32_t)n);
    pbox.resize(n);
    order = l.prim_sorted;
    SahTask root;
    root.lo = 0; root.hi = (uint32_t)n; root.parent = kSahNone; root.side = 0; root.depth = 0;
    root.box.reset();
    for (size_t i = 0; i < n; ++i) {
      Vec3 a, b;
      l.prims[i].aabb(a, b);
      pbox[i] = SahBox{{a.x, a.y, a.z}, {b.x, b.y, b.z}};
      root.box.grow(pbox[i]);
    }
    l.nodes.assign(n - 1, ptb_bvh_node{});
    std::vector<SahTask> large, small;
    (n > kSahSmall ? large : small).push_back(root);
    run_large(large, small);
    for (const SahTask& t : small) run_small(t);
    l.prim_sorted = order;
    refit(0);
  }
};

static inline SahStats sah_rebuild(Lbvh& l, int nbins, uint32_t max_depth = kSahMaxDepth) {
  SahBuilder b(l, nbins, max_depth);
  b.run();
  return b.stats;
}

}  // namespace ref
