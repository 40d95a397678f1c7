// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_math.hpp header).
// CPU statement of the device LBVH (the reference has no Morton/LBVH code: its builder is the top-down SAH
// of acceleration/mod.rs:97-160, restated in ref_bvh.hpp). This file defines, operation by operation, the
// tree the device must reproduce BIT-EXACTLY (Morton keys, sorted order, child/parent links, boxes), and the
// ordered, t-culled stack traversal whose mean node/triangle counts define the algorithmic bytes per ray
// (B_ray = 32 + 16 + V*64 + T*48).
//
// Build (Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees, and k-d trees"):
//   1. per primitive: AABB (sphere.rs:175-181 / triangle.rs:285-307), centroid = 0.5*(min+max)
//      (acceleration/mod.rs:29-41);
//   2. scene bounds over the centroids; q_axis = (uint)clamp((c-cmin)/(cmax-cmin)*1024, 0, 1023)
//      (0 when the extent is 0); 30-bit Morton code, x in the most significant interleave position;
//   3. stable sort by code (ties keep primitive-index order);
//   4. hierarchy with delta(i,j) = clz(code_i ^ code_j), ties broken by 32 + clz(i ^ j);
//   5. bottom-up refit: box(node) = union of its children's boxes; each node stores BOTH children's boxes.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "ref_scene.hpp"

namespace ref {

static inline uint32_t expand_bits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
static inline uint32_t quantise10(Float c, Float cmin, Float ext) {
  Float n = ext > 0.0f ? (c - cmin) / ext : 0.0f;
  Float s = fmin_(fmax_(n * 1024.0f, 0.0f), 1023.0f);
  return (uint32_t)s;
}
static inline int clz32(uint32_t x) { return x == 0 ? 32 : __builtin_clz(x); }

struct Lbvh {
  std::vector<Prim> prims;          // original (loader) order
  std::vector<uint32_t> morton;     // sorted
  std::vector<uint32_t> prim_sorted;  // sorted position -> original primitive id
  std::vector<ptb_bvh_node> nodes;  // n-1 internal nodes (1 when n == 1), root = 0
  Vec3 cmin, cmax;
  int delta(int64_t i, int64_t j) const {
    int64_t n = (int64_t)morton.size();
    if (j < 0 || j >= n) return -1;
    uint32_t a = morton[i], b = morton[j];
    if (a == b) return 32 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz32(a ^ b);
  }

  void build(const std::vector<Prim>& in) {
    prims = in;
    const size_t n = prims.size();
    morton.assign(n, 0);
    prim_sorted.assign(n, 0);
    nodes.clear();
    if (n == 0) return;
    std::vector<Vec3> bmin(n), bmax(n), cen(n);
    for (size_t i = 0; i < n; ++i) {
      prims[i].aabb(bmin[i], bmax[i]);
      cen[i] = 0.5f * (bmin[i] + bmax[i]);
    }
    cmin = cen[0]; cmax = cen[0];
    for (size_t i = 1; i < n; ++i) { cmin = cmin.min_by_component(cen[i]); cmax = cmax.max_by_component(cen[i]); }
    Vec3 ext = cmax - cmin;
    std::vector<uint32_t> code(n);
    for (size_t i = 0; i < n; ++i) {
      // one CUBIC grid for all axes (cell = largest extent / 1024): on flat scenes a per-axis grid spends Morton bits
      // on the thin axis and splits nodes into overlapping halves (measured: 10 % more node fetches on C3)
      const Float e = fmax_(ext.x, fmax_(ext.y, ext.z));
      uint32_t qx = quantise10(cen[i].x, cmin.x, e);
      uint32_t qy = quantise10(cen[i].y, cmin.y, e);
      uint32_t qz = quantise10(cen[i].z, cmin.z, e);
      code[i] = (expand_bits10(qx) << 2) | (expand_bits10(qy) << 1) | expand_bits10(qz);
    }
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
    for (size_t i = 0; i < n; ++i) { morton[i] = code[order[i]]; prim_sorted[i] = order[i]; }

    auto set_box = [](float* mn, float* mx, const Vec3& a, const Vec3& b) {
      mn[0] = a.x; mn[1] = a.y; mn[2] = a.z; mx[0] = b.x; mx[1] = b.y; mx[2] = b.z;
    };
    if (n == 1) {  // single primitive: both child slots reference leaf 0
      ptb_bvh_node nd{};
      set_box(nd.lmin, nd.lmax, bmin[order[0]], bmax[order[0]]);
      set_box(nd.rmin, nd.rmax, bmin[order[0]], bmax[order[0]]);
      nd.left = PTB_LEAF_BIT | 0u;
      nd.right = PTB_LEAF_BIT | 0u;
      nd.parent = 0xFFFFFFFFu;
      nodes.push_back(nd);
      return;
    }
    nodes.assign(n - 1, ptb_bvh_node{});
    std::vector<uint32_t> leaf_parent(n, 0xFFFFFFFFu);
    nodes[0].parent = 0xFFFFFFFFu;
    for (int64_t i = 0; i < (int64_t)n - 1; ++i) {
      int d = (delta(i, i + 1) - delta(i, i - 1)) < 0 ? -1 : 1;
      int dmin = delta(i, i - d);
      int64_t lmax = 2;
      while (delta(i, i + lmax * d) > dmin) lmax *= 2;
      int64_t l = 0;
      for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (delta(i, i + (l + t) * d) > dmin) l += t;
      int64_t j = i + l * d;
      int dnode = delta(i, j);
      int64_t s = 0;
      for (int64_t t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
      }
      int64_t gam = i + s * d + (d < 0 ? -1 : 0);
      int64_t lo = i < j ? i : j, hi = i < j ? j : i;
      if (lo == gam) { nodes[i].left = PTB_LEAF_BIT | (uint32_t)gam; leaf_parent[gam] = (uint32_t)i; }
      else { nodes[i].left = (uint32_t)gam; nodes[gam].parent = (uint32_t)i; }
      if (hi == gam + 1) { nodes[i].right = PTB_LEAF_BIT | (uint32_t)(gam + 1); leaf_parent[gam + 1] = (uint32_t)i; }
      else { nodes[i].right = (uint32_t)(gam + 1); nodes[gam + 1].parent = (uint32_t)i; }
    }
    // bottom-up refit (post-order recursion is equivalent to the device's atomic-flag walk: min/max are exact)
    std::vector<Vec3> nmin(n - 1), nmax(n - 1);
    std::vector<int> state(n - 1, 0);
    std::vector<uint32_t> stack;
    stack.push_back(0);
    while (!stack.empty()) {
      uint32_t i = stack.back();
      ptb_bvh_node& nd = nodes[i];
      bool ready = true;
      if (!(nd.left & PTB_LEAF_BIT) && state[nd.left] == 0) { stack.push_back(nd.left); ready = false; }
      if (!(nd.right & PTB_LEAF_BIT) && state[nd.right] == 0) { stack.push_back(nd.right); ready = false; }
      if (!ready) continue;
      Vec3 lmn, lmx, rmn, rmx;
      if (nd.left & PTB_LEAF_BIT) { uint32_t p = order[nd.left & ~PTB_LEAF_BIT]; lmn = bmin[p]; lmx = bmax[p]; }
      else { lmn = nmin[nd.left]; lmx = nmax[nd.left]; }
      if (nd.right & PTB_LEAF_BIT) { uint32_t p = order[nd.right & ~PTB_LEAF_BIT]; rmn = bmin[p]; rmx = bmax[p]; }
      else { rmn = nmin[nd.right]; rmx = nmax[nd.right]; }
      set_box(nd.lmin, nd.lmax, lmn, lmx);
      set_box(nd.rmin, nd.rmax, rmn, rmx);
      nmin[i] = lmn.min_by_component(rmn);
      nmax[i] = lmx.max_by_component(rmx);
      state[i] = 1;
      stack.pop_back();
    }
  }

  // The 32-byte nodes the device's traversal kernels read (lbvh_build.cu: k_qframe / k_quantise_nodes), restated: grid over
  // the scene box (union of the root's child boxes), q_step = (float)(extent / 65535) moved up until the grid reaches the
  // far side, boxes snapped outwards in f64. Words per node: left box x, y, z, right box x, y, z (low half min, high
  // half max), left, right. frame = origin xyz, step xyz.
  static uint32_t q_floor(float v, float mn, float step) {
    if (!(step > 0.0f)) return 0u;
    double q = std::floor(((double)v - (double)mn) / (double)step);
    q = q < 0.0 ? 0.0 : (q > 65535.0 ? 65535.0 : q);
    while (q > 0.0 && (double)mn + q * (double)step > (double)v) q -= 1.0;
    return (uint32_t)q;
  }
  static uint32_t q_ceil(float v, float mn, float step) {
    if (!(step > 0.0f)) return 0u;
    double q = std::ceil(((double)v - (double)mn) / (double)step);
    q = q < 0.0 ? 0.0 : (q > 65535.0 ? 65535.0 : q);
    while (q < 65535.0 && (double)mn + q * (double)step < (double)v) q += 1.0;
    return (uint32_t)q;
  }
  void quantise(float frame[6], std::vector<uint32_t>& words) const {
    words.assign(nodes.size() * 8, 0u);
    for (int k = 0; k < 6; ++k) frame[k] = 0.0f;
    if (nodes.empty()) return;
    const ptb_bvh_node& r = nodes[0];
    for (int k = 0; k < 3; ++k) {
      const float mn = fmin_(r.lmin[k], r.rmin[k]), mx = fmax_(r.lmax[k], r.rmax[k]);
      const double ext = (double)mx - (double)mn;
      float step = (float)(ext / 65535.0);
      if (ext > 0.0)
        while ((double)mn + 65535.0 * (double)step < (double)mx) step = std::nextafterf(step, INF_F);
      else
        step = 0.0f;
      frame[k] = mn;
      frame[3 + k] = step;
    }
    for (size_t i = 0; i < nodes.size(); ++i) {
      const ptb_bvh_node& nd = nodes[i];
      uint32_t* w = &words[8 * i];
      for (int k = 0; k < 3; ++k) {
        w[k] = q_floor(nd.lmin[k], frame[k], frame[3 + k]) | (q_ceil(nd.lmax[k], frame[k], frame[3 + k]) << 16);
        w[3 + k] = q_floor(nd.rmin[k], frame[k], frame[3 + k]) | (q_ceil(nd.rmax[k], frame[k], frame[3 + k]) << 16);
      }
      w[6] = nd.left;
      w[7] = nd.right;
    }
  }

  // The device's slab test (ptb_intersect.cuh box_entry), restated operation for operation: aabb.rs:22-57 with one fma
  // per plane (plane * dinv - o * dinv, the product hoisted per ray), near / far plane chosen by the direction's sign,
  // near distances moved down / far distances moved up by e_i = 2 eps |o_i dinv_i|, far side widened by 1 + 4 gamma(3).
  // The box is culled only when its entry distance less 32 eps max(|entry|, |exit|) exceeds best t; `tkey` is that key.
  struct SlabRay {
    Vec3 dinv, c_lo, c_hi;
  };
  static SlabRay make_slab_ray(const Ray& ray) {
    SlabRay r;
    const Float H = 1.0e30f;  // clamp: plane * inf - o * inf would be NaN (see ptb_intersect.cuh)
    r.dinv = Vec3(fmax_(fmin_(ray.d_inverse.x, H), -H), fmax_(fmin_(ray.d_inverse.y, H), -H), fmax_(fmin_(ray.d_inverse.z, H), -H));
    const Vec3 od = ray.origin * r.dinv;
    const Vec3 e = (2.0f * F32_EPS) * od.abs();
    r.c_lo = -od - e;
    r.c_hi = e - od;
    return r;
  }
  static inline bool box_hit(const float* mn, const float* mx, const SlabRay& r, Float best_t, Float& tkey) {
    const Float k = 1.0f + 4.0f * gamma(3);
    const bool sx = r.dinv.x < 0.0f, sy = r.dinv.y < 0.0f, sz = r.dinv.z < 0.0f;
    const Float lox = std::fmaf(sx ? mx[0] : mn[0], r.dinv.x, r.c_lo.x), hix = std::fmaf(sx ? mn[0] : mx[0], r.dinv.x, r.c_hi.x);
    const Float loy = std::fmaf(sy ? mx[1] : mn[1], r.dinv.y, r.c_lo.y), hiy = std::fmaf(sy ? mn[1] : mx[1], r.dinv.y, r.c_hi.y);
    const Float loz = std::fmaf(sz ? mx[2] : mn[2], r.dinv.z, r.c_lo.z), hiz = std::fmaf(sz ? mn[2] : mx[2], r.dinv.z, r.c_hi.z);
    const Float tmin = fmax_(fmax_(lox, loy), loz);
    const Float hmin = fmin_(fmin_(hix, hiy), hiz);
    tkey = std::fmaf(-(32.0f * F32_EPS), fmax_(std::fabs(tmin), std::fabs(hmin)), tmin);
    return hmin * k > fmax_(tmin, 0.0f) && tkey <= best_t;
  }

  // Ordered (near child first), t-culled closest hit. Ties on t go to the lower ORIGINAL primitive id.
  // nodes_fetched / prims_tested feed the algorithmic-bytes-per-ray figure.
  bool closest_hit(const Ray& ray, Hit& best, uint32_t& best_prim, uint64_t* nodes_fetched, uint64_t* prims_tested) const {
    best_prim = PTB_MISS;
    if (nodes.empty()) return false;
    Float best_t = INF_F;
    const SlabRay slab = make_slab_ray(ray);
    uint32_t stack[128];
    Float stack_t[128];  // entry distance of the deferred child: re-checked against best_t when popped
    int sp = 0;
    uint32_t cur = 0;
    Hit h;
    auto pop = [&](uint32_t& out) -> bool {
      while (sp > 0) {
        --sp;
        if (stack_t[sp] <= best_t) { out = stack[sp]; return true; }
      }
      return false;
    };
    for (;;) {
      if (cur & PTB_LEAF_BIT) {
        uint32_t slot = cur & ~PTB_LEAF_BIT;
        uint32_t pid = prim_sorted[slot];
        if (prims_tested) ++*prims_tested;
        if (prims[pid].get_int(ray, h) && h.t > 0.0f) {
          if (h.t < best_t || (h.t == best_t && pid < best_prim)) { best_t = h.t; best = h; best_prim = pid; }
        }
        if (!pop(cur)) break;
        continue;
      }
      const ptb_bvh_node& nd = nodes[cur];
      if (nodes_fetched) ++*nodes_fetched;
      Float tl, tr;
      bool hl = box_hit(nd.lmin, nd.lmax, slab, best_t, tl);
      bool hr = box_hit(nd.rmin, nd.rmax, slab, best_t, tr);
      if (hl && hr) {
        uint32_t nearc = nd.left, farc = nd.right;
        Float tfar = tr;
        if (tr < tl) { nearc = nd.right; farc = nd.left; tfar = tl; }
        stack[sp] = farc;
        stack_t[sp] = tfar;
        ++sp;
        cur = nearc;
      } else if (hl) cur = nd.left;
      else if (hr) cur = nd.right;
      else if (!pop(cur)) break;
    }
    return best_prim != PTB_MISS;
  }
};

}  // namespace ref
