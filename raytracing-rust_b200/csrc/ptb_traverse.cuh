// BVH2 traversal for ptb200 (device): ordered (near child first), t-culled, per-lane stack.
//   closest_hit : replaces get_intersection_candidates + check_hit
//                 (implementations/src/acceleration/mod.rs:199-224, 265-298). The reference keeps the minimum t > 0
//                 over every primitive in every leaf whose box the ray line crosses; an ordered traversal that only
//                 culls boxes entered beyond the current best t returns the same minimum. Exact-t ties go to the
//                 lower ORIGINAL primitive id (the reference: first found in its BFS order — quirk Q2).
//   occluded    : replaces the blocker scan of check_hit_index (acceleration/mod.rs:226-263) and the sky visibility
//                 test of sample_lights (integrators/mis.rs:104-115): any primitive != exclude with 0 < t < tmax.
// Node fetches are four 16-byte loads of one 64-byte node that carries BOTH children's boxes.
#pragma once
#include "ptb_intersect.cuh"

namespace ptb {

constexpr int kStackDepth = 64;  // LBVH depth <= 30 Morton bits + 32 index tie-break bits; one push per level

struct TraceResult {
  float t;       // 0 on miss (sky.rs:79-91)
  uint32_t ref;  // kNone on miss, else (kSphereBit?) | slot
};

PTB_DEV void load_node(const BvhNode* __restrict__ nodes, uint32_t idx, float4& n0, float4& n1, float4& n2, uint4& n3) {
  const float4* p = reinterpret_cast<const float4*>(nodes + idx);
  n0 = __ldg(p);
  n1 = __ldg(p + 1);
  n2 = __ldg(p + 2);
  n3 = __ldg(reinterpret_cast<const uint4*>(p + 3));
}

// COUNT: also report how many 64-byte nodes were fetched and how many primitives were tested (the V and T of the
// algorithmic-bytes-per-ray figure); compiled out otherwise.
template <bool COUNT>
PTB_DEV TraceResult closest_hit_t(const DevScene& sc, const Ray& ray, uint32_t& n_nodes, uint32_t& n_prims) {
  TraceResult res;
  res.t = 0.0f;
  res.ref = kNone;
  if (sc.n_prims == 0) return res;
  float best_t = __int_as_float(0x7f800000);
  uint32_t best_ref = kNone;
  uint32_t stack[kStackDepth];
  float stack_t[kStackDepth];
  int sp = 0;
  uint32_t cur = 0;
  for (;;) {
    if (cur & PTB_LEAF_BIT) {
      const float t = prim_t(sc, ray, cur);
      if (COUNT) ++n_prims;
      if (t > 0.0f) {
        if (t < best_t) {
          best_t = t;
          best_ref = cur;
        } else if (t == best_t) {
          const uint32_t a = __ldg(sc.slot_prim + (cur & kSlotMask));
          const uint32_t b = __ldg(sc.slot_prim + (best_ref & kSlotMask));
          if (a < b) best_ref = cur;
        }
      }
      bool popped = false;
      while (sp > 0) {
        --sp;
        if (stack_t[sp] <= best_t) { cur = stack[sp]; popped = true; break; }
      }
      if (!popped) break;
      continue;
    }
    float4 n0, n1, n2;
    uint4 n3;
    load_node(sc.nodes, cur, n0, n1, n2, n3);
    if (COUNT) ++n_nodes;
    float tl, tr;
    const bool hl = box_entry(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, ray, best_t, tl);
    const bool hr = box_entry(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, ray, best_t, tr);
    if (hl && hr) {
      uint32_t nearc = n3.x, farc = n3.y;
      float tfar = tr;
      if (tr < tl) { nearc = n3.y; farc = n3.x; tfar = tl; }
      stack[sp] = farc;
      stack_t[sp] = tfar;
      ++sp;
      cur = nearc;
    } else if (hl) {
      cur = n3.x;
    } else if (hr) {
      cur = n3.y;
    } else {
      bool popped = false;
      while (sp > 0) {
        --sp;
        if (stack_t[sp] <= best_t) { cur = stack[sp]; popped = true; break; }
      }
      if (!popped) break;
    }
  }
  if (best_ref != kNone) {
    res.t = best_t;
    res.ref = best_ref & ~PTB_LEAF_BIT;
  }
  return res;
}

PTB_DEV TraceResult closest_hit(const DevScene& sc, const Ray& ray) {
  uint32_t a = 0, b = 0;
  return closest_hit_t<false>(sc, ray, a, b);
}

// true when some primitive other than `exclude_slot` is hit with 0 < t < tmax
PTB_DEV bool occluded(const DevScene& sc, const Ray& ray, float tmax, uint32_t exclude_slot) {
  if (sc.n_prims == 0) return false;
  uint32_t stack[kStackDepth];
  int sp = 0;
  uint32_t cur = 0;
  for (;;) {
    if (cur & PTB_LEAF_BIT) {
      if ((cur & kSlotMask) != exclude_slot) {
        const float t = prim_t(sc, ray, cur);
        if (t > 0.0f && t < tmax) return true;
      }
      if (sp == 0) return false;
      cur = stack[--sp];
      continue;
    }
    float4 n0, n1, n2;
    uint4 n3;
    load_node(sc.nodes, cur, n0, n1, n2, n3);
    float tl, tr;
    const bool hl = box_entry(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, ray, tmax, tl);
    const bool hr = box_entry(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, ray, tmax, tr);
    if (hl && hr) {
      uint32_t nearc = n3.x, farc = n3.y;
      if (tr < tl) { nearc = n3.y; farc = n3.x; }
      stack[sp++] = farc;
      cur = nearc;
    } else if (hl) {
      cur = n3.x;
    } else if (hr) {
      cur = n3.y;
    } else {
      if (sp == 0) return false;
      cur = stack[--sp];
    }
  }
}

}  // namespace ptb
