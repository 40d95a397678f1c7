// BVH2 traversal for ptb200 (device): ordered (near child first), t-culled, per-lane stack, persistent warps.
//   closest hit : replaces get_intersection_candidates + check_hit
//                 (implementations/src/acceleration/mod.rs:199-224, 265-298). The reference keeps the minimum t > 0
//                 over every primitive in every leaf whose box the ray line crosses; an ordered traversal that only
//                 culls boxes entered beyond the current best t (less an error slack, see box_entry) returns the same
//                 minimum. Exact-t ties go to the lower ORIGINAL primitive id (the reference: first found in its BFS
//                 order — quirk Q2).
//   any hit     : replaces the blocker scan of check_hit_index (acceleration/mod.rs:226-263) and the sky visibility
//                 test of sample_lights (integrators/mis.rs:104-115): any primitive != exclude with 0 < t < tmax.
//
// Execution shape (the first ncu capture showed the naive one-ray-per-lane loop issue-bound at 11-15 active lanes of
// 32): every warp is persistent and keeps its 32 lanes busy —
//   * "while-while": a lane walks internal nodes until EVERY lane of the warp holds a postponed leaf, then all lanes
//     run the primitive test together (Aila & Laine 2009, speculative traversal);
//   * dynamic fetch: when fewer than kFetchThreshold lanes still have work, the warp leaves the traversal loop and
//     refills its idle lanes from the global ray queue (one warp-aggregated atomicAdd) instead of dragging a few long
//     rays along with 90 % of the lanes idle.
// Node fetches are four 16-byte loads of one 64-byte node that carries BOTH children's boxes.
#pragma once
#include "ptb_intersect.cuh"

namespace ptb {

constexpr int kStackDepth = 64;        // LBVH depth <= 30 Morton bits + 32 index tie-break bits; one push per level
constexpr int kFetchThreshold = 20;    // lanes; below this the warp refills from the queue
constexpr int kNodeBurst = 4;          // node steps per warp-level scheduling decision

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

// Per-lane traversal state. `cur`: internal node index, or a leaf reference (bit 31), or kNone when the stack ran dry.
// `leaf`: postponed leaf reference or kNone.
struct TravState {
  uint32_t cur, leaf;
  int sp;
  float best_t;       // closest hit so far (closest-hit) / tmax (any-hit)
  uint32_t best_ref;  // closest-hit: winning leaf ref; any-hit: kNone = unoccluded, 0 = occluded
  PTB_DEV bool done() const { return cur == kNone && leaf == kNone; }
};

PTB_DEV void trav_init(TravState& s, uint32_t n_prims, float tmax) {
  s.cur = n_prims ? 0u : kNone;
  s.leaf = kNone;
  s.sp = 0;
  s.best_t = tmax;
  s.best_ref = kNone;
}

// Stack entry = (node or leaf reference, cull key of its box) in one 8-byte local-memory word.
// Pops the next entry whose box can still hold a closer hit; a popped leaf is postponed when the slot is free and the
// pop continues, so on return `cur` is an internal node, a second leaf, or kNone (stack exhausted).
PTB_DEV void trav_pop(TravState& s, const uint2* stack) {
  for (;;) {
    if (s.sp == 0) { s.cur = kNone; return; }
    --s.sp;
    const uint2 e = stack[s.sp];
    if (__uint_as_float(e.y) <= s.best_t) {
      if ((e.x & PTB_LEAF_BIT) && s.leaf == kNone) { s.leaf = e.x; continue; }
      s.cur = e.x;
      return;
    }
  }
}

// One internal-node step of the lane: fetch the 64-byte node, test both child boxes, descend into the nearer hit child
// (deferring the other on the stack), and postpone the first leaf reached so the walk can continue. All pops of the step
// go through ONE loop (the profile showed two divergent pop sites running at 7 active lanes).
template <bool COUNT>
PTB_DEV void trav_node_step(const DevScene& sc, const Ray& ray, TravState& s, uint2* stack, uint32_t& n_nodes) {
  float4 n0, n1, n2;
  uint4 n3;
  load_node(sc.nodes, s.cur, n0, n1, n2, n3);
  if (COUNT) ++n_nodes;
  float tl, tr;
  const bool hl = box_entry(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, ray, s.best_t, tl);
  const bool hr = box_entry(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, ray, s.best_t, tr);
  bool want_pop = !(hl || hr);
  if (!want_pop) {
    const bool both = hl && hr;
    const bool right_first = both ? (tr < tl) : hr;
    if (both) {
      stack[s.sp] = make_uint2(right_first ? n3.x : n3.y, __float_as_uint(right_first ? tl : tr));
      ++s.sp;
    }
    s.cur = right_first ? n3.y : n3.x;
    if ((s.cur & PTB_LEAF_BIT) && s.leaf == kNone) {  // first leaf: postpone, keep walking
      s.leaf = s.cur;
      want_pop = true;
    }
  }
  if (want_pop) trav_pop(s, stack);
}

// One primitive step of the lane: test the postponed leaf; if the walk itself is parked on a leaf, that one is next.
template <bool ANYHIT, bool COUNT>
PTB_DEV void trav_prim_step(const DevScene& sc, const Ray& ray, TravState& s, uint2* stack, uint32_t exclude,
                            uint32_t& n_prims) {
  const uint32_t ref = s.leaf;
  s.leaf = kNone;
  if (ANYHIT) {
    if ((ref & kSlotMask) != exclude) {
      const float t = prim_t(sc, ray, ref);
      if (COUNT) ++n_prims;
      if (t > 0.0f && t < s.best_t) {  // blocker found: stop
        s.best_ref = 0u;
        s.cur = kNone;
        s.sp = 0;
        return;
      }
    }
  } else {
    const float t = prim_t(sc, ray, ref);
    if (COUNT) ++n_prims;
    if (t > 0.0f) {
      if (t < s.best_t) {
        s.best_t = t;
        s.best_ref = ref;
      } else if (t == s.best_t) {
        const uint32_t a = __ldg(sc.slot_prim + (ref & kSlotMask));
        const uint32_t b = __ldg(sc.slot_prim + (s.best_ref & kSlotMask));
        if (a < b) s.best_ref = ref;
      }
    }
  }
  if ((s.cur & PTB_LEAF_BIT) && s.cur != kNone) {  // the walk itself is parked on a leaf: it is next
    s.leaf = s.cur;
    trav_pop(s, stack);
  }
}

PTB_DEV TraceResult trav_result(const TravState& s) {
  TraceResult r;
  r.t = 0.0f;
  r.ref = kNone;
  if (s.best_ref != kNone) {
    r.t = s.best_t;
    r.ref = s.best_ref & ~PTB_LEAF_BIT;
  }
  return r;
}

// Persistent-warp driver, warp-synchronous: all 32 lanes run this loop in lock step (full-mask ballots only), so the
// SIMT efficiency is decided here and not by the compiler's reconvergence choices.
//   service : when fewer than kFetchThreshold lanes still have work, finished lanes are retired and idle lanes are
//             refilled from the global queue with one warp-aggregated atomicAdd;
//   phase   : each iteration runs EITHER a node step for the lanes parked on an internal node OR a primitive step for
//             the lanes holding a postponed leaf — whichever has more ready lanes.
// `fetch(i, ray, tmax, exclude)` loads work item i into the lane; `retire(fin, state)` is called by ALL 32 lanes together
// (fin = this lane just completed its item) so it may use warp-wide primitives.
template <bool ANYHIT, bool COUNT, class Fetch, class Retire>
PTB_DEV void persistent_trace(const DevScene& sc, uint32_t n, uint32_t* head, Fetch& fetch, Retire& retire,
                              uint32_t& cnt_nodes, uint32_t& cnt_prims, uint32_t& cnt_rays) {
  const uint32_t lane = threadIdx.x & 31u;
  uint2 stack[kStackDepth];
  TravState st;
  st.cur = st.leaf = kNone;
  st.sp = 0;
  st.best_t = 0.0f;
  st.best_ref = kNone;
  Ray ray;
  ray.o = ray.d = ray.dinv = ray.shear = mk(0.0f, 0.0f, 0.0f);
  ray.swap_xz = false;
  uint32_t exclude = kNone;
  bool has_ray = false, exhausted = false;
  for (;;) {
    // a lane with work is parked on an internal node, holds a postponed leaf, or both
    bool node_ready = has_ray && !(st.cur & PTB_LEAF_BIT);
    const bool leaf_ready = has_ray && st.leaf != kNone;
    const uint32_t m_node = __ballot_sync(0xffffffffu, node_ready);
    const uint32_t m_leaf = __ballot_sync(0xffffffffu, leaf_ready);
    if ((uint32_t)__popc(m_node | m_leaf) < (exhausted ? 1u : (uint32_t)sc.trace_fetch_threshold)) {
      // ---- service: retire finished items, refill idle lanes
      const bool fin = has_ray && !node_ready && !leaf_ready;
      retire(fin, st, ray);
      if (fin) has_ray = false;
      if (exhausted) {
        if (!__any_sync(0xffffffffu, has_ray)) break;
        continue;
      }
      const uint32_t idle = __ballot_sync(0xffffffffu, !has_ray);
      if (idle) {
        const uint32_t leader = __ffs(idle) - 1u, want = __popc(idle);
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(head, want);
        base = __shfl_sync(0xffffffffu, base, leader);
        const uint32_t mine = base + __popc(idle & ((1u << lane) - 1u));
        if (!has_ray && mine < n) {
          float tmax = __int_as_float(0x7f800000);
          fetch(mine, ray, tmax, exclude);
          trav_init(st, sc.n_prims, tmax);
          has_ray = true;
          if (COUNT) ++cnt_rays;
        }
        if (base + want >= n) exhausted = true;
      }
      continue;
    }
    if (__popc(m_node) >= __popc(m_leaf)) {
      // ---- node phase: a short burst of node steps amortises the warp-level bookkeeping above
#pragma unroll 1
      for (int burst = 0; burst < sc.trace_burst && node_ready; ++burst) {
        trav_node_step<COUNT>(sc, ray, st, stack, cnt_nodes);
        node_ready = !(st.cur & PTB_LEAF_BIT);
      }
    } else if (leaf_ready) {
      trav_prim_step<ANYHIT, COUNT>(sc, ray, st, stack, exclude, cnt_prims);
    }
  }
}

}  // namespace ptb
