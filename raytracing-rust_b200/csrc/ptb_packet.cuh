// Packet traversal of the binary LBVH (device): the 32 rays of a warp walk the tree TOGETHER, with one warp-uniform stack.
// For the camera rays of the window wavefront, whose 32 consecutive work items are samples of one pixel (or of the pixels
// of one 8 x 4 tile): they visit the same nodes down to the last levels.
//
// Why. The persistent phase machine of ptb_traverse.cuh is built for incoherent rays: every lane keeps its own stack in
// local memory and the warp alternates node and primitive phases. On the camera launch of C3 that machinery is most of
// the cost (ncu, round 2: 24.8 of 32 lanes per instruction, ALU pipe 72 %, issue 82 %, and 62 % of the launch's L1
// tag-stage wavefronts are per-lane stack traffic) although all 32 lanes want the same thing.
//
// What. Same per-ray result as check_hit (implementations/src/acceleration/mod.rs:265-298), bit for bit: a lane takes part
// in a subtree only if ITS OWN slab test (box_entry, against its own best hit) passed at every level above — exactly the
// nodes an ordered walk of that ray alone may visit, in a different order, and the minimum over the candidates (ties to the
// lower original primitive id, Q2) does not depend on the order.
//
// How. `cur` (node or leaf reference) and `mask` (the lanes taking part in it) are warp-uniform. A node step: ONE node
// fetch for the warp (every lane reads the same address: a broadcast), two slab tests per lane, two ballots; the child
// more lanes enter first is walked first, the other is pushed with its lane mask and the smallest entry distance among
// its lanes. The stack lives in REGISTERS: entry k is held by lane k & 31 (two register sets: 64 entries, the depth bound
// of the LBVH), a push is a predicated move, a pop three shuffles — no memory traffic, no divergence. An entry is skipped
// when its smallest entry distance lies behind every lane's best hit.
#pragma once
#include "ptb_intersect.cuh"

namespace ptb {

struct PacketStack {  // entry k: lane k & 31, register set k >> 5
  uint32_t ref0, ref1, key0, key1, mask0, mask1;
};
PTB_DEV void packet_push(PacketStack& s, int& sp, uint32_t lane, uint32_t ref, uint32_t key, uint32_t mask) {
  if (lane == ((uint32_t)sp & 31u)) {
    if (sp < 32) { s.ref0 = ref; s.key0 = key; s.mask0 = mask; }
    else { s.ref1 = ref; s.key1 = key; s.mask1 = mask; }
  }
  ++sp;
}
// order-preserving key of a cull distance: negative distances pass every test, as 0 does; non-negative floats order as uints
PTB_DEV uint32_t packet_key(float t) { return __float_as_uint(fmaxf(t, 0.0f)); }

// All 32 lanes call this together. `valid` lanes hold a ray; on return best_t / best_ref hold each lane's closest hit
// (best_ref == kNone: miss). n_nodes / n_prims count what the lane took part in (COUNT).
// OCT = the octant all valid lanes' directions lie in (bit 0 / 1 / 2: negative along x / y / z: the slab test then needs no
// sign selects), or -1 when they do not share one.
template <bool COUNT, int OCT>
PTB_DEV void packet_trace_oct(const DevScene& sc, const Ray& ray, const SlabRay& sr, bool valid, float& best_t, uint32_t& best_ref,
                              uint32_t& n_nodes, uint32_t& n_prims) {
  const uint32_t lane = threadIdx.x & 31u;
  best_t = __int_as_float(0x7f800000);
  best_ref = kNone;
  PacketStack st;
  st.ref0 = st.ref1 = st.key0 = st.key1 = st.mask0 = st.mask1 = 0u;
  int sp = 0;
  uint32_t mask = __ballot_sync(0xffffffffu, valid);
  uint32_t cur = (sc.n_prims && mask) ? 0u : kNone;
  uint32_t best_max = 0x7f800000u;  // largest best_t among the valid lanes, as bits (uniform)
  const uint32_t all = mask;
  while (cur != kNone) {
    const bool in = (mask >> lane) & 1u;
    bool pop = true;
    if (cur & PTB_LEAF_BIT) {
      if (in) {
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
      }
      best_max = __reduce_max_sync(0xffffffffu, ((all >> lane) & 1u) ? __float_as_uint(best_t) : 0u);
    } else {
      float4 n0, n1, n2, n3f;
      ldg256(sc.nodes + cur, n0, n1);
      ldg256(reinterpret_cast<const float4*>(sc.nodes + cur) + 2, n2, n3f);
      const uint32_t cl = __float_as_uint(n3f.x), cr = __float_as_uint(n3f.y);
      if (COUNT && in) ++n_nodes;
      float tl = 0.0f, tr = 0.0f;
      const bool hl = in && (OCT < 0 ? box_entry(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, sr, best_t, tl)
                                     : box_entry_oct<(OCT < 0 ? 0 : OCT)>(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, sr, best_t, tl));
      const bool hr = in && (OCT < 0 ? box_entry(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, sr, best_t, tr)
                                     : box_entry_oct<(OCT < 0 ? 0 : OCT)>(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, sr, best_t, tr));
      const uint32_t ml = __ballot_sync(0xffffffffu, hl), mr = __ballot_sync(0xffffffffu, hr);
      if (ml | mr) {
        pop = false;
        if (!mr) { cur = cl; mask = ml; }
        else if (!ml) { cur = cr; mask = mr; }
        else {
          // both children are entered by some lanes: the one more lanes would enter first goes first
          const uint32_t want_r = __ballot_sync(0xffffffffu, hr & (!hl | (tr < tl)));
          const uint32_t want_l = __ballot_sync(0xffffffffu, hl & (!hr | (tl <= tr)));
          const bool right_first = __popc(want_r) > __popc(want_l);
          const bool far_hit = right_first ? hl : hr;
          const uint32_t far_key = __reduce_min_sync(0xffffffffu, far_hit ? packet_key(right_first ? tl : tr) : 0xffffffffu);
          packet_push(st, sp, lane, right_first ? cl : cr, far_key, right_first ? ml : mr);
          cur = right_first ? cr : cl;
          mask = right_first ? mr : ml;
        }
      }
    }
    while (pop) {
      if (sp == 0) { cur = kNone; break; }
      --sp;
      const uint32_t src = (uint32_t)sp & 31u;
      const uint32_t r = __shfl_sync(0xffffffffu, sp < 32 ? st.ref0 : st.ref1, src);
      const uint32_t k = __shfl_sync(0xffffffffu, sp < 32 ? st.key0 : st.key1, src);
      const uint32_t m = __shfl_sync(0xffffffffu, sp < 32 ? st.mask0 : st.mask1, src);
      if (k <= best_max) { cur = r; mask = m; pop = false; }
    }
  }
}

template <bool COUNT>
PTB_DEV void packet_trace(const DevScene& sc, const Ray& ray, bool valid, float& best_t, uint32_t& best_ref, uint32_t& n_nodes,
                          uint32_t& n_prims) {
  const SlabRay sr = make_slab_ray(ray);
  // do the valid lanes share a direction octant? (the signs box_entry selects by: dinv < 0 after the clamp)
  const uint32_t all = __ballot_sync(0xffffffffu, valid);
  const uint32_t mx = __ballot_sync(0xffffffffu, valid && sr.dinv.x < 0.0f), my = __ballot_sync(0xffffffffu, valid && sr.dinv.y < 0.0f),
                 mz = __ballot_sync(0xffffffffu, valid && sr.dinv.z < 0.0f);
  int oct = -1;
  if ((mx == 0u || mx == all) && (my == 0u || my == all) && (mz == 0u || mz == all))
    oct = (mx ? 1 : 0) | (my ? 2 : 0) | (mz ? 4 : 0);
  switch (oct) {
    case 0: packet_trace_oct<COUNT, 0>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
    case 1: packet_trace_oct<COUNT, 1>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
    case 2: packet_trace_oct<COUNT, 2>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
    case 3: packet_trace_oct<COUNT, 3>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
    case 4: packet_trace_oct<COUNT, 4>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
    case 5: packet_trace_oct<COUNT, 5>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
    case 6: packet_trace_oct<COUNT, 6>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
    case 7: packet_trace_oct<COUNT, 7>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
    default: packet_trace_oct<COUNT, -1>(sc, ray, sr, valid, best_t, best_ref, n_nodes, n_prims); break;
  }
}

}  // namespace ptb