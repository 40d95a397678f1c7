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
// Execution shape. The first ncu capture showed the naive one-ray-per-lane loop issue-bound at 11-15 active lanes of 32,
// so every warp is persistent and all scheduling is warp-synchronous (full-mask ballots; Volta+ gives no lock-step
// guarantee to rely on):
//   * leaf queue: a lane does not test a primitive when it reaches a leaf; it queues the leaf (with the cull key of its
//     box) and keeps walking internal nodes, until its small queue is full (speculative "while-while", Aila & Laine 2009);
//   * phases: each iteration of the warp runs EITHER a burst of node steps for the lanes parked on an internal node OR one
//     primitive step for the lanes with queued leaves — whichever has more ready lanes; queued leaves whose box entry has
//     meanwhile fallen behind the best hit are dropped without a test;
//   * dynamic fetch: when fewer than `trace_fetch_threshold` lanes still have work, finished rays are retired and idle
//     lanes refilled from the global ray queue with one warp-aggregated atomicAdd.
// A node fetch is ONE 32-byte load: both children's boxes (16-bit grid coordinates) and both child references.
#pragma once
#include "ptb_coop.cuh"
#include "ptb_cwbvh.cuh"
#include "ptb_intersect.cuh"

namespace ptb {

#ifndef PTB_PREFETCH
#define PTB_PREFETCH 0  // 1: a node step requests both child nodes into L1 before testing their boxes; 2: leaf geometry too
#endif
PTB_DEV void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
constexpr int kStackDepth = 64;  // LBVH depth <= 30 Morton bits + 32 index tie-break bits: one push per binary level
#ifndef PTB_LEAF_QUEUE
#define PTB_LEAF_QUEUE 1         // postponed leaves per lane (power of two). Measured on C3: 1 -> 1911-1927 Mrays/s,
                                 // 2 -> 1872, 4 -> 1856, 8 -> 1825: deeper queues speculate more (V 32.9 -> 35.8 nodes/ray)
#endif
constexpr uint32_t kLeafQueue = PTB_LEAF_QUEUE;

struct TraceResult {
  float t;       // 0 on miss (sky.rs:79-91)
  uint32_t ref;  // kNone on miss, else (kSphereBit?) | slot
};

PTB_DEV void load_node(const BvhNode* __restrict__ nodes, uint32_t idx, float4& n0, float4& n1, float4& n2, uint4& n3) {
  float4 t;
  ldg256(nodes + idx, n0, n1);
  ldg256(reinterpret_cast<const float4*>(nodes + idx) + 2, n2, t);
  n3 = make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w));
}

// Per-lane traversal state. `cur`: internal node index, or a PARKED leaf reference (bit 31: the leaf queue was full), or
// kNone when the stack ran dry. The leaf queue is a ring of (reference, cull key) in local memory.
// (Round 1 also carried a shared-memory top-of-stack, a register top-of-stack and an uncompressed 4-wide step as
// compile-time options; all three measured slower, profiles/r1_sweeps.md, and left with the compressed 8-wide tree of
// ptb_cwbvh.cuh.)
constexpr int kTraceThreads = 256;  // block size of every kernel that runs persistent_trace

// Traversal stack: (L1-cached) local memory. Lanes sit at different depths, so an access of a warp touches one 128-byte
// line PER LANE (ncu: a third of the L1 wavefronts of the binary k_trace are stack traffic).
struct TravStack {
  uint2* local;   // kStackDepth entries (per lane)
};

struct TravState {
  uint32_t cur;
  float cur_key;      // cull key of a parked leaf
  int sp;
  uint32_t lq_head, lq_count;
  float best_t;       // closest hit so far (closest-hit) / tmax (any-hit)
  uint32_t best_ref;  // closest-hit: winning leaf ref; any-hit: kNone = unoccluded, 0 = occluded
  PTB_DEV bool done() const { return cur == kNone && lq_count == 0u; }
};

PTB_DEV void trav_init(TravState& s, uint32_t n_prims, float tmax) {
  s.cur = n_prims ? 0u : kNone;
  s.cur_key = 0.0f;
  s.sp = 0;
  s.lq_head = 0u;
  s.lq_count = 0u;
  s.best_t = tmax;
  s.best_ref = kNone;
}

PTB_DEV void lq_push(TravState& s, uint2* lq, uint32_t ref, float key) {
  lq[(s.lq_head + s.lq_count) & (kLeafQueue - 1u)] = make_uint2(ref, __float_as_uint(key));
  ++s.lq_count;
}

PTB_DEV bool stack_empty(const TravState& s) { return s.sp == 0; }
PTB_DEV void stack_push(TravState& s, const TravStack& k, uint2 e) { k.local[s.sp++] = e; }
PTB_DEV uint2 stack_pop(TravState& s, const TravStack& k) { return k.local[--s.sp]; }  // requires !stack_empty

// Stack entry = (node or leaf reference, cull key of its box) in one 8-byte word.
// Pops until an internal node is found (-> cur), the stack is empty (-> kNone), or a leaf turns up while the leaf queue is
// full (-> parked in cur). Entries whose box can no longer hold a closer hit are dropped; leaves go to the queue.
PTB_DEV void trav_pop(TravState& s, const TravStack& stack, uint2* lq) {
  for (;;) {
    if (stack_empty(s)) { s.cur = kNone; return; }
    const uint2 e = stack_pop(s, stack);
    const float key = __uint_as_float(e.y);
    if (key <= s.best_t) {
      if (e.x & PTB_LEAF_BIT) {
        if (s.lq_count < kLeafQueue) { lq_push(s, lq, e.x, key); continue; }
        s.cur_key = key;
      }
      s.cur = e.x;
      return;
    }
  }
}

// One internal-node step of the lane: fetch the 64-byte node, test both child boxes, descend into the nearer hit child
// (deferring the other on the stack); a leaf child is queued and the walk continues from the stack.
template <bool COUNT>
PTB_DEV void trav_node_step(const DevScene& sc, const BinRayCtx& ray, TravState& s, const TravStack& stack, uint2* lq,
                            uint32_t& n_nodes) {
#if PTB_QNODES
  // 32-byte node: {left box x, y, z, right box x} {right box y, z, left child, right child} (ptb_intersect.cuh: 16-bit boxes)
  uint4 qa, n3;
  ldg256u(sc.qnodes + 2u * (size_t)s.cur, qa, n3);
  if (COUNT) ++n_nodes;
  float tl, tr;
  const bool hl = box_entry_q(qa.x, qa.y, qa.z, ray, s.best_t, tl);
  const bool hr = box_entry_q(qa.w, n3.x, n3.y, ray, s.best_t, tr);
  n3.x = n3.z;  // left, right
  n3.y = n3.w;
#else
  float4 n0, n1, n2;
  uint4 n3;
  load_node(sc.nodes, s.cur, n0, n1, n2, n3);
  if (COUNT) ++n_nodes;
  float tl, tr;
  const bool hl = box_entry(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, ray, s.best_t, tl);
  const bool hr = box_entry(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, ray, s.best_t, tr);
#endif
#if PTB_PREFETCH >= 1 && PTB_QNODES
  // (measured: 13 - 17 % SLOWER on C3 — the walk is bound by LSU / issue throughput, not by latency)
  // The walk is a chain of dependent fetches (ncu: long-scoreboard is 58 % of a warp's time on the bounce launches). Both
  // children are requested into L1 the moment their references arrive, before the box tests decide which one is
  // entered: the next step's load then finds its line on the way or there. With 32-byte nodes the two requests cost the
  // bytes ONE 64-byte node used to.
  if (!(n3.x & PTB_LEAF_BIT)) prefetch_l1(sc.qnodes + 2u * (size_t)n3.x);
#if PTB_PREFETCH >= 2
  else prefetch_l1(sc.geom + 3u * (size_t)(n3.x & kSlotMask));
#endif
  if (!(n3.y & PTB_LEAF_BIT)) prefetch_l1(sc.qnodes + 2u * (size_t)n3.y);
#if PTB_PREFETCH >= 2
  else prefetch_l1(sc.geom + 3u * (size_t)(n3.y & kSlotMask));
#endif
#endif
  bool want_pop = !(hl || hr);
  if (!want_pop) {
    const bool right_first = hr & (!hl | (tr < tl));  // (no short circuit: it compiles to a divergent branch)
    if (hl & hr) stack_push(s, stack, make_uint2(right_first ? n3.x : n3.y, __float_as_uint(right_first ? tl : tr)));
    s.cur = right_first ? n3.y : n3.x;
    if (s.cur & PTB_LEAF_BIT) {
      const float key = right_first ? tr : tl;
      if (s.lq_count < kLeafQueue) {
        lq_push(s, lq, s.cur, key);
        want_pop = true;
      } else {
        s.cur_key = key;  // queue full: park on the leaf until the next primitive phase
      }
    }
  }
  if (want_pop) trav_pop(s, stack, lq);
}

// One primitive step of the lane: take the oldest queued leaf (the nearest, as the walk is near-first), drop it if its box
// has fallen behind the best hit, else test it; then move a parked leaf into the freed queue slot and resume the walk.
template <bool ANYHIT, bool COUNT>
PTB_DEV void trav_prim_step(const DevScene& sc, const Ray& ray, TravState& s, const TravStack& stack, uint2* lq, uint32_t exclude,
                            uint32_t& n_prims) {
  const uint2 e = lq[s.lq_head];
  s.lq_head = (s.lq_head + 1u) & (kLeafQueue - 1u);
  --s.lq_count;
  const uint32_t ref = e.x;
  if (__uint_as_float(e.y) <= s.best_t) {
    if (ANYHIT) {
      if ((ref & kSlotMask) != exclude) {
        const float t = prim_t(sc, ray, ref);
        if (COUNT) ++n_prims;
        if (t > 0.0f && t < s.best_t) {  // blocker found: stop
          s.best_ref = 0u;
          s.cur = kNone;
          s.sp = 0;
          s.lq_count = 0u;
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
  }
  if ((s.cur & PTB_LEAF_BIT) && s.cur != kNone) {  // a leaf was parked: it fits now
    lq_push(s, lq, s.cur, s.cur_key);
    trav_pop(s, stack, lq);
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

// occupancy targets of the persistent kernels (blocks of 256 threads per SM -> register cap), per tree
#ifndef PTB_TRACE_MIN_BLOCKS
#define PTB_TRACE_MIN_BLOCKS 6  // binary: 40 registers. Window mode, C3: 4 blocks (64 registers) 3611 Mrays/s, 5 (48) 3664, 6 (40) 3730
#endif
#ifndef PTB_API_MIN_BLOCKS
#define PTB_API_MIN_BLOCKS 6    // binary, same-run A/B (profiles/r1_sweeps.md): C5 4116 -> 4816 Mrays/s against unconstrained
#endif
#ifndef PTB_CW_TRACE_MIN_BLOCKS
#define PTB_CW_TRACE_MIN_BLOCKS 4  // wide: 64 registers (the node step holds 12 plane words + 8 header words + the ray)
#endif
#ifndef PTB_CW_API_MIN_BLOCKS
#define PTB_CW_API_MIN_BLOCKS 4
#endif

// ---- the two trees behind one interface: what persistent_trace / trace_lane need from a traversal
struct BinTrav {  // binary LBVH, one primitive per leaf (this file)
  static constexpr int kTraceMinBlocks = PTB_TRACE_MIN_BLOCKS, kApiMinBlocks = PTB_API_MIN_BLOCKS;
  static constexpr bool kCoop = true;   // ptb_coop.cuh walks this tree warp-cooperatively
  typedef TravState State;
  typedef BinRayCtx RayCtx;
  struct Scratch {
    uint2 stack[kStackDepth];
    uint2 lq[kLeafQueue];
  };
  PTB_DEV static RayCtx make(const DevScene& sc, const Ray& ray) { return make_bin_ray(sc, ray); }
  PTB_DEV static RayCtx idle() {
    RayCtx r;
#if PTB_QNODES
    r.a = r.b_lo = r.b_hi = mk(0.0f, 0.0f, 0.0f);
    r.sn_x = r.sn_y = r.sn_z = kSelLow;
#else
    r.dinv = r.c_lo = r.c_hi = mk(0.0f, 0.0f, 0.0f);
#endif
    return r;
  }
  PTB_DEV static void init(State& s, const RayCtx&, uint32_t n_prims, float tmax) { trav_init(s, n_prims, tmax); }
  PTB_DEV static bool node_ready(const State& s) { return !(s.cur & PTB_LEAF_BIT); }
  PTB_DEV static bool leaf_ready(const State& s) { return s.lq_count != 0u; }
  PTB_DEV static bool done(const State& s) { return s.done(); }
  template <bool COUNT>
  PTB_DEV static void node_step(const DevScene& sc, const RayCtx& rc, State& s, Scratch& k, uint32_t& n) {
    trav_node_step<COUNT>(sc, rc, s, TravStack{k.stack}, k.lq, n);
  }
  template <bool ANYHIT, bool COUNT>
  PTB_DEV static void prim_step(const DevScene& sc, const Ray& ray, State& s, Scratch& k, uint32_t exclude, uint32_t& n) {
    trav_prim_step<ANYHIT, COUNT>(sc, ray, s, TravStack{k.stack}, k.lq, exclude, n);
  }
  PTB_DEV static TraceResult result(const State& s) { return trav_result(s); }
};
struct CwTrav {  // compressed 8-wide tree, leaf groups of up to 3 primitives (ptb_cwbvh.cuh)
  static constexpr int kTraceMinBlocks = PTB_CW_TRACE_MIN_BLOCKS, kApiMinBlocks = PTB_CW_API_MIN_BLOCKS;
  static constexpr bool kCoop = false;
  typedef CwState State;
  typedef CwRay RayCtx;
  struct Scratch {
    uint2 stack[kCwStackDepth];
  };
  PTB_DEV static RayCtx make(const DevScene&, const Ray& ray) { return make_cw_ray(ray); }
  PTB_DEV static RayCtx idle() {
    RayCtx r;
    r.dinv = r.neg_od = r.ed = r.eo = mk(0.0f, 0.0f, 0.0f);
    r.oinv = 7u;
    r.neg = 0u;
    return r;
  }
  PTB_DEV static void init(State& s, const RayCtx& rc, uint32_t n_prims, float tmax) { cw_init(s, rc, n_prims, tmax); }
  PTB_DEV static bool node_ready(const State& s) { return s.node_ready(); }
  PTB_DEV static bool leaf_ready(const State& s) { return s.leaf_ready(); }
  PTB_DEV static bool done(const State& s) { return s.done(); }
  template <bool COUNT>
  PTB_DEV static void node_step(const DevScene& sc, const RayCtx& rc, State& s, Scratch& k, uint32_t& n) {
    cw_node_step<COUNT>(sc.cw_nodes, rc, s, k.stack, n);
  }
  template <bool ANYHIT, bool COUNT>
  PTB_DEV static void prim_step(const DevScene& sc, const Ray& ray, State& s, Scratch& k, uint32_t exclude, uint32_t& n) {
    cw_prim_step<ANYHIT, COUNT>(sc, ray, s, k.stack, exclude, n);
  }
  PTB_DEV static TraceResult result(const State& s) {
    TraceResult r;
    r.t = 0.0f;
    r.ref = kNone;
    if (s.best_ref != kNone) { r.t = s.best_t; r.ref = s.best_ref; }
    return r;
  }
};

#ifdef PTB_DRAIN_STATS  // tuning builds only: how long does a persistent launch run after its queue is empty?
// [0] first warp finds the queue empty (ns, abs; ~0 = not yet) [1] last warp leaves [2] first warp starts
// [3] sum of drains [4] sum of launch durations [5] launches folded (k_win_prepare folds and resets)
__device__ unsigned long long g_drain[6] = {~0ull, 0ull, ~0ull, 0ull, 0ull, 0ull};
PTB_DEV unsigned long long drain_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif
#ifdef PTB_LANE_STATS  // tuning builds only: where do the lanes of a warp spend their iterations?
// [0] loop iterations  [1] lanes with work (sum)  [2] node phases  [3] node-ready lanes in them  [4] primitive phases
// [5] leaf-ready lanes in them  [6] service passes  [7] node steps executed (lane level)
__device__ unsigned long long g_lane_stats[8];
#define PTB_LS(i, v) do { if (lane == 0u) ls[i] += (v); } while (0)
#else
#define PTB_LS(i, v) do { } while (0)
#endif

// Persistent-warp driver. `fetch(i, ray, tmax, exclude)` loads work item i into the lane; `retire(fin, result, ray)` is
// called by ALL 32 lanes together (fin = this lane just completed its item; for any-hit work result.ref == kNone means
// unoccluded) so it may use warp-wide primitives. TR = BinTrav or CwTrav.
template <class TR, bool ANYHIT, bool COUNT, class Fetch, class Retire>
PTB_DEV void persistent_trace(const DevScene& sc, uint32_t n, uint32_t* head, Fetch& fetch, Retire& retire,
                              uint32_t& cnt_nodes, uint32_t& cnt_prims, uint32_t& cnt_rays) {
  const uint32_t lane = threadIdx.x & 31u;
  typename TR::Scratch scratch;
  typename TR::State st;
  typename TR::RayCtx rc = TR::idle();
  TR::init(st, rc, 0u, 0.0f);
  Ray ray;
  ray.o = ray.d = ray.dinv = ray.shear = mk(0.0f, 0.0f, 0.0f);
  ray.swap_xz = false;
  uint32_t exclude = kNone;
  bool has_ray = false, exhausted = false;
  // Small launches (the tail of a render: a few thousand long paths) are latency bound: 32 rays in one warp run their
  // phases one after the other while most SMs idle. Cap the rays a warp holds so the launch spreads over every
  // resident warp; large launches keep all 32 lanes.
  const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);
  uint32_t cap = (n + total_warps - 1u) / total_warps;
  cap = cap < 1u ? 1u : (cap > 32u ? 32u : cap);
  const uint32_t fetch_below = (uint32_t)sc.trace_fetch_threshold < cap ? (uint32_t)sc.trace_fetch_threshold : cap;
#ifdef PTB_LANE_STATS
  unsigned long long ls[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
#ifdef PTB_DRAIN_STATS
  if (threadIdx.x == 0u) atomicMin(&g_drain[2], drain_now());
#endif
  for (;;) {
    // a lane with work can take a node step, a primitive step, or (binary tree: queued leaf while walking) both
    bool node_ready = has_ray && TR::node_ready(st);
    const bool leaf_ready = has_ray && TR::leaf_ready(st);
    const uint32_t m_node = __ballot_sync(0xffffffffu, node_ready);
    const uint32_t m_leaf = __ballot_sync(0xffffffffu, leaf_ready);
    PTB_LS(0, 1);
    PTB_LS(1, __popc(m_node | m_leaf));
    if ((uint32_t)__popc(m_node | m_leaf) < (exhausted ? 1u : fetch_below)) {
      PTB_LS(6, 1);
      // ---- service: retire finished items, refill idle lanes
      const bool fin = has_ray && !node_ready && !leaf_ready;
      retire(fin, TR::result(st), ray);
      if (fin) has_ray = false;
      if (exhausted) {
        if (!__any_sync(0xffffffffu, has_ray)) break;
        continue;
      }
      uint32_t idle = __ballot_sync(0xffffffffu, !has_ray);
      if (cap < 32u) {  // keep only as many idle lanes as the cap allows (lowest lanes first)
        const uint32_t busy = 32u - (uint32_t)__popc(idle);
        const uint32_t allow = busy < cap ? cap - busy : 0u;
        const uint32_t cut = __fns(idle, 0u, (int)allow + 1);  // position of the (allow+1)-th idle lane, or ~0u
        if (cut != 0xffffffffu) idle &= (1u << cut) - 1u;
      }
      if (idle) {
        const uint32_t leader = __ffs(idle) - 1u, want = __popc(idle);
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(head, want);
        base = __shfl_sync(0xffffffffu, base, leader);
        const uint32_t mine = base + __popc(idle & ((1u << lane) - 1u));
        if (((idle >> lane) & 1u) && mine < n) {
          float tmax = __int_as_float(0x7f800000);
          fetch(mine, ray, tmax, exclude);
          rc = TR::make(sc, ray);
          TR::init(st, rc, sc.n_prims, tmax);
          has_ray = true;
          if (COUNT) ++cnt_rays;
        }
        if (base + want >= n) exhausted = true;
#ifdef PTB_DRAIN_STATS
        if (lane == leader && base + want >= n) atomicMin(&g_drain[0], drain_now());
#endif
      }
      continue;
    }
    // node phase while it keeps at least as many lanes busy as a primitive phase would (blocked lanes = primitives
    // pending but nowhere to walk; `trace_prim_bias` shifts the balance towards batching more leaves per primitive phase)
    const int n_node = __popc(m_node), n_blocked = __popc(m_leaf & ~m_node), n_leaf = __popc(m_leaf);
    const bool do_node = sc.trace_prim_bias ? (n_node >= n_blocked * sc.trace_prim_bias) : (n_node >= n_leaf);
    if (do_node) {
      PTB_LS(2, 1);
      PTB_LS(3, n_node);
#pragma unroll 1
      for (int burst = 0; burst < sc.trace_burst && node_ready; ++burst) {
        TR::template node_step<COUNT>(sc, rc, st, scratch, cnt_nodes);
        node_ready = TR::node_ready(st);
#ifdef PTB_LANE_STATS
        atomicAdd(&g_lane_stats[7], 1ull);
#endif
      }
    } else if (leaf_ready) {
      TR::template prim_step<ANYHIT, COUNT>(sc, ray, st, scratch, exclude, cnt_prims);
    }
#ifdef PTB_LANE_STATS
    if (!do_node) { PTB_LS(4, 1); PTB_LS(5, n_leaf); }
#endif
  }
#ifdef PTB_LANE_STATS
  if (lane == 0u)
    for (int i = 0; i < 7; ++i) atomicAdd(&g_lane_stats[i], ls[i]);
#endif
#ifdef PTB_DRAIN_STATS
  if (lane == 0u) atomicMax(&g_drain[1], drain_now());
#endif
}

// One ray, one lane, start to finish (the fused tail kernel): the same steps without the warp-level scheduling.
template <class TR, bool ANYHIT>
PTB_DEV TraceResult trace_lane(const DevScene& sc, const Ray& ray, float tmax, uint32_t exclude, typename TR::Scratch& scratch) {
  typename TR::State st;
  const typename TR::RayCtx rc = TR::make(sc, ray);
  TR::init(st, rc, sc.n_prims, tmax);
  uint32_t unused = 0;
  while (!TR::done(st)) {
    if (TR::node_ready(st)) TR::template node_step<false>(sc, rc, st, scratch, unused);
    else TR::template prim_step<ANYHIT, false>(sc, ray, st, scratch, exclude, unused);
  }
  return TR::result(st);
}

}  // namespace ptb
