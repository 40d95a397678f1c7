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
// Node fetches are four 16-byte loads of one 64-byte node that carries BOTH children's boxes.
#pragma once
#include "ptb_intersect.cuh"

namespace ptb {

constexpr int kStackDepth = PTB_WIDE_BVH ? 96 : 64;  // LBVH depth <= 30 Morton bits + 32 index tie-break bits: one push per
                                                   // binary level, up to three per 4-wide level (two binary levels each)
#ifndef PTB_LEAF_QUEUE
#define PTB_LEAF_QUEUE 1         // postponed leaves per lane (power of two). Measured on C3: 1 -> 1911-1927 Mrays/s,
                                 // 2 -> 1872, 4 -> 1856, 8 -> 1825: deeper queues speculate more (V 32.9 -> 35.8 nodes/ray)
#endif
constexpr uint32_t kLeafQueue = PTB_LEAF_QUEUE;

struct TraceResult {
  float t;       // 0 on miss (sky.rs:79-91)
  uint32_t ref;  // kNone on miss, else (kSphereBit?) | slot
};

// 32 bytes per lane in one instruction (LDG.E.256, new on sm_100): incoherent traversal is bound by L1 wavefronts — one
// per distinct 128-byte line PER LOAD INSTRUCTION (ncu: l1tex throughput 80 % with four 16-byte loads per node) — so a
// 64-byte node costs two wavefronts instead of four. `p` must be 32-byte aligned.
PTB_DEV void ldg256(const void* p, float4& a, float4& b) {
#ifdef PTB_NO_LDG256
  a = __ldg(reinterpret_cast<const float4*>(p));
  b = __ldg(reinterpret_cast<const float4*>(p) + 1);
#else
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
#endif
}
#ifndef PTB_STREAM_HINTS
#define PTB_STREAM_HINTS 0  // 1 = path records are loaded / stored with the evict-first (.cs) policy so that they do not displace
                            // BVH nodes and triangles from L2
#endif
#if PTB_STREAM_HINTS
#define PTB_CS ".cs"
#else
#define PTB_CS ""
#endif
PTB_DEV void ldg256_rw(const void* p, float4& a, float4& b) {  // same, for data this launch sequence also writes (no .nc)
  asm volatile("ld.global" PTB_CS ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p) : "memory");
}
PTB_DEV void stg256(void* p, float4 a, float4 b) {
  asm volatile("st.global" PTB_CS ".v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
PTB_DEV void load_node(const BvhNode* __restrict__ nodes, uint32_t idx, float4& n0, float4& n1, float4& n2, uint4& n3) {
  float4 t;
  ldg256(nodes + idx, n0, n1);
  ldg256(reinterpret_cast<const float4*>(nodes + idx) + 2, n2, t);
  n3 = make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w));
}

// Per-lane traversal state. `cur`: internal node index, or a PARKED leaf reference (bit 31: the leaf queue was full), or
// kNone when the stack ran dry. The leaf queue is a ring of (reference, cull key) in local memory.
#ifndef PTB_SHARED_STACK
#define PTB_SHARED_STACK 0       // top-of-stack entries per lane kept in shared memory (power of two; 0 = local memory only).
                                 // Measured on B200, C3 (profiles/r1_sweeps.md): 0 -> 2228 Mrays/s, 4 -> 2011, 8 -> 2045,
                                 // 16 -> 2036: the extra index arithmetic and the smaller L1 cost more than the
                                 // conflict-free accesses save, so the local-memory stack stays the default.
#endif
constexpr int kSharedStack = PTB_SHARED_STACK;
constexpr int kTraceThreads = 256;  // block size of every kernel that runs persistent_trace

// Traversal stack: (L1-cached) local memory, optionally with the newest kSharedStack entries in shared memory.
// Lanes sit at different depths, so a local-memory access of a warp touches one 128-byte line PER LANE (ncu: a third of
// the L1 wavefronts of k_trace are stack traffic); the shared copy is indexed [depth][thread], its bank depends on the
// thread only, and an 8-byte access of a full warp is always two conflict-free wavefronts. Entries [lo, sp) are in
// shared memory at slot (index mod kSharedStack), entries [0, lo) in local memory.
struct TravStack {
  uint2* local;   // kStackDepth entries (per lane)
  uint2* shared;  // &smem[0][threadIdx.x], stride kTraceThreads
};

struct TravState {
  uint32_t cur;
  float cur_key;      // cull key of a parked leaf
  int sp;
  uint2 tos;          // PTB_TOS_REG: newest stack entry kept in registers (tos.x == kNone: empty)
  int lo;             // first stack index resident in shared memory
  uint32_t lq_head, lq_count;
  float best_t;       // closest hit so far (closest-hit) / tmax (any-hit)
  uint32_t best_ref;  // closest-hit: winning leaf ref; any-hit: kNone = unoccluded, 0 = occluded
  PTB_DEV bool done() const { return cur == kNone && lq_count == 0u; }
};

PTB_DEV void trav_init(TravState& s, uint32_t n_prims, float tmax) {
  s.cur = n_prims ? 0u : kNone;
  s.cur_key = 0.0f;
  s.sp = 0;
  s.tos = make_uint2(kNone, 0u);
  s.lo = 0;
  s.lq_head = 0u;
  s.lq_count = 0u;
  s.best_t = tmax;
  s.best_ref = kNone;
}

PTB_DEV void lq_push(TravState& s, uint2* lq, uint32_t ref, float key) {
  lq[(s.lq_head + s.lq_count) & (kLeafQueue - 1u)] = make_uint2(ref, __float_as_uint(key));
  ++s.lq_count;
}

#ifndef PTB_TOS_REG
#define PTB_TOS_REG 0   // 1 = keep the newest stack entry in registers ("push the far child, reach a leaf, pop it right
                        // back" then never touches local memory). Measured on B200, C3 window mode: k_trace 176 ms vs
                        // 160 ms without (3276 vs 3546 Mrays/s): the selects cost issue slots the kernel does not have.
#endif
PTB_DEV bool stack_empty(const TravState& s) { return s.sp == 0 && (!PTB_TOS_REG || s.tos.x == kNone); }
PTB_DEV void stack_push(TravState& s, const TravStack& k, uint2 e) {
  if (PTB_TOS_REG && kSharedStack == 0) {
    if (s.tos.x != kNone) k.local[s.sp++] = s.tos;
    s.tos = e;
    return;
  }
  if (kSharedStack == 0) { k.local[s.sp++] = e; return; }
  if (s.sp - s.lo == kSharedStack) {  // shared part full: its oldest entry moves to local memory
    k.local[s.lo] = k.shared[(s.lo & (kSharedStack - 1)) * kTraceThreads];
    ++s.lo;
  }
  k.shared[(s.sp & (kSharedStack - 1)) * kTraceThreads] = e;
  ++s.sp;
}
PTB_DEV uint2 stack_pop(TravState& s, const TravStack& k) {  // requires !stack_empty
  if (PTB_TOS_REG && kSharedStack == 0) {
    if (s.tos.x != kNone) {
      const uint2 e = s.tos;
      s.tos.x = kNone;
      return e;
    }
    return k.local[--s.sp];
  }
  --s.sp;
  if (kSharedStack == 0) return k.local[s.sp];
  if (s.sp < s.lo) {  // shared part empty: read the spilled entry in place
    s.lo = s.sp;
    return k.local[s.sp];
  }
  return k.shared[(s.sp & (kSharedStack - 1)) * kTraceThreads];
}

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
PTB_DEV void trav_node_step(const DevScene& sc, const SlabRay& ray, const Ray& full, TravState& s, const TravStack& stack, uint2* lq,
                            uint32_t& n_nodes) {
  float4 n0, n1, n2;
  uint4 n3;
  load_node(sc.nodes, s.cur, n0, n1, n2, n3);
  if (COUNT) ++n_nodes;
  float tl, tr;
#ifdef PTB_BOX_V1
  const bool hl = box_entry_v1(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, full, s.best_t, tl);
  const bool hr = box_entry_v1(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, full, s.best_t, tr);
#else
  const bool hl = box_entry(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, ray, s.best_t, tl);
  const bool hr = box_entry(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, ray, s.best_t, tr);
#endif
  bool want_pop = !(hl || hr);
  if (!want_pop) {
    const bool both = hl && hr;
    const bool right_first = both ? (tr < tl) : hr;
    if (both) stack_push(s, stack, make_uint2(right_first ? n3.x : n3.y, __float_as_uint(right_first ? tl : tr)));
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

// The same step over a 4-wide node (PTB_WIDE_BVH): four 32-byte loads, four slab tests, the hit children ranked by cull key
// (ties: lower slot first); the nearest becomes `cur`, the others go to the stack with the nearest on top — each hit child
// stores itself at the position its rank gives it, no sorting network. Identical decisions to Lbvh::closest_hit_wide.
template <bool COUNT>
PTB_DEV void trav_node_step4(const DevScene& sc, const SlabRay& ray, TravState& s, const TravStack& stack, uint2* lq,
                             uint32_t& n_nodes) {
  static_assert(!PTB_WIDE_BVH || kSharedStack == 0, "the wide step writes the local-memory stack directly");
  const float4* p = reinterpret_cast<const float4*>(sc.nodes4 + s.cur);
  float4 a0, a1, b0, b1, c0, c1, d0, d1;
  ldg256(p, a0, a1);
  ldg256(p + 2, b0, b1);
  ldg256(p + 4, c0, c1);
  ldg256(p + 6, d0, d1);
  if (COUNT) ++n_nodes;
  const uint32_t r0 = __float_as_uint(d0.x), r1 = __float_as_uint(d0.y), r2 = __float_as_uint(d0.z), r3 = __float_as_uint(d0.w);
  float t0, t1, t2, t3;
  const bool h0 = box_entry(a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, ray, s.best_t, t0);  // slot 0 is never empty
  const bool h1 = box_entry(a1.z, a1.w, b0.x, b0.y, b0.z, b0.w, ray, s.best_t, t1);  // nor slot 1
  const bool h2 = r2 != kNone && box_entry(b1.x, b1.y, b1.z, b1.w, c0.x, c0.y, ray, s.best_t, t2);
  const bool h3 = r3 != kNone && box_entry(c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, ray, s.best_t, t3);
  const int n_hit = (int)h0 + (int)h1 + (int)h2 + (int)h3;
  bool want_pop = n_hit == 0;
  if (!want_pop) {
    // rank of a hit child = number of hit children that come before it
    const int k0 = (int)(h1 && t1 < t0) + (int)(h2 && t2 < t0) + (int)(h3 && t3 < t0);
    const int k1 = (int)(h0 && t0 <= t1) + (int)(h2 && t2 < t1) + (int)(h3 && t3 < t1);
    const int k2 = (int)(h0 && t0 <= t2) + (int)(h1 && t1 <= t2) + (int)(h3 && t3 < t2);
    const int k3 = (int)(h0 && t0 <= t3) + (int)(h1 && t1 <= t3) + (int)(h2 && t2 <= t3);
    uint2* top = stack.local + s.sp + n_hit - 1;  // rank r (>= 1) lives at top[-r]
    uint32_t cur = r0;
    float key = t0;
    if (h0 && k0) top[-k0] = make_uint2(r0, __float_as_uint(t0));
    if (h1) { if (k1) top[-k1] = make_uint2(r1, __float_as_uint(t1)); else { cur = r1; key = t1; } }
    if (h2) { if (k2) top[-k2] = make_uint2(r2, __float_as_uint(t2)); else { cur = r2; key = t2; } }
    if (h3) { if (k3) top[-k3] = make_uint2(r3, __float_as_uint(t3)); else { cur = r3; key = t3; } }
    s.sp += n_hit - 1;
    s.cur = cur;
    if (cur & PTB_LEAF_BIT) {
      if (s.lq_count < kLeafQueue) {
        lq_push(s, lq, cur, key);
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
          s.tos.x = kNone;
          s.lo = 0;
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

#ifdef PTB_LANE_STATS  // tuning builds only: where do the lanes of a warp spend their iterations?
// [0] loop iterations  [1] lanes with work (sum)  [2] node phases  [3] node-ready lanes in them  [4] primitive phases
// [5] leaf-ready lanes in them  [6] service passes  [7] node steps executed (lane level)
__device__ unsigned long long g_lane_stats[8];
#define PTB_LS(i, v) do { if (lane == 0u) ls[i] += (v); } while (0)
#else
#define PTB_LS(i, v) do { } while (0)
#endif

// Persistent-warp driver. `fetch(i, ray, tmax, exclude)` loads work item i into the lane; `retire(fin, state, ray)` is
// called by ALL 32 lanes together (fin = this lane just completed its item) so it may use warp-wide primitives.
template <bool ANYHIT, bool COUNT, class Fetch, class Retire>
PTB_DEV void persistent_trace(const DevScene& sc, uint32_t n, uint32_t* head, Fetch& fetch, Retire& retire,
                              uint32_t& cnt_nodes, uint32_t& cnt_prims, uint32_t& cnt_rays) {
  const uint32_t lane = threadIdx.x & 31u;
  uint2 stack_local[kStackDepth];
  __shared__ uint2 stack_shared[(kSharedStack ? kSharedStack : 1) * kTraceThreads];
  const TravStack stack{stack_local, stack_shared + threadIdx.x};
  uint2 lq[kLeafQueue];
  TravState st;
  trav_init(st, 0u, 0.0f);
  Ray ray;
  ray.o = ray.d = ray.dinv = ray.shear = mk(0.0f, 0.0f, 0.0f);
  ray.swap_xz = false;
  SlabRay slab;
  slab.dinv = slab.c_lo = slab.c_hi = mk(0.0f, 0.0f, 0.0f);
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
  for (;;) {
    // a lane with work is parked on an internal node, holds queued leaves, or both
    bool node_ready = has_ray && !(st.cur & PTB_LEAF_BIT);
    const bool leaf_ready = has_ray && st.lq_count != 0u;
    const uint32_t m_node = __ballot_sync(0xffffffffu, node_ready);
    const uint32_t m_leaf = __ballot_sync(0xffffffffu, leaf_ready);
    PTB_LS(0, 1);
    PTB_LS(1, __popc(m_node | m_leaf));
    if ((uint32_t)__popc(m_node | m_leaf) < (exhausted ? 1u : fetch_below)) {
      PTB_LS(6, 1);
      // ---- service: retire finished items, refill idle lanes
      const bool fin = has_ray && !node_ready && !leaf_ready;
      retire(fin, st, ray);
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
          slab = make_slab_ray(ray);
          trav_init(st, sc.n_prims, tmax);
          has_ray = true;
          if (COUNT) ++cnt_rays;
        }
        if (base + want >= n) exhausted = true;
      }
      continue;
    }
    // node phase while it keeps at least as many lanes busy as a primitive phase would (blocked lanes = queued leaves
    // but nowhere to walk; `trace_prim_bias` shifts the balance towards batching more leaves per primitive phase)
    const int n_node = __popc(m_node), n_blocked = __popc(m_leaf & ~m_node), n_leaf = __popc(m_leaf);
    const bool do_node = sc.trace_prim_bias ? (n_node >= n_blocked * sc.trace_prim_bias) : (n_node >= n_leaf);
    if (do_node) {
      PTB_LS(2, 1);
      PTB_LS(3, n_node);
#pragma unroll 1
      for (int burst = 0; burst < sc.trace_burst && node_ready; ++burst) {
#if PTB_WIDE_BVH
        trav_node_step4<COUNT>(sc, slab, st, stack, lq, cnt_nodes);
#else
        trav_node_step<COUNT>(sc, slab, ray, st, stack, lq, cnt_nodes);
#endif
        node_ready = !(st.cur & PTB_LEAF_BIT);
#ifdef PTB_LANE_STATS
        atomicAdd(&g_lane_stats[7], 1ull);
#endif
      }
    } else if (leaf_ready) {
      trav_prim_step<ANYHIT, COUNT>(sc, ray, st, stack, lq, exclude, cnt_prims);
    }
#ifdef PTB_LANE_STATS
    if (!do_node) { PTB_LS(4, 1); PTB_LS(5, n_leaf); }
#endif
  }
#ifdef PTB_LANE_STATS
  if (lane == 0u)
    for (int i = 0; i < 7; ++i) atomicAdd(&g_lane_stats[i], ls[i]);
#endif
}

}  // namespace ptb
