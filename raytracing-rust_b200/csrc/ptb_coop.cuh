// Warp-cooperative traversal of the binary LBVH (device): up to kCoopRays rays of ONE warp share one frontier in shared
// memory and all 32 lanes work on it — the latency tool of ptb200, where ptb_traverse.cuh is the throughput tool.
//
// Why. A lane that walks a ray alone pays one dependent L2 round trip per node: ~150 node visits for a ray inside the
// 200 k-triangle glass sphere of C3 = ~60 us per bounce, and a path that bounces 50 times holds the end of a render for
// 3.5 ms while 99 % of the machine idles (profiles/r2_mid_c3.md, k_tail). The node visits of one ray are independent of
// each other except through the best hit so far, so a warp can take them 32 at a time: the number of dependent round
// trips drops from the number of nodes visited to about the depth of the tree.
//
// What. Same result as check_hit / check_hit_index (implementations/src/acceleration/mod.rs:226-298), bit for bit: the
// minimum t > 0 over every primitive whose boxes the ray crosses, exact ties to the lower ORIGINAL primitive id (Q2) —
// which is an order-independent definition, so the frontier may be processed in any order: (t, id) is packed into one
// 64-bit word and combined with atomicMin. Box culling uses the same conservative box_entry as the lane walk.
//
// How. The frontier is a stack of (node or leaf reference, cull key | ray tag) entries. A round pops the top <= 32
// entries (one per lane), drops those whose box has fallen behind their ray's best hit, loads a 64-byte node or a
// primitive per lane, and pushes the surviving children (far child first, so the nearest work stays on top: the order
// of a depth-first walk, 32 wide). Pushing is a ballot + popcount compaction; nothing is ever dropped: when the stack is
// nearly full the rounds shrink to one entry (a plain depth-first walk, which needs at most one more entry per level).
#pragma once
#include "ptb_intersect.cuh"

namespace ptb {

constexpr uint32_t kCoopRays = 16;    // rays that may share a warp's frontier (tag in the cull key's low 5 bits)
constexpr uint32_t kCoopTagMask = 31u;
#ifndef PTB_COOP_CAP
#define PTB_COOP_CAP 1024
#endif
constexpr uint32_t kCoopCap = PTB_COOP_CAP;   // stack entries per warp
constexpr uint32_t kCoopSoft = kCoopCap - 96u;  // above it: one entry per round (<= +1 entry per tree level, depth <= 62)

#ifdef PTB_TAIL_STATS  // tuning builds: [0] launches with work [1] paths [2] lane-walk bounces [3] paths handed to the cooperative
// phase [4] bounces [5] coop_trace calls [6] rounds [7] entries processed [9] first warp starts (ns, abs) [10] last warp ends
// [12] entries culled at pop [13] ns inside coop_trace (summed over warps)
__device__ unsigned long long g_tail_stats[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, ~0ull, 0, 0, 0, 0, 0, 0};
#define PTB_TS(i, v) atomicAdd(&g_tail_stats[i], (unsigned long long)(v))
#else
#define PTB_TS(i, v) do { } while (0)
#endif

struct CoopRay {  // what a node test and a primitive test need of a ray, 80 bytes
  float ox, oy, oz, dx, dy, dz, shx, shy, shz;
  uint32_t swap_xz;  // bit 0: Ray::swap_xz (PTB_QNODES: bits 1-3 = the direction is negative along x / y / z)
  float ax, ay, az, lx, ly, lz, hx, hy, hz;  // the slab constants: SlabRay dinv, c_lo, c_hi (PTB_QNODES: QSlabRay a, b_lo, b_hi)
  uint32_t exclude;                           // any-hit: slot that does not block (the surface the ray leaves), else kNone
};
struct CoopWarp {
  unsigned long long best[kCoopRays];  // (t bits << 32) | original primitive id; t bits = tmax, id = ~0 while nothing is hit
  uint32_t ref[kCoopRays];             // winning leaf reference (kSphereBit | slot), kNone while nothing is hit
  CoopRay ray[kCoopRays];
  uint2 stack[kCoopCap];
};

PTB_DEV void coop_set_ray(const DevScene& sc, CoopWarp& cw, uint32_t j, const Ray& ray, float tmax, uint32_t exclude) {
  const BinRayCtx s = make_bin_ray(sc, ray);
  CoopRay& r = cw.ray[j];
  r.ox = ray.o.x; r.oy = ray.o.y; r.oz = ray.o.z;
  r.dx = ray.d.x; r.dy = ray.d.y; r.dz = ray.d.z;
  r.shx = ray.shear.x; r.shy = ray.shear.y; r.shz = ray.shear.z;
#if PTB_QNODES
  r.ax = s.a.x; r.ay = s.a.y; r.az = s.a.z;
  r.lx = s.b_lo.x; r.ly = s.b_lo.y; r.lz = s.b_lo.z;
  r.hx = s.b_hi.x; r.hy = s.b_hi.y; r.hz = s.b_hi.z;
  r.swap_xz = (ray.swap_xz ? 1u : 0u) | (s.sn_x == kSelHigh ? 2u : 0u) | (s.sn_y == kSelHigh ? 4u : 0u) | (s.sn_z == kSelHigh ? 8u : 0u);
#else
  r.ax = s.dinv.x; r.ay = s.dinv.y; r.az = s.dinv.z;
  r.lx = s.c_lo.x; r.ly = s.c_lo.y; r.lz = s.c_lo.z;
  r.hx = s.c_hi.x; r.hy = s.c_hi.y; r.hz = s.c_hi.z;
  r.swap_xz = ray.swap_xz ? 1u : 0u;
#endif
  r.exclude = exclude;
  cw.best[j] = ((unsigned long long)__float_as_uint(tmax) << 32) | 0xFFFFFFFFull;
  cw.ref[j] = kNone;
}

// cull key of a frontier entry: the box's entry distance (less its slack) clamped to >= 0 — a negative key passes every
// cull test, as 0 does — with the ray's tag in the low mantissa bits (clearing them only lowers a non-negative float:
// conservative)
PTB_DEV uint32_t coop_key(float tkey, uint32_t j) { return (__float_as_uint(fmaxf(tkey, 0.0f)) & ~kCoopTagMask) | j; }

// All 32 lanes call this together. Rays 0 .. n_rays-1 of `cw` are set (coop_set_ray) and visible (__syncwarp by the
// caller). On return best / ref hold the answers: closest hit = (t, ref) or ref == kNone for a miss; any-hit:
// ref != kNone iff a primitive other than `exclude` lies at 0 < t < tmax.
template <bool ANYHIT>
PTB_DEV void coop_trace(const DevScene& sc, CoopWarp& cw, uint32_t n_rays) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t top = 0;
  if (sc.n_prims) {
    if (lane < n_rays) cw.stack[lane] = make_uint2(0u, lane);  // the root, key 0
    top = n_rays;
  }
  __syncwarp();
#ifdef PTB_TAIL_STATS
  if (lane == 0u) PTB_TS(5, 1);
  unsigned long long ts_t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_t0));
#endif
  while (top) {
    const uint32_t n = top > kCoopSoft ? 1u : (top < 32u ? top : 32u);
    const uint32_t base = top - n;
    bool has = lane < n;
    uint2 e = make_uint2(0u, 0u);
    if (has) e = cw.stack[base + lane];
    top = base;
    __syncwarp();  // every pop has been read before any push lands on it
    const uint32_t j = e.y & kCoopTagMask;
    float bt = 0.0f;
    if (has) {
      bt = __uint_as_float((uint32_t)(cw.best[j] >> 32));
      has = __uint_as_float(e.y & ~kCoopTagMask) <= bt;
    }
#ifdef PTB_TAIL_STATS
    {
      const uint32_t mh = __ballot_sync(0xffffffffu, has);
      if (lane == 0u) { PTB_TS(6, 1); PTB_TS(7, __popc(mh)); PTB_TS(12, n - __popc(mh)); }
    }
#endif
    const bool leaf = has && (e.x & PTB_LEAF_BIT);
    const bool node = has && !leaf;
    // ---- loads of both kinds first, so that their latencies overlap
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
    uint4 qa = make_uint4(0u, 0u, 0u, 0u), qb = qa;
    uint32_t orig = 0u;
    if (node) {
#if PTB_QNODES
      ldg256u(sc.qnodes + 2u * (size_t)e.x, qa, qb);
#else
      float4 d;
      ldg256(sc.nodes + e.x, a, b);
      ldg256(reinterpret_cast<const float4*>(sc.nodes + e.x) + 2, c, d);
      qb.z = __float_as_uint(d.x);
      qb.w = __float_as_uint(d.y);
#endif
    } else if (leaf) {
      orig = __ldg(sc.slot_prim + (e.x & kSlotMask));  // tie-break id, fetched beside the geometry rather than after the test
      const float4* g = sc.geom + 3u * (size_t)(e.x & kSlotMask);
      a = __ldg(g);
      if (!(e.x & kSphereBit)) { b = __ldg(g + 1); c = __ldg(g + 2); }
    }
    const CoopRay& r = cw.ray[j];
    uint2 p0 = make_uint2(0u, 0u), p1 = p0;
    uint32_t cnt = 0;
    unsigned long long packed = 0ull;
    bool candidate = false;
    if (node) {
      float tl, tr;
#if PTB_QNODES
      QSlabRay s;
      s.a = mk(r.ax, r.ay, r.az);
      s.b_lo = mk(r.lx, r.ly, r.lz);
      s.b_hi = mk(r.hx, r.hy, r.hz);
      s.sn_x = (r.swap_xz & 2u) ? kSelHigh : kSelLow;
      s.sn_y = (r.swap_xz & 4u) ? kSelHigh : kSelLow;
      s.sn_z = (r.swap_xz & 8u) ? kSelHigh : kSelLow;
      const bool hl = box_entry_q(qa.x, qa.y, qa.z, s, bt, tl);
      const bool hr = box_entry_q(qa.w, qb.x, qb.y, s, bt, tr);
#else
      SlabRay s;
      s.dinv = mk(r.ax, r.ay, r.az);
      s.c_lo = mk(r.lx, r.ly, r.lz);
      s.c_hi = mk(r.hx, r.hy, r.hz);
      const bool hl = box_entry(a.x, a.y, a.z, a.w, b.x, b.y, s, bt, tl);
      const bool hr = box_entry(b.z, b.w, c.x, c.y, c.z, c.w, s, bt, tr);
#endif
      const uint32_t cl = qb.z, cr = qb.w;
      if (hl && hr) {
        const bool right_first = tr < tl;
        p0 = right_first ? make_uint2(cl, coop_key(tl, j)) : make_uint2(cr, coop_key(tr, j));  // far child: below
        p1 = right_first ? make_uint2(cr, coop_key(tr, j)) : make_uint2(cl, coop_key(tl, j));  // near child: on top
        cnt = 2u;
      } else if (hl || hr) {
        p0 = hl ? make_uint2(cl, coop_key(tl, j)) : make_uint2(cr, coop_key(tr, j));
        cnt = 1u;
      }
    } else if (leaf) {
      const uint32_t slot = e.x & kSlotMask;
      if (!ANYHIT || slot != r.exclude) {
        Ray ray;
        ray.o = mk(r.ox, r.oy, r.oz);
        ray.d = mk(r.dx, r.dy, r.dz);
        ray.shear = mk(r.shx, r.shy, r.shz);
        ray.dinv = mk(0.0f, 0.0f, 0.0f);
        ray.swap_xz = (r.swap_xz & 1u) != 0u;
        const float t = (e.x & kSphereBit) ? sphere_t(ray, from4(a), a.w) : triangle_t(ray, from4(a), from4(b), from4(c));
        if (t > 0.0f && (ANYHIT ? t < bt : t <= bt)) {
          packed = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)orig;
          candidate = packed < atomicMin(&cw.best[j], packed);
        }
      }
    }
    // ---- push the surviving children (a lane's two entries stay adjacent, lanes in order)
    const uint32_t m1 = __ballot_sync(0xffffffffu, cnt >= 1u), m2 = __ballot_sync(0xffffffffu, cnt == 2u);
    const uint32_t pos = top + (uint32_t)__popc(m1 & lt) + (uint32_t)__popc(m2 & lt);
    if (cnt >= 1u) cw.stack[pos] = p0;
    if (cnt == 2u) cw.stack[pos + 1u] = p1;
    top += (uint32_t)__popc(m1) + (uint32_t)__popc(m2);
    __syncwarp();  // pushes and atomicMin visible to the warp
    if (candidate && cw.best[j] == packed) cw.ref[j] = e.x & ~PTB_LEAF_BIT;  // the round's winner names the primitive
  }
  __syncwarp();
#ifdef PTB_TAIL_STATS
  unsigned long long ts_t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_t1));
  if (lane == 0u) PTB_TS(13, ts_t1 - ts_t0);
#endif
}

}  // namespace ptb
