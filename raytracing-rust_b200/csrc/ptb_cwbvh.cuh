// Compressed 8-wide BVH traversal for ptb200 (device) — the wide tree of the north star ("optional SAH collapse to a wide
// BVH"). The reference has no counterpart: what it fixes is the result of check_hit
// (implementations/src/acceleration/mod.rs:265-298: minimum t > 0 over all primitives) and of check_hit_index (:226-263).
//
// Node (96 bytes = three 32-byte loads; Ylitie, Karras, Laine 2017, "Efficient incoherent ray traversal on GPUs through
// compressed wide BVHs"): grid origin + one power-of-two cell size per axis, eight child boxes quantised to 8 bits per
// plane, inner children contiguous (child index = child_base + popcount(imask below the slot)), the node's primitives one
// contiguous block of at most 24 (8 leaf groups of <= 3), one `meta` byte per slot. Children sit in the slot whose bits say
// on which side of the node centre they lie, so a ray with sign octant o visits its hit children in descending
// (slot ^ o ^ 7): near side first without sorting distances. The tree itself (SAH-driven greedy collapse of the LBVH, slot
// assignment, quantisation, layout) is built by cwbvh_build.cu; its CPU definition, which the device reproduces bit for
// bit, and the CPU statement of THIS traversal (equal node / primitive counts) are oracle/cwbvh_ref.hpp.
//
// Per-lane state: one node group (child_base, hit bits 24..31 | imask) and one primitive group (prim_base, hit bits 0..23)
// in registers, node groups with hits left on a local-memory stack of 8-byte entries (one push per node at most — a third
// of the binary walk's stack traffic). The warp-synchronous phase machine of ptb_traverse.cuh drives it: node phase for the
// lanes whose node group has a hit and whose primitive group is empty, primitive phase (one primitive per lane) otherwise.
#pragma once
#include "ptb_intersect.cuh"

namespace ptb {

constexpr int kCwStackDepth = 64;  // one push per wide level at most; the LBVH is at most 62 levels deep

struct CwRay {
  v3 dinv, neg_od;  // clamped 1 / d;  -(o * dinv)
  v3 ed, eo;        // 4 eps |dinv|;   4 eps |o * dinv|   (rounding of the per-node fold p * dinv - o * dinv)
  uint32_t oinv;    // sign octant ^ 7
  uint32_t neg;     // bit a: direction component a is negative (near plane = the box's max plane)
};
PTB_DEV CwRay make_cw_ray(const Ray& ray) {
  CwRay r;
  r.dinv = mk(fmaxf(fminf(ray.dinv.x, kDinvMax), -kDinvMax), fmaxf(fminf(ray.dinv.y, kDinvMax), -kDinvMax),
              fmaxf(fminf(ray.dinv.z, kDinvMax), -kDinvMax));
  const v3 od = ray.o * r.dinv;
  r.neg_od = -od;
  r.ed = (4.0f * kF32Eps) * vabs(r.dinv);
  r.eo = (4.0f * kF32Eps) * vabs(od);
  r.neg = (r.dinv.x < 0.0f ? 1u : 0u) | (r.dinv.y < 0.0f ? 2u : 0u) | (r.dinv.z < 0.0f ? 4u : 0u);
  r.oinv = r.neg ^ 7u;
  return r;
}

struct CwState {
  uint32_t gx, gy;    // node group: child_base, hit bits 24..31 | imask
  uint32_t tx, ty;    // primitive group: prim_base, hit bits 0..23
  int sp;
  float best_t;       // closest hit so far (closest-hit) / tmax (any-hit)
  uint32_t best_ref;  // closest-hit: winning slot | kSphereBit; any-hit: kNone = unoccluded, 0 = occluded
  PTB_DEV bool node_ready() const { return ty == 0u && (gy & 0xFF000000u) != 0u; }
  PTB_DEV bool leaf_ready() const { return ty != 0u; }
  PTB_DEV bool done() const { return ty == 0u && (gy & 0xFF000000u) == 0u; }
};
PTB_DEV void cw_init(CwState& s, const CwRay& r, uint32_t n_prims, float tmax) {
  // the root enters as a node group of one: child_base 0, imask bit of slot 0, hit bit of slot 0
  s.gx = 0u;
  s.gy = n_prims ? ((1u << (24u + r.oinv)) | 1u) : 0u;
  s.tx = s.ty = 0u;
  s.sp = 0;
  s.best_t = tmax;
  s.best_ref = kNone;
}
PTB_DEV void cw_pop(CwState& s, const uint2* stack) {
  if (s.sp > 0) {
    const uint2 e = stack[--s.sp];
    s.gx = e.x;
    s.gy = e.y;
  } else {
    s.gy = 0u;
  }
}

// Byte k of `w` as the float 1 + q * 2^-15: ONE byte permute drops the byte into mantissa bits 8..15 of 1.0f. (An integer
// to float conversion is a quarter-rate instruction, and 48 of them per node would be the kernel's bottleneck; the usual
// 2^23 + q trick would fold a constant of 2^23 cells into the plane offsets and lose half a cell to rounding — with
// 1 + q 2^-15 the folded constant is 2^15 cells and the rounding 1/512 of a cell, which `err` absorbs.)
PTB_DEV float cw_plane(uint32_t w, int k) { return __uint_as_float(__byte_perm(w, 0x3F800000u, 0x7604u | ((uint32_t)k << 4))); }

// Four children (one byte lane each of the six plane words) against the ray; ORs their bits into `mask`.
// ad / adk: cell * dinv * 2^15 (near) and that times k (far); clo / chik: the folded plane offsets minus ad / adk.
PTB_DEV void cw_test4(uint32_t qnx, uint32_t qny, uint32_t qnz, uint32_t qfx, uint32_t qfy, uint32_t qfz, uint32_t meta4,
                      const CwRay& r, v3 ad, v3 adk, v3 clo, v3 chik, float bound, uint32_t& mask) {
  // per byte: inner children have meta bits 4 and 3 set (index 24..31); their bit index is octant-ordered
  const uint32_t is_inner = (meta4 & (meta4 << 1)) & 0x10101010u;
  const uint32_t inner_ff = (is_inner >> 4) * 0xFFu;
  const uint32_t bit_index4 = (meta4 ^ ((r.oinv * 0x01010101u) & inner_ff)) & 0x1F1F1F1Fu;
  const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float tnx = fma_rn(cw_plane(qnx, k), ad.x, clo.x), tny = fma_rn(cw_plane(qny, k), ad.y, clo.y),
                tnz = fma_rn(cw_plane(qnz, k), ad.z, clo.z);
    const float tfx = fma_rn(cw_plane(qfx, k), adk.x, chik.x), tfy = fma_rn(cw_plane(qfy, k), adk.y, chik.y),
                tfz = fma_rn(cw_plane(qfz, k), adk.z, chik.z);
    const float lo = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), 0.0f);
    const float hi = fminf(fminf(tfx, tfy), fminf(tfz, bound));
    if (lo <= hi) mask |= ((child_bits4 >> (8 * k)) & 0xFFu) << ((bit_index4 >> (8 * k)) & 0xFFu);
  }
}

// One node step: take the nearest (octant order) hit child of the lane's node group, fetch it, test its eight children.
template <bool COUNT>
PTB_DEV void cw_node_step(const CwNode* __restrict__ nodes, const CwRay& r, CwState& s, uint2* stack, uint32_t& n_nodes) {
  const uint32_t bit = 31u - (uint32_t)__clz((int)s.gy);
  s.gy &= ~(1u << bit);
  if (s.gy & 0xFF000000u) stack[s.sp++] = make_uint2(s.gx, s.gy);
  const uint32_t slot = (bit - 24u) ^ r.oinv;
  const uint32_t rel = (uint32_t)__popc(s.gy & 0xFFu & ((1u << slot) - 1u));
  const float4* np = reinterpret_cast<const float4*>(nodes + (s.gx + rel));
  float4 h0, h1, a0, a1, b0, b1;
  ldg256(np, h0, h1);       // p.xyz e_imask | child_base prim_base meta[2]
  ldg256(np + 2, a0, a1);   // qlo x (2 words) qlo y (2) | qlo z (2) qhi x (2)
  ldg256(np + 4, b0, b1);   // qhi y (2) qhi z (2) | pad
  if (COUNT) ++n_nodes;
  const uint32_t e_imask = __float_as_uint(h0.w);
  const float k = 1.0f + 8.0f * gamma_n(3);
  // cell sizes times 2^15 (the plane floats are 1 + q 2^-15): exponent + 15, exact
  const v3 cell = mk(__uint_as_float(((e_imask & 0xFFu) + 15u) << 23), __uint_as_float((((e_imask >> 8) & 0xFFu) + 15u) << 23),
                     __uint_as_float((((e_imask >> 16) & 0xFFu) + 15u) << 23));
  const v3 ad = cell * r.dinv;
  const v3 adk = ad * k;
  const v3 base = mk(fma_rn(h0.x, r.dinv.x, r.neg_od.x), fma_rn(h0.y, r.dinv.y, r.neg_od.y), fma_rn(h0.z, r.dinv.z, r.neg_od.z));
  // rounding of the fold (4 eps (|p dinv| + |o dinv|)) + of the 2^15-cell offset (2^-9 cells, taken as 2^-23 |ad|)
  const float q8 = 1.0f / 8388608.0f;
  const v3 err = mk(fma_rn(fabsf(h0.x), r.ed.x, fma_rn(fabsf(ad.x), q8, r.eo.x)), fma_rn(fabsf(h0.y), r.ed.y, fma_rn(fabsf(ad.y), q8, r.eo.y)),
                    fma_rn(fabsf(h0.z), r.ed.z, fma_rn(fabsf(ad.z), q8, r.eo.z)));
  const v3 clo = (base - err) - ad;
  const v3 chik = (base + err) * k - adk;
  // the largest plane distance the node can produce: |base| + err + 255 cells
  const float c255 = 255.0f / 32768.0f;
  const float reach = fmaxf(fmaxf(fabsf(base.x) + fma_rn(c255, fabsf(adk.x), err.x), fabsf(base.y) + fma_rn(c255, fabsf(adk.y), err.y)),
                            fabsf(base.z) + fma_rn(c255, fabsf(adk.z), err.z));
  const float bound = fma_rn(32.0f * kF32Eps, reach, s.best_t);
  // near / far plane words per axis (slots 0..3 and 4..7)
  const uint32_t lx0 = __float_as_uint(a0.x), lx1 = __float_as_uint(a0.y), ly0 = __float_as_uint(a0.z), ly1 = __float_as_uint(a0.w);
  const uint32_t lz0 = __float_as_uint(a1.x), lz1 = __float_as_uint(a1.y), hx0 = __float_as_uint(a1.z), hx1 = __float_as_uint(a1.w);
  const uint32_t hy0 = __float_as_uint(b0.x), hy1 = __float_as_uint(b0.y), hz0 = __float_as_uint(b0.z), hz1 = __float_as_uint(b0.w);
  const bool nx = (r.neg & 1u) != 0u, ny = (r.neg & 2u) != 0u, nz = (r.neg & 4u) != 0u;
  uint32_t mask = 0u;
  cw_test4(nx ? hx0 : lx0, ny ? hy0 : ly0, nz ? hz0 : lz0, nx ? lx0 : hx0, ny ? ly0 : hy0, nz ? lz0 : hz0, __float_as_uint(h1.z),
           r, ad, adk, clo, chik, bound, mask);
  cw_test4(nx ? hx1 : lx1, ny ? hy1 : ly1, nz ? hz1 : lz1, nx ? lx1 : hx1, ny ? ly1 : hy1, nz ? lz1 : hz1, __float_as_uint(h1.w),
           r, ad, adk, clo, chik, bound, mask);
  s.gx = __float_as_uint(h1.x);
  s.gy = (mask & 0xFF000000u) | (e_imask >> 24);
  s.tx = __float_as_uint(h1.y);
  s.ty = mask & 0x00FFFFFFu;
  if (s.ty == 0u && !(s.gy & 0xFF000000u)) cw_pop(s, stack);
}

// t > 0 or -1 of the primitive in `slot`; `sphere` reports its kind (a sphere's record carries its radius in g0.w, a
// triangle's g0.w is 0: the leaf groups of the wide tree mix both, so the kind travels with the data, not the reference)
PTB_DEV float cw_prim_t(const DevScene& sc, const Ray& ray, uint32_t slot, bool& sphere) {
  const float4* g = sc.geom + 3u * (size_t)slot;
  const float4 g0 = __ldg(g);
  sphere = g0.w != 0.0f;
  if (sphere) return sphere_t(ray, from4(g0), g0.w);
  const float4 g1 = __ldg(g + 1), g2 = __ldg(g + 2);
  return triangle_t(ray, from4(g0), from4(g1), from4(g2));
}

// One primitive step: the lowest set bit of the lane's primitive group.
template <bool ANYHIT, bool COUNT>
PTB_DEV void cw_prim_step(const DevScene& sc, const Ray& ray, CwState& s, const uint2* stack, uint32_t exclude, uint32_t& n_prims) {
  const uint32_t bit = (uint32_t)__ffs((int)s.ty) - 1u;
  s.ty &= s.ty - 1u;
  const uint32_t slot = s.tx + bit;
  if (!ANYHIT || slot != exclude) {
    bool sphere;
    const float t = cw_prim_t(sc, ray, slot, sphere);
    if (COUNT) ++n_prims;
    if (ANYHIT) {
      if (t > 0.0f && t < s.best_t) {  // blocker found: stop
        s.best_ref = 0u;
        s.gy = s.ty = 0u;
        s.sp = 0;
        return;
      }
    } else if (t > 0.0f) {
      const uint32_t ref = slot | (sphere ? kSphereBit : 0u);
      if (t < s.best_t) {
        s.best_t = t;
        s.best_ref = ref;
      } else if (t == s.best_t) {  // exact tie: the lower ORIGINAL primitive id wins (quirk Q2)
        if (__ldg(sc.slot_prim + slot) < __ldg(sc.slot_prim + (s.best_ref & kSlotMask))) s.best_ref = ref;
      }
    }
  }
  if (s.ty == 0u && !(s.gy & 0xFF000000u)) cw_pop(s, stack);
}

}  // namespace ptb
