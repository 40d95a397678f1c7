// Primitive and box intersection for ptb200 (device).
//   sphere_t / sphere_hit     : implementations/src/primitives/sphere.rs:34-105
//   triangle_t / triangle_hit : implementations/src/primitives/triangle.rs:105-216 (PBRT-v3 watertight test with the
//                               reference's own axis permutation, f64 fallback and conservative t bound)
//   box_entry                 : implementations/src/acceleration/aabb.rs:22-57 (slab test, far side widened by
//                               1+2*gamma(3)), extended with the entry distance for ordered, t-culled traversal
// Operation order matches the reference statement by statement (see ptb_common.cuh for the arithmetic contract).
#pragma once
#include "ptb_common.cuh"

namespace ptb {

struct HitRec {  // rt_core/src/primitive.rs:3-10 minus uv (never consumed by any texture)
  float t;
  v3 point, error, normal;
  bool out;
  float b1, b2;  // triangle.rs:149-151 barycentrics (closest-hit API output only)
};

// ---- sphere -------------------------------------------------------------------------------------------------------
// returns t > 0 or -1 (no hit / behind)
PTB_DEV float sphere_t(const Ray& ray, v3 center, float radius) {
  v3 deltap = center - ray.o;
  float ddp = dot(ray.d, deltap);
  float deltapdot = dot(deltap, deltap);
  v3 remedy = deltap - ddp * ray.d;
  float discriminant = radius * radius - dot(remedy, remedy);
  if (!(discriminant > 0.0f)) return -1.0f;
  float sqrt_val = sqrtf(discriminant);
  float q = ddp > 0.0f ? ddp + sqrt_val : ddp - sqrt_val;
  float t0 = q;
  float t1 = (deltapdot - radius * radius) / q;
  if (t1 < t0) { float tmp = t0; t0 = t1; t1 = tmp; }
  if (t0 > 0.0f) return t0;
  if (t1 <= 0.0f) return -1.0f;
  return t1;  // NaN propagates as in the reference (callers test `t > 0`)
}
PTB_DEV bool sphere_hit(const Ray& ray, v3 center, float radius, HitRec& h) {
  float t = sphere_t(ray, center, radius);
  if (t == -1.0f) return false;
  v3 point = ray_at(ray, t);
  v3 normal = (point - center) / radius;
  bool out = true;
  if (dot(normal, ray.d) > 0.0f) { out = false; normal = -normal; }
  h.t = t;
  h.point = point;
  h.error = kEpsRt * mk(1.0f, 1.0f, 1.0f);
  h.normal = normal;
  h.out = out;
  h.b1 = 0.0f;
  h.b2 = 0.0f;
  return true;
}

// ---- triangle -----------------------------------------------------------------------------------------------------
struct TriCore {
  float t, b0, b1, b2;
};
PTB_DEV bool triangle_core(const Ray& ray, v3 p0, v3 p1, v3 p2, TriCore& c) {
  v3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
  if (ray.swap_xz) {
    float tmp;
    tmp = p0t.x; p0t.x = p0t.z; p0t.z = tmp;
    tmp = p1t.x; p1t.x = p1t.z; p1t.z = tmp;
    tmp = p2t.x; p2t.x = p2t.z; p2t.z = tmp;
  }
  p0t.x += ray.shear.x * p0t.z; p0t.y += ray.shear.y * p0t.z;
  p1t.x += ray.shear.x * p1t.z; p1t.y += ray.shear.y * p1t.z;
  p2t.x += ray.shear.x * p2t.z; p2t.y += ray.shear.y * p2t.z;

  float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    e0 = (float)__dsub_rn(__dmul_rn((double)p1t.x, (double)p2t.y), __dmul_rn((double)p1t.y, (double)p2t.x));
    e1 = (float)__dsub_rn(__dmul_rn((double)p2t.x, (double)p0t.y), __dmul_rn((double)p2t.y, (double)p0t.x));
    e2 = (float)__dsub_rn(__dmul_rn((double)p0t.x, (double)p1t.y), __dmul_rn((double)p0t.y, (double)p1t.x));
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = e0 + e1 + e2;
  if (det == 0.0f) return false;

  p0t = p0t * ray.shear.z;
  p1t = p1t * ray.shear.z;
  p2t = p2t * ray.shear.z;

  float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if ((det < 0.0f && t_scaled >= 0.0f) || (det > 0.0f && t_scaled <= 0.0f)) return false;

  float inv_det = 1.0f / det;
  float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
  float t = inv_det * t_scaled;

  float max_z_t = cmax3(fabsf(p0t.z), fabsf(p1t.z), fabsf(p2t.z));
  float delta_z = gamma_n(3) * max_z_t;
  float max_x_t = cmax3(fabsf(p0t.x), fabsf(p1t.x), fabsf(p2t.x));
  float max_y_t = cmax3(fabsf(p0t.y), fabsf(p1t.y), fabsf(p2t.y));
  float delta_x = gamma_n(5) * (max_x_t + max_z_t);
  float delta_y = gamma_n(5) * (max_y_t + max_z_t);
  float delta_e = 2.0f * (gamma_n(2) * max_x_t * max_y_t + delta_y * max_x_t + delta_x * max_y_t);
  float max_e = cmax3(fabsf(e0), fabsf(e1), fabsf(e2));
  float delta_t = 3.0f * (gamma_n(3) * max_e * max_z_t + delta_e * max_z_t + delta_z * max_e) * fabsf(inv_det);
  if (t < delta_t) return false;
  c.t = t; c.b0 = b0; c.b1 = b1; c.b2 = b2;
  return true;
}
// returns t > 0 or -1
PTB_DEV float triangle_t(const Ray& ray, v3 p0, v3 p1, v3 p2) {
  TriCore c;
  if (!triangle_core(ray, p0, p1, p2, c)) return -1.0f;
  return c.t;
}
PTB_DEV bool triangle_hit(const Ray& ray, v3 p0, v3 p1, v3 p2, v3 n0, v3 n1, v3 n2, HitRec& h) {
  TriCore c;
  if (!triangle_core(ray, p0, p1, p2, c)) return false;
  v3 normal = c.b0 * n0 + c.b1 * n1 + c.b2 * n2;
  bool out = true;
  if (dot(normal, ray.d) > 0.0f) { normal = -normal; out = false; }  // utility/mod.rs:6-13 check_side
  float x_abs_sum = fabsf(c.b0 * p0.x) + fabsf(c.b1 * p1.x) + fabsf(c.b2 * p2.x);
  float y_abs_sum = fabsf(c.b0 * p0.y) + fabsf(c.b1 * p1.y) + fabsf(c.b2 * p2.y);
  float z_abs_sum = fabsf(c.b0 * p0.z) + fabsf(c.b1 * p1.z) + fabsf(c.b2 * p2.z);
  h.error = gamma_n(7) * mk(x_abs_sum, y_abs_sum, z_abs_sum) + gamma_n(6) * mk(c.b2 * p2.x, c.b2 * p2.y, c.b2 * p2.z);
  h.point = c.b0 * p0 + c.b1 * p1 + c.b2 * p2;
  h.t = c.t;
  h.normal = normal;
  h.out = out;
  h.b1 = c.b1;
  h.b2 = c.b2;
  return true;
}

// ---- slot-level dispatch ---------------------------------------------------------------------------------------------
// `ref` = leaf reference (kSphereBit | slot)
PTB_DEV float prim_t(const DevScene& sc, const Ray& ray, uint32_t ref) {
  const uint32_t slot = ref & kSlotMask;
  const float4* g = sc.geom + 3u * (size_t)slot;
  const float4 g0 = __ldg(g);
  if (ref & kSphereBit) return sphere_t(ray, from4(g0), g0.w);
  const float4 g1 = __ldg(g + 1), g2 = __ldg(g + 2);
  return triangle_t(ray, from4(g0), from4(g1), from4(g2));
}
PTB_DEV bool prim_hit(const DevScene& sc, const Ray& ray, uint32_t ref, HitRec& h) {
  const uint32_t slot = ref & kSlotMask;
  const float4* g = sc.geom + 3u * (size_t)slot;
  const float4 g0 = __ldg(g);
  if (ref & kSphereBit) return sphere_hit(ray, from4(g0), g0.w, h);
  const float4 g1 = __ldg(g + 1), g2 = __ldg(g + 2);
  const float4* nn = sc.normals + 3u * (size_t)slot;
  const float4 n0 = __ldg(nn), n1 = __ldg(nn + 1), n2 = __ldg(nn + 2);
  return triangle_hit(ray, from4(g0), from4(g1), from4(g2), from4(n0), from4(n1), from4(n2), h);
}

// ---- box -----------------------------------------------------------------------------------------------------------------
// Slab test for ordered, t-culled traversal; replaces AABB::does_int (implementations/src/acceleration/aabb.rs:22-57).
// The reference evaluates (plane - o) * d_inverse per plane, widens the far side by 1 + 2*gamma(3) and accepts when
// tmax > max(tmin, 0); it never culls by t. Here each plane distance is ONE fma, plane * dinv - o * dinv, with the
// product o * dinv hoisted per ray, and the near / far plane of each axis is picked by the sign of the direction. The
// hoisted product is rounded once, so a distance can be off by eps/2 * |o_i * dinv_i| beyond the reference's own two
// roundings; the test stays a superset of the reference's by (a) moving the near distances down / the far distances up
// by e_i = 2 * eps * |o_i * dinv_i| (folded into the fma addends, free) and (b) widening the far side by
// 1 + 4 * gamma(3). `tkey` = entry distance minus the cull slack: a primitive's COMPUTED t may fall slightly before its
// box's computed entry (radius-1000 ground spheres of the shipped scenes: |t error| ~ 1e-5, i.e. relative to the
// sphere's size, not to t), so a box is culled only when tkey > best_t. The slack is kCullSlack * max(|entry|, |exit|)
// of this box: the exit distance of a box the ray hits a primitive in is at least of the primitive's own scale.
// (A per-ray slack from the scene extent is cheaper but wrong-sized: grazing rays, |dinv| ~ 1e4, then walk unculled.)
// |dinv| is clamped to 1e30: plane * inf - o * inf would be NaN, which fminf / fmaxf ignore, and a ray whose slab is
// ignored walks a whole slice of the tree (camera rays with an exactly-zero x component are not rare: the f32 camera
// arithmetic cancels to 0 within an ulp of the image centre column — measured: 6761 nodes for one such ray, 7 ms for the
// lane). With the clamp an axis-parallel ray is inside the slab iff near <= o <= far up to e_i, like the reference.
constexpr float kCullSlack = 32.0f * kF32Eps;
constexpr float kSlabErr = 2.0f * kF32Eps;
constexpr float kDinvMax = 1.0e30f;
struct SlabRay {
  v3 dinv;  // 1 / d
  v3 c_lo;  // -(o * dinv) - e   addend of the near planes
  v3 c_hi;  // -(o * dinv) + e   addend of the far planes
};
PTB_HD SlabRay make_slab_ray(const Ray& ray) {
  SlabRay r;
  r.dinv = mk(fmaxf(fminf(ray.dinv.x, kDinvMax), -kDinvMax), fmaxf(fminf(ray.dinv.y, kDinvMax), -kDinvMax),
              fmaxf(fminf(ray.dinv.z, kDinvMax), -kDinvMax));
  const v3 od = ray.o * r.dinv;
  const v3 e = kSlabErr * vabs(od);
  r.c_lo = -od - e;
  r.c_hi = e - od;
  return r;
}
PTB_HD float fma_rn(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
PTB_HD bool box_entry(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, const SlabRay& r, float best_t,
                      float& tkey) {
  const float k = 1.0f + 4.0f * gamma_n(3);
  const bool sx = r.dinv.x < 0.0f, sy = r.dinv.y < 0.0f, sz = r.dinv.z < 0.0f;
  const float lox = fma_rn(sx ? mxx : mnx, r.dinv.x, r.c_lo.x), hix = fma_rn(sx ? mnx : mxx, r.dinv.x, r.c_hi.x);
  const float loy = fma_rn(sy ? mxy : mny, r.dinv.y, r.c_lo.y), hiy = fma_rn(sy ? mny : mxy, r.dinv.y, r.c_hi.y);
  const float loz = fma_rn(sz ? mxz : mnz, r.dinv.z, r.c_lo.z), hiz = fma_rn(sz ? mnz : mxz, r.dinv.z, r.c_hi.z);
  // 3-input min/max (FMNMX3 on sm_100a)
  const float tmin = fmaxf(fmaxf(lox, loy), loz);
  const float hmin = fminf(fminf(hix, hiy), hiz);
  tkey = fma_rn(-kCullSlack, fmaxf(fabsf(tmin), fabsf(hmin)), tmin);
  return hmin * k > fmaxf(tmin, 0.0f) && tkey <= best_t;
}

// ---- 16-bit boxes ---------------------------------------------------------------------------------------------------------
// The traversal kernels read 32-byte nodes: both children's boxes as 16-bit grid coordinates over the scene box
// (plane(q) = q_min + q * q_step, a box snapped OUTWARDS by lbvh_build.cu: k_quantise_nodes) and the two child
// references — ONE 32-byte load per node visit instead of two (ncu, round 2: the bounce launches of k_trace sit at 87 %
// of the L1 data-pipe wavefront peak, node loads being most of it; the node array shrinks from 64 to 32 MB per million
// primitives). On the CPU statement of the walk the snapped boxes cost +0.7 % node visits and +5 % primitive tests.
// Decoding costs nothing: a word holds one axis of one box (low half = min, high half = max); ONE byte permute places
// the half chosen by the direction's sign (the selector is a per-ray constant) under the exponent of 2^23, which makes
// the float f = 2^23 + q exactly, and the plane distance is ONE fma, f * A + B, with A = q_step * dinv and
// B = (q_min - o) * dinv - 2^23 * A hoisted per ray — the byte permute replaces the sign select of the f32 test.
// Error budget (eps = 2^-23). Exact: t* = (q_min - o) * dinv + q * a, a = q_step * dinv. Computed: A = a (1 + d1);
// f * A + B cancels the 2^23 * A term exactly (B was made from the same rounded A), leaving q * a * d1 <= |a| / 256;
// B itself is rounded at magnitude ~2^23 |A|: |A| / 2; X = (q_min - o) * dinv carries two roundings: eps |X|. Near
// distances are therefore moved down and far distances up by e = 0.51 |A| + 4 eps |X| (about half a grid step),
// the final rounding of the fma is covered as before by the far side's 1 + 4 gamma(3) and by the cull slack.
// |dinv| is clamped to 1e24 here (f * A stays finite for scenes up to 1e12 across).
constexpr float kDinvMaxQ = 1.0e24f;
constexpr uint32_t kSelLow = 0x7610u, kSelHigh = 0x7632u, kSelFlip = 0x0022u;  // prmt(word, 0x4B000000, sel) = 2^23 + half
struct QSlabRay {
  v3 a;       // q_step * dinv
  v3 b_lo;    // addend of the near planes
  v3 b_hi;    // addend of the far planes
  uint32_t sn_x, sn_y, sn_z;  // selector of the NEAR plane's half per axis (far = sn ^ kSelFlip)
};
PTB_HD QSlabRay make_qslab_ray(const float q_min[3], const float q_step[3], const Ray& ray) {
  QSlabRay r;
  const v3 dinv = mk(fmaxf(fminf(ray.dinv.x, kDinvMaxQ), -kDinvMaxQ), fmaxf(fminf(ray.dinv.y, kDinvMaxQ), -kDinvMaxQ),
                     fmaxf(fminf(ray.dinv.z, kDinvMaxQ), -kDinvMaxQ));
  r.a = mk(q_step[0] * dinv.x, q_step[1] * dinv.y, q_step[2] * dinv.z);
  const v3 x = mk((q_min[0] - ray.o.x) * dinv.x, (q_min[1] - ray.o.y) * dinv.y, (q_min[2] - ray.o.z) * dinv.z);
  const v3 b = mk(x.x - 8388608.0f * r.a.x, x.y - 8388608.0f * r.a.y, x.z - 8388608.0f * r.a.z);
  const v3 e = 0.51f * vabs(r.a) + (4.0f * kF32Eps) * vabs(x);
  r.b_lo = b - e;
  r.b_hi = b + e;
  r.sn_x = dinv.x < 0.0f ? kSelHigh : kSelLow;
  r.sn_y = dinv.y < 0.0f ? kSelHigh : kSelLow;
  r.sn_z = dinv.z < 0.0f ? kSelHigh : kSelLow;
  return r;
}
#ifdef __CUDACC__
// wx / wy / wz: the three axis words of one child box
// 2^23 + (the half of `w` named by `sel`), as a float: prmt straight from PTX (__byte_perm masks its selector first)
PTB_DEV float q_plane(uint32_t w, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0x4B000000u), "r"(sel));
  return __uint_as_float(d);
}
PTB_DEV bool box_entry_q(uint32_t wx, uint32_t wy, uint32_t wz, const QSlabRay& r, float best_t, float& tkey) {
  const float k = 1.0f + 4.0f * gamma_n(3);
  const float lox = __fmaf_rn(q_plane(wx, r.sn_x), r.a.x, r.b_lo.x), hix = __fmaf_rn(q_plane(wx, r.sn_x ^ kSelFlip), r.a.x, r.b_hi.x);
  const float loy = __fmaf_rn(q_plane(wy, r.sn_y), r.a.y, r.b_lo.y), hiy = __fmaf_rn(q_plane(wy, r.sn_y ^ kSelFlip), r.a.y, r.b_hi.y);
  const float loz = __fmaf_rn(q_plane(wz, r.sn_z), r.a.z, r.b_lo.z), hiz = __fmaf_rn(q_plane(wz, r.sn_z ^ kSelFlip), r.a.z, r.b_hi.z);
  const float tmin = fmaxf(fmaxf(lox, loy), loz);
  const float hmin = fminf(fminf(hix, hiy), hiz);
  tkey = __fmaf_rn(-kCullSlack, fmaxf(fabsf(tmin), fabsf(hmin)), tmin);
  return hmin * k > fmaxf(tmin, 0.0f) && tkey <= best_t;
}
PTB_DEV void ldg256u(const void* p, uint4& a, uint4& b) {
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
      : "l"(p));
}
#endif

// Which node format the binary tree's traversal kernels read. Measured on B200 (profiles/r2_sweeps.md §6): the 32-byte
// nodes take the bounce launches of k_trace from 87 % to 60 % of the L1 data-pipe peak and make them 4 - 6 % faster, the
// (ALU / issue bound) camera launch 8 % slower — the C3 step and the C5 rate end up within 1 % of the 64-byte nodes, for 32 MB
// more per million primitives (the f32 nodes stay: export, wide collapse). Built, bit-exact, parity-green, not the default.
#if PTB_QNODES  // (the switch itself: ptb_common.cuh)
typedef QSlabRay BinRayCtx;
#else
typedef SlabRay BinRayCtx;
#endif
PTB_HD BinRayCtx make_bin_ray(const DevScene& sc, const Ray& ray) {
#if PTB_QNODES
  return make_qslab_ray(sc.q_min, sc.q_step, ray);
#else
  return make_slab_ray(ray);
#endif
}

// box_entry with the direction's signs known at compile time (camera packets: the 32 rays of a pixel share an octant):
// the six near / far selects disappear. OCT bit 0 / 1 / 2 = the direction is negative along x / y / z.
template <int OCT>
PTB_HD bool box_entry_oct(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, const SlabRay& r, float best_t,
                          float& tkey) {
  const float k = 1.0f + 4.0f * gamma_n(3);
  constexpr bool sx = (OCT & 1) != 0, sy = (OCT & 2) != 0, sz = (OCT & 4) != 0;
  const float lox = fma_rn(sx ? mxx : mnx, r.dinv.x, r.c_lo.x), hix = fma_rn(sx ? mnx : mxx, r.dinv.x, r.c_hi.x);
  const float loy = fma_rn(sy ? mxy : mny, r.dinv.y, r.c_lo.y), hiy = fma_rn(sy ? mny : mxy, r.dinv.y, r.c_hi.y);
  const float loz = fma_rn(sz ? mxz : mnz, r.dinv.z, r.c_lo.z), hiz = fma_rn(sz ? mnz : mxz, r.dinv.z, r.c_hi.z);
  const float tmin = fmaxf(fmaxf(lox, loy), loz);
  const float hmin = fminf(fminf(hix, hiy), hiz);
  tkey = fma_rn(-kCullSlack, fmaxf(fabsf(tmin), fabsf(hmin)), tmin);
  return hmin * k > fmaxf(tmin, 0.0f) && tkey <= best_t;
}

#ifdef PTB_BOX_V1  // experiment: the previous sub+mul slab test with the per-box slack
PTB_DEV bool box_entry_v1(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, const Ray& ray, float best_t,
                          float& tkey) {
  const float k = 1.0f + 2.0f * gamma_n(3);
  const float ax = (mnx - ray.o.x) * ray.dinv.x, bx = (mxx - ray.o.x) * ray.dinv.x;
  const float ay = (mny - ray.o.y) * ray.dinv.y, by = (mxy - ray.o.y) * ray.dinv.y;
  const float az = (mnz - ray.o.z) * ray.dinv.z, bz = (mxz - ray.o.z) * ray.dinv.z;
  const float lox = fminf(ax, bx), hix = fmaxf(ax, bx);
  const float loy = fminf(ay, by), hiy = fmaxf(ay, by);
  const float loz = fminf(az, bz), hiz = fmaxf(az, bz);
  const float tmin = fmaxf(fmaxf(lox, loy), loz);
  const float hmin = fminf(fminf(hix, hiy), hiz);
  const float tmax = hmin * k;
  float m = fmaxf(fmaxf(fmaxf(hix, hiy), hiz), -fminf(fminf(lox, loy), loz));
  if (!(m < 3.0e38f)) {
    const float mx = fmaxf(hix, -lox), my = fmaxf(hiy, -loy), mz = fmaxf(hiz, -loz);
    m = fmaxf(mx < 3.0e38f ? mx : 0.0f, fmaxf(my < 3.0e38f ? my : 0.0f, mz < 3.0e38f ? mz : 0.0f));
  }
  tkey = tmin - kCullSlack * m;
  return tmax > fmaxf(tmin, 0.0f) && tkey <= best_t;
}
#endif

}  // namespace ptb
