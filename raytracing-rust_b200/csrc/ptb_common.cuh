// ptb200 device-side common definitions (sm_100a).
//
// Arithmetic contract: this library is compiled with -fmad=false and without --use_fast_math, so every
// `a*b+c` below stays two IEEE-754 roundings exactly like rustc emits for the reference (and like the CPU
// oracle built with -ffp-contract=off); `/` and sqrtf are correctly rounded. Hit/miss sign decisions of the
// watertight triangle test and the robust sphere quadratic therefore agree bit-for-bit with the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ptb200.h"

namespace ptb {

#define PTB_DEV __device__ __forceinline__
#define PTB_HD __host__ __device__ __forceinline__

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTau = 6.28318530717958647692f;
constexpr float kEpsRt = 3.0e-4f;          // rt_core/src/lib.rs:28  EPSILON
constexpr float kF32Eps = 1.1920929e-7f;   // f32::EPSILON
#ifndef PTB_QNODES
#define PTB_QNODES 0  // 1: the binary tree's traversal kernels read 32-byte nodes with 16-bit boxes (ptb_intersect.cuh)
#endif
constexpr uint32_t kSphereBit = 0x40000000u;  // device-internal: leaf reference points at a sphere slot
constexpr uint32_t kSlotMask = 0x3FFFFFFFu;
constexpr uint32_t kNone = 0xFFFFFFFFu;

// ---------------------------------------------------------------- vec3 (rt_core/src/vec.rs:108-248)
struct v3 {
  float x, y, z;
};
PTB_HD v3 mk(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
PTB_HD v3 operator+(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
PTB_HD v3 operator-(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
PTB_HD v3 operator*(v3 a, v3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
PTB_HD v3 operator/(v3 a, v3 b) { return mk(a.x / b.x, a.y / b.y, a.z / b.z); }
PTB_HD v3 operator*(v3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
PTB_HD v3 operator*(float s, v3 a) { return mk(s * a.x, s * a.y, s * a.z); }
PTB_HD v3 operator/(v3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
PTB_HD v3 operator-(v3 a) { return mk(-a.x, -a.y, -a.z); }
PTB_HD float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PTB_HD v3 cross(v3 a, v3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PTB_HD float mag_sq(v3 a) { return dot(a, a); }
PTB_HD float mag(v3 a) { return sqrtf(dot(a, a)); }
PTB_HD v3 normalised(v3 a) { return a / mag(a); }
PTB_HD v3 vabs(v3 a) { return mk(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// Rust f32::min/max return the non-NaN operand == fminf/fmaxf
PTB_HD float cmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
PTB_HD v3 vmin(v3 a, v3 b) { return mk(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PTB_HD v3 vmax(v3 a, v3 b) { return mk(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
PTB_HD v3 reflected(v3 v, v3 n) { return 2.0f * dot(v, n) * n - v; }  // vec.rs:203-205
PTB_HD bool contains_nan(v3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }
PTB_HD bool any_finite(v3 a) { return isfinite(a.x) || isfinite(a.y) || isfinite(a.z); }  // vec.rs:245-247 (Q5)
PTB_HD bool is_zero(v3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
PTB_HD v3 from4(float4 f) { return mk(f.x, f.y, f.z); }

// ---------------------------------------------------------------- Ray (rt_core/src/ray.rs:4-46)
struct Ray {
  v3 o, d, dinv, shear;
  bool swap_xz;  // Axis::swap_z applies to X- and Y-dominant directions alike (quirk Q1)
};
// `d` must already be normalised (Ray::new normalises once, at creation; see make_ray_from_raw)
PTB_HD Ray make_ray(v3 o, v3 d) {
  Ray r;
  r.o = o;
  r.d = d;
  const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  r.swap_xz = (ax > ay && ax > az) || (ay > az);
  v3 s = d;
  if (r.swap_xz) { float t = s.x; s.x = s.z; s.z = t; }
  r.shear = mk(-s.x / s.z, -s.y / s.z, 1.0f / s.z);
  r.dinv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  return r;
}
PTB_HD Ray make_ray_from_raw(v3 o, v3 dir) { return make_ray(o, dir / mag(dir)); }
PTB_HD v3 ray_at(const Ray& r, float t) { return r.o + r.d * t; }

// ---------------------------------------------------------------- utility/mod.rs:51-117
PTB_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PTB_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
PTB_HD float next_float(float f) {
  if (isinf(f) && f > 0.0f) return f;
  if (f == -0.0f) f = 0.0f;
  return u2f(f >= 0.0f ? f2u(f) + 1u : f2u(f) - 1u);
}
PTB_HD float previous_float(float f) {
  if (isinf(f) && f < 0.0f) return f;
  if (f == 0.0f) f = -0.0f;
  return u2f(f <= 0.0f ? f2u(f) + 1u : f2u(f) - 1u);
}
PTB_HD float gamma_n(uint32_t n) {
  float nm = (float)n * 0.5f * kF32Eps;
  return nm / (1.0f - nm);
}
PTB_HD v3 offset_ray(v3 origin, v3 normal, v3 error, bool is_brdf) {
  float offset_val = dot(vabs(normal), error);
  v3 offset = offset_val * normal;
  if (!is_brdf) offset = -offset;
  v3 p = origin + offset;
  p.x = offset.x > 0.0f ? next_float(p.x) : previous_float(p.x);
  p.y = offset.y > 0.0f ? next_float(p.y) : previous_float(p.y);
  p.z = offset.z > 0.0f ? next_float(p.z) : previous_float(p.z);
  return p;
}
// utility/coord.rs:10-30 — to_coord(v) of the frame built from z
PTB_HD v3 onb_to_world(v3 z, v3 v) {
  v3 x;
  if (fabsf(z.x) > fabsf(z.y)) x = mk(-z.z, 0.0f, z.x) / sqrtf(z.x * z.x + z.z * z.z);
  else x = mk(0.0f, z.z, -z.y) / sqrtf(z.y * z.y + z.z * z.z);
  v3 y = cross(x, z);
  return v.x * x + v.y * y + v.z * z;
}
PTB_HD float power_heuristic(float pdf_a, float pdf_b) {  // rt_core/src/lib.rs:36-40
  float a_sq = pdf_a * pdf_a;
  return a_sq / (a_sq + pdf_b * pdf_b);
}

// ---------------------------------------------------------------- counter-based RNG
// Philox4x32-10; counter = (pixel, sample, (depth << 8) | purpose, block), key = seed. Shared with the oracle
// (oracle/ref_rng.hpp) so that both draw the same numbers for the same (pixel, sample, depth, purpose).
enum : uint32_t { RNG_JITTER = 0, RNG_NEE = 1, RNG_SCATTER = 2, RNG_RR = 3 };
PTB_HD uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
PTB_HD float u32_to_unit(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }
PTB_HD uint32_t rng_below(uint32_t u, uint32_t n) {
#ifdef __CUDA_ARCH__
  return __umulhi(u, n);
#else
  return (uint32_t)(((uint64_t)u * n) >> 32);
#endif
}

// ---------------------------------------------------------------- wide loads / stores
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
// ---------------------------------------------------------------- device scene
struct __align__(16) BvhNode {  // 64 B, same memory layout as ptb_bvh_node
  float4 n0;  // lmin.x lmin.y lmin.z lmax.x
  float4 n1;  // lmax.y lmax.z rmin.x rmin.y
  float4 n2;  // rmin.z rmax.x rmax.y rmax.z
  uint4 n3;   // left, right, parent, pad   (refs: bit31 leaf, bit30 sphere, low bits slot)
};
static_assert(sizeof(BvhNode) == 64 && sizeof(ptb_bvh_node) == 64, "node layout");
// Compressed 8-wide node (ptb_cwbvh.cuh has the layout's story; oracle/cwbvh_ref.hpp the CPU definition): 96 bytes.
struct __align__(32) CwNode {
  float p[3];            // grid origin = node box min
  uint32_t e_imask;      // biased cell exponents x | y << 8 | z << 16, imask << 24 (bit s: slot s holds an inner child)
  uint32_t child_base;   // index of the first inner child
  uint32_t prim_base;    // first slot of the node's primitives
  uint32_t meta[2];      // per slot one byte: 0 empty | inner 0x20 | (24 + s) | leaf group (2^count - 1) << 5 | offset
  uint32_t q[12];        // qlo x, y, z then qhi x, y, z: 8 bytes (slots 0..7) each
  uint32_t pad[4];
};
static_assert(sizeof(CwNode) == 96, "wide node layout");

struct DevMaterial {
  uint32_t kind, tex;
  float param, metallic;
  float ior[3];
  uint32_t _pad;
};
struct DevTexture {  // 48 B
  uint32_t kind;
  float a[3], b[3];
  uint32_t data_off;       // first word of this texture's bulk data in DevScene::tex_data (image pixels / perlin tables)
  uint32_t width, height;  // ImageTexture only
  uint32_t _pad[2];
};

// ---------------------------------------------------------------- textures (implementations/src/textures/mod.rs)
// One definition for the device (k_shade) and the host (sky table at commit). `words` = the texture's bulk data.
struct TexWords {
  const float* p;
  PTB_HD float f(uint32_t i) const {
#ifdef __CUDA_ARCH__
    return __ldg(p + i);
#else
    return p[i];
#endif
  }
  PTB_HD uint32_t u(uint32_t i) const { return f2u(f(i)); }
};
PTB_HD int32_t sat_i32(float f) {  // Rust `as i32`: saturating, NaN -> 0
  if (f != f) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return (-2147483647 - 1);
  return (int32_t)f;
}
PTB_HD uint32_t sat_index(float f, uint32_t hi) {  // Rust `as usize` then clamp(0, hi)
  if (!(f > 0.0f)) return 0u;
  if (f >= 4294967040.0f) return hi;
  const uint32_t i = (uint32_t)f;
  return i > hi ? hi : i;
}
// Perlin::noise + trilinear_lerp (textures/mod.rs:114-139, 161-179); words = 256 ran scalars | perm_x | perm_y | perm_z
PTB_HD float perlin_noise(TexWords words, v3 point) {
  const float fx = floorf(point.x), fy = floorf(point.y), fz = floorf(point.z);
  const float u = point.x - fx, v = point.y - fy, w = point.z - fz;
  const uint32_t i = (uint32_t)sat_i32(fx), j = (uint32_t)sat_i32(fy), k = (uint32_t)sat_i32(fz);
  const float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
  float value = 0.0f;
#pragma unroll
  for (uint32_t index = 0; index < 8u; ++index) {
    const uint32_t di = index >> 2, dj = (index >> 1) & 1u, dk = index & 1u;
    const uint32_t h = words.u(256u + ((i + di) & 255u)) ^ words.u(512u + ((j + dj) & 255u)) ^ words.u(768u + ((k + dk) & 255u));
    const float r = words.f(h & 255u);
    const float fi = (float)di, fj = (float)dj, fk = (float)dk;
    const v3 c = r * mk(1.0f, 1.0f, 1.0f);
    value += (fi * uu + (1.0f - fi) * (1.0f - uu)) * (fj * vv + (1.0f - fj) * (1.0f - vv)) * (fk * ww + (1.0f - fk) * (1.0f - ww)) *
             dot(c, mk(u - fi, v - fj, w - fk));
  }
  return value;
}
PTB_HD v3 texture_eval(uint32_t kind, v3 a, v3 b, uint32_t width, uint32_t height, TexWords words, v3 direction, v3 point) {
  if (kind == PTB_TEX_SOLID) return a;  // :193-200
  if (kind == PTB_TEX_LERP) {           // :283-291
    const float tt = direction.z * 0.5f + 0.5f;
    return a * tt + b * (1.0f - tt);
  }
  if (kind == PTB_TEX_CHECKERED) {      // :61-73
    const float sign = sinf(10.0f * point.x) * sinf(10.0f * point.y) * sinf(10.0f * point.z);
    return sign > 0.0f ? a : b;
  }
  if (kind == PTB_TEX_IMAGE) {          // :248-262: lat-long lookup by direction, nearest texel, dim = (w-1, h-1)
    const float phi = atan2f(direction.y, direction.x) + kPi;
    const float theta = acosf(direction.z);
    const float uvx = phi / (2.0f * kPi), uvy = theta / kPi;
    const uint32_t dim0 = width - 1u, dim1 = height - 1u;
    const uint32_t x_pixel = sat_index((float)dim0 * uvx, dim0), y_pixel = sat_index((float)dim1 * uvy, dim1);
    const uint32_t index = 3u * (y_pixel * (dim0 + 1u) + x_pixel);
    return mk(words.f(index), words.f(index + 1u), words.f(index + 2u));
  }
  if (kind == PTB_TEX_PERLIN) return 0.5f * mk(1.0f, 1.0f, 1.0f) * (1.0f + perlin_noise(words, point));  // :171-179
  return mk(1.0f, 1.0f, 1.0f);          // trait default :10-12
}

struct DevScene {
  // geometry in Morton-sorted slot order, 3 x float4 per slot:
  //   triangle: p0.xyz|0, p1.xyz|0, p2.xyz|0      sphere: c.xyz|r, 0, 0
  const float4* geom;
  const float4* normals;      // 3 x float4 per slot (triangles only): n0, n1, n2
  const uint32_t* slot_prim;  // slot -> original primitive id (loader order)
  const uint32_t* slot_mat;   // slot -> (material kind << 24) | material index
  const BvhNode* nodes;       // binary LBVH (leaf references: Morton position), f32 boxes: the build's product, exported as is
  const uint4* qnodes;        // the same tree as the traversal kernels read it: 32-byte nodes, 16-bit boxes (2 x uint4 per node)
  float q_min[3], q_step[3];  // their grid: plane(q) = q_min + q * q_step, exactly
  const CwNode* cw_nodes;     // compressed 8-wide tree over the same primitives; geometry is then in ITS primitive order
  const DevMaterial* materials;
  const DevTexture* textures;
  const float* tex_data;      // bulk texture data (image pixels, perlin tables), see DevTexture::data_off
  const uint32_t* lights;     // slots (with kSphereBit where applicable) whose material is_light()
  uint32_t n_prims, n_lights;
  // sky (implementations/src/sky.rs): Emit(texture, 1.0) + lat-long Distribution2D
  uint32_t sky_tex, sky_rx, sky_ry;
  const float* sky_ycdf;  // ry + 1
  const float* sky_ypdf;  // ry
  const float* sky_xcdf;  // ry * (rx + 1)
  const float* sky_xpdf;  // ry * rx
  // camera (implementations/src/camera.rs:57-63)
  v3 cam_origin, cam_lower_left, cam_horizontal, cam_vertical;
  // traversal scheduling knobs (ptb_traverse.cuh); defaults set in ptb_create, PTB_TRACE_BURST / PTB_TRACE_FETCH override
  int trace_burst, trace_fetch_threshold, trace_prim_bias;
};

}  // namespace ptb
