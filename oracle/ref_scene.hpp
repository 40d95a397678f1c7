// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_math.hpp header).
// Textures, materials, primitives, sky and the piecewise-constant distributions of the reference.
//
// Parity status: lambertian sample/pdf and Distribution1D/2D are pinned by the reference's chi-squared
// tests (statistics/bxdfs/lambertian.rs:30-48, statistics/distributions.rs:186-300), restated in
// tests/test_oracle_kats.py. Sphere/triangle intersection, materials and the sky have NO golden vectors in
// the reference ("parity unpinned" there); they are cross-checked against analytic answers instead.
#pragma once
#include <cstdint>
#include <vector>

#include "../include/ptb200.h"
#include "ref_math.hpp"
#include "ref_rng.hpp"

namespace ref {

// implementations/src/utility/mod.rs:41-44
static inline Float random_float() { return g_rng.next_float01(); }

// utility/mod.rs:15-25 — rejection loop in the unit ball, then normalise
static inline Vec3 random_unit_vector() {
  Float x = 1.0f, y = 1.0f, z = 1.0f;
  while (x * x + y * y + z * z > 1.0f) {
    x = g_rng.next_pm1();
    y = g_rng.next_pm1();
    z = g_rng.next_pm1();
  }
  return normalised(Vec3(x, y, z));
}

// ------------------------------------------------------------------ textures
// implementations/src/textures/mod.rs
struct Texture {
  uint32_t kind;
  Vec3 a, b;
  // ImageTexture: `data` = width*height RGB f32 rows (textures/mod.rs:202-245; dim = (width-1, height-1)).
  // Perlin: `data` = 256 ran_vecs scalars (every ran_vec is r*Vec3::one(), textures/mod.rs:96-99), `perm` = perm_x|perm_y|perm_z.
  const Float* data = nullptr;
  const uint32_t* perm = nullptr;
  uint32_t width = 0, height = 0;

  // textures/mod.rs:114-139, 161-179
  Float perlin_noise(const Vec3& point) const {
    Float fx = std::floor(point.x), fy = std::floor(point.y), fz = std::floor(point.z);
    Float u = point.x - fx, v = point.y - fy, w = point.z - fz;
    int32_t i = sat_i32(fx), j = sat_i32(fy), k = sat_i32(fz);
    Float c[8];
    for (int index = 0; index < 8; ++index) {
      int32_t di = index / 4, dj = (index / 2) % 2, dk = index % 2;
      uint32_t h = perm[((uint32_t)i + (uint32_t)di) & 255u] ^ perm[256u + (((uint32_t)j + (uint32_t)dj) & 255u)] ^
                   perm[512u + (((uint32_t)k + (uint32_t)dk) & 255u)];
      c[index] = data[h];
    }
    Float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
    Float value = 0.0f;
    for (int index = 0; index < 8; ++index) {
      int ii = index / 4, jj = (index / 2) % 2, kk = index % 2;
      Vec3 weight(u - (Float)ii, v - (Float)jj, w - (Float)kk);
      Vec3 cv = c[ii * 4 + jj * 2 + kk] * Vec3::one();
      value += ((Float)ii * uu + (1.0f - (Float)ii) * (1.0f - uu)) * ((Float)jj * vv + (1.0f - (Float)jj) * (1.0f - vv)) *
               ((Float)kk * ww + (1.0f - (Float)kk) * (1.0f - ww)) * cv.dot(weight);
    }
    return value;
  }

  Vec3 colour_value(const Vec3& direction, const Vec3& point) const {
    switch (kind) {
      case PTB_TEX_CHECKERED: {  // textures/mod.rs:61-73
        Float sign = std::sin(10.0f * point.x) * std::sin(10.0f * point.y) * std::sin(10.0f * point.z);
        return sign > 0.0f ? a : b;
      }
      case PTB_TEX_SOLID:  // textures/mod.rs:193-200
        return a;
      case PTB_TEX_LERP: {  // textures/mod.rs:283-291
        Float t = direction.z * 0.5f + 0.5f;
        return a * t + b * (1.0f - t);
      }
      case PTB_TEX_IMAGE: {  // textures/mod.rs:248-262 (lat-long lookup by DIRECTION, nearest texel)
        Float phi = std::atan2(direction.y, direction.x) + PI_F;
        Float theta = std::acos(direction.z);
        Float uvx = phi / (2.0f * PI_F), uvy = theta / PI_F;
        size_t dim0 = width - 1, dim1 = height - 1;
        size_t x_pixel = sat_usize((Float)dim0 * uvx), y_pixel = sat_usize((Float)dim1 * uvy);
        if (x_pixel > dim0) x_pixel = dim0;  // the reference would panic on an out-of-range index; unreachable for unit directions
        if (y_pixel > dim1) y_pixel = dim1;
        size_t index = y_pixel * (dim0 + 1) + x_pixel;
        return Vec3(data[3 * index], data[3 * index + 1], data[3 * index + 2]);
      }
      case PTB_TEX_PERLIN:  // textures/mod.rs:171-179
        return 0.5f * Vec3::one() * (1.0f + perlin_noise(point));
      default:  // trait default (textures/mod.rs:10-12)
        return Vec3(1.0f, 1.0f, 1.0f);
    }
  }
};

// textures/mod.rs:32-50
static inline std::vector<Float> generate_values(const Texture& tex, size_t res_x, size_t res_y) {
  std::vector<Float> values;
  Float step_x = 1.0f / (Float)res_x, step_y = 1.0f / (Float)res_y;
  for (size_t y = 0; y < res_y; ++y) {
    for (size_t x = 0; x < res_x; ++x) {
      Float u = ((Float)x + 0.5f) * step_x;
      Float v = ((Float)y + 0.5f) * step_y;
      Float phi = u * 2.0f * PI_F;
      Float theta = v * PI_F;
      Float sin_theta = std::sin(theta);
      Vec3 direction(std::cos(phi) * sin_theta, std::sin(phi) * sin_theta, std::cos(theta));
      Vec3 col = tex.colour_value(direction, Vec3::zero());
      values.push_back((0.2126f * col.x + 0.7152f * col.y + 0.0722f * col.z) * sin_theta);
    }
  }
  return values;
}

// ------------------------------------------------------------- distributions
// implementations/src/statistics/distributions.rs:11-72
struct Distribution1D {
  std::vector<Float> pdf, cdf;
  Distribution1D() {}
  explicit Distribution1D(const Float* values, size_t n) {
    std::vector<Float> intervals(1, 0.0f);
    for (size_t i = 1; i <= n; ++i) intervals.push_back(intervals[i - 1] + values[i - 1]);
    Float c = intervals[n];
    for (auto& v : intervals)
      if (c != 0.0f) v /= c;
    Float last = 0.0f;
    for (size_t i = 1; i <= n; ++i) {
      pdf.push_back(intervals[i] - last);
      last = intervals[i];
    }
    cdf = intervals;
  }
  // distributions.rs:51-72 (binary search for the first cdf entry > num, minus one, clamped)
  size_t sample_with(Float num) const {
    size_t first = 0, len = cdf.size();
    while (len > 0) {
      size_t half = len >> 1;
      size_t middle = first + half;
      if (cdf[middle] <= num) {
        first = middle + 1;
        len -= half + 1;
      } else {
        len = half;
      }
    }
    size_t r = first - 1;
    size_t hi = cdf.size() - 2;
    return r > hi ? hi : r;
  }
  size_t sample() const { return sample_with(g_rng.next_float01()); }
};

// distributions.rs:75-113
struct Distribution2D {
  std::vector<Distribution1D> x_distributions;
  Distribution1D y_distribution;
  size_t dim_x = 0, dim_y = 0;
  Distribution2D() {}
  Distribution2D(const std::vector<Float>& values, size_t width) {
    std::vector<Float> y_values;
    for (size_t off = 0; off + width <= values.size(); off += width) {
      x_distributions.emplace_back(&values[off], width);
      Float row_sum = 0.0f;
      for (size_t i = 0; i < width; ++i) row_sum += values[off + i];
      y_values.push_back(row_sum);
    }
    y_distribution = Distribution1D(y_values.data(), y_values.size());
    dim_x = width;
    dim_y = values.size() / width;
  }
  void sample(size_t& u, size_t& v) const {
    v = y_distribution.sample();
    u = x_distributions[v].sample();
  }
  Float pdf(Float u, Float v) const {
    size_t ui = sat_usize((Float)dim_x * u);
    if (ui > dim_x - 1) ui = dim_x - 1;
    size_t vi = sat_usize((Float)dim_y * v);
    if (vi > dim_y - 1) vi = dim_y - 1;
    return y_distribution.pdf[vi] * x_distributions[vi].pdf[ui];
  }
};

// ---------------------------------------------------------------- hit record
// rt_core/src/primitive.rs:3-10
struct Hit {
  Float t = 0.0f;
  Vec3 point, error, normal;
  bool out = false;
  Float b1 = 0.0f, b2 = 0.0f;  // uv of triangle.rs:179 (kept for the closest-hit comparison only)
};

// ----------------------------------------------------------------- materials
// implementations/src/statistics/bxdfs/lambertian.rs:5-22
namespace lambertian {
static inline Vec3 sample_local() {
  Float cos_theta = std::sqrt(1.0f - g_rng.next_float01());
  Float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
  Float phi = 2.0f * PI_F * g_rng.next_float01();
  return Vec3(std::cos(phi) * sin_theta, std::sin(phi) * sin_theta, cos_theta);
}
static inline Float pdf_local(const Vec3& outgoing) { return fmax_(outgoing.z, 0.0f) / PI_F; }
static inline Vec3 sample(const Vec3& normal) { return Coordinate::new_from_z(normal).to_coord(sample_local()); }
static inline Float pdf(const Vec3& outgoing, const Vec3& normal) { return fmax_(outgoing.dot(normal), 0.0f) / PI_F; }
}  // namespace lambertian

// materials/refract.rs:59-61
static inline Vec3 fresnel(Float cos, const Vec3& f0) { return f0 + (1.0f - f0) * std::pow(1.0f - cos, 5.0f); }

// statistics/bxdfs/trowbridge_reitz.rs:17-24,62-89 and trowbridge_reitz_vndf.rs (isotropic = anisotropic with a_x = a_y)
// Parity status: pinned by the reference's GGX tests (trowbridge_reitz.rs:128-230: projected area = 1, G1 integral =
// cos, weak white furnace = 1, G2 integral <= 1) and its chi-squared tests of the VNDF sampler against the pdf
// (trowbridge_reitz_vndf.rs:156-184), restated in tests/test_oracle_kats.py.
namespace tr {
static inline Float d(Float alpha, Float cos_theta) {  // trowbridge_reitz.rs:17-24
  if (cos_theta <= 0.0f) return 0.0f;
  Float a_sq = alpha * alpha;
  Float tmp = cos_theta * cos_theta * (a_sq - 1.0f) + 1.0f;
  return a_sq / (PI_F * tmp * tmp);
}
static inline Float g2(Float alpha, const Vec3& normal, const Vec3& h, const Vec3& incoming, const Vec3& outgoing) {  // :62-78
  if (incoming.dot(h) / incoming.dot(normal) <= 0.0f || outgoing.dot(h) / outgoing.dot(normal) <= 0.0f) return 0.0f;
  Float alpha_sq = alpha * alpha;
  Float one_minus_alpha_sq = 1.0f - alpha_sq;
  Float cos_i = normal.dot(incoming);
  Float cos_i_sq = cos_i * cos_i;
  Float tmp_a = alpha_sq + one_minus_alpha_sq * cos_i_sq;
  Float cos_o = normal.dot(outgoing);
  Float cos_o_sq = cos_o * cos_o;
  Float tmp_b = alpha_sq + one_minus_alpha_sq * cos_o_sq;
  return 2.0f * cos_i * cos_o / (cos_o * std::sqrt(tmp_a) + cos_i * std::sqrt(tmp_b));
}
static inline Float g1(Float alpha, const Vec3& normal, const Vec3& h, const Vec3& v) {  // :80-89
  if (v.dot(h) / v.dot(normal) <= 0.0f) return 0.0f;
  Float cos = normal.dot(v);
  Float cos_sq = cos * cos;
  Float alpha_sq = alpha * alpha;
  Float tmp = alpha_sq + (1.0f - alpha_sq) * cos_sq;
  return 2.0f * cos / (std::sqrt(tmp) + cos);
}
static inline Float vndf(Float a, const Vec3& h, const Vec3& incoming) {  // vndf.rs:9-15
  if (h.z < 0.0f) return 0.0f;
  return g1(a, Vec3(0.0f, 0.0f, 1.0f), h, incoming) * fmax_(incoming.dot(h), 0.0f) * d(a, h.z) / incoming.z;
}
static inline Vec3 sample_vndf(Float a_x, Float a_y, const Vec3& incoming) {  // vndf.rs:84-113
  Vec3 v_hemisphere = normalised(Vec3(a_x * incoming.x, a_y * incoming.y, incoming.z));
  Float len_sq = v_hemisphere.x * v_hemisphere.x + v_hemisphere.y * v_hemisphere.y;
  Vec3 basis_two = len_sq > 0.0f ? Vec3(-v_hemisphere.y, v_hemisphere.x, 0.0f) / std::sqrt(len_sq) : Vec3(1.0f, 0.0f, 0.0f);
  Vec3 basis_three = v_hemisphere.cross(basis_two);
  Float r = std::sqrt(g_rng.next_float01());
  Float phi = TAU_F * g_rng.next_float01();
  Float tx = r * std::cos(phi), ty = r * std::sin(phi);
  Float s = 0.5f * (1.0f + v_hemisphere.z);
  ty = (1.0f - s) * std::sqrt(1.0f - tx * tx) + s * ty;
  Vec3 h_hemisphere = tx * basis_two + ty * basis_three + std::sqrt(fmax_(1.0f - tx * tx - ty * ty, 0.0f)) * v_hemisphere;
  return normalised(Vec3(a_x * h_hemisphere.x, a_y * h_hemisphere.y, fmax_(h_hemisphere.z, 0.0f)));
}
static inline Vec3 sample_local(Float a, const Vec3& incoming) { return reflected(incoming, sample_vndf(a, a, incoming)); }
static inline Float pdf_local(Float alpha, const Vec3& incoming, const Vec3& outgoing) {  // vndf.rs:26-33
  Vec3 h = normalised(outgoing + incoming);
  if (h.z < 0.0f) h = -h;
  return vndf(alpha, h, incoming) / (4.0f * incoming.dot(h));
}
static inline Vec3 sample(Float a, const Vec3& incoming, const Vec3& normal) {  // vndf.rs:35-40
  Coordinate coord = Coordinate::new_from_z(normal);
  Coordinate inverse = coord.create_inverse();
  Vec3 h = coord.to_coord(sample_vndf(a, a, inverse.to_coord(incoming)));
  return reflected(incoming, h);
}
static inline Float pdf(Float alpha, const Vec3& incoming, const Vec3& outgoing, const Vec3& normal) {  // vndf.rs:42-52
  Coordinate inverse = Coordinate::new_from_z(normal).create_inverse();
  return pdf_local(alpha, inverse.to_coord(incoming), inverse.to_coord(outgoing));
}
}  // namespace tr

struct Material {
  uint32_t kind;
  const Texture* texture;
  Float param;
  Vec3 ior = Vec3(1.0f, 1.0f, 1.0f);  // TrowbridgeReitz only (trowbridge_reitz.rs:6-11); param = alpha
  Float metallic = 0.0f;

  // trowbridge_reitz.rs:26-31
  Vec3 tr_fresnel(const Hit& hit, const Vec3& wo, const Vec3& wi, const Vec3& h) const {
    Vec3 f0 = ((1.0f - ior) / (ior + 1.0f)).abs();
    f0 = f0 * f0;
    f0 = (1.0f - metallic) * f0 + metallic * texture->colour_value(wi, hit.point);  // lerp, trowbridge_reitz.rs:86-88
    return fresnel(wo.dot(h), f0);
  }

  // rt_core/src/material.rs:4-30 defaults + the overrides of materials/{emissive,lambertian,reflect,refract}.rs
  bool is_light() const { return kind == PTB_MAT_EMIT; }
  bool is_delta() const { return kind == PTB_MAT_REFLECT || kind == PTB_MAT_REFRACT; }

  static bool reflect_scatter(Float fuzz, Ray& ray, const Hit& hit) {  // reflect.rs:26-36
    Vec3 direction = reflected(-ray.direction, hit.normal);
    Vec3 point = offset_ray(hit.point, hit.normal, hit.error, true);
    ray = Ray(point, direction + fuzz * random_unit_vector(), ray.time);
    return false;
  }

  bool scatter_ray(Ray& ray, const Hit& hit) const {
    switch (kind) {
      case PTB_MAT_EMIT:  // emissive.rs:36-38
        return true;
      case PTB_MAT_LAMBERTIAN: {  // lambertian.rs:30-41
        Vec3 direction = lambertian::sample(hit.normal);
        Vec3 point = offset_ray(hit.point, hit.normal, hit.error, true);
        ray = Ray(point, direction, ray.time);
        return false;
      }
      case PTB_MAT_TROWBRIDGE_REITZ: {  // trowbridge_reitz.rs:38-50
        Vec3 direction = tr::sample(param, -ray.direction, hit.normal);
        Vec3 point = offset_ray(hit.point, hit.normal, hit.error, true);
        ray = Ray(point, direction, ray.time);
        return false;
      }
      case PTB_MAT_REFLECT:
        return reflect_scatter(param, ray, hit);
      case PTB_MAT_REFRACT: {  // refract.rs:27-50
        Float eta_fraction = 1.0f / param;
        if (!hit.out) eta_fraction = param;
        Float cos_theta = fmin_((-ray.direction).dot(hit.normal), 1.0f);
        Float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        bool cannot_refract = eta_fraction * sin_theta > 1.0f;
        Float f0s = (1.0f - eta_fraction) / (1.0f + eta_fraction);
        Vec3 f0 = f0s * f0s * Vec3::one();
        if (cannot_refract || fresnel(cos_theta, f0).x > random_float()) return reflect_scatter(0.0f, ray, hit);
        Vec3 perp = eta_fraction * (ray.direction + cos_theta * hit.normal);
        Vec3 para = -1.0f * std::sqrt(std::fabs(1.0f - perp.mag_sq())) * hit.normal;
        Vec3 direction = perp + para;
        Vec3 point = offset_ray(hit.point, hit.normal, hit.error, false);
        ray = Ray(point, direction, ray.time);
        return false;
      }
      default:
        return true;
    }
  }
  Float scattering_pdf(const Hit& hit, const Vec3& wo, const Vec3& wi) const {
    if (kind == PTB_MAT_LAMBERTIAN) return lambertian::pdf(wi, hit.normal);  // lambertian.rs:42-44
    if (kind == PTB_MAT_TROWBRIDGE_REITZ) {  // trowbridge_reitz.rs:51-59
      Float a = tr::pdf(param, -wo, wi, hit.normal);
      return a == 0.0f ? INF_F : a;
    }
    return 0.0f;  // trait default (material.rs:20-22): Reflect/Refract do not override (quirk Q4)
  }
  Vec3 eval(const Hit& hit, const Vec3& wo, const Vec3& wi) const {
    if (kind == PTB_MAT_LAMBERTIAN)  // lambertian.rs:45-47
      return texture->colour_value(wo, hit.point) * param * fmax_(hit.normal.dot(wi), 0.0f) / PI_F;
    if (kind == PTB_MAT_TROWBRIDGE_REITZ) {  // trowbridge_reitz.rs:60-73
      Vec3 wo_ = -wo;
      Vec3 h = normalised(wi + wo_);
      if (wi.dot(hit.normal) < 0.0f || h.dot(wo_) < 0.0f) return Vec3::zero();
      Vec3 f = tr_fresnel(hit, wo_, wi, h);
      Float g = tr::g2(param, hit.normal, h, wo_, wi);
      Float d = tr::d(param, hit.normal.dot(h));
      return f * g * d / (4.0f * std::fabs(wo_.dot(hit.normal)) * wi.dot(hit.normal));
    }
    return texture->colour_value(wo, hit.point);  // reflect.rs:37-39, refract.rs:51-53
  }
  Vec3 eval_over_scattering_pdf(const Hit& hit, const Vec3& wo, const Vec3& wi) const {
    if (kind == PTB_MAT_LAMBERTIAN) return texture->colour_value(wo, hit.point) * param;  // lambertian.rs:48-50
    if (kind == PTB_MAT_TROWBRIDGE_REITZ) {  // trowbridge_reitz.rs:74-87
      Vec3 wo_ = -wo;
      Vec3 h = normalised(wi + wo_);
      if (wo_.dot(h) < 0.0f || wi.dot(hit.normal) < 0.0f) return Vec3::zero();
      Vec3 f = tr_fresnel(hit, wo_, wi, h);
      Float g = tr::g2(param, hit.normal, h, wo_, wi);
      return f * g / tr::g1(param, hit.normal, h, wo_);
    }
    return eval(hit, wo, wi) / scattering_pdf(hit, wo, wi);  // material.rs:24-26  (colour / 0.0)
  }
  Vec3 get_emission(const Hit& hit, const Vec3& wo) const {
    if (kind == PTB_MAT_EMIT) {  // emissive.rs:23-26
      Vec3 point = offset_ray(hit.point, hit.normal, hit.error, true);
      return param * texture->colour_value(wo, point);
    }
    return Vec3::zero();  // material.rs:27-29
  }
};

// ---------------------------------------------------------------- primitives
struct Prim {
  uint32_t is_sphere;
  uint32_t orig_id;  // index in the loader's order (spheres first, then triangles)
  const Material* material;
  Vec3 center;
  Float radius;
  Vec3 p[3], n[3];

  // primitives/sphere.rs:34-105
  bool sphere_int(const Ray& ray, Hit& h) const {
    Vec3 dir = ray.direction;
    Vec3 orig = ray.origin;
    Vec3 deltap = center - orig;
    Float ddp = dir.dot(deltap);
    Float deltapdot = deltap.dot(deltap);
    Vec3 remedy_term = deltap - ddp * dir;
    Float discriminant = radius * radius - remedy_term.dot(remedy_term);
    if (!(discriminant > 0.0f)) return false;
    Float sqrt_val = std::sqrt(discriminant);
    Float q = ddp > 0.0f ? ddp + sqrt_val : ddp - sqrt_val;
    Float t0 = q;
    Float t1 = (deltapdot - radius * radius) / q;
    if (t1 < t0) { Float tmp = t0; t0 = t1; t1 = tmp; }
    Float t;
    if (t0 > 0.0f) t = t0;
    else {
      if (t1 <= 0.0f) return false;
      t = t1;
    }
    Vec3 point = ray.at(t);
    Vec3 normal = (point - center) / radius;
    bool out = true;
    if (normal.dot(dir) > 0.0f) { out = false; normal = -normal; }
    h.t = t;
    h.point = point;
    h.error = EPSILON_RT * Vec3::one();
    h.normal = normal;
    h.out = out;
    h.b1 = h.b2 = 0.0f;
    return true;
  }

  // primitives/triangle.rs:105-216
  bool triangle_int(const Ray& ray, Hit& h) const {
    Vec3 p0t = p[0] - ray.origin, p1t = p[1] - ray.origin, p2t = p[2] - ray.origin;
    // Axis::get_max_abs_axis + Axis::swap_z (primitives/mod.rs:62-82): X and Y both swap x<->z
    const Vec3& d = ray.direction;
    bool swap = (std::fabs(d.x) > std::fabs(d.y) && std::fabs(d.x) > std::fabs(d.z)) || (std::fabs(d.y) > std::fabs(d.z));
    if (swap) {
      Float tmp;
      tmp = p0t.x; p0t.x = p0t.z; p0t.z = tmp;
      tmp = p1t.x; p1t.x = p1t.z; p1t.z = tmp;
      tmp = p2t.x; p2t.x = p2t.z; p2t.z = tmp;
    }
    p0t.x += ray.shear.x * p0t.z; p0t.y += ray.shear.y * p0t.z;
    p1t.x += ray.shear.x * p1t.z; p1t.y += ray.shear.y * p1t.z;
    p2t.x += ray.shear.x * p2t.z; p2t.y += ray.shear.y * p2t.z;

    Float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    Float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    Float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
      e0 = (Float)((double)p1t.x * (double)p2t.y - (double)p1t.y * (double)p2t.x);
      e1 = (Float)((double)p2t.x * (double)p0t.y - (double)p2t.y * (double)p0t.x);
      e2 = (Float)((double)p0t.x * (double)p1t.y - (double)p0t.y * (double)p1t.x);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    Float det = e0 + e1 + e2;
    if (det == 0.0f) return false;

    p0t = p0t * ray.shear.z;
    p1t = p1t * ray.shear.z;
    p2t = p2t * ray.shear.z;

    Float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if ((det < 0.0f && t_scaled >= 0.0f) || (det > 0.0f && t_scaled <= 0.0f)) return false;

    Float inv_det = 1.0f / det;
    Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    Float t = inv_det * t_scaled;

    Float max_z_t = Vec3(std::fabs(p0t.z), std::fabs(p1t.z), std::fabs(p2t.z)).component_max();
    Float delta_z = gamma(3) * max_z_t;
    Float max_x_t = Vec3(std::fabs(p0t.x), std::fabs(p1t.x), std::fabs(p2t.x)).component_max();
    Float max_y_t = Vec3(std::fabs(p0t.y), std::fabs(p1t.y), std::fabs(p2t.y)).component_max();
    Float delta_x = gamma(5) * (max_x_t + max_z_t);
    Float delta_y = gamma(5) * (max_y_t + max_z_t);
    Float delta_e = 2.0f * (gamma(2) * max_x_t * max_y_t + delta_y * max_x_t + delta_x * max_y_t);
    Float max_e = Vec3(std::fabs(e0), std::fabs(e1), std::fabs(e2)).component_max();
    Float delta_t = 3.0f * (gamma(3) * max_e * max_z_t + delta_e * max_z_t + delta_z * max_e) * std::fabs(inv_det);
    if (t < delta_t) return false;

    Vec3 normal = b0 * n[0] + b1 * n[1] + b2 * n[2];
    bool out = check_side(normal, ray.direction);

    Float x_abs_sum = std::fabs(b0 * p[0].x) + std::fabs(b1 * p[1].x) + std::fabs(b2 * p[2].x);
    Float y_abs_sum = std::fabs(b0 * p[0].y) + std::fabs(b1 * p[1].y) + std::fabs(b2 * p[2].y);
    Float z_abs_sum = std::fabs(b0 * p[0].z) + std::fabs(b1 * p[1].z) + std::fabs(b2 * p[2].z);
    Vec3 point_error =
        gamma(7) * Vec3(x_abs_sum, y_abs_sum, z_abs_sum) + gamma(6) * Vec3(b2 * p[2].x, b2 * p[2].y, b2 * p[2].z);
    Vec3 point = b0 * p[0] + b1 * p[1] + b2 * p[2];

    h.t = t;
    h.point = point;
    h.error = point_error;
    h.normal = normal;
    h.out = out;
    // uv = b0*(0,0) + b1*(1,0) + b2*(1,1)  (triangle.rs:179); we keep b1, b2 themselves
    h.b1 = b1;
    h.b2 = b2;
    return true;
  }

  bool get_int(const Ray& ray, Hit& h) const { return is_sphere ? sphere_int(ray, h) : triangle_int(ray, h); }

  void aabb(Vec3& mn, Vec3& mx) const {
    if (is_sphere) {  // sphere.rs:175-181
      mn = center - radius * Vec3::one();
      mx = center + radius * Vec3::one();
    } else {  // triangle.rs:285-307
      mn = p[0].min_by_component(p[1].min_by_component(p[2]));
      mx = p[0].max_by_component(p[1].max_by_component(p[2]));
    }
  }

  Float area() const {
    if (is_sphere) return 4.0f * PI_F * radius * radius;                  // sphere.rs:168-170
    return 0.5f * (p[1] - p[0]).cross(p[2] - p[0]).mag();                // triangle.rs:249-257
  }

  // sphere.rs:112-117
  Vec3 sphere_get_sample() const {
    Float z = 1.0f - 2.0f * random_float();
    Float a = std::sqrt(fmax_(1.0f - z * z, 0.0f));
    Float b = 2.0f * PI_F * random_float();
    return center + radius * Vec3(a * std::cos(b), a * std::sin(b), z);
  }

  Vec3 sample_visible_from_point(const Vec3& in_point) const {
    if (is_sphere) {  // sphere.rs:118-154
      Float distance_sq = (in_point - center).mag_sq();
      Vec3 point;
      if (distance_sq <= radius * radius) {
        point = sphere_get_sample();
      } else {
        Float distance = std::sqrt(distance_sq);
        Float sin_theta_max_sq = radius * radius / distance_sq;
        Float cos_theta_max = std::sqrt(fmax_(1.0f - sin_theta_max_sq, 0.0f));
        Float r1 = random_float();
        Float cos_theta = (1.0f - r1) + r1 * cos_theta_max;
        Float sin_theta = std::sqrt(fmax_(1.0f - cos_theta * cos_theta, 0.0f));
        Float phi = 2.0f * random_float() * PI_F;
        Float ds = distance * cos_theta - std::sqrt(fmax_(radius * radius - distance_sq * sin_theta * sin_theta, 0.0f));
        Float cos_alpha = (distance_sq + radius * radius - ds * ds) / (2.0f * distance * radius);
        Float sin_alpha = std::sqrt(fmax_(1.0f - cos_alpha * cos_alpha, 0.0f));
        Coordinate cs = Coordinate::new_from_z(normalised(in_point - center));
        Vec3 vec(sin_alpha * std::cos(phi), sin_alpha * std::sin(phi), cos_alpha);
        vec = cs.to_coord(vec);
        point = center + radius * vec;
      }
      return normalised(point - in_point);
    }
    // MeshTriangle (triangle.rs:258-277): uv = (1 - sqrt(r1), sqrt(r1) * sqrt(r2))   (quirk Q8)
    Float s = std::sqrt(g_rng.next_float01());
    Float u0 = 1.0f - s;
    Float u1 = s * std::sqrt(g_rng.next_float01());
    Vec3 point = u0 * p[0] + u1 * p[1] + (1.0f - u0 - u1) * p[2];
    return normalised(point - in_point);
  }

  Float scattering_pdf(const Vec3& hit_point, const Vec3& wi, const Hit& sampled_hit) const {
    if (is_sphere) {  // sphere.rs:155-167
      Float rsq = radius * radius;
      Float dsq = (hit_point - center).mag_sq();
      if (dsq <= rsq) return (sampled_hit.point - hit_point).mag_sq() / (std::fabs(wi.dot(sampled_hit.normal)) * area());
      Float sin_theta_max_sq = rsq / dsq;
      Float cos_theta_max = std::sqrt(fmax_(1.0f - sin_theta_max_sq, 0.0f));
      return 1.0f / (2.0f * PI_F * (1.0f - cos_theta_max));
    }
    // triangle.rs:278-280
    return (sampled_hit.point - hit_point).mag_sq() / (std::fabs(wi.dot(sampled_hit.normal)) * area());
  }
};

// ----------------------------------------------------------------------- sky
// implementations/src/sky.rs:12-92
struct Sky {
  const Texture* texture = nullptr;
  Material mat;  // Emit(texture, 1.0)  (loader/src/misc.rs:26)
  Distribution2D distribution;
  bool has_distribution = false;
  size_t res_x = 0, res_y = 0;

  void init(const Texture* tex, size_t rx, size_t ry) {
    texture = tex;
    mat.kind = PTB_MAT_EMIT;
    mat.texture = tex;
    mat.param = 1.0f;
    res_x = rx;
    res_y = ry;
    std::vector<Float> values = generate_values(*tex, rx, ry);
    has_distribution = (rx | ry) != 0;
    if (has_distribution) distribution = Distribution2D(values, rx);
  }
  bool can_sample() const { return (res_x | res_y) != 0; }
  Float pdf(const Vec3& wi) const {  // sky.rs:43-60
    Float sin_theta = std::sqrt(1.0f - wi.z * wi.z);
    if (sin_theta <= 0.0f) return 0.0f;
    Float theta = std::acos(wi.z);
    Float phi = std::atan2(wi.y, wi.x);
    if (phi < 0.0f) phi += 2.0f * PI_F;
    Float u = phi / (2.0f * PI_F);
    Float v = theta / PI_F;
    return (Float)res_x * (Float)res_y * distribution.pdf(u, v) / (sin_theta * TAU_F * PI_F);
  }
  Vec3 sample() const {  // sky.rs:64-78
    size_t ui, vi;
    distribution.sample(ui, vi);
    Float u = next_float((Float)ui + random_float()) / (Float)res_x;
    Float v = next_float((Float)vi + random_float()) / (Float)res_y;
    Float phi = u * 2.0f * PI_F;
    Float theta = v * PI_F;
    Float st = std::sin(theta), ct = std::cos(theta), sp = std::sin(phi), cp = std::cos(phi);
    return Vec3(st * cp, st * sp, ct);  // Vec3::from_spherical (vec.rs:153-161)
  }
};

}  // namespace ref
