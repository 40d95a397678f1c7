// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of nonl4331/raytracing-rust's hot-path arithmetic (f32, strict IEEE: build with
// -O2 -fno-fast-math -ffp-contract=off, because rustc never contracts a*b+c into an FMA).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
// anything under oracle/. The product (libptb200.so) never links, loads or calls it.
//
// Parity status of this file: Vec3 / Ray::new / gamma / next_float / offset_ray / Coordinate are pinned
// by the reference's own property test (utility/coord.rs:39-49) and by hand-derived answers
// (SURVEY.md appendix B); the reference holds no golden vectors for them beyond that.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace ref {

typedef float Float;
// rt_core/src/lib.rs:23-29
static const Float EPSILON_RT = 3.0e-4f;
static const Float PI_F = 3.14159265358979323846f;
static const Float TAU_F = 6.28318530717958647692f;
static const Float F32_EPS = 1.1920929e-7f;  // f32::EPSILON
static const Float INF_F = std::numeric_limits<float>::infinity();

// Rust's f32::min / f32::max return the non-NaN operand; fminf/fmaxf have the same contract.
static inline Float fmin_(Float a, Float b) { return std::fmin(a, b); }
static inline Float fmax_(Float a, Float b) { return std::fmax(a, b); }

// rt_core/src/vec.rs:108-248
struct Vec3 {
  Float x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
  static Vec3 one() { return Vec3(1, 1, 1); }
  static Vec3 zero() { return Vec3(0, 0, 0); }
  Float dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }            // vec.rs:164-166
  Vec3 cross(const Vec3& o) const {                                                  // vec.rs:169-175
    return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
  }
  Float mag_sq() const { return dot(*this); }
  Float mag() const { return std::sqrt(dot(*this)); }
  Vec3 abs() const { return Vec3(std::fabs(x), std::fabs(y), std::fabs(z)); }
  Float component_min() const { return fmin_(x, fmin_(y, z)); }                      // vec.rs:211-213
  Float component_max() const { return fmax_(x, fmax_(y, z)); }                      // vec.rs:216-218
  Vec3 min_by_component(const Vec3& o) const { return Vec3(fmin_(x, o.x), fmin_(y, o.y), fmin_(z, o.z)); }
  Vec3 max_by_component(const Vec3& o) const { return Vec3(fmax_(x, o.x), fmax_(y, o.y), fmax_(z, o.z)); }
  bool contains_nan() const { return std::isnan(x) || std::isnan(y) || std::isnan(z); }
  // vec.rs:245-247 — true if ANY component is finite (quirk Q5)
  bool is_finite() const { return std::isfinite(x) || std::isfinite(y) || std::isfinite(z); }
  bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const Vec3& o) const { return !(*this == o); }
};
static inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline Vec3 operator*(const Vec3& a, const Vec3& b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline Vec3 operator/(const Vec3& a, const Vec3& b) { return Vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline Vec3 operator*(const Vec3& a, Float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
static inline Vec3 operator*(Float s, const Vec3& a) { return Vec3(s * a.x, s * a.y, s * a.z); }
static inline Vec3 operator/(const Vec3& a, Float s) { return Vec3(a.x / s, a.y / s, a.z / s); }
static inline Vec3 operator+(const Vec3& a, Float s) { return Vec3(a.x + s, a.y + s, a.z + s); }
static inline Vec3 operator-(Float s, const Vec3& a) { return Vec3(s - a.x, s - a.y, s - a.z); }
static inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
static inline Vec3 normalised(const Vec3& a) { return a / a.mag(); }                 // vec.rs:189-191
// vec.rs:203-205 — self points away from the surface
static inline Vec3 reflected(const Vec3& v, const Vec3& n) { return 2.0f * v.dot(n) * n - v; }

struct Vec2 {
  Float x, y;
  Vec2() : x(0), y(0) {}
  Vec2(Float x_, Float y_) : x(x_), y(y_) {}
};

// rt_core/src/ray.rs:4-46 (incl. quirk Q1: x<->z swap for BOTH x- and y-dominant directions)
struct Ray {
  Vec3 origin, direction, d_inverse, shear;
  Float time;
  Ray() : time(0) {}
  Ray(const Vec3& o, Vec3 d, Float t) {
    d = d / d.mag();  // direction.normalise()
    int max_axis;
    if (std::fabs(d.x) > std::fabs(d.y) && std::fabs(d.x) > std::fabs(d.z)) max_axis = 0;
    else if (std::fabs(d.y) > std::fabs(d.z)) max_axis = 1;
    else max_axis = 2;
    Vec3 s = d;
    if (max_axis == 0 || max_axis == 1) { Float tmp = s.x; s.x = s.z; s.z = tmp; }
    Float shear_x = -s.x / s.z;
    Float shear_y = -s.y / s.z;
    Float shear_z = 1.0f / s.z;
    origin = o;
    direction = d;
    d_inverse = Vec3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    shear = Vec3(shear_x, shear_y, shear_z);
    time = t;
  }
  Vec3 at(Float t) const { return origin + direction * t; }
};

static inline uint32_t f2u(Float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline Float u2f(uint32_t u) { Float f; std::memcpy(&f, &u, 4); return f; }

// implementations/src/utility/mod.rs:51-65
static inline Float next_float(Float f) {
  if (std::isinf(f) && f > 0.0f) return f;
  if (f == -0.0f) f = 0.0f;
  return u2f(f >= 0.0f ? f2u(f) + 1u : f2u(f) - 1u);
}
// utility/mod.rs:67-81
static inline Float previous_float(Float f) {
  if (std::isinf(f) && f < 0.0f) return f;
  if (f == 0.0f) f = -0.0f;
  return u2f(f <= 0.0f ? f2u(f) + 1u : f2u(f) - 1u);
}
// utility/mod.rs:83-86
static inline Float gamma(uint32_t n) {
  Float nm = (Float)n * 0.5f * F32_EPS;
  return nm / (1.0f - nm);
}
// utility/mod.rs:6-13
static inline bool check_side(Vec3& normal, const Vec3& ray_direction) {
  if (normal.dot(ray_direction) > 0.0f) { normal = -normal; return false; }
  return true;
}
// utility/mod.rs:88-117
static inline Vec3 offset_ray(const Vec3& origin, const Vec3& normal, const Vec3& error, bool is_brdf) {
  Float offset_val = normal.abs().dot(error);
  Vec3 offset = offset_val * normal;
  if (!is_brdf) offset = -offset;
  Vec3 p = origin + offset;
  p.x = offset.x > 0.0f ? next_float(p.x) : previous_float(p.x);
  p.y = offset.y > 0.0f ? next_float(p.y) : previous_float(p.y);
  p.z = offset.z > 0.0f ? next_float(p.z) : previous_float(p.z);
  return p;
}

// implementations/src/utility/coord.rs:10-30
struct Coordinate {
  Vec3 x, y, z;
  static Coordinate new_from_z(const Vec3& z) {
    Coordinate c;
    if (std::fabs(z.x) > std::fabs(z.y)) c.x = Vec3(-z.z, 0.0f, z.x) / std::sqrt(z.x * z.x + z.z * z.z);
    else c.x = Vec3(0.0f, z.z, -z.y) / std::sqrt(z.y * z.y + z.z * z.z);
    c.y = c.x.cross(z);
    c.z = z;
    return c;
  }
  Coordinate create_inverse() const {
    Coordinate c;
    c.x = Vec3(x.x, y.x, z.x);
    c.y = Vec3(x.y, y.y, z.y);
    c.z = Vec3(x.z, y.z, z.z);
    return c;
  }
  Vec3 to_coord(const Vec3& v) const { return v.x * x + v.y * y + v.z * z; }
};

// rt_core/src/lib.rs:36-40
static inline Float power_heuristic(Float pdf_a, Float pdf_b) {
  Float a_sq = pdf_a * pdf_a;
  return a_sq / (a_sq + pdf_b * pdf_b);
}

// Rust `as usize` on a float saturates (NaN -> 0, negative -> 0).
static inline size_t sat_usize(Float f) {
  if (!(f > 0.0f)) return 0;
  if (f >= 1.8446744e19f) return (size_t)-1;
  return (size_t)f;
}

// Rust `as i32` on a float saturates (NaN -> 0).
static inline int32_t sat_i32(Float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT32_MAX;
  if (f <= -2147483648.0f) return INT32_MIN;
  return (int32_t)f;
}

}  // namespace ref
