// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_math.hpp header).
// NaiveIntegrator, MisIntegrator, sample_lights, camera and the pass-per-sample driver of the reference.
// Follows implementations/src/integrators/{mod.rs,mis.rs}, camera.rs, samplers/random_sampler.rs.
//
// Parity status: pinned by the reference's documented (disabled) integration answers — furnace radiance
// (0.25,0.25,0.25) +- 0.001 and MIS == naive +- 0.001 (implementations/tests/sampling.rs:181-297) — which
// tests/test_oracle_kats.py runs against this restatement. No rendered-image golden exists in the reference.
#pragma once
#include <atomic>
#include <thread>
#include <vector>

#include "ref_bvh.hpp"

namespace ref {

static const uint32_t MAX_DEPTH = 50;                  // integrators/mod.rs:7
static const uint32_t RUSSIAN_ROULETTE_THRESHOLD = 3;  // integrators/mod.rs:8

struct IntegratorOpts {
  uint32_t max_depth = MAX_DEPTH;
  uint32_t rr_threshold = RUSSIAN_ROULETTE_THRESHOLD;
};

struct RayCounts {
  uint64_t reference = 0;  // the reference's own ray_count (integrators/mod.rs:34, mis.rs:38)
  uint64_t camera = 0, bounce = 0, shadow_light = 0, shadow_sky = 0;  // every traversal launched, by class
  // traversal statistics are OFF unless asked for (RenderOpts::count_traversal): the timed CPU baseline of bench.py
  // must not pay for counters the reference does not keep
  uint64_t nodes_visited = 0, prims_tested = 0;
  bool count_traversal = false;
  uint64_t* nv() { return count_traversal ? &nodes_visited : nullptr; }
  uint64_t* pt() { return count_traversal ? &prims_tested : nullptr; }
};

// integrators/mod.rs:22-78
static inline Vec3 naive_get_colour(Ray& ray, const Bvh& bvh, const IntegratorOpts& o, RayCounts& rc) {
  Vec3 throughput = Vec3::one(), output = Vec3::zero();
  uint32_t depth = 0;
  while (depth < o.max_depth) {
    Hit hit;
    const Material* mat;
    bvh.check_hit(ray, hit, mat, rc.nv(), rc.pt());
    rc.reference += 1;
    if (depth == 0) rc.camera += 1; else rc.bounce += 1;

    Vec3 wo = ray.direction;
    Vec3 emission = mat->get_emission(hit, wo);
    g_rng.select(depth, RNG_SCATTER);
    bool exit = mat->scatter_ray(ray, hit);

    if (depth == 0) {
      output = output + emission;
      if (exit) break;
    }
    if (exit) {
      output = output + throughput * emission;
      break;
    }
    if (!mat->is_delta()) throughput = throughput * mat->eval_over_scattering_pdf(hit, wo, ray.direction);
    else throughput = throughput * mat->eval(hit, wo, ray.direction);

    if (depth > o.rr_threshold) {
      Float p = throughput.component_max();
      g_rng.select(depth, RNG_RR);
      if (g_rng.next_float01() > p) break;
      throughput = throughput / p;
    }
    depth += 1;
  }
  if (output.contains_nan() || !output.is_finite()) return Vec3::zero();
  return output;
}

// integrators/mis.rs:95-157 — returns true with (l_wi, le, l_pdf) when a light sample is usable
static inline bool sample_lights(const Bvh& bvh, const Hit& hit, uint32_t depth, Vec3& l_wi, Vec3& le, Float& l_pdf,
                                 RayCounts& rc) {
  const Sky& sky = bvh.sky;
  size_t samplable_len = bvh.lights.size();
  bool sky_can_sample = sky.can_sample();
  g_rng.select(depth, RNG_NEE);

  auto sample_sky = [&](Float pdf_multiplier) -> bool {
    l_wi = sky.sample();
    Ray ray(hit.point + 0.0001f * hit.normal, l_wi, 0.0f);
    Hit sa;
    const Material* m;
    rc.shadow_sky += 1;
    size_t index = bvh.check_hit(ray, sa, m, rc.nv(), rc.pt());
    if (index == Bvh::MISS) {
      le = m->get_emission(hit, l_wi);
      l_pdf = sky.pdf(l_wi) * pdf_multiplier;
      return true;
    }
    return false;
  };
  auto sample_light = [&](Float pdf_multiplier, size_t li) -> bool {
    size_t index = bvh.lights[li];
    const Prim& light = bvh.primitives[index];
    l_wi = light.sample_visible_from_point(hit.point);
    Hit si;
    rc.shadow_light += 1;
    if (bvh.check_hit_index(Ray(hit.point + 0.0001f * hit.normal, l_wi, 0.0f), index, si)) {
      Float pdf = light.scattering_pdf(hit.point, l_wi, si);
      if (pdf > 0.0f) {
        le = light.material->get_emission(si, l_wi);
        l_pdf = pdf * pdf_multiplier;
        return true;
      }
    }
    return false;
  };

  if (samplable_len == 0 && !sky_can_sample) return false;
  if (samplable_len == 0) {
    // the reference draws no light index here (mis.rs:135-137); the device draws and discards one so that both
    // consume the NEE stream identically: draw 0 is always the light choice.
    (void)g_rng.next_u32();
    return sample_sky(1.0f);
  }
  if (!sky_can_sample) {
    Float multiplier = 1.0f / (Float)samplable_len;
    size_t li = g_rng.next_below((uint32_t)samplable_len);
    return sample_light(multiplier, li);
  }
  Float multiplier = 1.0f / (Float)(samplable_len + 1);
  size_t li = g_rng.next_below((uint32_t)(samplable_len + 1));
  if (li == samplable_len) return sample_sky(multiplier);
  return sample_light(multiplier, li);
}

// integrators/mis.rs:6-93
static inline Vec3 mis_get_colour(Ray& ray, const Bvh& bvh, const IntegratorOpts& o, RayCounts& rc) {
  Vec3 throughput = Vec3::one(), output = Vec3::zero();
  Hit hit;
  const Material* mat;
  rc.camera += 1;
  bvh.check_hit(ray, hit, mat, rc.nv(), rc.pt());
  Vec3 wo = ray.direction;
  Vec3 emission = mat->get_emission(hit, wo);
  {
    Ray scratch = ray;  // mis.rs:25 scatters a clone only to learn `exit`
    g_rng.select(0, RNG_SCATTER);
    bool exit = mat->scatter_ray(scratch, hit);
    output = output + emission;
    if (exit) return output;  // note: returned before the NaN check (mis.rs:29-31)
  }
  uint32_t depth = 1;
  while (depth < o.max_depth) {
    Vec3 l_wi, le;
    Float l_pdf;
    bool got = sample_lights(bvh, hit, depth, l_wi, le, l_pdf, rc);
    rc.reference += 1;
    if (got) {
      Float m_pdf = mat->scattering_pdf(hit, wo, l_wi);
      Float mis_weight = power_heuristic(l_pdf, m_pdf);
      output = output + throughput * mat->eval(hit, wo, l_wi) * mis_weight * le / l_pdf;
    }
    g_rng.select(depth, RNG_SCATTER);
    bool exit = mat->scatter_ray(ray, hit);
    if (exit) break;
    Vec3 m_wi = ray.direction;

    Hit next_hit;
    const Material* next_mat;
    rc.bounce += 1;
    size_t index = bvh.check_hit(ray, next_hit, next_mat, rc.nv(), rc.pt());

    Float m_pdf = mat->scattering_pdf(hit, wo, m_wi);
    Vec3 le2 = next_mat->get_emission(hit /* previous hit: quirk Q6 */, m_wi);
    throughput = throughput * mat->eval_over_scattering_pdf(hit, wo, m_wi);
    if (le2 != Vec3::zero()) {
      if ((bvh.is_samplable(index) && !mat->is_delta()) || (index == Bvh::MISS && bvh.sky.can_sample())) {
        Float lp = bvh.get_pdf_from_index(hit, next_hit, m_wi, index);
        Float mis_weight = power_heuristic(m_pdf, lp);
        output = output + throughput * le2 * mis_weight;
      } else {
        output = output + throughput * le2;
      }
    }
    if (next_mat->is_light()) break;

    if (depth > o.rr_threshold) {
      Float p = throughput.component_max();
      g_rng.select(depth, RNG_RR);
      if (g_rng.next_float01() > p) break;
      throughput = throughput / p;
    }
    wo = m_wi;
    hit = next_hit;
    mat = next_mat;
    depth += 1;
  }
  if (output.contains_nan() || !output.is_finite()) return Vec3::zero();
  return output;
}

// implementations/src/camera.rs:57-63 (the random `time` is never read afterwards and is not drawn)
struct Camera {
  Vec3 origin, lower_left, horizontal, vertical;
  Ray get_ray(Float u, Float v) const {
    return Ray(origin, lower_left + horizontal * u + vertical * v - origin, 0.0f);
  }
};

// camera.rs:20-53
static inline Camera make_camera(const Vec3& origin, const Vec3& lookat, const Vec3& vup, Float fov, Float aspect_ratio,
                                 Float /*aperture*/, Float focus_dist) {
  Float fov_rad = fov * (PI_F / 180.0f);  // f32::to_radians
  Float viewport_width = 2.0f * std::tan(fov_rad / 2.0f);
  Float viewport_height = viewport_width / aspect_ratio;
  Vec3 w = normalised(origin - lookat);
  Vec3 u = normalised(w.cross(vup));
  Vec3 v = u.cross(w);
  Vec3 horizontal = focus_dist * u * viewport_width;
  Vec3 vertical = focus_dist * v * viewport_height;
  Camera c;
  c.origin = origin;
  c.horizontal = horizontal;
  c.vertical = vertical;
  c.lower_left = origin - horizontal / 2.0f - vertical / 2.0f - focus_dist * w;
  return c;
}

struct RenderOpts {
  uint32_t width = 0, height = 0, spp = 0, sample_offset = 0, method = PTB_METHOD_MIS;
  uint64_t seed = 0;
  IntegratorOpts integ;
  unsigned threads = 0;  // 0 -> hardware_concurrency
  bool count_traversal = false;  // fill RayCounts::nodes_visited / prims_tested (costs time; off for timed baselines)
};

// One sample of one pixel: samplers/random_sampler.rs:50-74
static inline Vec3 sample_pixel(const Camera& cam, const Bvh& bvh, const RenderOpts& o, uint64_t pixel_i, uint32_t sample,
                                RayCounts& rc) {
  uint64_t x = pixel_i % o.width;
  uint64_t y = (pixel_i - x) / o.width;
  g_rng.path((uint32_t)pixel_i, sample);
  g_rng.select(0, RNG_JITTER);
  Float u = (g_rng.next_float01() + (Float)x) / (Float)(o.width - 1);
  Float v = 1.0f - (g_rng.next_float01() + (Float)y) / (Float)(o.height - 1);
  Ray ray = cam.get_ray(u, v);
  return o.method == PTB_METHOD_NAIVE ? naive_get_colour(ray, bvh, o.integ, rc) : mis_get_colour(ray, bvh, o.integ, rc);
}

// samplers/random_sampler.rs:23-99 + the running mean of src/main.rs:175-191, restated as a sum:
// `accum` (W*H*3, zero-initialised by the caller or carrying earlier samples) receives the SUM over the
// rendered samples; 10 000-pixel chunks are handed to worker threads like rayon's par_chunks_mut.
static inline RayCounts sample_image(const Camera& cam, const Bvh& bvh, const RenderOpts& o, float* accum) {
  const uint64_t pixel_num = (uint64_t)o.width * o.height;
  const uint64_t pixel_chunk_size = 10000;
  const uint64_t n_chunks = (pixel_num + pixel_chunk_size - 1) / pixel_chunk_size;
  unsigned nthreads = o.threads ? o.threads : std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  std::vector<RayCounts> per_thread(nthreads);
  for (auto& rc : per_thread) rc.count_traversal = o.count_traversal;
  for (uint32_t s = 0; s < o.spp; ++s) {
    std::atomic<uint64_t> next_chunk(0);
    auto worker = [&](unsigned tid) {
      g_rng.seed(o.seed);
      RayCounts& rc = per_thread[tid];
      for (;;) {
        uint64_t c = next_chunk.fetch_add(1);
        if (c >= n_chunks) break;
        uint64_t begin = c * pixel_chunk_size;
        uint64_t end = begin + pixel_chunk_size < pixel_num ? begin + pixel_chunk_size : pixel_num;
        for (uint64_t pixel_i = begin; pixel_i < end; ++pixel_i) {
          Vec3 rgb = sample_pixel(cam, bvh, o, pixel_i, o.sample_offset + s, rc);
          accum[pixel_i * 3 + 0] += rgb.x;
          accum[pixel_i * 3 + 1] += rgb.y;
          accum[pixel_i * 3 + 2] += rgb.z;
        }
      }
    };
    if (nthreads == 1) {
      worker(0);
    } else {
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
      for (auto& th : pool) th.join();
    }
  }
  RayCounts total;
  for (const auto& rc : per_thread) {
    total.reference += rc.reference;
    total.camera += rc.camera;
    total.bounce += rc.bounce;
    total.shadow_light += rc.shadow_light;
    total.shadow_sky += rc.shadow_sky;
    total.nodes_visited += rc.nodes_visited;
    total.prims_tested += rc.prims_tested;
  }
  return total;
}

}  // namespace ref
