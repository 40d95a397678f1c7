// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_math.hpp header).
// C entry points (ctypes-friendly) over the CPU restatement. Used by tests/, smoke() and bench.py's CPU legs.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

#include "cwbvh_ref.hpp"
#include "lbvh_ref.hpp"
#include "ploc_ref.hpp"
#include "sah_ref.hpp"
#include "ref_integrators.hpp"

namespace ref {
thread_local RngCursor g_rng;
}

using namespace ref;

struct orc_scene {
  std::vector<Texture> textures;
  std::vector<Material> materials;
  std::vector<std::vector<Float>> texture_words;  // owned copies of image pixels / perlin tables
  std::vector<Prim> prims_original;  // loader order: spheres first, then triangles
  Bvh bvh;
  Camera camera;
  Lbvh lbvh;
  bool lbvh_built = false;
  Cwbvh cw;
  double build_seconds = 0.0;
};

static inline Vec3 v3(const ptb_vec3& v) { return Vec3(v.x, v.y, v.z); }

template <class F>
static void parallel_for(size_t n, int threads, F f) {
  unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt == 1 || n < 1024) {
    f(0, (size_t)0, n);
    return;
  }
  std::atomic<size_t> next(0);
  const size_t grain = 4096;
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nt; ++t)
    pool.emplace_back([&, t]() {
      for (;;) {
        size_t b = next.fetch_add(grain);
        if (b >= n) break;
        size_t e = b + grain < n ? b + grain : n;
        f(t, b, e);
      }
    });
  for (auto& th : pool) th.join();
}

extern "C" {

orc_scene* orc_scene_create(const ptb_sphere* spheres, size_t ns, const ptb_triangle* tris, size_t nt,
                            const ptb_material* mats, size_t nm, const ptb_texture* texs, size_t ntex,
                            const ptb_camera* cam, const ptb_sky* sky, int split_type, const float* const* tex_data,
                            const uint32_t* tex_dims /* width, height, n_words per texture */) {
  orc_scene* s = new orc_scene();
  s->textures.resize(ntex);
  s->texture_words.resize(ntex);
  for (size_t i = 0; i < ntex; ++i) {
    s->textures[i].kind = texs[i].kind;
    s->textures[i].a = v3(texs[i].a);
    s->textures[i].b = v3(texs[i].b);
    if (tex_data && tex_data[i]) {
      s->texture_words[i].assign(tex_data[i], tex_data[i] + tex_dims[3 * i + 2]);
      s->textures[i].data = s->texture_words[i].data();
      s->textures[i].perm = reinterpret_cast<const uint32_t*>(s->texture_words[i].data()) + 256;
      s->textures[i].width = tex_dims[3 * i];
      s->textures[i].height = tex_dims[3 * i + 1];
    }
  }
  s->materials.resize(nm);
  for (size_t i = 0; i < nm; ++i) {
    s->materials[i].kind = mats[i].kind;
    s->materials[i].texture = &s->textures[mats[i].texture];
    s->materials[i].param = mats[i].param;
    s->materials[i].ior = v3(mats[i].ior);
    s->materials[i].metallic = mats[i].metallic;
  }
  s->prims_original.reserve(ns + nt);
  for (size_t i = 0; i < ns; ++i) {
    Prim p{};
    p.is_sphere = 1;
    p.orig_id = (uint32_t)i;
    p.material = &s->materials[spheres[i].material];
    p.center = v3(spheres[i].center);
    p.radius = spheres[i].radius;
    s->prims_original.push_back(p);
  }
  for (size_t i = 0; i < nt; ++i) {
    Prim p{};
    p.is_sphere = 0;
    p.orig_id = (uint32_t)(ns + i);
    p.material = &s->materials[tris[i].material];
    for (int k = 0; k < 3; ++k) { p.p[k] = v3(tris[i].p[k]); p.n[k] = v3(tris[i].n[k]); }
    s->prims_original.push_back(p);
  }
  if (cam) {
    s->camera.origin = v3(cam->origin);
    s->camera.lower_left = v3(cam->lower_left);
    s->camera.horizontal = v3(cam->horizontal);
    s->camera.vertical = v3(cam->vertical);
  }
  if (sky) s->bvh.sky.init(&s->textures[sky->texture], sky->sampler_res_x, sky->sampler_res_y);
  if (split_type >= 0) {
    auto t0 = std::chrono::steady_clock::now();
    s->bvh.build(s->prims_original, (SplitType)split_type);
    s->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return s;
}

void orc_scene_destroy(orc_scene* s) { delete s; }
size_t orc_bvh_num_nodes(const orc_scene* s) { return s->bvh.nodes.size(); }
size_t orc_bvh_depth(const orc_scene* s) { return s->bvh.max_depth_seen; }
double orc_bvh_build_seconds(const orc_scene* s) { return s->build_seconds; }
size_t orc_num_lights(const orc_scene* s) { return s->bvh.lights.size(); }
// BVH-order -> original primitive id (the permutation the reference discards, acceleration/mod.rs:79-82)
void orc_bvh_order(const orc_scene* s, uint32_t* out) {
  for (size_t i = 0; i < s->bvh.primitives.size(); ++i) out[i] = s->bvh.primitives[i].orig_id;
}

static inline void fill_hit(ptb_hit& o, bool have, const Hit& h, uint32_t prim) {
  if (have) { o.t = h.t; o.prim = prim; o.u = h.b1; o.v = h.b2; }
  else { o.t = 0.0f; o.prim = PTB_MISS; o.u = 0.0f; o.v = 0.0f; }
}

// check_hit with the reference's SAH tree + BFS candidates (acceleration/mod.rs:265-298)
void orc_closest_hit(const orc_scene* s, const ptb_ray* rays, size_t n, ptb_hit* out, int threads, uint64_t* counts2) {
  unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
  std::vector<uint64_t> nv(nt ? nt : 1, 0), pt(nt ? nt : 1, 0);
  parallel_for(n, threads, [&](unsigned tid, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      Ray ray(Vec3(rays[i].ox, rays[i].oy, rays[i].oz), Vec3(rays[i].dx, rays[i].dy, rays[i].dz), 0.0f);
      Hit h;
      const Material* m;
      size_t idx = s->bvh.check_hit(ray, h, m, &nv[tid], &pt[tid]);
      fill_hit(out[i], idx != Bvh::MISS, h, idx != Bvh::MISS ? s->bvh.primitives[idx].orig_id : PTB_MISS);
    }
  });
  if (counts2) {
    counts2[0] = counts2[1] = 0;
    for (size_t i = 0; i < nv.size(); ++i) { counts2[0] += nv[i]; counts2[1] += pt[i]; }
  }
}

// every primitive, no acceleration structure: the f32 ground truth of "min t over all primitives"
void orc_closest_hit_brute(const orc_scene* s, const ptb_ray* rays, size_t n, ptb_hit* out, int threads) {
  parallel_for(n, threads, [&](unsigned, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      Ray ray(Vec3(rays[i].ox, rays[i].oy, rays[i].oz), Vec3(rays[i].dx, rays[i].dy, rays[i].dz), 0.0f);
      Hit best, h;
      uint32_t bp = PTB_MISS;
      for (const Prim& p : s->prims_original) {
        if (p.get_int(ray, h) && h.t > 0.0f && (bp == PTB_MISS || h.t < best.t)) { best = h; bp = p.orig_id; }
      }
      fill_hit(out[i], bp != PTB_MISS, best, bp);
    }
  });
}

// INDEPENDENT intersector: double precision, a different formulation from the reference's — Möller &
// Trumbore 1997 ("Fast, minimum storage ray/triangle intersection": edge vectors, scalar triple products, no shear, no
// axis permutation, no error bounds) for triangles and the textbook quadratic for spheres — over EVERY primitive, no
// acceleration structure. It shares no arithmetic with triangle.rs:105-216 / sphere.rs:34-105 or with the slab test, so
// agreement on non-degenerate rays pins the f32 watertight path against something other than a copy of itself.
//   out[i]     = closest hit (t, original primitive id, barycentrics b1 b2 as the reference reports them), f64 rounded to f32
//   margin[i]  = how far the ray is from a decision the two formulations may legitimately take differently:
//                min( smallest barycentric coordinate of the winning triangle            (edge / vertex grazing),
//                     sqrt(discriminant) / radius of the winning sphere                  (silhouette grazing),
//                     relative t gap to the runner-up hit                                (near ties),
//                     -(largest "smallest barycentric" over the triangles just missed)   (a grazing near-miss) )
//                Tests compare ids only where margin > a stated threshold.
void orc_closest_hit_f64(const orc_scene* s, const ptb_ray* rays, size_t n, ptb_hit* out, float* margin, int threads) {
  struct D3 { double x, y, z; };
  auto sub = [](D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; };
  auto dot = [](D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
  auto cross = [](D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; };
  auto d3 = [](const Vec3& v) { return D3{(double)v.x, (double)v.y, (double)v.z}; };
  parallel_for(n, threads, [&](unsigned, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      const D3 o{rays[i].ox, rays[i].oy, rays[i].oz};
      D3 d{rays[i].dx, rays[i].dy, rays[i].dz};
      // the reference normalises in f32 (ray.rs:14) and reports t along that unit vector: do the same, then widen
      {
        Vec3 df = normalised(Vec3(rays[i].dx, rays[i].dy, rays[i].dz));
        d = d3(df);
      }
      double best_t = 1e300, second_t = 1e300, best_m = 1e300, near_miss = -1e300, bb1 = 0.0, bb2 = 0.0;
      uint32_t bp = PTB_MISS;
      for (const Prim& p : s->prims_original) {
        double t, m, b1 = 0.0, b2 = 0.0;
        if (p.is_sphere) {
          const D3 oc = sub(o, d3(p.center));
          const double r = (double)p.radius;
          const double hb = dot(oc, d), cc = dot(oc, oc) - r * r, dd = dot(d, d);
          const double disc = hb * hb - dd * cc;
          if (disc <= 0.0) { near_miss = std::max(near_miss, -std::sqrt(-disc) / (std::fabs(r) * std::sqrt(dd))); continue; }
          const double sq = std::sqrt(disc);
          double t0 = (-hb - sq) / dd, t1 = (-hb + sq) / dd;
          t = t0 > 0.0 ? t0 : t1;
          if (!(t > 0.0)) continue;
          m = sq / (std::fabs(r) * std::sqrt(dd));
        } else {
          const D3 v0 = d3(p.p[0]), e1 = sub(d3(p.p[1]), v0), e2 = sub(d3(p.p[2]), v0);
          const D3 pv = cross(d, e2);
          const double det = dot(e1, pv);
          if (det == 0.0) continue;
          const double inv = 1.0 / det;
          const D3 tv = sub(o, v0);
          const double u = dot(tv, pv) * inv;
          const D3 qv = cross(tv, e1);
          const double v = dot(d, qv) * inv;
          t = dot(e2, qv) * inv;
          const double w = 1.0 - u - v;
          const double mb = std::min(u, std::min(v, w));
          if (!(t > 0.0)) continue;
          if (mb < 0.0) { near_miss = std::max(near_miss, mb); continue; }
          m = mb;
          b1 = u; b2 = v;  // triangle.rs:149-151: b1, b2 weight p1, p2
        }
        if (t < best_t) { second_t = best_t; best_t = t; best_m = m; bp = p.orig_id; bb1 = b1; bb2 = b2; }
        else if (t < second_t) second_t = t;
      }
      double mg = -near_miss;
      if (bp != PTB_MISS) {
        mg = std::min(mg, best_m);
        if (second_t < 1e299) mg = std::min(mg, (second_t - best_t) / std::max(best_t, 1e-30));
        out[i].t = (float)best_t; out[i].prim = bp; out[i].u = (float)bb1; out[i].v = (float)bb2;
      } else {
        out[i].t = 0.0f; out[i].prim = PTB_MISS; out[i].u = out[i].v = 0.0f;
      }
      if (margin) margin[i] = (float)std::min(mg, 1e30);
    }
  });
}

// full hit record of one ray against the reference tree (point, normal, error, out) for unit tests
int orc_hit_record(const orc_scene* s, const ptb_ray* r, float out[12]) {
  Ray ray(Vec3(r->ox, r->oy, r->oz), Vec3(r->dx, r->dy, r->dz), 0.0f);
  Hit h;
  const Material* m;
  size_t idx = s->bvh.check_hit(ray, h, m);
  out[0] = h.t;
  out[1] = h.point.x; out[2] = h.point.y; out[3] = h.point.z;
  out[4] = h.normal.x; out[5] = h.normal.y; out[6] = h.normal.z;
  out[7] = h.error.x; out[8] = h.error.y; out[9] = h.error.z;
  out[10] = h.out ? 1.0f : 0.0f;
  out[11] = idx == Bvh::MISS ? -1.0f : (float)s->bvh.primitives[idx].orig_id;
  return idx != Bvh::MISS;
}

// RandomSampler::sample_image restated as a SUM into accum (W*H*3). counts: reference, camera, bounce,
// shadow_light, shadow_sky, nodes_visited, prims_tested. Returns wall seconds.
double orc_render(const orc_scene* s, const ptb_render_opts* o, float* accum, int threads, uint64_t counts[7]) {
  RenderOpts ro;
  ro.width = o->width;
  ro.height = o->height;
  ro.spp = o->samples_per_pixel;
  ro.sample_offset = o->sample_offset;
  ro.method = o->method;
  ro.seed = o->seed;
  ro.integ.max_depth = o->max_depth ? o->max_depth : MAX_DEPTH;
  ro.integ.rr_threshold = o->rr_threshold == PTB_RR_DEFAULT ? RUSSIAN_ROULETTE_THRESHOLD : o->rr_threshold;
  ro.threads = threads > 0 ? (unsigned)threads : 0;
  ro.count_traversal = (o->flags & 1u) != 0;  // ptb_render_opts::flags is reserved in the product ABI; the oracle uses bit 0
  auto t0 = std::chrono::steady_clock::now();
  RayCounts rc = sample_image(s->camera, s->bvh, ro, accum);
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (counts) {
    counts[0] = rc.reference; counts[1] = rc.camera; counts[2] = rc.bounce; counts[3] = rc.shadow_light;
    counts[4] = rc.shadow_sky; counts[5] = rc.nodes_visited; counts[6] = rc.prims_tested;
  }
  return dt;
}

// Mean radiance of ONE fixed ray over n samples (the shape of the reference's disabled furnace / MIS tests,
// implementations/tests/sampling.rs:181-297). Sample k uses the RNG path (pixel = k mod 2^20, sample = k >> 20).
void orc_radiance(const orc_scene* s, const ptb_ray* r, uint32_t method, uint64_t n, uint64_t seed, int threads,
                  double out[3]) {
  unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  std::vector<double> acc(3 * (size_t)nt, 0.0);
  IntegratorOpts io;
  parallel_for((size_t)n, threads, [&](unsigned tid, size_t b, size_t e) {
    g_rng.seed(seed);
    RayCounts rc;
    for (size_t k = b; k < e; ++k) {
      g_rng.path((uint32_t)(k & 0xFFFFFu), (uint32_t)(k >> 20));
      Ray ray(Vec3(r->ox, r->oy, r->oz), Vec3(r->dx, r->dy, r->dz), 0.0f);
      Vec3 c = method == PTB_METHOD_NAIVE ? naive_get_colour(ray, s->bvh, io, rc) : mis_get_colour(ray, s->bvh, io, rc);
      acc[3 * tid + 0] += c.x; acc[3 * tid + 1] += c.y; acc[3 * tid + 2] += c.z;
    }
  });
  out[0] = out[1] = out[2] = 0.0;
  for (unsigned t = 0; t < nt; ++t) { out[0] += acc[3 * t]; out[1] += acc[3 * t + 1]; out[2] += acc[3 * t + 2]; }
  out[0] /= (double)n; out[1] /= (double)n; out[2] /= (double)n;
}

// ------------------------------------------------------------------ LBVH
int orc_lbvh_build(orc_scene* s) {
  s->lbvh.build(s->prims_original);
  s->lbvh_built = true;
  return 0;
}
size_t orc_lbvh_num_nodes(const orc_scene* s) { return s->lbvh.nodes.size(); }
void orc_lbvh_quantise(const orc_scene* s, float frame[6], uint32_t* words) {
  std::vector<uint32_t> w;
  s->lbvh.quantise(frame, w);
  if (words) for (size_t i = 0; i < w.size(); ++i) words[i] = w[i];
}
// experiment: snaps every child box of the LBVH outwards onto a 65536^3 grid over the scene box, plus `extra` grid steps
// per side (what a 16-bit quantised node would let the traversal see)
void orc_lbvh_snap16(orc_scene* s, float extra) {
  Lbvh& l = s->lbvh;
  if (l.nodes.empty()) return;
  float mn[3], mx[3];
  const ptb_bvh_node& r = l.nodes[0];
  for (int k = 0; k < 3; ++k) { mn[k] = std::fmin(r.lmin[k], r.rmin[k]); mx[k] = std::fmax(r.lmax[k], r.rmax[k]); }
  for (ptb_bvh_node& nd : l.nodes) {
    float* lo[2] = {nd.lmin, nd.rmin};
    float* hi[2] = {nd.lmax, nd.rmax};
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < 3; ++k) {
        const double step = ((double)mx[k] - (double)mn[k]) / 65535.0;
        if (!(step > 0.0)) continue;
        const double ql = std::floor(((double)lo[c][k] - mn[k]) / step) - extra, qh = std::ceil(((double)hi[c][k] - mn[k]) / step) + extra;
        lo[c][k] = (float)(mn[k] + ql * step);
        hi[c][k] = (float)(mn[k] + qh * step);
      }
  }
}
// replaces the LBVH's Karras hierarchy by the PLOC hierarchy over the same Morton order (ploc_ref.hpp); returns the rounds
int orc_lbvh_ploc(orc_scene* s, int radius) {
  if (!s->lbvh_built) orc_lbvh_build(s);
  return ploc_rebuild(s->lbvh, radius);
}
// replaces the LBVH's hierarchy AND primitive order by the binned / swept SAH tree of sah_ref.hpp (the CPU statement of
// csrc/sah_build.cu); stats6: levels of large tasks, most large tasks in a level, small tasks, halving splits, depth of the
// deepest leaf, splits replaced by halving because of the depth bound
int orc_lbvh_sah(orc_scene* s, int nbins, uint32_t max_depth, uint32_t* stats4) {
  if (!s->lbvh_built) orc_lbvh_build(s);
  const SahStats st = sah_rebuild(s->lbvh, nbins, max_depth);
  if (stats4) { stats4[0] = st.levels; stats4[1] = st.max_tasks; stats4[2] = st.small_tasks; stats4[3] = st.fallbacks; stats4[4] = st.max_depth; stats4[5] = st.depth_limited; }
  return 0;
}
void orc_lbvh_export(const orc_scene* s, uint32_t* morton, uint32_t* prim_sorted, ptb_bvh_node* nodes) {
  const Lbvh& l = s->lbvh;
  if (morton) for (size_t i = 0; i < l.morton.size(); ++i) morton[i] = l.morton[i];
  if (prim_sorted) for (size_t i = 0; i < l.prim_sorted.size(); ++i) prim_sorted[i] = l.prim_sorted[i];
  if (nodes) for (size_t i = 0; i < l.nodes.size(); ++i) nodes[i] = l.nodes[i];
}
// counts2: nodes fetched, primitives tested (summed over the batch)
// nodes fetched by the ordered LBVH traversal, per ray (finds rays whose traversal is pathologically long)
void orc_lbvh_node_counts(const orc_scene* s, const ptb_ray* rays, size_t n, uint32_t* nodes_out, int threads) {
  parallel_for(n, threads, [&](unsigned, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      Ray ray(Vec3(rays[i].ox, rays[i].oy, rays[i].oz), Vec3(rays[i].dx, rays[i].dy, rays[i].dz), 0.0f);
      Hit h;
      uint32_t prim;
      uint64_t nv = 0, pt = 0;
      s->lbvh.closest_hit(ray, h, prim, &nv, &pt);
      nodes_out[i] = (uint32_t)nv;
    }
  });
}
void orc_lbvh_closest_hit(const orc_scene* s, const ptb_ray* rays, size_t n, ptb_hit* out, int threads, uint64_t* counts2) {
  unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
  std::vector<uint64_t> nv(nt ? nt : 1, 0), pt(nt ? nt : 1, 0);
  parallel_for(n, threads, [&](unsigned tid, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      Ray ray(Vec3(rays[i].ox, rays[i].oy, rays[i].oz), Vec3(rays[i].dx, rays[i].dy, rays[i].dz), 0.0f);
      Hit h;
      uint32_t prim;
      bool have = s->lbvh.closest_hit(ray, h, prim, &nv[tid], &pt[tid]);
      fill_hit(out[i], have, h, prim);
    }
  });
  if (counts2) {
    counts2[0] = counts2[1] = 0;
    for (size_t i = 0; i < nv.size(); ++i) { counts2[0] += nv[i]; counts2[1] += pt[i]; }
  }
}

// ------------------------------------------------------------------ compressed 8-wide BVH (cwbvh_ref.hpp)
size_t orc_cw_build(orc_scene* s, int max_leaf) {
  if (!s->lbvh_built) { s->lbvh.build(s->prims_original); s->lbvh_built = true; }
  s->cw.build(s->lbvh, max_leaf);
  return s->cw.nodes.size();
}
// nodes: n x 96 bytes in the device layout; slot_prim: final primitive order -> original id
void orc_cw_export(const orc_scene* s, void* nodes, uint32_t* slot_prim) {
  if (nodes) std::memcpy(nodes, s->cw.nodes.data(), s->cw.nodes.size() * sizeof(CwNode));
  if (slot_prim) for (size_t i = 0; i < s->cw.slot_prim.size(); ++i) slot_prim[i] = s->cw.slot_prim[i];
}
void orc_cw_closest_hit(const orc_scene* s, const ptb_ray* rays, size_t n, ptb_hit* out, int threads, uint64_t* counts2) {
  unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
  std::vector<uint64_t> nv(nt ? nt : 1, 0), pt(nt ? nt : 1, 0);
  parallel_for(n, threads, [&](unsigned tid, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      Ray ray(Vec3(rays[i].ox, rays[i].oy, rays[i].oz), Vec3(rays[i].dx, rays[i].dy, rays[i].dz), 0.0f);
      Hit h;
      uint32_t prim;
      bool have = s->cw.closest_hit(s->lbvh, ray, h, prim, &nv[tid], &pt[tid]);
      fill_hit(out[i], have, h, prim);
    }
  });
  if (counts2) {
    counts2[0] = counts2[1] = 0;
    for (size_t i = 0; i < nv.size(); ++i) { counts2[0] += nv[i]; counts2[1] += pt[i]; }
  }
}

// Ordered, t-culled traversal of the REFERENCE's SAH tree (not something the reference does — it is BFS un-culled):
// measures how many 2-child node fetches a SAH-quality tree would need for the same rays, to judge LBVH quality.
// counts2: internal nodes expanded, primitives tested.
void orc_sah_ordered_closest_hit(const orc_scene* s, const ptb_ray* rays, size_t n, ptb_hit* out, int threads, uint64_t* counts2) {
  unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
  std::vector<uint64_t> nv(nt ? nt : 1, 0), pt(nt ? nt : 1, 0);
  const Bvh& bvh = s->bvh;
  parallel_for(n, threads, [&](unsigned tid, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      Ray ray(Vec3(rays[i].ox, rays[i].oy, rays[i].oz), Vec3(rays[i].dx, rays[i].dy, rays[i].dz), 0.0f);
      Hit best, h;
      Float best_t = INF_F;
      uint32_t bp = PTB_MISS;
      const Lbvh::SlabRay slab = Lbvh::make_slab_ray(ray);
      size_t stack[256];
      Float stack_t[256];
      int sp = 0;
      size_t cur = 0;
      bool have_cur = !bvh.nodes.empty();
      while (have_cur) {
        const Node& nd = bvh.nodes[cur];
        have_cur = false;
        if (!nd.has_children) {
          for (size_t k = nd.primitive_offset; k < nd.primitive_offset + nd.number_primitives; ++k) {
            ++pt[tid];
            if (bvh.primitives[k].get_int(ray, h) && h.t > 0.0f && h.t < best_t) { best_t = h.t; best = h; bp = bvh.primitives[k].orig_id; }
          }
        } else {
          ++nv[tid];
          const Node& l = bvh.nodes[nd.children[0]];
          const Node& r = bvh.nodes[nd.children[1]];
          float lmn[3] = {l.bounds.min.x, l.bounds.min.y, l.bounds.min.z}, lmx[3] = {l.bounds.max.x, l.bounds.max.y, l.bounds.max.z};
          float rmn[3] = {r.bounds.min.x, r.bounds.min.y, r.bounds.min.z}, rmx[3] = {r.bounds.max.x, r.bounds.max.y, r.bounds.max.z};
          Float tl, tr;
          bool hl = Lbvh::box_hit(lmn, lmx, slab, best_t, tl), hr = Lbvh::box_hit(rmn, rmx, slab, best_t, tr);
          if (hl && hr) {
            size_t nearc = nd.children[0], farc = nd.children[1];
            Float tf = tr;
            if (tr < tl) { nearc = nd.children[1]; farc = nd.children[0]; tf = tl; }
            stack[sp] = farc; stack_t[sp] = tf; ++sp;
            cur = nearc; have_cur = true;
          } else if (hl) { cur = nd.children[0]; have_cur = true; }
          else if (hr) { cur = nd.children[1]; have_cur = true; }
        }
        while (!have_cur && sp > 0) {
          --sp;
          if (stack_t[sp] <= best_t) { cur = stack[sp]; have_cur = true; }
        }
      }
      fill_hit(out[i], bp != PTB_MISS, best, bp);
    }
  });
  if (counts2) {
    counts2[0] = counts2[1] = 0;
    for (size_t i = 0; i < nv.size(); ++i) { counts2[0] += nv[i]; counts2[1] += pt[i]; }
  }
}

// ------------------------------------------------------------- KAT hooks
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { Philox::block(ctr, key, out); }
void orc_sort_by_indices_u32(uint32_t* values, const uint64_t* indices, size_t n) {
  std::vector<uint32_t> v(values, values + n);
  std::vector<size_t> idx(indices, indices + n);
  sort_by_indices(v, idx);
  for (size_t i = 0; i < n; ++i) values[i] = v[i];
}
float orc_next_float(float f) { return next_float(f); }
float orc_previous_float(float f) { return previous_float(f); }
float orc_gamma(uint32_t n) { return gamma(n); }
void orc_offset_ray(const float o[3], const float nrm[3], const float err[3], int is_brdf, float out[3]) {
  Vec3 r = offset_ray(Vec3(o[0], o[1], o[2]), Vec3(nrm[0], nrm[1], nrm[2]), Vec3(err[0], err[1], err[2]), is_brdf != 0);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// Ray::new: out = direction(3), d_inverse(3), shear(3)
void orc_ray_new(const float o[3], const float d[3], float out[9]) {
  Ray r(Vec3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2]), 0.0f);
  out[0] = r.direction.x; out[1] = r.direction.y; out[2] = r.direction.z;
  out[3] = r.d_inverse.x; out[4] = r.d_inverse.y; out[5] = r.d_inverse.z;
  out[6] = r.shear.x; out[7] = r.shear.y; out[8] = r.shear.z;
}
// Coordinate: out1 = from(to(v)), out2 = to(from(v))   (utility/coord.rs:39-49)
void orc_coord_roundtrip(const float z[3], const float v[3], float out1[3], float out2[3]) {
  Coordinate to = Coordinate::new_from_z(Vec3(z[0], z[1], z[2]));
  Coordinate from = to.create_inverse();
  Vec3 vv(v[0], v[1], v[2]);
  Vec3 a = from.to_coord(to.to_coord(vv)), b = to.to_coord(from.to_coord(vv));
  out1[0] = a.x; out1[1] = a.y; out1[2] = a.z;
  out2[0] = b.x; out2[1] = b.y; out2[2] = b.z;
}
void orc_camera_make(const float origin[3], const float lookat[3], const float vup[3], float fov, float aspect,
                     float aperture, float focus, ptb_camera* out) {
  Camera c = make_camera(Vec3(origin[0], origin[1], origin[2]), Vec3(lookat[0], lookat[1], lookat[2]),
                         Vec3(vup[0], vup[1], vup[2]), fov, aspect, aperture, focus);
  out->origin = {c.origin.x, c.origin.y, c.origin.z};
  out->lower_left = {c.lower_left.x, c.lower_left.y, c.lower_left.z};
  out->horizontal = {c.horizontal.x, c.horizontal.y, c.horizontal.z};
  out->vertical = {c.vertical.x, c.vertical.y, c.vertical.z};
}
// camera ray for pixel coordinates (u, v): out = origin(3), normalised direction(3)
void orc_camera_ray(const orc_scene* s, float u, float v, float out[6]) {
  Ray r = s->camera.get_ray(u, v);
  out[0] = r.origin.x; out[1] = r.origin.y; out[2] = r.origin.z;
  out[3] = r.direction.x; out[4] = r.direction.y; out[5] = r.direction.z;
}
// lambertian::sample / pdf (statistics/bxdfs/lambertian.rs:5-22); sample k uses RNG path (k, 0), test stream
void orc_lambertian_sample(const float normal[3], uint64_t seed, size_t n, int local, float* dirs) {
  g_rng.seed(seed);
  Vec3 nn(normal[0], normal[1], normal[2]);
  for (size_t k = 0; k < n; ++k) {
    g_rng.path((uint32_t)k, (uint32_t)(k >> 32));
    g_rng.select(0, RNG_TEST);
    Vec3 d = local ? lambertian::sample_local() : lambertian::sample(nn);
    dirs[3 * k] = d.x; dirs[3 * k + 1] = d.y; dirs[3 * k + 2] = d.z;
  }
}
void orc_lambertian_pdf(const float normal[3], const float* dirs, size_t n, int local, float* pdf) {
  Vec3 nn(normal[0], normal[1], normal[2]);
  for (size_t k = 0; k < n; ++k) {
    Vec3 d(dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]);
    pdf[k] = local ? lambertian::pdf_local(d) : lambertian::pdf(d, nn);
  }
}
// Trowbridge-Reitz (statistics/bxdfs/trowbridge_reitz.rs, trowbridge_reitz_vndf.rs isotropic). which: 0 = sample_vndf /
// vndf (half vectors, local frame), 1 = sample_local / pdf_local, 2 = sample / pdf about `normal`.
void orc_tr_sample(float alpha, const float incoming[3], const float normal[3], uint64_t seed, size_t n, int which, float* dirs) {
  g_rng.seed(seed);
  Vec3 in(incoming[0], incoming[1], incoming[2]), nn(normal[0], normal[1], normal[2]);
  for (size_t k = 0; k < n; ++k) {
    g_rng.path((uint32_t)k, (uint32_t)(k >> 32));
    g_rng.select(0, RNG_TEST);
    Vec3 d = which == 0 ? tr::sample_vndf(alpha, alpha, in) : which == 1 ? tr::sample_local(alpha, in) : tr::sample(alpha, in, nn);
    dirs[3 * k] = d.x; dirs[3 * k + 1] = d.y; dirs[3 * k + 2] = d.z;
  }
}
void orc_tr_pdf(float alpha, const float incoming[3], const float normal[3], const float* dirs, size_t n, int which, float* pdf) {
  Vec3 in(incoming[0], incoming[1], incoming[2]), nn(normal[0], normal[1], normal[2]);
  for (size_t k = 0; k < n; ++k) {
    Vec3 d(dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]);
    pdf[k] = which == 0 ? tr::vndf(alpha, d, in) : which == 1 ? tr::pdf_local(alpha, in, d) : tr::pdf(alpha, in, d, nn);
  }
}
// integrands of the reference's GGX integration tests (trowbridge_reitz.rs:128-230), evaluated at `dirs`:
//   0 g1_cos_test  1 projected_area  2 weak_furnace  3 g2_test     (a = the fixed direction, normal = the frame's z)
void orc_tr_integrand(float alpha, const float a_[3], const float normal[3], const float* dirs, size_t n, int which, float* out) {
  Vec3 a(a_[0], a_[1], a_[2]), nn(normal[0], normal[1], normal[2]);
  for (size_t k = 0; k < n; ++k) {
    Vec3 b(dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]);
    Float v = 0.0f;
    if (which == 0) v = tr::g1(alpha, nn, b, a) * fmax_(a.dot(b), 0.0f) * tr::d(alpha, b.dot(nn));
    else if (which == 1) v = tr::d(alpha, b.dot(nn)) * b.dot(nn);
    else {
      Vec3 h = normalised(b + a);
      if (h.dot(nn) < 0.0f) h = -h;
      Float denom = 4.0f * std::fabs(a.dot(nn));
      if (denom >= 0.000000001f)
        v = (which == 2 ? tr::g1(alpha, nn, h, a) : tr::g2(alpha, nn, h, a, b)) * tr::d(alpha, h.dot(nn)) / denom;
    }
    out[k] = v;
  }
}
// Scatter methods of one material at a synthetic hit (normal, point): out = scattering_pdf, eval(3), eval_over_pdf(3)
void orc_material_terms(const orc_scene* s, uint32_t mat, const float normal[3], const float point[3], const float wo[3],
                        const float wi[3], float out[7]) {
  Hit h;
  h.normal = Vec3(normal[0], normal[1], normal[2]);
  h.point = Vec3(point[0], point[1], point[2]);
  Vec3 o(wo[0], wo[1], wo[2]), i(wi[0], wi[1], wi[2]);
  const Material& m = s->materials[mat];
  out[0] = m.scattering_pdf(h, o, i);
  Vec3 e = m.eval(h, o, i), r = m.eval_over_scattering_pdf(h, o, i);
  out[1] = e.x; out[2] = e.y; out[3] = e.z; out[4] = r.x; out[5] = r.y; out[6] = r.z;
}
void orc_random_unit_vectors(uint64_t seed, size_t n, float* dirs) {
  g_rng.seed(seed);
  for (size_t k = 0; k < n; ++k) {
    g_rng.path((uint32_t)k, (uint32_t)(k >> 32));
    g_rng.select(0, RNG_TEST);
    Vec3 d = random_unit_vector();
    dirs[3 * k] = d.x; dirs[3 * k + 1] = d.y; dirs[3 * k + 2] = d.z;
  }
}
// Distribution1D (statistics/distributions.rs:11-72)
void orc_dist1d(const float* values, size_t n, float* pdf_out, float* cdf_out, uint64_t seed, size_t nsamples, uint64_t* counts) {
  Distribution1D d(values, n);
  if (pdf_out) for (size_t i = 0; i < n; ++i) pdf_out[i] = d.pdf[i];
  if (cdf_out) for (size_t i = 0; i <= n; ++i) cdf_out[i] = d.cdf[i];
  if (counts) {
    g_rng.seed(seed);
    for (size_t k = 0; k < nsamples; ++k) {
      g_rng.path((uint32_t)k, (uint32_t)(k >> 32));
      g_rng.select(0, RNG_TEST);
      counts[d.sample()] += 1;
    }
  }
}
// Distribution2D (statistics/distributions.rs:75-113): counts is dim_y x dim_x row-major, pdf_out likewise
void orc_dist2d(const float* values, size_t n, size_t width, float* pdf_out, uint64_t seed, size_t nsamples, uint64_t* counts) {
  std::vector<Float> v(values, values + n);
  Distribution2D d(v, width);
  if (pdf_out)
    for (size_t y = 0; y < d.dim_y; ++y)
      for (size_t x = 0; x < d.dim_x; ++x)
        pdf_out[y * width + x] = d.pdf(((Float)x + 0.5f) / (Float)d.dim_x, ((Float)y + 0.5f) / (Float)d.dim_y);
  if (counts) {
    g_rng.seed(seed);
    for (size_t k = 0; k < nsamples; ++k) {
      g_rng.path((uint32_t)k, (uint32_t)(k >> 32));
      g_rng.select(0, RNG_TEST);
      size_t u, vv;
      d.sample(u, vv);
      counts[vv * width + u] += 1;
    }
  }
}
// Sky::sample / Sky::pdf (sky.rs:43-78)
void orc_sky_sample(const orc_scene* s, uint64_t seed, size_t n, float* dirs) {
  g_rng.seed(seed);
  for (size_t k = 0; k < n; ++k) {
    g_rng.path((uint32_t)k, (uint32_t)(k >> 32));
    g_rng.select(0, RNG_TEST);
    Vec3 d = s->bvh.sky.sample();
    dirs[3 * k] = d.x; dirs[3 * k + 1] = d.y; dirs[3 * k + 2] = d.z;
  }
}
void orc_sky_pdf(const orc_scene* s, const float* dirs, size_t n, float* pdf) {
  for (size_t k = 0; k < n; ++k) pdf[k] = s->bvh.sky.pdf(Vec3(dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]));
}
void orc_texture_colour(const orc_scene* s, uint32_t tex, const float dir[3], const float point[3], float out[3]) {
  Vec3 c = s->textures[tex].colour_value(Vec3(dir[0], dir[1], dir[2]), Vec3(point[0], point[1], point[2]));
  out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
// the sky table as the device must hold it: y cdf (ry+1), x cdfs (ry*(rx+1)), y pdf (ry), x pdfs (ry*rx)
void orc_sky_table(const orc_scene* s, float* ycdf, float* xcdf, float* ypdf, float* xpdf) {
  const Sky& sky = s->bvh.sky;
  if (!sky.has_distribution) return;
  const Distribution2D& d = sky.distribution;
  for (size_t i = 0; i <= d.dim_y; ++i) ycdf[i] = d.y_distribution.cdf[i];
  for (size_t i = 0; i < d.dim_y; ++i) ypdf[i] = d.y_distribution.pdf[i];
  for (size_t y = 0; y < d.dim_y; ++y) {
    for (size_t x = 0; x <= d.dim_x; ++x) xcdf[y * (d.dim_x + 1) + x] = d.x_distributions[y].cdf[x];
    for (size_t x = 0; x < d.dim_x; ++x) xpdf[y * d.dim_x + x] = d.x_distributions[y].pdf[x];
  }
}
unsigned orc_hardware_threads(void) { return std::thread::hardware_concurrency(); }

}  // extern "C"
