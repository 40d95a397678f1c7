"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so, the CPU restatement of the reference's hot path (see the headers of
oracle/ref_*.hpp for the reference file:line each function follows and for its parity-pinning status).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The Rust reference itself cannot be built in this image (no rustc/cargo, nightly features, crates.io
dependencies), so there is no oracle/_ref.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

SPLIT_SAH, SPLIT_MIDDLE, SPLIT_EQUAL_COUNTS = 0, 1, 2

ray_dtype = np.dtype([("o", "<f4", 3), ("_pad0", "<f4"), ("d", "<f4", 3), ("_pad1", "<f4")])
hit_dtype = np.dtype([("t", "<f4"), ("prim", "<u4"), ("u", "<f4"), ("v", "<f4")])
bvh_node_dtype = np.dtype([("lmin", "<f4", 3), ("lmax", "<f4", 3), ("rmin", "<f4", 3), ("rmax", "<f4", 3),
                           ("left", "<u4"), ("right", "<u4"), ("parent", "<u4"), ("_pad", "<u4")])


class RenderOpts(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples_per_pixel", C.c_uint32),
                ("sample_offset", C.c_uint32), ("method", C.c_uint32), ("max_depth", C.c_uint32),
                ("rr_threshold", C.c_uint32), ("flags", C.c_uint32), ("seed", C.c_uint64),
                ("row_begin", C.c_uint32), ("row_count", C.c_uint32)]


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def _load():
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.orc_scene_create.restype = P
    lib.orc_scene_create.argtypes = [P, C.c_size_t, P, C.c_size_t, P, C.c_size_t, P, C.c_size_t, P, P, C.c_int, P, P]
    lib.orc_scene_destroy.argtypes = [P]
    for n in ("orc_bvh_num_nodes", "orc_bvh_depth", "orc_num_lights", "orc_lbvh_num_nodes"):
        getattr(lib, n).restype = C.c_size_t
        getattr(lib, n).argtypes = [P]
    lib.orc_bvh_build_seconds.restype = C.c_double
    lib.orc_bvh_build_seconds.argtypes = [P]
    lib.orc_bvh_order.argtypes = [P, P]
    lib.orc_closest_hit.argtypes = [P, P, C.c_size_t, P, C.c_int, P]
    lib.orc_closest_hit_brute.argtypes = [P, P, C.c_size_t, P, C.c_int]
    lib.orc_closest_hit_f64.argtypes = [P, P, C.c_size_t, P, P, C.c_int]
    lib.orc_hit_record.argtypes = [P, P, P]
    lib.orc_hit_record.restype = C.c_int
    lib.orc_render.restype = C.c_double
    lib.orc_render.argtypes = [P, C.POINTER(RenderOpts), P, C.c_int, P]
    lib.orc_radiance.argtypes = [P, P, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, P]
    lib.orc_lbvh_build.argtypes = [P]
    lib.orc_lbvh_ploc.argtypes = [P, C.c_int]
    lib.orc_lbvh_snap16.argtypes = [P, C.c_float]
    lib.orc_lbvh_quantise.argtypes = [P, P, P]
    lib.orc_lbvh_ploc.restype = C.c_int
    lib.orc_lbvh_sah.argtypes = [P, C.c_int, C.c_uint32, P]
    lib.orc_lbvh_sah.restype = C.c_int
    lib.orc_lbvh_export.argtypes = [P, P, P, P]
    lib.orc_lbvh_closest_hit.argtypes = [P, P, C.c_size_t, P, C.c_int, P]
    lib.orc_lbvh_node_counts.argtypes = [P, P, C.c_size_t, P, C.c_int]
    lib.orc_cw_build.restype = C.c_size_t
    lib.orc_cw_build.argtypes = [P, C.c_int]
    lib.orc_cw_export.argtypes = [P, P, P]
    lib.orc_cw_closest_hit.argtypes = [P, P, C.c_size_t, P, C.c_int, P]
    lib.orc_sah_ordered_closest_hit.argtypes = [P, P, C.c_size_t, P, C.c_int, P]
    lib.orc_philox4x32_10.argtypes = [P, P, P]
    lib.orc_sort_by_indices_u32.argtypes = [P, P, C.c_size_t]
    for n in ("orc_next_float", "orc_previous_float"):
        getattr(lib, n).restype = C.c_float
        getattr(lib, n).argtypes = [C.c_float]
    lib.orc_gamma.restype = C.c_float
    lib.orc_gamma.argtypes = [C.c_uint32]
    lib.orc_offset_ray.argtypes = [P, P, P, C.c_int, P]
    lib.orc_ray_new.argtypes = [P, P, P]
    lib.orc_coord_roundtrip.argtypes = [P, P, P, P]
    lib.orc_camera_make.argtypes = [P, P, P, C.c_float, C.c_float, C.c_float, C.c_float, P]
    lib.orc_camera_ray.argtypes = [P, C.c_float, C.c_float, P]
    lib.orc_lambertian_sample.argtypes = [P, C.c_uint64, C.c_size_t, C.c_int, P]
    lib.orc_lambertian_pdf.argtypes = [P, P, C.c_size_t, C.c_int, P]
    lib.orc_random_unit_vectors.argtypes = [C.c_uint64, C.c_size_t, P]
    lib.orc_tr_sample.argtypes = [C.c_float, P, P, C.c_uint64, C.c_size_t, C.c_int, P]
    lib.orc_tr_pdf.argtypes = [C.c_float, P, P, P, C.c_size_t, C.c_int, P]
    lib.orc_tr_integrand.argtypes = [C.c_float, P, P, P, C.c_size_t, C.c_int, P]
    lib.orc_material_terms.argtypes = [P, C.c_uint32, P, P, P, P, P]
    lib.orc_dist1d.argtypes = [P, C.c_size_t, P, P, C.c_uint64, C.c_size_t, P]
    lib.orc_dist2d.argtypes = [P, C.c_size_t, C.c_size_t, P, C.c_uint64, C.c_size_t, P]
    lib.orc_sky_sample.argtypes = [P, C.c_uint64, C.c_size_t, P]
    lib.orc_sky_pdf.argtypes = [P, P, C.c_size_t, P]
    lib.orc_texture_colour.argtypes = [P, C.c_uint32, P, P, P]
    lib.orc_sky_table.argtypes = [P, P, P, P, P]
    lib.orc_hardware_threads.restype = C.c_uint
    return lib


lib = _load()


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def f3(v):
    return np.ascontiguousarray(v, dtype=np.float32)


class OracleScene:
    """The reference's Scene: primitives + materials + textures + camera + sky + SAH Bvh."""

    def __init__(self, host_scene, split_type: int = SPLIT_SAH):
        s = host_scene
        self._keep = [np.ascontiguousarray(a) for a in (s.spheres, s.triangles, s.materials, s.textures, s.camera, s.sky)]
        sp, tr, ma, te, cam, sky = self._keep
        self.n_prims = len(sp) + len(tr)
        tex_ptrs = (C.c_void_p * max(len(te), 1))()
        tex_dims = np.zeros(3 * max(len(te), 1), np.uint32)
        for i, (w, h, words) in getattr(s, "texture_data", {}).items():
            words = np.ascontiguousarray(words, dtype=np.float32)
            self._keep.append(words)
            tex_ptrs[i] = words.ctypes.data
            tex_dims[3 * i:3 * i + 3] = (w, h, words.size)
        self._h = lib.orc_scene_create(_p(sp), len(sp), _p(tr), len(tr), _p(ma), len(ma), _p(te), len(te), _p(cam), _p(sky),
                                       split_type, C.cast(tex_ptrs, C.c_void_p), _p(tex_dims))
        self.sky_res = (int(sky["sampler_res_x"][0]), int(sky["sampler_res_y"][0]))
        self._lbvh = False

    def __del__(self):
        if getattr(self, "_h", None):
            lib.orc_scene_destroy(self._h)
            self._h = None

    # -- reference BVH
    def num_nodes(self):
        return lib.orc_bvh_num_nodes(self._h)

    def depth(self):
        return lib.orc_bvh_depth(self._h)

    def build_seconds(self):
        return lib.orc_bvh_build_seconds(self._h)

    def num_lights(self):
        return lib.orc_num_lights(self._h)

    def bvh_order(self):
        out = np.zeros(self.n_prims, np.uint32)
        lib.orc_bvh_order(self._h, _p(out))
        return out

    def closest_hit(self, rays, threads=0, with_counts=False):
        rays = np.ascontiguousarray(rays, dtype=ray_dtype)
        out = np.zeros(len(rays), hit_dtype)
        counts = np.zeros(2, np.uint64)
        lib.orc_closest_hit(self._h, _p(rays), len(rays), _p(out), threads, _p(counts))
        return (out, counts) if with_counts else out

    def closest_hit_brute(self, rays, threads=0):
        rays = np.ascontiguousarray(rays, dtype=ray_dtype)
        out = np.zeros(len(rays), hit_dtype)
        lib.orc_closest_hit_brute(self._h, _p(rays), len(rays), _p(out), threads)
        return out

    def closest_hit_f64(self, rays, threads=0):
        """Independent f64 Moller-Trumbore / quadratic intersector over every primitive (no BVH): (hits, margin).
        `margin` is small where the f32 watertight path may legitimately decide differently (edge grazing, near ties)."""
        rays = np.ascontiguousarray(rays, dtype=ray_dtype)
        out = np.zeros(len(rays), hit_dtype)
        margin = np.zeros(len(rays), np.float32)
        lib.orc_closest_hit_f64(self._h, _p(rays), len(rays), _p(out), _p(margin), threads)
        return out, margin

    def sah_ordered_closest_hit(self, rays, threads=0):
        """Ordered, t-culled traversal of the reference's SAH tree: (hits, internal nodes expanded, prims tested)."""
        rays = np.ascontiguousarray(rays, dtype=ray_dtype)
        out = np.zeros(len(rays), hit_dtype)
        counts = np.zeros(2, np.uint64)
        lib.orc_sah_ordered_closest_hit(self._h, _p(rays), len(rays), _p(out), threads, _p(counts))
        return out, int(counts[0]), int(counts[1])

    def hit_record(self, origin, direction):
        r = np.zeros(1, ray_dtype)
        r["o"], r["d"] = origin, direction
        out = np.zeros(12, np.float32)
        hit = lib.orc_hit_record(self._h, _p(r), _p(out))
        return dict(hit=bool(hit), t=out[0], point=out[1:4].copy(), normal=out[4:7].copy(), error=out[7:10].copy(),
                    out=bool(out[10]), prim=int(out[11]))

    def render(self, width, height, spp, method, seed=0, sample_offset=0, threads=0, max_depth=50, rr_threshold=3,
               accum=None, count_traversal=False):
        """Returns (sum image (H,W,3) float32, counts dict, seconds). `count_traversal` fills nodes_visited /
        prims_tested (off by default: the reference keeps no such counters and timed legs must not pay for them)."""
        o = RenderOpts(width, height, spp, sample_offset, method, max_depth, rr_threshold, 1 if count_traversal else 0, seed)
        if accum is None:
            accum = np.zeros(width * height * 3, np.float32)
        counts = np.zeros(7, np.uint64)
        secs = lib.orc_render(self._h, C.byref(o), _p(accum), threads, _p(counts))
        names = ("reference", "camera", "bounce", "shadow_light", "shadow_sky", "nodes_visited", "prims_tested")
        return accum.reshape(height, width, 3), dict(zip(names, map(int, counts))), secs

    def radiance(self, origin, direction, method, n, seed=0, threads=0):
        r = np.zeros(1, ray_dtype)
        r["o"], r["d"] = origin, direction
        out = np.zeros(3, np.float64)
        lib.orc_radiance(self._h, _p(r), method, n, seed, threads, _p(out))
        return out

    def camera_ray(self, u, v):
        out = np.zeros(6, np.float32)
        lib.orc_camera_ray(self._h, u, v, _p(out))
        return out[:3].copy(), out[3:].copy()

    # -- sky / textures
    def sky_sample(self, n, seed=0):
        d = np.zeros((n, 3), np.float32)
        lib.orc_sky_sample(self._h, seed, n, _p(d))
        return d

    def sky_pdf(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32)
        out = np.zeros(len(dirs), np.float32)
        lib.orc_sky_pdf(self._h, _p(dirs), len(dirs), _p(out))
        return out

    def sky_table(self):
        """(ycdf, xcdf, ypdf, xpdf) of the sky's Distribution2D (sky.rs:20-37, distributions.rs:82-99)."""
        rx, ry = self.sky_res
        ycdf, xcdf = np.zeros(ry + 1, np.float32), np.zeros(ry * (rx + 1), np.float32)
        ypdf, xpdf = np.zeros(ry, np.float32), np.zeros(ry * rx, np.float32)
        lib.orc_sky_table(self._h, _p(ycdf), _p(xcdf), _p(ypdf), _p(xpdf))
        return ycdf, xcdf, ypdf, xpdf

    def material_terms(self, mat, normal, point, wo, wi):
        """(scattering_pdf, eval, eval_over_scattering_pdf) of material `mat` at a synthetic hit."""
        out = np.zeros(7, np.float32)
        lib.orc_material_terms(self._h, mat, _p(f3(normal)), _p(f3(point)), _p(f3(wo)), _p(f3(wi)), _p(out))
        return float(out[0]), out[1:4].copy(), out[4:7].copy()

    def texture_colour(self, tex, direction, point=(0, 0, 0)):
        out = np.zeros(3, np.float32)
        lib.orc_texture_colour(self._h, tex, _p(f3(direction)), _p(f3(point)), _p(out))
        return out

    # -- LBVH oracle
    def lbvh_build(self):
        lib.orc_lbvh_build(self._h)
        self._lbvh = True

    def lbvh_ploc(self, radius=16):
        """Replaces the LBVH's hierarchy by the PLOC hierarchy over the same Morton order (ploc_ref.hpp); returns the rounds."""
        self._lbvh = True
        return lib.orc_lbvh_ploc(self._h, radius)

    def lbvh_sah(self, nbins=8, max_depth=60):
        """Replaces the LBVH's hierarchy and primitive order by the SAH tree of sah_ref.hpp (CPU statement of the device's
        SAH builder). Returns (levels of large tasks, most large tasks in a level, small tasks, halving splits, depth of the
        deepest leaf, splits replaced by halving because of the depth bound)."""
        self._lbvh = True
        st = np.zeros(6, np.uint32)
        lib.orc_lbvh_sah(self._h, nbins, max_depth, _p(st))
        return tuple(int(x) for x in st)

    def lbvh_quantise(self):
        """(frame, (m, 8) uint32 words): the 32-byte traversal nodes of the LBVH, CPU definition (lbvh_ref.hpp quantise)."""
        if not self._lbvh:
            self.lbvh_build()
        m = lib.orc_lbvh_num_nodes(self._h)
        frame = np.zeros(6, np.float32)
        words = np.zeros((m, 8), np.uint32)
        lib.orc_lbvh_quantise(self._h, _p(frame), _p(words))
        return frame, words

    def lbvh_snap16(self, extra=0.0):
        """Experiment: child boxes snapped outwards onto a 65536^3 grid over the scene box, plus `extra` steps per side."""
        lib.orc_lbvh_snap16(self._h, extra)

    def lbvh_export(self):
        if not self._lbvh:
            self.lbvh_build()
        m = lib.orc_lbvh_num_nodes(self._h)
        morton = np.zeros(self.n_prims, np.uint32)
        prims = np.zeros(self.n_prims, np.uint32)
        nodes = np.zeros(m, bvh_node_dtype)
        lib.orc_lbvh_export(self._h, _p(morton), _p(prims), _p(nodes))
        return morton, prims, nodes

    def lbvh_closest_hit(self, rays, threads=0):
        """Returns (hits, nodes_fetched, prims_tested) of the ordered, t-culled traversal."""
        if not self._lbvh:
            self.lbvh_build()
        rays = np.ascontiguousarray(rays, dtype=ray_dtype)
        out = np.zeros(len(rays), hit_dtype)
        counts = np.zeros(2, np.uint64)
        lib.orc_lbvh_closest_hit(self._h, _p(rays), len(rays), _p(out), threads, _p(counts))
        return out, int(counts[0]), int(counts[1])


    # -- compressed 8-wide BVH oracle (cwbvh_ref.hpp)
    def cw_build(self, max_leaf=3):
        """Builds the compressed wide BVH from the LBVH; returns the number of 96-byte nodes."""
        self._lbvh = True
        self._cw_nodes = lib.orc_cw_build(self._h, max_leaf)
        return self._cw_nodes

    def cw_export(self):
        """(nodes as (n, 96) uint8 in the device layout, slot_prim: final primitive order -> original id)."""
        nodes = np.zeros((self._cw_nodes, 96), np.uint8)
        slot_prim = np.zeros(self.n_prims, np.uint32)
        lib.orc_cw_export(self._h, _p(nodes), _p(slot_prim))
        return nodes, slot_prim

    def cw_closest_hit(self, rays, threads=0):
        """(hits, wide nodes fetched, primitives tested) of the octant-ordered traversal of the compressed wide BVH."""
        rays = np.ascontiguousarray(rays, dtype=ray_dtype)
        out = np.zeros(len(rays), hit_dtype)
        counts = np.zeros(2, np.uint64)
        lib.orc_cw_closest_hit(self._h, _p(rays), len(rays), _p(out), threads, _p(counts))
        return out, int(counts[0]), int(counts[1])

    def lbvh_node_counts(self, rays, threads=0):
        """Nodes fetched per ray by the ordered LBVH traversal."""
        if not self._lbvh:
            self.lbvh_build()
        rays = np.ascontiguousarray(rays, dtype=ray_dtype)
        out = np.zeros(len(rays), np.uint32)
        lib.orc_lbvh_node_counts(self._h, _p(rays), len(rays), _p(out), threads)
        return out


# ---- free-standing KAT hooks -------------------------------------------------------------------------------------
def philox(ctr, key):
    c, k, o = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    lib.orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def sort_by_indices(values, indices):
    v = np.ascontiguousarray(values, np.uint32).copy()
    i = np.ascontiguousarray(indices, np.uint64)
    lib.orc_sort_by_indices_u32(_p(v), _p(i), len(v))
    return v


def next_float(f):
    return lib.orc_next_float(float(f))


def previous_float(f):
    return lib.orc_previous_float(float(f))


def gamma(n):
    return lib.orc_gamma(n)


def offset_ray(origin, normal, error, is_brdf):
    out = np.zeros(3, np.float32)
    lib.orc_offset_ray(_p(f3(origin)), _p(f3(normal)), _p(f3(error)), int(is_brdf), _p(out))
    return out


def ray_new(origin, direction):
    out = np.zeros(9, np.float32)
    lib.orc_ray_new(_p(f3(origin)), _p(f3(direction)), _p(out))
    return dict(direction=out[:3].copy(), d_inverse=out[3:6].copy(), shear=out[6:9].copy())


def coord_roundtrip(z, v):
    a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib.orc_coord_roundtrip(_p(f3(z)), _p(f3(v)), _p(a), _p(b))
    return a, b


def camera_make(origin, lookat, vup, fov, aspect, aperture, focus):
    out = np.zeros(12, np.float32)
    lib.orc_camera_make(_p(f3(origin)), _p(f3(lookat)), _p(f3(vup)), fov, aspect, aperture, focus, _p(out))
    return dict(origin=out[0:3].copy(), lower_left=out[3:6].copy(), horizontal=out[6:9].copy(), vertical=out[9:12].copy())


def lambertian_sample(normal, n, seed=0, local=False):
    d = np.zeros((n, 3), np.float32)
    lib.orc_lambertian_sample(_p(f3(normal)), seed, n, int(local), _p(d))
    return d


def lambertian_pdf(normal, dirs, local=False):
    dirs = np.ascontiguousarray(dirs, np.float32)
    out = np.zeros(len(dirs), np.float32)
    lib.orc_lambertian_pdf(_p(f3(normal)), _p(dirs), len(dirs), int(local), _p(out))
    return out


TR_H, TR_LOCAL, TR_WORLD = 0, 1, 2


def tr_sample(alpha, incoming, n, seed=0, which=TR_WORLD, normal=(0, 0, 1)):
    d = np.zeros((n, 3), np.float32)
    lib.orc_tr_sample(alpha, _p(f3(incoming)), _p(f3(normal)), seed, n, which, _p(d))
    return d


def tr_pdf(alpha, incoming, dirs, which=TR_WORLD, normal=(0, 0, 1)):
    dirs = np.ascontiguousarray(dirs, np.float32)
    out = np.zeros(len(dirs), np.float32)
    lib.orc_tr_pdf(alpha, _p(f3(incoming)), _p(f3(normal)), _p(dirs), len(dirs), which, _p(out))
    return out


def tr_integrand(alpha, a, dirs, which, normal=(0, 0, 1)):
    dirs = np.ascontiguousarray(dirs, np.float32)
    out = np.zeros(len(dirs), np.float32)
    lib.orc_tr_integrand(alpha, _p(f3(a)), _p(f3(normal)), _p(dirs), len(dirs), which, _p(out))
    return out


def random_unit_vectors(n, seed=0):
    d = np.zeros((n, 3), np.float32)
    lib.orc_random_unit_vectors(seed, n, _p(d))
    return d


def dist1d(values, nsamples=0, seed=0):
    v = np.ascontiguousarray(values, np.float32)
    pdf, cdf = np.zeros(len(v), np.float32), np.zeros(len(v) + 1, np.float32)
    counts = np.zeros(len(v), np.uint64)
    lib.orc_dist1d(_p(v), len(v), _p(pdf), _p(cdf), seed, nsamples, _p(counts) if nsamples else None)
    return pdf, cdf, counts


def dist2d(values, width, nsamples=0, seed=0):
    v = np.ascontiguousarray(values, np.float32).reshape(-1)
    pdf = np.zeros(len(v), np.float32)
    counts = np.zeros(len(v), np.uint64)
    lib.orc_dist2d(_p(v), len(v), width, _p(pdf), seed, nsamples, _p(counts) if nsamples else None)
    return pdf.reshape(-1, width), counts.reshape(-1, width)


def hardware_threads():
    return int(lib.orc_hardware_threads())
