"""The `--backend cuda` path as seen from Python: a thin, typed wrapper over the C ABI.

Names follow the reference's interface for this path:
  Bvh            <- implementations::acceleration::Bvh            (acceleration/mod.rs:43-93, check_hit :265-298)
  RenderOptions  <- implementations::samplers::RenderOptions       (samplers/mod.rs:22-41)
  RandomSampler  <- implementations::samplers::random_sampler::RandomSampler.sample_image (random_sampler.rs:10-99)
  Scene.render   <- frontend Scene::render                         (src/scene.rs:35-42)
Every compute call goes through libptb200.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .scene import HostScene


def _check(ctx_handle, rc: int):
    if rc != L.PTB_OK:
        msg = L.lib.ptb_last_error(ctx_handle)
        raise L.PtbError(rc, msg.decode() if msg else "")


@dataclass
class RenderOptions:
    """samplers/mod.rs:22-41 (defaults identical) + integrator constants + what sharding needs."""
    samples_per_pixel: int = 128
    render_method: int = L.METHOD_MIS
    width: int = 1920
    height: int = 1080
    gamma: float = 2.2
    sample_offset: int = 0
    max_depth: int = 50          # integrators/mod.rs:7
    rr_threshold: int = 3        # integrators/mod.rs:8
    seed: int = 0
    row_begin: int = 0           # image tile (multi-GPU second axis): rows [row_begin, row_begin + row_count); 0 = to the end
    row_count: int = 0

    def to_c(self) -> L.RenderOpts:
        return L.RenderOpts(self.width, self.height, self.samples_per_pixel, self.sample_offset, self.render_method,
                            self.max_depth, self.rr_threshold, 0, self.seed, self.row_begin, self.row_count)


class Context:
    """One ptb_ctx == one GPU."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = L.lib.ptb_create(device, C.byref(self._h))
        if rc != L.PTB_OK:
            msg = L.lib.ptb_last_error(None)
            raise L.PtbError(rc, msg.decode() if msg else "ptb_create failed")
        self.device = device
        self._progress_ref = None

    def close(self):
        if self._h:
            L.lib.ptb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- scene
    def set_stream(self, cuda_stream_ptr: int):
        _check(self._h, L.lib.ptb_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, option: int, value: int):
        _check(self._h, L.lib.ptb_set_option(self._h, option, value))

    def synchronize(self):
        _check(self._h, L.lib.ptb_synchronize(self._h))

    def upload(self, s: HostScene):
        """ptb_scene_set_* for every array (host -> library copy)."""
        h = self._h
        _check(h, L.lib.ptb_scene_set_textures(h, L.ptr(s.textures), len(s.textures)))
        for i, (w, hh, words) in s.texture_data.items():
            _check(h, L.lib.ptb_scene_set_texture_data(h, i, w, hh, L.ptr(words), words.size))
        _check(h, L.lib.ptb_scene_set_materials(h, L.ptr(s.materials), len(s.materials)))
        _check(h, L.lib.ptb_scene_set_spheres(h, L.ptr(s.spheres), len(s.spheres)))
        _check(h, L.lib.ptb_scene_set_triangles(h, L.ptr(s.triangles), len(s.triangles)))
        _check(h, L.lib.ptb_scene_set_camera(h, L.ptr(s.camera)))
        _check(h, L.lib.ptb_scene_set_sky(h, L.ptr(s.sky)))

    def commit(self, flags: int = 0):
        """Bvh::new: device build of the acceleration structure. L.BUILD_BINARY: Karras LBVH; L.BUILD_SAH: top-down SAH
        builder (sah_build.cu); L.BUILD_WIDE: the LBVH collapsed into the compressed 8-wide tree; L.BUILD_DEFAULT follows
        PTB_BVH=binary|lbvh|sah|wide, else the library default."""
        _check(self._h, L.lib.ptb_scene_commit(self._h, flags))

    def bvh_info(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self._h, L.lib.ptb_bvh_info(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bvh_export(self):
        n, m = self.bvh_info()
        morton = np.zeros(n, np.uint32)
        prims = np.zeros(n, np.uint32)
        nodes = np.zeros(m, L.bvh_node_dtype)
        _check(self._h, L.lib.ptb_bvh_export(self._h, L.ptr(morton), L.ptr(prims), L.ptr(nodes)))
        return morton, prims, nodes

    def bvh_export_quantised(self):
        """(frame: grid origin xyz + grid step xyz, nodes (m, 8) uint32: the 32-byte nodes the traversal kernels read)."""
        _, m = self.bvh_info()
        frame = np.zeros(6, np.float32)
        nodes = np.zeros((m, 8), np.uint32)
        _check(self._h, L.lib.ptb_bvh_export_quantised(self._h, L.ptr(frame), L.ptr(nodes)))
        return frame, nodes

    def bvh_builder(self):
        """(L.BUILD_BINARY | L.BUILD_SAH | L.BUILD_WIDE: which builder made the committed tree, levels the SAH build took)."""
        a, b = C.c_uint32(), C.c_uint32()
        _check(self._h, L.lib.ptb_bvh_builder(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bvh_wide_info(self):
        """(number of 96-byte nodes of the compressed 8-wide tree — 0 when the scene uses the binary tree, max leaf size)."""
        n, m = C.c_uint64(), C.c_uint32()
        _check(self._h, L.lib.ptb_bvh_wide_info(self._h, C.byref(n), C.byref(m)))
        return n.value, m.value

    def bvh_wide_export(self):
        """(nodes (n, 96) uint8, slot_prim: the wide tree's primitive order -> original primitive id)."""
        n, _ = self.bvh_wide_info()
        n_prims, _ = self.bvh_info()
        nodes = np.zeros((n, 96), np.uint8)
        slot_prim = np.zeros(n_prims, np.uint32)
        _check(self._h, L.lib.ptb_bvh_wide_export(self._h, L.ptr(nodes), L.ptr(slot_prim)))
        return nodes, slot_prim

    # ---- closest hit
    def closest_hit(self, rays: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """AccelerationStructure::check_hit for a batch (host buffers in, host buffers out). `out`: optional caller-owned
        hit buffer (e.g. pinned memory: the library overlaps upload, traversal and read-back of 4 Mi-ray batches)."""
        rays = np.ascontiguousarray(rays, dtype=L.ray_dtype)
        hits = np.zeros(len(rays), L.hit_dtype) if out is None else out
        assert hits.dtype == L.hit_dtype and len(hits) == len(rays) and hits.flags.c_contiguous
        _check(self._h, L.lib.ptb_closest_hit(self._h, L.ptr(rays), len(rays), L.ptr(hits)))
        return hits

    def closest_hit_device(self, d_rays_ptr: int, n: int, d_hits_ptr: int):
        _check(self._h, L.lib.ptb_closest_hit_device(self._h, C.c_void_p(d_rays_ptr), n, C.c_void_p(d_hits_ptr)))

    # ---- render
    def render(self, opts: RenderOptions, progress=None):
        cb = None
        if progress is not None:
            cb = L.PROGRESS_FN(lambda user, samples, rays: 1 if progress(samples, rays) else 0)
        self._progress_ref = cb
        o = opts.to_c()
        rc = L.lib.ptb_render(self._h, C.byref(o), C.cast(cb, C.c_void_p) if cb else None, None)
        self._progress_ref = None
        _check(self._h, rc)

    def render_passes(self, opts: RenderOptions, update):
        """The reference's per-pass contract (random_sampler.rs:31-98): `update(pass_image (H, W, 3), pass_number,
        rays_shot) -> bool` once per pass with that pass's single-sample image; True stops the render."""
        h, w = opts.height, opts.width

        def thunk(user, img, n, i, rays):
            a = np.ctypeslib.as_array(img, shape=(n,)).reshape(h, w, 3)
            return 1 if update(a, int(i), int(rays)) else 0

        cb = L.PASS_FN(thunk)
        o = opts.to_c()
        _check(self._h, L.lib.ptb_render_passes(self._h, C.byref(o), C.cast(cb, C.c_void_p), None))

    def accum_clear(self):
        _check(self._h, L.lib.ptb_accum_clear(self._h))

    def accum_read(self, width: int, height: int, normalise: bool = True, out: np.ndarray | None = None) -> np.ndarray:
        """`out`: optional caller-owned float32 buffer of W*H*3 elements (e.g. pinned memory) to read into."""
        if out is None:
            out = np.empty(width * height * 3, np.float32)
        out = out.reshape(-1)
        _check(self._h, L.lib.ptb_accum_read(self._h, L.ptr(out), out.size, 1 if normalise else 0))
        return out.reshape(height, width, 3)

    def accum_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(self._h, L.lib.ptb_accum_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def accum_set_samples(self, n: int):
        _check(self._h, L.lib.ptb_accum_set_samples(self._h, n))

    # ---- sampler test hook (the reference's chi-squared harness pointed at the device samplers)
    @staticmethod
    def _query(kind, alpha, normal, aux, light_index, seed) -> L.SamplerQuery:
        return L.SamplerQuery(kind, float(alpha), L.Vec3(*map(float, normal)), L.Vec3(*map(float, aux)), int(light_index), int(seed))

    def sample_only(self, kind: int, n: int, alpha=0.0, normal=(0, 0, 1), aux=(0, 0, 1), light_index=0, seed=0):
        """n directions from device sampler `kind` (L.SAMPLER_*) and the sampler's pdf of each: (dirs (n, 3), pdf (n,))."""
        q = self._query(kind, alpha, normal, aux, light_index, seed)
        dirs, pdf = np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        _check(self._h, L.lib.ptb_sample_only(self._h, C.byref(q), n, L.ptr(dirs), L.ptr(pdf)))
        return dirs, pdf

    def sampler_pdf(self, kind: int, dirs: np.ndarray, alpha=0.0, normal=(0, 0, 1), aux=(0, 0, 1), light_index=0):
        q = self._query(kind, alpha, normal, aux, light_index, 0)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        pdf = np.zeros(len(dirs), np.float32)
        _check(self._h, L.lib.ptb_sampler_pdf(self._h, C.byref(q), L.ptr(dirs), len(dirs), L.ptr(pdf)))
        return pdf

    def stats(self) -> L.Stats:
        s = L.Stats()
        _check(self._h, L.lib.ptb_stats_get(self._h, C.byref(s)))
        return s

    def stats_reset(self):
        _check(self._h, L.lib.ptb_stats_reset(self._h))


class Bvh:
    """Device acceleration structure over a host scene (Bvh::new + check_hit)."""

    def __init__(self, ctx: Context, scene: HostScene):
        self.ctx = ctx
        ctx.upload(scene)
        ctx.commit()
        self.n_prims, self.n_nodes = ctx.bvh_info()

    def number_nodes(self) -> int:  # acceleration/mod.rs:94-96
        return self.n_nodes

    def check_hit(self, rays: np.ndarray) -> np.ndarray:
        return self.ctx.closest_hit(rays)


class RandomSampler:
    """RandomSampler::sample_image on the device: `spp` samples of every pixel into the accumulator."""

    def sample_image(self, opts: RenderOptions, bvh: Bvh, update=None, presentation_update=None):
        """`presentation_update(pass_image, i, rays_shot) -> bool`: the reference's per-pass closure; `update(samples,
        rays) -> bool`: counters only (one device accumulator, no per-pass read-back)."""
        if presentation_update is not None:
            bvh.ctx.render_passes(opts, presentation_update)
        else:
            bvh.ctx.render(opts, progress=update)


class Scene:
    """frontend Scene (src/scene.rs:7-43) for `--backend cuda`."""

    def __init__(self, host_scene: HostScene, device: int = 0, ctx: Context | None = None):
        self.ctx = ctx or Context(device)
        self.host = host_scene
        self.acceleration = Bvh(self.ctx, host_scene)

    def render(self, opts: RenderOptions, update=None, presentation_update=None) -> np.ndarray:
        """Returns the running-mean image (H, W, 3) the TUI closure of src/main.rs:175-191 would hold."""
        self.ctx.accum_clear()
        RandomSampler().sample_image(opts, self.acceleration, update, presentation_update)
        return self.ctx.accum_read(opts.width, opts.height, normalise=True)


def render_multi(contexts, opts: RenderOptions) -> np.ndarray:
    """Scene::render across several GPUs of one box through ptb_render_multi: every context (one per GPU, same scene
    committed on each) renders its share of the samples on its own host thread inside the library, then one
    ncclReduce(sum) to contexts[0]. Returns the mean image."""
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    o = opts.to_c()
    _check(contexts[0]._h, L.lib.ptb_render_multi(C.cast(arr, C.c_void_p), len(contexts), C.byref(o)))
    return contexts[0].accum_read(opts.width, opts.height, normalise=True)


def make_rays(origins: np.ndarray, directions: np.ndarray) -> np.ndarray:
    r = np.zeros(len(origins), L.ray_dtype)
    r["o"] = origins
    r["d"] = directions
    return r
