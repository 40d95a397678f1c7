"""Host-side scene description (POD arrays) and the .ssml loader binding.

Mirrors what `loader::load_file_full` hands to `Bvh::new` + `Scene::new` in the reference
(crates/loader/src/lib.rs:196-243, src/parameters.rs:45-78): primitives (spheres first, then mesh
triangles), materials, textures, camera, sky. Parsing itself is native (host/ssml_loader.cpp).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L


def _camera_make_native(origin, lookat, vup, hfov_deg, aspect, aperture, focus_dist) -> np.ndarray:
    cam = np.zeros(1, L.camera_dtype)
    rc = L.lib.ptb_camera_make(L.Vec3(*map(float, origin)), L.Vec3(*map(float, lookat)), L.Vec3(*map(float, vup)),
                               float(hfov_deg), float(aspect), float(aperture), float(focus_dist), L.ptr(cam))
    if rc != L.PTB_OK:
        raise L.PtbError(rc, "ptb_camera_make")
    return cam


# SimpleCamera::new. bench.py's `--impl reference` arm swaps in the CPU checker's restatement (bit-equal,
# tests/ pin that) so that its process never maps libptb200.so; nothing in this package does.
CAMERA_MAKE = _camera_make_native


@dataclass
class HostScene:
    spheres: np.ndarray = field(default_factory=lambda: np.zeros(0, L.sphere_dtype))
    triangles: np.ndarray = field(default_factory=lambda: np.zeros(0, L.triangle_dtype))
    materials: np.ndarray = field(default_factory=lambda: np.zeros(0, L.material_dtype))
    textures: np.ndarray = field(default_factory=lambda: np.zeros(0, L.texture_dtype))
    camera: np.ndarray = field(default_factory=lambda: np.zeros(1, L.camera_dtype))
    sky: np.ndarray = field(default_factory=lambda: np.zeros(1, L.sky_dtype))
    # bulk data of ImageTexture / Perlin textures: texture index -> (width, height, float32 words)
    texture_data: dict = field(default_factory=dict)

    @property
    def n_primitives(self) -> int:
        return len(self.spheres) + len(self.triangles)

    def nbytes(self) -> int:
        """Bytes a ptb_scene_upload + commit moves host -> device."""
        return int(self.spheres.nbytes + self.triangles.nbytes + self.materials.nbytes + self.textures.nbytes
                   + self.camera.nbytes + self.sky.nbytes + sum(d.nbytes for _, _, d in self.texture_data.values()))

    # -- builders used by tests / generators ------------------------------------------------------------
    def add_texture(self, kind: int, a=(0, 0, 0), b=(0, 0, 0)) -> int:
        t = np.zeros(1, L.texture_dtype)
        t["kind"], t["a"], t["b"] = kind, a, b
        self.textures = np.concatenate([self.textures, t])
        return len(self.textures) - 1

    def add_image_texture(self, rgb: np.ndarray) -> int:
        """ImageTexture (textures/mod.rs:202-266) from an (H, W, 3) float array."""
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        h, w, _ = rgb.shape
        i = self.add_texture(L.TEX_IMAGE)
        self.texture_data[i] = (w, h, rgb.reshape(-1).copy())
        return i

    def add_perlin_texture(self, seed: int = 0) -> int:
        """Perlin (textures/mod.rs:75-180); tables are a pure function of `seed` (the reference uses OS entropy)."""
        i = self.add_texture(L.TEX_PERLIN)
        words = np.zeros(L.PERLIN_TABLE_WORDS, np.float32)
        rc = L.lib.ptb_perlin_tables(seed, L.ptr(words))
        if rc != L.PTB_OK:
            raise L.PtbError(rc, "ptb_perlin_tables")
        self.texture_data[i] = (0, 0, words)
        return i

    def add_material(self, kind: int, texture: int, param: float, ior=(1, 1, 1), metallic: float = 0.0) -> int:
        m = np.zeros(1, L.material_dtype)
        m["kind"], m["texture"], m["param"], m["ior"], m["metallic"] = kind, texture, param, ior, metallic
        self.materials = np.concatenate([self.materials, m])
        return len(self.materials) - 1

    def add_sphere(self, center, radius: float, material: int) -> int:
        s = np.zeros(1, L.sphere_dtype)
        s["center"], s["radius"], s["material"] = center, radius, material
        self.spheres = np.concatenate([self.spheres, s])
        return len(self.spheres) - 1

    def set_camera(self, origin, lookat, vup, hfov_deg, aspect=16.0 / 9.0, aperture=0.0, focus_dist=10.0):
        """SimpleCamera::new (implementations/src/camera.rs:20-53)."""
        self.camera = CAMERA_MAKE(origin, lookat, vup, hfov_deg, aspect, aperture, focus_dist)

    def set_sky(self, texture: int, sampler_res=(100, 100)):
        self.sky = np.zeros(1, L.sky_dtype)
        self.sky["texture"], self.sky["sampler_res_x"], self.sky["sampler_res_y"] = texture, sampler_res[0], sampler_res[1]


def _copy_array(handle, getter, dtype) -> np.ndarray:
    p = C.c_void_p()
    n = getter(handle, C.byref(p))
    if n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def _from_handle(handle) -> HostScene:
    s = HostScene()
    s.spheres = _copy_array(handle, L.lib.ptb_host_scene_spheres, L.sphere_dtype)
    s.triangles = _copy_array(handle, L.lib.ptb_host_scene_triangles, L.triangle_dtype)
    s.materials = _copy_array(handle, L.lib.ptb_host_scene_materials, L.material_dtype)
    s.textures = _copy_array(handle, L.lib.ptb_host_scene_textures, L.texture_dtype)
    for i in range(len(s.textures)):
        w, h, p = C.c_uint32(), C.c_uint32(), C.c_void_p()
        n = L.lib.ptb_host_scene_texture_data(handle, i, C.byref(w), C.byref(h), C.byref(p))
        if n:
            buf = (C.c_float * n).from_address(p.value)
            s.texture_data[i] = (w.value, h.value, np.frombuffer(buf, dtype=np.float32, count=n).copy())
    L.lib.ptb_host_scene_camera(handle, L.ptr(s.camera))
    L.lib.ptb_host_scene_sky(handle, L.ptr(s.sky))
    return s


def load_file(path: str) -> HostScene:
    """loader::load_file_full (crates/loader/src/lib.rs:196-243). Raises PtbError like LoadErr."""
    h = C.c_void_p()
    rc = L.lib.ptb_ssml_load_file(path.encode(), C.byref(h))
    if rc != L.PTB_OK:
        raise L.PtbError(rc, L.lib.ptb_host_last_error().decode())
    try:
        return _from_handle(h)
    finally:
        L.lib.ptb_host_scene_free(h)


def load_str(text: str, base_dir: str = ".") -> HostScene:
    """loader::load_str_full (crates/loader/src/lib.rs:245-288)."""
    h = C.c_void_p()
    rc = L.lib.ptb_ssml_load_str(text.encode(), base_dir.encode(), C.byref(h))
    if rc != L.PTB_OK:
        raise L.PtbError(rc, L.lib.ptb_host_last_error().decode())
    try:
        return _from_handle(h)
    finally:
        L.lib.ptb_host_scene_free(h)


def load_image(filename: str) -> np.ndarray:
    """ImageTexture::new's decode (textures/mod.rs:208-245): (H, W, 3) float32."""
    w, h, p = C.c_uint32(), C.c_uint32(), C.c_void_p()
    rc = L.lib.ptb_image_load(filename.encode(), C.byref(w), C.byref(h), C.byref(p))
    if rc != L.PTB_OK:
        raise L.PtbError(rc, L.lib.ptb_image_last_error().decode())
    try:
        buf = (C.c_float * (w.value * h.value * 3)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float32).reshape(h.value, w.value, 3).copy()
    finally:
        L.lib.ptb_image_free(p)


def save_image(filename: str, width: int, height: int, rgb: np.ndarray, gamma: float = 2.2):
    """output::save_data_to_image (crates/output/src/lib.rs:74-113)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32).reshape(-1)
    assert rgb.size == width * height * 3
    rc = L.lib.ptb_image_save(filename.encode(), width, height, L.ptr(rgb), float(gamma))
    if rc != L.PTB_OK:
        raise L.PtbError(rc, f"ptb_image_save({filename})")
