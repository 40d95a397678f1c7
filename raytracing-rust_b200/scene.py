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


@dataclass
class HostScene:
    spheres: np.ndarray = field(default_factory=lambda: np.zeros(0, L.sphere_dtype))
    triangles: np.ndarray = field(default_factory=lambda: np.zeros(0, L.triangle_dtype))
    materials: np.ndarray = field(default_factory=lambda: np.zeros(0, L.material_dtype))
    textures: np.ndarray = field(default_factory=lambda: np.zeros(0, L.texture_dtype))
    camera: np.ndarray = field(default_factory=lambda: np.zeros(1, L.camera_dtype))
    sky: np.ndarray = field(default_factory=lambda: np.zeros(1, L.sky_dtype))

    @property
    def n_primitives(self) -> int:
        return len(self.spheres) + len(self.triangles)

    def nbytes(self) -> int:
        """Bytes a ptb_scene_upload + commit moves host -> device."""
        return int(self.spheres.nbytes + self.triangles.nbytes + self.materials.nbytes + self.textures.nbytes
                   + self.camera.nbytes + self.sky.nbytes)

    # -- builders used by tests / generators ------------------------------------------------------------
    def add_texture(self, kind: int, a=(0, 0, 0), b=(0, 0, 0)) -> int:
        t = np.zeros(1, L.texture_dtype)
        t["kind"], t["a"], t["b"] = kind, a, b
        self.textures = np.concatenate([self.textures, t])
        return len(self.textures) - 1

    def add_material(self, kind: int, texture: int, param: float) -> int:
        m = np.zeros(1, L.material_dtype)
        m["kind"], m["texture"], m["param"], m["ior"] = kind, texture, param, (1, 1, 1)
        self.materials = np.concatenate([self.materials, m])
        return len(self.materials) - 1

    def add_sphere(self, center, radius: float, material: int) -> int:
        s = np.zeros(1, L.sphere_dtype)
        s["center"], s["radius"], s["material"] = center, radius, material
        self.spheres = np.concatenate([self.spheres, s])
        return len(self.spheres) - 1

    def set_camera(self, origin, lookat, vup, hfov_deg, aspect=16.0 / 9.0, aperture=0.0, focus_dist=10.0):
        """SimpleCamera::new (implementations/src/camera.rs:20-53)."""
        cam = np.zeros(1, L.camera_dtype)
        rc = L.lib.ptb_camera_make(L.Vec3(*map(float, origin)), L.Vec3(*map(float, lookat)), L.Vec3(*map(float, vup)),
                                   float(hfov_deg), float(aspect), float(aperture), float(focus_dist), L.ptr(cam))
        if rc != L.PTB_OK:
            raise L.PtbError(rc, "ptb_camera_make")
        self.camera = cam

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


def save_image(filename: str, width: int, height: int, rgb: np.ndarray, gamma: float = 2.2):
    """output::save_data_to_image (crates/output/src/lib.rs:74-113)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32).reshape(-1)
    assert rgb.size == width * height * 3
    rc = L.lib.ptb_image_save(filename.encode(), width, height, L.ptr(rgb), float(gamma))
    if rc != L.PTB_OK:
        raise L.PtbError(rc, f"ptb_image_save({filename})")
