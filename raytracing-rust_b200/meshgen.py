"""Deterministic synthetic inputs for BASELINE.json's configs (no RNG in the geometry, closed-form heights).

  c3_scene()          1 000 000 triangles: 800 000-triangle Lambertian terrain + 200 000-triangle dielectric UV sphere,
                      Lerp sky as scenes/rtweekend1.ssml, 16:9 camera                         (SURVEY.md §8d "C3")
  heightfield_scene() R x C quads in [-1,1]^2 x [-0.2,0.2]; 2500 x 2000 -> 10 000 000 triangles      ("C5")
  philox_rays()       incoherent rays from Philox4x32-10 (seed 0x5EED, counter = ray index): origin uniform in the
                      ball of radius 2 about the mesh centre, direction uniform on the sphere
  write_obj()/c3_ssml()  the same C3 mesh as an OBJ (v / vn / usemtl / f v//vn) + .ssml, for the loader path
All arithmetic is float32 numpy; the arrays ARE the input (oracle and device receive identical bytes).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .scene import HostScene

F = np.float32


def _grid_triangles(xs: np.ndarray, ys: np.ndarray, height_fn, normal_fn):
    """(len(xs)-1) x (len(ys)-1) quads -> 2 triangles each; returns positions (T,3,3) and normals (T,3,3)."""
    X, Y = np.meshgrid(xs.astype(F), ys.astype(F), indexing="ij")
    Z = height_fn(X, Y).astype(F)
    P = np.stack([X, Y, Z], axis=-1)              # (nx, ny, 3)
    N = normal_fn(X, Y).astype(F)
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    n00, n10, n01, n11 = N[:-1, :-1], N[1:, :-1], N[:-1, 1:], N[1:, 1:]
    t0 = np.stack([p00, p10, p11], axis=2)         # (nx-1, ny-1, 3, 3)
    t1 = np.stack([p00, p11, p01], axis=2)
    m0 = np.stack([n00, n10, n11], axis=2)
    m1 = np.stack([n00, n11, n01], axis=2)
    pos = np.stack([t0, t1], axis=2).reshape(-1, 3, 3)
    nrm = np.stack([m0, m1], axis=2).reshape(-1, 3, 3)
    return np.ascontiguousarray(pos, F), np.ascontiguousarray(nrm, F)


def _terrain_height(X, Y):
    return F(0.25) * np.sin(F(1.3) * X) * np.cos(F(0.9) * Y) + F(0.1) * np.sin(F(3.1) * X + F(1.7) * Y)


def _terrain_normal(X, Y):
    dzdx = F(0.25) * F(1.3) * np.cos(F(1.3) * X) * np.cos(F(0.9) * Y) + F(0.1) * F(3.1) * np.cos(F(3.1) * X + F(1.7) * Y)
    dzdy = -F(0.25) * F(0.9) * np.sin(F(1.3) * X) * np.sin(F(0.9) * Y) + F(0.1) * F(1.7) * np.cos(F(3.1) * X + F(1.7) * Y)
    n = np.stack([-dzdx, -dzdy, np.ones_like(dzdx)], axis=-1)
    return n / np.linalg.norm(n, axis=-1, keepdims=True)


def uv_sphere(center, radius: float, slices: int, stacks: int):
    """2 * slices * (stacks - 1) triangles (poles are fans); normals are radial."""
    c = np.asarray(center, F)
    theta = (np.arange(stacks + 1, dtype=F) / F(stacks)) * F(np.pi)        # 0 .. pi
    phi = (np.arange(slices + 1, dtype=F) / F(slices)) * F(2 * np.pi)
    st, ct = np.sin(theta), np.cos(theta)
    sp, cp = np.sin(phi), np.cos(phi)
    sp[-1], cp[-1] = sp[0], cp[0]                                           # close the seam exactly
    st[0] = st[-1] = F(0)
    D = np.stack([st[:, None] * cp[None, :], st[:, None] * sp[None, :], np.repeat(ct[:, None], slices + 1, 1)], -1).astype(F)
    P = c + F(radius) * D
    tris_p, tris_n = [], []
    for k in range(stacks):
        a, b = k, k + 1
        pa0, pa1, pb0, pb1 = P[a, :-1], P[a, 1:], P[b, :-1], P[b, 1:]
        na0, na1, nb0, nb1 = D[a, :-1], D[a, 1:], D[b, :-1], D[b, 1:]
        if k != 0:                # upper triangle degenerates at the north pole
            tris_p.append(np.stack([pa0, pb1, pa1], 1)); tris_n.append(np.stack([na0, nb1, na1], 1))
        if k != stacks - 1:       # lower triangle degenerates at the south pole
            tris_p.append(np.stack([pa0, pb0, pb1], 1)); tris_n.append(np.stack([na0, nb0, nb1], 1))
    return np.ascontiguousarray(np.concatenate(tris_p), F), np.ascontiguousarray(np.concatenate(tris_n), F)


def _base_scene(sky_like_rtweekend: bool = True) -> HostScene:
    s = HostScene()
    sky_tex = s.add_texture(L.TEX_LERP, (0.5, 0.7, 1.0), (1.0, 1.0, 1.0))   # scenes/rtweekend1.ssml:10-14
    grey = s.add_texture(L.TEX_SOLID, (0.5, 0.5, 0.5))
    white = s.add_texture(L.TEX_SOLID, (1.0, 1.0, 1.0))
    s.add_material(L.MAT_LAMBERTIAN, grey, 0.5)      # 0: ground
    s.add_material(L.MAT_REFRACT, white, 1.5)        # 1: glass
    s.set_sky(sky_tex, (100, 100))
    return s


def c3_scene(scale: float = 1.0) -> HostScene:
    """scale=1 -> exactly 1 000 000 triangles; scale<1 shrinks every grid dimension (tests)."""
    nx, ny = max(2, int(round(1000 * scale))), max(2, int(round(400 * scale)))
    slices, stacks = max(3, int(round(500 * scale))), max(3, int(round(201 * scale)))
    s = _base_scene()
    xs = np.linspace(-12.5, 12.5, nx + 1, dtype=np.float64).astype(F)
    ys = np.linspace(0.0, 10.0, ny + 1, dtype=np.float64).astype(F)
    tp, tn = _grid_triangles(xs, ys, _terrain_height, _terrain_normal)
    sp_, sn = uv_sphere((0.0, 4.0, 1.0), 0.8, slices, stacks)
    tri = np.zeros(len(tp) + len(sp_), L.triangle_dtype)
    tri["p"][: len(tp)], tri["n"][: len(tp)], tri["material"][: len(tp)] = tp, tn, 0
    tri["p"][len(tp):], tri["n"][len(tp):], tri["material"][len(tp):] = sp_, sn, 1
    s.triangles = tri
    s.set_camera((0.0, -1.5, 1.6), (0.0, 4.0, 0.7), (0.0, 0.0, 1.0), 60.0, 16.0 / 9.0, 0.0, 1.0)
    return s


def heightfield_scene(rows: int = 2500, cols: int = 2000) -> HostScene:
    """rows x cols quads -> 2*rows*cols triangles in [-1,1]^2 x [-0.2,0.2] (C5: 10 000 000)."""
    s = _base_scene()
    xs = np.linspace(-1.0, 1.0, rows + 1, dtype=np.float64).astype(F)
    ys = np.linspace(-1.0, 1.0, cols + 1, dtype=np.float64).astype(F)

    def h(X, Y):
        return F(0.1) * np.sin(F(9.0) * X) * np.cos(F(7.0) * Y) + F(0.1) * np.sin(F(23.0) * X + F(17.0) * Y)

    def nrm(X, Y):
        dx = F(0.9) * np.cos(F(9.0) * X) * np.cos(F(7.0) * Y) + F(2.3) * np.cos(F(23.0) * X + F(17.0) * Y)
        dy = -F(0.7) * np.sin(F(9.0) * X) * np.sin(F(7.0) * Y) + F(1.7) * np.cos(F(23.0) * X + F(17.0) * Y)
        n = np.stack([-dx, -dy, np.ones_like(dx)], -1)
        return n / np.linalg.norm(n, axis=-1, keepdims=True)

    tp, tn = _grid_triangles(xs, ys, h, nrm)
    tri = np.zeros(len(tp), L.triangle_dtype)
    tri["p"], tri["n"], tri["material"] = tp, tn, 0
    s.triangles = tri
    s.set_camera((0.0, -3.0, 1.5), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 50.0, 16.0 / 9.0, 0.0, 1.0)
    return s


# ---- Philox4x32-10 in numpy (vectorised over counters) ---------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & np.uint64(0xFFFFFFFF) for c in (c0, c1, c2, c3))
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    k0, k1 = np.uint64(k0 & 0xFFFFFFFF), np.uint64(k1 & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _unit(u32):
    return (u32 >> np.uint32(8)).astype(F) * F(1.0 / 16777216.0)


def philox_rays(n: int, first: int = 0, seed: int = 0x5EED, centre=(0.0, 0.0, 0.0), radius: float = 2.0) -> np.ndarray:
    """Rays [first, first+n) of the C5 stream."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    lo, hi = idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32)
    zero = np.zeros_like(idx)
    a = philox4x32_10(lo, hi, zero, zero, seed, 0)
    b = philox4x32_10(lo, hi, zero, zero + np.uint64(1), seed, 0)

    def sphere(u, v):
        z = F(1) - F(2) * u
        r = np.sqrt(np.maximum(F(0), F(1) - z * z))
        ph = F(2 * np.pi) * v
        return np.stack([r * np.cos(ph), r * np.sin(ph), z], -1).astype(F)

    od = sphere(_unit(a[0]), _unit(a[1]))
    rad = F(radius) * np.cbrt(_unit(a[2]))
    rays = np.zeros(n, L.ray_dtype)
    rays["o"] = np.asarray(centre, F) + od * rad[:, None]
    rays["d"] = sphere(_unit(b[0]), _unit(b[1]))
    return rays


# ---- OBJ / .ssml export of C3 (exercises the loader path) --------------------------------------------------------
def write_obj(path: str, scene: HostScene, names=("ground", "glass")):
    tri = scene.triangles
    with open(path, "w") as f:
        f.write("# generated by raytracing-rust_b200.meshgen\no mesh\n")
        p = tri["p"].reshape(-1, 3)
        n = tri["n"].reshape(-1, 3)
        for v in p:
            f.write(f"v {float(v[0])!r} {float(v[1])!r} {float(v[2])!r}\n")
        for v in n:
            f.write(f"vn {float(v[0])!r} {float(v[1])!r} {float(v[2])!r}\n")
        cur = None
        for i, m in enumerate(tri["material"]):
            if m != cur:
                f.write(f"usemtl {names[int(m)]}\n")
                cur = m
            a = 3 * i + 1
            f.write(f"f {a}//{a} {a + 1}//{a + 1} {a + 2}//{a + 2}\n")


def c3_ssml(obj_path: str) -> str:
    return f"""camera (
	origin   0 -1.5 1.6
	lookat   0 4 0.7
	vup      0 0 1
	fov      60.0
	aperture 0.0
	focus_dis 1.0
)

texture sky (
	type lerp
	primary 0.5 0.7 1.0
	secondary 1.0
)

sky (
	texture sky
)

texture grey (
	type solid
	colour 0.5
)

texture white (
	type solid
	colour 1.0
)

material ground (
	type lambertian
	texture grey
	albedo 0.5
)

material glass (
	type refract
	texture white
	eta 1.5
)

mesh (
	type mesh
	obj {obj_path}
)
"""
