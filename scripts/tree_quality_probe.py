#!/usr/bin/env python
"""CPU only: nodes fetched / primitives tested per ray by the ordered, t-culled walk on the C3 mesh for the Karras LBVH, the
PLOC hierarchy (oracle/ploc_ref.hpp) at several radii and the reference's SAH tree — camera rays, diffuse bounce rays
leaving the terrain, and rays inside the glass sphere. Usage: tree_quality_probe.py [scale] [n_rays]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ptb200  # noqa: E402  (meshgen only: the library is loaded lazily and never here)
import oracle as O  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
scene = ptb200.meshgen.c3_scene(scale)
orc = O.OracleScene(scene)
rng = np.random.default_rng(7)
# camera rays
u, v = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
cam = np.zeros(n, O.ray_dtype)
for i in range(n):
    cam["o"][i], cam["d"][i] = orc.camera_ray(float(u[i]), float(v[i]))
hits = orc.closest_hit(cam)
tri = scene.triangles
nt_terrain = int((tri["material"] == 0).sum())


def unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def bounce_rays(mask, inward):
    idx = np.nonzero(mask)[0]
    o = cam["o"][idx].astype(np.float64)
    d = cam["d"][idx].astype(np.float64)
    p = o + d * hits["t"][idx, None]
    P = tri["p"][hits["prim"][idx] - len(scene.spheres)]
    ng = unit(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]))
    flip = (np.sum(ng * d, -1) > 0) != inward
    ng = np.where(flip[:, None], -ng, ng)
    # cosine-weighted direction about ng
    r1, r2 = rng.random(len(idx)), rng.random(len(idx))
    phi = 2 * np.pi * r1
    loc = np.stack([np.cos(phi) * np.sqrt(r2), np.sin(phi) * np.sqrt(r2), np.sqrt(1 - r2)], -1)
    a = np.where(np.abs(ng[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    tx = unit(np.cross(ng, a))
    ty = np.cross(ng, tx)
    w = unit(loc[:, :1] * tx + loc[:, 1:2] * ty + loc[:, 2:3] * ng)
    out = np.zeros(len(idx), O.ray_dtype)
    q = p + 1e-4 * ng
    out["o"], out["d"] = q, w
    return out


hit = hits["prim"] != 0xFFFFFFFF
on_terrain = hit & (hits["prim"] - len(scene.spheres) < nt_terrain)
on_glass = hit & ~on_terrain
sets = {"camera": cam, "terrain bounce": bounce_rays(on_terrain, False), "inside glass": bounce_rays(on_glass, True)}
print({k: len(v) for k, v in sets.items()})


def report(label):
    line = f"{label:28s}"
    for k, r in sets.items():
        _, nv, pt = orc.lbvh_closest_hit(r)
        line += f" | {k}: V {nv / len(r):6.2f} T {pt / len(r):5.2f}"
    print(line, flush=True)


orc.lbvh_build()
report("Karras LBVH")
for radius in (int(x) for x in os.environ.get("RADII", "8,16,32").split(",") if x):
    t0 = time.time()
    rounds = orc.lbvh_ploc(radius)
    report(f"PLOC r={radius} ({rounds} rounds, {time.time() - t0:.1f}s)")
    orc.lbvh_build()
for nb in (int(x) for x in os.environ.get("SAH_BINS", "8,16").split(",") if x):
    orc.lbvh_build()
    t0 = time.time()
    st = orc.lbvh_sah(nb)
    report(f"SAH {nb} bins (levels {st[0]}, tasks <= {st[1]}, small {st[2]}, halvings {st[3]}, deepest leaf {st[4]}, depth-limited {st[5]}; {time.time() - t0:.1f}s)")
for extra in (0.0, 1.0, 2.0):
    if os.environ.get("SNAP", "1") == "0":
        break
    orc.lbvh_build()
    orc.lbvh_snap16(extra)
    report(f"LBVH, 16-bit boxes +{extra:g}")
orc.lbvh_build()
line = f"{'reference SAH (ordered walk)':28s}"
for k, r in sets.items():
    _, nv, pt = orc.sah_ordered_closest_hit(r)
    line += f" | {k}: V {nv / len(r):6.2f} T {pt / len(r):5.2f}"
print(line)
