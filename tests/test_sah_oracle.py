"""CPU: the definition of the device's SAH builder (oracle/sah_ref.hpp) is a valid tree over every primitive, finds the same
hits as the reference-semantics oracle and the LBVH, and needs fewer node fetches per ray than the LBVH (what it is for)."""
import numpy as np

from conftest import random_rays

LEAF = 0x80000000
MISS = 0xFFFFFFFF


def _check_tree(nodes, prims, n):
    assert np.array_equal(np.sort(prims), np.arange(n, dtype=np.uint32))
    refs = np.concatenate([nodes["left"], nodes["right"]])
    assert np.array_equal(np.sort(refs[refs >= LEAF] - LEAF), np.arange(n, dtype=np.uint32))
    assert np.array_equal(np.sort(refs[refs < LEAF]), np.arange(1, len(nodes), dtype=np.uint32))
    assert nodes["parent"][0] == MISS
    for side, mn, mx in (("left", "lmin", "lmax"), ("right", "rmin", "rmax")):
        ch = nodes[side]
        m = ch < LEAF
        assert np.array_equal(nodes["parent"][ch[m]], np.nonzero(m)[0].astype(np.uint32))
        c = nodes[ch[m]]
        assert np.array_equal(nodes[mn][m], np.minimum(c["lmin"], c["rmin"])) and np.array_equal(nodes[mx][m], np.maximum(c["lmax"], c["rmax"]))


def test_sah_definition_on_the_c3_mesh(ptb, orc):
    s = ptb.meshgen.c3_scene(0.1)
    rays = random_rays(ptb, 50_000, 31, centre=(0, 4, 1), radius=5.0)
    o = orc.OracleScene(s, split_type=-1)
    h0, v0, t0 = o.lbvh_closest_hit(rays)
    levels, max_tasks, small, fallbacks, max_depth, limited = o.lbvh_sah()
    assert levels > 5 and max_tasks > 10 and small > 100 and fallbacks == 0 and max_depth <= 60 and limited == 0
    _, prims, nodes = o.lbvh_export()
    _check_tree(nodes, prims, s.n_primitives)
    h1, v1, t1 = o.lbvh_closest_hit(rays)
    for f in ("prim", "t", "u", "v"):
        assert np.array_equal(h0[f].view(np.uint32), h1[f].view(np.uint32)), f
    assert v1 < 0.95 * v0 and t1 < 1.05 * t0, (v0, v1, t0, t1)
    r = orc.OracleScene(s).closest_hit(rays)          # the reference's own tree, test-all walk
    tie = (h1["prim"] != r["prim"]) & (h1["t"] == r["t"])
    assert tie.mean() < 1e-3 and np.array_equal(h1["prim"][~tie], r["prim"][~tie])


def test_sah_definition_small_and_degenerate(ptb, orc, rtweekend1, overshadowed):
    import copy
    for scene in (rtweekend1, overshadowed):
        o = orc.OracleScene(scene, split_type=-1)
        o.lbvh_sah()
        _, prims, nodes = o.lbvh_export()
        _check_tree(nodes, prims, scene.n_primitives)
    for n in (2, 3, 33, 700):                         # identical centroids: the halving fallback
        s = copy.deepcopy(rtweekend1)
        sp = np.zeros(n, ptb._lib.sphere_dtype)
        sp["radius"] = np.linspace(0.1, 2.0, n, dtype=np.float32)
        s.spheres = sp
        o = orc.OracleScene(s, split_type=-1)
        st = o.lbvh_sah()
        _, prims, nodes = o.lbvh_export()
        _check_tree(nodes, prims, n)
        assert st[3] > 0 or n <= 3


def test_sah_definition_depth_bound(ptb, orc):
    """The traversal stacks hold 64 entries: a split that would leave no room to finish by halving is replaced by the
    halving split. Lowering the bound (the test hook) exercises the rule; the tree stays valid and no leaf lies deeper."""
    s = ptb.meshgen.c3_scene(0.1)
    rays = random_rays(ptb, 20_000, 33, centre=(0, 4, 1), radius=5.0)
    o = orc.OracleScene(s, split_type=-1)
    h0, _, _ = o.lbvh_closest_hit(rays)
    for bound in (15, 14):
        st = o.lbvh_sah(8, bound)
        assert st[4] <= bound and st[5] > 0, st
        _, prims, nodes = o.lbvh_export()
        _check_tree(nodes, prims, s.n_primitives)
        h1, _, _ = o.lbvh_closest_hit(rays)
        assert np.array_equal(h0["prim"], h1["prim"]) and np.array_equal(h0["t"].view(np.uint32), h1["t"].view(np.uint32))
        o.lbvh_build()


import pytest


@pytest.mark.parametrize("seed", range(6))
def test_sah_definition_random_scenes(ptb, orc, rtweekend1, seed):
    """Random mixed scenes (spheres of very different sizes, slivers, duplicates, one flat axis): the tree is valid and its
    ordered walk finds the hits of the brute-force scan over every primitive."""
    import copy
    rng = np.random.default_rng(100 + seed)
    s = copy.deepcopy(rtweekend1)
    ns, nt = int(rng.integers(1, 200)), int(rng.integers(0, 400))
    sp = np.zeros(ns, ptb._lib.sphere_dtype)
    sp["center"] = rng.uniform(-3, 3, (ns, 3))
    sp["radius"] = 10.0 ** rng.uniform(-3, 0, ns)
    if seed % 2:
        sp["center"][:, 1] = 0.5                      # every centroid on one plane
        sp["center"][ns // 2:] = sp["center"][0]      # and half of them identical
    s.spheres = sp
    tri = np.zeros(nt, ptb._lib.triangle_dtype)
    if nt:
        base = rng.uniform(-3, 3, (nt, 1, 3))
        ext = 10.0 ** rng.uniform(-3, 0.3, (nt, 1, 1))
        tri["p"] = base + ext * rng.uniform(-1, 1, (nt, 3, 3)) * np.array([1.0, 1.0, 0.02 if seed % 3 == 0 else 1.0])
        n = np.cross(tri["p"][:, 1] - tri["p"][:, 0], tri["p"][:, 2] - tri["p"][:, 0])
        n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        tri["n"] = n[:, None, :]
    s.triangles = tri
    o = orc.OracleScene(s, split_type=-1)
    o.lbvh_sah()
    _, prims, nodes = o.lbvh_export()
    if s.n_primitives >= 2:
        _check_tree(nodes, prims, s.n_primitives)
    rays = random_rays(ptb, 20_000, 200 + seed, centre=(0, 0, 0), radius=4.0)
    h, _, _ = o.lbvh_closest_hit(rays)
    b = o.closest_hit_brute(rays)
    same = (h["prim"] == b["prim"]) | ((h["t"] == b["t"]) & (h["prim"] != MISS) & (b["prim"] != MISS))
    assert same.all(), np.nonzero(~same)[0][:5]
    m = h["prim"] != MISS
    assert np.array_equal(h["t"][m].view(np.uint32), b["t"][m].view(np.uint32))
