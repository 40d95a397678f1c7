"""K8 closest hit through the C ABI vs the oracle (reference semantics: SAH tree + BFS candidates + test-all).

Bar (BASELINE.json north_star): on non-degenerate rays the primitive id matches exactly and t agrees within
1e-5 relative. Because the device runs the reference's intersection arithmetic operation for operation, we also
require t to be BIT-identical wherever the primitive id matches.
"""
import numpy as np
import pytest

from conftest import random_rays

pytestmark = pytest.mark.gpu

T_RTOL = 1e-5


def degenerate(rays, hits_a, hits_b):
    """Rays excluded from the id-exactness requirement: y-dominant rays with a vanishing x
    direction component (the reference's axis permutation divides by dir.x), and exact-t ties."""
    d = rays["d"] / np.linalg.norm(rays["d"], axis=1, keepdims=True)
    ax, ay, az = np.abs(d[:, 0]), np.abs(d[:, 1]), np.abs(d[:, 2])
    ydom = ~((ax > ay) & (ax > az)) & (ay > az)
    q1 = ydom & (ax < 1e-3 * ay)
    tie = (hits_a["t"] == hits_b["t"]) & (hits_a["prim"] != hits_b["prim"])
    return q1 | tie


def check(ptb, orc, ctx, scene, rays, allow_frac=0.0):
    ctx.upload(scene)
    ctx.commit()
    g = ctx.closest_hit(rays)
    o = orc.OracleScene(scene)
    r = o.closest_hit(rays)
    ok = ~degenerate(rays, g, r)
    same = g["prim"][ok] == r["prim"][ok]
    bad = int((~same).sum())
    assert bad <= allow_frac * len(rays), f"{bad} primitive-id mismatches on {int(ok.sum())} non-degenerate rays"
    hit = ok & (g["prim"] == r["prim"]) & (g["prim"] != ptb.PTB_MISS)
    rel = np.abs(g["t"][hit] - r["t"][hit]) / np.maximum(np.abs(r["t"][hit]), 1e-30)
    assert rel.size == 0 or rel.max() <= T_RTOL
    assert np.array_equal(g["t"][hit].view(np.uint32), r["t"][hit].view(np.uint32)), "t not bit-identical"
    assert np.array_equal(g["u"][hit].view(np.uint32), r["u"][hit].view(np.uint32))
    assert np.array_equal(g["v"][hit].view(np.uint32), r["v"][hit].view(np.uint32))
    miss = ok & (g["prim"] == ptb.PTB_MISS)
    assert np.all(g["t"][miss] == 0.0)
    return g, r


def camera_rays(ptb, orc, scene, w, h):
    o = orc.OracleScene(scene, split_type=-1)
    rays = np.zeros(w * h, ptb.ray_dtype)
    k = 0
    for y in range(h):
        for x in range(w):
            org, d = o.camera_ray((x + 0.5) / (w - 1), 1 - (y + 0.5) / (h - 1))
            rays["o"][k], rays["d"][k] = org, d
            k += 1
    return rays


def test_rtweekend1_central_ray(ptb, gpu_ctx, rtweekend1):
    """dir (0,1,0) from the origin hits the small sphere (prim 1) at t = 0.5."""
    gpu_ctx.upload(rtweekend1)
    gpu_ctx.commit()
    h = gpu_ctx.closest_hit(ptb.make_rays(np.array([[0, 0, 0]], np.float32), np.array([[0, 1, 0]], np.float32)))
    assert h["prim"][0] == 1 and h["t"][0] == np.float32(0.5)


def test_rtweekend1(ptb, orc, gpu_ctx, rtweekend1):
    rays = np.concatenate([random_rays(ptb, 200_000, 11, centre=(0, 1, 0), radius=3.0),
                           camera_rays(ptb, orc, rtweekend1, 160, 90)])
    g, r = check(ptb, orc, gpu_ctx, rtweekend1, rays)
    assert (g["prim"] != ptb.PTB_MISS).mean() > 0.2


def test_overshadowed(ptb, orc, gpu_ctx, overshadowed):
    rays = np.concatenate([random_rays(ptb, 200_000, 12, centre=(-0.3, 0.3, -0.3), radius=1.5),
                           camera_rays(ptb, orc, overshadowed, 160, 90)])
    g, r = check(ptb, orc, gpu_ctx, overshadowed, rays)
    assert len(np.unique(g["prim"])) >= 6  # spheres, several cuboid faces, misses


@pytest.mark.parametrize("scale", [0.03, 0.1])
def test_c3_mesh(ptb, orc, gpu_ctx, scale):
    s = ptb.meshgen.c3_scene(scale)
    rays = np.concatenate([random_rays(ptb, 100_000, 13, centre=(0, 4, 1), radius=4.0),
                           camera_rays(ptb, orc, s, 128, 72)])
    check(ptb, orc, gpu_ctx, s, rays)


def test_brute_force_and_lbvh_oracle_agree(ptb, orc, gpu_ctx):
    """Three independent answers for the same rays: device, ordered CPU LBVH traversal, brute force over all prims."""
    s = ptb.meshgen.c3_scene(0.03)
    rays = random_rays(ptb, 50_000, 14, centre=(0, 4, 1), radius=4.0)
    gpu_ctx.upload(s)
    gpu_ctx.commit()
    g = gpu_ctx.closest_hit(rays)
    o = orc.OracleScene(s, split_type=-1)
    lb, nodes, prims = o.lbvh_closest_hit(rays)
    br = o.closest_hit_brute(rays)
    assert np.array_equal(g["prim"], lb["prim"]) and np.array_equal(g["t"].view(np.uint32), lb["t"].view(np.uint32))
    ok = ~degenerate(rays, g, br)
    assert np.array_equal(g["prim"][ok], br["prim"][ok])
    assert nodes > 0 and prims > 0


def test_empty_and_edge_inputs(ptb, gpu_ctx, rtweekend1):
    import copy
    gpu_ctx.upload(rtweekend1)
    gpu_ctx.commit()
    assert len(gpu_ctx.closest_hit(np.zeros(0, ptb.ray_dtype))) == 0       # empty batch
    s = copy.deepcopy(rtweekend1)                                           # empty scene: everything misses
    s.spheres = s.spheres[:0].copy()
    gpu_ctx.upload(s)
    gpu_ctx.commit()
    h = gpu_ctx.closest_hit(random_rays(ptb, 1000, 1))
    assert np.all(h["prim"] == ptb.PTB_MISS) and np.all(h["t"] == 0)
    # ragged batch sizes around the warp / block granularity
    gpu_ctx.upload(rtweekend1)
    gpu_ctx.commit()
    for n in (1, 31, 32, 33, 255, 257, 1025):
        assert len(gpu_ctx.closest_hit(random_rays(ptb, n, n))) == n


def test_host_batches_are_pipelined_without_changing_results(ptb, gpu_ctx, monkeypatch):
    """ptb_closest_hit splits a host ray array into batches over two staging pairs (upload | traverse | read back);
    results must not depend on the batch size, including a ragged last batch and a caller-owned output buffer."""
    s = ptb.meshgen.c3_scene(0.05)
    gpu_ctx.upload(s)
    gpu_ctx.commit()
    rays = random_rays(ptb, 50_001, 5, centre=(0, 4, 1), radius=5.0)
    whole = gpu_ctx.closest_hit(rays)
    monkeypatch.setenv("PTB_HIT_BATCH", "4096")
    out = np.zeros(len(rays), ptb.hit_dtype)
    parts = gpu_ctx.closest_hit(rays, out=out)
    assert parts is out and np.array_equal(whole, parts)
    assert len(gpu_ctx.closest_hit(rays[:0])) == 0


def test_full_size_c3_matches_lbvh_oracle(ptb, orc, gpu_ctx):
    """BASELINE config C3 at full size (1 000 000 triangles), 1 M incoherent rays: the device against the CPU traversal of the
    same (bit-exact) LBVH — identical primitive ids and bit-identical t — and the device's node count within the
    speculation margin of the sequential oracle."""
    s = ptb.meshgen.c3_scene(1.0)
    rays = random_rays(ptb, 1 << 20, 77, centre=(0, 4, 1), radius=6.0)
    gpu_ctx.upload(s)
    gpu_ctx.commit()
    g = gpu_ctx.closest_hit(rays)
    o = orc.OracleScene(s, split_type=-1)
    h, nodes, prims = o.lbvh_closest_hit(rays)
    assert np.array_equal(g["prim"], h["prim"])
    assert np.array_equal(g["t"].view(np.uint32), h["t"].view(np.uint32))
    assert 0.2 < float((g["prim"] != ptb.PTB_MISS).mean()) < 0.9


def test_full_size_heightfield_properties(ptb, gpu_ctx):
    """BASELINE config C5 geometry (10 000 000-triangle heightfield), 4 M rays of the C5 stream: size-independent properties —
    every hit lies on the surface z = h(x, y) (within the facet sag), t > 0, misses only where the ray leaves the slab
    [-1,1]^2 x [-0.2,0.2] untouched, and the answer does not depend on the order of the rays."""
    s = ptb.meshgen.heightfield_scene(2500, 2000)
    assert len(s.triangles) == 10_000_000
    gpu_ctx.upload(s)
    gpu_ctx.commit()
    n = 1 << 22
    rays = ptb.meshgen.philox_rays(n, first=0)
    hits = gpu_ctx.closest_hit(rays)
    hit = hits["prim"] != ptb.PTB_MISS
    assert 0.05 < hit.mean() < 0.95 and np.all(hits["t"][hit] > 0) and np.all(hits["t"][~hit] == 0)
    assert hits["prim"][hit].max() < 10_000_000
    d = rays["d"] / np.linalg.norm(rays["d"], axis=1, keepdims=True)
    p = rays["o"][hit].astype(np.float64) + d[hit].astype(np.float64) * hits["t"][hit, None]
    assert np.all(np.abs(p[:, 0]) <= 1 + 1e-5) and np.all(np.abs(p[:, 1]) <= 1 + 1e-5)
    X, Y = p[:, 0], p[:, 1]
    z = 0.1 * np.sin(9 * X) * np.cos(7 * Y) + 0.1 * np.sin(23 * X + 17 * Y)
    assert np.max(np.abs(p[:, 2] - z)) < 2e-4          # facets of 8e-4 x 1e-3 under curvature <= ~85: sag ~1e-5 .. 1e-4
    perm = np.random.default_rng(3).permutation(n)
    again = gpu_ctx.closest_hit(rays[perm])
    assert np.array_equal(again, hits[perm])


def test_batch_ordering_does_not_change_results(ptb, gpu_ctx, monkeypatch):
    """ptb_closest_hit orders batches of >= 65 536 rays (k_ray_sort_keys: reaches the scene box, Morton code of the entry
    point, octant) and traces them through an index array: hits[i] must still answer rays[i], bit for bit."""
    s = ptb.meshgen.c3_scene(0.1)
    gpu_ctx.upload(s)
    gpu_ctx.commit()
    rays = random_rays(ptb, 150_000, 31, centre=(0, 4, 1), radius=6.0)
    a = gpu_ctx.closest_hit(rays)
    monkeypatch.setenv("PTB_HIT_SORT", "0")
    b = gpu_ctx.closest_hit(rays)
    monkeypatch.delenv("PTB_HIT_SORT")
    assert 0.05 < np.mean(a["prim"] != ptb.PTB_MISS) < 0.95
    for f in ("t", "prim", "u", "v"):
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
