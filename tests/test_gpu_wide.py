"""The compressed 8-wide BVH (north_star "optional SAH collapse to a wide BVH") through the C ABI.

  * build: the device tree (cwbvh_build.cu) is BIT-EXACT against its CPU definition (oracle/cwbvh_ref.hpp): every byte of
    every 96-byte node, and the primitive order;
  * traversal: identical hits (primitive id, t, u, v bit for bit) to the binary LBVH traversal and to the reference-
    semantics SAH oracle, and EQUAL node / primitive counts to the CPU statement of the same traversal;
  * render: the image does not depend on which tree the kernels walk.
The whole `-m gpu` suite is also run with PTB_BVH=wide in the environment (scripts/gpu_suite.sh).
"""
import numpy as np
import pytest

from conftest import random_rays

pytestmark = pytest.mark.gpu

MISS = 0xFFFFFFFF


@pytest.fixture()
def wide_ctx(ptb):
    c = ptb.Context(0)
    yield c
    c.close()


def commit_wide(ptb, ctx, scene):
    ctx.upload(scene)
    ctx.commit(ptb._lib.BUILD_WIDE)


@pytest.mark.parametrize("max_leaf", [1, 3])
@pytest.mark.parametrize("which", ["rtweekend1", "overshadowed", "c3_small", "c3_mid", "one_sphere", "three_tris"])
def test_wide_tree_is_bit_exact(ptb, orc, wide_ctx, rtweekend1, overshadowed, which, max_leaf, monkeypatch):
    import copy
    if which == "one_sphere":
        scene = copy.deepcopy(rtweekend1)
        scene.spheres = scene.spheres[:1].copy()
    elif which == "three_tris":
        scene = ptb.meshgen.c3_scene(0.1)
        scene.triangles = scene.triangles[:3].copy()
    else:
        scene = {"rtweekend1": rtweekend1, "overshadowed": overshadowed}.get(which) or ptb.meshgen.c3_scene(0.05 if which == "c3_small" else 0.3)
    monkeypatch.setenv("PTB_WIDE_LEAF", str(max_leaf))
    commit_wide(ptb, wide_ctx, scene)
    monkeypatch.delenv("PTB_WIDE_LEAF")
    n, leaf = wide_ctx.bvh_wide_info()
    assert leaf == max_leaf
    nodes, slot_prim = wide_ctx.bvh_wide_export()
    o = orc.OracleScene(scene, split_type=-1)
    assert o.cw_build(max_leaf) == n
    want_nodes, want_prims = o.cw_export()
    assert np.array_equal(slot_prim, want_prims)
    assert sorted(slot_prim.tolist()) == list(range(scene.n_primitives))       # a permutation of the primitives
    assert np.array_equal(nodes, want_nodes), np.argwhere(nodes != want_nodes)[:5]
    # the LBVH underneath is still there and still bit-exact (ptb_bvh_export)
    morton, prims, bnodes = wide_ctx.bvh_export()
    wm, wp, wn = o.lbvh_export()
    assert np.array_equal(morton, wm) and np.array_equal(prims, wp)
    if len(wn):
        for f in ("left", "right", "parent", "lmin", "lmax", "rmin", "rmax"):
            assert np.array_equal(bnodes[f].view(np.uint32), wn[f].view(np.uint32)), f


@pytest.mark.parametrize("which", ["rtweekend1", "overshadowed", "c3"])
def test_wide_traversal_matches_binary_and_oracles(ptb, orc, gpu_ctx, wide_ctx, rtweekend1, overshadowed, which):
    scene, centre, radius = {"rtweekend1": (rtweekend1, (0, 1, 0), 3.0), "overshadowed": (overshadowed, (-0.3, 0.3, -0.3), 1.5),
                             "c3": (ptb.meshgen.c3_scene(0.3), (0, 4, 1), 5.0)}[which]
    rays = random_rays(ptb, 300_000, 51, centre=centre, radius=radius)
    gpu_ctx.upload(scene)
    gpu_ctx.commit(ptb._lib.BUILD_BINARY)
    b = gpu_ctx.closest_hit(rays)
    commit_wide(ptb, wide_ctx, scene)
    wide_ctx.set_option(ptb._lib.OPT_COUNT_TRAVERSAL, 1)
    wide_ctx.stats_reset()
    w = wide_ctx.closest_hit(rays)
    st = wide_ctx.stats()
    wide_ctx.set_option(ptb._lib.OPT_COUNT_TRAVERSAL, 0)
    for f in ("prim", "t", "u", "v"):
        assert np.array_equal(w[f].view(np.uint32), b[f].view(np.uint32)), f
    o = orc.OracleScene(scene)                      # SAH tree: reference semantics
    r = o.closest_hit(rays)
    tie = (w["prim"] != r["prim"]) & (w["t"] == r["t"])
    assert tie.mean() < 1e-3 and np.array_equal(w["prim"][~tie], r["prim"][~tie])
    assert np.array_equal(w["t"][~tie].view(np.uint32), r["t"][~tie].view(np.uint32))
    o.cw_build(3)
    h, nodes, prims = o.cw_closest_hit(rays)
    assert np.array_equal(h["prim"], w["prim"])
    assert st.rays_counted == len(rays)
    assert (st.nodes_fetched, st.prims_tested) == (nodes, prims)   # the device walks exactly the CPU statement's steps
    assert 0.05 < float((w["prim"] != MISS).mean()) < 0.95


def test_wide_full_size_c3_and_any_hit(ptb, orc, wide_ctx, overshadowed):
    """Full-size C3 closest hit through the wide tree vs the LBVH oracle; and the any-hit path (NEE shadow rays towards the
    light sphere of overshadowed) through a strict-MIS render against the oracle's counters."""
    s = ptb.meshgen.c3_scene(1.0)
    rays = random_rays(ptb, 1 << 19, 77, centre=(0, 4, 1), radius=6.0)
    commit_wide(ptb, wide_ctx, s)
    g = wide_ctx.closest_hit(rays)
    h, _, _ = orc.OracleScene(s, split_type=-1).lbvh_closest_hit(rays)
    assert np.array_equal(g["prim"], h["prim"]) and np.array_equal(g["t"].view(np.uint32), h["t"].view(np.uint32))
    n_nodes, _ = wide_ctx.bvh_wide_info()
    assert 100_000 < n_nodes < 400_000
    commit_wide(ptb, wide_ctx, overshadowed)
    wide_ctx.accum_clear()
    wide_ctx.stats_reset()
    wide_ctx.render(ptb.RenderOptions(samples_per_pixel=8, render_method=1, width=160, height=90, seed=3))
    st = wide_ctx.stats()
    img = wide_ctx.accum_read(160, 90)
    acc, counts, _ = orc.OracleScene(overshadowed).render(160, 90, 8, 1, seed=3)
    assert float(np.sqrt(np.mean((img - acc / 8) ** 2))) < 3e-2
    assert abs(st.rays_shadow_light - counts["shadow_light"]) <= 2e-3 * counts["shadow_light"]
    assert abs(st.rays_shadow_sky - counts["shadow_sky"]) <= 2e-3 * max(counts["shadow_sky"], 1)


@pytest.mark.parametrize("method", [0, 1])
def test_render_does_not_depend_on_the_tree(ptb, gpu_ctx, wide_ctx, overshadowed, method, monkeypatch):
    scenes = [overshadowed, ptb.meshgen.c3_scene(0.05)] if method == 0 else [overshadowed]
    for scene in scenes:
        o = ptb.RenderOptions(samples_per_pixel=16, render_method=method, width=160, height=90, seed=4)
        gpu_ctx.upload(scene)
        gpu_ctx.commit(ptb._lib.BUILD_BINARY)
        gpu_ctx.accum_clear()
        gpu_ctx.stats_reset()
        gpu_ctx.render(o)
        a, sa = gpu_ctx.accum_read(160, 90).copy(), gpu_ctx.stats()
        commit_wide(ptb, wide_ctx, scene)
        for tail in ("0", "65536"):                       # wavefront to the end, and with the fused tail kernel
            monkeypatch.setenv("PTB_TAIL_PATHS", tail)
            wide_ctx.accum_clear()
            wide_ctx.stats_reset()
            wide_ctx.render(o)
            b, sb = wide_ctx.accum_read(160, 90), wide_ctx.stats()
            assert np.allclose(a, b, rtol=1e-5, atol=1e-5)
            assert (sa.rays_camera, sa.rays_bounce, sa.rays_shadow_light, sa.rays_shadow_sky, sa.rays_reference, sa.paths) == \
                   (sb.rays_camera, sb.rays_bounce, sb.rays_shadow_light, sb.rays_shadow_sky, sb.rays_reference, sb.paths)
        monkeypatch.delenv("PTB_TAIL_PATHS")
