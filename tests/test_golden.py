"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle): closest hits of the
reference-semantics BVH, the LBVH definition (Morton codes, order, topology, boxes) and two small rendered images.
CPU half: the oracle still reproduces them bit for bit. GPU half (-m gpu): the device reproduces them through the C ABI."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as G  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = {name: (scene, centre, radius) for name, scene, centre, radius in G.scenes()}


def _check_lbvh(g, morton, order, nodes):
    assert np.array_equal(morton, g["morton"]) and np.array_equal(order, g["order"])
    if "nodes" in g:
        for k in ("lmin", "lmax", "rmin", "rmax", "left", "right", "parent"):
            assert np.array_equal(nodes[k], g["nodes"][k]), k
    else:
        assert np.array_equal(np.stack([nodes["left"], nodes["right"], nodes["parent"]], 1), g["node_children"])
        sums = np.array([nodes[k].astype(np.float64).sum() for k in ("lmin", "lmax", "rmin", "rmax")])
        assert np.array_equal(sums, g["node_box_sum"])


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_reproduces_golden(orc, name):
    scene, centre, radius = SCENES[name]
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    o = orc.OracleScene(scene)
    hits = o.closest_hit(G.golden_rays(centre, radius))
    assert np.array_equal(hits["prim"], g["hits"]["prim"])
    assert np.array_equal(hits["t"].view(np.uint32), g["hits"]["t"].view(np.uint32))
    _check_lbvh(g, *o.lbvh_export())


def test_oracle_reproduces_golden_images(orc, rtweekend1):
    g = np.load(os.path.join(GOLDEN, "render_rtweekend1_48x27x8.npz"))
    o = orc.OracleScene(rtweekend1)
    for method, tag in ((0, "naive"), (1, "mis")):
        acc, counts, _ = o.render(48, 27, 8, method, seed=11)
        assert np.allclose(acc / 8, g[tag], rtol=0, atol=1e-6)   # libm of another glibc may differ in the last ulp
        assert [counts["camera"], counts["bounce"], counts["shadow_sky"], counts["reference"]] == list(g[tag + "_rays"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_device_reproduces_golden(ptb, gpu_ctx, name):
    scene, centre, radius = SCENES[name]
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    gpu_ctx.upload(scene)
    gpu_ctx.commit(ptb._lib.BUILD_BINARY)   # the golden tree is the Karras LBVH's (the default builder is the SAH one)
    hits = gpu_ctx.closest_hit(G.golden_rays(centre, radius))
    assert np.array_equal(hits["prim"], g["hits"]["prim"])                       # primitive id exact
    m = hits["prim"] != ptb.PTB_MISS
    assert np.array_equal(hits["t"][m].view(np.uint32), g["hits"]["t"][m].view(np.uint32))   # t bit-identical
    _check_lbvh(g, *gpu_ctx.bvh_export())                                        # Morton sort + topology + boxes bit-exact


def _check_sah(g, order, nodes):
    assert np.array_equal(order, g["order"])
    leaf, sphere_bit = 0x80000000, np.uint32(0x40000000)
    kids = np.stack([nodes["left"], nodes["right"], nodes["parent"]], 1)
    kids[:, :2] = np.where(kids[:, :2] >= leaf, kids[:, :2] & ~sphere_bit, kids[:, :2])
    assert np.array_equal(kids, g["node_children"])
    sums = np.array([nodes[k].astype(np.float64).sum() for k in ("lmin", "lmax", "rmin", "rmax")])
    assert np.array_equal(sums, g["node_box_sum"])


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_reproduces_golden_sah_tree(orc, name):
    g = np.load(os.path.join(GOLDEN, f"sah_{name}.npz"))
    f = G.sah_fixture(SCENES[name][0])
    for k in ("order", "node_children", "node_box_sum"):
        assert np.array_equal(f[k], g[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_device_reproduces_golden_sah_tree(ptb, gpu_ctx, name):
    """The device SAH builder against the committed tree of its CPU definition (order, links; boxes by checksum)."""
    g = np.load(os.path.join(GOLDEN, f"sah_{name}.npz"))
    gpu_ctx.upload(SCENES[name][0])
    gpu_ctx.commit(ptb._lib.BUILD_SAH)
    _, order, nodes = gpu_ctx.bvh_export()
    _check_sah(g, order, nodes)


@pytest.mark.gpu
def test_device_reproduces_golden_images(ptb, gpu_ctx, rtweekend1):
    g = np.load(os.path.join(GOLDEN, "render_rtweekend1_48x27x8.npz"))
    sc = ptb.Scene(rtweekend1, ctx=gpu_ctx)
    for method, tag in ((0, "naive"), (1, "mis")):
        gpu_ctx.stats_reset()
        img = sc.render(ptb.RenderOptions(samples_per_pixel=8, render_method=method, width=48, height=27, seed=11))
        assert float(np.sqrt(np.mean((img - g[tag]) ** 2))) < 5e-3               # per-channel RMSE, linear radiance
        st = gpu_ctx.stats()
        assert st.rays_camera == g[tag + "_rays"][0]
        assert abs(int(st.rays_bounce) - int(g[tag + "_rays"][1])) <= 3e-3 * int(g[tag + "_rays"][1]) + 2
