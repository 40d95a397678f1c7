"""K2-K6 on the device vs oracle/lbvh_ref.hpp: Morton keys, sorted order, links and boxes must be BIT-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(ptb, orc, gpu_ctx, scene):
    gpu_ctx.upload(scene)
    gpu_ctx.commit(ptb._lib.BUILD_BINARY)
    gm, gp, gn = gpu_ctx.bvh_export()
    o = orc.OracleScene(scene, split_type=-1)
    om, op, on = o.lbvh_export()
    assert np.array_equal(gm, om), "Morton codes differ"
    assert np.array_equal(gp, op), "sorted primitive order differs"
    assert len(gn) == len(on)
    for f in ("left", "right", "parent"):
        assert np.array_equal(gn[f], on[f]), f"node field {f} differs"
    for f in ("lmin", "lmax", "rmin", "rmax"):
        assert np.array_equal(gn[f].view(np.uint32), on[f].view(np.uint32)), f"node boxes {f} differ (bitwise)"
    # sortedness + permutation (size-independent properties)
    assert np.all(np.diff(gm.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(gp), np.arange(scene.n_primitives, dtype=np.uint32))
    # the 32-byte nodes the traversal kernels read: grid and 16-bit boxes bit-exact, and every box contains its f32 box
    if not has_quantised_nodes(ptb, gpu_ctx):
        return
    gf, gq = gpu_ctx.bvh_export_quantised()
    of, oq = o.lbvh_quantise()
    assert np.array_equal(gf.view(np.uint32), of.view(np.uint32)), "quantisation grid differs"
    assert np.array_equal(gq, oq), "quantised nodes differ"
    check_quantised_contains(gn, gf, gq)


def has_quantised_nodes(ptb, ctx):
    """The 32-byte traversal nodes are a compile-time option of the library (-DPTB_QNODES=1; scripts/gpu_suite.sh runs the
    suite under that build too): the default build answers PTB_ERR_UNSUPPORTED."""
    try:
        ctx.bvh_export_quantised()
        return True
    except ptb.PtbError as e:
        assert e.code == 8  # PTB_ERR_UNSUPPORTED
        return False


def check_quantised_contains(nodes, frame, q):
    org, step = frame[:3].astype(np.float64), frame[3:].astype(np.float64)
    for side, mn, mx in ((0, "lmin", "lmax"), (3, "rmin", "rmax")):
        w = q[:, side:side + 3]
        lo = org + (w & 0xFFFF).astype(np.float64) * step
        hi = org + (w >> 16).astype(np.float64) * step
        assert np.all(lo <= nodes[mn].astype(np.float64)) and np.all(hi >= nodes[mx].astype(np.float64))
        flat = step == 0
        assert np.all((nodes[mn].astype(np.float64) - lo)[:, ~flat] < step[~flat]) and np.all((hi - nodes[mx].astype(np.float64))[:, ~flat] < step[~flat])
    assert np.array_equal(q[:, 6], nodes["left"]) and np.array_equal(q[:, 7], nodes["right"])


def test_rtweekend1(ptb, orc, gpu_ctx, rtweekend1):
    _compare(ptb, orc, gpu_ctx, rtweekend1)


def test_overshadowed(ptb, orc, gpu_ctx, overshadowed):
    _compare(ptb, orc, gpu_ctx, overshadowed)


def test_single_primitive(ptb, orc, gpu_ctx, rtweekend1):
    import copy
    s = copy.deepcopy(rtweekend1)
    s.spheres = s.spheres[:1].copy()
    _compare(ptb, orc, gpu_ctx, s)


@pytest.mark.parametrize("scale", [0.02, 0.11])
def test_c3_mesh(ptb, orc, gpu_ctx, scale):
    _compare(ptb, orc, gpu_ctx, ptb.meshgen.c3_scene(scale))


def test_duplicate_keys(ptb, orc, gpu_ctx, rtweekend1):
    """Many primitives with identical centroids (identical Morton keys): tie-break by index must match."""
    import copy
    s = copy.deepcopy(rtweekend1)
    sp = np.zeros(5000, ptb._lib.sphere_dtype)
    sp["center"] = np.repeat(np.array([[0, 1, 0], [1, 1, 0], [0, 2, 0], [0, 1, 0.5], [3, 3, 3]], np.float32), 1000, axis=0)
    sp["radius"] = np.tile(np.linspace(0.01, 0.5, 1000, dtype=np.float32), 5)
    sp["material"] = 0
    s.spheres = sp
    _compare(ptb, orc, gpu_ctx, s)


def test_full_size_properties(ptb, gpu_ctx):
    """BASELINE config C3 at full size (1 000 000 triangles): sortedness, permutation, tree consistency."""
    s = ptb.meshgen.c3_scene(1.0)
    assert len(s.triangles) == 1_000_000
    gpu_ctx.upload(s)
    gpu_ctx.commit(ptb._lib.BUILD_BINARY)
    gm, gp, gn = gpu_ctx.bvh_export()
    assert np.all(np.diff(gm.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(gp), np.arange(1_000_000, dtype=np.uint32))
    leaf = 0x80000000
    # every leaf referenced exactly once, every internal node (except the root) exactly once
    refs = np.concatenate([gn["left"], gn["right"]])
    leaves = np.sort(refs[refs >= leaf] - leaf)
    inner = np.sort(refs[refs < leaf])
    assert np.array_equal(leaves, np.arange(1_000_000, dtype=np.uint32))
    assert np.array_equal(inner, np.arange(1, len(gn), dtype=np.uint32))
    # parents' child boxes contain the children's own child boxes
    for side, mn, mx in (("left", "lmin", "lmax"), ("right", "rmin", "rmax")):
        ch = gn[side]
        m = ch < leaf
        c = gn[ch[m]]
        cmin = np.minimum(c["lmin"], c["rmin"])
        cmax = np.maximum(c["lmax"], c["rmax"])
        assert np.array_equal(gn[mn][m], cmin) and np.array_equal(gn[mx][m], cmax)
    if has_quantised_nodes(ptb, gpu_ctx):
        check_quantised_contains(gn, *gpu_ctx.bvh_export_quantised())
