#!/usr/bin/env python
"""Regenerates the fixtures in this directory from the ORACLE (oracle/, the CPU restatement of the reference).

The Rust reference cannot be built or run in this image, so these are not outputs of the reference
binary: they freeze the oracle's answers — which are pinned to the reference's own known answers by
tests/test_oracle_kats.py and tests/test_materials_textures.py — so that (a) a later change to the oracle that moves
its answers is caught on CPU and (b) the device is compared against committed numbers as well as against a live oracle.

    python tests/golden/make_golden.py        # rewrites *.npz next to this file
    python tests/golden/make_golden.py sah    # only the sah_*.npz fixtures (the SAH builder's tree)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402
import ptb200  # noqa: E402


def golden_rays(centre, radius):
    """The fixture's rays are not stored: they are the first 2048 of the Philox stream below."""
    return ptb200.meshgen.philox_rays(2048, seed=0x601D, centre=tuple(float(c) for c in centre), radius=float(radius))


def scenes():
    yield "rtweekend1", ptb200.load_file(os.path.join(ROOT, "scenes", "rtweekend1.ssml")), (0.0, 1.0, 0.0), 3.0
    yield "overshadowed", ptb200.load_file(os.path.join(ROOT, "scenes", "overshadowed.ssml")), (-0.3, 0.3, -0.3), 1.5
    yield "c3_small", ptb200.meshgen.c3_scene(0.03), (0.0, 4.0, 1.0), 5.0


def sah_fixture(scene):
    """The CPU definition of the device SAH builder (oracle/sah_ref.hpp): primitive order, topology, a checksum of the boxes."""
    o = orc.OracleScene(scene, split_type=-1)
    o.lbvh_sah()
    _, order, nodes = o.lbvh_export()
    return {"order": order, "node_children": np.stack([nodes["left"], nodes["right"], nodes["parent"]], 1),
            "node_box_sum": np.array([nodes[k].astype(np.float64).sum() for k in ("lmin", "lmax", "rmin", "rmax")])}


def main_sah():
    for name, scene, _, _ in scenes():
        np.savez_compressed(os.path.join(HERE, f"sah_{name}.npz"), **sah_fixture(scene))


def main():
    main_sah()
    for name, scene, centre, radius in scenes():
        o = orc.OracleScene(scene)
        rays = golden_rays(centre, radius)
        hits = o.closest_hit(rays)                       # reference semantics: SAH tree, BFS candidates, test-all
        morton, order, nodes = o.lbvh_export()           # the CPU definition of the device LBVH
        out = {"hits": hits, "morton": morton, "order": order, "centre": np.array(centre), "radius": np.array(radius)}
        if len(nodes) <= 4096:
            out["nodes"] = nodes
        else:                                            # large trees: topology + a box checksum instead of 64 B per node
            out["node_children"] = np.stack([nodes["left"], nodes["right"], nodes["parent"]], 1)
            out["node_box_sum"] = np.array([nodes[k].astype(np.float64).sum() for k in ("lmin", "lmax", "rmin", "rmax")])
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    # one small image per integrator (shared counter-based RNG: a pure function of scene, seed and sample range)
    rt = ptb200.load_file(os.path.join(ROOT, "scenes", "rtweekend1.ssml"))
    o = orc.OracleScene(rt)
    imgs = {}
    for method, tag in ((0, "naive"), (1, "mis")):
        acc, counts, _ = o.render(48, 27, 8, method, seed=11)
        imgs[tag] = (acc / 8).astype(np.float32)
        imgs[tag + "_rays"] = np.array([counts["camera"], counts["bounce"], counts["shadow_sky"], counts["reference"]], np.uint64)
    np.savez_compressed(os.path.join(HERE, "render_rtweekend1_48x27x8.npz"), **imgs)


if __name__ == "__main__":
    main_sah() if sys.argv[1:] == ["sah"] else main()
