"""Render small previews of the benchmark scenes on the GPU (sanity check by eye). Output: gpurun_out/*.png"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptb200

os.makedirs("gpurun_out", exist_ok=True)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
jobs = [("c3", ptb200.meshgen.c3_scene(1.0), ptb200.METHOD_NAIVE, 256),
        ("rtweekend1", ptb200.load_file(os.path.join(root, "scenes/rtweekend1.ssml")), ptb200.METHOD_MIS, 64),
        ("overshadowed_naive", ptb200.load_file(os.path.join(root, "scenes/overshadowed.ssml")), ptb200.METHOD_NAIVE, 512),
        ("overshadowed_mis", ptb200.load_file(os.path.join(root, "scenes/overshadowed.ssml")), ptb200.METHOD_MIS, 128)]
ctx = ptb200.Context(0)
for name, scene, method, spp in jobs:
    sc = ptb200.Scene(scene, ctx=ctx)
    img = sc.render(ptb200.RenderOptions(samples_per_pixel=spp, render_method=method, width=640, height=360, seed=1))
    ptb200.save_image(f"gpurun_out/{name}.png", 640, 360, img, 2.2)
    print(name, "mean", img.mean(axis=(0, 1)), "finite", bool(np.isfinite(img).all()), ctx.stats().render_ms, "ms")
