#!/usr/bin/env python
"""ptb_render when the device cannot provide the path pool: the pool is halved until it fits and the call runs in more
chunks, same image. Hogs the GPU with a torch tensor after the context has cached its memory budget. (B200, manual check)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ptb200

scene = ptb200.meshgen.c3_scene(0.05)
ctx = ptb200.Context(0)
sc = ptb200.Scene(scene, ctx=ctx)
small = ptb200.RenderOptions(samples_per_pixel=10, render_method=0, width=1920, height=1080, seed=3)  # 20.7 M paths: budget asked
big = ptb200.RenderOptions(samples_per_pixel=64, render_method=0, width=1920, height=1080, seed=3)    # 132.7 M paths: 9.2 GB
sc.render(small)  # caches the budget (half of ~178 GB free) and allocates 1.4 GB
free, total = torch.cuda.mem_get_info()
hog = torch.empty(int(free - 6.0e9), dtype=torch.uint8, device="cuda")  # leaves ~6 GB: the 9.2 GB pool cannot be allocated
ctx.stats_reset()
img = sc.render(big)
st = ctx.stats()
print("free before hog %.1f GB, while rendering %.1f GB" % (free / 1e9, torch.cuda.mem_get_info()[0] / 1e9))
del hog
torch.cuda.empty_cache()
ctx2 = ptb200.Context(0)
sc2 = ptb200.Scene(scene, ctx=ctx2)
ctx2.stats_reset()
ref = sc2.render(big)
print("iterations: unconstrained", ctx2.stats().wavefront_iterations, "constrained", st.wavefront_iterations,
      " max |diff|", float(np.max(np.abs(img - ref))))
assert st.wavefront_iterations > ctx2.stats().wavefront_iterations  # more chunks
assert np.allclose(img, ref, rtol=1e-5, atol=1e-5)
print("ok")
