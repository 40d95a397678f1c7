#!/usr/bin/env python
"""Window mode, block-level material sort in k_shade on / off (PTB_SHADE_SORT): the five-material showcase scene of
tests/test_gpu_materials.py at 1920x1080, naive and MIS. Prints per-class kernel times (CUDA events) and the rate."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ptb200
from test_gpu_materials import showcase_scene

spp = int(os.environ.get("SPP", "64"))
ctx = ptb200.Context(0)
sc = ptb200.Scene(showcase_scene(ptb200, sky="image"), ctx=ctx)
for method in (0, 1):
    for sort in ("0", "1"):
        os.environ["PTB_SHADE_SORT"] = sort
        o = ptb200.RenderOptions(samples_per_pixel=spp, render_method=method, width=1920, height=1080, seed=1)
        sc.render(o)  # warm
        ctx.set_option(ptb200._lib.OPT_TIME_KERNELS, 1)
        ctx.stats_reset()
        sc.render(o)
        st = ctx.stats()
        ctx.set_option(ptb200._lib.OPT_TIME_KERNELS, 0)
        print(f"method {method} sort {sort}: render {st.render_ms:8.2f} ms  {st.rays_total / st.render_ms / 1e3:8.0f} Mrays/s  "
              f"trace {st.ms_trace:7.2f} shade {st.ms_shade:7.2f} shadow {st.ms_shadow:7.2f} book {st.ms_generate:6.2f}", flush=True)
