#!/usr/bin/env python
"""How much of k_trace's cost is ray incoherence? C3 at 1080p with max_depth 1 (camera rays only), 2, 3, 50:
k_trace Mrays/s (CUDA events per kernel class) and nodes / primitives per ray (counted in a separate, untimed render)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ptb200

scene = ptb200.meshgen.c3_scene(1.0)
ctx = ptb200.Context(0)
sc = ptb200.Scene(scene, ctx=ctx)
spp = int(os.environ.get("SPP", "16"))
prev = None
for depth in (1, 2, 3, 5, 50):
    o = ptb200.RenderOptions(samples_per_pixel=spp, render_method=ptb200.METHOD_NAIVE, width=1920, height=1080, seed=1, max_depth=depth)
    sc.render(o)  # warm
    ctx.set_option(ptb200._lib.OPT_TIME_KERNELS, 1)
    ctx.stats_reset()
    sc.render(o)
    st = ctx.stats()
    rays, ms = st.rays_total, st.ms_trace
    ctx.set_option(ptb200._lib.OPT_TIME_KERNELS, 0)
    ctx.set_option(ptb200._lib.OPT_COUNT_TRAVERSAL, 1)
    ctx.stats_reset()
    sc.render(o)
    c = ctx.stats()
    ctx.set_option(ptb200._lib.OPT_COUNT_TRAVERSAL, 0)
    line = f"depth<={depth:2d}: rays {rays/1e6:8.1f} M  k_trace {ms:8.2f} ms  {rays/ms/1e3:7.0f} Mrays/s  V {c.nodes_fetched/c.rays_counted:6.2f}  T {c.prims_tested/c.rays_counted:5.2f}  shade {st.ms_shade:7.2f} ms gen {st.ms_generate:6.2f} ms launches {st.trace_launches}"
    if prev:
        dr, dm = rays - prev[0], ms - prev[1]
        line += f"   | marginal: {dr/1e6:7.1f} M rays in {dm:7.2f} ms = {dr/max(dm,1e-9)/1e3:6.0f} Mrays/s"
    print(line, flush=True)
    prev = (rays, ms)
