"""Wall-clock breakdown of one end-to-end step on the C3 workload (host buffers in, host image out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptb200

scene = ptb200.meshgen.c3_scene(1.0)
ctx = ptb200.Context(0)
w, h, spp = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 64
for it in range(3):
    t = [time.perf_counter()]
    ctx.upload(scene); t.append(time.perf_counter())
    ctx.commit(); ctx.synchronize(); t.append(time.perf_counter())
    ctx.accum_clear()
    ctx.render(ptb200.RenderOptions(samples_per_pixel=spp, render_method=0, width=w, height=h, sample_offset=it * spp)); t.append(time.perf_counter())
    img = ctx.accum_read(w, h, normalise=True); t.append(time.perf_counter())
    names = ["upload(set_*)", "commit(H2D+LBVH)", "render", "accum_read"]
    print(f"iter {it}: " + "  ".join(f"{n} {1e3 * (b - a):.1f} ms" for n, a, b in zip(names, t, t[1:])) +
          f"  | build_ms(device) {ctx.stats().build_ms:.2f}")
