import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ptb200
scene = ptb200.meshgen.c3_scene(1.0)
ctx = ptb200.Context(0)
sc = ptb200.Scene(scene, ctx=ctx)
ctx.set_option(ptb200._lib.OPT_TIME_KERNELS, 1)
for spp, off in ((16, 0), (16, 16), (16, 32), (16, 48), (8, 0), (8, 8), (32, 0), (32, 32), (64, 0), (7, 0), (9, 0)):
    o = ptb200.RenderOptions(samples_per_pixel=spp, sample_offset=off, render_method=0, width=1920, height=1080, seed=1, max_depth=1)
    ctx.accum_clear(); ctx.render(o)
    ctx.stats_reset(); ctx.accum_clear(); ctx.render(o)
    st = ctx.stats()
    print(f"spp {spp:3d} off {off:3d}: k_trace {st.ms_trace:7.2f} ms {st.rays_total/st.ms_trace/1e3:6.0f} Mrays/s launches {st.trace_launches} gen {st.ms_generate:.2f} shade {st.ms_shade:.2f}", flush=True)
