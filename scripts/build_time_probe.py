#!/usr/bin/env python
"""GPU: commit time (ptb_stats.build_ms, CUDA events inside the library) of the LBVH and the SAH builder on the C3 mesh and the
C5 heightfield; the second commit of each is the one without allocations. Usage: build_time_probe.py [c3|c5|both] [repeats]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = ptb200._lib
scenes = []
if which in ("c3", "both"):
    scenes.append(("c3 1M", ptb200.meshgen.c3_scene(1.0)))
if which in ("c5", "both"):
    scenes.append(("c5 10M", ptb200.meshgen.heightfield_scene()))
for name, scene in scenes:
    for label, flag in (("lbvh", L.BUILD_BINARY), ("sah", L.BUILD_SAH)):
        ctx = ptb200.Context(0)
        ctx.upload(scene)
        ms = []
        for _ in range(reps):
            ctx.commit(flag)
            ms.append(round(ctx.stats().build_ms, 3))
        print(name, label, "build_ms", ms, "builder/levels", ctx.bvh_builder(), flush=True)
        ctx.close()
