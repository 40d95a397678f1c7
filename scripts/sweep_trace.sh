#!/bin/bash
# parameter sweep helper (B200): traversal scheduling knobs (node-step burst x refill threshold x primitive-phase bias) on C3
for b in 2 4 8; do for f in 4 8 12 16 24; do
 echo -n "burst=$b fetch=$f: "; PTB_TRACE_BURST=$b PTB_TRACE_FETCH=$f python bench.py --steps 2 --warmup 2 --spp-per-step 64 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'V', round(r['nodes_per_ray'],2))"
done; done
for bias in 1 2 3; do
 echo -n "prim_bias=$bias: "; PTB_TRACE_PRIM_BIAS=$bias python bench.py --steps 2 --warmup 2 --spp-per-step 64 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'V', round(r['nodes_per_ray'],2))"
done
