#!/bin/bash
# run the C3 bench against each tuning build in build/variants (B200), optionally with PTB_TRACE_PRIM_BIAS values
for so in "" build/variants/libptb200_*.so; do for bias in 0 1 2; do
 echo -n "${so:-default} bias=$bias: "; PTB_TRACE_PRIM_BIAS=$bias PTB200_LIB=${so:+$PWD/$so} python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'V', round(r['nodes_per_ray'],2), 'T', round(r['prims_per_ray'],2))"
done; done
