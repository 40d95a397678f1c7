#!/bin/bash
# run the C3 bench (and rtweekend1 4K MIS) against each tuning build in build/variants (B200)
for so in "" build/variants/libptb200_*.so; do
 echo -n "${so:-default}: "; PTB200_LIB=${so:+$PWD/$so} python bench.py --steps 2 --warmup 3 --spp-per-step 64 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'gen', round(r['k_generate_ms']), end=' | ')"
 PTB200_LIB=${so:+$PWD/$so} python bench.py --workload rtweekend1 --steps 2 --spp-per-step 16 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('rt1', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'shadow', round(r['k_shadow_ms']), 'gen', round(r['k_generate_ms']))"
done
