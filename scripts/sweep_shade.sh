#!/bin/bash
# k_shade block size sweep on the sphere scenes (MIS) and C3 (naive), B200
for t in 64 128 256; do
 for wl in "rtweekend1 --spp-per-step 16" "overshadowed --spp-per-step 64" "c3 --spp-per-step 64"; do
  echo -n "threads=$t $wl: "; PTB_SHADE_THREADS=$t python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'shadow', round(r['k_shadow_ms']), 'gen', round(r['k_generate_ms']))"
 done
done
