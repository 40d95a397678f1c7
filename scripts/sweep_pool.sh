#!/bin/bash
# wavefront pool-size sweep on the sphere scenes (MIS), B200
for p in 262144 524288 1048576 2097152 4194304 16777216; do
 for wl in "rtweekend1 --spp-per-step 16" "overshadowed --spp-per-step 64"; do
  echo -n "pool=$p $wl: "; PTB_POOL_PATHS=$p python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'shadow', round(r['k_shadow_ms']), 'gen', round(r['k_generate_ms']), 'launches', d['gpu_launches'])"
 done
done
