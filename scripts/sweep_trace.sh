#!/bin/bash
# parameter sweep helper (B200): traversal scheduling knobs and wavefront pool size on the C3 workload
for p in 1048576 2097152 4194304 8388608 16777216; do for f in 6 12; do
 echo -n "pool=$p fetch=$f: "; PTB_POOL_PATHS=$p PTB_TRACE_BURST=4 PTB_TRACE_FETCH=$f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'gen', round(r['k_generate_ms']), 'launches', d['gpu_launches'])"
done; done
