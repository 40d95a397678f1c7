#!/bin/bash
# wavefront pool-size sweep (paths in flight) on C3 (naive, 64 and 256 spp per step) and rtweekend1 4K (MIS), B200
for p in ${POOLS:-16777216 33554432 67108864 134217728 268435456}; do
 echo -n "pool=$p: "; PTB_POOL_PATHS=$p bash scripts/quick_bench.sh 2>&1 | head -2 | tr '\n' '|'; 
 PTB_POOL_PATHS=$p python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('c3@256spp', round(d['value']), 'ms/step', round(d['ms_per_step'],1))"
done
