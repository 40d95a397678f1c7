#!/bin/bash
# both suites + binary / wide A/B (run under gpurun)
scripts/gpu_suite.sh r2e
rm -f gpurun_out/sweep_lines.jsonl
scripts/bench_sweep.sh "c3_256_bin:PTB_BVH=binary:--steps 3 --warmup 2" "c3_256_wide3:PTB_BVH=wide:--steps 3 --warmup 2" \
  "c3_256_wide1:PTB_BVH=wide PTB_WIDE_LEAF=1:--steps 3 --warmup 2" "c3_256_wide2:PTB_BVH=wide PTB_WIDE_LEAF=2:--steps 3 --warmup 2" \
  "c3_32_bin:PTB_BVH=binary:--steps 6 --warmup 2 --spp-per-step 32" "c3_32_wide3:PTB_BVH=wide:--steps 6 --warmup 2 --spp-per-step 32" \
  "rt1_64_bin:PTB_BVH=binary:--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" "rt1_64_wide3:PTB_BVH=wide:--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" \
  "rt1_64_wide1:PTB_BVH=wide PTB_WIDE_LEAF=1:--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" \
  "c2_64_bin:PTB_BVH=binary:--workload overshadowed --steps 3 --warmup 2 --spp-per-step 64" "c2_64_wide3:PTB_BVH=wide:--workload overshadowed --steps 3 --warmup 2 --spp-per-step 64" \
  "c2_64_wide1:PTB_BVH=wide PTB_WIDE_LEAF=1:--workload overshadowed --steps 3 --warmup 2 --spp-per-step 64" 2>&1 | tee gpurun_out/r2e_sweep.log
for t in binary wide; do
  PTB_BVH=$t python bench.py --workload closest_hit --no-cpu --no-e2e --steps 3 --warmup 2 --rays 50000000 2>/dev/null | tail -1 > gpurun_out/r2e_c5_$t.json
  python -c "
import json; d=json.load(open('gpurun_out/r2e_c5_$t.json')); r=d['roofline']; print('c5 $t', round(d['value']), 'frac', round(r['frac'],3), 'V', round(r['nodes_per_ray'],2), 'T', round(r['prims_per_ray'],2), 'build_ms', d['run'].get('build_ms'))"
done 2>&1 | tee -a gpurun_out/r2e_sweep.log
