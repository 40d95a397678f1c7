#!/bin/bash
# timeline of a 32-spp C3 render + stream-count A/B (run under gpurun)
PTB_TIMELINE=1 python bench.py --no-cpu --no-e2e --no-c5-leg --steps 1 --warmup 3 --spp-per-step 32 2> gpurun_out/r2g_timeline.log | tail -1 > gpurun_out/r2g_line.json
grep -c timeline gpurun_out/r2g_timeline.log
rm -f gpurun_out/sweep_lines.jsonl
scripts/bench_sweep.sh "c3_256_s1::--steps 3 --warmup 2" "c3_32_s1::--steps 6 --warmup 2 --spp-per-step 32" \
  "rt1_64_s2::--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" "rt1_64_s1:PTB_CHUNK_STREAMS=1:--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" \
  "c2_64_s2::--workload overshadowed --steps 3 --warmup 2 --spp-per-step 64" "c2_64_s1:PTB_CHUNK_STREAMS=1:--workload overshadowed --steps 3 --warmup 2 --spp-per-step 64" 2>&1 | tee gpurun_out/r2g_sweep.log
