#!/bin/bash
# one-chunk rule: full default suite + step-size sweep + the default bench line (run under gpurun)
(time python -m pytest tests -m gpu -q --durations=5) > gpurun_out/r2i_pytest.log 2>&1; tail -5 gpurun_out/r2i_pytest.log
rm -f gpurun_out/sweep_lines.jsonl
scripts/bench_sweep.sh "c3_256::--steps 3 --warmup 2" "c3_64::--steps 4 --warmup 2 --spp-per-step 64" "c3_32::--steps 6 --warmup 2 --spp-per-step 32" \
  "c3_16::--steps 8 --warmup 2 --spp-per-step 16" "c3_256_16GB:PTB_POOL_BYTES=17179869184:--steps 3 --warmup 2" \
  "rt1_64::--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" "c2_256::--workload overshadowed --steps 3 --warmup 2" 2>&1 | tee gpurun_out/r2i_sweep.log
python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 600 gpurun_out/r2i_bench.json; tail -3 gpurun_out/r2i_bench.err
