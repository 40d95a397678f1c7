#!/bin/bash
# two concurrent chunk streams: default suite + step-size sweep (run under gpurun)
(time python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/r2f_pytest.log 2>&1; tail -6 gpurun_out/r2f_pytest.log
rm -f gpurun_out/sweep_lines.jsonl
scripts/bench_sweep.sh "c3_256::--steps 3 --warmup 2" "c3_256_tail256k:PTB_TAIL_PATHS=262144:--steps 3 --warmup 2" \
  "c3_256_4GB:PTB_POOL_BYTES=4294967296:--steps 3 --warmup 2" "c3_256_16GB:PTB_POOL_BYTES=17179869184:--steps 3 --warmup 2" \
  "c3_64::--steps 4 --warmup 2 --spp-per-step 64" "c3_32::--steps 6 --warmup 2 --spp-per-step 32" "c3_16::--steps 8 --warmup 2 --spp-per-step 16" \
  "c3_32_tail256k:PTB_TAIL_PATHS=262144:--steps 6 --warmup 2 --spp-per-step 32" \
  "rt1_64::--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" "c2_64::--workload overshadowed --steps 3 --warmup 2 --spp-per-step 64" 2>&1 | tee gpurun_out/r2f_sweep.log
