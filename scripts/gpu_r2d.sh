#!/bin/bash
# tests + tail / pool sweep of the device-gated window pipeline (run under gpurun)
(time python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/r2d_pytest.log 2>&1; tail -8 gpurun_out/r2d_pytest.log
rm -f gpurun_out/sweep_lines.jsonl
scripts/bench_sweep.sh "c3_256::--steps 3 --warmup 2" "c3_256_tail16k:PTB_TAIL_PATHS=16384:--steps 3 --warmup 2" \
  "c3_256_tail256k:PTB_TAIL_PATHS=262144:--steps 3 --warmup 2" "c3_256_32GB:PTB_POOL_BYTES=80000000000:--steps 3 --warmup 2" \
  "c3_64::--steps 4 --warmup 2 --spp-per-step 64" "c3_32::--steps 6 --warmup 2 --spp-per-step 32" "c3_16::--steps 8 --warmup 2 --spp-per-step 16" \
  "c3_32_tail256k:PTB_TAIL_PATHS=262144:--steps 6 --warmup 2 --spp-per-step 32" "c3_32_tail16k:PTB_TAIL_PATHS=16384:--steps 6 --warmup 2 --spp-per-step 32" \
  "rt1_64::--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" "c2_64::--workload overshadowed --steps 3 --warmup 2 --spp-per-step 64" 2>&1 | tee gpurun_out/r2d_sweep.log
