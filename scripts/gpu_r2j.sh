#!/bin/bash
rm -f gpurun_out/sweep_lines.jsonl
scripts/bench_sweep.sh "c3_256::--steps 3 --warmup 2" "c3_64::--steps 4 --warmup 2 --spp-per-step 64" "c3_32::--steps 6 --warmup 2 --spp-per-step 32" \
  "c3_16::--steps 8 --warmup 2 --spp-per-step 16" "c3_32_t256k:PTB_TAIL_PATHS=262144:--steps 6 --warmup 2 --spp-per-step 32" "c3_32_t16k:PTB_TAIL_PATHS=16384:--steps 6 --warmup 2 --spp-per-step 32" 2>&1 | tee gpurun_out/r2j_sweep.log
(time python -m pytest tests/test_gpu_render.py -m gpu -q -x) 2>&1 | tail -4
