#!/bin/bash
# resource-light tail: tail size / block-count sweep (run under gpurun)
(time python -m pytest tests/test_gpu_render.py tests/test_gpu_wide.py tests/test_gpu_baseline_size.py -m gpu -q -x -k "tail or chunking or tree or c2_over or window") > gpurun_out/r2h_pytest.log 2>&1; tail -4 gpurun_out/r2h_pytest.log
rm -f gpurun_out/sweep_lines.jsonl
scripts/bench_sweep.sh "c3_256::--steps 3 --warmup 2" "c3_32::--steps 6 --warmup 2 --spp-per-step 32" "c3_16::--steps 8 --warmup 2 --spp-per-step 16" \
  "c3_32_t64k:PTB_TAIL_PATHS=65536:--steps 6 --warmup 2 --spp-per-step 32" "c3_32_t8k:PTB_TAIL_PATHS=8192:--steps 6 --warmup 2 --spp-per-step 32" \
  "c3_32_b74:PTB_TAIL_BLOCKS=74:--steps 6 --warmup 2 --spp-per-step 32" "c3_32_b296:PTB_TAIL_BLOCKS=296:--steps 6 --warmup 2 --spp-per-step 32" \
  "c3_256_t8k:PTB_TAIL_PATHS=8192:--steps 3 --warmup 2" "c3_256_b74:PTB_TAIL_BLOCKS=74:--steps 3 --warmup 2" \
  "rt1_64::--workload rtweekend1 --steps 3 --warmup 2 --spp-per-step 64" "c2_64::--workload overshadowed --steps 3 --warmup 2 --spp-per-step 64" 2>&1 | tee gpurun_out/r2h_sweep.log
