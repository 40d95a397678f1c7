#!/bin/bash
# One-line summaries of bench.py runs under different environments / step sizes (tuning aid; run under gpurun).
#   scripts/bench_sweep.sh "label1:ENV=.. ENV=..:--bench --flags" "label2::..." ...
for spec in "$@"; do
  label="${spec%%:*}"; rest="${spec#*:}"; envs="${rest%%:*}"; flags="${rest#*:}"
  line=$(env $envs python bench.py --no-cpu --no-e2e --no-c5-leg $flags 2>/dev/null | tail -1)
  echo "$line" >> gpurun_out/sweep_lines.jsonl
  echo "$line" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; m=r['ms_by_kernel']
print('$label', 'Mrays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'trace', round(m['k_trace']), 'shade', round(m['k_shade']), 'shadow', round(m['k_shadow']), 'book', round(m['bookkeeping']), 'tail', round(m.get('k_tail',0),1), 'launches', d['gpu_launches'], 'iters', d['run']['wavefront_iterations'])"
done
