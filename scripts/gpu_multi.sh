#!/bin/bash
# multi-GPU checks on an N-GPU box (gpurun --gpus N -- scripts/gpu_multi.sh N): the multi-GPU tests, then the strong-scaling
# bench at 1 and N ranks on the same box
N=${1:-2}
(time python -m pytest tests -m gpu -q -k "multi or gpus or tiles") > gpurun_out/multi${N}_pytest.log 2>&1; tail -4 gpurun_out/multi${N}_pytest.log
python bench.py --steps 4 --warmup 3 --no-cpu --no-c5-leg > gpurun_out/multi${N}_n1.json 2> gpurun_out/multi${N}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 \
  > gpurun_out/multi${N}_n${N}.json 2> gpurun_out/multi${N}_n${N}.err
python - <<PY
import json
a=json.loads(open('gpurun_out/multi${N}_n1.json').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/multi${N}_n${N}.json').read().strip().splitlines()[-1])
print('N=1', round(a['value']), 'ms/step', round(a['ms_per_step'],2), 'e2e', round(a['e2e']['value']))
print('N=${N}', round(b['value']), 'ms/step', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['value']), 'scaling', b['scaling'], 'efficiency', round(b['value']/a['value']/${N},3), 'image diff', b.get('multi_gpu_image_max_abs_diff'))
PY
