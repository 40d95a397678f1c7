#!/bin/bash
# Strong scaling on an N-GPU box (gpurun --gpus N -- scripts/gpu_scale.sh N): the multi-GPU tests, then the bench at 1 and N
# ranks on the same box — C3 (256 spp in total) and the C4-style config (rtweekend1 3840x2160 MIS, 4096 spp in total).
N=${1:-2}
(time python -m pytest tests -m gpu -q -k "multi or gpus or tiles") > gpurun_out/scale${N}_pytest.log 2>&1; tail -2 gpurun_out/scale${N}_pytest.log
run() {  # label, ranks, extra flags
  if [ $2 -eq 1 ]; then python bench.py --steps 3 --warmup 3 --no-cpu --no-c5-leg $3 > gpurun_out/scale${N}_$1_n1.json 2> gpurun_out/scale${N}_$1_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 3 --warmup 3 $3 \
    > gpurun_out/scale${N}_$1_n$2.json 2> gpurun_out/scale${N}_$1_n$2.err; fi
}
run c3 1 ""; run c3 $N ""
run c4 1 "--workload rtweekend1 --spp-per-step 4096"; run c4 $N "--workload rtweekend1 --spp-per-step 4096"
python - <<PY
import json
for w in ("c3", "c4"):
    a=json.loads(open('gpurun_out/scale${N}_%s_n1.json' % w).read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/scale${N}_%s_n${N}.json' % w).read().strip().splitlines()[-1])
    print(w, 'N=1', round(a['value']), 'ms/step', round(a['ms_per_step'],2), 'e2e', round(a['e2e']['value']))
    print(w, 'N=${N}', round(b['value']), 'ms/step', round(b['ms_per_step'],2), 'e2e', round(b['e2e']['value']), 'scaling', b['scaling'], 'efficiency', round(b['value']/a['value']/${N},3), 'e2e efficiency', round(b['e2e']['value']/a['e2e']['value']/${N},3), 'image diff', b.get('multi_gpu_image_max_abs_diff'))
PY
