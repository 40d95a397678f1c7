set -x
B="python bench.py --steps 3 --warmup 3"
$B --workload rtweekend1 --width 800 --height 450 --spp-per-step 64 --method mis > gpurun_out/r2fin_c1_mis.json 2>/dev/null
$B --workload rtweekend1 --width 800 --height 450 --spp-per-step 64 --method naive > gpurun_out/r2fin_c1_naive.json 2>/dev/null
$B --workload overshadowed > gpurun_out/r2fin_c2.json 2>/dev/null
$B --workload rtweekend1 --spp-per-step 512 --no-cpu > gpurun_out/r2fin_c4share.json 2>/dev/null
$B --workload closest_hit --rays 50331648 > gpurun_out/r2fin_c5.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2fin_ref.json 2>/dev/null
for f in c1_mis c1_naive c2 c4share c5 ref; do python - <<PY
import json
d=json.loads(open('gpurun_out/r2fin_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value'],2), round(d['ms_per_step'],2), d.get('e2e') and round(d['e2e']['value'],2), (d.get('cpu_baseline') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
PY
done
