python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for so in "" build/variants/libptb200_no256.so; do
  for spp in 64; do
    echo -n "${so:-default} spp=$spp: "; PTB200_LIB=${so:+$PWD/$so} python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --spp-per-step $spp 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'trace', round(r['k_trace_ms']), 'shade', round(r['k_shade_ms']), 'gen', round(r['k_generate_ms']), 'V', round(r['nodes_per_ray'],2), 'T', round(r['prims_per_ray'],2), d['clocks']['sm_mhz'])"
  done
done
SPP=64 python scripts/coherence_probe.py 2>&1 | tail -5
