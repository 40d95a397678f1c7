set -x
T=v
B="python bench.py --no-cpu --no-e2e --no-c5-leg"
NCU="ncu --set full --clock-control none "
python bench.py > gpurun_out/r2${T}_bench.json 2> gpurun_out/r2${T}_bench.err || exit 1
$B --steps 2 --warmup 1 > gpurun_out/r2${T}_plain.json 2> gpurun_out/r2${T}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2${T}_launches_c3.csv $B --steps 2 --warmup 1 > gpurun_out/r2${T}_ncu_l.log 2>&1
$B --steps 1 --warmup 0 --spp-per-step 32 > /dev/null 2>&1 || exit 1
$NCU -k regex:"k_trace|k_shade" -c 6 -o gpurun_out/prof_r2${T}_c3 -f $B --steps 1 --warmup 0 --spp-per-step 32 > gpurun_out/r2${T}_ncu_f.log 2>&1
ncu -i gpurun_out/prof_r2${T}_c3.ncu-rep --page raw --csv > gpurun_out/prof_r2${T}_c3.csv 2>/dev/null
rm -f gpurun_out/prof_r2${T}_c3.ncu-rep
tail -c 400 gpurun_out/r2${T}_bench.json
