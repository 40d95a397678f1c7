#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe: the plain command first, exit 0, then the same command under ncu).
# Reports land in gpurun_out/ (scratch); scripts/ncu_summary.py turns them into profiles/r2_*.md here.
set -x
B="python bench.py --no-cpu --no-e2e --no-c5-leg"
NCU="ncu --set full --clock-control none "
$B --steps 2 --warmup 1 > gpurun_out/r2p_plain.json 2> gpurun_out/r2p_plain.err || exit 1
# 1. launch list of the default bench command (kernel shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_c3.csv $B --steps 2 --warmup 1 > gpurun_out/r2p_ncu_l.log 2>&1
# 2. full capture: camera / first-bounce / second-bounce k_trace + k_shade (one 16-spp chunk)
$B --steps 1 --warmup 0 --spp-per-step 16 > /dev/null 2>&1 || exit 1
$NCU -k regex:"k_trace|k_shade" -c 6 -o gpurun_out/prof_r2_c3 -f $B --steps 1 --warmup 0 --spp-per-step 16 > gpurun_out/r2p_ncu_f.log 2>&1
# 3. the fused tail (its launches before the hand-over return at once: keep the first ten, the long one is the real one)
$NCU -k regex:"k_tail" -c 10 -o gpurun_out/prof_r2_tail -f $B --steps 1 --warmup 0 --spp-per-step 16 > gpurun_out/r2p_ncu_t.log 2>&1
# 4. the compressed 8-wide tree on the same first-bounce launch (why it is not the default)
PTB_BVH=wide $NCU -k regex:"k_trace" -c 3 -o gpurun_out/prof_r2_c3_wide -f $B --steps 1 --warmup 0 --spp-per-step 16 > gpurun_out/r2p_ncu_w.log 2>&1
# 5. rtweekend1 4K MIS: k_shade / k_shadow of the first two iterations
$B --workload rtweekend1 --steps 1 --warmup 0 --spp-per-step 8 > /dev/null 2>&1 || exit 1
$NCU -k regex:"k_shade|k_shadow|k_trace" -c 6 -o gpurun_out/prof_r2_rt1 -f $B --workload rtweekend1 --steps 1 --warmup 0 --spp-per-step 8 > gpurun_out/r2p_ncu_r.log 2>&1
# 6. C5 closest hit (10 M triangles, one 16 Mi-ray batch)
$B --workload closest_hit --rays 16777216 --steps 1 --warmup 1 > /dev/null 2>&1 || exit 1
$NCU -k regex:"k_closest_hit_api" --launch-skip 1 -c 1 -o gpurun_out/prof_r2_c5 -f $B --workload closest_hit --rays 16777216 --steps 1 --warmup 1 > gpurun_out/r2p_ncu_c5.log 2>&1
# the raw pages as CSV (small) travel back; the reports themselves only if they fit the 64 MiB return limit
for r in c3 tail c3_wide rt1 c5; do
  ncu -i gpurun_out/prof_r2_$r.ncu-rep --page raw --csv > gpurun_out/prof_r2_$r.csv 2>/dev/null
done
ls -la gpurun_out/prof_r2_*
du -sm gpurun_out
for r in rt1 tail c5 c3_wide c3; do
  if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_r2_$r.ncu-rep; fi
done
du -sm gpurun_out
