#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe: the plain command first, exit 0, then the same command under ncu).
# Reports land in gpurun_out/ (scratch); scripts/ncu_summary.py turns them into profiles/r2_*.md here.
#   gpu_profile_r2.sh [tag]   (tag: suffix of the output files, default "final")
set -x
T=${1:-final}
B="python bench.py --no-cpu --no-e2e --no-c5-leg"
NCU="ncu --set full --clock-control none "
# 0. the default bench command, whole (e2e, CPU baseline, extra.c5): the line that goes into profiles/r2_bench_lines.jsonl
python bench.py --steps 3 --warmup 3 > gpurun_out/r2${T}_bench.json 2> gpurun_out/r2${T}_bench.err || exit 1
$B --steps 2 --warmup 1 > gpurun_out/r2${T}_plain.json 2> gpurun_out/r2${T}_plain.err || exit 1
# 1. launch list of the bench command (kernel shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2${T}_launches_c3.csv $B --steps 2 --warmup 1 > gpurun_out/r2${T}_ncu_l.log 2>&1
# 2. full capture: camera (packets) / first-bounce / second-bounce traversal + k_shade (one 32-spp chunk: one pixel per warp)
$B --steps 1 --warmup 0 --spp-per-step 32 > /dev/null 2>&1 || exit 1
$NCU -k regex:"k_trace|k_shade" -c 6 -o gpurun_out/prof_r2${T}_c3 -f $B --steps 1 --warmup 0 --spp-per-step 32 > gpurun_out/r2${T}_ncu_f.log 2>&1
# 3. the fused tail (its launches before the hand-over return at once: keep the first ten, the long one is the real one)
$NCU -k regex:"k_tail" -c 10 -o gpurun_out/prof_r2${T}_tail -f $B --steps 1 --warmup 0 --spp-per-step 32 > gpurun_out/r2${T}_ncu_t.log 2>&1
# 4. rtweekend1 4K MIS: k_shade / k_shadow of the first two iterations
$B --workload rtweekend1 --steps 1 --warmup 0 --spp-per-step 8 > /dev/null 2>&1 || exit 1
$NCU -k regex:"k_shade|k_shadow|k_trace" -c 6 -o gpurun_out/prof_r2${T}_rt1 -f $B --workload rtweekend1 --steps 1 --warmup 0 --spp-per-step 8 > gpurun_out/r2${T}_ncu_r.log 2>&1
# 5. C5 closest hit (10 M triangles, one 16 Mi-ray batch)
$B --workload closest_hit --rays 16777216 --steps 1 --warmup 1 > /dev/null 2>&1 || exit 1
$NCU -k regex:"k_closest_hit_api" --launch-skip 1 -c 1 -o gpurun_out/prof_r2${T}_c5 -f $B --workload closest_hit --rays 16777216 --steps 1 --warmup 1 > gpurun_out/r2${T}_ncu_c5.log 2>&1
# the raw pages as CSV (small) travel back; the reports themselves are dropped (64 MiB return limit)
for r in c3 tail rt1 c5; do
  ncu -i gpurun_out/prof_r2${T}_$r.ncu-rep --page raw --csv > gpurun_out/prof_r2${T}_$r.csv 2>/dev/null
  rm -f gpurun_out/prof_r2${T}_$r.ncu-rep
done
ls -la gpurun_out/prof_r2${T}_*
