#!/bin/bash
# lane-occupancy statistics of persistent_trace (build/variants/libptb200_lanes.so = -DPTB_LANE_STATS), C3 16 spp, B200
for f in 4 8 16 24 30; do
 echo "fetch=$f:"; PTB200_LIB=build/variants/libptb200_lanes.so PTB_TRACE_FETCH=$f python bench.py --steps 1 --warmup 0 --spp-per-step 16 --no-cpu --no-e2e 2>&1 | grep lane_stats | tail -1
done
