#!/bin/bash
# window mode: launch-bounds variants (build/variants/libptb200_{tb4,tb6,sb3,sb4}.so) and the tiled pixel order, B200
q() { bash scripts/quick_bench.sh 2>&1 | head -2 | tr '\n' '|'; echo; }
echo -n "default: "; q
for v in tb4 tb6 sb3 sb4; do echo -n "$v: "; PTB200_LIB=build/variants/libptb200_$v.so q; done
echo -n "PTB_TILES=1: "; PTB_TILES=1 q
