#!/bin/bash
q() { bash scripts/quick_bench.sh 2>&1 | head -2 | tr '\n' '|'; echo; }
for v in sb4 sb5 sb6 tb6sb4 tb6sb5; do echo -n "$v: "; PTB200_LIB=build/variants/libptb200_$v.so q; echo -n "$v tiles: "; PTB_TILES=1 PTB200_LIB=build/variants/libptb200_$v.so q; done
