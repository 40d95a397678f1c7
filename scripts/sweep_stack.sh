#!/bin/bash
# traversal stack placement again, final build (window mode, k_trace L1-data-pipe bound on bounce launches), B200
q() { bash scripts/quick_bench.sh 2>&1 | head -2 | tr '\n' '|'; echo; }
echo -n "local stack (shipped): "; q
for v in ss2 ss4 ss8; do echo -n "$v: "; PTB200_LIB=build/variants/libptb200_$v.so q; done
