#!/bin/bash
# camera-path issue order: samples of one pixel issued back to back (PTB_SAMPLE_GROUP), B200
for g in ${GROUPS_:-1 4 8 32 64}; do
 echo -n "group=$g: "; PTB_SAMPLE_GROUP=$g bash scripts/quick_bench.sh 2>&1 | head -2 | tr '\n' '|'; echo
done
