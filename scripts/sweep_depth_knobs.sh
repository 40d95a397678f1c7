#!/bin/bash
# per-depth k_trace rate (window mode) under different scheduling knobs: are camera rays and bounce rays tuned alike? (B200)
for kn in "8 8" "4 8" "16 8" "8 16" "8 4" "16 16" "32 8"; do set -- $kn
 echo "== burst=$1 fetch=$2"; PTB_TRACE_BURST=$1 PTB_TRACE_FETCH=$2 SPP=16 python scripts/coherence_probe.py 2>&1 | cut -c1-130 | sed -n 1,3p
done
