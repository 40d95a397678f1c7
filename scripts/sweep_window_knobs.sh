#!/bin/bash
# window-mode knob sweep (B200): traversal burst / refill threshold, k_shade block size; C3 64 spp and rtweekend1 4K MIS 16 spp
q() { bash scripts/quick_bench.sh 2>&1 | head -2 | tr '\n' '|'; echo; }
for b in 2 4 8; do for f in 4 8 16; do echo -n "burst=$b fetch=$f: "; PTB_TRACE_BURST=$b PTB_TRACE_FETCH=$f q; done; done
for t in 64 128 256; do echo -n "shade_threads=$t: "; PTB_SHADE_THREADS=$t q; done
for bias in 1 2; do echo -n "prim_bias=$bias: "; PTB_TRACE_PRIM_BIAS=$bias q; done
