#!/bin/bash
REPS=9 python tools/bench_prefill.py 2>/dev/null | cut -c1-260
for w in 0.6 0.45; do for p in 0.3 0.4 0.55; do
HI_PAIR_HEAVY_TRIG=0.7 HI_PAIR_WHOLE_FRAC=$w HI_PAIR_PIECE_FRAC=$p REPS=9 python tools/bench_prefill.py 2>/dev/null | cut -c1-300
done; done
