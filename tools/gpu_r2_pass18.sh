#!/bin/bash
for w in 1.0 0.75; do for p in 0.3 0.4 0.55 0.7; do
HI_PAIR_WHOLE_FRAC=$w HI_PAIR_PIECE_FRAC=$p REPS=9 python tools/bench_prefill.py 2>/dev/null | cut -c1-260
done; done
