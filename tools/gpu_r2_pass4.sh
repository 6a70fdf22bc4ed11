#!/bin/bash
# same-box A/B: round-1 pair kernel (variant library) vs the current one
for i in 1 2; do
HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_r1pair.so python tools/bench_prefill.py 2>/dev/null
HI_PAIR_RING=44 HI_PAIR_TMA_STORE=0 python tools/bench_prefill.py 2>/dev/null
HI_PAIR_RING=44 python tools/bench_prefill.py 2>/dev/null
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
