#!/bin/bash
for i in 1 2; do
HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_r1pair.so python tools/bench_prefill.py 2>/dev/null
HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_nostage.so python tools/bench_prefill.py 2>/dev/null
python tools/bench_prefill.py 2>/dev/null
done
