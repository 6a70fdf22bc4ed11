#!/bin/bash
# timeline of CTA 0 of the pair kernel (trace variant of the library): per-role hand-off points + per-warp P arrivals
mkdir -p gpurun_out
for c in ${CASES:-pre1k cfg3p}; do
  HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_trace.so timeout 300 python tools/pair_trace.py $c --items ${ITEMS:-3} --raw > gpurun_out/trace_$c.txt 2>&1
  grep -v "^ *[0-9]* \(K-tma\|V-tma\|mma\|smx\)" gpurun_out/trace_$c.txt | head -${HEAD:-80}
done
