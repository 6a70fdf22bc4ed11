#!/bin/bash
mkdir -p gpurun_out
for il in 0 1; do
  HI_PAIR_INTERLEAVE=$il HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_trace.so timeout 300 python tools/pair_trace.py cfg3mix --items 3 --raw > gpurun_out/trace_cfg3mix_il$il.txt 2>&1
  echo "== il=$il"; grep -v "^ *[0-9]* \(K-tma\|V-tma\|mma\|smx\)" gpurun_out/trace_cfg3mix_il$il.txt | head -14 | cut -c1-250
done
