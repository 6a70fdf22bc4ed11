#!/bin/bash
# EW experiment: three-warpgroup kernel vs epilogue-warpgroup variants (spinning / sleeping waits), same box
for v in "" ew ewsleep; do
  if [ -n "$v" ]; then export HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_$v.so; fi
  echo "variant: ${v:-main}"
  REPS=9 timeout 300 python tools/bench_prefill.py 2>/dev/null | cut -c1-300
done
