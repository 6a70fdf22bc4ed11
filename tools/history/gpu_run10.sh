#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/configs_dbg.log
for dbg in 1 2 3 5 6 7; do
  HI_PAIR_DEBUG=$dbg timeout 600 python tools/bench_configs.py --only pre8k 2>&1 | sed "s/^/dbg=$dbg /" >> gpurun_out/configs_dbg.log
done
