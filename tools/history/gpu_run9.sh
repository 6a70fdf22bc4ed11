#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/configs.jsonl
for pf in 0 1 2; do
  HI_PAIR_POLY=$pf timeout 300 python tests/dev/first_light.py pair 2>&1 | sed "s/^/pf=$pf /" >> gpurun_out/first_light_pair.log
  HI_PAIR_POLY=$pf timeout 600 python tools/bench_configs.py --only cfg3p,pre1k,pre4k,pre8k,pre_mha2k,pre_mha8k 2>&1 | sed "s/^/pf=$pf /" >> gpurun_out/configs_poly.log
done
grep -c "bad 0.0000" gpurun_out/first_light_pair.log
