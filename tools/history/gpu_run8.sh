#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_pair -s 3 -c 1 -o gpurun_out/prof_pair8k -f python tools/bench_configs.py --only pre8k > gpurun_out/ncu_pair8k.log 2>&1
tail -n 3 gpurun_out/ncu_pair8k.log | cut -c1-200
