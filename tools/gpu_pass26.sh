#!/bin/bash
# GPU pass 26b: device plan with the warp scan: the plan tests + plan-less timings.
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -n 2
for dp in 1 0; do echo "== HI_PAIR_DEVICE_PLAN=$dp"; HI_PAIR_DEVICE_PLAN=$dp timeout 600 python tools/bench_configs.py --only cfg3p,cfg3mix,pre256,pre1k,pre4k,pre_mha2k 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep -E "pair|auto" | awk '{print $2, $8, $10}' | tr '\n' ';'; echo; done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:p2_plan -s 5 -c 3 python tools/bench_configs.py --only pre256 2>&1 | grep -E "p2_plan|gpu__time" | head -6
