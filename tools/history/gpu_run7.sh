#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/configs.jsonl
timeout 300 python tests/dev/first_light.py pair > gpurun_out/first_light_pair.log 2>&1
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q -x > gpurun_out/t_attn_all.log 2>&1
timeout 900 python tools/bench_configs.py --only cfg3p,cfg3mix,pre256,pre1k,pre4k,pre8k,pre_mha2k,pre_mha8k > gpurun_out/configs.log 2>&1
tail -n 30 gpurun_out/first_light_pair.log; tail -n 5 gpurun_out/t_attn_all.log
