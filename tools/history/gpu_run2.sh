#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/configs.jsonl
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q > gpurun_out/t_attn_all.log 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_stream -s 2 -c 1 -o gpurun_out/prof_stream -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_stream.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_tc -s 1 -c 1 -o gpurun_out/prof_tc -f python tools/bench_configs.py --only pre_mha2k > gpurun_out/ncu_tc.log 2>&1
timeout 900 python tools/bench_configs.py --only cfg2,cfg3d,cfg3p,pre4k,cfg4_4k --flashinfer > gpurun_out/configs_fi.log 2>&1
tail -n 3 gpurun_out/t_attn_all.log gpurun_out/bench.log
