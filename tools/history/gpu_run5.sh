#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/configs.jsonl
timeout 300 python tests/dev/first_light.py dec > gpurun_out/first_light_dec.log 2>&1
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q > gpurun_out/t_attn_all.log 2>&1
timeout 900 python tools/bench_configs.py --only cfg3d,cfg3p,cfg3mix,pre1k,pre4k,pre8k,pre_mha2k,pre_mha8k,cfg4_2k,cfg4_4k,cfg4_shard8 > gpurun_out/configs.log 2>&1
for sp in 1 2 4 8 16; do
  HI_DEC_SPLITS=$sp timeout 300 python tools/bench_configs.py --only cfg3d,cfg4_shard8 2>&1 | sed "s/^/dsp=$sp /" >> gpurun_out/configs_sweep_dec.log
done
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_tc -s 1 -c 1 -o gpurun_out/prof_tc2 -f python tools/bench_configs.py --only pre_mha2k > gpurun_out/ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_decode_tc -s 1 -c 1 -o gpurun_out/prof_dec -f python tools/bench_configs.py --only cfg4_4k > gpurun_out/ncu_dec.log 2>&1
tail -n 3 gpurun_out/t_attn_all.log gpurun_out/first_light_dec.log
