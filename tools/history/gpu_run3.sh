#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/configs.jsonl
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q -x > gpurun_out/t_attn_all.log 2>&1
timeout 900 python tools/bench_configs.py --only cfg2,cfg3d,cfg3mix,cfg4_2k,cfg4_4k,cfg4_shard8 > gpurun_out/configs.log 2>&1
for st in 1 2 3; do for sp in 1 2 4 8; do
  HI_TC_STAGES=$st HI_TC_SPLITS=$sp timeout 300 python tools/bench_configs.py --only cfg3d,cfg4_shard8 2>&1 | sed "s/^/st=$st sp=$sp /" >> gpurun_out/configs_sweep.log
done; done
for st in 1 2 3; do
  HI_TC_STAGES=$st timeout 300 python tools/bench_configs.py --only cfg4_4k,pre4k,cfg3p 2>&1 | sed "s/^/st=$st /" >> gpurun_out/configs_sweep.log
done
timeout 600 python tools/bench_migration.py > gpurun_out/migration.log 2>&1
tail -n 3 gpurun_out/t_attn_all.log
