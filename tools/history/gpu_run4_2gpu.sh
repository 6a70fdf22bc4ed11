#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_migration.py -m gpu -q > gpurun_out/t_migr_2gpu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x > gpurun_out/t_attn_all.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/bench_n1.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_migration.py > gpurun_out/migration_n2.log 2>&1
timeout 600 python tools/bench_configs.py --only cfg3p,cfg3mix,pre1k,pre4k,pre8k,pre_mha2k,pre_mha8k,cfg3d,cfg4_4k > gpurun_out/configs.log 2>&1
tail -n 3 gpurun_out/t_migr_2gpu.log gpurun_out/t_attn_all.log gpurun_out/bench_n2.log
