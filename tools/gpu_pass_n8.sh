#!/bin/bash
# 8-GPU pass, bounded: bench.py weak scaling at N=8 and the config-5 migration sweep (pairs / push / fanout / p2d) on two
# geometries and three request sizes.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
HI_MIG_GEOMS=llava7b,qwen2vl7b HI_MIG_SIZES=16,256,1024 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/bench_migration.py > gpurun_out/migration_n$N.log 2> gpurun_out/migration_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_cfg4_sharded.py > gpurun_out/cfg4_sharded_n$N.log 2>&1
cut -c 1-1800 gpurun_out/bench_n$N.json; tail -n 3 gpurun_out/bench_n$N.err; tail -n 4 gpurun_out/cfg4_sharded_n$N.log | cut -c1-400; grep -c bit_exact gpurun_out/migration_n$N.log; grep -c '"bit_exact": false' gpurun_out/migration_n$N.log; grep '"n_blocks": 1024\|"n_blocks": 256' gpurun_out/migration_n$N.log | cut -c 1-230; tail -n 5 gpurun_out/migration_n$N.err
