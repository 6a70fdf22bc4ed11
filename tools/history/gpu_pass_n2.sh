#!/bin/bash
# Multi-GPU pass (N = 4 / 8): bench.py weak scaling, config 4 sharded (strong scaling), config 5 migration sweep with pairs / fanout / p2d.
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_cfg4_sharded.py > gpurun_out/cfg4_sharded_n$N.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/bench_migration.py > gpurun_out/migration_n$N.log 2> gpurun_out/migration_n$N.err
cat gpurun_out/bench_n$N.json | cut -c 1-1500; tail -n 3 gpurun_out/bench_n$N.err; tail -n 4 gpurun_out/cfg4_sharded_n$N.log; grep -c bit_exact gpurun_out/migration_n$N.log; grep -c '"bit_exact": false' gpurun_out/migration_n$N.log; grep '"n_blocks": 4096\|"n_blocks": 256' gpurun_out/migration_n$N.log | cut -c 1-260; tail -n 5 gpurun_out/migration_n$N.err
