#!/bin/bash
# Multi-GPU pass: bench at N ranks (weak scaling over sequences + pair-wise migration), the config-5 migration sweep, and the gpu tests (which
# include the 2-GPU migration cases).  Usage: bash tools/gpu_pass_n.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/bench_migration.py > gpurun_out/migration_n$N.jsonl 2> gpurun_out/migration_n$N.err
timeout 600 python -m pytest tests/test_gpu_migration.py -m gpu -x -q > gpurun_out/pytest_migration_n$N.log 2>&1
cat gpurun_out/bench_n$N.json; tail -n 3 gpurun_out/bench_n$N.err; tail -n 30 gpurun_out/migration_n$N.jsonl | cut -c 1-400; tail -n 3 gpurun_out/pytest_migration_n$N.log
