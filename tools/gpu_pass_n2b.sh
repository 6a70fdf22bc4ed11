#!/bin/bash
# 2-GPU pass: the multi-GPU tests (peer-device migration, push, cross-process IPC + packed NCCL) and bench.py at N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_migration.py -m gpu -x -q > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n2.log; tail -n 4 gpurun_out/pytest_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "lines=$(wc -l < gpurun_out/bench_n2.json)"; cut -c1-200 gpurun_out/bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref lines=$(wc -l < gpurun_out/bench_ref_n2.json)"; cut -c1-160 gpurun_out/bench_ref_n2.json
