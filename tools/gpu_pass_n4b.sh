#!/bin/bash
# 4-GPU pass: bench.py at N=4 (copy-stream e2e leg).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "lines=$(wc -l < gpurun_out/bench_n4.json)"; cut -c1-200 gpurun_out/bench_n4.json; tail -n 2 gpurun_out/bench_n4.err
