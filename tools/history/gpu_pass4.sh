#!/bin/bash
# GPU pass 4: full gpu suite incl. the rotary/append fusion, rope bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_rope.py > gpurun_out/rope.log 2>&1
tail -n 15 gpurun_out/pytest_gpu.log; cat gpurun_out/rope.log
