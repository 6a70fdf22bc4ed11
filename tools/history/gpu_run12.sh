#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu_all.log 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -n 3 gpurun_out/t_gpu_all.log gpurun_out/smoke.log; tail -n 1 gpurun_out/bench.log | cut -c1-300
