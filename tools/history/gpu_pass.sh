#!/bin/bash
# One GPU pass: the gpu test suite, the driver bench (both arms), every BASELINE config per kernel path, ncu launch list.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
rm -f gpurun_out/configs.jsonl
timeout 900 python tools/bench_configs.py ${BENCH_CONFIGS_ARGS} > gpurun_out/configs.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/t_gpu.log gpurun_out/bench.log gpurun_out/bench_ref.log
tail -30 gpurun_out/configs.log | cut -c1-600
