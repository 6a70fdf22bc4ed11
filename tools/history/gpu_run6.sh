#!/bin/bash
# State check after container re-creation: full gpu suite, every BASELINE config through each path (+flashinfer comparator),
# driver bench (both arms), migration sweep on one GPU, launch list.
mkdir -p gpurun_out; rm -f gpurun_out/configs.jsonl
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu_all.log 2>&1
timeout 900 python tools/bench_configs.py --flashinfer > gpurun_out/configs.log 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 python tools/bench_migration.py > gpurun_out/migration.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
tail -n 3 gpurun_out/t_gpu_all.log gpurun_out/bench.log
