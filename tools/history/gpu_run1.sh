#!/bin/bash
# First GPU pass: first-light per kernel path, the gpu test suite per file, a short bench, and ncu captures.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > gpurun_out/env.txt 2>&1
timeout 1000 python tests/dev/first_light.py > gpurun_out/first_light.log 2>&1
timeout 600 python -m pytest tests/test_gpu_append.py -m gpu -q > gpurun_out/t_append.log 2>&1
timeout 600 python -m pytest tests/test_gpu_migration.py -m gpu -q > gpurun_out/t_migr.log 2>&1
HI_TEST_SKIP_TC=1 timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q > gpurun_out/t_attn_simt.log 2>&1
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q > gpurun_out/t_attn_all.log 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_simt -s 2 -c 2 -o gpurun_out/prof_simt -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/*.log
