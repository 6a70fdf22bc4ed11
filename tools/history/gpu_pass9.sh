#!/bin/bash
# GPU pass 9: pair kernel = index ring + per-row epilogue (the staged-epilogue / decoded-item-ring experiments lost on long
# sequences), direct tiles skip the merge, split chunk = 3/4 of an SM's share.  Full suite + bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py --flashinfer > gpurun_out/configs_p9.jsonl 2> gpurun_out/configs_p9.err; echo "configs rc=$?"
python tools/summarize_configs.py gpurun_out/configs_p9.jsonl | grep -E "cfg3|pre"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json
