#!/bin/bash
# GPU pass 5: new un-paged vision attention + image gather + migration layers/push first, then the whole gpu suite,
# smoke, the vision bench, the e2e breakdown and the bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vision_attention.py -x -q > gpurun_out/pytest_vision.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_vision.log
tail -n 25 gpurun_out/pytest_vision.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 8 gpurun_out/smoke.log
timeout 300 python tools/bench_vision.py --flash > gpurun_out/vision.jsonl 2> gpurun_out/vision.err; cat gpurun_out/vision.jsonl; tail -n 3 gpurun_out/vision.err
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.txt 2>&1; cat gpurun_out/e2e_breakdown.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
