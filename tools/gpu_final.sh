#!/bin/bash
# Final confirmation on the committed tree: full gpu suite, smoke, bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench.json)"; python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
