#!/bin/bash
# GPU pass 3 (re-entry confirmation): gpu tests, smoke, bench (both arms), every config with the flashinfer comparator, ncu launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
rm -f gpurun_out/configs.jsonl
timeout 900 python tools/bench_configs.py --flashinfer > gpurun_out/configs.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 5 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json; tail -n 5 gpurun_out/bench.err
