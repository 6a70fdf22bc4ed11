#!/bin/bash
# Round 2 close-out on one GPU: full gpu suite, smoke, both bench arms.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/final_pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log; tail -n 9 gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/final_bench_ref.json
bash tools/gpu_r2_bench.sh 1
