#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_vision_attention.py tests/test_gpu_cuda_graph.py -m gpu -q -x > gpurun_out/p17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p17_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/p17_pytest.log | tail -6
for i in 1 2; do
HI_PAIR_HEAVY_SPLIT=0 python tools/bench_prefill.py 2>/dev/null
python tools/bench_prefill.py 2>/dev/null
done
