#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_vision_attention.py tests/test_gpu_fuzz.py tests/test_gpu_cuda_graph.py -m gpu -q -x > gpurun_out/p3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p3_pytest.log; tail -n 4 gpurun_out/p3_pytest.log
HI_PAIR_RING=44 timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "pair or golden or mixed" > gpurun_out/p3_pytest44.log 2>&1; tail -n 2 gpurun_out/p3_pytest44.log
for i in 1 2; do
HI_PAIR_RING=44 HI_PAIR_TMA_STORE=0 python tools/bench_prefill.py 2>/dev/null
HI_PAIR_RING=44 python tools/bench_prefill.py 2>/dev/null
HI_PAIR_RING=53 python tools/bench_prefill.py 2>/dev/null
done
