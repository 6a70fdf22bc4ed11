#!/bin/bash
# Round 2, pass 20: pair kernel with an epilogue warpgroup (EW): correctness, then same-box A/B against the three-warpgroup instance
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_cuda_graph.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/p20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p20_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed\|rc=" gpurun_out/p20_pytest.log | tail -8
for i in 1 2; do
HI_PAIR_EPI_WG=0 REPS=9 timeout 300 python tools/bench_prefill.py 2>/dev/null | cut -c1-300
HI_PAIR_EPI_WG=1 REPS=9 timeout 300 python tools/bench_prefill.py 2>/dev/null | cut -c1-300
done
HI_PAIR_EPI_WG=0 timeout 300 python tools/bench_configs.py --only pre256,pre1k,pre4k,pre8k,pre_mha2k,cfg3p,cfg3mix 2>/dev/null | python tools/summarize_configs.py /dev/stdin | awk '{print $2, $8, $10, $18}' | tr '\n' ';'; echo
HI_PAIR_EPI_WG=1 timeout 300 python tools/bench_configs.py --only pre256,pre1k,pre4k,pre8k,pre_mha2k,cfg3p,cfg3mix 2>/dev/null | python tools/summarize_configs.py /dev/stdin | awk '{print $2, $8, $10, $18}' | tr '\n' ';'; echo
