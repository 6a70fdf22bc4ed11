#!/bin/bash
# GPU pass 14: merge kernel with one CTA per row walking the heads; split variants of config 3.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_cuda_graph.py -m gpu -x -q > gpurun_out/pytest_p14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_p14.log
tail -n 4 gpurun_out/pytest_p14.log
timeout 600 python tools/bench_configs.py --only cfg2_b1,cfg2_b8,cfg2,cfg3d,cfg3p,cfg3mix,cfg4_shard8 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep -E "simt|dec |plan"
for s in 1 2 3; do echo "== HI_TC_SPLITS=$s"; HI_TC_SPLITS=$s timeout 300 python tools/bench_configs.py --only cfg3p,cfg3mix 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep plan; done
