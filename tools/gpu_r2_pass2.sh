#!/bin/bash
# Round 2, pass 2: TMA-store epilogue of the pair kernel: correctness (attention + vision + fuzz suites) and A/B timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_vision_attention.py tests/test_gpu_fuzz.py tests/test_gpu_cuda_graph.py -m gpu -q -x > gpurun_out/p2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p2_pytest.log; tail -n 6 gpurun_out/p2_pytest.log
for tma in 0 1; do
  echo "== HI_PAIR_TMA_STORE=$tma"
  HI_PAIR_TMA_STORE=$tma timeout 600 python tools/bench_configs.py --only pre256,pre1k,pre4k,pre8k,cfg3p,cfg3mix,pre_mha2k 2>gpurun_out/p2_cfg_$tma.err | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    print(d['case'], {k: (round(v['ms'], 4), round(v['tc_frac'], 3)) for k, v in d.items() if isinstance(v, dict) and 'ms' in v})
"
done
