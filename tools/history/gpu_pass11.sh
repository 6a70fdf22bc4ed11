#!/bin/bash
# GPU pass 11: MMA warps decode the next work item during their last two steps; same-box A/B against the previous build.
# previous build (variant "base").
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_vision_attention.py -m gpu -x -q > gpurun_out/pytest_p11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_p11.log
tail -n 4 gpurun_out/pytest_p11.log
for rep in 1 2; do
for lib in hydrainfer_b200/lib/libhi_b200_base.so ""; do
  echo "=== rep $rep lib=${lib:-product}"
  HI_B200_LIB=$lib timeout 600 python tools/bench_configs.py --only cfg3p,cfg3mix,pre256,pre1k,pre4k,pre8k,pre_mha2k 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep "pair+plan"
done
done
