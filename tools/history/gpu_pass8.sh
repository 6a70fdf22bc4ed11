#!/bin/bash
# GPU pass 8: merge-skip for direct tiles (pair kernel) + warp-per-head merge; rope fixture fix; pair-kernel timelines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rope.py tests/test_gpu_attention.py -m gpu -x -q > gpurun_out/pytest_p8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_p8.log
tail -n 5 gpurun_out/pytest_p8.log
rm -f gpurun_out/configs.jsonl
timeout 600 python tools/bench_configs.py --only cfg2_b8,cfg3d,cfg3p,cfg3mix,pre256,pre1k,pre4k,pre_mha2k > gpurun_out/configs_p8.jsonl 2> gpurun_out/configs_p8.err; echo "configs rc=$?"
python tools/summarize_configs.py gpurun_out/configs_p8.jsonl
for s in 1 2 3 4 6; do echo "== HI_TC_SPLITS=$s"; HI_TC_SPLITS=$s timeout 300 python tools/bench_configs.py --only cfg3p,cfg3mix 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep plan; done
for c in pre1k cfg3p pre256; do
  HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_trace.so timeout 300 python tools/pair_trace.py $c --items 12 > gpurun_out/trace_$c.txt 2>&1
  head -n 24 gpurun_out/trace_$c.txt; tail -n 2 gpurun_out/trace_$c.txt
done
