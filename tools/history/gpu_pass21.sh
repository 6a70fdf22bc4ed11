#!/bin/bash
# GPU pass 21: (a) the decode kernel's new split rule, graph-timed; (b) the cp.async split-KV kernel's minimum chunk on small
# batches, graph-timed (the event-timed sweep of pass 10 measured the host for these sizes).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_cuda_graph.py -m gpu -x -q 2>&1 | tail -n 2
echo "== decode kernel, default rule (graph)"
timeout 300 python tools/bench_configs.py --graph --only cfg3d,gqa_b4_8k,gqa_b8_2k,gqa_b16_rag,gqa72_b64_rag,cfg4_shard8,cfg4_2k 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep "dec " | awk '{print $2, $10}' | tr '\n' ' '; echo
for m in 8 16 24 32 48 64; do echo "== graph, HI_SIMT_MIN_CHUNK_TILES=$m"; HI_SIMT_MIN_CHUNK_TILES=$m timeout 300 python tools/bench_configs.py --graph --only cfg2_b1,cfg2_b4,cfg2_b8,cfg2_b16,cfg2_b32 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep "simt" | awk '{print $2, $10}' | tr '\n' ' '; echo; done
