#!/bin/bash
# GPU pass 20b: GPU-side time (CUDA graph of 10 calls) of the decode paths on small / ragged batches, and the split sweep again.
echo "== events vs graph, default heuristics"
timeout 300 python tools/bench_configs.py --only cfg2_b1,cfg2_b4,cfg2_b8,cfg2_b16,cfg2,cfg3d,gqa_b4_8k,gqa_b8_2k,gqa_b16_rag,gqa72_b64_rag,cfg4_shard8 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep -E "simt|dec " | awk '{print $2, $8, $10}' | tr '\n' ';'; echo
timeout 300 python tools/bench_configs.py --graph --only cfg2_b1,cfg2_b4,cfg2_b8,cfg2_b16,cfg2,cfg3d,gqa_b4_8k,gqa_b8_2k,gqa_b16_rag,gqa72_b64_rag,cfg4_shard8 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep -E "simt|dec " | awk '{print $2, $8, $10}' | tr '\n' ';'; echo
for s in 1 2 3 4 6 8 12 16 32; do echo "== graph, HI_DEC_SPLITS=$s"; HI_DEC_SPLITS=$s timeout 300 python tools/bench_configs.py --graph --only cfg3d,gqa_b4_8k,gqa_b8_2k,gqa_b16_rag,gqa72_b64_rag,cfg4_shard8 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep "dec " | awk '{print $2, $10}' | tr '\n' ' '; echo; done
