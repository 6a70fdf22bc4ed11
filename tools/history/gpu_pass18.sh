#!/bin/bash
# GPU pass 18b: split sweep of the swapped-operand decode kernel on small and ragged grouped decode batches (warm-up case first).
for s in "" 1 2 3 4 6 8 12 16 24 32; do echo "== HI_DEC_SPLITS=${s:-auto}"; HI_DEC_SPLITS=$s timeout 300 python tools/bench_configs.py --only cfg2_b32,cfg3d,gqa_b4_8k,gqa_b8_2k,gqa_b16_rag,gqa72_b64_rag 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep "dec " | awk '{print $2, $10}' | tr '\n' ' '; echo; done
