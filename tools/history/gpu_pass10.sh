#!/bin/bash
# GPU pass 10b: small-batch decode — larger minimum split-KV chunks.
mkdir -p gpurun_out
for m in 16 24 32 48 64; do
  echo "== HI_SIMT_MIN_CHUNK_TILES=$m"
  HI_SIMT_MIN_CHUNK_TILES=$m timeout 300 python tools/bench_configs.py --only cfg2_b1,cfg2_b4,cfg2_b8,cfg2_b16 2>/dev/null | python tools/summarize_configs.py /dev/stdin | grep -E "simt"
done
