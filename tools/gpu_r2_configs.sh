#!/bin/bash
# Round 2, final tree: every BASELINE config and the side cases through every kernel path with flashinfer beside it (event-timed per call
# with an L2 flush, then as a CUDA graph of 10 calls), head_dim 64 / 96 prefills, the vision shapes.
mkdir -p gpurun_out
rm -f gpurun_out/configs.jsonl
timeout 1200 python tools/bench_configs.py --flashinfer > gpurun_out/r02_configs_final.jsonl 2> gpurun_out/r02_configs_final.err; echo "configs rc=$?"
python tools/summarize_configs.py gpurun_out/r02_configs_final.jsonl | grep -E "cfg2 |cfg3|pre|cfg4_2k|cfg4_4k" | cut -c1-120
timeout 600 python tools/bench_configs.py --graph > gpurun_out/r02_configs_graph_final.jsonl 2> gpurun_out/r02_configs_graph_final.err; echo "graph rc=$?"
python tools/summarize_configs.py gpurun_out/r02_configs_graph_final.jsonl | grep -E "cfg2|cfg3d|gqa|cfg4" | awk '{print $2, $8, $10}' | tr '\n' ';'; echo
HEAD_DIM=64 timeout 300 python tools/bench_configs.py --only pre1k,pre8k,cfg3p,cfg3mix > gpurun_out/r02_configs_d64.jsonl 2>/dev/null; python tools/summarize_configs.py gpurun_out/r02_configs_d64.jsonl | awk '{print $2, $8, $10, $18}' | tr '\n' ';'; echo
HEAD_DIM=96 timeout 300 python tools/bench_configs.py --only pre1k,pre8k,cfg3p,cfg3mix > gpurun_out/r02_configs_d96.jsonl 2>/dev/null; python tools/summarize_configs.py gpurun_out/r02_configs_d96.jsonl | awk '{print $2, $8, $10, $18}' | tr '\n' ';'; echo
timeout 300 python tools/bench_vision.py > gpurun_out/r02_vision.jsonl 2> gpurun_out/r02_vision.err; cut -c1-220 gpurun_out/r02_vision.jsonl
