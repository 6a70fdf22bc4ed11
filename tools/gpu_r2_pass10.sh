#!/bin/bash
show='
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    print(d["case"], {k: (round(v["ms"], 4), round(v["tc_frac"], 3)) for k, v in d.items() if isinstance(v, dict) and "ms" in v})
'
for p in 0 1 2; do echo "== head_dim 64 poly $p"; HI_PAIR_POLY=$p HEAD_DIM=64 python tools/bench_configs.py --only pre1k,pre8k,cfg3p 2>/dev/null | python -c "$show"; done
for p in 0 1 2; do echo "== head_dim 128 poly $p"; HI_PAIR_POLY=$p python tools/bench_configs.py --only pre1k,pre8k,cfg3p 2>/dev/null | python -c "$show"; done
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -k "head_dims_below or unsupported or golden" 2>&1 | tail -2
