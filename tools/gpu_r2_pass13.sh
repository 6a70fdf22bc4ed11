#!/bin/bash
python - <<'PY'
import sys, json
sys.path.insert(0, ".")
import tools.bench_configs as bc
bc.USE_GRAPH = True
lines = []
for b in (1, 2, 4, 8, 16, 32, 64, 128):
    bc.run_case(f"mha_b{b}_2k", [(1, 2048)] * b, 32, 32, [("simt", 1), ("dec", 3)], lines)
for b, L in ((4, 8192), (16, 8192), (64, 512), (64, 4096)):
    bc.run_case(f"mha_b{b}_{L}", [(1, L)] * b, 32, 32, [("simt", 1), ("dec", 3)], lines)
bc.run_case("mha_rag64", [(1, 100 + 60 * i) for i in range(64)], 32, 32, [("simt", 1), ("dec", 3)], lines)
bc.run_case("g2_b64_2k", [(1, 2048)] * 64, 32, 16, [("simt", 1), ("dec", 3)], lines)
for l in lines:
    print(l["case"], {k: round(v["ms"], 4) for k, v in l.items() if isinstance(v, dict) and "ms" in v})
PY
