#!/bin/bash
show='
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    print(d["case"], {k: (round(v["ms"], 4), round(v["hbm_frac"], 3)) for k, v in d.items() if isinstance(v, dict) and "ms" in v})
'
python - <<'PY'
# decode_tc (tokens on the MMA M side) on the MHA headline config: is the TMA-fed kernel faster than the cp.async one at group 1?
import sys
sys.path.insert(0, ".")
import tools.bench_configs as bc, torch, json
lines = []
bc.USE_GRAPH = True
bc.run_case("cfg2_mha_all_paths", [(1, 2048)] * 64, 32, 32, [("simt", 1), ("tc", 2), ("dec", 3)], lines)
bc.run_case("cfg2_b8_all_paths", [(1, 2048)] * 8, 32, 32, [("simt", 1), ("dec", 3)], lines)
PY
bash tools/gpu_r2_bench.sh 1 2>&1 | head -3
