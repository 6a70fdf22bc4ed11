#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_dropin_reference_layer.py -m gpu -q > gpurun_out/p15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p15_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/p15_pytest.log | tail -12
show='
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    print(d["case"], {k: (round(v["ms"], 4), round(v["tc_frac"], 3)) if "ms" in v else v.get("error", "")[:60] for k, v in d.items() if isinstance(v, dict)})
'
echo "== head_dim 256"; HEAD_DIM=256 python tools/bench_configs.py --only pre1k,pre4k,pre8k,cfg3p,pre_mha2k 2>/dev/null | python -c "$show"
