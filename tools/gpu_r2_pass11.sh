#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/p11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p11_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/p11_pytest.log | tail -12
for pdl in 0 1; do echo "== HI_PDL=$pdl"; HI_PDL=$pdl timeout 300 python tools/bench_eager_call.py 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line); print(d['case'], 'issue', round(d['eager_issue_us_per_call'], 1), 'wall', round(d['eager_wall_us_per_call'], 1), 'gpu(graph)', round(d['gpu_us_per_call_graph'], 2))
"; done
