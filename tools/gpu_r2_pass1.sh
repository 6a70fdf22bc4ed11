#!/bin/bash
# Round 2, pass 1: full gpu suite (incl. the oracle/_ref comparisons), smoke, bench line with the extras, eager-call latency.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/p1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/p1_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p1_pytest_gpu.log; tail -n 30 gpurun_out/p1_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/p1_smoke.log; tail -n 8 gpurun_out/p1_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/p1_bench.json 2> gpurun_out/p1_bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/p1_bench.json)"; tail -n 5 gpurun_out/p1_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/p1_bench.json').read())
    print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'serial', d['e2e']['serial_value'], 'launches', d['gpu_launches'], d['clocks'])
    for k, v in (d.get('prefill') or {}).items():
        if isinstance(v, dict): print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ('ms','kernel_ms','tflops','frac','frac_call')})
    print('cfg4', json.dumps(d.get('cfg4'))[:600])
    for p in (d.get('migrate_sweep') or {}).get('points', []):
        print(p['pool'], p['blocks_per_request'], round(p['gbs_per_pair'], 1), 'memcpy', round(p['memcpy_peer_gbs'], 1), p['bit_exact'])
    print('ref_gpu', d.get('reference_gpu_baseline'))
    print('extras_s', d.get('extras_seconds'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout 300 python tools/bench_eager_call.py > gpurun_out/p1_eager.log 2>&1; tail -n 5 gpurun_out/p1_eager.log
