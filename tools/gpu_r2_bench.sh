#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_n$N.err
fi
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$N.json').read())
print('value', round(d['value']), 'frac', d['roofline']['frac'], 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['clocks'])
for k, v in (d.get('prefill') or {}).items():
    if isinstance(v, dict): print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ('ms','kernel_ms','tflops','frac','frac_call')})
print('cfg4', {k: {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ('seqs_per_gpu','ms','eager_ms','tokens_per_s','hbm_frac')} for k, v in d['cfg4'].items() if isinstance(v, dict)})
for p in (d.get('migrate_sweep') or {}).get('points', []):
    print(p['pool'], p['blocks_per_request'], 'GB/s', round(p['gbs_per_pair'], 1), 'memcpy', round(p['memcpy_peer_gbs'], 1), p['bit_exact'])
for cap, v in d['migrate_under_decode']['caps'].items():
    print('cap', cap, {a: round(b, 3) if isinstance(b, float) else b for a, b in v.items() if a != 'note'})
print('ref_gpu', d.get('reference_gpu_baseline'))
print('extras_s', d.get('extras_seconds'))
PY
