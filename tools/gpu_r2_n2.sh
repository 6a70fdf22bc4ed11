#!/bin/bash
# Round 2, 2 GPUs: multi-GPU tests (peer pull / push, cross-process IPC, packed NCCL, second-device launches), smoke peer path, bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_migration.py tests/test_gpu_reference_native.py -m gpu -q > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n2_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed\|skipped" gpurun_out/n2_pytest.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/n2_smoke.log; tail -n 4 gpurun_out/n2_smoke.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/n2_bench.json)"; tail -n 3 gpurun_out/n2_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/n2_bench.json').read())
    print('value', d['value'], 'n', d['n_gpus'], 'e2e', d['e2e']['value'])
    print('cfg4', {k: {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()} for k, v in d['cfg4'].items() if isinstance(v, dict)})
    for p in (d.get('migrate_sweep') or {}).get('points', []):
        print(p['pool'], p['blocks_per_request'], 'GB/s', round(p['gbs_per_pair'], 1), 'memcpyPeer', round(p['memcpy_peer_gbs'], 1), 'frac900', round(p['frac_of_nvlink_900'], 3), p['bit_exact'])
    print('extras_s', d.get('extras_seconds'))
except Exception as e:
    print('bench parse failed', e)
PY
