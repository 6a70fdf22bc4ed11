#!/bin/bash
# pass 24 (2 GPUs): bulk-copy (TMA) migration kernel: bit-exactness tests, then the migration sweep and the under-decode leg with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_migration.py tests/test_gpu_multidevice.py tests/test_gpu_reference_native.py -m gpu -q -x -k "migrat or peer or push or ipc or nccl" > gpurun_out/p24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p24_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed\|rc=\|^E " gpurun_out/p24_pytest.log | tail -12
for b in 0 1; do
  HI_MIGRATE_BULK=$b timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/p24_bench_bulk$b.json 2> gpurun_out/p24_bench_bulk$b.err; echo "bulk=$b rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/p24_bench_bulk$b.json').read())
for p in d['migrate_sweep']['points']:
    print(p['pool'], p['blocks_per_request'], 'GB/s', round(p['gbs_per_pair'], 1), 'memcpy', round(p['memcpy_peer_gbs'], 1), p['bit_exact'])
for cap, v in d['migrate_under_decode']['caps'].items():
    print('cap', cap[:20], {a: round(b, 3) if isinstance(b, float) else b for a, b in v.items() if a != 'note'})
PY
done
