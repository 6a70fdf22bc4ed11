#!/bin/bash
# pass 25 (2 GPUs): bulk-copy migration, shared-memory stages per CTA 2 / 3 / 4 / 8
mkdir -p gpurun_out
for st in 2 3 4 8; do
  HI_MIGRATE_BULK_STAGES=$st timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/p25_bench_st$st.json 2> gpurun_out/p25_bench_st$st.err; echo "stages=$st rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/p25_bench_st$st.json').read())
print(' '.join(f"{p['pool'][:5]}/{p['blocks_per_request']}:{p['gbs_per_pair']:.0f}{'' if p['bit_exact'] else '!!'}" for p in d['migrate_sweep']['points']))
for cap, v in list(d['migrate_under_decode']['caps'].items())[:1]:
    print('   under decode', {a: round(b, 3) if isinstance(b, float) else b for a, b in v.items() if a != 'note'})
PY
done
