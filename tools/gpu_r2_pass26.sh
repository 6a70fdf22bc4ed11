#!/bin/bash
# pass 26 (2 GPUs): does the bulk-copy kernel also win on SMALL requests?  forced for every launch, 2 and 8 stages
mkdir -p gpurun_out
for st in 2 8; do
  HI_MIGRATE_BULK=2 HI_MIGRATE_BULK_STAGES=$st timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/p26_bench_st$st.json 2> gpurun_out/p26_bench_st$st.err; echo "forced bulk, stages=$st rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/p26_bench_st$st.json').read())
print(' '.join(f"{p['pool'][:5]}/{p['blocks_per_request']}:{p['gbs_per_pair']:.0f}{'' if p['bit_exact'] else '!!'}" for p in d['migrate_sweep']['points']))
PY
done
