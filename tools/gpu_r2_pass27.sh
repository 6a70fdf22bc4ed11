for st in 0 2 4 8; do
  if [ $st = 0 ]; then export HI_MIGRATE_BULK=0; else export HI_MIGRATE_BULK=2 HI_MIGRATE_BULK_STAGES=$st; fi
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/p27_$st.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open('gpurun_out/p27_$st.json').read())
print('stages', $st, ' '.join(f"{p['pool'][:5]}/{p['blocks_per_request']}:{p['gbs_per_pair']:.0f}/{p['memcpy_peer_gbs']:.0f}{'' if p['bit_exact'] else '!!'}" for p in d['migrate_sweep']['points']))
PY
done
