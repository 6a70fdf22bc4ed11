#!/bin/bash
# GPU pass 2: e2e breakdown, pair-kernel timelines (trace build), debug-bit floors, ncu full captures of the three hot kernels.
mkdir -p gpurun_out
python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1
for c in pre1k cfg3p pre256 pre8k; do
  HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_trace.so timeout 200 python tools/pair_trace.py $c --items 8 > gpurun_out/trace_$c.log 2>&1
done
HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_trace.so timeout 200 python tools/pair_trace.py pre1k --items 3 --raw > gpurun_out/trace_pre1k_raw.log 2>&1
rm -f gpurun_out/configs.jsonl
for dbg in 0 1 2 4 3 5 6 7; do
  echo "== HI_PAIR_DEBUG=$dbg" >> gpurun_out/pair_debug.log
  HI_PAIR_DEBUG=$dbg timeout 300 python tools/bench_configs.py --only pre1k,cfg3p,pre8k 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['case'], {k: round(v['ms'], 4) for k, v in d.items() if isinstance(v, dict) and 'ms' in v})
" >> gpurun_out/pair_debug.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_pair -s 3 -c 1 -o gpurun_out/prof_pair_pre8k -f python tools/bench_configs.py --only pre8k > gpurun_out/ncu_pair8k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_pair -s 3 -c 1 -o gpurun_out/prof_pair_pre1k -f python tools/bench_configs.py --only pre1k > gpurun_out/ncu_pair1k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_decode_tc -s 3 -c 1 -o gpurun_out/prof_dec_cfg4 -f python tools/bench_configs.py --only cfg4_4k > gpurun_out/ncu_dec.log 2>&1
tail -n 8 gpurun_out/e2e_breakdown.log
cat gpurun_out/pair_debug.log
tail -n 14 gpurun_out/trace_pre1k.log
