#!/bin/bash
# Round 2: ncu evidence.  (1) launch list of the bench command (cold-cache, serialised per-launch times: shares, not absolutes);
# (2) --set full of the FINAL pair kernel on pre8k, cfg3mix and pre1k; (3) --set full of the decode stream kernel (dram bytes per launch).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_launches_bench.csv)"
for c in pre8k cfg3mix pre1k; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_pair -s 4 -c 1 -f -o gpurun_out/r02_pair_$c python tools/bench_configs.py --only $c > gpurun_out/r02_ncu_pair_$c.log 2>&1; echo "$c rc=$?"; ls -la gpurun_out/r02_pair_$c.ncu-rep 2>/dev/null | awk '{print $5}'
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_stream -s 3 -c 1 -f -o gpurun_out/r02_stream_cfg2 python tools/bench_configs.py --only cfg2 > gpurun_out/r02_ncu_stream.log 2>&1; echo "stream rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_decode_tc -s 3 -c 1 -f -o gpurun_out/r02_dectc_cfg4 python tools/bench_configs.py --only cfg4_2k > gpurun_out/r02_ncu_dectc.log 2>&1; echo "dectc rc=$?"
