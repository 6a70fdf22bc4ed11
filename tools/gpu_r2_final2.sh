#!/bin/bash
# Round 2 close-out (one GPU): full gpu suite, smoke, memcheck of the changed paths, ncu of the final kernels, bench line.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/final_pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log; tail -n 2 gpurun_out/final_smoke.log
bash tools/gpu_r2_sanitizer.sh 2>&1 | tail -4
for c in pre8k cfg3mix pre1k cfg3p; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_pair -s 6 -c 1 -f -o gpurun_out/r02_pair_$c python tools/bench_configs.py --only $c > gpurun_out/r02_ncu_pair_$c.log 2>&1; echo "$c ncu rc=$?"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/final_bench_ref.json
bash tools/gpu_r2_bench.sh 1
