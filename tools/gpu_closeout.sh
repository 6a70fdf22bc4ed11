#!/bin/bash
# GPU pass 12 (round-1 close-out, re-run after the just-in-time claiming change): full gpu suite incl. the fuzz file, smoke, both bench arms, every config beside flashinfer,
# vision shapes, ncu launch list of the bench step and one --set full capture of the pair kernel on the config-3 mixed batch
# and on 1k prefill.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench.json)"; cut -c1-300 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
rm -f gpurun_out/configs.jsonl
timeout 900 python tools/bench_configs.py --flashinfer > gpurun_out/configs_p22.jsonl 2> gpurun_out/configs_p22.err; echo "configs rc=$?"
python tools/summarize_configs.py gpurun_out/configs_p22.jsonl | grep -E "cfg3|pre|cfg2_b"
timeout 300 python tools/bench_vision.py > gpurun_out/vision_p22.jsonl 2> gpurun_out/vision_p22.err; cut -c1-250 gpurun_out/vision_p22.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pass22.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_pair -s 50 -c 1 -f -o gpurun_out/pair_cfg3mix python tools/bench_configs.py --only cfg3mix > gpurun_out/ncu_pair_cfg3mix.log 2>&1; tail -n 2 gpurun_out/ncu_pair_cfg3mix.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_pair -s 30 -c 1 -f -o gpurun_out/pair_pre1k python tools/bench_configs.py --only pre1k > gpurun_out/ncu_pair_pre1k.log 2>&1; tail -n 2 gpurun_out/ncu_pair_pre1k.log | cut -c1-200
timeout 600 python tools/bench_configs.py --graph --only cfg2_b1,cfg2_b4,cfg2_b8,cfg2_b16,cfg2_b32,cfg2,cfg3d,gqa_b4_8k,gqa_b8_2k,gqa_b16_rag,gqa72_b64_rag,cfg4_shard8,cfg4_2k,cfg4_4k,cfg3p,cfg3mix,pre1k,pre8k > gpurun_out/configs_graph_p22.jsonl 2>/dev/null; python tools/summarize_configs.py gpurun_out/configs_graph_p22.jsonl | grep -E "simt|dec |plan" | awk '{print $2, $8, $10}' | tr '\n' ';'; echo
