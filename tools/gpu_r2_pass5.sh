#!/bin/bash
# Round 2, pass 5: split-KV merge folded into the last-arriving CTA (decode kernels): correctness, then timing sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py tests/test_gpu_cuda_graph.py tests/test_gpu_multidevice.py -m gpu -q -x > gpurun_out/p5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p5_pytest.log; tail -n 4 gpurun_out/p5_pytest.log
HI_MERGE_KERNEL=1 timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/p5_pytest_mk.log 2>&1; tail -n 2 gpurun_out/p5_pytest_mk.log
show='
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    print(d["case"], {k: (round(v["ms"], 4), round(v["hbm_frac"], 3)) for k, v in d.items() if isinstance(v, dict) and "ms" in v})
'
CASES=cfg2_b1,cfg2_b4,cfg2_b8,cfg2_b16,cfg2,gqa_b8_2k,gqa_b16_rag,cfg3d,cfg4_shard8_2k,cfg4_shard4_2k,cfg4_shard8,cfg4_2k
echo "== merge kernel (old)"; HI_MERGE_KERNEL=1 python tools/bench_configs.py --graph --only $CASES 2>/dev/null | python -c "$show"
echo "== in-kernel merge"; python tools/bench_configs.py --graph --only $CASES 2>/dev/null | python -c "$show"
for s in 1 2 3 4; do echo "== HI_DEC_SPLITS=$s"; HI_DEC_SPLITS=$s python tools/bench_configs.py --graph --only cfg4_shard8_2k,cfg4_shard4_2k,cfg4_shard8 2>/dev/null | python -c "$show"; done
echo "== flashinfer"; python tools/bench_configs.py --flashinfer --only cfg2_b4,cfg2_b8,gqa_b16_rag,cfg4_shard8_2k 2>/dev/null | python -c "$show"
