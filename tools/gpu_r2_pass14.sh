#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_cuda_graph.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_decode_tc -s 3 -c 1 -f -o gpurun_out/r02_dectc_cfg2 python tools/bench_configs.py --only cfg2 > gpurun_out/r02_ncu_dectc_cfg2.log 2>&1; echo "ncu rc=$?"
bash tools/gpu_r2_bench.sh 1 2>&1 | head -3
