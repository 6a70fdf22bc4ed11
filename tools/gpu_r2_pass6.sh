#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/p6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p6_pytest.log; tail -n 5 gpurun_out/p6_pytest.log
bash tools/gpu_r2_probe2.sh 2>&1 | head -8
python tools/bench_prefill.py 2>/dev/null
HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_r1pair.so python tools/bench_prefill.py 2>/dev/null
