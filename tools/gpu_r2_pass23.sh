#!/bin/bash
# pass 23: mha_varlen_fwd score options (softcap / window / alibi) against the oracle and the reference's compiled FlashAttention-2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_reference_native.py -m gpu -q -x -k "options or unsupported or fa2" > gpurun_out/p23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p23_pytest.log; grep -n "^FAILED\|^ERROR\|passed\|failed\|rc=\|^E " gpurun_out/p23_pytest.log | tail -20
