#!/bin/bash
T="tests/test_gpu_multidevice.py::test_two_threads_on_two_streams_do_not_share_scratch"
for env in "A=1" "HI_MERGE_KERNEL=1" "HI_PAIR_TMA_STORE=0" "HI_MERGE_KERNEL=1 HI_PAIR_TMA_STORE=0" "HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_r1pair.so" "HI_PAIR_STATIC=1" "HI_PAIR_DEVICE_PLAN=0"; do
  echo "== $env"; env $env timeout 300 python -m pytest "$T" -m gpu -q -x 2>&1 | tail -n 3 | head -2
done
