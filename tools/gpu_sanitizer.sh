#!/bin/bash
# compute-sanitizer over a subset of the gpu tests: memcheck (out-of-bounds / misaligned accesses) on the append, rope,
# migration and attention paths, racecheck (shared-memory hazards) on the CUDA-core decode kernel.
mkdir -p gpurun_out
export HI_TEST_SKIP_TC=${HI_TEST_SKIP_TC:-0}
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_append.py tests/test_rope.py "tests/test_gpu_migration.py::test_migrate_blocks_same_device" tests/test_gpu_attention.py::test_ragged_decode tests/test_gpu_attention.py::test_chunked_prefill_shapes tests/test_gpu_attention.py::test_golden_layer_forward "tests/test_gpu_fuzz.py" tests/test_gpu_vision_attention.py::test_head_dims -m gpu -x -q > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -n 6 gpurun_out/sanitizer_memcheck.log | cut -c1-200; grep -c "Invalid\|misaligned\|out of bounds" gpurun_out/sanitizer_memcheck.log
HI_TEST_SKIP_TC=1 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_attention.py::test_ragged_decode tests/test_gpu_attention.py::test_golden_layer_forward -m gpu -x -q > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -n 6 gpurun_out/sanitizer_racecheck.log | cut -c1-200
