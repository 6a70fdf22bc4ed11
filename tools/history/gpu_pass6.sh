#!/bin/bash
# GPU pass 6: CUDA-graph decode runner tests + bench, ncu of the un-paged vision kernel, launch list of the bench step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cuda_graph.py tests/test_gpu_attention.py::test_unsupported_arguments_raise -x -q > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_graph.log
tail -n 25 gpurun_out/pytest_graph.log
timeout 600 python tools/bench_graph.py > gpurun_out/graph.jsonl 2> gpurun_out/graph.err; cat gpurun_out/graph.jsonl; tail -n 5 gpurun_out/graph.err
for c in clip qwen; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_pair -s 3 -c 1 -f -o gpurun_out/vision_$c python tools/ncu_vision_case.py $c > gpurun_out/ncu_vision_$c.log 2>&1
  tail -n 2 gpurun_out/ncu_vision_$c.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pass6.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -n 2 gpurun_out/bench_under_ncu.log | cut -c1-300
