#!/bin/bash
# GPU pass 13: bench.py with the e2e leg's copies on copy streams (double-buffered).
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python bench.py > gpurun_out/bench_p13_$i.json 2> gpurun_out/bench_p13_$i.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_p13_$i.json)"; python -c "
import json; d=json.loads(open('gpurun_out/bench_p13_$i.json').read()); print(d['value'], d['ms_per_step'], d['e2e'])"; tail -n 2 gpurun_out/bench_p13_$i.err; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 1 --steps 20 --warmup 3 2>/dev/null | cut -c1-200
