#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/dev/peer_copy_probe.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -16
timeout 900 python -m pytest tests/test_gpu_migration.py tests/test_gpu_reference_native.py -m gpu -q 2>&1 | tail -3
