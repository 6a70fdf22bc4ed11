"""Measurement tool (GPU box): SURVEY §8 config 4 — Qwen2-VL-72B shape (64q/8kv heads, d=128) decode, batch 256, sequences
sharded `seq i -> GPU i mod N` over N ranks (STRONG scaling: the batch is fixed, each rank holds 256/N sequences and their pages).

    python tools/bench_cfg4_sharded.py                                                  # N = 1
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/bench_cfg4_sharded.py

No data-path collective: every rank runs the single-GPU layer call (KV append + paged attention) on its own pool.  Timing: CUDA
events on each rank between barriers, MAX over ranks; aggregate tokens/s = 256 / that time.  One JSON line per context length."""
from __future__ import annotations

import json
import os
import statistics
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig  # noqa: E402
from hydrainfer_b200.memory import KVCache  # noqa: E402
from hydrainfer_b200.workloads import make_batch, shard_round_robin  # noqa: E402

HQ, HKV, D, BS, BATCH = 64, 8, 128, 16, 256
HBM_PEAK = 6537.0
try:
    HBM_PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def main() -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lines = []
    for ctx in (2048, 4096):
        mine = shard_round_robin(BATCH, rank, world)
        batch = make_batch([(1, ctx)] * len(mine), HQ, HKV, D, BS, dtype=torch.bfloat16, device=dev, gen_device=dev, seed=rank)
        layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(HQ, HKV, D))
        builder = AttentionParametersBuilder(HQ, HKV, D, BS, dev)
        for req in batch.requests():
            builder.add_request(*req)
        builder.add_kv_cache(KVCache(batch.key_cache, batch.value_cache))
        params = builder.build_attention_parameters()[0]
        q, k, v = batch.query, batch.key, batch.value

        def step():
            return layer(q, k, v, params).o

        for _ in range(5):
            step()
        ts = []
        inner = 20  # back-to-back layer calls per timed region: a single call is shorter than the host's launch path
        for _ in range(10):
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(inner):
                step()
            e.record()
            torch.cuda.synchronize(dev)
            ts.append(s.elapsed_time(e) / inner)
        t = torch.tensor([statistics.median(ts)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        bytes_per_token = 2 * ctx * HKV * D * 2 + 2 * HQ * D * 2
        per_gpu_gbs = len(mine) * bytes_per_token / ms / 1e6
        line = {"case": f"cfg4_ctx{ctx}", "n_gpus": world, "batch_total": BATCH, "seqs_per_gpu": len(mine), "ms_max_over_ranks": round(ms, 5),
                "tokens_per_s_aggregate": round(BATCH / ms * 1e3, 1), "per_gpu_gbs": round(per_gpu_gbs, 1),
                "per_gpu_frac_hbm_measured": round(per_gpu_gbs / HBM_PEAK, 3), "scaling": "strong", "collective": "none"}
        if rank == 0:
            print(json.dumps(line), flush=True)
            lines.append(line)
        del batch, params, builder
        torch.cuda.empty_cache()
    if rank == 0:
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        with open(out / f"cfg4_sharded_n{world}.jsonl", "w") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
