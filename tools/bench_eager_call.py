"""Measurement tool (GPU box): host cost of ONE eager attention-layer call (KV append + paged attention through the public
layer API) at small batch, where the call is launch-bound: wall time per call over a long back-to-back loop, against the GPU
time of the same call from a CUDA graph.  VERDICT r01 item 6 (38.7 us per eager call at batch 1 through ctypes)."""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig  # noqa: E402
from hydrainfer_b200.memory import KVCache  # noqa: E402
from hydrainfer_b200.workloads import make_batch  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    out = []
    for name, hq, hkv, batch_size, ctx in (("llava_b1", 32, 32, 1, 2048), ("llava_b8", 32, 32, 8, 2048), ("qwen_b1", 28, 4, 1, 2048), ("qwen72_b32", 64, 8, 32, 2048)):
        batch = make_batch([(1, ctx)] * batch_size, hq, hkv, 128, 16, dtype=torch.bfloat16, device=dev, gen_device=dev, seed=0)
        layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(hq, hkv, 128))
        builder = AttentionParametersBuilder(hq, hkv, 128, 16, dev)
        for req in batch.requests():
            builder.add_request(*req)
        builder.add_kv_cache(KVCache(batch.key_cache, batch.value_cache))
        params = builder.build_attention_parameters()[0]
        q, k, v = batch.query, batch.key, batch.value
        for _ in range(20):
            layer(q, k, v, params)
        torch.cuda.synchronize()
        n = 2000
        t0 = time.perf_counter()
        for _ in range(n):
            layer(q, k, v, params)
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        # GPU time of the same call
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            layer(q, k, v, params)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(10):
                    layer(q, k, v, params)
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        line = {"case": name, "batch": batch_size, "ctx": ctx, "heads": [hq, hkv], "eager_issue_us_per_call": t_issue / n * 1e6,
                "eager_wall_us_per_call": t_all / n * 1e6, "gpu_us_per_call_graph": s.elapsed_time(e) / 100 * 1e3}
        print(json.dumps(line), flush=True)
        out.append(line)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    with open(ROOT / "gpurun_out" / "eager_call.jsonl", "a") as f:
        for line in out:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
