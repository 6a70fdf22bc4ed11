"""Dev tool (GPU box): a 32-layer decode step of the hot path (per layer: KV append + paged attention on that layer's cache), eager
launches vs CUDA-graph replay over static metadata (hydrainfer_b200/model_runner).  Wall time per step including the
per-step metadata build + upload and a host read of the result; JSON lines."""
import json
import statistics
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig  # noqa: E402
from hydrainfer_b200.memory import KVCache  # noqa: E402
from hydrainfer_b200.model_runner import CudaGraphModelRunner  # noqa: E402

dev = torch.device("cuda:0")
HQ, HKV, D, BS, LAYERS = 32, 32, 128, 16, 32


def run(batch, ctx):
    n_blocks = batch * ((ctx + BS - 1) // BS) + 8
    # one physical pool shared by the "layers" keeps the footprint small; each layer call still streams the full KV of the batch
    kc = torch.randn(n_blocks, BS, HKV, D, device=dev).to(torch.bfloat16)
    vc = torch.randn(n_blocks, BS, HKV, D, device=dev).to(torch.bfloat16)
    caches = [KVCache(kc, vc) for _ in range(LAYERS)]
    layers = [CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(HQ, HKV, D)) for _ in range(LAYERS)]
    qw, kw = HQ * D, HKV * D

    def model(hidden, pos, params):
        q, k, v = hidden[:, :qw], hidden[:, qw:qw + kw], hidden[:, qw + kw:]
        o = None
        for layer, p in zip(layers, params):
            o = layer(q, k, v, p).o
        return o

    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(n_blocks, generator=g).tolist()
    nb = (ctx + BS - 1) // BS
    reqs = []
    for b in range(batch):
        table = perm[b * nb:(b + 1) * nb]
        reqs.append((1, ctx, [table[(ctx - 1) // BS] * BS + (ctx - 1) % BS], table))
    hidden = torch.randn(batch, qw + 2 * kw, device=dev).to(torch.bfloat16)
    pos = torch.full((batch,), ctx - 1, dtype=torch.int32, device=dev)
    out_host = torch.empty((batch, qw), dtype=torch.bfloat16).pin_memory()
    runner = CudaGraphModelRunner(model, torch.bfloat16, dev, BS, qw + 2 * kw, qw, caches, HQ, HKV, D,
                                  cuda_graph_max_batch_size=batch, cuda_graph_max_seq_len=ctx, batch_sizes=[batch])

    def builder():
        b = AttentionParametersBuilder(HQ, HKV, D, BS, dev)
        for r in reqs:
            b.add_request(*r)
        return b

    def eager_step():
        b = builder()
        for c in caches:
            b.add_kv_cache(c)
        out_host.copy_(model(hidden, pos, b.build_attention_parameters()), non_blocking=True)
        torch.cuda.synchronize()

    def graph_step():
        out_host.copy_(runner(hidden, pos, builder()), non_blocking=True)
        torch.cuda.synchronize()

    res = {}
    for name, fn in (("eager", eager_step), ("graph", graph_step)):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        res[name] = statistics.median(ts)
    assert runner.replays >= 35
    print(json.dumps({"case": f"decode step, {LAYERS} layers, batch {batch}, ctx {ctx}, 32/32 heads d128 bf16", "eager_ms_per_step": round(res["eager"], 4),
                      "graph_ms_per_step": round(res["graph"], 4), "eager_us_per_layer": round(res["eager"] * 1e3 / LAYERS, 2),
                      "graph_us_per_layer": round(res["graph"] * 1e3 / LAYERS, 2), "speedup": round(res["eager"] / res["graph"], 2)}), flush=True)


if __name__ == "__main__":
    for batch, ctx in ((1, 2048), (8, 2048), (64, 512), (64, 2048)):
        run(batch, ctx)
