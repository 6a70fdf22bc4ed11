"""Dev tool (GPU box): where the end-to-end decode step of bench.py spends its wall time (host metadata build, H2D, layer call,
D2H), each phase timed alone with a synchronize on both sides, next to the whole step."""
import statistics
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig  # noqa: E402
from hydrainfer_b200.memory import KVCache  # noqa: E402
from hydrainfer_b200.workloads import make_batch  # noqa: E402

HQ, HKV, D, BS, BATCH, CTX = 32, 32, 128, 16, 64, 2048
dev = torch.device("cuda:0")
batch = make_batch([(1, CTX)] * BATCH, HQ, HKV, D, BS, dtype=torch.bfloat16, device=dev, gen_device=dev, seed=0)
kv_cache = KVCache(batch.key_cache, batch.value_cache)
layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(HQ, HKV, D))
requests = batch.requests()


def build_params():
    builder = AttentionParametersBuilder(HQ, HKV, D, BS, dev)
    for req in requests:
        builder.add_request(*req)
    builder.add_kv_cache(kv_cache)
    return builder.build_attention_parameters()[0]


q_host, k_host, v_host = (t.cpu().pin_memory() for t in (batch.query, batch.key, batch.value))
o_host = torch.empty((BATCH, HQ * D), dtype=torch.bfloat16).pin_memory()
params = build_params()
state = {}


def timed(fn, n=200):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    return statistics.median(ts)


def host_only():
    builder = AttentionParametersBuilder(HQ, HKV, D, BS, dev)
    for req in requests:
        builder.add_request(*req)


def h2d():
    state["q"] = q_host.to(dev, non_blocking=True)
    state["k"] = k_host.to(dev, non_blocking=True)
    state["v"] = v_host.to(dev, non_blocking=True)


def layer_call():
    state["o"] = layer(batch.query, batch.key, batch.value, params).o


def d2h():
    o_host.copy_(state["o"], non_blocking=True)


def e2e():
    q = q_host.to(dev, non_blocking=True)
    k = k_host.to(dev, non_blocking=True)
    v = v_host.to(dev, non_blocking=True)
    p = build_params()
    o = layer(q, k, v, p).o
    o_host.copy_(o, non_blocking=True)


print(f"add_request x{BATCH} (host only)     {timed(host_only):8.1f} us")
print(f"build_params (host + upload)      {timed(build_params):8.1f} us")
print(f"H2D q,k,v                         {timed(h2d):8.1f} us")
print(f"layer call (append + attention)   {timed(layer_call):8.1f} us")
print(f"D2H o                             {timed(d2h):8.1f} us")
print(f"whole step                        {timed(e2e):8.1f} us")
