"""GPU: seeded random batches through the whole layer (append + attention) on every kernel path that takes the shape.

The hand-picked cases of test_gpu_attention.py follow the reference's test grid; this file walks the space between them:
random head geometries (MHA / GQA / MQA), dtypes, block sizes, fused-qkv strides and ragged mixes of decode rows, square
prefills and chunked prefills, each with and without the host plan, each compared with the CPU oracle:
KV caches bit-exact (torch.equal on the whole pool), outputs within the north-star tolerance of the fp32 recompute
(2e-2 + 1e-2 * |fp32| for 16-bit dtypes).  Sizes are small enough that the oracle runs the 24 cases in a few seconds."""
import random

import pytest
import torch

from hydrainfer_b200.workloads import make_batch
from oracle import paged_kv_oracle as oracle
from test_gpu_attention import DEV, assert_close_to_fp32, oracle_fp32, paths_of

pytestmark = pytest.mark.gpu

HEADS = [(32, 32), (28, 4), (8, 1), (16, 2), (64, 8), (12, 12), (7, 7), (40, 8)]


def _random_case(seed: int):
    rng = random.Random(seed)
    hq, hkv = rng.choice(HEADS)
    d = rng.choice([128, 128, 128, 64])
    bs = rng.choice([16, 16, 16, 8, 32])
    dtype = rng.choice([torch.bfloat16, torch.bfloat16, torch.float16])
    seq_lens = []
    for _ in range(rng.randint(1, 6)):
        kind = rng.random()
        if kind < 0.45:                                  # decode row
            seq_lens.append((1, rng.randint(1, 1500)))
        elif kind < 0.7:                                 # square prefill
            q = rng.randint(2, 300)
            seq_lens.append((q, q))
        else:                                            # chunked prefill: part of the context is already cached
            q = rng.randint(2, 260)
            seq_lens.append((q, q + rng.randint(1, 1200)))
    return dict(hq=hq, hkv=hkv, d=d, bs=bs, dtype=dtype, seq_lens=seq_lens, fused=rng.random() < 0.3)


@pytest.mark.parametrize("seed", range(24))
def test_random_batch_through_the_layer(seed):
    run_case(_random_case(1000 + seed), seed)


def run_case(c: dict, seed: int):
    """One random batch through the layer on every path that takes it (also driven by tests/dev/fuzz_campaign.py with a wider space)."""
    from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
    from hydrainfer_b200.memory import KVCache
    batch = make_batch(c["seq_lens"], c["hq"], c["hkv"], c["d"], c["bs"], dtype=c["dtype"], seed=2000 + seed, fused_qkv=c["fused"])
    t = batch.n_tokens
    # oracle: append, then the fp32 recompute over the appended caches
    kc_ref, vc_ref = batch.clone_caches()
    oracle.set_kv_cache(torch.tensor(batch.new_cache_slots, dtype=torch.int32), batch.key.view(t, c["hkv"], c["d"]),
                        batch.value.view(t, c["hkv"], c["d"]), kc_ref, vc_ref)
    fp32 = oracle_fp32(batch, kc_ref, vc_ref)
    what = f"seed {seed}: heads {c['hq']}/{c['hkv']} d={c['d']} bs={c['bs']} {c['dtype']} seqs={c['seq_lens']} fused={c['fused']}"
    if c["fused"]:
        qkv = torch.cat([batch.query, batch.key, batch.value], dim=1).to(DEV)
        wq, wk = batch.query.shape[1], batch.key.shape[1]
        q_d, k_d, v_d = qkv[:, :wq], qkv[:, wq:wq + wk], qkv[:, wq + wk:]
    else:
        q_d, k_d, v_d = batch.query.to(DEV), batch.key.to(DEV), batch.value.to(DEV)
    for path in paths_of(batch):
        kc, vc = batch.key_cache.to(DEV), batch.value_cache.to(DEV)
        builder = AttentionParametersBuilder(c["hq"], c["hkv"], c["d"], c["bs"], torch.device(DEV))
        for req in batch.requests():
            builder.add_request(*req)
        builder.add_kv_cache(KVCache(kc, vc))
        params = builder.build_attention_parameters()[0]
        layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(c["hq"], c["hkv"], c["d"]))
        layer.handler.path = path
        out = layer(q_d, k_d, v_d, params).o
        torch.cuda.synchronize()
        assert torch.equal(kc.cpu(), kc_ref) and torch.equal(vc.cpu(), vc_ref), f"{what} path={path}: KV append is not bit-exact"
        assert_close_to_fp32(out, fp32, c["dtype"], f"{what} path={path}")
