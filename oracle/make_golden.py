"""Generates tests/golden/*.npz by running the REAL reference (imported from /root/reference) on seeded inputs, and
checks oracle/paged_kv_oracle.py against it on the way.  Run in the build container only:

    python oracle/make_golden.py            # writes fixtures, prints the oracle-vs-reference report
    python oracle/make_golden.py --only-rope   # the rotary fixtures only

The reference's Python torch path runs on CPU once four plotting modules are stubbed (hydrainfer/utils/statistic.py:2-7
imports matplotlib / seaborn, which this image lacks; nothing on the path uses them).  Functions exercised:
  CausalGroupedQueryPageAttention.forward            hydrainfer/layer/causal_attention.py:394-406
    -> KVCache.set_kv_cache (Python fallback)        hydrainfer/memory/kv_cache.py:44-50
    -> TorchCausalGroupedQueryPageAttentionHandler   hydrainfer/layer/causal_attention.py:307-374
  AttentionParametersBuilder                         hydrainfer/layer/causal_attention.py:110-210
  TokenCache.set_caches (CPU path)                   hydrainfer/memory/token_cache.py:53-56
  BlockAllocator                                     hydrainfer/memory/block_allocator.py:11-39
  TokenCacheBlockManager.v2p                         hydrainfer/memory/token_cache_manger.py:126-133
  RotaryEmbedding -> TorchRotaryEmbeddingHandler     hydrainfer/layer/rotary_embedding.py:19-99, 136-148
  ROPECausalGroupedQueryPageAttention.forward        hydrainfer/model/model_forward.py:66-86
  MultiHeadAttention -> TorchMultiHeadAttentionHandler          hydrainfer/layer/multihead_attention.py:40-73, 163-176
  QwenMultiHeadAttention -> QwenTorchMultiHeadAttentionHandler  hydrainfer/layer/multihead_attention.py:235-270
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))

from oracle import paged_kv_oracle as oracle  # noqa: E402
from hydrainfer_b200.workloads import make_batch  # noqa: E402


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.colors"].LogNorm = object
    sys.path.insert(0, "/root/reference")
    import hydrainfer.layer.causal_attention as ca
    import hydrainfer.memory as mem
    return ca, mem


def to_np(t: torch.Tensor) -> np.ndarray:
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.contiguous().view(torch.int16).numpy()
    return t.contiguous().numpy()


DTYPE_NAMES = {torch.float32: "float32", torch.float16: "float16", torch.bfloat16: "bfloat16"}

# (name, seq_lens [(q, kv)], Hq, Hkv, d, block_size, dtype, fused_qkv)
ATTENTION_CASES = [
    # cfg-1 flavour: MHA fp32 decode-only, ragged context lengths
    ("mha_fp32_decode", [(1, 1), (1, 16), (1, 17), (1, 100)], 4, 4, 64, 16, torch.float32, False),
    # reference test grid (tests/layer/test_attention.py:42) scaled down: decode + square prefill + chunked prefill
    ("gqa_bf16_mixed", [(1, 37), (15, 15), (21, 50), (1, 130)], 8, 2, 128, 16, torch.bfloat16, False),
    # Qwen2-VL-7B head geometry (28 q heads, 4 kv heads): group size 7, q/k/v are slices of a fused qkv tensor
    ("qwen_gqa7_bf16_fused", [(1, 40), (9, 33), (1, 16)], 28, 4, 128, 16, torch.bfloat16, True),
    ("mha_fp16_prefill", [(33, 33), (5, 70)], 4, 4, 128, 16, torch.float16, False),
    ("mqa_fp16_d256", [(1, 20), (7, 19)], 4, 1, 256, 16, torch.float16, False),
    ("gqa_bf16_d64_bs8", [(1, 9), (12, 30)], 4, 2, 64, 8, torch.bfloat16, False),
]


# (name, n_tokens, Hq, Hkv, d, rotary_dim, theta, interleaved, max_positions, dtype, table dtype) — the reference's grid
# (tests/layer/test_rotary_embedding.py:66-75) scaled down, plus partial rotation and a rotary_dim the vector kernel cannot take
ROPE_CASES = [
    ("f32_half", 16, 8, 2, 128, 128, 100000., False, 4096, torch.float32, torch.float32),
    ("f16_interleaved", 9, 8, 1, 128, 128, 500000., True, 8192, torch.float16, torch.float16),
    ("bf16_partial", 7, 4, 2, 64, 32, 10000., False, 2048, torch.bfloat16, torch.bfloat16),
    ("bf16_f32table_interleaved_rd24", 5, 4, 4, 32, 24, 10000., True, 512, torch.bfloat16, torch.float32),
    ("f16_f32table_half", 3, 8, 8, 128, 128, 1000000., False, 4096, torch.float16, torch.float32),
]


# (name, batch, seq_len, n_heads, head_dim, dtype) — tests/layer/test_multihead_attention-style equal-length image batches:
# CLIP-L (d 64, 577 tokens, scaled down), SigLIP (d 72), a d=128 tower; seq lengths straddle the 64-key step and 128-row tile
MHA_CASES = [
    ("clip_d64_bf16", 3, 150, 4, 64, torch.bfloat16),
    ("siglip_d72_fp16", 2, 81, 3, 72, torch.float16),
    ("d128_bf16", 2, 200, 2, 128, torch.bfloat16),
]
# (name, segment lengths, n_heads, head_dim, dtype) — Qwen2-VL vision tower: packed images of different sizes, d = 80
QWEN_MHA_CASES = [
    ("qwen_d80_bf16", [64, 196, 100, 1, 130], 4, 80, torch.bfloat16),
    ("qwen_d80_fp16", [256, 36], 2, 80, torch.float16),
]


def mha_goldens() -> None:
    """Un-paged vision attention: the reference's two Torch handlers (the CPU end of its handler chains) vs the oracle."""
    import hydrainfer.layer.multihead_attention as ref_mha
    for idx, (name, batch, seq_len, heads, d, dtype) in enumerate(MHA_CASES):
        g = torch.Generator().manual_seed(300 + idx)
        q, k, v = (torch.randn(batch, seq_len, heads * d, generator=g).to(dtype) for _ in range(3))
        # the chain's fused handlers have no device check (flash_attn imports here and would be called with CPU tensors), so the
        # Torch handler — the chain's last link and the reference's own test oracle — is run directly
        module = ref_mha.TorchMultiHeadAttentionHandler(ref_mha.MultiHeadAttentionConfig(n_heads=heads, head_dim=d))
        ref = module(q, k, v, ref_mha.MultiHeadAttentionParameters(return_scores=False)).o
        mine = oracle.multi_head_attention(q, k, v, heads, d)
        assert torch.equal(mine, ref), f"mha {name}: oracle differs from the reference"
        cu = list(range(0, (batch + 1) * seq_len, seq_len))
        fp32 = oracle.varlen_attention_fp32(q.view(-1, heads, d), k.view(-1, heads, d), v.view(-1, heads, d), cu, cu)
        np.savez_compressed(GOLDEN / f"mha_{name}.npz", dtype=DTYPE_NAMES[dtype], geometry=np.array([batch, seq_len, heads, d]), seed=300 + idx,
                            query=to_np(q), key=to_np(k), value=to_np(v), ref_out=to_np(ref))
        print(f"  mha_{name:24s} max |reference(dtype) - fp32 recompute| = {float((ref.float().view(-1, heads * d) - fp32).abs().max()):.3e}")
    for idx, (name, segs, heads, d, dtype) in enumerate(QWEN_MHA_CASES):
        g = torch.Generator().manual_seed(400 + idx)
        total = sum(segs)
        q, k, v = (torch.randn(total, heads, d, generator=g).to(dtype) for _ in range(3))
        cu = torch.tensor([0] + list(np.cumsum(segs)), dtype=torch.int32)
        module = ref_mha.QwenTorchMultiHeadAttentionHandler(ref_mha.MultiHeadAttentionConfig(n_heads=heads, head_dim=d))
        ref = module(q, k, v, total, cu)
        mine = oracle.qwen_multi_head_attention(q, k, v, total, cu.tolist(), d)
        assert torch.equal(mine, ref), f"qwen mha {name}: oracle differs from the reference"
        fp32 = oracle.varlen_attention_fp32(q, k, v, cu.tolist(), cu.tolist())
        np.savez_compressed(GOLDEN / f"mha_{name}.npz", dtype=DTYPE_NAMES[dtype], geometry=np.array([len(segs), total, heads, d]), seed=400 + idx,
                            cu_seqlens=cu.numpy(), query=to_np(q), key=to_np(k), value=to_np(v), ref_out=to_np(ref))
        print(f"  mha_{name:24s} max |reference(dtype) - fp32 recompute| = {float((ref.float() - fp32).abs().max()):.3e}")


def rope_goldens(ca, mem) -> None:
    """Rotary embedding alone and the ROPE attention module, reference vs oracle, frozen as fixtures."""
    import hydrainfer.layer.rotary_embedding as ref_rope
    import hydrainfer.model.model_forward as ref_mf

    for idx, (name, t, hq, hkv, d, rd, theta, inter, max_pos, dtype, table_dtype) in enumerate(ROPE_CASES):
        g = torch.Generator().manual_seed(500 + idx)
        q = torch.randn(t, hq, d, generator=g).to(dtype)
        k = torch.randn(t, hkv, d, generator=g).to(dtype)
        pos = torch.randint(0, max_pos, (t,), generator=g, dtype=torch.int32)
        inv_freq = ref_rope.compute_default_inv_freq(rotary_dim=rd, theta=theta)
        emb = ref_rope.RotaryEmbedding(rotary_dim=rd, max_position_embeddings=max_pos, inv_freq=inv_freq, interleaved=inter)
        for h in emb.handlers:  # handlers live in a plain list (rotary_embedding.py:139-142): convert each like model.to(dtype) would a registered one
            h.to(table_dtype)
        with torch.inference_mode():
            ref_q, ref_k = emb(q.clone(), k.clone(), pos)
        table = oracle.rotary_cos_sin_table(rd, max_pos, inv_freq)
        # the torch handler's cache holds the same numbers as the fused layout: cos part = first rd columns
        ref_cache = emb.handlers[1].cos_sin_cache.float()
        n = rd // 2
        ref_cos = ref_cache[:, :rd:2] if inter else ref_cache[:, :n]
        ref_sin = ref_cache[:, rd::2] if inter else ref_cache[:, rd:rd + n]
        tt = table.to(table_dtype)
        assert torch.equal(tt[:, 0].float(), ref_cos) and torch.equal(tt[:, 1].float(), ref_sin), f"rope {name}: cos/sin table differs"
        my_q, my_k = oracle.apply_rotary(q, k, pos, tt, rd, inter)
        assert torch.equal(my_q, ref_q) and torch.equal(my_k, ref_k), f"rope {name}: oracle differs from the reference"
        np.savez_compressed(GOLDEN / f"rope_{name}.npz", dtype=DTYPE_NAMES[dtype], table_dtype=DTYPE_NAMES[table_dtype],
                            geometry=np.array([t, hq, hkv, d, rd, max_pos]), theta=theta, interleaved=int(inter),
                            query=to_np(q), key=to_np(k), positions=pos.numpy(), ref_query=to_np(ref_q), ref_key=to_np(ref_k),
                            table_checksum=np.array([float(tt.double().sum())]),
                            # the cos/sin rows the case reads, as the reference's cache holds them: libm's cosf/sinf differ by an ulp
                            # between host CPUs, so the GPU tests take these rows from here instead of recomputing them on the box
                            table_rows=to_np(tt[pos.long()]))

    # ---- the ROPE attention module, driven like a decoder layer drives it (identity projections: q/k/v are slices of a fused tensor)
    name, seq_lens, hq, hkv, d, bs, dtype = "ropeattn_qwen_bf16", [(1, 40), (9, 33), (1, 16)], 28, 4, 128, 16, torch.bfloat16
    theta, max_pos = 1000000., 256
    batch = make_batch(seq_lens, hq, hkv, d, bs, dtype=dtype, seed=700, fused_qkv=True)
    positions = torch.tensor([p for q_len, kv_len in seq_lens for p in range(kv_len - q_len, kv_len)], dtype=torch.int32)
    qkv = torch.cat([batch.query, batch.key, batch.value], dim=1).contiguous()
    inv_freq = ref_rope.compute_default_inv_freq(rotary_dim=d, theta=theta)
    emb = ref_rope.RotaryEmbedding(rotary_dim=d, max_position_embeddings=max_pos, inv_freq=inv_freq, interleaved=False)
    for h in emb.handlers:
        h.to(dtype)
    module = ref_mf.ROPECausalGroupedQueryPageAttention(n_qo_heads=hq, n_kv_heads=hkv, head_dim=d, rotary_emb=emb, qkv_proj=torch.nn.Identity())
    kc, vc = batch.clone_caches()
    builder = ca.AttentionParametersBuilder(hq, hkv, d, bs, torch.device("cpu"))
    for req in batch.requests():
        builder.add_request(*req)
    builder.add_kv_cache(mem.KVCache(kc, vc))
    params = builder.build_attention_parameters()[0]
    with torch.inference_mode():
        ref_out = module.forward(qkv.clone(), positions, params)
    table = oracle.rotary_cos_sin_table(d, max_pos, inv_freq).to(dtype)
    my_kc, my_vc = batch.clone_caches()
    meta = oracle.build_metadata(batch.requests(), bs)
    out, q_rot, k_rot = oracle.rope_attention_layer_forward(
        batch.query, batch.key, batch.value, positions, table, d, False, my_kc, my_vc, torch.tensor(meta.new_cache_slots, dtype=torch.int32),
        meta.q_cu_seq_lens, meta.kv_cu_seq_lens, torch.tensor(meta.block_tables, dtype=torch.int32), meta.cu_blocks_lens, hq, hkv, d)
    assert torch.equal(my_kc, kc) and torch.equal(my_vc, vc), "ROPE attention: oracle caches differ from the reference"
    assert torch.equal(out, ref_out), "ROPE attention: oracle output differs from the reference"
    fp32 = oracle.paged_attention_fp32(q_rot, my_kc, my_vc, meta.q_cu_seq_lens, meta.kv_cu_seq_lens, torch.tensor(meta.block_tables, dtype=torch.int32),
                                       meta.cu_blocks_lens, hq, hkv, d)
    blocks = torch.tensor(sorted(set(batch.block_tables)), dtype=torch.long)
    np.savez_compressed(GOLDEN / f"{name}.npz", dtype=DTYPE_NAMES[dtype], seq_lens=np.array(seq_lens, dtype=np.int64),
                        geometry=np.array([hq, hkv, d, bs, batch.n_blocks, max_pos]), theta=theta, seed=700, qkv=to_np(qkv), positions=positions.numpy(),
                        key_cache=to_np(batch.key_cache), value_cache=to_np(batch.value_cache), owned_blocks=blocks.numpy(),
                        ref_key_cache_owned=to_np(kc[blocks]), ref_value_cache_owned=to_np(vc[blocks]), ref_out=to_np(ref_out),
                        ref_fp32=fp32.numpy(), ref_query_rot=to_np(q_rot), table=to_np(table),
                        new_cache_slots=np.array(meta.new_cache_slots), block_tables=np.array(meta.block_tables),
                        q_cu_seq_lens=np.array(meta.q_cu_seq_lens), cu_blocks_lens=np.array(meta.cu_blocks_lens))


def run_reference_layer(ca, mem, batch):
    """Drive the reference layer exactly as its engine does: builder -> AttentionParameters -> layer.forward."""
    device = torch.device("cpu")
    kc, vc = batch.clone_caches()
    builder = ca.AttentionParametersBuilder(batch.n_qo_heads, batch.n_kv_heads, batch.head_dim, batch.block_size, device)
    for q_len, kv_len, slots, table in batch.requests():
        builder.add_request(q_len, kv_len, slots, table)
    builder.add_kv_cache(mem.KVCache(kc, vc))
    params = builder.build_attention_parameters()[0]
    layer = ca.CausalGroupedQueryPageAttention(ca.CausalGroupedQueryPageAttentionConfig(batch.n_qo_heads, batch.n_kv_heads, batch.head_dim))
    with torch.inference_mode():
        out = layer(batch.query, batch.key, batch.value, params).o
    return out, kc, vc, params


def main() -> None:
    torch.manual_seed(0)
    ca, mem = import_reference()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    report = []
    if "--only-rope" in sys.argv:  # rewrite rope_*.npz / ropeattn_*.npz only (every case is seeded on its own)
        rope_goldens(ca, mem)
        return

    for idx, (name, seq_lens, hq, hkv, d, bs, dtype, fused) in enumerate(ATTENTION_CASES):
        batch = make_batch(seq_lens, hq, hkv, d, bs, dtype=dtype, seed=100 + idx, fused_qkv=fused)
        ref_out, ref_kc, ref_vc, params = run_reference_layer(ca, mem, batch)

        # oracle on the same inputs
        kc, vc = batch.clone_caches()
        meta = oracle.build_metadata(batch.requests(), bs)
        out = oracle.attention_layer_forward(batch.query, batch.key, batch.value, kc, vc,
                                             torch.tensor(meta.new_cache_slots, dtype=torch.int32), meta.q_cu_seq_lens,
                                             meta.kv_cu_seq_lens, torch.tensor(meta.block_tables, dtype=torch.int32),
                                             meta.cu_blocks_lens, hq, hkv, d)
        assert torch.equal(kc, ref_kc) and torch.equal(vc, ref_vc), f"{name}: oracle KV append differs from the reference"
        assert torch.equal(out, ref_out), f"{name}: oracle attention differs from the reference"
        for field in ("q_cu_seq_lens", "kv_cu_seq_lens", "paged_kv_last_page_len", "new_cache_slots", "block_tables", "cu_blocks_lens"):
            assert getattr(params, field).tolist() == getattr(meta, field), f"{name}: metadata {field} differs"
        assert params.num_sequences == meta.num_sequences and params.all_sequences_decode == meta.all_sequences_decode
        assert params.q_max_seq_len == meta.q_max_seq_len and params.kv_max_seq_len == meta.kv_max_seq_len
        fp32 = oracle.paged_attention_fp32(batch.query.view(-1, hq, d), kc, vc, meta.q_cu_seq_lens, meta.kv_cu_seq_lens,
                                           torch.tensor(meta.block_tables, dtype=torch.int32), meta.cu_blocks_lens, hq, hkv, d)
        report.append((name, float((ref_out.float() - fp32).abs().max())))

        # touched-slot view of the caches keeps the fixture small: rows of the blocks this batch owns, after the append
        blocks = torch.tensor(sorted(set(batch.block_tables)), dtype=torch.long)
        np.savez_compressed(
            GOLDEN / f"attn_{name}.npz",
            dtype=DTYPE_NAMES[dtype], seq_lens=np.array(seq_lens, dtype=np.int64), geometry=np.array([hq, hkv, d, bs, batch.n_blocks]),
            seed=100 + idx, fused_qkv=int(fused),
            query=to_np(batch.query), key=to_np(batch.key), value=to_np(batch.value),
            key_cache=to_np(batch.key_cache), value_cache=to_np(batch.value_cache),
            q_cu_seq_lens=np.array(meta.q_cu_seq_lens), kv_cu_seq_lens=np.array(meta.kv_cu_seq_lens),
            paged_kv_last_page_len=np.array(meta.paged_kv_last_page_len), new_cache_slots=np.array(meta.new_cache_slots),
            block_tables=np.array(meta.block_tables), cu_blocks_lens=np.array(meta.cu_blocks_lens),
            owned_blocks=blocks.numpy(), ref_key_cache_owned=to_np(ref_kc[blocks]), ref_value_cache_owned=to_np(ref_vc[blocks]),
            ref_out=to_np(ref_out), ref_fp32=fp32.numpy())

    # ---- set_image_cache via the reference's TokenCache CPU path ------------------------------------------------
    g = torch.Generator().manual_seed(7)
    n_blocks, bs, heads, d = 3, 576, 2, 64
    cache = torch.randn(n_blocks, bs, heads, d, generator=g).to(torch.float16)
    tokens = torch.randn(40, heads, d, generator=g).to(torch.float16)
    slots = torch.randperm(n_blocks * bs, generator=g)[:40].to(torch.int32)
    ref_cache = cache.clone()
    mem.TokenCache([ref_cache]).set_caches(slots, [tokens])
    mine = cache.clone()
    oracle.set_image_cache(slots, tokens, mine)
    assert torch.equal(mine, ref_cache), "oracle set_image_cache differs from the reference"
    touched = torch.unique(slots.long() // bs)
    np.savez_compressed(GOLDEN / "image_cache.npz", geometry=np.array([n_blocks, bs, heads, d]), slots=slots.numpy(),
                        tokens=to_np(tokens), cache_seed=7, ref_rows=to_np(ref_cache.view(-1, heads, d)[slots.long()]),
                        ref_checksum=np.array([int(ref_cache.view(torch.int16).to(torch.int64).sum())]),
                        cache=to_np(cache), touched_blocks=touched.numpy())

    # ---- allocator + v2p known answers -----------------------------------------------------------------------------
    trace = []
    ref_alloc, my_alloc = mem.BlockAllocator(40), oracle.BlockAllocator(40)
    rng = np.random.default_rng(3)
    held: list[int] = []
    for _ in range(30):
        if held and rng.random() < 0.4:
            k = int(rng.integers(1, len(held) + 1))
            give, held = held[:k], held[k:]
            ref_alloc.free(list(give))
            my_alloc.free(list(give))
            trace.append(("free", give, []))
        else:
            n = int(rng.integers(0, 9))
            a, b = ref_alloc.allocate(n), my_alloc.allocate(n)
            assert a == b, "oracle BlockAllocator differs from the reference"
            held += a
            trace.append(("allocate", [n], a))
    table = ref_alloc.allocate(5)
    assert my_alloc.allocate(5) == table
    vids = [0, 1, 15, 16, 17, 31, 47, 64, 79]
    fake_self = types.SimpleNamespace(block_size=16)
    fake_cache = types.SimpleNamespace(block_table=table)
    ref_slots = mem.TokenCacheBlockManager.v2p(fake_self, fake_cache, vids)
    assert oracle.v2p(table, 16, vids) == ref_slots
    np.savez_compressed(GOLDEN / "allocator.npz",
                        ops=np.array([t[0] for t in trace]), args=np.array([",".join(map(str, t[1])) for t in trace]),
                        results=np.array([",".join(map(str, t[2])) for t in trace]),
                        v2p_table=np.array(table), v2p_vids=np.array(vids), v2p_slots=np.array(ref_slots))

    rope_goldens(ca, mem)
    mha_goldens()

    print("oracle == reference on every case (bit-exact outputs, caches, metadata, allocator, v2p, rotary, ROPE attention, vision attention)")
    for name, err in report:
        print(f"  {name:28s} max |reference(dtype) - fp32 recompute| = {err:.3e}")
    total = sum(p.stat().st_size for p in GOLDEN.glob("*.npz"))
    print(f"fixtures: {len(list(GOLDEN.glob('*.npz')))} files, {total / 1024:.0f} KiB in {GOLDEN}")


if __name__ == "__main__":
    main()
