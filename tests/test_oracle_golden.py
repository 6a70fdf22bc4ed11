"""CPU: the oracle reproduces the reference's frozen outputs (tests/golden, made by oracle/make_golden.py from the
real reference).  This is the pin that lets the GPU tests trust oracle/paged_kv_oracle.py where /root/reference is absent."""
import numpy as np
import torch

from conftest import GOLDEN, _from_np
from oracle import paged_kv_oracle as oracle


def test_attention_layer_matches_reference(golden_attention):
    g = golden_attention
    kc, vc = g.key_cache.clone(), g.value_cache.clone()
    out = oracle.attention_layer_forward(
        g.query, g.key, g.value, kc, vc, torch.tensor(g.new_cache_slots, dtype=torch.int32), g.q_cu_seq_lens, g.kv_cu_seq_lens,
        torch.tensor(g.block_tables, dtype=torch.int32), g.cu_blocks_lens, g.n_qo_heads, g.n_kv_heads, g.head_dim)
    owned = torch.tensor(g.owned_blocks)
    assert torch.equal(kc[owned], g.ref_key_cache_owned), "KV append differs from the reference"
    assert torch.equal(vc[owned], g.ref_value_cache_owned)
    # blocks the batch does not own are untouched
    mask = torch.ones(g.n_blocks, dtype=torch.bool)
    mask[owned] = False
    assert torch.equal(kc[mask], g.key_cache[mask]) and torch.equal(vc[mask], g.value_cache[mask])
    assert torch.equal(out, g.ref_out), "attention output differs from the reference"
    fp32 = oracle.paged_attention_fp32(g.query.view(-1, g.n_qo_heads, g.head_dim), kc, vc, g.q_cu_seq_lens, g.kv_cu_seq_lens,
                                       torch.tensor(g.block_tables, dtype=torch.int32), g.cu_blocks_lens, g.n_qo_heads, g.n_kv_heads, g.head_dim)
    assert torch.equal(fp32, g.ref_fp32)


def test_metadata_matches_reference(golden_attention):
    g = golden_attention
    meta = oracle.build_metadata(g.requests(), g.block_size)
    for name in ("q_cu_seq_lens", "kv_cu_seq_lens", "paged_kv_last_page_len", "new_cache_slots", "block_tables", "cu_blocks_lens"):
        assert getattr(meta, name) == getattr(g, name), name
    assert meta.num_sequences == len(g.seq_lens)
    assert meta.all_sequences_decode == all(q == 1 for q, _ in g.seq_lens)
    assert meta.q_max_seq_len == max(q for q, _ in g.seq_lens) and meta.kv_max_seq_len == max(kv for _, kv in g.seq_lens)


def test_image_cache_matches_reference():
    z = np.load(GOLDEN / "image_cache.npz")
    n_blocks, bs, heads, d = (int(v) for v in z["geometry"])
    cache = _from_np(z["cache"], torch.float16)
    tokens = _from_np(z["tokens"], torch.float16)
    slots = torch.from_numpy(z["slots"])
    before = cache.clone()
    oracle.set_image_cache(slots, tokens, cache)
    flat = cache.view(-1, heads, d)
    assert torch.equal(flat[slots.long()], _from_np(z["ref_rows"], torch.float16))
    assert int(cache.view(torch.int16).to(torch.int64).sum()) == int(z["ref_checksum"][0])
    untouched = torch.ones(n_blocks * bs, dtype=torch.bool)
    untouched[slots.long()] = False
    assert torch.equal(flat[untouched], before.view(-1, heads, d)[untouched])


def _ints(s) -> list[int]:
    s = str(s)
    return [int(v) for v in s.split(",")] if s else []


def test_allocator_and_v2p_match_reference():
    z = np.load(GOLDEN / "allocator.npz")
    alloc = oracle.BlockAllocator(40)
    for op, args, result in zip(z["ops"], z["args"], z["results"]):
        if str(op) == "free":
            alloc.free(_ints(args))
        else:
            assert alloc.allocate(_ints(args)[0]) == _ints(result)
    assert oracle.v2p([int(v) for v in z["v2p_table"]], 16, [int(v) for v in z["v2p_vids"]]) == [int(v) for v in z["v2p_slots"]]


def test_migrate_blocks_restatement_properties():
    """No reference fixture exists for migration (parity unpinned); check the restated index arithmetic by properties."""
    g = torch.Generator().manual_seed(5)
    src = torch.randn(3, 2, 7, 4, 2, 8, generator=g)
    dst = torch.randn(3, 2, 5, 4, 2, 8, generator=g)
    before = dst.clone()
    oracle.migrate_blocks([6, 0, 3], [1, 4, 2], src, dst)
    for s, d in zip([6, 0, 3], [1, 4, 2]):
        assert torch.equal(dst[:, :, d], src[:, :, s])
    assert torch.equal(dst[:, :, [0, 3]], before[:, :, [0, 3]])
    # round trip restores the moved blocks
    back = torch.zeros_like(src)
    oracle.migrate_blocks([1, 4, 2], [6, 0, 3], dst, back)
    assert torch.equal(back[:, :, [6, 0, 3]], src[:, :, [6, 0, 3]])


def _mha_cases(qwen: bool):
    return sorted(p for p in GOLDEN.glob("mha_*.npz") if ("qwen" in p.stem) == qwen)


def test_vision_attention_matches_reference():
    """Un-paged attention of the vision towers: oracle vs the frozen outputs of the reference's Torch handlers
    (multihead_attention.py:40-73 equal-length batches, :235-256 packed cu_seqlens form)."""
    from conftest import _TORCH_DTYPES
    assert _mha_cases(False) and _mha_cases(True)
    for path in _mha_cases(False):
        z = np.load(path)
        dtype = _TORCH_DTYPES[str(z["dtype"])]
        batch, seq_len, heads, d = (int(v) for v in z["geometry"])
        q, k, v, ref = (_from_np(z[n], dtype) for n in ("query", "key", "value", "ref_out"))
        assert torch.equal(oracle.multi_head_attention(q, k, v, heads, d), ref), path.stem
        cu = list(range(0, (batch + 1) * seq_len, seq_len))
        fp32 = oracle.varlen_attention_fp32(q.view(-1, heads, d), k.view(-1, heads, d), v.view(-1, heads, d), cu, cu)
        assert (ref.float().view(-1, heads * d) - fp32).abs().max() < 2e-2  # the dtype-rounded reference sits inside the GPU tolerance
    for path in _mha_cases(True):
        z = np.load(path)
        dtype = _TORCH_DTYPES[str(z["dtype"])]
        _, total, heads, d = (int(v) for v in z["geometry"])
        q, k, v, ref = (_from_np(z[n], dtype) for n in ("query", "key", "value", "ref_out"))
        cu = [int(x) for x in z["cu_seqlens"]]
        assert torch.equal(oracle.qwen_multi_head_attention(q, k, v, total, cu, d), ref), path.stem
        fp32 = oracle.varlen_attention_fp32(q, k, v, cu, cu)
        assert (ref.float() - fp32).abs().max() < 3e-2


def test_varlen_fp32_causal_agrees_with_paged_oracle():
    """The causal form of the un-paged recompute is the same function as the (pinned) paged recompute on contiguous pages."""
    g = torch.Generator().manual_seed(11)
    hq, hkv, d, bs = 4, 2, 32, 16
    lens = [(5, 20), (16, 16), (1, 33)]
    q = torch.randn(sum(a for a, _ in lens), hq, d, generator=g)
    n_blocks = sum((kv + bs - 1) // bs for _, kv in lens)
    kc, vc = torch.randn(n_blocks, bs, hkv, d, generator=g), torch.randn(n_blocks, bs, hkv, d, generator=g)
    q_cu, kv_cu, tables, cu_blocks, ks, vs, blk = [0], [0], [], [0], [], [], 0
    for ql, kv in lens:
        nb = (kv + bs - 1) // bs
        tables += list(range(blk, blk + nb))
        ks.append(kc[blk:blk + nb].reshape(-1, hkv, d)[:kv])
        vs.append(vc[blk:blk + nb].reshape(-1, hkv, d)[:kv])
        blk += nb
        q_cu.append(q_cu[-1] + ql), kv_cu.append(kv_cu[-1] + kv), cu_blocks.append(cu_blocks[-1] + nb)
    paged = oracle.paged_attention_fp32(q, kc, vc, q_cu, kv_cu, torch.tensor(tables, dtype=torch.int32), cu_blocks, hq, hkv, d)
    flat = oracle.varlen_attention_fp32(q, torch.cat(ks), torch.cat(vs), q_cu, kv_cu, causal=True)
    assert torch.allclose(paged, flat, atol=1e-6, rtol=1e-6)


def _tiny_paged_case(seed=12):
    g = torch.Generator().manual_seed(seed)
    hq, hkv, d, bs = 4, 2, 32, 16
    lens = [(5, 40), (16, 16), (1, 33), (3, 3)]
    q = torch.randn(sum(a for a, _ in lens), hq, d, generator=g)
    n_blocks = sum((kv + bs - 1) // bs for _, kv in lens)
    kc, vc = torch.randn(n_blocks, bs, hkv, d, generator=g), torch.randn(n_blocks, bs, hkv, d, generator=g)
    q_cu, kv_cu, tables, cu_blocks, blk = [0], [0], [], [0], 0
    for ql, kv in lens:
        nb = (kv + bs - 1) // bs
        tables += list(range(blk, blk + nb))
        blk += nb
        q_cu.append(q_cu[-1] + ql), kv_cu.append(kv_cu[-1] + kv), cu_blocks.append(cu_blocks[-1] + nb)
    return q, kc, vc, q_cu, kv_cu, torch.tensor(tables, dtype=torch.int32), cu_blocks, hq, hkv, d, lens


def test_score_options_oracle_reduces_to_the_pinned_function():
    """`paged_attention_options_fp32` with mha_varlen_fwd's defaults (softcap 0, window (-1, 0), no alibi) IS the pinned restatement of
    the reference's torch handler; a window wider than every sequence and a zero alibi slope change nothing either."""
    q, kc, vc, q_cu, kv_cu, tables, cu_blocks, hq, hkv, d, _ = _tiny_paged_case()
    scale = 1.0 / d ** 0.5
    pinned = oracle.paged_attention_fp32(q, kc, vc, q_cu, kv_cu, tables, cu_blocks, hq, hkv, d)
    for kwargs in ({}, {"window_size_left": 1000}, {"alibi_slopes": torch.zeros(hq)}, {"alibi_slopes": torch.zeros(len(q_cu) - 1, hq)}):
        got = oracle.paged_attention_options_fp32(q, kc, vc, q_cu, kv_cu, tables, cu_blocks, hq, hkv, d, scale, **kwargs)
        assert torch.allclose(got, pinned, atol=1e-6, rtol=1e-6), kwargs


def test_score_options_oracle_hand_checked_rows():
    """Rows whose result follows from the definitions (flash_api.cpp:93-111, src/mask.h:54-62, 183-186) without running a softmax:
    window_left = 0 and window_right = 0 leave exactly the diagonal key, so the output row is that key's V row; a huge alibi slope
    does the same through the bias; softcap -> 0+ flattens every visible score, so the row is the mean of the visible V rows."""
    q, kc, vc, q_cu, kv_cu, tables, cu_blocks, hq, hkv, d, lens = _tiny_paged_case(13)
    scale = 1.0 / d ** 0.5
    group = hq // hkv
    diag = oracle.paged_attention_options_fp32(q, kc, vc, q_cu, kv_cu, tables, cu_blocks, hq, hkv, d, scale, window_size_left=0, window_size_right=0)
    steep = oracle.paged_attention_options_fp32(q, kc, vc, q_cu, kv_cu, tables, cu_blocks, hq, hkv, d, scale, alibi_slopes=torch.full((hq,), 1e4))
    flat = oracle.paged_attention_options_fp32(q, kc, vc, q_cu, kv_cu, tables, cu_blocks, hq, hkv, d, scale, softcap=1e-6)
    blk = 0
    for b, (ql, kv) in enumerate(lens):
        nb = (kv + 15) // 16
        v = vc[blk:blk + nb].reshape(-1, hkv, d)[:kv].repeat_interleave(group, dim=1)  # [kv, hq, d]
        blk += nb
        for i in range(ql):
            i_abs = i + kv - ql
            row = q_cu[b] + i
            assert torch.allclose(diag[row].view(hq, d), v[i_abs], atol=1e-6)
            assert torch.allclose(steep[row].view(hq, d), v[i_abs], atol=1e-5)
            assert torch.allclose(flat[row].view(hq, d), v[: i_abs + 1].mean(dim=0), atol=1e-4)
    # an empty window (the row sees no key at all) yields zeros: q_len > kv_len rows under a causal window
    out = oracle.paged_attention_options_fp32(q[:2], kc, vc, [0, 2], [0, 1], tables[:1], [0, 1], hq, hkv, d, scale)
    assert torch.equal(out[0], torch.zeros(hq * d)) and bool(out[1].abs().sum() > 0)


def test_get_image_cache_restatement():
    z = np.load(GOLDEN / "image_cache.npz")
    n_blocks, bs, heads, d = (int(v) for v in z["geometry"])
    cache = _from_np(z["cache"], torch.float16)
    tokens = _from_np(z["tokens"], torch.float16)
    slots = torch.from_numpy(z["slots"])
    oracle.set_image_cache(slots, tokens, cache)
    # read-back of what the (pinned) scatter wrote: parameters_builder.py:50-54
    assert torch.equal(oracle.get_image_cache(slots, cache), tokens.view(-1, heads * d))
