"""GPU: paged attention kernels (through the reference-shaped Python API, which calls the C ABI) vs the CPU oracle.

Tolerances (BASELINE.json north_star): 16-bit dtypes |ours - fp32 recompute| <= 2e-2 + 1e-2 * |fp32 recompute|;
fp32 (CPU-reference config 1) 1e-4 / 1e-4.  Shape grid from the reference's tests/layer/test_attention.py:42-48."""
import math

import pytest
import torch

from hydrainfer_b200.workloads import make_batch
from oracle import paged_kv_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SIMT, TC, DEC, PAIR = 1, 2, 3, 4


def tolerances(dtype):
    return (1e-4, 1e-4) if dtype == torch.float32 else (2e-2, 1e-2)


def assert_close_to_fp32(out: torch.Tensor, fp32: torch.Tensor, dtype, what=""):
    atol, rtol = tolerances(dtype)
    out = out.float().cpu()
    assert out.shape == fp32.shape, (out.shape, fp32.shape)
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    err = (out - fp32).abs()
    bound = atol + rtol * fp32.abs()
    worst = (err - bound).max().item()
    assert worst <= 0, f"{what}: max |err| {err.max().item():.4e}, exceeds atol {atol} + rtol {rtol} by {worst:.3e}"


def tc_supported(head_dim, dtype, block_size):
    return head_dim == 128 and dtype in (torch.float16, torch.bfloat16) and block_size in (8, 16, 32, 64, 128)


def pair_supported(head_dim, dtype, block_size):
    return head_dim % 16 == 0 and 16 <= head_dim <= 128 and dtype in (torch.float16, torch.bfloat16) and block_size in (8, 16, 32, 64)


def run_attention(query3d, key_cache, value_cache, q_cu, kv_cu, block_tables, cu_blocks, q_max, kv_max, head_dim, path):
    """mha_varlen_fwd exactly as FlashAttentionCausalGroupedQueryPageAttentionHandler calls it (causal_attention.py:274-291)."""
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    out = torch.empty_like(query3d) if query3d.is_contiguous() else torch.empty(query3d.shape, dtype=query3d.dtype, device=query3d.device)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    mha_varlen_fwd(out, query3d, key_cache, value_cache, i32(q_cu), i32(kv_cu), i32(block_tables), i32(cu_blocks), None,
                   q_max, kv_max, 1.0 / math.sqrt(head_dim), 0, -1, 0, 0, path)
    torch.cuda.synchronize()
    return out


def oracle_fp32(batch, kc, vc):
    return oracle.paged_attention_fp32(batch.query.view(-1, batch.n_qo_heads, batch.head_dim), kc, vc, batch.q_cu_seq_lens,
                                       batch.kv_cu_seq_lens, torch.tensor(batch.block_tables, dtype=torch.int32),
                                       batch.cu_blocks_lens, batch.n_qo_heads, batch.n_kv_heads, batch.head_dim)


def check_batch(batch, paths, what=""):
    """Append on CPU (oracle) so attention is tested in isolation, then compare each kernel path with the fp32 recompute."""
    kc, vc = batch.clone_caches()
    t = batch.n_tokens
    oracle.set_kv_cache(torch.tensor(batch.new_cache_slots, dtype=torch.int32), batch.key.view(t, batch.n_kv_heads, batch.head_dim),
                        batch.value.view(t, batch.n_kv_heads, batch.head_dim), kc, vc)
    fp32 = oracle_fp32(batch, kc, vc)
    q_d = batch.query.to(DEV)
    if not batch.query.is_contiguous():  # keep the fused-qkv row stride on the device too
        full = batch.query._base.to(DEV) if batch.query._base is not None else None
        if full is not None:
            q_d = full[:, : batch.n_qo_heads * batch.head_dim]
    q3 = q_d.view(t, batch.n_qo_heads, batch.head_dim)
    kc_d, vc_d = kc.to(DEV), vc.to(DEV)
    for path in paths:
        out = run_attention(q3, kc_d, vc_d, batch.q_cu_seq_lens, batch.kv_cu_seq_lens, batch.block_tables, batch.cu_blocks_lens,
                            batch.q_max, batch.kv_max, batch.head_dim, path)
        assert_close_to_fp32(out.reshape(t, -1), fp32, batch.dtype, f"{what} path={path}")
    return fp32


def paths_for(head_dim, dtype, block_size=16, group=1):
    """Every kernel path that covers the shape: 1 split-KV CUDA-core, 2 tcgen05 tile, 3 tcgen05 swapped-operand decode,
    4 tcgen05 pair-tile (prefill), 0 auto."""
    import os
    if os.environ.get("HI_TEST_SKIP_TC") == "1":  # dev switch: validate the split-KV kernel alone
        return [SIMT]
    if head_dim == 256 and dtype in (torch.float16, torch.bfloat16) and block_size in (8, 16, 32, 64, 128):
        return [SIMT, TC, 0]  # the tile kernel's head_dim-256 instance (one CTA per SM, S | O = 128 + 256 TMEM columns)
    if not tc_supported(head_dim, dtype, block_size):
        if pair_supported(head_dim, dtype, block_size):  # head_dim 64 / 96: the pair-tile kernel's variable-head-dim mode
            return ([SIMT] if head_dim == 64 else []) + [PAIR, 0]
        return [SIMT, 0]
    return [SIMT, TC, DEC, PAIR, 0] if group <= 16 else [SIMT, TC, PAIR, 0]


def paths_of(batch):
    return paths_for(batch.head_dim, batch.dtype, batch.block_size, batch.n_qo_heads // batch.n_kv_heads)


# ---- golden fixtures: the reference's own outputs ---------------------------------------------------------------------
def test_golden_layer_forward(golden_attention):
    """Full layer (append + attend) through the reference-shaped API against the reference's frozen output."""
    from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
    from hydrainfer_b200.memory import KVCache
    g = golden_attention
    for path in paths_for(g.head_dim, g.dtype, g.block_size, g.n_qo_heads // g.n_kv_heads):
        kc, vc = g.key_cache.to(DEV), g.value_cache.to(DEV)
        builder = AttentionParametersBuilder(g.n_qo_heads, g.n_kv_heads, g.head_dim, g.block_size, torch.device(DEV))
        for req in g.requests():
            builder.add_request(*req)
        builder.add_kv_cache(KVCache(kc, vc))
        params = builder.build_attention_parameters()[0]
        layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(g.n_qo_heads, g.n_kv_heads, g.head_dim))
        layer.handler.path = path
        if g.fused_qkv:
            qkv = torch.cat([g.query, g.key, g.value], dim=1).to(DEV)
            w_q, w_k = g.query.shape[1], g.key.shape[1]
            q, k, v = qkv[:, :w_q], qkv[:, w_q:w_q + w_k], qkv[:, w_q + w_k:]
        else:
            q, k, v = g.query.to(DEV), g.key.to(DEV), g.value.to(DEV)
        out = layer(q, k, v, params).o
        torch.cuda.synchronize()
        owned = torch.tensor(g.owned_blocks)
        assert torch.equal(kc.cpu()[owned], g.ref_key_cache_owned) and torch.equal(vc.cpu()[owned], g.ref_value_cache_owned), "KV append not bit-exact"
        assert out.shape == g.ref_out.shape and out.dtype == g.dtype
        assert_close_to_fp32(out, g.ref_fp32, g.dtype, f"{g.name} path={path}")
        # and against the reference's own rounded output (informational bound: two roundings apart)
        atol, rtol = tolerances(g.dtype)
        assert torch.allclose(out.float().cpu(), g.ref_out.float(), atol=2 * atol, rtol=2 * rtol)


# ---- reference test grid ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("head_size", [64, 128, 256])
@pytest.mark.parametrize("num_heads", [(8, 8), (8, 4), (8, 2), (8, 1)])
def test_reference_grid(num_heads, head_size, dtype):
    seq_lens = [(1, 100), (15, 15), (111, 234), (1, 1024)]  # tests/layer/test_attention.py:42
    batch = make_batch(seq_lens, num_heads[0], num_heads[1], head_size, 16, n_blocks=120, dtype=dtype, seed=42)
    check_batch(batch, paths_of(batch), f"grid heads={num_heads} d={head_size}")


@pytest.mark.parametrize("dtype", [torch.float32])
def test_cpu_reference_config_fp32(dtype):
    # BASELINE config 1 geometry (32 heads, d=128, batch 8, ctx 512, block 16, fp32), decode
    batch = make_batch([(1, 512)] * 8, 32, 32, 128, 16, dtype=dtype, seed=1)
    check_batch(batch, [SIMT, 0], "cfg1 fp32")


# ---- ragged / edge cases ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("heads", [(32, 32), (28, 4), (64, 8), (8, 1), (6, 2)])
def test_ragged_decode(heads):
    g = torch.Generator().manual_seed(heads[0])
    lens = [1, 2, 15, 16, 17, 31, 32, 33, 127, 128, 129, 255, 256, 257, 600, 1025]
    lens += torch.randint(1, 900, (8,), generator=g).tolist()
    batch = make_batch([(1, L) for L in lens], heads[0], heads[1], 128, 16, dtype=torch.bfloat16, seed=7)
    check_batch(batch, paths_of(batch), f"ragged decode heads={heads}")


@pytest.mark.parametrize("heads", [(8, 8), (28, 4), (16, 2)])
def test_chunked_prefill_shapes(heads):
    # q < kv: the bottom-right aligned mask; q tiles that start mid-sequence; kv not a multiple of the 128-token tile
    seq_lens = [(128, 128), (129, 300), (64, 1000), (200, 200), (1, 77), (37, 165), (256, 513)]
    batch = make_batch(seq_lens, heads[0], heads[1], 128, 16, dtype=torch.bfloat16, seed=9)
    check_batch(batch, paths_of(batch), f"chunked prefill heads={heads}")


def test_mixed_batch_qwen_geometry_fp16_fused():
    # config 3 flavour (28 q / 4 kv heads): decode rows + chunked prefill rows in one call, q a strided qkv slice
    seq_lens = [(1, 300), (1, 45), (512, 2048), (1, 1999), (300, 300), (1, 16)]
    batch = make_batch(seq_lens, 28, 4, 128, 16, dtype=torch.float16, seed=13, fused_qkv=True)
    check_batch(batch, paths_of(batch), "mixed qwen fp16")


@pytest.mark.parametrize("block_size", [8, 32, 64])
def test_other_block_sizes(block_size):
    seq_lens = [(1, 70), (50, 131), (1, 5), (130, 130)]
    batch = make_batch(seq_lens, 8, 2, 128, block_size, dtype=torch.bfloat16, seed=21)
    check_batch(batch, paths_of(batch), f"block_size={block_size}")


def test_small_block_size_simt_only():
    batch = make_batch([(1, 9), (5, 30)], 4, 2, 64, 4, dtype=torch.float16, seed=22)
    check_batch(batch, [SIMT, 0], "block_size=4")


def test_single_token_single_sequence():
    batch = make_batch([(1, 1)], 8, 8, 128, 16, dtype=torch.bfloat16, seed=2)
    fp32 = check_batch(batch, paths_of(batch), "one token")
    # softmax over one key is 1: the output is the value row itself
    t = 0
    v_row = batch.value.view(1, 8, 128).float().reshape(1, -1)
    assert torch.allclose(fp32, v_row, atol=1e-6)


def test_garbage_in_unused_slots_does_not_leak():
    """Slots past kv_len inside the last page, and unowned blocks, may hold NaN/Inf (an uninitialised pool)."""
    seq_lens = [(1, 5), (20, 37), (1, 129)]
    batch = make_batch(seq_lens, 8, 2, 128, 16, dtype=torch.bfloat16, seed=23)
    kc, vc = batch.clone_caches()
    t = batch.n_tokens
    oracle.set_kv_cache(torch.tensor(batch.new_cache_slots, dtype=torch.int32), batch.key.view(t, 2, 128), batch.value.view(t, 2, 128), kc, vc)
    fp32 = oracle_fp32(batch, kc, vc)
    poison_k, poison_v = kc.clone(), vc.clone()
    used = torch.zeros(batch.n_blocks * 16, dtype=torch.bool)
    for (q, kv), table in zip(batch.seq_lens, batch.per_seq_block_tables):
        for pos in range(kv):
            used[table[pos // 16] * 16 + pos % 16] = True
    poison_k.view(-1, 2, 128)[~used] = float("nan")
    poison_v.view(-1, 2, 128)[~used] = float("inf")
    q3 = batch.query.to(DEV).view(t, 8, 128)
    for path in paths_of(batch):
        out = run_attention(q3, poison_k.to(DEV), poison_v.to(DEV), batch.q_cu_seq_lens, batch.kv_cu_seq_lens, batch.block_tables,
                            batch.cu_blocks_lens, batch.q_max, batch.kv_max, 128, path)
        assert_close_to_fp32(out.reshape(t, -1), fp32, torch.bfloat16, f"poisoned pool path={path}")


def test_large_score_magnitudes_stay_finite():
    # scores of a few hundred: exercises the running-max logic (and the lazy rescale of the tile kernel)
    batch = make_batch([(1, 700), (140, 400)], 8, 4, 128, 16, dtype=torch.bfloat16, seed=24)
    batch.query.mul_(6.0)
    batch.key_cache.mul_(4.0)
    batch.key.mul_(4.0)
    check_batch(batch, paths_of(batch), "large scores")


@pytest.mark.parametrize("stages", ["1", "2", "3"])
@pytest.mark.parametrize("splits", ["1", "2", "3", "8"])
def test_tile_kernel_split_kv_and_ring_depth_variants(monkeypatch, splits, stages):
    """Split-KV partials + merge and every K/V ring depth of the tile kernel give the same answer (tuning overrides)."""
    monkeypatch.setenv("HI_TC_SPLITS", splits)
    monkeypatch.setenv("HI_TC_STAGES", stages)
    monkeypatch.setenv("HI_DEC_SPLITS", splits)
    seq_lens = [(1, 1300), (1, 17), (1, 128), (1, 129), (1, 640), (3, 700), (1, 2049)]
    for heads in ((28, 4), (8, 8)):
        batch = make_batch(seq_lens, heads[0], heads[1], 128, 16, dtype=torch.bfloat16, seed=40)
        check_batch(batch, [TC, DEC, PAIR], f"splits={splits} stages={stages} heads={heads}")


@pytest.mark.parametrize("splits", ["2", "3", "5"])
def test_pair_kernel_direct_and_split_tiles_share_a_launch(monkeypatch, splits):
    """With split-KV on, a query tile whose keys all fall into the first chunk is written straight to `out` and skipped by the
    merge; tiles that span chunks go through the fp32 partials.  One launch holds both kinds (early vs late tiles of the same
    prefill, short vs long sequences); the workspace is filled with NaN bit patterns first, so a merge that read a partial
    nobody wrote, or skipped a row that needed it, cannot pass."""
    from hydrainfer_b200._C.kernel import flash_attn
    monkeypatch.setenv("HI_TC_SPLITS", splits)
    seq_lens = [(600, 600), (300, 2000), (1, 3000), (64, 64), (1, 40), (200, 1100)]
    for heads in ((28, 4), (8, 8), (16, 1)):
        batch = make_batch(seq_lens, heads[0], heads[1], 128, 16, dtype=torch.bfloat16, seed=43)
        flash_attn.workspace(0).fill_(0xFF)
        check_batch(batch, [PAIR], f"direct+split tiles splits={splits} heads={heads}")
        flash_attn.workspace(0).fill_(0xFF)
        check_batch(batch, [TC, DEC] if heads[0] // heads[1] <= 16 else [TC], f"merge of the other tcgen05 paths splits={splits} heads={heads}")


@pytest.mark.parametrize("static", ["0", "1"])
@pytest.mark.parametrize("ctas", ["1", "3", "148"])
def test_pair_kernel_item_walk_variants(monkeypatch, static, ctas):
    """The persistent pair kernel walks its work items with a dynamic counter (items claimed just in time) or, without a
    workspace counter, in a static boustrophedon order; a launch may have far fewer CTAs than items (many rounds per CTA).
    Same answer in every combination, with and without the host plan."""
    from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
    from hydrainfer_b200.memory import KVCache
    monkeypatch.setenv("HI_PAIR_STATIC", static)
    monkeypatch.setenv("HI_PAIR_CTAS", ctas)
    seq_lens = [(1, 700), (130, 130), (75, 900), (1, 64), (300, 300), (1, 1), (40, 1300)]
    for heads in ((28, 4), (8, 8)):
        batch = make_batch(seq_lens, heads[0], heads[1], 128, 16, dtype=torch.bfloat16, seed=44)
        fp32 = check_batch(batch, [PAIR], f"item walk static={static} ctas={ctas} heads={heads} (no plan)")
        builder = AttentionParametersBuilder(heads[0], heads[1], 128, 16, torch.device(DEV))
        for req in batch.requests():
            builder.add_request(*req)
        builder.add_kv_cache(KVCache(batch.key_cache.to(DEV), batch.value_cache.to(DEV)))
        params = builder.build_attention_parameters()[0]
        layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(heads[0], heads[1], 128))
        layer.handler.path = PAIR
        out = layer(batch.query.to(DEV), batch.key.to(DEV), batch.value.to(DEV), params).o
        torch.cuda.synchronize()
        assert_close_to_fp32(out, fp32, batch.dtype, f"item walk static={static} ctas={ctas} heads={heads} (plan)")


@pytest.mark.parametrize("device_plan", ["1", "0"])
def test_plan_built_on_the_device_for_the_bare_seam(monkeypatch, device_plan):
    """mha_varlen_fwd without a host plan (the reference's own call, 16 positional arguments): the tile list is built by a
    one-CTA kernel from the device-side sequence lengths (counting sort, heaviest first, padded with empty items).  More
    sequences than the kernel has warps, a prefill with more pairs than a warp has lanes, empty and one-token sequences."""
    monkeypatch.setenv("HI_PAIR_DEVICE_PLAN", device_plan)
    seq_lens = [(1, 37 + 61 * i) for i in range(40)] + [(1300, 1300), (1, 1), (2, 2), (90, 400), (700, 2500)]
    for heads in ((28, 4), (8, 8)):
        batch = make_batch(seq_lens, heads[0], heads[1], 128, 16, dtype=torch.bfloat16, seed=45)
        check_batch(batch, [PAIR, 0], f"device plan={device_plan} heads={heads}")
    for splits in ("2", "3"):
        monkeypatch.setenv("HI_TC_SPLITS", splits)
        batch = make_batch(seq_lens[30:], 28, 4, 128, 16, dtype=torch.bfloat16, seed=46)
        check_batch(batch, [PAIR], f"device plan={device_plan} splits={splits}")


def test_host_plan_ragged_prefill_through_the_layer(monkeypatch):
    """AttentionParametersBuilder's plan (cost-sorted work items + work hint) drives the pair kernel's grid and split-KV chunk:
    a ragged chunked-prefill + decode batch must give the same answer with and without it, for forced split counts too."""
    from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
    from hydrainfer_b200.memory import KVCache
    seq_lens = [(1, 900), (37, 37), (200, 1500), (1, 64), (130, 130), (512, 2048), (1, 1)]
    for heads in ((28, 4), (8, 8)):
        batch = make_batch(seq_lens, heads[0], heads[1], 128, 16, dtype=torch.bfloat16, seed=41)
        kc, vc = batch.clone_caches()
        t = batch.n_tokens
        oracle.set_kv_cache(torch.tensor(batch.new_cache_slots, dtype=torch.int32), batch.key.view(t, heads[1], 128), batch.value.view(t, heads[1], 128), kc, vc)
        fp32 = oracle_fp32(batch, kc, vc)
        for splits in (None, "1", "3"):
            if splits is None:
                monkeypatch.delenv("HI_TC_SPLITS", raising=False)
            else:
                monkeypatch.setenv("HI_TC_SPLITS", splits)
            builder = AttentionParametersBuilder(heads[0], heads[1], 128, 16, torch.device(DEV))
            for req in batch.requests():
                builder.add_request(*req)
            builder.add_kv_cache(KVCache(batch.key_cache.to(DEV), batch.value_cache.to(DEV)))
            params = builder.build_attention_parameters()[0]
            assert params.work_items is not None and params.work_items.shape[1] == 2 and params.work_tile_tokens == 2 * (128 // (heads[0] // heads[1]))
            # the plan covers every query tile exactly once
            tiles = {(int(b), int(tl)) for b, tl in params.work_items.cpu().tolist()}
            want = {(b, tl) for b, (q, _) in enumerate(seq_lens) for tl in range(-(-q // params.work_tile_tokens))}
            assert tiles == want and len(tiles) == params.work_items.shape[0]
            layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(heads[0], heads[1], 128))
            layer.handler.path = PAIR
            out = layer(batch.query.to(DEV), batch.key.to(DEV), batch.value.to(DEV), params).o
            torch.cuda.synchronize()
            assert_close_to_fp32(out, fp32, batch.dtype, f"plan heads={heads} splits={splits}")


# ---- full-size configs: sampled oracle rows + size-independent properties -------------------------------------------------
def _sampled_check(batch_dev, sample_seqs, path, what):
    """Oracle on a subset of sequences (copied to CPU); the kernels ran on the whole batch."""
    t = batch_dev.n_tokens
    q3 = batch_dev.query.view(t, batch_dev.n_qo_heads, batch_dev.head_dim)
    out = run_attention(q3, batch_dev.key_cache, batch_dev.value_cache, batch_dev.q_cu_seq_lens, batch_dev.kv_cu_seq_lens,
                        batch_dev.block_tables, batch_dev.cu_blocks_lens, batch_dev.q_max, batch_dev.kv_max, batch_dev.head_dim, path)
    out2 = out.reshape(t, -1)
    for b in sample_seqs:
        q0, q1 = batch_dev.q_cu_seq_lens[b], batch_dev.q_cu_seq_lens[b + 1]
        table = batch_dev.per_seq_block_tables[b]
        idx = torch.tensor(table, device=DEV)
        kc = batch_dev.key_cache[idx].cpu()
        vc = batch_dev.value_cache[idx].cpu()
        kv_len = batch_dev.seq_lens[b][1]
        fp32 = oracle.paged_attention_fp32(q3[q0:q1].cpu(), kc, vc, [0, q1 - q0], [0, kv_len], torch.arange(len(table), dtype=torch.int32),
                                           [0, len(table)], batch_dev.n_qo_heads, batch_dev.n_kv_heads, batch_dev.head_dim)
        assert_close_to_fp32(out2[q0:q1], fp32, batch_dev.dtype, f"{what} seq {b} path={path}")
    return out2


def test_config2_llava_decode_full_size():
    # BASELINE config 2: 32 heads, d=128, batch 64, ctx 2048, bf16 (2.1 GB of KV)
    batch = make_batch([(1, 2048)] * 64, 32, 32, 128, 16, dtype=torch.bfloat16, device=DEV, gen_device=DEV, seed=0)
    out = _sampled_check(batch, [0, 31, 63], 0, "cfg2")
    # property: attention output is a convex combination of value rows -> bounded by the per-dim min/max of V
    assert out.float().abs().max().item() <= batch.value_cache.float().abs().max().item() + 1e-2
    # property: deterministic (split-KV merge order is fixed)
    out_again = _sampled_check(batch, [], 0, "cfg2 repeat")
    assert torch.equal(out, out_again)


def test_config3_qwen_mixed_full_size():
    # BASELINE config 3: 28/4 heads, 48 decode seqs L~U[256,8192] + 4 chunked-prefill seqs q=512
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(256, 8193, (48,), generator=g).tolist()
    seq_lens = [(1, L) for L in lens] + [(512, 512), (512, 2048), (512, 4096), (512, 8192)]
    batch = make_batch(seq_lens, 28, 4, 128, 16, dtype=torch.bfloat16, device=DEV, gen_device=DEV, seed=3)
    for path in paths_of(batch):
        _sampled_check(batch, [0, 17, 48, 49, 51], path, "cfg3")


def test_config4_qwen72b_shape_decode():
    # BASELINE config 4 per-GPU shard at N=8: 64/8 heads, 32 sequences of ctx 4096
    batch = make_batch([(1, 4096)] * 32, 64, 8, 128, 16, dtype=torch.bfloat16, device=DEV, gen_device=DEV, seed=4)
    for path in paths_of(batch):
        _sampled_check(batch, [0, 15, 31], path, "cfg4")


def test_duplicate_sequences_give_identical_rows():
    # two sequences sharing one block table and one query must produce bit-identical outputs (scheduling independence)
    batch = make_batch([(1, 900), (1, 900)], 32, 32, 128, 16, dtype=torch.bfloat16, seed=30)
    table = batch.per_seq_block_tables[0]
    tables = table + table
    q = batch.query.clone()
    q[1] = q[0]
    q3 = q.to(DEV).view(2, 32, 128)
    for path in paths_of(batch):
        out = run_attention(q3, batch.key_cache.to(DEV), batch.value_cache.to(DEV), [0, 1, 2], [0, 900, 1800], tables,
                            [0, len(table), 2 * len(table)], 1, 900, 128, path)
        assert torch.equal(out[0], out[1]), f"path={path}"


# ---- API behaviour ------------------------------------------------------------------------------------------------------
def test_unsupported_arguments_raise():
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    batch = make_batch([(1, 20)], 4, 2, 128, 16, dtype=torch.bfloat16, device=DEV, seed=1)
    q3 = batch.query.view(1, 4, 128)
    out = torch.empty_like(q3)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    args = [out, q3, batch.key_cache, batch.value_cache, i32([0, 1]), i32([0, 20]), i32(batch.block_tables), i32([0, 2]), None, 1, 20, 0.088, 0, -1, 0, 0]
    mha_varlen_fwd(*args)
    bad = list(args); bad[6] = None  # block_table None selects the un-paged form, where a 4-D paged cache is not a valid k
    with pytest.raises(RuntimeError, match=r"must be \[n_tokens"):
        mha_varlen_fwd(*bad)
    bad = list(args); bad[7] = None
    with pytest.raises(RuntimeError, match="cu_block_lens"):
        mha_varlen_fwd(*bad)
    bad = list(args); bad[12] = -30.0
    with pytest.raises(RuntimeError, match="softcap"):
        mha_varlen_fwd(*bad)
    bad = list(args); bad[8] = torch.ones(4, dtype=torch.float16, device=DEV)  # flash_api.cpp:204
    with pytest.raises(RuntimeError, match="fp32"):
        mha_varlen_fwd(*bad)
    bad = list(args); bad[8] = torch.ones(3, dtype=torch.float32, device=DEV)  # flash_api.cpp:207
    with pytest.raises(RuntimeError, match="num_heads"):
        mha_varlen_fwd(*bad)
    # the score options run on the CUDA-core path: asking for a tcgen05 path with them is an error, not a silent switch
    with pytest.raises(RuntimeError, match="CUDA-core"):
        mha_varlen_fwd(*args[:12], 30.0, -1, 0, 0, 4)
    bad = list(args); bad[4] = i32([0, 1]).long()
    with pytest.raises(RuntimeError, match="int32"):
        mha_varlen_fwd(*bad)
    bad = list(args); bad[2] = batch.key_cache.float()
    with pytest.raises(RuntimeError, match="dtype"):
        mha_varlen_fwd(*bad)
    # head_dim the kernels do not cover -> RuntimeError from the C ABI status, not a crash
    b72 = make_batch([(1, 20)], 4, 2, 72, 16, dtype=torch.bfloat16, device=DEV, seed=1)
    q72 = b72.query.view(1, 4, 72)
    with pytest.raises(RuntimeError, match="head_dim"):
        mha_varlen_fwd(torch.empty_like(q72), q72, b72.key_cache, b72.value_cache, i32([0, 1]), i32([0, 20]), i32(b72.block_tables), i32([0, 2]),
                       None, 1, 20, 0.1, 0, -1, 0, 0)
    # forcing the tile kernel on a shape it does not support is an error, not a silent fallback
    b64 = make_batch([(4, 20)], 4, 2, 64, 16, dtype=torch.bfloat16, device=DEV, seed=1)
    q64 = b64.query.view(4, 4, 64)
    with pytest.raises(RuntimeError, match="tcgen05"):
        mha_varlen_fwd(torch.empty_like(q64), q64, b64.key_cache, b64.value_cache, i32([0, 4]), i32([0, 20]), i32(b64.block_tables), i32([0, 2]),
                       None, 4, 20, 0.1, 0, -1, 0, 0, TC)


@pytest.mark.parametrize("path", [TC, PAIR])
def test_split_prefill_is_reproducible_run_to_run(path):
    """Same inputs, same launch -> same bits, whichever work item used a CTA's buffers before.  (Padding rows of a tile hold
    scores of stale Q rows; when they took part in the warp's lazy-rescale vote, the rounding of the warp's real rows depended on
    the item order of the dynamic scheduler: within tolerance, but different from run to run once a launch had many splits.)"""
    batch = make_batch([(300, 4000), (64, 64)], 28, 4, 128, 16, dtype=torch.bfloat16, seed=32).to(DEV)
    q3 = batch.query.view(batch.n_tokens, 28, 128)
    outs = [run_attention(q3, batch.key_cache, batch.value_cache, batch.q_cu_seq_lens, batch.kv_cu_seq_lens, batch.block_tables,
                          batch.cu_blocks_lens, batch.q_max, batch.kv_max, 128, path) for _ in range(5)]
    for i, o in enumerate(outs[1:], 1):
        assert torch.equal(outs[0], o), f"path {path}: run {i} differs from run 0 in {int((outs[0] != o).sum())} elements"


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("head_dim", [64, 96, 32])
def test_tcgen05_prefill_for_head_dims_below_128(head_dim, dtype):
    """The pair-tile kernel's variable-head-dim mode on PAGED caches: the reference's fused backend covers head_dim 64 / 96 / 128 /
    256 (csrc/kernel/flash_attn/src/static_switch.h:70-85, flash_api.cpp:126-136).  Prefill, chunked prefill and decode rows in
    one batch, MHA and GQA, forced PAIR path and AUTO (which must pick it for batches with prefill rows)."""
    seq_lens = [(1, 300), (130, 130), (75, 900), (1, 17), (300, 300), (40, 1300), (1, 1)]
    for heads in ((8, 8), (12, 4), (14, 2)):
        batch = make_batch(seq_lens, heads[0], heads[1], head_dim, 16, dtype=dtype, seed=61)
        check_batch(batch, [PAIR, 0], f"head_dim {head_dim} heads {heads}")
    # decode-only batch of a head dim the CUDA-core kernels do not cover: AUTO has to route it to the tile kernel
    if head_dim == 96:
        batch = make_batch([(1, 700), (1, 64), (1, 2000)], 8, 2, head_dim, 16, dtype=dtype, seed=62)
        check_batch(batch, [0], "decode-only head_dim 96")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_tcgen05_prefill_for_head_dim_256(dtype):
    """head_dim 256 (the last of the reference's fused head dims, static_switch.h:70-85) on the tcgen05 tile kernel: prefill,
    chunked prefill and decode rows, MHA / GQA / MQA, unsplit and split (few tiles -> split-KV + the 256-wide merge)."""
    seq_lens = [(1, 300), (130, 130), (75, 900), (1, 17), (300, 300), (40, 1300), (1, 1)]
    for heads in ((4, 4), (8, 2), (16, 1)):
        batch = make_batch(seq_lens, heads[0], heads[1], 256, 16, dtype=dtype, seed=71)
        check_batch(batch, [TC, 0], f"head_dim 256 heads {heads}")
    batch = make_batch([(200, 3000)], 4, 2, 256, 16, dtype=dtype, seed=72)  # 2 tiles x 2 heads: split-KV
    check_batch(batch, [TC], "head_dim 256 split")


# ---- mha_varlen_fwd's score options: softcap, sliding window, alibi (flash_api.cpp:93-111, 197-213; src/mask.h) ---------------------
OPTION_CASES = [
    # (softcap, window_left, window_right, alibi: None | "head" | "batch")
    (30.0, -1, 0, None), (0.0, 48, 0, None), (0.0, 0, 0, None), (0.0, -1, -1, None), (0.0, 32, 16, None), (0.0, -1, 7, None),
    (0.0, -1, 0, "head"), (0.0, -1, -1, "batch"), (20.0, 64, 0, "head"), (5.0, 17, 3, "batch"),
]


@pytest.mark.parametrize("case", OPTION_CASES, ids=lambda c: f"cap{c[0]}_w{c[1]}_{c[2]}_alibi{c[3]}")
@pytest.mark.parametrize("geom", [(torch.bfloat16, 8, 2, 128, 16), (torch.float16, 4, 4, 64, 16), (torch.bfloat16, 2, 1, 256, 32), (torch.float32, 4, 2, 128, 16)],
                         ids=["bf16_gqa_d128", "f16_mha_d64", "bf16_gqa_d256_bs32", "f32_d128"])
def test_score_options_match_oracle(geom, case):
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    dtype, hq, hkv, d, bs = geom
    softcap, wl, wr, alibi = case
    seq_lens = [(1, 300), (37, 37), (5, 130), (64, 200), (1, 1), (9, 9)]  # decode rows, prefills, chunked prefills, one-token sequences
    batch = make_batch(seq_lens, hq, hkv, d, bs, dtype=dtype, device=DEV, seed=3)
    t = batch.n_tokens
    q3 = batch.query.view(t, hq, d)
    out = torch.full_like(q3, float("nan"))
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    slopes = None
    if alibi == "head":
        slopes = (2.0 ** -torch.arange(1, hq + 1, dtype=torch.float32)).to(DEV)
    elif alibi == "batch":
        slopes = (torch.rand(len(seq_lens), hq, generator=torch.Generator().manual_seed(5)) * 0.2).to(DEV)
    scale = 1.0 / math.sqrt(d)
    mha_varlen_fwd(out, q3, batch.key_cache, batch.value_cache, i32(batch.q_cu_seq_lens), i32(batch.kv_cu_seq_lens), i32(batch.block_tables),
                   i32(batch.cu_blocks_lens), slopes, batch.q_max, batch.kv_max, scale, softcap, wl, wr, 0)
    ref = oracle.paged_attention_options_fp32(q3.cpu(), batch.key_cache.cpu(), batch.value_cache.cpu(), batch.q_cu_seq_lens, batch.kv_cu_seq_lens,
                                              torch.tensor(batch.block_tables), batch.cu_blocks_lens, hq, hkv, d, scale, softcap, wl, wr,
                                              None if slopes is None else slopes.cpu())
    got = out.float().cpu().reshape(t, hq * d)
    assert torch.isfinite(got).all()
    atol, rtol = (1e-4, 1e-3) if dtype == torch.float32 else (2e-2, 1e-2)  # fp32: tanh.approx carries ~2^-11 relative error
    if dtype == torch.float32 and softcap > 0:
        atol = 2e-3
    err = (got - ref).abs()
    assert bool((err <= atol + rtol * ref.abs()).all()), float(err.max())


def test_score_options_default_equals_plain_causal():
    """window (-1, 0), softcap 0, no alibi through the option-taking entry == the ordinary launch, bit for bit (same kernel path forced)."""
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    batch = make_batch([(1, 200), (20, 50)], 8, 2, 128, 16, dtype=torch.bfloat16, device=DEV, seed=4)
    t = batch.n_tokens
    q3 = batch.query.view(t, 8, 128)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    meta = (i32(batch.q_cu_seq_lens), i32(batch.kv_cu_seq_lens), i32(batch.block_tables), i32(batch.cu_blocks_lens))
    a, b = torch.empty_like(q3), torch.empty_like(q3)
    mha_varlen_fwd(a, q3, batch.key_cache, batch.value_cache, *meta, None, batch.q_max, batch.kv_max, 0.088, 0.0, -1, 0, 0, 1)
    # a window wider than any sequence masks nothing beyond causal: the option instance must agree with the plain one within rounding
    mha_varlen_fwd(b, q3, batch.key_cache, batch.value_cache, *meta, None, batch.q_max, batch.kv_max, 0.088, 0.0, 100000, 0, 0)
    assert torch.allclose(a.float(), b.float(), atol=1e-2, rtol=1e-2)


def test_score_options_through_the_ctypes_mirror_of_the_c_abi():
    """The same options through `hydrainfer_b200._lib.HiAttnArgs` (the ctypes mirror of include/hi_b200.h, ABI 6): field offsets, flag
    bits and error statuses, independent of the compiled binding."""
    import ctypes
    from hydrainfer_b200 import _lib
    dev = torch.device(DEV)
    hq, hkv, d = 8, 2, 128
    b = make_batch([(1, 300), (20, 90)], hq, hkv, d, 16, dtype=torch.bfloat16, device=DEV, seed=21)
    t = b.n_tokens
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    q3 = b.query.view(t, hq, d)
    out = torch.full_like(q3, float("nan"))
    meta = [i32(b.q_cu_seq_lens), i32(b.kv_cu_seq_lens), i32(b.block_tables), i32(b.cu_blocks_lens)]
    slopes = (2.0 ** -torch.arange(1, hq + 1, dtype=torch.float32)).to(DEV)
    scale = 1 / math.sqrt(d)

    def call(**extra):
        a = _lib.HiAttnArgs(q=q3.data_ptr(), out=out.data_ptr(), key_cache=b.key_cache.data_ptr(), value_cache=b.value_cache.data_ptr(),
                            q_row_stride=hq * d, out_row_stride=hq * d, q_cu_seq_lens=meta[0].data_ptr(), kv_cu_seq_lens=meta[1].data_ptr(),
                            block_tables=meta[2].data_ptr(), cu_blocks_lens=meta[3].data_ptr(), n_seqs=2, n_tokens=t, max_q_len=b.q_max, max_kv_len=b.kv_max,
                            n_qo_heads=hq, n_kv_heads=hkv, head_dim=d, block_size=16, n_blocks=b.key_cache.shape[0], dtype=_lib.HI_BF16,
                            softmax_scale=scale, workspace=None, workspace_bytes=0, path=0, device=0, **extra)
        return _lib.lib.hi_paged_attention(ctypes.byref(a), _lib.current_stream_ptr(dev))

    opts = _lib.HI_ATTN_OPT_WINDOW | _lib.HI_ATTN_OPT_SOFTCAP | _lib.HI_ATTN_OPT_ALIBI
    assert call(options=opts, window_left=40, window_right=0, softcap=25.0, alibi_slopes=slopes.data_ptr(), alibi_batch_stride=0) == 0, _lib.lib.hi_last_error()
    torch.cuda.synchronize()
    ref = oracle.paged_attention_options_fp32(q3.cpu(), b.key_cache.cpu(), b.value_cache.cpu(), b.q_cu_seq_lens, b.kv_cu_seq_lens, torch.tensor(b.block_tables),
                                              b.cu_blocks_lens, hq, hkv, d, scale, 25.0, 40, 0, slopes.cpu())
    got = out.float().cpu().reshape(t, hq * d)
    assert bool(((got - ref).abs() <= 2e-2 + 1e-2 * ref.abs()).all())
    # statuses: unknown bits, softcap <= 0 with its flag, alibi flag without slopes -> invalid argument; a tcgen05 path with options -> unsupported
    assert call(options=8) == -1 and b"option" in _lib.lib.hi_last_error()
    assert call(options=_lib.HI_ATTN_OPT_SOFTCAP, softcap=0.0) == -1
    assert call(options=_lib.HI_ATTN_OPT_ALIBI) == -1
    a_bad = dict(options=_lib.HI_ATTN_OPT_WINDOW, window_left=3, window_right=0)
    assert call(**a_bad) == 0
    # (the rejection of a forced tcgen05 path with options is covered through the binding in test_unsupported_arguments_raise)
