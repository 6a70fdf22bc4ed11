"""GPU: un-paged varlen attention (the vision-encoder form of mha_varlen_fwd) through the reference-shaped
modules MultiHeadAttention / QwenMultiHeadAttention (which call the C ABI's hi_varlen_attention) vs the CPU oracle.

Tolerance as for the paged path (BASELINE.json north_star): |ours - fp32 recompute| <= 2e-2 + 1e-2 * |fp32 recompute| for the
16-bit dtypes.  Fixtures: the reference's Torch handlers frozen by oracle/make_golden.py."""
import math

import numpy as np
import pytest
import torch

from conftest import GOLDEN, _TORCH_DTYPES, _from_np
from oracle import paged_kv_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL, RTOL = 2e-2, 1e-2


def assert_close(out: torch.Tensor, fp32: torch.Tensor, what: str):
    out = out.float().cpu().reshape(fp32.shape)
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    err = (out - fp32).abs()
    worst = (err - (ATOL + RTOL * fp32.abs())).max().item()
    assert worst <= 0, f"{what}: max |err| {err.max().item():.4e} exceeds atol {ATOL} + rtol {RTOL} by {worst:.3e}"


def varlen(q, k, v, cu_q, cu_k, max_q, max_k, causal=False):
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    out = torch.empty((q.shape[0], q.shape[1], q.shape[2]), dtype=q.dtype, device=q.device)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)
    # exactly the reference's call (multihead_attention.py:140-157): block_table None, window (-1, -1) = not causal
    mha_varlen_fwd(out, q, k, v, i32(cu_q), i32(cu_k), None, None, None, max_q, max_k, 1.0 / math.sqrt(q.shape[-1]), 0, -1, 0 if causal else -1, 0)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("path", sorted(p for p in GOLDEN.glob("mha_*.npz") if "qwen" not in p.stem), ids=lambda p: p.stem)
def test_multi_head_attention_golden(path):
    from hydrainfer_b200.layer import MultiHeadAttention, MultiHeadAttentionConfig, MultiHeadAttentionParameters
    z = np.load(path)
    dtype = _TORCH_DTYPES[str(z["dtype"])]
    batch, seq_len, heads, d = (int(x) for x in z["geometry"])
    q, k, v, ref = (_from_np(z[n], dtype) for n in ("query", "key", "value", "ref_out"))
    cu = list(range(0, (batch + 1) * seq_len, seq_len))
    fp32 = oracle.varlen_attention_fp32(q.view(-1, heads, d), k.view(-1, heads, d), v.view(-1, heads, d), cu, cu)
    module = MultiHeadAttention(MultiHeadAttentionConfig(n_heads=heads, head_dim=d))
    out = module(q.to(DEV), k.to(DEV), v.to(DEV), MultiHeadAttentionParameters(return_scores=False))
    torch.cuda.synchronize()
    assert out.attention_scores is None and out.o.shape == q.shape and out.o.dtype == dtype
    assert_close(out.o, fp32, path.stem)
    # and against the reference's own (dtype-rounded) output: two roundings of the same number
    assert (out.o.float().cpu() - ref.float()).abs().max().item() <= 2 * ATOL


@pytest.mark.parametrize("path", sorted(p for p in GOLDEN.glob("mha_*.npz") if "qwen" in p.stem), ids=lambda p: p.stem)
def test_qwen_multi_head_attention_golden(path):
    from hydrainfer_b200.layer import MultiHeadAttentionConfig, QwenMultiHeadAttention
    z = np.load(path)
    dtype = _TORCH_DTYPES[str(z["dtype"])]
    _, total, heads, d = (int(x) for x in z["geometry"])
    q, k, v, ref = (_from_np(z[n], dtype) for n in ("query", "key", "value", "ref_out"))
    cu = [int(x) for x in z["cu_seqlens"]]
    fp32 = oracle.varlen_attention_fp32(q, k, v, cu, cu)
    module = QwenMultiHeadAttention(MultiHeadAttentionConfig(n_heads=heads, head_dim=d))
    out = module(q.to(DEV), k.to(DEV), v.to(DEV), total, torch.tensor(cu, dtype=torch.int32, device=DEV))
    torch.cuda.synchronize()
    assert out.shape == (total, heads * d) and out.dtype == dtype
    assert_close(out, fp32, path.stem)
    assert (out.float().cpu() - ref.float()).abs().max().item() <= 3 * ATOL


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("head_dim", [64, 72, 80, 96, 128, 32, 8])
def test_head_dims(head_dim, dtype):
    g = torch.Generator().manual_seed(head_dim)
    heads, lens = 3, [1, 63, 64, 65, 130, 257]
    total = sum(lens)
    q, k, v = (torch.randn(total, heads, head_dim, generator=g).to(dtype) for _ in range(3))
    cu = [0] + list(np.cumsum(lens))
    fp32 = oracle.varlen_attention_fp32(q, k, v, cu, cu)
    out = varlen(q.to(DEV), k.to(DEV), v.to(DEV), cu, cu, max(lens), max(lens))
    assert_close(out, fp32, f"d={head_dim} {dtype}")


@pytest.mark.parametrize("seq_len", [1, 2, 127, 128, 129, 255, 256, 257, 577, 1024])
def test_equal_length_images_through_fused_qkv(seq_len):
    """CLIP-style: q, k, v are column slices of one fused projection output (row stride 3 * hidden), 16 heads of 64."""
    from hydrainfer_b200.layer import MultiHeadAttention, MultiHeadAttentionConfig, MultiHeadAttentionParameters
    heads, d, batch = 16, 64, 2
    g = torch.Generator().manual_seed(seq_len)
    qkv = torch.randn(batch, seq_len, 3 * heads * d, generator=g).to(torch.bfloat16)
    q, k, v = qkv.split(heads * d, dim=-1)
    cu = list(range(0, (batch + 1) * seq_len, seq_len))
    fp32 = oracle.varlen_attention_fp32(q.reshape(-1, heads, d), k.reshape(-1, heads, d), v.reshape(-1, heads, d), cu, cu)
    qkv_d = qkv.to(DEV)
    qd, kd, vd = qkv_d.split(heads * d, dim=-1)
    out = MultiHeadAttention(MultiHeadAttentionConfig(heads, d))(qd, kd, vd, MultiHeadAttentionParameters()).o
    torch.cuda.synchronize()
    assert_close(out, fp32, f"seq_len={seq_len}")


@pytest.mark.parametrize("group", [1, 2, 7])
def test_causal_and_grouped_unpaged(group):
    """window (-1, 0): the bottom-right aligned causal mask with q_len != kv_len, and GQA sharing of un-paged K/V."""
    hkv, d = 2, 128
    lens = [(1, 70), (40, 40), (33, 100), (128, 129), (5, 5)]
    g = torch.Generator().manual_seed(group)
    q = torch.randn(sum(a for a, _ in lens), hkv * group, d, generator=g).to(torch.bfloat16)
    k, v = (torch.randn(sum(b for _, b in lens), hkv, d, generator=g).to(torch.bfloat16) for _ in range(2))
    cu_q = [0] + list(np.cumsum([a for a, _ in lens]))
    cu_k = [0] + list(np.cumsum([b for _, b in lens]))
    for causal in (True, False):
        fp32 = oracle.varlen_attention_fp32(q, k, v, cu_q, cu_k, causal=causal)
        out = varlen(q.to(DEV), k.to(DEV), v.to(DEV), cu_q, cu_k, 128, 129, causal=causal)
        assert_close(out, fp32, f"group={group} causal={causal}")


def test_garbage_after_the_last_sequence_is_ignored():
    """Rows of k / v beyond cu_seqlens[-1] (and the neighbouring sequence's rows inside a 64-key step) never leak: NaN there."""
    heads, d, lens = 2, 80, [70, 50]
    g = torch.Generator().manual_seed(3)
    total = sum(lens)
    q, k, v = (torch.randn(total, heads, d, generator=g).to(torch.float16) for _ in range(3))
    cu = [0, 70, 120]
    fp32 = oracle.varlen_attention_fp32(q, k, v, cu, cu)
    pad = torch.full((40, heads, d), float("nan"), dtype=torch.float16)
    kd, vd = torch.cat([k, pad]).to(DEV), torch.cat([v, pad]).to(DEV)
    out = varlen(q.to(DEV), kd, vd, cu, cu, 70, 70)
    assert_close(out, fp32, "nan tail")


def test_llava_clip_full_size_properties():
    """LLaVA-1.5 vision tower shape at serving size: 16 images x 577 tokens x 16 heads x 64 (bf16).  Sampled images against the
    oracle, plus a size-independent property of non-causal attention: permuting a sequence's keys (with their values) leaves the output unchanged."""
    batch, seq_len, heads, d = 16, 577, 16, 64
    g = torch.Generator(device=DEV).manual_seed(1)
    q, k, v = (torch.randn(batch * seq_len, heads, d, generator=g, device=DEV).to(torch.bfloat16) for _ in range(3))
    cu = list(range(0, (batch + 1) * seq_len, seq_len))
    out = varlen(q, k, v, cu, cu, seq_len, seq_len)
    assert torch.isfinite(out.float()).all()
    for b in (0, 7, 15):
        sl = slice(b * seq_len, (b + 1) * seq_len)
        fp32 = oracle.varlen_attention_fp32(q[sl].cpu(), k[sl].cpu(), v[sl].cpu(), [0, seq_len], [0, seq_len])
        assert_close(out[sl], fp32, f"image {b}")
    perm = torch.cat([b * seq_len + torch.randperm(seq_len, device=DEV) for b in range(batch)])
    out_p = varlen(q, k[perm], v[perm], cu, cu, seq_len, seq_len)
    assert (out_p.float() - out.float()).abs().max().item() <= 2e-2
    # deterministic
    assert torch.equal(varlen(q, k, v, cu, cu, seq_len, seq_len), out)


def test_qwen2vl_packed_full_size():
    """Qwen2-VL vision tower shape: 16 heads of 80, packed images of very different sizes (64 .. 4096 patches), max_seqlen passed as
    the packed total exactly like the reference (multihead_attention.py:203-204)."""
    from hydrainfer_b200.layer import MultiHeadAttentionConfig, QwenMultiHeadAttention
    heads, d = 16, 80
    lens = [4096, 64, 1024, 300, 2500, 16, 784]
    total = sum(lens)
    g = torch.Generator(device=DEV).manual_seed(2)
    q, k, v = (torch.randn(total, heads, d, generator=g, device=DEV).to(torch.bfloat16) for _ in range(3))
    cu = [0] + [int(x) for x in np.cumsum(lens)]
    out = QwenMultiHeadAttention(MultiHeadAttentionConfig(heads, d))(q, k, v, total, torch.tensor(cu, dtype=torch.int32, device=DEV))
    torch.cuda.synchronize()
    assert out.shape == (total, heads * d)
    for b in (1, 3, 5, 6):
        sl = slice(cu[b], cu[b + 1])
        fp32 = oracle.varlen_attention_fp32(q[sl].cpu(), k[sl].cpu(), v[sl].cpu(), [0, lens[b]], [0, lens[b]])
        assert_close(out[sl], fp32, f"image {b}")
    # the 4096-patch image: a few rows against an fp32 recompute on the GPU (the CPU oracle at this size takes too long)
    sl = slice(0, 4096)
    rows = torch.tensor([0, 1, 2047, 4095], device=DEV)
    s = torch.einsum("qhd,khd->hqk", q[sl][rows].float(), k[sl].float()) / math.sqrt(d)
    ref = torch.einsum("hqk,khd->qhd", torch.softmax(s, dim=-1), v[sl].float()).reshape(len(rows), -1)
    assert_close(out[sl][rows], ref.cpu(), "4096-patch image")


def test_rejects_what_the_reference_rejects():
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    q = torch.randn(8, 2, 64, device=DEV)
    cu = torch.tensor([0, 8], dtype=torch.int32, device=DEV)
    with pytest.raises(RuntimeError):  # fp32: flash_api.cpp:236
        mha_varlen_fwd(torch.empty_like(q), q, q, q, cu, cu, None, None, None, 8, 8, 0.125, 0, -1, -1, 0)
    h = q.half()
    with pytest.raises(RuntimeError):  # sliding window
        mha_varlen_fwd(torch.empty_like(h), h, h, h, cu, cu, None, None, None, 8, 8, 0.125, 0, 4, -1, 0)
    with pytest.raises(RuntimeError):  # CPU tensors: no fallback
        mha_varlen_fwd(torch.empty_like(h.cpu()), h.cpu(), h.cpu(), h.cpu(), cu.cpu(), cu.cpu(), None, None, None, 8, 8, 0.125, 0, -1, -1, 0)


def test_get_image_cache_bit_exact():
    from hydrainfer_b200._C.kernel.cache_kernels import get_image_cache, set_image_cache
    for dtype, (nb, bs, heads, d) in ((torch.float16, (3, 576, 32, 128)), (torch.bfloat16, (5, 16, 3, 72)), (torch.float32, (2, 8, 1, 5))):
        g = torch.Generator().manual_seed(nb)
        cache = torch.randn(nb, bs, heads, d, generator=g).to(dtype)
        slots = torch.randperm(nb * bs, generator=g)[: max(1, nb * bs // 3)].to(torch.int32)
        ref = oracle.get_image_cache(slots, cache)
        out = get_image_cache(slots.to(DEV), cache.to(DEV))
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), ref)
        # scatter -> gather round trip
        tokens = torch.randn(slots.shape[0], heads, d, generator=g).to(dtype)
        cache_d = cache.to(DEV)
        set_image_cache(slots.to(DEV), tokens.to(DEV), cache_d)
        assert torch.equal(get_image_cache(slots.to(DEV), cache_d).cpu(), tokens.view(slots.shape[0], -1))
