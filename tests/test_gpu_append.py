"""GPU: KV append kernels vs the oracle, bit-exact.  Grid from the reference's
tests/memory/test_kv_cache.py:6-13 (block 4/8/16, heads 8/4/2/1, head 64/128/256, tokens 1/15/64/100, fp16/bf16/fp32)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, _from_np
from oracle import paged_kv_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _kernels():
    from hydrainfer_b200._C.kernel.kv_cache_kernels import set_kv_cache
    from hydrainfer_b200._C.kernel.cache_kernels import set_image_cache
    return set_kv_cache, set_image_cache


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("block_size", [4, 8, 16])
@pytest.mark.parametrize("n_tokens", [1, 15, 64, 100])
def test_set_kv_cache_grid(dtype, block_size, n_tokens):
    set_kv_cache, _ = _kernels()
    g = torch.Generator().manual_seed(n_tokens * 31 + block_size)
    num_blocks = 100
    for heads in (8, 4, 2, 1):
        for head_size in (64, 128, 256):
            kc = torch.randn(num_blocks, block_size, heads, head_size, generator=g).to(dtype)
            vc = torch.randn(num_blocks, block_size, heads, head_size, generator=g).to(dtype)
            slots = torch.randperm(num_blocks * block_size, generator=g)[:n_tokens].to(torch.int32)
            k = torch.randn(n_tokens, heads, head_size, generator=g).to(dtype)
            v = torch.randn(n_tokens, heads, head_size, generator=g).to(dtype)
            kc_ref, vc_ref = kc.clone(), vc.clone()
            oracle.set_kv_cache(slots, k, v, kc_ref, vc_ref)
            kc_d, vc_d = kc.to(DEV), vc.to(DEV)
            set_kv_cache(slots.to(DEV), k.to(DEV), v.to(DEV), kc_d, vc_d)
            assert torch.equal(kc_d.cpu(), kc_ref) and torch.equal(vc_d.cpu(), vc_ref), (heads, head_size)


def test_set_kv_cache_strided_fused_qkv_rows():
    # keys/values are column slices of one fused qkv projection (model_forward.py:69-73): row stride > row width
    set_kv_cache, _ = _kernels()
    g = torch.Generator().manual_seed(3)
    hq, hkv, d, bs, nb, t = 28, 4, 128, 16, 40, 37
    qkv = torch.randn(t, (hq + 2 * hkv) * d, generator=g).to(torch.bfloat16)
    kc = torch.randn(nb, bs, hkv, d, generator=g).to(torch.bfloat16)
    vc = torch.randn(nb, bs, hkv, d, generator=g).to(torch.bfloat16)
    slots = torch.randperm(nb * bs, generator=g)[:t].to(torch.int32)
    kc_ref, vc_ref = kc.clone(), vc.clone()
    k_cpu = qkv[:, hq * d:(hq + hkv) * d].view(t, hkv, d)
    v_cpu = qkv[:, (hq + hkv) * d:].view(t, hkv, d)
    oracle.set_kv_cache(slots, k_cpu, v_cpu, kc_ref, vc_ref)
    qkv_d = qkv.to(DEV)
    k_d = qkv_d[:, hq * d:(hq + hkv) * d].view(t, hkv, d)
    v_d = qkv_d[:, (hq + hkv) * d:].view(t, hkv, d)
    assert not k_d.is_contiguous()
    kc_d, vc_d = kc.to(DEV), vc.to(DEV)
    set_kv_cache(slots.to(DEV), k_d, v_d, kc_d, vc_d)
    assert torch.equal(kc_d.cpu(), kc_ref) and torch.equal(vc_d.cpu(), vc_ref)


def test_set_kv_cache_odd_alignment_falls_back_to_narrow_vectors():
    # head_dim 8 * 1 head fp16 = 16-byte rows, but a source view offset by one element is only 2-byte aligned
    set_kv_cache, _ = _kernels()
    g = torch.Generator().manual_seed(4)
    buf = torch.randn(10 * 24 + 1, generator=g).to(torch.float16)
    k_cpu = torch.as_strided(buf, (10, 1, 8), (24, 8, 1), storage_offset=1)
    kc = torch.zeros(4, 4, 1, 8, dtype=torch.float16)
    vc = torch.zeros(4, 4, 1, 8, dtype=torch.float16)
    slots = torch.tensor([3, 0, 7, 9, 15, 1, 2, 4, 8, 11], dtype=torch.int32)
    kc_ref, vc_ref = kc.clone(), vc.clone()
    oracle.set_kv_cache(slots, k_cpu, k_cpu, kc_ref, vc_ref)
    buf_d = buf.to(DEV)
    k_d = torch.as_strided(buf_d, (10, 1, 8), (24, 8, 1), storage_offset=1)
    kc_d, vc_d = kc.to(DEV), vc.to(DEV)
    set_kv_cache(slots.to(DEV), k_d, k_d, kc_d, vc_d)
    assert torch.equal(kc_d.cpu(), kc_ref) and torch.equal(vc_d.cpu(), vc_ref)


def test_set_kv_cache_empty_batch_is_a_noop():
    set_kv_cache, _ = _kernels()
    kc = torch.randn(2, 16, 2, 64, device=DEV, dtype=torch.float16)
    vc = kc.clone()
    before = kc.clone()
    set_kv_cache(torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros(0, 2, 64, device=DEV, dtype=torch.float16),
                 torch.zeros(0, 2, 64, device=DEV, dtype=torch.float16), kc, vc)
    torch.cuda.synchronize()
    assert torch.equal(kc, before)


def test_set_kv_cache_golden(golden_attention):
    set_kv_cache, _ = _kernels()
    g = golden_attention
    t = g.key.shape[0]
    kc, vc = g.key_cache.to(DEV), g.value_cache.to(DEV)
    set_kv_cache(torch.tensor(g.new_cache_slots, dtype=torch.int32, device=DEV), g.key.to(DEV).view(t, g.n_kv_heads, g.head_dim),
                 g.value.to(DEV).view(t, g.n_kv_heads, g.head_dim), kc, vc)
    owned = torch.tensor(g.owned_blocks)
    assert torch.equal(kc.cpu()[owned], g.ref_key_cache_owned) and torch.equal(vc.cpu()[owned], g.ref_value_cache_owned)
    mask = torch.ones(g.n_blocks, dtype=torch.bool)
    mask[owned] = False
    assert torch.equal(kc.cpu()[mask], g.key_cache[mask]) and torch.equal(vc.cpu()[mask], g.value_cache[mask])


def test_set_image_cache_golden():
    _, set_image_cache = _kernels()
    z = np.load(GOLDEN / "image_cache.npz")
    n_blocks, bs, heads, d = (int(v) for v in z["geometry"])
    cache = _from_np(z["cache"], torch.float16)
    tokens = _from_np(z["tokens"], torch.float16)
    slots = torch.from_numpy(z["slots"])
    ref = cache.clone()
    oracle.set_image_cache(slots, tokens, ref)
    cache_d = cache.to(DEV)
    set_image_cache(slots.to(DEV), tokens.to(DEV), cache_d)
    assert torch.equal(cache_d.cpu(), ref)
    assert int(cache_d.cpu().view(torch.int16).to(torch.int64).sum()) == int(z["ref_checksum"][0])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_set_image_cache_llava_geometry(dtype):
    # LLaVA image pool: block_size 576, 32 heads x 128; one full image = 576 tokens
    _, set_image_cache = _kernels()
    from hydrainfer_b200.memory import TokenCache
    g = torch.Generator().manual_seed(11)
    cache = torch.randn(3, 576, 32, 128, generator=g).to(dtype)
    tokens = torch.randn(576, 32, 128, generator=g).to(dtype)
    slots = (torch.arange(576) + 576).to(torch.int32)  # block 1
    ref = cache.clone()
    oracle.set_image_cache(slots, tokens, ref)
    cache_d = cache.to(DEV)
    TokenCache([cache_d]).set_caches(slots.to(DEV), [tokens.to(DEV)])
    assert torch.equal(cache_d.cpu(), ref)


def test_layout_violations_raise():
    set_kv_cache, set_image_cache = _kernels()
    kc = torch.zeros(2, 16, 2, 64, device=DEV, dtype=torch.float16)
    k = torch.zeros(3, 2, 64, device=DEV, dtype=torch.float16)
    slots = torch.zeros(3, dtype=torch.int32, device=DEV)
    with pytest.raises(RuntimeError, match="contiguous over"):
        set_kv_cache(slots, k.transpose(1, 2).contiguous().transpose(1, 2), k, kc, kc.clone())
    with pytest.raises(RuntimeError, match="dtype mismatch"):
        set_kv_cache(slots, k.float(), k.float(), kc, kc.clone())
    with pytest.raises(RuntimeError, match="int32"):
        set_kv_cache(slots.long(), k, k, kc, kc.clone())
    with pytest.raises(RuntimeError, match="contiguous"):
        set_image_cache(slots, k, kc[:, ::2])
