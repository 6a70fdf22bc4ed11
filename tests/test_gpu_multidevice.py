"""GPU: per-device and per-thread state of the library.

  * one process driving TWO GPUs: the >48 KiB dynamic shared-memory opt-in (cudaFuncSetAttribute) and the SM count are per
    device; a launch on cuda:1 must work after cuda:0 configured the kernels, and must not change the caller's current device;
  * two THREADS on two streams of one GPU (the engine thread and the image-embed thread of SURVEY §8b, executor.py:238-285):
    attention calls in flight on different streams use different workspaces (partials + work counter);
  * the workspace contract: hi_attention_workspace_bytes() covers what the split rule asks for and a smaller workspace is
    HI_ERR_WORKSPACE, never a quietly different split count.
"""
import ctypes
import math
import threading

import pytest
import torch

from hydrainfer_b200.workloads import make_batch
from oracle import paged_kv_oracle as oracle

pytestmark = pytest.mark.gpu
needs_2gpu = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
SIMT, TC, DEC, PAIR = 1, 2, 3, 4


def _fp32(batch, kc, vc):
    return oracle.paged_attention_fp32(batch.query.view(-1, batch.n_qo_heads, batch.head_dim), kc, vc, batch.q_cu_seq_lens, batch.kv_cu_seq_lens,
                                       torch.tensor(batch.block_tables, dtype=torch.int32), batch.cu_blocks_lens, batch.n_qo_heads, batch.n_kv_heads, batch.head_dim)


def _layer_forward(batch, dev, path):
    from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
    from hydrainfer_b200.memory import KVCache
    kc, vc = batch.key_cache.to(dev), batch.value_cache.to(dev)
    builder = AttentionParametersBuilder(batch.n_qo_heads, batch.n_kv_heads, batch.head_dim, batch.block_size, torch.device(dev))
    for req in batch.requests():
        builder.add_request(*req)
    builder.add_kv_cache(KVCache(kc, vc))
    params = builder.build_attention_parameters()[0]
    layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(batch.n_qo_heads, batch.n_kv_heads, batch.head_dim))
    layer.handler.path = path
    out = layer(batch.query.to(dev), batch.key.to(dev), batch.value.to(dev), params).o
    return out, kc, vc


@needs_2gpu
@pytest.mark.parametrize("path", [SIMT, TC, DEC, PAIR, 0])
def test_attention_on_a_second_device_of_the_same_process(path):
    batch = make_batch([(1, 700), (40, 170), (1, 17), (130, 130)], 28, 4, 128, 16, dtype=torch.bfloat16, seed=21)
    kc_ref, vc_ref = batch.clone_caches()
    t = batch.n_tokens
    oracle.set_kv_cache(torch.tensor(batch.new_cache_slots, dtype=torch.int32), batch.key.view(t, 4, 128), batch.value.view(t, 4, 128), kc_ref, vc_ref)
    want = _fp32(batch, kc_ref, vc_ref)
    assert torch.cuda.current_device() == 0
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        out, kc, vc = _layer_forward(batch, dev, path)
        torch.cuda.synchronize(dev)
        assert torch.cuda.current_device() == 0, f"a launch on {dev} changed the caller's current device"
        assert torch.equal(kc.cpu(), kc_ref) and torch.equal(vc.cpu(), vc_ref), f"{dev}: KV append not bit-exact"
        err = (out.float().cpu() - want).abs()
        assert bool((err <= 2e-2 + 1e-2 * want.abs()).all()), f"{dev} path {path}: max |err| {err.max().item():.3e}"


@needs_2gpu
def test_rope_scatter_and_vision_attention_on_a_second_device():
    from hydrainfer_b200._C.kernel.cache_kernels import get_image_cache, set_image_cache
    from hydrainfer_b200.layer import MultiHeadAttentionConfig, QwenMultiHeadAttention
    dev = torch.device("cuda:1")
    g = torch.Generator().manual_seed(4)
    cache = torch.zeros(3, 64, 2, 64, dtype=torch.float16, device=dev)
    tokens = torch.randn(50, 2, 64, generator=g).to(torch.float16).to(dev)
    slots = torch.randperm(3 * 64, generator=g)[:50].to(torch.int32).to(dev)
    set_image_cache(slots, tokens, cache)
    back = get_image_cache(slots, cache)
    torch.cuda.synchronize(dev)
    assert torch.equal(back.view(50, 2, 64), tokens) and torch.cuda.current_device() == 0
    vq, vk, vv = (torch.randn(230, 4, 80, generator=g).to(torch.bfloat16) for _ in range(3))
    cu = [0, 100, 101, 230]
    out = QwenMultiHeadAttention(MultiHeadAttentionConfig(4, 80))(vq.to(dev), vk.to(dev), vv.to(dev), 230, torch.tensor(cu, dtype=torch.int32, device=dev))
    torch.cuda.synchronize(dev)
    want = oracle.varlen_attention_fp32(vq, vk, vv, cu, cu)
    err = (out.float().cpu() - want).abs()
    assert bool((err <= 2e-2 + 1e-2 * want.abs()).all())


def test_two_threads_on_two_streams_do_not_share_scratch():
    """Engine thread: split-KV decode steps (partials + merge).  Image-embed thread: set_image_cache + a split prefill on its own
    stream.  Both loop concurrently; every result must equal the one computed alone."""
    from hydrainfer_b200._C.kernel.cache_kernels import set_image_cache
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    dev = torch.device("cuda:0")
    dec = make_batch([(1, 3000)] * 4, 32, 32, 128, 16, dtype=torch.bfloat16, seed=31).to(dev)          # few rows: split-KV + merge
    pre = make_batch([(300, 4000), (64, 64)], 28, 4, 128, 16, dtype=torch.bfloat16, seed=32).to(dev)   # few tiles: split pair kernel
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)

    def args_of(b):
        return (b.query.view(b.n_tokens, b.n_qo_heads, 128), b.key_cache, b.value_cache, i32(b.q_cu_seq_lens), i32(b.kv_cu_seq_lens), i32(b.block_tables),
                i32(b.cu_blocks_lens), None, b.q_max, b.kv_max, 1 / math.sqrt(128), 0, -1, 0, 0)

    dec_args, pre_args = args_of(dec), args_of(pre)
    dec_ref, pre_ref = torch.empty_like(dec_args[0]), torch.empty_like(pre_args[0])
    mha_varlen_fwd(dec_ref, *dec_args)
    mha_varlen_fwd(pre_ref, *pre_args)
    torch.cuda.synchronize()
    img_cache = torch.zeros(2, 576, 8, 128, dtype=torch.bfloat16, device=dev)
    img_tokens = torch.randn(576, 8, 128, device=dev).to(torch.bfloat16)
    img_slots = torch.arange(576, dtype=torch.int32, device=dev)
    errors, n_iter = [], 60

    def engine():
        try:
            s = torch.cuda.Stream(dev)
            with torch.cuda.stream(s):
                for _ in range(n_iter):
                    out = torch.empty_like(dec_ref)
                    mha_varlen_fwd(out, *dec_args)
                    s.synchronize()
                    if not torch.equal(out, dec_ref):
                        errors.append("decode result changed while another stream was running attention")
                        return
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def image_embed():
        try:
            s = torch.cuda.Stream(dev)
            with torch.cuda.stream(s):
                for _ in range(n_iter):
                    set_image_cache(img_slots, img_tokens, img_cache)
                    out = torch.empty_like(pre_ref)
                    mha_varlen_fwd(out, *pre_args)
                    s.synchronize()
                    if not torch.equal(out, pre_ref):
                        errors.append("prefill result changed while another stream was running attention")
                        return
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=engine), threading.Thread(target=image_embed)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors
    assert torch.equal(img_cache[0], img_tokens)


def test_workspace_contract():
    """Too small a workspace for the split count the rule picks is HI_ERR_WORKSPACE; the advertised size is enough."""
    from hydrainfer_b200 import _lib
    dev = torch.device("cuda:0")
    need_any = _lib.lib.hi_attention_workspace_bytes(0, 0, 128, 0)
    need_small = _lib.lib.hi_attention_workspace_bytes(4, 32, 128, 3000)
    need_big = _lib.lib.hi_attention_workspace_bytes(2048, 28, 128, 8192)
    assert 0 < need_small <= need_big <= need_any and need_small % 256 == 0
    b = make_batch([(1, 3000)] * 4, 28, 4, 128, 16, dtype=torch.bfloat16, seed=33).to(dev)  # 16 row-heads of work: every path splits
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    q3 = b.query.view(4, 28, 128)
    out = torch.empty_like(q3)
    meta = [i32(b.q_cu_seq_lens), i32(b.kv_cu_seq_lens), i32(b.block_tables), i32(b.cu_blocks_lens)]

    def call(path, ws):
        a = _lib.HiAttnArgs(q=q3.data_ptr(), out=out.data_ptr(), key_cache=b.key_cache.data_ptr(), value_cache=b.value_cache.data_ptr(),
                            q_row_stride=28 * 128, out_row_stride=28 * 128, q_cu_seq_lens=meta[0].data_ptr(), kv_cu_seq_lens=meta[1].data_ptr(),
                            block_tables=meta[2].data_ptr(), cu_blocks_lens=meta[3].data_ptr(), n_seqs=4, n_tokens=4, max_q_len=1, max_kv_len=3000,
                            n_qo_heads=28, n_kv_heads=4, head_dim=128, block_size=16, n_blocks=b.n_blocks, dtype=_lib.HI_BF16,
                            softmax_scale=1 / math.sqrt(128), workspace=ws.data_ptr() if ws is not None else None,
                            workspace_bytes=ws.numel() if ws is not None else 0, path=path, device=0, kv_blocks_hint=len(b.block_tables))
        return _lib.lib.hi_paged_attention(ctypes.byref(a), _lib.current_stream_ptr(dev))

    tiny = torch.empty(4096, dtype=torch.uint8, device=dev)
    enough = torch.empty(_lib.lib.hi_attention_workspace_bytes(4, 28, 128, 3000), dtype=torch.uint8, device=dev)
    for path in (SIMT, TC, DEC, PAIR):
        assert call(path, tiny) == -4, f"path {path}: a 4 KiB workspace must be rejected (HI_ERR_WORKSPACE), got {_lib.lib.hi_last_error()}"
        assert b"workspace" in _lib.lib.hi_last_error()
        assert call(path, None) == -4
        assert call(path, enough) == 0, _lib.lib.hi_last_error()
    torch.cuda.synchronize()


def test_rows_without_visible_keys_get_zeros():
    """kv_len < q_len is malformed metadata; the split-KV kernel writes zeros for such rows instead of leaving `out` as it was."""
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    dev = torch.device("cuda:0")
    b = make_batch([(1, 40), (1, 33)], 4, 4, 128, 16, dtype=torch.float16, seed=34).to(dev)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    q3 = b.query.view(2, 4, 128)
    out = torch.full_like(q3, float("nan"))
    kv_cu = [0, 0, 33]  # sequence 0 claims no cached token at all
    mha_varlen_fwd(out, q3, b.key_cache, b.value_cache, i32(b.q_cu_seq_lens), i32(kv_cu), i32(b.block_tables), i32(b.cu_blocks_lens), None, 1, 40,
                   1 / math.sqrt(128), 0, -1, 0, 0, SIMT)
    torch.cuda.synchronize()
    assert bool((out[0] == 0).all()) and bool(torch.isfinite(out[1]).all())
