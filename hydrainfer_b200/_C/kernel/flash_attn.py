"""Mirror of the reference pybind module `hydrainfer._C.kernel.flash_attn`
(csrc/kernel/flash_attn/flash_api.cpp:216-355, stub hydrainfer/_C/kernel/flash_attn/__init__.pyi:23-40):
`mha_varlen_fwd` with the same 16 positional arguments, writing `out` in place.  Backed by hi_paged_attention
(split-KV CUDA-core kernel for decode rows, tcgen05/TMEM tile kernel for prefill)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from ... import _lib

_workspaces: dict[torch.device, Tensor] = {}


_workspace_need: dict[int, int] = {}


def _workspace(dev: torch.device, head_dim: int) -> Tensor:
    need = _workspace_need.get(head_dim)
    if need is None:
        need = _workspace_need[head_dim] = int(_lib.lib.hi_attention_workspace_bytes(0, 0, head_dim, 0))
    ws = _workspaces.get(dev)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)  # persistent, like flashinfer's workspace (executor.py:99)
        _workspaces[dev] = ws
    return ws


def mha_varlen_fwd(out: Tensor, q: Tensor, k: Tensor, v: Tensor, cu_seqlens_q: Tensor, cu_seqlens_k: Tensor,
                   block_table_: Optional[Tensor], cu_block_lens: Optional[Tensor], alibi_slopes: Optional[Tensor],
                   max_seqlen_q: int, max_seqlen_k: int, softmax_scale: float, softcap: float, window_size_left: int,
                   window_size_right: int, num_splits: int, path: int = _lib.HI_ATTN_AUTO,
                   work_items: Optional[Tensor] = None, work_tile_tokens: int = 0, qk_work_hint: int = 0) -> None:
    """Paged causal varlen attention. q/out [T, Hq, d] (row stride free), k/v caches [NB, bs, Hkv, d] contiguous,
    int32 cu_seqlens_q/k [B+1], flattened block table + cu_block_lens [B+1] (flash_api.cpp:216-232).

    `path`, `work_items` (int32 device tensor [n_items, 2] = (sequence, tile), heaviest first, tiles of `work_tile_tokens`
    query tokens) and `qk_work_hint` are optional extensions after the reference's 16 positional arguments: the host-side
    plan AttentionParametersBuilder builds next to the metadata.

    Two forms, as in the reference: paged (block_table + cu_block_lens, window (-1, 0) == causal; the attention layer,
    causal_attention.py:274-291) and un-paged (block_table None, k/v [T, Hkv, d], window (-1, -1) or (-1, 0); the vision
    encoders, multihead_attention.py:140-157, 194-211).  No alibi, no softcap, no sliding window: those raise
    RuntimeError like TORCH_CHECK."""
    if block_table_ is None:
        # the vision encoders' form (multihead_attention.py:140-157, 194-211): k/v are plain [T, H, d], window (-1, -1)
        if alibi_slopes is not None or softcap != 0 or window_size_left != -1 or window_size_right not in (-1, 0):
            raise RuntimeError("mha_varlen_fwd: alibi / softcap / sliding window are not implemented")
        return _varlen_fwd(out, q, k, v, cu_seqlens_q, cu_seqlens_k, int(max_seqlen_q), int(max_seqlen_k), float(softmax_scale),
                           causal=(window_size_right == 0))
    if cu_block_lens is None:
        raise RuntimeError("mha_varlen_fwd: a block_table needs cu_block_lens (flattened CSR block table)")
    if alibi_slopes is not None or softcap != 0 or window_size_left != -1 or window_size_right != 0:
        raise RuntimeError("mha_varlen_fwd: alibi / softcap / sliding window are not used by the paged attention layer and are not implemented")
    dev = _lib.require_cuda(out, q, k, v, cu_seqlens_q, cu_seqlens_k, block_table_, cu_block_lens)
    # (this function runs once per layer per step: the checks below are written to cost a few microseconds in all)
    q_shape, k_shape = q.shape, k.shape
    if len(q_shape) != 3 or out.shape != q_shape:
        raise RuntimeError("mha_varlen_fwd: q and out must be [n_tokens, n_heads, head_dim], contiguous over the last two dims")
    n_tokens, n_qo_heads, head_dim = q_shape
    if not (q.is_contiguous() or (q.stride(2) == 1 and q.stride(1) == head_dim)) or not (out.is_contiguous() or (out.stride(2) == 1 and out.stride(1) == head_dim)):
        raise RuntimeError("mha_varlen_fwd: q and out must be [n_tokens, n_heads, head_dim], contiguous over the last two dims")
    if len(k_shape) != 4 or k_shape != v.shape or not k.is_contiguous() or not v.is_contiguous():
        raise RuntimeError("mha_varlen_fwd: k and v must be contiguous paged caches [n_blocks, block_size, n_kv_heads, head_dim]")
    dtype = q.dtype
    if out.dtype != dtype or k.dtype != dtype or v.dtype != dtype:
        raise RuntimeError("mha_varlen_fwd: dtype mismatch")
    i32 = torch.int32
    if (cu_seqlens_q.dtype != i32 or cu_seqlens_k.dtype != i32 or block_table_.dtype != i32 or cu_block_lens.dtype != i32
            or not (cu_seqlens_q.is_contiguous() and cu_seqlens_k.is_contiguous() and block_table_.is_contiguous() and cu_block_lens.is_contiguous())):
        raise RuntimeError("mha_varlen_fwd: cu_seqlens_q, cu_seqlens_k, block_table and cu_block_lens must be contiguous int32 tensors")
    n_blocks, block_size, n_kv_heads, kd = k_shape
    if kd != head_dim or n_qo_heads % n_kv_heads != 0:
        raise RuntimeError(f"mha_varlen_fwd: head mismatch q {tuple(q_shape)} cache {tuple(k_shape)}")
    n_seqs = cu_seqlens_q.shape[0] - 1
    if cu_seqlens_k.shape[0] != n_seqs + 1 or cu_block_lens.shape[0] != n_seqs + 1:
        raise RuntimeError("mha_varlen_fwd: cu_seqlens_q, cu_seqlens_k and cu_block_lens must all have batch + 1 entries")
    if work_items is not None and (work_items.dtype != i32 or work_items.device != dev or work_items.dim() != 2 or work_items.shape[1] != 2 or not work_items.is_contiguous()):
        raise RuntimeError("mha_varlen_fwd: work_items must be a contiguous int32 device tensor of shape [n_items, 2]")
    ws = _workspace(dev, head_dim)
    row = n_qo_heads * head_dim
    args = _lib.HiAttnArgs(
        q=q.data_ptr(), out=out.data_ptr(), key_cache=k.data_ptr(), value_cache=v.data_ptr(),
        q_row_stride=q.stride(0) if n_tokens > 1 else row, out_row_stride=out.stride(0) if n_tokens > 1 else row,
        q_cu_seq_lens=cu_seqlens_q.data_ptr(), kv_cu_seq_lens=cu_seqlens_k.data_ptr(),
        block_tables=block_table_.data_ptr(), cu_blocks_lens=cu_block_lens.data_ptr(),
        n_seqs=n_seqs, n_tokens=n_tokens, max_q_len=int(max_seqlen_q), max_kv_len=int(max_seqlen_k),
        n_qo_heads=n_qo_heads, n_kv_heads=n_kv_heads, head_dim=head_dim, block_size=block_size, n_blocks=n_blocks,
        dtype=_lib.dtype_code(q.dtype), softmax_scale=float(softmax_scale),
        workspace=ws.data_ptr(), workspace_bytes=ws.numel(), path=int(path), device=dev.index or 0,
        kv_blocks_hint=int(block_table_.numel()),
        work_items=work_items.data_ptr() if work_items is not None else None, qk_work_hint=int(qk_work_hint),
        n_work_items=int(work_items.shape[0]) if work_items is not None else 0, work_tile_tokens=int(work_tile_tokens))
    _lib.check(_lib.lib.hi_paged_attention(args, _lib.current_stream_ptr(dev)))


def _varlen_fwd(out: Tensor, q: Tensor, k: Tensor, v: Tensor, cu_seqlens_q: Tensor, cu_seqlens_k: Tensor,
                max_seqlen_q: int, max_seqlen_k: int, softmax_scale: float, causal: bool) -> None:
    """Un-paged varlen attention: q/out [Tq, Hq, d], k/v [Tk, Hkv, d] (row strides free), sequence b = rows
    cu_seqlens[b] .. cu_seqlens[b + 1]; hi_varlen_attention (tcgen05 pair-tile kernel, head_dim % 8 == 0, <= 128)."""
    dev = _lib.require_cuda(out, q, k, v, cu_seqlens_q, cu_seqlens_k)
    for name, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        if t.dim() != 3 or t.stride(-1) != 1 or t.stride(-2) != t.size(-1):
            raise RuntimeError(f"mha_varlen_fwd: {name} must be [n_tokens, n_heads, head_dim], contiguous over the last two dims")
    if out.shape != q.shape or k.shape != v.shape or k.shape[-1] != q.shape[-1] or q.shape[1] % k.shape[1] != 0:
        raise RuntimeError(f"mha_varlen_fwd: shape mismatch q {tuple(q.shape)} out {tuple(out.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
    if not (q.dtype == out.dtype == k.dtype == v.dtype):
        raise RuntimeError("mha_varlen_fwd: dtype mismatch")
    if q.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError("mha_varlen_fwd: only fp16 and bf16 are supported")  # flash_api.cpp:236
    for name, t in (("cu_seqlens_q", cu_seqlens_q), ("cu_seqlens_k", cu_seqlens_k)):
        if t.dtype != torch.int32 or not t.is_contiguous() or t.dim() != 1:
            raise RuntimeError(f"mha_varlen_fwd: {name} must be a contiguous int32 vector")
    n_seqs = cu_seqlens_q.shape[0] - 1
    if cu_seqlens_k.shape[0] != n_seqs + 1:
        raise RuntimeError("mha_varlen_fwd: cu_seqlens_q and cu_seqlens_k must both have batch + 1 entries")
    n_q, n_qo_heads, head_dim = q.shape
    n_k, n_kv_heads, _ = k.shape
    ws = _workspace(dev, 128)

    def row_stride(t: Tensor) -> int:
        return t.stride(0) if t.shape[0] > 1 else t.shape[1] * t.shape[2]

    args = _lib.HiVarlenArgs(
        q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), out=out.data_ptr(),
        q_row_stride=row_stride(q), k_row_stride=row_stride(k), v_row_stride=row_stride(v), out_row_stride=row_stride(out),
        cu_seqlens_q=cu_seqlens_q.data_ptr(), cu_seqlens_k=cu_seqlens_k.data_ptr(),
        n_seqs=n_seqs, n_q_tokens=n_q, n_k_tokens=n_k, max_q_len=max_seqlen_q, max_kv_len=max_seqlen_k,
        n_qo_heads=n_qo_heads, n_kv_heads=n_kv_heads, head_dim=head_dim, dtype=_lib.dtype_code(q.dtype), causal=int(causal),
        softmax_scale=softmax_scale, device=dev.index or 0, workspace=ws.data_ptr(), workspace_bytes=ws.numel())
    _lib.check(_lib.lib.hi_varlen_attention(args, _lib.current_stream_ptr(dev)))


def last_launch_count() -> int:
    return _lib.lib.hi_last_launch_count()
