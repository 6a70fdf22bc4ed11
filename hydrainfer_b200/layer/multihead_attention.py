"""Un-paged (vision-encoder) multi-head attention with the reference's interface
(hydrainfer/layer/multihead_attention.py): MultiHeadAttentionConfig / Parameters / Output (:22-37), MultiHeadAttention
(:163-176: CLIP / SigLIP towers, equal-length image sequences) and QwenMultiHeadAttention (:258-270: Qwen2-VL tower,
packed variable-length sequences delimited by cu_seqlens).

The reference chains FlashAttention(csrc mha_varlen_fwd) -> flash_attn pip -> Torch.  Here there is one handler per
module, backed by hi_varlen_attention (the tcgen05 pair-tile kernel in its un-paged, non-causal mode); it keeps the
handler protocol (`next_handler`) so it can head the reference's own chain.  CPU tensors and `return_scores=True` (which
the fused backends of the reference also refuse, :86-87, :130-131) go to `next_handler` if one is linked and raise
otherwise: this package ships no CPU implementation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor, nn

from .._C.kernel.flash_attn import mha_varlen_fwd


@dataclass
class MultiHeadAttentionConfig:
    n_heads: int
    head_dim: int


@dataclass
class MultiHeadAttentionParameters:
    return_scores: bool = False


@dataclass
class MultiHeadAttentionOutput:
    o: Tensor
    attention_scores: Optional[Tensor]


class _CuSeqlensCache:
    """cu_seqlens = arange(0, (batch + 1) * seq_len, seq_len) is rebuilt on the device by the reference for every layer call
    (multihead_attention.py:138-139); a vision tower calls it with the same (batch, seq_len) for all of its layers."""

    def __init__(self):
        self.entries: dict[tuple, Tensor] = {}

    def get(self, batch_size: int, seq_len: int, device: torch.device) -> Tensor:
        key = (batch_size, seq_len, device)
        t = self.entries.get(key)
        if t is None:
            if len(self.entries) > 256:
                self.entries.clear()
            t = torch.arange(0, (batch_size + 1) * seq_len, seq_len, dtype=torch.int32, device=device)
            self.entries[key] = t
        return t


_cu_seqlens_cache = _CuSeqlensCache()


class B200MultiHeadAttentionHandler(nn.Module):
    """query/key/value [batch, seq_len, n_heads * head_dim] (any row stride: slices of a fused qkv projection are fine)
    -> o [batch, seq_len, hidden]; non-causal softmax(q k^T / sqrt(d)) v per (image, head)."""

    def __init__(self, config: MultiHeadAttentionConfig):
        super().__init__()
        self.n_heads = config.n_heads
        self.head_dim = config.head_dim
        self.next_handler: Optional[nn.Module] = None

    def forward(self, query: Tensor, key: Tensor, value: Tensor, params: MultiHeadAttentionParameters) -> MultiHeadAttentionOutput:
        if query.device.type != "cuda" or params.return_scores:
            if self.next_handler is not None:
                return self.next_handler(query, key, value, params)
            raise RuntimeError("hydrainfer_b200: MultiHeadAttention needs CUDA tensors and return_scores=False; no fallback handler is linked")
        batch_size, seq_len, hidden_size = query.shape
        q = self._rows(query, batch_size, seq_len)
        k = self._rows(key, batch_size, seq_len)
        v = self._rows(value, batch_size, seq_len)
        o = torch.empty((batch_size * seq_len, self.n_heads, self.head_dim), dtype=query.dtype, device=query.device)
        cu = _cu_seqlens_cache.get(batch_size, seq_len, query.device)
        try:
            mha_varlen_fwd(o, q, k, v, cu, cu, None, None, None, seq_len, seq_len, 1.0 / math.sqrt(self.head_dim), 0, -1, -1, 0)
        except RuntimeError as e:  # a geometry the kernel does not cover goes down the chain when one is linked
            if self.next_handler is not None and "hi_b200 error -2" in str(e):
                return self.next_handler(query, key, value, params)
            raise
        return MultiHeadAttentionOutput(o=o.view(batch_size, seq_len, hidden_size), attention_scores=None)

    def _rows(self, t: Tensor, batch_size: int, seq_len: int) -> Tensor:
        """[batch, seq, hidden] -> [batch * seq, n_heads, head_dim] without a copy when the rows are evenly strided."""
        if t.stride(-1) != 1 or (batch_size > 1 and t.stride(0) != seq_len * t.stride(1)):
            t = t.contiguous()
        return t.as_strided((batch_size * seq_len, self.n_heads, self.head_dim), (t.stride(1), self.head_dim, 1), t.storage_offset())


class MultiHeadAttention(nn.Module):
    def __init__(self, config: MultiHeadAttentionConfig):
        super().__init__()
        self.handlers = [B200MultiHeadAttentionHandler(config)]
        self.handler = self.handlers[0]

    def forward(self, query: Tensor, key: Tensor, value: Tensor, params: MultiHeadAttentionParameters) -> MultiHeadAttentionOutput:
        return self.handler(query, key, value, params)


class B200QwenMultiHeadAttentionHandler(nn.Module):
    """q/k/v [seq_length, n_heads, head_dim] packed over images, cu_seqlens int32 [n_images + 1] -> [seq_length, hidden]
    (QwenFlashAttentionMutliHeadAttentionHandler2.forward, multihead_attention.py:183-211).  The reference passes
    max_seqlen = seq_length (the packed total); so does this handler: it is only an upper bound for the work list."""

    def __init__(self, config: MultiHeadAttentionConfig):
        super().__init__()
        self.n_heads = config.n_heads
        self.head_dim = config.head_dim
        self.next_handler: Optional[nn.Module] = None

    def forward(self, q: Tensor, k: Tensor, v: Tensor, seq_length: int, cu_seqlens: Tensor, max_seqlen: Optional[int] = None) -> Tensor:
        if q.device.type != "cuda":
            if self.next_handler is not None:
                return self.next_handler(q, k, v, seq_length, cu_seqlens)
            raise RuntimeError("hydrainfer_b200: QwenMultiHeadAttention needs CUDA tensors; no CPU handler is linked")
        if cu_seqlens.dtype != torch.int32:
            cu_seqlens = cu_seqlens.to(torch.int32)
        attn_output = torch.empty((seq_length, self.n_heads, self.head_dim), dtype=q.dtype, device=q.device)
        bound = int(max_seqlen) if max_seqlen is not None else int(seq_length)
        try:
            mha_varlen_fwd(attn_output, q, k, v, cu_seqlens, cu_seqlens, None, None, None, bound, bound,
                           1.0 / math.sqrt(self.head_dim), 0, -1, -1, 0)
        except RuntimeError as e:
            if self.next_handler is not None and "hi_b200 error -2" in str(e):
                return self.next_handler(q, k, v, seq_length, cu_seqlens)
            raise
        return attn_output.reshape(seq_length, -1)


class QwenMultiHeadAttention(nn.Module):
    def __init__(self, config: MultiHeadAttentionConfig):
        super().__init__()
        self.handlers = [B200QwenMultiHeadAttentionHandler(config)]
        self.handler = self.handlers[0]

    def forward(self, query: Tensor, key: Tensor, value: Tensor, seq_length: int, cu_seqlens: Tensor) -> Tensor:
        return self.handler(query, key, value, seq_length, cu_seqlens)
