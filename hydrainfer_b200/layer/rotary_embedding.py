"""Rotary embedding with the reference's interface (hydrainfer/layer/rotary_embedding.py): `compute_default_inv_freq`
(:13-17) and `RotaryEmbedding(rotary_dim, max_position_embeddings, inv_freq, interleaved).forward(query, key,
position_ids)` (:136-148), one handler backed by hi_rope_append.

The table is the fused handler's `cos_sin_cache` [max_positions, 2, rotary_dim/2] (:113-115), computed on the CPU in
fp32 exactly as the reference does (t x inv_freq outer product, cos/sin) and registered as a non-persistent buffer, so
`module.to(dtype/device)` converts it the way it converts the reference's.  The kernel reproduces the reference's
rounding for either table dtype (include/hi_b200.h), so results are bit-identical to TorchRotaryEmbeddingHandler.
CPU tensors go to `next_handler` if one is linked and raise otherwise (no CPU implementation ships here)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from .._C.kernel.position_embedding import apply_rotary_pos_emb, rope_set_kv_cache


def compute_default_inv_freq(rotary_dim: int, theta: float) -> Tensor:
    assert rotary_dim % 2 == 0, "rotary_dim must be even"
    exponents = torch.arange(0, rotary_dim, 2, dtype=torch.float)
    return 1. / torch.pow(theta, exponents / rotary_dim)


class B200RotaryEmbeddingHandler(nn.Module):
    def __init__(self, rotary_dim: int, max_position_embeddings: int, inv_freq: Tensor, interleaved: bool):
        super().__init__()
        self.next_handler: Optional[nn.Module] = None
        self.rotary_dim = rotary_dim
        self.max_position_embeddings = max_position_embeddings
        self.inv_freq = inv_freq
        self.interleaved = interleaved
        t = torch.arange(max_position_embeddings, dtype=torch.float)
        freqs = torch.einsum("i,j->ij", t, inv_freq.to(torch.float).cpu())  # (max_positions, rotary_dim / 2)
        cos_sin = torch.cat([freqs.cos()[:, None, :], freqs.sin()[:, None, :]], dim=1)
        self.register_buffer(name="cos_sin_cache", tensor=cos_sin, persistent=False)

    def forward(self, query: Tensor, key: Tensor, position_ids: Tensor) -> tuple[Tensor, Tensor]:
        """In place, like the reference's fused handler (:123-133)."""
        if query.device.type != "cuda":
            if self.next_handler is not None:
                return self.next_handler(query, key, position_ids)
            raise RuntimeError("hydrainfer_b200: rotary embedding needs CUDA tensors; no CPU handler is linked")
        try:
            apply_rotary_pos_emb(query, key, position_ids, self.cos_sin_cache, self.rotary_dim, self.interleaved)
        except RuntimeError as e:  # a dtype / layout the kernel does not cover goes down the chain when one is linked
            if self.next_handler is not None and "hi_b200 error -2" in str(e):
                return self.next_handler(query, key, position_ids)
            raise
        return query, key

    def forward_and_cache(self, query: Tensor, key: Tensor, value: Tensor, position_ids: Tensor, slot_ids: Tensor,
                          key_cache: Tensor, value_cache: Tensor) -> Tensor:
        """Fused form: rotates query in place and writes the rotated key and the value into their cache slots."""
        rope_set_kv_cache(query, key, value, position_ids, self.cos_sin_cache, self.rotary_dim, self.interleaved,
                          slot_ids, key_cache, value_cache)
        return query


class RotaryEmbedding(nn.Module):
    def __init__(self, rotary_dim: int, max_position_embeddings: int, inv_freq: Tensor, interleaved: bool):
        super().__init__()
        self.handlers = nn.ModuleList([B200RotaryEmbeddingHandler(rotary_dim, max_position_embeddings, inv_freq, interleaved)])
        self.handler = self.handlers[0]

    def forward(self, query: Tensor, key: Tensor, position_ids: Tensor) -> tuple[Tensor, Tensor]:
        return self.handler(query, key, position_ids)
