"""ROPECausalGroupedQueryPageAttention with the reference's interface (hydrainfer/model/model_forward.py:40-86): the
caller of the attention layer in every language model (llama.py:38, qwen2_vl.py, deepseek_v3.py:178).

forward(hidden_states, position_ids, attention_param): optional q/k/v (or fused qkv) projections, rotary embedding,
KV append, paged attention, optional output projection.  The reference runs rotary (1 launch, in place), set_kv_cache
(1 launch) and the attention backend; here rotation and append are ONE launch (hi_rope_append): q is rotated in place
and the rotated k and v go straight to their cache slots, then the attention handler runs on the appended cache.
The projections stay the caller's `nn.Linear`s (cuBLAS), exactly as in the reference."""
from __future__ import annotations

from typing import Optional

from torch import Tensor, nn

from ..layer.causal_attention import AttentionParameters, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
from ..layer.rotary_embedding import RotaryEmbedding


class ROPECausalGroupedQueryPageAttention:
    def __init__(self, n_qo_heads: int, n_kv_heads: int, head_dim: int, rotary_emb: Optional[RotaryEmbedding] = None,
                 q_proj: Optional[nn.Linear] = None, k_proj: Optional[nn.Linear] = None, v_proj: Optional[nn.Linear] = None,
                 qkv_proj: Optional[nn.Linear] = None, o_proj: Optional[nn.Linear] = None):
        self.q_proj = q_proj
        self.k_proj = k_proj
        self.v_proj = v_proj
        self.o_proj = o_proj
        self.qkv_proj = qkv_proj
        self.rotary_emb = rotary_emb
        self.n_qo_heads = n_qo_heads
        self.n_kv_heads = n_kv_heads
        self.head_dim = head_dim
        self.attention = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(n_qo_heads=n_qo_heads, n_kv_heads=n_kv_heads, head_dim=head_dim))
        self.fuse_rope_append = True  # False: rotary and set_kv_cache as two launches, like the reference (tests compare both)

    def forward(self, hidden_states: Tensor, position_ids: Tensor, attention_param: AttentionParameters) -> Tensor:
        q_width, kv_width = self.n_qo_heads * self.head_dim, self.n_kv_heads * self.head_dim
        if self.qkv_proj is not None:
            qkv = self.qkv_proj(hidden_states)
            query, key, value = qkv[:, :q_width], qkv[:, q_width:q_width + kv_width], qkv[:, q_width + kv_width:]
        else:
            query = self.q_proj(hidden_states) if self.q_proj is not None else hidden_states
            key = self.k_proj(hidden_states)
            value = self.v_proj(hidden_states)
        query = query.view(-1, self.n_qo_heads, self.head_dim)
        key = key.view(-1, self.n_kv_heads, self.head_dim)
        value = value.view(-1, self.n_kv_heads, self.head_dim)
        if self.rotary_emb is not None and self.fuse_rope_append and query.device.type == "cuda":
            key_cache, value_cache = attention_param.kv_cache.get_kv_cache()
            self.rotary_emb.handler.forward_and_cache(query, key, value, position_ids, attention_param.new_cache_slots, key_cache, value_cache)
            hidden_states = self.attention.handler(query, attention_param).o
        else:
            if self.rotary_emb is not None:
                query, key = self.rotary_emb(query, key, position_ids)
            hidden_states = self.attention(query, key, value, attention_param).o
        if self.o_proj is not None:
            hidden_states = self.o_proj(hidden_states)
        return hidden_states

    __call__ = forward
