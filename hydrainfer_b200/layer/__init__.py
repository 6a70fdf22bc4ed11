from .causal_attention import (AttentionParameters, AttentionParametersBuilder, B200CausalGroupedQueryPageAttentionHandler,
                               CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig,
                               CausalGroupedQueryPageAttentionOutput)
from .multihead_attention import (B200MultiHeadAttentionHandler, B200QwenMultiHeadAttentionHandler, MultiHeadAttention,
                                  MultiHeadAttentionConfig, MultiHeadAttentionOutput, MultiHeadAttentionParameters, QwenMultiHeadAttention)
from .rotary_embedding import B200RotaryEmbeddingHandler, RotaryEmbedding, compute_default_inv_freq

__all__ = ["B200MultiHeadAttentionHandler", "B200QwenMultiHeadAttentionHandler", "MultiHeadAttention", "MultiHeadAttentionConfig", "MultiHeadAttentionOutput",
           "MultiHeadAttentionParameters", "QwenMultiHeadAttention", "B200RotaryEmbeddingHandler", "RotaryEmbedding", "compute_default_inv_freq", "AttentionParameters", "AttentionParametersBuilder", "B200CausalGroupedQueryPageAttentionHandler",
           "CausalGroupedQueryPageAttention", "CausalGroupedQueryPageAttentionConfig", "CausalGroupedQueryPageAttentionOutput"]
