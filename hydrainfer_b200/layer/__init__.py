from .causal_attention import (AttentionParameters, AttentionParametersBuilder, B200CausalGroupedQueryPageAttentionHandler,
                               CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig,
                               CausalGroupedQueryPageAttentionOutput)

__all__ = ["AttentionParameters", "AttentionParametersBuilder", "B200CausalGroupedQueryPageAttentionHandler",
           "CausalGroupedQueryPageAttention", "CausalGroupedQueryPageAttentionConfig", "CausalGroupedQueryPageAttentionOutput"]
