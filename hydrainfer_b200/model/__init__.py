from .model_forward import ROPECausalGroupedQueryPageAttention

__all__ = ["ROPECausalGroupedQueryPageAttention"]
