from .cuda_graph_model_runner import CudaGraphModelRunner, StaticAttentionMetadata

__all__ = ["CudaGraphModelRunner", "StaticAttentionMetadata"]
