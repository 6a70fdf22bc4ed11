"""TokenCache / VirtualTokenCache with the reference's interface (hydrainfer/memory/token_cache.py:15-66).
CUDA tensors go through the hi_set_image_cache kernel; there is no Python fallback loop in the product path."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

from torch import Tensor

from .._C.kernel.cache_kernels import set_image_cache


class TokenCache:
    """KV cache: caches = [key_cache, value_cache]; image-embedding cache: caches = [image_cache];
    every cache is [n_blocks, block_size, n_heads, head_size] (token_cache.py:16-22)."""

    def __init__(self, caches: list[Tensor]):
        first = caches[0]
        for cache in caches:
            assert cache.dim() == 4, f"cache must be (n_blocks, block_size, n_heads, head_size), got {tuple(cache.shape)}"
            assert cache.shape == first.shape and cache.dtype == first.dtype and cache.device == first.device, \
                "all caches of a TokenCache share shape, dtype and device"
        self.caches = caches
        self.block_size = first.shape[1]
        self.dtype = first.dtype
        self.device = first.device

    def get_caches(self) -> list[Tensor]:
        return self.caches

    def set_caches(self, slot_ids: Tensor, values: list[Tensor]) -> None:
        """caches[k][slot // bs, slot % bs] = values[k][t] for every cache (token_cache.py:37-56)."""
        assert slot_ids.dim() == 1
        for value in values:
            assert value.dim() == 3 and value.shape[0] == slot_ids.shape[0] and value.device == slot_ids.device
        for cache, value in zip(self.caches, values):
            set_image_cache(slot_ids, value, cache)


@dataclass
class VirtualTokenCache:
    """A request's view of the pool (token_cache.py:59-66); pickled across processes during migration."""
    vid: int
    n_blocks_of_cache_manager: int  # n_blocks of the pool that owns block_table (index math on the remote side)
    n_cache_tokens: int = 0
    block_table: list[int] = field(default_factory=list)
    memory_handle: Optional[list[int]] = None  # cudaIpcMemHandle_t bytes of the owning pool
    rank: int = -1  # torch.distributed rank of the owning process (NCCL backend)
