"""KVCache with the reference's interface (hydrainfer/memory/kv_cache.py:14-56): per-layer (key_cache, value_cache)
views of the pool, each [n_blocks, block_size, n_kv_heads, head_size]; set_kv_cache scatters new K/V rows by slot id
through ONE hi_set_kv_cache launch.  No Python per-token loop: the oracle holds that restatement."""
from __future__ import annotations

from torch import Tensor

from .._C.kernel.kv_cache_kernels import set_kv_cache as set_kv_cache_kernel
from .token_cache import TokenCache


class KVCache:
    def __init__(self, key_cache: Tensor, value_cache: Tensor):
        assert key_cache.dim() == 4, f"key cache must be 4-D, got {tuple(key_cache.shape)}"
        assert value_cache.shape == key_cache.shape, f"key/value cache shapes differ: {tuple(key_cache.shape)} {tuple(value_cache.shape)}"
        self.key_cache = key_cache
        self.value_cache = value_cache
        self.dtype = key_cache.dtype
        self.device = key_cache.device
        self.block_size = key_cache.shape[1]

    def get_kv_cache(self) -> tuple[Tensor, Tensor]:
        return (self.key_cache, self.value_cache)

    def set_kv_cache(self, slot_ids: Tensor, keys: Tensor, values: Tensor) -> None:
        """slot_ids int32 [T]; keys/values [T, n_kv_heads, head_size] (kv_cache.py:27-50)."""
        assert slot_ids.shape[0] == keys.shape[0] == values.shape[0], f"{slot_ids.shape} {keys.shape} {values.shape}"
        assert slot_ids.device == keys.device == values.device, f"{slot_ids.device} {keys.device} {values.device}"
        set_kv_cache_kernel(slot_ids, keys, values, self.key_cache, self.value_cache)

    @classmethod
    def from_token_cache(cls, token_cache: TokenCache) -> "KVCache":
        tensors = token_cache.get_caches()
        assert len(tensors) == 2, f"a KV cache has exactly two tensors, got {len(tensors)}"
        return cls(tensors[0], tensors[1])
