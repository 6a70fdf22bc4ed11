"""Mirror of the reference pybind module `hydrainfer._C.kernel.kv_cache_kernels`
(csrc/kernel/kv_cache_kernels/kv_cache_kernels_pybind.cpp:7-10, stub hydrainfer/_C/kernel/kv_cache_kernels/__init__.pyi).
Same name, positional arguments and in-place semantics; backed by hi_set_kv_cache."""
from __future__ import annotations

import torch
from torch import Tensor

from ... import _lib


def set_kv_cache(slot_ids: Tensor, keys: Tensor, values: Tensor, key_cache: Tensor, value_cache: Tensor) -> None:
    """key_cache[slot // bs, slot % bs] = keys[t]; same for values (kv_cache_kernels.cu:60-95).

    slot_ids int32 [T]; keys/values [T, Hkv, d], contiguous over the last two dims, any row stride;
    caches [NB, bs, Hkv, d] contiguous.  Layout violations raise RuntimeError (the reference CHECK-aborts).
    """
    dev = _lib.require_cuda(slot_ids, keys, values, key_cache, value_cache)
    if keys.dim() != 3 or values.shape != keys.shape:
        raise RuntimeError(f"set_kv_cache: keys/values must be [n_tokens, n_heads, head_dim], got {tuple(keys.shape)} {tuple(values.shape)}")
    if not (keys.is_contiguous() or (keys.stride(-1) == 1 and keys.stride(-2) == keys.size(-1))) or \
            not (values.is_contiguous() or (values.stride(-1) == 1 and values.stride(-2) == values.size(-1))):
        raise RuntimeError("set_kv_cache: keys and values must be contiguous over (n_heads, head_dim)")
    if key_cache.dim() != 4 or key_cache.shape != value_cache.shape or not key_cache.is_contiguous() or not value_cache.is_contiguous():
        raise RuntimeError("set_kv_cache: caches must be contiguous [n_blocks, block_size, n_heads, head_dim] of equal shape")
    if key_cache.shape[-2:] != keys.shape[-2:]:
        raise RuntimeError(f"set_kv_cache: cache heads/dim {tuple(key_cache.shape[-2:])} differ from keys {tuple(keys.shape[-2:])}")
    if not (keys.dtype == values.dtype == key_cache.dtype == value_cache.dtype):
        raise RuntimeError("set_kv_cache: dtype mismatch between keys, values and caches")
    if slot_ids.dtype != torch.int32 or slot_ids.dim() != 1 or not slot_ids.is_contiguous() or slot_ids.shape[0] != keys.shape[0]:
        raise RuntimeError("set_kv_cache: slot_ids must be a contiguous int32 vector with one entry per token")
    n_tokens = keys.shape[0]
    _lib.check(_lib.lib.hi_set_kv_cache(
        slot_ids.data_ptr(), keys.data_ptr(), values.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
        n_tokens, keys.shape[1] * keys.shape[2], keys.stride(0) if n_tokens > 1 else keys.shape[1] * keys.shape[2],
        values.stride(0) if n_tokens > 1 else keys.shape[1] * keys.shape[2],
        _lib.dtype_code(keys.dtype), dev.index or 0, _lib.current_stream_ptr(dev)))
