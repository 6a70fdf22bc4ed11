"""Mirror of the reference pybind module `hydrainfer._C.kernel.cache_kernels`
(csrc/kernel/cache_kernels/cache_kernels_pybind.cpp:7-10); backed by hi_set_image_cache."""
from __future__ import annotations

import torch
from torch import Tensor

from ... import _lib


def set_image_cache(slot_ids: Tensor, image_tokens: Tensor, image_cache: Tensor) -> None:
    """image_cache[slot // bs, slot % bs] = image_tokens[t] (cache_kernels.cu:55-83).

    slot_ids int32 [T]; image_tokens [T, H, d] contiguous over (H, d); image_cache [NB, bs, H, d] contiguous."""
    dev = _lib.require_cuda(slot_ids, image_tokens, image_cache)
    if image_tokens.dim() != 3 or image_tokens.stride(-1) != 1 or image_tokens.stride(-2) != image_tokens.size(-1):
        raise RuntimeError("set_image_cache: image_tokens must be [n_tokens, n_heads, head_dim], contiguous over the last two dims")
    if image_cache.dim() != 4 or not image_cache.is_contiguous() or tuple(image_cache.shape[-2:]) != tuple(image_tokens.shape[-2:]):
        raise RuntimeError("set_image_cache: image_cache must be contiguous [n_blocks, block_size, n_heads, head_dim] matching the tokens")
    if image_tokens.dtype != image_cache.dtype:
        raise RuntimeError("set_image_cache: dtype mismatch")
    if slot_ids.dtype != torch.int32 or slot_ids.dim() != 1 or not slot_ids.is_contiguous() or slot_ids.shape[0] != image_tokens.shape[0]:
        raise RuntimeError("set_image_cache: slot_ids must be a contiguous int32 vector with one entry per token")
    n_tokens = image_tokens.shape[0]
    row = image_tokens.shape[1] * image_tokens.shape[2]
    _lib.check(_lib.lib.hi_set_image_cache(
        slot_ids.data_ptr(), image_tokens.data_ptr(), image_cache.data_ptr(), n_tokens, row,
        image_tokens.stride(0) if n_tokens > 1 else row, _lib.dtype_code(image_tokens.dtype), dev.index or 0,
        _lib.current_stream_ptr(dev)))


def get_image_cache(slot_ids: Tensor, image_cache: Tensor) -> Tensor:
    """image_cache.view(-1, H * d)[slot_ids, :] -> [T, H * d]: the read side of the image-embedding cache, the gather
    LanguageModelParametersBuilder.add does with advanced indexing (hydrainfer/engine/parameters_builder.py:48-55).
    Extension of the reference module (it has no such function); bit-exact, one launch."""
    dev = _lib.require_cuda(slot_ids, image_cache)
    if image_cache.dim() != 4 or not image_cache.is_contiguous():
        raise RuntimeError("get_image_cache: image_cache must be contiguous [n_blocks, block_size, n_heads, head_dim]")
    if slot_ids.dtype != torch.int32 or slot_ids.dim() != 1 or not slot_ids.is_contiguous():
        raise RuntimeError("get_image_cache: slot_ids must be a contiguous int32 vector")
    n_tokens = slot_ids.shape[0]
    row = image_cache.shape[2] * image_cache.shape[3]
    out = torch.empty((n_tokens, row), dtype=image_cache.dtype, device=dev)
    _lib.check(_lib.lib.hi_get_image_cache(
        slot_ids.data_ptr(), image_cache.data_ptr(), out.data_ptr(), n_tokens, row, row,
        _lib.dtype_code(image_cache.dtype), dev.index or 0, _lib.current_stream_ptr(dev)))
    return out
