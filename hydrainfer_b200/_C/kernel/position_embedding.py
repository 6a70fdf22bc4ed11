"""Mirror of the reference pybind module `hydrainfer._C.kernel.position_embedding`
(csrc/kernel/position_embedding/position_embedding_pybind.cpp, imported by hydrainfer/layer/rotary_embedding.py:7).
`apply_rotary_pos_emb` keeps the reference's name, positional arguments and in-place semantics; `rope_set_kv_cache` is
the fused form (rotation + KV append in one launch) the B200 layer uses.  Both are backed by hi_rope_append."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from ... import _lib


def _check_qk(name: str, query: Tensor, key: Tensor, positions: Tensor, cos_sin: Tensor, rotary_dim: int) -> None:
    if query.dim() != 3 or key.dim() != 3 or query.shape[0] != key.shape[0] or query.shape[2] != key.shape[2]:
        raise RuntimeError(f"{name}: query/key must be [n_tokens, n_heads, head_dim] with equal n_tokens and head_dim, got {tuple(query.shape)} {tuple(key.shape)}")
    # rope.cu:100-101 (CHECK-aborts in the reference)
    if query.stride(-1) != 1 or query.stride(-2) != query.size(-1) or key.stride(-1) != 1 or key.stride(-2) != key.size(-1):
        raise RuntimeError(f"{name}: query and key must be contiguous over (n_heads, head_dim)")
    if key.dtype != query.dtype:
        raise RuntimeError(f"{name}: query/key dtype mismatch")
    if positions.dim() != 1 or positions.shape[0] != query.shape[0] or not positions.is_contiguous() or positions.dtype not in (torch.int32, torch.int64):
        raise RuntimeError(f"{name}: positions must be a contiguous int32/int64 vector with one entry per token")
    if rotary_dim % 2 != 0 or rotary_dim > query.shape[2]:
        raise RuntimeError(f"{name}: rotary_dim {rotary_dim} must be even and <= head_dim {query.shape[2]}")
    if not cos_sin.is_contiguous() or cos_sin.dim() != 3 or cos_sin.shape[1] != 2 or cos_sin.shape[2] * 2 != rotary_dim:
        raise RuntimeError(f"{name}: cos_sin must be contiguous [max_positions, 2, rotary_dim/2], got {tuple(cos_sin.shape)}")
    if cos_sin.dtype not in (query.dtype, torch.float32):
        raise RuntimeError(f"{name}: cos_sin must be float32 or have the query dtype, got {cos_sin.dtype} for {query.dtype}")


def _launch(query: Tensor, key: Tensor, value: Optional[Tensor], positions: Tensor, cos_sin: Tensor, rotary_dim: int, interleaved: bool,
            slot_ids: Optional[Tensor], key_cache: Optional[Tensor], value_cache: Optional[Tensor], write_back_k: bool, force_scalar: bool) -> None:
    dev = query.device
    n_tokens, n_heads, head_dim = query.shape
    row = lambda t: t.stride(0) if t.shape[0] > 1 else t.shape[1] * t.shape[2]
    args = _lib.HiRopeArgs()
    args.q, args.k = query.data_ptr(), key.data_ptr()
    args.q_row_stride, args.k_row_stride = row(query), row(key)
    args.positions, args.cos_sin = positions.data_ptr(), cos_sin.data_ptr()
    if slot_ids is not None:
        args.v, args.v_row_stride = value.data_ptr(), row(value)
        args.slot_ids, args.key_cache, args.value_cache = slot_ids.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr()
    args.n_tokens = n_tokens
    args.n_qo_heads, args.n_kv_heads, args.head_dim, args.rotary_dim = n_heads, key.shape[1], head_dim, rotary_dim
    args.dtype, args.cos_sin_dtype = _lib.dtype_code(query.dtype), _lib.dtype_code(cos_sin.dtype)
    args.positions_int64 = 1 if positions.dtype == torch.int64 else 0
    args.interleaved = 1 if interleaved else 0
    args.write_back_k = 1 if write_back_k else 0
    args.force_scalar = 1 if force_scalar else 0
    args.device = dev.index or 0
    _lib.check(_lib.lib.hi_rope_append(args, _lib.current_stream_ptr(dev)))


def apply_rotary_pos_emb(query: Tensor, key: Tensor, positions: Tensor, cos_sin: Tensor, rotary_dim: int, interleaved: bool) -> None:
    """Rotates query [T, Hq, d] and key [T, Hkv, d] in place (rope.cu:90-117).  positions int32 (or int64) [T];
    cos_sin [max_positions, 2, rotary_dim/2] in the query dtype (as the reference kernel reads it) or float32."""
    _lib.require_cuda(query, key, positions, cos_sin)
    _check_qk("apply_rotary_pos_emb", query, key, positions, cos_sin, rotary_dim)
    _launch(query, key, None, positions, cos_sin, rotary_dim, interleaved, None, None, None, True, False)


def rope_set_kv_cache(query: Tensor, key: Tensor, value: Tensor, positions: Tensor, cos_sin: Tensor, rotary_dim: int, interleaved: bool,
                      slot_ids: Tensor, key_cache: Tensor, value_cache: Tensor, write_back_k: bool = False, force_scalar: bool = False) -> None:
    """apply_rotary_pos_emb + set_kv_cache (kv_cache_kernels.cu:60-95) in one launch: query is rotated in place, the
    rotated key and the value go to key_cache / value_cache[slot_ids]; key itself is rewritten only if write_back_k."""
    _lib.require_cuda(query, key, value, positions, cos_sin, slot_ids, key_cache, value_cache)
    _check_qk("rope_set_kv_cache", query, key, positions, cos_sin, rotary_dim)
    if value.shape != key.shape or value.dtype != key.dtype or value.stride(-1) != 1 or value.stride(-2) != value.size(-1):
        raise RuntimeError("rope_set_kv_cache: value must have the key's shape and dtype and be contiguous over (n_heads, head_dim)")
    if key_cache.dim() != 4 or key_cache.shape != value_cache.shape or not key_cache.is_contiguous() or not value_cache.is_contiguous():
        raise RuntimeError("rope_set_kv_cache: caches must be contiguous [n_blocks, block_size, n_heads, head_dim] of equal shape")
    if tuple(key_cache.shape[-2:]) != tuple(key.shape[-2:]) or key_cache.dtype != key.dtype or value_cache.dtype != key.dtype:
        raise RuntimeError("rope_set_kv_cache: cache geometry / dtype differs from the keys")
    if slot_ids.dtype != torch.int32 or slot_ids.dim() != 1 or not slot_ids.is_contiguous() or slot_ids.shape[0] != key.shape[0]:
        raise RuntimeError("rope_set_kv_cache: slot_ids must be a contiguous int32 vector with one entry per token")
    _launch(query, key, value, positions, cos_sin, rotary_dim, interleaved, slot_ids, key_cache, value_cache, write_back_k, force_scalar)
