"""ctypes binding of libhi_b200.so — the C ABI declared in include/hi_b200.h.

The product path has no CPU fallback: if the shared library cannot be loaded (or built from the in-tree sources)
importing this module raises, and every entry point raises RuntimeError on a non-zero status.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libhi_b200.so"

HI_F32, HI_F16, HI_BF16 = 0, 1, 2
HI_ATTN_AUTO, HI_ATTN_SIMT, HI_ATTN_TCGEN05, HI_ATTN_TCGEN05_DECODE, HI_ATTN_TCGEN05_PAIR = 0, 1, 2, 3, 4
HI_ATTN_OPT_WINDOW, HI_ATTN_OPT_SOFTCAP, HI_ATTN_OPT_ALIBI = 1, 2, 4

_DTYPES = {torch.float32: HI_F32, torch.float16: HI_F16, torch.bfloat16: HI_BF16}


class HiAttnArgs(Structure):
    _fields_ = [
        ("q", c_void_p), ("out", c_void_p), ("key_cache", c_void_p), ("value_cache", c_void_p),
        ("q_row_stride", c_int64), ("out_row_stride", c_int64),
        ("q_cu_seq_lens", c_void_p), ("kv_cu_seq_lens", c_void_p), ("block_tables", c_void_p), ("cu_blocks_lens", c_void_p),
        ("n_seqs", c_int32), ("n_tokens", c_int32), ("max_q_len", c_int32), ("max_kv_len", c_int32),
        ("n_qo_heads", c_int32), ("n_kv_heads", c_int32), ("head_dim", c_int32), ("block_size", c_int32),
        ("n_blocks", c_int64),
        ("dtype", c_int32), ("softmax_scale", c_float),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("path", c_int32), ("device", c_int32), ("kv_blocks_hint", c_int32), ("reserved", c_int32 * 3),
        ("work_items", c_void_p), ("qk_work_hint", c_int64), ("n_work_items", c_int32), ("work_tile_tokens", c_int32),
        ("options", c_int32), ("window_left", c_int32), ("window_right", c_int32), ("softcap", c_float),
        ("alibi_slopes", c_void_p), ("alibi_batch_stride", c_int64),
    ]


class HiRopeArgs(Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p),
        ("q_row_stride", c_int64), ("k_row_stride", c_int64), ("v_row_stride", c_int64),
        ("positions", c_void_p), ("cos_sin", c_void_p), ("slot_ids", c_void_p), ("key_cache", c_void_p), ("value_cache", c_void_p),
        ("n_tokens", c_int64),
        ("n_qo_heads", c_int32), ("n_kv_heads", c_int32), ("head_dim", c_int32), ("rotary_dim", c_int32),
        ("dtype", c_int32), ("cos_sin_dtype", c_int32), ("positions_int64", c_int32), ("interleaved", c_int32),
        ("write_back_k", c_int32), ("force_scalar", c_int32), ("device", c_int32), ("reserved", c_int32),
    ]


class HiVarlenArgs(Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("out", c_void_p),
        ("q_row_stride", c_int64), ("k_row_stride", c_int64), ("v_row_stride", c_int64), ("out_row_stride", c_int64),
        ("cu_seqlens_q", c_void_p), ("cu_seqlens_k", c_void_p),
        ("n_seqs", c_int32), ("n_q_tokens", c_int32), ("n_k_tokens", c_int32), ("max_q_len", c_int32), ("max_kv_len", c_int32),
        ("n_qo_heads", c_int32), ("n_kv_heads", c_int32), ("head_dim", c_int32), ("dtype", c_int32), ("causal", c_int32),
        ("softmax_scale", c_float), ("device", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_int64), ("reserved", c_int64 * 2),
    ]


class HiPoolGeom(Structure):
    _fields_ = [("n_layers", c_int64), ("n_tokens", c_int64), ("n_blocks", c_int64), ("run_bytes", c_int64)]


# name -> (restype, argtypes); also the list the symbol-export test checks against include/hi_b200.h
SIGNATURES = {
    "hi_last_error": (c_char_p, []),
    "hi_abi_version": (c_int, []),
    "hi_last_launch_count": (c_int, []),
    "hi_event_create": (c_int, [POINTER(c_void_p)]),
    "hi_event_destroy": (c_int, [c_void_p]),
    "hi_event_record": (c_int, [c_void_p, c_void_p]),
    "hi_event_elapsed_ms": (c_int, [c_void_p, c_void_p, POINTER(c_float)]),
    "hi_set_kernel_timing_events": (c_int, [c_void_p, c_void_p]),
    "hi_set_kv_cache": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                c_int, c_int, c_void_p]),
    "hi_set_image_cache": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "hi_get_image_cache": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "hi_rope_append": (c_int, [POINTER(HiRopeArgs), c_void_p]),
    "hi_varlen_attention": (c_int, [POINTER(HiVarlenArgs), c_void_p]),
    "hi_attention_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "hi_paged_attention": (c_int, [POINTER(HiAttnArgs), c_void_p]),
    "hi_attention_tile_tokens": (c_int32, [c_int32, c_int32]),
    "hi_migrate_blocks": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, HiPoolGeom, HiPoolGeom, c_int, c_void_p]),
    "hi_migrate_blocks_layers": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, HiPoolGeom, HiPoolGeom, c_int64, c_int64, c_int, c_void_p]),
    "hi_migrate_blocks_host_tables": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, HiPoolGeom, HiPoolGeom, c_int64, c_int64, c_int, c_void_p]),
    "hi_migrate_inline_table_blocks": (c_int, []),
    "hi_migrate_set_max_ctas": (c_int, [c_int]),
    "hi_ipc_get_handle": (c_int, [c_void_p, POINTER(c_uint8), POINTER(c_int64), c_int]),
    "hi_ipc_open_handle": (c_int, [POINTER(c_uint8), c_int64, c_int, POINTER(c_void_p)]),
    "hi_ipc_close_all": (c_int, []),
    "hi_enable_peer_access": (c_int, [c_int, c_int]),
    "hi_peer_copy": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p]),
}


def _load() -> ctypes.CDLL:
    from . import build as _build

    override = os.environ.get("HI_B200_LIB")  # dev only: an instrumented variant built by build.build_variant()
    if override:
        lib = ctypes.CDLL(override)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        return lib
    force = os.environ.get("HI_B200_REBUILD") == "1"
    if force or not LIB_PATH.exists() or not _build.is_current():
        try:
            _build.build(force=force)  # needs nvcc
        except Exception as e:
            if not LIB_PATH.exists():
                raise ImportError(f"hydrainfer_b200: {LIB_PATH} is missing and cannot be built: {e}") from e
            raise
    try:
        lib = ctypes.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover - exercised only on a broken install
        raise ImportError(f"hydrainfer_b200: cannot load the CUDA extension {LIB_PATH}: {e}") from e
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def check(status: int) -> None:
    """Translate a HiStatus into the RuntimeError the reference's pybind modules raise (TORCH_CHECK)."""
    if status != 0:
        raise RuntimeError(f"hi_b200 error {status}: {lib.hi_last_error().decode()}")


def dtype_code(dtype: torch.dtype) -> int:
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise RuntimeError(f"hi_b200: dtype {dtype} is not supported (float32, float16, bfloat16)") from None


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors on one CUDA device (the shims call this on every launch: integer device indices, no torch.device
    objects, keep it to a fraction of a microsecond per tensor)."""
    first = tensors[0]
    if not first.is_cuda:
        raise RuntimeError("hi_b200: the CUDA extension only accepts CUDA tensors; there is no CPU fallback")
    idx = first.get_device()
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("hi_b200: the CUDA extension only accepts CUDA tensors; there is no CPU fallback")
        if t.get_device() != idx:
            raise RuntimeError(f"hi_b200: tensors on different devices ({t.device} vs {first.device})")
    return first.device


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def current_stream_ptr(device: torch.device) -> int:
    """cudaStream_t of torch's current stream on `device` (what at::cuda::getCurrentCUDAStream() is to the reference)."""
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(device).cuda_stream
