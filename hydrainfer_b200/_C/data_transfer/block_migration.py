"""Mirror of the reference pybind module `hydrainfer._C.data_transfer.block_migration`
(csrc/data_transfer/block_migration_pybind.cpp:10-15, stub hydrainfer/_C/data_transfer/block_migration/__init__.pyi).

Same three functions and argument meaning.  Differences, all inside the contract:
  * the peer pool is mapped once per process and cached (the reference re-opens the handle on every call,
    block_migration.cpp:213-215);
  * all (layer, K/V, block) runs of a request move in one gather launch on the current stream instead of
    n_layers * n_tokens * n_blocks cudaMemcpyAsync calls (:222-244);
  * a handle exported by THIS process resolves to the local pointer (cudaIpcOpenMemHandle cannot open a handle in
    the process that created it), which also makes single-process tests possible;
  * a pool that does not start at the beginning of its CUDA allocation is supported: the handle list then carries
    8 extra ints with the byte offset (the reference silently assumes offset 0);
  * errors raise RuntimeError instead of printf + exit(-1) (:9-15).
"""
from __future__ import annotations

import ctypes

import torch
from torch import Tensor

from ... import _lib

cudaMemoryIpcHandle = list  # list[int]: 64 handle bytes (+ 8 little-endian offset bytes when the offset is non-zero)

# handles exported by this process: key -> (device pointer, device index)
_local_exports: dict[tuple[int, ...], tuple[int, int]] = {}
# reference API: register_ipc_mem_handle returns an index into a process-global vector (block_migration.cpp:61-80)
_registered: list[int] = []


def _split(handle: list[int]) -> tuple[bytes, int]:
    if len(handle) not in (64, 72):
        raise RuntimeError(f"ipc handle must have 64 (or 72) entries, got {len(handle)}")
    raw = bytes(int(b) & 0xFF for b in handle[:64])
    offset = int.from_bytes(bytes(int(b) & 0xFF for b in handle[64:]), "little") if len(handle) == 72 else 0
    return raw, offset


def get_ipc_mem_handle(tensor: Tensor) -> cudaMemoryIpcHandle:
    """cudaIpcGetMemHandle of the tensor's storage as a list of ints, one per byte (block_migration.cpp:34-40, 55-59)."""
    dev = _lib.require_cuda(tensor)
    out = (ctypes.c_uint8 * 64)()
    offset = ctypes.c_int64(0)
    _lib.check(_lib.lib.hi_ipc_get_handle(tensor.data_ptr(), out, ctypes.byref(offset), dev.index or 0))
    handle = [int(b) for b in out]
    if offset.value != 0:
        handle += [int(b) for b in int(offset.value).to_bytes(8, "little")]
    _local_exports[tuple(handle)] = (tensor.data_ptr(), dev.index or 0)
    return handle


def _resolve(handle: list[int], device_index: int) -> int:
    """Device pointer of the pool named by `handle`, usable from `device_index`."""
    local = _local_exports.get(tuple(handle))
    if local is not None:
        ptr, src_dev = local
        if src_dev != device_index:
            _lib.check(_lib.lib.hi_enable_peer_access(device_index, src_dev))
        return ptr
    raw, offset = _split(handle)
    buf = (ctypes.c_uint8 * 64).from_buffer_copy(raw)
    ptr = ctypes.c_void_p(0)
    _lib.check(_lib.lib.hi_ipc_open_handle(buf, offset, device_index, ctypes.byref(ptr)))
    return int(ptr.value)


def register_ipc_mem_handle(kv_cache_handle_vec: cudaMemoryIpcHandle) -> int:
    """Map a peer pool and return its index; -1 if peer access is unsupported (block_migration.cpp:69-80)."""
    device_index = torch.cuda.current_device()
    try:
        ptr = _resolve(list(kv_cache_handle_vec), device_index)
    except RuntimeError as e:
        if "error -5" in str(e):  # HI_ERR_PEER_UNSUPPORTED
            return -1
        raise
    _registered.append(ptr)
    return len(_registered) - 1


def _check_tables(src_block_table: list[int], dst_block_table: list[int], src_n_blocks: int, dst_n_blocks: int) -> int:
    if len(src_block_table) != len(dst_block_table):
        raise RuntimeError(f"migrate_blocks: block tables differ in length ({len(src_block_table)} vs {len(dst_block_table)})")
    n = len(dst_block_table)
    if n and (max(dst_block_table) >= dst_n_blocks or min(dst_block_table) < 0 or max(src_block_table) >= src_n_blocks or min(src_block_table) < 0):
        raise RuntimeError("migrate_blocks: block id out of range")
    return n


def _launch(src_block_table: list[int], dst_block_table: list[int], src_ptr: int, dst_ptr: int, shape: tuple, element_size: int,
            src_n_blocks: int, dst_n_blocks: int, layer_begin: int, layer_end: int, dev: torch.device) -> None:
    n_layers, n_tokens, _, block_size, n_heads, head_size = shape
    run_bytes = block_size * n_heads * head_size * element_size
    tables = torch.tensor([src_block_table, dst_block_table], dtype=torch.int32, device=dev)
    src_geom = _lib.HiPoolGeom(n_layers, n_tokens, int(src_n_blocks), run_bytes)
    dst_geom = _lib.HiPoolGeom(n_layers, n_tokens, int(dst_n_blocks), run_bytes)
    _lib.check(_lib.lib.hi_migrate_blocks_layers(tables[0].data_ptr(), tables[1].data_ptr(), len(dst_block_table), src_ptr, dst_ptr,
                                                 src_geom, dst_geom, int(layer_begin), int(layer_end), dev.index or 0,
                                                 _lib.current_stream_ptr(dev)))
    # `tables` may be freed by Python before the kernel runs; the caching allocator keeps the block alive for the
    # stream it was allocated on, and record_stream covers a migrate stream that differs from the allocation stream.
    tables.record_stream(torch.cuda.current_stream(dev))


def migrate_blocks(src_block_table: list[int], dst_block_table: list[int], src_cache: cudaMemoryIpcHandle,
                   dst_cache: Tensor, src_cache_n_blocks: int) -> None:
    """Copy blocks src_block_table[i] -> dst_block_table[i] for every (layer, K/V) plane of the pools
    (block_migration.cpp:194-245).  dst_cache is the local 6-D pool
    (n_layers, n_tokens, n_blocks, block_size, n_heads, head_size); the source pool has the same geometry except
    n_blocks == src_cache_n_blocks.  Asynchronous on the current stream."""
    migrate_blocks_layers(src_block_table, dst_block_table, src_cache, dst_cache, src_cache_n_blocks, 0, dst_cache.shape[0] if dst_cache.dim() == 6 else 0)


def migrate_blocks_layers(src_block_table: list[int], dst_block_table: list[int], src_cache: cudaMemoryIpcHandle,
                          dst_cache: Tensor, src_cache_n_blocks: int, layer_begin: int, layer_end: int) -> None:
    """migrate_blocks restricted to layers [layer_begin, layer_end) — extension of the reference module: a decode node can
    pull a prefill's pages layer by layer while later layers are still being computed."""
    dev = _lib.require_cuda(dst_cache)
    if dst_cache.dim() != 6 or not dst_cache.is_contiguous():
        raise RuntimeError("migrate_blocks: dst_cache must be a contiguous 6-D pool")
    n = _check_tables(src_block_table, dst_block_table, src_cache_n_blocks, dst_cache.shape[2])
    if n == 0:
        return
    src_ptr = _resolve(list(src_cache), dev.index or 0)
    _launch(src_block_table, dst_block_table, src_ptr, dst_cache.data_ptr(), tuple(dst_cache.shape), dst_cache.element_size(),
            src_cache_n_blocks, dst_cache.shape[2], layer_begin, layer_end, dev)


def push_blocks(src_block_table: list[int], dst_block_table: list[int], src_cache: Tensor, dst_cache: cudaMemoryIpcHandle,
                dst_cache_n_blocks: int, layer_begin: int = 0, layer_end: int = -1) -> None:
    """The same copy issued by the SENDER: src_cache is the local pool, dst_cache the IPC handle of the receiver's pool; the
    kernel runs on the source GPU and writes through the peer mapping (posted NVLink writes).  Extension of the reference
    module (its IPC backend is pull-only, communication.py:33-34): lets a prefill node ship pages as soon as they exist
    without a round trip to the receiver.  Asynchronous on the current stream of the source device."""
    dev = _lib.require_cuda(src_cache)
    if src_cache.dim() != 6 or not src_cache.is_contiguous():
        raise RuntimeError("push_blocks: src_cache must be a contiguous 6-D pool")
    n = _check_tables(src_block_table, dst_block_table, src_cache.shape[2], dst_cache_n_blocks)
    if n == 0:
        return
    if layer_end < 0:
        layer_end = src_cache.shape[0]
    dst_ptr = _resolve(list(dst_cache), dev.index or 0)
    _launch(src_block_table, dst_block_table, src_cache.data_ptr(), dst_ptr, tuple(src_cache.shape), src_cache.element_size(),
            src_cache.shape[2], dst_cache_n_blocks, layer_begin, layer_end, dev)
