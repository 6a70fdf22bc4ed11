"""Migration backends with the reference's interface (hydrainfer/memory/communication.py:18-123).

  * IPCHandleMemoryBackend — same-host path: the receiver pulls the sender's pages through a CUDA-IPC peer mapping
    with ONE gather kernel over NVLink (reference: L*2*n_blocks cudaMemcpyAsync, communication.py:23-45 +
    block_migration.cpp:194-245).  `is_send=True` is a no-op exactly as in the reference (:33-34).
  * NCCLBackend — cross-host path: the reference posts one P2POp per (block, layer, K/V) view (:65-74).  Here the
    sender packs the request's pages into one contiguous staging buffer with the same gather kernel, a single
    send/recv moves it, and the receiver scatters it into its pool: 2 kernels + 1 NCCL message per request.
  * CommunicationBackendManager — backend choice by rank2host, unchanged (:102-123).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal

import torch
import torch.distributed as dist
from torch import Tensor

from .._C.data_transfer import block_migration
from .token_cache import VirtualTokenCache


class CommunicationBackend:
    def migrate_blocks(self, src_virtual_cache: VirtualTokenCache, dst_virtual_cache: VirtualTokenCache, is_send: bool):
        raise NotImplementedError()


class IPCHandleMemoryBackend(CommunicationBackend):
    def __init__(self, migrate_stream: "torch.cuda.Stream", cache: Tensor, n_blocks: int, debug: bool = False):
        self.migrate_stream = migrate_stream
        self.cache = cache
        self.n_blocks = n_blocks
        self.debug = debug

    def migrate_blocks(self, src_virtual_cache: VirtualTokenCache, dst_virtual_cache: VirtualTokenCache, is_send: bool):
        if is_send:
            return  # pull model: the sender does nothing (communication.py:33-34)
        assert src_virtual_cache.memory_handle is not None, "the source virtual cache carries no IPC handle"
        with torch.cuda.stream(self.migrate_stream):
            block_migration.migrate_blocks(
                src_virtual_cache.block_table,
                dst_virtual_cache.block_table,
                src_virtual_cache.memory_handle,
                self.cache,
                src_virtual_cache.n_blocks_of_cache_manager,
            )


    def migrate_layers(self, src_virtual_cache: VirtualTokenCache, dst_virtual_cache: VirtualTokenCache, layer_begin: int, layer_end: int,
                       wait_event: "torch.cuda.Event | None" = None) -> "torch.cuda.Event":
        """Layer-pipelined pull (extension): moves layers [layer_begin, layer_end) of the request on the migrate
        stream, after `wait_event` (recorded by the producer once those layers' KV exists), and returns an event that
        completes when they have landed — the caller releases the source pages / starts decode on it instead of the
        reference's end-of-step stream synchronize."""
        assert src_virtual_cache.memory_handle is not None, "the source virtual cache carries no IPC handle"
        with torch.cuda.stream(self.migrate_stream):
            if wait_event is not None:
                self.migrate_stream.wait_event(wait_event)
            block_migration.migrate_blocks_layers(
                src_virtual_cache.block_table, dst_virtual_cache.block_table, src_virtual_cache.memory_handle, self.cache,
                src_virtual_cache.n_blocks_of_cache_manager, layer_begin, layer_end)
            done = torch.cuda.Event()
            done.record(self.migrate_stream)
        return done

    def push_blocks(self, src_virtual_cache: VirtualTokenCache, dst_virtual_cache: VirtualTokenCache, layer_begin: int = 0, layer_end: int = -1,
                    wait_event: "torch.cuda.Event | None" = None) -> "torch.cuda.Event":
        """Sender-side variant: this process owns the SOURCE pool (self.cache) and writes the pages into the receiver's pool
        through its IPC mapping (dst_virtual_cache.memory_handle).  Returns the completion event on the migrate stream."""
        assert dst_virtual_cache.memory_handle is not None, "the destination virtual cache carries no IPC handle"
        with torch.cuda.stream(self.migrate_stream):
            if wait_event is not None:
                self.migrate_stream.wait_event(wait_event)
            block_migration.push_blocks(
                src_virtual_cache.block_table, dst_virtual_cache.block_table, self.cache, dst_virtual_cache.memory_handle,
                dst_virtual_cache.n_blocks_of_cache_manager, layer_begin, layer_end)
            done = torch.cuda.Event()
            done.record(self.migrate_stream)
        return done


class NCCLBackend(CommunicationBackend):
    """Packed send/recv over torch.distributed. Works with any backend that supports send/recv on the pool's device
    (nccl on GPUs; the packing logic is covered on CPU tensors with gloo in tests through `pack`/`unpack` hooks)."""

    def __init__(self, migrate_stream: "torch.cuda.Stream", cache: Tensor, debug: bool = False, gather_fn=None):
        self.migrate_stream = migrate_stream
        self.cache = cache
        self.debug = debug
        # (src_pool, dst_pool, src_blocks, dst_blocks) -> None; tests inject a CPU stand-in to cover the protocol on gloo
        self._gather = gather_fn if gather_fn is not None else self._gather_cuda

    def _staging(self, n: int) -> Tensor:
        n_layers, n_tokens, _, block_size, n_heads, head_size = self.cache.shape
        return torch.empty((n_layers, n_tokens, n, block_size, n_heads, head_size), dtype=self.cache.dtype, device=self.cache.device)

    def _gather_cuda(self, src: Tensor, dst: Tensor, src_blocks: list[int], dst_blocks: list[int]) -> None:
        block_migration.copy_blocks(src_blocks, dst_blocks, src, dst)

    def migrate_blocks(self, src_virtual_cache: VirtualTokenCache, dst_virtual_cache: VirtualTokenCache, is_send: bool):
        block_table = src_virtual_cache.block_table if is_send else dst_virtual_cache.block_table
        peer = dst_virtual_cache.rank if is_send else src_virtual_cache.rank
        n = len(block_table)
        if n == 0:
            return
        with torch.cuda.stream(self.migrate_stream):
            staging = self._staging(n)
            if is_send:
                self._gather(self.cache, staging, block_table, list(range(n)))
                dist.send(staging, dst=peer)
            else:
                dist.recv(staging, src=peer)
                self._gather(staging, self.cache, list(range(n)), block_table)
            if staging.is_cuda:
                staging.record_stream(torch.cuda.current_stream(self.cache.device))


@dataclass
class CommunicationBackendManagerContext:
    migrate_stream: "torch.cuda.Stream"
    cache: Tensor  # (n_layers, n_tokens, n_blocks, block_size, n_heads, head_size)
    n_blocks: int
    rank2host: dict[int, str]


@dataclass
class CommunicationBackendManagerConfig:
    intranode_migrate_backend: Literal["auto", "ipc", "nccl"] = "auto"
    internode_migrate_backend: Literal["nccl"] = "nccl"
    debug: bool = False


def get_migrate_backend(backend: str, context: CommunicationBackendManagerContext, debug: bool = False) -> CommunicationBackend:
    if backend == "ipc":
        return IPCHandleMemoryBackend(context.migrate_stream, context.cache, context.n_blocks, debug)
    if backend == "nccl":
        return NCCLBackend(context.migrate_stream, context.cache, debug)
    raise ValueError(f"invalid migrate backend {backend}")


class CommunicationBackendManager(CommunicationBackend):
    def __init__(self, config: CommunicationBackendManagerConfig, context: CommunicationBackendManagerContext):
        self.context = context
        self.rank2host = context.rank2host
        intranode = "ipc" if config.intranode_migrate_backend == "auto" else config.intranode_migrate_backend
        self.intranode_backend = get_migrate_backend(intranode, context, config.debug)
        self.internode_backend = get_migrate_backend(config.internode_migrate_backend, context, config.debug)

    def in_same_machine(self, rank1: int, rank2: int) -> bool:
        if rank1 not in self.rank2host or rank2 not in self.rank2host:
            return False
        return self.rank2host[rank1] == self.rank2host[rank2]

    def migrate_blocks(self, src_virtual_cache: VirtualTokenCache, dst_virtual_cache: VirtualTokenCache, is_send: bool):
        assert src_virtual_cache.n_cache_tokens == dst_virtual_cache.n_cache_tokens, \
            f"{src_virtual_cache.n_cache_tokens} {dst_virtual_cache.n_cache_tokens}"
        if self.in_same_machine(src_virtual_cache.rank, dst_virtual_cache.rank):
            self.intranode_backend.migrate_blocks(src_virtual_cache, dst_virtual_cache, is_send)
        else:
            self.internode_backend.migrate_blocks(src_virtual_cache, dst_virtual_cache, is_send)
