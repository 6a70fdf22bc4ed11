"""TokenCacheBlockManager with the reference's allocator contract (hydrainfer/memory/token_cache_manger.py:51-179;
the module keeps the reference's file name, typo included, so imports stay drop-in).

Pool: ONE tensor (n_layers, n_tokens, n_blocks, block_size, n_heads, head_size) (:65); layer views
pool[layer, token] are the [n_blocks, block_size, n_heads, head_size] caches the kernels see (:161-162).
Blocks come from the LIFO BlockAllocator first and, once that is dry, from the set of unpinned blocks (:97-103).

Out of scope: prefix-cache hash matching (shared_cache.py).  `allocate_virtual_cache(hashes)`
accepts hashes for signature compatibility and treats them as a miss; the pin / unpin reference counting that the
allocator contract depends on is kept.  The reference also materialises n_blocks*n_layers*n_tokens tensor views
for its NCCL path (:68-72); the packed NCCL path here does not need them.
"""
from __future__ import annotations

from array import array
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .._C.data_transfer.block_migration import get_ipc_mem_handle
from .block_allocator import BlockAllocator, BlockAllocatorMetrics
from .communication import CommunicationBackendManager, CommunicationBackendManagerConfig, CommunicationBackendManagerContext
from .token_cache import TokenCache, VirtualTokenCache

_DTYPES = {"fp16": torch.float16, "fp32": torch.float32, "bf16": torch.bfloat16}


@dataclass
class TokenCacheManagerMetrics:
    allocator_metrics: BlockAllocatorMetrics
    cache_hit_rate: float


@dataclass
class TokenCacheBlockManagerConfig:
    communication_backend_manager_config: CommunicationBackendManagerConfig
    n_layers: int = 32
    n_tokens: int = 2
    n_blocks: int = 1024
    block_size: int = 16
    n_heads: int = 32
    head_size: int = 128
    dtype: str = "fp16"  # the reference accepts fp16 / fp32 (utils/torch_utils.py:13-18); bf16 is added here
    device: str = "cuda:0"


@dataclass
class TokenCacheBlockManagerContext:
    rank: int
    rank2host: dict[int, str]


class _PinnedBlocks:
    """Reference-counted reuse pool: the part of SharedCache the allocator contract needs (shared_cache.py:20-70)."""

    def __init__(self, n_blocks: int):
        self.ref_count = [0] * n_blocks
        self.evictable: set[int] = set()

    def pin(self, block_ids: list[int]) -> None:
        for b in block_ids:
            self.ref_count[b] += 1
            self.evictable.discard(b)

    def unpin(self, block_ids: list[int]) -> None:
        for b in block_ids:
            self.ref_count[b] -= 1
            assert self.ref_count[b] >= 0, f"block {b} unpinned more often than pinned"
            if self.ref_count[b] == 0:
                self.evictable.add(b)

    def allocate(self, n_blocks: int) -> list[int]:
        return [self.evictable.pop() for _ in range(min(n_blocks, len(self.evictable)))]

    def get_num_avaiable_blocks(self) -> int:
        return len(self.evictable)


class TokenCacheBlockManager:
    def __init__(self, config: TokenCacheBlockManagerConfig, context: TokenCacheBlockManagerContext):
        self.config = config
        self.context = context
        self.n_layers = config.n_layers
        self.n_tokens = config.n_tokens
        self.n_blocks = config.n_blocks
        self.block_size = config.block_size
        self.n_heads = config.n_heads
        self.head_size = config.head_size
        self.dtype = _DTYPES[config.dtype]
        self.device = torch.device(config.device)
        self.rank = context.rank
        if self.device.type != "cuda":
            raise RuntimeError("TokenCacheBlockManager: the pool lives in GPU memory; there is no CPU path")

        self.cache_tensor = torch.randn(
            size=(self.n_layers, self.n_tokens, self.n_blocks, self.block_size, self.n_heads, self.head_size),
            dtype=self.dtype, device=self.device)
        self.memory_handle: list[int] = get_ipc_mem_handle(self.cache_tensor)
        self.block_allocator = BlockAllocator(self.n_blocks)
        self._next_vid = 0
        self.migrate_stream = torch.cuda.Stream(device=self.device)
        self.migrate_manager = CommunicationBackendManager(
            config.communication_backend_manager_config,
            CommunicationBackendManagerContext(migrate_stream=self.migrate_stream, cache=self.cache_tensor,
                                               n_blocks=self.n_blocks, rank2host=context.rank2host))
        self.shared_cache = _PinnedBlocks(self.n_blocks)
        self.total_block_queried = 0.0
        self.total_block_matched = 0.0

    def get_num_avaiable_blocks(self) -> int:
        """Free blocks + unpinned (evictable) blocks.  Deviation from the reference: its SharedCache starts with every block
        evictable, so a fresh reference pool reports 2 * n_blocks here (shared_cache.py:27-31 + :88); this one reports n_blocks,
        the number that can really be allocated."""
        return self.block_allocator.get_num_avaiable_blocks() + self.shared_cache.get_num_avaiable_blocks()

    def _allocate_new_blocks(self, n_blocks: int) -> list[int]:
        block_ids = self.block_allocator.allocate(n_blocks)
        if len(block_ids) < n_blocks:
            block_ids += self.shared_cache.allocate(n_blocks - len(block_ids))
        assert len(block_ids) == n_blocks, "not enough blocks"
        self.shared_cache.pin(block_ids)
        return block_ids

    def allocate_virtual_cache(self, hashes: Optional[list[int]] = None) -> VirtualTokenCache:
        if hashes is not None:
            self.total_block_queried += len(hashes)  # prefix matching is out of scope: every lookup is a miss
        self._next_vid += 1
        return VirtualTokenCache(vid=self._next_vid, n_cache_tokens=0, block_table=[], memory_handle=self.memory_handle,
                                 rank=self.rank, n_blocks_of_cache_manager=self.n_blocks)

    def v2p(self, virtual_cache: VirtualTokenCache, virtual_cache_ids: list[int]) -> list[int]:
        """Virtual token index -> physical slot: block_table[v // bs] * bs + v % bs (:126-133).  One id (a decode row) is
        two integer operations; long lists (a prefill chunk) go through numpy instead of a per-token Python loop."""
        bs = self.block_size
        table = virtual_cache.block_table
        n = len(virtual_cache_ids)
        if n == 1:
            v = virtual_cache_ids[0]
            return [table[v // bs] * bs + v % bs]
        if n >= 64:
            ids = np.asarray(virtual_cache_ids, dtype=np.int64)
            return (np.asarray(table, dtype=np.int64)[ids // bs] * bs + ids % bs).tolist()
        return [table[v // bs] * bs + v % bs for v in virtual_cache_ids]

    def v2p_range(self, virtual_cache: VirtualTokenCache, begin: int, end: int) -> array:
        """Physical slots of the CONTIGUOUS virtual ids [begin, end) as an int32 `array` - what every step actually asks for
        (the new tokens of a request are its last q positions, engine/parameters_builder.py:63-69).  Built block by block (one
        `range` per page touched), it goes into AttentionParametersBuilder.add_request unchanged: no per-token Python and no
        intermediate list (extension of the reference class)."""
        bs = self.block_size
        table = virtual_cache.block_table
        out = array("i")
        v = begin
        while v < end:
            block, off = divmod(v, bs)
            run = min(end - v, bs - off)
            base = table[block] * bs + off
            out.extend(range(base, base + run))
            v += run
        return out

    def set_blocks(self, virtual_cache: VirtualTokenCache, virtual_block_ids: list[int], hashes: list[int]) -> None:
        """Prefix-cache insertion (:135-138; the reference's executor calls it for every request whose Fill carries hashes,
        engine/executor.py:127).  Prefix matching is out of scope: accepted and ignored, so every later lookup stays
        a miss and the engine runs unchanged with prefix caching switched on."""
        assert len(virtual_block_ids) == len(hashes)

    def set(self, virtual_cache: VirtualTokenCache, virtual_cache_ids: list[int], hashes: list[int]) -> None:
        """Token-granular variant of set_blocks (:140-148): accepted and ignored, see set_blocks."""
        assert len(virtual_cache_ids) == len(hashes)

    def realloc(self, virtual_cache: VirtualTokenCache, n_tokens: int) -> None:
        """Grow by whole blocks or shrink and release the tail blocks (:150-159)."""
        n_need_blocks = (n_tokens + self.block_size - 1) // self.block_size
        if n_tokens > virtual_cache.n_cache_tokens:
            virtual_cache.block_table += self._allocate_new_blocks(n_need_blocks - len(virtual_cache.block_table))
        else:
            self.shared_cache.unpin(virtual_cache.block_table[n_need_blocks:])
            virtual_cache.block_table = virtual_cache.block_table[:n_need_blocks]
        virtual_cache.n_cache_tokens = n_tokens

    def get_layer_cache(self, layer_id: int) -> TokenCache:
        return TokenCache([self.cache_tensor[layer_id, token_id] for token_id in range(self.n_tokens)])

    def migrate_blocks(self, src_virtual_cache: VirtualTokenCache, dst_virtual_cache: VirtualTokenCache, is_send: bool = False):
        self.migrate_manager.migrate_blocks(src_virtual_cache, dst_virtual_cache, is_send)

    def record_migration_done(self) -> "torch.cuda.Event":
        """Event that fires when every migration issued so far has landed.  The reference only offers a stream
        synchronize at the end of the step and frees the source blocks before it (epdnode.py:399 vs :299-302);
        hand this event to whoever releases the source pages instead."""
        ev = torch.cuda.Event()
        ev.record(self.migrate_stream)
        return ev

    def synchronize(self) -> None:
        self.migrate_stream.synchronize()

    @classmethod
    def compute_n_blocks(cls, config: TokenCacheBlockManagerConfig, memory: int) -> int:
        itemsize = torch.empty((), dtype=_DTYPES[config.dtype]).element_size()
        return memory // (config.n_layers * config.n_tokens * config.block_size * config.n_heads * config.head_size * itemsize)

    def get_metrics(self) -> TokenCacheManagerMetrics:
        rate = self.total_block_matched / self.total_block_queried if self.total_block_queried else 0.0
        return TokenCacheManagerMetrics(allocator_metrics=self.block_allocator.get_metrics(), cache_hit_rate=rate)
