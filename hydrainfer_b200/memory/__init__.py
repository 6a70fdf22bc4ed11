"""Host-side mirror of `hydrainfer.memory` for the paged-KV hot path (same public names as the reference's
hydrainfer/memory/__init__.py:1-6, minus the prefix-hash helpers that are out of scope)."""
from .block_allocator import BlockAllocator, BlockAllocatorMetrics
from .token_cache import TokenCache, VirtualTokenCache
from .communication import (CommunicationBackendManager, CommunicationBackendManagerConfig,
                            CommunicationBackendManagerContext, IPCHandleMemoryBackend, NCCLBackend)
from .token_cache_manger import TokenCacheBlockManager, TokenCacheBlockManagerConfig, TokenCacheBlockManagerContext
from .kv_cache import KVCache

__all__ = [
    "BlockAllocator", "BlockAllocatorMetrics", "TokenCache", "VirtualTokenCache", "CommunicationBackendManager",
    "CommunicationBackendManagerConfig", "CommunicationBackendManagerContext", "IPCHandleMemoryBackend", "NCCLBackend",
    "TokenCacheBlockManager", "TokenCacheBlockManagerConfig", "TokenCacheBlockManagerContext", "KVCache",
]
