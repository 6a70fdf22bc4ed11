"""Free-list block allocator with the reference's contract (hydrainfer/memory/block_allocator.py:11-39):
ids are handed out LIFO from the tail of a list initialised to [n-1, ..., 0], `allocate(n)` returns AT MOST n ids
(a partial list when the pool runs dry, [] for n == 0), `free` appends, so the most recently freed blocks are
reused first.  KATs from the reference's tests/memory/test_block_allocator.py:5-22, 34-38 are in tests/."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class BlockAllocatorMetrics:
    n_used_blocks: int
    n_total_blocks: int
    block_usage: float


class BlockAllocator:
    def __init__(self, total_blocks: int):
        self.total_blocks = total_blocks
        self.free_blocks: list[int] = list(range(total_blocks - 1, -1, -1))

    def get_metrics(self) -> BlockAllocatorMetrics:
        used = self.total_blocks - len(self.free_blocks)
        return BlockAllocatorMetrics(n_used_blocks=used, n_total_blocks=self.total_blocks, block_usage=used / self.total_blocks)

    def allocate(self, n_blocks: int) -> list[int]:
        take = min(n_blocks, len(self.free_blocks))
        if take <= 0:
            return []
        cut = len(self.free_blocks) - take
        blocks = self.free_blocks[cut:]
        del self.free_blocks[cut:]
        return blocks

    def free(self, blocks: list[int]) -> None:
        self.free_blocks.extend(blocks)
        assert len(self.free_blocks) <= self.total_blocks, "more blocks freed than the pool holds"

    def get_num_avaiable_blocks(self) -> int:  # (sic) the reference's spelling is part of the interface
        return len(self.free_blocks)
