"""Seeded synthetic batches for the paged-KV attention path (tests, smoke and bench share this generator).

Conventions follow SURVEY §8(d): q/k/v and pools ~ N(0,1) cast to dtype, block tables are a seeded random permutation
of pool blocks (non-contiguous pages), every sequence owns ceil(L/block_size) distinct blocks, new_cache_slots are the
last q_i logical positions mapped through the block table (reference tests/layer/test_attention.py:73-82)."""
from __future__ import annotations

from dataclasses import dataclass
from itertools import accumulate

import torch
from torch import Tensor


@dataclass
class SyntheticBatch:
    # geometry
    n_qo_heads: int
    n_kv_heads: int
    head_dim: int
    block_size: int
    n_blocks: int
    dtype: torch.dtype
    # per-sequence (q_len, kv_len)
    seq_lens: list[tuple[int, int]]
    # host metadata (python lists, as AttentionParametersBuilder accumulates them)
    q_cu_seq_lens: list[int]
    kv_cu_seq_lens: list[int]
    new_cache_slots: list[int]
    block_tables: list[int]
    cu_blocks_lens: list[int]
    per_seq_block_tables: list[list[int]]
    # tensors (on `device`)
    query: Tensor        # [T, Hq*d]   (a strided slice of qkv when fused_qkv)
    key: Tensor          # [T, Hkv*d]
    value: Tensor        # [T, Hkv*d]
    key_cache: Tensor    # [NB, bs, Hkv, d]
    value_cache: Tensor  # [NB, bs, Hkv, d]

    @property
    def n_tokens(self) -> int:
        return self.q_cu_seq_lens[-1]

    @property
    def q_max(self) -> int:
        return max(q for q, _ in self.seq_lens)

    @property
    def kv_max(self) -> int:
        return max(kv for _, kv in self.seq_lens)

    def requests(self) -> list[tuple[int, int, list[int], list[int]]]:
        """(q_len, kv_len, new_cache_slots, block_table) per sequence — the arguments of add_request."""
        out = []
        for i, (q, kv) in enumerate(self.seq_lens):
            out.append((q, kv, self.new_cache_slots[self.q_cu_seq_lens[i]: self.q_cu_seq_lens[i + 1]], self.per_seq_block_tables[i]))
        return out

    def int_tensor(self, values: list[int], device=None) -> Tensor:
        return torch.tensor(values, dtype=torch.int32, device=device if device is not None else self.query.device)

    def to(self, device) -> "SyntheticBatch":
        kw = dict(self.__dict__)
        for name in ("query", "key", "value", "key_cache", "value_cache"):
            kw[name] = kw[name].to(device)
        return SyntheticBatch(**kw)

    def clone_caches(self) -> tuple[Tensor, Tensor]:
        return self.key_cache.clone(), self.value_cache.clone()


def make_batch(seq_lens: list[tuple[int, int]], n_qo_heads: int, n_kv_heads: int, head_dim: int, block_size: int = 16,
               n_blocks: int | None = None, dtype: torch.dtype = torch.bfloat16, device="cpu", seed: int = 0,
               fused_qkv: bool = False, gen_device=None) -> SyntheticBatch:
    """Random batch.  Data is generated with a seeded generator on `gen_device` (default CPU, so CPU oracle and GPU
    kernels can be fed bit-identical inputs) and moved to `device`."""
    gen_device = torch.device(gen_device if gen_device is not None else "cpu")
    g = torch.Generator(device=gen_device)
    g.manual_seed(seed)
    need_blocks = sum((kv + block_size - 1) // block_size for _, kv in seq_lens)
    if n_blocks is None:
        n_blocks = need_blocks + 3
    assert n_blocks >= need_blocks, f"pool of {n_blocks} blocks cannot hold {need_blocks}"
    perm = torch.randperm(n_blocks, generator=torch.Generator().manual_seed(seed + 1)).tolist()

    q_lens = [q for q, _ in seq_lens]
    kv_lens = [kv for _, kv in seq_lens]
    for q, kv in seq_lens:
        assert 1 <= q <= kv, f"need 1 <= q_len <= kv_len, got ({q}, {kv})"
    n_tokens = sum(q_lens)

    def randn(*shape):
        return torch.randn(*shape, generator=g, device=gen_device, dtype=torch.float32).to(dtype)

    if fused_qkv:
        width = (n_qo_heads + 2 * n_kv_heads) * head_dim
        qkv = randn(n_tokens, width).to(device)
        query = qkv[:, : n_qo_heads * head_dim]
        key = qkv[:, n_qo_heads * head_dim: (n_qo_heads + n_kv_heads) * head_dim]
        value = qkv[:, (n_qo_heads + n_kv_heads) * head_dim:]
    else:
        query = randn(n_tokens, n_qo_heads * head_dim).to(device)
        key = randn(n_tokens, n_kv_heads * head_dim).to(device)
        value = randn(n_tokens, n_kv_heads * head_dim).to(device)
    key_cache = randn(n_blocks, block_size, n_kv_heads, head_dim).to(device)
    value_cache = randn(n_blocks, block_size, n_kv_heads, head_dim).to(device)

    new_cache_slots: list[int] = []
    block_tables: list[int] = []
    per_seq: list[list[int]] = []
    cu_blocks = [0]
    cursor = 0
    for q, kv in seq_lens:
        nb = (kv + block_size - 1) // block_size
        table = perm[cursor: cursor + nb]
        cursor += nb
        per_seq.append(table)
        block_tables += table
        cu_blocks.append(cu_blocks[-1] + nb)
        for pos in range(kv - q, kv):
            new_cache_slots.append(table[pos // block_size] * block_size + pos % block_size)

    return SyntheticBatch(
        n_qo_heads=n_qo_heads, n_kv_heads=n_kv_heads, head_dim=head_dim, block_size=block_size, n_blocks=n_blocks, dtype=dtype,
        seq_lens=list(seq_lens), q_cu_seq_lens=[0] + list(accumulate(q_lens)), kv_cu_seq_lens=[0] + list(accumulate(kv_lens)),
        new_cache_slots=new_cache_slots, block_tables=block_tables, cu_blocks_lens=cu_blocks, per_seq_block_tables=per_seq,
        query=query, key=key, value=value, key_cache=key_cache, value_cache=value_cache)


def shard_round_robin(n_items: int, rank: int, world_size: int) -> list[int]:
    """Indices owned by `rank` when items are dealt out round-robin: item i -> rank i % world_size.
    Sequences are independent units, so this is the whole multi-GPU partitioning of the attention path."""
    assert 0 <= rank < world_size
    return list(range(rank, n_items, world_size))
