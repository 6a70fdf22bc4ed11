"""CPU: host-side mirrors of the reference interface (allocator contract, metadata builder, backend selection)."""
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from hydrainfer_b200.layer.causal_attention import (AttentionParametersBuilder, B200CausalGroupedQueryPageAttentionHandler,
                                                    CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig)
from hydrainfer_b200.memory import (BlockAllocator, CommunicationBackendManager, CommunicationBackendManagerConfig,
                                    CommunicationBackendManagerContext, TokenCacheBlockManager, TokenCacheBlockManagerConfig,
                                    TokenCacheBlockManagerContext, VirtualTokenCache)
from hydrainfer_b200.memory.token_cache_manger import _PinnedBlocks


def _ints(s) -> list[int]:
    s = str(s)
    return [int(v) for v in s.split(",")] if s else []


def test_allocator_reference_trace():
    z = np.load(GOLDEN / "allocator.npz")
    alloc = BlockAllocator(40)
    for op, args, result in zip(z["ops"], z["args"], z["results"]):
        if str(op) == "free":
            alloc.free(_ints(args))
        else:
            assert alloc.allocate(_ints(args)[0]) == _ints(result)


def test_allocator_random_allocate():
    # reference tests/memory/test_block_allocator.py:5-22
    total = 100
    alloc = BlockAllocator(total)
    ref = list(reversed(range(total)))
    rng = random.Random(0)
    for _ in range(10):
        n = rng.randint(1, 10)
        expect = [ref.pop() for _ in range(n)]
        expect.reverse()
        assert alloc.allocate(n) == expect
    assert alloc.allocate(0) == []


def test_allocator_free_kat():
    # reference tests/memory/test_block_allocator.py:34-38
    alloc = BlockAllocator(10)
    assert alloc.allocate(3) == [2, 1, 0]
    alloc.free([1, 0])
    assert alloc.allocate(3) == [3, 1, 0]


def test_allocator_partial_when_dry():
    # the current reference returns what is left (block_allocator.py:25-32); its stale test_out_of_memory expects []
    alloc = BlockAllocator(4)
    assert alloc.allocate(6) == [3, 2, 1, 0]
    assert alloc.allocate(1) == []
    assert alloc.get_num_avaiable_blocks() == 0
    m = alloc.get_metrics()
    assert (m.n_used_blocks, m.n_total_blocks, m.block_usage) == (4, 4, 1.0)


def test_builder_metadata_matches_reference(golden_attention):
    g = golden_attention
    builder = AttentionParametersBuilder(g.n_qo_heads, g.n_kv_heads, g.head_dim, g.block_size, torch.device("cpu"))
    for req in g.requests():
        builder.add_request(*req)
    builder.add_kv_cache(object())
    builder.add_kv_cache(object())
    params = builder.build_attention_parameters()
    assert len(params) == 2 and params[0].block_tables is params[1].block_tables or torch.equal(params[0].block_tables, params[1].block_tables)
    p = params[0]
    for name in ("q_cu_seq_lens", "kv_cu_seq_lens", "paged_kv_last_page_len", "new_cache_slots", "block_tables", "cu_blocks_lens"):
        t = getattr(p, name)
        assert t.dtype == torch.int32 and t.tolist() == getattr(g, name), name
        assert t.data_ptr() % 16 == 0, f"{name} slice is not 16-byte aligned"
    assert p.num_sequences == len(g.seq_lens)
    assert p.all_sequences_decode == all(q == 1 for q, _ in g.seq_lens)
    assert p.q_max_seq_len == max(q for q, _ in g.seq_lens) and p.kv_max_seq_len == max(kv for _, kv in g.seq_lens)


def test_handler_rejects_cpu_without_fallback():
    cfg = CausalGroupedQueryPageAttentionConfig(4, 2, 64)
    handler = B200CausalGroupedQueryPageAttentionHandler(cfg)
    with pytest.raises(RuntimeError, match="no CPU handler"):
        handler(torch.zeros(1, 4, 64), None)

    class Next(torch.nn.Module):
        def forward(self, q, p):
            return "fell through"

    handler.next_handler = Next()
    assert handler(torch.zeros(1, 4, 64), None) == "fell through"
    with pytest.raises(AssertionError):
        CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(6, 4, 64))


def test_kernels_reject_cpu_tensors():
    from hydrainfer_b200._C.kernel.kv_cache_kernels import set_kv_cache
    from hydrainfer_b200._C.kernel.cache_kernels import set_image_cache
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        set_kv_cache(torch.zeros(1, dtype=torch.int32), torch.zeros(1, 1, 8), torch.zeros(1, 1, 8), torch.zeros(1, 4, 1, 8), torch.zeros(1, 4, 1, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        set_image_cache(torch.zeros(1, dtype=torch.int32), torch.zeros(1, 1, 8), torch.zeros(1, 4, 1, 8))


def test_block_manager_has_no_cpu_pool():
    cfg = TokenCacheBlockManagerConfig(CommunicationBackendManagerConfig(), n_layers=1, n_blocks=4, n_heads=1, head_size=8, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        TokenCacheBlockManager(cfg, TokenCacheBlockManagerContext(rank=0, rank2host={0: "a"}))


def test_compute_n_blocks():
    cfg = TokenCacheBlockManagerConfig(CommunicationBackendManagerConfig(), n_layers=32, n_tokens=2, block_size=16, n_heads=32, head_size=128, dtype="fp16")
    per_block = 32 * 2 * 16 * 32 * 128 * 2  # 8 MiB
    assert per_block == 8 << 20
    assert TokenCacheBlockManager.compute_n_blocks(cfg, 10 * per_block + 5) == 10


def test_pinned_blocks_reuse():
    pool = _PinnedBlocks(6)
    pool.pin([0, 1, 2])
    pool.pin([1])
    pool.unpin([0, 1, 2])
    assert pool.get_num_avaiable_blocks() == 2 and sorted(pool.evictable) == [0, 2]
    got = pool.allocate(5)
    assert sorted(got) == [0, 2] and pool.get_num_avaiable_blocks() == 0
    with pytest.raises(AssertionError):
        pool.unpin([5])


class _Recorder:
    def __init__(self):
        self.calls = []

    def migrate_blocks(self, s, d, is_send):
        self.calls.append((s.rank, d.rank, is_send))


def test_backend_selection_by_host():
    ctx = CommunicationBackendManagerContext(migrate_stream=None, cache=torch.zeros(1, 1, 1, 1, 1, 1), n_blocks=1,
                                             rank2host={0: "hostA", 1: "hostA", 2: "hostB"})
    mgr = CommunicationBackendManager(CommunicationBackendManagerConfig(), ctx)
    mgr.intranode_backend, mgr.internode_backend = _Recorder(), _Recorder()
    v = lambda rank, n=4: VirtualTokenCache(vid=1, n_blocks_of_cache_manager=8, n_cache_tokens=n, rank=rank)
    mgr.migrate_blocks(v(0), v(1), False)
    mgr.migrate_blocks(v(0), v(2), True)
    mgr.migrate_blocks(v(0), v(7), False)  # unknown rank -> treated as another machine (communication.py:112-115)
    assert mgr.intranode_backend.calls == [(0, 1, False)]
    assert mgr.internode_backend.calls == [(0, 2, True), (0, 7, False)]
    with pytest.raises(AssertionError):
        mgr.migrate_blocks(v(0, 4), v(1, 5), False)
    assert mgr.in_same_machine(0, 1) and not mgr.in_same_machine(1, 2)


def _built_tables(requests, bs=16):
    builder = AttentionParametersBuilder(4, 4, 128, bs, torch.device("cpu"))
    for req in requests:
        builder.add_request(*req)
    builder.add_kv_cache(None)
    p = builder.build_attention_parameters()[0]
    return p.block_tables.tolist(), p.cu_blocks_lens.tolist(), p.new_cache_slots.tolist()


def test_block_table_image_cache_follows_the_list():
    """The builder keeps int32 images of the per-request block_table lists between steps (same list object every step, as
    in the reference engine); the metadata must follow every kind of change to those lists."""
    t0, t1 = [7, 3, 9], [5]
    assert _built_tables([(1, 40, [9 * 16 + 7], t0), (1, 3, [5 * 16 + 2], t1)]) == ([7, 3, 9, 5], [0, 3, 4], [151, 82])
    # unchanged lists: served from the cache
    assert _built_tables([(1, 40, [151], t0), (1, 3, [82], t1)])[0] == [7, 3, 9, 5]
    # grown in place by realloc (token_cache_manger.py:150-159)
    t0 += [11, 2]
    t1.append(6)
    assert _built_tables([(1, 70, [0], t0), (1, 17, [0], t1)])[:2] == ([7, 3, 9, 11, 2, 5, 6], [0, 5, 7])
    # shrunk, then edited in place at the same length: a stale image must never be used
    del t0[2:]
    assert _built_tables([(1, 20, [0], t0)])[0] == [7, 3]
    t0[1] = 99
    assert _built_tables([(1, 20, [0], t0)])[0] == [7, 99]
    # grown AND edited inside the old prefix
    t1[0] = 4
    t1.append(8)
    assert _built_tables([(1, 40, [0], t1)])[0] == [4, 6, 8]
    # an image handed out earlier is not modified by later growth of the list
    builder = AttentionParametersBuilder(4, 4, 128, 16, torch.device("cpu"))
    builder.add_request(1, 40, [0], t1)
    t1.append(1)
    builder.add_request(1, 50, [0], t1)
    builder.add_kv_cache(None)
    p = builder.build_attention_parameters()[0]
    assert p.block_tables.tolist() == [4, 6, 8, 4, 6, 8, 1] and p.cu_blocks_lens.tolist() == [0, 3, 7]
    # tuples / other sequences are accepted too (converted every time)
    assert _built_tables([(2, 20, (3, 4), (1, 2))])[0] == [1, 2]


def test_v2p_fast_paths_and_prefix_cache_stubs():
    """v2p keeps the reference's arithmetic (token_cache_manger.py:126-133) on every internal path; v2p_range is the same map
    over a contiguous id range; set_blocks / set (called by the reference's executor when prefix caching is on) are accepted."""
    from hydrainfer_b200.memory.token_cache import VirtualTokenCache

    class _Geometry:
        block_size = 16

    vc = VirtualTokenCache(vid=1, n_blocks_of_cache_manager=10, n_cache_tokens=100, block_table=[7, 2, 9, 4, 0, 1, 3], memory_handle=None, rank=0)
    ids = list(range(5, 100))
    want = [vc.block_table[v // 16] * 16 + v % 16 for v in ids]
    v2p, v2p_range = TokenCacheBlockManager.v2p, TokenCacheBlockManager.v2p_range
    assert v2p(_Geometry, vc, ids) == want                      # numpy path (>= 64 ids)
    assert v2p(_Geometry, vc, ids[:10]) == want[:10]            # short list
    assert v2p(_Geometry, vc, [33]) == [9 * 16 + 1]             # decode row
    assert list(v2p_range(_Geometry, vc, 5, 100)) == want and v2p_range(_Geometry, vc, 5, 100).typecode == "i"
    assert list(v2p_range(_Geometry, vc, 99, 100)) == want[-1:] and len(v2p_range(_Geometry, vc, 7, 7)) == 0
    TokenCacheBlockManager.set_blocks(_Geometry, vc, [0, 1], [123, 456])
    TokenCacheBlockManager.set(_Geometry, vc, [0, 1, 2], [1, 2, 3])
    with pytest.raises(AssertionError):
        TokenCacheBlockManager.set_blocks(_Geometry, vc, [0, 1], [123])
    # the int32 array goes into the builder unchanged
    b = AttentionParametersBuilder(4, 4, 64, 16, torch.device("cpu"))
    b.add_request(95, 100, v2p_range(_Geometry, vc, 5, 100), vc.block_table)
    assert list(b.new_cache_slots) == want
