"""GPU: block_migration (IPC handle plumbing + the gather kernel) vs the oracle restatement, bit-exact, including the
regions of the destination pool that must stay untouched.  The reference has no migration test: geometry and
patterns follow SURVEY §8 config 5 scaled to test size."""
import os
import socket

import pytest
import torch

from oracle import paged_kv_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bm():
    from hydrainfer_b200._C.data_transfer import block_migration
    return block_migration


def _pools(shape_src, shape_dst, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(*shape_src, generator=g).to(dtype)
    dst = torch.randn(*shape_dst, generator=g).to(dtype)
    return src, dst


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("geom", [
    # (n_layers, n_tokens, block_size, n_heads, head_size, src_blocks, dst_blocks, n_move)
    (4, 2, 16, 4, 128, 40, 33, 1),      # GQA-like 16 KiB runs
    (4, 2, 16, 4, 128, 40, 33, 16),
    (3, 2, 16, 32, 128, 24, 30, 20),    # LLaVA-like 128 KiB runs
    (1, 1, 576, 2, 64, 6, 5, 4),        # image pool: one plane, 144 KiB runs
    (2, 2, 4, 1, 8, 300, 280, 257),     # tiny 64-byte runs, many blocks
])
def test_migrate_blocks_same_device(geom, dtype):
    bm = _bm()
    L, T, bs, H, d, nb_src, nb_dst, n = geom
    src, dst = _pools((L, T, nb_src, bs, H, d), (L, T, nb_dst, bs, H, d), dtype, seed=n)
    g = torch.Generator().manual_seed(99)
    src_bt = torch.randperm(nb_src, generator=g)[:n].tolist()
    dst_bt = torch.randperm(nb_dst, generator=g)[:n].tolist()
    ref = dst.clone()
    oracle.migrate_blocks(src_bt, dst_bt, src, ref)
    src_d, dst_d = src.to(DEV), dst.to(DEV)
    handle = bm.get_ipc_mem_handle(src_d)
    assert len(handle) in (64, 72) and all(0 <= b < 256 for b in handle)
    bm.migrate_blocks(src_bt, dst_bt, handle, dst_d, nb_src)
    torch.cuda.synchronize()
    assert torch.equal(dst_d.cpu(), ref), "destination pool differs from the oracle (moved blocks or untouched regions)"
    assert torch.equal(src_d.cpu(), src), "source pool was modified"


@pytest.mark.parametrize("geom", [
    (4, 2, 16, 4, 128, 40, 33, 16),     # 16-KiB runs = exactly one piece, block tables inside the kernel parameters
    (3, 2, 16, 32, 128, 24, 30, 20),    # 128-KiB runs = 8 pieces
    (1, 1, 576, 2, 64, 6, 5, 4),        # 144-KiB runs = 9 pieces
    (2, 2, 4, 1, 8, 300, 280, 257),     # 64-byte runs: bulk copies of 64 bytes
    (2, 2, 16, 3, 40, 700, 650, 500),   # 3840-byte runs (not a power of two), more blocks than an inline table holds: device tables
])
def test_bulk_copy_kernel_same_device(geom, monkeypatch):
    """The TMA bulk-copy kernel (chosen by itself only for large transfers from / to peer memory) forced onto ordinary launches:
    same bytes as the oracle, untouched regions untouched."""
    monkeypatch.setenv("HI_MIGRATE_BULK", "2")
    bm = _bm()
    L, T, bs, H, d, nb_src, nb_dst, n = geom
    src, dst = _pools((L, T, nb_src, bs, H, d), (L, T, nb_dst, bs, H, d), torch.bfloat16, seed=n + 1)
    g = torch.Generator().manual_seed(98)
    src_bt = torch.randperm(nb_src, generator=g)[:n].tolist()
    dst_bt = torch.randperm(nb_dst, generator=g)[:n].tolist()
    ref = dst.clone()
    oracle.migrate_blocks(src_bt, dst_bt, src, ref)
    src_d, dst_d = src.to(DEV), dst.to(DEV)
    bm.migrate_blocks(src_bt, dst_bt, bm.get_ipc_mem_handle(src_d), dst_d, nb_src)
    torch.cuda.synchronize()
    assert torch.equal(dst_d.cpu(), ref), "destination pool differs from the oracle (moved blocks or untouched regions)"
    # a layer range through the same kernel
    dst_d2 = dst.to(DEV)
    bm.migrate_blocks_layers(src_bt, dst_bt, bm.get_ipc_mem_handle(src_d), dst_d2, nb_src, 0, L)
    torch.cuda.synchronize()
    assert torch.equal(dst_d2.cpu(), ref)


def test_layer_ranges_compose_to_the_whole_request():
    """migrate_blocks_layers over a partition of the layers == one migrate_blocks; each call touches only its layers."""
    bm = _bm()
    L, T, bs, H, d, nb_src, nb_dst, n = 6, 2, 16, 4, 128, 30, 25, 11
    src, dst = _pools((L, T, nb_src, bs, H, d), (L, T, nb_dst, bs, H, d), torch.bfloat16, seed=4)
    g = torch.Generator().manual_seed(5)
    src_bt = torch.randperm(nb_src, generator=g)[:n].tolist()
    dst_bt = torch.randperm(nb_dst, generator=g)[:n].tolist()
    ref = dst.clone()
    oracle.migrate_blocks(src_bt, dst_bt, src, ref)
    src_d, dst_d = src.to(DEV), dst.to(DEV)
    handle = bm.get_ipc_mem_handle(src_d)
    bm.migrate_blocks_layers(src_bt, dst_bt, handle, dst_d, nb_src, 2, 5)
    torch.cuda.synchronize()
    got = dst_d.cpu()
    assert torch.equal(got[2:5], ref[2:5]) and torch.equal(got[:2], dst[:2]) and torch.equal(got[5:], dst[5:])
    bm.migrate_blocks_layers(src_bt, dst_bt, handle, dst_d, nb_src, 0, 2)
    bm.migrate_blocks_layers(src_bt, dst_bt, handle, dst_d, nb_src, 5, 6)
    bm.migrate_blocks_layers(src_bt, dst_bt, handle, dst_d, nb_src, 3, 3)  # empty range: no-op
    torch.cuda.synchronize()
    assert torch.equal(dst_d.cpu(), ref)
    with pytest.raises(RuntimeError):
        bm.migrate_blocks_layers(src_bt, dst_bt, handle, dst_d, nb_src, 4, 7)


def test_push_from_the_sender_matches_pull():
    bm = _bm()
    L, T, bs, H, d, nb_src, nb_dst, n = 3, 2, 16, 8, 128, 20, 26, 9
    src, dst = _pools((L, T, nb_src, bs, H, d), (L, T, nb_dst, bs, H, d), torch.float16, seed=8)
    g = torch.Generator().manual_seed(6)
    src_bt = torch.randperm(nb_src, generator=g)[:n].tolist()
    dst_bt = torch.randperm(nb_dst, generator=g)[:n].tolist()
    ref = dst.clone()
    oracle.migrate_blocks(src_bt, dst_bt, src, ref)
    src_d, dst_d = src.to(DEV), dst.to(DEV)
    bm.push_blocks(src_bt, dst_bt, src_d, bm.get_ipc_mem_handle(dst_d), nb_dst)
    torch.cuda.synchronize()
    assert torch.equal(dst_d.cpu(), ref) and torch.equal(src_d.cpu(), src)


def test_layer_pipelined_backend_events():
    """IPCHandleMemoryBackend.migrate_layers: per-layer pulls gated on producer events, completion events instead of a stream sync."""
    from hydrainfer_b200.memory.communication import IPCHandleMemoryBackend
    from hydrainfer_b200.memory.token_cache import VirtualTokenCache
    L, T, bs, H, d, nb = 4, 2, 16, 4, 128, 12
    src = torch.zeros(L, T, nb, bs, H, d, device=DEV, dtype=torch.bfloat16)
    dst = torch.zeros(L, T, nb, bs, H, d, device=DEV, dtype=torch.bfloat16)
    bm = _bm()
    src_vc = VirtualTokenCache(vid=0, n_blocks_of_cache_manager=nb, n_cache_tokens=3 * bs, block_table=[7, 2, 9], memory_handle=bm.get_ipc_mem_handle(src), rank=0)
    dst_vc = VirtualTokenCache(vid=1, n_blocks_of_cache_manager=nb, n_cache_tokens=3 * bs, block_table=[1, 0, 5], memory_handle=None, rank=0)
    backend = IPCHandleMemoryBackend(torch.cuda.Stream(), dst, nb)
    compute = torch.cuda.Stream()
    done = []
    for layer in range(L):  # the "prefill" writes layer l on its own stream, then that layer is shipped while l + 1 is produced
        with torch.cuda.stream(compute):
            src[layer, :, [7, 2, 9]] = float(layer + 1)
            ready = torch.cuda.Event()
            ready.record(compute)
        done.append(backend.migrate_layers(src_vc, dst_vc, layer, layer + 1, wait_event=ready))
    for ev in done:
        ev.synchronize()
    for layer in range(L):
        assert bool((dst[layer, :, [1, 0, 5]] == layer + 1).all()), f"layer {layer} did not arrive after its producer"
    untouched = [b for b in range(nb) if b not in (1, 0, 5)]
    assert bool((dst[:, :, untouched] == 0).all())


def test_round_trip_restores_blocks():
    bm = _bm()
    L, T, bs, H, d = 4, 2, 16, 8, 128
    a = torch.randn(L, T, 50, bs, H, d, device=DEV).to(torch.bfloat16)
    b = torch.zeros(L, T, 64, bs, H, d, device=DEV, dtype=torch.bfloat16)
    c = torch.zeros_like(a)
    g = torch.Generator().manual_seed(1)
    a_bt = torch.randperm(50, generator=g)[:32].tolist()
    b_bt = torch.randperm(64, generator=g)[:32].tolist()
    bm.migrate_blocks(a_bt, b_bt, bm.get_ipc_mem_handle(a), b, 50)
    bm.migrate_blocks(b_bt, a_bt, bm.get_ipc_mem_handle(b), c, 64)
    torch.cuda.synchronize()
    assert torch.equal(c[:, :, a_bt], a[:, :, a_bt])
    rest = [i for i in range(50) if i not in a_bt]
    assert c[:, :, rest].abs().max().item() == 0


def test_pool_inside_a_larger_allocation_carries_an_offset():
    bm = _bm()
    big = torch.randn(2 * 2 * 9 * 16 * 2 * 64 + 4096, device=DEV).to(torch.float16)
    src = big[4096:].view(2, 2, 9, 16, 2, 64)  # does not start at the allocation base
    dst = torch.zeros(2, 2, 4, 16, 2, 64, device=DEV, dtype=torch.float16)
    handle = bm.get_ipc_mem_handle(src)
    assert len(handle) == 72, "a pool at a non-zero offset must carry the offset in the handle list"
    bm.migrate_blocks([8, 0], [1, 3], handle, dst, 9)
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :, 1], src[:, :, 8]) and torch.equal(dst[:, :, 3], src[:, :, 0])
    assert dst[:, :, [0, 2]].abs().max().item() == 0


def test_argument_errors_raise():
    bm = _bm()
    pool = torch.zeros(1, 2, 4, 16, 2, 64, device=DEV, dtype=torch.float16)
    h = bm.get_ipc_mem_handle(pool)
    with pytest.raises(RuntimeError, match="length"):
        bm.migrate_blocks([0, 1], [0], h, pool, 4)
    with pytest.raises(RuntimeError, match="out of range"):
        bm.migrate_blocks([0], [7], h, pool, 4)
    with pytest.raises(RuntimeError, match="contiguous"):
        bm.migrate_blocks([0], [0], h, pool[:, :, ::2], 4)
    with pytest.raises(RuntimeError, match="64"):
        bm.migrate_blocks([0], [0], h[:10], pool, 4)
    bm.migrate_blocks([], [], h, pool, 4)  # empty request is a no-op


def test_block_manager_end_to_end_migration():
    """Prefill node -> decode node through the reference-shaped manager API (token_cache_manger.py + communication.py)."""
    from hydrainfer_b200.memory import (CommunicationBackendManagerConfig, TokenCacheBlockManager, TokenCacheBlockManagerConfig,
                                        TokenCacheBlockManagerContext)
    def mk(rank, n_blocks):
        cfg = TokenCacheBlockManagerConfig(CommunicationBackendManagerConfig(), n_layers=3, n_tokens=2, n_blocks=n_blocks, block_size=16,
                                           n_heads=4, head_size=128, dtype="fp16", device=DEV)
        return TokenCacheBlockManager(cfg, TokenCacheBlockManagerContext(rank=rank, rank2host={0: "node", 1: "node"}))
    prefill, decode = mk(0, 20), mk(1, 12)
    assert len(prefill.memory_handle) in (64, 72)
    src = prefill.allocate_virtual_cache()
    prefill.realloc(src, 70)                      # 5 blocks, LIFO: [4, 3, 2, 1, 0]
    assert src.block_table == [4, 3, 2, 1, 0] and src.n_cache_tokens == 70
    assert prefill.v2p(src, [0, 15, 16, 69]) == [64, 79, 48, 5]
    dst = decode.allocate_virtual_cache()
    decode.realloc(dst, 70)
    before = decode.cache_tensor.clone()
    decode.migrate_blocks(src, dst, is_send=False)
    prefill.migrate_blocks(src, dst, is_send=True)  # no-op on the IPC path
    done = decode.record_migration_done()
    decode.synchronize()
    assert done.query()
    for s, d in zip(src.block_table, dst.block_table):
        assert torch.equal(decode.cache_tensor[:, :, d], prefill.cache_tensor[:, :, s])
    untouched = [i for i in range(12) if i not in dst.block_table]
    assert torch.equal(decode.cache_tensor[:, :, untouched], before[:, :, untouched])
    # shrink releases tail blocks for reuse (realloc :155-158)
    prefill.realloc(src, 20)
    assert src.block_table == [4, 3] and prefill.get_num_avaiable_blocks() == 15 + 3
    layer = decode.get_layer_cache(1).get_caches()
    assert len(layer) == 2 and layer[0].shape == (12, 16, 4, 128) and layer[0].data_ptr() == decode.cache_tensor[1, 0].data_ptr()


# ---- multi-GPU ---------------------------------------------------------------------------------------------------------
needs_2gpu = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")


@needs_2gpu
def test_same_process_peer_device_migration():
    bm = _bm()
    src = torch.randn(2, 2, 16, 16, 4, 128, device="cuda:0").to(torch.bfloat16)
    dst = torch.zeros(2, 2, 16, 16, 4, 128, device="cuda:1", dtype=torch.bfloat16)
    handle = bm.get_ipc_mem_handle(src)
    bm.migrate_blocks([3, 5, 7], [0, 1, 2], handle, dst, 16)
    torch.cuda.synchronize("cuda:1")
    assert torch.equal(dst[:, :, [0, 1, 2]].cpu(), src[:, :, [3, 5, 7]].cpu())


@needs_2gpu
def test_same_process_peer_device_push():
    """Sender-side variant over NVLink: the kernel runs on cuda:0 and writes cuda:1's pool through peer access."""
    bm = _bm()
    src = torch.randn(2, 2, 16, 16, 4, 128, device="cuda:0").to(torch.bfloat16)
    dst = torch.zeros(2, 2, 12, 16, 4, 128, device="cuda:1", dtype=torch.bfloat16)
    bm.push_blocks([3, 5, 7], [0, 11, 2], src, bm.get_ipc_mem_handle(dst), 12, 1, 2)
    torch.cuda.synchronize("cuda:0")
    assert torch.equal(dst[1][:, [0, 11, 2]].cpu(), src[1][:, [3, 5, 7]].cpu())
    assert dst[0].abs().max().item() == 0


def _ipc_worker(rank, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=2, device_id=torch.device(f"cuda:{rank}"))
    try:
        from hydrainfer_b200._C.data_transfer import block_migration as bm
        from hydrainfer_b200.memory import NCCLBackend, VirtualTokenCache
        g = torch.Generator().manual_seed(5)
        ref_src = torch.randn(2, 2, 10, 16, 4, 128, generator=g).to(torch.bfloat16)  # both ranks can rebuild rank 0's pool
        pool = ref_src.to(f"cuda:{rank}") if rank == 0 else torch.zeros(2, 2, 8, 16, 4, 128, device=f"cuda:{rank}", dtype=torch.bfloat16)
        handle = [bm.get_ipc_mem_handle(pool) if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        src_vc = VirtualTokenCache(vid=1, n_blocks_of_cache_manager=10, n_cache_tokens=48, block_table=[9, 2, 4], memory_handle=handle[0], rank=0)
        dst_vc = VirtualTokenCache(vid=2, n_blocks_of_cache_manager=8, n_cache_tokens=48, block_table=[7, 0, 3], rank=1)
        if rank == 1:  # pull through the CUDA-IPC peer mapping (NVLink)
            assert bm.register_ipc_mem_handle(handle[0]) >= 0
            bm.migrate_blocks(src_vc.block_table, dst_vc.block_table, handle[0], pool, 10)
            torch.cuda.synchronize()
            assert torch.equal(pool[:, :, [7, 0, 3]].cpu(), ref_src[:, :, [9, 2, 4]])
            pool.zero_()
        dist.barrier()
        # packed NCCL path (cross-host backend) between the same two ranks
        backend = NCCLBackend(torch.cuda.Stream(), pool)
        backend.migrate_blocks(src_vc, dst_vc, is_send=(rank == 0))
        backend.migrate_stream.synchronize()
        if rank == 1:
            assert torch.equal(pool[:, :, [7, 0, 3]].cpu(), ref_src[:, :, [9, 2, 4]])
            assert pool[:, :, [1, 2, 4, 5, 6]].abs().max().item() == 0
        dist.barrier()
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@needs_2gpu
def test_cross_process_ipc_and_packed_nccl(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_ipc_worker, args=(port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
