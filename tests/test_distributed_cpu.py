"""CPU, world_size 2 over gloo: the host logic of the N>1 paths — round-robin sequence sharding (attention shards by
sequence with no data-path collective) and the packed send/recv migration protocol of NCCLBackend."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hydrainfer_b200.workloads import shard_round_robin


def test_round_robin_partition_is_exact():
    for n in (0, 1, 7, 64, 257):
        for world in (1, 2, 4, 8):
            owned = [shard_round_robin(n, r, world) for r in range(world)]
            flat = sorted(i for part in owned for i in part)
            assert flat == list(range(n))
            assert max(len(p) for p in owned) - min(len(p) for p in owned) <= 1


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, result_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import paged_kv_oracle as oracle
        from hydrainfer_b200.memory import NCCLBackend, VirtualTokenCache

        # every rank owns a pool with a different number of blocks
        n_blocks = 6 + rank
        g = torch.Generator().manual_seed(10 + rank)
        pool = torch.randn(2, 2, n_blocks, 4, 2, 8, generator=g)
        before = pool.clone()
        gather = lambda src, dst, sb, db: oracle.migrate_blocks(sb, db, src, dst)
        backend = NCCLBackend(migrate_stream=None, cache=pool, gather_fn=gather)

        # rank 0 (prefill role) sends blocks [4, 1, 3] to rank 1 (decode role), which stores them at [0, 6, 2]
        src_vc = VirtualTokenCache(vid=1, n_blocks_of_cache_manager=6, n_cache_tokens=10, block_table=[4, 1, 3], rank=0)
        dst_vc = VirtualTokenCache(vid=9, n_blocks_of_cache_manager=7, n_cache_tokens=10, block_table=[0, 6, 2], rank=1)
        # the control plane pickles virtual caches between processes (epdnode.py:412-442)
        box = [src_vc if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert box[0] == src_vc
        backend.migrate_blocks(src_vc, dst_vc, is_send=(rank == 0))

        # ship rank 0's pool to rank 1 for the check
        ref = torch.zeros(2, 2, 6, 4, 2, 8)
        if rank == 0:
            assert torch.equal(pool, before), "the sender's pool must not change"
            dist.send(pool, dst=1)
        else:
            dist.recv(ref, src=0)
            for s, d in zip([4, 1, 3], [0, 6, 2]):
                assert torch.equal(pool[:, :, d], ref[:, :, s])
            for d in (1, 3, 4, 5):
                assert torch.equal(pool[:, :, d], before[:, :, d]), "an untouched block changed"
        # weak-scaling aggregation used by bench.py: MAX of per-rank time, SUM of per-rank tokens
        t = torch.tensor([1.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = torch.tensor([64.0])
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        assert t.item() == 2.0 and n.item() == 128.0
        open(os.path.join(result_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_packed_migration_protocol_over_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
