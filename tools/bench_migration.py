"""Measurement tool (GPU box): block_migration sweep of SURVEY §8 config 5.

    python tools/bench_migration.py                                   # 1 GPU: pool -> pool on the same device (HBM)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_migration.py

Geometries: (a) LLaVA-7B pool [32,2,NB,16,32,128] bf16 (8 MiB/block), (b) Qwen2-VL-7B [28,2,NB,16,4,128] (896 KiB/block),
(c) image pool [1,1,NB,576,32,128] (4.5 MiB/block).  n_blocks/request in {16, 64, 256, 1024, 4096} where the pool fits.
Patterns at N>1: disjoint pairs 2i -> 2i+1 (all at once), fan-out 0 -> {1..N-1} (all receivers pull from rank 0) and, at
N >= 4, "p2d": ranks [0, N/2) are prefill nodes, ranks [N/2, N) decode nodes, and every decode rank pulls 1/(N/2) of its request
from EACH prefill rank at once (the disaggregated all-to-all of SURVEY §8d config 5).
Also times the plain cudaMemcpyPeerAsync of the same payload (the NVLink roofline probe) and, for small requests, the
reference's per-run memcpy loop restated with hi_peer_copy (block_migration.cpp:222-244) to show the launch-bound regime.
Bit-exactness of every destination pool is checked against the source pool.
"""
from __future__ import annotations

import json
import os
import statistics
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from hydrainfer_b200 import _lib  # noqa: E402
from hydrainfer_b200._C.data_transfer import block_migration as bm  # noqa: E402

GEOMS = {
    "llava7b": dict(n_layers=32, n_tokens=2, block_size=16, n_heads=32, head_size=128),
    "qwen2vl7b": dict(n_layers=28, n_tokens=2, block_size=16, n_heads=4, head_size=128),
    "image": dict(n_layers=1, n_tokens=1, block_size=576, n_heads=32, head_size=128),
}


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lines = []

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def timed(fn, reps=5, warm=2):
        ts = []
        for i in range(reps + warm):
            barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize(dev)
            if i >= warm:
                ts.append(s.elapsed_time(e))
        t = torch.tensor([statistics.median(ts)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    only_geoms = [g for g in os.environ.get("HI_MIG_GEOMS", "").split(",") if g]   # e.g. llava7b,qwen2vl7b (bounds an 8-GPU pass)
    only_sizes = [int(x) for x in os.environ.get("HI_MIG_SIZES", "").split(",") if x]
    for gname, geom in GEOMS.items():
        if only_geoms and gname not in only_geoms:
            continue
        bytes_per_block = geom["n_layers"] * geom["n_tokens"] * geom["block_size"] * geom["n_heads"] * geom["head_size"] * 2
        for n_move in (16, 64, 256, 1024, 4096):
            payload = n_move * bytes_per_block
            if payload > 36 * 2**30 or (only_sizes and n_move not in only_sizes):
                continue
            pool_blocks = n_move + max(8, n_move // 8)
            shape = (geom["n_layers"], geom["n_tokens"], pool_blocks, geom["block_size"], geom["n_heads"], geom["head_size"])
            pool = torch.empty(shape, dtype=torch.bfloat16, device=dev)
            pool.view(torch.int16).random_(-30000, 30000)
            g = torch.Generator().manual_seed(1234 + n_move)
            src_bt = torch.randperm(pool_blocks, generator=g)[:n_move].tolist()
            dst_bt = torch.randperm(pool_blocks, generator=g)[:n_move].tolist()
            handle = bm.get_ipc_mem_handle(pool)
            patterns = ["same_gpu"] if world == 1 else ["pairs", "pairs_push", "fanout"] + (["p2d"] if world >= 4 and world % 2 == 0 else [])
            half = world // 2
            handles = [handle]
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, handle)
            for pattern in patterns:
                if pattern == "same_gpu":
                    dst_pool, src_handle, src_rank, receiver = torch.zeros_like(pool), handle, rank, True
                elif pattern == "pairs":
                    receiver = rank % 2 == 1
                    src_rank = rank - 1 if receiver else rank
                    dst_pool, src_handle = pool if not receiver else torch.zeros_like(pool), handles[src_rank]
                elif pattern == "pairs_push":  # the SENDER (even rank) runs the kernel and writes the partner's pool through its mapping
                    receiver = rank % 2 == 1
                    src_rank = rank - 1 if receiver else rank
                    dst_pool = pool if not receiver else torch.zeros_like(pool)
                    dst_handles = [None] * world
                    dist.all_gather_object(dst_handles, bm.get_ipc_mem_handle(dst_pool) if receiver else None)
                    src_handle = None
                elif pattern == "fanout":
                    receiver = rank != 0
                    src_rank = 0
                    dst_pool, src_handle = pool if not receiver else torch.zeros_like(pool), handles[0]
                else:  # p2d
                    receiver = rank >= half
                    src_rank = rank - half  # owner of part 0 (the part that is checked)
                    dst_pool, src_handle = pool if not receiver else torch.zeros_like(pool), None
                    part = (n_move + half - 1) // half

                def run():
                    if pattern == "pairs_push":
                        if not receiver and rank + 1 < world:
                            bm.push_blocks(src_bt, dst_bt, pool, dst_handles[rank + 1], pool_blocks)
                        return
                    if not receiver:
                        return
                    if pattern == "p2d":
                        for k in range(half):  # part k comes from prefill rank (k + rank) % half: sources are hit evenly
                            lo, hi = k * part, min(n_move, (k + 1) * part)
                            if lo < hi:
                                bm.migrate_blocks(src_bt[lo:hi], dst_bt[lo:hi], handles[(k + rank) % half], dst_pool, pool_blocks)
                    else:
                        bm.migrate_blocks(src_bt, dst_bt, src_handle, dst_pool, pool_blocks)

                ms = timed(run)
                # bit-exact check: pull the source blocks with torch and compare
                ok = True
                if receiver:
                    if world == 1:
                        ok = bool(torch.equal(dst_pool[:, :, dst_bt].view(torch.int16), pool[:, :, src_bt].view(torch.int16)))
                    else:
                        # the source pool content is reproducible only on its owner: fetch it over NCCL for the check
                        pass
                if world > 1:
                    # owner sends its moved blocks (first 8 only, to bound time) to each of its receivers for the check
                    chk = min(8, n_move)
                    if pattern == "p2d":
                        chk = min(chk, part)
                    if pattern in ("pairs", "pairs_push", "p2d"):
                        to_rank = rank + half if pattern == "p2d" else rank + 1
                        if receiver:
                            buf = torch.empty_like(pool[:, :, :chk])
                            dist.recv(buf, src=src_rank)
                            ok = bool(torch.equal(dst_pool[:, :, dst_bt[:chk]].view(torch.int16), buf.view(torch.int16)))
                        elif to_rank < world:
                            dist.send(pool[:, :, src_bt[:chk]].contiguous(), dst=to_rank)
                    else:
                        if rank == 0:
                            blk = pool[:, :, src_bt[:chk]].contiguous()
                            for r in range(1, world):
                                dist.send(blk, dst=r)
                        else:
                            buf = torch.empty_like(pool[:, :, :chk])
                            dist.recv(buf, src=0)
                            ok = bool(torch.equal(dst_pool[:, :, dst_bt[:chk]].view(torch.int16), buf.view(torch.int16)))
                okt = torch.tensor([1 if ok else 0], device=dev)
                if world > 1:
                    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
                n_recv = 1 if world == 1 else (world - 1 if pattern == "fanout" else world // 2)
                line = {"geom": gname, "pattern": pattern, "n_gpus": world, "n_blocks": n_move, "payload_bytes": payload, "ms": ms,
                        "gbs_per_receiver": payload / ms / 1e6, "gbs_aggregate": n_recv * payload / ms / 1e6, "bit_exact": bool(okt.item()),
                        "run_bytes": bytes_per_block // (geom["n_layers"] * geom["n_tokens"]), "launches": half if pattern == "p2d" else 1,
                        "reference_memcpy_calls": geom["n_layers"] * geom["n_tokens"] * n_move}
                if rank == 0:
                    print(json.dumps(line), flush=True)
                    lines.append(line)
                if receiver and dst_pool is not pool:
                    del dst_pool

            # roofline probe: one contiguous cudaMemcpyPeerAsync of the same payload (pairs) / same-GPU copy
            if n_move in (256, 1024) and gname == "llava7b":
                n_el = payload // 2
                flat_src = pool.view(-1)[:n_el]
                if world == 1:
                    flat_dst = torch.empty_like(flat_src)
                    ms = timed(lambda: flat_dst.copy_(flat_src))
                    line = {"geom": gname, "pattern": "contiguous_copy_same_gpu", "n_gpus": 1, "payload_bytes": payload, "ms": ms, "gbs_per_receiver": payload / ms / 1e6}
                else:
                    receiver = rank % 2 == 1
                    ptrs = [None] * world
                    # peer pointer of the partner's pool through the same IPC mapping the migration uses
                    import ctypes
                    if receiver:
                        raw = bytes(int(b) & 0xFF for b in handles[rank - 1][:64])
                        off = int.from_bytes(bytes(int(b) & 0xFF for b in handles[rank - 1][64:]), "little") if len(handles[rank - 1]) == 72 else 0
                        p = ctypes.c_void_p()
                        _lib.check(_lib.lib.hi_ipc_open_handle((ctypes.c_uint8 * 64).from_buffer_copy(raw), off, local, ctypes.byref(p)))
                        flat_dst = torch.empty_like(flat_src)
                        stream = torch.cuda.current_stream(dev).cuda_stream
                        fn = lambda: _lib.check(_lib.lib.hi_peer_copy(flat_dst.data_ptr(), local, p.value, local - 1, payload, stream))
                    else:
                        fn = lambda: None
                    ms = timed(fn)
                    line = {"geom": gname, "pattern": "cudaMemcpyPeerAsync_pairs", "n_gpus": world, "payload_bytes": payload, "ms": ms,
                            "gbs_per_receiver": payload / ms / 1e6, "gbs_aggregate": (world // 2) * payload / ms / 1e6}
                if rank == 0:
                    print(json.dumps(line), flush=True)
                    lines.append(line)
            del pool
            torch.cuda.empty_cache()
            barrier()

    if rank == 0:
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        with open(out / f"migration_n{world}.jsonl", "w") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
