"""Dev probe (2 GPUs, torchrun): cudaMemcpyPeerAsync through a CUDA-IPC mapping at several sizes, and small migrations."""
import ctypes
import os
import statistics
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200 import _lib  # noqa: E402
from hydrainfer_b200._C.data_transfer import block_migration as bm  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
pool = torch.empty((28, 2, 300, 16, 4, 128), dtype=torch.bfloat16, device=dev).normal_()
handle = bm.get_ipc_mem_handle(pool)
handles = [None, None]
dist.all_gather_object(handles, handle)
dst = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=7):
    ts = []
    for i in range(reps + 2):
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if rank == 1:
            fn()
        e.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(s.elapsed_time(e))
    return statistics.median(ts), min(ts)


src_ptr = ctypes.c_void_p(0)
if rank == 1:
    h = handles[0]
    raw = (ctypes.c_uint8 * 64).from_buffer_copy(bytes(int(b) & 0xFF for b in h[:64]))
    _lib.check(_lib.lib.hi_ipc_open_handle(raw, 0, local, ctypes.byref(src_ptr)))
stream = _lib.current_stream_ptr(dev)
for mb in (1, 4, 14, 16, 32, 64, 128):
    n = mb << 20
    med, best = timed(lambda: _lib.check(_lib.lib.hi_peer_copy(dst.data_ptr(), local, src_ptr.value, local - 1, n, stream)))
    if rank == 1:
        print(f"peer copy {mb} MiB: median {med * 1e3:.1f} us ({n / med / 1e6:.0f} GB/s), best {best * 1e3:.1f} us")
for n_move in (1, 4, 16, 64, 256):
    g = torch.Generator().manual_seed(n_move)
    sb = torch.randperm(300, generator=g)[:n_move].tolist()
    db = torch.randperm(300, generator=g)[:n_move].tolist()
    med, best = timed(lambda: bm.migrate_blocks(sb, db, handles[0], pool, 300))
    if rank == 1:
        b = n_move * 28 * 2 * 16 * 4 * 128 * 2
        print(f"migrate {n_move} qwen blocks ({b >> 10} KiB): median {med * 1e3:.1f} us ({b / med / 1e6:.0f} GB/s), best {best * 1e3:.1f} us")
dist.barrier()
dist.destroy_process_group()
