"""Dev tool (GPU box, library built with HI_BUILD_DEFINES=-DHI_MBAR_DEBUG): run small pair-kernel cases and report the
first mbarrier wait that timed out (barrier index, parity, thread, block) instead of trapping."""
import ctypes
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200 import _lib  # noqa: E402
from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd  # noqa: E402
from hydrainfer_b200.workloads import make_batch  # noqa: E402
from oracle import paged_kv_oracle as oracle  # noqa: E402

NK = NV = 4
KBARS = 2 * 32768 + (NK + NV) * 16384
NAMES = (["qfull0", "qfull1"] + [f"kfull{i}" for i in range(NK)] + [f"kempty{i}" for i in range(NK)] + [f"vfull{i}" for i in range(NV)]
         + [f"vempty{i}" for i in range(NV)] + ["sfull00", "sfull01", "sfull10", "sfull11", "pfull00", "pfull01", "pfull10", "pfull11", "pvdone0", "pvdone1",
                                                 "ofull0", "ofull1", "vtail", "qempty0", "qempty1", "oempty0", "oempty1", "itemfull0", "itemfull1", "itemfull2", "itemfull3", "itemempty0", "itemempty1", "itemempty2", "itemempty3"])
cases = [("q128 kv128", [(128, 128)], 1, 1), ("2 heads q128 kv128", [(128, 128)], 2, 2), ("mha4 ragged", [(300, 300), (1, 77), (64, 500), (700, 1500)], 4, 4), ("q256 kv256", [(256, 256)], 1, 1), ("q300 kv300", [(300, 300)], 1, 1),
         ("q700 kv1500", [(700, 1500)], 1, 1), ("gqa7", [(1, 300), (40, 170), (1, 17), (200, 513)], 28, 4)]
dev = "cuda:0"
fn = _lib.lib.hi_debug_mbar_timeout
fn.argtypes = [ctypes.POINTER(ctypes.c_uint * 64)]
for label, seq_lens, hq, hkv in cases:
    d = 128
    batch = make_batch(seq_lens, hq, hkv, d, 16, dtype=torch.bfloat16, seed=1)
    fp32 = oracle.paged_attention_fp32(batch.query.view(-1, hq, d), batch.key_cache, batch.value_cache, batch.q_cu_seq_lens, batch.kv_cu_seq_lens,
                                       torch.tensor(batch.block_tables, dtype=torch.int32), batch.cu_blocks_lens, hq, hkv, d)
    t = batch.n_tokens
    q3 = batch.query.to(dev).view(t, hq, d)
    out = torch.full_like(q3, float("nan"))
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    mha_varlen_fwd(out, q3, batch.key_cache.to(dev), batch.value_cache.to(dev), i32(batch.q_cu_seq_lens), i32(batch.kv_cu_seq_lens),
                   i32(batch.block_tables), i32(batch.cu_blocks_lens), None, batch.q_max, batch.kv_max, 1 / math.sqrt(d), 0, -1, 0, 0, 4)
    buf = (ctypes.c_uint * 64)()
    rc = fn(ctypes.byref(buf))
    err = (out.float().cpu().reshape(t, -1) - fp32).abs().max().item()
    print(f"{label:14s} rc={rc} max_err={err:.3e}", flush=True)
    for w in range(16):
        if buf[4 * w]:
            off = buf[4 * w + 1] % (1 << 20)
            idx = ((off - KBARS) % 1024) // 8  # barriers sit right after the staging buffers in the 1024-aligned block
            print(f"    warp {w:2d}: TIMEOUT {NAMES[idx] if idx < len(NAMES) else idx} parity={buf[4 * w + 2]} block=({buf[4 * w + 3] & 0xffff},{buf[4 * w + 3] >> 16})", flush=True)
