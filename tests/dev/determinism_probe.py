"""Dev probe (GPU box): is a split prefill bit-identical across streams / under a concurrent decode stream?"""
import math
import sys
import threading
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd  # noqa: E402
from hydrainfer_b200.workloads import make_batch  # noqa: E402

dev = torch.device("cuda:0")
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)


def args_of(b):
    return (b.query.view(b.n_tokens, b.n_qo_heads, 128), b.key_cache, b.value_cache, i32(b.q_cu_seq_lens), i32(b.kv_cu_seq_lens), i32(b.block_tables),
            i32(b.cu_blocks_lens), None, b.q_max, b.kv_max, 1 / math.sqrt(128), 0, -1, 0, 0)


pre = make_batch([(300, 4000), (64, 64)], 28, 4, 128, 16, dtype=torch.bfloat16, seed=32).to(dev)
dec = make_batch([(1, 3000)] * 4, 32, 32, 128, 16, dtype=torch.bfloat16, seed=31).to(dev)
pre_args, dec_args = args_of(pre), args_of(dec)
ref = torch.empty_like(pre_args[0])
mha_varlen_fwd(ref, *pre_args)
torch.cuda.synchronize()


def diff(out, what):
    d = (out.float() - ref.float()).abs()
    bad = (out != ref)
    rows = bad.any(dim=2).any(dim=1).nonzero().flatten().tolist()
    print(f"{what}: equal={bool(torch.equal(out, ref))} max|d|={d.max().item():.3e} differing elems={int(bad.sum())} rows={rows[:12]}{'...' if len(rows) > 12 else ''}")


for i in range(3):
    out = torch.empty_like(ref)
    mha_varlen_fwd(out, *pre_args)
    torch.cuda.synchronize()
    diff(out, f"default stream run {i}")
s = torch.cuda.Stream(dev)
with torch.cuda.stream(s):
    for i in range(3):
        out = torch.empty_like(ref)
        mha_varlen_fwd(out, *pre_args)
        s.synchronize()
        diff(out, f"side stream run {i}")
stop = False


def engine():
    s2 = torch.cuda.Stream(dev)
    with torch.cuda.stream(s2):
        while not stop:
            o = torch.empty_like(dec_args[0])
            mha_varlen_fwd(o, *dec_args)
            s2.synchronize()


th = threading.Thread(target=engine)
th.start()
with torch.cuda.stream(s):
    for i in range(6):
        out = torch.empty_like(ref)
        mha_varlen_fwd(out, *pre_args)
        s.synchronize()
        diff(out, f"side stream + concurrent decode run {i}")
stop = True
th.join()
