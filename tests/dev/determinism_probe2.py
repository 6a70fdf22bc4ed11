"""Dev probe (GPU box): run-to-run bit-determinism of the pair kernel under env variations."""
import math
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd  # noqa: E402
from hydrainfer_b200.workloads import make_batch  # noqa: E402

dev = torch.device("cuda:0")
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
seqs = eval(os.environ.get("SEQS", "[(300, 4000), (64, 64)]"))
pre = make_batch(seqs, 28, 4, 128, 16, dtype=torch.bfloat16, seed=32).to(dev)
args = (pre.query.view(pre.n_tokens, 28, 128), pre.key_cache, pre.value_cache, i32(pre.q_cu_seq_lens), i32(pre.kv_cu_seq_lens), i32(pre.block_tables),
        i32(pre.cu_blocks_lens), None, pre.q_max, pre.kv_max, 1 / math.sqrt(128), 0, -1, 0, 0, int(os.environ.get("PATH_ID", "4")))
outs = []
for i in range(6):
    out = torch.empty_like(args[0])
    mha_varlen_fwd(out, *args)
    torch.cuda.synchronize()
    outs.append(out)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("HI_") or k in ("SEQS", "PATH_ID"))
eq = [bool(torch.equal(outs[0], o)) for o in outs[1:]]
bad = (outs[0] != outs[1])
rows = bad.any(dim=2).any(dim=1).nonzero().flatten().tolist()
heads = bad.any(dim=2).any(dim=0).nonzero().flatten().tolist()
print(f"[{tag}] vs run0: {eq}  max|d| {(outs[0].float() - outs[1].float()).abs().max().item():.2e} rows {rows[:8]} n_rows {len(rows)} heads {heads[:10]}")
