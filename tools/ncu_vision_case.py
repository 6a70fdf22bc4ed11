"""One vision-attention launch pattern for ncu: a few calls of hi_varlen_attention on a named shape (see tools/bench_vision.py)."""
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "clip"
lens, heads, d = {"clip": ([577] * 64, 16, 64), "qwen": ([4096] * 4, 16, 80), "d128": ([2048] * 8, 16, 128)}[name]
dev = torch.device("cuda:0")
total = sum(lens)
q, k, v = (torch.randn(total, heads, d, device=dev).to(torch.bfloat16) for _ in range(3))
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
out = torch.empty_like(q)
for _ in range(6):
    mha_varlen_fwd(out, q, k, v, cu, cu, None, None, None, max(lens), max(lens), 1.0 / math.sqrt(d), 0, -1, -1, 0)
torch.cuda.synchronize()
