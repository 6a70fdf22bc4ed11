"""Dev tool (GPU box): small attention cases per kernel path with error breakdowns, each group in its own subprocess so a
trapped kernel cannot take the other groups down.  Usage: python tests/dev/first_light.py [group]"""
import math
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

GROUPS = {
    "simt": dict(path=1, cases=[
        ("mha decode", [(1, 1), (1, 16), (1, 17), (1, 300)], 8, 8),
        ("gqa4 mixed", [(1, 100), (15, 15), (40, 234)], 8, 2),
        ("gqa7", [(1, 40), (9, 33)], 28, 4),
    ]),
    "tc_basic": dict(path=2, cases=[
        ("1 head q128 kv128", [(128, 128)], 1, 1),
        ("1 head q1 kv128", [(1, 128)], 1, 1),
        ("1 head q1 kv16", [(1, 16)], 1, 1),
        ("1 head q128 kv256", [(128, 256)], 1, 1),
        ("1 head q16 kv40", [(16, 40)], 1, 1),
        ("1 head q300 kv300", [(300, 300)], 1, 1),
    ]),
    "dec": dict(path=3, cases=[
        ("gqa8 one row kv128", [(1, 128)], 8, 1),
        ("gqa8 one row kv16", [(1, 16)], 8, 1),
        ("gqa8 one row kv300", [(1, 300)], 8, 1),
        ("gqa4 ragged", [(1, 100), (1, 15), (1, 234), (1, 1024)], 8, 2),
        ("gqa7 qwen mixed", [(1, 300), (40, 170), (1, 17)], 28, 4),
        ("gqa16", [(1, 2000)] * 2, 32, 2),
        ("mha", [(1, 500), (3, 140)], 4, 4),
    ]),
    "pair": dict(path=4, cases=[
        ("1 head q128 kv128", [(128, 128)], 1, 1),
        ("1 head q256 kv256", [(256, 256)], 1, 1),
        ("1 head q1 kv16", [(1, 16)], 1, 1),
        ("1 head q16 kv40", [(16, 40)], 1, 1),
        ("1 head q300 kv300", [(300, 300)], 1, 1),
        ("1 head q700 kv1500", [(700, 1500)], 1, 1),
        ("mha 8 heads mixed", [(1, 100), (15, 15), (111, 234), (1, 1024)], 8, 8),
        ("gqa4", [(1, 100), (15, 15), (111, 234), (1, 1024)], 8, 2),
        ("gqa7 qwen", [(1, 300), (40, 170), (1, 17), (200, 513)], 28, 4),
    ]),
    "tc_gqa": dict(path=2, cases=[
        ("mha 8 heads mixed", [(1, 100), (15, 15), (111, 234), (1, 1024)], 8, 8),
        ("gqa4", [(1, 100), (15, 15), (111, 234), (1, 1024)], 8, 2),
        ("gqa7 qwen", [(1, 300), (40, 170), (1, 17), (200, 513)], 28, 4),
        ("gqa8 decode", [(1, 2000)] * 4, 64, 8),
    ]),
}


def run_group(name: str) -> None:
    import torch
    from hydrainfer_b200.workloads import make_batch
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    from oracle import paged_kv_oracle as oracle
    spec = GROUPS[name]
    dev = "cuda:0"
    for label, seq_lens, hq, hkv in spec["cases"]:
        d = 128
        batch = make_batch(seq_lens, hq, hkv, d, 16, dtype=torch.bfloat16, seed=1)
        fp32 = oracle.paged_attention_fp32(batch.query.view(-1, hq, d), batch.key_cache, batch.value_cache, batch.q_cu_seq_lens, batch.kv_cu_seq_lens,
                                           torch.tensor(batch.block_tables, dtype=torch.int32), batch.cu_blocks_lens, hq, hkv, d)
        t = batch.n_tokens
        q3 = batch.query.to(dev).view(t, hq, d)
        out = torch.full_like(q3, float("nan"))
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        mha_varlen_fwd(out, q3, batch.key_cache.to(dev), batch.value_cache.to(dev), i32(batch.q_cu_seq_lens), i32(batch.kv_cu_seq_lens),
                       i32(batch.block_tables), i32(batch.cu_blocks_lens), None, batch.q_max, batch.kv_max, 1 / math.sqrt(d), 0, -1, 0, 0, spec["path"])
        torch.cuda.synchronize()
        o = out.float().cpu().reshape(t, hq, d)
        f = fp32.reshape(t, hq, d)
        err = (o - f).abs()
        bad = ~(err <= 2e-2 + 1e-2 * f.abs())
        print(f"[{name}] {label:22s} max_err {err.max().item():.4e} bad {bad.float().mean().item():.4f} nan {torch.isnan(o).float().mean().item():.4f}", flush=True)
        if bad.any():
            print("    bad fraction by token  :", [round(x, 2) for x in bad.float().mean(dim=(1, 2)).tolist()[:24]])
            print("    bad fraction by head   :", [round(x, 2) for x in bad.float().mean(dim=(0, 2)).tolist()[:28]])
            print("    bad fraction by dim/16 :", [round(x, 2) for x in bad.float().mean(dim=(0, 1)).reshape(8, 16).mean(dim=1).tolist()])
            print("    out[0,0,:8] ", [round(x, 3) for x in o[0, 0, :8].tolist()])
            print("    ref[0,0,:8] ", [round(x, 3) for x in f[0, 0, :8].tolist()])
            print("    out[-1,-1,64:72]", [round(x, 3) for x in o[-1, -1, 64:72].tolist()])
            print("    ref[-1,-1,64:72]", [round(x, 3) for x in f[-1, -1, 64:72].tolist()])
            ratio = (o / f)[~torch.isnan(o)]
            if ratio.numel():
                print("    median out/ref ratio   :", round(ratio.median().item(), 4))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_group(sys.argv[1])
    else:
        for g in GROUPS:
            for env_extra in ({},):
                tag = g + ("+serialize" if env_extra else "")
                print(f"===== {tag}", flush=True)
                try:
                    r = subprocess.run([sys.executable, __file__, g], env={**os.environ, **env_extra}, timeout=240, capture_output=True, text=True)
                    print(r.stdout[-6000:])
                    if r.returncode != 0:
                        print(f"[{tag}] exit code {r.returncode}\n{r.stderr[-3000:]}")
                except subprocess.TimeoutExpired as e:
                    print(f"[{tag}] TIMEOUT\n{(e.stdout or b'')[-3000:]}")
