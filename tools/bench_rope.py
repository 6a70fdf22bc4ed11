"""Dev/measurement tool (GPU box): the step in front of the attention kernels — rotary embedding + KV append as
the reference issues it (two launches: apply_rotary_pos_emb in place, then set_kv_cache) against the fused hi_rope_append
launch, device-timed, with the HBM roofline.  One JSON object per line (stdout + gpurun_out/rope.jsonl).

Algorithmic bytes per token (fused): read q, k, v once, write q and the two cache rows = itemsize*d*(2*Hq + 4*Hkv)
(+ cos/sin row and ids, ignored).  The two-launch form moves itemsize*d*(2*Hq + 2*Hkv) (rotary, in place) +
itemsize*d*4*Hkv (append) = itemsize*d*(2*Hq + 6*Hkv).
Inputs rotate over enough copies to exceed the 126 MB L2."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from hydrainfer_b200._C.kernel.kv_cache_kernels import set_kv_cache  # noqa: E402
from hydrainfer_b200._C.kernel.position_embedding import apply_rotary_pos_emb, rope_set_kv_cache  # noqa: E402
from hydrainfer_b200.layer.rotary_embedding import B200RotaryEmbeddingHandler  # noqa: E402

DEV = "cuda:0"
HBM_PEAK = 6537.0
try:
    HBM_PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass

CASES = [  # name, tokens, Hq, Hkv
    ("cfg2_decode_llava7b", 64, 32, 32),
    ("cfg3_mixed_qwen7b", 2096, 28, 4),
    ("cfg4_decode_qwen72b", 256, 64, 8),
    ("prefill8k_qwen7b", 8192, 28, 4),
    ("prefill8k_llava7b", 8192, 32, 32),
    ("prefill32k_qwen72b", 32768, 64, 8),
]


def time_rotating(fns, iters=30, warmup=3):
    n = len(fns)
    for i in range(warmup * n):
        fns[i % n]()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (s, e) in enumerate(evs):
        s.record()
        fns[i % n]()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2]


def main() -> None:
    d, bs, dtype, max_pos = 128, 16, torch.bfloat16, 32768
    inv_freq = 1. / torch.pow(torch.tensor(1e6), torch.arange(0, d, 2, dtype=torch.float) / d)
    table = B200RotaryEmbeddingHandler(d, max_pos, inv_freq, False).cos_sin_cache.to(dtype).to(DEV)
    out_path = ROOT / "gpurun_out" / "rope.jsonl"
    out_path.parent.mkdir(exist_ok=True)
    lines = []
    for name, t, hq, hkv in CASES:
        row_bytes = (hq + 2 * hkv) * d * 2
        copies = max(3, int(300e6 // (t * row_bytes)) + 1)
        copies = min(copies, 64)
        n_blocks = (t + bs - 1) // bs + 8
        g = torch.Generator(device=DEV).manual_seed(0)
        sets = []
        for _ in range(copies):
            qkv = torch.randn(t, (hq + 2 * hkv) * d, generator=g, device=DEV, dtype=torch.float32).to(dtype)
            kc = torch.zeros(n_blocks, bs, hkv, d, dtype=dtype, device=DEV)
            vc = torch.zeros_like(kc)
            slots = torch.randperm(n_blocks * bs, generator=g, device=DEV)[:t].to(torch.int32)
            pos = torch.randint(0, max_pos, (t,), generator=g, device=DEV, dtype=torch.int32)
            q = qkv[:, :hq * d].view(t, hq, d)
            k = qkv[:, hq * d:(hq + hkv) * d].view(t, hkv, d)
            v = qkv[:, (hq + hkv) * d:].view(t, hkv, d)
            sets.append((q, k, v, kc, vc, slots, pos))

        def two_launch(s):
            q, k, v, kc, vc, slots, pos = s
            apply_rotary_pos_emb(q, k, pos, table, d, False)
            set_kv_cache(slots, k, v, kc, vc)

        def fused(s):
            q, k, v, kc, vc, slots, pos = s
            rope_set_kv_cache(q, k, v, pos, table, d, False, slots, kc, vc)

        def rotary_only(s):
            q, k, v, kc, vc, slots, pos = s
            apply_rotary_pos_emb(q, k, pos, table, d, False)

        ms_two = time_rotating([lambda s=s: two_launch(s) for s in sets])
        ms_fused = time_rotating([lambda s=s: fused(s) for s in sets])
        ms_rot = time_rotating([lambda s=s: rotary_only(s) for s in sets])
        bytes_fused = t * 2 * d * (2 * hq + 4 * hkv)
        bytes_two = t * 2 * d * (2 * hq + 6 * hkv)
        line = {"case": name, "tokens": t, "heads": f"{hq}/{hkv}", "copies_rotated": copies,
                "two_launch_ms": round(ms_two, 5), "fused_ms": round(ms_fused, 5), "rotary_only_ms": round(ms_rot, 5),
                "speedup_fused_vs_two_launch": round(ms_two / ms_fused, 3),
                "fused_gbs": round(bytes_fused / ms_fused / 1e6, 1), "fused_frac_hbm": round(bytes_fused / ms_fused / 1e6 / HBM_PEAK, 3),
                "two_launch_gbs": round(bytes_two / ms_two / 1e6, 1), "algorithmic_bytes_fused": bytes_fused}
        print(json.dumps(line), flush=True)
        lines.append(json.dumps(line))
        del sets
        torch.cuda.empty_cache()
    out_path.write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
