"""Dev tool (GPU box): un-paged vision attention (hi_varlen_attention through the layer modules) vs the installed flash_attn
(the reference's second handler, multihead_attention.py:76-112, 214-233), device-timed.  Prints JSON lines.

    python tools/bench_vision.py [--flash]
"""
import json
import math
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd  # noqa: E402

dev = torch.device("cuda:0")
PEAK_TF = 1632.4
try:
    PEAK_TF = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops"])
except Exception:
    pass

# (name, sequence lengths, heads, head_dim)
CASES = [
    ("llava_clip_b16", [577] * 16, 16, 64),     # LLaVA-1.5 CLIP ViT-L/14-336: 577 tokens per image
    ("llava_clip_b64", [577] * 64, 16, 64),
    ("siglip_b16", [729] * 16, 16, 72),         # SigLIP-so400m: 729 patches, 16 heads of 72
    ("qwen2vl_mixed", [4096, 64, 1024, 300, 2500, 16, 784, 1600], 16, 80),   # Qwen2-VL ViT: packed images, 16 heads of 80
    ("qwen2vl_4x4096", [4096] * 4, 16, 80),
    ("d128_8x2048", [2048] * 8, 16, 128),
]


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts)


def main():
    use_flash = "--flash" in sys.argv
    fa = None
    if use_flash:
        try:
            import flash_attn
            fa = flash_attn
        except Exception as e:  # pragma: no cover
            print(json.dumps({"flash_attn": f"unavailable: {e}"}))
    for name, lens, heads, d in CASES:
        total = sum(lens)
        g = torch.Generator(device=dev).manual_seed(0)
        q, k, v = (torch.randn(total, heads, d, generator=g, device=dev).to(torch.bfloat16) for _ in range(3))
        cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
        out = torch.empty_like(q)
        mx = max(lens)
        flops = 4 * heads * d * sum(n * n for n in lens)
        scale = 1.0 / math.sqrt(d)

        def ours():
            mha_varlen_fwd(out, q, k, v, cu, cu, None, None, None, mx, mx, scale, 0, -1, -1, 0)

        ms = timeit(ours)
        row = {"case": name, "tokens": total, "heads": heads, "head_dim": d, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
               "frac_bf16_peak": round(flops / ms / 1e9 / PEAK_TF, 3)}
        if fa is not None:
            def theirs():
                fa.flash_attn_varlen_func(q, k, v, cu, cu, mx, mx, causal=False)
            try:
                ref = fa.flash_attn_varlen_func(q, k, v, cu, cu, mx, mx, causal=False)
                ours()
                torch.cuda.synchronize()
                row["max_abs_diff_vs_flash_attn"] = float((ref.float() - out.float()).abs().max())
                ms_fa = timeit(theirs)
                row["flash_attn_ms"] = round(ms_fa, 4)
                row["flash_attn_tflops"] = round(flops / ms_fa / 1e9, 1)
                row["speedup_vs_flash_attn"] = round(ms_fa / ms, 2)
            except Exception as e:  # pragma: no cover
                row["flash_attn"] = f"failed: {type(e).__name__}: {e}"[:200]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
