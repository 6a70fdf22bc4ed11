"""Turns gpurun_out/configs.jsonl (tools/bench_configs.py) into the markdown table kept under profiles/.

    python tools/summarize_configs.py gpurun_out/configs.jsonl > profiles/rNN_configs.md
"""
import json
import sys


def main(path: str) -> None:
    print("| case | heads (q/kv) | new tokens | path | ms (median of 20) | GB/s | % HBM (measured copy) | TFLOP/s | % bf16 peak (measured) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for raw in open(path):
        d = json.loads(raw)
        for name, v in d.items():
            if not isinstance(v, dict):
                continue
            if "ms" not in v:
                print(f"| {d['case']} | {d['heads'][0]}/{d['heads'][1]} | {d['tokens']} | {name} | error: {v.get('error', '')[:60]} | | | | |")
                continue
            print(f"| {d['case']} | {d['heads'][0]}/{d['heads'][1]} | {d['tokens']} | {name} | {v['ms']:.4f} | {v['GBs']:.0f} | {100 * v['hbm_frac']:.1f} | "
                  f"{v['TFLOPs']:.0f} | {100 * v['tc_frac']:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
