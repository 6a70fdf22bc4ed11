"""Measurement tool (GPU box): bench.py's `prefill` extra alone (pre1k / pre8k / cfg3p / cfg3mix through the layer API, tcgen05
kernel timed by its own events, L2 flushed between calls) - the numbers the driver records, for A/B runs under tuning
environment variables (HI_PAIR_RING, HI_PAIR_TMA_STORE, ...)."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

if __name__ == "__main__":
    res = bench.measure_prefill(torch.device("cuda:0"), reps=int(os.environ.get("REPS", "15")))
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("HI_"))
    row = {k: (round(v["kernel_ms"], 4), round(v["frac"], 3), round(v["ms"], 4)) for k, v in res.items() if isinstance(v, dict)}
    print(tag or "default", json.dumps(row))
