"""Dev/measurement tool (GPU box): every BASELINE config through each kernel path, device-timed, with roofline numbers.
Writes one JSON object per line to stdout (and gpurun_out/configs.jsonl).  Not the driver's bench (that is bench.py).

    python tools/bench_configs.py [--only cfg2,cfg3d,...] [--flashinfer]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd  # noqa: E402
from hydrainfer_b200.workloads import make_batch  # noqa: E402

DEV = "cuda:0"
HBM_PEAK, TC_PEAK = 6537.0, 1629.8
try:
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    HBM_PEAK, TC_PEAK = float(peaks["hbm_gbs"]), float(peaks["bf16_tflops"])
except Exception:
    pass


def algo_bytes(seq_lens, hq, hkv, d):
    kv = sum(2 * L * hkv * d * 2 for _, L in seq_lens)
    qo = sum(2 * q * hq * d * 2 for q, _ in seq_lens)
    return kv + qo


def algo_flops(seq_lens, hq, d):
    return sum(4 * hq * d * (q * (L - q) + q * (q + 1) / 2) for q, L in seq_lens)


def time_call(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2], ts[0]


def time_graph(fn, calls=10, replays=7):
    """GPU time per call with the host out of the picture: `calls` invocations captured into one CUDA graph, replayed
    `replays` times, median.  For launches shorter than the Python + launch cost of a call (small decode batches) the
    per-call event timing above measures the host, not the kernels."""
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(calls):
                fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(replays):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        graph.replay()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / calls)
    return sorted(ts)[len(ts) // 2]


USE_GRAPH = False


def run_case(name, seq_lens, hq, hkv, paths, out_lines, flashinfer_cmp=False, dtype=torch.bfloat16):
    d, bs = int(os.environ.get("HEAD_DIM", "128")), 16  # HEAD_DIM=64: the variable-head-dim mode of the pair kernel
    batch = make_batch(seq_lens, hq, hkv, d, bs, dtype=dtype, device=DEV, gen_device=DEV, seed=0)
    t = batch.n_tokens
    q3 = batch.query.view(t, hq, d)
    out = torch.empty_like(q3)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    q_cu, kv_cu, bt, cu_b = i32(batch.q_cu_seq_lens), i32(batch.kv_cu_seq_lens), i32(batch.block_tables), i32(batch.cu_blocks_lens)
    nbytes, flops = algo_bytes(seq_lens, hq, hkv, d), algo_flops(seq_lens, hq, d)
    results = {}
    from hydrainfer_b200 import _lib
    tt = int(_lib.lib.hi_attention_tile_tokens(hq, hkv))
    items = sorted(((L - q + min(q, (tile + 1) * tt), b, tile) for b, (q, L) in enumerate(seq_lens) for tile in range((q + tt - 1) // tt)), reverse=True)
    plan = i32([x for _, b, tile in items for x in (b, tile)]).view(-1, 2)
    work = sum(cost for cost, _, _ in items)
    paths = list(paths) + [(p + "+plan", c) for p, c in paths if p == "pair"]
    for pname, path in paths:
        hints = (plan, tt, work) if pname.endswith("+plan") else (None, 0, 0)
        def fn():
            mha_varlen_fwd(out, q3, batch.key_cache, batch.value_cache, q_cu, kv_cu, bt, cu_b, None, batch.q_max, batch.kv_max,
                           1 / math.sqrt(d), 0, -1, 0, 0, path, *hints)
        try:
            med, best = time_call(fn)
        except RuntimeError as e:
            results[pname] = {"error": str(e)[:200]}
            continue
        if USE_GRAPH:
            try:
                med = best = time_graph(fn)
            except RuntimeError as e:
                results[pname] = {"error": "graph: " + str(e)[:200]}
                continue
        results[pname] = {"ms": med, "ms_best": best, "GBs": nbytes / med / 1e6, "hbm_frac": nbytes / med / 1e6 / HBM_PEAK,
                          "TFLOPs": flops / med / 1e9, "tc_frac": flops / med / 1e9 / TC_PEAK, "tokens_per_s": t / med * 1e3,
                          "timing": "cuda graph of 10 calls" if USE_GRAPH else "events around each call"}
        results[pname + "_out"] = out.float().abs().mean().item()
    if flashinfer_cmp:
        try:
            results["flashinfer"] = flashinfer_case(batch, q3, seq_lens, hq, hkv, d, bs, nbytes, flops)
        except Exception as e:  # comparator only
            results["flashinfer"] = {"error": repr(e)[:300]}
    line = {"case": name, "n_seqs": len(seq_lens), "tokens": t, "heads": [hq, hkv], "algo_bytes": nbytes, "algo_flops": flops, **results}
    print(json.dumps(line), flush=True)
    out_lines.append(line)
    del batch
    torch.cuda.empty_cache()


def flashinfer_case(batch, q3, seq_lens, hq, hkv, d, bs, nbytes, flops):
    import flashinfer
    ws = torch.empty(128 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    last = i32([(L + bs - 1) % bs + 1 for _, L in seq_lens])
    decode = all(q == 1 for q, _ in seq_lens)
    t0 = time.time()
    if decode:
        w = flashinfer.BatchDecodeWithPagedKVCacheWrapper(ws, "NHD", use_tensor_cores=True)
        w.plan(i32(batch.cu_blocks_lens), i32(batch.block_tables), last, hq, hkv, d, bs, q_data_type=torch.bfloat16, kv_data_type=torch.bfloat16)
    else:
        w = flashinfer.BatchPrefillWithPagedKVCacheWrapper(ws, "NHD")
        w.plan(i32(batch.q_cu_seq_lens), i32(batch.cu_blocks_lens), i32(batch.block_tables), last, hq, hkv, d, bs, causal=True,
               q_data_type=torch.bfloat16, kv_data_type=torch.bfloat16)
    fn = lambda: w.run(q3, (batch.key_cache, batch.value_cache))
    fn()
    torch.cuda.synchronize()
    jit_s = time.time() - t0
    med, best = time_call(fn)
    return {"ms": med, "ms_best": best, "GBs": nbytes / med / 1e6, "hbm_frac": nbytes / med / 1e6 / HBM_PEAK, "TFLOPs": flops / med / 1e9,
            "tc_frac": flops / med / 1e9 / TC_PEAK, "first_call_s": jit_s, "version": flashinfer.__version__}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--flashinfer", action="store_true")
    ap.add_argument("--graph", action="store_true", help="time our paths as a CUDA graph of 10 calls (GPU time without host launch cost)")
    args = ap.parse_args()
    global USE_GRAPH
    USE_GRAPH = args.graph
    only = set(args.only.split(",")) if args.only else None
    SIMT, TC, AUTO, DEC, PAIR = ("simt", 1), ("tc", 2), ("auto", 0), ("dec", 3), ("pair", 4)
    g = torch.Generator().manual_seed(0)
    cfg3_dec = [(1, int(L)) for L in torch.randint(256, 8193, (48,), generator=g).tolist()]
    cfg3_pre = [(512, 512), (512, 2048), (512, 4096), (512, 8192)]
    cases = [
        ("cfg2", [(1, 2048)] * 64, 32, 32, [SIMT, TC, DEC, AUTO]),
        ("cfg2_b8", [(1, 2048)] * 8, 32, 32, [SIMT, TC]),
        ("cfg2_b1", [(1, 2048)] * 1, 32, 32, [SIMT]),
        ("cfg2_b4", [(1, 2048)] * 4, 32, 32, [SIMT]),
        ("cfg2_b16", [(1, 2048)] * 16, 32, 32, [SIMT]),
        ("cfg2_b32", [(1, 2048)] * 32, 32, 32, [SIMT]),
        ("cfg2_ctx8k", [(1, 8192)] * 16, 32, 32, [SIMT, TC]),
        ("cfg3d", cfg3_dec, 28, 4, [SIMT, TC, DEC]),
        ("gqa_b4_8k", [(1, 8192)] * 4, 28, 4, [DEC]),
        ("gqa_b8_2k", [(1, 2048)] * 8, 28, 4, [DEC]),
        ("gqa_b16_rag", [(1, 300 + 500 * i) for i in range(16)], 28, 4, [DEC]),
        ("gqa72_b64_rag", [(1, 256 + 120 * i) for i in range(64)], 64, 8, [DEC]),
        ("cfg3p", cfg3_pre, 28, 4, [TC, PAIR]),
        ("cfg3mix", cfg3_dec + cfg3_pre, 28, 4, [AUTO, TC, PAIR]),
        ("pre256", [(256, 256)] * 32, 28, 4, [TC, PAIR]),
        ("pre1k", [(1024, 1024)] * 8, 28, 4, [TC, PAIR]),
        ("pre4k", [(4096, 4096)] * 2, 28, 4, [TC, PAIR]),
        ("pre8k", [(8192, 8192)] * 1, 28, 4, [TC, PAIR]),
        ("pre_mha2k", [(2048, 2048)] * 4, 32, 32, [TC, PAIR]),
        ("pre_mha8k", [(8192, 8192)] * 1, 32, 32, [TC, PAIR]),
        ("cfg4_2k", [(1, 2048)] * 256, 64, 8, [SIMT, TC, DEC]),
        ("cfg4_4k", [(1, 4096)] * 256, 64, 8, [SIMT, TC, DEC]),
        ("cfg4_shard8", [(1, 4096)] * 32, 64, 8, [SIMT, TC, DEC]),
        ("cfg4_shard8_2k", [(1, 2048)] * 32, 64, 8, [DEC]),
        ("cfg4_shard4_2k", [(1, 2048)] * 64, 64, 8, [DEC]),
        ("cfg4_shard2_2k", [(1, 2048)] * 128, 64, 8, [DEC]),
    ]
    lines = []
    for name, seq_lens, hq, hkv, paths in cases:
        if only and name not in only:
            continue
        run_case(name, seq_lens, hq, hkv, paths, lines, flashinfer_cmp=args.flashinfer)
    outdir = ROOT / "gpurun_out"
    outdir.mkdir(exist_ok=True)
    with open(outdir / "configs.jsonl", "a") as f:
        for line in lines:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
