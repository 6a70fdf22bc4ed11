#!/usr/bin/env python
"""bench.py — the paged-KV attention hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): LLaVA-1.5-7B LM shape — 32 q / 32 kv heads, head_dim 128, block 16, bf16 — decode
batch 64 at context 2048, one attention-layer call per step: KV append of the 64 new tokens (set_kv_cache) followed by
paged decode attention over the 2048 cached tokens of every sequence.  Synthetic N(0,1) data, seeded random block tables.
Multi-GPU: sequences are independent, every rank runs its own batch of 64 on its own pool (weak scaling, no data-path
collective); value = tokens of all ranks / max-over-ranks device time.

Prints ONE JSON line (rank 0): metric decode-attn tokens/s (+ roofline on the split-KV kernel, cpu_baseline, e2e,
clocks, migration GB/s as an extra).  `--impl reference` times the CPU restatement of the reference's torch path
(oracle/, kind "port") on the box's host cores instead.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# ---- workload: BASELINE.json configs[1] ----------------------------------------------------------------------------------
HQ, HKV, D, BS = 32, 32, 128, 16
BATCH, CTX = 64, 2048
DTYPE = torch.bfloat16
ALGO_BYTES_PER_TOKEN = 2 * CTX * HKV * D * 2 + 2 * HQ * D * 2  # K+V read once + q read + o written = 33 570 816
WORKLOAD = "LLaVA-1.5-7B decode attention layer call: 32q/32kv heads d=128 block16 bf16, batch 64, ctx 2048 (576 image + text), KV append + paged attention"


def load_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
            return
        # nvidia-smi needs a moment to attach; wait for its first line so the timed region is covered
        t0 = time.time()
        while time.time() - t0 < 3.0 and os.path.getsize(self.file.name) == 0:
            time.sleep(0.05)

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        rows = [r.strip().split(",") for r in Path(self.file.name).read_text().splitlines() if r.strip()]
        os.unlink(self.file.name)
        clocks, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                clocks.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.strip().lower() == "active":
                    reasons.add(name)
        if clocks:
            out["sm_mhz"] = statistics.median(clocks)
        out["samples"] = len(clocks)
        out["reasons"] = sorted(reasons)
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---- the reference arm / cpu baseline: oracle port on the host cores ------------------------------------------------------
def cpu_reference_run(n_seqs: int, steps: int, warmup: int, budget_s: float = 1e9) -> dict:
    """Times the reference's CPU path restated in oracle/ (TorchCausalGroupedQueryPageAttentionHandler + Python
    set_kv_cache, hydrainfer/layer/causal_attention.py:307-374, 394-406) on `n_seqs` sequences of the workload."""
    from hydrainfer_b200.workloads import make_batch
    from oracle import paged_kv_oracle as oracle
    torch.set_num_threads(os.cpu_count() or 1)
    batch = make_batch([(1, CTX)] * n_seqs, HQ, HKV, D, BS, dtype=DTYPE, device="cpu", seed=0)
    slots = torch.tensor(batch.new_cache_slots, dtype=torch.int32)
    tables = torch.tensor(batch.block_tables, dtype=torch.int32)
    times = []
    t_begin = time.perf_counter()
    with torch.inference_mode():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle.attention_layer_forward(batch.query, batch.key, batch.value, batch.key_cache, batch.value_cache, slots,
                                           batch.q_cu_seq_lens, batch.kv_cu_seq_lens, tables, batch.cu_blocks_lens, HQ, HKV, D)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_begin > budget_s and len(times) >= 3:  # slow hosts: keep the run within minutes
                break
    steps = len(times)
    med = statistics.median(times)
    return {"value": n_seqs / med, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_seqs} of the {BATCH} sequences of the workload per step (same ctx {CTX}, heads, dtype), median of {steps} steps, {warmup} warm-up",
            "ms_per_step": med * 1e3, "cpu_count": os.cpu_count(), "steps": steps}


_RESULT_FD = 1


def emit_result(line: dict) -> None:
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


def run_reference_arm(args) -> None:
    rank, world, _ = dist_env()
    if rank != 0:
        return
    warmup = args.warmup
    n_seqs = BATCH  # the whole step: ~0.7 s of CPU work on the GPU box's 16 host threads
    # K steps as asked, unless the host is so slow that the run would not end within a few minutes (then as many as fit)
    res = cpu_reference_run(n_seqs, max(1, args.steps), warmup, budget_s=150.0)
    steps = res["steps"]
    line = {
        "impl": "reference", "metric": "decode-attn tokens/s", "value": res["value"], "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "inputs_larger_than_l2", "step": f"{n_seqs} sequences per step (the whole workload step)"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_result(line)


# ---- our arm ---------------------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from hydrainfer_b200 import _lib
    from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
    from hydrainfer_b200.memory import KVCache
    from hydrainfer_b200.workloads import make_batch
    from hydrainfer_b200._C.kernel.flash_attn import last_launch_count

    # every rank owns a different batch (seed = rank): weak scaling over independent sequences
    batch = make_batch([(1, CTX)] * BATCH, HQ, HKV, D, BS, dtype=DTYPE, device=dev, gen_device=dev, seed=rank)
    kv_cache = KVCache(batch.key_cache, batch.value_cache)
    layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(HQ, HKV, D))

    def build_params():
        builder = AttentionParametersBuilder(HQ, HKV, D, BS, dev)
        for req in batch.requests():
            builder.add_request(*req)
        builder.add_kv_cache(kv_cache)
        return builder.build_attention_parameters()[0]

    requests = batch.requests()
    params = build_params()
    stream = torch.cuda.current_stream(dev)

    def step():
        return layer(batch.query, batch.key, batch.value, params).o

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- kernel-only leg: inputs resident in HBM, device timed ----------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    n_events = args.steps
    evs = []
    for _ in range(2 * n_events):
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.hi_event_create(ctypes.byref(h)))
        evs.append(h)
    launches = 0
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for i in range(args.steps):
        # the append launch happens first; arm the pair so the attention call records around its split-KV kernel
        _lib.lib.hi_set_kernel_timing_events(evs[2 * i], evs[2 * i + 1])
        step()
        launches += 1 + last_launch_count()  # set_kv_cache (1 launch) + attention kernels of the last call
    t_stop.record(stream)
    barrier()
    ms_total = t_start.elapsed_time(t_stop)
    kernel_ms = []
    for i in range(args.steps):
        ms = ctypes.c_float()
        _lib.check(_lib.lib.hi_event_elapsed_ms(evs[2 * i], evs[2 * i + 1], ctypes.byref(ms)))
        kernel_ms.append(ms.value)
    for h in evs:
        _lib.lib.hi_event_destroy(h)

    # ---- end-to-end leg: host buffers in, host result out, through the layer API ---------------------------------------------
    q_host = batch.query.cpu().pin_memory()
    k_host = batch.key.cpu().pin_memory()
    v_host = batch.value.cpu().pin_memory()
    o_host = torch.empty((BATCH, HQ * D), dtype=DTYPE).pin_memory()

    o_hosts = [o_host, torch.empty_like(o_host).pin_memory()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    # double buffering for the pipelined leg: device-side q/k/v per parity, one stream for H2D and one for D2H, so the copies
    # of neighbouring steps run beside the kernels instead of in front of / behind them on the compute stream
    copy_in, copy_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    qkv_dev = [tuple(torch.empty_like(t, device=dev) for t in (q_host, k_host, v_host)) for _ in range(2)]
    in_ready = [torch.cuda.Event(), torch.cuda.Event()]
    computed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_enqueue(i: int) -> None:
        """One whole step through the public layer API: pinned host q/k/v -> HBM, per-step metadata upload (one pinned int32
        buffer), KV append + attention, result -> pinned host buffer i % 2, completion event."""
        q = q_host.to(dev, non_blocking=True)
        k = k_host.to(dev, non_blocking=True)
        v = v_host.to(dev, non_blocking=True)
        p = build_params()
        o = layer(q, k, v, p).o
        o_hosts[i % 2].copy_(o, non_blocking=True)
        done[i % 2].record(stream)

    def e2e_enqueue_overlapped(i: int) -> None:
        """The same step with its copies on the copy streams.  Buffers of parity i % 2 are free: the caller has read the result
        of step i - 2 (done[i % 2] synchronized) before it enqueues step i."""
        b = i % 2
        q, k, v = qkv_dev[b]
        with torch.cuda.stream(copy_in):
            q.copy_(q_host, non_blocking=True)
            k.copy_(k_host, non_blocking=True)
            v.copy_(v_host, non_blocking=True)
            in_ready[b].record(copy_in)
        p = build_params()                      # metadata: one pinned buffer, uploaded on the compute stream
        stream.wait_event(in_ready[b])
        o = layer(q, k, v, p).o
        computed[b].record(stream)
        o.record_stream(copy_out)
        with torch.cuda.stream(copy_out):
            copy_out.wait_event(computed[b])
            o_hosts[b].copy_(o, non_blocking=True)
            done[b].record(copy_out)

    def e2e_serial_step():
        e2e_enqueue(0)
        done[0].synchronize()  # the caller reads the result before it prepares the next step

    def e2e_pipelined(n: int) -> None:
        # One step in flight: the host builds and enqueues step i + 1 while the GPU runs step i, then reads step i's result
        # (what an engine serving more than one micro-batch / layer stream does).  Every step still does all its copies.
        e2e_enqueue_overlapped(0)
        for i in range(1, n):
            e2e_enqueue_overlapped(i)
            done[(i - 1) % 2].synchronize()
        done[(n - 1) % 2].synchronize()

    for _ in range(max(3, args.warmup // 2)):
        e2e_serial_step()
    barrier()
    e2e_steps = max(5, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_serial_step()
    torch.cuda.synchronize(dev)
    e2e_serial_s = time.perf_counter() - t0
    e2e_pipelined(4)
    barrier()
    t0 = time.perf_counter()
    e2e_pipelined(e2e_steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    # the host buffers of both parities hold what the resident-input path computes (the kernels are deterministic)
    o_ref = step().cpu()
    e2e_checked = bool(torch.equal(o_hosts[0], o_ref) and torch.equal(o_hosts[1], o_ref))
    if not e2e_checked:
        raise SystemExit("bench.py: the end-to-end leg's host result differs from the resident-input result")
    h2d = q_host.numel() * 2 + k_host.numel() * 2 + v_host.numel() * 2 + 4 * (
        len(batch.q_cu_seq_lens) + len(batch.kv_cu_seq_lens) + BATCH + len(batch.new_cache_slots) + len(batch.block_tables) + len(batch.cu_blocks_lens))
    d2h = o_host.numel() * 2
    # the timed region is ~10 ms; keep the identical step running for ~1.5 s so nvidia-smi (100 ms period) sees the clocks
    # the kernels actually run at (same kernels, same inputs; not part of any reported time)
    t_end = time.time() + 1.5
    while time.time() < t_end:
        for _ in range(50):
            step()
        torch.cuda.synchronize(dev)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- migration extra: KV-migrate GB/s (same metric family, BASELINE.json) -----------------------------------------------
    migrate = measure_migration(rank, world, local, dev)

    # ---- aggregate over ranks: MAX time, SUM tokens ---------------------------------------------------------------------------
    stats = torch.tensor([ms_total, e2e_s, statistics.mean(kernel_ms), e2e_serial_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_total_max, e2e_s_max, kernel_ms_max, e2e_serial_s_max = (float(x) for x in stats.tolist())
    tokens_per_step = BATCH * world
    value = tokens_per_step * args.steps / (ms_total_max * 1e-3)
    e2e_value = tokens_per_step * e2e_steps / e2e_s_max

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = BATCH * ALGO_BYTES_PER_TOKEN / (kernel_ms_max * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("paged_attn_stream_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        # the reference's CPU path on this box's host cores: rank 0, N = 1 only (at N > 1 the other ranks' processes share the cores)
        cpu = cpu_reference_run(BATCH, 10, 3, budget_s=45.0) if world == 1 else None
        line = {
            "metric": "decode-attn tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "tokens_per_step_per_gpu": BATCH, "l2": "inputs_larger_than_l2 (2.1 GB of KV per step vs 126 MB L2)",
                       "parallelism": f"sequences sharded, {world} independent rank(s), no collective", "kernel": "scatter_rows_kernel + paged_attn_stream_kernel<bf16,128,1> (cp.async split-KV) + merge_partials_kernel"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0, "kernel_ms": kernel_ms_max,
                         "algorithmic_bytes_per_launch": BATCH * ALGO_BYTES_PER_TOKEN},
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu is not None else None,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": e2e_s_max / e2e_steps * 1e3, "mode": "one step in flight, double-buffered: step i+1 is built and enqueued while step i runs, its H2D copies and step i's D2H run on copy streams beside the kernels; every step does all its copies and every result is read on the host",
                    "serial_ms_per_step": e2e_serial_s_max / e2e_steps * 1e3, "serial_value": tokens_per_step * e2e_steps / e2e_serial_s_max,
                    "result_equals_resident_path": e2e_checked},
            "gpu_launches": launches,
            "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": clocks.get("reasons", []), "samples": clocks.get("samples", 0)},
            "migrate": migrate,
        }
        emit_result(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_migration(rank: int, world: int, local: int, dev: torch.device) -> dict:
    """block_migration GB/s.  N=1: pool -> pool on the same GPU (HBM bound).  N>1: disjoint pairs (2i -> 2i+1) pull
    through CUDA-IPC peer mappings over NVLink, all pairs at once; payload bytes one direction / max device time."""
    import torch.distributed as dist
    from hydrainfer_b200._C.data_transfer import block_migration as bm
    geom = dict(n_layers=32, n_tokens=2, block_size=16, n_heads=32, head_size=128)  # LLaVA-7B pool: 8 MiB per block
    pool_blocks, n_move = 320, 256  # 2 GiB moved per request
    bytes_per_block = geom["n_layers"] * geom["n_tokens"] * geom["block_size"] * geom["n_heads"] * geom["head_size"] * 2
    pool = torch.empty((geom["n_layers"], geom["n_tokens"], pool_blocks, geom["block_size"], geom["n_heads"], geom["head_size"]), dtype=DTYPE, device=dev)
    pool.normal_()
    g = torch.Generator().manual_seed(100 + rank)
    src_bt = torch.randperm(pool_blocks, generator=g)[:n_move].tolist()
    dst_bt = torch.randperm(pool_blocks, generator=g)[:n_move].tolist()
    handle = bm.get_ipc_mem_handle(pool)
    pattern = "same-GPU pool->pool (HBM)"
    is_receiver = True
    src_handle = handle
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        pattern = f"{world // 2} disjoint pair(s) 2i->2i+1 over NVLink P2P (CUDA IPC pull)"
        is_receiver = (rank % 2 == 1)
        src_handle = handles[rank - 1] if is_receiver else handle
        dst_pool = pool
    if world == 1:
        dst_pool = torch.empty_like(pool)
    times = []
    reps = 5
    for i in range(reps + 2):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if is_receiver:
            bm.migrate_blocks(src_bt, dst_bt, src_handle, dst_pool, pool_blocks)
        e.record()
        torch.cuda.synchronize(dev)
        if i >= 2:
            times.append(s.elapsed_time(e))
    ok = True
    if is_receiver and world == 1:
        ok = bool(torch.equal(dst_pool[:, :, dst_bt[:4]], pool[:, :, src_bt[:4]]))
    t = torch.tensor([statistics.median(times) if is_receiver else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    pairs = max(1, world // 2)
    payload = n_move * bytes_per_block
    del pool
    return {"gbs_per_pair": payload / (ms * 1e-3) / 1e9, "gbs_aggregate": pairs * payload / (ms * 1e-3) / 1e9, "ms": ms, "pattern": pattern,
            "blocks_per_request": n_move, "bytes_per_request": payload, "bit_exact_spot_check": ok,
            "nvlink_peer_copy_reference_gbs": 770.0 if world > 1 else None}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
    # NCCL_DEBUG is set on the box), so file descriptor 1 is pointed at stderr for the run and the line goes to the real stdout.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
