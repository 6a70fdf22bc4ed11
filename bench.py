#!/usr/bin/env python
"""bench.py — the paged-KV attention hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): LLaVA-1.5-7B LM shape — 32 q / 32 kv heads, head_dim 128, block 16, bf16 — decode
batch 64 at context 2048, one attention-layer call per step: KV append of the 64 new tokens (set_kv_cache) followed by
paged decode attention over the 2048 cached tokens of every sequence.  Synthetic N(0,1) data, seeded random block tables.
Multi-GPU: sequences are independent, every rank runs its own batch of 64 on its own pool (weak scaling, no data-path
collective); value = tokens of all ranks / max-over-ranks device time.

Prints ONE JSON line (rank 0): metric decode-attn tokens/s (+ roofline on the decode attention kernel, cpu_baseline, e2e,
clocks) and, as extras on the same line, the other halves of BASELINE.json's metric and configs:
  "prefill"        N = 1: BASELINE config 3 (Qwen2-VL-7B 28q/4kv) — pre1k, pre8k, cfg3p (chunked prefill), cfg3mix (48 decode rows +
                   4 chunked prefills) through the layer API; ms, TFLOP/s from 4*Hq*d*sum[q(L-q)+q(q+1)/2], fraction of the measured
                   bf16 peak (burst and sustained), the tcgen05 kernel alone via hi_set_kernel_timing_events
  "cfg4"           BASELINE config 4 (Qwen2-VL-72B 64q/8kv decode, batch 256, ctx 2048 / 4096): batch 256 at N = 1, 256/N sequences per
                   rank at N > 1 (STRONG scaling, sequence i -> rank i mod N), aggregate tokens/s and per-GPU HBM fraction
  "migrate_sweep"  BASELINE config 5: {16, 256, 4096} blocks per request x {LLaVA-7B, Qwen2-VL-7B} pools; same GPU at N = 1, disjoint
                   NVLink pairs at N > 1, next to a cudaMemcpyPeerAsync of the same bytes timed in the same run; every moved block
                   is verified against checksums of the source blocks
  "migrate"        the 256-block LLaVA point of that sweep (kept as its own key: SCALE_r01 carried it)
  "migrate_under_decode"  the receiver decodes (the bench workload) WHILE it pulls 2 GiB requests on a side stream, per CTA cap of
                   the migration kernel: decode tokens/s alone / under the pull, pull GB/s alone / under decode
`--impl reference` times the CPU restatement of the reference's torch path (oracle/, kind "port") on the box's host cores instead.
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
        # the append launch happens first; arm the pair so the attention call records around its decode kernel
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
            p = build_params()                  # metadata: one pinned buffer, uploaded on the copy stream with q / k / v
            in_ready[b].record(copy_in)
        p.q_cu_seq_lens.record_stream(stream)   # the six metadata arrays are views of one buffer allocated on the copy stream
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

    # ---- extras: the other configs of BASELINE.json (not part of `value`) -----------------------------------------------------
    del batch, kv_cache, params, qkv_dev
    torch.cuda.empty_cache()
    extras_t0 = time.time()
    prefill = None
    if world == 1 and not args.no_extras:
        # the tensor-bound kernels are the clock-sensitive ones: sample the clocks of THIS leg too, so that a box that throttles under
        # the tensor load (sw_power_cap / thermal) is visible next to the fractions it produced
        pre_sampler = ClockSampler(local)
        pre_sampler.start()
        prefill = measure_prefill(dev)
        prefill["clocks"] = pre_sampler.stop()
    cfg4 = measure_cfg4(rank, world, dev) if not args.no_extras else None
    migrate, migrate_sweep = measure_migration_sweep(rank, world, local, dev) if not args.no_extras else (None, None)
    migrate_under_decode = measure_migration_under_decode(rank, world, local, dev) if not args.no_extras else None
    extras_s = time.time() - extras_t0

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
                traffic = json.loads(tp.read_text()).get("bench_decode_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        # the reference's CPU path on this box's host cores: rank 0, N = 1 only (at N > 1 the other ranks' processes share the cores)
        cpu = cpu_reference_run(BATCH, 10, 3, budget_s=45.0) if world == 1 else None
        ref_gpu = reference_gpu_baseline(dev, args.steps, args.warmup) if world == 1 and not args.no_extras else None
        line = {
            "metric": "decode-attn tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "tokens_per_step_per_gpu": BATCH, "l2": "inputs_larger_than_l2 (2.1 GB of KV per step vs 126 MB L2)",
                       "parallelism": f"sequences sharded, {world} independent rank(s), no collective", "kernel": "scatter_rows_kernel + paged_decode_tc_kernel<bf16,1,8> (TMA-fed KV stream, tokens on the tcgen05 M side; one CTA per (row, KV head), unsplit at this size)"},
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
            "migrate": migrate, "migrate_sweep": migrate_sweep, "migrate_under_decode": migrate_under_decode, "prefill": prefill, "cfg4": cfg4,
            "reference_gpu_baseline": ref_gpu, "extras_seconds": extras_s,
        }
        emit_result(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---- extras ----------------------------------------------------------------------------------------------------------------------
def load_tensor_peaks() -> tuple[float, float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json bf16_tflops / bf16_tflops_sustained)"
        except Exception:
            pass
    return 1600.0, 1400.0, "fallback (B200_PROFILING.md)"


def _hi_events(n: int):
    from hydrainfer_b200 import _lib
    evs = []
    for _ in range(n):
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.hi_event_create(ctypes.byref(h)))
        evs.append(h)
    return evs


def _layer_case(seq_lens, hq, hkv, dev, seed):
    """A batch of the given (q, kv) lengths + the layer call over it (KV append + attention through the public layer API)."""
    from hydrainfer_b200.layer import AttentionParametersBuilder, CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
    from hydrainfer_b200.memory import KVCache
    from hydrainfer_b200.workloads import make_batch
    batch = make_batch(seq_lens, hq, hkv, D, BS, dtype=DTYPE, device=dev, gen_device=dev, seed=seed)
    layer = CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(hq, hkv, D))
    builder = AttentionParametersBuilder(hq, hkv, D, BS, dev)
    for req in batch.requests():
        builder.add_request(*req)
    builder.add_kv_cache(KVCache(batch.key_cache, batch.value_cache))
    params = builder.build_attention_parameters()[0]

    def step():
        return layer(batch.query, batch.key, batch.value, params).o

    return batch, step


def cfg3_lengths():
    """48 decode sequences with L ~ U{256..8192} (seed 0) + 4 chunked-prefill sequences with q = 512."""
    g = torch.Generator().manual_seed(0)
    dec = [(1, int(L)) for L in torch.randint(256, 8193, (48,), generator=g).tolist()]
    pre = [(512, 512), (512, 2048), (512, 4096), (512, 8192)]
    return dec, pre


def measure_prefill(dev: torch.device, reps: int = 10, warm: int = 3) -> dict:
    """BASELINE config 3 (Qwen2-VL-7B head geometry) on one GPU: the tensor-bound half of the metric."""
    from hydrainfer_b200 import _lib
    hq, hkv = 28, 4
    dec, pre = cfg3_lengths()
    cases = {"pre1k": [(1024, 1024)] * 8, "pre8k": [(8192, 8192)], "cfg3p": pre, "cfg3mix": dec + pre}
    burst, sustained, src = load_tensor_peaks()
    hbm, _ = load_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB of L2, written between timed calls
    out = {"peak_tflops_burst": burst, "peak_tflops_sustained": sustained, "peak_source": src,
           "flops": "4*Hq*d*sum_i[q_i*(L_i-q_i) + q_i*(q_i+1)/2] (causal, QK^T + PV, 2 flop/MAC; SURVEY §8d)",
           "timing": f"CUDA events around each layer call (KV append + attention), median of {reps} after {warm} warm-up, L2 flushed (256 MiB write) between calls; kernel_ms = the tcgen05 kernel alone (hi_set_kernel_timing_events)",
           "geometry": "Qwen2-VL-7B: 28 q / 4 kv heads, d=128, block 16, bf16"}
    for name, seq_lens in cases.items():
        batch, step = _layer_case(seq_lens, hq, hkv, dev, seed=0)
        flops = sum(4 * hq * D * (q * (L - q) + q * (q + 1) / 2) for q, L in seq_lens)
        decode_bytes = sum(2 * L * hkv * D * 2 + 2 * hq * D * 2 for q, L in seq_lens if q == 1)
        for _ in range(warm):
            step()
        torch.cuda.synchronize(dev)
        call_ms, kern_ms = [], []
        stream = torch.cuda.current_stream(dev)
        for _ in range(reps):
            flush.zero_()
            evs = _hi_events(2)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            _lib.lib.hi_set_kernel_timing_events(evs[0], evs[1])
            s.record(stream)
            step()
            e.record(stream)
            torch.cuda.synchronize(dev)
            call_ms.append(s.elapsed_time(e))
            ms = ctypes.c_float()
            _lib.check(_lib.lib.hi_event_elapsed_ms(evs[0], evs[1], ctypes.byref(ms)))
            kern_ms.append(ms.value)
            for h in evs:
                _lib.lib.hi_event_destroy(h)
        cm, km = statistics.median(call_ms), statistics.median(kern_ms)
        entry = {"seqs": len(seq_lens), "tokens": batch.n_tokens, "ms": cm, "kernel_ms": km, "flops": flops,
                 "tflops": flops / km / 1e9, "frac": flops / km / 1e9 / burst, "frac_sustained": flops / km / 1e9 / sustained,
                 "tflops_call": flops / cm / 1e9, "frac_call": flops / cm / 1e9 / burst, "tokens_per_s": batch.n_tokens / cm * 1e3}
        # the two rooflines of the launch side by side: tensor time of its FLOPs and HBM time of its algorithmic bytes (every sequence's
        # K / V once, q in, out out); a mixed batch is bound by neither alone, the launch can at best overlap the two
        algo_bytes = sum(2 * L * hkv * D * 2 + 2 * q * hq * D * 2 for q, L in seq_lens)
        tensor_ms, hbm_ms = flops / burst / 1e9, algo_bytes / (hbm * 1e6)
        entry.update({"algorithmic_bytes": algo_bytes, "tensor_bound_ms": tensor_ms, "hbm_bound_ms": hbm_ms,
                      "frac_of_binding_roofline": max(tensor_ms, hbm_ms) / km, "frac_if_the_two_could_not_overlap": (tensor_ms + hbm_ms) / km})
        if decode_bytes:
            entry["decode_rows_kv_bytes"] = decode_bytes
            entry["note"] = "mixed batch: the 48 decode rows stream %.0f MB of KV inside the same launch (%.0f us at the measured HBM rate)" % (
                decode_bytes / 1e6, decode_bytes / (hbm * 1e3))
        out[name] = entry
        del batch, step
        torch.cuda.empty_cache()
    return out


def measure_cfg4(rank: int, world: int, dev: torch.device) -> dict:
    """BASELINE config 4: Qwen2-VL-72B head geometry (64q / 8kv), decode, batch 256 in total, sequence i -> rank i mod N."""
    import torch.distributed as dist
    from hydrainfer_b200.workloads import shard_round_robin
    hq, hkv, total = 64, 8, 256
    hbm, _ = load_peaks()
    out = {"geometry": "Qwen2-VL-72B: 64 q / 8 kv heads, d=128, block 16, bf16; one attention layer call (KV append + paged decode attention)",
           "batch_total": total, "scaling": "strong", "n_gpus": world,
           "timing": "GPU time per layer call from a CUDA graph of 8 calls (replayed 7 times, median), MAX over ranks; eager_ms = the same call issued from Python, 20 back to back"}
    for ctx in (2048, 4096):
        mine = shard_round_robin(total, rank, world)
        batch, step = _layer_case([(1, ctx)] * len(mine), hq, hkv, dev, seed=1000 + rank)
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        # eager
        inner, ts = 20, []
        for _ in range(5):
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(inner):
                step()
            e.record()
            torch.cuda.synchronize(dev)
            ts.append(s.elapsed_time(e) / inner)
        eager_ms = statistics.median(ts)
        # graph
        calls = 8
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            step()
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(calls):
                    step()
        torch.cuda.synchronize(dev)
        graph.replay()
        ts = []
        for _ in range(7):
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            graph.replay()
            e.record()
            torch.cuda.synchronize(dev)
            ts.append(s.elapsed_time(e) / calls)
        graph_ms = statistics.median(ts)
        tms = torch.tensor([graph_ms, eager_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        graph_ms, eager_ms = (float(x) for x in tms.tolist())
        bytes_per_token = 2 * ctx * hkv * D * 2 + 2 * hq * D * 2
        per_gpu_gbs = len(mine) * bytes_per_token / graph_ms / 1e6
        out[f"ctx{ctx}"] = {"seqs_per_gpu": len(mine), "ms": graph_ms, "eager_ms": eager_ms, "tokens_per_s": total / graph_ms * 1e3,
                            "tokens_per_s_eager": total / eager_ms * 1e3, "per_gpu_gbs": per_gpu_gbs, "hbm_frac": per_gpu_gbs / hbm,
                            "bytes_per_token": bytes_per_token}
        del graph, batch, step
        torch.cuda.empty_cache()
    return out


MIGRATION_POOLS = {
    "llava7b": dict(n_layers=32, n_tokens=2, block_size=16, n_heads=32, head_size=128),   # 128 KiB runs, 8 MiB per block
    "qwen2vl7b": dict(n_layers=28, n_tokens=2, block_size=16, n_heads=4, head_size=128),  # 16 KiB runs, 896 KiB per block
}


def _block_checksums(pool: torch.Tensor, blocks: list[int]) -> torch.Tensor:
    """One int64 per block: position-weighted wrap-around sum over every 32-bit word of the block in every (layer, K/V) plane.
    Any missing, misplaced or altered word changes it; computed in slabs so 32 GiB requests do not need 64 GiB of temporaries."""
    L, T, NB = pool.shape[0], pool.shape[1], pool.shape[2]
    words = pool.view(torch.int32).view(L, T, NB, -1)
    W = words.shape[-1]
    w_word = (torch.arange(W, device=pool.device, dtype=torch.int64) % 8191) + 1
    w_plane = (torch.arange(L * T, device=pool.device, dtype=torch.int64) * 1000003 + 17).view(L, T, 1)
    idx = torch.tensor(blocks, dtype=torch.int64, device=pool.device)
    out = torch.empty(len(blocks), dtype=torch.int64, device=pool.device)
    slab = max(1, (256 << 20) // (L * T * W * 8))
    for i in range(0, len(blocks), slab):
        sel = words[:, :, idx[i:i + slab]].to(torch.int64)          # [L, T, n, W]
        out[i:i + slab] = ((sel * w_word).sum(dim=-1) * w_plane).sum(dim=(0, 1))
    return out


def measure_migration(rank: int, world: int, local: int, dev: torch.device, pool_name: str, n_move: int, reps: int = 5) -> dict:
    """block_migration GB/s for one (pool geometry, blocks per request) point.  N=1: pool -> pool on the same GPU (HBM bound).
    N>1: disjoint pairs (2i -> 2i+1) pull through CUDA-IPC peer mappings over NVLink, all pairs at once; payload bytes one
    direction / max-over-ranks device time.  The same payload is also moved by ONE cudaMemcpyPeerAsync (N>1) / cudaMemcpyAsync
    (N=1) in the same run - the copy-engine figure this kernel is compared with - and every moved block is verified."""
    import torch.distributed as dist
    from hydrainfer_b200 import _lib
    from hydrainfer_b200._C.data_transfer import block_migration as bm
    geom = MIGRATION_POOLS[pool_name]
    pool_blocks = n_move + max(16, n_move // 8)
    shape = (geom["n_layers"], geom["n_tokens"], pool_blocks, geom["block_size"], geom["n_heads"], geom["head_size"])
    bytes_per_block = geom["n_layers"] * geom["n_tokens"] * geom["block_size"] * geom["n_heads"] * geom["head_size"] * 2
    payload = n_move * bytes_per_block
    pool = torch.empty(shape, dtype=DTYPE, device=dev)
    pool.view(torch.int32).random_(-2**31, 2**31 - 1, generator=torch.Generator(device=dev).manual_seed(100 + rank))  # every bit pattern, incl. NaNs
    g = torch.Generator().manual_seed(200 + n_move)  # the same tables on every rank
    src_bt = torch.randperm(pool_blocks, generator=g)[:n_move].tolist()
    dst_bt = torch.randperm(pool_blocks, generator=g)[:n_move].tolist()
    handle = bm.get_ipc_mem_handle(pool)
    is_receiver, peer_local = True, local
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        is_receiver = (rank % 2 == 1)
        src_handle = handles[rank - 1] if is_receiver else handle
        peer_local = local - 1
        dst_pool = pool
        pattern = f"{world // 2} disjoint pair(s) 2i->2i+1 over NVLink P2P (CUDA IPC pull), all at once"
    else:
        src_handle = handle
        dst_pool = torch.empty_like(pool)
        dst_pool.view(torch.int32).random_(-2**31, 2**31 - 1, generator=torch.Generator(device=dev).manual_seed(7))
        pattern = "same-GPU pool->pool (HBM)"

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    per_receiver: list[list[float]] = []  # one entry per timed leg: the median of every receiving rank (the figure is their max)

    def timed(fn) -> float:
        ts = []
        for i in range(reps + 2):
            sync_all()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if is_receiver:
                fn()
            e.record()
            torch.cuda.synchronize(dev)
            if i >= 2:
                ts.append(s.elapsed_time(e))
        t = torch.tensor([statistics.median(ts) if is_receiver else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            every = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            per_receiver.append([round(float(x.item()), 5) for x in every[1::2]])
            return max(per_receiver[-1])
        per_receiver.append([round(float(t.item()), 5)])
        return float(t.item())

    # verification material BEFORE the copy: checksums of the source blocks, computed by the rank that owns them
    src_sums = _block_checksums(pool, src_bt)
    if world > 1:
        gathered = [torch.empty_like(src_sums) for _ in range(world)]
        dist.all_gather(gathered, src_sums)
        expect = gathered[rank - 1] if is_receiver else None
    else:
        expect = src_sums
    ms = timed(lambda: bm.migrate_blocks(src_bt, dst_bt, src_handle, dst_pool, pool_blocks))
    ok = True
    if is_receiver:
        ok = bool(torch.equal(_block_checksums(dst_pool, dst_bt), expect))
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())

    # the copy-engine comparator: the same number of bytes as ONE contiguous copy, same source GPU, same destination GPU
    flat_dst = torch.empty(min(payload, pool.numel() * 2), dtype=torch.uint8, device=dev)
    copy_bytes = flat_dst.numel()
    src_ptr = ctypes.c_void_p(pool.data_ptr())
    if world > 1 and is_receiver:
        raw = (ctypes.c_uint8 * 64).from_buffer_copy(bytes(int(b) & 0xFF for b in src_handle[:64]))
        off = int.from_bytes(bytes(int(b) & 0xFF for b in src_handle[64:]), "little") if len(src_handle) == 72 else 0
        _lib.check(_lib.lib.hi_ipc_open_handle(raw, off, local, ctypes.byref(src_ptr)))
    stream_ptr = _lib.current_stream_ptr(dev)
    copy_ms = timed(lambda: _lib.check(_lib.lib.hi_peer_copy(flat_dst.data_ptr(), local, src_ptr.value, peer_local, copy_bytes, stream_ptr)))
    pairs = max(1, world // 2)
    del pool, dst_pool, flat_dst
    torch.cuda.empty_cache()
    gbs = payload / (ms * 1e-3) / 1e9
    copy_gbs = copy_bytes / (copy_ms * 1e-3) / 1e9
    return {"pool": pool_name, "blocks_per_request": n_move, "bytes_per_request": payload, "run_bytes": bytes_per_block // (geom["n_layers"] * geom["n_tokens"]),
            "ms": ms, "gbs_per_pair": gbs, "gbs_aggregate": pairs * gbs, "pattern": pattern,
            "bit_exact": ok, "verified": f"all {n_move} moved blocks: position-weighted 64-bit checksums of the destination blocks == those the source rank computed",
            "memcpy_peer_gbs": copy_gbs, "memcpy_peer_ms": copy_ms,
            "ms_per_receiver": per_receiver[0], "memcpy_peer_ms_per_receiver": per_receiver[1],
            "memcpy_peer_what": ("cudaMemcpyPeerAsync" if world > 1 else "cudaMemcpyAsync (same GPU)") + f" of {copy_bytes} contiguous bytes, timed in this run",
            "frac_of_memcpy_peer": gbs / copy_gbs, "frac_of_nvlink_900": (gbs / 900.0) if world > 1 else None}


def measure_migration_sweep(rank: int, world: int, local: int, dev: torch.device) -> tuple[dict, dict]:
    sweep = {"unit": "GB/s per receiving GPU, payload bytes one direction", "points": []}
    headline = None
    for pool_name in ("llava7b", "qwen2vl7b"):
        for n_move in (16, 256, 4096):
            point = measure_migration(rank, world, local, dev, pool_name, n_move, reps=3 if n_move >= 4096 and pool_name == "llava7b" else 5)
            sweep["points"].append(point)
            if pool_name == "llava7b" and n_move == 256:
                headline = point
    return headline, sweep


def measure_migration_under_decode(rank: int, world: int, local: int, dev: torch.device) -> dict:
    """What the disaggregated protocol actually does (hydrainfer/cluster/epdnode.py:362-447): the decode node pulls a request's
    pages on its migrate stream WHILE it runs decode steps.  Pairs 2i -> 2i+1; the receiver runs the bench workload's decode layer
    call back to back on the compute stream and, beside it, pulls 256 LLaVA-7B blocks (2 GiB) over and over on a side stream.
    Reported per CTA cap of the migration kernel: decode tokens/s of the receiver (alone and under the pull) and the pull's GB/s
    (alone and under decode).  N = 1: the same on one GPU (pool -> pool), where the two streams share HBM instead of NVLink."""
    import torch.distributed as dist
    from hydrainfer_b200._C.data_transfer import block_migration as bm
    geom = MIGRATION_POOLS["llava7b"]
    n_move, pool_blocks = 256, 288
    shape = (geom["n_layers"], geom["n_tokens"], pool_blocks, geom["block_size"], geom["n_heads"], geom["head_size"])
    payload = n_move * geom["n_layers"] * geom["n_tokens"] * geom["block_size"] * geom["n_heads"] * geom["head_size"] * 2
    pool = torch.empty(shape, dtype=DTYPE, device=dev).normal_()
    g = torch.Generator().manual_seed(300)
    src_bt = torch.randperm(pool_blocks, generator=g)[:n_move].tolist()
    dst_bt = torch.randperm(pool_blocks, generator=g)[:n_move].tolist()
    handle = bm.get_ipc_mem_handle(pool)
    is_receiver = True
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        is_receiver = rank % 2 == 1
        src_handle, dst_pool = (handles[rank - 1] if is_receiver else handle), pool
    else:
        src_handle, dst_pool = handle, torch.empty_like(pool)
    batch, step = _layer_case([(1, CTX)] * BATCH, HQ, HKV, dev, seed=2000 + rank)
    compute, side = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
    n_steps, n_pulls = 40, 3

    def run(decode: bool, pull: bool) -> tuple[float, float]:
        """-> (ms per decode step, ms per pull), each timed on its own stream while the other (if any) runs beside it."""
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if is_receiver:
            if pull:
                side.wait_stream(compute)
                with torch.cuda.stream(side):
                    p0.record(side)
                    for _ in range(n_pulls):
                        bm.migrate_blocks(src_bt, dst_bt, src_handle, dst_pool, pool_blocks)
                    p1.record(side)
            if decode:
                d0.record(compute)
                for _ in range(n_steps):
                    step()
                d1.record(compute)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        if not is_receiver:
            return 0.0, 0.0
        return (d0.elapsed_time(d1) / n_steps if decode else 0.0), (p0.elapsed_time(p1) / n_pulls if pull else 0.0)

    out = {"pattern": (f"{world // 2} pair(s) 2i->2i+1 over NVLink, receiver decodes while it pulls" if world > 1 else "one GPU: pool->pool pull beside the decode step (shared HBM)"),
           "decode_workload": WORKLOAD, "pull": f"{n_move} LLaVA-7B blocks ({payload} bytes) x {n_pulls} back to back on a side stream", "caps": {}}
    for _ in range(2):
        run(True, True)  # warm-up (IPC mapping, workspaces)
    alone_decode_ms, _ = run(True, False)
    for cap in (0, 1184, 296, 148, 64):
        bm.set_max_ctas(cap)
        _, alone_pull_ms = run(False, True)
        both_decode_ms, both_pull_ms = run(True, True)
        vals = torch.tensor([alone_decode_ms, alone_pull_ms, both_decode_ms, both_pull_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        a_d, a_p, b_d, b_p = (float(x) for x in vals.tolist())
        out["caps"]["auto (large transfer from / to peer memory: TMA bulk-copy kernel, 1 CTA of one warp per SM; else load / store kernel, 8 CTAs per SM)" if cap == 0 else str(cap) + " CTAs of the load / store kernel"] = {
            "decode_tokens_per_s_alone": BATCH / a_d * 1e3, "decode_tokens_per_s_under_pull": BATCH / b_d * 1e3, "decode_slowdown": b_d / a_d,
            "pull_gbs_alone": payload / a_p / 1e6, "pull_gbs_under_decode": payload / b_p / 1e6,
            "note": "the pull outlasts or ends inside the 40 timed decode steps depending on its rate; both figures are per-stream event times"}
    bm.set_max_ctas(0)
    del pool, dst_pool, batch, step
    torch.cuda.empty_cache()
    return out


def reference_gpu_baseline(dev: torch.device, steps: int, warmup: int) -> dict:
    """Comparator only (baseline leg, next to cpu_baseline): the REFERENCE's own layer — its python files and its own csrc
    (set_kv_cache kernel + vendored FlashAttention-2 mha_varlen_fwd) compiled unmodified into oracle/_ref — on this GPU on the
    bench workload.  Nothing of hydrainfer_b200 is on that path."""
    try:
        from oracle import reference_tree
        if not (reference_tree.available("flash_attn") and reference_tree.available("kv_cache_kernels")):
            return {"unavailable": "oracle/_ref was not built with the reference's flash_attn (python oracle/build_ref.py --fa2)"}
        from hydrainfer_b200.workloads import make_batch
        ca, mem = reference_tree.import_reference_package(native="ref")
        batch = make_batch([(1, CTX)] * BATCH, HQ, HKV, D, BS, dtype=DTYPE, device=dev, gen_device=dev, seed=0)
        builder = ca.AttentionParametersBuilder(num_qo_heads=HQ, num_kv_heads=HKV, head_dim=D, block_size=BS, device=dev)
        for req in batch.requests():
            builder.add_request(*req)
        builder.add_kv_cache(mem.KVCache(batch.key_cache, batch.value_cache))
        params = builder.build_attention_parameters()[0]
        layer = ca.CausalGroupedQueryPageAttention(ca.CausalGroupedQueryPageAttentionConfig(n_qo_heads=HQ, n_kv_heads=HKV, head_dim=D))
        for _ in range(warmup):
            layer(batch.query, batch.key, batch.value, params)
        torch.cuda.synchronize(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            layer(batch.query, batch.key, batch.value, params)
        e.record()
        torch.cuda.synchronize(dev)
        ms = s.elapsed_time(e) / steps
        return {"value": BATCH / ms * 1e3, "unit": "tokens/s", "ms_per_step": ms, "kind": "reference (oracle/_ref: the reference's python layer + its own kv_cache_kernels.cu and FlashAttention-2 csrc, nvcc 12.9 sm_100, unmodified)",
                "hbm_gbs": BATCH * ALGO_BYTES_PER_TOKEN / ms / 1e6, "steps": steps}
    except Exception as ex:  # comparator only: never fails the bench
        return {"unavailable": repr(ex)[:300]}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-extras", action="store_true", help="skip the prefill / cfg4 / migration-sweep extras (the headline metric only)")
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
