"""CUDA-graph replay of decode steps over static metadata buffers — the reference's (currently unwired)
hydrainfer/model_runner/cuda_graph_model_runner.py:7-76, rebuilt around this package's metadata path.

A decode step of the hot path is 2-3 short launches per layer (fused rope + append, split-KV attention, split merge): for
small batches the host cannot issue them as fast as the GPU retires them.  The runner captures one graph per batch size
over STATIC buffers: the six int32 metadata arrays of AttentionParameters live at fixed offsets of one device buffer
(the reference keeps six separate static tensors, :21-27), the per-step upload is ONE pinned -> device copy, then
`graph.replay()`.  As in the reference, graphs are captured with kv_max_seq_len = cuda_graph_max_seq_len (:41) and a
step is replayed only if every sequence is a decode row and fits that bound (:60-61); anything else runs eagerly.

What makes the kernels replay-safe: every length, block id and slot is read from the device metadata; the host-side values
baked into a captured launch (grid size, split-KV chunking, tensor maps) depend only on the batch size, the pools and
the kv_max_seq_len BOUND, never on the actual lengths (include/hi_b200.h: kv_blocks_hint / work plan "never affect
results"); the library allocates nothing (caller-owned workspace) and its one memset is a graph node.
"""
from __future__ import annotations

from array import array
from typing import Callable, Optional, Sequence

import torch
from torch import Tensor

from ..layer.causal_attention import AttentionParameters, AttentionParametersBuilder
from ..memory.kv_cache import KVCache


class StaticAttentionMetadata:
    """The six metadata arrays at fixed offsets of one device int32 buffer sized for (max_batch_size, max_blocks)."""

    def __init__(self, max_batch_size: int, max_blocks: int, device: torch.device, staging_slots: int = 4):
        self.max_batch_size = max_batch_size
        self.max_blocks = max_blocks
        self.device = torch.device(device)
        b = max_batch_size
        sizes = [b + 1, b + 1, b, b, max_blocks, b + 1]  # q_cu, kv_cu, last_page_len, new_cache_slots, block_tables, cu_blocks
        self.offsets, total = [], 0
        for n in sizes:
            self.offsets.append(total)
            total += (n + 3) & ~3  # 16-byte aligned slices
        self.total = total
        self.buffer = torch.zeros(total, dtype=torch.int32, device=self.device)
        self.staging = [torch.zeros(total, dtype=torch.int32).pin_memory() for _ in range(staging_slots)] if self.device.type == "cuda" else [torch.zeros(total, dtype=torch.int32)]
        self.views = [memoryview(s.numpy()).cast("B").cast("i") for s in self.staging]
        self.events: list[Optional[torch.cuda.Event]] = [None] * len(self.staging)
        self.cursor = 0

    def fits(self, builder: AttentionParametersBuilder) -> bool:
        return builder.num_sequences <= self.max_batch_size and len(builder.block_tables) <= self.max_blocks

    def fill(self, builder: AttentionParametersBuilder) -> None:
        """Copies the builder's arrays into the static buffer (one pinned staging write + one H2D on the current stream)."""
        assert self.fits(builder), "batch does not fit the static metadata buffers"
        i = self.cursor
        self.cursor = (i + 1) % len(self.staging)
        if self.events[i] is not None:
            self.events[i].synchronize()
        view = self.views[i]
        parts = [builder.q_cu_seq_lens, builder.kv_cu_seq_lens, builder.paged_kv_last_page_len, builder.new_cache_slots, builder.block_tables, builder.cu_blocks_lens]
        end = 0
        for part, o in zip(parts, self.offsets):
            if len(part):
                view[o:o + len(part)] = memoryview(part if isinstance(part, array) else array("i", part))
                end = max(end, o + len(part))
        # the block-table region is the long one: copy only up to the last entry in use
        self.buffer[:end].copy_(self.staging[i][:end], non_blocking=True)
        if self.device.type == "cuda":
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.events[i] = ev

    def parameters(self, batch_size: int, kv_caches: Sequence[KVCache], kv_max_seq_len: int, n_blocks_view: Optional[int] = None) -> list[AttentionParameters]:
        """AttentionParameters (one per layer) whose tensors are views of the static buffer for a decode batch of `batch_size`."""
        o = self.offsets
        nb = self.max_blocks if n_blocks_view is None else n_blocks_view
        buf = self.buffer
        return [AttentionParameters(
            kv_cache=kv_cache,
            q_cu_seq_lens=buf[o[0]:o[0] + batch_size + 1], kv_cu_seq_lens=buf[o[1]:o[1] + batch_size + 1],
            paged_kv_last_page_len=buf[o[2]:o[2] + batch_size], new_cache_slots=buf[o[3]:o[3] + batch_size],
            block_tables=buf[o[4]:o[4] + nb], cu_blocks_lens=buf[o[5]:o[5] + batch_size + 1],
            num_sequences=batch_size, all_sequences_decode=True, q_max_seq_len=1, kv_max_seq_len=kv_max_seq_len,
        ) for kv_cache in kv_caches]


class CudaGraphModelRunner:
    """model_runner(hidden [B, width_in], position_ids [B] int32, attention_params: list[AttentionParameters]) -> Tensor [B, width_out]
    is captured once per batch size; forward() replays it for decode-only steps that fit the captured bounds."""

    def __init__(self, model_runner: Callable[[Tensor, Tensor, list], Tensor], dtype: torch.dtype, device: torch.device, block_size: int,
                 width_in: int, width_out: int, kv_caches: Sequence[KVCache], n_qo_heads: int, n_kv_heads: int, head_dim: int,
                 cuda_graph_max_batch_size: int = 64, cuda_graph_max_seq_len: int = 1024, batch_sizes: Optional[Sequence[int]] = None):
        self.model_runner = model_runner
        self.dtype = dtype
        self.device = torch.device(device)
        self.block_size = block_size
        self.kv_caches = list(kv_caches)
        self.geometry = (n_qo_heads, n_kv_heads, head_dim)
        self.max_batch_size = cuda_graph_max_batch_size
        self.max_seq_len = cuda_graph_max_seq_len
        max_blocks_per_seq = (cuda_graph_max_seq_len + block_size - 1) // block_size
        self.metadata = StaticAttentionMetadata(cuda_graph_max_batch_size, cuda_graph_max_batch_size * max_blocks_per_seq, self.device)
        self.static_hidden = torch.zeros((cuda_graph_max_batch_size, width_in), dtype=dtype, device=self.device)
        self.static_position_ids = torch.zeros(cuda_graph_max_batch_size, dtype=torch.int32, device=self.device)
        self.static_out = torch.zeros((cuda_graph_max_batch_size, width_out), dtype=dtype, device=self.device)
        self.graphs: dict[int, torch.cuda.CUDAGraph] = {}
        self.replays = 0
        self.eager_calls = 0
        self._capture(list(batch_sizes) if batch_sizes is not None else list(range(1, cuda_graph_max_batch_size + 1)))

    def _warmup_metadata(self, batch_size: int) -> None:
        """Valid metadata for the capture-time runs: each sequence has one token in its own block (block ids 0 .. B-1)."""
        n_qo, n_kv, d = self.geometry
        builder = AttentionParametersBuilder(n_qo, n_kv, d, self.block_size, self.device)
        for b in range(batch_size):
            builder.add_request(1, 1, [b * self.block_size], [b])
        self.metadata.fill(builder)

    def _capture(self, batch_sizes: list[int]) -> None:
        # The capture-time runs append their (zero) K/V rows to slot 0 of blocks 0 .. B-1 (see _warmup_metadata); those rows are
        # saved and restored so that capturing next to a live pool does not disturb it (the reference captures on whatever
        # its uninitialised static slot buffer holds, cuda_graph_model_runner.py:25, 46-51).
        n_touch = min(max(batch_sizes), min(c.key_cache.shape[0] for c in self.kv_caches)) if batch_sizes else 0
        saved = [(c.key_cache[:n_touch, 0].clone(), c.value_cache[:n_touch, 0].clone()) for c in self.kv_caches]
        pool = None
        for batch_size in sorted(batch_sizes, reverse=True):  # largest first: smaller graphs reuse its memory pool
            self._warmup_metadata(batch_size)
            params = self.metadata.parameters(batch_size, self.kv_caches, self.max_seq_len)
            hidden, pos = self.static_hidden[:batch_size], self.static_position_ids[:batch_size]
            # eager runs first: one-time work (cudaFuncSetAttribute, tensor-map cache, workspace allocation) must not be captured
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.model_runner(hidden, pos, params)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self.static_out[:batch_size] = self.model_runner(hidden, pos, params)
            if pool is None:
                pool = g.pool()
            self.graphs[batch_size] = g
        torch.cuda.synchronize(self.device)
        for c, (k0, v0) in zip(self.kv_caches, saved):
            c.key_cache[:n_touch, 0].copy_(k0)
            c.value_cache[:n_touch, 0].copy_(v0)
        torch.cuda.synchronize(self.device)

    def forward(self, hidden: Tensor, position_ids: Tensor, builder: AttentionParametersBuilder) -> Tensor:
        """One step.  `builder` has had add_request() called for every sequence (add_kv_cache is not needed: the runner owns
        the caches).  Decode-only batches within the captured bounds are replayed; others run eagerly on freshly built params."""
        batch_size = builder.num_sequences
        if (builder.all_sequences_decode and batch_size in self.graphs and builder.kv_max_seq_len <= self.max_seq_len
                and hidden.shape[0] == batch_size and self.metadata.fits(builder)):
            self.metadata.fill(builder)
            self.static_hidden[:batch_size].copy_(hidden, non_blocking=True)
            self.static_position_ids[:batch_size].copy_(position_ids, non_blocking=True)
            self.graphs[batch_size].replay()
            self.replays += 1
            return self.static_out[:batch_size]
        self.eager_calls += 1
        builder.kv_caches = list(self.kv_caches)
        return self.model_runner(hidden, position_ids, builder.build_attention_parameters())

    __call__ = forward
