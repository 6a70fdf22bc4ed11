"""Paged causal grouped-query attention layer with the reference's interface
(hydrainfer/layer/causal_attention.py): AttentionParameters (:31-68), AttentionParametersBuilder (:110-210),
CausalGroupedQueryPageAttentionConfig / Output (:213-222) and CausalGroupedQueryPageAttention.forward (:394-406).

The reference dispatches through a chain FlashInfer -> FlashAttention(csrc) -> Torch (:385-392).  Here there is one
handler, B200CausalGroupedQueryPageAttentionHandler, backed by the sm_100a kernels; it keeps the handler protocol
(`next_handler`, `forward(query, attention_params)`), so it can also be inserted at the head of the reference's own
chain.  CPU tensors are passed to `next_handler` if one is linked and rejected otherwise: this
package ships no CPU implementation (the fp32 restatement lives in oracle/ and is test infrastructure only).
"""
from __future__ import annotations

import math
from array import array
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor, nn

from .._C.kernel.flash_attn import append_and_attend, mha_varlen_fwd
from ..memory.kv_cache import KVCache


@dataclass
class AttentionParameters:
    """Per-step attention metadata; one instance per layer, all sharing the same int32 device tensors."""
    kv_cache: KVCache
    q_cu_seq_lens: Tensor = None           # int32 [n_seqs + 1]  cumulative new tokens
    kv_cu_seq_lens: Tensor = None          # int32 [n_seqs + 1]  cumulative cached + new tokens
    paged_kv_last_page_len: Tensor = None  # int32 [n_seqs]      tokens in each sequence's last block
    new_cache_slots: Tensor = None         # int32 [n_tokens]    physical slot of every new token
    block_tables: Tensor = None            # int32 [sum blocks]  flattened per-sequence block ids
    cu_blocks_lens: Tensor = None          # int32 [n_seqs + 1]
    num_sequences: int = None
    all_sequences_decode: bool = False
    q_max_seq_len: int = 128
    kv_max_seq_len: int = 128
    flash_infer_handler: Optional[object] = None  # kept for signature compatibility; unused
    # host-side plan for the B200 prefill kernel (the analogue of flashinfer's plan(), causal_attention.py:171-195): the
    # query tiles of the batch as (sequence, tile) pairs sorted by cost, and the total number of keys they walk.
    # Results never depend on them.
    work_items: Tensor = None              # int32 [n_items, 2]; None for decode-only batches
    work_tile_tokens: int = 0
    qk_work: int = 0

    _TENSOR_FIELDS = ("q_cu_seq_lens", "kv_cu_seq_lens", "paged_kv_last_page_len", "new_cache_slots", "block_tables", "cu_blocks_lens", "work_items")

    def to(self, device: torch.device) -> None:
        for name in self._TENSOR_FIELDS:
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.to(device))


class _PinnedStaging:
    """Per-device ring of pinned int32 staging buffers for the per-step metadata upload.  pin_memory() on a fresh tensor
    costs a cudaHostAlloc per step; the ring allocates once and an event per slot guards reuse while a copy is in flight."""

    _rings: dict = {}

    def __init__(self, device: torch.device, slots: int = 4):
        self.device = device
        self.buffers = [torch.empty(0, dtype=torch.int32) for _ in range(slots)]
        self.views = [memoryview(b"").cast("i") for _ in range(slots)]  # int32 memoryviews of the pinned buffers
        self.events = [None] * slots
        self.cursor = 0

    @classmethod
    def get(cls, device: torch.device) -> "_PinnedStaging":
        ring = cls._rings.get(device)
        if ring is None:
            ring = cls._rings[device] = cls(device)
        return ring

    def upload(self, parts: list, offsets: list[int], total: int) -> Tensor:
        """Writes the int32 arrays `parts` at `offsets` of one pinned buffer and starts its H2D copy."""
        i = self.cursor
        self.cursor = (i + 1) % len(self.buffers)
        if self.events[i] is not None:
            self.events[i].synchronize()  # the copy that last used this slot has finished
        if self.buffers[i].numel() < total:
            self.buffers[i] = torch.empty(max(total, 4096, 2 * self.buffers[i].numel()), dtype=torch.int32).pin_memory()
            self.views[i] = memoryview(self.buffers[i].numpy()).cast("B").cast("i")
        view = self.views[i]
        for part, o in zip(parts, offsets):
            if len(part):
                view[o:o + len(part)] = memoryview(part)  # memcpy
        dev = self.buffers[i][:total].to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.events[i] = ev
        return dev


class _BlockTableCache:
    """int32 images of the per-sequence block-table lists.

    The engine passes `virtual_kv_cache.block_table` to add_request every step (reference engine/parameters_builder.py:
    63-69): the SAME list object for the lifetime of a request, grown in place by whole blocks
    (token_cache_manger.py:150-159).  Converting its Python ints to int32 again every step is the largest host cost of a
    decode step (8192 ints for batch 64 at context 2048), so the converted image is kept per list object and only the
    appended tail is converted.  An entry is used only after its snapshot compares equal to the list (a C-speed list
    compare), so a recycled id() or an in-place edit can never produce a stale table."""

    def __init__(self, capacity: int = 8192):
        self.capacity = capacity
        self.entries: dict[int, tuple[list[int], array]] = {}

    def image(self, table: list[int]) -> array:
        key = id(table)
        hit = self.entries.get(key)
        if hit is not None:
            snap, img = hit
            n = len(snap)
            if len(table) == n:
                if table == snap:
                    return img
            elif len(table) > n and table[:n] == snap:  # grown by whole blocks since the last step
                tail = table[n:]
                img = img + array("i", tail)  # a new array: images handed out earlier stay intact
                self.entries[key] = (snap + tail, img)
                return img
        img = array("i", table)
        if len(self.entries) >= self.capacity:
            self.entries.clear()
        self.entries[key] = (list(table), img)
        return img


_block_table_cache = _BlockTableCache()


class AttentionParametersBuilder:
    """add_request() per sequence, add_kv_cache() per layer, then build_attention_parameters()
    (causal_attention.py:110-210).  The six metadata arrays are uploaded ONCE as a single pinned int32 buffer and
    sliced on the device (the reference does six torch.tensor(list) H2D copies, :163-168); they are accumulated as
    int32 `array`s, not Python lists, so assembling the buffer is a handful of memcpys."""

    def __init__(self, num_qo_heads: int, num_kv_heads: int, head_dim: int, block_size: int, device: torch.device,
                 flash_infer_batch_prefill_handler=None, flash_infer_batch_decode_handler=None):
        self.num_qo_heads = num_qo_heads
        self.num_kv_heads = num_kv_heads
        self.head_dim = head_dim
        self.block_size = block_size
        self.device = torch.device(device)
        self.kv_caches: list[KVCache] = []
        self.q_cu_seq_lens = array("i", [0])
        self.kv_cu_seq_lens = array("i", [0])
        self.paged_kv_last_page_len = array("i")
        self.new_cache_slots = array("i")
        self.block_tables = array("i")
        self.cu_blocks_lens = array("i", [0])
        self.num_sequences = 0
        self.all_sequences_decode = True
        self.q_max_seq_len = 0
        self.kv_max_seq_len = 0
        self.seq_lens: list[tuple[int, int]] = []  # (q, kv) per sequence, for the host-side plan

    def add_request(self, q_seq_len: int, kv_seq_len: int, new_cache_slots: list[int], block_table: list[int]) -> None:
        self.q_cu_seq_lens.append(self.q_cu_seq_lens[-1] + q_seq_len)
        self.kv_cu_seq_lens.append(self.kv_cu_seq_lens[-1] + kv_seq_len)
        self.paged_kv_last_page_len.append((kv_seq_len + self.block_size - 1) % self.block_size + 1)
        self.new_cache_slots.fromlist(new_cache_slots) if type(new_cache_slots) is list else self.new_cache_slots.extend(new_cache_slots)
        self.block_tables.extend(_block_table_cache.image(block_table) if type(block_table) is list else array("i", block_table))
        self.cu_blocks_lens.append(self.cu_blocks_lens[-1] + len(block_table))
        self.num_sequences += 1
        if q_seq_len != 1:
            self.all_sequences_decode = False
        if q_seq_len > self.q_max_seq_len:
            self.q_max_seq_len = q_seq_len
        if kv_seq_len > self.kv_max_seq_len:
            self.kv_max_seq_len = kv_seq_len
        self.seq_lens.append((q_seq_len, kv_seq_len))

    def add_kv_cache(self, kv_cache: KVCache) -> None:
        self.kv_caches.append(kv_cache)

    def build_attention_parameters(self) -> list[AttentionParameters]:
        work_flat, tile_tokens, qk_work = self._plan()
        parts = [self.q_cu_seq_lens, self.kv_cu_seq_lens, self.paged_kv_last_page_len, self.new_cache_slots,
                 self.block_tables, self.cu_blocks_lens, work_flat]
        # each slice starts on a 16-byte boundary (4 int32) so the kernels' vector paths never see a misaligned table
        offsets, total = [], 0
        for part in parts:
            offsets.append(total)
            total += (len(part) + 3) & ~3
        if self.device.type == "cuda":
            dev = _PinnedStaging.get(self.device).upload(parts, offsets, total)
        else:
            dev = torch.zeros(total, dtype=torch.int32)
            for part, o in zip(parts, offsets):
                if len(part):
                    dev[o:o + len(part)] = torch.frombuffer(part, dtype=torch.int32)
        views = [dev[o:o + len(part)] for o, part in zip(offsets, parts)]
        return [AttentionParameters(
            kv_cache=kv_cache,
            q_cu_seq_lens=views[0], kv_cu_seq_lens=views[1], paged_kv_last_page_len=views[2], new_cache_slots=views[3],
            block_tables=views[4], cu_blocks_lens=views[5],
            num_sequences=self.num_sequences, all_sequences_decode=self.all_sequences_decode,
            q_max_seq_len=self.q_max_seq_len, kv_max_seq_len=self.kv_max_seq_len, flash_infer_handler=None,
            work_items=views[6].view(-1, 2) if len(work_flat) else None, work_tile_tokens=tile_tokens, qk_work=qk_work,
        ) for kv_cache in self.kv_caches]

    def _plan(self) -> tuple[array, int, int]:
        """Host-side plan for batches with prefill rows: every run of `tile_tokens` query tokens of a sequence is one work
        item whose cost is the number of keys its last token sees; items go out heaviest first (longest-processing-time
        order), so the launch ends on light tiles, and the grid holds exactly the tiles that exist."""
        empty = array("i")
        if self.all_sequences_decode or self.device.type != "cuda":
            return empty, 0, 0
        from .. import _lib
        tile_tokens = int(_lib.lib.hi_attention_tile_tokens(self.num_qo_heads, self.num_kv_heads))
        if tile_tokens <= 0:
            return empty, 0, 0
        items = []
        for b, (q, kv) in enumerate(self.seq_lens):
            for tile in range((q + tile_tokens - 1) // tile_tokens):
                items.append((kv - q + min(q, (tile + 1) * tile_tokens), b, tile))
        items.sort(reverse=True)
        flat = array("i", [x for _, b, tile in items for x in (b, tile)])
        return flat, tile_tokens, sum(cost for cost, _, _ in items)


@dataclass
class CausalGroupedQueryPageAttentionConfig:
    n_qo_heads: int
    n_kv_heads: int
    head_dim: int


@dataclass
class CausalGroupedQueryPageAttentionOutput:
    o: Tensor


class B200CausalGroupedQueryPageAttentionHandler(nn.Module):
    """Handler backed by hi_paged_attention.  query [n_tokens, n_qo_heads, head_dim] (rows may be strided),
    KV already appended; returns o [n_tokens, n_qo_heads * head_dim] in the query dtype."""

    def __init__(self, config: CausalGroupedQueryPageAttentionConfig):
        super().__init__()
        assert config.n_qo_heads % config.n_kv_heads == 0, f"n_qo_heads {config.n_qo_heads} is not divisible by n_kv_heads {config.n_kv_heads}"
        self.n_qo_heads = config.n_qo_heads
        self.n_kv_heads = config.n_kv_heads
        self.head_dim = config.head_dim
        self.next_handler: Optional[nn.Module] = None
        self.path = 0  # HiAttnPath: 0 auto, 1 split-KV CUDA-core kernel, 2 tcgen05 tile kernel (tests force each)

    def forward(self, query: Tensor, attention_params: AttentionParameters) -> CausalGroupedQueryPageAttentionOutput:
        if query.device.type != "cuda":
            if self.next_handler is not None:
                return self.next_handler(query, attention_params)
            raise RuntimeError("hydrainfer_b200: attention needs CUDA tensors; no CPU handler is linked")
        key_cache, value_cache = attention_params.kv_cache.get_kv_cache()
        output = torch.empty((query.shape[0], self.n_qo_heads, self.head_dim), dtype=query.dtype, device=query.device)
        try:
            mha_varlen_fwd(
                output, query, key_cache, value_cache,
                attention_params.q_cu_seq_lens, attention_params.kv_cu_seq_lens,
                attention_params.block_tables, attention_params.cu_blocks_lens,
                None, attention_params.q_max_seq_len, attention_params.kv_max_seq_len,
                1.0 / math.sqrt(self.head_dim), 0, -1, 0, 0, self.path,
                getattr(attention_params, "work_items", None), getattr(attention_params, "work_tile_tokens", 0),
                getattr(attention_params, "qk_work", 0),
            )
        except RuntimeError as e:
            # a geometry the kernels do not cover (HI_ERR_UNSUPPORTED: head_dim outside {64, 128, 256}, odd strides, ...)
            # goes down the chain when this handler heads the reference's own list; with no next handler it is an error
            if self.next_handler is not None and _is_unsupported(e):
                return self.next_handler(query, attention_params)
            raise
        return CausalGroupedQueryPageAttentionOutput(o=output.view(-1, self.n_qo_heads * self.head_dim))


def _is_unsupported(e: RuntimeError) -> bool:
    return "hi_b200 error -2" in str(e)


class CausalGroupedQueryPageAttention(nn.Module):
    """forward(query [T, Hq*d], key [T, Hkv*d], value [T, Hkv*d], attention_params) -> Output(o [T, Hq*d]):
    view, append K/V to the paged cache, attend (causal_attention.py:394-406)."""

    def __init__(self, config: CausalGroupedQueryPageAttentionConfig):
        super().__init__()
        assert config.n_qo_heads % config.n_kv_heads == 0, f"n_qo_heads {config.n_qo_heads} is not divisible by n_kv_heads {config.n_kv_heads}"
        self.n_qo_heads = config.n_qo_heads
        self.n_kv_heads = config.n_kv_heads
        self.head_dim = config.head_dim
        self.softmax_scale = 1.0 / math.sqrt(config.head_dim)
        self.handlers = [B200CausalGroupedQueryPageAttentionHandler(config)]
        self.handler = self.handlers[0]

    def forward(self, query: Tensor, key: Tensor, value: Tensor, attention_params: AttentionParameters) -> CausalGroupedQueryPageAttentionOutput:
        handler = self.handler
        if type(handler) is B200CausalGroupedQueryPageAttentionHandler and handler.next_handler is None and query.is_cuda:
            # view + append + attend in ONE call into the compiled module (the steps of :396-405 back to back on the current stream)
            p = attention_params
            kv = p.kv_cache
            o = append_and_attend(
                query, key, value, p.new_cache_slots, kv.key_cache, kv.value_cache, p.q_cu_seq_lens, p.kv_cu_seq_lens, p.block_tables, p.cu_blocks_lens,
                p.q_max_seq_len, p.kv_max_seq_len, self.softmax_scale, handler.path, p.work_items, p.work_tile_tokens, p.qk_work)
            return CausalGroupedQueryPageAttentionOutput(o=o if o.dim() == 2 else o.view(o.shape[0], -1))  # 3-D in -> 3-D out of the extension
        n_tokens = query.shape[0]
        query = query.view(n_tokens, self.n_qo_heads, self.head_dim)
        key = key.view(n_tokens, self.n_kv_heads, self.head_dim)
        value = value.view(n_tokens, self.n_kv_heads, self.head_dim)
        attention_params.kv_cache.set_kv_cache(attention_params.new_cache_slots, key, value)
        return handler(query, attention_params)
