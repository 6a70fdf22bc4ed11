"""CPU oracle for hydrainfer's paged-KV attention hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain torch-on-CPU / Python, what the reference computes on this path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it; nothing under
hydrainfer_b200/ does (the product path has no CPU fallback).

Parity pin: oracle/make_golden.py imports the REAL reference from /root/reference (Python torch path, runnable on CPU)
and checks every function below against it on seeded inputs, then freezes the reference's outputs as fixtures in
tests/golden/*.npz; tests/test_oracle_golden.py re-checks the oracle against those fixtures wherever the reference is
absent (the GPU box).  Pinned this way: attention (Torch handler), set_kv_cache / set_image_cache (Python fallbacks),
BlockAllocator, AttentionParametersBuilder metadata, v2p, rotary embedding (Torch handler, both table dtypes), the
ROPE attention module (model_forward.py) and the un-paged vision attention (both Torch handlers of
multihead_attention.py).  NOT pinned ("parity unpinned"): migrate_blocks — the
reference implementation is CUDA-only C++ with no test or fixture anywhere in the reference tree; its
index arithmetic is restated from the source alone.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
from torch import Tensor


# ---------------------------------------------------------------------------------------------------------------
# KV append
# ---------------------------------------------------------------------------------------------------------------
def set_kv_cache(slot_ids: Tensor, keys: Tensor, values: Tensor, key_cache: Tensor, value_cache: Tensor) -> None:
    """hydrainfer/memory/kv_cache.py:44-50 (Python fallback of KVCache.set_kv_cache), same result as
    csrc/kernel/kv_cache_kernels/kv_cache_kernels.cu:17-58:
        cache[slot // block_size, slot % block_size, :, :] = rows[i, :, :]   for i in order
    Written token by token so duplicate slots resolve the way the reference's loop resolves them (last write wins)."""
    block_size = key_cache.shape[1]
    for i in range(slot_ids.shape[0]):
        slot = int(slot_ids[i])
        block_id, block_offset = slot // block_size, slot % block_size
        key_cache[block_id, block_offset, :, :] = keys[i, :, :]
        value_cache[block_id, block_offset, :, :] = values[i, :, :]


def set_image_cache(slot_ids: Tensor, image_tokens: Tensor, image_cache: Tensor) -> None:
    """hydrainfer/memory/token_cache.py:53-56 (slot_view[slot_ids] = value) ==
    csrc/kernel/cache_kernels/cache_kernels.cu:17-53."""
    slot_view = image_cache.view(-1, image_cache.shape[-2], image_cache.shape[-1])
    slot_view[slot_ids.long(), :, :] = image_tokens


def get_image_cache(slot_ids: Tensor, image_cache: Tensor) -> Tensor:
    """hydrainfer/engine/parameters_builder.py:50-54: image_token_cache.view(-1, n_heads * head_dim)[slot_ids, :]."""
    return image_cache.view(-1, image_cache.shape[-2] * image_cache.shape[-1])[slot_ids.long(), :]


# ---------------------------------------------------------------------------------------------------------------
# Un-paged vision attention
# ---------------------------------------------------------------------------------------------------------------
def multi_head_attention(query: Tensor, key: Tensor, value: Tensor, n_heads: int, head_dim: int) -> Tensor:
    """TorchMultiHeadAttentionHandler.forward (hydrainfer/layer/multihead_attention.py:48-73): [batch, seq, hidden] inputs
    cast to fp32, q scaled by 1/sqrt(d) BEFORE the product (:60), bmm, softmax, bmm, cast back to the input dtype (:66)."""
    batch_size, seq_len, hidden_size = query.shape
    dtype = query.dtype

    def heads(t: Tensor) -> Tensor:
        return t.view(-1, seq_len, n_heads, head_dim).transpose(1, 2).contiguous().to(torch.float).view(-1, seq_len, head_dim)

    q, k, v = heads(query), heads(key), heads(value)
    q = q * (1. / math.sqrt(head_dim))
    score = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
    o = torch.bmm(score, v).view(batch_size, n_heads, seq_len, head_dim).transpose(1, 2).contiguous()
    return o.view(batch_size, seq_len, hidden_size).to(dtype)


def qwen_multi_head_attention(q: Tensor, k: Tensor, v: Tensor, seq_length: int, cu_seqlens, head_dim: int) -> Tensor:
    """QwenTorchMultiHeadAttentionHandler.forward (hydrainfer/layer/multihead_attention.py:241-256): packed
    [seq_length, n_heads, head_dim] inputs, block-diagonal additive mask of finfo.min outside each cu_seqlens segment
    (:243-245), products in the INPUT dtype, softmax in fp32 then cast back (:251), -> [seq_length, hidden]."""
    mask = torch.full([1, seq_length, seq_length], torch.finfo(q.dtype).min, dtype=q.dtype)
    for i in range(1, len(cu_seqlens)):
        mask[..., int(cu_seqlens[i - 1]):int(cu_seqlens[i]), int(cu_seqlens[i - 1]):int(cu_seqlens[i])] = 0
    qh, kh, vh = q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1)
    w = torch.matmul(qh, kh.transpose(1, 2)) / math.sqrt(head_dim)
    w = w + mask
    w = torch.nn.functional.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    return torch.matmul(w, vh).transpose(0, 1).reshape(seq_length, -1)


def varlen_attention_fp32(q: Tensor, k: Tensor, v: Tensor, cu_seqlens_q, cu_seqlens_k, causal: bool = False) -> Tensor:
    """fp32 recompute of what mha_varlen_fwd computes in its un-paged form (csrc/kernel/flash_attn/flash_api.cpp:216-355
    with block_table None): q [Tq, Hq, d], k/v [Tk, Hkv, d]; per sequence softmax(q k^T / sqrt(d)) v, optionally with the
    bottom-right aligned causal mask.  The tolerance target of the GPU tests.  Returns [Tq, Hq * d] fp32."""
    tq, hq, d = q.shape
    group = hq // k.shape[1]
    out = torch.zeros(tq, hq, d, dtype=torch.float32)
    for b in range(len(cu_seqlens_q) - 1):
        q0, q1, k0, k1 = int(cu_seqlens_q[b]), int(cu_seqlens_q[b + 1]), int(cu_seqlens_k[b]), int(cu_seqlens_k[b + 1])
        if q1 == q0 or k1 == k0:
            continue
        qb = q[q0:q1].float()
        kb = k[k0:k1].float().repeat_interleave(group, dim=1)
        vb = v[k0:k1].float().repeat_interleave(group, dim=1)
        s = torch.einsum("qhd,khd->hqk", qb, kb) / math.sqrt(d)
        if causal:
            i = torch.arange(q1 - q0)[:, None]
            j = torch.arange(k1 - k0)[None, :]
            s = s.masked_fill((j - i) > ((k1 - k0) - (q1 - q0)), float("-inf"))
        out[q0:q1] = torch.einsum("hqk,khd->qhd", torch.softmax(s, dim=-1), vb)
    return out.view(tq, hq * d)


# ---------------------------------------------------------------------------------------------------------------
# Rotary embedding (the step in front of the append)
# ---------------------------------------------------------------------------------------------------------------
def rotary_cos_sin_table(rotary_dim: int, max_position_embeddings: int, inv_freq: Tensor) -> Tensor:
    """fp32 [max_positions, 2, rotary_dim/2]: cos then sin of position x inv_freq — the fused handler's cache layout
    (hydrainfer/layer/rotary_embedding.py:110-115); the torch handler holds the same numbers repeated per pair (:30-41).
    Cast it to the model dtype to mimic `module.to(dtype)`."""
    t = torch.arange(max_position_embeddings, dtype=torch.float)
    freqs = torch.einsum("i,j->ij", t, inv_freq.float())
    return torch.stack([freqs.cos(), freqs.sin()], dim=1)


def apply_rotary(query: Tensor, key: Tensor, positions: Tensor, cos_sin: Tensor, rotary_dim: int, interleaved: bool) -> tuple[Tensor, Tensor]:
    """TorchRotaryEmbeddingHandler.forward (hydrainfer/layer/rotary_embedding.py:44-83), same result as the CUDA kernel
    csrc/kernel/position_embedding/rope.cu:9-31: for each pair (x, y) of a head — (2i, 2i+1) when interleaved (:58-70,
    rotate_every_two :85-93), else (i, i + rotary_dim/2) (:71-81, rotate_half :95-99) —
        x' = x*cos + (-y)*sin        y' = y*cos + x*sin
    as separate elementwise torch ops, so a 16-bit table rounds every product and the sum to 16 bits and an fp32 table
    promotes to fp32 and rounds once in the final `.to(dtype)` (:83).  Dims >= rotary_dim pass through (:51-54, :81-82).
    Returns new tensors [T, H, d] in the input dtype."""
    n = rotary_dim // 2
    cs = cos_sin[positions.long()]                      # [T, 2, n]
    c, s = cs[:, 0, None, :], cs[:, 1, None, :]         # broadcast over heads

    def rotate(t: Tensor) -> Tensor:
        rot, rest = t[:, :, :rotary_dim], t[:, :, rotary_dim:]
        if interleaved:
            x, y = rot[:, :, 0::2], rot[:, :, 1::2]
        else:
            x, y = rot[:, :, :n], rot[:, :, n:]
        xo = x * c + (-y) * s
        yo = y * c + x * s
        out = torch.stack([xo, yo], dim=-1).flatten(start_dim=-2) if interleaved else torch.cat([xo, yo], dim=-1)
        return torch.cat([out, rest.to(out.dtype)], dim=-1).to(t.dtype)

    return rotate(query), rotate(key)


def rope_attention_layer_forward(query: Tensor, key: Tensor, value: Tensor, positions: Tensor, cos_sin: Tensor, rotary_dim: int,
                                 interleaved: bool, key_cache: Tensor, value_cache: Tensor, new_cache_slots: Tensor, q_cu_seq_lens,
                                 kv_cu_seq_lens, block_tables: Tensor, cu_blocks_lens, n_qo_heads: int, n_kv_heads: int, head_dim: int):
    """ROPECausalGroupedQueryPageAttention.forward without projections (hydrainfer/model/model_forward.py:78-83): rotary,
    then the attention layer (append + attend).  Returns (o [T, Hq*d], rotated query, rotated key)."""
    q = query.reshape(-1, n_qo_heads, head_dim)
    k = key.reshape(-1, n_kv_heads, head_dim)
    q, k = apply_rotary(q, k, positions, cos_sin, rotary_dim, interleaved)
    out = attention_layer_forward(q, k, value, key_cache, value_cache, new_cache_slots, q_cu_seq_lens, kv_cu_seq_lens, block_tables,
                                  cu_blocks_lens, n_qo_heads, n_kv_heads, head_dim)
    return out, q, k


# ---------------------------------------------------------------------------------------------------------------
# Attention
# ---------------------------------------------------------------------------------------------------------------
def paged_attention_fp32(query: Tensor, key_cache: Tensor, value_cache: Tensor, q_cu_seq_lens, kv_cu_seq_lens,
                         block_tables: Tensor, cu_blocks_lens, n_qo_heads: int, n_kv_heads: int, head_dim: int) -> Tensor:
    """fp32 result of TorchCausalGroupedQueryPageAttentionHandler.forward BEFORE the final cast
    (hydrainfer/layer/causal_attention.py:307-374).  query [T, Hq, d]; caches [NB, bs, Hkv, d]; returns fp32 [T, Hq*d].

    Per sequence (:314): gather pages by block table and flatten to [n_pages*bs, Hkv, d] (:315-317), keep the first
    L rows (:318-319), cast q/k/v to fp32 (:318-320), repeat KV heads to Hq (:324-326), scores = q.k * 1/sqrt(d)
    (:331-335), mask where (j - i) > (L - q) (:339-342), softmax over keys (:365), o = scores.v (:366)."""
    q_cu = [int(x) for x in q_cu_seq_lens]
    kv_cu = [int(x) for x in kv_cu_seq_lens]
    cu_blocks = [int(x) for x in cu_blocks_lens]
    block_tables = block_tables.long()
    group = n_qo_heads // n_kv_heads
    sm_scale = 1.0 / math.sqrt(head_dim)
    outputs = []
    for b in range(len(q_cu) - 1):
        table = block_tables[cu_blocks[b]: cu_blocks[b + 1]]
        kv_len = kv_cu[b + 1] - kv_cu[b]
        k = key_cache[table].reshape(-1, n_kv_heads, head_dim)[:kv_len].to(torch.float)
        v = value_cache[table].reshape(-1, n_kv_heads, head_dim)[:kv_len].to(torch.float)
        q = query[q_cu[b]: q_cu[b + 1]].to(torch.float)
        q_len = q.shape[0]
        k = k.repeat_interleave(repeats=group, dim=1)
        v = v.repeat_interleave(repeats=group, dim=1)
        scores = torch.einsum("qhd,khd->hqk", q, k) * sm_scale
        j = torch.arange(kv_len)[None, None, :]
        i = torch.arange(q_len)[None, :, None]
        scores = scores.masked_fill((j - i) > (kv_len - q_len), float("-inf"))
        scores = torch.softmax(scores, dim=-1)
        outputs.append(torch.einsum("hqk,khd->qhd", scores, v))
    return torch.cat(outputs, dim=0).reshape(-1, n_qo_heads * head_dim)


def paged_attention_options_fp32(query: Tensor, key_cache: Tensor, value_cache: Tensor, q_cu_seq_lens, kv_cu_seq_lens,
                                 block_tables: Tensor, cu_blocks_lens, n_qo_heads: int, n_kv_heads: int, head_dim: int,
                                 softmax_scale: float, softcap: float = 0.0, window_size_left: int = -1, window_size_right: int = 0,
                                 alibi_slopes: Tensor | None = None) -> Tensor:
    """fp32 result of the reference's fused backend `mha_varlen_fwd` (paged form) WITH its score options - what no Python handler of the
    reference computes, restated from its vendored FlashAttention-2 (csrc/kernel/flash_attn):

      * softcap > 0: the kernel turns q.k into tanh(q.k * softmax_scale / softcap) (flash_api.cpp:93-96 sets params.softcap =
        softmax_scale / softcap, src/flash_fwd_kernel.h:282-284 applies it) and then scales by softcap instead of softmax_scale
        (flash_api.cpp:95-96), i.e. score = softcap * tanh(q.k * softmax_scale / softcap);
      * alibi: score -= slope[b, h] * |i_abs - j| with i_abs = i + kv_len - q_len (src/mask.h:183-186; the slope is divided by the
        softmax scale there because it is added before the scaling, flash_fwd_kernel.h:244), after the softcap (:283-288);
      * windows: key j is visible iff i_abs - window_left <= j <= i_abs + window_right (src/mask.h:54-62 with the ACTUAL sequence
        lengths, flash_fwd_kernel.h:245); a negative bound means unlimited (flash_api.cpp:104-111: (-1, 0) is causal, (-1, -1) no
        mask).  A row without a visible key yields zeros.

    PARITY: pinned on the GPU against the reference's compiled FlashAttention-2 (oracle/_ref, tests/test_gpu_reference_native.py)."""
    q_cu = [int(x) for x in q_cu_seq_lens]
    kv_cu = [int(x) for x in kv_cu_seq_lens]
    cu_blocks = [int(x) for x in cu_blocks_lens]
    block_tables = block_tables.long()
    group = n_qo_heads // n_kv_heads
    outputs = []
    for b in range(len(q_cu) - 1):
        table = block_tables[cu_blocks[b]: cu_blocks[b + 1]]
        kv_len = kv_cu[b + 1] - kv_cu[b]
        k = key_cache[table].reshape(-1, n_kv_heads, head_dim)[:kv_len].to(torch.float).repeat_interleave(repeats=group, dim=1)
        v = value_cache[table].reshape(-1, n_kv_heads, head_dim)[:kv_len].to(torch.float).repeat_interleave(repeats=group, dim=1)
        q = query[q_cu[b]: q_cu[b + 1]].to(torch.float)
        q_len = q.shape[0]
        scores = torch.einsum("qhd,khd->hqk", q, k) * softmax_scale
        if softcap > 0:
            scores = softcap * torch.tanh(scores / softcap)
        j = torch.arange(kv_len)[None, None, :]
        i_abs = torch.arange(q_len)[None, :, None] + (kv_len - q_len)
        if alibi_slopes is not None:
            slopes = alibi_slopes.to(torch.float).cpu()
            slopes = slopes[b] if slopes.dim() == 2 else slopes
            scores = scores - slopes[:, None, None] * (i_abs - j).abs().to(torch.float)
        hidden = torch.zeros_like(scores, dtype=torch.bool)
        if window_size_right >= 0:
            hidden |= j > i_abs + window_size_right
        if window_size_left >= 0:
            hidden |= j < i_abs - window_size_left
        scores = scores.masked_fill(hidden, float("-inf"))
        probs = torch.softmax(scores, dim=-1)
        probs = torch.where(hidden.all(dim=-1, keepdim=True), torch.zeros_like(probs), probs)  # rows without a visible key
        outputs.append(torch.einsum("hqk,khd->qhd", probs, v))
    return torch.cat(outputs, dim=0).reshape(-1, n_qo_heads * head_dim)


def paged_attention(query: Tensor, key_cache: Tensor, value_cache: Tensor, q_cu_seq_lens, kv_cu_seq_lens,
                    block_tables: Tensor, cu_blocks_lens, n_qo_heads: int, n_kv_heads: int, head_dim: int) -> Tensor:
    """The handler's return value: the fp32 result rounded to the query dtype (causal_attention.py:370-372)."""
    return paged_attention_fp32(query, key_cache, value_cache, q_cu_seq_lens, kv_cu_seq_lens, block_tables,
                                cu_blocks_lens, n_qo_heads, n_kv_heads, head_dim).to(query.dtype)


def attention_layer_forward(query: Tensor, key: Tensor, value: Tensor, key_cache: Tensor, value_cache: Tensor,
                            new_cache_slots: Tensor, q_cu_seq_lens, kv_cu_seq_lens, block_tables: Tensor, cu_blocks_lens,
                            n_qo_heads: int, n_kv_heads: int, head_dim: int) -> Tensor:
    """CausalGroupedQueryPageAttention.forward (causal_attention.py:394-406): view q/k/v to [T, H, d], append the new
    K/V rows to the cache, then attend over old + new tokens."""
    n_tokens = query.shape[0]
    q = query.view(n_tokens, n_qo_heads, head_dim)
    k = key.view(n_tokens, n_kv_heads, head_dim)
    v = value.view(n_tokens, n_kv_heads, head_dim)
    set_kv_cache(new_cache_slots, k, v, key_cache, value_cache)
    return paged_attention(q, key_cache, value_cache, q_cu_seq_lens, kv_cu_seq_lens, block_tables, cu_blocks_lens,
                           n_qo_heads, n_kv_heads, head_dim)


# ---------------------------------------------------------------------------------------------------------------
# Allocator contract and per-step metadata
# ---------------------------------------------------------------------------------------------------------------
class BlockAllocator:
    """hydrainfer/memory/block_allocator.py:11-39: free list [n-1 .. 0]; allocate(n) takes the LAST n entries in list
    order (:25-32, at most what is left, [] for 0); free appends (:34-36)."""

    def __init__(self, total_blocks: int):
        self.total_blocks = total_blocks
        self.free_blocks = [b for b in reversed(range(total_blocks))]

    def allocate(self, n_blocks: int) -> list[int]:
        if n_blocks == 0:
            return []
        n_blocks = min(n_blocks, len(self.free_blocks))
        blocks, self.free_blocks = self.free_blocks[-n_blocks:], self.free_blocks[:-n_blocks]
        return blocks

    def free(self, blocks: list[int]) -> None:
        self.free_blocks += blocks


def v2p(block_table: list[int], block_size: int, virtual_cache_ids: list[int]) -> list[int]:
    """hydrainfer/memory/token_cache_manger.py:126-133: slot = block_table[v // bs] * bs + v % bs."""
    return [block_table[v // block_size] * block_size + v % block_size for v in virtual_cache_ids]


@dataclass
class Metadata:
    """The lists AttentionParametersBuilder accumulates (causal_attention.py:135-157) before they become int32 tensors."""
    q_cu_seq_lens: list[int] = field(default_factory=lambda: [0])
    kv_cu_seq_lens: list[int] = field(default_factory=lambda: [0])
    paged_kv_last_page_len: list[int] = field(default_factory=list)
    new_cache_slots: list[int] = field(default_factory=list)
    block_tables: list[int] = field(default_factory=list)
    cu_blocks_lens: list[int] = field(default_factory=lambda: [0])
    num_sequences: int = 0
    all_sequences_decode: bool = True
    q_max_seq_len: int = 0
    kv_max_seq_len: int = 0


def build_metadata(requests: list[tuple[int, int, list[int], list[int]]], block_size: int) -> Metadata:
    """requests: (q_seq_len, kv_seq_len, new_cache_slots, block_table) per sequence, in batch order
    (AttentionParametersBuilder.add_request, causal_attention.py:147-157)."""
    m = Metadata()
    for q_len, kv_len, slots, table in requests:
        m.q_cu_seq_lens.append(m.q_cu_seq_lens[-1] + q_len)
        m.kv_cu_seq_lens.append(m.kv_cu_seq_lens[-1] + kv_len)
        m.paged_kv_last_page_len.append((kv_len + block_size - 1) % block_size + 1)
        m.new_cache_slots += slots
        m.block_tables += table
        m.cu_blocks_lens.append(m.cu_blocks_lens[-1] + len(table))
        m.num_sequences += 1
        m.all_sequences_decode = m.all_sequences_decode and q_len == 1
        m.q_max_seq_len = max(m.q_max_seq_len, q_len)
        m.kv_max_seq_len = max(m.kv_max_seq_len, kv_len)
    return m


# ---------------------------------------------------------------------------------------------------------------
# Migration (parity unpinned: restated from source only)
# ---------------------------------------------------------------------------------------------------------------
def migrate_blocks(src_block_table: list[int], dst_block_table: list[int], src_cache: Tensor, dst_cache: Tensor) -> None:
    """csrc/data_transfer/block_migration.cpp:222-244: for every (layer, token/kv plane, i) copy the contiguous run
    src[layer, plane, src_bt[i], :, :, :] -> dst[layer, plane, dst_bt[i], :, :, :] (INDEX_6D, :26-27).
    Pools are 6-D (n_layers, n_tokens, n_blocks, block_size, n_heads, head_size) and may differ in n_blocks only."""
    assert len(src_block_table) == len(dst_block_table)
    n_layers, n_tokens = dst_cache.shape[0], dst_cache.shape[1]
    for layer_id in range(n_layers):
        for token_id in range(n_tokens):
            for s, d in zip(src_block_table, dst_block_table):
                dst_cache[layer_id, token_id, d].copy_(src_cache[layer_id, token_id, s])
