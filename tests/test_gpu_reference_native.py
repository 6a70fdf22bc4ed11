"""GPU: hydrainfer_b200's kernels against the REFERENCE'S OWN compiled CUDA, bit for bit.

oracle/_ref holds the reference's csrc compiled unmodified for sm_100 (oracle/build_ref.py): `set_kv_cache`
(csrc/kernel/kv_cache_kernels/kv_cache_kernels.cu:17-95), `set_image_cache` (csrc/kernel/cache_kernels/cache_kernels.cu:17-83),
`apply_rotary_pos_emb` (csrc/kernel/position_embedding/rope.cu:33-117) and `migrate_blocks`
(csrc/data_transfer/block_migration.cpp:194-245, index math `INDEX_6D` :26-27).  These are the only definitions of the expected
bytes the reference has for the CUDA side of the path (it stores no fixtures and has no migration test), so this file is what
pins a4 / a5 / a10 of SURVEY §8 to the reference rather than to a restatement of it.
"""
import os
import socket

import pytest
import torch

from oracle import paged_kv_oracle as oracle
from oracle import reference_tree as ref_tree

pytestmark = [pytest.mark.gpu]
DEV = "cuda:0"


def _need(name):
    if not ref_tree.available(name):
        pytest.skip(f"oracle/_ref/{name} did not travel to this box (git-ignored build product of `python oracle/build_ref.py`): the comparison with the reference's compiled CUDA cannot run")
    return ref_tree.load_native(name)


# ---- set_kv_cache ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("geom", [
    # (n_blocks, block_size, n_kv_heads, head_dim, n_tokens)   reference grid tests/memory/test_kv_cache.py:6-13 + the BASELINE shapes
    (100, 4, 8, 64, 15), (100, 8, 2, 256, 100), (100, 16, 1, 128, 1), (100, 16, 4, 128, 64),
    (300, 16, 32, 128, 64),     # config 2: LLaVA-1.5-7B decode step
    (600, 16, 4, 128, 2096),    # config 3: Qwen2-VL-7B mixed batch
    (600, 16, 8, 128, 256),     # config 4: Qwen2-VL-72B decode step
])
def test_set_kv_cache_matches_reference_cuda(geom, dtype):
    ref = _need("kv_cache_kernels")
    from hydrainfer_b200._C.kernel.kv_cache_kernels import set_kv_cache
    nb, bs, h, d, t = geom
    g = torch.Generator().manual_seed(nb + t)
    kc = torch.randn(nb, bs, h, d, generator=g).to(dtype).to(DEV)
    vc = torch.randn(nb, bs, h, d, generator=g).to(dtype).to(DEV)
    k = torch.randn(t, h, d, generator=g).to(dtype).to(DEV)
    v = torch.randn(t, h, d, generator=g).to(dtype).to(DEV)
    slots = torch.randperm(nb * bs, generator=g)[:t].to(torch.int32).to(DEV)
    kc_ref, vc_ref = kc.clone(), vc.clone()
    ref.set_kv_cache(slots, k, v, kc_ref, vc_ref)
    set_kv_cache(slots, k, v, kc, vc)
    torch.cuda.synchronize()
    assert torch.equal(kc, kc_ref) and torch.equal(vc, vc_ref)
    # and the oracle restatement agrees with the reference's CUDA kernel (pins the oracle from the device side too)
    kc_o, vc_o = torch.zeros(nb, bs, h, d, dtype=dtype), torch.zeros(nb, bs, h, d, dtype=dtype)
    kc_z, vc_z = torch.zeros_like(kc), torch.zeros_like(vc)
    ref.set_kv_cache(slots, k, v, kc_z, vc_z)
    oracle.set_kv_cache(slots.cpu(), k.cpu(), v.cpu(), kc_o, vc_o)
    assert torch.equal(kc_z.cpu(), kc_o) and torch.equal(vc_z.cpu(), vc_o)


def test_set_kv_cache_fused_qkv_rows_match_reference_cuda():
    """keys / values as column slices of a fused qkv projection (model_forward.py:69-73): stride(-3) > row width (:75-76)."""
    ref = _need("kv_cache_kernels")
    from hydrainfer_b200._C.kernel.kv_cache_kernels import set_kv_cache
    hq, hkv, d, bs, nb, t = 28, 4, 128, 16, 40, 37
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(t, (hq + 2 * hkv) * d, generator=g).to(torch.bfloat16).to(DEV)
    k = qkv[:, hq * d:(hq + hkv) * d].view(t, hkv, d)
    v = qkv[:, (hq + hkv) * d:].view(t, hkv, d)
    kc = torch.randn(nb, bs, hkv, d, generator=g).to(torch.bfloat16).to(DEV)
    vc = torch.randn(nb, bs, hkv, d, generator=g).to(torch.bfloat16).to(DEV)
    slots = torch.randperm(nb * bs, generator=g)[:t].to(torch.int32).to(DEV)
    kc_ref, vc_ref = kc.clone(), vc.clone()
    ref.set_kv_cache(slots, k, v, kc_ref, vc_ref)
    set_kv_cache(slots, k, v, kc, vc)
    torch.cuda.synchronize()
    assert torch.equal(kc, kc_ref) and torch.equal(vc, vc_ref)


# ---- set_image_cache ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("geom", [(3, 576, 32, 128, 576 * 2), (5, 64, 2, 64, 77), (2, 576, 28, 128, 1)])
def test_set_image_cache_matches_reference_cuda(geom, dtype):
    ref = _need("cache_kernels")
    from hydrainfer_b200._C.kernel.cache_kernels import set_image_cache
    nb, bs, h, d, t = geom
    g = torch.Generator().manual_seed(t)
    cache = torch.randn(nb, bs, h, d, generator=g).to(dtype).to(DEV)
    tokens = torch.randn(t, h, d, generator=g).to(dtype).to(DEV)
    slots = torch.randperm(nb * bs, generator=g)[:t].to(torch.int32).to(DEV)
    cache_ref = cache.clone()
    ref.set_image_cache(slots, tokens, cache_ref)
    set_image_cache(slots, tokens, cache)
    torch.cuda.synchronize()
    assert torch.equal(cache, cache_ref)


# ---- apply_rotary_pos_emb ----------------------------------------------------------------------------------------------
def _rope_inputs(geom, dtype):
    t, hq, hkv, d, rd = geom
    g = torch.Generator().manual_seed(t * 7 + rd)
    q = torch.randn(t, hq, d, generator=g).to(dtype).to(DEV)
    k = torch.randn(t, hkv, d, generator=g).to(dtype).to(DEV)
    pos = torch.randint(0, 4096, (t,), generator=g, dtype=torch.int32).to(DEV)
    inv = 1.0 / torch.pow(torch.tensor(10000.0), torch.arange(0, rd, 2, dtype=torch.float) / rd)
    freqs = torch.einsum("i,j->ij", torch.arange(4096, dtype=torch.float), inv)
    cos_sin = torch.cat([freqs.cos()[:, None, :], freqs.sin()[:, None, :]], dim=1).to(dtype).to(DEV)  # rotary_embedding.py:113-116
    return q, k, pos, cos_sin


@pytest.mark.parametrize("interleaved", [False, True])
@pytest.mark.parametrize("geom", [(64, 32, 32, 128, 128), (37, 28, 4, 128, 128), (9, 8, 2, 64, 32)])
def test_rotary_matches_reference_cuda_fp16(geom, interleaved):
    """The reference kernel computes `x*c - y*s` in native `half` (dispatch.h:21-24, rope.cu:27-28), which nvcc contracts into a
    half FMA: ONE rounding where the reference's torch handler (rotary_embedding.py:47-99; the definition the oracle, the golden
    fixtures and hi_rope_append follow bit for bit) has two.  So the two reference paths themselves differ by an ulp; against the
    compiled kernel the bar is the rounding of the two products (2^-10 of their magnitudes), with most elements identical."""
    ref = _need("position_embedding")
    from hydrainfer_b200._C.kernel.position_embedding import apply_rotary_pos_emb
    q, k, pos, cos_sin = _rope_inputs(geom, torch.float16)
    rd = geom[4]
    q_in, k_in = q.clone(), k.clone()
    q_ref, k_ref = q.clone(), k.clone()
    ref.apply_rotary_pos_emb(q_ref, k_ref, pos, cos_sin, rd, interleaved)
    apply_rotary_pos_emb(q, k, pos, cos_sin, rd, interleaved)
    torch.cuda.synchronize()
    n = rd // 2
    for ours, theirs, x in ((q, q_ref, q_in.float()), (k, k_ref, k_in.float())):
        rot = x[..., :rd]
        partner = rot.reshape(*rot.shape[:-1], n, 2).flip(-1).reshape(rot.shape) if interleaved else torch.cat([rot[..., n:], rot[..., :n]], dim=-1)
        # each product is at most |x| or |y| (|cos|, |sin| <= 1) and is rounded to 11 bits in one path and not in the other; the sum once more
        bound = 2.0 ** -10 * (rot.abs() + partner.abs()) + 2.0 ** -24
        diff = (ours.float() - theirs.float())[..., :rd].abs()
        assert bool((diff <= bound).all()), f"max excess {(diff - bound).max().item():.3e} over the two-rounding bound"
        assert (ours == theirs).float().mean().item() > 0.5
        assert torch.equal(ours[..., rd:], theirs[..., rd:])  # the pass-through dims are copies


def test_reference_cuda_rotary_has_no_bf16():
    """dispatch.h:12-29 knows Float and Half only: the reference's compiled rotary kernel rejects bf16 (its handler chain then has
    no fused path for bf16 models); hi_rope_append covers bf16, pinned by the torch-handler fixtures of tests/test_rope.py."""
    ref = _need("position_embedding")
    q, k, pos, cos_sin = _rope_inputs((9, 8, 2, 64, 32), torch.bfloat16)
    with pytest.raises(RuntimeError, match="dispatch"):
        ref.apply_rotary_pos_emb(q, k, pos, cos_sin, 32, False)


def test_rotary_matches_reference_cuda_fp32():
    """fp32: nvcc contracts the reference's `x*c - y*s` into an FMA (rope.cu:27-28, default -fmad), the reference's torch handler
    (rotary_embedding.py:47-99, which the oracle and hi_rope_append follow bit for bit) does not: one rounding of difference."""
    ref = _need("position_embedding")
    from hydrainfer_b200._C.kernel.position_embedding import apply_rotary_pos_emb
    t, hq, hkv, d, rd = 16, 8, 2, 128, 128
    g = torch.Generator().manual_seed(1)
    q = torch.randn(t, hq, d, generator=g).to(DEV)
    k = torch.randn(t, hkv, d, generator=g).to(DEV)
    pos = torch.randint(0, 2048, (t,), generator=g, dtype=torch.int32).to(DEV)
    inv = 1.0 / torch.pow(torch.tensor(100000.0), torch.arange(0, rd, 2, dtype=torch.float) / rd)
    freqs = torch.einsum("i,j->ij", torch.arange(2048, dtype=torch.float), inv)
    cos_sin = torch.cat([freqs.cos()[:, None, :], freqs.sin()[:, None, :]], dim=1).to(DEV)
    q_ref, k_ref = q.clone(), k.clone()
    ref.apply_rotary_pos_emb(q_ref, k_ref, pos, cos_sin, rd, False)
    apply_rotary_pos_emb(q, k, pos, cos_sin, rd, False)
    torch.cuda.synchronize()
    torch.testing.assert_close(q, q_ref, atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(k, k_ref, atol=1e-6, rtol=1e-6)


# ---- migrate_blocks ----------------------------------------------------------------------------------------------------
# The reference opens the IPC handle on every call (block_migration.cpp:213-215) and cudaIpcOpenMemHandle cannot open a handle in
# the process that exported it, so its migrate_blocks only runs with the source pool owned by ANOTHER process: the parent owns
# the source pool, the child runs the reference's migrate_blocks and ours on the same inputs and compares the destination pools.
MIGRATION_GEOMETRIES = [
    # name, (n_layers, n_tokens, block_size, n_heads, head_size), src_blocks, dst_blocks, n_move        SURVEY §8d config 5
    ("llava7b", (32, 2, 16, 32, 128), 40, 36, 24),        # 128 KiB runs, 8 MiB per block
    ("qwen2vl7b", (28, 2, 16, 4, 128), 300, 280, 256),    # 16 KiB runs, 896 KiB per block
    ("qwen2vl72b_8l", (8, 2, 16, 8, 128), 200, 180, 64),  # 32 KiB runs (8 of the 80 layers)
    ("image_pool", (1, 1, 576, 32, 128), 12, 10, 7),      # one 4.5 MiB run per block
    ("one_block", (28, 2, 16, 4, 128), 300, 280, 1),
]


def _migration_child(handle_port, result_path):
    import pickle
    from multiprocessing.connection import Client
    conn = Client(("127.0.0.1", handle_port))
    try:
        ref_bm = ref_tree.load_native("block_migration")
        from hydrainfer_b200._C.data_transfer import block_migration as bm
        results = {}
        while True:
            msg = conn.recv()
            if msg is None:
                break
            name, shape, nb_src, nb_dst, src_bt, dst_bt, handle_ours, handle_ref, dtype_name = msg
            dtype = getattr(torch, dtype_name)
            L, T, bs, H, d = shape
            g = torch.Generator().manual_seed(17)
            dst0 = torch.randn(L, T, nb_dst, bs, H, d, generator=g).to(dtype)
            dst_ref, dst_ours = dst0.to(DEV), dst0.to(DEV)
            # handles cross over: the reference opens the handle OUR get_ipc_mem_handle exported, we open the reference's
            ref_bm.migrate_blocks(src_bt, dst_bt, handle_ours, dst_ref, nb_src)
            bm.migrate_blocks(src_bt, dst_bt, handle_ref, dst_ours, nb_src)
            torch.cuda.synchronize()
            same = torch.equal(dst_ref, dst_ours)
            moved = bool((dst_ref.cpu() != dst0).any())
            untouched = [b for b in range(nb_dst) if b not in dst_bt]
            untouched_ok = torch.equal(dst_ours[:, :, untouched].cpu(), dst0[:, :, untouched])
            results[name] = (same, moved, untouched_ok, dst_ref[:, :, dst_bt].cpu().view(torch.uint8).sum(dtype=torch.int64).item())
            conn.send("done")
        with open(result_path, "wb") as f:
            pickle.dump(results, f)
    finally:
        conn.close()


def test_migrate_blocks_matches_reference_cuda_across_processes(tmp_path):
    import pickle
    import torch.multiprocessing as mp
    from multiprocessing.connection import Listener
    _need("block_migration")
    from hydrainfer_b200._C.data_transfer import block_migration as bm
    ref_bm = ref_tree.load_native("block_migration")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    listener = Listener(("127.0.0.1", port))
    result_path = str(tmp_path / "result.pkl")
    ctx = mp.get_context("spawn")
    child = ctx.Process(target=_migration_child, args=(port, result_path))
    child.start()
    conn = listener.accept()
    expected = {}
    try:
        for name, shape, nb_src, nb_dst, n_move in MIGRATION_GEOMETRIES:
            L, T, bs, H, d = shape
            g = torch.Generator().manual_seed(n_move)
            torch.cuda.empty_cache()  # a fresh cudaMalloc per pool: offset 0 inside its allocation, as the reference assumes
            src = torch.randn(L, T, nb_src, bs, H, d, generator=g).to(torch.bfloat16).to(DEV)
            src_bt = torch.randperm(nb_src, generator=g)[:n_move].tolist()
            dst_bt = torch.randperm(nb_dst, generator=g)[:n_move].tolist()
            handle = bm.get_ipc_mem_handle(src)
            # same format as the reference's get_ipc_mem_handle: 64 ints, one per byte of cudaIpcMemHandle_t (block_migration.cpp:34-40, 55-59)
            assert len(handle) == 64, "the source pool must sit at the base of its allocation for the reference (it assumes offset 0)"
            handle_ref = [int(b) for b in ref_bm.get_ipc_mem_handle(src)]
            assert len(handle_ref) == 64 and all(0 <= b < 256 for b in handle + handle_ref)
            torch.cuda.synchronize()
            conn.send((name, shape, nb_src, nb_dst, src_bt, dst_bt, handle, handle_ref, "bfloat16"))
            assert conn.recv() == "done"
            expected[name] = src[:, :, src_bt].cpu().view(torch.uint8).sum(dtype=torch.int64).item()
            del src
        conn.send(None)
    finally:
        child.join(timeout=300)
        conn.close()
        listener.close()
    assert child.exitcode == 0, "the child process (reference migrate_blocks + ours) failed"
    with open(result_path, "rb") as f:
        results = pickle.load(f)
    for name, *_ in MIGRATION_GEOMETRIES:
        same, moved, untouched_ok, checksum = results[name]
        assert same, f"{name}: destination pool differs from the reference's migrate_blocks"
        assert moved and untouched_ok, name
        assert checksum == expected[name], f"{name}: moved blocks do not carry the source bytes"


def test_oracle_migrate_blocks_agrees_with_reference_cuda(tmp_path):
    """Pins oracle.migrate_blocks (block_migration.cpp:222-244 restated) to the reference's compiled code: the child runs the
    reference on a pool exported by this process, this process runs the oracle on CPU copies."""
    import torch.multiprocessing as mp
    _need("block_migration")
    from hydrainfer_b200._C.data_transfer import block_migration as bm
    shape, nb_src, nb_dst = (3, 2, 8, 2, 64), 2100, 1950   # 25 MiB straight from cudaMalloc (empty_cache below): its own allocation, offset 0
    L, T, bs, H, d = shape
    g = torch.Generator().manual_seed(23)
    src = torch.randn(L, T, nb_src, bs, H, d, generator=g).to(torch.float16)
    dst0 = torch.randn(L, T, nb_dst, bs, H, d, generator=g).to(torch.float16)
    src_bt = torch.randperm(nb_src, generator=g)[:129].tolist()
    dst_bt = torch.randperm(nb_dst, generator=g)[:129].tolist()
    torch.cuda.empty_cache()  # the pool must not be carved out of a cached segment: the reference assumes offset 0 inside the allocation
    src_d = src.to(DEV)
    handle = bm.get_ipc_mem_handle(src_d)
    assert len(handle) == 64
    torch.cuda.synchronize()
    out = str(tmp_path / "dst.pt")
    ctx = mp.get_context("spawn")
    child = ctx.Process(target=_oracle_pin_child, args=(handle, shape, nb_src, nb_dst, src_bt, dst_bt, out))
    child.start()
    child.join(timeout=300)
    assert child.exitcode == 0
    got = torch.load(out)
    want = dst0.clone()
    oracle.migrate_blocks(src_bt, dst_bt, src, want)
    assert torch.equal(got, want)


def _oracle_pin_child(handle, shape, nb_src, nb_dst, src_bt, dst_bt, out):
    ref_bm = ref_tree.load_native("block_migration")
    L, T, bs, H, d = shape
    g = torch.Generator().manual_seed(23)
    torch.randn(L, T, nb_src, bs, H, d, generator=g)  # advance the generator past the source pool
    dst = torch.randn(L, T, nb_dst, bs, H, d, generator=g).to(torch.float16).to(DEV)
    ref_bm.migrate_blocks(src_bt, dst_bt, handle, dst, nb_src)
    torch.cuda.synchronize()
    torch.save(dst.cpu(), out)


# ---- mha_varlen_fwd with its score options: ours against the reference's compiled FlashAttention-2 -----------------------------------
@pytest.mark.parametrize("case", [
    # (softcap, window_left, window_right, alibi)
    (0.0, -1, 0, False), (30.0, -1, 0, False), (0.0, 48, 0, False), (0.0, 32, 16, False), (0.0, -1, -1, False), (0.0, -1, 0, True), (20.0, 64, 0, True),
], ids=lambda c: f"cap{c[0]}_w{c[1]}_{c[2]}_alibi{int(c[3])}")
@pytest.mark.parametrize("geom", [(torch.bfloat16, 8, 2, 128), (torch.float16, 4, 4, 64), (torch.bfloat16, 4, 1, 256)], ids=["bf16_d128", "f16_d64", "bf16_d256"])
def test_paged_mha_varlen_fwd_options_match_reference_fa2(geom, case):
    """The reference's own `mha_varlen_fwd` (csrc/kernel/flash_attn/flash_api.cpp:216-355, compiled unmodified with its vendored
    FlashAttention-2) and ours on the same paged inputs, including the options the paged layer never passes (softcap, local windows,
    alibi: flash_api.cpp:93-111, 197-213)."""
    import math
    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd as ours
    from hydrainfer_b200.workloads import make_batch
    theirs = _need("flash_attn").mha_varlen_fwd
    dtype, hq, hkv, d = geom
    softcap, wl, wr, alibi = case
    seq_lens = [(1, 300), (37, 37), (5, 130), (64, 200), (9, 9)]
    batch = make_batch(seq_lens, hq, hkv, d, 16, dtype=dtype, device=DEV, seed=11)
    t = batch.n_tokens
    q3 = batch.query.view(t, hq, d).contiguous()
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    meta = (i32(batch.q_cu_seq_lens), i32(batch.kv_cu_seq_lens), i32(batch.block_tables), i32(batch.cu_blocks_lens))
    slopes = (2.0 ** -torch.arange(1, hq + 1, dtype=torch.float32)).to(DEV) if alibi else None
    scale = 1.0 / math.sqrt(d)
    a, b = torch.empty_like(q3), torch.empty_like(q3)
    ours(a, q3, batch.key_cache, batch.value_cache, *meta, slopes, batch.q_max, batch.kv_max, scale, softcap, wl, wr, 0)
    theirs(b, q3, batch.key_cache, batch.value_cache, *meta, slopes, batch.q_max, batch.kv_max, scale, softcap, wl, wr, 0)
    torch.cuda.synchronize()
    err = (a.float() - b.float()).abs()
    # two 16-bit implementations of the same sum: the reference's own cross-backend tolerance (tests/layer/test_attention.py:95-96)
    assert bool((err <= 1e-2 + 1e-2 * b.float().abs()).all()), float(err.max())
