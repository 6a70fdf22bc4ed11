"""Rotary embedding + fused KV append.

CPU: the oracle reproduces the reference's frozen outputs (tests/golden/rope_*.npz and ropeattn_*.npz, made by
oracle/make_golden.py from TorchRotaryEmbeddingHandler and ROPECausalGroupedQueryPageAttention).
GPU: hi_rope_append through the `position_embedding` shim and the layer/model mirrors, BIT-EXACT against the oracle and
the fixtures (the kernel reproduces the reference's per-operation rounding); attention output of the ROPE module within
the north-star tolerance of the fp32 recompute.  Grid from the reference's tests/layer/test_rotary_embedding.py:66-75."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, _TORCH_DTYPES, _from_np
from oracle import paged_kv_oracle as oracle

DEV = "cuda:0"
ROPE_FIXTURES = sorted(GOLDEN.glob("rope_*.npz"))


def _inv_freq(rotary_dim: int, theta: float) -> torch.Tensor:
    return 1. / torch.pow(torch.tensor(theta), torch.arange(0, rotary_dim, 2, dtype=torch.float) / rotary_dim)


class GoldenRope:
    def __init__(self, path):
        z = np.load(path)
        self.dtype = _TORCH_DTYPES[str(z["dtype"])]
        self.table_dtype = _TORCH_DTYPES[str(z["table_dtype"])]
        self.t, self.hq, self.hkv, self.d, self.rd, self.max_pos = (int(v) for v in z["geometry"])
        self.theta = float(z["theta"])
        self.interleaved = bool(int(z["interleaved"]))
        for name in ("query", "key", "ref_query", "ref_key"):
            setattr(self, name, _from_np(z[name], self.dtype))
        self.positions = torch.from_numpy(z["positions"])
        self.table_checksum = float(z["table_checksum"][0])
        # The table is rebuilt on this host, then the rows the case reads are replaced by the reference's own (frozen in the
        # fixture): libm's cosf/sinf may differ by an ulp between host CPUs, and a bit-exact GPU check must not depend on that.
        self.local_table = oracle.rotary_cos_sin_table(self.rd, self.max_pos, _inv_freq(self.rd, self.theta)).to(self.table_dtype)
        self.table_rows = _from_np(z["table_rows"], self.table_dtype)
        self.table = self.local_table.clone()
        self.table[self.positions.long()] = self.table_rows


@pytest.fixture(params=ROPE_FIXTURES, ids=lambda p: p.stem[len("rope_"):])
def golden_rope(request) -> GoldenRope:
    return GoldenRope(request.param)


# ------------------------------------------------------------------------------------------------ CPU: oracle pin
def test_oracle_rotary_matches_reference(golden_rope):
    g = golden_rope
    # the oracle's table builder against the reference's rows: identical on the host that made the fixture, within an ulp of
    # libm elsewhere
    torch.testing.assert_close(g.local_table[g.positions.long()].float(), g.table_rows.float(), rtol=1e-2 if g.table_dtype != torch.float32 else 1e-5, atol=1e-6)
    q, k = oracle.apply_rotary(g.query, g.key, g.positions, g.table, g.rd, g.interleaved)
    assert torch.equal(q, g.ref_query) and torch.equal(k, g.ref_key)


def _load_ropeattn():
    z = np.load(GOLDEN / "ropeattn_qwen_bf16.npz")
    dtype = _TORCH_DTYPES[str(z["dtype"])]
    hq, hkv, d, bs, n_blocks, max_pos = (int(v) for v in z["geometry"])
    seq_lens = [tuple(int(v) for v in row) for row in z["seq_lens"]]
    t = {name: _from_np(z[name], dtype) for name in ("qkv", "key_cache", "value_cache", "ref_key_cache_owned", "ref_value_cache_owned", "ref_out", "ref_query_rot")}
    t["ref_fp32"] = torch.from_numpy(z["ref_fp32"])
    t["table"] = _from_np(z["table"], dtype)  # the reference's cos/sin cache (see GoldenRope)
    t["positions"] = torch.from_numpy(z["positions"])
    t["owned_blocks"] = torch.from_numpy(z["owned_blocks"])
    for name in ("new_cache_slots", "block_tables", "q_cu_seq_lens", "cu_blocks_lens"):
        t[name] = [int(v) for v in z[name]]
    return dict(dtype=dtype, hq=hq, hkv=hkv, d=d, bs=bs, n_blocks=n_blocks, max_pos=max_pos, seq_lens=seq_lens, theta=float(z["theta"]), seed=int(z["seed"]), **t)


def _ropeattn_requests(g):
    out = []
    for i, (q, kv) in enumerate(g["seq_lens"]):
        out.append((q, kv, g["new_cache_slots"][g["q_cu_seq_lens"][i]: g["q_cu_seq_lens"][i + 1]], g["block_tables"][g["cu_blocks_lens"][i]: g["cu_blocks_lens"][i + 1]]))
    return out


def test_oracle_rope_attention_matches_reference():
    g = _load_ropeattn()
    requests = _ropeattn_requests(g)
    meta = oracle.build_metadata(requests, g["bs"])
    hq, hkv, d = g["hq"], g["hkv"], g["d"]
    qkv = g["qkv"]
    query, key, value = qkv[:, :hq * d], qkv[:, hq * d:(hq + hkv) * d], qkv[:, (hq + hkv) * d:]
    table = g["table"]
    kc, vc = g["key_cache"].clone(), g["value_cache"].clone()
    out, q_rot, _ = oracle.rope_attention_layer_forward(
        query, key, value, g["positions"], table, d, False, kc, vc, torch.tensor(meta.new_cache_slots, dtype=torch.int32), meta.q_cu_seq_lens,
        meta.kv_cu_seq_lens, torch.tensor(meta.block_tables, dtype=torch.int32), meta.cu_blocks_lens, hq, hkv, d)
    owned = g["owned_blocks"]
    assert torch.equal(kc[owned], g["ref_key_cache_owned"]) and torch.equal(vc[owned], g["ref_value_cache_owned"])
    assert torch.equal(q_rot, g["ref_query_rot"])
    assert torch.equal(out, g["ref_out"])


# ------------------------------------------------------------------------------------------------ GPU: kernel parity
def _shim():
    from hydrainfer_b200._C.kernel.position_embedding import apply_rotary_pos_emb, rope_set_kv_cache
    return apply_rotary_pos_emb, rope_set_kv_cache


@pytest.mark.gpu
def test_apply_rotary_matches_reference_fixture(golden_rope):
    apply_rotary_pos_emb, _ = _shim()
    g = golden_rope
    q, k = g.query.to(DEV), g.key.to(DEV)
    apply_rotary_pos_emb(q, k, g.positions.to(DEV), g.table.to(DEV), g.rd, g.interleaved)
    assert torch.equal(q.cpu(), g.ref_query) and torch.equal(k.cpu(), g.ref_key)
    # int64 positions (what torch.arange gives a caller) take the same path
    q, k = g.query.to(DEV), g.key.to(DEV)
    apply_rotary_pos_emb(q, k, g.positions.long().to(DEV), g.table.to(DEV), g.rd, g.interleaved)
    assert torch.equal(q.cpu(), g.ref_query) and torch.equal(k.cpu(), g.ref_key)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("table_f32", [False, True])
@pytest.mark.parametrize("interleaved", [False, True])
def test_apply_rotary_reference_grid(dtype, table_f32, interleaved):
    """tests/layer/test_rotary_embedding.py:66-75: 1000 tokens x (8|1) kv heads x 128, theta 1e5/5e5, max positions 8192."""
    apply_rotary_pos_emb, _ = _shim()
    g = torch.Generator().manual_seed(17 + int(interleaved) + 2 * int(table_f32))
    for n_tokens, hq, hkv, d, rd, theta in ((1000, 8, 8, 128, 128, 100000.), (257, 8, 1, 128, 128, 500000.), (33, 28, 4, 128, 64, 1000000.),
                                             (19, 4, 2, 64, 64, 10000.), (5, 2, 2, 256, 256, 10000.), (11, 3, 1, 40, 24, 10000.)):
        max_pos = 8192
        table = oracle.rotary_cos_sin_table(rd, max_pos, _inv_freq(rd, theta)).to(torch.float32 if table_f32 else dtype)
        q = torch.randn(n_tokens, hq, d, generator=g).to(dtype)
        k = torch.randn(n_tokens, hkv, d, generator=g).to(dtype)
        pos = torch.randint(0, max_pos, (n_tokens,), generator=g, dtype=torch.int32)
        ref_q, ref_k = oracle.apply_rotary(q, k, pos, table, rd, interleaved)
        q_d, k_d = q.to(DEV), k.to(DEV)
        apply_rotary_pos_emb(q_d, k_d, pos.to(DEV), table.to(DEV), rd, interleaved)
        assert torch.equal(q_d.cpu(), ref_q), (n_tokens, hq, hkv, d, rd)
        assert torch.equal(k_d.cpu(), ref_k), (n_tokens, hq, hkv, d, rd)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("force_scalar", [False, True])
@pytest.mark.parametrize("write_back_k", [False, True])
def test_fused_rope_append_equals_rotary_then_set_kv_cache(dtype, force_scalar, write_back_k):
    """One launch == the reference's two (rotary in place, then set_kv_cache), on q/k/v slices of a fused qkv row."""
    _, rope_set_kv_cache = _shim()
    g = torch.Generator().manual_seed(23)
    for t, hq, hkv, d, rd, bs, nb, interleaved in ((64, 32, 32, 128, 128, 16, 40, False), (37, 28, 4, 128, 128, 16, 30, False),
                                                    (1, 28, 4, 128, 128, 16, 8, False), (300, 8, 2, 64, 32, 8, 60, True), (9, 4, 1, 256, 128, 4, 16, True)):
        max_pos = 4096
        table = oracle.rotary_cos_sin_table(rd, max_pos, _inv_freq(rd, 10000.)).to(dtype)
        qkv = torch.randn(t, (hq + 2 * hkv) * d, generator=g).to(dtype)
        kc = torch.randn(nb, bs, hkv, d, generator=g).to(dtype)
        vc = torch.randn(nb, bs, hkv, d, generator=g).to(dtype)
        slots = torch.randperm(nb * bs, generator=g)[:t].to(torch.int32)
        pos = torch.randint(0, max_pos, (t,), generator=g, dtype=torch.int32)

        def views(x):
            return (x[:, :hq * d].view(t, hq, d), x[:, hq * d:(hq + hkv) * d].view(t, hkv, d), x[:, (hq + hkv) * d:].view(t, hkv, d))

        q, k, v = views(qkv)
        ref_q, ref_k = oracle.apply_rotary(q, k, pos, table, rd, interleaved)
        kc_ref, vc_ref = kc.clone(), vc.clone()
        oracle.set_kv_cache(slots, ref_k, v, kc_ref, vc_ref)

        qkv_d = qkv.to(DEV)
        q_d, k_d, v_d = views(qkv_d)
        kc_d, vc_d = kc.to(DEV), vc.to(DEV)
        rope_set_kv_cache(q_d, k_d, v_d, pos.to(DEV), table.to(DEV), rd, interleaved, slots.to(DEV), kc_d, vc_d, write_back_k, force_scalar)
        assert torch.equal(kc_d.cpu(), kc_ref) and torch.equal(vc_d.cpu(), vc_ref), (t, hq, hkv, d)
        assert torch.equal(q_d.cpu(), ref_q)
        assert torch.equal(k_d.cpu(), ref_k if write_back_k else k), "k must be rewritten only on request"
        assert torch.equal(v_d.cpu(), v)


@pytest.mark.gpu
def test_rotary_module_and_errors():
    from hydrainfer_b200.layer import RotaryEmbedding, compute_default_inv_freq
    rd, max_pos = 128, 2048
    emb = RotaryEmbedding(rotary_dim=rd, max_position_embeddings=max_pos, inv_freq=compute_default_inv_freq(rd, 1e6), interleaved=False)
    emb.to(torch.bfloat16).to(DEV)
    g = torch.Generator().manual_seed(3)
    q = torch.randn(12, 28, 128, generator=g).to(torch.bfloat16)
    k = torch.randn(12, 4, 128, generator=g).to(torch.bfloat16)
    pos = torch.arange(100, 112, dtype=torch.int32)
    table = oracle.rotary_cos_sin_table(rd, max_pos, compute_default_inv_freq(rd, 1e6)).to(torch.bfloat16)
    assert torch.equal(emb.handler.cos_sin_cache.cpu(), table)
    ref_q, ref_k = oracle.apply_rotary(q, k, pos, table, rd, False)
    out_q, out_k = emb(q.to(DEV), k.to(DEV), pos.to(DEV))
    assert torch.equal(out_q.cpu(), ref_q) and torch.equal(out_k.cpu(), ref_k)
    # no CPU path: CPU tensors raise instead of silently computing somewhere else
    with pytest.raises(RuntimeError):
        emb(q, k, pos)
    # layout the reference CHECK-aborts on (rope.cu:100-101) raises here
    apply_rotary_pos_emb, _ = _shim()
    with pytest.raises(RuntimeError):
        apply_rotary_pos_emb(q.to(DEV).transpose(0, 1), k.to(DEV).transpose(0, 1), pos.to(DEV), table.to(DEV), rd, False)
    with pytest.raises(RuntimeError):
        apply_rotary_pos_emb(q.to(DEV), k.to(DEV), pos.to(DEV), table.to(DEV).float().half(), rd, False)


@pytest.mark.gpu
@pytest.mark.parametrize("fuse", [True, False])
def test_rope_attention_module_matches_reference_fixture(fuse):
    """ROPECausalGroupedQueryPageAttention (model_forward.py:66-86) on the reference's frozen inputs: caches and rotated
    query bit-exact, attention output within 2e-2 + 1e-2*|fp32| (north star) of the fp32 recompute."""
    from hydrainfer_b200.layer import AttentionParametersBuilder, RotaryEmbedding, compute_default_inv_freq
    from hydrainfer_b200.memory import KVCache
    from hydrainfer_b200.model import ROPECausalGroupedQueryPageAttention
    g = _load_ropeattn()
    requests = _ropeattn_requests(g)
    hq, hkv, d, bs = g["hq"], g["hkv"], g["d"], g["bs"]
    emb = RotaryEmbedding(rotary_dim=d, max_position_embeddings=g["max_pos"], inv_freq=compute_default_inv_freq(d, g["theta"]), interleaved=False)
    emb.to(g["dtype"]).to(DEV)
    # the reference's frozen table (an ulp of the host's cosf/sinf can flip a bf16 rounding; see GoldenRope)
    torch.testing.assert_close(emb.handler.cos_sin_cache.cpu().float(), g["table"].float(), rtol=1e-2, atol=1e-6)
    emb.handler.cos_sin_cache.copy_(g["table"].to(DEV))
    module = ROPECausalGroupedQueryPageAttention(n_qo_heads=hq, n_kv_heads=hkv, head_dim=d, rotary_emb=emb, qkv_proj=torch.nn.Identity())
    module.fuse_rope_append = fuse
    kc, vc = g["key_cache"].to(DEV), g["value_cache"].to(DEV)
    builder = AttentionParametersBuilder(hq, hkv, d, bs, torch.device(DEV))
    for req in requests:
        builder.add_request(*req)
    builder.add_kv_cache(KVCache(kc, vc))
    params = builder.build_attention_parameters()[0]
    qkv = g["qkv"].to(DEV)
    out = module.forward(qkv, g["positions"].to(DEV), params)
    owned = g["owned_blocks"]
    assert torch.equal(kc.cpu()[owned], g["ref_key_cache_owned"]) and torch.equal(vc.cpu()[owned], g["ref_value_cache_owned"])
    mask = torch.ones(g["n_blocks"], dtype=torch.bool)
    mask[owned] = False
    assert torch.equal(kc.cpu()[mask], g["key_cache"][mask]) and torch.equal(vc.cpu()[mask], g["value_cache"][mask])
    assert torch.equal(qkv[:, :hq * d].cpu().view(-1, hq, d), g["ref_query_rot"])
    ref = g["ref_fp32"]
    err = (out.float().cpu() - ref).abs()
    assert bool((err <= 2e-2 + 1e-2 * ref.abs()).all()), float(err.max())
