"""GPU: decode steps replayed from a CUDA graph over static metadata buffers (hydrainfer_b200/model_runner, the rebuilt
hydrainfer/model_runner/cuda_graph_model_runner.py) give the same results as the eager path and the oracle, for steps whose
lengths, block tables and slots differ from the ones seen at capture time."""
import pytest
import torch

from oracle import paged_kv_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _stack(hq, hkv, d, n_layers):
    from hydrainfer_b200.layer import CausalGroupedQueryPageAttention, CausalGroupedQueryPageAttentionConfig
    layers = [CausalGroupedQueryPageAttention(CausalGroupedQueryPageAttentionConfig(hq, hkv, d)) for _ in range(n_layers)]
    qw, kw = hq * d, hkv * d

    def model_runner(hidden, position_ids, attention_params):
        # hidden [B, (Hq + 2 Hkv) d]: fused qkv of the step; every layer attends with its own cache, outputs are summed in fp32
        q, k, v = hidden[:, :qw], hidden[:, qw:qw + kw], hidden[:, qw + kw:]
        acc = None
        for layer, params in zip(layers, attention_params):
            o = layer(q, k, v, params).o.float()
            acc = o if acc is None else acc + o
        return acc.to(hidden.dtype)

    return model_runner


def _requests(lens, block_size, n_blocks, seed):
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(n_blocks, generator=g).tolist()
    reqs, used = [], 0
    for kv in lens:
        nb = (kv + block_size - 1) // block_size
        table = perm[used:used + nb]
        used += nb
        slot = table[(kv - 1) // block_size] * block_size + (kv - 1) % block_size
        reqs.append((1, kv, [slot], table))
    return reqs


@pytest.mark.parametrize("hq,hkv", [(8, 8), (28, 4)])
def test_graph_replay_matches_eager_and_oracle(hq, hkv):
    from hydrainfer_b200.layer import AttentionParametersBuilder
    from hydrainfer_b200.memory import KVCache
    from hydrainfer_b200.model_runner import CudaGraphModelRunner
    d, bs, n_blocks, n_layers, max_seq = 128, 16, 96, 3, 512
    g = torch.Generator().manual_seed(0)
    caches_cpu = [(torch.randn(n_blocks, bs, hkv, d, generator=g).to(torch.bfloat16), torch.randn(n_blocks, bs, hkv, d, generator=g).to(torch.bfloat16)) for _ in range(n_layers)]
    kv_caches = [KVCache(k.to(DEV), v.to(DEV)) for k, v in caches_cpu]
    width = (hq + 2 * hkv) * d
    runner = CudaGraphModelRunner(_stack(hq, hkv, d, n_layers), torch.bfloat16, DEV, bs, width, hq * d, kv_caches, hq, hkv, d,
                                  cuda_graph_max_batch_size=4, cuda_graph_max_seq_len=max_seq, batch_sizes=[1, 3, 4])
    eager_model = _stack(hq, hkv, d, n_layers)
    for step, lens in enumerate([[17, 300, 64], [512, 1, 33, 200], [5], [40, 41, 42], [100, 90]]):
        reqs = _requests(lens, bs, n_blocks, seed=10 + step)
        hidden = torch.randn(len(lens), width, generator=g).to(torch.bfloat16)
        pos = torch.tensor([kv - 1 for kv in lens], dtype=torch.int32)
        # oracle on CPU copies of the caches (append + attention per layer, summed in fp32)
        slots = torch.tensor([r[2][0] for r in reqs], dtype=torch.int32)
        meta = oracle.build_metadata(reqs, bs)
        ref = torch.zeros(len(lens), hq * d)
        q, k, v = hidden[:, :hq * d], hidden[:, hq * d:(hq + hkv) * d], hidden[:, (hq + hkv) * d:]
        for kc, vc in caches_cpu:
            oracle.set_kv_cache(slots, k.reshape(-1, hkv, d), v.reshape(-1, hkv, d), kc, vc)
            ref += oracle.paged_attention_fp32(q.reshape(-1, hq, d), kc, vc, meta.q_cu_seq_lens, meta.kv_cu_seq_lens,
                                               torch.tensor(meta.block_tables, dtype=torch.int32), meta.cu_blocks_lens, hq, hkv, d)
        builder = AttentionParametersBuilder(hq, hkv, d, bs, DEV)
        for r in reqs:
            builder.add_request(*r)
        replays_before = runner.replays
        out = runner(hidden.to(DEV), pos.to(DEV), builder).clone()
        torch.cuda.synchronize()
        assert (runner.replays > replays_before) == (len(lens) in (1, 3, 4)), "graph used exactly for the captured batch sizes"
        for (kc, vc), cache in zip(caches_cpu, kv_caches):  # the append inside the graph is bit-exact
            assert torch.equal(cache.key_cache.cpu(), kc) and torch.equal(cache.value_cache.cpu(), vc)
        err = (out.float().cpu() - ref).abs()
        assert bool((err <= n_layers * (2e-2 + 1e-2 * ref.abs())).all()), f"step {step}: max |err| {err.max().item():.3e}"
        # eager path on the same (already appended) caches: same kernels and inputs; only the split-KV chunking may differ (the graph
        # was captured for the kv_max_seq_len bound), i.e. the fp32 summation order
        eb = AttentionParametersBuilder(hq, hkv, d, bs, DEV)
        for r in reqs:
            eb.add_request(*r)
        for c in kv_caches:
            eb.add_kv_cache(c)
        eager = eager_model(hidden.to(DEV), pos.to(DEV), eb.build_attention_parameters())
        torch.cuda.synchronize()
        assert (eager.float() - out.float()).abs().max().item() <= 2e-2, f"step {step}: graph replay differs from the eager path"
    assert runner.replays == 4 and runner.eager_calls == 1


def test_steps_outside_the_captured_bounds_run_eagerly():
    from hydrainfer_b200.layer import AttentionParametersBuilder
    from hydrainfer_b200.memory import KVCache
    from hydrainfer_b200.model_runner import CudaGraphModelRunner
    hq, hkv, d, bs, n_blocks = 4, 4, 128, 16, 64
    kv_caches = [KVCache(torch.randn(n_blocks, bs, hkv, d, device=DEV).to(torch.float16), torch.randn(n_blocks, bs, hkv, d, device=DEV).to(torch.float16))]
    width = (hq + 2 * hkv) * d
    runner = CudaGraphModelRunner(_stack(hq, hkv, d, 1), torch.float16, DEV, bs, width, hq * d, kv_caches, hq, hkv, d,
                                  cuda_graph_max_batch_size=2, cuda_graph_max_seq_len=64)
    for lens, q_lens in (([65, 3], [1, 1]), ([20, 30], [1, 4]), ([10, 11, 12], [1, 1, 1])):  # too long / prefill row / batch too large
        builder = AttentionParametersBuilder(hq, hkv, d, bs, DEV)
        used = 0
        for kv, ql in zip(lens, q_lens):
            nb = (kv + bs - 1) // bs
            table = list(range(used, used + nb))
            used += nb
            slots = [table[p // bs] * bs + p % bs for p in range(kv - ql, kv)]
            builder.add_request(ql, kv, slots, table)
        hidden = torch.randn(sum(q_lens), width, device=DEV).to(torch.float16)
        out = runner(hidden, torch.zeros(sum(q_lens), dtype=torch.int32, device=DEV), builder)
        torch.cuda.synchronize()
        assert out.shape == (sum(q_lens), hq * d) and torch.isfinite(out.float()).all()
    assert runner.replays == 0 and runner.eager_calls == 3
