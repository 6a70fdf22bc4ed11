"""Runs the REFERENCE's own python layer + memory classes (the files under oracle/_ref/hydrainfer, copied unmodified from
/root/reference by oracle/build_ref.py) on cuda:0 and dumps what they produce.  Not a test module: tests/
test_gpu_dropin_reference_layer.py launches it as a subprocess once per configuration, because the package name
`hydrainfer` can be bound to only one set of native modules per process.

    python tests/ref_layer_runner.py --native ours --out /tmp/ours.pt     hydrainfer._C.* = hydrainfer_b200's compiled modules
    python tests/ref_layer_runner.py --native ref  --out /tmp/ref.pt      hydrainfer._C.* = the reference's own csrc, compiled

Per case it records: the output of CausalGroupedQueryPageAttention.forward (causal_attention.py:394-406, handler chain
FlashInfer -> FlashAttention(csrc mha_varlen_fwd) -> Torch; no flashinfer wrappers are planned, so the chain lands on the
`mha_varlen_fwd` of whichever native module set is installed), the appended caches, the output of the reference's Torch
handler (:307-374) on the same device tensors, and a TokenCacheBlockManager scenario (allocator + v2p + IPC export +
migration) when the native set supports it in one process.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from hydrainfer_b200.workloads import make_batch  # noqa: E402
from oracle import reference_tree  # noqa: E402

# (name, seq_lens [(q, kv)], Hq, Hkv, d, block_size, dtype, fused_qkv)
CASES = [
    ("llava_decode", [(1, 37), (1, 512), (1, 2048), (1, 16)], 32, 32, 128, 16, torch.bfloat16, False),
    ("qwen_mixed_fused", [(1, 300), (64, 64), (200, 1000), (1, 17), (512, 2048)], 28, 4, 128, 16, torch.bfloat16, True),
    ("qwen72b_decode", [(1, 4096), (1, 100), (1, 1000)], 64, 8, 128, 16, torch.float16, False),
    ("mha_prefill_fp16", [(33, 33), (5, 70), (300, 300)], 8, 8, 128, 16, torch.float16, False),
    ("gqa_d64", [(1, 90), (40, 130)], 8, 2, 64, 16, torch.bfloat16, False),
    ("mqa_d256", [(1, 50), (20, 77)], 4, 1, 256, 16, torch.float16, False),
]


def run(native: str) -> dict:
    ca, mem = reference_tree.import_reference_package(native=native)
    dev = torch.device("cuda:0")
    results = {"native": native, "cases": {}}
    for name, seq_lens, hq, hkv, d, bs, dtype, fused in CASES:
        batch = make_batch(seq_lens, hq, hkv, d, bs, dtype=dtype, seed=11, fused_qkv=fused).to(dev)
        kc, vc = batch.key_cache.clone(), batch.value_cache.clone()
        builder = ca.AttentionParametersBuilder(num_qo_heads=hq, num_kv_heads=hkv, head_dim=d, block_size=bs, device=dev)
        for req in batch.requests():
            builder.add_request(*req)
        builder.add_kv_cache(mem.KVCache(kc, vc))
        params = builder.build_attention_parameters()[0]
        layer = ca.CausalGroupedQueryPageAttention(ca.CausalGroupedQueryPageAttentionConfig(n_qo_heads=hq, n_kv_heads=hkv, head_dim=d))
        out = layer(batch.query, batch.key, batch.value, params).o
        torch.cuda.synchronize()
        # the reference's own Torch handler on the same (already appended) caches: the a6 oracle, run by the reference itself
        q3 = batch.query.view(-1, hq, d)
        torch_out = layer.handlers[2](q3, params).o
        torch.cuda.synchronize()
        results["cases"][name] = {"out": out.cpu(), "torch_handler_out": torch_out.cpu(), "key_cache": kc.cpu(), "value_cache": vc.cpu()}

    # TokenCacheBlockManager: pool allocation + IPC export happen in the constructor (token_cache_manger.py:65-74)
    def manager(rank: int, n_blocks: int):
        cfg = mem.TokenCacheBlockManagerConfig(mem.CommunicationBackendManagerConfig(), n_layers=3, n_tokens=2, n_blocks=n_blocks, block_size=16,
                                               n_heads=4, head_size=128, dtype="fp16", device="cuda:0")
        return mem.TokenCacheBlockManager(cfg, mem.TokenCacheBlockManagerContext(rank=rank, rank2host={0: "node", 1: "node"}))

    torch.manual_seed(3)
    prefill, decode = manager(0, 400), manager(1, 300)
    src = prefill.allocate_virtual_cache()
    prefill.realloc(src, 70)
    dst = decode.allocate_virtual_cache()
    decode.realloc(dst, 70)
    scenario = {"handle_len": len(prefill.memory_handle), "src_table": list(src.block_table), "dst_table": list(dst.block_table),
                "v2p": prefill.v2p(src, [0, 15, 16, 69]), "available": prefill.get_num_avaiable_blocks()}
    if native == "ours":
        # same-process migration: the reference's own migrate_blocks cannot open a handle exported by its own process
        # (block_migration.cpp:213-215), ours resolves it locally; the cross-process comparison is tests/test_gpu_reference_native.py
        before = decode.cache_tensor.clone()
        decode.migrate_blocks(src, dst, is_send=False)
        prefill.migrate_blocks(src, dst, is_send=True)
        decode.synchronize()
        scenario["moved_ok"] = all(torch.equal(decode.cache_tensor[:, :, d_], prefill.cache_tensor[:, :, s_]) for s_, d_ in zip(src.block_table, dst.block_table))
        untouched = [i for i in range(300) if i not in dst.block_table]
        scenario["untouched_ok"] = torch.equal(decode.cache_tensor[:, :, untouched], before[:, :, untouched])
        layer_cache = decode.get_layer_cache(1).get_caches()
        scenario["layer_cache_ptr_ok"] = layer_cache[0].data_ptr() == decode.cache_tensor[1, 0].data_ptr()
    results["manager"] = scenario

    # which native functions the reference modules ended up bound to
    import hydrainfer.memory.kv_cache as ref_kv
    import hydrainfer.memory.token_cache as ref_tc
    import hydrainfer.memory.communication as ref_comm
    results["bound"] = {
        "set_kv_cache": getattr(ref_kv.set_kv_cache_kernel, "__module__", None) or type(ref_kv.set_kv_cache_kernel).__name__,
        "mha_varlen_fwd_file": sys.modules[ca.mha_varlen_fwd.__module__].__file__ if ca.mha_varlen_fwd is not None and ca.mha_varlen_fwd.__module__ in sys.modules else str(ca.mha_varlen_fwd),
        "set_image_cache_is_native": ref_tc.set_image_cache is not None,
        "block_migration_file": getattr(ref_comm.block_migration, "__file__", None),
    }
    return results


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--native", choices=["ours", "ref"], required=True)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    torch.save(run(args.native), args.out)
    print("RUNNER-OK")
