"""CPU, build container only: the mirror modules slot in under the reference's import names.
Skipped where /root/reference does not exist (the GPU box); never part of the -m gpu run."""
import subprocess
import sys
from pathlib import Path

import pytest

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parents[1]

SCRIPT = r'''
import sys, types
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "seaborn"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib.colors"].LogNorm = object
sys.path.insert(0, "/root/reference")
sys.path.insert(0, %r)
import hydrainfer_b200._C.kernel.kv_cache_kernels as _kv
import hydrainfer_b200._C.kernel.cache_kernels as _img
import hydrainfer_b200._C.kernel.flash_attn as _fa
import hydrainfer_b200._C.data_transfer.block_migration as _bm
sys.modules["hydrainfer._C.kernel.kv_cache_kernels"] = _kv
sys.modules["hydrainfer._C.kernel.cache_kernels"] = _img
sys.modules["hydrainfer._C.kernel.flash_attn"] = _fa
sys.modules["hydrainfer._C.data_transfer.block_migration"] = _bm
import hydrainfer._C.data_transfer as _dt
_dt.block_migration = _bm
import hydrainfer.memory.kv_cache as ref_kv
import hydrainfer.memory.token_cache as ref_tc
import hydrainfer.layer.causal_attention as ref_ca
import hydrainfer.memory.communication as ref_comm
import hydrainfer.memory.token_cache_manger as ref_mgr
assert ref_kv.set_kv_cache_kernel is _kv.set_kv_cache
assert ref_tc.set_image_cache is _img.set_image_cache
assert ref_ca.mha_varlen_fwd is _fa.mha_varlen_fwd
assert ref_comm.block_migration is _bm and ref_mgr.get_ipc_mem_handle is _bm.get_ipc_mem_handle
# the native handler accepts the reference's own config object and chains in front of its handlers
from hydrainfer_b200.layer import B200CausalGroupedQueryPageAttentionHandler
cfg = ref_ca.CausalGroupedQueryPageAttentionConfig(n_qo_heads=8, n_kv_heads=2, head_dim=128)
attn = ref_ca.CausalGroupedQueryPageAttention(cfg)
ours = B200CausalGroupedQueryPageAttentionHandler(cfg)
ours.next_handler = attn.handlers[0]
attn.handlers.insert(0, ours); attn.handler = ours
# CPU tensors fall through the whole chain to the reference's torch handler
import torch
from hydrainfer.memory import KVCache
kc = torch.randn(4, 16, 2, 128); vc = torch.randn(4, 16, 2, 128)
b = ref_ca.AttentionParametersBuilder(8, 2, 128, 16, torch.device("cpu"))
b.add_request(3, 3, [16, 17, 18], [1]); b.add_kv_cache(KVCache(kc, vc))
p = b.build_attention_parameters()[0]
o = attn(torch.randn(3, 8 * 128), torch.randn(3, 2 * 128), torch.randn(3, 2 * 128), p).o
assert o.shape == (3, 1024)
print("WIRING-OK")
'''


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present")
def test_mirror_modules_resolve_under_reference_names():
    res = subprocess.run([sys.executable, "-c", SCRIPT % str(ROOT)], capture_output=True, text=True, timeout=300)
    assert "WIRING-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
