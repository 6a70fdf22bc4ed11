"""GPU: the REFERENCE's own `CausalGroupedQueryPageAttention`, `KVCache`, `AttentionParametersBuilder` and
`TokenCacheBlockManager` (python files copied unmodified into oracle/_ref by oracle/build_ref.py) running on the B200 with
`hydrainfer._C.*` resolved to hydrainfer_b200's compiled modules (hydrainfer_b200/dropin.py) — SURVEY §7 step 7, the drop-in
proof — and, beside it, the same reference code on its OWN compiled csrc (kv_cache_kernels.cu, the vendored FlashAttention-2
`mha_varlen_fwd`, ...), so the two native module sets are compared through the reference's own layer on identical inputs.

Each configuration runs in a subprocess (tests/ref_layer_runner.py): one `hydrainfer` package per process."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from oracle import reference_tree as ref_tree

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _run(native: str, tmp_path) -> dict:
    out = tmp_path / f"{native}.pt"
    res = subprocess.run([sys.executable, str(ROOT / "tests" / "ref_layer_runner.py"), "--native", native, "--out", str(out)],
                         capture_output=True, text=True, timeout=900)
    assert "RUNNER-OK" in res.stdout, f"runner ({native}) failed:\n{res.stdout[-2000:]}\n{res.stderr[-6000:]}"
    return torch.load(out)


@pytest.fixture(scope="module")
def ours(tmp_path_factory):
    if not ref_tree.available():
        pytest.skip("oracle/_ref did not travel to this box (git-ignored build product of `python oracle/build_ref.py`)")
    return _run("ours", tmp_path_factory.mktemp("dropin"))


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    if not ref_tree.available("kv_cache_kernels"):
        pytest.skip("oracle/_ref native modules did not travel to this box")
    return _run("ref", tmp_path_factory.mktemp("refnative"))


def test_reference_modules_bind_to_our_compiled_extensions(ours):
    bound = ours["bound"]
    assert "hydrainfer_b200/_C/kernel/flash_attn" in bound["mha_varlen_fwd_file"], bound
    assert "hydrainfer_b200/_C/data_transfer/block_migration" in bound["block_migration_file"], bound
    assert bound["set_image_cache_is_native"]


def test_reference_layer_on_our_kernels_matches_its_own_torch_handler(ours):
    """north_star tolerance against the reference's Torch handler (fp32 math, rounded to the query dtype) run by the reference
    itself on the same device tensors: |ours - torch| <= 2e-2 + 1e-2 |torch| (one extra 16-bit rounding of the comparand)."""
    for name, case in ours["cases"].items():
        out, want = case["out"].float(), case["torch_handler_out"].float()
        assert out.shape == want.shape, name
        err = (out - want).abs()
        assert bool((err <= 2e-2 + 1e-2 * want.abs()).all()), f"{name}: max |err| {err.max().item():.3e}"


def test_reference_layer_kv_append_is_identical_under_both_native_sets(ours, ref):
    for name in ours["cases"]:
        assert torch.equal(ours["cases"][name]["key_cache"], ref["cases"][name]["key_cache"]), name
        assert torch.equal(ours["cases"][name]["value_cache"], ref["cases"][name]["value_cache"]), name


def test_reference_layer_output_matches_the_reference_fa2_backend(ours, ref):
    """Same reference layer, same inputs: our `mha_varlen_fwd` against the reference's vendored FlashAttention-2 (a7).  Both
    round a fp32 accumulation to 16 bits; the reference's own cross-backend tolerance is atol = rtol = 1e-2
    (tests/layer/test_attention.py:102-106)."""
    if not ref_tree.available("flash_attn"):
        pytest.skip("the reference's flash_attn was not built into oracle/_ref (python oracle/build_ref.py --fa2)")
    assert "oracle/_ref" in ref["bound"]["mha_varlen_fwd_file"], ref["bound"]
    for name in ours["cases"]:
        a, b = ours["cases"][name]["out"].float(), ref["cases"][name]["out"].float()
        torch.testing.assert_close(a, b, atol=1e-2, rtol=1e-2, msg=lambda m: f"{name}: {m}")
        # and the reference's own FA2 agrees with its own Torch handler to the same tolerance (sanity of the comparand)
        torch.testing.assert_close(b, ref["cases"][name]["torch_handler_out"].float(), atol=2e-2, rtol=1e-2, msg=lambda m: f"{name} (reference FA2 vs Torch): {m}")


def test_reference_block_manager_on_our_block_migration(ours, ref):
    m = ours["manager"]
    assert m["handle_len"] in (64, 72)
    assert m["src_table"] == ref["manager"]["src_table"] and m["dst_table"] == ref["manager"]["dst_table"]
    assert m["v2p"] == ref["manager"]["v2p"] and m["available"] == ref["manager"]["available"]
    assert m["moved_ok"] and m["untouched_ok"] and m["layer_cache_ptr_ok"]
