"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/hi_b200.h declares."""
import ctypes
import re
from pathlib import Path

from hydrainfer_b200 import _lib

HEADER = Path(__file__).resolve().parents[1] / "include" / "hi_b200.h"


def declared_functions() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(hi_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_functions()
    for required in ("hi_set_kv_cache", "hi_set_image_cache", "hi_rope_append", "hi_paged_attention", "hi_migrate_blocks", "hi_ipc_get_handle", "hi_ipc_open_handle"):
        assert required in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/hi_b200.h but missing from {_lib.LIB_PATH.name}"


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_abi_version_and_error_string():
    assert _lib.lib.hi_abi_version() == 6
    assert isinstance(_lib.lib.hi_last_error(), bytes)


def test_struct_layout_matches_header():
    # HiAttnArgs: 4 ptrs + 2 i64 + 4 ptrs + 8 i32 + i64 + i32 + f32 + ptr + i64 + 2 i32 + 4 i32 (natural alignment, no packing)
    assert ctypes.sizeof(_lib.HiAttnArgs) == 4 * 8 + 2 * 8 + 4 * 8 + 8 * 4 + 8 + 4 + 4 + 8 + 8 + 2 * 4 + 4 * 4 + 8 + 8 + 4 + 4 + 4 * 4 + 8 + 8 == 224
    assert ctypes.sizeof(_lib.HiPoolGeom) == 32
    # HiRopeArgs: 3 ptrs + 3 i64 + 5 ptrs + i64 + 12 i32
    assert ctypes.sizeof(_lib.HiRopeArgs) == 3 * 8 + 3 * 8 + 5 * 8 + 8 + 12 * 4
    # HiVarlenArgs: 4 ptrs + 4 i64 + 2 ptrs + 10 i32 + f32 + i32 + ptr + i64 + 2 i64 (api.cu static_asserts the same 160)
    assert ctypes.sizeof(_lib.HiVarlenArgs) == 4 * 8 + 4 * 8 + 2 * 8 + 10 * 4 + 4 + 4 + 8 + 8 + 16 == 160


def test_argument_validation_needs_no_gpu():
    # null args are rejected before any CUDA call
    assert _lib.lib.hi_paged_attention(None, None) == -1
    assert b"null args" in _lib.lib.hi_last_error()
    assert _lib.lib.hi_varlen_attention(None, None) == -1
    assert b"null args" in _lib.lib.hi_last_error()
    assert _lib.lib.hi_attention_workspace_bytes(64, 32, 128, 2048) > 0


# ---- the compiled Python boundary: the reference's five pybind modules ------------------------------------------
REFERENCE_MODULES = {
    # module -> (package directory, functions the reference imports from it)
    "kv_cache_kernels": ("_C/kernel", ["set_kv_cache"]),                          # hydrainfer/memory/kv_cache.py:9
    "cache_kernels": ("_C/kernel", ["set_image_cache"]),                          # hydrainfer/memory/token_cache.py:9
    "flash_attn": ("_C/kernel", ["mha_varlen_fwd"]),                              # hydrainfer/layer/causal_attention.py:14
    "position_embedding": ("_C/kernel", ["apply_rotary_pos_emb"]),                # hydrainfer/layer/rotary_embedding.py:7
    "block_migration": ("_C/data_transfer", ["get_ipc_mem_handle", "register_ipc_mem_handle", "migrate_blocks"]),  # communication.py:10-11
}


def test_compiled_modules_export_the_reference_init_symbols():
    """Each module is a real extension with the `PyInit_<name>` entry point the reference's CMake target produces
    (csrc/CMakeLists.txt:4-11, PYBIND11_MODULE(PY_MODULE_NAME, m)), not a Python shim."""
    import importlib
    import importlib.machinery

    from hydrainfer_b200 import build
    for name, path in build.binding_paths().items():
        assert path.exists(), path
        assert path.name.endswith(tuple(importlib.machinery.EXTENSION_SUFFIXES)), path.name
        lib = ctypes.CDLL(str(path))
        assert hasattr(lib, f"PyInit_{name}"), f"{path.name} does not export PyInit_{name}"
    for name, (sub, functions) in REFERENCE_MODULES.items():
        mod = importlib.import_module(f"hydrainfer_b200.{sub.replace('/', '.')}.{name}")
        assert isinstance(mod.__loader__, importlib.machinery.ExtensionFileLoader), f"{name} is not a compiled extension"
        for fn in functions:
            assert callable(getattr(mod, fn)), f"{name}.{fn}"


def test_mha_varlen_fwd_keeps_the_reference_positional_signature():
    """16 positional arguments (hydrainfer/_C/kernel/flash_attn/__init__.pyi:22-40); a call with 15 is a TypeError, CPU tensors
    a RuntimeError (no CPU fallback), a negative softcap a RuntimeError like TORCH_CHECK (softcap / windows / alibi themselves are
    accepted by the paged form: tests/test_gpu_attention.py::test_score_options_match_oracle)."""
    import pytest
    import torch

    from hydrainfer_b200._C.kernel.flash_attn import mha_varlen_fwd
    q = torch.zeros(2, 4, 128, dtype=torch.float16)
    kc = torch.zeros(3, 16, 4, 128, dtype=torch.float16)
    cu = torch.tensor([0, 2], dtype=torch.int32)
    bt = torch.tensor([0], dtype=torch.int32)
    cub = torch.tensor([0, 1], dtype=torch.int32)
    with pytest.raises(TypeError):
        mha_varlen_fwd(q, q, kc, kc, cu, cu, bt, cub, None, 2, 2, 0.1, 0, -1, 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mha_varlen_fwd(q, q, kc, kc, cu, cu, bt, cub, None, 2, 2, 0.1, 0, -1, 0, 0)
    with pytest.raises(RuntimeError, match="softcap"):
        mha_varlen_fwd(q, q, kc, kc, cu, cu, bt, cub, None, 2, 2, 0.1, -30.0, -1, 0, 0)
    with pytest.raises(RuntimeError, match="cu_block_lens"):
        mha_varlen_fwd(q, q, kc, kc, cu, cu, bt, None, None, 2, 2, 0.1, 0, -1, 0, 0)


def test_dropin_install_aliases_the_reference_names():
    """hydrainfer_b200.dropin.install() in a fresh interpreter: the five `hydrainfer._C...` names resolve to our extensions."""
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from hydrainfer_b200 import dropin\n"
        "mods = dropin.install()\n"
        "assert len(mods) == 5\n"
        "import importlib\n"
        "for name, mod in mods.items():\n"
        "    assert sys.modules[name] is mod and 'hydrainfer_b200/_C' in mod.__file__, (name, mod)\n"
        "print('ALIASED')\n" % str(root))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "ALIASED" in res.stdout, res.stderr[-3000:]
