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
    assert _lib.lib.hi_abi_version() == 4
    assert isinstance(_lib.lib.hi_last_error(), bytes)


def test_struct_layout_matches_header():
    # HiAttnArgs: 4 ptrs + 2 i64 + 4 ptrs + 8 i32 + i64 + i32 + f32 + ptr + i64 + 2 i32 + 4 i32 (natural alignment, no packing)
    assert ctypes.sizeof(_lib.HiAttnArgs) == 4 * 8 + 2 * 8 + 4 * 8 + 8 * 4 + 8 + 4 + 4 + 8 + 8 + 2 * 4 + 4 * 4 + 8 + 8 + 4 + 4
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
