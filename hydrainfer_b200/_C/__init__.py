"""The compiled Python boundary: pybind extensions with the reference's module names and `PyInit_<name>` entry points
(`hydrainfer._C.kernel.{kv_cache_kernels,cache_kernels,flash_attn,position_embedding}`,
`hydrainfer._C.data_transfer.block_migration`; csrc/CMakeLists.txt:4-11), built from csrc/torch_binding.cpp over the C ABI.

Importing this package loads libhi_b200.so first (so the extensions' DT_NEEDED entry resolves to the in-tree library) and
(re)builds the extensions when they are missing or older than their source.  No fallback: if neither a current binary nor a
compiler is available the import fails."""
from .. import _lib  # noqa: F401  dlopens (or builds) lib/libhi_b200.so
from .. import build as _build

if not _build.binding_is_current():
    try:
        _build.build_binding()
    except Exception as e:  # pragma: no cover - a broken install
        raise ImportError(f"hydrainfer_b200: the compiled modules under _C/ are missing or stale and cannot be built: {e}") from e
