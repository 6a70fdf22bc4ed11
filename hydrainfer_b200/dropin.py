"""Puts hydrainfer_b200's compiled modules under the reference's import names.

The reference imports its native code as
    hydrainfer._C.kernel.kv_cache_kernels      (hydrainfer/memory/kv_cache.py:9)
    hydrainfer._C.kernel.cache_kernels         (hydrainfer/memory/token_cache.py:9)
    hydrainfer._C.kernel.flash_attn            (hydrainfer/layer/causal_attention.py:14, multihead_attention.py:16)
    hydrainfer._C.kernel.position_embedding    (hydrainfer/layer/rotary_embedding.py:7)
    hydrainfer._C.data_transfer.block_migration (hydrainfer/memory/communication.py:10-11, token_cache_manger.py:18-19)
each a pybind extension with positional-only functions.  `install()` registers hydrainfer_b200's extensions of the same
names (`PyInit_<name>`, built from csrc/torch_binding.cpp over the C ABI of include/hi_b200.h) in `sys.modules` BEFORE the
reference's layer / memory modules are imported, so `CausalGroupedQueryPageAttention`, `KVCache`, `TokenCache`,
`TokenCacheBlockManager` and the IPC backend run unchanged on the B200 kernels.  The alternative with no import hook is to
copy the five `.so` files into `hydrainfer/_C/kernel/` and `hydrainfer/_C/data_transfer/` (where the reference's CMake puts
its own, csrc/CMakeLists.txt:4-11) with `hydrainfer_b200/lib` on LD_LIBRARY_PATH.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path
from typing import Optional

MODULES = {
    "hydrainfer._C.kernel.kv_cache_kernels": "hydrainfer_b200._C.kernel.kv_cache_kernels",
    "hydrainfer._C.kernel.cache_kernels": "hydrainfer_b200._C.kernel.cache_kernels",
    "hydrainfer._C.kernel.flash_attn": "hydrainfer_b200._C.kernel.flash_attn",
    "hydrainfer._C.kernel.position_embedding": "hydrainfer_b200._C.kernel.position_embedding",
    "hydrainfer._C.data_transfer.block_migration": "hydrainfer_b200._C.data_transfer.block_migration",
}


def install(reference_root: Optional[Path | str] = None) -> dict[str, object]:
    """Alias the five native modules; `reference_root` (the directory that holds the `hydrainfer` package) is put on
    sys.path when given.  Must run before `hydrainfer.layer` / `hydrainfer.memory` are first imported: their
    `from hydrainfer._C... import ...` statements bind the functions at import time."""
    late = [m for m in ("hydrainfer.memory.kv_cache", "hydrainfer.layer.causal_attention", "hydrainfer.memory.communication") if m in sys.modules]
    if late:
        raise RuntimeError(f"hydrainfer_b200.dropin.install() must run before {late} are imported")
    if reference_root is not None and str(reference_root) not in sys.path:
        sys.path.insert(0, str(reference_root))
    installed = {}
    for ref_name, our_name in MODULES.items():
        mod = importlib.import_module(our_name)
        sys.modules[ref_name] = mod
        installed[ref_name] = mod
    # `from hydrainfer._C.data_transfer import block_migration` (communication.py:11) reads the attribute off the parent package
    for ref_name, mod in installed.items():
        parent_name, _, leaf = ref_name.rpartition(".")
        try:
            parent = importlib.import_module(parent_name)
        except ImportError:
            continue  # no reference tree on sys.path yet: sys.modules entries are enough for `from a.b.c import f`
        setattr(parent, leaf, mod)
    return installed
