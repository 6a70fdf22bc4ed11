"""TEST INFRASTRUCTURE — loaders for `oracle/_ref/` (the reference "as built", see oracle/build_ref.py).

Two ways in:
  * `load_native(name)`: one of the reference's OWN compiled pybind modules (kv_cache_kernels, cache_kernels,
    position_embedding, block_migration, flash_attn), loaded by file path under a private module name so that it can sit in
    the same process as hydrainfer_b200's modules of the same name.  This is what the `-m gpu` parity tests compare against.
  * `import_reference_package(native=...)`: puts oracle/_ref on sys.path and imports the reference's python package
    `hydrainfer` (layer + memory).  native="ref" leaves its `_C` modules as built from its own sources; native="ours"
    aliases `hydrainfer._C.*` to hydrainfer_b200's compiled modules first — the drop-in configuration of INTEGRATION.md §1.
    One configuration per process (the package name is fixed), so tests run each in a subprocess.

Only tests/, smoke() and bench.py's reference legs may import this file.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import sys
import sysconfig
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF_TREE = ROOT / "oracle" / "_ref"
PKG = REF_TREE / "hydrainfer"
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")

NATIVE_DIRS = {
    "kv_cache_kernels": "_C/kernel",
    "cache_kernels": "_C/kernel",
    "position_embedding": "_C/kernel",
    "flash_attn": "_C/kernel",
    "block_migration": "_C/data_transfer",
}


def native_path(name: str) -> Path:
    return PKG / NATIVE_DIRS[name] / f"{name}{EXT_SUFFIX}"


def available(name: str | None = None) -> bool:
    """True when oracle/_ref holds the python tree (and, if given, the named compiled module)."""
    if not (PKG / "layer" / "causal_attention.py").exists():
        return False
    return name is None or native_path(name).exists()


_loaded: dict[str, types.ModuleType] = {}


def load_native(name: str) -> types.ModuleType:
    """The reference's own compiled module `name` (PyInit_<name>), imported from its file under oracle/_ref."""
    if name in _loaded:
        return _loaded[name]
    import torch  # noqa: F401  (the modules link against libtorch; it must be loaded first)

    path = native_path(name)
    if not path.exists():
        raise FileNotFoundError(f"{path} is missing: run `python oracle/build_ref.py` in the build container")
    loader = importlib.machinery.ExtensionFileLoader(name, str(path))
    spec = importlib.util.spec_from_file_location(name, str(path), loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _loaded[name] = mod
    return mod


def stub_plotting_modules() -> None:
    """hydrainfer/utils/statistic.py:2-7 imports matplotlib / seaborn, which this image lacks; nothing on the path uses them."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.colors"].LogNorm = object


def import_reference_package(native: str = "ref"):
    """Imports the reference's `hydrainfer.layer.causal_attention` and `hydrainfer.memory` from oracle/_ref.

    native="ref":  `hydrainfer._C.*` are the reference's own compiled kernels.
    native="ours": `hydrainfer._C.*` resolve to hydrainfer_b200's compiled modules (drop-in configuration).
    native="none": neither (pure torch fallbacks; what oracle/make_golden.py pins the oracle against).
    """
    if "hydrainfer" in sys.modules:
        raise RuntimeError("the reference package is already imported in this process; use one configuration per process")
    if not available():
        raise FileNotFoundError(f"{PKG} is missing: run `python oracle/build_ref.py` in the build container")
    stub_plotting_modules()
    if native == "ours":
        if str(ROOT) not in sys.path:
            sys.path.insert(0, str(ROOT))
        from hydrainfer_b200 import dropin
        dropin.install(reference_root=REF_TREE)
    elif native == "none":
        for name, sub in NATIVE_DIRS.items():
            sys.modules[f"hydrainfer.{sub.replace('/', '.')}.{name}"] = None  # import -> ImportError -> the torch fallbacks
        sys.path.insert(0, str(REF_TREE))
    else:
        sys.path.insert(0, str(REF_TREE))
    import hydrainfer.layer.causal_attention as ca
    import hydrainfer.memory as mem
    return ca, mem
