"""Builds the two native artefacts in-tree; both travel with the source tree to the GPU box.

1. lib/libhi_b200.so — the C-ABI CUDA library (include/hi_b200.h), nvcc for sm_100a.  No torch headers: plain CUDA behind
   `extern "C"`, a handful of `nvcc -c` calls plus one link.
2. the compiled Python boundary — csrc/torch_binding.cpp (pybind11 + ATen, g++ only, no device code) linked against (1) and
   installed under the reference's five module names with their `PyInit_<name>` entry points:
   _C/kernel/{kv_cache_kernels,cache_kernels,flash_attn,position_embedding}<EXT_SUFFIX> and
   _C/data_transfer/block_migration<EXT_SUFFIX> (the reference's CMake puts its own there, csrc/CMakeLists.txt:4-11).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libhi_b200.so"
OBJ_DIR = CSRC / "build"
SOURCES = ["api.cu", "scatter.cu", "rope.cu", "attn_simt.cu", "attn_tc.cu", "attn_tc2.cu", "attn_decode_tc.cu", "migrate.cu"]
HEADERS = ["common.cuh", "ptx_sm100.cuh", "attn_common.cuh", "tma_maps.h", "../../include/hi_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


# Extra -D flags for debug builds, e.g. HI_BUILD_DEFINES="-DHI_MBAR_DEBUG" (part of the build stamp).
NVCC_FLAGS += [f for f in os.environ.get("HI_BUILD_DEFINES", "").split() if f]


BINDING_SRC = CSRC / "torch_binding.cpp"
BINDING_LIB = LIB_DIR / "hi_b200_torch.so"
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")
# module name -> package directory (relative to hydrainfer_b200/) that holds <name><EXT_SUFFIX>
BINDING_MODULES = {
    "kv_cache_kernels": "_C/kernel",
    "cache_kernels": "_C/kernel",
    "flash_attn": "_C/kernel",
    "position_embedding": "_C/kernel",
    "block_migration": "_C/data_transfer",
}


def binding_paths() -> dict[str, Path]:
    return {name: PKG_DIR / sub / f"{name}{EXT_SUFFIX}" for name, sub in BINDING_MODULES.items()}


def _binding_stamp() -> str:
    import torch

    h = hashlib.sha256()
    h.update(BINDING_SRC.read_bytes())
    h.update((PKG_DIR.parent / "include" / "hi_b200.h").read_bytes())
    h.update(f"{torch.__version__}{EXT_SUFFIX}".encode())
    return h.hexdigest()


def binding_is_current() -> bool:
    stamp_file = LIB_DIR / "binding.stamp"
    return stamp_file.exists() and stamp_file.read_text() == _binding_stamp() and all(p.exists() for p in binding_paths().values())


def build_binding(force: bool = False) -> dict[str, Path]:
    """g++ csrc/torch_binding.cpp -> lib/hi_b200_torch.so (links libhi_b200.so by $ORIGIN-relative rpath), then one copy per
    module name.  Needs libhi_b200.so to exist (link step)."""
    if not force and binding_is_current():
        return binding_paths()
    import torch

    cxx = os.environ.get("CXX") or shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        raise RuntimeError("g++ not found; the compiled Python boundary cannot be built")
    torch_dir = Path(torch.__file__).parent
    cuda_inc = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda")) / "include"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-w",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           str(BINDING_SRC),
           f"-I{torch_dir / 'include'}", f"-I{torch_dir / 'include' / 'torch' / 'csrc' / 'api' / 'include'}",
           f"-I{sysconfig.get_paths()['include']}", f"-I{cuda_inc}",
           f"-L{torch_dir / 'lib'}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python",
           f"-L{LIB_DIR}", "-l:libhi_b200.so",
           "-Wl,-rpath,$ORIGIN/../../lib", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_dir / 'lib'}",
           "-o", str(BINDING_LIB)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"g++ failed for torch_binding.cpp:\n{res.stdout}\n{res.stderr[-6000:]}")
    paths = binding_paths()
    for path in paths.values():
        path.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(BINDING_LIB, path)
    (LIB_DIR / "binding.stamp").write_text(_binding_stamp())
    return paths


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def _stamp() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    """True when lib/libhi_b200.so was built from the sources as they are now."""
    stamp_file = LIB_DIR / "build.stamp"
    return LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == _stamp()


def build_variant(name: str, defines: list[str]) -> Path:
    """Dev builds (tracing / debug instrumentation): lib/libhi_b200_<name>.so compiled with extra -D flags.  Loaded
    instead of the product library when HI_B200_LIB points at it (hydrainfer_b200/_lib.py); never the default."""
    nvcc = _nvcc()
    obj_dir = CSRC / f"build_{name}"
    obj_dir.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    out = LIB_DIR / f"libhi_b200_{name}.so"
    objs = []
    for src in SOURCES:
        obj = obj_dir / (Path(src).stem + ".o")
        res = subprocess.run([nvcc, *NVCC_FLAGS, *defines, "-c", str(CSRC / src), "-o", str(obj)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        objs.append(str(obj))
    res = subprocess.run([nvcc, "-shared", "-o", str(out), *objs, "-cudart", "static", "-Xlinker", "-soname=libhi_b200.so"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile (if sources changed) and return the path of the shared library."""
    stamp_file = LIB_DIR / "build.stamp"
    stamp = _stamp()
    if not force and LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB_PATH
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)

    def compile_one(name: str) -> tuple[str, str]:
        obj = OBJ_DIR / (Path(name).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / name), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{res.stdout}\n{res.stderr}")
        return str(obj), res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    (OBJ_DIR / "ptxas.log").write_text("\n".join(log for _, log in results))
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *[o for o, _ in results], "-cudart", "static", "-Xlinker", "-soname=libhi_b200.so"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp_file.write_text(stamp)
    return LIB_PATH


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m hydrainfer_b200.build --variant trace -DHI_PAIR_TRACE
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        path = build(force="--force" in sys.argv, verbose=True)
        print(path)
        for p in build_binding(force="--force" in sys.argv).values():
            print(p)
